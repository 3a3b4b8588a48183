"""Launch-order independence of the programmatic-dependent-launch scheme (csrc/dd_common.cuh, rules 3-5).

The filter kernels request their first tile and per-row columns of the batch AHEAD of griddepcontrol.wait.  That is only correct if
the kernel that produced the batch cannot still be writing it, whatever a caller launches in front: the producers (synthesis,
dark-channel prior) never release their dependents early.  Here the producer is launched IMMEDIATELY in front of the consumer on the
same stream, into the SAME buffers with NEW contents every iteration (a consumer that read early would see the previous iteration's
values), and the result is compared bit for bit with a run that synchronises the device between the two calls.

(Checked once with a library whose producers DO release early: the test still passed on a B200 -- the synthesis fills every thread
slot of the GPU, so the consumer's CTAs only become resident when it is nearly done.  The rule does not rely on that; this test pins
the contract, it is not a proof of the rule's necessity.)
"""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(B=8, H=640, W=640):
    import dedark_yolo_b200 as dd

    torch.manual_seed(7)
    dev = torch.device("cuda", 0)
    m = dd.lowlight_recovery(3).to(dev).train()
    pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=15.0)
    feat = (torch.randn(B, 15, device=dev) * 0.3).contiguous()
    return dd, dev, pipe, feat


@pytest.mark.parametrize("direction", ["forward", "backward"])
def test_filter_kernels_right_behind_the_synthesis(direction):
    from dedark_yolo_b200 import _lib
    from dedark_yolo_b200.pipeline import _p

    dd, dev, pipe, feat = _setup()
    B, H, W = pipe.B, pipe.H, pipe.W
    st = torch.cuda.current_stream(dev).cuda_stream
    g = torch.randn(B, 3, H, W, device=dev)
    dfeat = torch.empty(B, 15, device=dev)
    outs = {}
    for mode in ("sync", "back_to_back"):
        gen = torch.Generator(device=dev).manual_seed(11)
        res = []
        for it in range(12):
            src = torch.rand(B, 3, H, W, generator=gen, device=dev)   # new contents, same destination buffers (pipe.dark)
            # the synthesis kernel alone (no recovery loss, hence no finalize kernel between it and the consumer)
            _lib.check(_lib.lib.dd_synth_fwd(_p(src), _lib.SRC_F32, C.c_float(15.0), None, None, None, _p(pipe.dark), None, None,
                                             C.c_longlong(src.numel()), None, 0, st))
            if mode == "sync":
                torch.cuda.synchronize(dev)
            if direction == "forward":
                _lib.check(_lib.lib.dd_recovery_fwd(_p(pipe.dark), None, None, _p(feat), _p(pipe.y), B, H, W, st))
                res.append(pipe.y.clone())
            else:
                _lib.check(_lib.lib.dd_recovery_bwd(_p(pipe.dark), None, None, _p(feat), _p(g), _p(dfeat), None, B, H, W,
                                                    _p(pipe._ws_rb), pipe._ws_rb.numel(), st))
                res.append(dfeat.clone())
        torch.cuda.synchronize(dev)
        outs[mode] = res
    for a, b in zip(outs["sync"], outs["back_to_back"]):
        assert torch.equal(a, b)


def test_filter_forward_right_behind_the_dark_channel_prior():
    from dedark_yolo_b200 import ops

    dd, dev, pipe, feat = _setup(B=4, H=320, W=320)
    B, H, W = pipe.B, pipe.H, pipe.W
    outs = {}
    for mode in ("sync", "back_to_back"):
        gen = torch.Generator(device=dev).manual_seed(5)
        res = []
        for it in range(8):
            src = torch.randint(0, 256, (B, 3, H, W), generator=gen, device=dev, dtype=torch.uint8)
            x = torch.rand(B, 3, H, W, generator=gen, device=dev)
            torch.cuda.synchronize(dev)
            A, IcA = ops.dark_prior(src, 15.0)          # writes A and IcA (image-sized) ...
            if mode == "sync":
                torch.cuda.synchronize(dev)
            y = ops.filters_forward(x, feat, A=A, IcA=IcA)   # ... which the filter kernel reads ahead of its grid dependency
            res.append(y.clone())
        torch.cuda.synchronize(dev)
        outs[mode] = res
    for a, b in zip(outs["sync"], outs["back_to_back"]):
        assert torch.equal(a, b)
