"""SURVEY.md section 8(f) N2 -- uint8 batch -> synthesis -> filter chain without the darkened fp32 batch in HBM
(models/yolo/detect/train.py:72-109).  The filter kernels look the darkened value up in the 256-entry table in their stage phase:
everything must be BIT-IDENTICAL to the two-step route (dd_synth_fwd materialises the batch, dd_recovery_fwd/bwd read it)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(B, H, W, seed, custom):
    gen = torch.Generator().manual_seed(seed)
    dev = torch.device("cuda:0")
    u8 = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=gen).to(dev)
    feat = (torch.randn(B, 15, generator=gen) * 0.8).to(dev)
    g = torch.randn(B, 3, H, W, generator=gen).to(dev)
    A = (0.4 + 0.5 * torch.rand(B, 3, generator=gen)).to(dev) if custom else None
    IcA = torch.rand(B, 1, H, W, generator=gen).to(dev) if custom else None
    return u8, feat, g, A, IcA


@pytest.mark.parametrize("B,H,W,p,custom", [(2, 96, 80, 5.0, False), (2, 96, 80, 1.7, True), (3, 200, 132, 7.5, True), (1, 13, 16, 2.0, False),
                                            (16, 640, 640, 15.0, False), (4, 640, 640, 2.2, False)])
def test_u8_chain_bit_identical_to_materialised_batch(B, H, W, p, custom):
    import dedark_yolo_b200 as dd
    from dedark_yolo_b200 import ops

    u8, feat, g, A, IcA = _inputs(B, H, W, B * H + W, custom)
    _, dark, _, _ = ops.synth_forward(u8, p, want_clean=False, want_rec=False)
    tab = ops.dark_table(p, u8.device)
    k = torch.arange(256, dtype=torch.uint8, device=u8.device)
    assert torch.equal(tab, ops.synth_forward(k, p, want_clean=False, want_rec=False)[1])   # the table is the synthesis' own bits
    n0 = dd.launch_count()
    y8 = ops.filters_forward_u8(u8, tab, feat, A, IcA)
    d8 = ops.filters_backward_u8(u8, tab, feat, g, A, IcA)
    assert dd.launch_count() - n0 == 3
    y = ops.filters_forward(dark, feat, A, IcA)
    d, _ = ops.filters_backward(dark, feat, g, A, IcA)
    torch.cuda.synchronize()
    assert torch.equal(y8, y), float((y8 - y).abs().max())
    assert torch.equal(d8, d), (d8, d)
    assert float(y.abs().max()) > 0 and float(d.abs().max()) > 0


def test_u8_chain_rejects_unaligned_width():
    from dedark_yolo_b200 import ops

    u8, feat, g, _, _ = _inputs(1, 20, 22, 3, False)
    tab = ops.dark_table(5.0, u8.device)
    with pytest.raises(ValueError):
        ops.filters_forward_u8(u8, tab, feat)


@pytest.mark.parametrize("overlapped", [False, True])
def test_pipeline_without_dark_batch_is_bit_identical(overlapped):
    import dedark_yolo_b200 as dd

    dev = torch.device("cuda:0")
    B, H, W = 4, 320, 320
    torch.manual_seed(0)
    m = dd.lowlight_recovery(3).to(dev).train()
    gen = torch.Generator().manual_seed(11)
    srcs = [torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=gen).to(dev) for _ in range(3)]
    gs = [torch.randn(B, 3, H, W, generator=gen).to(dev) for _ in range(3)]
    outs = {}
    for mat in (True, False):
        pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=5.0, src_dtype=torch.uint8, materialize_dark=mat)
        assert pipe.fused_resize
        res = []
        if overlapped:
            pipe.prime(srcs[0])
            for i in range(3):
                y, rec, fg = pipe.step_overlapped(srcs[i + 1] if i < 2 else None, gs[i])
                res.append((y.clone(), rec.clone(), fg.clone()))
        else:
            for i in range(3):
                y, rec, fg = pipe.step(srcs[i], gs[i])
                res.append((y.clone(), rec.clone(), fg.clone()))
        torch.cuda.synchronize()
        outs[mat] = res
        assert (pipe.dark is None) == (not mat)
    for (y0, r0, f0), (y1, r1, f1) in zip(outs[True], outs[False]):
        assert torch.equal(y0, y1) and torch.equal(r0, r1) and torch.equal(f0, f1)
        assert float(f0.abs().max()) > 0
