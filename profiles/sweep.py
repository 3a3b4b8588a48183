"""BASELINE configs[4]: resolution / batch sweep of the hot path at 1 / 2 / 4 / 8 GPUs (CUDA-graph replay of the step, CUDA
events; per-stage times from eager C-ABI calls).

    python profiles/sweep.py > profiles/r02_sweep_1gpu.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/sweep.py [short] \
        > profiles/r02_sweep_Ngpu.json          # B is the batch PER GPU (weak scaling); one NCCL all-reduce of the gradient per step

For every (B, H, W): images/s of the full step (synthesis + fwd + bwd) replayed from a CUDA graph, and the algorithmic
HBM bandwidth of the two filter kernels as a fraction of the measured peak (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dedark_yolo_b200 as dd  # noqa: E402
from dedark_yolo_b200 import _lib  # noqa: E402
from dedark_yolo_b200.pipeline import _p  # noqa: E402

from dedark_yolo_b200.dist import init_from_env, max_over_ranks  # noqa: E402

rank, local_rank, world = init_from_env("nccl")
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
peak = 6550.4
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
torch.manual_seed(0)
m = dd.lowlight_recovery(3).to(dev).train()
cases = [(1, 320, 320), (16, 320, 320), (64, 320, 320), (256, 320, 320), (16, 480, 480), (1, 640, 640), (4, 640, 640), (16, 640, 640),
         (64, 640, 640), (256, 640, 640), (16, 960, 960), (1, 1280, 1280), (8, 1280, 1280), (32, 1280, 1280), (16, 333, 517)]
if "short" in sys.argv:  # the multi-GPU runs: box time is charged per GPU
    cases = [(1, 640, 640), (16, 320, 320), (16, 640, 640), (64, 640, 640), (256, 640, 640), (16, 1280, 1280), (32, 1280, 1280)]
out = []


def barrier():
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize(dev)


for B, H, W in cases:
    pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=15.0, allreduce=world > 1)
    gen = torch.Generator(device=dev).manual_seed(1)
    ring = 2 if B * H * W * 12 > 100e6 else 4
    cl = [torch.rand(B, 3, H, W, generator=gen, device=dev) for _ in range(ring)]
    gs = [torch.randn(B, 3, H, W, generator=gen, device=dev) for _ in range(ring)]
    for i in range(ring):
        pipe.capture(i, cl[i], gs[i])
    steps = max(20, min(200, int(3e9 / (B * H * W * 12 * 8))))
    for i in range(5):
        pipe.replay(i % ring)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        pipe.replay(i % ring)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1), dev) / steps
    # stage times (eager)
    st = torch.cuda.current_stream(dev).cuda_stream
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(10)]
    for i in range(10):
        pipe.synth(cl[i % ring], st)
        pipe.forward(st)
        ev[i][0].record()
        _lib.check(_lib.lib.dd_recovery_fwd(_p(pipe.dark), None, None, _p(pipe.feat), _p(pipe.y), B, H, W, st))
        ev[i][1].record()
        _lib.check(_lib.lib.dd_recovery_bwd(_p(pipe.dark), None, None, _p(pipe.feat), _p(gs[i % ring]), _p(pipe.dfeat), None, B, H, W,
                                            _p(pipe._ws_rb), pipe._ws_rb.numel(), st))
        ev[i][2].record()
    torch.cuda.synchronize()
    fwd_us = 1e3 * sum(ev[i][0].elapsed_time(ev[i][1]) for i in range(2, 10)) / 8
    bwd_us = 1e3 * sum(ev[i][1].elapsed_time(ev[i][2]) for i in range(2, 10)) / 8
    byt = 24 * B * H * W
    out.append({"B_per_gpu": B, "H": H, "W": W, "images_per_s": world * B / (ms * 1e-3), "ms_per_step": ms, "filters_fwd_us": fwd_us,
                "filters_bwd_us": bwd_us, "fwd_frac_of_hbm_peak": byt / (fwd_us * 1e-6) / 1e9 / peak,
                "bwd_frac_of_hbm_peak": byt / (bwd_us * 1e-6) / 1e9 / peak})
    if rank == 0:
        print(f"# {world} GPU(s), B={B}/GPU {H}x{W}: {out[-1]['images_per_s']:.0f} img/s, {ms:.3f} ms/step, fwd {fwd_us:.0f} us, bwd {bwd_us:.0f} us", file=sys.stderr)
    del pipe, cl, gs
    torch.cuda.empty_cache()
if rank == 0:
    print(json.dumps({"peak_gbs": peak, "n_gpus": world,
                      "unit": f"images/s over {world} B200 (fp32, CUDA-graph replay of synthesis + fwd + bwd per rank"
                              + (", + one NCCL all-reduce(sum) of the 164 943 gradients per step; max over ranks)" if world > 1 else ")"),
                      "cases": out}, indent=1))
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
