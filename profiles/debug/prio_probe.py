"""Does the synthesis of batch n+1 really run BESIDE the predictor backward of batch n?  (CUPTI timeline: no -- its persistent
CTAs fill every thread slot, so the backward's kernels start when it ends.)  Times the pipelined step (CUDA graph replay,
16x3x640x640) for: persistent / short synthesis CTAs x the step captured on a default / high-priority stream (the side stream
that carries the synthesis keeps the default, i.e. lower, priority).

    python profiles/debug/prio_probe.py

Measured (profiles/r02_prio_probe.txt): 373 us default, 388 us with the step on a high-priority stream; with the synthesis cut into
9600 short CTAs instead of 1184 persistent ones (a build that is not kept) 385 / 388 us.  Neither gives the backward's kernels
room beside the synthesis; the shipped step keeps default priorities and persistent synthesis CTAs.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dedark_yolo_b200 as dd

dev = torch.device("cuda", 0)
B, H, W, RING = 16, 640, 640, 4
torch.manual_seed(0)
m = dd.lowlight_recovery(3).to(dev).train()
U8 = os.environ.get("SRC", "f32") == "u8"
if U8:
    cleans = [torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, device=dev) for _ in range(RING)]
else:
    cleans = [torch.rand(B, 3, H, W, device=dev) for _ in range(RING)]
gs = [torch.randn(B, 3, H, W, device=dev) for _ in range(RING)]


def run(high_prio):
    pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=15.0, src_dtype=torch.uint8 if U8 else torch.float32)
    pipe.enable_overlap()
    if high_prio:
        pipe.capture_stream = torch.cuda.Stream(dev, priority=-1)
    for i in range(RING):
        pipe.capture_overlapped(("o", i), cleans[(i + 1) % RING], gs[i], slot=i % 2)
    pipe._cur = 0
    pipe.prime(cleans[0])
    for n in range(12):
        pipe.replay_overlapped(("o", n % RING))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for n in range(12, 212):
        pipe.replay_overlapped(("o", n % RING))
    e1.record()
    e1.synchronize()
    print(f"synth CTAs {os.environ.get('DEDARK_SYNTH_CTAS', 'persistent')}, src {'u8' if U8 else 'f32'}, step captured on a "
          f"{'high-priority' if high_prio else 'default'} stream: {e0.elapsed_time(e1) / 200 * 1e3:.1f} us per step, "
          f"|grad| = {float(pipe.flat_grad.norm()):.6e}, rec = {float(pipe._slots[0].rec):.8f}", flush=True)


run(False)
run(True)
run(False)
run(True)
