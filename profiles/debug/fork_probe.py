"""Where should the synthesis of batch n+1 fork off?  Times the software-pipelined step (CUDA graph replay, 16x3x640x640 fp32)
with the fork behind the filter backward (shipped) and before it, the latter also with the step on a high-priority stream so
that the filter backward's persistent CTAs are placed before the synthesis CTAs.

    python profiles/debug/fork_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dedark_yolo_b200 as dd  # noqa: E402

dev = torch.device("cuda:0")
B, H, W, RING = 16, 640, 640, 4
torch.manual_seed(0)
m = dd.lowlight_recovery(3).to(dev).train()
gen = torch.Generator(device=dev).manual_seed(1234)
cleans = [torch.rand(B, 3, H, W, generator=gen, device=dev) for _ in range(RING)]
gs = [torch.randn(B, 3, H, W, generator=gen, device=dev) for _ in range(RING)]


def run(early, high_prio, side_prio=0):
    stream = torch.cuda.Stream(dev, priority=-1 if high_prio else 0)
    with torch.cuda.stream(stream):
        pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=15.0)
        pipe.fork_before_filters_bwd = early
        pipe.enable_overlap()
        if side_prio:
            pipe._side = torch.cuda.Stream(dev, priority=side_prio)
        for i in range(RING):
            pipe.capture_overlapped(("o", i), cleans[(i + 1) % RING], gs[i], slot=i % 2)
        pipe._cur = 0
        pipe.prime(cleans[0])
        for n in range(12):
            pipe.replay_overlapped(("o", n % RING))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for n in range(12, 212):
            pipe.replay_overlapped(("o", n % RING))
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / 200
        ref = pipe.flat_grad.clone()
    print(f"fork {'before' if early else 'behind'} the filter backward, main stream priority {'high' if high_prio else 'default'}: "
          f"{ms * 1e3:.1f} us per step, |grad| = {float(ref.norm()):.6e}", flush=True)


run(False, False)
run(True, False)
run(True, True)
run(False, True)
