"""Tiny driver for ncu: a few eager steps of the bench workload (configs[1]: 16x3x640x640 fp32 per GPU).

    python profiles/prof_step.py [steps] [B] [H] [W]

Run plain first, then under ncu (B200_PROFILING.md); numbers printed under ncu are never bench values."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dedark_yolo_b200 as dd  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, H, W = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (16, 640, 640)
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = dd.lowlight_recovery(3).to(dev).train()
U8 = os.environ.get("PROF_SRC", "f32") == "u8"  # the e2e path: uint8 batch, synthesis fused with the resize
BF16 = os.environ.get("PROF_IO", "f32") == "bf16"  # bf16 I/O mode: tensor-core filter kernels
N2 = os.environ.get("PROF_N2", "0") == "1"         # uint8 batch read by the filter kernels (no dark batch)
pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=15.0, src_dtype=torch.uint8 if U8 else torch.float32,
                           io_dtype=torch.bfloat16 if BF16 else torch.float32, materialize_dark=not N2)
gen = torch.Generator(device=dev).manual_seed(1234)
clean = (torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, device=dev) if U8
         else torch.rand(B, 3, H, W, generator=gen, device=dev))
g = torch.randn(B, 3, H, W, generator=gen, device=dev)
if BF16:
    g = g.bfloat16()
n0 = dd.launch_count()
for _ in range(steps):
    pipe.step(clean, g)
torch.cuda.synchronize()
print(f"ok: {steps} steps, {dd.launch_count() - n0} launches, rec={float(pipe.rec):.6f}, |grad|={float(pipe.flat_grad.norm()):.4e}")
