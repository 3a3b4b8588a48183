"""Where the host time of one step through the reference-shaped API goes (VERDICT r1 'What's weak' 6): cProfile of
preprocess_batch -> lowlight_recovery(nn.Module) forward -> autograd backward at 16x3x640x640, GPU work enqueued
asynchronously (one sync at the end), plus wall-clock per step with and without a per-step sync.

    python profiles/debug/host_overhead.py
"""
import cProfile
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dedark_yolo_b200 as dd  # noqa: E402

dev = torch.device("cuda:0")
B, H, W = 16, 640, 640
torch.manual_seed(0)
m = dd.lowlight_recovery(3).to(dev).train()
params = list(m.parameters())
gen = torch.Generator(device=dev).manual_seed(1)
u8 = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=gen, device=dev)
g = torch.randn(B, 3, H, W, generator=gen, device=dev)


def step():
    for p in params:
        p.grad = None
    batch = dd.preprocess_batch({"img": u8}, dev, dark_param=15.0, dedark_FLAG=False)
    y = m(batch["img"])
    y.backward(g)
    return batch["recovery_loss_batch"]


for _ in range(10):
    step()
torch.cuda.synchronize()
N = 200
t0 = time.perf_counter()
for _ in range(N):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue time per step: {1e6 * (t1 - t0) / N:.0f} us; with the final drain: {1e6 * (t2 - t0) / N:.0f} us per step")
t0 = time.perf_counter()
for _ in range(N):
    step()
    torch.cuda.synchronize()
print(f"synchronous step: {1e6 * (time.perf_counter() - t0) / N:.0f} us")

# forward and backward separately
torch.cuda.synchronize()
batch = dd.preprocess_batch({"img": u8}, dev, dark_param=15.0, dedark_FLAG=False)
t_pre = t_f = t_b = 0.0
for _ in range(N):
    for p in params:
        p.grad = None
    a = time.perf_counter()
    batch = dd.preprocess_batch({"img": u8}, dev, dark_param=15.0, dedark_FLAG=False)
    b = time.perf_counter()
    y = m(batch["img"])
    c = time.perf_counter()
    y.backward(g)
    d = time.perf_counter()
    t_pre += b - a
    t_f += c - b
    t_b += d - c
    if _ % 20 == 19:
        torch.cuda.synchronize()
print(f"host time per call: preprocess_batch {1e6 * t_pre / N:.0f} us, module forward {1e6 * t_f / N:.0f} us, backward {1e6 * t_b / N:.0f} us")

pr = cProfile.Profile()
pr.enable()
for _ in range(100):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
