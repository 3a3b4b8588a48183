"""GPU timeline (CUPTI via torch.profiler) of one software-pipelined step: which kernels of the two streams overlap."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import profile, ProfilerActivity
import dedark_yolo_b200 as dd

dev = torch.device("cuda", 0)
B, H, W = 16, 640, 640
m = dd.lowlight_recovery(3).to(dev).train()
U8 = os.environ.get("SRC", "f32") == "u8"
pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=15.0, src_dtype=torch.uint8 if U8 else torch.float32)
gen = torch.Generator(device=dev).manual_seed(1)
if U8:
    cleans = [torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, device=dev) for _ in range(4)]
else:
    cleans = [torch.rand(B, 3, H, W, generator=gen, device=dev) for _ in range(4)]
PLAIN = os.environ.get("PLAIN") == "1"
gs = [torch.randn(B, 3, H, W, generator=gen, device=dev) for _ in range(4)]
pipe.prime(cleans[0])
for i in range(6):
    pipe.step_overlapped(cleans[(i + 1) % 4], gs[i % 4])
torch.cuda.synchronize()
GRAPH = os.environ.get("GRAPH") == "1"   # plain steps as CUDA-graph replays: the gaps between the kernels of a captured step
if GRAPH:
    for k in range(4):
        pipe.capture(k, cleans[k], gs[k])
    for i in range(8):
        pipe.replay(i % 4)
    torch.cuda.synchronize()
OVL_GRAPH = os.environ.get("OVL_GRAPH") == "1"   # the headline configuration of bench.py: the pipelined step as one graph per step
if OVL_GRAPH:
    for k in range(4):
        pipe.capture_overlapped(("ovl", k), cleans[(k + 1) % 4], gs[k], slot=k % 2)
    pipe._cur = 0
    pipe.prime(cleans[0])
    for i in range(8):
        pipe.replay_overlapped(("ovl", i % 4))
    torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(8, 11):
        if OVL_GRAPH:
            pipe.replay_overlapped(("ovl", i % 4))
        elif GRAPH:
            pipe.replay(i % 4)
        elif PLAIN:
            pipe.step(cleans[i % 4], gs[i % 4])
        else:
            pipe.step_overlapped(cleans[(i + 1) % 4], gs[i % 4])
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
prev_end = None
for e in evs:
    gap = "" if prev_end is None else f" gap {e.time_range.start - prev_end:6.1f}"
    prev_end = max(prev_end or 0, e.time_range.end)
    print(gap, end=" ")
    print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:7.1f} us  stream? {getattr(e, 'device_resource_id', '?')!s:4}  {e.name[:70]}")
