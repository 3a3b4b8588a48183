"""CUPTI timeline (rank 0) of the bucketed data-parallel step: where the time between the single-GPU step and the N-GPU step goes.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 profiles/debug/ddp_timeline.py
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dedark_yolo_b200 as dd
from dedark_yolo_b200.dist import init_from_env

rank, local_rank, world = init_from_env("nccl")
dev = torch.device("cuda", local_rank)
torch.cuda.set_device(dev)
B, H, W, RING = 16, 640, 640, 4
torch.manual_seed(0)
m = dd.lowlight_recovery(3).to(dev).train()
pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=15.0, allreduce=world > 1)
gen = torch.Generator(device=dev).manual_seed(1 + rank)
cleans = [torch.rand(B, 3, H, W, generator=gen, device=dev) for _ in range(RING)]
gs = [torch.randn(B, 3, H, W, generator=gen, device=dev) for _ in range(RING)]
for k in range(RING):
    pipe.capture_overlapped(("ovl", k), cleans[(k + 1) % RING], gs[k], slot=k % 2)
pipe._cur = 0
pipe.prime(cleans[0])
for i in range(12):
    pipe.replay_overlapped(("ovl", i % RING))
torch.cuda.synchronize()
if world > 1:
    torch.distributed.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(12, 16):
        pipe.replay_overlapped(("ovl", i % RING))
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    for e in evs:
        print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:90]}")
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
