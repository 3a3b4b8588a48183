"""2-GPU probe: predictor backward + gradient exchange, fused (peer memory) vs dd_predictor_bwd + NCCL all-reduce.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 profiles/debug/exchange_probe.py"""
import ctypes as C, os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dedark_yolo_b200 as dd
from dedark_yolo_b200 import _lib
from dedark_yolo_b200.dist import init_from_env, GradExchange
from dedark_yolo_b200.pipeline import _p

rank, local_rank, world = init_from_env("nccl")
dev = torch.device("cuda", local_rank); torch.cuda.set_device(dev)
torch.manual_seed(0)
m = dd.lowlight_recovery(3).to(dev).train()
B, H, W = 16, 640, 640
ex = GradExchange(dev)
pipe = dd.RecoveryPipeline(m, B, H, W, exchange=ex)
gen = torch.Generator(device=dev).manual_seed(rank)
clean = torch.rand(B, 3, H, W, generator=gen, device=dev); g = torch.randn(B, 3, H, W, generator=gen, device=dev)
st = torch.cuda.current_stream(dev).cuda_stream
pipe.synth(clean, st); pipe.forward(st)
_lib.check(_lib.lib.dd_recovery_bwd(_p(pipe.dark), None, None, _p(pipe.feat), _p(g), _p(pipe.dfeat), None, B, H, W, _p(pipe._ws_rb), pipe._ws_rb.numel(), st))

def fused():
    _lib.check(_lib.lib.dd_predictor_bwd_allreduce(_p(pipe.r), C.byref(pipe._w), _p(pipe.acts), _p(pipe.dfeat), C.byref(pipe._g), B, _p(pipe._ws_pb), pipe._ws_pb.numel(), C.byref(ex.px), st))
def plain():
    _lib.check(_lib.lib.dd_predictor_bwd(_p(pipe.r), C.byref(pipe._w), _p(pipe.acts), _p(pipe.dfeat), C.byref(pipe._g), None, B, _p(pipe._ws_pb), pipe._ws_pb.numel(), st))
def nccl():
    plain(); dist.all_reduce(pipe.flat_grad)
def nccl_only():
    dist.all_reduce(pipe.flat_grad)

def timeit(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

res = {k: timeit(f) for k, f in (("plain (no exchange)", plain), ("fused peer exchange", fused), ("plain + NCCL all-reduce", nccl),
                                 ("NCCL all-reduce only", nccl_only))}
fused(); torch.cuda.synchronize(); a = pipe.flat_grad.clone()
nccl(); torch.cuda.synchronize(); b = pipe.flat_grad.clone()
if rank == 0:
    print({k: round(v, 1) for k, v in res.items()}, "us per call; max |fused - nccl| / max:", float((a - b).abs().max() / b.abs().max()), "how:", ex.how)
dist.barrier(); dist.destroy_process_group()
