"""Where the e2e step's time goes: uint8 pipeline graphs (a) replayed back to back, (b) with a stream sync per step,
(c) with the H2D of the next batch in flight as well."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import dedark_yolo_b200 as dd

dev = torch.device("cuda", 0)
B, H, W = 16, 640, 640
m = dd.lowlight_recovery(3).to(dev).train()
pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=15.0, src_dtype=torch.uint8)
gen = torch.Generator(device=dev).manual_seed(1)
gs = [torch.randn(B, 3, H, W, generator=gen, device=dev) for _ in range(2)]
host = [torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
slots = [h.to(dev) for h in host]
host_out = torch.empty(2).pin_memory()
res = torch.empty(2, device=dev)
def readback(y, rec, flat):
    res[0].copy_(rec); torch.linalg.vector_norm(flat, out=res[1]); host_out.copy_(res, non_blocking=True)
for k in range(2):
    pipe.capture_overlapped(k, slots[(k + 1) % 2], gs[k], slot=k, epilogue=readback)
pipe._cur = 0
pipe.prime(slots[0])
copy_stream = torch.cuda.Stream(dev)
N = 200
def run(sync, h2d):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        pipe.replay_overlapped(i % 2)
        if h2d:
            with torch.cuda.stream(copy_stream):
                slots[i % 2].copy_(host[i % 2], non_blocking=True)   # (racy on purpose: only the traffic matters here)
        if sync:
            torch.cuda.current_stream().synchronize()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / N * 1e6
for sync, h2d in ((False, False), (True, False), (False, True), (True, True)):
    run(sync, h2d)
    print(f"sync={sync} h2d={h2d}: {run(sync, h2d):.1f} us/step")
