"""cProfile of the eager e2e step (host overhead of the drop-in module path).  timeout 120 python profiles/debug/profile_e2e.py"""
import cProfile, os, pstats, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dedark_yolo_b200 as dd

dev = torch.device("cuda:0")
B, H, W = 16, 640, 640
torch.manual_seed(0)
module = dd.lowlight_recovery(3).to(dev).train()
params = list(module.parameters())
host = [torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
g = torch.randn(B, 3, H, W, device=dev)
host_out = torch.empty(2).pin_memory()
pf = dd.HostBatchPrefetcher(dev)
pf.submit(host[0])

def step(i):
    for p in params:
        p.grad = None
    src = pf.get()
    pf.submit(host[(i + 1) % 2])
    batch = dd.preprocess_batch({"img": src}, dev, dark_param=15.0)
    y = module(batch["img"])
    y.backward(g)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    res = torch.stack([batch["recovery_loss_batch"], flat.norm()])
    host_out.copy_(res, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()

for i in range(5):
    step(i)
t0 = time.perf_counter()
for i in range(50):
    step(i)
print("ms/step", (time.perf_counter() - t0) / 50 * 1e3)
# host-only cost: same calls without the per-step sync (launch-bound rate)
pr = cProfile.Profile()
pr.enable()
for i in range(50):
    step(i)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
