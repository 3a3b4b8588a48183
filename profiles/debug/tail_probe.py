"""Times dd_predictor_fwd (B = 16) with the fused tail (conv4 + conv5 + fc in one cluster launch) and with the three-launch
tensor-core tail (DEDARK_TAIL=tc), in bursts of back-to-back calls.  timeout 120 python profiles/debug/tail_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dedark_yolo_b200 as dd
from dedark_yolo_b200 import ops

B = 16
torch.manual_seed(0)
m = dd.lowlight_recovery(3).to("cuda")
params = [p.detach() for p in m.extractor.ordered_parameters()]
r = torch.rand(B, 3, 256, 256, device="cuda")
out = {}
for mode in ("fused", "tc", "fused", "tc"):  # DEDARK_TAIL: "fused" = the experiment, anything else = the shipped three launches
    os.environ["DEDARK_TAIL"] = mode
    for _ in range(5):
        feat, acts = ops.predictor_forward(r, params)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(12):
            feat, acts = ops.predictor_forward(r, params)
        b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / 12)
    ts.sort()
    out[mode] = feat.clone()
    print(f"tail {mode}: predictor fwd median {ts[len(ts)//2]:.1f} us (best {ts[0]:.1f})", flush=True)
print("max |feat fused - feat tc| =", (out["fused"] - out["tc"]).abs().max().item())

# ---- the predictor backward (all launches of dd_predictor_bwd): 12 calls captured in one CUDA graph (no host time between them)
os.environ.pop("DEDARK_TAIL", None)
feat, acts = ops.predictor_forward(r, params)
dfeat = torch.randn(B, 15, device="cuda")
flat = torch.empty(sum(p.numel() for p in params), device="cuda")
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for _ in range(3):
        ops.predictor_backward(r, params, acts, dfeat, flat_grad=flat)
    side.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(12):
            ops.predictor_backward(r, params, acts, dfeat, flat_grad=flat)
    for _ in range(3):
        g.replay()
    side.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(side)
        g.replay()
        b.record(side); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / 12)
ts.sort()
print(f"predictor bwd (graph of 12) median {ts[len(ts)//2]:.1f} us (best {ts[0]:.1f})", flush=True)
