"""Per-layer check of the predictor (tensor-core convs) against torch fp64 on the same GPU.  Debug aid:
    timeout 120 python profiles/debug/check_predictor.py [B]"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dedark_yolo_b200 as dd
from dedark_yolo_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
dev = "cuda"
m = dd.lowlight_recovery(3).to(dev)
params = [p.detach() for p in m.extractor.ordered_parameters()]
r = torch.rand(B, 3, 256, 256, device=dev)
dfeat = torch.randn(B, 15, device=dev)
feat, acts = ops.predictor_forward(r, params)
torch.cuda.synchronize()
p64 = [p.double().requires_grad_(True) for p in params]
x = r.double()
sizes = [(16, 128), (32, 64), (32, 32), (32, 16), (32, 8)]
off = 0
ref_acts = []
for l in range(5):
    x = F.leaky_relu(F.conv2d(x, p64[2 * l], p64[2 * l + 1], stride=2, padding=1), 0.1)
    x.retain_grad()
    ref_acts.append(x)
    c, h = sizes[l]
    n = B * c * h * h
    got = acts[off:off + n].view(B, c, h, h).double()
    off += n
    err = (got - x).abs().max().item() / x.abs().max().item()
    print(f"act{l + 1}: rel-to-max {err:.3e}")
hh = F.leaky_relu(F.linear(x.reshape(B, -1), p64[10], p64[11]), 0.1)
f64 = F.linear(hh, p64[12], p64[13])
print(f"feat: rel-to-max {(feat.double() - f64).abs().max().item() / f64.abs().max().item():.3e}")
f64.backward(dfeat.double())
grads, _ = ops.predictor_backward(r, params, acts, dfeat)
torch.cuda.synchronize()
names = [f"conv{l + 1}.{k}" for l in range(5) for k in ("w", "b")] + ["fc1.w", "fc1.b", "fc2.w", "fc2.b"]
for nme, g, p in zip(names, grads, p64):
    e = (g.double() - p.grad).abs().max().item() / p.grad.abs().max().item()
    print(f"grad {nme}: rel-to-max {e:.3e}")
print("launches", dd.launch_count())
for idx in (8, 9):
    g, p = grads[idx], p64[idx]
    print(names[idx], "got max", g.abs().max().item(), "ref max", p.grad.abs().max().item())
    print(" got", g.flatten()[:8].tolist())
    print(" ref", p.grad.flatten()[:8].tolist())
