"""Round-2 bring-up of the tensor-core blur engine (dd_blur_tc.cuh): numerics of the bare blur (dd_debug_blur_tc, 3xTF32 and
1xTF32) against an fp64 reflect-padded separable Gaussian, the fused forward / backward with DEDARK_BLUR=tc against the
CUDA-core kernels (DEDARK_BLUR=cc) and the fp64 oracle, and CUDA-event timings of both at 16x3x640x640.

    python profiles/debug/tc_blur_check.py [quick]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import dedark_yolo_b200 as dd  # noqa: E402
from dedark_yolo_b200 import ops  # noqa: E402
from oracle import lowlight_oracle as O  # noqa: E402

dev = torch.device("cuda:0")


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def blur_ref(x):
    return O.blur_separable(x.double())


def check_blur():
    for shape in [(1, 3, 96, 80), (2, 3, 64, 128), (1, 3, 200, 132), (1, 3, 13, 16), (1, 3, 640, 640), (1, 3, 333, 516)]:
        g = torch.Generator().manual_seed(sum(shape))
        x = torch.rand(shape, generator=g) * 2 - 0.5
        ref = blur_ref(x)
        for x3 in (True, False):
            t0 = time.time()
            try:
                y = ops.debug_blur_tc(x.to(dev), x3=x3)
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print(f"blur {shape} x3={x3}: FAILED {e!r}", flush=True)
                raise
            err = rel(y.cpu(), ref)
            bad = (y.cpu().double() - ref).abs()
            idx = torch.nonzero(bad == bad.max())[0].tolist()
            print(f"blur {shape} x3={int(x3)}: rel-to-max {err:.3e} (worst at {idx}) finite={bool(torch.isfinite(y).all())} "
                  f"[{time.time() - t0:.2f}s]", flush=True)


def fused(mode, x, feat, g, A=None, IcA=None, need_dx=False):
    os.environ["DEDARK_BLUR"] = mode
    y = ops.filters_forward(x, feat, A, IcA)
    dfeat, dx = ops.filters_backward(x, feat, g, A, IcA, need_dx=need_dx)
    torch.cuda.synchronize()
    return y, dfeat, dx


def check_fused():
    for (B, H, W, custom) in [(2, 96, 80, False), (2, 96, 80, True), (3, 200, 132, True), (16, 640, 640, False)]:
        gen = torch.Generator().manual_seed(B * H + W)
        x = torch.rand(B, 3, H, W, generator=gen)
        g = torch.randn(B, 3, H, W, generator=gen)
        feat = torch.randn(B, 15, generator=gen) * 0.8
        A = (0.4 + 0.5 * torch.rand(B, 3, generator=gen)) if custom else None
        IcA = torch.rand(B, 1, H, W, generator=gen) if custom else None
        xd, gd, fd = x.to(dev), g.to(dev), feat.to(dev)
        Ad, Id = (None if A is None else A.to(dev)), (None if IcA is None else IcA.to(dev))
        y_cc, df_cc, dx_cc = fused("cc", xd, fd, gd, Ad, Id, need_dx=True)
        y_tc, df_tc, dx_tc = fused("tc", xd, fd, gd, Ad, Id, need_dx=True)
        line = (f"fused B={B} {H}x{W} custom={custom}: y tc-vs-cc {rel(y_tc, y_cc):.2e}, dfeat tc-vs-cc "
                f"{rel(df_tc, df_cc):.2e}, dx tc-vs-cc {rel(dx_tc, dx_cc):.2e}")
        if B * H * W <= 3 * 200 * 132:  # fp64 truth for the small cases
            xr = x.double().requires_grad_(True)
            fr = feat.double().requires_grad_(True)
            yr = O.filter_chain(xr, fr, None if A is None else A.double(), None if IcA is None else IcA.double(), dense_blur=False)
            yr.backward(g.double())
            line += (f" | vs fp64: y tc {rel(y_tc.cpu(), yr.detach()):.2e} cc {rel(y_cc.cpu(), yr.detach()):.2e}, dfeat tc "
                     f"{rel(df_tc.cpu(), fr.grad):.2e} cc {rel(df_cc.cpu(), fr.grad):.2e}, dx tc {rel(dx_tc.cpu(), xr.grad):.2e} "
                     f"cc {rel(dx_cc.cpu(), xr.grad):.2e}")
        print(line, flush=True)


def timing():
    B, H, W = 16, 640, 640
    gen = torch.Generator(device=dev).manual_seed(1)
    xs = [torch.rand(B, 3, H, W, generator=gen, device=dev) for _ in range(4)]
    gs = [torch.randn(B, 3, H, W, generator=gen, device=dev) for _ in range(4)]
    feat = torch.randn(B, 15, generator=gen, device=dev) * 0.3
    for mode in ("cc", "tc"):
        os.environ["DEDARK_BLUR"] = mode
        for name, fn in (("fwd", lambda i: ops.filters_forward(xs[i % 4], feat)),
                         ("bwd", lambda i: ops.filters_backward(xs[i % 4], feat, gs[i % 4]))):
            for i in range(5):
                fn(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(20):
                fn(i)
            e1.record()
            torch.cuda.synchronize()
            print(f"timing {mode} {name}: {1e3 * e0.elapsed_time(e1) / 20:.1f} us per call (eager, incl. allocation)", flush=True)


if __name__ == "__main__":
    check_blur()
    check_fused()
    if "quick" not in sys.argv:
        timing()


def timing_bf16():
    """Filter forward / backward in the bf16 I/O mode (bf16 x, y, g) at 16x3x640x640."""
    B, H, W = 16, 640, 640
    gen = torch.Generator(device=dev).manual_seed(1)
    xs = [torch.rand(B, 3, H, W, generator=gen, device=dev).to(torch.bfloat16) for _ in range(4)]
    gs = [torch.randn(B, 3, H, W, generator=gen, device=dev).to(torch.bfloat16) for _ in range(4)]
    feat = torch.randn(B, 15, generator=gen, device=dev) * 0.3
    for name, fn in (("fwd", lambda i: ops.filters_forward(xs[i % 4], feat, out_dtype=torch.bfloat16)),
                     ("bwd", lambda i: ops.filters_backward(xs[i % 4], feat, gs[i % 4]))):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        print(f"timing bf16 {name}: {1e3 * e0.elapsed_time(e1) / 20:.1f} us per call (eager, incl. allocation)", flush=True)


if __name__ == "__main__" and "quick" not in sys.argv:
    timing_bf16()
