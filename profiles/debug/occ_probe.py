"""Times dd_recovery_fwd / dd_recovery_bwd of an arbitrary build of the library (path given on the command line) on the bench
workload (16x3x640x640 fp32): used to compare experimental builds (other occupancy, ring sizes) with the shipped one without
touching the package.  Results of an experimental build may be numerically wrong on purpose; only the time is read.

    python profiles/debug/occ_probe.py dedark_yolo_b200/lib/libdedark_b200.so build/exp/libX.so ...
"""
import ctypes as C
import os
import sys

import torch


def load(path):
    lib = C.CDLL(path)
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.dd_workspace_bytes.restype = sz
    lib.dd_workspace_bytes.argtypes = [i, i, i, i]
    lib.dd_last_error.restype = C.c_char_p
    lib.dd_recovery_fwd.argtypes = [vp, vp, vp, vp, vp, i, i, i, vp]
    lib.dd_recovery_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, vp, sz, vp]
    if hasattr(lib, "dd_recovery_fwd_ex"):
        lib.dd_recovery_fwd_ex.argtypes = [vp, i, vp, vp, vp, vp, i, i, i, i, vp]
        lib.dd_recovery_bwd_ex.argtypes = [vp, i, vp, vp, vp, vp, i, vp, vp, i, i, i, vp, sz, vp]
    return lib


def timeit(fn, iters=8, warm=5, burst=12):
    """median / best time of one call, measured over bursts of back-to-back calls (the launch queue stays full, so the host's
    per-call overhead does not enter the device time)"""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(burst):
            fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / burst)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    B, H, W = 16, 640, 640
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    xs = [torch.rand(B, 3, H, W, device=dev) for _ in range(3)]   # rotate: 3 x 78.6 MB > L2
    gs = [torch.randn(B, 3, H, W, device=dev) for _ in range(3)]
    feat = torch.randn(B, 15, device=dev) * 0.1
    y = torch.empty(B, 3, H, W, device=dev)
    dfeat = torch.empty(B, 15, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for path in sys.argv[1:]:
        lib = load(path)
        nws = lib.dd_workspace_bytes(3, B, H, W)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        k = [0]

        def fwd():
            k[0] += 1
            e = lib.dd_recovery_fwd(xs[k[0] % 3].data_ptr(), None, None, feat.data_ptr(), y.data_ptr(), B, H, W, st)
            assert e == 0, lib.dd_last_error()

        def bwd():
            k[0] += 1
            e = lib.dd_recovery_bwd(xs[k[0] % 3].data_ptr(), None, None, feat.data_ptr(), gs[k[0] % 3].data_ptr(), dfeat.data_ptr(), None,
                                    B, H, W, ws.data_ptr(), nws, st)
            assert e == 0, lib.dd_last_error()

        for mode in ("cc", "tc"):
            os.environ["DEDARK_BLUR"] = mode
            mf, bf = timeit(fwd)
            mb, bb = timeit(bwd)
            print(f"{path} [fp32, blur {mode}]: fwd median {mf:.1f} us (best {bf:.1f}), bwd+finalize median {mb:.1f} us (best {bb:.1f}); "
                  f"y finite {bool(torch.isfinite(y).all())}, dfeat finite {bool(torch.isfinite(dfeat).all())}", flush=True)
        os.environ["DEDARK_BLUR"] = "cc"
        if hasattr(lib, "dd_recovery_fwd_ex"):   # bf16 I/O mode: x, y, g bf16 (DD_BF16 = 2)
            xb = [t.bfloat16() for t in xs]
            gb = [t.bfloat16() for t in gs]
            yb = torch.empty(B, 3, H, W, device=dev, dtype=torch.bfloat16)

            def fwd16():
                k[0] += 1
                e = lib.dd_recovery_fwd_ex(xb[k[0] % 3].data_ptr(), 2, None, None, feat.data_ptr(), yb.data_ptr(), 2, B, H, W, st)
                assert e == 0, lib.dd_last_error()

            def bwd16():
                k[0] += 1
                e = lib.dd_recovery_bwd_ex(xb[k[0] % 3].data_ptr(), 2, None, None, feat.data_ptr(), gb[k[0] % 3].data_ptr(), 2, dfeat.data_ptr(), None,
                                           B, H, W, ws.data_ptr(), nws, st)
                assert e == 0, lib.dd_last_error()

            mf, bf = timeit(fwd16)
            mb, bb = timeit(bwd16)
            print(f"{path} [bf16 I/O]: fwd median {mf:.1f} us (best {bf:.1f}), bwd+finalize median {mb:.1f} us (best {bb:.1f})", flush=True)


if __name__ == "__main__":
    main()
