"""Do two streams run kernels concurrently on this box?  Two syntheses at 2 CTAs/SM each (DEDARK_SYNTH_CTAS=2) on separate buffers."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from dedark_yolo_b200 import ops
print({k: v for k, v in os.environ.items() if k.startswith(("CUDA", "NCCL", "NVIDIA"))})
dev = torch.device("cuda", 0)
x = [torch.rand(16, 3, 640, 640, device=dev) for _ in range(2)]
s = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
import ctypes as C
from dedark_yolo_b200 import _lib
from dedark_yolo_b200._lib import lib, check
out = [torch.empty_like(x[0]) for _ in range(2)]
def synth(k, st):
    check(lib.dd_synth_fwd(C.c_void_p(x[k].data_ptr()), _lib.SRC_F32, 15.0, None, None, None, C.c_void_p(out[k].data_ptr()), None, None, x[k].numel(), None, 0, st.cuda_stream))
def sinop(k, st):
    with torch.cuda.stream(st):
        torch.sin(x[k], out=out[k])
for name, fn in (("synth", synth), ("torch.sin", sinop)):
    for mode in ("one", "two"):
        ts = []
        for it in range(20):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            main = torch.cuda.current_stream()
            e0.record(main)
            for k in range(2 if mode == "two" else 1):
                s[k].wait_stream(main)
                fn(k, s[k])
            for k in range(2 if mode == "two" else 1):
                main.wait_stream(s[k])
            e1.record(main)
            torch.cuda.synchronize()
            if it >= 5:
                ts.append(e0.elapsed_time(e1) * 1e3)
        print(name, mode, round(sum(ts) / len(ts), 1), "us")
