"""Timeline of the software-pipelined step: predictor backward (main stream) vs synthesis+resize (side stream), alone and together.
CUDA events on each stream; eager launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import dedark_yolo_b200 as dd

CORUN = os.environ.get("CORUN", "pred_bwd")
dev = torch.device("cuda", 0)
B, H, W = 16, 640, 640
m = dd.lowlight_recovery(3).to(dev).train()
pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=15.0)
pipe.enable_overlap()
gen = torch.Generator(device=dev).manual_seed(1)
cleans = [torch.rand(B, 3, H, W, generator=gen, device=dev) for _ in range(4)]
gs = [torch.randn(B, 3, H, W, generator=gen, device=dev) for _ in range(4)]
main = torch.cuda.current_stream(dev)
side = pipe._side
E = lambda: torch.cuda.Event(enable_timing=True)

def run(mode, iters=30):
    acc = {}
    for it in range(iters + 5):
        src, g = cleans[it % 4], gs[it % 4]
        st = main.cuda_stream
        pipe._cur = 0
        pipe.synth(src, st); pipe.resize(st)
        pipe.forward(st, resize=False)
        pipe.backward_filters(g, st)
        e = {k: E() for k in ("f", "p0", "p1", "s0", "s1", "end")}
        e["f"].record(main)
        if mode in ("both", "synth"):
            side.wait_event(e["f"])
            e["s0"].record(side)
            pipe.synth(cleans[(it + 1) % 4], side.cuda_stream, slot=1)
            pipe.resize(side.cuda_stream, slot=1)
            e["s1"].record(side)
        if mode in ("both", "pred"):
            e["p0"].record(main)
            if CORUN == "pred_bwd":
                pipe.backward_predictor(st)
            elif CORUN == "pred_fwd":
                pipe.forward(st, resize=False)
            elif CORUN == "pfwd_only":
                import ctypes as C
                from dedark_yolo_b200 import _lib
                from dedark_yolo_b200.pipeline import _p
                _lib.check(_lib.lib.dd_predictor_fwd(_p(pipe.r), C.byref(pipe._w), _p(pipe.acts), _p(pipe.feat), B, st))
            elif CORUN == "resize":
                for _ in range(5):
                    pipe.resize(st, slot=0)
            elif CORUN == "torch":
                for _ in range(5):
                    torch.sin(pipe.y, out=pipe.y)
            e["p1"].record(main)
        if mode in ("both", "synth"):
            main.wait_event(e["s1"])
        e["end"].record(main)
        torch.cuda.synchronize()
        if it >= 5:
            for a, b, name in (("p0", "p1", "pred_bwd"), ("s0", "s1", "synth+resize"), ("f", "end", "total"), ("f", "s0", "fork->synth start")):
                try:
                    v = e[a].elapsed_time(e[b]) * 1e3
                except Exception:
                    continue
                acc.setdefault(name, []).append(v)
    print(mode, {k: round(sum(v) / len(v), 1) for k, v in acc.items()})

for mode in ("pred", "synth", "both"):
    run(mode)
