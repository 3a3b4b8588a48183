#!/usr/bin/env python
"""bench.py -- headline benchmark of the low-light hot path (BASELINE.json: "lowlight_recovery fwd+bwd
images/sec @640^2 and % HBM roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE configs[1]): per GPU, batch 16x3x640x640 fp32 synthetic; one *step* =
    synthesis clean -> dark (= clean ** 15) fused with the recovery-loss reduction
    -> bilinear resize 256^2 -> predictor CNN forward -> fused filter chain forward (y)
    -> fused filter chain backward (cotangent g, read from HBM) -> predictor backward (164 943 gradients)
    -> [N > 1] one NCCL all-reduce(sum) of the flat predictor gradient.
Weak scaling: every rank runs its own 16 images; the only collective is the gradient all-reduce.

Timed numbers
  value       images/s, whole job; inputs resident in HBM; K steps replayed from CUDA graphs; CUDA events on the
              launching stream; barrier + synchronize on both sides; max over ranks.
  e2e         same metric from HOST buffers: the uint8 batch sits in pinned host memory, its H2D copy (double-buffered,
              overlapping the previous step) and the D2H of the recovery loss + gradient norm are inside the timed
              region, every step.  Headline: the C-ABI pipeline object (`RecoveryPipeline`, graph-replayed); beside it
              (`module_api`) the drop-in `lowlight_recovery` nn.Module with autograd, which is bound by Python host work.
  roofline    for the dominant kernel (largest stage time): algorithmic bytes per launch
              / its mean duration, measured with CUDA events around each stage over a second eager pass of K steps.
  cpu_baseline  the UNMODIFIED reference (baseline/_ref, a byte-for-byte copy of its ultralytics package made by
              baseline/install_reference.py) on the host cores, on a bounded sample; `--impl reference` runs the same
              on the full 16-image batch per step.
  reference_gpu_eager  the same unmodified reference, PyTorch eager on this GPU, TF32 off and on (SURVEY.md section 0.1:
              "the bar is PyTorch eager on the same B200").
L2 hygiene: inputs rotate over a ring of sets much larger than the 126 MB L2 (see config.l2).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "lowlight_recovery fwd+bwd images/sec @640x640 (synthesis + recovery loss + fwd + bwd)"
UNIT = "images/s"
B_PER_GPU, H, W = 16, 640, 640
DARK_PARAM = 15.0
N_PIX = H * W
# algorithmic bytes per image
BYTES_FWD = 12 * N_PIX + 12 * N_PIX          # read x, write y (filter-chain kernel only)
BYTES_BWD = 12 * N_PIX + 12 * N_PIX          # read g, read x (no dx)
BYTES_SYNTH = 12 * N_PIX + 12 * N_PIX        # fp32 clean in, dark out
BYTES_RESIZE_TAPS = 3 * 256 * 256 * 4 * 4


def load_traffic(kernel: str):
    """Measured DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture,
    profiles/r02_traffic.json -- r01_traffic.json as second choice --, written from the committed ncu summaries); None when no
    capture exists."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                v = json.load(f)["dram_bytes_per_launch"].get(kernel)
            if v is not None:
                return v
        except Exception:
            continue
    return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _sample(self):
        nv = self.nv
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        for k, bit in names.items():
            if r & bit:
                self.reasons.add(k)

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self._sample()
            except Exception:
                break
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        if self.ok and not self.samples:
            try:
                self._sample()
            except Exception:
                pass
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


# ---------------------------------------------------------------------------------------------------------------
# Reference arms.  The UNMODIFIED reference (baseline/_ref: a byte-for-byte copy of the reference's ultralytics package,
# baseline/install_reference.py) is driven through its own public API -- torch.pow / F.mse_loss as in
# models/yolo/detect/train.py:103,108 and lowlight_recovery.forward (nn/modules/llie.py:17-53) + autograd.  When the copy
# is absent the oracle's fp32 port (same op mix) stands in and the line says kind = "port".
# ---------------------------------------------------------------------------------------------------------------
WORKLOAD = ("configs[1]: lowlight_recovery fwd+bwd + recovery_loss, batch 16x3x640x640 fp32 synthetic per GPU "
            "(synthesis clean**15 + mse, resize, predictor fwd, filter chain fwd, filter chain bwd, predictor bwd; "
            "N > 1: + sum over ranks of the 164943 predictor gradients)")


def shared_config(world: int):
    """The part of `config` both arms print identically (the workload, not how an arm executes it)."""
    return {"workload": WORKLOAD, "global_batch": world * B_PER_GPU, "parallelism": f"dp{world}",
            "inputs": "clean = U[0,1) fp32 [16,3,640,640], cotangent g = N(0,1); module weights torch.manual_seed(0)",
            "l2": "GPU arm: inputs rotate over a ring of 4 (clean, g) sets = 629 MB + 157 MB of outputs per step (L2 = 126 MB), "
                  "no explicit flush; CPU arm: one 16-image set (157 MB, far beyond the host caches)"}


def reference_step_fn(torch, device):
    """Returns (step(clean, g) -> rec, kind).  One call = synthesis + recovery loss + module forward + backward."""
    import torch.nn.functional as F
    from baseline import reference_runtime as R
    if R.available():
        ns = R.load_modules()
        torch.manual_seed(0)
        m = ns.lowlight_recovery(3).to(device).train()

        def step(clean, g):
            for q in m.parameters():
                q.grad = None
            dark = torch.pow(clean, DARK_PARAM)          # train.py:103
            rec = F.mse_loss(dark, clean)                # train.py:108
            y = m(dark)                                  # llie.py:17-53
            y.backward(g)
            return rec
        return step, "reference"
    from oracle import lowlight_oracle as O
    weights = {k: v.to(device) for k, v in O.cast_weights(O.init_weights(0), torch.float32, requires_grad=True).items()}

    def step(clean, g):
        for v in weights.values():
            v.grad = None
        dark = O.synth_darken(clean, DARK_PARAM)
        rec = O.recovery_mse(dark, clean)
        y = O.recovery_forward(dark, weights, dense_blur=True)
        y.backward(g)
        return rec
    return step, "port"


def cpu_inputs(torch, n):
    gen = torch.Generator().manual_seed(1234)
    return torch.rand(n, 3, H, W, generator=gen), torch.randn(n, 3, H, W, generator=gen)


def cpu_baseline(sample_images: int, reps: int):
    """The reference on the host cores, bounded: `reps` steps of `sample_images` images after one warm-up."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = reference_step_fn(torch, torch.device("cpu"))
    clean, g = cpu_inputs(torch, sample_images)
    times = []
    for i in range(reps + 1):
        t0 = time.perf_counter()
        step(clean, g)
        if i > 0 or reps == 0:
            times.append(time.perf_counter() - t0)
    t = statistics.median(times)
    return {
        "value": sample_images / t, "unit": UNIT, "cores": cores, "kind": kind,
        "sample": f"{sample_images} image(s) of the 16x3x640x640 workload per step (pow + mse + lowlight_recovery fwd + bwd, fp32, "
                  f"torch CPU, {cores} threads), median of {len(times)} after 1 warm-up",
    }, t


def run_reference(args):
    """--impl reference: the unmodified reference on all host cores.  A timed step is the FULL configs[1] batch (16 images,
    ~10-30 s on the box's cores); the W warm-up steps run on a 2-image sample (they only page the code and the allocator in),
    so that the default driver invocation (--steps 20 --warmup 5) ends within minutes.  DEDARK_REF_SAMPLE=n overrides the
    images per timed step.  Rank 0 only under torchrun."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = reference_step_fn(torch, torch.device("cpu"))
    sample = int(os.environ.get("DEDARK_REF_SAMPLE", B_PER_GPU))
    clean, g = cpu_inputs(torch, B_PER_GPU)
    for _ in range(args.warmup):
        step(clean[:2], g[:2])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(clean[:sample], g[:sample])
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    what = "full 16-image batch" if sample == B_PER_GPU else f"{sample}-image sample of the 16-image batch"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": shared_config(max(1, args.gpus)),
        "execution": f"{'unmodified reference (baseline/_ref)' if kind == 'reference' else 'oracle port'} on the host CPU, "
                     f"{cores} torch threads, one process; {what} per timed step",
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{what} per step x {args.steps} steps (warm-up steps on 2 images)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def reference_gpu_eager(torch, dev, clean, g, steps=10):
    """SURVEY.md section 0.1 / 8(d) timing (ii): the unmodified reference module, PyTorch eager, on THIS GPU -- the bar a
    PyTorch user starts from.  TF32 off (parity-comparable) and on (cuDNN's default, ~1e-3 accurate).  CUDA events."""
    from baseline import reference_runtime as R
    if not R.available():
        return {"unavailable": "baseline/_ref missing (run baseline/install_reference.py in the build container)"}
    step, _ = reference_step_fn(torch, dev)
    out = {}
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, flag in (("tf32_off", False), ("tf32_on", True)):
            torch.backends.cudnn.allow_tf32 = flag
            torch.backends.cuda.matmul.allow_tf32 = flag
            for _ in range(3):
                step(clean, g)
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step(clean, g)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"ms_per_step": ms, "value": clean.shape[0] / (ms * 1e-3), "unit": UNIT}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    out["what"] = ("unmodified reference (baseline/_ref): torch.pow + F.mse_loss + lowlight_recovery fwd + autograd bwd, PyTorch "
                   f"{torch.__version__} eager on this GPU, same 16x3x640x640 batch, CUDA events over {steps} steps after 3 warm-ups")
    return out


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import dedark_yolo_b200 as dd
    from dedark_yolo_b200.dist import init_from_env, max_over_ranks

    rank, local_rank, world = init_from_env("nccl")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    B = B_PER_GPU
    torch.manual_seed(0)
    module = dd.lowlight_recovery(3).to(dev).train()
    # N > 1: one NCCL all-reduce(sum) of the flat 164 943-float gradient per step.  DEDARK_EXCHANGE=peer selects the exchange
    # fused into the predictor backward over peer memory instead (dd_predictor_bwd_allreduce): measured at parity with NCCL
    # on 2 and 4 GPUs but with one unexplained slow 4-GPU run, so it is opt-in for now.
    exchange, exchange_how = None, "none (single GPU)"
    if world > 1 and os.environ.get("DEDARK_EXCHANGE", "nccl") != "peer":
        exchange_how = ("NCCL all-reduce(sum) of the flat gradient in two buckets: the fc gradients (80 % of the bytes) right after the fc "
                        "backward, beside the convolution backward and the next batch's synthesis; the conv gradients at the end of the step")
    elif world > 1:
        ok = torch.zeros(1, device=dev)
        try:
            from dedark_yolo_b200.dist import GradExchange
            exchange = GradExchange(dev)
            exchange_how = f"fused into the predictor backward over peer memory ({exchange.how}), no collective library call"
            ok += 1
        except Exception as e:  # pragma: no cover
            print(f"[bench] rank {rank}: peer exchange unavailable ({e!r}); using NCCL all-reduce", file=sys.stderr)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() < 1:
            exchange, exchange_how = None, "NCCL all-reduce(sum) of the flat gradient (peer buffers could not be shared)"
    pipe = dd.RecoveryPipeline(module, B, H, W, dark_param=DARK_PARAM, allreduce=world > 1, exchange=exchange)

    # input ring: RING distinct (clean, g) sets; one step touches ~390 MB, so a set is long gone from L2 when reused
    RING = 4
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    cleans = [torch.rand(B, 3, H, W, generator=gen, device=dev) for _ in range(RING)]
    gs = [torch.randn(B, 3, H, W, generator=gen, device=dev) for _ in range(RING)]

    use_graph = not args.no_graph
    overlap = not args.no_overlap

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed_run(step_fn, first):
        """W warm-up + K timed calls of step_fn(n), n counting on from `first`; returns ms for the K steps (max over ranks)."""
        n = first
        for _ in range(max(args.warmup, 3)):
            step_fn(n)
            n += 1
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step_fn(n)
            n += 1
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev), sampler.stop(), n

    # (1) plain steps: synthesis -> ... -> predictor backward back to back on one stream
    launches_per_step = None
    seq_graph = use_graph
    if seq_graph:
        try:
            for i in range(RING):
                n0 = dd.launch_count()
                pipe.capture(i, cleans[i], gs[i])
                launches_per_step = (dd.launch_count() - n0) // 2  # warm-up + captured pass
        except Exception as e:  # pragma: no cover
            print(f"[bench] CUDA graph capture failed ({e!r}); falling back to eager launches", file=sys.stderr)
            seq_graph = use_graph = False
            pipe.graphs.clear()
    if launches_per_step is None:
        n0 = dd.launch_count()
        pipe.step(cleans[0], gs[0])
        launches_per_step = dd.launch_count() - n0

    def plain_step(n):
        if seq_graph:
            pipe.replay(n % RING)
        else:
            pipe.step(cleans[n % RING], gs[n % RING])

    ms_plain, clocks, _ = timed_run(plain_step, 0)

    # (2) software-pipelined steps (the headline): step n consumes the batch synthesised during step n-1 and synthesises batch
    # n+1 on a side stream under its own predictor backward.  One synthesis and one of everything else per step, as in (1).
    ms_total = ms_plain
    if overlap:
        if use_graph:
            for i in range(RING):  # RING is even: step n uses graph n % RING and buffer set n % 2
                pipe.capture_overlapped(("ovl", i), cleans[(i + 1) % RING], gs[i], slot=i % 2)
        pipe._cur = 0
        pipe.prime(cleans[0])

        def overlapped_step(n):
            if use_graph:
                pipe.replay_overlapped(("ovl", n % RING))
            else:
                pipe.step_overlapped(cleans[(n + 1) % RING], gs[n % RING])

        ms_total, clocks, _ = timed_run(overlapped_step, 0)
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- stage timing (eager, CUDA events around each C-ABI call, same stream) -> roofline of the dominant kernel
    stages = ["synth", "resize+predictor_fwd", "filters_fwd", "filters_bwd", "predictor_bwd"]
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(stages) + 1)] for _ in range(args.steps)]
    st = torch.cuda.current_stream(dev).cuda_stream
    import ctypes as C
    from dedark_yolo_b200 import _lib
    from dedark_yolo_b200.pipeline import _p
    for i in range(3):
        pipe.step(cleans[i % RING], gs[i % RING])
    torch.cuda.synchronize(dev)
    for i in range(args.steps):
        src, g = cleans[i % RING], gs[i % RING]
        ev[i][0].record()
        pipe.synth(src, st)  # synthesis + recovery loss (fp32 source: the resize stays a separate pass, see pipeline.py)
        ev[i][1].record()
        if not pipe.fused_resize:
            _lib.check(_lib.lib.dd_resize256(_p(pipe.dark), _p(pipe.r), B, H, W, st))
        _lib.check(_lib.lib.dd_predictor_fwd(_p(pipe.r), C.byref(pipe._w), _p(pipe.acts), _p(pipe.feat), B, st))
        ev[i][2].record()
        _lib.check(_lib.lib.dd_recovery_fwd(_p(pipe.dark), None, None, _p(pipe.feat), _p(pipe.y), B, H, W, st))
        ev[i][3].record()
        _lib.check(_lib.lib.dd_recovery_bwd(_p(pipe.dark), None, None, _p(pipe.feat), _p(g), _p(pipe.dfeat), None, B, H, W,
                                            _p(pipe._ws_rb), pipe._ws_rb.numel(), st))
        ev[i][4].record()
        _lib.check(_lib.lib.dd_predictor_bwd(_p(pipe.r), C.byref(pipe._w), _p(pipe.acts), _p(pipe.dfeat), C.byref(pipe._g), None, B,
                                             _p(pipe._ws_pb), pipe._ws_pb.numel(), st))
        ev[i][5].record()
    torch.cuda.synchronize(dev)
    stage_us = {s: 1e3 * statistics.mean(ev[i][k].elapsed_time(ev[i][k + 1]) for i in range(args.steps))
                for k, s in enumerate(stages)}
    # the two filter kernels alone: bursts of 12 back-to-back launches (the queue stays full, so neither the host's launch work nor
    # the event records of the eager loop above enter), inputs rotating over RING x 78.6 MB buffers (> L2).  dd_recovery_bwd is two
    # launches (main kernel + finalize): its time covers both.
    burst_us = {}
    if cleans[0].dtype == torch.float32:
        def _fwd(k):
            _lib.check(_lib.lib.dd_recovery_fwd(_p(cleans[k]), None, None, _p(pipe.feat), _p(pipe.y), B, H, W, st))

        def _bwd(k):
            _lib.check(_lib.lib.dd_recovery_bwd(_p(cleans[k]), None, None, _p(pipe.feat), _p(gs[k]), _p(pipe.dfeat), None, B, H, W,
                                                _p(pipe._ws_rb), pipe._ws_rb.numel(), st))

        for name, fn in (("filters_fwd", _fwd), ("filters_bwd", _bwd)):
            for i in range(4):
                fn(i % RING)
            ts = []
            for rep in range(5):
                b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                b0.record()
                for i in range(12):
                    fn(i % RING)
                b1.record()
                b1.synchronize()
                ts.append(1e3 * b0.elapsed_time(b1) / 12)
            burst_us[name] = statistics.median(ts)
    peak, peak_src = load_peaks()
    alg = {"filters_fwd": BYTES_FWD * B, "filters_bwd": BYTES_BWD * B, "synth": BYTES_SYNTH * B}
    # whole-step fraction: the quantity the north star targets (22 806 528 B per image fwd+bwd at
    # 640^2 = x + y + resize taps + g + x; + 9 830 400 B when the fp32 synthesis pass is counted) over the timed step
    step_bytes = (BYTES_FWD + BYTES_RESIZE_TAPS + BYTES_BWD) * B
    step_s = ms_per_step * 1e-3
    dominant = max(("filters_fwd", "filters_bwd"), key=lambda s: stage_us[s])
    kernel_us = burst_us.get(dominant, stage_us[dominant])   # the launch(es) alone; the eager stage time is reported beside it
    achieved = alg[dominant] / (kernel_us * 1e-6) / 1e9
    roofline = {
        "bound": "hbm", "kernel": {"filters_fwd": "recovery_fwd_kernel", "filters_bwd": "recovery_bwd_kernel (+finalize)"}[dominant],
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": load_traffic({"filters_fwd": "recovery_fwd_kernel", "filters_bwd": "recovery_bwd_kernel"}[dominant]),
        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg[dominant],
        "kernel_us": kernel_us,
        "kernel_us_how": ("median of 5 bursts of 12 back-to-back launches, CUDA events on the launching stream, inputs rotating over "
                          f"{RING} x {B * 3 * H * W * 4 / 1e6:.1f} MB (> L2)") if dominant in burst_us else "eager stage time",
        "kernel_us_by_kernel": burst_us,
        "frac_eager_stage": alg[dominant] / (stage_us[dominant] * 1e-6) / 1e9 / peak,
        "all_stages_us": stage_us,
        "frac_by_stage": {s: alg[s] / (burst_us.get(s, stage_us[s]) * 1e-6) / 1e9 / peak for s in alg},
        "frac_vs_nominal_8TBs": achieved / 8000.0,
        "step_frac": step_bytes / step_s / 1e9 / peak,
        "step_frac_with_synthesis_bytes": (step_bytes + BYTES_SYNTH * B) / step_s / 1e9 / peak,
        "step_algorithmic_bytes": step_bytes, "step_target_frac": 0.70,
    }

    # ---- e2e: user-facing API, uint8 batch in pinned host memory, H2D + D2H inside the timed region
    cpu_gen = torch.Generator().manual_seed(4321 + rank)
    host_u8 = [torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=cpu_gen) for _ in range(2)]
    host_kind = "pinned (torch pin_memory)"
    if os.environ.get("DEDARK_HOST_WC", "0") == "1":
        # experiment: write-combined pinned staging buffers (cudaHostAllocWriteCombined): the DMA engine's reads
        # of uncached memory need no snoop of the CPU caches
        import ctypes as _C
        _rt = _C.CDLL("libcudart.so.12")
        wc = []
        for t in host_u8:
            ptr = _C.c_void_p()
            rc = _rt.cudaHostAlloc(_C.byref(ptr), _C.c_size_t(t.numel()), _C.c_uint(0x04))
            assert rc == 0, f"cudaHostAlloc(write-combined) failed: {rc}"
            buf = torch.frombuffer((_C.c_uint8 * t.numel()).from_address(ptr.value), dtype=torch.uint8).view(t.shape)
            buf.copy_(t)
            wc.append(buf)
        host_u8 = wc
        host_kind = "pinned, write-combined (cudaHostAllocWriteCombined)"
    else:
        host_u8 = [t.pin_memory() for t in host_u8]
    host_out = torch.empty(2, dtype=torch.float32).pin_memory()
    params = [p for p in module.parameters()]

    prefetch = dd.HostBatchPrefetcher(dev)
    prefetch.submit(host_u8[0])

    def e2e_step(i):
        for p in params:
            p.grad = None
        src = prefetch.get()                      # H2D of this step's batch (issued while the previous step ran)
        prefetch.submit(host_u8[(i + 1) % 2])     # next step's H2D overlaps this step's kernels
        batch = dd.preprocess_batch({"img": src}, dev, dark_param=DARK_PARAM, dedark_FLAG=False)
        y = module(batch["img"])
        y.backward(gs[i % RING])
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        if world > 1:
            dist.all_reduce(flat)
        res = torch.stack([batch["recovery_loss_batch"], flat.norm()])
        host_out.copy_(res, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()  # the step's result is on the host
        return host_out

    e2e_steps = max(3, min(args.steps, 50))
    for i in range(3):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0, dev)

    # the same loop with module.use_cuda_graphs = True: forward and backward launch sequences replayed as CUDA graphs
    module.use_cuda_graphs = True
    for i in range(8):   # first sightings run eagerly, second sightings capture
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_graph_s = max_over_ranks(time.perf_counter() - t0, dev)
    module.use_cuda_graphs = False

    # same loop with the host reading each step's result one step late (the D2H of step i is awaited while step i+1 is
    # already enqueued): host-side launch work then overlaps the kernels of the previous step.  Reported beside the
    # synchronous number, which stays the headline.
    host_out2 = torch.empty(2, 2, dtype=torch.float32).pin_memory()
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_step_lagged(i):
        for p in params:
            p.grad = None
        src = prefetch.get()
        prefetch.submit(host_u8[(i + 1) % 2])
        batch = dd.preprocess_batch({"img": src}, dev, dark_param=DARK_PARAM, dedark_FLAG=False)
        y = module(batch["img"])
        y.backward(gs[i % RING])
        flat = torch.cat([p.grad.reshape(-1) for p in params])
        if world > 1:
            dist.all_reduce(flat)
        host_out2[i % 2].copy_(torch.stack([batch["recovery_loss_batch"], flat.norm()]), non_blocking=True)
        done[i % 2].record()
        if i > 0:
            done[(i - 1) % 2].synchronize()  # result of the previous step is on the host
        return host_out2[(i - 1) % 2]

    for i in range(3):
        e2e_step_lagged(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step_lagged(i)
    barrier()  # includes the last step's result
    e2e_lag_s = max_over_ranks(time.perf_counter() - t0, dev)
    # the same host-to-host step through the C-ABI pipeline object (RecoveryPipeline): the uint8 batch lands in one of the
    # prefetcher's two device slots, the step that reads that slot is a captured CUDA graph (one graph per slot), the
    # result (recovery loss, gradient norm) is copied back and awaited every step.  No autograd / Python per kernel.
    pipe8 = dd.RecoveryPipeline(module, B, H, W, dark_param=DARK_PARAM, src_dtype=torch.uint8, allreduce=False, exchange=exchange)
    pf2 = dd.HostBatchPrefetcher(dev)
    for k in range(2):  # allocate both slots, then capture one graph per slot
        pf2.submit(host_u8[k])
    slots = [pf2.get(), pf2.get()]
    torch.cuda.synchronize(dev)
    res_dev = torch.empty(2, dtype=torch.float32, device=dev)
    # the read-back of the step's result (recovery loss, gradient norm -> pinned host memory) is part of the captured graph when
    # no collective library call has to run between the step and it (1 GPU, or the peer-memory exchange inside the backward)
    readback_in_graph = overlap and (world == 1 or exchange is not None)

    def readback(y, rec, flat):
        res_dev[0].copy_(rec)
        torch.linalg.vector_norm(flat, out=res_dev[1])
        host_out.copy_(res_dev, non_blocking=True)

    if overlap:
        # step i: forward/backward of batch i (synthesised during step i-1) + synthesis of batch i+1 from staging slot (i+1) % 2,
        # whose H2D was submitted during step i-1; the H2D of batch i+2 is submitted right after this step's graph launch.
        for k in range(2):
            pipe8.capture_overlapped(("u8", k), slots[(k + 1) % 2], gs[k], slot=k, epilogue=readback if readback_in_graph else None)
        pipe8._cur = 0
        pf2.submit(host_u8[0])
        pipe8.prime(pf2.get())  # staging slot 0
        pf2.submit(host_u8[1])  # -> staging slot 1
    else:
        for k in range(2):
            pipe8.capture(("u8", k), slots[k], gs[k])
        pf2.submit(host_u8[0])

    if overlap:
        pf2.get()  # orders the stream behind the H2D of batch 1 (staging slot 1), which step 0 synthesises

    def e2e_step_pipeline(i):
        if overlap:
            out = pipe8.replay_overlapped(("u8", i % 2))
            pf2.submit(host_u8[i % 2])  # H2D of batch i+2 -> staging slot i % 2 (read last by the synthesis inside step i-1)
            nxt = pf2.get()             # orders the stream -- the next step's graph -- behind that copy; no host wait
            assert nxt.data_ptr() == slots[i % 2].data_ptr()
        else:
            src = pf2.get()
            assert src.data_ptr() == slots[i % 2].data_ptr()
            pf2.submit(host_u8[(i + 1) % 2])
            pipe8.graphs[("u8", i % 2)].replay()
            out = (pipe8.y, pipe8.rec, pipe8.flat_grad)
        if world > 1 and exchange is None:
            dist.all_reduce(pipe8.flat_grad)
        if not readback_in_graph:
            readback(*out)
        torch.cuda.current_stream(dev).synchronize()  # the step's result is on the host

    for i in range(4):
        e2e_step_pipeline(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(4, 4 + e2e_steps):
        e2e_step_pipeline(i)
    barrier()
    e2e_pipe_s = max_over_ranks(time.perf_counter() - t0, dev)

    # The same loop with the host reading each step's result ONE STEP LATE (it waits for step i-1 while step i runs), as a training
    # loop that logs its loss asynchronously does: every step still copies its batch host->device and its result device->host inside
    # the timed region; what disappears is the host's wake-up + launch turn-around between two steps (reported beside the
    # headline, not instead of it).
    e2e_pipe_lag_s = None
    if overlap and readback_in_graph:
        done = [torch.cuda.Event(), torch.cuda.Event()]
        lag_results = []

        def e2e_step_pipeline_lagged(i):
            pipe8.replay_overlapped(("u8", i % 2))
            done[i % 2].record()
            pf2.submit(host_u8[i % 2])
            nxt = pf2.get()
            assert nxt.data_ptr() == slots[i % 2].data_ptr()
            if i > i_first:
                done[(i - 1) % 2].synchronize()   # step i-1's result is on the host
                lag_results.append(float(host_out[0]))

        i_first = 4 + e2e_steps
        for i in range(i_first, i_first + 4):
            e2e_step_pipeline_lagged(i)
        torch.cuda.synchronize(dev)
        barrier()
        t0 = time.perf_counter()
        i_first = i_first + 4
        for i in range(i_first, i_first + e2e_steps):
            e2e_step_pipeline_lagged(i)
        torch.cuda.current_stream(dev).synchronize()   # the last step's result
        barrier()
        e2e_pipe_lag_s = max_over_ranks(time.perf_counter() - t0, dev)

    # the same step on device-resident uint8 batches with and without the darkened fp32 batch in HBM
    # (plain captured steps, CUDA events; y / gradients are bit-identical, tests/test_gpu_u8_chain.py)
    n2 = {}
    try:
        for mat in (True, False):
            pn = dd.RecoveryPipeline(module, B, H, W, dark_param=DARK_PARAM, src_dtype=torch.uint8, allreduce=False, materialize_dark=mat)
            for k in range(2):
                pn.capture(("n2", k), slots[k], gs[k])
            for i in range(6):
                pn.graphs[("n2", i % 2)].replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(40):
                pn.graphs[("n2", i % 2)].replay()
            e1.record()
            torch.cuda.synchronize(dev)
            n2["with_dark_batch_ms" if mat else "uint8_only_ms"] = e0.elapsed_time(e1) / 40
            del pn
        n2["hbm_bytes_saved_per_step"] = B * 3 * H * W * (4 + 4 + 4 - 1 - 1)   # dark write + two fp32 reads -> two uint8 reads
        n2["what"] = ("plain (not software-pipelined) captured step from a device-resident uint8 batch: synthesis+resize+loss, predictor, "
                      "filter chain fwd+bwd; uint8_only = RecoveryPipeline(materialize_dark=False), the filter kernels read the uint8 "
                      "batch through the darkening table")
    except Exception as e:  # noqa: BLE001  (reported, not fatal: the headline does not depend on this leg)
        n2 = {"error": repr(e)}

    # bf16 I/O mode (north_star "fp32/bf16 batch"): the same plain captured step with the darkened batch, y and the
    # cotangent in bf16 (forward: 1xTF32 tensor-core blur; backward: CUDA-core kernel with bf16 loads), and its filter kernels alone
    bf = {}
    try:
        pb = dd.RecoveryPipeline(module, B, H, W, dark_param=DARK_PARAM, io_dtype=torch.bfloat16, allreduce=False)
        gb = [gs[k].bfloat16() for k in range(2)]
        for k in range(2):
            pb.capture(("bf", k), cleans[k], gb[k])
        for i in range(6):
            pb.graphs[("bf", i % 2)].replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40):
            pb.graphs[("bf", i % 2)].replay()
        e1.record()
        torch.cuda.synchronize(dev)
        bf["ms_per_step_plain"] = e0.elapsed_time(e1) / 40
        stq = torch.cuda.current_stream(dev).cuda_stream
        for name, fn in (("filters_fwd_us", lambda k: pb._filters_fwd(stq, None, None)), ("filters_bwd_us", lambda k: pb._filters_bwd(gb[k], stq))):
            for i in range(3):
                fn(i % 2)
            e0.record()
            for i in range(12):      # back-to-back launches: the queue stays full, host overhead does not enter
                fn(i % 2)
            e1.record()
            torch.cuda.synchronize(dev)
            bf[name] = 1e3 * e0.elapsed_time(e1) / 12
        nb = B * 3 * H * W * (2 + 2)   # bf16 in + bf16 out (forward); bf16 g + bf16 x (backward)
        bf["algorithmic_bytes_per_kernel"] = nb
        bf["frac_of_hbm_peak"] = {"filters_fwd": nb / (bf["filters_fwd_us"] * 1e-6) / 1e9 / peak, "filters_bwd": nb / (bf["filters_bwd_us"] * 1e-6) / 1e9 / peak}
        bf["note"] = "half the bytes of the fp32 mode at about the same kernel times: the filter kernels are FMA-pipe bound, not HBM bound (DESIGN.md 4.5, 4.6)"
        del pb
    except Exception as e:  # noqa: BLE001
        bf = {"error": repr(e)}

    # the pinned-host -> device copy alone (what bounds the overlapped pipeline): 5 copies of one batch, CUDA events
    h2d0, h2d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dev_u8 = torch.empty_like(host_u8[0], device=dev)
    dev_u8.copy_(host_u8[0], non_blocking=True)
    h2d0.record()
    for _ in range(5):
        dev_u8.copy_(host_u8[0], non_blocking=True)
    h2d1.record()
    torch.cuda.synchronize(dev)
    h2d_gbs = 5 * host_u8[0].numel() / (h2d0.elapsed_time(h2d1) * 1e-3) / 1e9
    e2e = {"value": world * B * e2e_steps / e2e_pipe_s, "unit": UNIT, "h2d_bytes_per_step": B * 3 * H * W,
           "d2h_bytes_per_step": 8, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_pipe_s / e2e_steps,
           "api": "HostBatchPrefetcher(uint8 pinned host batch; H2D runs ahead of the step that reads it) -> RecoveryPipeline (the C-ABI "
                  "calls of one step, captured in a CUDA graph per staging slot" + ("; software-pipelined as config.pipelining" if overlap else "")
                  + ") -> D2H(recovery loss, grad norm)" + (" as the last nodes of that graph" if readback_in_graph else "")
                  + " + stream sync, every step",
           "module_api": {"value": world * B * e2e_steps / e2e_s, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                          "api": "HostBatchPrefetcher -> preprocess_batch -> lowlight_recovery(nn.Module) fwd -> autograd bwd -> "
                                 "D2H(recovery loss, grad norm) + stream sync every step (eager launches; host-bound: ~0.4 ms of "
                                 "Python / autograd / launch work per step)"},
           "module_api_cuda_graphs": {"value": world * B * e2e_steps / e2e_graph_s, "ms_per_step": 1e3 * e2e_graph_s / e2e_steps,
                                      "api": "the same loop with lowlight_recovery.use_cuda_graphs = True (the module replays its forward "
                                             "and backward launch sequences as CUDA graphs keyed by the buffers' addresses)"},
           "module_api_result_read_one_step_late": {"value": world * B * e2e_steps / e2e_lag_s, "ms_per_step": 1e3 * e2e_lag_s / e2e_steps},
           "pipeline_result_read_one_step_late": (None if e2e_pipe_lag_s is None else
                                                  {"value": world * B * e2e_steps / e2e_pipe_lag_s, "ms_per_step": 1e3 * e2e_pipe_lag_s / e2e_steps,
                                                   "what": "the headline e2e loop with the host waiting for step i-1's result while step i runs "
                                                           "(same H2D and D2H copies per step; paced by the H2D copy)"}),
           "no_dark_batch_n2": n2,
           "bf16_io_mode": bf,
           "host_buffers": host_kind,
           "h2d_gbs_measured": h2d_gbs, "h2d_ms_per_step_alone": B * 3 * H * W / (h2d_gbs * 1e9) * 1e3}

    if rank == 0:
        cpu, ref_gpu = None, None
        if not args.skip_cpu and world == 1:
            cpu, _ = cpu_baseline(sample_images=2, reps=3)
            del pipe8, pf2, prefetch
            torch.cuda.empty_cache()
            try:
                ref_gpu = reference_gpu_eager(torch, dev, cleans[0], gs[0])
            except Exception as e:  # pragma: no cover
                ref_gpu = {"unavailable": repr(e)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": shared_config(world),
            "execution": {
                "gradient_exchange": exchange_how,
                "launch": "cuda_graph" if use_graph else "eager",
                "pipelining": ("step n = predictor fwd, filters fwd, filters bwd, predictor bwd of batch n + synthesis/resize of batch n+1 on a "
                               "side stream under that predictor bwd (weight-independent data preparation); one synthesis per step")
                              if overlap else "none: the stages of a batch run back to back on one stream",
                "without_pipelining": {"ms_per_step": ms_plain / args.steps, "value": world * B * args.steps / (ms_plain * 1e-3)},
                "e2e_input": "uint8 batch (train.py:72) in pinned host memory",
            },
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step, "roofline": roofline,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if ref_gpu is not None:
            line["reference_gpu_eager"] = ref_gpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-graph", action="store_true", help="launch eagerly instead of replaying CUDA graphs")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-overlap", action="store_true", help="do not overlap the synthesis of batch n+1 with the predictor backward of batch n")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps == 200 and args.warmup == 20:  # defaults are sized for the GPU arm
            args.steps, args.warmup = 5, 1
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
