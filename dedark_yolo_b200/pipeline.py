"""Device-resident training step of the hot path with every buffer pre-allocated.

    synthesis (+recovery loss, + the 256x256 resize in the same pass)  ->  predictor fwd  ->  fused filters fwd
                                ->  fused filters bwd (cotangent g)  ->  predictor bwd
                                ->  [NCCL all-reduce(sum) of the flat predictor gradient when world_size > 1]

This is what ``bench.py`` times: no allocation, no host sync, all launches on the current stream, so a step
can also be captured into a CUDA graph (``capture``).

Software-pipelined mode (``prime`` / ``step_overlapped``): the synthesis (+resize) of batch i+1 does not depend on the
weights, so it is issued on a side stream as soon as the fused filter backward of batch i has been enqueued and runs in
the shadow of the predictor backward -- eight small-grid, latency-bound launches that leave most SMs idle.  dark / r /
rec are double-buffered for it; every step still executes exactly one synthesis and one of everything else.  It shares the parameter tensors of a
``lowlight_recovery`` module; gradients land in one flat fp32 buffer (164 943 floats) whose 14 views are laid
out in state-dict order -- the buffer the single all-reduce runs on.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import PredictorTensors, check, lib
from .llie import lowlight_recovery, ordered_parameters


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


# NVTX ranges around the five stages of a step (SURVEY.md section 5 "tracing"): DEDARK_NVTX=1; off by default (a range costs
# ~1 us of host time, the eager step has ~25 of them)
_NVTX = os.environ.get("DEDARK_NVTX", "0") == "1"


class _Range:
    __slots__ = ("name",)

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


class RecoveryPipeline:
    def __init__(self, module: lowlight_recovery, B: int, H: int, W: int, dark_param: float = 15.0,
                 src_dtype: torch.dtype = torch.float32, device=None, process_group=None, allreduce: bool = False,
                 exchange=None, keep_clean: bool = False, io_dtype: torch.dtype = torch.float32, materialize_dark: bool = True):
        dev = torch.device(device if device is not None else next(module.parameters()).device)
        if dev.type != "cuda":
            raise RuntimeError("RecoveryPipeline needs a CUDA device (no CPU fallback)")
        self.dev, self.B, self.H, self.W, self.p = dev, B, H, W, float(dark_param)
        self.src_dtype = src_dtype
        # bf16 I/O mode: the darkened batch, y and the cotangent g are bf16 (TF32 blur on the tensor cores, 2e-2 gate); r, the
        # predictor, feat / dfeat and the gradients stay fp32
        if io_dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("io_dtype must be float32 or bfloat16")
        self.io_dtype = io_dtype
        self._dt = _lib.DT_BF16 if io_dtype == torch.bfloat16 else _lib.DT_F32
        # uint8 sources: the fp32 clean image (train.py:72, ``batch["clean_img"]``) is only an operand of the recovery loss, which
        # the synthesis pass has already reduced -- it is materialised (one more full-size write) only on request
        self.keep_clean = keep_clean
        self.materialize_dark = bool(materialize_dark)   # validated below (needs fused_resize)
        self.params = [q.detach() for q in ordered_parameters(module.extractor)]
        for q in self.params:
            assert q.is_cuda and q.dtype == torch.float32 and q.is_contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        n_par = sum(q.numel() for q in self.params)
        self.flat_grad = torch.zeros(n_par, **f32)
        self.grads, off = [], 0
        for q in self.params:
            self.grads.append(self.flat_grad[off:off + q.numel()].view(q.shape))
            off += q.numel()
        self._slots = [self._new_slot()]  # [1] is added by enable_overlap()
        self._cur = 0
        self._side = None
        self.acts = torch.empty(_lib.workspace_bytes(_lib.WS_PREDICTOR_ACTS, B) // 4, **f32)
        self.feat = torch.empty(B, 15, **f32)
        self.y = torch.empty(B, 3, H, W, dtype=io_dtype, device=dev)
        self.dfeat = torch.empty(B, 15, **f32)
        self._ws_pb = torch.empty(_lib.workspace_bytes(_lib.WS_PREDICTOR_BWD, B), dtype=torch.uint8, device=dev)
        self._ws_rb = torch.empty(_lib.workspace_bytes(_lib.WS_RECOVERY_BWD, B, H, W), dtype=torch.uint8, device=dev)
        self._w = PredictorTensors.from_tensors(self.params)
        self._g = PredictorTensors.from_tensors(self.grads)
        self.allreduce = allreduce
        self.pg = process_group
        # exchange: a dist.GradExchange (or any object with a ``px`` descriptor).  With it the predictor backward itself
        # exchanges and sums the gradients over peer memory (dd_predictor_bwd_allreduce) and no NCCL call is made.
        self.exchange = exchange
        # uint8 batches: synthesis and the module's 256x256 resize in ONE pass when the size allows it (W % 4 == 0, band fits in
        # shared memory) -- that pass is HBM bound and the fusion saves the re-read of the dark batch (e2e +7 %).  fp32 sources stay
        # on two passes: their synthesis is issue bound on the exact powf and the fused kernel measured 8 us slower than the pair.
        self.fused_resize = src_dtype == torch.uint8 and io_dtype == torch.float32 and bool(lib.dd_synth_resize_supported(H, W))
        # uint8 batch -> synthesis -> filter chain without the darkened fp32 batch in HBM.  The fused
        # synthesis pass then only produces r and the recovery loss, and the filter kernels read the uint8 batch through the
        # 256-entry darkening table (y and the gradients are bit-identical).  The uint8 batch of a step must stay alive and
        # unchanged until that step's backward has run (HostBatchPrefetcher holds a slot exactly that long).
        if not self.materialize_dark:
            if not self.fused_resize:
                raise ValueError("materialize_dark=False needs a uint8 source, fp32 I/O and a size the fused synthesis + resize pass "
                                 "supports (W % 4 == 0)")
            self._dark_tab = torch.empty(256, **f32)
            with torch.cuda.device(dev):
                check(lib.dd_dark_table(self.p, None, _p(self._dark_tab), torch.cuda.current_stream(dev).cuda_stream))
        # experiment switch (profiles/debug/fork_probe.py): fork the next batch's synthesis before the filter backward instead of
        # behind it
        self.fork_before_filters_bwd = False
        # bucketed data-parallel step, experiment switch: DEDARK_DDP_SCHEDULE=behind synthesises the next batch BEHIND the
        # convolution backward (three graphs per step; the synthesis then runs beside the last all-reduce) instead of beside it
        # (two graphs, the default).  Same step time at 2 GPUs (0.4003 ms either way).
        self.synth_behind_conv_bwd = os.environ.get("DEDARK_DDP_SCHEDULE", "beside") == "behind"
        self.capture_stream = None   # stream the overlapped step is captured on (None: torch's capture stream); see enable_overlap
        self.graphs = {}

    # -- per-batch buffers: what the synthesis writes and the rest of the step reads ---------------------
    class _Slot:
        __slots__ = ("clean", "dark", "rec", "r", "ws_syn", "src")

    def _new_slot(self):
        f32 = dict(dtype=torch.float32, device=self.dev)
        B, H, W = self.B, self.H, self.W
        s = self._Slot()
        s.clean = torch.empty(B, 3, H, W, **f32) if (self.src_dtype == torch.uint8 and self.keep_clean) else None
        s.dark = torch.empty(B, 3, H, W, dtype=self.io_dtype, device=self.dev) if self.materialize_dark else None
        s.src = None   # materialize_dark=False: the uint8 batch this slot's step reads
        s.rec = torch.zeros((), **f32)
        s.r = torch.empty(B, 3, 256, 256, **f32)
        s.ws_syn = torch.empty(_lib.workspace_bytes(_lib.WS_SYNTH, B), dtype=torch.uint8, device=self.dev)
        return s

    # the buffers of the batch the next forward/backward works on
    clean = property(lambda self: self._slots[self._cur].clean)
    dark = property(lambda self: self._slots[self._cur].dark)
    rec = property(lambda self: self._slots[self._cur].rec)
    r = property(lambda self: self._slots[self._cur].r)

    # -- individual stages (each is one C-ABI call) -------------------------------------------------------
    def synth(self, src, st, slot=None):
        with _Range("dedark/synthesis+recovery_loss"):
            self._synth(src, st, slot)

    def _synth(self, src, st, slot=None):
        s = self._slots[self._cur if slot is None else slot]
        is_u8 = src.dtype == torch.uint8
        if self.fused_resize:  # also produces r: forward() then skips dd_resize256
            if not self.materialize_dark:
                s.src = src
            check(lib.dd_synth_resize_fwd(_p(src), _lib.SRC_U8 if is_u8 else _lib.SRC_F32, self.p, None, None,
                                          _p(s.clean), _p(s.dark), _p(s.r), _p(s.rec), self.B, self.H,
                                          self.W, _p(s.ws_syn), s.ws_syn.numel(), st))
            return
        check(lib.dd_synth_fwd_ex(_p(src), _lib.SRC_U8 if is_u8 else _lib.SRC_F32, self.p, None, None, _p(s.clean),
                                  _p(s.dark), self._dt, None, _p(s.rec), src.numel(), _p(s.ws_syn), s.ws_syn.numel(), st))

    def resize(self, st, slot=None):
        s = self._slots[self._cur if slot is None else slot]
        if not self.fused_resize:
            check(lib.dd_resize256_ex(_p(s.dark), self._dt, _p(s.r), self.B, self.H, self.W, st))

    def forward(self, st, A=None, IcA=None, resize=True):
        B, H, W = self.B, self.H, self.W
        if resize:
            with _Range("dedark/resize256"):
                self.resize(st)
        with _Range("dedark/predictor_fwd"):
            check(lib.dd_predictor_fwd(_p(self.r), C.byref(self._w), _p(self.acts), _p(self.feat), B, st))
        with _Range("dedark/filters_fwd"):
            self._filters_fwd(st, A, IcA)

    def _filters_fwd(self, st, A, IcA):
        B, H, W = self.B, self.H, self.W
        if not self.materialize_dark:
            check(lib.dd_recovery_fwd_u8(_p(self._slots[self._cur].src), _p(self._dark_tab), _p(A), _p(IcA), _p(self.feat), _p(self.y), B, H, W, st))
            return
        check(lib.dd_recovery_fwd_ex(_p(self.dark), self._dt, _p(A), _p(IcA), _p(self.feat), _p(self.y), self._dt, B, H, W, st))

    def backward_filters(self, g, st, A=None, IcA=None):
        with _Range("dedark/filters_bwd"):
            self._filters_bwd(g, st, A, IcA)

    def _filters_bwd(self, g, st, A=None, IcA=None):
        B, H, W = self.B, self.H, self.W
        if not self.materialize_dark:
            check(lib.dd_recovery_bwd_u8(_p(self._slots[self._cur].src), _p(self._dark_tab), _p(A), _p(IcA), _p(self.feat), _p(g), _p(self.dfeat),
                                         B, H, W, _p(self._ws_rb), self._ws_rb.numel(), st))
            return
        check(lib.dd_recovery_bwd_ex(_p(self.dark), self._dt, _p(A), _p(IcA), _p(self.feat), _p(g), self._dt, _p(self.dfeat), None, B, H, W,
                                     _p(self._ws_rb), self._ws_rb.numel(), st))

    # flat gradient layout (state-dict order): [conv w/b x5 | fc1.w fc1.b fc2.w fc2.b]; the fc tail is final after part 1
    FC_OFFSET = 448 + 4640 + 3 * 9248

    def backward_predictor_part(self, st, part: int):
        """``dd_predictor_bwd_part``: 1 = fully connected layers (their gradients, flat_grad[FC_OFFSET:], are final afterwards),
        2 = the convolutions."""
        check(lib.dd_predictor_bwd_part(_p(self.r), C.byref(self._w), _p(self.acts), _p(self.dfeat), C.byref(self._g), None, self.B,
                                        _p(self._ws_pb), self._ws_pb.numel(), int(part), st))

    def allreduce_fc_async(self):
        """Start the sum over ranks of the fc gradients (80 % of the bytes) on the collective's own stream, ordered after
        everything enqueued on the current stream so far; returns the work handle (``wait()`` it before reading)."""
        return torch.distributed.all_reduce(self.flat_grad[self.FC_OFFSET:], group=self.pg, async_op=True)

    def allreduce_conv_async(self):
        return torch.distributed.all_reduce(self.flat_grad[:self.FC_OFFSET], group=self.pg, async_op=True)

    def backward_predictor(self, st):
        with _Range("dedark/predictor_bwd"):
            self._predictor_bwd(st)

    def _predictor_bwd(self, st):
        B = self.B
        if self.exchange is not None:
            check(lib.dd_predictor_bwd_allreduce(_p(self.r), C.byref(self._w), _p(self.acts), _p(self.dfeat), C.byref(self._g), B,
                                                 _p(self._ws_pb), self._ws_pb.numel(), C.byref(self.exchange.px), st))
        else:
            check(lib.dd_predictor_bwd(_p(self.r), C.byref(self._w), _p(self.acts), _p(self.dfeat), C.byref(self._g), None, B,
                                       _p(self._ws_pb), self._ws_pb.numel(), st))

    def backward(self, g, st, A=None, IcA=None):
        self.backward_filters(g, st, A, IcA)
        self.backward_predictor(st)

    def step(self, src: torch.Tensor, g: torch.Tensor):
        """One full pass: ``src`` is the clean batch (uint8 or fp32 [B,3,H,W]), ``g`` the cotangent dL/dy."""
        assert src.shape == (self.B, 3, self.H, self.W) and g.shape == src.shape and g.dtype == self.io_dtype
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream(self.dev).cuda_stream
            self.synth(src, st)
            self.forward(st)
            self.backward(g, st)
            if self.allreduce and self.exchange is None:
                torch.distributed.all_reduce(self.flat_grad, group=self.pg)
        return self.y, self.rec, self.flat_grad

    # -- software-pipelined steps ---------------------------------------------------------------------
    def enable_overlap(self):
        """Allocate the second set of per-batch buffers and the side stream of ``step_overlapped``."""
        if len(self._slots) == 1:
            self._slots.append(self._new_slot())
        if self._side is None:
            self._side = torch.cuda.Stream(self.dev)
            self._ev_fork, self._ev_join = torch.cuda.Event(), torch.cuda.Event()

    def prime(self, src: torch.Tensor):
        """Synthesise the first batch (current stream); the next ``step_overlapped`` consumes it."""
        self.enable_overlap()
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream(self.dev).cuda_stream
            self.synth(src, st)
            self.resize(st)

    def step_overlapped(self, src_next: Optional[torch.Tensor], g: torch.Tensor):
        """Forward + backward of the batch synthesised by the previous call (or ``prime``), with the synthesis (+resize) of
        ``src_next`` overlapped with the predictor backward on a side stream.  Returns ``(y, rec, flat_grad)`` of the batch
        consumed; ``rec`` is that batch's recovery loss.  ``src_next=None`` ends the sequence (nothing is synthesised)."""
        assert self._side is not None, "call prime() first"
        assert g.shape == (self.B, 3, self.H, self.W) and g.dtype == self.io_dtype
        with torch.cuda.device(self.dev):
            main = torch.cuda.current_stream(self.dev)
            st = main.cuda_stream
            cur = self._cur
            rec = self._slots[cur].rec
            self.forward(st, resize=False)
            early = self.fork_before_filters_bwd and src_next is not None
            if not early:
                self.backward_filters(g, st)
            if src_next is not None:
                assert src_next.shape == (self.B, 3, self.H, self.W)
                self._ev_fork.record(main)
                self._side.wait_event(self._ev_fork)
                sst = self._side.cuda_stream
                self.synth(src_next, sst, slot=cur ^ 1)
                self.resize(sst, slot=cur ^ 1)
                self._ev_join.record(self._side)
            if early:
                self.backward_filters(g, st)
            bucketed = self.allreduce and self.exchange is None
            if bucketed:
                # two buckets: the fc gradients are reduced over the ranks while the convolution backward (and the synthesis
                # of the next batch) still run; only the small conv bucket (131 KB) is exposed at the end of the step
                self.backward_predictor_part(st, 1)
                w1 = self.allreduce_fc_async()
                self.backward_predictor_part(st, 2)
            else:
                self.backward_predictor(st)
            if src_next is not None:
                main.wait_event(self._ev_join)
            if bucketed:
                w2 = self.allreduce_conv_async()
                w1.wait()
                w2.wait()
            self._cur = cur ^ 1
        return self.y, rec, self.flat_grad

    def capture_overlapped(self, key, src_next: torch.Tensor, g: torch.Tensor, slot: int, epilogue=None):
        """Capture ``step_overlapped(src_next, g)`` for the step that consumes buffer set ``slot`` (steps alternate 0, 1, 0, ...;
        the collective stays outside the graph).  ``replay_overlapped(key)`` must be called in that alternation.
        ``epilogue(y, rec, flat_grad)``: optional work captured behind the step in the same graph (reading the step's results
        back to pinned host memory, say) -- only meaningful when no collective has to run between the two."""
        self.enable_overlap()
        if self.allreduce and self.exchange is None:
            return self._capture_overlapped_bucketed(key, src_next, g, slot)
        ar, self.allreduce = self.allreduce, False
        try:
            self._cur = slot
            out = self.step_overlapped(src_next, g)  # warm-up outside capture
            if epilogue is not None:
                epilogue(*out)
            torch.cuda.synchronize(self.dev)
            self._cur = slot
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=self.capture_stream):
                out = self.step_overlapped(src_next, g)
                if epilogue is not None:
                    epilogue(*out)
            self.graphs[key] = (graph, slot)
        finally:
            self.allreduce = ar
        return graph

    def _step_halves(self, src_next, g, half: int):
        """The overlapped step in pieces around the points where a gradient bucket is final: half 0 = forward, filter
        backward, fc backward; then either half 1 = convolution backward with the synthesis of the next batch forked beside it
        (the default; the synthesis fills the GPU, so that fork overlaps little -- DESIGN.md section 6 (xxix) -- and the last
        all-reduce is exposed at the end of the step), or (``synth_behind_conv_bwd``) half 1 = convolution backward and half 2 =
        synthesis + resize of the next batch, which then runs beside the all-reduce of the convolution bucket."""
        main = torch.cuda.current_stream(self.dev)
        st = main.cuda_stream
        if half == 0:
            self.forward(st, resize=False)
            self.backward_filters(g, st)
            self.backward_predictor_part(st, 1)
            return
        if self.synth_behind_conv_bwd:
            if half == 1:
                self.backward_predictor_part(st, 2)
            else:
                self.synth(src_next, st, slot=self._cur ^ 1)
                self.resize(st, slot=self._cur ^ 1)
            return
        self._ev_fork.record(main)
        self._side.wait_event(self._ev_fork)
        sst = self._side.cuda_stream
        self.synth(src_next, sst, slot=self._cur ^ 1)
        self.resize(sst, slot=self._cur ^ 1)
        self._ev_join.record(self._side)
        self.backward_predictor_part(st, 2)
        main.wait_event(self._ev_join)

    def _capture_overlapped_bucketed(self, key, src_next, g, slot):
        """Data-parallel variant of ``capture_overlapped``: two graphs per step with the NCCL all-reduce of the fc bucket issued
        between them (collectives stay outside the graphs: capturing them hung at teardown)."""
        with torch.cuda.device(self.dev):
            graphs = []
            for half in ((0, 1, 2) if self.synth_behind_conv_bwd else (0, 1)):
                self._cur = slot
                self._step_halves(src_next, g, half)  # warm-up outside capture
                torch.cuda.synchronize(self.dev)
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr):
                    self._step_halves(src_next, g, half)
                graphs.append(gr)
            self.graphs[key] = (tuple(graphs), slot)
        return graphs

    def replay_overlapped(self, key):
        graph, slot = self.graphs[key]
        assert slot == self._cur, "overlapped graphs must be replayed in the order they alternate buffer sets"
        rec = self._slots[slot].rec
        if isinstance(graph, tuple):  # bucketed data-parallel step: fc bucket reduced beside the second half
            graph[0].replay()
            w1 = self.allreduce_fc_async()
            graph[1].replay()
            w2 = self.allreduce_conv_async()
            if len(graph) > 2:
                graph[2].replay()   # the next batch's synthesis, beside the all-reduce of the convolution bucket
            w1.wait()
            w2.wait()
        else:
            graph.replay()
            if self.allreduce and self.exchange is None:
                torch.distributed.all_reduce(self.flat_grad, group=self.pg)
        self._cur = slot ^ 1
        return self.y, rec, self.flat_grad

    # -- CUDA graphs ------------------------------------------------------------------------------------
    def capture(self, key, src: torch.Tensor, g: torch.Tensor):
        """Capture ``step(src, g)`` (without the collective) into a CUDA graph replayable with ``replay(key)``."""
        ar, self.allreduce = self.allreduce, False
        try:
            self.step(src, g)  # warm-up outside capture (sets kernel attributes, primes allocations)
            torch.cuda.synchronize(self.dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.step(src, g)
            self.graphs[key] = graph
        finally:
            self.allreduce = ar
        return graph

    def replay(self, key):
        self.graphs[key].replay()
        if self.allreduce and self.exchange is None:
            torch.distributed.all_reduce(self.flat_grad, group=self.pg)
        return self.y, self.rec, self.flat_grad
