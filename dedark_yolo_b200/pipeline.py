"""Device-resident training step of the hot path with every buffer pre-allocated.

    synthesis (+recovery loss, + the 256x256 resize in the same pass)  ->  predictor fwd  ->  fused filters fwd
                                ->  fused filters bwd (cotangent g)  ->  predictor bwd
                                ->  [NCCL all-reduce(sum) of the flat predictor gradient when world_size > 1]

This is what ``bench.py`` times: no allocation, no host sync, all launches on the current stream, so a step
can also be captured into a CUDA graph (``capture``).  It shares the parameter tensors of a
``lowlight_recovery`` module; gradients land in one flat fp32 buffer (164 943 floats) whose 14 views are laid
out in state-dict order -- the buffer the single all-reduce runs on.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import PredictorTensors, check, lib
from .llie import lowlight_recovery


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class RecoveryPipeline:
    def __init__(self, module: lowlight_recovery, B: int, H: int, W: int, dark_param: float = 15.0,
                 src_dtype: torch.dtype = torch.float32, device=None, process_group=None, allreduce: bool = False,
                 exchange=None):
        dev = torch.device(device if device is not None else next(module.parameters()).device)
        if dev.type != "cuda":
            raise RuntimeError("RecoveryPipeline needs a CUDA device (no CPU fallback)")
        self.dev, self.B, self.H, self.W, self.p = dev, B, H, W, float(dark_param)
        self.src_dtype = src_dtype
        self.params = [q.detach() for q in module.extractor.ordered_parameters()]
        for q in self.params:
            assert q.is_cuda and q.dtype == torch.float32 and q.is_contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        n_par = sum(q.numel() for q in self.params)
        self.flat_grad = torch.zeros(n_par, **f32)
        self.grads, off = [], 0
        for q in self.params:
            self.grads.append(self.flat_grad[off:off + q.numel()].view(q.shape))
            off += q.numel()
        self.clean = torch.empty(B, 3, H, W, **f32) if src_dtype == torch.uint8 else None
        self.dark = torch.empty(B, 3, H, W, **f32)
        self.rec = torch.zeros((), **f32)
        self.r = torch.empty(B, 3, 256, 256, **f32)
        self.acts = torch.empty(_lib.workspace_bytes(_lib.WS_PREDICTOR_ACTS, B) // 4, **f32)
        self.feat = torch.empty(B, 15, **f32)
        self.y = torch.empty(B, 3, H, W, **f32)
        self.dfeat = torch.empty(B, 15, **f32)
        self._ws_syn = torch.empty(_lib.workspace_bytes(_lib.WS_SYNTH, B), dtype=torch.uint8, device=dev)
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
        self.fused_resize = src_dtype == torch.uint8 and bool(lib.dd_synth_resize_supported(H, W))
        self.graphs = {}

    # -- individual stages (each is one C-ABI call) -------------------------------------------------------
    def synth(self, src, st):
        is_u8 = src.dtype == torch.uint8
        if self.fused_resize:  # also produces self.r: forward() then skips dd_resize256
            check(lib.dd_synth_resize_fwd(_p(src), _lib.SRC_U8 if is_u8 else _lib.SRC_F32, self.p, None, None,
                                          _p(self.clean) if is_u8 else None, _p(self.dark), _p(self.r), _p(self.rec), self.B, self.H,
                                          self.W, _p(self._ws_syn), self._ws_syn.numel(), st))
            return
        check(lib.dd_synth_fwd(_p(src), _lib.SRC_U8 if is_u8 else _lib.SRC_F32, self.p, None, None, _p(self.clean) if is_u8 else None,
                               _p(self.dark), None, _p(self.rec), src.numel(), _p(self._ws_syn), self._ws_syn.numel(), st))

    def forward(self, st, A=None, IcA=None):
        B, H, W = self.B, self.H, self.W
        if not self.fused_resize:
            check(lib.dd_resize256(_p(self.dark), _p(self.r), B, H, W, st))
        check(lib.dd_predictor_fwd(_p(self.r), C.byref(self._w), _p(self.acts), _p(self.feat), B, st))
        check(lib.dd_recovery_fwd(_p(self.dark), _p(A), _p(IcA), _p(self.feat), _p(self.y), B, H, W, st))

    def backward(self, g, st, A=None, IcA=None):
        B, H, W = self.B, self.H, self.W
        check(lib.dd_recovery_bwd(_p(self.dark), _p(A), _p(IcA), _p(self.feat), _p(g), _p(self.dfeat), None, B, H, W,
                                  _p(self._ws_rb), self._ws_rb.numel(), st))
        if self.exchange is not None:
            check(lib.dd_predictor_bwd_allreduce(_p(self.r), C.byref(self._w), _p(self.acts), _p(self.dfeat), C.byref(self._g), B,
                                                 _p(self._ws_pb), self._ws_pb.numel(), C.byref(self.exchange.px), st))
        else:
            check(lib.dd_predictor_bwd(_p(self.r), C.byref(self._w), _p(self.acts), _p(self.dfeat), C.byref(self._g), None, B,
                                       _p(self._ws_pb), self._ws_pb.numel(), st))

    def step(self, src: torch.Tensor, g: torch.Tensor):
        """One full pass: ``src`` is the clean batch (uint8 or fp32 [B,3,H,W]), ``g`` the cotangent dL/dy."""
        assert src.shape == (self.B, 3, self.H, self.W) and g.shape == src.shape and g.dtype == torch.float32
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream(self.dev).cuda_stream
            self.synth(src, st)
            self.forward(st)
            self.backward(g, st)
            if self.allreduce and self.exchange is None:
                torch.distributed.all_reduce(self.flat_grad, group=self.pg)
        return self.y, self.rec, self.flat_grad

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
