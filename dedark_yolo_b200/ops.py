"""Thin PyTorch host layer over the C-ABI: tensors in, raw pointers + current stream out.

torch is used for device memory, streams and autograd bookkeeping only; every arithmetic step is a
kernel of ``libdedark_b200.so``.  All functions require CUDA tensors and raise otherwise (there is
no CPU path).
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import PredictorTensors, check, lib

NUM_FEATURES = 15
RESIZE = 256


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dedark_yolo_b200: CUDA tensor required (no CPU fallback in this build)")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


_DT = {torch.float32: _lib.DT_F32, torch.bfloat16: _lib.DT_BF16}


def bf16_io_supported(H: int, W: int) -> bool:
    """Shapes the bf16 I/O kernels accept (16-byte rows; the tensor-core backward needs the borders > one blur radius apart)."""
    return W % 4 == 0 and H >= 14 and W >= 14


def _img(t: torch.Tensor) -> torch.Tensor:
    """An image operand as the kernels read it: contiguous fp32 or bf16 (anything else is promoted to fp32)."""
    t = t.detach()
    if t.dtype not in _DT:
        t = t.to(torch.float32)
    return t.contiguous()


# ---- a1 + a2 ---------------------------------------------------------------------------------------
def synth_forward(src: torch.Tensor, p: float, lut: Optional[torch.Tensor] = None,
                  clean_lut: Optional[torch.Tensor] = None, want_clean: bool = True,
                  want_dark: bool = True, want_u8: bool = False, want_rec: bool = True, dark_dtype=torch.float32):
    """Low-light synthesis + recovery-loss scalar in one pass (train.py:72,79,103,108-109).

    ``src``: uint8 or float32 CUDA tensor of any shape.  ``lut`` / ``clean_lut``: optional 256-entry tables that
    override the device-computed ``pow(k/255, p)`` / ``k/255`` (see ``reference_cpu_tables``).  Returns ``(clean, dark, dark_u8, rec)``; entries not
    requested are None.  For a float32 source ``clean`` is ``src`` itself.  ``dark_dtype=torch.bfloat16`` (bf16 I/O mode) stores the
    darkened image rounded to bf16; the loss is always reduced from the exact fp32 values."""
    _need_cuda(src, lut)
    if dark_dtype not in _DT:
        raise TypeError(f"synth_forward: dark_dtype must be float32 or bfloat16, got {dark_dtype}")
    if src.dtype not in (torch.uint8, torch.float32):
        raise TypeError(f"synth_forward: uint8 or float32 source expected, got {src.dtype}")
    src = src.contiguous()
    is_u8 = src.dtype == torch.uint8
    dev, n = src.device, src.numel()
    with torch.cuda.device(dev):
        clean = torch.empty(src.shape, dtype=torch.float32, device=dev) if (is_u8 and want_clean) else None
        dark = torch.empty(src.shape, dtype=dark_dtype, device=dev) if want_dark else None
        dark_u8 = torch.empty(src.shape, dtype=torch.uint8, device=dev) if want_u8 else None
        rec = torch.empty((), dtype=torch.float32, device=dev) if want_rec else None
        ws_bytes = _lib.workspace_bytes(_lib.WS_SYNTH, 1)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if want_rec else None
        luts = []
        for t in (lut, clean_lut):
            if t is not None:
                _need_cuda(t)
                t = _f32c(t)
                if t.numel() != 256:
                    raise ValueError("lookup tables must have 256 entries")
            luts.append(t)
        lut, clean_lut = luts
        check(lib.dd_synth_fwd_ex(_ptr(src), _lib.SRC_U8 if is_u8 else _lib.SRC_F32, float(p), _ptr(lut), _ptr(clean_lut), _ptr(clean),
                                  _ptr(dark), _DT[dark_dtype], _ptr(dark_u8), _ptr(rec), n, _ptr(ws), ws_bytes if want_rec else 0,
                                  _stream(dev)))
    if not is_u8 and want_clean:
        clean = src
    return clean, dark, dark_u8, rec


def synth_resize_supported(H: int, W: int) -> bool:
    return bool(lib.dd_synth_resize_supported(int(H), int(W)))


def synth_resize_forward(src: torch.Tensor, p: float, lut: Optional[torch.Tensor] = None,
                         clean_lut: Optional[torch.Tensor] = None, want_clean: bool = True, want_rec: bool = True):
    """Synthesis fused with the module's 256x256 bilinear resize (train.py:72,103,108 + llie.py:43) in one pass:
    ``src`` uint8 or float32 ``[B,3,H,W]`` -> ``(clean, dark, r, rec)`` with ``r = resize256(dark)`` bit for bit, without
    reading the dark batch back from HBM.  Needs ``synth_resize_supported(H, W)``."""
    _need_cuda(src, lut, clean_lut)
    if src.dtype not in (torch.uint8, torch.float32) or src.dim() != 4 or src.shape[1] != 3:
        raise TypeError(f"synth_resize_forward: uint8 or float32 [B,3,H,W] expected, got {src.dtype} {tuple(src.shape)}")
    src = src.contiguous()
    B, _, H, W = src.shape
    is_u8 = src.dtype == torch.uint8
    dev = src.device
    with torch.cuda.device(dev):
        clean = torch.empty(src.shape, dtype=torch.float32, device=dev) if (is_u8 and want_clean) else None
        dark = torch.empty(src.shape, dtype=torch.float32, device=dev)
        r = torch.empty(B, 3, RESIZE, RESIZE, dtype=torch.float32, device=dev)
        rec = torch.empty((), dtype=torch.float32, device=dev) if want_rec else None
        ws_bytes = _lib.workspace_bytes(_lib.WS_SYNTH, B)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if want_rec else None
        lut = None if lut is None else _f32c(lut)
        clean_lut = None if clean_lut is None else _f32c(clean_lut)
        check(lib.dd_synth_resize_fwd(_ptr(src), _lib.SRC_U8 if is_u8 else _lib.SRC_F32, float(p), _ptr(lut), _ptr(clean_lut), _ptr(clean),
                                      _ptr(dark), _ptr(r), _ptr(rec), B, H, W, _ptr(ws), ws_bytes if want_rec else 0, _stream(dev)))
    if not is_u8 and want_clean:
        clean = src
    return clean, dark, r, rec


def reference_cpu_tables(p: float, device):
    """The reference's CPU bits for uint8-sourced data: ``clean = k / 255`` (true division) and
    ``dark = torch.pow(clean, p)`` evaluated by torch on the host for k = 0..255 (train.py:72,79 run on CPU).
    Pass them as ``clean_lut`` / ``lut`` to make the GPU synthesis bit-identical to a CPU run of the reference."""
    k = torch.arange(256, dtype=torch.uint8)
    clean = k.float() / 255
    return torch.pow(clean, float(p)).to(device), clean.to(device)


# ---- SURVEY.md section 8(f) N3 ------------------------------------------------------------------------
def dark_prior(src_u8: torch.Tensor, p: float, lut: Optional[torch.Tensor] = None):
    """The dark-channel prior of the darkened batch on the GPU (train.py:42-68,81-97 as intended; ``dd_dark_prior``):
    ``src_u8`` uint8 ``[B,3,H,W]`` -> ``(dedark_A [B,3] in the image's [0,1] units, IcA [B,1,H,W])``, both fp32."""
    _need_cuda(src_u8, lut)
    if src_u8.dtype != torch.uint8 or src_u8.dim() != 4 or src_u8.shape[1] != 3:
        raise TypeError(f"dark_prior: uint8 [B,3,H,W] expected, got {src_u8.dtype} {tuple(src_u8.shape)}")
    src_u8 = src_u8.contiguous()
    B, _, H, W = src_u8.shape
    dev = src_u8.device
    with torch.cuda.device(dev):
        A = torch.empty(B, 3, dtype=torch.float32, device=dev)
        ica = torch.empty(B, 1, H, W, dtype=torch.float32, device=dev)
        ws_bytes = _lib.workspace_bytes(_lib.WS_DARK_PRIOR, B)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        lut = None if lut is None else _f32c(lut)
        check(lib.dd_dark_prior(_ptr(src_u8), float(p), _ptr(lut), _ptr(A), _ptr(ica), B, H, W, _ptr(ws), ws_bytes, _stream(dev)))
    return A, ica


# ---- a4 / a5 ---------------------------------------------------------------------------------------
def resize256(x: torch.Tensor) -> torch.Tensor:
    """[B,3,H,W] fp32 or bf16 -> [B,3,256,256] fp32 (llie.py:43)."""
    _need_cuda(x)
    x = _img(x)
    B, Cc, H, W = x.shape
    assert Cc == 3
    r = torch.empty(B, 3, RESIZE, RESIZE, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.dd_resize256_ex(_ptr(x), _DT[x.dtype], _ptr(r), B, H, W, _stream(x.device)))
    return r


def predictor_forward(r: torch.Tensor, params: Sequence[torch.Tensor]):
    """``r`` [B,3,256,256] -> (feat [B,15], acts workspace).  ``params``: the 14 tensors in state-dict order."""
    _need_cuda(r, *params)
    B = r.shape[0]
    dev = r.device
    acts = torch.empty(_lib.workspace_bytes(_lib.WS_PREDICTOR_ACTS, B) // 4, dtype=torch.float32, device=dev)
    feat = torch.empty(B, NUM_FEATURES, dtype=torch.float32, device=dev)
    w = PredictorTensors.from_tensors(params)
    with torch.cuda.device(dev):
        check(lib.dd_predictor_fwd(_ptr(r), C.byref(w), _ptr(acts), _ptr(feat), B, _stream(dev)))
    return feat, acts


def predictor_backward(r, params, acts, dfeat, need_dr: bool = False, flat_grad: Optional[torch.Tensor] = None):
    """Returns (list of 14 gradient tensors -- views into one flat buffer, dr or None)."""
    _need_cuda(r, acts, dfeat, *params)
    B, dev = r.shape[0], r.device
    sizes = [p.numel() for p in params]
    if flat_grad is None:
        flat_grad = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
    grads, off = [], 0
    for p, n in zip(params, sizes):
        grads.append(flat_grad[off:off + n].view(p.shape))
        off += n
    dr = torch.empty_like(r) if need_dr else None
    ws_bytes = _lib.workspace_bytes(_lib.WS_PREDICTOR_BWD, B)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    w, g = PredictorTensors.from_tensors(params), PredictorTensors.from_tensors(grads)
    with torch.cuda.device(dev):
        check(lib.dd_predictor_bwd(_ptr(r), C.byref(w), _ptr(acts), _ptr(dfeat), C.byref(g), _ptr(dr), B, _ptr(ws),
                                   ws_bytes, _stream(dev)))
    return grads, dr


# ---- a6..a12, a14 ----------------------------------------------------------------------------------
def filters_forward(x, feat, A=None, IcA=None, out_dtype=None) -> torch.Tensor:
    """The fused filter chain.  ``x`` fp32 or bf16; ``out_dtype`` fp32 (default, the reference's output dtype) or bf16.  With a
    bf16 operand the call runs in the bf16 I/O mode (``dd_recovery_fwd_ex``: TF32 blur on the tensor cores, 2e-2 gate)."""
    _need_cuda(x, feat, A, IcA)
    B, _, H, W = x.shape
    y = torch.empty(x.shape, dtype=out_dtype or torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.dd_recovery_fwd_ex(_ptr(x), _DT[x.dtype], _ptr(A), _ptr(IcA), _ptr(feat), _ptr(y), _DT[y.dtype], B, H, W, _stream(x.device)))
    return y


def filters_backward(x, feat, g, A=None, IcA=None, need_dx: bool = False):
    """Returns (dfeat [B,15], dx or None).  ``x`` / ``g`` fp32 or bf16 (bf16 I/O mode when either is bf16); dx has x's dtype."""
    _need_cuda(x, feat, g, A, IcA)
    B, _, H, W = x.shape
    dev = x.device
    dfeat = torch.empty(B, NUM_FEATURES, dtype=torch.float32, device=dev)
    dx = torch.empty_like(x) if need_dx else None
    ws_bytes = _lib.workspace_bytes(_lib.WS_RECOVERY_BWD, B, H, W)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.dd_recovery_bwd_ex(_ptr(x), _DT[x.dtype], _ptr(A), _ptr(IcA), _ptr(feat), _ptr(g), _DT[g.dtype], _ptr(dfeat), _ptr(dx), B, H, W,
                                     _ptr(ws), ws_bytes, _stream(dev)))
    return dfeat, dx


# ---- the chain on the uint8 batch (no darkened fp32 batch in HBM) ----------------------------------------
def dark_table(p: float, device, lut: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The 256 darkened values ``pow(k/255, p)`` as ``dd_synth_fwd`` produces them (``lut`` overrides, e.g. the reference's CPU bits)."""
    dev = torch.device(device)
    with torch.cuda.device(dev):
        t = torch.empty(256, dtype=torch.float32, device=dev)
        lut = None if lut is None else _f32c(lut.to(dev))
        check(lib.dd_dark_table(float(p), _ptr(lut), _ptr(t), _stream(dev)))
    return t


def _u8_batch(src):
    _need_cuda(src)
    if src.dtype != torch.uint8 or src.dim() != 4 or src.shape[1] != 3:
        raise TypeError(f"uint8 [B,3,H,W] expected, got {src.dtype} {tuple(src.shape)}")
    return src.contiguous()


def filters_forward_u8(src_u8, table, feat, A=None, IcA=None) -> torch.Tensor:
    """``filters_forward(table[src_u8], feat, A, IcA)`` without materialising the darkened batch (``dd_recovery_fwd_u8``)."""
    src_u8 = _u8_batch(src_u8)
    _need_cuda(table, feat, A, IcA)
    B, _, H, W = src_u8.shape
    y = torch.empty(src_u8.shape, dtype=torch.float32, device=src_u8.device)
    with torch.cuda.device(src_u8.device):
        check(lib.dd_recovery_fwd_u8(_ptr(src_u8), _ptr(table), _ptr(A), _ptr(IcA), _ptr(feat), _ptr(y), B, H, W, _stream(src_u8.device)))
    return y


def filters_backward_u8(src_u8, table, feat, g, A=None, IcA=None) -> torch.Tensor:
    """dfeat of ``filters_backward(table[src_u8], ...)`` (``dd_recovery_bwd_u8``; a uint8 source has no dx)."""
    src_u8 = _u8_batch(src_u8)
    _need_cuda(table, feat, g, A, IcA)
    B, _, H, W = src_u8.shape
    dev = src_u8.device
    dfeat = torch.empty(B, NUM_FEATURES, dtype=torch.float32, device=dev)
    ws_bytes = _lib.workspace_bytes(_lib.WS_RECOVERY_BWD, B, H, W)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.dd_recovery_bwd_u8(_ptr(src_u8), _ptr(table), _ptr(A), _ptr(IcA), _ptr(feat), _ptr(_f32c(g)), _ptr(dfeat), B, H, W,
                                     _ptr(ws), ws_bytes, _stream(dev)))
    return dfeat


def debug_blur_tc(x: torch.Tensor, x3: bool = True) -> torch.Tensor:
    """Unit-test hook: the reflect-padded 25x25 Gaussian of ``x`` through the tensor-core engine (``dd_debug_blur_tc``)."""
    _need_cuda(x)
    x = _f32c(x)
    B, Cc, H, W = x.shape
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(lib.dd_debug_blur_tc(_ptr(x), _ptr(y), B * Cc // 3, H, W, int(bool(x3)), _stream(x.device)))
    return y


def resize256_backward_(dr: torch.Tensor, dx: torch.Tensor) -> None:
    """dx += resize^T(dr), in place."""
    B, _, H, W = dx.shape
    with torch.cuda.device(dx.device):
        check(lib.dd_resize256_bwd(_ptr(dr), _ptr(dx), B, H, W, _stream(dx.device)))


def check_image_shape(x: torch.Tensor) -> None:
    """Raise what the reference raises for the same input (SURVEY.md section 8(b) 'Error conventions')."""
    if x.dim() != 4:
        raise RuntimeError(f"lowlight_recovery expects a 4-D NCHW tensor, got {tuple(x.shape)}")
    B, Cc, H, W = x.shape
    if Cc != 3:
        raise RuntimeError(f"lowlight_recovery expects 3 channels (common.py:59 Conv2d(3, 16)), got {Cc}")
    if x.dtype == torch.float64:
        raise RuntimeError("lowlight_recovery: float64 input is not supported (filtersB.py:155 builds a float32 kernel)")
    if W < 3:
        raise IndexError(f"index 2 is out of bounds for dimension 3 with size {W} (rgb2lum, util_filters.py:270-273)")
    if H <= 12 or W <= 12:
        raise RuntimeError(f"reflect padding of 12 needs H, W > 12, got {H} x {W} (filtersB.py:167)")


# ---- the module's autograd node ---------------------------------------------------------------------------------------
_SIZES = (432, 16, 4608, 32, 9216, 32, 9216, 32, 9216, 32, 131072, 64, 960, 15)   # the 14 tensors, state-dict order
_SHAPES = ((16, 3, 3, 3), (16,), (32, 16, 3, 3), (32,), (32, 32, 3, 3), (32,), (32, 32, 3, 3), (32,), (32, 32, 3, 3), (32,),
           (64, 2048), (64,), (15, 64), (15,))
_OFFS = tuple(sum(_SIZES[:k]) for k in range(15))
_MAX_PLANS = 4


class _Plan:
    """Buffers of one (device, B, H, W) call shape of one module instance, reused from call to call so that a forward +
    backward costs a handful of C-ABI calls and two allocations (y and the flat gradient) of host work.  ``version``
    counts the forwards that have written the activation buffers: a backward whose forward is no longer the latest one
    (two forwards before a backward) recomputes resize + predictor forward first -- correctness never depends on the cache."""
    __slots__ = ("r", "acts", "feat", "dfeat", "ws_pb", "ws_rb", "version", "wkey", "w", "tick", "graphs", "seen", "cap_stream", "captures", "replays", "graphs_off")

    def __init__(self, dev, B, H, W):
        f32 = dict(dtype=torch.float32, device=dev)
        self.r = torch.empty(B, 3, RESIZE, RESIZE, **f32)
        self.acts = torch.empty(_lib.workspace_bytes(_lib.WS_PREDICTOR_ACTS, B) // 4, **f32)
        self.feat = torch.empty(B, NUM_FEATURES, **f32)
        self.dfeat = torch.empty(B, NUM_FEATURES, **f32)
        self.ws_pb = self.ws_rb = None   # allocated by the first backward (inference never needs them)
        self.version = 0
        self.wkey, self.w = None, None
        self.tick = 0
        self.graphs, self.seen, self.cap_stream = {}, {}, None   # CUDA-graph replay of the launch sequences (module.use_cuda_graphs)
        self.captures = self.replays = 0
        self.graphs_off = False   # set when captures do not pay off (addresses that do not recur)


_PLANS = weakref.WeakKeyDictionary()   # module instance -> {(device index, B, H, W): _Plan}
_TICK = [0]


def _plan_for(owner, dev, B, H, W) -> _Plan:
    plans = _PLANS.get(owner)
    if plans is None:
        plans = _PLANS[owner] = {}
    key = (dev.index, B, H, W)
    pl = plans.get(key)
    if pl is None:
        if len(plans) >= _MAX_PLANS:   # drop the least recently used shape (validation batches, ragged last batches)
            del plans[min(plans, key=lambda k: plans[k].tick)]
        with torch.cuda.device(dev):
            pl = plans[key] = _Plan(dev, B, H, W)
    _TICK[0] += 1
    pl.tick = _TICK[0]
    return pl


def _weights_struct(pl: _Plan, pd):
    key = tuple(t.data_ptr() for t in pd)
    if pl.wkey != key:
        pl.wkey, pl.w = key, PredictorTensors.from_tensors(pd)
    return pl.w


def _grad_struct(base_ptr: int) -> PredictorTensors:
    g = PredictorTensors()
    for i in range(5):
        g.conv_w[i] = base_ptr + 4 * _OFFS[2 * i]
        g.conv_b[i] = base_ptr + 4 * _OFFS[2 * i + 1]
    g.fc1_w, g.fc1_b, g.fc2_w, g.fc2_b = (base_ptr + 4 * _OFFS[k] for k in (10, 11, 12, 13))
    return g


def _ready(t: torch.Tensor, dev) -> bool:
    return t.device == dev and t.dtype == torch.float32 and t.is_contiguous()


def _ready_img(t: torch.Tensor, dev) -> bool:
    """An image the kernels can read in place: fp32, or bf16 on a shape the bf16 I/O kernels accept."""
    if t.device != dev or not t.is_contiguous():
        return False
    return t.dtype == torch.float32 or (t.dtype == torch.bfloat16 and bf16_io_supported(t.shape[2], t.shape[3]))


_MAX_GRAPHS = 24


def _launch(pl: _Plan, device, key, fn, use_graphs: bool):
    """Run ``fn()`` (a fixed sequence of C-ABI launches on the current stream) -- directly, or, with ``use_graphs``, as the replay
    of a CUDA graph captured the second time the same ``key`` (every pointer and shape the launches depend on) is seen.  One
    replay costs the host ~10 us where the 9-10 launches of a forward or backward cost ~40.  Nothing in ``fn`` allocates."""
    if not use_graphs or pl.graphs_off or torch.cuda.is_current_stream_capturing():
        fn()
        return
    g = pl.graphs.get(key)
    if g is None:
        if pl.captures >= 8 and pl.replays < 2 * pl.captures:
            # a capture costs about a millisecond; when the buffers' addresses do not recur (a caller that allocates fresh tensors
            # from a growing pool every step) replaying never amortises it: go back to plain launches for this call shape
            pl.graphs_off = True
            pl.graphs.clear()
            fn()
            return
        n = pl.seen.get(key, 0)
        if n == 0:                                   # first sighting: plain launches (also sets the kernels' attributes)
            if len(pl.seen) > 4 * _MAX_GRAPHS:
                pl.seen.clear()
            pl.seen[key] = 1
            fn()
            return
        if len(pl.graphs) >= _MAX_GRAPHS:            # pointers that never repeat (no caching allocator?): stop capturing
            pl.graphs.pop(next(iter(pl.graphs)))
        if pl.cap_stream is None:
            pl.cap_stream = torch.cuda.Stream(device)
        cur = torch.cuda.current_stream(device)
        pl.cap_stream.wait_stream(cur)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(pl.cap_stream):
            g.capture_begin()
            try:
                fn()
            finally:
                g.capture_end()
        cur.wait_stream(pl.cap_stream)
        pl.graphs[key] = g
        pl.captures += 1
    else:
        pl.replays += 1
    g.replay()


class RecoveryFunction(torch.autograd.Function):
    """lowlight_recovery.forward as one autograd node: resize -> predictor -> fused filter chain.

    Inputs may live on the CPU (DetectionModel.__init__ probes the model with CPU zeros, nn/tasks.py:290-291):
    they are staged through ``device`` and the result is returned on x's device.  The work itself always runs on
    the GPU through the C-ABI.  ``owner`` (the module instance) keys the buffer cache, see ``_Plan``."""

    @staticmethod
    def forward(ctx, owner, device, x, A, IcA, *params):
        # the reference propagates gradients into dedark_A / IcA through DeDarkFilter (filtersB.py:211-214); no caller of the
        # module asks for them (they are data, train.py:95-96) and the fused backward does not produce them: refuse loudly
        # rather than return a silent None
        if (A is not None and A.requires_grad) or (IcA is not None and IcA.requires_grad):
            raise NotImplementedError("lowlight_recovery (dedark_yolo_b200): gradients w.r.t. dedark_A / IcA are not implemented; "
                                      "detach them (the reference trainer passes constants)")
        out_dev = x.device
        # bf16 input on a supported shape: read in place (bf16 I/O mode); anything else is promoted to fp32 as in the reference
        # (a bf16 x that requires grad -- ad-hoc tests only, x is data in training -- takes the fp32 route: the resize adjoint
        # accumulates into an fp32 dx)
        in_place = _ready_img(x, device) and not (x.dtype == torch.bfloat16 and ctx.needs_input_grad[2])
        xd = x.detach() if in_place else _f32c(x.to(device))
        B, _, H, W = xd.shape
        out_dtype = getattr(owner, "out_dtype", None)
        if out_dtype not in (None, torch.float32, torch.bfloat16):
            raise TypeError(f"lowlight_recovery.out_dtype must be None, float32 or bfloat16, got {out_dtype}")
        if out_dtype == torch.bfloat16 and not bf16_io_supported(H, W):
            out_dtype = None  # shapes the bf16 kernels do not take: fp32 kernels, fp32 output
        Ad = None if A is None else _f32c(A.to(device)).reshape(B, 3)
        Id = None if IcA is None else _f32c(IcA.to(device)).expand(B, 1, H, W).contiguous()
        pd = [q.detach() if _ready(q, device) else _f32c(q.to(device)) for q in params]
        pl = _plan_for(owner, device, B, H, W)
        # preprocess_batch computes the 256x256 resize in the synthesis pass itself (dd_synth_resize_fwd) and leaves it on the
        # darkened tensor; it is valid while that tensor has not been written since
        stash = getattr(x, "_dd_resize256", None) if in_place else None
        r = pl.r
        if stash is not None and stash[1] == x._version and stash[0].device == device and tuple(stash[0].shape) == (B, 3, RESIZE, RESIZE):
            r = stash[0]
        use_graphs = bool(getattr(owner, "use_cuda_graphs", False))
        prev = torch.cuda.current_device()
        if prev != device.index:  # kernels launch on the current device: switch for the duration of the call only
            torch.cuda.set_device(device)
        try:
            w = _weights_struct(pl, pd)
            y = torch.empty(xd.shape, dtype=out_dtype or torch.float32, device=device)
            px, pr, pA, pI, py, dtx, dty = _ptr(xd), _ptr(r), _ptr(Ad), _ptr(Id), _ptr(y), _DT[xd.dtype], _DT[y.dtype]
            resize = r is pl.r

            def run():
                st = _stream(device)
                if resize:
                    check(lib.dd_resize256_ex(px, dtx, pr, B, H, W, st))
                check(lib.dd_predictor_fwd(pr, C.byref(w), _ptr(pl.acts), _ptr(pl.feat), B, st))
                check(lib.dd_recovery_fwd_ex(px, dtx, pA, pI, _ptr(pl.feat), py, dty, B, H, W, st))

            key = ("f", xd.data_ptr(), r.data_ptr(), y.data_ptr(), 0 if Ad is None else Ad.data_ptr(), 0 if Id is None else Id.data_ptr(),
                   pl.wkey, dtx, dty)
            _launch(pl, device, key, run, use_graphs)
        finally:
            if prev != device.index:
                torch.cuda.set_device(prev)
        pl.version += 1
        ctx.plan, ctx.version, ctx.st_dev = pl, pl.version, device
        ctx.r = r                      # the resized batch the predictor read (the plan's buffer, or preprocess_batch's)
        ctx.use_graphs = use_graphs
        ctx.save_for_backward(xd, Ad, Id, *pd)
        ctx.out_dev = out_dev
        ctx.param_meta = None if all(q.device == device and q.dtype == torch.float32 for q in params) \
            else [(q.device, q.dtype) for q in params]
        return y if out_dev == device else y.to(out_dev)

    @staticmethod
    def backward(ctx, g):
        xd, Ad, Id, *pd = ctx.saved_tensors
        pl, dev = ctx.plan, ctx.st_dev
        B, _, H, W = xd.shape
        need_dx = ctx.needs_input_grad[2]
        gd = g if _ready_img(g, dev) else _f32c(g.to(dev))
        prev = torch.cuda.current_device()
        if prev != dev.index:
            torch.cuda.set_device(dev)
        try:
            return RecoveryFunction._backward(ctx, pl, dev, xd, Ad, Id, pd, gd, need_dx, B, H, W)
        finally:
            if prev != dev.index:
                torch.cuda.set_device(prev)

    @staticmethod
    def _backward(ctx, pl, dev, xd, Ad, Id, pd, gd, need_dx, B, H, W):
        st = _stream(dev)
        w = _weights_struct(pl, pd)
        r = ctx.r
        if pl.version != ctx.version:  # another forward of this shape ran in between: its activations replaced ours
            if r is pl.r:
                check(lib.dd_resize256_ex(_ptr(xd), _DT[xd.dtype], _ptr(pl.r), B, H, W, st))
            check(lib.dd_predictor_fwd(_ptr(r), C.byref(w), _ptr(pl.acts), _ptr(pl.feat), B, st))
            pl.version += 1
            ctx.version = pl.version
        if pl.ws_pb is None:
            pl.ws_pb = torch.empty(_lib.workspace_bytes(_lib.WS_PREDICTOR_BWD, B), dtype=torch.uint8, device=dev)
            pl.ws_rb = torch.empty(_lib.workspace_bytes(_lib.WS_RECOVERY_BWD, B, H, W), dtype=torch.uint8, device=dev)
        flat = torch.empty(_OFFS[14], dtype=torch.float32, device=dev)
        dx = torch.empty_like(xd) if need_dx else None
        dr = torch.empty_like(r) if need_dx else None
        gs = _grad_struct(flat.data_ptr())
        px, pA, pI, pg, pdx, pdr, dtx, dtg = _ptr(xd), _ptr(Ad), _ptr(Id), _ptr(gd), _ptr(dx), _ptr(dr), _DT[xd.dtype], _DT[gd.dtype]

        def run():
            s2 = _stream(dev)
            check(lib.dd_recovery_bwd_ex(px, dtx, pA, pI, _ptr(pl.feat), pg, dtg, _ptr(pl.dfeat), pdx, B, H, W, _ptr(pl.ws_rb), pl.ws_rb.numel(), s2))
            check(lib.dd_predictor_bwd(_ptr(r), C.byref(w), _ptr(pl.acts), _ptr(pl.dfeat), C.byref(gs), pdr, B, _ptr(pl.ws_pb), pl.ws_pb.numel(), s2))
            if need_dx:
                check(lib.dd_resize256_bwd(pdr, pdx, B, H, W, s2))

        key = ("b", xd.data_ptr(), r.data_ptr(), gd.data_ptr(), flat.data_ptr(), 0 if Ad is None else Ad.data_ptr(),
               0 if Id is None else Id.data_ptr(), pl.wkey, dtx, dtg)
        _launch(pl, dev, key, run, ctx.use_graphs and not need_dx)
        if need_dx:
            dx = dx.to(ctx.out_dev)
        grads = [t.view(sh) for t, sh in zip(flat.split(_SIZES), _SHAPES)]
        if ctx.param_meta is not None:
            grads = [gr.to(device=d, dtype=t) for gr, (d, t) in zip(grads, ctx.param_meta)]
        return (None, None, dx, None, None, *grads)
