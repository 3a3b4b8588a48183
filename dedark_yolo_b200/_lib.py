"""ctypes binding of ``libdedark_b200.so`` (the C-ABI declared in ``include/dedark_b200.h``).

There is deliberately no fallback: if the library is missing this module raises at import, and every
wrapper raises if a call fails.  Nothing here touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

DD_OK, DD_ERR_INVALID, DD_ERR_REFLECT_PAD, DD_ERR_WIDTH_LT3, DD_ERR_WORKSPACE, DD_ERR_CUDA = range(6)
WS_SYNTH, WS_PREDICTOR_ACTS, WS_PREDICTOR_BWD, WS_RECOVERY_BWD, WS_DARK_PRIOR = range(5)
SRC_U8, SRC_F32 = 0, 1
DT_F32, DT_BF16 = 1, 2   # element types of the *_ex entry points

EXPORTS = (
    "dd_version", "dd_last_error", "dd_launch_count", "dd_workspace_bytes", "dd_synth_fwd", "dd_resize256",
    "dd_resize256_bwd", "dd_predictor_fwd", "dd_predictor_bwd", "dd_recovery_fwd", "dd_recovery_bwd",
    "dd_synth_resize_fwd", "dd_synth_resize_supported", "dd_predictor_bwd_allreduce", "dd_predictor_bwd_part", "dd_debug_blur_tc", "dd_synth_fwd_ex", "dd_resize256_ex",
    "dd_recovery_fwd_ex", "dd_recovery_bwd_ex", "dd_dark_prior", "dd_dark_table", "dd_recovery_fwd_u8", "dd_recovery_bwd_u8", "dd_exchange_bytes",
)
MAX_PEERS = 8


class PredictorTensors(C.Structure):
    """``dd_predictor_tensors``: 14 device pointers in the reference's state-dict order."""
    _fields_ = [("conv_w", C.c_void_p * 5), ("conv_b", C.c_void_p * 5), ("fc1_w", C.c_void_p),
                ("fc1_b", C.c_void_p), ("fc2_w", C.c_void_p), ("fc2_b", C.c_void_p)]

    @classmethod
    def from_tensors(cls, tensors):
        """``tensors``: the 14 tensors in state-dict order (conv0.w, conv0.b, ..., fc1.w, fc1.b, fc2.w, fc2.b)."""
        t = cls()
        for i in range(5):
            t.conv_w[i] = tensors[2 * i].data_ptr()
            t.conv_b[i] = tensors[2 * i + 1].data_ptr()
        t.fc1_w, t.fc1_b, t.fc2_w, t.fc2_b = (tensors[k].data_ptr() for k in (10, 11, 12, 13))
        return t


class PeerExchange(C.Structure):
    """``dd_peer_exchange``: this rank, the world size and every rank's exchange buffer as addressed from this process."""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("buf", C.c_void_p * MAX_PEERS)]

    @classmethod
    def from_pointers(cls, rank: int, ptrs):
        if not (1 <= len(ptrs) <= MAX_PEERS and 0 <= rank < len(ptrs)):
            raise ValueError(f"peer exchange: rank {rank} / world {len(ptrs)} (at most {MAX_PEERS} peers)")
        px = cls()
        px.rank, px.world = rank, len(ptrs)
        for i, q in enumerate(ptrs):
            px.buf[i] = int(q)
        return px


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python dedark_yolo_b200/build.py` (needs nvcc). "
            "dedark_yolo_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i, f, ll, sz = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_size_t
    lib.dd_version.restype = i
    lib.dd_last_error.restype = C.c_char_p
    lib.dd_launch_count.restype = C.c_ulonglong
    lib.dd_workspace_bytes.restype = sz
    lib.dd_workspace_bytes.argtypes = [i, i, i, i]
    lib.dd_synth_fwd.argtypes = [vp, i, f, vp, vp, vp, vp, vp, vp, ll, vp, sz, vp]
    lib.dd_resize256.argtypes = [vp, vp, i, i, i, vp]
    lib.dd_resize256_bwd.argtypes = [vp, vp, i, i, i, vp]
    lib.dd_predictor_fwd.argtypes = [vp, C.POINTER(PredictorTensors), vp, vp, i, vp]
    lib.dd_predictor_bwd.argtypes = [vp, C.POINTER(PredictorTensors), vp, vp, C.POINTER(PredictorTensors), vp, i, vp, sz, vp]
    lib.dd_predictor_bwd_part.argtypes = [vp, C.POINTER(PredictorTensors), vp, vp, C.POINTER(PredictorTensors), vp, i, vp, sz, i, vp]
    lib.dd_recovery_fwd.argtypes = [vp, vp, vp, vp, vp, i, i, i, vp]
    lib.dd_recovery_bwd.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, vp, sz, vp]
    lib.dd_synth_resize_fwd.argtypes = [vp, i, f, vp, vp, vp, vp, vp, vp, i, i, i, vp, sz, vp]
    lib.dd_synth_resize_supported.argtypes = [i, i]
    lib.dd_predictor_bwd_allreduce.argtypes = [vp, C.POINTER(PredictorTensors), vp, vp, C.POINTER(PredictorTensors), i, vp, sz,
                                               C.POINTER(PeerExchange), vp]
    lib.dd_debug_blur_tc.argtypes = [vp, vp, i, i, i, i, vp]
    lib.dd_synth_fwd_ex.argtypes = [vp, i, f, vp, vp, vp, vp, i, vp, vp, ll, vp, sz, vp]
    lib.dd_resize256_ex.argtypes = [vp, i, vp, i, i, i, vp]
    lib.dd_recovery_fwd_ex.argtypes = [vp, i, vp, vp, vp, vp, i, i, i, i, vp]
    lib.dd_recovery_bwd_ex.argtypes = [vp, i, vp, vp, vp, vp, i, vp, vp, i, i, i, vp, sz, vp]
    lib.dd_dark_prior.argtypes = [vp, f, vp, vp, vp, i, i, i, vp, sz, vp]
    lib.dd_dark_table.argtypes = [f, vp, vp, vp]
    lib.dd_recovery_fwd_u8.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, vp]
    lib.dd_recovery_bwd_u8.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, i, vp, sz, vp]
    lib.dd_exchange_bytes.restype = sz
    lib.dd_exchange_bytes.argtypes = []
    for name in EXPORTS[4:-1]:
        getattr(lib, name).restype = i
    return lib


lib = _load()


def check(code: int) -> None:
    """Translate a DD_ERR_* code into the exception the reference raises for the same condition."""
    if code == DD_OK:
        return
    msg = (lib.dd_last_error() or b"").decode("utf-8", "replace")
    if code == DD_ERR_WIDTH_LT3:
        raise IndexError(msg)            # reference: IndexError from rgb2lum (util_filters.py:270-273)
    if code == DD_ERR_INVALID:
        raise ValueError(msg)
    raise RuntimeError(msg)              # reflect pad (filtersB.py:167), workspace, CUDA


def workspace_bytes(kind: int, B: int, H: int = 0, W: int = 0) -> int:
    return int(lib.dd_workspace_bytes(kind, B, H, W))


def launch_count() -> int:
    return int(lib.dd_launch_count())
