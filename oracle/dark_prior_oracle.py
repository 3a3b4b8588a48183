"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the dark-channel prior that ``dd_dark_prior`` computes (SURVEY.md
section 8(f) N3).  Only tests/ may import this file; the product path never does.

Follows DetectionTrainer.preprocess_batch (models/yolo/detect/train.py:81-97) and its helpers:

  DarkChannel (train.py:42-45)   dc = min(r, g, b) per pixel of the quantised darkened image
                                 ``(pow(u8/255, p) * 255).astype(np.uint8)`` (train.py:84)
  AtmLight   (train.py:47-61)    numpx = max(floor(HW / 1000), 1); the numpx pixels of largest dc are selected by argsort and
                                 the colours of all but the first of them (``range(1, numpx)``) are summed and divided by numpx
  DarkIcA    (train.py:63-67)    intended: dark channel of im / A per channel

PARITY: ``atm_light`` is PINNED against the reference's own AtmLight on inputs whose selected dark-channel values are
distinct (tests/golden/prior.npz, recorded by tests/golden/generate_prior.py by executing the reference's function bodies).
With ties at the selection threshold numpy's argsort picks an implementation-defined subset; this restatement (and the
kernel) let the tied pixels contribute in equal shares instead.  ``DarkIcA`` is PARITY UNPINNED: the reference indexes the
first axis of an HWC array (rows 0..2) into an uninitialised uint8 buffer (``np.empty``), so its output is not a function of its
input; what is restated here is the documented intent (per-channel divide, then channel minimum), in float32, with the
atmospheric light floored at 1 (uint8 units) so that a black image does not divide by zero.
"""
from __future__ import annotations

import numpy as np


def dark_table_u8(dark_lut: np.ndarray) -> np.ndarray:
    """train.py:84: float32 darkened value in [0, 1] -> ``(d * 255).astype(uint8)`` (truncation), for the 256 source values."""
    d = np.asarray(dark_lut, dtype=np.float32).reshape(256)
    return (d * np.float32(255.0)).astype(np.uint8)


def atm_light(dark_u8_chw: np.ndarray) -> np.ndarray:
    """AtmLight in uint8 units for one image [3, H, W] of quantised darkened values -> float32 [3]."""
    im = np.asarray(dark_u8_chw)
    hw = im.shape[1] * im.shape[2]
    dc = im.min(axis=0).reshape(-1).astype(np.int64)
    cols = im.reshape(3, -1).astype(np.int64)
    numpx = max(hw // 1000, 1)
    need = numpx - 1
    count = np.bincount(dc, minlength=256)
    sums = [np.bincount(dc, weights=cols[c].astype(np.float64), minlength=256) for c in range(3)]   # exact: integers < 2^53
    s = np.zeros(3, dtype=np.float64)
    for v in range(255, -1, -1):
        if need <= 0:
            break
        n = int(count[v])
        if n == 0:
            continue
        if n <= need:
            for c in range(3):
                s[c] += sums[c][v]
            need -= n
        else:
            for c in range(3):
                s[c] += sums[c][v] * float(need) / float(n)
            need = 0
    return (s / float(numpx)).astype(np.float32)


def dark_prior(src_u8: np.ndarray, dark_lut: np.ndarray):
    """src_u8 [B,3,H,W] uint8, dark_lut: 256 float32 darkened values -> (A [B,3] in [0,1] units, IcA [B,1,H,W]) float32."""
    src = np.asarray(src_u8)
    B, _, H, W = src.shape
    tab = dark_table_u8(dark_lut)
    A = np.zeros((B, 3), dtype=np.float32)
    ica = np.zeros((B, 1, H, W), dtype=np.float32)
    for b in range(B):
        q = tab[src[b]]                         # [3,H,W] quantised darkened image
        au = atm_light(q)
        A[b] = au / np.float32(255.0)
        den = np.maximum(au, np.float32(1.0))
        ratio = tab.astype(np.float32)[None, :] / den[:, None]          # [3,256] float32 division
        ica[b, 0] = np.minimum(np.minimum(ratio[0][src[b, 0]], ratio[1][src[b, 1]]), ratio[2][src[b, 2]])
    return A, ica
