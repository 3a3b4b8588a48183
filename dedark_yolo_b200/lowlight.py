"""Low-light synthesis and the recovery-loss term around the module (host-side mirrors of the reference).

  preprocess_batch   <- DetectionTrainer.preprocess_batch   (models/yolo/detect/train.py:70-111)
  apply_lowlight     <- apply_lowlight_and_save, pre-encode  (utils/lowlight_process.py:57-74)
  add_recovery_term  <- RcoveryDetectionLoss.__call__        (utils/loss.py:393-416)

The arithmetic (``/255``, ``pow``, ``mse``, truncating uint8 writer) is one CUDA pass (``dd_synth_fwd``).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def preprocess_batch(batch: dict, device, dark_param: float = 15.0, lowlight_FLAG: bool = True,
                     dedark_FLAG: bool = False, lut: Optional[torch.Tensor] = None,
                     clean_lut: Optional[torch.Tensor] = None) -> dict:
    """Same keys and semantics as the reference trainer hook:

    ``batch['img']`` uint8 NCHW  ->  ``clean_img`` (= u8/255), ``img`` (= clean ** dark_param when
    ``lowlight_FLAG``) and ``recovery_loss_batch`` (= mse(img, clean_img), a constant without grad).
    With ``dedark_FLAG and lowlight_FLAG`` the reference overwrites clean_img with the darkened image, so
    img is clean_img and the loss is exactly 0 (train.py:79,100,108); that branch's CPU dark-channel loop
    (train.py:81-97) produces ``dedark_A``/``IcA`` that no consumer reads and is not
    reproduced here.
    """
    src = batch["img"].to(device, non_blocking=True)
    if not lowlight_FLAG:
        clean, _, _, _ = ops.synth_forward(src, 1.0, clean_lut=clean_lut, want_dark=False, want_rec=False) if src.dtype == torch.uint8 \
            else (src.float(), None, None, None)
        batch["clean_img"] = clean
        batch["img"] = clean
        batch["recovery_loss_batch"] = torch.zeros((), dtype=torch.float32, device=src.device)
        return batch
    clean, dark, _, rec = ops.synth_forward(src, dark_param, lut=lut, clean_lut=clean_lut)
    if dedark_FLAG:
        batch["clean_img"] = dark
        batch["img"] = dark
        batch["recovery_loss_batch"] = torch.zeros((), dtype=torch.float32, device=src.device)
    else:
        batch["clean_img"] = clean
        batch["img"] = dark
        batch["recovery_loss_batch"] = rec
    return batch


class HostBatchPrefetcher:
    """Double-buffered host -> device staging of the dataloader's pinned uint8 batches (train.py:72 does the same
    ``.to(device, non_blocking=True)`` on the compute stream, where it serialises with the step).

    ``submit(host_u8)`` enqueues the copy on a private copy stream; ``get()`` orders the current stream after the
    oldest outstanding copy and returns that device tensor.  Submitting batch i+1 right after ``get()`` of batch i
    overlaps its PCIe transfer (19.7 MB for 16x3x640x640) with the kernels of step i.  A slot is only overwritten
    after everything enqueued up to the following ``get()`` has finished reading it.
    """

    def __init__(self, device, depth: int = 2):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("HostBatchPrefetcher needs a CUDA device")
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.buf = [None] * depth
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [None] * depth
        self.n_sub = 0
        self.n_get = 0

    def submit(self, host_u8: torch.Tensor) -> None:
        if self.n_sub - self.n_get >= self.depth:
            raise RuntimeError("HostBatchPrefetcher: all slots are in flight; call get() first")
        k = self.n_sub % self.depth
        if self.buf[k] is None or self.buf[k].shape != host_u8.shape or self.buf[k].dtype != host_u8.dtype:
            self.buf[k] = torch.empty(host_u8.shape, dtype=host_u8.dtype, device=self.dev)
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[k] is not None:
                self.copy_stream.wait_event(self.consumed[k])
            self.buf[k].copy_(host_u8, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self.n_sub += 1

    def get(self) -> torch.Tensor:
        if self.n_get >= self.n_sub:
            raise RuntimeError("HostBatchPrefetcher: nothing submitted")
        cur = torch.cuda.current_stream(self.dev)
        if self.n_get > 0:  # the previous slot has been consumed by everything enqueued so far
            j = (self.n_get - 1) % self.depth
            self.consumed[j] = torch.cuda.Event()
            self.consumed[j].record(cur)
        k = self.n_get % self.depth
        cur.wait_event(self.ready[k])
        self.n_get += 1
        return self.buf[k]


def apply_lowlight(u8: torch.Tensor, lowlight_param: float = 7.5, lut: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 RGB NCHW -> darkened uint8 (``(pow(u8/255, p) * 255).astype(uint8)``, truncation).  This is the array
    the reference hands to ``cv2.imwrite`` (before the RGB->BGR flip and the JPEG encoder)."""
    _, _, q, _ = ops.synth_forward(u8, lowlight_param, lut=lut, want_clean=False, want_dark=False, want_u8=True,
                                   want_rec=False)
    return q


def add_recovery_term(loss: torch.Tensor, loss_items: torch.Tensor, batch: dict, lrl: float = 2.0):
    """utils/loss.py:393-416: ``loss += lrl * rec`` and the same amount is folded into the cls column, keeping
    ``loss_items`` at shape [3].  ``rec`` is a constant (no gradient flows through it)."""
    box, cls, dfl = loss_items
    rec = batch.get("recovery_loss_batch")
    if rec is not None:
        if rec.ndim > 0:
            rec = rec.mean()
        rec = rec.to(loss.device)
        cls = cls + lrl * rec
        loss = loss + lrl * rec
    items = torch.stack([box.detach(), cls.detach(), dfl.detach()]).to(loss.device)
    return loss, items
