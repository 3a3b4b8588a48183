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
                     dedark_FLAG: bool = True, lut: Optional[torch.Tensor] = None,
                     clean_lut: Optional[torch.Tensor] = None, dedark_prior: bool = False) -> dict:
    """Same keys and semantics as the reference trainer hook (defaults = cfg/default.yaml:32-34: ``dark_param: 15.0``,
    ``lowlight_FLAG: True``, ``dedark_FLAG: True``):

    ``batch['img']`` uint8 NCHW  ->  ``clean_img`` (= u8/255), ``img`` (= clean ** dark_param when
    ``lowlight_FLAG``) and ``recovery_loss_batch`` (= mse(img, clean_img), a constant without grad).
    With ``dedark_FLAG and lowlight_FLAG`` the reference overwrites clean_img with the darkened image, so
    img is clean_img and the loss is exactly 0 (train.py:79,100,108); that branch's CPU dark-channel loop
    (train.py:81-97) produces ``dedark_A``/``IcA`` from uninitialised memory (train.py:65-67) that no training-mode
    consumer reads (tasks.py:107-110): the keys are emitted as None, which ``_predict_once``
    (tasks.py:87-91) and the module treat as "use the defaults A = 0.8, IcA = 0.5".

    ``dedark_prior=True`` (opt-in, a behaviour change -- SURVEY.md section 8(f) N3) fills those two keys with the dark-channel
    prior computed on the GPU (``ops.dark_prior``: channel minimum, atmospheric light from the brightest 0.1 % of the dark
    channel, per-channel divide), i.e. what train.py:42-68,81-97 set out to compute.

    ``batch['img']`` must be uint8 (what the dataloader's collate hands over, data/dataset.py:172-188): the reference
    divides whatever it gets by 255 (train.py:72), so a float image in [0, 1] would silently be darkened from
    [0, 1/255]; float sources are rejected instead of guessing.
    """
    if batch["img"].dtype != torch.uint8:
        raise TypeError(f"preprocess_batch: uint8 NCHW batch expected (train.py:72 divides by 255), got {batch['img'].dtype}; "
                        "use ops.synth_forward for float sources already scaled to [0, 1]")
    src = batch["img"].to(device, non_blocking=True)
    if not lowlight_FLAG:
        clean, _, _, _ = ops.synth_forward(src, 1.0, clean_lut=clean_lut, want_dark=False, want_rec=False)
        batch["clean_img"] = clean
        batch["img"] = clean
        batch["recovery_loss_batch"] = torch.zeros((), dtype=torch.float32, device=src.device)
        return batch
    if src.dim() == 4 and src.shape[1] == 3 and ops.synth_resize_supported(src.shape[2], src.shape[3]):
        # one pass: darkening + recovery loss + the 256x256 resize the module applies first (llie.py:43); the resized batch rides
        # on the darkened tensor and lowlight_recovery.forward picks it up instead of re-reading the full-size batch
        clean, dark, r, rec = ops.synth_resize_forward(src, dark_param, lut=lut, clean_lut=clean_lut)
        dark._dd_resize256 = (r, dark._version)
    else:
        clean, dark, _, rec = ops.synth_forward(src, dark_param, lut=lut, clean_lut=clean_lut)
    if dedark_FLAG:
        batch["clean_img"] = dark
        batch["img"] = dark
        if dedark_prior:
            batch["dedark_A"], batch["IcA"] = ops.dark_prior(src, dark_param, lut=lut)
        else:
            batch["dedark_A"] = None
            batch["IcA"] = None
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
    overlaps its PCIe transfer (19.7 MB for 16x3x640x640) with the kernels of step i.

    Slot life cycle: a slot is *outstanding* from ``submit`` to the ``get`` that returns it, then *held* by the consumer
    until the next ``get()`` (or an explicit ``release()``), at which point an event recorded on the consumer's stream
    marks the end of its reads; a later copy into the slot waits for that event.  ``submit`` refuses to touch a held
    slot: with ``depth`` slots at most ``depth - 1`` submits may be outstanding while one is held.
    """

    def __init__(self, device, depth: int = 2):
        self.dev = torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("HostBatchPrefetcher needs a CUDA device")
        if depth < 2:
            raise ValueError("HostBatchPrefetcher needs at least two slots")
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.buf = [None] * depth
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [None] * depth
        self.n_sub = 0
        self.n_get = 0
        self._held = None  # slot handed out by the latest get() and not yet released

    def submit(self, host_u8: torch.Tensor) -> None:
        in_flight = (self.n_sub - self.n_get) + (0 if self._held is None else 1)
        if in_flight >= self.depth:
            raise RuntimeError("HostBatchPrefetcher: all slots are in flight (outstanding copies + the batch held by the "
                               "consumer); call get() or release() first")
        k = self.n_sub % self.depth
        fresh = self.buf[k] is None or self.buf[k].shape != host_u8.shape or self.buf[k].dtype != host_u8.dtype
        if fresh:
            self.buf[k] = torch.empty(host_u8.shape, dtype=host_u8.dtype, device=self.dev)
            # the caching allocator may hand back memory whose last use is still queued on the allocating stream
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.dev))
            self.consumed[k] = None
        with torch.cuda.stream(self.copy_stream):
            if self.consumed[k] is not None:
                self.copy_stream.wait_event(self.consumed[k])
            self.buf[k].copy_(host_u8, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self.n_sub += 1

    def release(self) -> None:
        """Everything enqueued on the current stream so far is the last use of the batch returned by the latest ``get()``."""
        if self._held is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.dev))
            self.consumed[self._held] = ev
            self._held = None

    def get(self) -> torch.Tensor:
        if self.n_get >= self.n_sub:
            raise RuntimeError("HostBatchPrefetcher: nothing submitted")
        self.release()  # the previous batch has been consumed by everything enqueued so far
        k = self.n_get % self.depth
        torch.cuda.current_stream(self.dev).wait_event(self.ready[k])
        self.n_get += 1
        self._held = k
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
