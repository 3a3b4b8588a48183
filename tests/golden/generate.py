"""Golden-vector generator -- runs the REAL reference (cvYouTian/Dedark-YOLO, /root/reference) on CPU.

Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/generate.py

It imports the unmodified reference modules through ``oracle/reference_loader.py`` (Recipe A of
SURVEY.md section 8(c)), executes them in fp32 on seeded inputs and writes small ``.npz`` fixtures
next to this file.  The fixtures pin ``oracle/lowlight_oracle.py`` (tests/test_oracle.py) and are
the ground truth for the ``-m gpu`` parity tests.  Nothing here is copied from the reference: only
its *outputs* are recorded.

Fixtures
  weights_seed0.npz      state-dict of ``lowlight_recovery(3)`` built after ``torch.manual_seed(0)``
  case_*.npz             x, (A, IcA), cotangent g, reference y / feat / parameter grads / dx
  synth.npz              darkening LUTs ``pow(k/255, p)`` for k=0..255, mse, truncating u8 writer
  bus640.npz             bus.jpg -> 640x640 RGB u8 (BASELINE config 1) + reference output checksums
  loss_term.npz          RcoveryDetectionLoss recovery term (utils/loss.py:393-416) on fixed numbers
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.reference_loader import REFERENCE_ROOT, load_reference  # noqa: E402

FC1_STRIDE = 16  # fc1.weight grads are stored sub-sampled (every 16th element) + sums


def _np(t):
    return t.detach().cpu().numpy()


def make_module(ref, fc2_scale=1.0):
    torch.manual_seed(0)
    m = ref.lowlight_recovery(3)
    if fc2_scale != 1.0:
        with torch.no_grad():
            m.extractor.fc2.weight.mul_(fc2_scale)
            m.extractor.fc2.bias.mul_(fc2_scale)
    return m


def pack_grads(m):
    out = {}
    for k, p in m.named_parameters():
        g = _np(p.grad).astype(np.float32)
        if k == "extractor.fc1.weight":
            out["grad_sub." + k] = g.reshape(-1)[::FC1_STRIDE].copy()
            out["grad_sum." + k] = np.array([g.astype(np.float64).sum(), np.abs(g.astype(np.float64)).sum()])
        else:
            out["grad." + k] = g
    return out


def run_case(ref, name, x, A=None, IcA=None, fc2_scale=1.0, g_seed=4321, with_dx=True):
    m = make_module(ref, fc2_scale)
    m.train()
    x = x.clone().requires_grad_(with_dx)
    feats = {}
    hook = m.extractor.register_forward_hook(lambda mod, i, o: feats.__setitem__("f", o))
    y = m(x) if A is None and IcA is None else m(x, A, IcA)
    hook.remove()
    gen = torch.Generator().manual_seed(g_seed)
    g = torch.randn(y.shape, generator=gen)
    feat = feats["f"]
    feat.retain_grad()
    y.backward(g)
    rec = {
        "x": _np(x).astype(np.float32), "g": _np(g), "y": _np(y), "feat": _np(feat), "dfeat": _np(feat.grad),
        "fc2_scale": np.float32(fc2_scale),
    }
    if A is not None:
        rec["A"] = _np(A)
    if IcA is not None:
        rec["IcA"] = _np(IcA)
    if with_dx:
        rec["dx"] = _np(x.grad)
    rec.update(pack_grads(m))
    np.savez_compressed(os.path.join(HERE, f"case_{name}.npz"), **rec)
    print(f"case_{name}: y range [{float(y.min()):.3f}, {float(y.max()):.3f}]  |dfeat|max {float(feat.grad.abs().max()):.3e}")


def gen_cases(ref):
    m = make_module(ref)
    np.savez_compressed(os.path.join(HERE, "weights_seed0.npz"), **{k: _np(v) for k, v in m.state_dict().items()})

    gen = torch.Generator().manual_seed(11)
    run_case(ref, "small_default", torch.rand(2, 3, 96, 80, generator=gen))

    gen = torch.Generator().manual_seed(7)
    x = torch.rand(2, 3, 96, 80, generator=gen)
    A = 0.4 + 0.5 * torch.rand(2, 3, generator=gen)
    IcA = torch.rand(2, 1, 96, 80, generator=gen)
    run_case(ref, "small_custom", x, A, IcA, fc2_scale=8.0)

    gen = torch.Generator().manual_seed(13)
    run_case(ref, "min13", torch.rand(1, 3, 13, 13, generator=gen), fc2_scale=8.0)

    gen = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 64, 72, generator=gen) ** 7.5  # a darkened batch: most pixels hit the gamma clamp
    run_case(ref, "dark", x, fc2_scale=4.0)

    gen = torch.Generator().manual_seed(17)
    x = torch.rand(1, 3, 150, 200, generator=gen)  # larger than one blur tile in both directions
    A = 0.4 + 0.5 * torch.rand(1, 3, generator=gen)
    IcA = torch.rand(1, 1, 150, 200, generator=gen)
    run_case(ref, "wide_custom", x, A, IcA, fc2_scale=8.0)


def gen_synth():
    rec = {}
    k = torch.arange(256, dtype=torch.uint8)
    clean = k.float() / 255
    gen = torch.Generator().manual_seed(2024)
    u8 = torch.randint(0, 256, (2, 3, 40, 48), dtype=torch.uint8, generator=gen)
    rec["u8"] = _np(u8)
    for p in (5.0, 7.5, 10.0, 15.0):
        lut = torch.pow(clean, p)  # train.py:79 / lowlight_process.py:68
        rec[f"lut_{p}"] = _np(lut)
        c = u8.float() / 255
        d = torch.pow(c, p)
        rec[f"mse_{p}"] = _np(torch.nn.functional.mse_loss(d, c))  # train.py:108
        rec[f"q_{p}"] = (_np(d) * 255).astype(np.uint8)  # lowlight_process.py:74
    gen = torch.Generator().manual_seed(1234)
    cf = torch.rand(2, 3, 32, 32, generator=gen)
    rec["clean_f32"] = _np(cf)
    rec["dark_f32_15.0"] = _np(torch.pow(cf, 15.0))
    np.savez_compressed(os.path.join(HERE, "synth.npz"), **rec)
    print("synth: ok")


def gen_bus(ref):
    import cv2
    bgr = cv2.imread(os.path.join(REFERENCE_ROOT, "bus.jpg"))
    assert bgr is not None and bgr.shape == (1080, 810, 3)
    rgb = cv2.cvtColor(cv2.resize(bgr, (640, 640), interpolation=cv2.INTER_LINEAR), cv2.COLOR_BGR2RGB)
    u8 = np.ascontiguousarray(rgb.transpose(2, 0, 1))
    x = (torch.from_numpy(u8).float() / 255)[None].repeat(4, 1, 1, 1)
    m = make_module(ref).eval()
    with torch.no_grad():
        y = m(x)
    np.savez_compressed(
        os.path.join(HERE, "bus640.npz"), u8=u8,
        y_sum=np.float64(y.double().sum()), y_mean_abs=np.float64(y.double().abs().mean()),
        y_sub=_np(y[0, :, ::8, ::8]), y_rows=_np(y[0, :, 317:323, :]))
    print(f"bus640: sum(y)={float(y.double().sum()):.2f}  mean|y|={float(y.double().abs().mean()):.6f}")


def gen_loss_term():
    """Exercise utils/loss.py:393-416 for real: stub the plotting deps, import the loss module, replace the
    parent's ``__call__`` (the detector loss, out of scope) by fixed numbers."""
    from unittest.mock import MagicMock
    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        sys.modules.setdefault(name, MagicMock())
    sys.path.insert(0, REFERENCE_ROOT)
    argv, sys.argv = sys.argv, sys.argv[:1]
    try:
        from ultralytics.utils import loss as ref_loss
    finally:
        sys.argv = argv
    fixed_loss, fixed_items = torch.tensor(24.5), torch.tensor([3.25, 5.0, 4.125])
    rec_cases = {"scalar": torch.tensor(0.2481), "vector": torch.tensor([0.1, 0.3, 0.5]), "zero": torch.tensor(0.0)}
    out = {"base_loss": _np(fixed_loss), "base_items": _np(fixed_items), "lrl": np.float32(2.0)}
    orig = ref_loss.v8DetectionLoss.__call__
    ref_loss.v8DetectionLoss.__call__ = lambda self, preds, batch: (fixed_loss.clone(), fixed_items.clone())
    try:
        obj = ref_loss.RcoveryDetectionLoss.__new__(ref_loss.RcoveryDetectionLoss)
        obj.recovery_weight = 2.0
        for name, rec in rec_cases.items():
            loss, items = obj(None, {"recovery_loss_batch": rec})
            out[f"rec_{name}"], out[f"loss_{name}"], out[f"items_{name}"] = _np(rec), _np(loss), _np(items)
        loss, items = obj(None, {})
        out["loss_absent"], out["items_absent"] = _np(loss), _np(items)
    finally:
        ref_loss.v8DetectionLoss.__call__ = orig
    np.savez_compressed(os.path.join(HERE, "loss_term.npz"), **out)
    print("loss_term: ok", {k: v.tolist() for k, v in out.items() if k.startswith("loss_")})


if __name__ == "__main__":
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = load_reference()
    gen_cases(ref)
    gen_synth()
    gen_bus(ref)
    try:
        gen_loss_term()
    except Exception as e:  # the loss import drags most of ultralytics in; keep the rest usable
        print("loss_term: SKIPPED:", repr(e))
