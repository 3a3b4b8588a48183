"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the Dedark-YOLO low-light hot path.

Nothing under ``dedark_yolo_b200/`` may import this file.  Its only consumers are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
(where it is the thing being *timed as the CPU baseline*, never the product).

What it restates (all paths relative to the reference checkout, cvYouTian/Dedark-YOLO):

  synthesis + recovery loss   ultralytics/models/yolo/detect/train.py:72,79,103,108-109
                              ultralytics/utils/lowlight_process.py:57,68,74
  recovery-loss term          ultralytics/utils/loss.py:393-416
  module forward              ultralytics/nn/modules/llie.py:17-53
  predictor CNN               ultralytics/nn/modules/common.py:9-23,52-78
  filter slots / ranges       ultralytics/nn/modules/filter_cfg.py:17-44,65-75
  regressors + filters        ultralytics/nn/modules/filtersB.py:32-37,144-259,289-303
  rgb2lum / tanh_range / lerp ultralytics/nn/modules/util_filters.py:270-273,295-304,316-317

Parity status: PINNED.  The reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the pin is made by executing the real reference modules in the build
container (``oracle/reference_loader.py``) and committing their inputs/outputs as
``tests/golden/*.npz`` (generator: ``tests/golden/generate.py``).  ``tests/test_oracle.py``
checks every function here against those vectors.

The restatement is dtype-generic torch code so that the same functions serve as
  * the fp32 "port" (``dense_blur=True`` reproduces the reference's op mix: reflect pad + dense
    25x25 conv2d) -- the CPU baseline and the 1e-5 forward gate;
  * the fp64 "truth" (``dense_blur=False``: separable blur, explicit reflect gather) -- the
    gradient gate (autograd in float64).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

# ---- frozen constants (filter_cfg.py:17-44, filtersB.py:156,163,247) --------------------------
NUM_FEATURES = 15
SLOT_DEDARK, SLOT_WB, SLOT_GAMMA, SLOT_TONE, SLOT_CONTRAST, SLOT_USM = 0, 1, 4, 5, 13, 14
DEFOG_RANGE = (0.1, 1.0)
GAMMA_RANGE = 3.0
USM_RANGE = (0.0, 5.0)
LOG_WB_RANGE = 0.5
LUM_W = (0.27, 0.67, 0.06)
WB_EPS = 1e-5
CONTRAST_EPS = 1e-6
TX_MIN = 0.01
GAMMA_CLAMP = 1e-4
BLUR_RADIUS = 12
BLUR_SIGMA = 5.0
RESIZE = 256
LEAKY = 0.1
DEFAULT_A = 0.8
DEFAULT_ICA = 0.5

STATE_KEYS = tuple(
    [f"extractor.conv_layers.{i}.conv_block.0.{n}" for i in range(5) for n in ("weight", "bias")]
    + ["extractor.fc1.weight", "extractor.fc1.bias", "extractor.fc2.weight", "extractor.fc2.bias"]
)
STATE_SHAPES = (
    (16, 3, 3, 3), (16,), (32, 16, 3, 3), (32,), (32, 32, 3, 3), (32,), (32, 32, 3, 3), (32,),
    (32, 32, 3, 3), (32,), (64, 2048), (64,), (15, 64), (15,),
)


# ---- a1/a2: synthesis and recovery loss --------------------------------------------------------
def synth_clean_from_u8(u8: torch.Tensor) -> torch.Tensor:
    """train.py:72 -- ``batch['img'].float() / 255`` (true division, fp32)."""
    return u8.float() / 255


def synth_darken(clean: torch.Tensor, p: float) -> torch.Tensor:
    """train.py:79,103 and lowlight_process.py:68 -- ``torch.pow(clean, p)`` with a python float."""
    return torch.pow(clean, p)


def synth_quantize_u8(dark: torch.Tensor) -> torch.Tensor:
    """lowlight_process.py:74 -- ``(img * 255).astype(np.uint8)``: truncation, not rounding."""
    return (dark * 255).to(torch.uint8)


def recovery_mse(img: torch.Tensor, clean: torch.Tensor) -> torch.Tensor:
    """train.py:108 -- ``F.mse_loss(batch['img'], batch['clean_img'])`` (mean over all elements)."""
    return F.mse_loss(img, clean)


def recovery_loss_term(loss: torch.Tensor, loss_items: torch.Tensor, rec, lrl: float):
    """loss.py:393-416 -- fold ``lrl * rec`` into the total and into the cls column.

    ``rec`` may be None (term skipped) or a tensor; non-scalar tensors are ``.mean()``-ed first.
    Returns ``(loss, loss_items)`` with ``loss_items`` still of shape [3]."""
    box, cls, dfl = loss_items
    if rec is not None:
        if rec.ndim > 0:
            rec = rec.mean()
        cls = cls + lrl * rec
        loss = loss + lrl * rec
    items = torch.stack([box.detach(), cls.detach(), dfl.detach()]).to(loss.device)
    return loss, items


# ---- a4: bilinear resize to 256x256 (llie.py:43) ----------------------------------------------
def _src_index(out_size: int, in_size: int, dtype):
    """ATen ``upsample_bilinear2d`` source index, align_corners=False, no antialias."""
    scale = in_size / out_size
    i = torch.arange(out_size, dtype=dtype)
    s = (i + 0.5) * scale - 0.5
    s = torch.clamp(s, min=0.0)
    i0 = torch.floor(s).to(torch.int64)
    i0 = torch.clamp(i0, max=in_size - 1)
    i1 = torch.clamp(i0 + 1, max=in_size - 1)
    lam = s - i0.to(dtype)
    return i0, i1, lam


def resize256(x: torch.Tensor) -> torch.Tensor:
    H, W = x.shape[-2:]
    # the fp32 path computes indices/lambdas in fp32 exactly like ATen does for float input
    y0, y1, ly = _src_index(RESIZE, H, x.dtype)
    x0, x1, lx = _src_index(RESIZE, W, x.dtype)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    ly = ly.view(-1, 1)
    rows0 = top[..., x0] * (1 - lx) + top[..., x1] * lx
    rows1 = bot[..., x0] * (1 - lx) + bot[..., x1] * lx
    return rows0 * (1 - ly) + rows1 * ly


# ---- a5: predictor CNN (common.py:52-78) -------------------------------------------------------
def predictor_forward(r: torch.Tensor, w: Dict[str, torch.Tensor], return_acts: bool = False):
    """5x (conv3x3 s2 p1 + bias + LeakyReLU 0.1) -> flatten 2048 -> fc1 + LeakyReLU -> fc2."""
    acts = [r]
    net = r
    for i in range(5):
        net = F.conv2d(net, w[f"extractor.conv_layers.{i}.conv_block.0.weight"],
                       w[f"extractor.conv_layers.{i}.conv_block.0.bias"], stride=2, padding=1)
        net = F.leaky_relu(net, LEAKY)
        acts.append(net)
    flat = net.contiguous().view(-1, 2048)
    h = F.leaky_relu(F.linear(flat, w["extractor.fc1.weight"], w["extractor.fc1.bias"]), LEAKY)
    acts.append(h)
    feat = F.linear(h, w["extractor.fc2.weight"], w["extractor.fc2.bias"])
    return (feat, acts) if return_acts else feat


# ---- a6/a7: regressors (filtersB.py:151-152,186-187,227-229,246-256,296-297) -------------------
def _tanh_range(v, lo, hi):
    """util_filters.py:295-304 (``initial`` is ignored by the reference)."""
    return torch.tanh(v) * (hi - lo) / 2.0 + (hi + lo) / 2.0


def regress(feat: torch.Tensor) -> Dict[str, torch.Tensor]:
    """[B,15] raw features -> the seven filter parameters (per image)."""
    w = _tanh_range(feat[:, SLOT_DEDARK:SLOT_DEDARK + 1], *DEFOG_RANGE)                 # [B,1]
    mask = torch.tensor([[0.0, 1.0, 1.0]], dtype=feat.dtype)
    cs = torch.exp(_tanh_range(feat[:, SLOT_WB:SLOT_WB + 3] * mask, -LOG_WB_RANGE, LOG_WB_RANGE))
    s = cs * 1.0 / (WB_EPS + LUM_W[0] * cs[:, 0] + LUM_W[1] * cs[:, 1] + LUM_W[2] * cs[:, 2])[:, None]
    lg = math.log(GAMMA_RANGE)
    gamma = torch.exp(_tanh_range(feat[:, SLOT_GAMMA:SLOT_GAMMA + 1], -lg, lg))          # [B,1]
    c = torch.tanh(feat[:, SLOT_CONTRAST:SLOT_CONTRAST + 1])                              # [B,1]
    p = _tanh_range(feat[:, SLOT_USM:SLOT_USM + 1], *USM_RANGE)                           # [B,1]
    return {"w": w, "s": s, "gamma": gamma, "c": c, "p": p}


# ---- a8..a12: the five filters -----------------------------------------------------------------
def f_dedark(x, w, A, IcA):
    """filtersB.py:211-214."""
    tx = 1 - w[:, :, None, None] * IcA
    tx3 = tx.repeat(1, 3, 1, 1)
    return (x - A[:, :, None, None]) / torch.clamp(tx3, min=TX_MIN) + A[:, :, None, None]


def f_wb(x, s):
    """filtersB.py:259."""
    return x * s[:, :, None, None]


def f_gamma(x, gamma):
    """filtersB.py:232-233."""
    g3 = gamma.repeat(1, 3)
    return torch.pow(torch.clamp(x, GAMMA_CLAMP), g3[:, :, None, None])


def lum_quirk(x):
    """util_filters.py:270-273 applied to an NCHW tensor (filtersB.py:300): the "luminance" mixes the
    first three COLUMNS of each row, per channel -> [B,3,H,1]."""
    return (LUM_W[0] * x[:, :, :, 0] + LUM_W[1] * x[:, :, :, 1] + LUM_W[2] * x[:, :, :, 2])[:, :, :, None]


def f_contrast(x, c):
    """filtersB.py:300-303 + lerp (util_filters.py:316-317)."""
    lum = torch.clamp(lum_quirk(x), 0.0, 1.0)
    cl = -torch.cos(math.pi * lum) * 0.5 + 0.5
    ci = x / (lum + CONTRAST_EPS) * cl
    l = c[:, :, None, None]
    return (1 - l) * x + l * ci


def gaussian_taps(dtype=torch.float32) -> torch.Tensor:
    """filtersB.py:155-161 -- 25 taps, sigma 5, radius 12, normalised; built in fp32 like the reference
    and then cast (so the fp64 truth uses the *same* kernel values the reference convolves with)."""
    xs = torch.arange(-BLUR_RADIUS, BLUR_RADIUS + 1, dtype=torch.float32)
    k = torch.exp(-0.5 * torch.square(xs / BLUR_SIGMA))
    k = k / torch.sum(k)
    return k.to(dtype)


def blur_dense(x):
    """filtersB.py:163-173 -- reflect pad 12 + dense 25x25 cross-correlation, channel by channel."""
    k1 = gaussian_taps(torch.float32)
    k2 = (torch.unsqueeze(k1, 1) * k1).to(x.dtype)[None, None]
    padded = F.pad(x, (BLUR_RADIUS,) * 4, mode="reflect")
    outs = [F.conv2d(padded[:, ch:ch + 1], k2, stride=1) for ch in range(3)]
    return torch.cat(outs, dim=1)


def _reflect_index(n: int) -> torch.Tensor:
    """Indices of ``F.pad(..., mode='reflect')`` with pad 12 on both sides: -i -> i, (n-1)+i -> (n-1)-i."""
    if n <= BLUR_RADIUS:
        raise RuntimeError("reflect padding needs H, W > 12 (same failure as the reference, filtersB.py:167)")
    idx = torch.arange(-BLUR_RADIUS, n + BLUR_RADIUS)
    idx = torch.where(idx < 0, -idx, idx)
    idx = torch.where(idx > n - 1, 2 * (n - 1) - idx, idx)
    return idx


def blur_separable(x):
    """Same operator as ``blur_dense`` written as two 1-D passes (k2d = outer(k1d, k1d), filtersB.py:161).
    Note the 2-D kernel used by the reference is the fp32-rounded outer product; the separable form
    differs from it by <= 1 ulp(fp32) per tap, far inside the 1e-5 gate."""
    k = gaussian_taps(x.dtype)
    H, W = x.shape[-2:]
    xi = x[..., _reflect_index(H), :]
    v = sum(k[j] * xi[..., j:j + H, :] for j in range(2 * BLUR_RADIUS + 1))
    vi = v[..., _reflect_index(W)]
    return sum(k[j] * vi[..., j:j + W] for j in range(2 * BLUR_RADIUS + 1))


def f_usm(x, p, dense_blur: bool = True):
    """filtersB.py:154-175 -- unsharp mask: (x - blur(x)) * p + x."""
    blur = blur_dense(x) if dense_blur else blur_separable(x)
    return (x - blur) * p[:, :, None, None] + x


def filter_chain(x, feat, A=None, IcA=None, dense_blur: bool = True, return_stages: bool = False):
    """llie.py:34-40,49-52 -- defaults + DeDark -> WB -> Gamma -> Contrast -> USM."""
    if x.dim() != 4 or x.shape[1] != 3:
        raise RuntimeError(f"expected [B,3,H,W], got {tuple(x.shape)}")
    B, _, H, W = x.shape
    if W < 3:
        raise IndexError("rgb2lum quirk indexes columns 0..2 (util_filters.py:270-273)")
    if A is None:
        A = torch.ones(B, 3, dtype=x.dtype) * DEFAULT_A
    if IcA is None:
        IcA = torch.ones(B, 1, H, W, dtype=x.dtype) * DEFAULT_ICA
    prm = regress(feat)
    x1 = f_dedark(x, prm["w"], A, IcA)
    x2 = f_wb(x1, prm["s"])
    x3 = f_gamma(x2, prm["gamma"])
    x4 = f_contrast(x3, prm["c"])
    y = f_usm(x4, prm["p"], dense_blur)
    if return_stages:
        return y, {"x1": x1, "x2": x2, "x3": x3, "x4": x4, **prm}
    return y


# ---- a3: the whole module forward --------------------------------------------------------------
def recovery_forward(x, weights: Dict[str, torch.Tensor], A=None, IcA=None,
                     dense_blur: bool = True, return_feat: bool = False):
    """llie.py:17-53.  ``x`` and ``weights`` must share a dtype (fp32 port or fp64 truth)."""
    r = resize256(x)
    feat = predictor_forward(r, weights)
    y = filter_chain(x, feat, A, IcA, dense_blur)
    return (y, feat) if return_feat else y


def cast_weights(weights: Dict[str, torch.Tensor], dtype, requires_grad: bool = False):
    out = {}
    for k in STATE_KEYS:
        t = weights[k].detach().to(dtype).clone()
        t.requires_grad_(requires_grad)
        out[k] = t
    return out


def recovery_forward_backward(x, weights, g, A=None, IcA=None, dtype=torch.float64,
                              dense_blur: bool = False, need_dx: bool = False):
    """Truth for a14: autograd of the restated forward.  Returns (y, feat, dfeat, grads{key}, dx|None).

    ``g`` is the cotangent dL/dy.  ``dfeat`` is dL/d(raw fc2 output) -- the interface between the
    fused filter backward and the predictor backward."""
    x = x.detach().to(dtype).clone().requires_grad_(need_dx)
    w = cast_weights(weights, dtype, requires_grad=True)
    A = None if A is None else A.to(dtype)
    IcA = None if IcA is None else IcA.to(dtype)
    r = resize256(x)
    feat = predictor_forward(r, w)
    feat.retain_grad()
    y = filter_chain(x, feat, A, IcA, dense_blur)
    y.backward(g.to(dtype))
    grads = {k: w[k].grad for k in STATE_KEYS}
    return y.detach(), feat.detach(), feat.grad, grads, (x.grad if need_dx else None)


def init_weights(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Build the predictor's torch layers in the reference's construction order (common.py:58-66) under
    ``torch.manual_seed(seed)`` so RNG consumption -- hence the weights -- match ``lowlight_recovery(3)``."""
    torch.manual_seed(seed)
    ch = [3, 16, 32, 32, 32, 32]
    out = {}
    for i in range(5):
        conv = torch.nn.Conv2d(ch[i], ch[i + 1], kernel_size=3, stride=2, padding=1)
        out[f"extractor.conv_layers.{i}.conv_block.0.weight"] = conv.weight.detach().clone()
        out[f"extractor.conv_layers.{i}.conv_block.0.bias"] = conv.bias.detach().clone()
    fc1 = torch.nn.Linear(2048, 64)
    fc2 = torch.nn.Linear(64, NUM_FEATURES)
    out["extractor.fc1.weight"], out["extractor.fc1.bias"] = fc1.weight.detach().clone(), fc1.bias.detach().clone()
    out["extractor.fc2.weight"], out["extractor.fc2.bias"] = fc2.weight.detach().clone(), fc2.bias.detach().clone()
    return out
