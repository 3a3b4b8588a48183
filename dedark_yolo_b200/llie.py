"""Drop-in ``lowlight_recovery`` module: the reference's interface, B200 kernels underneath.

Mirrors ``ultralytics/nn/modules/llie.py:11-53`` (class, constructor, forward signature) and the child
layout of ``ultralytics/nn/modules/common.py:9-23,52-78`` so that the 14 state-dict keys

    extractor.conv_layers.{0..4}.conv_block.0.{weight,bias}, extractor.fc1.{weight,bias}, extractor.fc2.{weight,bias}

match the reference exactly (checkpoints load both ways), and the torch layers are constructed in the same
order (same RNG consumption: same seed -> same initial weights).  The torch layers only *hold* the parameters:
``forward`` never calls them.  It runs resize -> predictor -> fused filter chain through the C-ABI as a single
autograd node (``ops.RecoveryFunction``).

Behaviour kept from the reference:
  * ``forward(x, dedark_A=None, IcA=None)``; defaults A = 0.8, IcA = 0.5 (llie.py:34-40);
  * the module follows the input's device (``self.to(x.device)``, llie.py:28) when x is a CUDA tensor;
  * output is a new fp32 tensor on x's device, x is not modified, fp16/bf16 inputs are promoted;
  * errors: H or W <= 12 -> RuntimeError, W < 3 -> IndexError, C != 3 -> RuntimeError, fp64 -> RuntimeError;
  * CPU input (the stride probe in DetectionModel.__init__, nn/tasks.py:290-291) is accepted: it is staged
    through the GPU.  Without any CUDA device the call raises -- there is no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

__all__ = ("lowlight_recovery", "ExtractParameters2", "ConvBlock", "FILTER_ORDER")

# filter_cfg.py:65-75 -- the chain the fused kernel implements, in order
FILTER_ORDER = ("DF", "W", "G", "Ct", "UF")
_FILTER_SLOTS = {"DF": (0, 1), "W": (1, 3), "G": (4, 1), "Ct": (13, 1), "UF": (14, 1)}  # filter_cfg.py:19-25


class ConvBlock(nn.Module):
    """Parameter holder with the reference's child names (common.py:9-23, bn=False)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv_block = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=2, padding=1),
            nn.LeakyReLU(0.1, inplace=False),
        )


class ExtractParameters2(nn.Module):
    """Parameter holder for the predictor CNN (common.py:52-66): 3->16->32->32->32->32, fc 2048->64->15."""

    def __init__(self):
        super().__init__()
        self.output_dim = 15
        self.channels = 16
        c = self.channels
        self.conv_layers = nn.Sequential(
            ConvBlock(3, c), ConvBlock(c, 2 * c), ConvBlock(2 * c, 2 * c), ConvBlock(2 * c, 2 * c), ConvBlock(2 * c, 2 * c))
        self.fc1 = nn.Linear(2048, 64)
        self.fc2 = nn.Linear(64, self.output_dim)

    def ordered_parameters(self):
        """The 14 tensors in state-dict order (the order of ``dd_predictor_tensors``)."""
        out = []
        for blk in self.conv_layers:
            out += [blk.conv_block[0].weight, blk.conv_block[0].bias]
        return out + [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias]

    def forward(self, r):
        """[B,3,256,256] -> [B,15] through the CUDA predictor (inference helper; no autograd)."""
        from . import ops
        feat, _ = ops.predictor_forward(r.float().contiguous(), [p.detach().float().contiguous() for p in self.ordered_parameters()])
        return feat


class _FilterSlot(nn.Module):
    """Stateless stand-in for one entry of the reference's ``cfg.filters`` (filtersB.py): keeps the module tree
    shape (``len(m.filters) == 5``, short names, parameter slots).  The arithmetic lives in the fused kernel."""

    def __init__(self, short_name):
        super().__init__()
        self.short_name = short_name
        self.begin_filter_parameter, self.num_filter_parameters = _FILTER_SLOTS[short_name]

    def get_short_name(self):
        return self.short_name

    def forward(self, *a, **k):
        raise RuntimeError("filters are fused into dd_recovery_fwd; call lowlight_recovery.forward")


class lowlight_recovery(nn.Module):
    def __init__(self, in_channels=3, out_channels=3):
        super().__init__()
        self.extractor = ExtractParameters2()
        self.filters = nn.ModuleList(_FilterSlot(n) for n in FILTER_ORDER)

    def _compute_device(self, x):
        if x.is_cuda:
            return x.device
        if not torch.cuda.is_available():
            raise RuntimeError(
                "lowlight_recovery (dedark_yolo_b200): no CUDA device available -- this build has no CPU fallback")
        p = next(self.parameters())
        return p.device if p.is_cuda else torch.device("cuda", torch.cuda.current_device())

    def forward(self, x, dedark_A=None, IcA=None):
        from . import ops
        ops.check_image_shape(x)
        if x.is_cuda and any(p.device != x.device for p in self.parameters()):
            self.to(x.device)  # llie.py:28: the module follows its input (a no-op traversal is skipped: 0.2 ms of host time)
        dev = self._compute_device(x)
        return ops.RecoveryFunction.apply(dev, x, dedark_A, IcA, *self.extractor.ordered_parameters())
