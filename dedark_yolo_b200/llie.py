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

import copy

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


def ordered_parameters(extractor: nn.Module):
    """The 14 tensors of a predictor in state-dict order (the order of ``dd_predictor_tensors``).  Works on this package's
    ``ExtractParameters2`` and on the reference's (a checkpoint written while installed holds reference children)."""
    out = []
    for blk in extractor.conv_layers:
        out += [blk.conv_block[0].weight, blk.conv_block[0].bias]
    return out + [extractor.fc1.weight, extractor.fc1.bias, extractor.fc2.weight, extractor.fc2.bias]


class ExtractParameters2(nn.Module):
    """Parameter holder for the predictor CNN (common.py:52-66): 3->16->32->32->32->32, fc 2048->64->15.
    ``cfg`` is accepted for signature parity with the reference (common.py:53) and ignored: the filter configuration is
    frozen into the kernels (filter_cfg.py:17-44)."""

    def __init__(self, cfg=None):
        super().__init__()
        self.output_dim = 15
        self.channels = 16
        c = self.channels
        self.conv_layers = nn.Sequential(
            ConvBlock(3, c), ConvBlock(c, 2 * c), ConvBlock(2 * c, 2 * c), ConvBlock(2 * c, 2 * c), ConvBlock(2 * c, 2 * c))
        self.fc1 = nn.Linear(2048, 64)
        self.fc2 = nn.Linear(64, self.output_dim)

    def ordered_parameters(self):
        return ordered_parameters(self)

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
    # bf16 I/O mode (the reference trains under autocast, engine/trainer.py:330).  A bf16 input is read
    # in place (half the bytes; the blur then runs as TF32 tensor-core GEMMs, 2e-2 gate); the OUTPUT stays fp32 -- the dtype
    # the reference returns, because its fp32 defaults promote -- unless ``out_dtype = torch.bfloat16`` is set, in which case
    # y and therefore the cotangent autograd hands back are bf16 as well.  An extra attribute, not part of the state-dict.
    out_dtype = None
    # Opt-in: replay the 9-10 kernel launches of a forward (and of a backward) as one CUDA graph, captured the second time the
    # same buffers (input, output, cotangent and gradient pointers -- the caching allocator cycles through a handful) come by.
    # Cuts ~60 us of host time per step; off by default because a graph pins the addresses it was captured with (a tensor
    # that is freed and re-allocated elsewhere simply causes a new capture, up to 24 per call shape).
    use_cuda_graphs = False

    def __init__(self, in_channels=3, out_channels=3):
        super().__init__()
        self.extractor = ExtractParameters2()
        self.filters = nn.ModuleList(_FilterSlot(n) for n in FILTER_ORDER)

    def _tensors(self):
        """The 14 parameters in state-dict order through plain dict look-ups (no nn.Module.__getattr__: 2 us instead of 30)."""
        ex = self._modules["extractor"]._modules
        out = []
        for blk in ex["conv_layers"]._modules.values():
            c = blk._modules["conv_block"]._modules["0"]._parameters
            out.append(c["weight"])
            out.append(c["bias"])
        f1, f2 = ex["fc1"]._parameters, ex["fc2"]._parameters
        out += [f1["weight"], f1["bias"], f2["weight"], f2["bias"]]
        return out

    def forward(self, x, dedark_A=None, IcA=None):
        from . import ops
        ops.check_image_shape(x)
        params = self._tensors()
        pdev = params[-1].device
        if x.is_cuda:
            if pdev != x.device:
                self.to(x.device)  # llie.py:28: the module follows its input (the no-op traversal is skipped: 0.2 ms of host time)
                params = self._tensors()
            dev = x.device
        else:
            if not torch.cuda.is_available():
                raise RuntimeError(
                    "lowlight_recovery (dedark_yolo_b200): no CUDA device available -- this build has no CPU fallback")
            dev = pdev if pdev.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        return ops.RecoveryFunction.apply(self, dev, x, dedark_A, IcA, *params)

    # -- copies and checkpoints ----------------------------------------------------------------------------------------------
    def __deepcopy__(self, memo):
        """Plain deep copy (ModelEMA, torch_utils.py:353; save_model, trainer.py:413): never goes through ``__reduce_ex__``."""
        new = type(self).__new__(type(self))
        memo[id(self)] = new
        new.__dict__.update(copy.deepcopy(self.__dict__, memo))
        return new

    def __reduce_ex__(self, protocol):
        """While ``integrate.install()`` is active the module pickles itself as a genuine reference module under the
        reference's qualified class name (see integrate.py, "Checkpoints"): ultralytics checkpoints hold whole objects
        (trainer.py:408-433), and one written with the drop-in must load in a vanilla Dedark-YOLO process."""
        from . import integrate
        if integrate.installed() and type(self) is lowlight_recovery:
            import copyreg
            return copyreg.__newobj__, (type(self),), integrate.export_reference_module(self).__dict__
        return super().__reduce_ex__(protocol)
