"""Swap the reference's ``lowlight_recovery`` for the B200 one inside an *unmodified* Dedark-YOLO checkout.

    import ultralytics                                   # the unmodified reference
    import dedark_yolo_b200.integrate as it
    it.install()                 # before building DetectionModel('yolov8l.yaml')
    ...
    it.uninstall()               # restores the reference class

``parse_model`` looks the class up by name through ``globals()`` of ``ultralytics.nn.tasks`` (tasks.py:844),
compares by identity (tasks.py:888) and by ``isinstance`` (tasks.py:107); checkpoints pickle it by qualified
name ``ultralytics.nn.modules.llie.lowlight_recovery``.  ``install`` therefore rebinds the name in all three
modules and gives our class that qualified name.

Checkpoints.  ultralytics pickles the WHOLE model object (trainer.py:408-433), not a state-dict.  While installed, the
drop-in therefore pickles itself as a *genuine reference module*: ``__reduce_ex__`` (llie.py) builds an instance of the
original class with the same 14 tensors (``export_reference_module``) and writes that object's ``__dict__`` under the
qualified class name.  Such a checkpoint un-pickles in a vanilla Dedark-YOLO process (no dedark_yolo_b200 installed, no
GPU) as the reference class with reference children, and in a process that called ``install()`` as the drop-in (whose
forward only reads the 14 parameters, whatever class holds them).
"""
from __future__ import annotations

import importlib
import sys

from .llie import lowlight_recovery

_PATCHED = ("ultralytics.nn.modules.llie", "ultralytics.nn.modules", "ultralytics.nn.tasks")
_original = None  # the reference class, while installed


def installed() -> bool:
    return _original is not None


def install(verbose: bool = False):
    """Rebind ``lowlight_recovery`` in ultralytics (must already be importable).  Returns the replaced class."""
    global _original
    llie = importlib.import_module("ultralytics.nn.modules.llie")
    current = getattr(llie, "lowlight_recovery")
    if current is lowlight_recovery:
        return _original
    _original = current
    lowlight_recovery.__module__ = "ultralytics.nn.modules.llie"
    lowlight_recovery.__qualname__ = "lowlight_recovery"
    for name in _PATCHED:
        mod = sys.modules.get(name) or importlib.import_module(name)
        if hasattr(mod, "lowlight_recovery"):
            setattr(mod, "lowlight_recovery", lowlight_recovery)
            if verbose:
                print(f"dedark_yolo_b200: patched {name}.lowlight_recovery")
    return _original


def uninstall() -> None:
    """Undo ``install``: the reference class is bound again and the drop-in gets its own qualified name back."""
    global _original
    if _original is None:
        return
    for name in _PATCHED:
        mod = sys.modules.get(name)
        if mod is not None and getattr(mod, "lowlight_recovery", None) is lowlight_recovery:
            setattr(mod, "lowlight_recovery", _original)
    lowlight_recovery.__module__ = "dedark_yolo_b200.llie"
    lowlight_recovery.__qualname__ = "lowlight_recovery"
    _original = None


def export_reference_module(m: lowlight_recovery):
    """A genuine reference ``lowlight_recovery`` (the class ``install`` replaced) holding copies of ``m``'s 14 tensors, on
    the same device, in the same dtype and train/eval mode, with the attributes ``parse_model`` attaches (i, f, type, np;
    tasks.py:911-912)."""
    if _original is None:
        raise RuntimeError("export_reference_module needs install() (the reference class must be importable)")
    ref = _original(3)
    ref.load_state_dict(m.state_dict())
    p = next(m.parameters())
    ref.to(device=p.device, dtype=p.dtype)
    ref.train(m.training)
    for k, v in m.__dict__.items():
        if not k.startswith("_") and k != "training":
            setattr(ref, k, v)
    return ref
