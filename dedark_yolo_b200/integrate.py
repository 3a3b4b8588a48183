"""Swap the reference's ``lowlight_recovery`` for the B200 one inside an *unmodified* Dedark-YOLO checkout.

    import dedark_yolo_b200.integrate as it
    it.install()                 # before building DetectionModel('yolov8l.yaml')

``parse_model`` looks the class up by name through ``globals()`` of ``ultralytics.nn.tasks`` (tasks.py:844),
compares by identity (tasks.py:888) and by ``isinstance`` (tasks.py:107); checkpoints pickle it by qualified
name ``ultralytics.nn.modules.llie.lowlight_recovery``.  ``install`` therefore rebinds the name in all three
modules and gives our class that qualified name.
"""
from __future__ import annotations

import importlib
import sys

from .llie import lowlight_recovery


def install(verbose: bool = False):
    """Rebind ``lowlight_recovery`` in ultralytics (must already be importable).  Returns the replaced class."""
    llie = importlib.import_module("ultralytics.nn.modules.llie")
    original = getattr(llie, "lowlight_recovery")
    lowlight_recovery.__module__ = "ultralytics.nn.modules.llie"
    lowlight_recovery.__qualname__ = "lowlight_recovery"
    for name in ("ultralytics.nn.modules.llie", "ultralytics.nn.modules", "ultralytics.nn.tasks"):
        mod = sys.modules.get(name) or importlib.import_module(name)
        if hasattr(mod, "lowlight_recovery"):
            setattr(mod, "lowlight_recovery", lowlight_recovery)
            if verbose:
                print(f"dedark_yolo_b200: patched {name}.lowlight_recovery")
    return original
