"""TEST INFRASTRUCTURE ONLY -- loads the *real* reference modules for golden-vector generation.

This file imports the unmodified reference hot-path files from ``/root/reference`` (read-only
checkout of cvYouTian/Dedark-YOLO).  That tree exists only in the build container, never on the
GPU box, so nothing in ``tests/ -m gpu``, ``bench.py`` or ``__graft_entry__.smoke()`` may call
this at run time.  Its single consumer is ``tests/golden/generate.py`` (and the
``test_oracle_vs_reference_live`` test, which skips when the tree is absent).

Two shims are needed ("Recipe A"):
  * ``easydict`` is not installed            -> a tiny attribute-dict stand-in;
  * ``filter_cfg.py:6-7`` calls ``argparse.parse_args()`` at import time
                                              -> ``sys.argv`` is trimmed while importing;
  * ``ultralytics/__init__`` pulls matplotlib -> the five files are imported through a synthetic
    package whose ``__path__`` points straight at ``ultralytics/nn/modules``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DEDARK_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_ROOT, "ultralytics")):
    # the byte-for-byte copy made by baseline/install_reference.py (git-ignored; it travels to the GPU box)
    REFERENCE_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
_MODULES_DIR = os.path.join(REFERENCE_ROOT, "ultralytics", "nn", "modules")
_PKG = "_dedark_reference_modules"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_MODULES_DIR, "llie.py"))


class _AttrDict(dict):
    """Stand-in for easydict.EasyDict (attribute access on a dict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def load_reference():
    """Return a namespace with the reference's ``llie``, ``common``, ``filtersB``, ``util_filters``,
    ``filter_cfg`` modules, imported byte-for-byte from ``/root/reference``."""
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    if _PKG + ".llie" in sys.modules:
        pkg = sys.modules[_PKG]
    else:
        if "easydict" not in sys.modules:
            shim = types.ModuleType("easydict")
            shim.EasyDict = _AttrDict
            sys.modules["easydict"] = shim
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [_MODULES_DIR]
        sys.modules[_PKG] = pkg
        argv = sys.argv
        sys.argv = argv[:1]
        try:
            for name in ("util_filters", "filtersB", "filter_cfg", "common", "llie"):
                setattr(pkg, name, importlib.import_module(f"{_PKG}.{name}"))
        finally:
            sys.argv = argv
    ns = types.SimpleNamespace()
    for name in ("util_filters", "filtersB", "filter_cfg", "common", "llie"):
        setattr(ns, name, sys.modules[f"{_PKG}.{name}"])
    ns.lowlight_recovery = ns.llie.lowlight_recovery
    return ns
