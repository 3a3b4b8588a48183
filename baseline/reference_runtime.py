"""Imports the UNMODIFIED reference from ``baseline/_ref`` (see ``install_reference.py``) -- the reference arm of
``bench.py`` and the integration tests.  Never imported by the product package (``dedark_yolo_b200``).

Two recipes:
  A. ``load_modules()``  -- the five hot-path files (llie, common, filter_cfg, filtersB, util_filters) through a synthetic
     package, ~1 s.  Shims: ``easydict`` (not installed) and ``sys.argv`` (filter_cfg.py:6-7 parses argv at import).
  B. ``load_ultralytics()`` -- the whole ``ultralytics`` package (DetectionModel, RcoveryDetectionLoss, get_cfg), ~15 s.
     Additional shims: matplotlib / seaborn / thop are not installed and only used for plotting / FLOP counting.
"""
from __future__ import annotations

import importlib
import os
import sys
import types
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
_PKG = "_dedark_reference_modules"
HOT_FILES = ("util_filters", "filtersB", "filter_cfg", "common", "llie")


def available(root: str | None = None) -> bool:
    return os.path.isfile(os.path.join(root or REF_ROOT, "ultralytics", "nn", "modules", "llie.py"))


class _AttrDict(dict):
    """Stand-in for easydict.EasyDict (attribute access on a dict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _shim_easydict():
    if "easydict" not in sys.modules:
        try:
            importlib.import_module("easydict")
        except ImportError:
            shim = types.ModuleType("easydict")
            shim.EasyDict = _AttrDict
            sys.modules["easydict"] = shim


def load_modules(root: str | None = None):
    """Recipe A.  Returns a namespace with the reference's llie / common / filtersB / util_filters / filter_cfg modules
    and ``lowlight_recovery``."""
    root = root or REF_ROOT
    if not available(root):
        raise FileNotFoundError(f"reference not found under {root}: run baseline/install_reference.py in the build container")
    if _PKG + ".llie" not in sys.modules:
        _shim_easydict()
        pkg = types.ModuleType(_PKG)
        pkg.__path__ = [os.path.join(root, "ultralytics", "nn", "modules")]
        sys.modules[_PKG] = pkg
        argv = sys.argv
        sys.argv = argv[:1]
        try:
            for name in HOT_FILES:
                setattr(pkg, name, importlib.import_module(f"{_PKG}.{name}"))
        finally:
            sys.argv = argv
    ns = types.SimpleNamespace(**{name: sys.modules[f"{_PKG}.{name}"] for name in HOT_FILES})
    ns.lowlight_recovery = ns.llie.lowlight_recovery
    ns.root = sys.modules[_PKG].__path__[0]
    return ns


def load_ultralytics(root: str | None = None):
    """Recipe B.  ``import ultralytics`` from the reference copy; returns a namespace with DetectionModel, get_cfg,
    DEFAULT_CFG, the yaml path of the (only buildable) scale-l model and the package itself."""
    root = root or REF_ROOT
    if not available(root):
        raise FileNotFoundError(f"reference not found under {root}: run baseline/install_reference.py in the build container")
    if "ultralytics" in sys.modules:
        u = sys.modules["ultralytics"]
        if not os.path.abspath(getattr(u, "__file__", "")).startswith(os.path.abspath(root)):
            raise RuntimeError(f"another ultralytics is already imported from {u.__file__}")
    else:
        _shim_easydict()
        for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "seaborn", "thop"):
            if name not in sys.modules:
                try:
                    importlib.import_module(name)
                except Exception:  # noqa: BLE001 -- not installed in this image: plotting / FLOP counting only
                    sys.modules[name] = MagicMock()
        os.environ.setdefault("YOLO_VERBOSE", "false")
        sys.path.insert(0, root)
        argv = sys.argv
        sys.argv = argv[:1]
        try:
            importlib.import_module("ultralytics")
            importlib.import_module("ultralytics.nn.tasks")
        finally:
            sys.argv = argv
    tasks = importlib.import_module("ultralytics.nn.tasks")
    cfg = importlib.import_module("ultralytics.cfg")
    utils = importlib.import_module("ultralytics.utils")
    llie = importlib.import_module("ultralytics.nn.modules.llie")
    return types.SimpleNamespace(
        ultralytics=sys.modules["ultralytics"], tasks=tasks, DetectionModel=tasks.DetectionModel, get_cfg=cfg.get_cfg,
        DEFAULT_CFG=utils.DEFAULT_CFG, llie=llie, lowlight_recovery=llie.lowlight_recovery,
        yaml_l=os.path.join(root, "ultralytics", "cfg", "models", "v8", "yolov8l.yaml"), root=root)
