"""pytest config: registers the ``gpu`` marker and shared golden-fixture helpers."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


def golden_weights(fc2_scale=1.0):
    w = {k: torch.from_numpy(v) for k, v in load_golden("weights_seed0.npz").items()}
    if fc2_scale != 1.0:
        w["extractor.fc2.weight"] = w["extractor.fc2.weight"] * fc2_scale
        w["extractor.fc2.bias"] = w["extractor.fc2.bias"] * fc2_scale
    return w


CASES = ["small_default", "small_custom", "min13", "dark", "wide_custom"]


def load_case(name):
    c = load_golden(f"case_{name}.npz")
    t = {k: torch.from_numpy(v) for k, v in c.items() if not k.startswith("grad") and k != "fc2_scale"}
    t["grads"] = {k: torch.from_numpy(v) for k, v in c.items() if k.startswith("grad")}
    t["fc2_scale"] = float(c["fc2_scale"])
    t["A"] = t.get("A")
    t["IcA"] = t.get("IcA")
    return t


def rel_to_max(a, b):
    """max|a-b| / max|b| -- the parity metric of SURVEY.md section 8(d)."""
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
