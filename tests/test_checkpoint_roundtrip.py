"""ultralytics checkpoints pickle the whole model, so a checkpoint written with the drop-in installed
must load in a vanilla Dedark-YOLO process -- one that cannot import dedark_yolo_b200 at all -- as the reference class
with reference children.  CPU only (two helper processes, tests/_ckpt_roundtrip.py)."""
import os
import subprocess
import sys

import pytest
import torch

from baseline import reference_runtime as R
from conftest import golden_weights, rel_to_max
from oracle import lowlight_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    res = subprocess.run([sys.executable, os.path.join(HERE, "_ckpt_roundtrip.py"), *args], capture_output=True, text=True, env=env,
                         timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    return res.stdout


def test_checkpoint_from_dropin_loads_in_vanilla_reference(tmp_path):
    if not R.available():
        pytest.skip("baseline/_ref missing: run baseline/install_reference.py in the build container")
    ckpt, out = str(tmp_path / "ckpt.pkl"), str(tmp_path / "out.pt")
    assert "written by ultralytics.nn.modules.llie dedark_yolo_b200.llie" in _run("write", ckpt)
    assert "read as ultralytics.nn.modules.llie" in _run("read", ckpt, out)
    got = torch.load(out)
    # the weights survived (through the fp16 the trainer saves in) and the un-pickled reference module computes with them
    w = golden_weights()
    for k, v in got["state"].items():
        assert torch.equal(v, w[k].half().float()), k
    y_ref = O.recovery_forward(got["x"].double(), {k: v.half().double() for k, v in w.items()}, dense_blur=False)
    assert rel_to_max(got["y"], y_ref) <= 1e-5
