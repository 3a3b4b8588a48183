"""Builds ``lib/libdedark_b200.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python dedark_yolo_b200/build.py [--force] [-v]   (run as a script: importing the package needs the built library)

The library links the CUDA runtime statically (nvcc 12.9 is newer than torch's bundled 12.8 runtime;
both sit on the same driver context, and streams are passed explicitly), so it has no dependency on
torch and loads through ``ctypes`` even on a machine without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdedark_b200.so")
SOURCES = ["dd_api.cu", "dd_synth.cu", "dd_predictor.cu", "dd_recovery_fwd.cu", "dd_recovery_bwd.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
LINK_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]
OBJ_DIR = os.path.join(HERE, "..", "build", "obj")


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(cand):
        raise RuntimeError("nvcc not found: cannot build libdedark_b200.so")
    return cand


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "dedark_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source (one nvcc process per file, in parallel) and link one shared library."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    for src in SOURCES:
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *(["-Xptxas", "-v"] if verbose else []), "-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for src, obj, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose:
            print(out)
        objs.append(obj)
    res = subprocess.run([nvcc, *LINK_FLAGS, "-o", LIB_PATH + ".tmp", *objs], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    os.replace(LIB_PATH + ".tmp", LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
