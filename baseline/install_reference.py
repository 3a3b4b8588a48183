"""Puts the UNMODIFIED reference (cvYouTian/Dedark-YOLO) under git-ignored ``baseline/_ref/`` so that it travels to the
GPU box with the gpurun snapshot (SURVEY.md section 7 operational notes).

    python baseline/install_reference.py            # run in the build container (needs /root/reference)

The reference has no setup.py / pyproject, so "installing" it is copying its ``ultralytics``
package byte for byte (1.7 MB, pure Python + yaml + bus.jpg).  Nothing is edited; ``MANIFEST.json`` records the sha256
of every file so a reader can check that the timed reference is the stock one.  ``baseline/_ref/`` is listed in
``.gitignore`` (the sources never enter this repository's history) but not in ``.gpurunignore``.
``__graft_entry__.build()`` calls ``install()`` whenever ``/root/reference`` is present.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
REF_SRC = os.environ.get("DEDARK_REFERENCE_ROOT", "/root/reference")


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def install(force: bool = False) -> str | None:
    """Copy ``<reference>/ultralytics`` to ``baseline/_ref/ultralytics``.  Returns the destination, or None when the
    reference checkout is not on this machine (the GPU box: it then uses the copy that travelled with the snapshot)."""
    src = os.path.join(REF_SRC, "ultralytics")
    if not os.path.isdir(src):
        return REF_DST if os.path.isdir(os.path.join(REF_DST, "ultralytics")) else None
    dst = os.path.join(REF_DST, "ultralytics")
    manifest_path = os.path.join(REF_DST, "MANIFEST.json")
    if os.path.isfile(manifest_path) and not force:
        return REF_DST
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(REF_DST, exist_ok=True)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    files = {}
    for root, _, names in os.walk(dst):
        for n in sorted(names):
            p = os.path.join(root, n)
            files[os.path.relpath(p, REF_DST)] = _sha(p)
    with open(manifest_path, "w") as f:
        json.dump({"source": REF_SRC, "files": files, "modified": False}, f, indent=1, sort_keys=True)
    return REF_DST


if __name__ == "__main__":
    print(install(force=True))
