"""Golden vector for the dark-channel prior -- executes the REAL reference functions.

    python tests/golden/generate_prior.py        (build container only: reads /root/reference)

DetectionTrainer.DarkChannel / AtmLight (models/yolo/detect/train.py:42-61) are compiled from the reference file's own AST
(the module itself cannot be imported without the whole ultralytics tree and its plotting dependencies) and run on crafted
images whose selected dark-channel values are distinct, so that the result does not depend on how argsort breaks ties.
Only inputs and OUTPUTS are recorded (tests/golden/prior.npz); no reference source is copied.
"""
import ast
import math
import os

import cv2
import numpy as np

REF = "/root/reference/ultralytics/models/yolo/detect/train.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_functions():
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DetectionTrainer")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("DarkChannel", "AtmLight", "DarkIcA")]
    mod = ast.Module(body=fns, type_ignores=[])
    ns = {"cv2": cv2, "np": np, "math": math}
    exec(compile(mod, REF, "exec"), ns)
    return ns


def craft(rng, H, W, n_top):
    """HWC uint8 image: background <= 200, n_top pixels whose channel minimum takes the distinct values 255, 254, ..."""
    im = rng.integers(0, 201, size=(H, W, 3), dtype=np.uint8)
    pos = rng.choice(H * W, size=n_top, replace=False)
    for i, q in enumerate(pos):
        v = 255 - i
        px = rng.integers(v, 256, size=3)
        px[rng.integers(0, 3)] = v
        im[q // W, q % W] = px.astype(np.uint8)
    return im


def main():
    ns = reference_functions()

    class Self:
        DarkChannel = staticmethod(lambda im: ns["DarkChannel"](None, im))

    rng = np.random.default_rng(20240607)
    out = {}
    for name, (H, W) in {"a": (100, 100), "b": (64, 250), "c": (13, 16)}.items():
        numpx = max(H * W // 1000, 1)
        im = craft(rng, H, W, numpx + 3)
        dark = ns["DarkChannel"](Self, im)
        A = ns["AtmLight"](Self, im, dark)        # [1,3] float64, uint8 units
        out[f"{name}.img_chw"] = np.ascontiguousarray(im.transpose(2, 0, 1))
        out[f"{name}.A_u8"] = np.asarray(A, dtype=np.float64).reshape(3)
    np.savez_compressed(os.path.join(HERE, "prior.npz"), **out)
    for k, v in out.items():
        if k.endswith("A_u8"):
            print(k, v)


if __name__ == "__main__":
    main()
