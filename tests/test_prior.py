"""SURVEY.md section 8(f) N3 -- the dark-channel prior.

CPU: the numpy restatement (oracle/dark_prior_oracle.py) against the reference's own AtmLight outputs (tests/golden/prior.npz,
recorded by tests/golden/generate_prior.py from the real train.py:42-61) and its defining properties.
GPU: ``dd_dark_prior`` through the C-ABI against the restatement, bit for bit.
"""
import os

import numpy as np
import pytest
import torch

from oracle import dark_prior_oracle as P

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prior.npz")


def identity_lut():
    # (k + 0.5) / 255 * 255 truncates to k for every k: the "darkened" image is the source itself
    return ((np.arange(256, dtype=np.float64) + 0.5) / 255.0).astype(np.float32)


def test_oracle_atm_light_matches_reference_outputs():
    g = np.load(GOLD)
    for name in ("a", "b", "c"):
        img = g[f"{name}.img_chw"]
        assert np.array_equal(P.dark_table_u8(identity_lut()), np.arange(256, dtype=np.uint8))
        got = P.atm_light(img).astype(np.float64)
        ref = g[f"{name}.A_u8"]
        assert np.allclose(got, ref, rtol=0, atol=1e-4), (name, got, ref)   # float32 of an exact rational


def test_oracle_tie_shares_and_black_image():
    # all pixels equal: every pixel ties at the threshold -> A = value * (numpx - 1) / numpx
    img = np.full((3, 40, 50), 200, dtype=np.uint8)
    numpx = 40 * 50 // 1000
    assert np.allclose(P.atm_light(img), 200.0 * (numpx - 1) / numpx)
    A, ica = P.dark_prior(np.zeros((1, 3, 20, 24), dtype=np.uint8), identity_lut())
    assert np.all(A == 0) and np.all(ica == 0) and ica.shape == (1, 1, 20, 24)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,p", [((2, 3, 100, 100), 1.0), ((3, 3, 64, 252), 5.0), ((2, 3, 37, 53), 2.5), ((16, 3, 640, 640), 15.0)])
def test_gpu_dark_prior_matches_restatement(shape, p):
    import dedark_yolo_b200 as dd
    from dedark_yolo_b200 import ops

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(sum(shape))
    src = torch.randint(0, 256, shape, dtype=torch.uint8, generator=gen)
    # the device's own darkening table (bit-identical to torch.pow on this GPU, tests/test_gpu_parity.py) feeds the restatement,
    # so a 1-ulp CPU/GPU difference in pow cannot flip a truncation
    k = torch.arange(256, dtype=torch.uint8, device=dev)
    lut = torch.pow(k.float() / 255, p)
    n0 = dd.launch_count()
    A, ica = ops.dark_prior(src.to(dev), p)
    A2, ica2 = ops.dark_prior(src.to(dev), p, lut=lut)
    torch.cuda.synchronize()
    assert dd.launch_count() - n0 == 6
    refA, refI = P.dark_prior(src.numpy(), lut.cpu().numpy())
    assert torch.equal(A, A2) and torch.equal(ica, ica2)
    assert np.array_equal(A.cpu().numpy(), refA), (A.cpu().numpy(), refA)
    assert np.array_equal(ica.cpu().numpy(), refI)


@pytest.mark.gpu
def test_gpu_dark_prior_golden_and_preprocess_batch():
    import dedark_yolo_b200 as dd
    from dedark_yolo_b200 import ops

    dev = torch.device("cuda:0")
    g = np.load(GOLD)
    lut = torch.from_numpy(identity_lut()).to(dev)
    for name in ("a", "b"):
        img = torch.from_numpy(g[f"{name}.img_chw"])[None].to(dev)
        A, _ = ops.dark_prior(img, 1.0, lut=lut)
        assert np.allclose(A.cpu().numpy()[0] * 255.0, g[f"{name}.A_u8"], rtol=0, atol=1e-3)
    u8 = torch.randint(0, 256, (2, 3, 64, 80), dtype=torch.uint8)
    plain = dd.preprocess_batch({"img": u8.clone()}, dev, dark_param=5.0)
    assert plain["dedark_A"] is None and plain["IcA"] is None           # default: the module's A = 0.8, IcA = 0.5
    b = dd.preprocess_batch({"img": u8.clone()}, dev, dark_param=5.0, dedark_prior=True)
    assert b["dedark_A"].shape == (2, 3) and b["IcA"].shape == (2, 1, 64, 80) and b["IcA"].dtype == torch.float32
    m = dd.lowlight_recovery(3).to(dev).eval()
    with torch.no_grad():
        y = m(b["img"], b["dedark_A"], b["IcA"])                           # the eval-mode call of tasks.py:107-110
    assert y.shape == (2, 3, 64, 80) and torch.isfinite(y).all()
