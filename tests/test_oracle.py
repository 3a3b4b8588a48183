"""Pins ``oracle/lowlight_oracle.py`` against golden vectors recorded from the real reference
(tests/golden/generate.py).  CPU only.

Tolerances:
  forward, fp32 port vs reference fp32 ........ max|d|/max|ref| <= 1e-5
  forward, fp64 truth vs reference fp32 ....... <= 1e-5
  gradients: the reference's own fp32 autograd is ~2e-5 (rel-to-max) away from fp64 truth, so the
  fp64 truth must agree with the recorded fp32 reference grads to 2e-4 rel-to-max.
  darkening LUT ............................... bit-exact (same torch CPU kernel)
"""
import numpy as np
import pytest
import torch

from conftest import CASES, golden_weights, load_case, load_golden, rel_to_max
from oracle import lowlight_oracle as O

FC1 = "extractor.fc1.weight"


def test_init_weights_match_reference_rng():
    ref = golden_weights()
    mine = O.init_weights(0)
    assert tuple(mine.keys()) == O.STATE_KEYS
    for k, shp in zip(O.STATE_KEYS, O.STATE_SHAPES):
        assert tuple(mine[k].shape) == shp
        assert torch.equal(mine[k], ref[k]), k


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dense", [True, False])
def test_forward_fp32_port(name, dense):
    c = load_case(name)
    w = golden_weights(c["fc2_scale"])
    y, feat = O.recovery_forward(c["x"], w, c["A"], c["IcA"], dense_blur=dense, return_feat=True)
    assert y.dtype == torch.float32 and y.shape == c["y"].shape
    assert rel_to_max(feat, c["feat"]) <= 1e-5
    assert rel_to_max(y, c["y"]) <= 1e-5


@pytest.mark.parametrize("name", CASES)
def test_forward_backward_fp64_truth(name):
    c = load_case(name)
    w = golden_weights(c["fc2_scale"])
    y, feat, dfeat, grads, dx = O.recovery_forward_backward(
        c["x"], w, c["g"], c["A"], c["IcA"], dtype=torch.float64, dense_blur=False, need_dx=True)
    assert rel_to_max(y, c["y"]) <= 1e-5
    assert rel_to_max(dfeat, c["dfeat"]) <= 2e-4
    assert rel_to_max(dx, c["dx"]) <= 2e-4
    # columns 1 (masked WB slot) and 5..12 (tone, not in the chain) never receive gradient
    assert float(dfeat[:, [1, 5, 6, 7, 8, 9, 10, 11, 12]].abs().max()) == 0.0
    for k in O.STATE_KEYS:
        g = grads[k].float()
        if k == FC1:
            sub = g.reshape(-1)[::16]
            assert rel_to_max(sub, c["grads"]["grad_sub." + k]) <= 2e-4
            s = c["grads"]["grad_sum." + k]
            assert abs(float(grads[k].sum()) - float(s[0])) <= 2e-4 * float(s[1])
        else:
            assert rel_to_max(g, c["grads"]["grad." + k]) <= 2e-4, k


def test_resize_matches_aten():
    gen = torch.Generator().manual_seed(3)
    for shp in [(1, 3, 13, 13), (2, 3, 96, 80), (1, 3, 640, 640), (1, 3, 300, 517)]:
        x = torch.rand(shp, generator=gen)
        ref = torch.nn.functional.interpolate(x, size=(256, 256), mode="bilinear", align_corners=False)
        assert rel_to_max(O.resize256(x), ref) <= 2e-6


def test_blur_dense_equals_separable_fp64():
    gen = torch.Generator().manual_seed(4)
    x = torch.rand(1, 3, 40, 29, generator=gen, dtype=torch.float64)
    assert rel_to_max(O.blur_separable(x), O.blur_dense(x)) <= 1e-7  # 2-D taps are fp32-rounded products
    taps = O.gaussian_taps()
    assert abs(float(taps.sum()) - 1.0) < 1e-6
    assert abs(float(taps[12]) - 0.08077993) < 1e-7 and abs(float(taps[24]) - 0.00453456) < 1e-7


def test_size_and_channel_errors():
    w = golden_weights()
    with pytest.raises(RuntimeError):
        O.recovery_forward(torch.rand(1, 3, 12, 12), w, dense_blur=False)
    with pytest.raises(RuntimeError):
        O.recovery_forward(torch.rand(1, 3, 12, 12), w, dense_blur=True)
    with pytest.raises(RuntimeError):
        O.filter_chain(torch.rand(1, 4, 32, 32), torch.zeros(1, 15))
    with pytest.raises(IndexError):
        O.filter_chain(torch.rand(1, 3, 32, 2), torch.zeros(1, 15))


def test_synthesis_lut_bit_exact_and_mse():
    s = load_golden("synth.npz")
    u8 = torch.from_numpy(s["u8"])
    k = torch.arange(256, dtype=torch.uint8)
    for p in (5.0, 7.5, 10.0, 15.0):
        lut = O.synth_darken(O.synth_clean_from_u8(k), p)
        assert np.array_equal(lut.numpy().view(np.int32), s[f"lut_{p}"].view(np.int32))
        clean = O.synth_clean_from_u8(u8)
        dark = O.synth_darken(clean, p)
        assert torch.equal(dark, lut[u8.long()])  # the LUT *is* the darkening for u8-sourced data
        assert abs(float(O.recovery_mse(dark, clean)) - float(s[f"mse_{p}"])) <= 1e-6 * float(s[f"mse_{p}"])
        assert np.array_equal(O.synth_quantize_u8(dark).numpy(), s[f"q_{p}"])
    cf = torch.from_numpy(s["clean_f32"])
    assert np.array_equal(O.synth_darken(cf, 15.0).numpy().view(np.int32), s["dark_f32_15.0"].view(np.int32))


def test_recovery_loss_term():
    t = load_golden("loss_term.npz")
    base_loss, base_items, lrl = torch.tensor(t["base_loss"]), torch.from_numpy(t["base_items"]), float(t["lrl"])
    for name in ("scalar", "vector", "zero"):
        loss, items = O.recovery_loss_term(base_loss.clone(), base_items.clone(), torch.from_numpy(t[f"rec_{name}"]), lrl)
        assert items.shape == (3,)
        assert abs(float(loss) - float(t[f"loss_{name}"])) <= 1e-6 * abs(float(t[f"loss_{name}"]))
        assert np.allclose(items.numpy(), t[f"items_{name}"], rtol=1e-6, atol=0)
    loss, items = O.recovery_loss_term(base_loss.clone(), base_items.clone(), None, lrl)
    assert float(loss) == float(t["loss_absent"]) and np.array_equal(items.numpy(), t["items_absent"])


def test_bus_known_answer():
    """BASELINE config 1: bus.jpg -> 640x640, B=4, weights seed 0; checksums from BASELINE.md section 2."""
    b = load_golden("bus640.npz")
    x = (torch.from_numpy(b["u8"]).float() / 255)[None]
    y = O.recovery_forward(x, golden_weights(), dense_blur=False)
    assert abs(float(b["y_sum"]) - 1901632.45) < 0.5 and abs(float(b["y_mean_abs"]) - 0.571195) < 1e-6
    assert abs(4 * float(y.double().sum()) - float(b["y_sum"])) <= 1e-5 * float(b["y_sum"])
    assert abs(float(y.double().abs().mean()) - float(b["y_mean_abs"])) <= 1e-5
    assert rel_to_max(y[0, :, ::8, ::8], torch.from_numpy(b["y_sub"])) <= 1e-5
    assert rel_to_max(y[0, :, 317:323, :], torch.from_numpy(b["y_rows"])) <= 1e-5


def test_oracle_vs_reference_live():
    """When the reference tree is present (build container), compare on a fresh random case as well."""
    from oracle.reference_loader import load_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree absent (GPU box)")
    ref = load_reference()
    torch.manual_seed(0)
    m = ref.lowlight_recovery(3).eval()
    gen = torch.Generator().manual_seed(99)
    x = torch.rand(1, 3, 50, 61, generator=gen)
    with torch.no_grad():
        yr = m(x)
    w = {k: v.detach() for k, v in m.state_dict().items()}
    assert rel_to_max(O.recovery_forward(x, w, dense_blur=True), yr) <= 1e-5
