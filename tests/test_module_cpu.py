"""Host-side contract of the drop-in module, checked without a GPU."""
import copy
import io
import pickle

import numpy as np
import pytest
import torch

from conftest import golden_weights, load_golden
from oracle import lowlight_oracle as O

import dedark_yolo_b200 as dd


def test_state_dict_keys_shapes_and_rng_parity():
    torch.manual_seed(0)
    m = dd.lowlight_recovery(3)
    sd = m.state_dict()
    assert tuple(sd.keys()) == O.STATE_KEYS
    ref = golden_weights()
    for k, shp in zip(O.STATE_KEYS, O.STATE_SHAPES):
        assert tuple(sd[k].shape) == shp
        assert torch.equal(sd[k], ref[k]), f"{k}: same seed must give the reference's initial weights"
    assert sum(p.numel() for p in m.parameters()) == 164943
    assert [p.shape for p in m.extractor.ordered_parameters()] == [torch.Size(s) for s in O.STATE_SHAPES]


def test_constructor_signature_and_children():
    for args in [(), (3,), (3, 3)]:
        m = dd.lowlight_recovery(*args)
        assert len(m.filters) == 5 and [f.get_short_name() for f in m.filters] == ["DF", "W", "G", "Ct", "UF"]
        assert isinstance(m.extractor, dd.ExtractParameters2)
    m = dd.lowlight_recovery(in_channels=3, out_channels=3)
    m.i, m.f, m.type, m.np = 0, -1, "lowlight_recovery", 164943  # attributes attached by parse_model (tasks.py:911-912)


def test_load_reference_state_dict_and_roundtrips():
    m = dd.lowlight_recovery(3)
    missing, unexpected = m.load_state_dict(golden_weights(), strict=True)
    assert not missing and not unexpected
    m2 = copy.deepcopy(m)
    buf = io.BytesIO()
    pickle.dump(m, buf)
    m3 = pickle.loads(buf.getvalue())
    for a in (m2, m3):
        for (k1, v1), (k2, v2) in zip(m.state_dict().items(), a.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2)
    h = copy.deepcopy(m).half()
    assert all(p.dtype == torch.float16 for p in h.parameters())
    f = h.float()
    assert all(p.dtype == torch.float32 for p in f.parameters())
    # optimizer grouping by name/type (trainer.py:638-646): biases vs weights are ordinary nn.Parameters
    assert sum(1 for n, _ in m.named_parameters() if n.endswith("bias")) == 7


def test_shape_errors_match_reference():
    m = dd.lowlight_recovery(3)
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 12, 12))
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 640, 12))
    with pytest.raises(IndexError):
        m(torch.rand(1, 3, 32, 2))
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 4, 32, 32))
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 3, 32, 32, dtype=torch.float64))


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without CUDA")
def test_no_cpu_fallback():
    m = dd.lowlight_recovery(3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.rand(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        dd.preprocess_batch({"img": torch.zeros(1, 3, 16, 16, dtype=torch.uint8)}, "cpu")


def test_add_recovery_term_matches_reference_numbers():
    t = load_golden("loss_term.npz")
    base_loss, base_items, lrl = torch.tensor(t["base_loss"]), torch.from_numpy(t["base_items"]), float(t["lrl"])
    for name in ("scalar", "vector", "zero"):
        loss, items = dd.add_recovery_term(base_loss.clone(), base_items.clone(),
                                           {"recovery_loss_batch": torch.from_numpy(t[f"rec_{name}"])}, lrl)
        assert items.shape == (3,)
        assert abs(float(loss) - float(t[f"loss_{name}"])) <= 1e-6 * abs(float(t[f"loss_{name}"]))
        assert np.allclose(items.numpy(), t[f"items_{name}"], rtol=1e-6, atol=0)
    loss, items = dd.add_recovery_term(base_loss.clone(), base_items.clone(), {}, lrl)
    assert float(loss) == float(t["loss_absent"]) and np.array_equal(items.numpy(), t["items_absent"])


def test_shard_range_partitions_the_batch():
    from dedark_yolo_b200.dist import shard_range
    for n in (1, 7, 16, 64, 257):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)
