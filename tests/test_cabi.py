"""CPU-side checks of the C-ABI boundary: the library loads without a GPU, exports every symbol the header
declares, sizes workspaces sanely and rejects bad shapes with the reference's error classes -- no kernel runs."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "dedark_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dd_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_header_symbols():
    from dedark_yolo_b200 import _lib
    names = _declared()
    assert len(names) >= 11
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/dedark_b200.h but not exported"
    assert set(_lib.EXPORTS) == set(names)
    assert _lib.lib.dd_version() >= 100
    assert _lib.launch_count() >= 0


def test_workspace_sizes():
    from dedark_yolo_b200 import _lib
    acts = _lib.workspace_bytes(_lib.WS_PREDICTOR_ACTS, 16)
    per_img = 16 * 128 * 128 + 32 * 64 * 64 + 32 * 32 * 32 + 32 * 16 * 16 + 32 * 8 * 8 + 64
    # + the prepared tensor-core weights of conv2..conv5: (18 + 32) * Cin * Cout floats per layer (hi and lo halves of the
    # forward and data-gradient operand layouts), independent of the batch size
    prep = (18 + 32) * (16 * 32 + 3 * 32 * 32)
    assert acts == (16 * per_img + prep) * 4
    assert _lib.workspace_bytes(_lib.WS_PREDICTOR_ACTS, 1) == (per_img + prep) * 4
    assert _lib.workspace_bytes(_lib.WS_PREDICTOR_BWD, 16) > 16 * per_img * 4
    assert _lib.workspace_bytes(_lib.WS_SYNTH, 1) >= 8 * 148
    small = _lib.workspace_bytes(_lib.WS_RECOVERY_BWD, 1, 13, 13)
    big = _lib.workspace_bytes(_lib.WS_RECOVERY_BWD, 16, 640, 640)
    assert 0 < small < big
    assert big >= 16 * 3 * 640 * 5 * 4  # per-row partial sums: B*3*H*strips floats
    assert _lib.workspace_bytes(99, 1, 1, 1) == 0 and _lib.workspace_bytes(_lib.WS_SYNTH, 0) == 0


def test_shape_errors_mirror_reference_without_touching_the_gpu():
    from dedark_yolo_b200 import _lib
    one = C.c_void_p(16)  # never dereferenced: validation fails first
    with pytest.raises(RuntimeError, match="reflect"):
        _lib.check(_lib.lib.dd_recovery_fwd(one, None, None, one, C.c_void_p(32), 1, 12, 12, None))
    with pytest.raises(RuntimeError, match="reflect"):
        _lib.check(_lib.lib.dd_recovery_bwd(one, None, None, one, one, one, None, 1, 640, 12, one, 1 << 30, None))
    with pytest.raises(IndexError):
        _lib.check(_lib.lib.dd_recovery_fwd(one, None, None, one, C.c_void_p(32), 1, 32, 2, None))
    with pytest.raises(ValueError):
        _lib.check(_lib.lib.dd_recovery_fwd(None, None, None, one, one, 1, 32, 32, None))
    with pytest.raises(ValueError):
        _lib.check(_lib.lib.dd_synth_fwd(one, 7, 1.0, None, None, None, one, None, None, 16, None, 0, None))
    with pytest.raises(RuntimeError, match="workspace"):
        _lib.check(_lib.lib.dd_recovery_bwd(one, None, None, one, one, one, None, 1, 64, 64, one, 8, None))
    assert b"workspace" in _lib.lib.dd_last_error()


def test_peer_exchange_descriptor_is_validated_without_touching_the_gpu():
    from dedark_yolo_b200 import _lib
    assert _lib.lib.dd_exchange_bytes() == 256 + 2 * 8 * 164944 * 8  # header + [2 parities][8 peers] slots of {value, tag} words
    one = C.c_void_p(16)  # never dereferenced: validation fails first
    t = _lib.PredictorTensors.from_tensors([type("T", (), {"data_ptr": lambda self: 16})() for _ in range(14)])
    bad = _lib.PeerExchange()
    bad.rank, bad.world = 3, 2  # rank outside the world
    bad.buf[0], bad.buf[1] = 16, 16
    with pytest.raises(ValueError):
        _lib.check(_lib.lib.dd_predictor_bwd_allreduce(one, C.byref(t), one, one, C.byref(t), 1, one, 1 << 30, C.byref(bad), None))
    hole = _lib.PeerExchange.from_pointers(0, [16, 16])
    hole.buf[1] = None  # a peer buffer that was never mapped
    with pytest.raises(ValueError):
        _lib.check(_lib.lib.dd_predictor_bwd_allreduce(one, C.byref(t), one, one, C.byref(t), 1, one, 1 << 30, C.byref(hole), None))
    with pytest.raises(ValueError):
        _lib.PeerExchange.from_pointers(0, list(range(1, 10)))  # more than 8 peers
