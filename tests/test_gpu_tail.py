"""The fused CUDA-core tail of the predictor forward (DEDARK_TAIL=fused, dd_predictor_tail.cuh: conv4 + conv5 + fc1 + fc2 in one
cluster launch, plain fp32 FMA chains) against the shipped tensor-core tail (3xTF32 GEMMs) and torch fp64: two independent
implementations of the same layers (nn/modules/common.py:52-78) agree within the predictor's 2e-6 gate."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B", [1, 3, 16])
def test_fused_tail_matches_tensor_core_tail_and_fp64(B):
    import dedark_yolo_b200 as dd
    from dedark_yolo_b200 import ops

    torch.manual_seed(B)
    m = dd.lowlight_recovery(3).to("cuda")
    params = [p.detach() for p in m.extractor.ordered_parameters()]
    r = torch.rand(B, 3, 256, 256, device="cuda")
    old = os.environ.get("DEDARK_TAIL")
    try:
        os.environ["DEDARK_TAIL"] = "fused"
        n0 = dd.launch_count()
        feat_f, acts_f = ops.predictor_forward(r, params)
        n_fused = dd.launch_count() - n0
        feat_f, acts_f = feat_f.clone(), acts_f.clone()
        os.environ["DEDARK_TAIL"] = "tc"
        n0 = dd.launch_count()
        feat_t, acts_t = ops.predictor_forward(r, params)
        n_tc = dd.launch_count() - n0
    finally:
        if old is None:
            os.environ.pop("DEDARK_TAIL", None)
        else:
            os.environ["DEDARK_TAIL"] = old
    assert n_tc - n_fused == 2   # one launch instead of three
    x = r.double()
    for l in range(5):
        x = F.leaky_relu(F.conv2d(x, params[2 * l].double(), params[2 * l + 1].double(), stride=2, padding=1), 0.1)
    h = F.leaky_relu(F.linear(x.reshape(B, -1), params[10].double(), params[11].double()), 0.1)
    ref = F.linear(h, params[12].double(), params[13].double())
    scale = ref.abs().max().item()
    assert (feat_f.double() - ref).abs().max().item() <= 2e-6 * scale
    assert (feat_t.double() - ref).abs().max().item() <= 2e-6 * scale
    # the activations the backward reads (a1..a5, h: the head of the workspace) agree between the two engines
    n = B * (16 * 128 * 128 + 32 * 64 * 64 + 32 * 32 * 32 + 32 * 16 * 16 + 32 * 8 * 8 + 64)
    amax = acts_t[:n].abs().max().item()
    assert (acts_f[:n] - acts_t[:n]).abs().max().item() <= 4e-6 * amax
