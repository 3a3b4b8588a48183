"""bf16 I/O mode (north_star: "reads the NCHW fp32/bf16 batch ... within 1e-5 relative (fp32) or 2e-2 (bf16)"; SURVEY.md
section 8(d): x and g read as bf16, y fp32 like the reference unless the caller opts into bf16).

Gate: 2e-2 rel-to-max against the fp64 oracle evaluated on the SAME bf16-rounded inputs (the reference module fed a bf16
tensor promotes it to fp32 and computes in fp32, llie.py:34-40) -- what is left is the TF32 blur and, when requested, the bf16
rounding of y.  The golden cases of the fp32 suite are reused (96x80, caller-supplied A / IcA, the darkened batch, ...).
"""
import pytest
import torch

from conftest import CASES, golden_weights, load_case, rel_to_max
from oracle import lowlight_oracle as O

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2


@pytest.fixture(scope="module")
def dd():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dedark_yolo_b200 as pkg
    return pkg


def _bf(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("name", [c for c in CASES if c != "min13"])   # 13x13: W % 4 != 0, the module promotes to fp32 (tested below)
@pytest.mark.parametrize("out_bf16", [False, True])
def test_filters_bf16_io_vs_fp64_oracle(dd, name, out_bf16):
    from dedark_yolo_b200 import ops
    c = load_case(name)
    x16 = _bf(c["x"])
    if not ops.bf16_io_supported(x16.shape[2], x16.shape[3]):
        pytest.skip("shape not taken by the bf16 kernels")
    feat = c["feat"]
    A, IcA = c["A"], c["IcA"]
    cu = lambda t: None if t is None else t.cuda()  # noqa: E731
    y = ops.filters_forward(cu(x16), cu(feat), cu(A), cu(IcA), out_dtype=torch.bfloat16 if out_bf16 else torch.float32)
    assert y.dtype == (torch.bfloat16 if out_bf16 else torch.float32)
    g16 = _bf(c["g"])
    dfeat, dx = ops.filters_backward(cu(x16), cu(feat), cu(g16) if out_bf16 else cu(g16.float()), cu(A), cu(IcA), need_dx=True)
    assert dx.dtype == torch.bfloat16
    # fp64 truth on the bf16-rounded operands
    xr = x16.double().requires_grad_(True)
    fr = feat.double().requires_grad_(True)
    yr = O.filter_chain(xr, fr, None if A is None else A.double(), None if IcA is None else IcA.double(), dense_blur=False)
    yr.backward(g16.double())
    e_y, e_f, e_x = rel_to_max(y.float().cpu(), yr.detach()), rel_to_max(dfeat.cpu(), fr.grad), rel_to_max(dx.float().cpu(), xr.grad)
    print(f"[bf16] {name} out_bf16={out_bf16}: y {e_y:.2e}, dfeat {e_f:.2e}, dx {e_x:.2e} (gate {BF16_TOL:.0e})")
    assert e_y <= BF16_TOL and e_f <= BF16_TOL and e_x <= BF16_TOL


def test_module_bf16_input_and_bf16_output(dd):
    """The drop-in with a bf16 batch: fp32 output by default (the reference's promotion), bf16 output on request; gradients of
    the 14 predictor tensors within the bf16 gate of the fp64 oracle on the same rounded input."""
    torch.manual_seed(0)
    m = dd.lowlight_recovery(3).cuda().train()
    w = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(21)
    x16 = _bf(torch.rand(2, 3, 96, 80, generator=gen))
    g16 = _bf(torch.randn(2, 3, 96, 80, generator=gen))
    y = m(x16.cuda())
    assert y.dtype == torch.float32
    y_ref, _, _, grads, _ = O.recovery_forward_backward(x16.float(), w, g16.float(), dtype=torch.float64)
    assert rel_to_max(y.detach().cpu(), y_ref) <= BF16_TOL
    m.out_dtype = torch.bfloat16
    try:
        m.zero_grad()
        y2 = m(x16.cuda())
        assert y2.dtype == torch.bfloat16
        y2.backward(g16.cuda())
        assert rel_to_max(y2.detach().float().cpu(), y_ref) <= BF16_TOL
        worst = max(rel_to_max(p.grad.cpu(), grads[k]) for k, p in m.named_parameters())
        print(f"[bf16] module: worst predictor-gradient rel-to-max {worst:.2e}")
        assert worst <= BF16_TOL
    finally:
        m.out_dtype = None
    # shapes the bf16 kernels do not take (W % 4 != 0) are promoted to fp32 like in the reference
    x13 = _bf(torch.rand(1, 3, 13, 13, generator=gen))
    y13 = m(x13.cuda())
    assert y13.dtype == torch.float32 and rel_to_max(y13.detach().cpu(), O.recovery_forward(x13.double(), {k: v.double() for k, v in w.items()}, dense_blur=False)) <= 1e-5


def test_pipeline_bf16_step_full_size(dd):
    """RecoveryPipeline(io_dtype=bf16) at BASELINE configs[1] size: bf16 dark / y / g.  Reference: the fp32 kernels fed the SAME
    bf16-rounded darkened batch and cotangent (what the reference module computes when it is handed bf16 tensors)."""
    from dedark_yolo_b200 import ops
    torch.manual_seed(0)
    m = dd.lowlight_recovery(3).cuda().train()
    params = [p.detach() for p in m.extractor.ordered_parameters()]
    B, H, W = 16, 640, 640
    gen = torch.Generator(device="cuda").manual_seed(5)
    clean = torch.rand(B, 3, H, W, generator=gen, device="cuda")
    g16 = torch.randn(B, 3, H, W, generator=gen, device="cuda").to(torch.bfloat16)
    p16 = dd.RecoveryPipeline(m, B, H, W, dark_param=5.0, io_dtype=torch.bfloat16)
    y16, rec16, flat16 = p16.step(clean, g16)
    torch.cuda.synchronize()
    assert y16.dtype == torch.bfloat16 and p16.dark.dtype == torch.bfloat16
    # the darkened batch is the exact fp32 value rounded to bf16; the loss is reduced from the fp32 values
    dark32 = torch.pow(clean, 5.0)
    assert torch.equal(p16.dark, dark32.to(torch.bfloat16))
    rec_ref = torch.nn.functional.mse_loss(dark32.double(), clean.double())
    assert abs(float(rec16) - float(rec_ref)) <= 1e-6 * float(rec_ref)
    # fp32 kernels on the same rounded operands
    x32, g32 = p16.dark.float(), g16.float()
    r = ops.resize256(x32)
    feat, acts = ops.predictor_forward(r, params)
    y32 = ops.filters_forward(x32, feat)
    dfeat, _ = ops.filters_backward(x32, feat, g32)
    grads, _ = ops.predictor_backward(r, params, acts, dfeat)
    flat32 = torch.cat([t.reshape(-1) for t in grads])
    e_y, e_g = rel_to_max(y16.float().cpu(), y32.cpu()), rel_to_max(flat16.cpu(), flat32.cpu())
    print(f"[bf16] pipeline 16x3x640x640: y {e_y:.2e}, flat gradient {e_g:.2e} vs the fp32 kernels on the same bf16-rounded operands")
    assert e_y <= BF16_TOL and e_g <= BF16_TOL
