"""GPU parity tests (run on the B200 box with ``-m gpu``): every CUDA entry point, called through the C-ABI,
against the oracle / the golden vectors recorded from the reference.

Tolerances:
  darkening .......... bit-exact vs torch.pow on the same GPU; bit-exact vs the reference CPU bits with a host LUT
  rec scalar ......... 1e-6 relative
  forward fp32 ....... max|d| / max|ref| <= 1e-5
  gradients .......... vs the fp64 oracle, <= 1e-4 rel-to-max (the reference's own fp32 autograd is ~2e-5 off)
"""
import numpy as np
import pytest
import torch

from conftest import CASES, golden_weights, load_case, load_golden, rel_to_max
from oracle import lowlight_oracle as O

pytestmark = pytest.mark.gpu

FWD_TOL = 1e-5
GRAD_TOL = 1e-4


@pytest.fixture(scope="module")
def dd():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import dedark_yolo_b200 as pkg
    return pkg


@pytest.fixture(scope="module")
def ops(dd):
    from dedark_yolo_b200 import ops as o
    return o


def cuda_params(weights):
    return [weights[k].cuda().contiguous() for k in O.STATE_KEYS]


def ops_tables(p):
    from dedark_yolo_b200 import ops as o
    return o.reference_cpu_tables(p, "cuda")


def grad_tol(ref_fp32, truth64):
    """error <= max(1e-4 rel-to-max, 2x the reference's own fp32 error vs the fp64 truth)."""
    return max(GRAD_TOL, 2.0 * rel_to_max(ref_fp32, truth64))


def report(name, got, ref, tol):
    err = rel_to_max(got.detach().cpu(), ref.detach().cpu())
    print(f"[parity] {name}: rel-to-max {err:.3e} (tol {tol:.0e})")
    assert err <= tol, f"{name}: {err:.3e} > {tol:.0e}"
    return err


# ---- a1 / a2 ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", [5.0, 7.5, 10.0, 15.0, 2.0, 3.0, 0.5])
def test_synth_u8_bit_exact_vs_torch_cuda(ops, p):
    gen = torch.Generator().manual_seed(2024)
    u8 = torch.randint(0, 256, (2, 3, 250, 333), dtype=torch.uint8, generator=gen).cuda()  # numel % 16 != 0
    clean, dark, q, rec = ops.synth_forward(u8, p, want_u8=True)
    ref_clean = u8.float() / 255
    ref_dark = torch.pow(ref_clean, p)
    assert torch.equal(clean.view(torch.int32), ref_clean.view(torch.int32))
    assert torch.equal(dark.view(torch.int32), ref_dark.view(torch.int32)), "darkening must be bit-exact"
    assert torch.equal(q, (ref_dark * 255).to(torch.uint8))
    ref_rec = torch.nn.functional.mse_loss(ref_dark.double(), ref_clean.double())
    assert abs(float(rec) - float(ref_rec)) <= 1e-6 * float(ref_rec)


@pytest.mark.parametrize("p", [5.0, 15.0, 2.0])
def test_synth_f32_bit_exact_vs_torch_cuda(ops, p):
    gen = torch.Generator().manual_seed(1234)
    clean = torch.rand(2, 3, 101, 77, generator=gen).cuda()
    c2, dark, q, rec = ops.synth_forward(clean, p, want_u8=True)
    ref = torch.pow(clean, p)
    assert c2 is clean or torch.equal(c2, clean)
    assert torch.equal(dark.view(torch.int32), ref.view(torch.int32))
    assert torch.equal(q, (ref * 255).to(torch.uint8))
    ref_rec = torch.nn.functional.mse_loss(ref.double(), clean.double())
    assert abs(float(rec) - float(ref_rec)) <= 1e-6 * float(ref_rec)


@pytest.mark.parametrize("p", [15.0, 5.0, 7.5, 10.0, 1.5, 2.2, 0.45, 4.999999, 1e-3, 977.0, 2.0 ** -20, 2.0 ** 20])
def test_synth_f32_every_float_in_unit_interval_bit_exact(ops, p):
    """The fp32 synthesis evaluates the main path of powf inline (dd_synth.cu: powf_unit).  EVERY bit pattern from +0 to 2.0
    -- all of [0, 1], the denormals and the (1, 2) values that take the library call -- must give torch.pow's bits."""
    chunk = 1 << 26
    for k in range(16):
        bits = torch.arange(k * chunk, (k + 1) * chunk, dtype=torch.int32, device="cuda")
        x = bits.view(torch.float32)
        _, dark, _, _ = ops.synth_forward(x, p)
        ref = torch.pow(x, p)
        same = dark.view(torch.int32) == ref.view(torch.int32)
        if not bool(same.all()):
            bad = (~same).nonzero()[:4].flatten().tolist()
            raise AssertionError(f"p={p}: {int((~same).sum())} mismatches in chunk {k}, e.g. " +
                                 ", ".join(f"x={float(x[i])!r} got {float(dark[i])!r} ref {float(ref[i])!r}" for i in bad))


@pytest.mark.parametrize("p", [15.0, 2.5, 3.0, 0.5, -1.0, -2.0, -0.5, 0.0, -3.3, float("inf")])
def test_synth_f32_out_of_domain_inputs_follow_torch_pow(ops, p):
    """Negative, > 1, denormal, infinite and NaN inputs (never produced by an image, but legal) and non-positive or special
    exponents take the library path: same bits as torch.pow, NaN for NaN."""
    gen = torch.Generator().manual_seed(7)
    x = torch.cat([torch.randn(4096, generator=gen) * 3, torch.rand(4096, generator=gen),
                   torch.tensor([0.0, -0.0, 1.0, -1.0, float("inf"), -float("inf"), float("nan"), 1e-40, -1e-40, 1e-38, 2.0, 0.5]),
                   torch.zeros(4)]).cuda()
    _, dark, _, _ = ops.synth_forward(x, p)
    ref = torch.pow(x, p)
    nan = torch.isnan(ref)
    assert torch.equal(torch.isnan(dark), nan)
    assert torch.equal(dark.view(torch.int32)[~nan], ref.view(torch.int32)[~nan])


def test_synth_host_lut_reproduces_reference_cpu_bits(ops):
    s = load_golden("synth.npz")
    u8 = torch.from_numpy(s["u8"]).cuda()
    for p in (5.0, 7.5, 10.0, 15.0):
        lut, clean_lut = ops.reference_cpu_tables(p, "cuda")
        assert torch.equal(lut.cpu(), torch.from_numpy(s[f"lut_{p}"]))
        clean, dark, q, rec = ops.synth_forward(u8, p, lut=lut, clean_lut=clean_lut, want_u8=True)
        ref = torch.from_numpy(s[f"lut_{p}"])[torch.from_numpy(s["u8"]).long()]
        assert torch.equal(clean.cpu(), torch.from_numpy(s["u8"]).float() / 255)  # CPU true division bits
        assert torch.equal(dark.cpu().view(torch.int32), ref.view(torch.int32))
        assert np.array_equal(q.cpu().numpy(), s[f"q_{p}"])
        assert abs(float(rec) - float(s[f"mse_{p}"])) <= 1e-6 * float(s[f"mse_{p}"])
        # device-computed table vs the reference's CPU table: report the ULP gap, do not gate
        _, dark_dev, _, _ = ops.synth_forward(u8, p)
        ulp = (dark_dev.cpu().view(torch.int32) - ref.view(torch.int32)).abs().max().item()
        print(f"[parity] synth p={p}: CUDA powf vs CPU torch.pow max ULP distance {ulp}")


def test_preprocess_batch_and_offline_darkener(dd):
    s = load_golden("synth.npz")
    u8 = torch.from_numpy(s["u8"])
    batch = dd.preprocess_batch({"img": u8.clone()}, "cuda", dark_param=15.0, dedark_FLAG=False)
    assert set(batch) >= {"img", "clean_img", "recovery_loss_batch"}
    assert torch.equal(batch["clean_img"], u8.cuda().float() / 255)  # the reference's bits when it runs on CUDA
    assert torch.equal(batch["img"], torch.pow(u8.cuda().float() / 255, 15.0))
    assert batch["recovery_loss_batch"].ndim == 0 and not batch["recovery_loss_batch"].requires_grad
    assert abs(float(batch["recovery_loss_batch"]) - float(s["mse_15.0"])) <= 2e-6 * float(s["mse_15.0"])
    b2 = dd.preprocess_batch({"img": u8.clone()}, "cuda", dark_param=15.0)  # defaults = cfg/default.yaml:32-34 (dedark branch)
    assert b2["img"] is b2["clean_img"] and float(b2["recovery_loss_batch"]) == 0.0
    assert torch.equal(b2["img"], batch["img"]) and b2["dedark_A"] is None and b2["IcA"] is None
    with pytest.raises(TypeError):
        dd.preprocess_batch({"img": torch.rand(1, 3, 16, 16)}, "cuda")  # train.py:72 would divide a [0,1] image by 255
    b3 = dd.preprocess_batch({"img": u8.clone()}, "cuda", lowlight_FLAG=False)
    assert torch.equal(b3["img"], u8.cuda().float() / 255) and float(b3["recovery_loss_batch"]) == 0.0
    lut, _ = ops_tables(7.5)
    assert np.array_equal(dd.apply_lowlight(u8.cuda(), 7.5, lut=lut).cpu().numpy(), s["q_7.5"])


# ---- a4 / a5 ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 3, 13, 13), (2, 3, 96, 80), (1, 3, 640, 640), (1, 3, 300, 517), (1, 3, 1280, 1280)])
def test_resize256(ops, shape):
    gen = torch.Generator().manual_seed(3)
    x = torch.rand(shape, generator=gen)
    report(f"resize {shape}", ops.resize256(x.cuda()), O.resize256(x.double()), 2e-6)


@pytest.mark.parametrize("B,scale", [(1, 1.0), (3, 8.0), (16, 1.0)])
def test_predictor_forward_backward(ops, B, scale):
    w = golden_weights(scale)
    gen = torch.Generator().manual_seed(21)
    r = torch.rand(B, 3, 256, 256, generator=gen)
    dfeat = torch.randn(B, 15, generator=gen)
    w64 = O.cast_weights(w, torch.float64, requires_grad=True)
    r64 = r.double().requires_grad_(True)
    feat64, acts64 = O.predictor_forward(r64, w64, return_acts=True)
    feat64.backward(dfeat.double())
    params = cuda_params(w)
    feat, acts = ops.predictor_forward(r.cuda(), params)
    report(f"predictor feat B={B}", feat, feat64, 2e-6)
    grads, dr = ops.predictor_backward(r.cuda(), params, acts, dfeat.cuda(), need_dr=True)
    for k, g in zip(O.STATE_KEYS, grads):
        report(f"predictor grad {k} B={B}", g, w64[k].grad, 2e-5)
    report(f"predictor dr B={B}", dr, r64.grad, 2e-5)


# ---- a6..a12: fused filters, forward ----------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_filters_forward_golden(ops, name):
    c = load_case(name)
    A = None if c["A"] is None else c["A"].cuda()
    IcA = None if c["IcA"] is None else c["IcA"].cuda()
    y = ops.filters_forward(c["x"].cuda(), c["feat"].cuda(), A, IcA)
    report(f"filters fwd {name}", y, c["y"], FWD_TOL)


@pytest.mark.parametrize("name", CASES)
def test_filters_backward_vs_fp64_oracle(ops, name):
    c = load_case(name)
    x64 = c["x"].double().requires_grad_(True)
    f64 = c["feat"].double().requires_grad_(True)
    A64 = None if c["A"] is None else c["A"].double()
    I64 = None if c["IcA"] is None else c["IcA"].double()
    y64 = O.filter_chain(x64, f64, A64, I64, dense_blur=False)
    y64.backward(c["g"].double())
    A = None if c["A"] is None else c["A"].cuda()
    IcA = None if c["IcA"] is None else c["IcA"].cuda()
    dfeat, dx = ops.filters_backward(c["x"].cuda(), c["feat"].cuda(), c["g"].cuda(), A, IcA, need_dx=True)
    print("[parity] dfeat gpu", dfeat[0].cpu().numpy().round(4), "\n[parity] dfeat ref", f64.grad[0].numpy().round(4))
    report(f"filters bwd dfeat {name}", dfeat, f64.grad, GRAD_TOL)
    # dx only: the reference's own fp32 dx (golden) also carries the predictor path, so compare its error on the
    # full module below; here the filter-chain dx is held to 2e-4 (1/tx and x3/x2c amplify fp32 rounding)
    report(f"filters bwd dx {name}", dx, x64.grad, 2e-4)
    assert float(dfeat[:, [1, 5, 6, 7, 8, 9, 10, 11, 12]].abs().max()) == 0.0


# ---- a3 + a14: the drop-in module ----------------------------------------------------------------------------------
def make_module(dd, scale=1.0):
    m = dd.lowlight_recovery(3)
    m.load_state_dict(golden_weights(scale))
    return m.cuda()


@pytest.mark.parametrize("name", CASES)
def test_module_forward_backward_golden(dd, name):
    c = load_case(name)
    m = make_module(dd, c["fc2_scale"]).train()
    x = c["x"].cuda().requires_grad_(True)
    args = () if c["A"] is None else (c["A"].cuda(), c["IcA"].cuda())
    y = m(x, *args)
    assert y.dtype == torch.float32 and y.shape == x.shape and y.device == x.device
    report(f"module fwd {name}", y, c["y"], FWD_TOL)
    y.backward(c["g"].cuda())
    # truth for the gradients: fp64 oracle (the reference's recorded fp32 grads are checked to the same bar)
    _, _, dfeat64, grads64, dx64 = O.recovery_forward_backward(
        c["x"], golden_weights(c["fc2_scale"]), c["g"], c["A"], c["IcA"], dtype=torch.float64, need_dx=True)
    for k, p in m.named_parameters():
        assert p.grad is not None, f"{k} got no gradient (DDP needs all 14)"
        report(f"module grad {k} {name}", p.grad, grads64[k], GRAD_TOL)
        if "grad." + k in c["grads"]:
            report(f"module grad-vs-reference-fp32 {k} {name}", p.grad, c["grads"]["grad." + k], 2e-4)
    report(f"module dx {name}", x.grad, dx64, grad_tol(c["dx"], dx64))


def test_module_cpu_input_is_staged_through_gpu(dd):
    """DetectionModel.__init__ probes with CPU zeros(1,3,256,256) in train mode, grad enabled (tasks.py:290-291)."""
    torch.manual_seed(0)
    m = dd.lowlight_recovery(3)  # parameters on the CPU, like at construction time
    y = m(torch.zeros(1, 3, 256, 256))
    assert y.device.type == "cpu" and y.shape == (1, 3, 256, 256) and y.requires_grad
    y.sum().backward()
    assert all(p.grad is not None and p.grad.device.type == "cpu" for p in m.parameters())
    c = load_case("small_default")
    m.load_state_dict(golden_weights())
    report("module fwd (cpu tensors)", m(c["x"]), c["y"], FWD_TOL)


def test_module_promotes_half_inputs_and_follows_device(dd):
    c = load_case("small_default")
    m = dd.lowlight_recovery(3)
    m.load_state_dict(golden_weights())
    y = m(c["x"].cuda().to(torch.bfloat16))
    assert y.dtype == torch.float32 and next(m.parameters()).is_cuda  # llie.py:28 semantics; fp32 output like the reference
    ref = O.recovery_forward(c["x"].to(torch.bfloat16).float(), golden_weights(), dense_blur=False)
    # a bf16 batch is read in place (bf16 I/O mode: TF32 blur on the tensor cores): the north star's bf16 gate, 2e-2
    report("module fwd (bf16 input, fp32 output)", y, ref, 2e-2)
    # fp16 has no native path: promoted to fp32 as in the reference, fp32 gate
    y16 = m(c["x"].cuda().half())
    assert y16.dtype == torch.float32
    report("module fwd (fp16 input promoted)", y16, O.recovery_forward(c["x"].half().float(), golden_weights(), dense_blur=False), FWD_TOL)
    with torch.no_grad():
        y2 = m.eval()(c["x"].cuda())
    assert not y2.requires_grad
    report("module fwd eval/no_grad", y2, c["y"], FWD_TOL)


def test_bus_640_known_answer(dd):
    """BASELINE config 1: bus.jpg -> 640x640, B=4, weights seed 0."""
    b = load_golden("bus640.npz")
    x = (torch.from_numpy(b["u8"]).float() / 255)[None].repeat(4, 1, 1, 1).cuda()
    m = make_module(dd).eval()
    with torch.no_grad():
        y = m(x)
    s, ma = float(y.double().sum()), float(y.double().abs().mean())
    print(f"[parity] bus640: sum(y)={s:.2f} (ref {float(b['y_sum']):.2f})  mean|y|={ma:.6f} (ref {float(b['y_mean_abs']):.6f})")
    assert abs(s - float(b["y_sum"])) <= 1e-5 * float(b["y_sum"])
    assert abs(ma - float(b["y_mean_abs"])) <= 1e-5
    report("bus640 y_sub", y[0, :, ::8, ::8], torch.from_numpy(b["y_sub"]), FWD_TOL)
    report("bus640 y_rows", y[0, :, 317:323, :], torch.from_numpy(b["y_rows"]), FWD_TOL)
    assert torch.equal(y[0], y[3])  # images are independent: identical inputs -> identical outputs


# ---- full-size, size-independent properties --------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,W", [(16, 640, 640), (2, 1280, 1280), (3, 333, 517)])
def test_full_size_properties(dd, ops, B, H, W):
    gen = torch.Generator().manual_seed(1234)
    x = torch.rand(B, 3, H, W, generator=gen).cuda()
    g = torch.randn(B, 3, H, W, generator=gen).cuda()
    m = make_module(dd, 4.0).train()
    params = [p.detach() for p in m.extractor.ordered_parameters()]
    feat, _ = ops.predictor_forward(ops.resize256(x), params)
    y = ops.filters_forward(x, feat)
    # (1) batch independence: permuting the batch permutes the output.  Bit for bit on the CUDA-core kernels (W % 4 != 0); the
    #     tensor-core blur sums a row's 25 taps in k-steps of 8 staged rows whose alignment follows the CTA's work range, so an
    #     image at another batch position may differ in the last bits (far inside the 1e-5 gate); run to run it is bit-exact.
    perm = torch.arange(B - 1, -1, -1).cuda()
    y_p = ops.filters_forward(x[perm].contiguous(), feat[perm].contiguous())
    if W % 4:
        assert torch.equal(y_p, y[perm])
    else:
        report(f"batch permutation {H}x{W}", y_p, y[perm], 2e-6)
    assert torch.equal(ops.filters_forward(x, feat), y), "forward must be bit-reproducible"
    # (2) forward and backward of image 0 and image B-1 against the fp64 oracle (one image at a time keeps the CPU
    #     cost at seconds); covers multi-strip / multi-segment decompositions at BASELINE sizes
    d1, _ = ops.filters_backward(x, feat, g)
    for b in sorted({0, B - 1}):
        x64 = x[b:b + 1].cpu().double()
        f64 = feat[b:b + 1].cpu().double().requires_grad_(True)
        y64 = O.filter_chain(x64, f64, dense_blur=False)
        report(f"filters fwd full {H}x{W} img {b}", y[b:b + 1], y64, FWD_TOL)
        y64.backward(g[b:b + 1].cpu().double())
        print("[parity] dfeat gpu", d1[b].cpu().numpy()[[0, 2, 3, 4, 13, 14]], "ref", f64.grad[0].numpy()[[0, 2, 3, 4, 13, 14]])
        report(f"filters bwd full {H}x{W} img {b}", d1[b:b + 1], f64.grad, GRAD_TOL)
    # (3) backward: run-to-run determinism (fixed-order reductions) and linearity in the cotangent
    d2, _ = ops.filters_backward(x, feat, g)
    assert torch.equal(d1, d2), "backward must be bit-reproducible"
    d3, _ = ops.filters_backward(x, feat, 2.0 * g)
    report(f"bwd linearity {H}x{W}", d3, 2.0 * d1, 1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.uint8])
def test_pipeline_overlapped_steps_equal_plain_steps(dd, dtype):
    """Software-pipelined mode (synthesis of batch i+1 on a side stream under the predictor backward of batch i): the same bits
    as the plain step for every batch of a sequence, eagerly and replayed from CUDA graphs."""
    gen = torch.Generator().manual_seed(99)
    B, H, W, n = 2, 160, 200, 6
    if dtype == torch.uint8:
        srcs = [torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=gen).cuda() for _ in range(n)]
    else:
        srcs = [torch.rand(B, 3, H, W, generator=gen).cuda() for _ in range(n)]
    gs = [torch.randn(B, 3, H, W, generator=gen).cuda() for _ in range(n)]
    m = make_module(dd, 4.0).train()
    plain = dd.RecoveryPipeline(m, B, H, W, dark_param=5.0, src_dtype=dtype)
    want = [tuple(t.clone() for t in plain.step(srcs[i], gs[i])) for i in range(n)]

    pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=5.0, src_dtype=dtype)
    pipe.prime(srcs[0])
    for i in range(n):
        y, rec, flat = pipe.step_overlapped(srcs[i + 1] if i + 1 < n else None, gs[i])
        torch.cuda.synchronize()
        assert torch.equal(y, want[i][0]) and torch.equal(rec, want[i][1]) and torch.equal(flat, want[i][2]), f"eager step {i}"

    # graphs: step i consumes buffer set i % 2 and synthesises batch i + 1; inputs are read from fixed staging tensors
    stage_src = [torch.empty_like(srcs[0]) for _ in range(2)]
    stage_g = [torch.empty_like(gs[0]) for _ in range(2)]
    for k in range(2):
        pipe.capture_overlapped(k, stage_src[(k + 1) % 2], stage_g[k], slot=k)
    pipe._cur = 0
    stage_src[0].copy_(srcs[0])
    pipe.prime(stage_src[0])
    for i in range(n):
        stage_src[(i + 1) % 2].copy_(srcs[(i + 1) % n])
        stage_g[i % 2].copy_(gs[i])
        y, rec, flat = pipe.replay_overlapped(i % 2)
        torch.cuda.synchronize()
        assert torch.equal(y, want[i][0]) and torch.equal(rec, want[i][1]) and torch.equal(flat, want[i][2]), f"graph step {i}"


def test_pipeline_step_matches_module(dd):
    gen = torch.Generator().manual_seed(77)
    B, H, W = 4, 160, 200
    clean = torch.rand(B, 3, H, W, generator=gen).cuda()
    g = torch.randn(B, 3, H, W, generator=gen).cuda()
    m = make_module(dd, 4.0).train()
    pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=5.0)
    y, rec, flat = pipe.step(clean, g)
    dark = torch.pow(clean, 5.0)
    assert torch.equal(pipe.dark, dark)
    assert abs(float(rec) - float(torch.nn.functional.mse_loss(dark.double(), clean.double()))) <= 1e-6 * float(rec)
    y2 = m(dark)
    assert torch.equal(y, y2)
    y2.backward(g)
    flat2 = torch.cat([p.grad.reshape(-1) for p in m.extractor.ordered_parameters()])
    assert torch.equal(flat, flat2)
    # uint8 source + CUDA graph replay give the same bits as eager launches
    u8 = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=gen).cuda()
    pipe8 = dd.RecoveryPipeline(m, B, H, W, dark_param=7.5, src_dtype=torch.uint8)
    y_e, rec_e, flat_e = (t.clone() for t in pipe8.step(u8, g))
    pipe8.capture("k", u8, g)
    pipe8.y.zero_(); pipe8.flat_grad.zero_()
    y_g, rec_g, flat_g = pipe8.replay("k")
    torch.cuda.synchronize()
    assert torch.equal(y_e, y_g) and torch.equal(flat_e, flat_g) and torch.equal(rec_e, rec_g)


# ---- host staging + determinism ---------------------------------------------------------------------------------------
def test_host_batch_prefetcher_overlapped_copies_are_exact(dd):
    """Every batch comes out of the double-buffered H2D staging bit-identical and in order, also when a slot is
    reused while later work is still enqueued."""
    gen = torch.Generator().manual_seed(9)
    host = [torch.randint(0, 256, (2, 3, 40, 48), dtype=torch.uint8, generator=gen).pin_memory() for _ in range(5)]
    pf = dd.HostBatchPrefetcher("cuda")
    pf.submit(host[0])
    outs = []
    for i in range(5):
        src = pf.get()
        if i + 1 < 5:
            pf.submit(host[i + 1])
        batch = dd.preprocess_batch({"img": src}, "cuda", dark_param=5.0, dedark_FLAG=False)
        outs.append(batch["clean_img"].clone())
    torch.cuda.synchronize()
    for i in range(5):
        assert torch.equal(outs[i], host[i].cuda().float() / 255), f"batch {i}"  # the reference's expression on this GPU
    with pytest.raises(RuntimeError):
        pf.get()  # nothing outstanding


def test_host_batch_prefetcher_never_overwrites_the_batch_in_use(dd):
    """get(); submit(); submit() with two slots must not copy into the buffer the consumer still reads."""
    gen = torch.Generator().manual_seed(8)
    host = [torch.randint(0, 256, (4, 3, 256, 256), dtype=torch.uint8, generator=gen).pin_memory() for _ in range(3)]
    pf = dd.HostBatchPrefetcher("cuda", depth=2)
    pf.submit(host[0])
    cur = pf.get()
    pf.submit(host[1])
    with pytest.raises(RuntimeError):
        pf.submit(host[2])           # slot 0 is held by the consumer, slot 1 is outstanding
    torch.cuda.synchronize()
    assert torch.equal(cur.cpu(), host[0])
    pf.release()                     # explicit end of use: the slot may be refilled now
    pf.submit(host[2])
    assert torch.equal(pf.get().cpu(), host[1]) and torch.equal(pf.get().cpu(), host[2])
    # three slots: two copies may run ahead of the batch in use
    pf3 = dd.HostBatchPrefetcher("cuda", depth=3)
    pf3.submit(host[0]); a = pf3.get(); pf3.submit(host[1]); pf3.submit(host[2])
    torch.cuda.synchronize()
    assert torch.equal(a.cpu(), host[0])


def test_predictor_is_bit_reproducible(ops):
    """Fixed-order reductions everywhere (slices summed in index order, no float atomics): two runs agree bit for bit."""
    w = golden_weights(1.0)
    gen = torch.Generator().manual_seed(33)
    r = torch.rand(5, 3, 256, 256, generator=gen).cuda()
    dfeat = torch.randn(5, 15, generator=gen).cuda()
    params = cuda_params(w)
    runs = []
    for _ in range(2):
        feat, acts = ops.predictor_forward(r, params)
        grads, _ = ops.predictor_backward(r, params, acts, dfeat)
        runs.append((feat.clone(), [g.clone() for g in grads]))
    assert torch.equal(runs[0][0], runs[1][0])
    for a, b in zip(runs[0][1], runs[1][1]):
        assert torch.equal(a, b)


# ---- 8(e): predictor backward fused with the gradient exchange over peer memory ------------------------------------------
def test_fused_gradient_exchange_two_virtual_ranks(ops):
    """dd_predictor_bwd_allreduce with world = 2 emulated on ONE GPU (two exchange buffers, two streams): both ranks end with
    the bitwise identical sum of the two per-rank gradients (slots are added in rank order), for three consecutive steps
    (the slot parity alternates), and it equals the separately computed gradients added on the host side."""
    import ctypes as C
    from dedark_yolo_b200 import _lib
    from dedark_yolo_b200._lib import PeerExchange, PredictorTensors, check, lib
    dev = torch.device("cuda")
    w = golden_weights(1.0)
    params = cuda_params(w)
    nbytes = int(lib.dd_exchange_bytes())
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    pxs = [PeerExchange.from_pointers(k, [b.data_ptr() for b in bufs]) for k in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    B = 3
    ws_bytes = _lib.workspace_bytes(_lib.WS_PREDICTOR_BWD, B)
    gen = torch.Generator().manual_seed(77)
    for step in range(3):
        rs = [torch.rand(B, 3, 256, 256, generator=gen).cuda() for _ in range(2)]
        dfs = [torch.randn(B, 15, generator=gen).cuda() for _ in range(2)]
        fwd = [ops.predictor_forward(rs[k], params) for k in range(2)]
        ref = [ops.predictor_backward(rs[k], params, fwd[k][1], dfs[k])[0] for k in range(2)]
        expect = [a + b for a, b in zip(ref[0], ref[1])]
        outs, keep = [], []
        torch.cuda.synchronize()
        for k in range(2):
            flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
            grads, off = [], 0
            for p in params:
                grads.append(flat[off:off + p.numel()].view(p.shape))
                off += p.numel()
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            wt, gt = PredictorTensors.from_tensors(params), PredictorTensors.from_tensors(grads)
            keep += [ws, wt, gt]
            with torch.cuda.stream(streams[k]):
                check(lib.dd_predictor_bwd_allreduce(C.c_void_p(rs[k].data_ptr()), C.byref(wt), C.c_void_p(fwd[k][1].data_ptr()),
                                                     C.c_void_p(dfs[k].data_ptr()), C.byref(gt), B, C.c_void_p(ws.data_ptr()), ws_bytes,
                                                     C.byref(pxs[k]), streams[k].cuda_stream))
            outs.append(grads)
        torch.cuda.synchronize()
        for i, key in enumerate(O.STATE_KEYS):
            assert torch.equal(outs[0][i], outs[1][i]), f"step {step}: ranks disagree on {key}"
            assert torch.equal(outs[0][i], expect[i]), f"step {step}: {key} is not the sum of the per-rank gradients"


# ---- BASELINE configs[3]: synthesis + recovery inference at 1280x1280, batch 32, bit-exact darkening ------------------------
@pytest.mark.parametrize("p", [7.5, 15.0])
def test_config4_synthesis_and_inference_1280_batch32(dd, p):
    """uint8 batch 32x3x1280x1280, darkening compared bit for bit with the reference's own
    expression evaluated by torch on the same GPU, then module inference with size-independent checks (finite,
    batch-permutation equivariance bit for bit, one image against the fp64 oracle)."""
    B, H, W = 32, 1280, 1280
    gen = torch.Generator(device="cuda").manual_seed(2024)
    u8 = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, generator=gen, device="cuda")
    batch = dd.preprocess_batch({"img": u8}, "cuda", dark_param=p, dedark_FLAG=False)
    clean_ref = u8.float() / 255                                  # train.py:72 on this GPU
    assert torch.equal(batch["clean_img"].view(torch.int32), clean_ref.view(torch.int32))
    dark_ref = torch.pow(clean_ref, p)                            # train.py:103
    assert torch.equal(batch["img"].view(torch.int32), dark_ref.view(torch.int32)), "darkening is not bit-exact"
    rec_ref = torch.nn.functional.mse_loss(dark_ref.double(), clean_ref.double())
    assert abs(float(batch["recovery_loss_batch"]) - float(rec_ref)) <= 1e-6 * float(rec_ref)
    del clean_ref, dark_ref
    m = make_module(dd, 4.0).eval()
    with torch.no_grad():
        y = m(batch["img"])
        assert y.shape == (B, 3, H, W) and y.dtype == torch.float32 and bool(torch.isfinite(y).all())
        perm = torch.randperm(B, generator=torch.Generator().manual_seed(1)).cuda()
        y_p = m(batch["img"][perm].contiguous())
        assert torch.equal(y_p, y[perm])
    if p == 7.5:  # one image against the fp64 oracle (the CPU side of this costs a few seconds)
        b = 17
        weights = O.cast_weights({k: v.detach().cpu() for k, v in m.state_dict().items()}, torch.float64)
        y64 = O.recovery_forward(batch["img"][b:b + 1].cpu().double(), weights, dense_blur=False)
        report(f"config4 fwd 1280x1280 img {b}", y[b:b + 1], y64, FWD_TOL)


# ---- a1 + a2 + a4 in one pass ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,dtype", [((2, 3, 640, 640), torch.float32), ((3, 3, 640, 640), torch.uint8), ((1, 3, 1280, 1280), torch.uint8),
                                         ((2, 3, 96, 80), torch.float32), ((2, 3, 300, 516), torch.uint8), ((1, 3, 2048, 1024), torch.float32),
                                         ((2, 3, 13, 16), torch.uint8)])
def test_synth_resize_fused_equals_separate_passes(ops, shape, dtype):
    """dd_synth_resize_fwd == dd_synth_fwd followed by dd_resize256: clean, dark and r bit for bit, rec within 1e-6 of the fp64
    value (the partial sums are grouped differently); covers down- and up-sampling, scales below and above 2, non-square images."""
    gen = torch.Generator().manual_seed(11)
    src = (torch.randint(0, 256, shape, dtype=torch.uint8, generator=gen) if dtype == torch.uint8 else torch.rand(shape, generator=gen)).cuda()
    assert ops.synth_resize_supported(shape[2], shape[3])
    clean_a, dark_a, _, rec_a = ops.synth_forward(src, 7.5)
    r_a = ops.resize256(dark_a)
    clean_b, dark_b, r_b, rec_b = ops.synth_resize_forward(src, 7.5)
    assert torch.equal(dark_a, dark_b) and torch.equal(r_a, r_b)
    if dtype == torch.uint8:
        assert torch.equal(clean_a, clean_b)
    clean64 = (clean_b if dtype == torch.uint8 else src).double()
    rec64 = float(torch.nn.functional.mse_loss(dark_b.double(), clean64))
    assert abs(float(rec_b) - rec64) <= 1e-6 * rec64 and abs(float(rec_a) - rec64) <= 1e-6 * rec64
    assert not ops.synth_resize_supported(640, 641)  # W % 4 != 0 stays on the two-pass path


# ---- reflect borders of the backward by mirror extension: shapes where both borders share a block / strip -----
@pytest.mark.parametrize("B,H,W", [(2, 14, 14), (1, 14, 40), (2, 20, 24), (1, 31, 33), (2, 45, 140), (1, 70, 132), (1, 33, 300), (1, 160, 156),
                                   (1, 13, 40), (1, 40, 13)])
def test_backward_border_shapes_vs_fp64_oracle(dd, ops, B, H, W):
    """Every combination the mirror extension distinguishes: top and bottom mirror rows in one block (H < 32), left and right
    mirror columns in one strip (W < 128), a strip whose right halo lies outside the image (W = 132, 140), the TMA path
    (W >= 156), unaligned widths (register staging), and 13-pixel images, which keep the fold-back form.  dfeat and dx against
    the fp64 oracle (autograd through the reflect-padded dense-equivalent blur)."""
    gen = torch.Generator().manual_seed(100 * H + W)
    x = torch.rand(B, 3, H, W, generator=gen)
    g = torch.randn(B, 3, H, W, generator=gen)
    feat = torch.randn(B, 15, generator=gen) * 0.8
    x64 = x.double().requires_grad_(True)
    f64 = feat.double().requires_grad_(True)
    y64 = O.filter_chain(x64, f64, dense_blur=False)
    y64.backward(g.double())
    y = ops.filters_forward(x.cuda(), feat.cuda())
    dfeat, dx = ops.filters_backward(x.cuda(), feat.cuda(), g.cuda(), need_dx=True)
    report(f"border fwd {H}x{W}", y, y64, FWD_TOL)
    report(f"border dfeat {H}x{W}", dfeat, f64.grad, GRAD_TOL)
    report(f"border dx {H}x{W}", dx, x64.grad, 2 * GRAD_TOL)
    d2, dx2 = ops.filters_backward(x.cuda(), feat.cuda(), g.cuda(), need_dx=True)
    assert torch.equal(dfeat, d2) and torch.equal(dx, dx2), "backward must be bit-reproducible"
