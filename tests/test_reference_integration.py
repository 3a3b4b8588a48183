"""The drop-in inside the REAL, unmodified reference (BASELINE configs[2] at test size).

``baseline/_ref`` holds a byte-for-byte copy of the reference's ``ultralytics`` package (baseline/install_reference.py;
git-ignored, shipped to the GPU box).  These tests build the reference ``DetectionModel('yolov8l.yaml', nc=3)`` twice --
stock, and with layer 0 swapped by ``integrate.install()`` -- load the same weights and compare a train-mode
``model(batch)`` + backward: ``parse_model`` by-name lookup and identity (nn/tasks.py:844,888), the CPU stride probe
(tasks.py:290-291), ``_predict_once`` (tasks.py:107-110), ``RcoveryDetectionLoss`` (utils/loss.py:388-416), and
deepcopy / pickle / ``.half().float()`` of the whole model (trainer.py:408-433, torch_utils.py:353).
"""
import copy
import io
import pickle

import pytest
import torch

from baseline import reference_runtime as R

pytestmark = pytest.mark.gpu

# fp32 everywhere (no TF32): the reference's cuDNN convs default to TF32 on GPU, which is ~1e-3
LOSS_RTOL = 2e-4
# Predictor gradients travel through 26 fp32 detector layers (BatchNorm on a 2-image batch) before they reach the module, so
# the yardstick is the stock model's OWN response to an input perturbation at the module's output tolerance (1e-6 relative):
# the drop-in must stay within GRAD_YARD_FACTOR x that, with GRAD_FLOOR as the floor (measured on B200: drop-in 2.5e-3).
GRAD_FLOOR = 2e-3
GRAD_YARD_FACTOR = 5.0


@pytest.fixture(scope="module")
def ref():
    if not R.available():
        pytest.skip("baseline/_ref missing: run baseline/install_reference.py in the build container")
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    ns = R.load_ultralytics()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return ns


def make_batch(dev, B=2, S=320, seed=11):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, S, S, generator=g)
    return {
        "img": img.to(dev),
        "cls": torch.tensor([[float(i % 3)] for i in range(B)], device=dev),
        "bboxes": torch.tensor([[0.5, 0.5, 0.3, 0.3]] * B, device=dev),
        "batch_idx": torch.arange(B, dtype=torch.float32, device=dev),
        "recovery_loss_batch": torch.tensor(0.2481, device=dev),
    }


def build(ns, dev, state=None):
    torch.manual_seed(0)
    model = ns.DetectionModel(ns.yaml_l, nc=3, verbose=False)
    if state is not None:
        model.load_state_dict(state)
    model.args = ns.get_cfg(ns.DEFAULT_CFG)
    return model.to(dev).train()


def step(model, batch):
    model.zero_grad(set_to_none=True)
    loss, items = model(dict(batch))
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters() if k.startswith("model.0.")}
    return loss.detach(), items.detach(), grads


def test_dropin_inside_reference_detection_model(ref):
    import dedark_yolo_b200 as dd
    import dedark_yolo_b200.integrate as it
    dev = torch.device("cuda:0")
    batch = make_batch(dev)

    stock = build(ref, dev)
    assert type(stock.model[0]).__module__ == "ultralytics.nn.modules.llie" and not isinstance(stock.model[0], dd.lowlight_recovery)
    state = {k: v.detach().clone() for k, v in stock.state_dict().items()}
    loss_ref, items_ref, grads_ref = step(stock, batch)
    noisy = dict(batch)
    noisy["img"] = batch["img"] * (1.0 + 1e-6 * torch.randn(batch["img"].shape, generator=torch.Generator().manual_seed(5)).to(dev))
    _, _, grads_noisy = step(stock, noisy)
    yard = max(float((grads_noisy[k] - g).abs().max() / g.abs().max().clamp_min(1e-20)) for k, g in grads_ref.items())
    stock.eval()
    with torch.no_grad():
        y_eval_ref = stock.model[0](batch["img"])
    del stock

    original = it.install()
    try:
        assert original is ref.lowlight_recovery or original is not None
        swapped = build(ref, dev, state)      # parse_model finds the class by name; __init__ probes with CPU zeros
        m0 = swapped.model[0]
        assert isinstance(m0, dd.lowlight_recovery) and type(m0).__module__ == "ultralytics.nn.modules.llie"
        assert (m0.i, m0.f) == (0, -1) and len(m0.filters) == 5
        assert len(grads_ref) == 14 and sorted(k for k in swapped.state_dict() if k.startswith("model.0.")) == sorted(
            k for k in state if k.startswith("model.0."))
        n0 = dd.launch_count()
        loss, items, grads = step(swapped, batch)
        assert dd.launch_count() > n0, "the swapped model did not launch any kernel of libdedark_b200.so"
        assert abs(float(loss) - float(loss_ref)) <= LOSS_RTOL * abs(float(loss_ref)), (float(loss), float(loss_ref))
        assert torch.allclose(items.cpu(), items_ref.cpu(), rtol=LOSS_RTOL, atol=1e-5), (items, items_ref)
        worst = 0.0
        for k, gr in grads_ref.items():
            scale = float(gr.abs().max().clamp_min(1e-20))
            worst = max(worst, float((grads[k] - gr).abs().max()) / scale)
        print(f"integrated step: loss {float(loss):.6f} vs {float(loss_ref):.6f}, worst predictor-grad rel-to-max {worst:.2e} "
              f"(the stock model's own response to a 1e-6 input perturbation: {yard:.2e})")
        assert worst <= max(GRAD_FLOOR, GRAD_YARD_FACTOR * yard), (worst, yard)
        # every one of the 14 tensors received a gradient (DDP with find_unused_parameters=False, trainer.py:223)
        assert all(g is not None and torch.isfinite(g).all() for g in grads.values())

        # eval path: _predict_once hands (x, dedark_A, IcA) to the module (tasks.py:107-108)
        swapped.eval()
        with torch.no_grad():
            y_eval = swapped.model[0](batch["img"], None, None)
            det = swapped(batch["img"])
        assert float((y_eval - y_eval_ref).abs().max() / y_eval_ref.abs().max()) <= 1e-5
        assert isinstance(det, (tuple, list, torch.Tensor))

        # EMA / checkpoint mechanics on the whole model
        ema = copy.deepcopy(swapped)
        assert isinstance(ema.model[0], dd.lowlight_recovery)
        half = copy.deepcopy(swapped).half()
        assert next(half.model[0].parameters()).dtype == torch.float16
        buf = io.BytesIO()
        pickle.dump({"model": half}, buf)
        back = pickle.loads(buf.getvalue())["model"].float().to(dev).eval()
        assert isinstance(back.model[0], dd.lowlight_recovery)
        with torch.no_grad():
            y_back = back.model[0](batch["img"])
        # weights went through fp16: compare against the same module evaluated with the rounded weights
        ref_half = copy.deepcopy(swapped.model[0]).half().float()
        with torch.no_grad():
            y_half = ref_half(batch["img"])
        assert float((y_back - y_half).abs().max() / y_half.abs().max()) <= 1e-5
    finally:
        it.uninstall()
    assert ref.tasks.lowlight_recovery is original


def test_checkpoint_written_with_dropin_loads_as_reference_module(ref):
    """trainer.py:408-433 pickles whole objects: while installed the drop-in writes itself as a genuine reference module."""
    import dedark_yolo_b200 as dd
    import dedark_yolo_b200.integrate as it
    dev = torch.device("cuda:0")
    original = it.install()
    try:
        torch.manual_seed(0)
        m = dd.lowlight_recovery(3).to(dev)
        m.i, m.f, m.type, m.np = 0, -1, "ultralytics.nn.modules.llie.lowlight_recovery", 164943
        blob = pickle.dumps(m)
        x = torch.rand(2, 3, 96, 80, device=dev)
        with torch.no_grad():
            y = m(x)
    finally:
        it.uninstall()
    back = pickle.loads(blob)              # not installed any more: the name resolves to the reference class
    assert type(back) is original and (back.i, back.f, back.np) == (0, -1, 164943)
    assert type(back.extractor).__module__ == "ultralytics.nn.modules.common"
    with torch.no_grad():
        y_ref = back(x)
    assert float((y - y_ref).abs().max() / y_ref.abs().max()) <= 1e-5
