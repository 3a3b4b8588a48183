"""The reference-shaped API at speed: preprocess_batch hands the module the 256x256 resize it computed in
the synthesis pass, and ``module.use_cuda_graphs`` replays the launch sequences of forward / backward as CUDA graphs.  Both
must be invisible: bit-identical outputs and gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _step(m, u8, g, dev, p=5.0):
    import dedark_yolo_b200 as dd
    for q in m.parameters():
        q.grad = None
    batch = dd.preprocess_batch({"img": u8.clone()}, dev, dark_param=p, dedark_FLAG=False)
    y = m(batch["img"])
    y.backward(g)
    return batch, y.detach().clone(), torch.cat([q.grad.reshape(-1) for q in m.parameters()]).clone()


def test_resize_stash_is_used_and_invisible():
    import dedark_yolo_b200 as dd
    from dedark_yolo_b200 import ops

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = dd.lowlight_recovery(3).to(dev).train()
    gen = torch.Generator().manual_seed(3)
    u8 = torch.randint(0, 256, (2, 3, 96, 80), dtype=torch.uint8, generator=gen)
    g = torch.randn(2, 3, 96, 80, generator=gen).to(dev)
    batch, y, grads = _step(m, u8, g, dev)
    r, ver = batch["img"]._dd_resize256
    assert ver == batch["img"]._version and torch.equal(r, ops.resize256(batch["img"]))
    n0 = dd.launch_count()
    y2 = m(batch["img"])                      # stash valid: no resize launch
    n1 = dd.launch_count()
    y3 = m(batch["img"].clone())              # no stash on the clone: resize launch
    n2 = dd.launch_count()
    assert (n2 - n1) - (n1 - n0) == 1
    assert torch.equal(y2, y) and torch.equal(y3, y)
    # an in-place write invalidates the stash
    img = batch["img"]
    img.mul_(0.5)
    y4 = m(img)
    assert torch.equal(y4, m(img.clone()))
    # gradients with and without the stash
    for q in m.parameters():
        q.grad = None
    x = batch["clean_img"].clone()
    m(x).backward(g)
    gref = torch.cat([q.grad.reshape(-1) for q in m.parameters()])
    for q in m.parameters():
        q.grad = None
    x2 = batch["clean_img"].clone()
    x2._dd_resize256 = (ops.resize256(x2), x2._version)
    m(x2).backward(g)
    assert torch.equal(gref, torch.cat([q.grad.reshape(-1) for q in m.parameters()]))


def test_cuda_graph_replay_is_bit_identical():
    import dedark_yolo_b200 as dd

    dev = torch.device("cuda:0")
    gen = torch.Generator().manual_seed(5)
    u8s = [torch.randint(0, 256, (4, 3, 160, 192), dtype=torch.uint8, generator=gen) for _ in range(3)]
    gs = [torch.randn(4, 3, 160, 192, generator=gen).to(dev) for _ in range(3)]
    outs = {}
    for use in (False, True):
        torch.manual_seed(0)
        m = dd.lowlight_recovery(3).to(dev).train()
        m.use_cuda_graphs = use
        res = []
        for i in range(12):                    # the allocator cycles through a few addresses: graphs get captured and replayed
            _, y, gr = _step(m, u8s[i % 3], gs[i % 3], dev)
            res.append((y, gr))
            with torch.no_grad():              # an optimizer-like in-place update: same addresses, new values
                for q in m.parameters():
                    q.add_(1e-3 * q.grad)
        torch.cuda.synchronize()
        outs[use] = res
        if use:
            from dedark_yolo_b200 import ops
            plans = ops._PLANS[m]
            assert any(len(pl.graphs) > 0 for pl in plans.values()), "no graph was ever captured"
    for (y0, g0), (y1, g1) in zip(outs[False], outs[True]):
        assert torch.equal(y0, y1) and torch.equal(g0, g1)
