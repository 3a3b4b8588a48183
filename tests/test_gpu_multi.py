"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): image sharding + the bucketed NCCL all-reduce(sum) of the flat
predictor gradient reproduces the full-batch gradient, bitwise identical on every rank -- the reference's DDP semantics
(engine/trainer.py:223,334-335: DDP averages, the trainer pre-multiplies the loss by world_size).  Exercises the captured
two-graph step of ``RecoveryPipeline`` (fc bucket reduced beside the convolution backward) and the plain eager step."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import dedark_yolo_b200 as dd
    from dedark_yolo_b200.dist import init_from_env, shard_range
    init_from_env("nccl")
    dev = torch.device("cuda", rank)
    torch.manual_seed(0)
    m = dd.lowlight_recovery(3).to(dev).train()
    Bt, H, W = 4 * world, 96, 128
    gen = torch.Generator().manual_seed(99)
    clean = torch.rand(Bt, 3, H, W, generator=gen)
    g = torch.randn(Bt, 3, H, W, generator=gen)
    lo, hi = shard_range(Bt, rank, world)
    mine, gm = clean[lo:hi].to(dev), g[lo:hi].to(dev)
    B = hi - lo
    # (1) eager overlapped step with the bucketed all-reduce
    pipe = dd.RecoveryPipeline(m, B, H, W, dark_param=5.0, allreduce=True)
    pipe.prime(mine)
    _, _, flat = pipe.step_overlapped(mine, gm)
    flat_eager = flat.clone()
    # (2) the same step replayed from the two captured graphs
    pipe.capture_overlapped("k", mine, gm, slot=pipe._cur)
    pipe.flat_grad.zero_()
    _, _, flat = pipe.replay_overlapped("k")
    torch.cuda.synchronize(dev)
    flat_graph = flat.clone()
    gathered = [torch.empty_like(flat_graph) for _ in range(world)]
    dist.all_gather(gathered, flat_graph)
    if rank == 0:
        # full batch on one GPU, no collective
        full = dd.RecoveryPipeline(m, Bt, H, W, dark_param=5.0)
        _, _, flat_full = full.step(clean.to(dev), g.to(dev))
        torch.save({"eager": flat_eager.cpu(), "graph": flat_graph.cpu(), "ranks": [t.cpu() for t in gathered], "full": flat_full.cpu()},
                   os.path.join(out_dir, "out.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_bucketed_allreduce_equals_full_batch(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    out = torch.load(os.path.join(tmp_path, "out.pt"))
    assert torch.equal(out["eager"], out["graph"]), "captured two-graph step differs from the eager step"
    for r in range(1, world):
        assert torch.equal(out["ranks"][0], out["ranks"][r]), "ranks disagree on the reduced gradient"
    err = float((out["graph"].double() - out["full"].double()).abs().max() / out["full"].double().abs().max())
    print(f"[multi] sum over {world} ranks vs full batch: rel-to-max {err:.2e}")
    assert err <= 1e-5
