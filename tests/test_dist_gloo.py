"""world_size-2 gloo test of the multi-process logic (CPU): image sharding + one all-reduce(sum) of the flat
predictor gradient reproduces the full-batch gradient (the reference's DDP semantics: average x world_size)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden_weights, load_case, rel_to_max
from oracle import lowlight_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from dedark_yolo_b200.dist import allreduce_flat_, init_from_env, max_over_ranks, shard_range
    r, _, w = init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    torch.set_num_threads(2)
    c = load_case("small_default")
    lo, hi = shard_range(c["x"].shape[0], rank, world)
    weights = golden_weights()
    # per-rank gradient of the per-rank loss (the oracle stands in for the GPU kernels on this CPU-only box)
    _, _, _, grads, _ = O.recovery_forward_backward(c["x"][lo:hi], weights, c["g"][lo:hi], dtype=torch.float64)
    flat = torch.cat([grads[k].reshape(-1) for k in O.STATE_KEYS]).float()
    allreduce_flat_(flat)
    assert max_over_ranks(float(rank)) == world - 1
    if rank == 0:
        torch.save(flat, os.path.join(out_dir, "flat.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_gradient_sum_equals_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    flat = torch.load(os.path.join(tmp_path, "flat.pt"))
    c = load_case("small_default")
    _, _, _, grads, _ = O.recovery_forward_backward(c["x"], golden_weights(), c["g"], dtype=torch.float64)
    full = torch.cat([grads[k].reshape(-1) for k in O.STATE_KEYS]).float()
    assert flat.numel() == 164943
    assert rel_to_max(flat, full) <= 1e-6
