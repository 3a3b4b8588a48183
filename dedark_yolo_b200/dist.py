"""Data-parallel plumbing for the hot path (one process per GPU, torch.distributed).

The path shards over images with no data-path collective (every statistic is per image);
the only exchange is one all-reduce(sum) of the flat predictor gradient (164 943 fp32 = 659 772 B) per step.
That mirrors the reference: DDP averages gradients and the trainer pre-multiplies the loss by world_size
(engine/trainer.py:223,334-335), i.e. the net effect is the SUM over ranks of per-rank gradients.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_images: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, as-even-as-possible image range [lo, hi) owned by ``rank`` (first ranks take the remainder)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(n_images, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: str = "nccl"):
    """Initialise the default process group from torchrun's environment; returns (rank, local_rank, world_size)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            kw["device_id"] = torch.device("cuda", local_rank)
            if os.environ.get("DEDARK_NCCL_HIGH_PRIORITY", "1") == "1":
                # the collectives of the bucketed step are small and latency-bound and start beside kernels that fill the GPU
                # (the next batch's synthesis): their CTAs must not queue behind those
                try:
                    opts = dist.ProcessGroupNCCL.Options()
                    opts.is_high_priority_stream = True
                    kw["pg_options"] = opts
                except Exception:  # pragma: no cover - older torch
                    pass
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, local_rank, world


class GradExchange:
    """Peer-memory exchange buffers for ``dd_predictor_bwd_allreduce``.

    Every rank allocates one exchange buffer (``dd_exchange_bytes()``, zero-filled) that all ranks of the node map into
    their address space: through ``torch.distributed._symmetric_memory`` when it is available, otherwise through CUDA IPC
    handles of a plain allocation.  ``self.px`` is the descriptor the C-ABI takes; the predictor backward then stores its
    gradients straight into the peers' buffers and sums the world's contributions itself -- no collective library call.
    Raises -- on every rank together, the ranks agree on each step of the set-up -- if the buffers cannot be shared (the
    caller then keeps the NCCL all-reduce and says so).

    Scope: the receiving kernel polls for its peers' words with a bounded spin and TRAPS (the launch fails, the context is
    lost) if a peer has not arrived after about a minute.  That is the right behaviour for lock-step data-parallel steps
    (bench.py, RecoveryPipeline) and the wrong one for a job in which one rank may legitimately stall for minutes between
    the backward passes of a step (rank-0 validation, checkpointing inside a step): such jobs keep the NCCL path."""

    def __init__(self, device, group=None):
        from . import _lib
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("GradExchange needs an initialised process group")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError(f"GradExchange: world size {self.world} > {_lib.MAX_PEERS}")
        dev = torch.device(device)
        nbytes = int(_lib.lib.dd_exchange_bytes())
        self.how, self._keep = None, []

        def all_ok(flag: bool) -> bool:
            """Every rank takes the same branch: a path is used only if it worked on ALL ranks (a rank that entered the
            fallback's collective alone would hang the others)."""
            t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
            return bool(t.item())

        ptrs, errs = None, []
        # (1) symmetric memory: local allocation first, agreement, then the collective rendezvous, agreement again
        buf = None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
            buf.zero_()
        except Exception as e:  # pragma: no cover - depends on driver / fabric support
            errs.append(f"symmetric memory allocation: {e!r}")
            buf = None
        if all_ok(buf is not None):
            try:
                hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                self._keep += [buf, hdl]
            except Exception as e:  # pragma: no cover
                errs.append(f"symmetric memory rendezvous: {e!r}")
                ptrs = None
            if all_ok(ptrs is not None):
                self.how = "symmetric_memory"
            else:
                ptrs = None
        # (2) CUDA IPC handles of a plain allocation, exchanged with all_gather_object (all ranks or none)
        if self.how is None:
            self._keep = []
            info = None
            try:
                buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
                info = buf.untyped_storage()._share_cuda_()
            except Exception as e:  # pragma: no cover
                errs.append(f"CUDA IPC export: {e!r}")
            if all_ok(info is not None):
                infos = [None] * self.world
                dist.all_gather_object(infos, info, group=group)
                try:
                    ptrs = []
                    for r, inf in enumerate(infos):
                        if r == self.rank:
                            ptrs.append(buf.data_ptr())
                        else:
                            st = torch.UntypedStorage._new_shared_cuda(*inf)
                            self._keep.append(st)
                            ptrs.append(st.data_ptr())
                    self._keep.append(buf)
                except Exception as e:  # pragma: no cover
                    errs.append(f"CUDA IPC import: {e!r}")
                    ptrs = None
                if all_ok(ptrs is not None):
                    self.how = "cuda_ipc"
        if self.how is None:  # every rank raises together: the caller keeps the NCCL all-reduce
            raise RuntimeError("GradExchange: the exchange buffers cannot be shared on every rank (" + "; ".join(errs or ["a peer failed"]) + ")")
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)  # every buffer is zero-filled and mapped before anyone stores into it
        self.px = _lib.PeerExchange.from_pointers(self.rank, ptrs)


def allreduce_flat_(flat_grad: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce(sum) of the flat gradient buffer; a no-op for a single process."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def max_over_ranks(value: float, device=None) -> float:
    """Max of a host scalar over ranks (timing is reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
