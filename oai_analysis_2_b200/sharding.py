"""By-volume sharding of knees over the GPUs of one box: knees are independent, so every rank processes
its own subset and the only communication is a host-side gather of small per-knee records.  No data-path collective.

The reference distributes the same way with Dask delayed task chains, one per knee
(notebooks/DaskComputationCoiled.ipynb cell 3; oai_analysis/dask_processing.py)."""
import os

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_process_group(backend=None):
    """One process per GPU; rendezvous from MASTER_ADDR/MASTER_PORT/RANK/WORLD_SIZE (torchrun)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_indices(n_items, rank, world):
    """Static round-robin deal: item i goes to rank i % world."""
    return list(range(rank, n_items, world))


def gather_records(records, rank, world, dst=0):
    """Host-side gather of picklable per-knee records to `dst` (returns the merged list there, None elsewhere)."""
    if world == 1:
        return list(records)
    out = [None] * world if rank == dst else None
    dist.gather_object(list(records), out, dst=dst)
    if rank != dst:
        return None
    merged = [r for part in out for r in part]
    merged.sort(key=lambda r: r.get("index", 0) if isinstance(r, dict) else 0)
    return merged


def max_over_ranks(value, world, device=None):
    """Max-reduce a python float over ranks (timing is reported as the slowest rank)."""
    if world == 1:
        return float(value)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        dist.barrier()
