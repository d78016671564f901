"""Multi-GPU plumbing: one process per GPU, torch.distributed over NCCL.

Training is data parallel exactly like the reference (DistributedSampler + DDP,
movenet/dataset.py:78-87, movenet/trainer.py:223-238): clips are sharded over ranks, weights are
replicated, gradients are averaged -- here with ONE all-reduce of the flat gradient buffer per
backward (``WaveNet.enable_data_parallel``).  Generation shards independent clips with no
communication at all.
"""
import os

import torch


def shard_range(n_items: int, rank: int, world: int):
    """contiguous [lo, hi) slice of n_items owned by rank (sizes differ by at most one)"""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init_from_env(backend: str = "nccl"):
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*); returns (rank, local_rank, world)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local, world
