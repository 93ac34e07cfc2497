"""Multi-GPU sharding of independent restorations (images x hyper-parameter grid points).

The path shards with no exchange inside the loop: each rank (one process per GPU,
torchrun) takes a contiguous block of the flattened item list, runs it resident on its GPU, and the
per-item traces/finals are gathered ONCE at the end with a single all-gather (NCCL over NVLink on
GPUs; gloo in the CPU tests of this host-side logic).
"""
from __future__ import annotations

import os

import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous block [lo, hi) of items for `rank`; blocks differ by at most one item."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def dist_info():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_distributed(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = dist_info()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend)
    return rank, local_rank, world


def gather_rows(local_rows: np.ndarray, n_items: int, device=None) -> np.ndarray:
    """All-gather per-item rows.  local_rows: [items_on_this_rank, width] float32/float64 for the block
    shard_range() gave this rank.  Returns [n_items, width] on every rank.  The last blocks are padded
    to the largest block so one all_gather_into_tensor suffices."""
    import torch
    import torch.distributed as dist
    rank, _, world = dist_info()
    if world == 1 or not dist.is_initialized():
        return np.asarray(local_rows)
    width = local_rows.shape[1]
    per = (n_items + world - 1) // world
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    send = torch.zeros((per, width), dtype=torch.float64, device=dev)
    if local_rows.shape[0]:
        send[: local_rows.shape[0]] = torch.from_numpy(np.ascontiguousarray(local_rows, dtype=np.float64)).to(dev)
    recv = torch.empty((world * per, width), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(recv, send)
    recv = recv.cpu().numpy().reshape(world, per, width)
    out = np.empty((n_items, width), dtype=np.float64)
    for r in range(world):
        lo, hi = shard_range(n_items, r, world)
        out[lo:hi] = recv[r, : hi - lo]
    return out
