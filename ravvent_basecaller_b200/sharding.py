"""Read sharding across the GPUs of one box: snippets / reads are independent, so each
rank takes a contiguous range, runs the whole hot path on its own GPU, and the results are gathered
host-side in input order.  There is no data-path collective."""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous [lo, hi) of rank; sizes differ by at most one, earlier ranks take the remainder."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_in_order(local: np.ndarray, n_items: int, group=None):
    """All ranks contribute their shard (first axis); every rank gets the concatenation in input order.
    Host-side gather over torch.distributed (gloo or nccl object path); shards may be ragged."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, np.ascontiguousarray(local), group=group)
    out = np.concatenate(parts, axis=0)
    if out.shape[0] != n_items:
        raise RuntimeError(f"gathered {out.shape[0]} items, expected {n_items}")
    return out


def run_sharded(fn, inputs, rank: int, world: int, group=None):
    """fn(shard_inputs) -> ndarray per item.  inputs: array or tuple of arrays sharing the first axis."""
    first = inputs[0] if isinstance(inputs, (tuple, list)) else inputs
    n = int(first.shape[0])
    lo, hi = shard_range(n, rank, world)
    shard = tuple(x[lo:hi] for x in inputs) if isinstance(inputs, (tuple, list)) else inputs[lo:hi]
    return gather_in_order(np.asarray(fn(shard)), n, group)
