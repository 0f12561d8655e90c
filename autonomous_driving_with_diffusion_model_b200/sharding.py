"""Batch sharding of independent planning requests over the GPUs of one box: contiguous split, packed
weights replicated per GPU, no collective on the sampling path."""
from __future__ import annotations

from typing import List, Tuple


def shard_bounds(batch: int, world_size: int) -> List[Tuple[int, int]]:
    """[lo, hi) of each rank; the first ``batch % world_size`` ranks take one extra trajectory."""
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    q, r = divmod(batch, world_size)
    out, lo = [], 0
    for rank in range(world_size):
        hi = lo + q + (1 if rank < r else 0)
        out.append((lo, hi))
        lo = hi
    return out


def shard(t, rank: int, world_size: int, dim: int = 0):
    """Slice of tensor ``t`` along ``dim`` owned by ``rank`` (None passes through)."""
    if t is None:
        return None
    lo, hi = shard_bounds(t.shape[dim], world_size)[rank]
    return t.narrow(dim, lo, hi - lo)
