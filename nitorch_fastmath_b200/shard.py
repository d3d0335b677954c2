"""Multi-GPU partitioning of the batch (voxel) axis.

Every matrix is independent, so the path shards with no exchange step: rank
``r`` of ``w`` owns one contiguous slab of the flattened batch and nothing is
communicated on the data path (NCCL is not used).
Slab boundaries are multiples of ``align`` matrices so that every slab keeps
the 16-byte alignment the TMA fast path needs for any record length.
"""
from __future__ import annotations

from typing import Tuple


def shard_bounds(batch: int, world: int, rank: int, align: int = 1024) -> Tuple[int, int]:
    """[begin, end) of rank's slab; slabs differ by at most ``align`` matrices."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank {rank} for world size {world}")
    if batch < 0:
        raise ValueError("negative batch")
    units = -(-batch // align)                      # ceil
    base, extra = divmod(units, world)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    begin = min(first * align, batch)
    end = min((first + count) * align, batch)
    return begin, end


def shard_sizes(batch: int, world: int, align: int = 1024):
    return [e - b for b, e in (shard_bounds(batch, world, r, align) for r in range(world))]
