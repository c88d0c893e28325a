"""Tile sharding across the GPUs of one box: tiles are independent, so rank r of W
takes a contiguous slice of the batch and there is no collective on the data path."""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n units for `rank`; the first n % world ranks get one extra."""
    if world < 1 or not 0 <= rank < world or n < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
