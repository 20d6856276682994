"""Data-parallel plumbing: one process per GPU, segments sharded in contiguous ranges (fingerprint
generation has no collective at all -- segments are independent in eval mode); the only exchanges
on the path are the embedding all-gather inside ntxent_loss_distributed and the gradient
all-reduce of the train step."""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n items owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
