"""Batch sharding across the GPUs of one box: witnesses are independent,
so rank r of W takes a contiguous slice of the batch and nothing is exchanged on the data path.
torch.distributed is only plumbing (barrier, gathering the small per-lane verdicts)."""
from __future__ import annotations

import numpy as np


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous, balanced slice [lo, hi) of `total` lanes for `rank`."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_lanes(local: np.ndarray, total: int, dist=None):
    """Concatenate per-rank result arrays (status / first_bad / public) in lane order on every rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == total
    return out
