"""Read-chunk sharding for N GPUs: the unit is a read with all its candidate
chains; every rank holds the full reference; no collective on the data path -- only the final
gather of per-read records, in read order."""
from __future__ import annotations

import numpy as np


def shard_bounds(weights: np.ndarray, world: int) -> np.ndarray:
    """Contiguous read ranges balanced by `weights` (read lengths).  Returns world+1 boundaries."""
    w = np.asarray(weights, dtype=np.float64)
    c = np.concatenate([[0.0], np.cumsum(w)])
    targets = c[-1] * np.arange(1, world) / world
    cuts = np.searchsorted(c, targets, side="left")
    b = np.concatenate([[0], cuts, [len(w)]]).astype(np.int64)
    return np.maximum.accumulate(b)


def shard_chains(chain_read_id: np.ndarray, bounds: np.ndarray, rank: int) -> np.ndarray:
    """Indices of the chains whose read falls in this rank's read range."""
    lo, hi = bounds[rank], bounds[rank + 1]
    return np.flatnonzero((chain_read_id >= lo) & (chain_read_id < hi))


def merge_in_read_order(parts):
    """parts: list over ranks of lists of (read_id, payload).  Returns payloads ordered by read, keeping
    each rank's internal order (results are returned to the host in read order)."""
    out = [x for p in parts for x in p]
    out.sort(key=lambda x: x[0])
    return out
