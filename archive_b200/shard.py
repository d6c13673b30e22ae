"""Multi-GPU partitioning of independent compressed units: contiguous unit ranges per rank,
balanced by compressed bytes, each rank decoding straight into its slice of the final output so that ONE in-place
all-gather (equal slices) or all-gather-v (ragged) reassembles the byte stream in order.  Host logic only."""
from __future__ import annotations

import numpy as np


def split_units(in_len, world: int):
    """-> list of (lo, hi) unit ranges, one per rank, contiguous, balanced by compressed bytes."""
    in_len = np.asarray(in_len, dtype=np.int64)
    n = len(in_len)
    if world <= 1 or n == 0:
        return [(0, n)] + [(n, n)] * (world - 1)
    csum = np.concatenate([[0], np.cumsum(in_len)])
    total = int(csum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        k = int(np.searchsorted(csum, target, side="left"))
        k = min(max(k, cuts[-1]), n)
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def out_slices(out_len, ranges):
    """Byte ranges of each rank's decoded output inside the final stream, given per-unit output sizes."""
    out_len = np.asarray(out_len, dtype=np.int64)
    csum = np.concatenate([[0], np.cumsum(out_len)])
    return [(int(csum[lo]), int(csum[hi])) for lo, hi in ranges]
