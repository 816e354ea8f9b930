"""Host-side sharding of the hot path over the GPUs of one box (one process per GPU).

Extraction: clips are independent (the delta stencil never leaves a clip, lib.rs:212-228), so each rank takes a
contiguous range of clips balanced by sample count and no collective is needed.  Training: batch-parallel -- every
global batch of the shuffled order (lib.rs:601-602) is cut into `world` slices, each rank runs forward/backward on its
slice and the flattened gradient (+ surviving-window count and loss) is all-reduced once per step, which makes the
update identical to the single-GPU step on the whole batch.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_clips(lengths: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous clip ranges [begin, end) per rank with near-equal total sample counts."""
    lengths = np.asarray(lengths, dtype=np.int64)
    total = int(lengths.sum())
    cum = np.concatenate([[0], np.cumsum(lengths)])
    bounds = [0]
    for r in range(1, world):
        target = total * r / world
        idx = int(np.searchsorted(cum, target, side="left"))
        if idx > 0 and abs(cum[idx - 1] - target) <= abs(cum[min(idx, len(cum) - 1)] - target):
            idx -= 1
        bounds.append(min(max(idx, bounds[-1]), len(lengths)))
    bounds.append(len(lengths))
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def shard_windows(n_windows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous window ranges [begin, end) per rank for ONE long clip (identification sweep, BASELINE configs[4]): the
    ranges differ by at most one window; each rank re-reads only a two-frame halo either side of its range."""
    n = int(n_windows)
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def shard_batches(perm: np.ndarray, batch: int, rank: int, world: int) -> Tuple[np.ndarray, List[int]]:
    """Slice every global batch of `perm` for `rank`.  Returns the concatenated local order and the local batch sizes
    (one per global step, possibly 0 for a rank when the last batch is short).  Every rank gets the same number of steps."""
    perm = np.asarray(perm)
    batch = max(1, int(batch))
    parts, sizes = [], []
    for s in range(0, len(perm), batch):
        chunk = perm[s:s + batch]
        lo, hi = len(chunk) * rank // world, len(chunk) * (rank + 1) // world
        parts.append(chunk[lo:hi])
        sizes.append(hi - lo)
    local = np.concatenate(parts) if parts else perm[:0]
    return local, sizes
