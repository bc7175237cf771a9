"""Sharding of a batch of read pairs across GPUs and the end-of-run STAT merge.

Pairs are independent (the reference resets its k-mer table per pair, assembler.c:113-116), so a batch is cut into
contiguous, near-equal slices, one per rank, with no collective on the data path.  The only cross-rank step is the
merge of the counter vectors: every entry adds, except the longest overlap which takes the maximum
(assembler.h:59-67,75-76; the reference prints one STAT block per worker and never merges, pool.c:86-104)."""
from __future__ import annotations

import numpy as np

from . import C_LONGEST, PB_NCOUNTERS


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) of rank's slice of n pairs."""
    return n * rank // world, n * (rank + 1) // world


def merge_counters(vectors) -> np.ndarray:
    """Host-side merge of per-shard counter vectors."""
    vectors = [np.asarray(v, dtype=np.int64) for v in vectors]
    out = np.sum(vectors, axis=0).astype(np.int64)
    out[C_LONGEST] = max(int(v[C_LONGEST]) for v in vectors)
    return out


def dist_merge_counters(t, dist):
    """All-rank merge of a torch int64 counter tensor (any backend: nccl on GPUs, gloo in the CPU tests)."""
    assert t.numel() == PB_NCOUNTERS
    longest = t[C_LONGEST].clone()
    total = t.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM)
    dist.all_reduce(longest, op=dist.ReduceOp.MAX)
    total[C_LONGEST] = longest
    return total
