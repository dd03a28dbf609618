"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
GPU box, gloo in CPU tests).  The hot path has exactly two exchange steps:

* RDF shards the sampled frames across ranks -> one all-reduce(sum, int64) of the
  ``[n_pairs][nbins]`` histograms;
* MSD / ACF / ionic current shard atoms across ranks -> all-reduce(sum, float64) of the series.

Everything else (unwrap, fits, coordination numbers) needs no communication.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


import contextlib

_local_only = False


@contextlib.contextmanager
def local_only():
    """Inside this context the calculators treat the process as a single rank: the caller has
    already given every rank its own shard (bench.py's weak-scaling run) and reduces itself."""
    global _local_only
    old, _local_only = _local_only, True
    try:
        yield
    finally:
        _local_only = old


def _dist():
    import torch.distributed as dist

    if _local_only:
        return None
    return dist if (dist.is_available() and dist.is_initialized()) else None


def world_size() -> int:
    d = _dist()
    return d.get_world_size() if d else 1


def rank() -> int:
    d = _dist()
    return d.get_rank() if d else 0


def shard_frames(frames: np.ndarray, r: int = None, w: int = None) -> np.ndarray:
    """Round-robin frame shard of this rank (cost per frame is uniform)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    return np.asarray(frames)[r::w]


def shard_atoms(a_lo: int, a_hi: int, r: int = None, w: int = None) -> Tuple[int, int]:
    """Contiguous atom block of [a_lo, a_hi) owned by this rank (block sizes differ by <= 1)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    n = a_hi - a_lo
    base, extra = divmod(n, w)
    lo = a_lo + r * base + min(r, extra)
    return lo, lo + base + (1 if r < extra else 0)


def all_reduce_sum_(tensors: List):
    """In-place sum over ranks of every tensor in ``tensors`` (no-op for one rank).  Issued on
    the current stream right behind the kernels that produced them."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return tensors
    for t in tensors:
        if t is not None:
            d.all_reduce(t, op=d.ReduceOp.SUM)
    return tensors
