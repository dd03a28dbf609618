"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink on the
GPU box, gloo in CPU tests).  The trajectory store is ATOM-SHARDED across ranks (store.py): a
rank keeps, uploads and writes back only its own atom block of every per-species dataset.  The
hot path has three exchange steps:

* RDF shards the sampled frames across ranks: the ranks swap (atom block x sampled frame)
  slabs with one all-to-all over NVLink so that each holds all atoms of its frames
  (``exchange_frames``), then one all-reduce(sum, int64) of the ``[n_pairs][nbins]`` histograms;
* MSD / ACF / ionic current run on the rank's own atom block -> all-reduce(sum, float64) of
  the series;
* unwrap needs no communication (per-atom scan over the rank's block).

Fits and coordination numbers run replicated on every rank.  Persistent state (dataset files,
index, result database) is created and written by rank 0 only (``is_root``), the other ranks
wait on a barrier; the cache-hit decision is broadcast so that all ranks take the same branch.
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


def _host_staged(d) -> bool:
    """gloo has no CUDA all-to-all and only some CUDA collectives: with that backend (CPU tests,
    and the two-ranks-on-one-GPU check of tests/multigpu_check.py) device tensors take a detour
    through the host.  NCCL, the production backend, works on the device tensors directly."""
    return d.get_backend() == "gloo"


def all_reduce_sum_(tensors: List):
    """In-place sum over ranks of every tensor in ``tensors`` (no-op for one rank).  Issued on
    the current stream right behind the kernels that produced them."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return tensors
    staged = _host_staged(d)
    for t in tensors:
        if t is None:
            continue
        if staged and t.is_cuda:
            h = t.cpu()
            d.all_reduce(h, op=d.ReduceOp.SUM)
            t.copy_(h)
        else:
            d.all_reduce(t, op=d.ReduceOp.SUM)
    return tensors


def is_root() -> bool:
    return rank() == 0


def barrier():
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.barrier()


def broadcast_object(obj, src: int = 0):
    """Rank ``src``'s picklable object on every rank (cache decisions, small results)."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return obj
    box = [obj if d.get_rank() == src else None]
    d.broadcast_object_list(box, src=src)
    return box[0]


def gather_rows(local, n_rows_total: int):
    """All ranks' contiguous row blocks (shard_atoms order) of a host array -> the whole array
    on every rank.  Collective; used by host reads of a sharded in-memory dataset."""
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return local
    parts = [None] * d.get_world_size()
    d.all_gather_object(parts, local)
    out = np.concatenate(parts, axis=0)
    assert out.shape[0] == n_rows_total
    return out


def _all_to_all(d, send, out_splits, in_splits):
    if _host_staged(d) and send.is_cuda:
        h_send = send.cpu()
        h_recv = h_send.new_empty(sum(out_splits))
        d.all_to_all_single(h_recv, h_send, out_splits, in_splits)
        return h_recv.to(send.device)
    recv = send.new_empty(sum(out_splits))
    d.all_to_all_single(recv, send, out_splits, in_splits)
    return recv


def exchange_frames(local, rows_per_rank, n_frames: int):
    """All-to-all of sampled frames: ``local`` is this rank's atom block of ALL sampled frames,
    a tensor [A_local][n_frames][3]; ``rows_per_rank[q]`` is the block size of rank q.  The
    result is [sum(rows_per_rank)][F_mine][3] holding every atom of the frames
    ``shard_frames(arange(n_frames))`` this rank owns.  Blocks are contiguous and in rank
    order, so the received pieces are already in atom order."""
    import torch

    d = _dist()
    if d is None or d.get_world_size() == 1:
        return local
    w, r = d.get_world_size(), d.get_rank()
    a_loc = local.shape[0]
    if a_loc != rows_per_rank[r]:
        raise ValueError("exchange_frames: local block does not match rows_per_rank")
    f_of = [len(range(q, n_frames, w)) for q in range(w)]
    send = torch.cat([local[:, q::w].reshape(-1) for q in range(w)])
    in_splits = [a_loc * f * 3 for f in f_of]
    out_splits = [int(n) * f_of[r] * 3 for n in rows_per_rank]
    recv = _all_to_all(d, send.contiguous(), out_splits, in_splits)
    return recv.view(int(sum(rows_per_rank)), f_of[r], 3)


def exchange_frame_groups(local, rows_per_rank, frames_per_rank):
    """As ``exchange_frames`` for an arbitrary frame ownership: the frames along axis 1 of
    ``local`` [A_local][F][3] are grouped by owning rank (``frames_per_rank[q]`` consecutive
    frames go to rank q); returns [sum(rows_per_rank)][frames_per_rank[rank]][3]."""
    import torch

    d = _dist()
    if d is None or d.get_world_size() == 1:
        return local
    w, r = d.get_world_size(), d.get_rank()
    a_loc = local.shape[0]
    if a_loc != rows_per_rank[r] or local.shape[1] != int(sum(frames_per_rank)):
        raise ValueError("exchange_frame_groups: local block does not match the layout")
    starts = np.concatenate([[0], np.cumsum(frames_per_rank)]).astype(int)
    send = torch.cat([local[:, starts[q]:starts[q + 1]].reshape(-1) for q in range(w)])
    in_splits = [a_loc * int(f) * 3 for f in frames_per_rank]
    out_splits = [int(n) * int(frames_per_rank[r]) * 3 for n in rows_per_rank]
    recv = _all_to_all(d, send.contiguous(), out_splits, in_splits)
    return recv.view(int(sum(rows_per_rank)), int(frames_per_rank[r]), 3)
