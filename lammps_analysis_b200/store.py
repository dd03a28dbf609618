"""Trajectory store: per-species arrays keyed ``"{species}/{Property}"`` with shape
``(n_atoms, n_frames, n_dims)`` float32 -- the MDSuite simulation database layout
(mdsuite/database/simulation_database.py:333-690: atom-major, time contiguous per atom,
float32 on disk).  h5py is not available in this environment, so datasets are ``.npy`` files
(memory-mapped on load) under ``<experiment>/database/``; the *interface* the hot path needs is
kept: ``add_dataset / add_data / check_existence / get_data_size / load_data``.

HBM residency: ``device(path, ...)`` returns a CUDA tensor of (a row range of) a dataset and
keeps it cached, so that consecutive calculators do not re-upload; uploads go through pinned
host staging in bounded chunks.

Multi-rank (one process per GPU): the store is ATOM-SHARDED.  Every per-species dataset keeps
its global shape, but a rank owns the contiguous row block ``distributed.shard_atoms`` assigns
to it: an in-memory store holds (page-locks, uploads, writes back) only that block; a directory
store maps the whole file -- created by rank 0, the others wait on a barrier -- and every rank
reads and writes only its own rows of it.  ``Observables/*`` datasets are replicated.  Row
arguments and results of the public methods are always GLOBAL row indices.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np

from . import distributed as D
from . import trace
from .config import config


def _cuda_available() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # pragma: no cover
        return False


class _PinnedBlock:
    """Exactly-sized page-locked host block: page-aligned memory registered with
    cudaHostRegister.  (torch's pinned allocator rounds every request up to a power of two --
    a 12 GB dataset would lock 16 GiB -- and page-locking costs ~0.4 s per GB on the GPU boxes,
    so the size matters.)"""

    PAGE = 4096

    def __init__(self, nbytes: int):
        import torch

        self.nbytes = int(nbytes)
        raw = np.empty(self.nbytes + self.PAGE, dtype=np.uint8)
        off = (-raw.ctypes.data) % self.PAGE
        self._raw = raw
        self.bytes = raw[off:off + self.nbytes]
        self._registered = False
        if self.nbytes:
            rc = torch.cuda.cudart().cudaHostRegister(self.bytes.ctypes.data, self.nbytes, 0)
            if int(rc) != 0:
                from ._lib import MdkError

                raise MdkError(f"cudaHostRegister of {self.nbytes} bytes failed with code {int(rc)}")
            self._registered = True

    def tensor(self, shape):
        import torch

        n = int(np.prod(shape))
        return torch.from_numpy(self.bytes[:4 * n].view(np.float32).reshape(shape))

    def __del__(self):
        if getattr(self, "_registered", False):
            try:
                import torch

                torch.cuda.cudart().cudaHostUnregister(self.bytes.ctypes.data)
            except Exception:  # interpreter shutdown
                pass
            self._registered = False


class _PinnedPool:
    """Free list of page-locked blocks keyed by size: a dataset that is removed or replaced
    hands its block back, the next dataset of that size (a transformation re-run, the next
    experiment of the same shape) takes it without paying the page-locking again."""

    def __init__(self):
        self.free: Dict[int, list] = {}
        self.cached_bytes = 0

    def acquire(self, nbytes: int) -> _PinnedBlock:
        blocks = self.free.get(nbytes)
        if blocks:
            self.cached_bytes -= nbytes
            return blocks.pop()
        return _PinnedBlock(nbytes)

    def release(self, block: _PinnedBlock):
        limit = config.pinned_pool_bytes
        if block.nbytes == 0 or self.cached_bytes + block.nbytes > limit:
            return                      # dropped: unregistered when the last reference goes
        self.free.setdefault(block.nbytes, []).append(block)
        self.cached_bytes += block.nbytes

    def clear(self):
        self.free.clear()
        self.cached_bytes = 0


pinned_pool = _PinnedPool()


def join_path(*args) -> str:
    """meta_functions.join_path (:73-93): database paths always use '/'."""
    return "/".join(args)


class TrajectoryStore:
    def __init__(self, directory: Optional[str] = None, pinned: Optional[bool] = None,
                 sharded: Optional[bool] = None):
        self.directory = directory
        # the shard geometry is fixed at construction: (rank, world) of the process group, or
        # (0, 1) for an unsharded store (single process, or sharded=False: the caller keeps a
        # full private copy per rank, as bench.py's weak-scaling replicas do)
        if sharded is None:
            sharded = D.world_size() > 1
        self.rank, self.world = (D.rank(), D.world_size()) if sharded else (0, 1)
        self._arrays: Dict[str, np.ndarray] = {}
        self._rows: Dict[str, int] = {}     # global row count of a dataset
        self._row0: Dict[str, int] = {}     # global index of the first row of the host array
        # in-memory stores live in page-locked host memory when a GPU is present: uploads and
        # downloads are then single DMA transfers without a staging copy
        if pinned is None:
            pinned = directory is None and _cuda_available()
        self.pinned = bool(pinned) and directory is None
        self._pinned: Dict[str, object] = {}
        self._blocks: Dict[str, _PinnedBlock] = {}
        self._device_cache: "OrderedDict[Tuple, object]" = OrderedDict()
        self._device_bytes = 0
        self.h2d_bytes = 0  # bytes uploaded so far (bench.py reads this)
        # device -> host write-backs in flight on the copy stream, per dataset: whoever reads a
        # dataset's host copy (or replaces it) drains them first (_drain)
        self._copy_stream = None
        self._pending: Dict[str, list] = {}
        # device copies that are still being filled block by block (an upload in flight on the
        # upload stream, or a transformation writing row blocks): cache key -> [(r0, r1, event)]
        # with rows relative to the cached tensor.  device() waits for them; device_blocks() hands
        # them to a consumer that wants to start on the first rows while the rest arrives.
        self._upload_stream = None
        self._block_events: Dict[Tuple, list] = {}
        if directory is not None:
            os.makedirs(directory, exist_ok=True)
            self._load_index()

    # ---- persistence -------------------------------------------------------------------
    def _index_path(self):
        return os.path.join(self.directory, "index.json")

    def _file_of(self, path: str) -> str:
        return os.path.join(self.directory, path.replace("/", "__") + ".npy")

    def _load_index(self):
        if os.path.exists(self._index_path()):
            with open(self._index_path()) as fh:
                for path in json.load(fh):
                    if os.path.exists(self._file_of(path)):
                        self._map_file(path)

    def _map_file(self, path: str):
        arr = np.load(self._file_of(path), mmap_mode="r+")
        self._arrays[path] = arr
        self._rows[path] = arr.shape[0]
        self._row0[path] = 0

    def _save_index(self):
        if self.directory is not None and self._is_writer():
            with open(self._index_path(), "w") as fh:
                json.dump(sorted(self._arrays), fh)

    # ---- sharding ---------------------------------------------------------------------
    def _is_writer(self) -> bool:
        """Rank that creates files and writes the index of a shared directory store."""
        return self.rank == 0

    def _barrier(self):
        if self.world > 1:
            self._check_group()
            D.barrier()

    def _check_group(self):
        if self.world != D.world_size():
            from ._lib import MdkError

            raise MdkError(f"store sharded over {self.world} rank(s) used in a process group of "
                           f"{D.world_size()} (inside distributed.local_only()?)")

    def is_sharded(self, path: str) -> bool:
        return self.world > 1 and not path.startswith("Observables/")

    def owned_rows(self, path: str) -> Tuple[int, int]:
        """Global row range [lo, hi) of ``path`` this rank reads, uploads and writes."""
        n = self._rows[path]
        if not self.is_sharded(path):
            return 0, n
        self._check_group()
        return D.shard_atoms(0, n, self.rank, self.world)

    def rows_per_rank(self, path: str):
        """[(lo, hi)] of every rank, in rank order."""
        n = self._rows[path]
        if not self.is_sharded(path):
            return [(0, n)]
        return [D.shard_atoms(0, n, r, self.world) for r in range(self.world)]

    def _local(self, path: str, lo: int, hi: int) -> np.ndarray:
        """Host view of the global rows [lo, hi); they must be held by this rank."""
        arr, r0 = self._arrays[path], self._row0[path]
        if hi > lo and (lo < r0 or hi > r0 + arr.shape[0]):
            from ._lib import MdkError

            raise MdkError(f"rows [{lo}, {hi}) of {path} are not held by rank {self.rank} "
                           f"(it holds [{r0}, {r0 + arr.shape[0]}))")
        return arr[lo - r0:hi - r0]

    # ---- simulation_database.Database interface ------------------------------------------
    def check_existence(self, path: str) -> bool:
        """simulation_database.py:546-572."""
        return path in self._arrays

    def add_dataset(self, path: str, shape: Tuple[int, int, int]):
        """simulation_database.py:452-497: float32 dataset of (n_rows, n_frames, n_dims)."""
        if path in self._arrays:
            raise ValueError(f"dataset {path} already exists")
        shape = tuple(int(s) for s in shape)
        self._rows[path] = shape[0]
        if self.directory is not None:
            # one file for the whole dataset: rank 0 creates it, the others map it afterwards
            if self._is_writer():
                arr = np.lib.format.open_memmap(self._file_of(path), mode="w+", dtype=np.float32,
                                                shape=shape)
                arr.flush()
            self._barrier()
            if not self._is_writer():
                arr = np.load(self._file_of(path), mmap_mode="r+")
            self._row0[path] = 0
        else:
            lo, hi = self.owned_rows(path)
            local = (hi - lo,) + shape[1:]
            self._row0[path] = lo
            if self.pinned:
                # not zero-filled: every writer (ingest, transformations) covers the whole
                # dataset, and a memset of gigabytes of page-locked memory costs as much as the
                # transfer
                block = pinned_pool.acquire(int(np.prod(local)) * 4)
                t = block.tensor(local)
                self._blocks[path] = block
                self._pinned[path] = t
                arr = t.numpy()
            else:
                arr = np.zeros(local, dtype=np.float32)
        self._arrays[path] = arr
        self._save_index()
        return arr

    def _drop_pinned(self, path: str):
        """Forget the page-locked block behind ``path`` and return it to the pool."""
        self._pinned.pop(path, None)
        block = self._blocks.pop(path, None)
        if block is not None:
            pinned_pool.release(block)

    def _drain(self, path: Optional[str] = None):
        """Wait for the asynchronous write-backs of one dataset (or of all of them)."""
        keys = [path] if path is not None else list(self._pending)
        for k in keys:
            for ev in self._pending.pop(k, []):
                ev.synchronize()

    def flush(self):
        """Block until every asynchronous device -> host write-back has landed."""
        self._drain(None)

    def resize_dataset(self, path: str, n_frames: int):
        """simulation_database.py:380-420 (extend along the frame axis)."""
        self._drain(path)
        old = self._arrays[path]
        if old.shape[1] >= n_frames:
            return old
        lo, hi = self.owned_rows(path)
        data = np.array(self._local(path, lo, hi))      # this rank's rows survive the resize
        n_rows = self._rows[path]
        self._barrier()                                  # everyone has read the old file
        del self._arrays[path], old
        self._drop_pinned(path)
        self.invalidate(path)
        self.add_dataset(path, (n_rows, n_frames, data.shape[2]))
        self._local(path, lo, hi)[:, : data.shape[1]] = data
        return self._arrays[path]

    def add_data(self, path: str, data, start: int = 0, rows: Optional[Tuple[int, int]] = None):
        """Write ``data`` (n_rows, k, n_dims) at frame offset ``start``; values are rounded to
        float32 exactly as the HDF5 store does (simulation_database.py:333-378).  ``data``
        covers the whole dataset (every rank keeps its own block of it), this rank's block, or
        the global row range ``rows``."""
        self._drain(path)
        data = np.asarray(data)
        lo, hi = self.owned_rows(path)
        if rows is not None:
            # data holds the global rows [rows[0], rows[1]): keep what this rank owns of them
            g0, g1 = int(rows[0]), int(rows[1])
            lo, hi = max(lo, g0), min(hi, g1)
            data = data[max(lo - g0, 0):max(hi - g0, 0)]
        elif data.shape[0] == self._rows[path]:
            data = data[lo:hi]                           # global array: every rank reads the source
        elif data.shape[0] != hi - lo:
            raise ValueError(f"add_data({path}): {data.shape[0]} rows given, expected the whole "
                             f"dataset ({self._rows[path]}) or this rank's block ({hi - lo})")
        shared_replica = self.directory is not None and self.world > 1 and \
            not self.is_sharded(path)
        if hi > lo and (not shared_replica or self._is_writer()):
            self._local(path, lo, hi)[:, start:start + data.shape[1]] = \
                data.astype(np.float32, copy=False)
        if shared_replica:
            self._barrier()        # one file for all ranks: rank 0 wrote it
        self.invalidate(path)

    def put(self, path: str, array):
        """Create-or-replace a whole dataset from an array (ScriptInput-style ingest)."""
        array = np.asarray(array)
        if array.ndim != 3:
            raise ValueError("datasets are (n_rows, n_frames, n_dims)")
        if path in self._arrays:
            self._drain(path)
            self._barrier()
            del self._arrays[path]
            self._drop_pinned(path)
            self.invalidate(path)
        self.add_dataset(path, array.shape)
        self.add_data(path, array)
        return self._arrays[path]

    def get_data_size(self, path: str):
        """(n_rows, n_configurations, n_bytes) -- simulation_database.py:683-690."""
        n, t, d = self.shape(path)
        return n, t, int(n * t * d * 4)

    def shape(self, path: str):
        """Global shape (n_rows, n_frames, n_dims)."""
        a = self._arrays[path]
        return (self._rows[path],) + tuple(a.shape[1:])

    def load_data(self, path: str, select_slice=np.s_[:]) -> np.ndarray:
        """Host read as float64 (simulation_database.py:594-639 casts to tf.float64)."""
        self._drain(path)
        if self.is_sharded(path):
            # collective: every rank contributes its block (in-memory store) or waits for the
            # other ranks' writes to the shared file (directory store)
            if self.directory is None:
                lo, hi = self.owned_rows(path)
                whole = D.gather_rows(np.array(self._local(path, lo, hi)), self._rows[path])
                return np.asarray(whole[select_slice], dtype=np.float64)
            self._barrier()
        return np.asarray(self._arrays[path][select_slice], dtype=np.float64)

    def host(self, path: str) -> np.ndarray:
        """Host array of the rows this rank owns (the whole dataset for one rank)."""
        self._drain(path)
        lo, hi = self.owned_rows(path)
        return self._local(path, lo, hi)

    def paths(self):
        return sorted(self._arrays)

    # ---- HBM residency ------------------------------------------------------------------------
    def _budget(self) -> int:
        if config.device_cache_bytes is not None:
            return int(config.device_cache_bytes)
        import torch

        free, _total = torch.cuda.mem_get_info()
        config.device_cache_bytes = int(0.6 * free)
        return config.device_cache_bytes

    def invalidate(self, path: Optional[str] = None):
        """Drop cached device copies (of one dataset, or all)."""
        for key in [k for k in self._device_cache if path is None or k[0] == path]:
            t = self._device_cache.pop(key)
            self._device_bytes -= t.numel() * t.element_size()
            self._block_events.pop(key, None)

    def device(self, path: str, rows: Optional[Tuple[int, int]] = None,
               row_index: Optional[np.ndarray] = None, device=None):
        """CUDA float32 tensor of the global dataset rows [lo, hi) (default: the rows this rank
        owns) or of a fancy row selection (global indices); the rows must be held by this rank."""
        import torch

        from ._lib import MdkError

        if not torch.cuda.is_available():
            raise MdkError("CUDA device required: the trajectory store has no CPU compute path")
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        own = self.owned_rows(path)
        if row_index is not None:
            row_index = np.asarray(row_index)
            key = (path, "idx", hash(row_index.tobytes()), str(dev))
        else:
            lo, hi = own if rows is None else (int(rows[0]), int(rows[1]))
            key = (path, lo, hi, str(dev))
        hit = self._device_cache.get(key)
        if hit is not None:
            self._device_cache.move_to_end(key)
            self._wait_blocks(key)
            return hit
        if row_index is None:
            wkey = (path, own[0], own[1], str(dev))
            whole = self._device_cache.get(wkey)
            if whole is not None and own[0] <= lo and hi <= own[1]:
                self._wait_blocks(wkey)
                return whole[lo - own[0]:hi - own[0]]   # rows are the leading axis: a view
        self._drain(path)                    # the host copy is read from here on
        r0 = self._row0[path]
        if row_index is not None:
            if len(row_index) and (row_index.min() < r0
                                   or row_index.max() >= r0 + self._arrays[path].shape[0]):
                raise MdkError(f"row selection of {path} reaches outside the rows held by rank "
                               f"{self.rank}")
            src = self._arrays[path][row_index - r0]
        else:
            src = self._local(path, lo, hi)
        nbytes = int(np.prod(src.shape)) * 4
        budget = self._budget()
        while self._device_cache and self._device_bytes + nbytes > budget:
            okey, old = self._device_cache.popitem(last=False)
            self._device_bytes -= old.numel() * old.element_size()
            self._block_events.pop(okey, None)
        out = torch.empty(src.shape, dtype=torch.float32, device=dev)
        pin = self._pinned.get(path)
        if pin is not None and row_index is None:
            out.copy_(pin[lo - r0:hi - r0], non_blocking=True)   # one DMA from page-locked memory
            torch.cuda.current_stream().synchronize()
            self.h2d_bytes += nbytes
        else:
            self._upload(src, out)
        self._device_cache[key] = out
        self._device_bytes += nbytes
        return out

    def upload_stream(self, device=None):
        """The one stream all block-wise host -> device copies of this store are queued on: they
        share the link anyway, and FIFO order means the dataset requested first is complete
        first (its consumer starts while the next dataset is still on the wire)."""
        import torch

        if self._upload_stream is None:
            self._upload_stream = torch.cuda.Stream(device=device)
        return self._upload_stream

    def _wait_blocks(self, key):
        """The current stream waits until a block-wise filled device copy is complete."""
        import torch

        blocks = self._block_events.pop(key, None)
        if blocks:
            cur = torch.cuda.current_stream()
            for _, _, ev in blocks:
                if ev is not None:
                    cur.wait_event(ev)

    def device_blocks(self, path: str, block_bytes: Optional[int] = None, device=None):
        """The rows this rank owns as a CUDA tensor that may still be filling up, plus the row
        blocks it arrives in: (tensor, [(r0, r1, event | None), ...]) with rows relative to the
        tensor.  A consumer launches its kernel on rows [r0, r1) after ``wait_event(event)``
        and so overlaps with the host -> device copy (or the transformation) that produces the
        following blocks.  Page-locked datasets are uploaded block by block on a side stream;
        anything else falls back to ``device()`` and one block."""
        import torch

        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if block_bytes is None:
            block_bytes = config.upload_block_bytes
        lo, hi = self.owned_rows(path)
        key = (path, lo, hi, str(dev))
        hit = self._device_cache.get(key)
        if hit is not None:
            self._device_cache.move_to_end(key)
            return hit, list(self._block_events.get(key) or [(0, hi - lo, None)])
        pin = self._pinned.get(path)
        row_bytes = int(np.prod(self._arrays[path].shape[1:])) * 4
        rows_per = max(1, int(block_bytes // max(row_bytes, 1)))
        if pin is None or hi - lo <= rows_per:
            return self.device(path, device=dev), [(0, hi - lo, None)]
        self._drain(path)
        nbytes = (hi - lo) * row_bytes
        budget = self._budget()
        while self._device_cache and self._device_bytes + nbytes > budget:
            okey, old = self._device_cache.popitem(last=False)
            self._device_bytes -= old.numel() * old.element_size()
            self._block_events.pop(okey, None)
        out = torch.empty((hi - lo,) + tuple(pin.shape[1:]), dtype=torch.float32, device=dev)
        up = self.upload_stream(dev)
        # `out` may be recycled memory whose previous users were queued on the current stream:
        # the upload is ordered behind what is queued there NOW (not behind later work)
        fence = torch.cuda.Event()
        fence.record()
        up.wait_event(fence)
        r0 = self._row0[path]
        blocks = []
        with torch.cuda.stream(up):
            for b0 in range(0, hi - lo, rows_per):
                b1 = min(hi - lo, b0 + rows_per)
                out[b0:b1].copy_(pin[lo - r0 + b0:lo - r0 + b1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                blocks.append((b0, b1, ev))
                trace.event(f"{path} upload rows [{b0}, {b1}) done")
        out.record_stream(up)
        self.h2d_bytes += nbytes
        self._device_cache[key] = out
        self._device_bytes += nbytes
        self._block_events[key] = blocks
        return out, list(blocks)

    def pinned_tensor(self, path: str):
        """The page-locked host tensor behind this rank's rows of a dataset of an in-memory
        store (or None).  Row 0 of the tensor is global row ``owned_rows(path)[0]``."""
        self._drain(path)
        return self._pinned.get(path)

    def is_resident(self, path: str) -> bool:
        lo, hi = self.owned_rows(path)
        return any(k[0] == path and k[1] == lo and k[2] == hi
                   for k in self._device_cache if k[1] != "idx")

    def adopt_device(self, path: str, tensor, blocks=None):
        """Register a device tensor that already holds this rank's rows of the dataset ``path``
        (e.g. the output a transformation just produced), so that the next calculator does not
        re-upload it.  ``blocks`` = [(r0, r1, event)]: the row blocks the producer writes it in
        (kernels still queued): see device_blocks()."""
        self.invalidate(path)
        lo, hi = self.owned_rows(path)
        if tensor.shape[0] != hi - lo:
            raise ValueError("adopt_device: tensor does not hold the rows this rank owns")
        key = (path, lo, hi, str(tensor.device))
        self._device_cache[key] = tensor
        self._device_bytes += tensor.numel() * tensor.element_size()
        if blocks:
            self._block_events[key] = list(blocks)

    def device_frames(self, path: str, frames, row_index=None, device=None):
        """CUDA float32 [rows][len(frames)][dims] holding only the selected frames of the rows
        this rank owns (or of the global row selection ``row_index`` within them); not cached:
        the RDF samples a few frames of a long trajectory."""
        import torch

        self._drain(path)
        lo, hi = self.owned_rows(path)
        arr = self._local(path, lo, hi)
        frames = np.asarray(frames, dtype=np.int64)
        src = arr[:, frames] if row_index is None else \
            arr[np.asarray(row_index) - lo][:, frames]
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        out = torch.empty(src.shape, dtype=torch.float32, device=dev)
        self._upload(np.ascontiguousarray(src), out)
        return out

    def write_from_device(self, path: str, tensor, t0: int = 0, row0: Optional[int] = None):
        """Device tensor (n_rows, k, n_dims) -> frames [t0, t0 + k) of the global rows
        [row0, row0 + n_rows) of the host dataset (default: the rows this rank owns)."""
        import torch

        lo, hi = self.owned_rows(path)
        row0 = lo if row0 is None else int(row0)
        arr = self._local(path, row0, row0 + tensor.shape[0])
        k = tensor.shape[1]
        pin = self._pinned.get(path)
        if pin is not None:
            pin = pin[row0 - self._row0[path]:row0 - self._row0[path] + tensor.shape[0]]
        if pin is not None:
            # asynchronous write-back on a side stream: PCIe is full duplex, so the copy overlaps
            # the uploads and kernels of whatever runs next; readers of the host copy drain it
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=tensor.device)
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(ready)
                pin[:, t0:t0 + k].copy_(tensor, non_blocking=True)
                done = torch.cuda.Event()
                done.record()
                trace.event(f"{path} write-back rows from {row0} done")
            tensor.record_stream(self._copy_stream)
            self._pending.setdefault(path, []).append(done)
        else:
            arr[:, t0:t0 + k] = tensor.cpu().numpy()

    def remove(self, path: str):
        """Delete a dataset (host, disk and device copies)."""
        self._drain(path)
        self.invalidate(path)
        if self.directory is not None:
            self._barrier()
        arr = self._arrays.pop(path, None)
        self._drop_pinned(path)
        self._rows.pop(path, None)
        self._row0.pop(path, None)
        del arr
        if self.directory is not None and self._is_writer() and \
                os.path.exists(self._file_of(path)):
            os.remove(self._file_of(path))
        self._save_index()

    def _upload(self, src: np.ndarray, dst, chunk_bytes: int = 256 << 20):
        """Host -> device through a pinned staging buffer, double buffered."""
        import torch

        n_rows = src.shape[0]
        if n_rows == 0:
            return
        row_bytes = int(np.prod(src.shape[1:])) * 4
        rows_per = max(1, chunk_bytes // max(row_bytes, 1))
        stage = [torch.empty((min(rows_per, n_rows),) + tuple(src.shape[1:]), dtype=torch.float32,
                             pin_memory=True) for _ in range(2)]
        events = [None, None]
        for i, lo in enumerate(range(0, n_rows, rows_per)):
            hi = min(n_rows, lo + rows_per)
            s = stage[i & 1]
            if events[i & 1] is not None:
                events[i & 1].synchronize()
            s[: hi - lo].numpy()[...] = src[lo:hi]
            dst[lo:hi].copy_(s[: hi - lo], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            events[i & 1] = ev
            self.h2d_bytes += (hi - lo) * row_bytes
        torch.cuda.current_stream().synchronize()
