"""Device-side entry points: torch CUDA tensors in, libmdk.so (include/mdk.h) underneath.

PyTorch is only the buffer carrier here (allocation, streams, host<->device copies); all
arithmetic on the hot path happens in the hand-written sm_100a kernels of ``csrc/``.
There is no CPU fallback: every function raises when CUDA or libmdk.so is missing.

Each function names the reference code it replaces (paths relative to the MDSuite repo).
"""
from __future__ import annotations

import ctypes as C
import os
import functools
import itertools
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from ._lib import MdkError, check

# number of libmdk kernel launches issued through this module (bench.py reports it)
launch_count = 0


def _count(n=1):
    global launch_count
    launch_count += n


def _ptr(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t: torch.Tensor, dtype, what: str):
    if not t.is_cuda:
        raise MdkError(f"{what}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise MdkError(f"{what}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise MdkError(f"{what}: tensor must be contiguous")


def _need_device_readable(t: torch.Tensor, dtype, what: str):
    """CUDA tensor, or page-locked host tensor (mapped into the device address space under
    unified addressing: a gather kernel reads it in place over PCIe / NVLink-C2C)."""
    if not (t.is_cuda or t.is_pinned()):
        raise MdkError(f"{what}: expected a CUDA or pinned host tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise MdkError(f"{what}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise MdkError(f"{what}: tensor must be contiguous")


def read_back(*tensors: torch.Tensor):
    """Device tensors -> NumPy arrays through ``mdk_store_mapped``: the SMs write the values
    into mapped page-locked memory on the CURRENT stream and the host waits for that stream
    only.  A ``.cpu()`` would go through a copy engine, where it queues behind every bulk
    upload already handed to the engine -- the streamed calculators read their results while
    the next dataset is on the wire."""
    lib = _lib.load()
    outs = []
    for t in tensors:
        _need_cuda(t, t.dtype, "read_back source")
        t = t.contiguous()
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        check(lib.mdk_store_mapped(_ptr(t), C.c_void_p(host.data_ptr()),
                                   t.numel() * t.element_size(), _stream()), "mdk_store_mapped")
        _count(1)
        outs.append((host, t))          # keep the source alive until the stream has run
    ev = torch.cuda.Event()
    ev.record()
    ev.synchronize()
    return [h.numpy().copy() for h, _ in outs]


def sm_count() -> int:
    return _lib.load().mdk_sm_count()


def peak_fp32(packed: bool = True, iters: int = 20000) -> float:
    """Measured FP32 FMA throughput in TFLOP/s (roofline denominator of the RDF kernel)."""
    out = C.c_double(0.0)
    check(_lib.load().mdk_peak_fp32(int(packed), int(iters), C.byref(out)), "mdk_peak_fp32")
    return out.value


# --------------------------------------------------------------------------------------
# RDF
# --------------------------------------------------------------------------------------
def rdf_thresholds(cutoff: float, nbins: int):
    """Cached front end of :func:`_rdf_thresholds` (the table depends on (cutoff, nbins) only;
    repeated calculator runs reuse it)."""
    thr, cut2 = _rdf_thresholds(float(np.float32(cutoff)), int(nbins))
    return thr.copy(), cut2


@functools.lru_cache(maxsize=32)
def _rdf_thresholds(cutoff: float, nbins: int):
    """Host table of fp32 thresholds on d^2 reproducing tf.histogram_fixed_width exactly.

    Returns (thr float32[nbins + 1], cut2 float).  Replaces bin_minibatch
    (radial_distribution_function.py:616-645) + apply_system_cutoff (utils/linalg.py:125-136).
    """
    thr = np.empty(nbins + 1, dtype=np.float32)
    cut2 = C.c_float(0.0)
    check(
        _lib.load().mdk_rdf_thresholds(
            C.c_float(cutoff), int(nbins), thr.ctypes.data_as(C.c_void_p), C.byref(cut2)
        ),
        "mdk_rdf_thresholds",
    )
    return thr, float(cut2.value)


@dataclass
class RdfLayout:
    """Frame-major SoA layout of the packed positions: species blocks padded to the tile."""

    counts: list          # atoms per species that enter the pair pass
    tile: int
    sp_lo: np.ndarray = field(init=False)
    sp_hi: np.ndarray = field(init=False)
    n_pad: int = field(init=False)

    def __post_init__(self):
        lo, hi, cur = [], [], 0
        for n in self.counts:
            lo.append(cur)
            hi.append(cur + n)
            cur += ((n + self.tile - 1) // self.tile) * self.tile
        self.sp_lo = np.asarray(lo, dtype=np.int32)
        self.sp_hi = np.asarray(hi, dtype=np.int32)
        self.n_pad = max(cur, self.tile)

    @property
    def n_species(self):
        return len(self.counts)

    @property
    def n_pairs(self):
        s = self.n_species
        return s * (s + 1) // 2

    def pair_keys(self):
        return list(itertools.combinations_with_replacement(range(self.n_species), 2))


def rdf_layout(counts) -> RdfLayout:
    return RdfLayout(list(int(c) for c in counts), _lib.load().mdk_rdf_tile())


def rdf_pack(traj: torch.Tensor, frames: torch.Tensor, out: torch.Tensor, layout: RdfLayout,
             species_index: int, atom_first: int, atom_count: int):
    """Gather frames of one species' atom-major [A][T][3] array into out[k][3][n_pad].

    Replaces data_manager.py:195-201 (frame fancy index) and _format_data
    (radial_distribution_function.py:535-563).
    """
    _need_device_readable(traj, torch.float32, "rdf_pack traj")
    _need_cuda(frames, torch.int32, "rdf_pack frames")
    _need_cuda(out, torch.float32, "rdf_pack out")
    A, T, D = traj.shape
    if D != 3:
        raise MdkError("rdf_pack: trajectory must be [A][T][3]")
    nf = frames.numel()
    if out.numel() < nf * 3 * layout.n_pad:
        raise MdkError("rdf_pack: output buffer too small")
    lo = int(layout.sp_lo[species_index])
    span = ((atom_count + layout.tile - 1) // layout.tile) * layout.tile
    lib = _lib.load()
    for k0 in range(0, nf, 65535):
        k1 = min(nf, k0 + 65535)
        check(
            lib.mdk_rdf_pack(
                _ptr(traj), A, T, atom_first, atom_count,
                C.c_void_p(frames.data_ptr() + 4 * k0), k1 - k0,
                C.c_void_p(out.data_ptr() + 4 * k0 * 3 * layout.n_pad),
                layout.n_pad, lo, span, _stream(),
            ),
            "mdk_rdf_pack",
        )
        _count()


def gather_frames(traj: torch.Tensor, frames: torch.Tensor, out: torch.Tensor = None):
    """out[a][k] = traj[a][frames[k]] for an atom block [A][T][3] (CUDA or pinned host memory);
    returns the CUDA tensor [A][len(frames)][3].  Replaces data_manager.py:195-201."""
    _need_device_readable(traj, torch.float32, "gather_frames traj")
    _need_cuda(frames, torch.int32, "gather_frames frames")
    A, T, D = traj.shape
    if D != 3:
        raise MdkError("gather_frames: trajectory must be [A][T][3]")
    nf = frames.numel()
    if out is None:
        out = torch.empty(A, nf, 3, dtype=torch.float32, device=frames.device)
    _need_cuda(out, torch.float32, "gather_frames out")
    if out.numel() != A * nf * 3:
        raise MdkError("gather_frames: output must hold A * n_frames * 3 values")
    check(_lib.load().mdk_gather_frames(_ptr(traj), A, T, _ptr(frames), nf, _ptr(out), _stream()),
          "mdk_gather_frames")
    _count()
    return out


def coord_extent(pos_soa: torch.Tensor, n_frames: int, n_pad: int) -> np.ndarray:
    """Per-dimension [min xyz, max xyz] of a packed frame array (synchronises)."""
    _need_cuda(pos_soa, torch.float32, "coord_extent")
    mm = torch.tensor([np.inf] * 3 + [-np.inf] * 3, dtype=torch.float32, device=pos_soa.device)
    check(_lib.load().mdk_coord_extent(_ptr(pos_soa), n_frames, n_pad, _ptr(mm), _stream()),
          "mdk_coord_extent")
    _count()
    return mm.cpu().numpy()


RDF_SUBTILE = int(os.environ.get("MDK_RDF_SUBTILE", 32))  # MDK_RDF_SUBTILE (env: A/B builds)


def rdf_sort_workspace(max_atoms: int) -> int:
    return int(_lib.load().mdk_rdf_sort_workspace(int(max_atoms)))


def rdf_pack_sorted(traj: torch.Tensor, frame: int, out: torch.Tensor, k: int, layout: RdfLayout,
                    species_index: int, atom_first: int, atom_count: int, box,
                    workspace: torch.Tensor):
    """Morton-ordered pack of one frame of one species into slab k of ``out`` ([k][3][n_pad])."""
    _need_device_readable(traj, torch.float32, "rdf_pack_sorted traj")
    _need_cuda(out, torch.float32, "rdf_pack_sorted out")
    A, T, D = traj.shape
    if D != 3:
        raise MdkError("rdf_pack_sorted: trajectory must be [A][T][3]")
    lo = int(layout.sp_lo[species_index])
    span = ((atom_count + layout.tile - 1) // layout.tile) * layout.tile
    box32 = np.asarray(box, dtype=np.float32)
    check(
        _lib.load().mdk_rdf_pack_sorted(
            _ptr(traj), A, T, atom_first, atom_count, int(frame),
            C.c_void_p(out.data_ptr() + 4 * k * 3 * layout.n_pad), layout.n_pad, lo, span,
            box32.ctypes.data_as(C.c_void_p), _ptr(workspace),
            workspace.numel() * workspace.element_size(), _stream(),
        ),
        "mdk_rdf_pack_sorted",
    )
    _count(3)


def rdf_sort_batch_workspace(max_atoms: int, n_frames: int) -> int:
    """Scratch bytes of rdf_pack_sorted_batch, or -1 when the batch is too large for it."""
    return int(_lib.load().mdk_rdf_sort_batch_workspace(int(max_atoms), int(n_frames)))


def rdf_pack_sorted_batch(traj: torch.Tensor, frames: torch.Tensor, out: torch.Tensor,
                          layout: RdfLayout, species_index: int, atom_first: int, atom_count: int,
                          box, workspace: torch.Tensor):
    """Hilbert-ordered pack of a batch of frames of one species (one radix sort for the batch)
    into the slabs 0 .. len(frames) - 1 of ``out`` ([F][3][n_pad]).  frames: CUDA int32."""
    _need_device_readable(traj, torch.float32, "rdf_pack_sorted_batch traj")
    _need_cuda(out, torch.float32, "rdf_pack_sorted_batch out")
    _need_cuda(frames, torch.int32, "rdf_pack_sorted_batch frames")
    A, T, D = traj.shape
    if D != 3:
        raise MdkError("rdf_pack_sorted_batch: trajectory must be [A][T][3]")
    lo = int(layout.sp_lo[species_index])
    span = ((atom_count + layout.tile - 1) // layout.tile) * layout.tile
    box32 = np.asarray(box, dtype=np.float32)
    check(
        _lib.load().mdk_rdf_pack_sorted_batch(
            _ptr(traj), A, T, atom_first, atom_count, _ptr(frames), frames.numel(), _ptr(out),
            layout.n_pad, lo, span, box32.ctypes.data_as(C.c_void_p), _ptr(workspace),
            workspace.numel() * workspace.element_size(), _stream(),
        ),
        "mdk_rdf_pack_sorted_batch",
    )
    _count(3)


def rdf_bbox(pos_soa: torch.Tensor, n_frames: int, layout: RdfLayout, bbox: torch.Tensor):
    """Bounding boxes [n_frames][n_pad / RDF_SUBTILE][6] of a packed frame array."""
    _need_cuda(pos_soa, torch.float32, "rdf_bbox pos")
    _need_cuda(bbox, torch.float32, "rdf_bbox out")
    if bbox.numel() < n_frames * (layout.n_pad // RDF_SUBTILE) * 6:
        raise MdkError("rdf_bbox: output too small")
    check(_lib.load().mdk_rdf_bbox(_ptr(pos_soa), n_frames, layout.n_pad, _ptr(bbox), _stream()),
          "mdk_rdf_bbox")
    _count()


def rdf_hist(pos_soa: torch.Tensor, n_frames: int, layout: RdfLayout, box, cutoff: float,
             nbins: int, thr_dev: torch.Tensor, cut2: float, hist: torch.Tensor,
             work_counter: torch.Tensor, exact_div: bool = False, tuning: int = 0,
             bbox: torch.Tensor | None = None, wrapped: bool = False):
    """hist[pair][bin] += counts of all minimum-image pair distances below the cutoff.

    Replaces get_partial_triu_indices / apply_minimum_image (utils/linalg.py:84-122),
    get_dij and bin_minibatch (radial_distribution_function.py:616-689) and the batch /
    minibatch / species-pair loops (:422-524, :846-885).
    """
    _need_cuda(pos_soa, torch.float32, "rdf_hist pos")
    _need_cuda(thr_dev, torch.float32, "rdf_hist thr")
    _need_cuda(hist, torch.int64, "rdf_hist hist")
    if hist.numel() != layout.n_pairs * nbins:
        raise MdkError("rdf_hist: hist must hold n_pairs * nbins int64 values")
    if pos_soa.numel() < n_frames * 3 * layout.n_pad:
        raise MdkError("rdf_hist: position buffer too small")
    box32 = np.asarray(box, dtype=np.float32)
    flags = ((_lib.MDK_RDF_EXACT_DIV if exact_div else 0)
             | (_lib.MDK_RDF_WRAPPED if wrapped and not exact_div else 0) | int(tuning))
    check(
        _lib.load().mdk_rdf_hist(
            _ptr(pos_soa), int(n_frames), layout.n_pad,
            layout.sp_lo.ctypes.data_as(C.c_void_p), layout.sp_hi.ctypes.data_as(C.c_void_p),
            layout.n_species, box32.ctypes.data_as(C.c_void_p), C.c_float(cut2),
            C.c_float(cutoff), int(nbins), _ptr(thr_dev), _ptr(hist), _ptr(work_counter),
            _ptr(bbox) if bbox is not None else None, flags, _stream(),
        ),
        "mdk_rdf_hist",
    )
    _count()


def rdf_tie_count(pos_frame: torch.Tensor, layout: RdfLayout, n_rows: int, box, cutoff: float,
                  nbins: int, thr_dev: torch.Tensor, cut2: float, out: torch.Tensor,
                  exact_div: bool = False):
    """out[0] += in-cutoff pairs of the first n_rows atoms of one packed frame, out[1] += those
    whose reference bin differs from a plain fp32 bin (bin-edge ties; mdk.h)."""
    _need_cuda(pos_frame, torch.float32, "rdf_tie_count pos")
    _need_cuda(thr_dev, torch.float32, "rdf_tie_count thr")
    _need_cuda(out, torch.int64, "rdf_tie_count out")
    if pos_frame.numel() < 3 * layout.n_pad or out.numel() != 2 or thr_dev.numel() != nbins + 1:
        raise MdkError("rdf_tie_count: bad shapes")
    box32 = np.asarray(box, dtype=np.float32)
    check(_lib.load().mdk_rdf_tie_count(_ptr(pos_frame), layout.n_pad, int(min(n_rows, layout.n_pad)),
                                        box32.ctypes.data_as(C.c_void_p), C.c_float(cut2),
                                        C.c_float(cutoff), int(nbins), _ptr(thr_dev),
                                        _lib.MDK_RDF_EXACT_DIV if exact_div else 0, _ptr(out),
                                        _stream()), "mdk_rdf_tie_count")
    _count()


# --------------------------------------------------------------------------------------
# Angular distribution function
# --------------------------------------------------------------------------------------
def adf_workspace(n_atoms: int, n_frames: int, box, cutoff: float) -> int:
    box32 = np.asarray(box, dtype=np.float32)
    n = int(_lib.load().mdk_adf_workspace(int(n_atoms), int(n_frames),
                                          box32.ctypes.data_as(C.c_void_p), C.c_float(cutoff)))
    if n < 0:
        raise MdkError("adf_workspace: bad arguments")
    return n


def adf_hist(pos: torch.Tensor, sp_hi, box, cutoff: float, nbins: int, range_hi: float,
             norm_power: float, capacity: int, hist_w: torch.Tensor, hist_c: torch.Tensor,
             overflow: torch.Tensor, workspace: torch.Tensor):
    """hist_w / hist_c [n_combos][nbins] += weights / counts of the triplet angles of a batch of
    frames ``pos`` [F][N][3]; ``overflow`` (int32[1], zeroed by the caller) reports a
    neighbour count beyond ``capacity``.

    Replaces utils/neighbour_list.py:53-177, utils/linalg.py:30-81 and
    angular_distribution_function.py:302-403.
    """
    _need_cuda(pos, torch.float32, "adf_hist pos")
    _need_cuda(hist_w, torch.float64, "adf_hist hist_w")
    _need_cuda(hist_c, torch.int64, "adf_hist hist_c")
    _need_cuda(overflow, torch.int32, "adf_hist overflow")
    _need_cuda(workspace, torch.uint8, "adf_hist workspace")
    F, N, D = pos.shape
    if D != 3:
        raise MdkError("adf_hist: positions must be [F][N][3]")
    sp = np.asarray(sp_hi, dtype=np.int32)
    ns = len(sp)
    n_combos = ns * (ns + 1) * (ns + 2) // 6
    if hist_w.numel() != n_combos * nbins or hist_c.numel() != n_combos * nbins:
        raise MdkError("adf_hist: histograms must hold n_combos * nbins values")
    box32 = np.asarray(box, dtype=np.float32)
    check(_lib.load().mdk_adf_hist(_ptr(pos), F, N, sp.ctypes.data_as(C.c_void_p), ns,
                                   box32.ctypes.data_as(C.c_void_p), C.c_float(cutoff),
                                   int(nbins), C.c_double(range_hi), C.c_double(norm_power),
                                   int(capacity), _ptr(hist_w), _ptr(hist_c), _ptr(overflow),
                                   _ptr(workspace), workspace.numel(), _stream()),
          "mdk_adf_hist")
    _count(4)


# --------------------------------------------------------------------------------------
# Einstein MSD / Green-Kubo ACF
# --------------------------------------------------------------------------------------
def msd_windowed(traj: torch.Tensor, a_lo: int, a_hi: int, t0: int, W: int, ct: int,
                 tau_dev: torch.Tensor, span: int, msd_sum: torch.Tensor, dense: bool = False):
    """msd_sum[k] += sum over windows, atoms, dims of (x(s+tau_k) - x(s))^2.

    Replaces einstein_diffusion_coefficients.py:168-190 and the window loop :230-244.
    """
    _need_cuda(traj, torch.float32, "msd traj")
    _need_cuda(tau_dev, torch.int32, "msd tau")
    _need_cuda(msd_sum, torch.float64, "msd out")
    A, T, D = traj.shape
    if D != 3 or msd_sum.numel() != tau_dev.numel():
        raise MdkError("msd_windowed: bad shapes")
    if dense and ct == 1:
        # tau = 0 .. n-1: register-ring kernel
        check(
            _lib.load().mdk_msd_dense(_ptr(traj), A, T, a_lo, a_hi, t0, W, tau_dev.numel(),
                                      _ptr(msd_sum), _stream()),
            "mdk_msd_dense",
        )
        _count()
        return
    check(
        _lib.load().mdk_msd_windowed(_ptr(traj), A, T, a_lo, a_hi, t0, W, ct, _ptr(tau_dev),
                                     tau_dev.numel(), span, _ptr(msd_sum), _stream()),
        "mdk_msd_windowed",
    )
    _count()


def acf_windowed(traj: torch.Tensor, a_lo: int, a_hi: int, t0: int, B: int, N: int, W: int,
                 ct: int, acf_sum: torch.Tensor, acf_win: torch.Tensor | None,
                 scratch: torch.Tensor | None = None):
    """acf_sum[m] += sum_w S_w[m]; acf_win[w][m] = S_w[m] where S_w is the atom- and
    dimension-summed unbiased autocorrelation of window w (tfp.stats.auto_correlation).

    Replaces green_kubo_self_diffusion_coefficients.py:191-199 and
    green_kubo_ionic_conductivity.py:201-203 plus their window loops.
    """
    _need_cuda(traj, torch.float32, "acf traj")
    _need_cuda(acf_sum, torch.float64, "acf out")
    A, T, D = traj.shape
    if D != 3 or acf_sum.numel() != N:
        raise MdkError("acf_windowed: bad shapes")
    if acf_win is not None:
        _need_cuda(acf_win, torch.float64, "acf per-window out")
        if acf_win.numel() != W * N:
            raise MdkError("acf_windowed: acf_win must be [W][N]")
    if scratch is None or scratch.numel() < B * N:
        scratch = torch.empty(B * N, dtype=torch.float64, device=traj.device)
    P = scratch[: B * N]
    P.zero_()
    acf_accumulate(traj, a_lo, a_hi, t0, B, N, P)
    acf_finish(P, B, N, W, ct, acf_sum, acf_win)
    return scratch


def acf_accumulate(traj: torch.Tensor, a_lo: int, a_hi: int, t0: int, B: int, N: int,
                   P: torch.Tensor):
    """P[t][m] += sum over the atoms [a_lo, a_hi) and dimensions of v(t) v(t + m): the lag
    products of one frame batch.  Additive over atom ranges, so a caller may feed the rows in
    blocks (as they arrive on the device) before ``acf_finish``."""
    _need_cuda(traj, torch.float32, "acf traj")
    _need_cuda(P, torch.float64, "acf lag products")
    A, T, D = traj.shape
    if D != 3 or P.numel() < B * N:
        raise MdkError("acf_accumulate: bad shapes")
    check(_lib.load().mdk_acf_lagprod(_ptr(traj), A, T, a_lo, a_hi, t0, B, N, _ptr(P), _stream()),
          "mdk_acf_lagprod")
    _count(1)


def acf_finish(P: torch.Tensor, B: int, N: int, W: int, ct: int, acf_sum: torch.Tensor,
               acf_win: torch.Tensor | None):
    """Prefix sums of P along t (in place), then acf_sum[m] += sum_w S_w[m] and
    acf_win[w][m] = S_w[m]."""
    _need_cuda(acf_sum, torch.float64, "acf out")
    check(
        _lib.load().mdk_acf_windows(_ptr(P), B, N, W, ct, _ptr(acf_sum),
                                    _ptr(acf_win) if acf_win is not None else None, _stream()),
        "mdk_acf_windows",
    )
    _count(2)


# --------------------------------------------------------------------------------------
# Transformations
# --------------------------------------------------------------------------------------
def unwrap(pos: torch.Tensor, box, carry_pos: torch.Tensor | None, carry_img: torch.Tensor,
           have_carry: bool, out: torch.Tensor):
    """Box-jump unwrapping along time with carry-over between batches.

    Replaces transformations/unwrap_coordinates.py:51-81.
    """
    _need_cuda(pos, torch.float32, "unwrap pos")
    _need_cuda(out, torch.float32, "unwrap out")
    _need_cuda(carry_img, torch.float64, "unwrap carry_img")
    A, T, D = pos.shape
    if D != 3 or out.shape != pos.shape or carry_img.numel() != A * 3:
        raise MdkError("unwrap: bad shapes")
    if carry_pos is not None:
        _need_cuda(carry_pos, torch.float32, "unwrap carry_pos")
    box64 = np.asarray(box, dtype=np.float64)
    check(
        _lib.load().mdk_unwrap(_ptr(pos), A, T, box64.ctypes.data_as(C.c_void_p),
                               _ptr(carry_pos) if carry_pos is not None else None,
                               int(have_carry), _ptr(carry_img), _ptr(out), _stream()),
        "mdk_unwrap",
    )
    _count()


def unwrap_indices(pos: torch.Tensor, img: torch.Tensor, box, out: torch.Tensor):
    """out = pos + img * L.  Replaces transformations/unwrap_via_indices.py:49-57."""
    _need_cuda(pos, torch.float32, "unwrap_indices pos")
    _need_cuda(img, torch.float32, "unwrap_indices img")
    _need_cuda(out, torch.float32, "unwrap_indices out")
    if pos.shape != img.shape or pos.shape != out.shape or pos.shape[-1] != 3:
        raise MdkError("unwrap_indices: bad shapes")
    box64 = np.asarray(box, dtype=np.float64)
    check(
        _lib.load().mdk_unwrap_indices(_ptr(pos), _ptr(img), pos.numel() // 3,
                                       box64.ctypes.data_as(C.c_void_p), _ptr(out), _stream()),
        "mdk_unwrap_indices",
    )
    _count()


def velocity_from_positions(pos: torch.Tensor, dt: float, out: torch.Tensor):
    """out[a][t] = (pos[a][t + 1] - pos[a][t]) / dt (fp32), last frame repeated.
    Replaces transformations/velocity_from_positions.py:62-77."""
    _need_cuda(pos, torch.float32, "velocity_from_positions pos")
    _need_cuda(out, torch.float32, "velocity_from_positions out")
    if pos.shape != out.shape or pos.shape[-1] != 3:
        raise MdkError("velocity_from_positions: bad shapes")
    A, T, _ = pos.shape
    check(_lib.load().mdk_velocity_from_positions(_ptr(pos), A, T, C.c_float(dt), _ptr(out),
                                                  _stream()), "mdk_velocity_from_positions")
    _count()


def ionic_current(vel: torch.Tensor, charge, J: torch.Tensor):
    """J[t][d] += sum_a q_a v[a][t][d].  charge: python float, [A] tensor or [A][T] tensor.

    Replaces transformations/ionic_current.py:48-58.
    """
    _need_cuda(vel, torch.float32, "ionic_current vel")
    _need_cuda(J, torch.float64, "ionic_current J")
    A, T, D = vel.shape
    if D != 3 or J.numel() != T * 3:
        raise MdkError("ionic_current: bad shapes")
    lib = _lib.load()
    if isinstance(charge, torch.Tensor):
        _need_cuda(charge, torch.float32, "ionic_current charge")
        if charge.numel() == A:
            mode = 1
        elif charge.numel() == A * T:
            mode = 2
        else:
            raise MdkError("ionic_current: charge must have A or A*T elements")
        check(lib.mdk_ionic_current(_ptr(vel), A, T, _ptr(charge), mode, _ptr(J), _stream()),
              "mdk_ionic_current")
    else:
        q = C.c_double(float(charge))
        check(lib.mdk_ionic_current(_ptr(vel), A, T, C.cast(C.byref(q), C.c_void_p), 0, _ptr(J),
                                    _stream()),
              "mdk_ionic_current")
    _count()


def flux_sum(x: torch.Tensor, J: torch.Tensor, comp0: int = 0, w1: torch.Tensor | None = None,
             w2: torch.Tensor | None = None):
    """J[t][k] += sum_a w(a, t) x[a][t][comp0 + k] (k < 3), w = 1 or w1 (+ w2), fp64 sums.

    Replaces transformations/momentum_flux.py:45-55 and integrated_heat_current.py:49-60.
    """
    _need_cuda(x, torch.float32, "flux_sum x")
    _need_cuda(J, torch.float64, "flux_sum J")
    A, T, ncomp = x.shape
    if J.numel() != T * 3:
        raise MdkError("flux_sum: J must hold T * 3 values")
    for w in (w1, w2):
        if w is not None:
            _need_cuda(w, torch.float32, "flux_sum weight")
            if w.numel() != A * T:
                raise MdkError("flux_sum: weights must have A * T elements")
    check(_lib.load().mdk_flux_sum(_ptr(x), A, T, int(ncomp), int(comp0),
                                   _ptr(w1) if w1 is not None else None,
                                   _ptr(w2) if w2 is not None else None, _ptr(J), _stream()),
          "mdk_flux_sum")
    _count()


def thermal_flux(stress: torch.Tensor, vel: torch.Tensor, ke: torch.Tensor, pe: torch.Tensor,
                 J: torch.Tensor):
    """J[t][k] += sum_a (KE + PE) v_k - (S v)_k.  Replaces transformations/thermal_flux.py:51-92."""
    _need_cuda(stress, torch.float32, "thermal_flux stress")
    _need_cuda(vel, torch.float32, "thermal_flux vel")
    _need_cuda(ke, torch.float32, "thermal_flux ke")
    _need_cuda(pe, torch.float32, "thermal_flux pe")
    _need_cuda(J, torch.float64, "thermal_flux J")
    A, T, D = vel.shape
    if D != 3 or tuple(stress.shape) != (A, T, 6) or ke.numel() != A * T or pe.numel() != A * T \
            or J.numel() != T * 3:
        raise MdkError("thermal_flux: bad shapes")
    check(_lib.load().mdk_thermal_flux(_ptr(stress), _ptr(vel), _ptr(ke), _ptr(pe), A, T, _ptr(J),
                                       _stream()), "mdk_thermal_flux")
    _count()
