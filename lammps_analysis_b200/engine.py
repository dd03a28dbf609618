"""Batch drivers between the calculators and the kernels: what is resident in HBM, in which
layout, and which kernel runs over it.

* RDF: per-species atom-major stores ``[A_s][T][3]`` fp32 -> frame batches packed into the
  frame-major SoA layout ``[F_b][3][n_pad]`` (species blocks padded to the column tile with
  NaN) -> one ``mdk_rdf_hist`` launch per frame batch -> u64 histograms stay on the device
  until the last batch.
* MSD / ACF: the atom-major store is already the layout the kernels stream (time contiguous
  per atom); the reference batch plan (which windows exist) is applied as
  launch parameters, not as host loops over windows.

No function here falls back to the CPU.
"""
from __future__ import annotations

import numpy as np
import torch

from . import kernels as K
from . import trace
from ._lib import MdkError


def _device(device=None):
    if not torch.cuda.is_available():
        raise MdkError("CUDA device required: lammps_analysis_b200 has no CPU path")
    return torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")


def to_device_f32(x, device=None) -> torch.Tensor:
    """Host array / tensor -> contiguous float32 CUDA tensor (pinned staging for numpy)."""
    dev = _device(device)
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float32).contiguous()
    a = np.ascontiguousarray(x, dtype=np.float32)
    t = torch.from_numpy(a)
    return t.to(dev, non_blocking=False)


class RdfEngine:
    """All-pairs histogram over sampled frames for every species pair.

    Restates RadialDistributionFunction.run_calculator's data movement
    (radial_distribution_function.py:828-887): the reference's frame batches, atom
    minibatches and per-species-pair masked passes only partition an integer sum, so the
    counts are independent of that plan and one fused pass per HBM-sized
    frame batch yields the same integers.
    """

    # species at least this large are Morton-ordered per frame so that whole blocks of pairs
    # beyond the cutoff can be skipped; below it a tile spans too much of the box to gain
    TIE_ROWS = 1024
    TIE_PAIRS = 250_000_000
    SORT_BATCH_FRAMES = 32
    SORT_MIN_ATOMS = 40_000    # measured on B200 (round 2: 32-atom boxes, one sort per batch):
                               # -2 % at 30k atoms, +4 % at 50k, +10 % at 65k, +18 % at 100k,
                               # +60 % at 10^6

    def __init__(self, counts, box, cutoff: float, nbins: int, drop_first: bool = True,
                 device=None, max_batch_bytes: int = 2 << 30, spatial_sort=None):
        self.device = _device(device)
        self.full_counts = [int(c) for c in counts]
        # Q1: the reference's strict index masks drop the first atom of every species
        # (radial_distribution_function.py:637-638)
        self.atom_first = 1 if drop_first else 0
        eff = [max(c - self.atom_first, 0) for c in self.full_counts]
        self.layout = K.rdf_layout(eff)
        self.eff_counts = eff
        self.box = np.asarray(box, dtype=np.float32)
        self.cutoff = float(np.float32(cutoff))
        self.nbins = int(nbins)
        thr, self.cut2 = K.rdf_thresholds(self.cutoff, self.nbins)
        self.thr = torch.from_numpy(thr).to(self.device)
        self.hist = torch.zeros(self.layout.n_pairs * self.nbins, dtype=torch.int64,
                                device=self.device)
        self.counter = torch.zeros(2, dtype=torch.int64, device=self.device)
        self.frame_bytes = 3 * self.layout.n_pad * 4
        self.max_frames = max(1, int(max_batch_bytes // self.frame_bytes))
        self._buf = None
        self.exact_div = bool(self.cutoff >= float(self.box.min()) / 2)
        self.frames_done = 0
        if spatial_sort is None:
            spatial_sort = max(eff, default=0) >= self.SORT_MIN_ATOMS
        self.spatial_sort = bool(spatial_sort) and not self.exact_div
        self._work = None
        self._bbox = None
        self.record_events = False      # bench.py: CUDA events around every mdk_rdf_hist launch
        self.kernel_events = []
        # bin-edge tie census on the first packed frame (rows sampled: the first TIE_ROWS atoms,
        # fewer when that would be more than TIE_PAIRS pair distances -- 1024 rows of a 10^6-atom
        # frame cost 5 ms, 2 % of a one-frame RDF call --
        # against all atoms); 0 disables it
        self.tie_rows = int(min(self.TIE_ROWS,
                                max(64, self.TIE_PAIRS // max(self.layout.n_pad, 1))))
        self.tie_counts = torch.zeros(2, dtype=torch.int64, device=self.device)
        self._tie_done = False

    # pair-distance evaluations per frame (all i<j pairs of the full system)
    def pairs_per_frame(self) -> int:
        n = sum(self.full_counts)
        return n * (n - 1) // 2

    def _buffer(self, n_frames):
        need = n_frames * 3 * self.layout.n_pad
        if self._buf is None or self._buf.numel() < need:
            self._buf = torch.empty(need, dtype=torch.float32, device=self.device)
        return self._buf

    def add_frames(self, species_traj, frames, check_extent: bool = True, tuning: int = 0):
        """species_traj: list of CUDA float32 [A_s][T][3]; frames: int array of frame ids."""
        frames = np.asarray(frames, dtype=np.int64)
        # the sorted pack is one host call per frame and species: short launch batches let the
        # host enqueue the packs of the next batch while the pair kernel of this one runs
        step = min(self.max_frames, self.SORT_BATCH_FRAMES) if self.spatial_sort \
            else self.max_frames
        for k0 in range(0, len(frames), step):
            sel = frames[k0:k0 + step]
            buf = self._buffer(len(sel))
            for s, traj in enumerate(species_traj):
                if traj.shape[0] != self.full_counts[s]:
                    raise MdkError("RdfEngine: species array does not match the declared count")
            bbox = None
            if self.spatial_sort:
                self.pack_sorted(species_traj, sel, buf)
                bbox = self.boxes(buf, len(sel))
            else:
                fdev = torch.from_numpy(sel.astype(np.int32)).to(self.device)
                for s, traj in enumerate(species_traj):
                    K.rdf_pack(traj, fdev, buf, self.layout, s, self.atom_first,
                               self.eff_counts[s])
            self.add_packed(buf, len(sel), check_extent=check_extent, tuning=tuning, bbox=bbox)

    def pack_sorted(self, species_traj, frames, buf):
        """Hilbert-ordered pack of the given frames: one sort per species for the whole batch
        when it is small enough (per-frame sorts of ~10^5 atoms are launch-bound), otherwise one
        sort per frame and species."""
        n_max = max(self.eff_counts)
        need = K.rdf_sort_batch_workspace(n_max, len(frames)) if len(frames) > 1 else -1
        if need >= 0:
            if self._work is None or self._work.numel() < need:
                self._work = torch.empty(need, dtype=torch.uint8, device=self.device)
            fdev = torch.from_numpy(np.asarray(frames, dtype=np.int32)).to(self.device)
            for s, traj in enumerate(species_traj):
                K.rdf_pack_sorted_batch(traj, fdev, buf, self.layout, s, self.atom_first,
                                        self.eff_counts[s], self.box, self._work)
            return
        nbytes = K.rdf_sort_workspace(n_max)
        if self._work is None or self._work.numel() < nbytes:
            self._work = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        for k, f in enumerate(frames):
            for s, traj in enumerate(species_traj):
                K.rdf_pack_sorted(traj, int(f), buf, k, self.layout, s, self.atom_first,
                                  self.eff_counts[s], self.box, self._work)

    def boxes(self, buf, n_frames):
        need = n_frames * (self.layout.n_pad // K.RDF_SUBTILE) * 6
        if self._bbox is None or self._bbox.numel() < need:
            self._bbox = torch.empty(need, dtype=torch.float32, device=self.device)
        K.rdf_bbox(buf, n_frames, self.layout, self._bbox)
        return self._bbox

    def add_packed(self, pos_soa, n_frames, check_extent: bool = True, tuning: int = 0,
                   bbox=None):
        exact = self.exact_div
        wrapped = False
        if check_extent and not exact:
            mm = K.coord_extent(pos_soa, n_frames, self.layout.n_pad)
            span = mm[3:] - mm[:3]
            # the fast minimum image (r*invL, fma) is bit-identical to the reference only
            # while |rint(r/L)| <= 2
            if np.any(span >= 2.5 * self.box):
                exact = True
            # coordinates inside one box length: min(|d|, L - |d|) is the minimum image
            wrapped = bool(np.all(span < self.box))
        if self.tie_rows and not self._tie_done and n_frames:
            K.rdf_tie_count(pos_soa, self.layout, self.tie_rows, self.box, self.cutoff,
                            self.nbins, self.thr, self.cut2, self.tie_counts, exact_div=exact)
            self._tie_done = True
        if self.record_events:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        K.rdf_hist(pos_soa, n_frames, self.layout, self.box, self.cutoff, self.nbins, self.thr,
                   self.cut2, self.hist, self.counter, exact_div=exact, tuning=tuning,
                   bbox=None if exact else bbox, wrapped=wrapped and not exact)
        if self.record_events:
            e1.record()
            self.kernel_events.append((e0, e1))
        self.frames_done += n_frames

    def tie_report(self) -> dict:
        """{"pairs_checked", "ties", ...}: see mdk_rdf_tie_count (synchronises)."""
        checked, ties = (int(v) for v in self.tie_counts.cpu().tolist())
        return {"pairs_checked": checked, "ties": ties,
                "sample": f"first packed frame, pairs (i, j > i) of the first {self.tie_rows} "
                          "atoms against all atoms, inside the cutoff",
                "definition": "pairs whose bin under the reference rule (double-step "
                              "histogram_fixed_width on the correctly rounded fp32 distance; "
                              "what the kernel counts) differs from plain fp32 binning "
                              "floor(sqrt(d2) * float(nbins / cutoff))"}

    def counts(self) -> np.ndarray:
        """int64 [n_pairs][nbins] (device -> host, synchronises)."""
        return self.hist.view(self.layout.n_pairs, self.nbins).cpu().numpy()


class AdfEngine:
    """Triplet-angle histograms per frame batch for every species triple (centre, j, k) with
    a <= b <= c.  Restates AngularDistributionFunction._build_histograms
    (angular_distribution_function.py:365-403) over the cell-list kernels of csrc/adf.cu: one
    ``add_batch`` per reference batch returns that batch's un-normalised weight sums and triple
    counts (the reference density-normalises every batch on its own before summing)."""

    BIN_RANGE_HI = 3.15          # "a chemist's pi" (:192)

    def __init__(self, counts, box, cutoff: float, nbins: int, norm_power, device=None,
                 capacity: int = None):
        self.device = _device(device)
        self.counts = [int(c) for c in counts]
        self.sp_hi = np.cumsum(self.counts).astype(np.int32)
        self.n_atoms = int(self.sp_hi[-1]) if len(self.counts) else 0
        self.box = np.asarray(box, dtype=np.float32)
        self.cutoff = float(cutoff)
        self.nbins = int(nbins)
        self.norm_power = float(norm_power)
        ns = len(self.counts)
        self.n_combos = ns * (ns + 1) * (ns + 2) // 6
        if capacity is None:
            # three times the mean neighbour count of a uniform system, at least 64
            rho = self.n_atoms / float(np.prod(self.box.astype(np.float64)))
            capacity = max(64, int(3 * rho * 4.19 * self.cutoff**3))
        self.capacity = int(min(max(capacity, 1), 8192))
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._work = None
        self.max_batch_bytes = 1 << 30

    def add_batch(self, pos: torch.Tensor):
        """pos: CUDA float32 [F][N][3] (species concatenated).  Returns (weights float64
        [n_combos][nbins], counts int64 [n_combos][nbins]) of this batch, on the device."""
        F, N, _ = pos.shape
        if N != self.n_atoms:
            raise MdkError("AdfEngine: positions do not match the declared species counts")
        hw = torch.zeros(self.n_combos * self.nbins, dtype=torch.float64, device=self.device)
        hc = torch.zeros(self.n_combos * self.nbins, dtype=torch.int64, device=self.device)
        step = max(1, int(self.max_batch_bytes // max(N * 16, 1)))
        for f0 in range(0, F, step):
            chunk = pos[f0:f0 + step].contiguous()
            while True:
                need = K.adf_workspace(N, chunk.shape[0], self.box, self.cutoff)
                if self._work is None or self._work.numel() < need:
                    self._work = torch.empty(need, dtype=torch.uint8, device=self.device)
                w, c = torch.zeros_like(hw), torch.zeros_like(hc)
                self.overflow.zero_()
                K.adf_hist(chunk, self.sp_hi, self.box, self.cutoff, self.nbins,
                           self.BIN_RANGE_HI, self.norm_power, self.capacity, w, c,
                           self.overflow, self._work)
                over = int(self.overflow.item())
                if over == 0:
                    hw += w
                    hc += c
                    break
                if over > 8192:
                    raise MdkError(f"AdfEngine: {over} neighbours within the cutoff of one atom "
                                   "exceed the supported 8192")
                self.capacity = min(8192, max(over, 2 * self.capacity))   # redo this chunk
        return hw.view(self.n_combos, self.nbins), hc.view(self.n_combos, self.nbins)


def plan_windows(plan: dict, data_range: int, correlation_time: int, n_atoms: int):
    """Expand a reference batch plan (oracle-free restatement of data_manager.py:156-339)
    into launch descriptors (a_lo, a_hi, t0, B, W)."""
    out = []
    bs, nb, rem = plan["batch_size"], plan["n_batches"], plan["remainder"]
    if not plan["minibatch"]:
        sizes = [bs] * nb + ([rem] if rem > 0 else [])
        for b, size in enumerate(sizes):
            if size < data_range:
                continue  # short remainder windows are filtered by shape (:240-241)
            W = int(np.clip((size - data_range) / correlation_time, 1, None))
            out.append((0, n_atoms, b * bs, size, W))
    else:
        ab = plan["atom_batch_size"]
        nab = plan["n_atom_batches"]
        if plan["atom_remainder"]:
            raise MdkError("atom mini-batch plans with an atom remainder are not reproducible "
                           "(data_manager.py:266-267); choose a divisible atom count")
        for a in range(nab):
            a_lo, a_hi = int(a * ab), int(a * ab + ab)
            for b in range(nb):
                if bs < data_range:
                    continue
                W = int(np.clip((bs - data_range) / correlation_time, 1, None))
                out.append((a_lo, a_hi, b * bs, bs, W))
    return out


def _local_rows(a_lo, a_hi, a_shard, row_offset, n_rows):
    lo, hi = a_lo, a_hi
    if a_shard is not None:
        lo, hi = max(lo, a_shard[0]), min(hi, a_shard[1])
    lo, hi = lo - row_offset, hi - row_offset
    if hi > lo and (lo < 0 or hi > n_rows):
        raise MdkError("atom range outside the rows resident on this device")
    return lo, hi


def _row_blocks(lo, hi, blocks):
    """[lo, hi) cut along the row blocks a device copy arrives in: yields (l0, l1, event)."""
    if not blocks:
        yield lo, hi, None
        return
    for b0, b1, ev in blocks:
        l0, l1 = max(lo, b0), min(hi, b1)
        if l1 > l0:
            yield l0, l1, ev


def _merge_blocks(blocks, max_launches: int):
    """Coarsen a block list to at most ``max_launches`` runs of consecutive blocks (each run
    waits for its last event): for kernels whose per-launch cost is not negligible."""
    if not blocks or len(blocks) <= max_launches:
        return blocks
    per = -(-len(blocks) // max_launches)
    out = []
    for i in range(0, len(blocks), per):
        run = blocks[i:i + per]
        out.append((run[0][0], run[-1][1], run[-1][2]))
    return out


def msd_series(traj: torch.Tensor, launches, data_range: int, correlation_time: int,
               tau_values, a_shard=None, row_offset: int = 0, blocks=None):
    """Returns (msd_sum device float64 [n_tau], count).

    ``launches`` are in global atom indices; ``traj`` holds the global rows
    [row_offset, row_offset + traj.shape[0]); ``a_shard`` = (lo, hi) restricts the atoms this
    rank processes (multi-GPU atom sharding).  count is always the full-plan count, computed
    analytically (einstein_diffusion_coefficients.py:184, :244).

    ``blocks`` = [(r0, r1, event)] (rows of ``traj``; store.device_blocks): the kernel runs per
    row block as soon as the block's event has fired -- the sums are additive over atoms -- so
    that it overlaps with the copy / transformation still producing the later blocks."""
    tau_host = np.asarray(tau_values, dtype=np.int32)
    dense = bool(np.array_equal(tau_host, np.arange(len(tau_host))))
    tau = torch.as_tensor(tau_host, device=traj.device)
    out = torch.zeros(len(tau_values), dtype=torch.float64, device=traj.device)
    count = 0
    stream = torch.cuda.current_stream()
    for a_lo, a_hi, t0, B, W in launches:
        count += W * ((a_hi - a_lo) + 1)
        lo, hi = _local_rows(a_lo, a_hi, a_shard, row_offset, traj.shape[0])
        if hi > lo:
            for l0, l1, ev in _row_blocks(lo, hi, blocks):
                if ev is not None:
                    stream.wait_event(ev)
                K.msd_windowed(traj, l0, l1, t0, W, correlation_time, tau, data_range, out,
                               dense=dense)
                trace.event(f"msd rows [{l0}, {l1}) done")
    return out, count


def acf_series(traj: torch.Tensor, launches, data_range: int, correlation_time: int,
               per_window: bool = True, a_shard=None, row_offset: int = 0, blocks=None):
    """Returns (acf_sum device [N], count, [acf_win device [W][N] per launch], [A_sel]).

    count follows green_kubo_self_diffusion_coefficients.py:196, :334 (A + 1 per window).
    ``blocks``: as in msd_series (the lag products are additive over atoms; prefix sums and
    window sums run once per launch after the last block).
    """
    N = data_range
    out = torch.zeros(N, dtype=torch.float64, device=traj.device)
    wins, sizes = [], []
    count = 0
    scratch = None
    stream = torch.cuda.current_stream()
    # every lag-product launch ends with one fp64 atomicAdd per tile element and atom group
    # (tens of millions per launch): a few launches per dataset, not one per upload block
    blocks = _merge_blocks(blocks, 4)
    for a_lo, a_hi, t0, B, W in launches:
        count += W * ((a_hi - a_lo) + 1)
        lo, hi = _local_rows(a_lo, a_hi, a_shard, row_offset, traj.shape[0])
        win = torch.zeros(W, N, dtype=torch.float64, device=traj.device) if per_window else None
        if hi > lo:
            if scratch is None or scratch.numel() < B * N:
                scratch = torch.empty(B * N, dtype=torch.float64, device=traj.device)
            P = scratch[: B * N]
            P.zero_()
            for l0, l1, ev in _row_blocks(lo, hi, blocks):
                if ev is not None:
                    stream.wait_event(ev)
                K.acf_accumulate(traj, l0, l1, t0, B, N, P)
                trace.event(f"acf rows [{l0}, {l1}) done")
            K.acf_finish(P, B, N, W, correlation_time, out, win)
        wins.append(win)
        sizes.append(a_hi - a_lo)
    return out, count, wins, sizes
