"""
Oracle restatement of the AngularDistributionFunction calculator (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module;
the product never does.

Follows:
  mdsuite/calculators/angular_distribution_function.py:229-268 (_prepare_data_structure),
      :302-328 (_compute_rijk_matrices), :330-363 (_compute_angles: species masks),
      :365-403 (_build_histograms), :405-444 (_compute_adfs), :506-527 (_correct_batch_properties)
  mdsuite/utils/neighbour_list.py:37-50 (get_triu_indicies), :53-112 (get_neighbour_list),
      :116-177 (get_triplets: float16 cutoff test, roll-by-shift enumeration)
  mdsuite/utils/linalg.py:30-48 (unit_vector, angle_between), :51-81 (get_angles)
  numpy.histogram(weights=..., density=True) as called at :388-394

Third-party numerics restated (parity for these is pinned to the restatement):
tf.norm (sqrt of the left-to-right fp32 sum of squares), tf.einsum("ij,ij->i") (products
rounded, summed left to right), tf.math.acos (numpy float32 arccos), the float16 cast
(round to nearest even).
"""
from __future__ import annotations

import itertools

import numpy as np

from oracle.planner import MemoryManager

F32 = np.float32
BIN_RANGE = (0.0, 3.15)   # :192 "from 0 to a chemists pi"


# --- neighbour_list.py:37-50 -------------------------------------------------
def get_triu_indices(n_atoms: int) -> np.ndarray:
    """tf.where(~band_part(ones, -1, 0)): (row, col) with col > row, row-major."""
    r, c = np.triu_indices(n_atoms, k=1)
    return np.stack([r, c]).astype(np.int32)


# --- neighbour_list.py:53-112 -------------------------------------------------
def get_neighbour_list(positions: np.ndarray, cell) -> np.ndarray:
    """positions (T, n, 3) float32 -> r_ij_flat (T, n(n-1)/2, 3): p_row - p_col for row < col,
    minimum image when a cell is given (tf converts the cell list to float32)."""
    positions = np.asarray(positions, dtype=F32)
    triu = get_triu_indices(positions.shape[1])
    r = positions[:, triu[0]] - positions[:, triu[1]]
    if cell:
        cell32 = np.asarray(cell, dtype=F32)
        r = r - np.rint(r / cell32) * cell32
    return r


# --- angular_distribution_function.py:302-328 ---------------------------------
def rij_matrix(positions: np.ndarray, cell) -> np.ndarray:
    """(T, n, n, 3): scatter of the upper triangle minus its transpose."""
    T, n, _ = positions.shape
    flat = get_neighbour_list(positions, cell)
    triu = get_triu_indices(n)
    mat = np.zeros((T, n, n, 3), dtype=F32)
    mat[:, triu[0], triu[1]] = flat
    return mat - np.transpose(mat, (0, 2, 1, 3))


def _norm(v: np.ndarray) -> np.ndarray:
    sq = v * v
    return np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])


# --- neighbour_list.py:116-177 -------------------------------------------------
def get_triplets(full_r_ij: np.ndarray, r_cut: float, n_atoms: int, n_batches=200) -> np.ndarray:
    """(n_triples, 4) int64 rows (t, i, j, k): j != k, both within r_cut of i -- compared in
    FLOAT16, zero distances (the diagonal, coincident atoms) excluded.  Order: shift batches,
    then tf.where's row-major (t, shift, i, j)."""
    if n_batches >= n_atoms:
        n_batches = n_atoms - 1
    r_ij = _norm(np.asarray(full_r_ij, dtype=F32)).astype(np.float16)
    rc16 = np.float16(r_cut)
    r_ij = np.where(r_ij == 0, rc16, r_ij)
    near = r_ij < rc16
    triples = []
    if n_atoms < 2:
        return np.zeros((0, 4), dtype=np.int64)
    for batch in np.array_split(np.arange(1, n_atoms), n_batches):
        rolled = np.stack([np.roll(near, -int(a), axis=2) for a in batch], axis=1)  # (t, n, i, j)
        t, n, i, j = np.nonzero(near[:, None] & rolled)
        k = j + n + int(batch[0])
        k = np.where(k >= n_atoms, k - n_atoms, k)
        triples.append(np.stack([t, i, j, k], axis=1))
    return np.concatenate(triples, axis=0).astype(np.int64)


# --- linalg.py:30-81 -------------------------------------------------------------
def get_angles(r_ij_mat: np.ndarray, indices: np.ndarray):
    """Returns (angles float32, |r_ij| * |r_ik| float32) for the triples (t, i, j, k)."""
    r_ij = r_ij_mat[indices[:, 0], indices[:, 1], indices[:, 2]]
    r_ik = r_ij_mat[indices[:, 0], indices[:, 1], indices[:, 3]]
    n_ij, n_ik = _norm(r_ij), _norm(r_ik)
    u1 = r_ij / n_ij[:, None]
    u2 = r_ik / n_ik[:, None]
    pr = u1 * u2
    cos = (pr[:, 0] + pr[:, 1]) + pr[:, 2]
    return np.arccos(np.clip(cos, F32(-1.0), F32(1.0))).astype(F32), n_ij * n_ik


# --- angular_distribution_function.py:229-268 -------------------------------------
def species_indices(species: list, n_particles: dict):
    out, start = [], 0
    for sp in species:
        stop = start + int(n_particles[sp])
        out.append((sp, start, stop))
        start = stop
    return out


# --- _prepare_managers + _correct_batch_properties (:506-527) ----------------------
def adf_plan(species_shapes: dict, n_configs_total: int, number_of_configurations: int,
             memory: float, memory_fraction: float = 0.5, override_n_batches=None) -> int:
    """Number of batches the sampled configurations are split into (each batch is
    density-normalised on its own, so this changes the result)."""

    class _DB:
        def get_data_size(self, item):
            n = species_shapes[item]
            return n, n_configs_total, n * n_configs_total * 3 * 4

    mm = MemoryManager(data_path=list(species_shapes.keys()), database=_DB(),
                       memory_fraction=memory_fraction,
                       scale_function={"quadratic": {"outer_scale_factor": 10}}, memory=memory)
    batch_size, n_batches, _ = mm.get_batch_size()
    _, minibatch = mm.get_ensemble_loop(1, 1)
    if batch_size > number_of_configurations:
        n_batches = 1
    else:
        n_batches = int(number_of_configurations / batch_size)
    if override_n_batches is not None:
        n_batches = override_n_batches
    if minibatch:
        n_batches = number_of_configurations
    return n_batches


# --- _build_histograms (:365-403) + run_calculator (:584-609) -----------------------
def adf_histograms(positions_by_species: dict, species: list, box_array, sample_frames,
                   cutoff: float, number_of_bins: int, norm_power, n_batches: int,
                   return_counts: bool = False) -> dict:
    """{"A-B-C": float32 [nbins]} summed over the batches; with ``return_counts`` also the
    un-normalised per-batch (weight sums float64, triple counts int64) per key."""
    n_particles = {s: positions_by_species[s].shape[0] for s in species}
    sp_idx = species_indices(species, n_particles)
    n_atoms = sum(n_particles.values())
    cell = [float(b) for b in box_array]
    angles, raw = {}, {}
    for frames in np.array_split(np.asarray(sample_frames), n_batches):
        pos = np.concatenate([np.asarray(positions_by_species[s], dtype=F32)[:, frames]
                              for s in species], axis=0)
        tmp = np.transpose(pos, (1, 0, 2))                      # (timesteps, atoms, 3)
        r_ij_mat = rij_matrix(tmp, cell)
        trip = get_triplets(r_ij_mat, cutoff, n_atoms, n_batches=n_atoms)
        for combo in itertools.combinations_with_replacement(sp_idx, 3):
            (i_n, i0, i1), (j_n, j0, j1), (k_n, k0, k1) = combo
            name = f"{i_n}-{j_n}-{k_n}"
            cond = ((trip[:, 1] >= i0) & (trip[:, 1] < i1) & (trip[:, 2] >= j0)
                    & (trip[:, 2] < j1) & (trip[:, 3] >= k0) & (trip[:, 3] < k1))
            sel = trip[cond]
            angle_vals, pre = get_angles(r_ij_mat, sel)
            with np.errstate(divide="ignore"):
                weights = (F32(1) / pre**norm_power).astype(F32)
            with np.errstate(invalid="ignore", divide="ignore"):
                hist, _ = np.histogram(angle_vals, bins=number_of_bins, range=list(BIN_RANGE),
                                       weights=weights, density=True)
            hist = hist.astype(F32)
            angles[name] = angles[name] + hist if name in angles else hist
            if return_counts:
                w, _ = np.histogram(angle_vals, bins=number_of_bins, range=list(BIN_RANGE),
                                    weights=weights.astype(np.float64))
                c, _ = np.histogram(angle_vals, bins=number_of_bins, range=list(BIN_RANGE))
                raw.setdefault(name, []).append((w, c.astype(np.int64)))
    return (angles, raw) if return_counts else angles


# --- _compute_adfs (:405-444) ----------------------------------------------------------
def adf_finish(angles: dict, number_of_bins: int) -> dict:
    out = {}
    axis = np.linspace(BIN_RANGE[0] * (180 / 3.14159), BIN_RANGE[1] * (180 / 3.14159),
                       number_of_bins)
    for name, hist in angles.items():
        out[name.replace("-", "_")] = {"max_peak": axis[int(np.argmax(hist))],
                                       "angle": axis.tolist(), "adf": hist.tolist()}
    return out
