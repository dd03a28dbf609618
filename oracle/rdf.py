"""
Oracle restatement of the reference RDF path (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows:
  mdsuite/calculators/radial_distribution_function.py:215-279 (check_input, sampling)
  mdsuite/calculators/radial_distribution_function.py:299-382 (prefactor, g(r))
  mdsuite/calculators/radial_distribution_function.py:395-420 (_correct_batch_properties)
  mdsuite/calculators/radial_distribution_function.py:422-524 (minibatch loop, species masks)
  mdsuite/calculators/radial_distribution_function.py:616-689 (bin_minibatch, get_dij)
  mdsuite/calculators/radial_distribution_function.py:719-826 (ideal_correction)
  mdsuite/calculators/radial_distribution_function.py:828-887 (run_calculator)
  mdsuite/utils/linalg.py:84-136 (minimum image, triu indices, cutoff)
  mdsuite/utils/meta_functions.py:468-490 (split_array)

TF op semantics restated (tensorflow is an unpinned third-party dependency):
tf.gather / tf.boolean_mask -> numpy fancy indexing; tf.math.rint -> np.rint
(half to even); tf.linalg.norm -> sqrt(reduce_sum(x*x)) with products rounded to
fp32 before the adds, summed left to right; tf.histogram_fixed_width -> the CPU
functor of tensorflow/core/kernels/histogram_op.cc (double step, truncation).
"""
from __future__ import annotations

import itertools

import numpy as np

from oracle.planner import MemoryManager

F32 = np.float32


# --- utils/linalg.py:102-122 -------------------------------------------------
def get_partial_triu_indices(n_atoms: int, m_atoms: int, idx: int) -> np.ndarray:
    """~band_part(ones(m, n), -1, idx) -> tf.where -> int32 (2, x).

    band_part keeps (row, col) with col - row <= idx, the negation keeps
    col > row + idx, i.e. global j > global i for a minibatch starting at idx.
    tf.where returns row-major order, as np.nonzero does.
    """
    rows = np.arange(m_atoms, dtype=np.int64)[:, None]
    cols = np.arange(n_atoms, dtype=np.int64)[None, :]
    r, c = np.nonzero((cols - rows) > idx)
    return np.stack([r, c]).astype(np.int32)


# --- utils/linalg.py:84-99 ---------------------------------------------------
def apply_minimum_image(r_ij: np.ndarray, box_array: np.ndarray) -> np.ndarray:
    assert r_ij.dtype == F32 and box_array.dtype == F32
    return r_ij - np.rint(r_ij / box_array) * box_array


# --- radial_distribution_function.py:647-689 ---------------------------------
def get_dij(indices, positions_tensor, atoms, box_array):
    _positions = positions_tensor[indices[1]]  # tf.gather(positions_tensor, indices[1])
    atoms_position = atoms[indices[0]]  # tf.gather(atoms, indices[0])
    r_ij = _positions - atoms_position
    if box_array is not None:
        r_ij = apply_minimum_image(r_ij, box_array)
    sq = r_ij * r_ij  # fp32 products, each rounded
    d2 = (sq[..., 0] + sq[..., 1]) + sq[..., 2]  # Eigen inner-dim scalar reduction order
    return np.sqrt(d2)  # correctly rounded fp32 sqrt


# --- tensorflow/core/kernels/histogram_op.cc (CPU functor) -------------------
def histogram_fixed_width(values, value_range, nbins) -> np.ndarray:
    values = np.asarray(values, dtype=F32).ravel()
    lo, hi = F32(value_range[0]), F32(value_range[1])
    step = float(F32(hi - lo)) / float(nbins)
    shifted = (np.maximum(values, lo) - lo).astype(np.float64)
    index_to_bin = np.minimum(shifted / step, float(nbins - 1)).astype(np.int32)
    return np.bincount(index_to_bin, minlength=nbins).astype(np.int64)


# --- radial_distribution_function.py:616-645 + linalg.py:125-136 -------------
def bin_minibatch(start, stop, indices, d_ij, bin_range, number_of_bins, cutoff):
    mask_1 = (indices[:, 0] > start[0]) & (indices[:, 0] < stop[0])  # strict: Q1
    mask_2 = (indices[:, 1] > start[1]) & (indices[:, 1] < stop[1])
    values_species = d_ij[mask_1 & mask_2]
    values = values_species[values_species < cutoff]  # apply_system_cutoff
    return histogram_fixed_width(values, bin_range, number_of_bins)


def default_cutoff(box_array) -> float:
    """:226-229 -- Python float arithmetic on box_array[0]."""
    return float(box_array[0]) / 2 - 0.1


def default_number_of_bins(cutoff: float) -> int:
    """:239-242."""
    return int(cutoff / 0.01)


def sample_configurations(start, stop, number_of_configurations):
    """:264-269."""
    return np.linspace(start, stop, number_of_configurations, dtype=int)


def rdf_plan(species_shapes, n_configs_total, number_of_configurations, memory,
             memory_fraction=0.5, override_n_batches=None):
    """_prepare_managers (quadratic scale :119-121) + _correct_batch_properties :395-420."""

    class _DB:
        def get_data_size(self, item):
            n = species_shapes[item]
            return n, n_configs_total, n * n_configs_total * 3 * 4

    mm = MemoryManager(
        data_path=list(species_shapes.keys()),
        database=_DB(),
        memory_fraction=memory_fraction,
        scale_function={"quadratic": {"outer_scale_factor": 10, "inner_scale_factor": 5}},
        memory=memory,
    )
    batch_size, n_batches, remainder = mm.get_batch_size()
    _, minibatch = mm.get_ensemble_loop(1, 1)
    if batch_size > number_of_configurations:
        batch_size = number_of_configurations
        n_batches = 1
    else:
        n_batches = int(number_of_configurations / batch_size)
    if override_n_batches is not None:
        n_batches = override_n_batches
    if minibatch:
        batch_size = 1
        n_batches = number_of_configurations
    return batch_size, n_batches


def rdf_counts(
    positions_by_species: dict,
    species: list,
    box_array,
    sample_frames: np.ndarray,
    cutoff: float,
    number_of_bins: int,
    rdf_minibatch: int,
    n_batches: int,
) -> dict:
    """run_calculator :828-887 up to (not including) the normalisation.

    positions_by_species[s] : (n_s, n_frames, 3) float32-valued array (the HDF5
    store is float32, re-read as float64 and cast to self.dtype=float32 in
    _format_data :556-563).
    Returns {"A_B": int64[nbins]}; the reference accumulates in int32 (Q3) --
    ``rdf_int32_overflow`` reports whether that would have overflowed.
    """
    box_f32 = np.asarray(box_array, dtype=F32)
    cutoff_f32 = F32(cutoff)
    bin_range = [F32(0), cutoff_f32]
    index_list = list(range(len(species)))
    key_list = [
        f"{species[a]}_{species[b]}"
        for a, b in itertools.combinations_with_replacement(index_list, 2)
    ]
    particles_list = [positions_by_species[s].shape[0] for s in species]
    total = {name: np.zeros(number_of_bins, dtype=np.int64) for name in key_list}

    split_arr = np.array_split(np.asarray(sample_frames), n_batches)
    for frames in split_arr:
        # data_manager.py:195-201 (fancy frame index) + _format_data :535-563
        positions_tensor = np.concatenate(
            [np.asarray(positions_by_species[s][:, frames, :], dtype=F32) for s in species],
            axis=0,
        )
        n_atoms = positions_tensor.shape[0]
        minibatch_start = 0
        stop = 0
        rdf = {name: np.zeros(number_of_bins, dtype=np.int64) for name in key_list}
        for lo in range(0, n_atoms, rdf_minibatch):  # per_atoms_ds.batch(rdf_minibatch)
            atoms = positions_tensor[lo : lo + rdf_minibatch]
            atoms_per_batch = atoms.shape[0]
            # run_minibatch_loop :422-468
            stop += atoms_per_batch
            indices = get_partial_triu_indices(n_atoms, atoms_per_batch, minibatch_start)
            d_ij = get_dij(indices, positions_tensor, atoms, box_f32)
            # compute_species_values :470-524
            ind_t = indices.T
            for a, b in itertools.combinations_with_replacement(index_list, 2):
                name = f"{species[a]}_{species[b]}"
                start_ = np.array(
                    [sum(particles_list[:a]) - minibatch_start, sum(particles_list[:b])]
                )
                stop_ = start_ + np.array([particles_list[a], particles_list[b]])
                rdf[name] = rdf[name] + bin_minibatch(
                    start_, stop_, ind_t, d_ij, bin_range, number_of_bins, cutoff_f32
                )
            minibatch_start = stop
        for key in total:
            total[key] += rdf[key]
    return total


def rdf_int32_overflow(counts: dict) -> bool:
    return any(int(v.max()) > np.iinfo(np.int32).max for v in counts.values())


# --- radial_distribution_function.py:719-826 ---------------------------------
def _split_array(data, condition):
    initial_split = [data[condition], data[~condition]]
    if len(initial_split[1]) == 0:
        return [data[condition]]
    return list(initial_split)


def ideal_correction(cutoff: float, number_of_bins: int, box0: float) -> np.ndarray:
    def _spherical_symmetry(data):
        return 4 * np.pi * (data**2)

    def _correction_1(data):
        return 2 * np.pi * data * (3 - 4 * data)

    def _correction_2(data):
        arctan_1 = np.arctan(np.sqrt(4 * (data**2) - 2))
        arctan_2 = (
            8
            * data
            * np.arctan(
                (2 * data * (4 * (data**2) - 3))
                / (np.sqrt(4 * (data**2) - 2) * (4 * (data**2) + 1))
            )
        )
        return 2 * data * (3 * np.pi - 12 * arctan_1 + arctan_2)

    def _piecewise(data):
        lower_bound = box0 / 2
        middle_bound = np.sqrt(2) * box0 / 2
        split_1 = _split_array(data, data <= lower_bound)
        if len(split_1) == 1:
            return _spherical_symmetry(split_1[0])
        split_2 = _split_array(split_1[1], split_1[1] < middle_bound)
        if len(split_2) == 1:
            return np.concatenate(
                (_spherical_symmetry(split_1[0]), _correction_1(split_2[0]))
            )
        return np.concatenate(
            (
                _spherical_symmetry(split_1[0]),
                _correction_1(split_2[0]),
                _correction_2(split_2[1]),
            )
        )

    bin_width = cutoff / number_of_bins
    bin_edges = np.linspace(0.0, cutoff, number_of_bins)
    return _piecewise(np.array(bin_edges)) * bin_width


# --- radial_distribution_function.py:299-382 ---------------------------------
def rdf_normalise(
    counts: dict,
    n_particles: dict,
    box_array,
    cutoff: float,
    number_of_bins: int,
    number_of_configurations: int,
    length_unit: float,
) -> dict:
    """Returns {"A_B": {"x": [...nm], "y": [...g(r)]}}."""
    volume = float(np.prod(np.asarray(box_array, dtype=float)))  # experiment_database.py:430-433
    corr = ideal_correction(cutoff, number_of_bins, float(box_array[0]))
    out = {}
    with np.errstate(divide="ignore", invalid="ignore"):
        for names, hist in counts.items():
            a, b = names.split("_")
            species_scale_factor = 2 if a == b else 1
            rho = n_particles[b] / volume
            denominator = number_of_configurations * rho * corr * n_particles[a]
            prefactor = species_scale_factor / denominator
            y = np.array(hist, dtype=float) * prefactor
            x = (length_unit / 1e-9) * np.linspace(0.0, cutoff, number_of_bins)
            out[names] = {"x": x.tolist(), "y": y.tolist()}
    return out


# --- independent cross-check of the binning rule (SURVEY.md A.1 last paragraph)
def rdf_counts_direct(positions: np.ndarray, offsets, counts_n, box_array, cutoff, nbins):
    """All-pairs, frame-by-frame evaluation with the same fp32 arithmetic but no
    minibatch / gather structure.  positions: (N_tot, F, 3) float32.  Used by the
    tests to check that ``rdf_counts`` is independent of the batch plan."""
    box = np.asarray(box_array, dtype=F32)
    cutoff_f32 = F32(cutoff)
    ns = len(offsets)
    out = {}
    N, F, _ = positions.shape
    step = float(cutoff_f32) / float(nbins)
    for a in range(ns):
        for b in range(a, ns):
            h = np.zeros(nbins, dtype=np.int64)
            ia = np.arange(offsets[a] + 1, offsets[a] + counts_n[a])
            jb = np.arange(offsets[b] + 1, offsets[b] + counts_n[b])
            for f in range(F):
                P = positions[:, f, :]
                r = P[jb][None, :, :] - P[ia][:, None, :]
                r = r - np.rint(r / box) * box
                sq = r * r
                d = np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2])
                if a == b:
                    iu = np.triu_indices(len(ia), k=1)
                    d = d[iu]
                else:
                    d = d.ravel()
                d = d[d < cutoff_f32]
                k = np.minimum(d.astype(np.float64) / step, nbins - 1).astype(np.int32)
                h += np.bincount(k, minlength=nbins)
            out[(a, b)] = h
    return out


def rdf_counts_slab(rows: np.ndarray, cols: np.ndarray, box_array, cutoff, nbins,
                    chunk_rows: int = 8) -> np.ndarray:
    """Histogram of ALL (row, column) pair distances of one frame -- a rows x columns slab of
    the pair matrix, evaluated with get_dij's fp32 arithmetic (:647-689, linalg.py:84-99) and
    binned by histogram_fixed_width, a few rows at a time so that a 10^6-column slab fits in
    memory.  rows: (R, 3), cols: (C, 3) float32.  Used to check the pair kernel against the
    oracle at full system size (the species-pair block A x B of a frame is such a slab)."""
    box = np.asarray(box_array, dtype=F32)
    cutoff_f32 = F32(cutoff)
    rows = np.asarray(rows, dtype=F32)
    cols = np.asarray(cols, dtype=F32)
    h = np.zeros(nbins, dtype=np.int64)
    for r0 in range(0, len(rows), chunk_rows):
        r = cols[None, :, :] - rows[r0:r0 + chunk_rows, None, :]
        r = apply_minimum_image(r, box)
        sq = r * r
        d = np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2]).ravel()
        h += histogram_fixed_width(d[d < cutoff_f32], [0, cutoff], nbins)
    return h
