"""AngularDistributionFunction: distribution of the angles j-i-k between the neighbours of
every atom, for all species triples.

Mirrors mdsuite/calculators/angular_distribution_function.py (Args :51-69, __call__ :141-214,
check_input :216-236, _prepare_data_structure :268-300, _build_histograms :365-403,
_compute_adfs :405-444, _correct_batch_properties :506-527, run_calculator :584-609) and
mdsuite/utils/neighbour_list.py.  The reference builds the dense (n, n, 3) r_ij matrix and an
(n, n, n) float16 roll-and-compare to enumerate triples; here ``engine.AdfEngine`` runs a
cell-list neighbour search and one warp per centre atom (csrc/adf.cu), with the reference's
arithmetic: float16 cutoff test, ordered neighbour pairs, species triples (centre, j, k) in
combinations_with_replacement order, numpy's float32 bin edges, and one density normalisation
PER BATCH (so the batch plan is part of the result).  Batches shard across ranks.
"""
from __future__ import annotations

import itertools
from dataclasses import dataclass

import numpy as np
import torch

from .. import distributed as D
from .. import kernels as K
from ..engine import AdfEngine
from ..planner import plan_batches
from ..store import join_path
from .calculator import TrajectoryCalculator, call


@dataclass
class Args:
    number_of_bins: int
    number_of_configurations: int
    correlation_time: int
    atom_selection: object
    data_range: int
    cutoff: float
    start: int
    norm_power: object
    stop: int
    species: list
    molecules: bool


class AngularDistributionFunction(TrajectoryCalculator):
    analysis_name = "Angular_Distribution_Function"
    loaded_property = "Positions"
    result_keys = ["max_peak"]
    result_series_keys = ["angle", "adf"]
    scale_function = {"quadratic": {"outer_scale_factor": 10}}
    bin_range = [0.0, 3.15]

    @call
    def __call__(self, batch_size: int = 1, minibatch: int = -1,
                 number_of_configurations: int = 5, cutoff: float = 6.0, start: int = 1,
                 stop: int = None, number_of_bins: int = 500, species: list = None,
                 use_tf_function: bool = False, molecules: bool = False,
                 atom_selection=np.s_[:], plot: bool = True, norm_power: int = 4, **kwargs):
        self.args = Args(number_of_bins=number_of_bins, cutoff=cutoff, start=start, stop=stop,
                         atom_selection=atom_selection, data_range=1, correlation_time=1,
                         molecules=molecules, species=species,
                         number_of_configurations=number_of_configurations,
                         norm_power=norm_power)
        self.plot = plot
        self.adf_minibatch = minibatch          # accepted for API parity (memory knob upstream)
        self.override_n_batches = kwargs.get("batches")

    def check_input(self):
        exp = self.experiment
        if self.args.molecules:
            raise NotImplementedError("molecule ADFs need the molecule-mapping subsystem "
                                      "(out of the hot-path scope)")
        if isinstance(self.args.atom_selection, dict):
            raise NotImplementedError("AngularDistributionFunction: atom_selection dictionaries "
                                      "are not supported")
        if self.args.stop is None:
            self.args.stop = exp.number_of_configurations - 1
        if self.args.species is None:
            self.args.species = list(exp.species)
        self.sample_configurations = np.linspace(self.args.start, self.args.stop,
                                                 self.args.number_of_configurations, dtype=int)

    def number_of_batches(self) -> int:
        """_prepare_managers + _correct_batch_properties (:506-527)."""
        store = self.experiment.store
        paths = [join_path(s, self.loaded_property) for s in self.args.species]
        plan = plan_batches([store.get_data_size(p) for p in paths], 1, 1, self.scale_function)
        n_cfg = self.args.number_of_configurations
        n_batches = 1 if plan.batch_size > n_cfg else int(n_cfg / plan.batch_size)
        if self.override_n_batches is not None:
            n_batches = int(self.override_n_batches)
        if plan.minibatch:
            n_batches = n_cfg
        return n_batches

    # -- hot path ----------------------------------------------------------------------------------
    def _frames_on_device(self, frames: np.ndarray) -> torch.Tensor:
        """[F][N][3] float32 CUDA tensor of the given frames, species concatenated.  With an
        atom-sharded store the ranks first swap their atom blocks of the frames they need."""
        store = self.experiment.store
        fdev = torch.from_numpy(np.asarray(frames, dtype=np.int32)).cuda()
        parts = []
        for sp in self.args.species:
            path = join_path(sp, self.loaded_property)
            pin = store.pinned_tensor(path)
            if store.is_resident(path):
                local = store.device(path).index_select(1, fdev.long())
            elif pin is not None:
                local = K.gather_frames(pin, fdev)
            else:
                local = store.device_frames(path, frames)
            parts.append(local)
        return torch.cat(parts, dim=0).permute(1, 0, 2).contiguous()

    def compute_histograms(self):
        """Per batch: (weights float64 [n_combos][nbins], counts int64) on the host, for the
        batches this rank owns, keyed by batch index."""
        exp = self.experiment
        store = exp.store
        species = self.args.species
        counts = [exp.species[s].n_particles for s in species]
        self.engine = AdfEngine(counts, exp.box_array, self.args.cutoff, self.args.number_of_bins,
                                self.args.norm_power)
        batches = np.array_split(self.sample_configurations, self.number_of_batches())
        paths = [join_path(s, self.loaded_property) for s in species]
        sharded = any(store.is_sharded(p) for p in paths)
        out = {}
        if not sharded:
            for b in range(D.rank(), len(batches), D.world_size()):
                if len(batches[b]):
                    w, c = self.engine.add_batch(self._frames_on_device(batches[b]))
                    out[b] = (w.cpu().numpy(), c.cpu().numpy())
            return out, len(batches)
        # atom-sharded store: every rank gathers its atom block of all sampled frames, one
        # all-to-all per species hands each rank all atoms of the frames of ITS batches
        w_size, r = D.world_size(), D.rank()
        owner = np.concatenate([np.full(len(fr), b % w_size) for b, fr in enumerate(batches)])
        order = np.argsort(owner, kind="stable")          # frames grouped by owning rank
        frames_all = np.concatenate(batches)[order]
        per_rank = np.bincount(owner, minlength=w_size)
        fdev = torch.from_numpy(frames_all.astype(np.int32)).cuda()
        full = []
        for sp, path in zip(species, paths):
            pin = store.pinned_tensor(path)
            if store.is_resident(path):
                local = store.device(path).index_select(1, fdev.long())
            elif pin is not None:
                local = K.gather_frames(pin, fdev)
            else:
                local = store.device_frames(path, frames_all)
            rows = [hi - lo for lo, hi in store.rows_per_rank(path)]
            full.append(D.exchange_frame_groups(local, rows, per_rank))
        mine = torch.cat(full, dim=0).permute(1, 0, 2).contiguous()   # [F_mine][N][3]
        k = 0
        for b, fr in enumerate(batches):
            if b % w_size != r:
                continue
            if len(fr):
                w, c = self.engine.add_batch(mine[k:k + len(fr)])
                out[b] = (w.cpu().numpy(), c.cpu().numpy())
            k += len(fr)
        return out, len(batches)

    def run_calculator(self):
        self.check_input()
        per_batch, n_batches = self.compute_histograms()
        nbins = self.args.number_of_bins
        n_combos = self.engine.n_combos
        # numpy.histogram(..., density=True): n / diff(edges) / n.sum() per batch, cast to
        # float32, summed over the batches (:388-400)
        edges = np.linspace(self.bin_range[0], self.bin_range[1], nbins + 1).astype(np.float32)
        db = np.diff(edges).astype(float)
        total = np.zeros((n_combos, nbins), dtype=np.float32)
        triples = np.zeros(n_combos, dtype=np.int64)
        for b in sorted(per_batch):
            w, c = per_batch[b]
            with np.errstate(invalid="ignore", divide="ignore"):
                dens = w / db / w.sum(axis=1, keepdims=True)
            total += dens.astype(np.float32)
            triples += c.sum(axis=1)
        if D.world_size() > 1:
            t = torch.from_numpy(total).cuda()
            n = torch.from_numpy(triples).cuda()
            D.all_reduce_sum_([t, n])
            total, triples = t.cpu().numpy(), n.cpu().numpy()
        axis = np.linspace(self.bin_range[0] * (180 / 3.14159), self.bin_range[1] * (180 / 3.14159),
                           nbins)
        self.triple_counts = {}
        combos = itertools.combinations_with_replacement(self.args.species, 3)
        for p, names in enumerate(combos):
            hist = total[p]
            self.triple_counts["_".join(names)] = int(triples[p])
            data = {"max_peak": float(axis[int(np.argmax(hist))]), "angle": axis.tolist(),
                    "adf": hist.tolist()}
            self.queue_data(data=data, subjects=list(names))
        self.queue_metadata(triples=self.triple_counts, n_batches=int(n_batches),
                            neighbour_capacity=int(self.engine.capacity))
