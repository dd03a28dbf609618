"""GreenKuboDiffusionCoefficients: windowed velocity autocorrelation -> D = integral / 3.

Mirrors mdsuite/calculators/green_kubo_self_diffusion_coefficients.py (Args :58-69, __call__
:111-167, ensemble_operation :179-206, postprocessing :270-300, run_calculator :302-337).
The reference runs ``tfp.stats.auto_correlation`` (complex128 FFT) per window, atom and
dimension; here one ``mdk_acf_lagprod`` pass forms the lag products
P[t][m] = sum_a,d v(t) v(t+m) of a whole batch and ``mdk_acf_windows`` turns prefix sums of P
into every window's unbiased ACF (needed for the SEM) and their sum.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Union

import numpy as np
from scipy.integrate import cumulative_trapezoid

from .. import distributed as D
from .. import kernels as K
from .. import trace
from ..engine import acf_series, plan_windows
from ..store import join_path
from .calculator import TrajectoryCalculator, call


@dataclass
class Args:
    data_range: int
    correlation_time: int
    tau_values: object
    molecules: bool
    species: list
    atom_selection: object
    integration_range: int


class GreenKuboDiffusionCoefficients(TrajectoryCalculator):
    analysis_name = "Green_Kubo_Self_Diffusion"
    loaded_property = "Velocities"
    scale_function = {"linear": {"scale_factor": 150}}
    result_keys = ["diffusion_coefficient", "uncertainty"]
    result_series_keys = ["time", "acf", "integral", "integral_uncertainty"]

    @call
    def __call__(self, plot: bool = True, species: list = None, data_range: int = 500,
                 correlation_time: int = 1, atom_selection=np.s_[:], molecules: bool = False,
                 tau_values: Union[int, list, slice] = np.s_[:], integration_range: int = None):
        if species is None:
            species = list(self.experiment.species)
        if integration_range is None:
            integration_range = data_range - 1
        self.args = Args(data_range=data_range, correlation_time=correlation_time,
                         atom_selection=atom_selection, tau_values=tau_values,
                         molecules=molecules, species=species,
                         integration_range=integration_range)
        self.plot = plot
        self.time = self._handle_tau_values() * self.experiment.units.time

    def check_input(self):
        if self.args.molecules:
            raise NotImplementedError("molecule diffusion needs the molecule-mapping subsystem")
        if self.data_resolution != self.args.data_range:
            # upstream adds a full-length ACF to zeros(data_resolution) and fails (:311, :332)
            raise ValueError("GreenKuboDiffusionCoefficients needs the full tau range "
                             "(tau_values=np.s_[:])")
        self._run_dependency_check()

    # -- hot path ----------------------------------------------------------------------------------
    def compute_acf(self, species: str):
        """Returns (acf_sum [N], count, per-window atom-summed ACFs [W_total][N], A_sel per
        window [W_total]) on the host, in simulation units (no length^2/time^2 factor)."""
        path = join_path(species, self.loaded_property)
        self._prepare_managers([path])
        import torch

        traj, n_atoms, shard, offset, blocks = self._device_row_blocks(path, species)
        launches = plan_windows(self.plan.as_dict(), self.args.data_range,
                                self.args.correlation_time, n_atoms)
        # the lag-product kernel follows the row blocks of the velocity upload
        cur, side = torch.cuda.current_stream(), self._side_stream()
        with torch.cuda.stream(side):
            acf, count, wins, sizes = acf_series(traj, launches, self.args.data_range,
                                                 self.args.correlation_time, per_window=True,
                                                 a_shard=shard, row_offset=offset, blocks=blocks)
            D.all_reduce_sum_([acf] + wins)
        trace.mark(f"GreenKubo[{species}] kernels enqueued")
        # the host is about to wait for these kernels and to post-process: queue the upload of
        # the next species behind this one so that the link stays busy meanwhile
        self._prefetch_next(species)
        with torch.cuda.stream(side):
            got = K.read_back(acf, *wins)
            acf_host = got[0]
            win_host = np.concatenate(got[1:], axis=0) if wins else \
                np.zeros((0, self.args.data_range))
        cur.wait_stream(side)
        a_sel = np.concatenate([np.full(w.shape[0], s, dtype=float) for w, s in zip(wins, sizes)]) \
            if wins else np.zeros(0)
        trace.mark(f"GreenKubo[{species}] windows on the host")
        return acf_host, count, win_host, a_sel

    def _prefetch_next(self, species: str):
        names = list(self.args.species)
        k = names.index(species) + 1
        if k < len(names) and not isinstance(self.args.atom_selection, dict):
            self.experiment.store.device_blocks(join_path(names[k], self.loaded_property))

    def postprocessing(self, acf_sum, count, win, a_sel) -> dict:
        units = self.experiment.units
        scale = units.length**2 / units.time**2
        acf = scale * np.asarray(acf_sum) / count
        # sigmas: cumulative trapezoid of (sum_d mean_a vacf) per window (:200-204)
        sigmas = cumulative_trapezoid(scale * win / a_sel[:, None], x=self.time, axis=1)
        sigma = cumulative_trapezoid(acf, x=self.time)
        sigma_sem = np.std(sigmas, axis=0) / np.sqrt(len(sigmas))
        ir = self.args.integration_range
        return {"diffusion_coefficient": [1 / 3 * sigma[ir - 1]],
                "uncertainty": [1 / 3 * sigma_sem[ir - 1]], "time": self.time.tolist(),
                "acf": acf.tolist(), "integral": sigma.tolist(),
                "integral_uncertainty": sigma_sem.tolist()}

    def run_calculator(self):
        self.check_input()
        for species in self.args.species:
            acf_sum, count, win, a_sel = self.compute_acf(species)
            self.queue_data(data=self.postprocessing(acf_sum, count, win, a_sel),
                            subjects=[species])
            trace.mark(f"GreenKubo[{species}] post-processing done")
