"""GreenKuboIonicConductivity: autocorrelation of the system ionic current.

Mirrors mdsuite/calculators/green_kubo_ionic_conductivity.py (Args :56-64, __call__ :113-155,
prefactor :167-186, ensemble_operation :188-206, post-processing :208-231, run_calculator
:286-310).  The current ``Observables/Ionic_Current`` (1, T, 3) is produced by the IonicCurrent
transformation when missing; its windowed ACF uses the same kernels as the velocity ACF with a
single row (4.75e6 updates at config 3 -- not sharded, every rank computes it).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Union

import numpy as np
from scipy.integrate import cumulative_trapezoid

from ..engine import acf_series
from ..planner import frame_batches, window_count
from ..store import join_path
from ..units import boltzmann_constant, elementary_charge
from .calculator import TrajectoryCalculator, call


@dataclass
class Args:
    data_range: int
    correlation_time: int
    tau_values: object
    atom_selection: object
    integration_range: int


class GreenKuboIonicConductivity(TrajectoryCalculator):
    analysis_name = "Green_Kubo_Ionic_Conductivity"
    loaded_property = "Ionic_Current"
    system_property = True
    scale_function = {"linear": {"scale_factor": 5}}
    result_keys = ["ionic_conductivity", "uncertainty"]
    result_series_keys = ["time", "acf", "integral", "integral_uncertainty"]

    @call
    def __call__(self, plot: bool = True, data_range: int = 500, correlation_time: int = 1,
                 tau_values: Union[int, list, slice] = np.s_[:], integration_range: int = None):
        self.plot = plot
        if integration_range is None:
            integration_range = data_range - 1
        self.args = Args(data_range=data_range, correlation_time=correlation_time,
                         tau_values=tau_values, atom_selection=np.s_[:],
                         integration_range=integration_range)
        self.time = self._handle_tau_values()

    def check_input(self):
        if self.data_resolution != self.args.data_range:
            raise ValueError("GreenKuboIonicConductivity is implemented for the full tau range "
                             "(tau_values=np.s_[:])")
        self._run_dependency_check()

    def _calculate_prefactor(self) -> float:
        exp, u = self.experiment, self.experiment.units
        numerator = elementary_charge**2 * u.length**2
        denominator = (3 * boltzmann_constant * exp.temperature * exp.volume * u.volume * u.time)
        return numerator / denominator

    def compute_acf(self):
        """Returns (acf_sum [N], count = number of windows, per-window ACFs [W][N])."""
        store = self.experiment.store
        path = join_path("Observables", self.loaded_property)
        self._prepare_managers([path])
        batches = frame_batches(self.plan)
        if len(batches) != 1:
            # Q7: the reference slices axis 0 of the (1, T, 3) dataset with the frame range, so
            # system observables only work when everything fits one batch
            raise ValueError("system observable requested with more than one batch (the "
                             "reference cannot do this either: data_manager.py:204-205)")
        J = store.device(path)
        (t0, t1), N, ct = batches[0], self.args.data_range, self.args.correlation_time
        B = t1 - t0
        if B < N:
            raise ValueError("data_range exceeds the number of configurations")
        W = window_count(B, N, ct)
        acf, _, wins, _ = acf_series(J, [(0, 1, t0, B, W)], N, ct, per_window=True)
        return acf.cpu().numpy(), W, wins[0].cpu().numpy()

    def run_calculator(self):
        self.check_input()
        prefactor = self._calculate_prefactor()
        acf_sum, count, win = self.compute_acf()
        acf = acf_sum / count
        sigmas = cumulative_trapezoid(win, x=self.time, axis=1)
        sigma = cumulative_trapezoid(acf, x=self.time)
        sigma_sem = np.std(sigmas, axis=0) / np.sqrt(len(sigmas))
        ir = self.args.integration_range
        data = {"ionic_conductivity": [prefactor * sigma[ir - 1]],
                "uncertainty": [prefactor * sigma_sem[ir - 1]], "time": self.time.tolist(),
                "acf": acf.tolist(), "integral": sigma.tolist(),
                "integral_uncertainty": sigma_sem.tolist()}
        self.queue_data(data=data, subjects=["System"])
