"""GreenKuboThermalConductivity and GreenKuboViscosity: windowed autocorrelation of a system flux.

SURVEY.md 8f-2: both reuse the ACF kernels (``mdk_acf_lagprod`` + ``mdk_acf_windows``) on a
one-row observable -- ``Observables/Thermal_Flux`` (ThermalFlux transformation) and
``Observables/Momentum_Flux`` (MomentumFlux transformation).  They mirror
mdsuite/calculators/green_kubo_thermal_conductivity.py (Args :43-52, __call__ :100-140,
prefactor :152-176, ensemble_operation :188-213, post-processing :215-248, run_calculator
:250-281) and green_kubo_viscosity.py (same structure; prefactor :146-171, results :214-227),
including two properties of the reference a user may not expect:

* every window's series is ``data_range * sum_dims tfp.auto_correlation(window)`` (unbiased,
  divided by N - m) and the stored ``acf`` is the SUM of those series over the windows
  (``_apply_averaging_factor`` is a no-op, :178-186);
* the reported value and "uncertainty" are ``prefactor * sigma[0]`` and ``prefactor * sigma[1]``:
  the trapezoidal integrals of the FIRST TWO windows (:222-227), so at least two windows are
  needed (the reference raises IndexError otherwise).

The reference's own integration tests for these calculators are disabled
(CI/integration_tests/calculators/_test_green_kubo_thermal_conductivity.py, _test_green_kubo_viscosity.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Union

import numpy as np

from ..engine import acf_series
from ..planner import frame_batches, window_count
from ..store import join_path
from .calculator import TrajectoryCalculator, call


@dataclass
class Args:
    data_range: int
    correlation_time: int
    tau_values: object
    atom_selection: object
    integration_range: int


class _GreenKuboFlux(TrajectoryCalculator):
    system_property = True
    scale_function = {"linear": {"scale_factor": 5}}
    value_key: str = None
    result_series_keys = ["time", "acf"]

    @call
    def __call__(self, plot: bool = False, data_range: int = 500,
                 tau_values: Union[int, list, slice] = np.s_[:], correlation_time: int = 1,
                 integration_range: int = None):
        self.plot = plot
        if integration_range is None:
            integration_range = data_range
        self.args = Args(data_range=data_range, correlation_time=correlation_time,
                         tau_values=tau_values, atom_selection=np.s_[:],
                         integration_range=integration_range)
        self.time = self._handle_tau_values()

    def check_input(self):
        if self.data_resolution != self.args.data_range:
            raise ValueError(f"{type(self).__name__} is implemented for the full tau range "
                             "(tau_values=np.s_[:])")
        self._run_dependency_check()

    def _calculate_prefactor(self) -> float:
        raise NotImplementedError

    def compute_acf(self):
        """Returns (sum over windows of the unbiased ACF [N], per-window ACFs [W][N])."""
        store = self.experiment.store
        path = join_path("Observables", self.loaded_property)
        self._prepare_managers([path])
        batches = frame_batches(self.plan)
        if len(batches) != 1:
            raise ValueError("system observable requested with more than one batch (the "
                             "reference cannot do this either: data_manager.py:204-205)")
        J = store.device(path)
        (t0, t1), N, ct = batches[0], self.args.data_range, self.args.correlation_time
        B = t1 - t0
        if B < N:
            raise ValueError("data_range exceeds the number of configurations")
        W = window_count(B, N, ct)
        acf, _, wins, _ = acf_series(J, [(0, 1, t0, B, W)], N, ct, per_window=True)
        return acf.cpu().numpy(), wins[0].cpu().numpy()

    def _stored_acf(self, acf_sum: np.ndarray, win: np.ndarray) -> np.ndarray:
        return self.args.data_range * acf_sum                 # self.jacf += jacf, per window

    def run_calculator(self):
        self.check_input()
        prefactor = self._calculate_prefactor()
        acf_sum, win = self.compute_acf()
        N, ir = self.args.data_range, self.args.integration_range
        jacf = self._stored_acf(acf_sum, win)
        trapz = getattr(np, "trapezoid", None) or np.trapz   # renamed in NumPy 2
        sigma = trapz(N * win[:, :ir], x=self.time[:ir], axis=1)
        result = prefactor * sigma
        if len(result) < 2:
            raise IndexError("the reference reports the integrals of the first two windows: "
                             "at least two windows are needed")
        self.queue_data(data={self.value_key: result[0], "uncertainty": result[1],
                              "time": self.time.tolist(), "acf": jacf.tolist()},
                        subjects=["System"])


class GreenKuboThermalConductivity(_GreenKuboFlux):
    analysis_name = "Green_Kubo_Thermal_Conductivity"
    loaded_property = "Thermal_Flux"
    value_key = "computation_results"     # sic (green_kubo_thermal_conductivity.py:225)
    result_keys = ["computation_results", "uncertainty"]

    def _calculate_prefactor(self) -> float:
        exp, u = self.experiment, self.experiment.units
        denominator = (3 * (self.args.data_range - 1) * exp.temperature**2 * u.boltzmann
                       * exp.volume)
        return (1 / denominator) * (u.energy / u.length / u.time)


class GreenKuboViscosityFlux(_GreenKuboFlux):
    """green_kubo_viscosity_flux.py:55-290: the autocorrelation of the off-diagonal pressure
    tensor ``Observables/Stress_Visc`` (pxy, pxz, pyz of a LAMMPS log, read by
    ``LAMMPSFluxFile``; no transformation produces it).  Restated as the reference computes it:
    the prefactor carries the volume in the NUMERATOR (:149-165), every window adds the single
    value ``jacf[data_range - 1]`` to the whole stored series (:199) and the series is then
    divided by its maximum (:167-175) -- the stored "acf" is therefore constant 1 (NaN when the
    sum is 0); the reported value / "uncertainty" are the integrals of the first two windows."""

    analysis_name = "Viscosity_Flux"
    loaded_property = "Stress_Visc"
    value_key = "viscosity"
    result_keys = ["viscosity", "uncertainty"]

    def _calculate_prefactor(self) -> float:
        exp, u = self.experiment, self.experiment.units
        denominator = 3 * (self.args.data_range - 1) * exp.temperature * u.boltzmann
        return (exp.volume / denominator) * (u.pressure**2 * u.volume * u.time / u.energy)

    def _stored_acf(self, acf_sum: np.ndarray, win: np.ndarray) -> np.ndarray:
        N = self.args.data_range
        jacf = np.zeros(self.data_resolution) + (N * win[:, N - 1]).sum()   # broadcast add
        with np.errstate(invalid="ignore", divide="ignore"):
            return jacf / np.max(jacf)


class GreenKuboViscosity(_GreenKuboFlux):
    analysis_name = "Green_Kubo_Viscosity"
    loaded_property = "Momentum_Flux"
    value_key = "viscosity"
    result_keys = ["viscosity", "uncertainty"]

    def _calculate_prefactor(self) -> float:
        exp, u = self.experiment, self.experiment.units
        denominator = 3 * (self.args.data_range - 1) * exp.temperature * u.boltzmann * exp.volume
        return (1 / denominator) * (u.pressure**2 * u.volume * u.time / u.energy)
