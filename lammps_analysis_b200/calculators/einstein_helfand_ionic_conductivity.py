"""EinsteinHelfandIonicConductivity: MSD of the translational dipole moment of the system.

SURVEY.md 8f-2: reuses the MSD kernels on the one-row observable
``Observables/Translational_Dipole_Moment`` (produced by the TranslationalDipoleMoment
transformation, itself the ionic-current reduction applied to unwrapped positions).  Mirrors
mdsuite/calculators/einstein_helfand_ionic_conductivity.py (Args :43-51, __call__ :112-156,
prefactor :167-186, ensemble_operation :192-209, post-processing :211-231, run_calculator
:233-258).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Union

import numpy as np

from ..engine import msd_series
from ..planner import frame_batches, window_count
from ..store import join_path
from ..units import boltzmann_constant, elementary_charge
from .calculator import TrajectoryCalculator, call
from .einstein_diffusion_coefficients import fit_einstein_curve


@dataclass
class Args:
    data_range: int
    correlation_time: int
    tau_values: object
    atom_selection: object
    fit_range: int


class EinsteinHelfandIonicConductivity(TrajectoryCalculator):
    analysis_name = "Einstein_Helfand_Ionic_Conductivity"
    loaded_property = "Translational_Dipole_Moment"
    system_property = True
    scale_function = {"linear": {"scale_factor": 5}}
    result_keys = ["ionic_conductivity", "uncertainty"]
    result_series_keys = ["time", "msd"]

    @call
    def __call__(self, plot: bool = True, data_range: int = 500, correlation_time: int = 1,
                 tau_values: Union[int, list, slice] = np.s_[:], fit_range: int = -1):
        if fit_range == -1:
            fit_range = int(data_range - 1)
        self.args = Args(data_range=data_range, correlation_time=correlation_time,
                         tau_values=tau_values, atom_selection=np.s_[:], fit_range=fit_range)
        self.plot = plot
        self.time = self._handle_tau_values()

    def check_input(self):
        self._run_dependency_check()

    def _calculate_prefactor(self) -> float:
        exp, u = self.experiment, self.experiment.units
        numerator = u.length**2 * elementary_charge**2
        denominator = u.time * exp.volume * u.volume * exp.temperature * boltzmann_constant
        return numerator / denominator

    def compute_msd(self):
        """Returns (msd_sum [n_tau] host, number of windows)."""
        store = self.experiment.store
        path = join_path("Observables", self.loaded_property)
        self._prepare_managers([path])
        batches = frame_batches(self.plan)
        if len(batches) != 1:
            raise ValueError("system observable requested with more than one batch (the "
                             "reference cannot do this either: data_manager.py:204-205)")
        M = store.device(path)
        (t0, t1), N, ct = batches[0], self.args.data_range, self.args.correlation_time
        if t1 - t0 < N:
            raise ValueError("data_range exceeds the number of configurations")
        W = window_count(t1 - t0, N, ct)
        msd, _ = msd_series(M, [(0, 1, t0, t1 - t0, W)], N, ct, self.args.tau_values)
        return msd.cpu().numpy(), W

    def run_calculator(self):
        self.check_input()
        prefactor = self._calculate_prefactor()
        msd_sum, W = self.compute_msd()
        # _apply_averaging_factor (:188-190): / (n_batches * ensemble_loop)
        msd = prefactor * msd_sum / (int(self.plan.n_batches) * W)
        popt, pcov, _, _ = fit_einstein_curve(self.time, msd, self.args.fit_range)
        if len(popt) == 0:
            raise ValueError("fit_range lies before the linear regime found by the spline")
        error = np.sqrt(np.diag(pcov))[0]
        self.queue_data(data={"ionic_conductivity": 1 / 6 * popt[0], "uncertainty": 1 / 6 * error,
                              "time": np.asarray(self.time).tolist(), "msd": msd.tolist()},
                        subjects=["System"])
