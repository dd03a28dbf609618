"""EinsteinHelfandThermalConductivity: MSD of the integrated heat current of the system.

SURVEY.md 8f-2: the MSD kernels on the one-row observable
``Observables/Integrated_Heat_Current`` (IntegratedHeatCurrent transformation: unwrapped
positions weighted by the per-atom energies).  Mirrors
mdsuite/calculators/einstein_helfand_thermal_conductivity.py (__call__ :105-141, prefactor
:152-174, averaging :176-185, ensemble_operation :187-203, post-processing :205-229); the window
loop is the one of EinsteinHelfandIonicConductivity.
"""
from __future__ import annotations

import numpy as np

from .einstein_diffusion_coefficients import fit_einstein_curve
from .einstein_helfand_ionic_conductivity import EinsteinHelfandIonicConductivity


class EinsteinHelfandThermalConductivity(EinsteinHelfandIonicConductivity):
    analysis_name = "Einstein Helfand Thermal Conductivity"
    loaded_property = "Integrated_Heat_Current"
    result_keys = ["thermal_conductivity", "uncertainty"]

    def _calculate_prefactor(self) -> float:
        exp, u = self.experiment, self.experiment.units
        denominator = exp.volume * exp.temperature * u.boltzmann
        return (1 / denominator) * (u.energy / u.length / u.time / u.temperature)

    def run_calculator(self):
        self.check_input()
        prefactor = self._calculate_prefactor()
        msd_sum, W = self.compute_msd()
        msd = prefactor * msd_sum / (int(self.plan.n_batches) * W)   # :176-185
        popt, pcov, _, _ = fit_einstein_curve(self.time, msd, self.args.fit_range)
        if len(popt) == 0:
            raise ValueError("fit_range lies before the linear regime found by the spline")
        error = np.sqrt(np.diag(pcov))[0]
        self.queue_data(data={"thermal_conductivity": 1 / 6 * popt[0], "uncertainty": 1 / 6 * error,
                              "time": np.asarray(self.time).tolist(), "msd": msd.tolist()},
                        subjects=["System"])
