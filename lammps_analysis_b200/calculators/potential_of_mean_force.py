"""PotentialOfMeanForce: w(r) = -k_B T ln g(r) and its value at the minima between peaks.

Host post-processing of the RDF result, mirroring
mdsuite/calculators/potential_of_mean_force.py (Args :46-57, __call__ :126-181,
_calculate_potential_of_mean_force :183-200, get_pomf_peaks :222-262, _find_minimum :264-292,
_get_pomf_values :294-325, run_calculator :327-349).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.signal import find_peaks, savgol_filter

from ..project import Computation
from ..units import boltzmann_constant
from .calculator import Calculator, call
from .coordination_number_calculation import golden_section_search


@dataclass
class Args:
    savgol_order: int
    savgol_window_length: int
    number_of_bins: int
    number_of_configurations: int
    cutoff: float
    number_of_shells: int


class PotentialOfMeanForce(Calculator):
    analysis_name = "Potential_of_Mean_Force"
    result_series_keys = ["r", "pomf"]

    @call
    def __call__(self, rdf_data: Computation = None, plot=True, savgol_order: int = 2,
                 savgol_window_length: int = 17, number_of_shells: int = 1):
        if isinstance(rdf_data, Computation):
            self.rdf_data = rdf_data
        else:
            self.rdf_data = self.experiment.run.RadialDistributionFunction(plot=False)
        self.plot = plot
        par = self.rdf_data.computation_parameter
        self.args = Args(savgol_order=savgol_order, savgol_window_length=savgol_window_length,
                         number_of_bins=par["number_of_bins"], cutoff=par["cutoff"],
                         number_of_configurations=par["number_of_configurations"],
                         number_of_shells=number_of_shells)

    def run_calculator(self):
        a = self.args
        for selected_species, vals in self.rdf_data.data_dict.items():
            radii = np.array(vals["x"]).astype(float)[1:]
            rdf = np.array(vals["y"]).astype(float)[1:]
            with np.errstate(divide="ignore", invalid="ignore"):
                pomf = -1 * boltzmann_constant * self.experiment.temperature * np.log(rdf)
            pomf = pomf * 6.242e8  # "convert to eV" (:200, reproduced literally)
            filtered = savgol_filter(pomf, a.savgol_window_length, a.savgol_order)
            peaks = find_peaks(filtered)[0]
            if len(peaks) < a.number_of_shells + 1:
                raise ValueError("Not enough peaks were detecting in the RDF to perform the "
                                 "desired analysis.")
            data = {"r": radii[1:].tolist(), "pomf": pomf.tolist()}
            for i in range(a.number_of_shells):
                lo, hi = golden_section_search([radii, pomf], radii[peaks[i + 1]], radii[peaks[i]])
                idx = [int(np.where(radii == v)[0][0]) for v in (lo, hi)]
                lower, upper = pomf[idx[0]], pomf[idx[1]]
                data[f"POMF_{i + 1}"] = float(np.mean([lower, upper]))
                data[f"POMF_{i + 1}_error"] = float(np.std([lower, upper]) / np.sqrt(2))
            self.queue_data(data=data, subjects=selected_species.split("_"))
