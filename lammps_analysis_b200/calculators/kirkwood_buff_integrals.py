"""KirkwoodBuffIntegral: G(r) = 4 pi int (g(r') - 1) r'^2 dr' on the (smoothed) RDF.

Host post-processing of the RDF result, mirroring
mdsuite/calculators/kirkwood_buff_integrals.py (Args :42-50, __call__ :115-158,
_calculate_kb_integral :160-187, run_calculator :189-203).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.integrate import cumulative_trapezoid
from scipy.signal import savgol_filter

from ..project import Computation
from .calculator import Calculator, call


@dataclass
class Args:
    savgol_order: int
    savgol_window_length: int
    number_of_bins: int
    number_of_configurations: int
    cutoff: float


class KirkwoodBuffIntegral(Calculator):
    analysis_name = "Kirkwood-Buff_Integral"
    result_series_keys = ["r", "kb_integral"]

    @call
    def __call__(self, rdf_data: Computation = None, plot=True, savgol_order: int = 2,
                 savgol_window_length: int = 17):
        if isinstance(rdf_data, Computation):
            self.rdf_data = rdf_data
        else:
            self.rdf_data = self.experiment.run.RadialDistributionFunction(plot=False)
        self.plot = plot
        par = self.rdf_data.computation_parameter
        self.args = Args(savgol_order=savgol_order, savgol_window_length=savgol_window_length,
                         number_of_bins=par["number_of_bins"], cutoff=par["cutoff"],
                         number_of_configurations=par["number_of_configurations"])

    def run_calculator(self):
        for selected_species, vals in self.rdf_data.data_dict.items():
            radii = np.array(vals["x"]).astype(float)[1:]
            rdf = np.array(vals["y"]).astype(float)[1:]
            filtered = savgol_filter(rdf, self.args.savgol_window_length, self.args.savgol_order)
            integral = cumulative_trapezoid(y=(filtered[1:] - 1) * radii[1:] ** 2, x=radii[1:])
            self.queue_data(data={"r": radii[1:].tolist(),
                                  "kb_integral": (4 * np.pi * integral).tolist()},
                            subjects=selected_species.split("_"))
