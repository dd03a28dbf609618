"""CoordinationNumbers: integrates the RDF and reads the value at the minima between peaks.

Host-side post-processing of the RDF result (no kernel), mirroring
mdsuite/calculators/coordination_number_calculation.py:59-81, 156-359 and
mdsuite/utils/meta_functions.py:327-437 (savgol filter, golden-section search).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
from scipy.integrate import cumulative_trapezoid
from scipy.signal import find_peaks, savgol_filter

from ..project import Computation
from ..units import golden_ratio
from .calculator import Calculator, call


class CannotPerformThisAnalysis(Exception):
    """utils/exceptions.py."""


@dataclass
class Args:
    savgol_order: int
    savgol_window_length: int
    number_of_shells: int


def _closest(data, value):
    data = np.asarray(data)
    return data[np.argmin(np.abs(data - value))]


def golden_section_search(data, a, b, tol=1e-5):
    """meta_functions.py:376-437, iteratively: bracket the minimum of data[1] over the grid
    data[0] between a and b."""
    xs, ys = np.asarray(data[0]), np.asarray(data[1])
    phi_a, phi_b = 1 / golden_ratio, 1 / golden_ratio**2
    a, b = min(a, b), max(a, b)
    h = b - a
    c = d = fc = fd = None
    while h > tol:
        a, b = min(a, b), max(a, b)  # the reference re-orders the bracket at every level
        if c is None:
            c = _closest(xs, a + phi_b * h)
            fc = ys[np.where(xs == c)]
        if d is None:
            d = _closest(xs, a + phi_a * h)
            fd = ys[np.where(xs == d)]
        if fc < fd:
            b, d, fd, c, fc = d, c, fc, None, None
        else:
            a, c, fc, d, fd = c, d, fd, None, None
        h = h * phi_a
    return min(a, b), max(a, b)


class CoordinationNumbers(Calculator):
    analysis_name = "Coordination_Numbers"

    @call
    def __call__(self, rdf_data: Computation = None, plot: bool = True, savgol_order: int = 2,
                 savgol_window_length: int = 17, number_of_shells: int = 1):
        if isinstance(rdf_data, Computation):
            self.rdf_data = rdf_data
        else:
            self.rdf_data = self.experiment.run.RadialDistributionFunction(plot=False)
        self.args = Args(savgol_order=savgol_order, savgol_window_length=savgol_window_length,
                         number_of_shells=number_of_shells)
        self.plot = plot

    def _get_density(self, species: str) -> float:
        """:208-225 -- particles of the first species / volume in nm^3."""
        sp = species.split("_")
        volume_si = self.experiment.volume * self.experiment.units.length**3
        return self.experiment.species[sp[0]].n_particles / (volume_si / 1e-9**3)

    def run_calculator(self):
        a = self.args
        for selected_species, vals in self.rdf_data.data_dict.items():
            radii = np.array(vals["x"]).astype(float)[1:]   # drops the nan bin (Q2)
            rdf = np.array(vals["y"]).astype(float)[1:]
            density = self._get_density(selected_species)
            integral = 4 * np.pi * density * cumulative_trapezoid(
                y=radii[1:] ** 2 * rdf[1:], x=radii[1:])
            filtered = savgol_filter(rdf, a.savgol_window_length, a.savgol_order)
            peaks = find_peaks(filtered, height=1.0)[0]
            if len(peaks) < a.number_of_shells + 1:
                raise CannotPerformThisAnalysis(
                    "We have detected too few peaks for this analysis; the g(r) may be too "
                    "noisy or the simulation too small")
            data = {"r": radii[1:].tolist(), "cn": integral.tolist()}
            for i in range(a.number_of_shells):
                lo, hi = golden_section_search([radii, rdf], radii[peaks[i + 1]], radii[peaks[i]])
                idx = [int(np.where(radii == v)[0][0]) for v in (lo, hi)]
                lower, upper = integral[idx[0]], integral[idx[1]]
                data[f"CN_{i + 1}"] = float(np.mean([lower, upper]))
                data[f"CN_{i + 1}_error"] = float(np.std([lower, upper]) / np.sqrt(2))
            self.queue_data(data=data, subjects=selected_species.split("_"))
