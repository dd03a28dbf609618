"""
Oracle restatement of CoordinationNumbers post-processing (TEST INFRASTRUCTURE).

Follows:
  mdsuite/calculators/coordination_number_calculation.py:59-81   (_integrate_rdf)
  mdsuite/calculators/coordination_number_calculation.py:208-359 (density, peaks, minima, CN)
  mdsuite/utils/meta_functions.py:327-437 (apply_savgol_filter, closest_point,
                                           golden_section_search)
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import cumulative_trapezoid
from scipy.signal import find_peaks, savgol_filter

golden_ratio = 1.618033988749895  # utils/units.py:42


class CannotPerformThisAnalysis(Exception):
    pass


def closest_point(data, value):
    return min(data, key=lambda x: abs(x - value))


def golden_section_search(data, a, b, tol=1e-5, h=None, c=None, d=None, fc=None, fd=None):
    phi_a = 1 / golden_ratio
    phi_b = 1 / (golden_ratio**2)
    (a, b) = (min(a, b), max(a, b))
    if h is None:
        h = b - a
    if h <= tol:
        return a, b
    if c is None:
        c = closest_point(data[0], a + phi_b * h)
    if d is None:
        d = closest_point(data[0], a + phi_a * h)
    if fc is None:
        fc = data[1][np.where(data[0] == c)]
    if fd is None:
        fd = data[1][np.where(data[0] == d)]
    if fc < fd:
        return golden_section_search(data, a, d, tol, h * phi_a, c=None, fc=None, d=c, fd=fc)
    return golden_section_search(data, c, b, tol, h * phi_a, c=d, fc=fd, d=None, fd=None)


def integrate_rdf(radii_data, rdf_data, density):
    integral_data = cumulative_trapezoid(
        y=radii_data[1:] ** 2 * rdf_data[1:], x=radii_data[1:]
    )
    return 4 * np.pi * density * integral_data


def coordination_numbers(rdf_data_dict: dict, n_particles: dict, volume_nm3: float,
                         savgol_order=2, savgol_window_length=17, number_of_shells=1):
    """run_calculator :334-359.  Returns {"A_B": {"r","cn","CN_1","CN_1_error",...}}."""
    out = {}
    for selected_species, vals in rdf_data_dict.items():
        radii = np.array(vals["x"]).astype(float)[1:]
        rdf = np.array(vals["y"]).astype(float)[1:]
        sp = selected_species.split("_")
        density = n_particles[sp[0]] / volume_nm3  # _get_density :220-225
        integral_data = integrate_rdf(radii, rdf, density)
        filtered = savgol_filter(rdf, savgol_window_length, savgol_order)
        peaks = find_peaks(filtered, height=1.0)[0]
        if len(peaks) < number_of_shells + 1:
            raise CannotPerformThisAnalysis("Not enough peaks")
        data = {"r": radii[1:].tolist(), "cn": integral_data.tolist()}
        for i in range(number_of_shells):
            rng = golden_section_search([radii, rdf], radii[peaks[i + 1]], radii[peaks[i]])
            idx = [int(np.where(radii == rng[j])[0][0]) for j in range(2)]
            lower, upper = integral_data[idx[0]], integral_data[idx[1]]
            data[f"CN_{i + 1}"] = np.mean([lower, upper])
            data[f"CN_{i + 1}_error"] = np.std([lower, upper]) / np.sqrt(2)
        out[selected_species] = data
    return out


# --- potential_of_mean_force.py:183-349 ---------------------------------------
def potential_of_mean_force(rdf_data_dict: dict, temperature: float, savgol_order=2,
                            savgol_window_length=17, number_of_shells=1):
    boltzmann_constant = 1.380649e-23
    out = {}
    for selected_species, vals in rdf_data_dict.items():
        radii = np.array(vals["x"]).astype(float)[1:]
        rdf = np.array(vals["y"]).astype(float)[1:]
        with np.errstate(divide="ignore", invalid="ignore"):
            pomf = -1 * boltzmann_constant * temperature * np.log(rdf)
        pomf = pomf * 6.242e8
        filtered = savgol_filter(pomf, savgol_window_length, savgol_order)
        peaks = find_peaks(filtered)[0]
        if len(peaks) < number_of_shells + 1:
            raise ValueError("Not enough peaks")
        data = {"r": radii[1:].tolist(), "pomf": pomf.tolist()}
        for i in range(number_of_shells):
            rng = golden_section_search([radii, pomf], radii[peaks[i + 1]], radii[peaks[i]])
            idx = [int(np.where(radii == rng[j])[0][0]) for j in range(2)]
            lower, upper = pomf[idx[0]], pomf[idx[1]]
            data[f"POMF_{i + 1}"] = np.mean([lower, upper])
            data[f"POMF_{i + 1}_error"] = np.std([lower, upper]) / np.sqrt(2)
        out[selected_species] = data
    return out


# --- kirkwood_buff_integrals.py:160-203 ----------------------------------------
def kirkwood_buff_integral(rdf_data_dict: dict, savgol_order=2, savgol_window_length=17):
    out = {}
    for selected_species, vals in rdf_data_dict.items():
        radii = np.array(vals["x"]).astype(float)[1:]
        rdf = np.array(vals["y"]).astype(float)[1:]
        filtered = savgol_filter(rdf, savgol_window_length, savgol_order)
        integral = cumulative_trapezoid(y=(filtered[1:] - 1) * radii[1:] ** 2, x=radii[1:])
        out[selected_species] = {"r": radii[1:].tolist(),
                                 "kb_integral": (4 * np.pi * integral).tolist()}
    return out
