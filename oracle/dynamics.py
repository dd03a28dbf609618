"""
Oracle restatement of the Einstein / Green-Kubo calculators (TEST INFRASTRUCTURE).

Follows:
  mdsuite/calculators/einstein_diffusion_coefficients.py:168-248
  mdsuite/calculators/green_kubo_self_diffusion_coefficients.py:179-206, 270-337
  mdsuite/calculators/green_kubo_ionic_conductivity.py:167-231, 286-310
  mdsuite/calculators/trajectory_calculator.py:196-228 (_handle_tau_values)
  mdsuite/utils/calculator_helper_methods.py:41-107 (fit_einstein_curve)
  tensorflow_probability/python/stats/sample_stats.py auto_correlation (unpinned
  third-party dependency; restated from its published algorithm)
"""
from __future__ import annotations

import numpy as np
from scipy.integrate import cumulative_trapezoid
from scipy.interpolate import UnivariateSpline
from scipy.optimize import curve_fit

from oracle.planner import iter_batches, iter_ensembles

elementary_charge = 1.602176634e-19  # utils/units.py:31
boltzmann_constant = 1.380649e-23  # utils/units.py:32


# --- trajectory_calculator.py:196-228 ----------------------------------------
def handle_tau_values(tau_values, data_range, time_step, sample_rate):
    """Returns (tau_values ndarray, data_range, data_resolution, times)."""
    if isinstance(tau_values, (int, np.integer)):
        data_resolution = int(tau_values)
        tau_values = np.linspace(0, data_range - 1, tau_values, dtype=int)
    if isinstance(tau_values, (list, np.ndarray)):
        data_resolution = len(tau_values)
        data_range = tau_values[-1] + 1
    if isinstance(tau_values, slice):
        tau_values = np.linspace(0, data_range - 1, data_range, dtype=int)[tau_values]
        data_resolution = len(tau_values)
    times = np.asarray(tau_values) * time_step * sample_rate
    return np.asarray(tau_values), int(data_range), data_resolution, times


# --- tfp.stats.auto_correlation(x, axis=1, normalize=False, center=False) ----
def tfp_auto_correlation(x: np.ndarray) -> np.ndarray:
    """x: (A, N, D) float64.  FFT-based, unbiased (divide lag m by N-m)."""
    x = np.asarray(x, dtype=np.float64)
    x_rot = np.moveaxis(x, 1, -1)  # rotate `axis` to the end
    n = x_rot.shape[-1]
    target_length = int(2.0 ** np.ceil(np.log(float(n) * 2) / np.log(2.0)))
    xc = x_rot.astype(np.complex128)
    fx = np.fft.fft(xc, n=target_length, axis=-1)  # zero pad to target_length
    shifted_product = np.fft.ifft(fx * np.conj(fx), axis=-1).real
    shifted_product = shifted_product[..., :n]
    denominator = n - np.arange(0.0, float(n))
    out = shifted_product / denominator
    return np.moveaxis(out, -1, 1)


# --- einstein_diffusion_coefficients.py:168-248 ------------------------------
def einstein_msd(data: np.ndarray, plan: dict, data_range: int, correlation_time: int,
                 tau_values: np.ndarray):
    """data: (A, T, 3) unwrapped positions (fp32-valued), processed in float64.

    Returns (msd_sum[n_tau] float64, count int) *before* the division -- the
    division `/ count`, unit scaling and the fit are in ``einstein_finish``.
    """
    msd_array = np.zeros(len(tau_values))
    count = 0
    for atom_sel, start, stop, data_size in iter_batches(plan):
        batch = np.asarray(data[atom_sel, start:stop], dtype=np.float64)
        for s, e in iter_ensembles(data_size, data_range, correlation_time):
            ensemble = batch[:, s:e]
            if not ensemble.shape[1] == data_range:  # :240-241
                continue
            # ensemble_operation :168-190
            msd = (ensemble[:, tau_values] - ensemble[:, None, 0]) ** 2
            count += msd.shape[0]
            msd_array += msd.sum(axis=0).sum(axis=-1)
            count += 1  # :244
    return msd_array, count


def fit_einstein_curve(x_data, y_data, fit_max_index):
    """utils/calculator_helper_methods.py:41-107."""
    popt, pcov = [], []

    def func(x, m, a):
        return m * x + a

    spline_data = UnivariateSpline(x_data, y_data, s=0, k=4)
    derivatives = spline_data.derivative(n=2)(x_data)
    derivatives[abs(derivatives) < 1e-5] = 0
    start_index = np.argmin(abs(derivatives))
    gradients, gradient_errors = [], []
    for i in range(start_index + 2, len(y_data)):
        popt_temp, pcov_temp = curve_fit(
            func, xdata=x_data[start_index:i], ydata=y_data[start_index:i]
        )
        gradients.append(popt_temp[0])
        gradient_errors.append(np.sqrt(np.diag(pcov_temp))[0])
        if i == fit_max_index:
            popt, pcov = popt_temp, pcov_temp
    return popt, pcov, gradients, gradient_errors


def einstein_finish(msd_sum, count, times, units_length, units_time, fit_range):
    """fit_diff_coeff :192-215."""
    msd = np.array(msd_sum, dtype=float) / count
    msd = msd * units_length**2
    time = np.array(times, dtype=float) * units_time
    fit_values, covariance, gradients, gradient_errors = fit_einstein_curve(
        x_data=time, y_data=msd, fit_max_index=fit_range
    )
    error = np.sqrt(np.diag(covariance))[0]
    return {
        "diffusion_coefficient": 1 / 6.0 * fit_values[0],
        "uncertainty": 1 / 6.0 * error,
        "gradient": fit_values[0],
        "intercept": fit_values[1],
        "time": time.tolist(),
        "msd": msd.tolist(),
        "gradients": (np.array(gradients) / 6).tolist(),
        "gradient_errors": (np.array(gradient_errors) / 6).tolist(),
    }


# --- green_kubo_self_diffusion_coefficients.py:179-206, 302-337 --------------
def gk_diffusion_acf(data: np.ndarray, plan: dict, data_range: int, correlation_time: int,
                     time: np.ndarray, units_length: float, units_time: float):
    """data: (A, T, 3) velocities.  Returns (acf_sum[N], count, sigmas[W, N-1])."""
    acf_array = np.zeros(data_range)
    count = 0
    sigmas = []
    scale = units_length**2 / units_time**2
    for atom_sel, start, stop, data_size in iter_batches(plan):
        batch = np.asarray(data[atom_sel, start:stop], dtype=np.float64)
        for s, e in iter_ensembles(data_size, data_range, correlation_time):
            ensemble = batch[:, s:e]
            if not ensemble.shape[1] == data_range:  # :329-330
                continue
            vacf = scale * tfp_auto_correlation(ensemble)
            count += vacf.shape[0]
            acf_array += vacf.sum(axis=0).sum(axis=-1)
            sigmas.append(cumulative_trapezoid(vacf.mean(axis=0).sum(axis=-1), x=time))
            count += 1  # :334
    return acf_array, count, np.array(sigmas)


def gk_diffusion_finish(acf_sum, count, sigmas, time, integration_range):
    """postprocessing :270-300."""
    acf = np.array(acf_sum, dtype=float) / count
    sigma = cumulative_trapezoid(acf, x=time)
    sigma_SEM = np.std(sigmas, axis=0) / np.sqrt(len(sigmas))
    return {
        "diffusion_coefficient": [1 / 3 * sigma[integration_range - 1]],
        "uncertainty": [1 / 3 * sigma_SEM[integration_range - 1]],
        "time": np.asarray(time).tolist(),
        "acf": acf.tolist(),
        "integral": sigma.tolist(),
        "integral_uncertainty": sigma_SEM.tolist(),
    }


# --- green_kubo_ionic_conductivity.py:167-231, 286-310 -----------------------
def gk_ionic_prefactor(units_length, units_time, temperature, volume):
    numerator = (elementary_charge**2) * (units_length**2)
    denominator = (
        3 * boltzmann_constant * temperature * volume * units_length**3 * units_time
    )
    return numerator / denominator


def gk_ionic_acf(current: np.ndarray, plan: dict, data_range: int, correlation_time: int,
                 tau_values: np.ndarray, time: np.ndarray):
    """current: (1, T, 3) Observables/Ionic_Current.  No short-window filter (Q8).

    Q7: the system-property slice ``np.s_[start:stop]`` acts on axis 0 of the
    (1, T, 3) dataset (data_manager.py:204-205), so every batch sees the *whole*
    series; ``data_size`` is still the planned batch size.
    """
    acf_array = np.zeros((data_range,))
    count = 0
    sigmas = []
    for _atom_sel, start, stop, data_size in iter_batches(plan, system=True):
        batch = np.asarray(current[start:stop], dtype=np.float64)  # axis-0 slice (Q7)
        if batch.shape[0] == 0:
            # tf.squeeze(axis=0) on a (0, N) tensor raises upstream: system
            # observables only work with a single batch (Q7).
            raise ValueError("system observable requested with more than one batch (Q7)")
        for s, e in iter_ensembles(data_size, data_range, correlation_time):
            ensemble = batch[:, s:e][:, tau_values]
            jacf = tfp_auto_correlation(ensemble)
            jacf = jacf.sum(axis=-1)[0]
            sigmas.append(cumulative_trapezoid(jacf, x=time))
            acf_array += jacf
            count += 1
    return acf_array, count, np.array(sigmas)


def gk_ionic_finish(acf_sum, count, sigmas, time, prefactor, integration_range):
    acf = np.array(acf_sum, dtype=float) / count
    sigma = cumulative_trapezoid(acf, x=time)
    sigma_SEM = np.std(sigmas, axis=0) / np.sqrt(len(sigmas))
    return {
        "ionic_conductivity": [prefactor * sigma[integration_range - 1]],
        "uncertainty": [prefactor * sigma_SEM[integration_range - 1]],
        "time": np.asarray(time).tolist(),
        "acf": acf.tolist(),
        "integral": sigma.tolist(),
        "integral_uncertainty": sigma_SEM.tolist(),
    }


# --- einstein_helfand_ionic_conductivity.py:167-258 ---------------------------
def eh_ionic_prefactor(units_length, units_time, temperature, volume):
    numerator = (units_length**2) * (elementary_charge**2)
    denominator = units_time * volume * units_length**3 * temperature * boltzmann_constant
    return numerator / denominator


def eh_ionic_msd(dipole: np.ndarray, plan: dict, data_range: int, correlation_time: int,
                 tau_values: np.ndarray, prefactor: float):
    """dipole: (1, T, 3) Observables/Translational_Dipole_Moment.  Returns the averaged msd
    (ensemble_operation :192-209, _apply_averaging_factor :188-190).  Same axis-0 slicing of
    the system observable as the ionic ACF (Q7)."""
    msd_array = np.zeros(len(tau_values))
    ensemble_loop = None
    for _atom_sel, start, stop, data_size in iter_batches(plan, system=True):
        batch = np.asarray(dipole[start:stop], dtype=np.float64)
        if batch.shape[0] == 0:
            raise ValueError("system observable requested with more than one batch (Q7)")
        windows = list(iter_ensembles(data_size, data_range, correlation_time))
        ensemble_loop = len(windows)
        for s, e in windows:
            ensemble = batch[:, s:e]
            msd = (ensemble[:, tau_values] - ensemble[:, None, 0]) ** 2
            msd_array += (prefactor * msd.sum(axis=2))[0, :]
    return msd_array / (int(plan["n_batches"]) * ensemble_loop)


# --- green_kubo_thermal_conductivity.py:152-248, green_kubo_viscosity.py:146-227 ------------
_trapz = getattr(np, "trapezoid", None) or np.trapz   # np.trapz was renamed in NumPy 2


def gk_thermal_prefactor(units, temperature, volume, data_range):
    denominator = 3 * (data_range - 1) * temperature**2 * units.boltzmann * volume
    return (1 / denominator) * (units.energy / units.length / units.time)


def gk_viscosity_prefactor(units, temperature, volume, data_range):
    denominator = 3 * (data_range - 1) * temperature * units.boltzmann * volume
    return (1 / denominator) * (units.pressure**2 * units.length**3 * units.time / units.energy)


def gk_flux(flux: np.ndarray, plan: dict, data_range: int, correlation_time: int,
            time: np.ndarray, integration_range: int, prefactor: float, value_key: str):
    """flux: (1, T, 3) system observable.  Per window: jacf = data_range * sum_dims
    tfp.auto_correlation(window, axis=0); self.jacf += jacf; sigma.append(trapz(...)).  The
    result is prefactor * sigma[0] with "uncertainty" prefactor * sigma[1] (the first two
    windows), and the stored acf is the un-averaged sum (reference behaviour, restated as is)."""
    jacf_sum = np.zeros(data_range)
    sigma = []
    for _atom_sel, start, stop, data_size in iter_batches(plan, system=True):
        batch = np.asarray(flux[start:stop], dtype=np.float64)
        if batch.shape[0] == 0:
            raise ValueError("system observable requested with more than one batch (Q7)")
        for s, e in iter_ensembles(data_size, data_range, correlation_time):
            ensemble = batch[0, s:e]                       # (N, 3)
            acf = tfp_auto_correlation(ensemble[None])[0]  # (N, 3), axis 0 correlated
            jacf = data_range * acf.sum(axis=-1)
            jacf_sum += jacf
            sigma.append(_trapz(jacf[:integration_range], x=time[:integration_range]))
    result = prefactor * np.array(sigma)
    return {value_key: result[0], "uncertainty": result[1], "time": np.asarray(time).tolist(),
            "acf": jacf_sum.tolist()}


# --- einstein_helfand_thermal_conductivity.py:152-229 ----------------------------------------
def eh_thermal_prefactor(units, temperature, volume):
    denominator = volume * temperature * units.boltzmann
    return (1 / denominator) * (units.energy / units.length / units.time / units.temperature)


# --- green_kubo_viscosity_flux.py:149-225 -----------------------------------------------------
def gk_viscosity_flux_prefactor(units, temperature, volume, data_range):
    denominator = 3 * (data_range - 1) * temperature * units.boltzmann
    return (volume / denominator) * (units.pressure**2 * units.volume * units.time / units.energy)


def gk_viscosity_flux(flux: np.ndarray, plan: dict, data_range: int, correlation_time: int,
                      time: np.ndarray, integration_range: int, prefactor: float):
    """As ``gk_flux`` except for the stored series: ``self.jacf += jacf[data_range - 1:]`` adds
    ONE value per window to every element (:199), then ``self.jacf /= max(self.jacf)`` (:167-175)."""
    jacf_sum = np.zeros(data_range)
    sigma = []
    for _atom_sel, start, stop, data_size in iter_batches(plan, system=True):
        batch = np.asarray(flux[start:stop], dtype=np.float64)
        if batch.shape[0] == 0:
            raise ValueError("system observable requested with more than one batch (Q7)")
        for s, e in iter_ensembles(data_size, data_range, correlation_time):
            ensemble = batch[0, s:e]
            acf = tfp_auto_correlation(ensemble[None])[0]
            jacf = data_range * acf.sum(axis=-1)
            jacf_sum += jacf[int(data_range - 1):]
            sigma.append(_trapz(jacf[:integration_range], x=time[:integration_range]))
    with np.errstate(invalid="ignore", divide="ignore"):
        jacf_sum = jacf_sum / max(jacf_sum)
    result = prefactor * np.array(sigma)
    return {"viscosity": result[0], "uncertainty": result[1], "time": np.asarray(time).tolist(),
            "acf": jacf_sum.tolist()}
