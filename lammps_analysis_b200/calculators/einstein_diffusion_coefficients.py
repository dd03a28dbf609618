"""EinsteinDiffusionCoefficients: windowed single-origin MSD -> linear fit -> D = slope / 6.

Mirrors mdsuite/calculators/einstein_diffusion_coefficients.py (Args :50-61, __call__
:112-166, ensemble_operation :168-190, fit_diff_coeff :192-215, run_calculator :217-248) and
mdsuite/utils/calculator_helper_methods.py:41-107 (fit_einstein_curve).  The per-window
Python loop of the reference becomes launch parameters of ``mdk_msd_windowed``: every window
of a planned batch is one launch; atoms shard across ranks.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Union

import numpy as np
from scipy.interpolate import UnivariateSpline

from .. import distributed as D
from .. import kernels as K
from .. import trace
from ..engine import msd_series, plan_windows
from ..store import join_path
from .calculator import TrajectoryCalculator, call


@dataclass
class Args:
    data_range: int
    correlation_time: int
    atom_selection: object
    tau_values: object
    molecules: bool
    species: list
    fit_range: int


def fit_einstein_curve(x_data: np.ndarray, y_data: np.ndarray, fit_max_index: int):
    """calculator_helper_methods.py:41-107.  The reference calls scipy ``curve_fit`` on the
    line m*x + a for every end index i in (start+2 .. len); a straight line has a closed-form
    least-squares solution (and covariance s^2 (X^T X)^-1, s^2 = RSS / (n - 2)), so all fits
    come from prefix sums.  Returns (popt, pcov, gradients, gradient_errors)."""
    x_data = np.asarray(x_data, dtype=float)
    y_data = np.asarray(y_data, dtype=float)
    spline = UnivariateSpline(x_data, y_data, s=0, k=4)
    deriv = spline.derivative(n=2)(x_data)
    deriv[np.abs(deriv) < 1e-5] = 0
    start = int(np.argmin(np.abs(deriv)))
    n_pts = len(y_data)
    if start + 2 >= n_pts:
        return [], [], [], []
    # normalise for conditioning, fit, scale back
    xs = np.max(np.abs(x_data[start:])) or 1.0
    ys = np.max(np.abs(y_data[start:])) or 1.0
    x = x_data[start:] / xs
    y = y_data[start:] / ys
    x0, y0 = x[0], y[0]
    dx, dy = x - x0, y - y0        # shift to the first point: exact, reduces cancellation
    c1 = np.arange(1, len(x) + 1, dtype=float)
    sx, sy = np.cumsum(dx), np.cumsum(dy)
    sxx, sxy, syy = np.cumsum(dx * dx), np.cumsum(dx * dy), np.cumsum(dy * dy)
    ends = np.arange(start + 2, n_pts)          # slice x[start:i], n = i - start points
    k = ends - start - 1                        # index of the last included point
    n = c1[k]
    Sxx = sxx[k] - sx[k] ** 2 / n
    Sxy = sxy[k] - sx[k] * sy[k] / n
    Syy = syy[k] - sy[k] ** 2 / n
    with np.errstate(divide="ignore", invalid="ignore"):
        m = Sxy / Sxx
        a = (sy[k] - m * sx[k]) / n + y0 - m * x0
        rss = np.maximum(Syy - m * Sxy, 0.0)
        s2 = np.where(n > 2, rss / (n - 2), np.inf)
        var_m = s2 / Sxx
        mean_x = sx[k] / n + x0
        var_a = s2 * (1.0 / n + mean_x**2 / Sxx)
        cov_ma = -s2 * mean_x / Sxx
    m_out = m * ys / xs
    a_out = a * ys
    gradients = m_out.tolist()
    gradient_errors = (np.sqrt(var_m) * ys / xs).tolist()
    popt, pcov = [], []
    sel = np.nonzero(ends == fit_max_index)[0]
    if len(sel):
        j = sel[0]
        popt = np.array([m_out[j], a_out[j]])
        pcov = np.array([[var_m[j] * (ys / xs) ** 2, cov_ma[j] * ys * ys / xs],
                         [cov_ma[j] * ys * ys / xs, var_a[j] * ys * ys]])
    return popt, pcov, gradients, gradient_errors


class EinsteinDiffusionCoefficients(TrajectoryCalculator):
    analysis_name = "Einstein_Self_Diffusion_Coefficients"
    loaded_property = "Unwrapped_Positions"
    scale_function = {"linear": {"scale_factor": 150}}
    result_keys = ["diffusion_coefficient", "uncertainty", "gradient", "intercept"]
    result_series_keys = ["time", "msd", "gradients", "gradient_errors"]

    @call
    def __call__(self, plot: bool = True, species: list = None, data_range: int = 100,
                 correlation_time: int = 1, atom_selection: Union[slice, dict] = np.s_[:],
                 molecules: bool = False, tau_values: Union[int, list, slice] = np.s_[:],
                 fit_range: int = -1):
        if species is None:
            species = list(self.experiment.species)
        if fit_range == -1:
            fit_range = int(data_range - 1)
        self.args = Args(data_range=data_range, correlation_time=correlation_time,
                         atom_selection=atom_selection, tau_values=tau_values,
                         molecules=molecules, species=species, fit_range=fit_range)
        self.plot = plot
        self.system_property = False

    def check_input(self):
        if self.args.molecules:
            raise NotImplementedError("molecule diffusion needs the molecule-mapping subsystem")
        self._run_dependency_check()

    # -- hot path ----------------------------------------------------------------------------------
    def compute_msd(self, species: str):
        """Returns (msd_sum float64 [n_tau] on the host, count) for one species."""
        path = join_path(species, self.loaded_property)
        self._prepare_managers([path])
        # only this rank's atom block is uploaded (it is already resident when the unwrap
        # transformation has just produced it)
        # (and the kernel follows the row blocks of that transformation / of the upload)
        import torch

        traj, n_atoms, shard, offset, blocks = self._device_row_blocks(path, species)
        launches = plan_windows(self.plan.as_dict(), self.args.data_range,
                                self.args.correlation_time, n_atoms)
        # kernels, reduction and read-back all run on the consumer stream: the current stream
        # may still hold the unwrap pipeline of the NEXT species, and a read-back queued there
        # would only complete after its last block
        cur, side = torch.cuda.current_stream(), self._side_stream()
        with torch.cuda.stream(side):
            msd, count = msd_series(traj, launches, self.args.data_range,
                                    self.args.correlation_time, self.args.tau_values,
                                    a_shard=shard, row_offset=offset, blocks=blocks)
            D.all_reduce_sum_([msd])
            trace.mark(f"Einstein[{species}] MSD kernels enqueued")
            host, = K.read_back(msd)
        cur.wait_stream(side)
        trace.mark(f"Einstein[{species}] MSD on the host")
        return host, count

    # -- :192-215 --------------------------------------------------------------------------------------
    def fit_diff_coeff(self, msd_sum: np.ndarray, count: int, time: np.ndarray) -> dict:
        units = self.experiment.units
        msd = np.array(msd_sum) / count
        msd = msd * units.length**2
        t = np.array(time) * units.time
        popt, pcov, gradients, gradient_errors = fit_einstein_curve(t, msd, self.args.fit_range)
        if len(popt) == 0:
            raise ValueError("fit_range lies before the linear regime found by the spline "
                             "(the reference fails here as well)")
        error = np.sqrt(np.diag(pcov))[0]
        return {"diffusion_coefficient": 1 / 6 * popt[0], "uncertainty": 1 / 6 * error,
                "gradient": popt[0], "intercept": popt[1], "time": t.tolist(),
                "msd": msd.tolist(), "gradients": (np.array(gradients) / 6).tolist(),
                "gradient_errors": (np.array(gradient_errors) / 6).tolist()}

    def run_calculator(self):
        trace.mark("Einstein start")
        self.check_input()
        trace.mark("Einstein dependencies resolved")
        for species in self.args.species:
            time = self._handle_tau_values()
            msd_sum, count = self.compute_msd(species)
            self.queue_data(data=self.fit_diff_coeff(msd_sum, count, time), subjects=[species])
            trace.mark(f"Einstein[{species}] fit done")
