"""Calculator shell: the ``@call`` cache/run/store protocol and the trajectory-calculator base
(dependency resolution, tau handling, batch plan).

Mirrors mdsuite/calculators/calculator.py:52-317 and
mdsuite/calculators/trajectory_calculator.py:117-297.  Plotting is dropped (``plot=`` is
accepted and ignored: bokeh / matplotlib are outside the hot path).
"""
from __future__ import annotations

import functools
import logging
from collections import OrderedDict
from typing import List

import numpy as np

from ..planner import BatchPlan, plan_batches
from ..project import Computation, args_to_parameters, subjects_key
from ..store import join_path
from .. import trace

log = logging.getLogger("mdsuite_b200")


def call(func):
    """Decorator of ``Calculator.__call__`` (calculator.py:52-148): per experiment, fill
    ``self.args``; return the stored computation when one with identical args and experiment
    version exists; otherwise run, store, re-apply the user args and return the stored
    object.  Project-level calls return ``{experiment_name: Computation}``."""

    @functools.wraps(func)
    def inner(self, *args, **kwargs):
        return_dict = self.experiment is None
        out = {}
        for experiment in self.experiments:
            cls = self.__class__(experiment=experiment)
            func(cls, *args, **kwargs)
            data = cls.get_computation_data()
            if data is None:
                trace.mark(f"{cls.analysis_name}: cache miss, running")
                cls.prepare_db_entry()
                cls.save_computation_args()
                cls.run_analysis()
                trace.mark(f"{cls.analysis_name}: analysis done")
                cls.save_db_data()
                func(cls, *args, **kwargs)
                data = cls.get_computation_data()
                trace.mark(f"{cls.analysis_name}: stored and read back")
            out[experiment.name] = data
        return out if return_dict else out[self.experiment.name]

    return inner


class Calculator:
    analysis_name = "Calculator"
    result_keys: list = None
    result_series_keys: list = None

    def __init__(self, experiment=None, experiments=None, **kwargs):
        self.experiment = experiment
        if experiments is None:
            experiments = [experiment]
        self.experiments = experiments
        self.args = None
        self.plot = False
        self._queued_data: List = []
        self._queued_metadata: dict = {}
        self._saved_parameters = None

    # -- cache protocol (database/calculator_database.py:91-248) ---------------------------------
    def _parameters(self) -> dict:
        return args_to_parameters(self.args, self.experiment.version)

    def get_computation_data(self) -> Computation:
        return self.experiment.project.find_computation(self.analysis_name, self.experiment.name,
                                                        self._parameters())

    def prepare_db_entry(self):
        self._queued_data = []
        self._queued_metadata = {}

    def save_computation_args(self):
        # stored *before* the run, i.e. with the user's -1 / None defaults unresolved
        self._saved_parameters = self._parameters()

    def queue_data(self, data: dict, subjects: list):
        self._queued_data.append((subjects_key(list(subjects)), data))

    def queue_metadata(self, **items):
        """Run diagnostics stored next to the result (``Computation.metadata``), outside the
        reference's ``data_dict``."""
        self._queued_metadata.update(items)

    def save_db_data(self):
        results = OrderedDict(self._queued_data)
        self.experiment.project.store_computation(self.analysis_name, self.experiment.name,
                                                  self._saved_parameters, results,
                                                  metadata=self._queued_metadata)

    def run_analysis(self):
        """calculator.py:310-317."""
        self.run_calculator()

    def run_calculator(self):
        raise NotImplementedError


class TrajectoryCalculator(Calculator):
    """trajectory_calculator.py:52-406 without the tf.data plumbing."""

    loaded_property: str = None       # dataset name, e.g. "Unwrapped_Positions"
    system_property: bool = False
    scale_function: dict = None
    dependency: str = None

    def __init__(self, experiment=None, experiments=None, **kwargs):
        super().__init__(experiment=experiment, experiments=experiments, **kwargs)
        self.data_resolution = None
        self.plan: BatchPlan = None

    # -- dependencies (:117-194) ---------------------------------------------------------------------
    def _run_dependency_check(self):
        """The loaded property is produced by its transformation when it is missing -- or when
        the experiment has grown since it was written (the transformation then extends it)."""
        store = self.experiment.store
        n_cfg = self.experiment.number_of_configurations

        def stale(path):
            return not store.check_existence(path) or store.shape(path)[1] < n_cfg

        if self.system_property:
            if stale(join_path("Observables", self.loaded_property)):
                self._resolve_dependencies(self.loaded_property)
            return
        for sp in self.args.species:
            if stale(join_path(sp, self.loaded_property)):
                self._resolve_dependencies(self.loaded_property)
                break

    def _resolve_dependencies(self, dependency: str):
        run = self.experiment.run
        if dependency == "Unwrapped_Positions":
            # _unwrap_choice (:181-194): box images available -> indices, else box hopping
            first = next(iter(self.experiment.species))
            if self.experiment.store.check_existence(join_path(first, "Box_Images")):
                run.UnwrapViaIndices()
            else:
                run.CoordinateUnwrapper()
        elif dependency == "Ionic_Current":
            run.IonicCurrent()
        elif dependency == "Translational_Dipole_Moment":
            run.TranslationalDipoleMoment()
        elif dependency == "Integrated_Heat_Current":   # transformations_reference.py:27-34
            run.IntegratedHeatCurrent()
        elif dependency == "Thermal_Flux":
            run.ThermalFlux()
        elif dependency == "Momentum_Flux":
            run.MomentumFlux()
        else:
            raise KeyError("Data not in database and cannot be generated.")  # :171-174

    # -- tau values (:196-228) ----------------------------------------------------------------------------
    def _handle_tau_values(self) -> np.ndarray:
        tau = self.args.tau_values
        if isinstance(tau, (int, np.integer)):
            self.data_resolution = int(tau)
            self.args.tau_values = np.linspace(0, self.args.data_range - 1, int(tau), dtype=int)
        if isinstance(self.args.tau_values, (list, np.ndarray)):
            self.data_resolution = len(self.args.tau_values)
            self.args.data_range = int(self.args.tau_values[-1] + 1)
        if isinstance(self.args.tau_values, slice):
            self.args.tau_values = np.linspace(0, self.args.data_range - 1, self.args.data_range,
                                               dtype=int)[self.args.tau_values]
            self.data_resolution = len(self.args.tau_values)
        return (np.asarray(self.args.tau_values) * self.experiment.time_step
                * self.experiment.sample_rate)

    # -- device residency of one species' rows -------------------------------------------------------------
    def _device_rows(self, path: str, species: str):
        """The atom rows of ``path`` this rank processes, on the device.

        Returns (traj, n_atoms, a_shard, row_offset): ``traj`` holds the selected atoms
        [a_shard[0], a_shard[1]) of the species (indices into the atom selection, or global atom
        indices when everything is selected), ``n_atoms`` is the size of the whole selection
        (what the reference's plan and normalisation see), ``row_offset`` the selection index of
        traj's first row.  Atoms shard across ranks along the store's row blocks."""
        store = self.experiment.store
        sel = self.args.atom_selection
        lo, hi = store.owned_rows(path)
        if isinstance(sel, dict):
            idx = np.asarray(sel[species])
            i0, i1 = 0, len(idx)
            if store.is_sharded(path):
                if np.any(np.diff(idx) < 0):
                    raise ValueError("atom_selection must be sorted when atoms shard across ranks")
                i0, i1 = (int(v) for v in np.searchsorted(idx, [lo, hi]))
            return store.device(path, row_index=idx[i0:i1]), len(idx), (i0, i1), i0
        return store.device(path, rows=(lo, hi)), store.shape(path)[0], (lo, hi), lo

    def _device_row_blocks(self, path: str, species: str):
        """``_device_rows`` plus the row blocks the device copy arrives in (store.device_blocks):
        (traj, n_atoms, a_shard, row_offset, blocks).  ``blocks`` is None when the rows are a
        fancy selection (uploaded in one piece)."""
        store = self.experiment.store
        if isinstance(self.args.atom_selection, dict):
            return self._device_rows(path, species) + (None,)
        import torch

        lo, hi = store.owned_rows(path)
        traj, blocks = store.device_blocks(path)
        # a block without an event was produced by work already queued on the current stream
        # (or is complete): one event for "everything queued so far" stands in, so that the
        # consumer stream never has to wait for the whole current stream -- which, behind an
        # unwrap pipeline, would mean waiting for its last block
        fence = None
        out = []
        for b0, b1, ev in blocks:
            if ev is None:
                if fence is None:
                    fence = torch.cuda.Event()
                    fence.record()
                ev = fence
            out.append((b0, b1, ev))
        return traj, store.shape(path)[0], (lo, hi), lo, out

    @staticmethod
    def _side_stream():
        """Stream on which a calculator consumes row blocks: the unwrap pipeline queues its
        kernels on the current stream, so a consumer on the same stream would only start after
        the last block; on its own stream it follows the blocks as they complete."""
        import torch

        if TrajectoryCalculator._consumer_stream is None:
            TrajectoryCalculator._consumer_stream = torch.cuda.Stream()
        return TrajectoryCalculator._consumer_stream

    _consumer_stream = None

    # -- batch plan (:243-297) ------------------------------------------------------------------------------
    def _prepare_managers(self, data_path: list, correct: bool = False) -> BatchPlan:
        sizes = [self.experiment.store.get_data_size(p) for p in data_path]
        self.plan = plan_batches(sizes, self.args.data_range, self.args.correlation_time,
                                 self.scale_function)
        return self.plan
