"""
Oracle restatement of the reference batch / ensemble planner (TEST INFRASTRUCTURE).

Follows, line by line:
  mdsuite/memory_management/memory_manager.py:69-133   (constructor state)
  mdsuite/memory_management/memory_manager.py:179-219  (get_batch_size)
  mdsuite/memory_management/memory_manager.py:257-340  (_compute_atomwise_minibatch)
  mdsuite/memory_management/memory_manager.py:342-372  (get_ensemble_loop)
  mdsuite/utils/scale_functions.py:30-117              (scale functions)
  mdsuite/database/data_manager.py:118-341             (batch / atom / ensemble generators)
  mdsuite/calculators/trajectory_calculator.py:243-297 (_prepare_managers)

The only deliberate change: the amount of machine memory is an explicit input
(``memory``) instead of ``psutil.virtual_memory().available`` so that a plan is
reproducible.
"""
from __future__ import annotations

import numpy as np


# --- mdsuite/utils/scale_functions.py:30-117 ---------------------------------
def linear_scale_function(memory_usage, scale_factor=1):
    return memory_usage * scale_factor


def linearithmic_scale_function(memory_usage, scale_factor=1):
    return scale_factor * memory_usage * np.log(memory_usage)


def quadratic_scale_function(memory_usage, inner_scale_factor=1, outer_scale_factor=1):
    return outer_scale_factor * (memory_usage * inner_scale_factor) ** 2


def polynomial_scale_function(
    memory_usage, inner_scale_factor=1, outer_scale_factor=1, order=3
):
    return outer_scale_factor * (memory_usage * inner_scale_factor) ** order


_SWITCHER = {
    "linear": linear_scale_function,
    "log-linear": linearithmic_scale_function,
    "quadratic": quadratic_scale_function,
    "polynomial": polynomial_scale_function,
}


class MemoryManager:
    """memory_manager.py:47-372 with ``memory`` (bytes) made explicit."""

    def __init__(
        self,
        data_path=None,
        database=None,
        memory_fraction=0.5,
        scale_function=None,
        memory=60e9,
        offset=0,
    ):
        if scale_function is None:
            scale_function = {"linear": {"scale_factor": 10}}
        self.data_path = data_path
        self.database = database
        # memory_manager.py:108 -- the ctor argument is ignored upstream and the
        # global config.memory_fraction (default 0.5) is used; callers pass that.
        self.memory_fraction = memory_fraction
        self.offset = offset
        self.machine_properties = {"memory": memory}
        self.batch_size = None
        self.n_batches = None
        self.remainder = None
        self.atom_batch_size = None
        self.n_atom_batches = None
        self.atom_remainder = None
        self.minibatch = False
        self.scale_function, self.scale_function_parameters = self._select_scale_function(
            scale_function
        )

    @staticmethod
    def _select_scale_function(input_dict):
        # memory_manager.py:135-177
        key = list(input_dict.keys())[0]
        try:
            fn = _SWITCHER[key]
        except KeyError:
            raise KeyError("Invalid choice")
        return fn, input_dict[key]

    def get_batch_size(self):
        # memory_manager.py:179-219
        if self.data_path is None:
            raise ValueError("No tensor_values have been requested.")
        per_configuration_memory = 0.0
        n_configs = None
        for item in self.data_path:
            n_particles, n_configs, n_bytes = self.database.get_data_size(item)
            per_configuration_memory += n_bytes / n_configs
        per_configuration_memory = self.scale_function(
            per_configuration_memory, **self.scale_function_parameters
        )
        maximum_loaded_configurations = int(
            np.clip(
                (self.memory_fraction * self.machine_properties["memory"])
                / per_configuration_memory,
                1,
                n_configs - self.offset,
            )
        )
        batch_size = maximum_loaded_configurations  # _get_optimal_batch_size :239-255
        number_of_batches, remainder = divmod((n_configs - self.offset), batch_size)
        self.batch_size = batch_size
        self.n_batches = number_of_batches
        self.remainder = remainder
        return batch_size, number_of_batches, remainder

    def _compute_atomwise_minibatch(self, data_range):
        # memory_manager.py:257-340
        per_atom_memory = 0
        per_configuration_memory = 0
        total_rows = 0
        n_particles = n_configs = None
        for item in self.data_path:
            n_particles, n_configs, n_bytes = self.database.get_data_size(item)
            per_configuration_memory += n_bytes / n_configs
            per_atom_memory += per_configuration_memory / n_particles
            total_rows += n_particles
        per_atom_memory = self.scale_function(
            per_atom_memory, **self.scale_function_parameters
        )
        fractions = [1 / 2, 1 / 4, 1 / 8, 1 / 20, 1 / 100, 1 / 200, 0]
        batch_size = None
        for fraction in fractions:
            if fraction == 0:
                batch_size = int(
                    np.clip(
                        self.memory_fraction
                        * self.machine_properties["memory"]
                        / per_atom_memory,
                        1,
                        n_configs,
                    )
                )
                self.atom_batch_size = 1
                break
            atom_batch_memory = fraction * per_atom_memory
            batch_size = int(
                np.clip(
                    self.memory_fraction
                    * self.machine_properties["memory"]
                    / atom_batch_memory,
                    1,
                    n_configs,
                )
            )
            if batch_size > data_range:
                self.atom_batch_size = n_particles * fraction
                break
        self.batch_size = batch_size
        self.n_batches = int(n_configs / batch_size)
        self.remainder = int(n_configs % batch_size)
        self.n_atom_batches = int(n_particles / self.atom_batch_size)
        self.atom_remainder = int(n_particles % self.atom_batch_size)

    def get_ensemble_loop(self, data_range, correlation_time=1):
        # memory_manager.py:342-372
        final_window = self.batch_size - data_range
        if final_window < 0:
            self._compute_atomwise_minibatch(data_range)
            final_window = self.batch_size - data_range
            self.minibatch = True
            return int(np.clip(final_window / correlation_time, 1, None)), True
        return int(np.clip(final_window / correlation_time, 1, None)), False


class ArrayDatabase:
    """Minimal stand-in for simulation_database.Database over in-memory arrays.

    ``get_data_size`` follows simulation_database.py:683-690 and reports the byte
    size of the *float32* dataset (h5py default dtype).
    """

    def __init__(self, arrays: dict):
        self.arrays = arrays

    def get_data_size(self, path):
        a = self.arrays[path]
        return a.shape[0], a.shape[1], int(a.shape[0] * a.shape[1] * a.shape[2] * 4)

    def check_existence(self, path):
        return path in self.arrays


def plan_trajectory_calculator(
    database, data_path, data_range, correlation_time, scale_function, memory,
    memory_fraction=0.5,
):
    """trajectory_calculator.py:243-297 (_prepare_managers) -> dict plan."""
    mm = MemoryManager(
        data_path=data_path,
        database=database,
        memory_fraction=memory_fraction,
        scale_function=scale_function,
        memory=memory,
    )
    batch_size, n_batches, remainder = mm.get_batch_size()
    ensemble_loop, minibatch = mm.get_ensemble_loop(data_range, correlation_time)
    if minibatch:
        batch_size, n_batches, remainder = mm.batch_size, mm.n_batches, mm.remainder
    return dict(
        batch_size=batch_size,
        n_batches=n_batches,
        remainder=remainder,
        ensemble_loop=ensemble_loop,
        minibatch=minibatch,
        atom_batch_size=mm.atom_batch_size,
        n_atom_batches=mm.n_atom_batches,
        atom_remainder=mm.atom_remainder,
    )


def iter_batches(plan, n_configs_offset=0, system=False):
    """data_manager.py:156-281: yield (atom_slice, start, stop, data_size).

    ``generator`` (no atom minibatching) loops the remainder batch when
    ``remainder > 0`` (:184-194); ``atom_generator`` is called without the
    ``remainder`` flag by trajectory_calculator.py:354 so it never does.
    """
    bs, nb, rem = plan["batch_size"], plan["n_batches"], plan["remainder"]
    if not plan["minibatch"]:
        loop_rem = rem > 0
        for batch in range(nb + int(loop_rem)):
            start = int(batch * bs) + n_configs_offset
            stop = int(start + bs)
            data_size = bs
            if batch == nb:
                stop = int(start + rem)
                data_size = rem
            yield slice(None), start, stop, data_size
    else:
        abs_, nab, arem = (
            plan["atom_batch_size"], plan["n_atom_batches"], plan["atom_remainder"],
        )
        _atom_remainder = 1 if arem else 0
        start = 0
        for atom_batch in range(nab + _atom_remainder):
            atom_start = atom_batch * abs_
            atom_stop = atom_start + abs_
            if atom_batch == nab:
                atom_stop = start + arem  # data_manager.py:266-267 (upstream bug kept)
            for batch in range(nb):
                start = int(batch * bs) + n_configs_offset
                stop = int(start + bs)
                yield slice(int(atom_start), int(atom_stop)), start, stop, bs


def iter_ensembles(data_size, data_range, correlation_time):
    """data_manager.py:309-339: window (start, stop) pairs inside one batch."""
    ensemble_loop = int(np.clip((data_size - data_range) / correlation_time, 1, None))
    for ensemble in range(ensemble_loop):
        start = ensemble * correlation_time
        yield start, start + data_range
