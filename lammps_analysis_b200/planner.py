"""Batch / window planning: how many frames form a batch, how many correlation windows a batch
holds, and when atoms are mini-batched.

The plan is *semantics*, not performance, here: correlation windows never cross a batch
boundary and the normalisation counts windows per atom batch, so the GPU path must follow the
same plan as the reference to produce the same numbers.  What is
physically resident in HBM is decided elsewhere (store.py / engine.py).

Restates mdsuite/memory_management/memory_manager.py:179-372,
mdsuite/utils/scale_functions.py:30-117, mdsuite/database/data_manager.py:118-341 and
mdsuite/calculators/trajectory_calculator.py:243-297.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

from .config import config, machine_memory

_ATOM_FRACTIONS = (1 / 2, 1 / 4, 1 / 8, 1 / 20, 1 / 100, 1 / 200, 0)


def scale(memory_usage: float, spec: Optional[dict]) -> float:
    """Evaluate a scale-function spec such as {"linear": {"scale_factor": 150}}
    (scale_functions.py:30-117; default memory_manager.py:100-101)."""
    if spec is None:
        spec = {"linear": {"scale_factor": 10}}
    kind, par = next(iter(spec.items()))
    if kind == "linear":
        return memory_usage * par.get("scale_factor", 1)
    if kind == "log-linear":
        return par.get("scale_factor", 1) * memory_usage * math.log(memory_usage)
    if kind == "quadratic":
        return par.get("outer_scale_factor", 1) * (memory_usage * par.get("inner_scale_factor", 1)) ** 2
    if kind == "polynomial":
        return par.get("outer_scale_factor", 1) * (
            memory_usage * par.get("inner_scale_factor", 1)) ** par.get("order", 3)
    raise KeyError("Invalid choice")


def _clip_int(x: float, lo: float, hi: Optional[float]) -> int:
    if hi is not None:
        x = min(x, hi)
    return int(max(x, lo))


@dataclass
class BatchPlan:
    batch_size: int
    n_batches: int
    remainder: int
    ensemble_loop: int
    minibatch: bool = False
    atom_batch_size: Optional[float] = None
    n_atom_batches: Optional[int] = None
    atom_remainder: Optional[int] = None
    memory: float = 0.0
    memory_fraction: float = 0.5

    def as_dict(self) -> dict:
        return dict(self.__dict__)


def plan_batches(data_sizes: List[Tuple[int, int, int]], data_range: int, correlation_time: int,
                 scale_function: Optional[dict], offset: int = 0, memory: Optional[float] = None,
                 memory_fraction: Optional[float] = None) -> BatchPlan:
    """data_sizes: one (n_rows, n_configurations, n_bytes) triple per loaded dataset
    (simulation_database.get_data_size, :683-690; n_bytes of the float32 dataset)."""
    mem = machine_memory() if memory is None else float(memory)
    frac = config.memory_fraction if memory_fraction is None else float(memory_fraction)
    if not data_sizes:
        raise ValueError("No tensor_values have been requested.")
    budget = frac * mem
    per_cfg = sum(nb / nc for (_, nc, nb) in data_sizes)
    n_configs = data_sizes[-1][1]
    batch = _clip_int(budget / scale(per_cfg, scale_function), 1, n_configs - offset)
    n_batches, remainder = divmod(n_configs - offset, batch)
    plan = BatchPlan(batch, n_batches, remainder, 1, memory=mem, memory_fraction=frac)
    if batch - data_range < 0:
        # atom-wise mini-batching (memory_manager.py:257-340)
        per_cfg_acc, per_atom = 0.0, 0.0
        n_rows = 0
        for rows, nc, nb in data_sizes:
            per_cfg_acc += nb / nc
            per_atom += per_cfg_acc / rows
            n_rows, n_configs = rows, nc
        per_atom = scale(per_atom, scale_function)
        for fraction in _ATOM_FRACTIONS:
            if fraction == 0:
                batch = _clip_int(budget / per_atom, 1, n_configs)
                atom_batch = 1
                break
            batch = _clip_int(budget / (fraction * per_atom), 1, n_configs)
            if batch > data_range:
                atom_batch = n_rows * fraction
                break
        plan.batch_size = batch
        plan.n_batches = int(n_configs / batch)
        plan.remainder = int(n_configs % batch)
        plan.atom_batch_size = atom_batch
        plan.n_atom_batches = int(n_rows / atom_batch)
        plan.atom_remainder = int(n_rows % atom_batch)
        plan.minibatch = True
    plan.ensemble_loop = window_count(plan.batch_size, data_range, correlation_time)
    return plan


def window_count(batch_size: int, data_range: int, correlation_time: int) -> int:
    """int(clip((B - N) / ct, 1, None)) -- the last admissible window is skipped (Q5)."""
    return _clip_int((batch_size - data_range) / correlation_time, 1, None)


def frame_batches(plan: BatchPlan, offset: int = 0) -> List[Tuple[int, int]]:
    """[start, stop) frame ranges of the plain batch generator (data_manager.py:156-221)."""
    out = [(b * plan.batch_size + offset, (b + 1) * plan.batch_size + offset)
           for b in range(plan.n_batches)]
    if plan.remainder > 0 and not plan.minibatch:
        s = plan.n_batches * plan.batch_size + offset
        out.append((s, s + plan.remainder))
    return out
