"""Global configuration (mirrors mdsuite/utils/config.py:31-59 and helpers.py:34-83).

``memory_fraction`` and the amount of machine memory decide the batch plan, and the batch plan
decides which correlation windows exist.  The reference reads free host RAM
at call time; here ``planner_memory_bytes`` can pin it so that plans are reproducible.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass
from typing import Optional


@dataclass
class Config:
    memory_fraction: float = 0.5
    # None -> psutil.virtual_memory().available, as the reference does
    planner_memory_bytes: Optional[float] = None
    # device cache budget for store arrays (bytes); None -> 60 % of the free HBM at first use
    device_cache_bytes: Optional[int] = None
    # page-locked host blocks kept for reuse after a dataset is removed or replaced (store.py)
    pinned_pool_bytes: int = 96 << 30
    # row-block size of block-wise host -> device copies (store.device_blocks): consumers start
    # on the first block while the rest is on the wire
    upload_block_bytes: int = 512 << 20
    jupyter: bool = False


config = Config()


def machine_memory() -> float:
    """meta_functions.get_machine_properties()['memory'] (:132-158)."""
    if config.planner_memory_bytes is not None:
        return float(config.planner_memory_bytes)
    import psutil

    return float(psutil.virtual_memory().available)


@contextlib.contextmanager
def change_memory_fraction(desired_memory: float = None):
    """helpers.change_memory_fraction (:59-83): make the planner believe only
    ``desired_memory`` GB are usable.  With a pinned ``planner_memory_bytes`` the pinned value
    is what the fraction refers to."""
    if desired_memory is None:
        yield
        return
    old = config.memory_fraction
    config.memory_fraction = desired_memory * 1e9 / machine_memory()
    try:
        yield
    finally:
        config.memory_fraction = old
