"""Opt-in host timeline (MDK_TRACE=1): named marks with the host clock, printed on demand.
A debug aid for the end-to-end paths, where the question is which call the host blocks in;
no cost when the switch is off."""
from __future__ import annotations

import os
import sys
import time

ENABLED = os.environ.get("MDK_TRACE", "") not in ("", "0")
_marks = []


def mark(name: str):
    if ENABLED:
        _marks.append((time.perf_counter(), name))


_events = []


def event(name: str):
    """Timestamp on the CURRENT CUDA stream (a timing event): when the device gets there."""
    if ENABLED:
        import torch

        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        _events.append((ev, name))


def dump(stream=None, reset: bool = True):
    """Print the marks since the last dump as ``+ms since first  (+ms since previous)  name``."""
    if not ENABLED or not _marks:
        return
    stream = stream or sys.stderr
    t0, prev = _marks[0][0], _marks[0][0]
    for t, name in _marks:
        stream.write(f"[mdk-trace] {1e3 * (t - t0):9.2f} ms  (+{1e3 * (t - prev):8.2f})  {name}\n")
        prev = t
    if _events:
        import torch

        torch.cuda.synchronize()
        e0 = _events[0][0]
        rows = sorted((e0.elapsed_time(ev), name) for ev, name in _events)
        for t, name in rows:
            stream.write(f"[mdk-trace-gpu] {t:9.2f} ms  {name}\n")
    if reset:
        _marks.clear()
        _events.clear()
