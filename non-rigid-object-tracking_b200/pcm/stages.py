"""Wall-clock accounting of the HOST stages around the GPU hot path (sweep / sequence driver).

    with stages.stage("fit"): ...        # adds the elapsed seconds of this thread to "fit"

Seconds are summed over all threads of the process (so they can exceed the wall time of a
multi-threaded run); `snapshot()` returns {stage: seconds}.  Used for the per-stage table of
the grid sweep (bench.py `sweep.per_stage_host_s`)."""
import threading
import time
from contextlib import contextmanager

_lock = threading.Lock()
_totals = {}
_counts = {}


@contextmanager
def stage(name):
    t0 = time.perf_counter()
    try:
        yield
    finally:
        dt = time.perf_counter() - t0
        with _lock:
            _totals[name] = _totals.get(name, 0.0) + dt
            _counts[name] = _counts.get(name, 0) + 1


def add(name, seconds, n=1):
    with _lock:
        _totals[name] = _totals.get(name, 0.0) + seconds
        _counts[name] = _counts.get(name, 0) + n


def reset():
    with _lock:
        _totals.clear()
        _counts.clear()


def snapshot():
    with _lock:
        return dict(_totals)


def counts():
    with _lock:
        return dict(_counts)
