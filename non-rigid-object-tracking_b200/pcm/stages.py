"""Wall-clock accounting of the HOST stages around the GPU hot path (sweep / sequence driver).

    with stages.stage("fit"): ...        # adds the elapsed seconds of this thread to "fit"

Seconds are summed over all threads of the process (so they can exceed the wall time of a
multi-threaded run); `snapshot()` returns {stage: seconds}.  Used for the per-stage table of
the grid sweep (bench.py `sweep.per_stage_host_s`).

PCM_STAGE_TIMELINE=<path>: every stage interval is also kept as (stage, thread, start, end) and `dump_timeline()`
(called by the sweep when it ends) writes them as JSON -- who waited for what, which the sums cannot tell."""
import json
import os
import threading
import time
from contextlib import contextmanager

_lock = threading.Lock()
_totals = {}
_counts = {}
_timeline_path = os.environ.get("PCM_STAGE_TIMELINE")
_events = []
_t_origin = time.perf_counter()


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
            if _timeline_path:
                _events.append((name, threading.get_ident(), t0 - _t_origin, t0 + dt - _t_origin))


def add(name, seconds, n=1):
    with _lock:
        _totals[name] = _totals.get(name, 0.0) + seconds
        _counts[name] = _counts.get(name, 0) + n


def reset():
    global _t_origin
    with _lock:
        _totals.clear()
        _counts.clear()
        del _events[:]
        _t_origin = time.perf_counter()


def dump_timeline(suffix=""):
    """Write the recorded intervals (seconds since the last reset()) to PCM_STAGE_TIMELINE + suffix; no-op when unset."""
    if not _timeline_path:
        return None
    with _lock:
        events = list(_events)
    threads = {t: k for k, t in enumerate(sorted({e[1] for e in events}))}
    path = _timeline_path + suffix
    with open(path, "w") as f:
        json.dump([dict(stage=n, thread=threads[t], start=round(a, 6), end=round(b, 6)) for n, t, a, b in events], f)
    return path


def snapshot():
    with _lock:
        return dict(_totals)


def counts():
    with _lock:
        return dict(_counts)
