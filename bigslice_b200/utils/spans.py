"""Host-side span recorder for finding scheduling / control-plane overhead.

Enabled with ``BIGSLICE_SPANS=1`` (or :func:`enable`); otherwise ``span`` is a
no-op context manager costing one attribute test.  Spans are (name, thread,
start, end) in ``time.perf_counter`` seconds; :func:`summary` aggregates them
per name over a window.  This is the host-time complement of the CUDA-event
timings in the exchange log and of the Chrome trace (``exec/tracer.py``,
reference ``exec/tracer.go``): it answers "where does a Session.run spend
wall-clock outside kernels".
"""
from __future__ import annotations

import os
import threading
import time
from typing import Dict, List, Tuple

_enabled = os.environ.get("BIGSLICE_SPANS", "") not in ("", "0")
_lock = threading.Lock()
_spans: List[Tuple[str, str, float, float]] = []


def enable(on: bool = True) -> None:
    global _enabled
    _enabled = on


def enabled() -> bool:
    return _enabled


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _Null()


class _Span:
    __slots__ = ("name", "t0")

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        self.t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        t1 = time.perf_counter()
        with _lock:
            _spans.append((self.name, threading.current_thread().name, self.t0, t1))
        return False


def span(name: str):
    """``with span("phase"):`` records the enclosed wall-clock when enabled."""
    return _Span(name) if _enabled else _NULL


def mark() -> int:
    """Position in the span log (use with ``since=`` to window a summary)."""
    with _lock:
        return len(_spans)


def clear() -> None:
    with _lock:
        del _spans[:]


def raw(since: int = 0) -> List[Tuple[str, str, float, float]]:
    with _lock:
        return list(_spans[since:])


def summary(since: int = 0, per: int = 1) -> List[Dict]:
    """Per-name totals (ms), divided by ``per`` (e.g. the number of steps)."""
    agg: Dict[str, List[float]] = {}
    for name, _th, t0, t1 in raw(since):
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += (t1 - t0) * 1e3
    out = [{"span": k, "count_per": c / per, "ms_per": ms / per} for k, (c, ms) in agg.items()]
    out.sort(key=lambda d: -d["ms_per"])
    return out


def format_summary(since: int = 0, per: int = 1) -> str:
    rows = summary(since, per)
    lines = ["%-34s %8s %10s" % ("span", "n/step", "ms/step")]
    for r in rows:
        lines.append("%-34s %8.1f %10.3f" % (r["span"], r["count_per"], r["ms_per"]))
    return "\n".join(lines)
