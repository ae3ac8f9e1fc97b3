"""Test kit: run slices on a local session and inspect the rows.

Capability parity: reference ``slicetest/run.go:24-93`` (Run, RunErr, ScanAll,
RunAndScan) and ``slicetest/print.go:20-122`` (deterministic, deep-sorted
Print).
"""
from __future__ import annotations

import io
import sys
from typing import Any, Callable, List, Optional, Sequence, Tuple

from .. import exec as bsexec
from ..func import Func, FuncValue
from ..slices import Slice


def _as_func(slice_or_func) -> Tuple[FuncValue, tuple]:
    if isinstance(slice_or_func, FuncValue):
        return slice_or_func, ()
    s = slice_or_func
    return Func(lambda: s), ()


def run_err(slice_or_func, *args, options: Sequence = (), session=None):
    """Run on a fresh local session; returns ``(result, error)``."""
    fn, _ = _as_func(slice_or_func)
    sess = session or bsexec.start(bsexec.Local, *options)
    try:
        return sess.run(fn, *args), None
    except BaseException as e:  # noqa: BLE001 - the kit reports any failure
        return None, e


def run(slice_or_func, *args, options: Sequence = (), session=None):
    """Run and return the Result, raising on failure (``slicetest.Run``)."""
    res, err = run_err(slice_or_func, *args, options=options, session=session)
    if err is not None:
        raise err
    return res


def scan_all(result) -> List[list]:
    """All rows as a list of columns (python lists) (``ScanAll``)."""
    rows = result.rows()
    n = result.num_out()
    cols: List[list] = [[] for _ in range(n)]
    for r in rows:
        for c, v in zip(cols, r):
            c.append(v)
    return cols


def run_and_scan(slice_or_func, *args, options: Sequence = (), session=None) -> List[list]:
    return scan_all(run(slice_or_func, *args, options=options, session=session))


def sorted_rows(result) -> List[tuple]:
    """Rows deep-sorted for order-independent comparison."""
    def norm(v):
        if isinstance(v, list):
            return tuple(sorted((norm(x) for x in v), key=_key))
        return v
    return sorted((tuple(norm(v) for v in r) for r in result.rows()), key=_key)


def _key(v):
    if isinstance(v, tuple):
        return tuple(_key(x) for x in v)
    return (type(v).__name__, v) if not isinstance(v, (int, float, bool)) else ("num", v)


def print_(result, w=None) -> None:
    """Deterministically print rows (sorted), tab-separated (``Print``)."""
    w = w or sys.stdout
    for r in sorted_rows(result):
        w.write(" ".join(_fmt(v) for v in r) + "\n")


def _fmt(v) -> str:
    if isinstance(v, tuple):
        return "[" + " ".join(_fmt(x) for x in v) + "]"
    return str(v)


Run, RunErr, ScanAll, RunAndScan, Print = run, run_err, scan_all, run_and_scan, print_
