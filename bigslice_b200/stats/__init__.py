"""Named integer counters used for internal runtime statistics.

Capability parity: reference ``stats/stats.go:19-107`` (Map, Int, Values,
``Values.String``/``Copy``): per-task and per-rank counters such as
``read``/``write``/``inrecords``/``readDuration``/``writeDuration``.
"""
from __future__ import annotations

import threading
from typing import Dict


class Int:
    """An atomically updated integer."""

    __slots__ = ("_v", "_lock")

    def __init__(self):
        self._v = 0
        self._lock = threading.Lock()

    def add(self, delta: int) -> None:
        with self._lock:
            self._v += int(delta)

    def set(self, v: int) -> None:
        with self._lock:
            self._v = int(v)

    def get(self) -> int:
        with self._lock:
            return self._v

    Add, Set, Get = add, set, get


class Values(dict):
    """A snapshot of counters."""

    def copy(self) -> "Values":
        return Values(self)

    def __str__(self):
        return " ".join("%s:%d" % (k, self[k]) for k in sorted(self))

    Copy = copy


class Map:
    """A collection of named counters."""

    def __init__(self):
        self._lock = threading.Lock()
        self._ints: Dict[str, Int] = {}

    def int(self, name: str) -> Int:
        with self._lock:
            v = self._ints.get(name)
            if v is None:
                v = self._ints[name] = Int()
            return v

    def add_all(self, vals: Values) -> None:
        """Write the current counter values into ``vals``."""
        with self._lock:
            items = list(self._ints.items())
        for k, v in items:
            vals[k] = v.get()

    def values(self) -> Values:
        v = Values()
        self.add_all(v)
        return v

    Int_, AddAll = int, add_all


def NewMap() -> Map:
    return Map()
