"""Chrome trace-event tracer.

Records begin/end events for task runs (and, on the GPU executors, CUDA-event
timed kernel phases) and renders them in the Chrome ``chrome://tracing`` /
Perfetto JSON format: one *pid* per machine/rank (pid 0 is the evaluator),
virtual *tids* drawn from a per-pid free list so concurrent tasks stack as
rows, B/E pairs coalesced into complete ("X") events, orphans dropped.  The
schema matches the reference so ``slicetrace`` analyses carry over.

Capability parity: reference ``exec/tracer.go:29-242`` and
``internal/trace/trace.go:14-41``.
"""
from __future__ import annotations

import json
import threading
import time
from typing import Any, Dict, List, Optional, Tuple


class Event:
    """One trace event (``trace.Event``)."""

    __slots__ = ("pid", "tid", "ts", "ph", "dur", "name", "cat", "args")

    def __init__(self, pid, tid, ts, ph, name, cat="", args=None, dur=0):
        self.pid, self.tid, self.ts, self.ph = pid, tid, ts, ph
        self.name, self.cat, self.args, self.dur = name, cat, args or {}, dur

    def to_json(self) -> Dict[str, Any]:
        d = {"pid": self.pid, "tid": self.tid, "ts": self.ts, "ph": self.ph,
             "name": self.name, "cat": self.cat, "args": self.args}
        if self.ph == "X":
            d["dur"] = self.dur
        return d


class Tracer:
    def __init__(self):
        self._lock = threading.Lock()
        self.events: List[Event] = []
        self._pids: Dict[Any, int] = {}
        self._free: Dict[int, List[int]] = {}
        self._next_tid: Dict[int, int] = {}
        self._assigned: Dict[Tuple[int, str], int] = {}
        self._t0 = time.time()
        self._names: Dict[int, str] = {0: "evaluator"}

    def _pid(self, machine) -> int:
        if machine is None:
            return 0
        if machine not in self._pids:
            self._pids[machine] = len(self._pids) + 1
            self._names[self._pids[machine]] = str(machine)
        return self._pids[machine]

    def _now_us(self) -> int:
        return int((time.time() - self._t0) * 1e6)

    def event(self, machine, subject, ph: str, ts_us: Optional[int] = None, **args) -> None:
        """Record a B (begin) or E (end) event for ``subject`` (a task, an
        invocation or any object with a printable name) on ``machine``."""
        name = str(getattr(subject, "name", subject))
        if hasattr(subject, "name") and hasattr(subject.name, "inv_index"):
            name = "inv%d:%s" % (subject.name.inv_index, subject.name) \
                if not str(subject.name).startswith("inv") else str(subject.name)
        ts = self._now_us() if ts_us is None else ts_us
        with self._lock:
            pid = self._pid(machine)
            key = (pid, name)
            if ph == "B":
                free = self._free.setdefault(pid, [])
                if free:
                    tid = free.pop(0)
                else:
                    tid = self._next_tid.get(pid, 0)
                    self._next_tid[pid] = tid + 1
                self._assigned[key] = tid
            elif ph == "E":
                if key not in self._assigned:
                    return
                tid = self._assigned.pop(key)
                self._free.setdefault(pid, []).append(tid)
                self._free[pid].sort()
            else:
                tid = 0
            self.events.append(Event(pid, tid, ts, ph, name, "task" if ph in "BE" else "", args))

    def complete(self, machine, name: str, ts_us: int, dur_us: int, cat: str = "kernel",
                 tid: int = 100, **args) -> None:
        """Record a complete (X) event, e.g. a CUDA-event timed kernel phase."""
        with self._lock:
            self.events.append(Event(self._pid(machine), tid, ts_us, "X", name, cat, args, dur_us))

    def marshal(self) -> str:
        """JSON with B/E pairs coalesced into X events (orphans dropped)."""
        with self._lock:
            events = list(self.events)
            names = dict(self._names)
        out: List[Dict[str, Any]] = []
        open_: Dict[Tuple[int, int, str], Event] = {}
        for e in events:
            if e.ph == "B":
                open_[(e.pid, e.tid, e.name)] = e
            elif e.ph == "E":
                b = open_.pop((e.pid, e.tid, e.name), None)
                if b is None:
                    continue
                args = dict(b.args)
                args.update(e.args)
                out.append(Event(e.pid, e.tid, b.ts, "X", e.name, b.cat, args,
                                 max(e.ts - b.ts, 0)).to_json())
            else:
                out.append(e.to_json())
        for pid, n in names.items():
            out.append({"pid": pid, "tid": 0, "ts": 0, "ph": "M", "name": "process_name",
                        "cat": "", "args": {"name": n}})
        return json.dumps({"traceEvents": out})

    def write(self, path: str) -> None:
        with open(path, "w") as fh:
            fh.write(self.marshal())

    Event_, Marshal = event, marshal


def new_tracer() -> Tracer:
    return Tracer()
