"""slicetrace: offline analysis of a session trace (Chrome trace JSON).

    python -m bigslice_b200.cmd.slicetrace trace.json

Prints, per invocation, a table of ops with start offset, duration, total task
time, read/write time and Tukey quartiles (min/q1/q2/q3/max) of task durations.

Capability parity: reference ``cmd/slicetrace/main.go:18-171``,
``session.go:21-266`` (task-name regex ``inv(\\d+)_([^@]*)@(\\d+):(\\d+)``),
``quartile.go:11-41``.
"""
from __future__ import annotations

import json
import re
import sys
from collections import defaultdict
from typing import Dict, List, Sequence, Tuple

TASK_RE = re.compile(r"inv(\d+)_([^@]*)@(\d+):(\d+)")


def quartiles(xs: Sequence[float]) -> Tuple[float, float, float, float, float]:
    """Tukey's five-number summary (hinges use the median-inclusive method)."""
    s = sorted(xs)
    n = len(s)
    if n == 0:
        return (0.0, 0.0, 0.0, 0.0, 0.0)

    def median(a):
        m = len(a)
        return a[m // 2] if m % 2 else (a[m // 2 - 1] + a[m // 2]) / 2.0
    q2 = median(s)
    half = (n + 1) // 2
    q1 = median(s[:half])
    q3 = median(s[n - half:])
    return (s[0], q1, q2, q3, s[-1])


class TaskEvent:
    __slots__ = ("inv", "op", "nshard", "shard", "ts", "dur", "read", "write", "err", "pid")

    def __init__(self, inv, op, nshard, shard, ts, dur, read, write, err, pid):
        self.inv, self.op, self.nshard, self.shard = inv, op, nshard, shard
        self.ts, self.dur, self.read, self.write, self.err, self.pid = ts, dur, read, write, err, pid


def parse_trace(data: dict) -> List[TaskEvent]:
    out = []
    for e in data.get("traceEvents", []):
        if e.get("ph") != "X":
            continue
        m = TASK_RE.search(e.get("name", ""))
        if not m:
            continue
        args = e.get("args", {}) or {}
        out.append(TaskEvent(int(m.group(1)), m.group(2), int(m.group(3)), int(m.group(4)),
                             float(e.get("ts", 0)), float(e.get("dur", 0)),
                             float(args.get("readDuration", 0)), float(args.get("writeDuration", 0)),
                             args.get("error_type"), e.get("pid", 0)))
    return out


def summarize(events: List[TaskEvent]) -> Dict[int, List[dict]]:
    by_inv: Dict[int, Dict[str, List[TaskEvent]]] = defaultdict(lambda: defaultdict(list))
    for ev in events:
        by_inv[ev.inv][ev.op].append(ev)
    out: Dict[int, List[dict]] = {}
    for inv, ops in sorted(by_inv.items()):
        t0 = min(e.ts for evs in ops.values() for e in evs)
        rows = []
        for op, evs in ops.items():
            start = min(e.ts for e in evs)
            end = max(e.ts + e.dur for e in evs)
            rows.append({"op": op, "tasks": len(evs), "start": (start - t0) / 1e6,
                         "duration": (end - start) / 1e6, "total": sum(e.dur for e in evs) / 1e6,
                         "read": sum(e.read for e in evs) / 1e6,
                         "write": sum(e.write for e in evs) / 1e6,
                         "lost": sum(1 for e in evs if e.err == "lost"),
                         "quartiles": tuple(q / 1e6 for q in quartiles([e.dur for e in evs]))})
        rows.sort(key=lambda r: r["start"])
        out[inv] = rows
    return out


def render(summary: Dict[int, List[dict]], w=None) -> None:
    w = w or sys.stdout
    for inv, rows in summary.items():
        w.write("invocation %d\n" % inv)
        w.write("  %-40s %6s %9s %9s %9s %9s %9s  %s\n" % (
            "op", "tasks", "start", "duration", "total", "read", "write", "min/q1/q2/q3/max"))
        for r in rows:
            w.write("  %-40s %6d %9.3f %9.3f %9.3f %9.3f %9.3f  %s\n" % (
                r["op"][:40], r["tasks"], r["start"], r["duration"], r["total"], r["read"],
                r["write"], "/".join("%.3f" % q for q in r["quartiles"])))


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) != 1:
        sys.stderr.write("usage: slicetrace trace.json\n")
        return 2
    with open(argv[0]) as fh:
        data = json.load(fh)
    render(summarize(parse_trace(data)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
