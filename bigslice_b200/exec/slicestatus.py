"""Slice-level status: per-slice task state counts, and periodic invocation logs.

For every slice of an invocation the status line shows
``tasks idle/running/done`` aggregated over the tasks that compute it
(pipelined slices share tasks); a background logger prints a one-line summary
per invocation every minute while it runs.

Capability parity: reference ``exec/slicestatus.go:19-202`` (sliceStatus,
maintainSliceGroup, logInvocation).
"""
from __future__ import annotations

import logging
import threading
from typing import Dict, List, Optional, Sequence

from ..status import Group, Status
from .task import Task, TaskState, TaskSubscriber, iter_tasks

log = logging.getLogger("bigslice")


class SliceStatus:
    """Counts of task states for one slice."""

    def __init__(self, name: str):
        self.name = name
        self.counts = {s: 0 for s in TaskState}

    def idle(self) -> int:
        return self.counts[TaskState.INIT] + self.counts[TaskState.WAITING] + self.counts[TaskState.LOST]

    def running(self) -> int:
        return self.counts[TaskState.RUNNING]

    def done(self) -> int:
        return self.counts[TaskState.OK]

    def errors(self) -> int:
        return self.counts[TaskState.ERR]

    def __str__(self):
        s = "tasks idle/running/done: %d/%d/%d" % (self.idle(), self.running(), self.done())
        if self.errors():
            s += " errors: %d" % self.errors()
        return s


class SliceGroupMaintainer(TaskSubscriber):
    """Keeps a status group with one line per slice up to date."""

    def __init__(self, group: Optional[Group], tasks: Sequence[Task]):
        self.group = group
        self.lock = threading.Lock()
        self.task_state: Dict[int, TaskState] = {}
        self.slices: Dict[str, SliceStatus] = {}
        self.by_task: Dict[int, List[str]] = {}
        self.lines = {}
        all_tasks: List[Task] = []
        iter_tasks(tasks, all_tasks.append)
        self.tasks = all_tasks
        for t in all_tasks:
            names = [str(s.name()) for s in t.slices] or [str(t.name.op)]
            self.by_task[id(t)] = names
            st = t.state()
            self.task_state[id(t)] = st
            for n in names:
                self.slices.setdefault(n, SliceStatus(n)).counts[st] += 1
        if group is not None:
            for n in sorted(self.slices):
                self.lines[n] = group.start(n)
        self._render()
        for t in all_tasks:
            t.subscribe(self)

    def notify(self, task: Task) -> None:
        with self.lock:
            old = self.task_state.get(id(task))
            new = task.state()
            if old == new:
                return
            self.task_state[id(task)] = new
            for n in self.by_task.get(id(task), []):
                ss = self.slices[n]
                if old is not None:
                    ss.counts[old] -= 1
                ss.counts[new] += 1
        self._render()

    def _render(self):
        for n, line in self.lines.items():
            line.print(str(self.slices[n]))

    def summary(self) -> Dict[str, str]:
        with self.lock:
            return {n: str(s) for n, s in self.slices.items()}

    def close(self):
        for t in self.tasks:
            t.unsubscribe(self)
        for line in self.lines.values():
            line.done()


def log_invocation(inv_index: int, location: str, maint: SliceGroupMaintainer,
                   stop: threading.Event, period: float = 60.0) -> None:
    """Log per-slice progress every ``period`` seconds until ``stop`` is set."""
    while not stop.wait(period):
        for n, s in sorted(maint.summary().items()):
            log.info("invocation %d (%s): %s: %s", inv_index, location, n, s)
