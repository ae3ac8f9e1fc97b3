"""Hierarchical status display: a Status holds Groups which hold Tasks.

Used by sessions to show per-slice and per-task progress (rendered to a
terminal, ``/debug/status`` or logs).  Plays the role ``grailbio/base/status``
plays in the reference (``exec/slicestatus.go``, ``docs/index.md:223-278``).
"""
from __future__ import annotations

import io
import threading
import time
from typing import List, Optional


class StatusTask:
    def __init__(self, group: "Group", title: str):
        self.group = group
        self._title = title
        self._status = ""
        self.begin = time.time()
        self.end: Optional[float] = None

    def title(self, t: str) -> None:
        self._title = t

    def print(self, s: str) -> None:
        self._status = s
        self.group._bump()

    def printf(self, fmt: str, *args) -> None:
        self.print(fmt % args if args else fmt)

    def done(self) -> None:
        self.end = time.time()
        self.group._remove(self)

    Print, Printf, Done, Title = print, printf, done, title


class Group:
    def __init__(self, status: "Status", name: str):
        self.status = status
        self.name = name
        self._status = ""
        self._tasks: List[StatusTask] = []
        self._lock = threading.Lock()

    def print(self, s: str) -> None:
        self._status = s
        self._bump()

    def printf(self, fmt: str, *args) -> None:
        self.print(fmt % args if args else fmt)

    def start(self, title: str = "") -> StatusTask:
        t = StatusTask(self, title)
        with self._lock:
            self._tasks.append(t)
        return t

    def startf(self, fmt: str, *args) -> StatusTask:
        return self.start(fmt % args if args else fmt)

    def tasks(self) -> List[StatusTask]:
        with self._lock:
            return list(self._tasks)

    def _remove(self, t: StatusTask) -> None:
        with self._lock:
            if t in self._tasks:
                self._tasks.remove(t)
        self._bump()

    def _bump(self) -> None:
        self.status._bump()

    def done(self) -> None:
        self.status._remove(self)

    Print, Printf, Start, Startf, Done = print, printf, start, startf, done


class Status:
    def __init__(self):
        self._groups: List[Group] = []
        self._lock = threading.Lock()
        self.version = 0

    def group(self, name: str) -> Group:
        g = Group(self, name)
        with self._lock:
            self._groups.append(g)
        return g

    def groupf(self, fmt: str, *args) -> Group:
        return self.group(fmt % args if args else fmt)

    def groups(self) -> List[Group]:
        with self._lock:
            return list(self._groups)

    def _remove(self, g: Group) -> None:
        with self._lock:
            if g in self._groups:
                self._groups.remove(g)
        self._bump()

    def _bump(self) -> None:
        self.version += 1

    def marshal(self, w=None) -> str:
        """Render all groups and tasks as text."""
        b = io.StringIO()
        now = time.time()
        for g in self.groups():
            b.write("%s: %s\n" % (g.name, g._status))
            for t in g.tasks():
                el = (t.end or now) - t.begin
                b.write("  %-40s %-30s %s\n" % (t._title + ":", t._status, _dur(el)))
        s = b.getvalue()
        if w is not None:
            w.write(s)
        return s

    Group_, Groupf, Marshal = group, groupf, marshal


def _dur(s: float) -> str:
    if s < 60:
        return "%.0fs" % s
    m, s = divmod(int(s), 60)
    if m < 60:
        return "%dm%ds" % (m, s)
    h, m = divmod(m, 60)
    return "%dh%dm%ds" % (h, m, s)
