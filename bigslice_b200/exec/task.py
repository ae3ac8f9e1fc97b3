"""Tasks: nodes of the executable DAG and their runtime state machine.

``INIT -> WAITING -> RUNNING -> {OK, ERR, LOST}``.  ``LOST`` means the task's
output (or the rank/GPU holding it) went away and the evaluator will recompute
it; ``ERR`` is fatal for the evaluation.

Capability parity: reference ``exec/task.go:38-510`` (TaskState, TaskName,
TaskDep, Task, TaskSubscriber, WriteGraph/All) — re-designed around a
condition variable rather than broadcast channels.
"""
from __future__ import annotations

import enum
import io
import threading
from typing import Any, Callable, Dict, List, Optional, Sequence

from .. import errors, metrics
from ..slices import Pragma, Pragmas, Slice


class TaskState(enum.IntEnum):
    INIT = 0      # initial state
    WAITING = 1   # scheduled, waiting for resources
    RUNNING = 2   # running on an executor
    OK = 3        # completed, output available
    ERR = 4       # failed with a fatal error
    LOST = 5      # output lost; will be recomputed

    def __str__(self):
        return ["INIT", "WAITING", "RUNNING", "OK", "ERROR", "LOST"][int(self)]


TaskInit, TaskWaiting, TaskRunning, TaskOk, TaskErr, TaskLost = (
    TaskState.INIT, TaskState.WAITING, TaskState.RUNNING, TaskState.OK,
    TaskState.ERR, TaskState.LOST)


class TaskName:
    """Unique task name: ``inv<idx>_<ops>@<nshard>:<shard>``."""

    __slots__ = ("inv_index", "op", "shard", "num_shard")

    def __init__(self, inv_index: int, op: str, shard: int, num_shard: int):
        self.inv_index, self.op, self.shard, self.num_shard = inv_index, op, shard, num_shard

    def __str__(self):
        if self.num_shard == 0:
            return self.op + "_combiner"
        return "%s@%d:%d" % (self.op, self.num_shard, self.shard)

    __repr__ = __str__

    def is_combiner(self) -> bool:
        return self.num_shard == 0

    def __eq__(self, o):
        return isinstance(o, TaskName) and str(self) == str(o) and self.inv_index == o.inv_index

    def __hash__(self):
        return hash((self.inv_index, self.op, self.shard, self.num_shard))


class TaskDep:
    """A dependency of a task on partition ``partition`` of the task group
    headed by ``head`` (``exec.TaskDep``)."""

    __slots__ = ("head", "partition", "expand", "combine_key")

    def __init__(self, head: "Task", partition: int = 0, expand: bool = False,
                 combine_key: str = ""):
        self.head = head
        self.partition = partition
        self.expand = expand
        self.combine_key = combine_key

    def num_task(self) -> int:
        return len(self.head.group) if self.head.group else 1

    def task(self, i: int) -> "Task":
        return self.head.group[i] if self.head.group else self.head

    def tasks(self) -> List["Task"]:
        return list(self.head.group) if self.head.group else [self.head]

    NumTask, Task = num_task, task


class TaskSubscriber:
    """Receives task state changes; override :meth:`notify`."""

    def notify(self, task: "Task") -> None:
        pass


class Task:
    """One shard of a pipelined group of slice operations."""

    #: how the planner lowered this task's output exchange ("frames",
    #: "reduce-table"); see exec/compile.py
    lowering = "frames"

    def __init__(self, typ: Slice, name: TaskName, invocation=None,
                 do: Optional[Callable] = None, deps: Optional[List[TaskDep]] = None,
                 partitioner=None, num_partition: int = 1, combiner=None,
                 combine_key: str = "", pragma: Optional[Pragma] = None,
                 slices: Optional[Sequence[Slice]] = None):
        self.type = typ
        self.name = name
        self.invocation = invocation
        self.do = do
        self.deps: List[TaskDep] = deps or []
        self.partitioner = partitioner
        self.num_partition = num_partition
        self.combiner = combiner
        self.combine_key = combine_key
        self.pragma: Pragma = pragma if pragma is not None else Pragmas()
        self.slices: Sequence[Slice] = slices or ()
        self.group: Optional[List["Task"]] = None
        self.scope = metrics.Scope()
        self.stats: Dict[str, int] = {}
        self.status: Any = None  # status.Task display hook
        self._cond = threading.Condition()
        self._state = TaskState.INIT
        self._err: Optional[BaseException] = None
        self._subs: List[TaskSubscriber] = []
        self.consecutive_lost = 0

    # -- types ---------------------------------------------------------------
    def num_out(self) -> int:
        return self.type.num_out()

    def out(self, i: int):
        return self.type.out(i)

    def prefix(self) -> int:
        return self.type.prefix()

    # -- state ---------------------------------------------------------------
    def state(self) -> TaskState:
        # a single attribute read is atomic; taking the condition's lock here made
        # wide shuffles (consumers x producers state checks) contend on it
        return self._state

    def err(self) -> Optional[BaseException]:
        with self._cond:
            return self._err if self._state == TaskState.ERR else None

    def state_err(self):
        with self._cond:
            return self._state, self._err

    def set(self, state: TaskState, err: Optional[BaseException] = None) -> None:
        """Set the state (and error for ERR) and notify waiters/subscribers."""
        with self._cond:
            if state == TaskState.LOST and self._state != TaskState.LOST:
                self.consecutive_lost += 1  # counted once per loss, not per observer
            elif state == TaskState.OK:
                self.consecutive_lost = 0
            self._state = state
            self._err = err if state == TaskState.ERR else None
            subs = list(self._subs)
            self._cond.notify_all()
        if self.status is not None:
            try:
                self.status.print(str(state))
            except Exception:
                pass
        for s in subs:
            s.notify(self)

    def try_start(self) -> bool:
        """Atomically claim the task for running: LOST/INIT -> WAITING.

        Returns True for exactly one caller; concurrent evaluations of a shared
        graph then merely wait on the task (reference ``exec/eval.go:110-127``).
        """
        with self._cond:
            if self._state == TaskState.LOST:
                self._state = TaskState.INIT
            if self._state != TaskState.INIT:
                return False
            self._state = TaskState.WAITING
            subs = list(self._subs)
            self._cond.notify_all()
        for s in subs:
            s.notify(self)
        return True

    def error(self, err: BaseException) -> None:
        self.set(TaskState.ERR, err)

    def errorf(self, fmt: str, *args) -> None:
        self.error(errors.Error(fmt % args if args else fmt))

    def subscribe(self, s: TaskSubscriber) -> None:
        with self._cond:
            self._subs.append(s)

    def unsubscribe(self, s: TaskSubscriber) -> None:
        with self._cond:
            self._subs = [x for x in self._subs if x is not s]

    def wait_state(self, min_state: TaskState, timeout: Optional[float] = None) -> TaskState:
        """Block until the task reaches at least ``min_state``."""
        with self._cond:
            self._cond.wait_for(lambda: self._state >= min_state, timeout)
            return self._state

    Set, Error, Errorf, State, Err = set, error, errorf, state, err
    Subscribe, Unsubscribe, WaitState = subscribe, unsubscribe, wait_state

    # -- graph ---------------------------------------------------------------
    def head(self) -> "Task":
        return self.group[0] if self.group else self

    def phase(self) -> List["Task"]:
        """All tasks evaluated together with this one."""
        return list(self.group) if self.group else [self]

    Phase = phase

    def all(self) -> List["Task"]:
        """Every task reachable from this one, sorted by name."""
        seen: Dict[int, Task] = {}
        stack = [self]
        while stack:
            t = stack.pop()
            if id(t) in seen:
                continue
            seen[id(t)] = t
            for d in t.deps:
                stack.extend(d.tasks())
        return sorted(seen.values(), key=lambda t: str(t.name))

    All = all

    def write_graph(self, w) -> None:
        """Schematic rendering of the graph rooted here (``WriteGraph``)."""
        w.write("tasks:\n")
        rows = []
        for t in self.all():
            outs = ",".join(str(t.out(i)) for i in range(t.num_out()))
            rows.append((str(t.name), outs, "%d [%s]" % (t.num_partition, t.state())))
        w0 = max([len(r[0]) for r in rows] + [0]) + 1
        w1 = max([len(r[1]) for r in rows] + [0]) + 1
        for r in rows:
            w.write("    %s%s%s\n" % (r[0].ljust(w0), r[1].ljust(w1), r[2]))
        w.write("dependencies:\n")
        self._write_deps(w, set())

    def _write_deps(self, w, seen) -> None:
        for d in self.deps:
            for t in d.tasks():
                w.write("    %s:\t%s[%d]\n" % (self.name, t.name, d.partition))
                if id(t) not in seen:
                    seen.add(id(t))
                    t._write_deps(w, seen)

    def graph_string(self) -> str:
        b = io.StringIO()
        self.write_graph(b)
        return b.getvalue()

    WriteGraph, GraphString = write_graph, graph_string

    def __str__(self):
        return "%s(%s)" % (self.name, self.state())

    __repr__ = __str__


def iter_tasks(roots: Sequence[Task], fn: Callable[[Task], None]) -> None:
    """Visit every task reachable from ``roots`` exactly once."""
    seen = set()
    stack = list(roots)
    while stack:
        t = stack.pop()
        if id(t) in seen:
            continue
        seen.add(id(t))
        fn(t)
        for d in t.deps:
            stack.extend(d.tasks())


def make_graph(tasks: Sequence[Task]) -> str:
    """Node-list + edge-list text form used by the golden compile tests
    (same layout as the reference ``exec/testdata/*.graph``)."""
    nodes, edges = set(), set()

    def visit(t: Task):
        nodes.add(str(t.name))
        for d in t.deps:
            for dt in d.tasks():
                edges.add("%s -> %s" % (t.name, dt.name))

    iter_tasks(tasks, visit)
    return "\n".join(sorted(nodes) + sorted(edges)) + "\n"
