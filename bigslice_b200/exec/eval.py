"""Evaluator: schedules a task DAG onto an executor, tolerating task loss.

``evaluate(executor, roots)`` returns when every root task is ``OK`` or raises
when a task fails fatally.  Tasks are handed to ``Executor.run`` as soon as all
their dependencies are ``OK``; ``LOST`` tasks (and their lost dependencies)
are resubmitted, giving lineage-based recomputation; a task lost
``MAX_CONSECUTIVE_LOST`` times in a row fails with ``TooManyTries``.

Dependency bookkeeping is *phase-aware*: a consumer waits on a whole producer
group (one counter per group, not one edge per producer task), so graphs with
thousands of shards per shuffle stay O(tasks).  The bookkeeping core has a C++
implementation (``csrc/native.cpp: EvalCore``) used when built; this module
holds the python fallback with identical behaviour.

Capability parity: reference ``exec/eval.go:26-451`` (Executor interface,
Eval, state/waitlist machine, maxConsecutiveLost) and multiple concurrent
evaluations over a shared graph.
"""
from __future__ import annotations

import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, List, Optional, Sequence, Set

from .. import _native, errors
from .task import Task, TaskState, TaskSubscriber

#: A task lost this many times in a row fails the evaluation (eval.go:30).
MAX_CONSECUTIVE_LOST = 5


class Executor:
    """What the evaluator needs from an execution backend (``exec.Executor``)."""

    def name(self) -> str:
        return type(self).__name__

    def start(self, session) -> Optional[Callable[[], None]]:
        """Start the executor; returns an optional shutdown function."""
        return None

    def run(self, task: Task) -> None:
        """Run ``task`` (state WAITING on entry): set RUNNING, then OK, ERR
        (fatal) or LOST.  Called on a worker thread."""
        raise NotImplementedError

    def reader(self, task: Task, partition: int):
        """A reader over partition ``partition`` of an OK task's output."""
        raise NotImplementedError

    def discard(self, task: Task) -> None:
        """Drop the task's stored output; the task becomes LOST."""

    def eventer(self):
        return None

    def handle_debug(self, mux) -> None:
        pass

    def max_workers(self) -> int:
        return 64


class _PyCore:
    """Pure-python dependency bookkeeping (mirror of native EvalCore)."""

    def __init__(self):
        self.counts: Dict[int, int] = {}
        self.group_missing: Dict[int, Set[int]] = {}
        self.group_waiters: Dict[int, List[int]] = {}
        self.task_groups: Dict[int, List[int]] = {}
        self.todo: List[int] = []

    def set_count(self, task: int, n: int) -> None:
        self.counts[task] = n
        if n == 0:
            self.todo.append(task)

    def has_group(self, g: int) -> bool:
        return g in self.group_missing

    def add_group(self, g: int, missing: Sequence[int]) -> None:
        self.group_missing[g] = set(missing)
        for t in missing:
            self.task_groups.setdefault(t, []).append(g)

    def wait_on(self, task: int, g: int) -> None:
        self.group_waiters.setdefault(g, []).append(task)

    def schedule(self, task: int) -> None:
        self.todo.append(task)

    def done_ok(self, task: int) -> None:
        for g in self.task_groups.pop(task, []):
            s = self.group_missing.get(g)
            if s is None:
                continue
            s.discard(task)
            if not s:
                del self.group_missing[g]
                for c in self.group_waiters.pop(g, []):
                    self.counts[c] -= 1
                    if self.counts[c] == 0:
                        self.todo.append(c)

    def take_todo(self) -> List[int]:
        t, self.todo = self.todo, []
        return t


def _new_core():
    nat = _native.lib()
    if nat is not None and hasattr(nat, "EvalCore"):
        return nat.EvalCore()
    return _PyCore()


class _State:
    def __init__(self):
        self.core = _new_core()
        self.tasks: Dict[int, Task] = {}
        self.pending: Set[int] = set()   # scheduled (todo or submitted)
        self.waiting: Set[int] = set()   # blocked on dependencies
        self.err: Optional[BaseException] = None

    def enqueue(self, task: Task) -> None:
        tid = id(task)
        if tid in self.pending or tid in self.waiting:
            return
        st, err = task.state_err()
        if st == TaskState.OK:
            return
        if st == TaskState.ERR:
            if self.err is None:
                self.err = err or errors.Error("task %s failed" % task.name)
            return
        self.tasks[tid] = task
        if st in (TaskState.WAITING, TaskState.RUNNING):
            self.pending.add(tid)
            self.core.schedule(tid)
            return
        # INIT or LOST: make sure all dependencies are (being) computed.
        nunmet = 0
        for dep in task.deps:
            deps = dep.tasks()
            g = id(dep.head)
            if self.core.has_group(g):
                if any(t.state() != TaskState.OK for t in deps):
                    nunmet += 1
                    self.core.wait_on(tid, g)
                continue
            missing = [t for t in deps if t.state() != TaskState.OK]
            if not missing:
                continue
            nunmet += 1
            self.core.add_group(g, [id(t) for t in missing])
            self.core.wait_on(tid, g)
            for t in missing:
                self.enqueue(t)
                if self.err is not None:
                    return
        if nunmet == 0:
            self.pending.add(tid)
        else:
            self.waiting.add(tid)
        self.core.set_count(tid, nunmet)

    def runnable(self) -> List[Task]:
        out = []
        for tid in self.core.take_todo():
            self.waiting.discard(tid)
            self.pending.add(tid)
            out.append(self.tasks[tid])
        return out

    def return_(self, task: Task) -> None:
        tid = id(task)
        self.pending.discard(tid)
        st, err = task.state_err()
        if st == TaskState.OK:
            self.core.done_ok(tid)
        elif st == TaskState.ERR:
            self.err = err or errors.Error("task %s failed" % task.name)
        elif st == TaskState.LOST:
            # Re-traverse dependencies: lost producers are recomputed too.
            self.enqueue(task)
        else:
            # Restarted by a concurrent evaluation of the same graph since the
            # waiter fired: wait for (or resubmit) it again.
            self.enqueue(task)

    def idle(self) -> bool:
        return not self.pending


class _Waiter(TaskSubscriber):
    """One-shot: posts the task on ``q`` once it reaches a terminal state."""

    def __init__(self, task: Task, q: "queue.Queue"):
        self.task, self.q = task, q
        self._fired = False
        self._lock = threading.Lock()

    def arm(self):
        self.task.subscribe(self)
        self.notify(self.task)

    def notify(self, task: Task) -> None:
        if task.state() < TaskState.OK:
            return
        with self._lock:
            if self._fired:
                return
            self._fired = True
        task.unsubscribe(self)
        self.q.put(task)


def evaluate(executor: Executor, roots: Sequence[Task], group=None,
             pool: Optional[ThreadPoolExecutor] = None,
             cancel: Optional[threading.Event] = None) -> None:
    """Evaluate ``roots`` (and everything they depend on) with ``executor``."""
    state = _State()
    for r in roots:
        state.enqueue(r)
    own_pool = pool is None
    if own_pool:
        pool = ThreadPoolExecutor(max_workers=max(4, executor.max_workers()),
                                  thread_name_prefix="bigslice-eval")
    doneq: "queue.Queue[Task]" = queue.Queue()
    eventer = executor.eventer()
    try:
        while True:
            if state.err is not None:
                raise state.err
            for task in state.runnable():
                if task.try_start():
                    pool.submit(_run, executor, task)
                _Waiter(task, doneq).arm()
            if state.idle():
                if state.waiting:
                    raise errors.fatal("evaluator stuck: %d tasks wait on dependencies "
                                       "that are not scheduled" % len(state.waiting))
                break
            _update_status(group, state)
            while True:
                try:
                    task = doneq.get(timeout=0.2)
                    break
                except queue.Empty:
                    if cancel is not None and cancel.is_set():
                        raise errors.Error("evaluation canceled", errors.Kind.Canceled)
            while True:
                st = task.state()
                if st == TaskState.LOST and task.consecutive_lost >= MAX_CONSECUTIVE_LOST:
                    task.error(errors.TooManyTriesError(
                        "task %s lost on %d consecutive attempts" % (task.name, task.consecutive_lost)))
                if eventer is not None:
                    eventer.event("bigslice:taskComplete", "name", str(task.name),
                                  "state", str(task.state()))
                state.return_(task)
                try:
                    task = doneq.get_nowait()
                except queue.Empty:
                    break
        if state.err is not None:
            raise state.err
        for r in roots:
            if r.state() != TaskState.OK:
                # A root was lost after we saw it complete: go around again.
                return evaluate(executor, roots, group, None if own_pool else pool, cancel)
    finally:
        _update_status(group, state, final=True)
        if own_pool:
            pool.shutdown(wait=False)


def _run(executor: Executor, task: Task) -> None:
    try:
        executor.run(task)
    except BaseException as e:  # executor bugs must not hang the evaluation
        if task.state() < TaskState.OK:
            task.error(errors.fatal("executor %s failed running %s" % (executor.name(), task.name), e))


def _update_status(group, state: _State, final: bool = False) -> None:
    if group is None:
        return
    try:
        group.printf("tasks: waiting %d, running/scheduled %d%s",
                     len(state.waiting), len(state.pending), " (done)" if final else "")
    except Exception:
        pass


Eval = evaluate
