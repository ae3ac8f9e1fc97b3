"""Sessions: the entry point that runs slice computations.

``start(...)`` creates a :class:`Session` bound to an executor; ``Session.
run(func, *args)`` invokes a registered ``Func``, compiles the slice it
returns, evaluates the task graph and returns a :class:`Result`.  A Result is
itself a Slice and may be passed to later ``run`` calls (iterative
computation re-uses its tasks; its tensors stay resident in HBM until
``discard``).

Capability parity: reference ``exec/session.go:68-461`` (Start, options
Local/Bigmachine/Parallelism/MaxLoad/Status/Eventer/TracePath/
MachineCombiners, Session.{Run,Must,Discard,Shutdown,Parallelism,MaxLoad,
Status,HandleDebug}, Result.{Scanner,Scope,Discard}).
"""
from __future__ import annotations

import logging
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, Dict, List, Optional

from .. import errors, metrics, typecheck
from ..frame import Frame, concat
from ..func import FuncValue, Invocation
from ..utils.spans import span
from ..sliceio import MultiReader, Reader, Scanner
from ..slices import Slice, make_name
from ..status import Status
from .compile import ExecInvocation, compile_slice
from .eval import Executor, evaluate
from .local import LocalExecutor
from .task import Task, TaskState, iter_tasks
from .tracer import Tracer

log = logging.getLogger("bigslice")

DefaultMaxLoad = 0.95


class NopEventer:
    def event(self, typ: str, *kv) -> None:
        pass


class ListEventer(NopEventer):
    """Collects events in memory (tests, debugging)."""

    def __init__(self):
        self.events: List[tuple] = []
        self._lock = threading.Lock()

    def event(self, typ: str, *kv) -> None:
        with self._lock:
            self.events.append((typ, dict(zip(kv[::2], kv[1::2]))))


class Option:
    def __init__(self, fn: Callable[["Session"], None]):
        self.fn = fn

    def __call__(self, sess: "Session") -> None:
        self.fn(sess)


def _set_executor(factory):
    def apply(s: "Session"):
        s._executor_factory = factory
    return Option(apply)


#: Run everything in-process on CPU tensors (``exec.Local``).
Local = _set_executor(lambda s: LocalExecutor(s.p, None, store=s.store_))


def LocalDevice(device) -> Option:
    """In-process executor with frames resident on ``device`` (one B200)."""
    return _set_executor(lambda s: LocalExecutor(s.p, device, store=s.store_))


def Spmd(**kw) -> Option:
    """One rank per GPU over ``torch.distributed`` (the multi-GPU executor;
    takes the place of ``exec.Bigmachine(system)``)."""
    def factory(s):
        from .spmd import SpmdExecutor
        return SpmdExecutor(parallelism=s.p, **kw)
    return _set_executor(factory)


Bigmachine = Spmd


def Parallelism(p: int) -> Option:
    if p <= 0:
        raise ValueError("parallelism must be positive")
    return Option(lambda s: setattr(s, "p", p))


def MaxLoad(f: float) -> Option:
    if not (0 < f <= 1.0):
        raise ValueError("max load must be in (0, 1]")
    return Option(lambda s: setattr(s, "max_load_", f))


def StatusOpt(status: Status) -> Option:
    return Option(lambda s: setattr(s, "status_", status))


def EventerOpt(e) -> Option:
    return Option(lambda s: setattr(s, "eventer_", e))


def TracePath(path: str) -> Option:
    return Option(lambda s: setattr(s, "trace_path", path))


#: Share combiners per machine/rank rather than per task (experimental in the
#: reference, ``exec/session.go:166-176``).  On the Local executors the planner tags
#: the producers with a shared combine key.  On the SPMD executor this is how the
#: Reduce fast path *always* works — all producer tasks of a rank aggregate into
#: one per-rank HBM table (``SpmdExecutor._group_table``) — so the option changes
#: nothing there.
MachineCombiners = Option(lambda s: setattr(s, "machine_combiners", True))


def StoreOpt(store_or_path) -> Option:
    """Keep task outputs in a :class:`~bigslice_b200.exec.store.Store` (a path
    selects a ``FileStore``): outputs survive in files, are read back with
    resumable readers and can be lost/discarded individually (``exec/store.go``)."""
    def apply(s: "Session"):
        from .store import FileStore
        s.store_ = FileStore(store_or_path) if isinstance(store_or_path, str) else store_or_path
    return Option(apply)


def Journal(directory: Optional[str] = None) -> Option:
    """Record every invocation (Func index, arguments, ``Result`` references) in a
    compressed on-disk journal (``exec/invocation.py: InvDiskCache``), the way the
    reference's driver keeps encoded invocations for its workers
    (``exec/bigmachine_inv_disk_cache.go``).  An argument that cannot be encoded
    fails the run immediately, naming the type."""
    def apply(s: "Session"):
        from .invocation import InvDiskCache
        s.journal_ = InvDiskCache(directory)
    return Option(apply)


class Session:
    """Runs slice computations on an executor."""

    _index_lock = threading.Lock()
    _next_index = 0

    def __init__(self):
        with Session._index_lock:
            self.index = Session._next_index
            Session._next_index += 1
        self.p = 0
        self.max_load_ = 0.0
        self.status_: Optional[Status] = None
        self.eventer_ = NopEventer()
        self.trace_path = ""
        self.machine_combiners = False
        self.store_ = None
        self.journal_ = None
        self._executor_factory = None
        self.executor: Optional[Executor] = None
        self.tracer = Tracer()
        self.roots: Dict[int, Task] = {}
        self._lock = threading.Lock()
        self._shutdown: Optional[Callable[[], None]] = None
        self._status_lock = threading.Lock()
        self._results: List["Result"] = []
        self._pool: Optional[ThreadPoolExecutor] = None

    # -- lifecycle -----------------------------------------------------------
    def _start(self) -> None:
        if self.p == 0:
            self.p = 1
        if self.max_load_ == 0:
            self.max_load_ = DefaultMaxLoad
        if self._executor_factory is None:
            self._executor_factory = lambda s: LocalExecutor(s.p, None, store=s.store_)
        self.executor = self._executor_factory(self)
        self._shutdown = self.executor.start(self)
        self.eventer_.event("bigslice:sessionStart", "executorType", self.executor.name(),
                            "parallelism", self.p, "maxLoad", self.max_load_)

    def _eval_pool(self) -> ThreadPoolExecutor:
        """Worker threads shared by all evaluations of this session (thread
        start-up would otherwise dominate short invocations)."""
        with self._lock:
            if self._pool is None:
                self._pool = ThreadPoolExecutor(max_workers=max(4, self.executor.max_workers()),
                                                thread_name_prefix="bigslice-eval")
            return self._pool

    def shutdown(self) -> None:
        """Release executor resources and write the trace if configured."""
        with self._lock:
            pool, self._pool = self._pool, None
        if pool is not None:
            pool.shutdown(wait=False)
        if self._shutdown is not None:
            s, self._shutdown = self._shutdown, None
            s()
        if self.trace_path:
            try:
                self.tracer.write(self.trace_path)
            except OSError as e:
                log.error("writing trace to %s: %s", self.trace_path, e)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.shutdown()

    def parallelism(self) -> int:
        return self.p

    def max_load(self) -> float:
        return self.max_load_

    def status(self) -> Optional[Status]:
        return self.status_

    # -- running -------------------------------------------------------------
    def run(self, funcv, *args) -> "Result":
        """Invoke ``funcv`` with ``args`` and evaluate the resulting slice."""
        file, line = typecheck.caller_location()
        return self._run(funcv, args, file, line)

    def must(self, funcv, *args) -> "Result":
        """Like run; errors are fatal to the caller (``Session.Must``)."""
        file, line = typecheck.caller_location()
        try:
            return self._run(funcv, args, file, line)
        except typecheck.TypecheckError:
            raise
        except BaseException as e:
            log.error("session.must: %s", e)
            raise

    def _run(self, funcv, args, file: str, line: int) -> "Result":
        if not isinstance(funcv, FuncValue):
            raise typecheck.TypecheckError(
                "session.run: %r is not a bigslice.Func" % (funcv,), file, line)
        location = "%s:%d" % (file, line)
        with typecheck.location(file, line), span("run/compile"):
            with self._status_lock:
                inv = funcv.invocation(location, *args)
            einv = ExecInvocation(inv)
            if self.journal_ is not None:
                from .invocation import encode_invocation
                try:
                    self.journal_.get_or_create(inv.index, lambda w: w.write(encode_invocation(einv))).close()
                except errors.Error:
                    raise
                except Exception as e:  # noqa: BLE001
                    raise errors.fatal("session.run: invocation %d cannot be encoded: %s" % (inv.index, e), e)

            def do_compile(ei):
                sl = ei.invoke()
                return sl, compile_slice(ei, sl, self.machine_combiners)
            hook = getattr(self.executor, "compile_invocation", None)
            slice_, tasks = hook(einv, do_compile) if hook is not None else do_compile(einv)
        einv.env.freeze()
        group = None
        status_tasks = []
        if self.status_ is not None:
            group = self.status_.groupf("run %s [%d] tasks", location, inv.index)
            seen: List[Task] = []
            iter_tasks(tasks, seen.append)
            for t in sorted(seen, key=lambda t: str(t.name)):
                t.status = group.start(str(t.name))
                status_tasks.append(t.status)
        maint = stop_log = slice_group = None
        if self.status_ is not None:
            from .slicestatus import SliceGroupMaintainer, log_invocation
            slice_group = self.status_.groupf("run %s [%d] slices", location, inv.index)
            maint = SliceGroupMaintainer(slice_group, tasks)
            stop_log = threading.Event()
            threading.Thread(target=log_invocation, daemon=True,
                             args=(inv.index, location, maint, stop_log)).start()
        with self._lock:
            for t in tasks:
                self.roots[id(t)] = t
        t0 = time.time()
        begin = getattr(self.executor, "begin_invocation", None)
        if begin is not None:
            with span("run/begin_invocation"):
                begin(tasks)
        failed = True
        try:
            with span("run/evaluate"):
                evaluate(self.executor, tasks, group, pool=self._eval_pool())
            failed = False
        finally:
            end = getattr(self.executor, "end_invocation", None)
            if end is not None:
                end(tasks, failed)
            log.debug("invocation %d (%s) evaluated in %.3fs", inv.index, location, time.time() - t0)
            if group is not None:
                for st in status_tasks:
                    st.done()
                group.done()
            if maint is not None:
                stop_log.set()
                maint.close()
                slice_group.done()
        res = Result(slice_, self, inv.index, tasks)
        with self._lock:
            self._results.append(res)
        return res

    Run, Must, Shutdown, Parallelism, MaxLoad, Status = run, must, shutdown, parallelism, max_load, status

    def discard(self) -> None:
        """Discard the outputs of every task this session has run."""
        with self._lock:
            roots = list(self.roots.values())
        seen: List[Task] = []
        iter_tasks(roots, seen.append)
        with ThreadPoolExecutor(max_workers=64) as pool:
            list(pool.map(self.executor.discard, seen))

    Discard = discard

    def handle_debug(self, mux) -> None:
        """Register ``/debug/*`` handlers on ``mux`` (a DebugMux)."""
        from . import graph
        graph.register(self, mux)
        self.executor.handle_debug(mux)

    HandleDebug = handle_debug


class Result(Slice):
    """The outcome of ``Session.run``: readable, and reusable as a Func arg."""

    def __init__(self, slice_: Slice, sess: Session, inv_index: int, tasks: List[Task]):
        Slice.__init__(self, make_name("result"), slice_.typ, slice_.num_shard())
        self.slice = slice_
        self.sess = sess
        self.inv_index = inv_index
        self.tasks = tasks
        self._scope: Optional[metrics.Scope] = None
        self._scope_lock = threading.Lock()

    def reader(self, shard, deps):
        return self.sess.executor.reader(self.tasks[shard], 0)

    def _open(self) -> Reader:
        ex = self.sess.executor
        open_result = getattr(ex, "result_reader", None)
        if open_result is not None:
            return open_result(self)
        return MultiReader([_EvalReader(self.sess, t) for t in self.tasks])

    def scanner(self) -> Scanner:
        """A scanner over all rows, shard after shard (``Result.Scanner``)."""
        return Scanner(self._open(), self.typ)

    def frames(self) -> List[Frame]:
        """All output batches (device frames stay on device)."""
        r = self._open()
        out = [f for f in r]
        r.close()
        return out

    def to_frame(self, device=None) -> Frame:
        """The whole result concatenated into one frame."""
        fr = [f.to(device) if device is not None else f for f in self.frames()]
        return concat(fr, self.typ) if fr else Frame.empty(self.typ)

    def rows(self) -> List[tuple]:
        return [r for f in self.frames() for r in f.to("cpu").rows()]

    def scope(self) -> metrics.Scope:
        """Merged metrics of all tasks of this result (``Result.Scope``)."""
        with self._scope_lock:
            if self._scope is None:
                sc = metrics.Scope()
                seen: List[Task] = []
                iter_tasks(self.tasks, seen.append)
                for t in seen:
                    sc.merge(t.scope)
                self._scope = sc
            return self._scope

    def discard(self) -> None:
        """Drop this result's task outputs; tasks become LOST and are
        recomputed if needed again (``Result.Discard``)."""
        seen: List[Task] = []
        iter_tasks(self.tasks, seen.append)
        for t in seen:
            self.sess.executor.discard(t)

    Scanner_, Scope, Discard = scanner, scope, discard


class _EvalReader(Reader):
    """Opens a task's output lazily, re-evaluating the task if its output was
    lost since the run (fault-tolerant scans, reference
    ``exec/bigmachine.go:1500-1513``)."""

    def __init__(self, sess: Session, task: Task):
        self.sess, self.task = sess, task
        self.r: Optional[Reader] = None
        self.tries = 0

    def read(self, n=None):
        while True:
            if self.r is None:
                if self.task.state() != TaskState.OK:
                    evaluate(self.sess.executor, [self.task])
                try:
                    self.r = self.sess.executor.reader(self.task, 0)
                except errors.TaskLostError:
                    self.tries += 1
                    if self.tries > 5:
                        raise
                    self.task.set(TaskState.LOST)
                    continue
            return self.r.read(n)

    def close(self):
        if self.r is not None:
            self.r.close()


def start(*options: Option) -> Session:
    """Create and start a session (``exec.Start``)."""
    s = Session()
    for o in options:
        o(s)
    s._start()
    return s


Start = start
