"""Local executor: runs tasks in-process on CPU tensors or on one B200.

Each task's output is partitioned in memory into a :class:`TaskBuffer`;
consumers read partitions directly (the "shuffle" is a pointer hand-off).
``parallelism`` bounds concurrently running tasks (procs), honouring
``Procs(n)`` and ``Exclusive`` pragmas.  With ``device='cuda'`` frames live in
HBM and partitioning, combining, sorting and compaction use the sm_100a
kernels; this is the single-GPU execution path.  The multi-GPU SPMD executor
(:mod:`.spmd`) reuses this class for per-rank task execution.

Failure model matches the reference: a fatal error (user panic, application
error) puts the task in ``ERR``; anything else — including a missing or
discarded dependency — marks it ``LOST`` so the evaluator recomputes it.

Capability parity: reference ``exec/local.go:24-265`` (localExecutor.Run/
Reader/Discard, depReaders, bufferOutput) and ``exec/buffer.go:19-95``
(taskBuffer).
"""
from __future__ import annotations

import os
import threading
import time
import traceback
from typing import Dict, List, Optional

import torch

from .. import context as _context
from .. import errors, slicefunc
from ..frame import Frame
from ..sliceio import EmptyReader, FrameReader, MultiReader, Reader
from .combiner import Combiner
from ..utils.spans import span
from .eval import Executor
from .task import Task, TaskState


_NO_FRAMES: tuple = ()


class _SparseParts:
    """``parts[p]`` -> frames of partition ``p`` (an empty tuple when none);
    iterates like a dense list of per-partition frame sequences."""

    __slots__ = ("n", "d")

    def __init__(self, n: int):
        self.n = n
        self.d: Dict[int, List[Frame]] = {}

    def add(self, p: int, f: Frame) -> None:
        if not 0 <= p < self.n:
            raise IndexError("partition %d out of range [0, %d)" % (p, self.n))
        lst = self.d.get(p)
        if lst is None:
            self.d[p] = [f]
        else:
            lst.append(f)

    def __getitem__(self, p: int):
        if not 0 <= p < self.n:
            raise IndexError("partition %d out of range [0, %d)" % (p, self.n))
        return self.d.get(p, _NO_FRAMES)

    def __len__(self) -> int:
        return self.n

    def __iter__(self):
        for p in range(self.n):
            yield self.d.get(p, _NO_FRAMES)


class TaskBuffer:
    """In-memory output of one task: frames per partition (``taskBuffer``)."""

    def __init__(self, npartition: int):
        # Sparse: with hundreds of shards almost every (task, partition) pair is
        # empty, and hundreds of thousands of empty lists are what the garbage
        # collector then walks on every full collection.
        self.parts: "_SparseParts" = _SparseParts(npartition)
        self.combined_unique = False
        self.nbytes = 0
        self.nrows = 0

    def append(self, partition: int, f: Frame) -> None:
        if len(f):
            self.parts.add(partition, f)
            self.nrows += len(f)

    def reader(self, partition: int) -> Reader:
        frames = self.parts[partition]
        if not frames:
            return EmptyReader()
        return _BufferReader(frames)

    def rows(self, partition: int) -> int:
        return sum(len(f) for f in self.parts[partition])


class _BufferReader(Reader):
    """Streams buffered frames; reading after close is a bug (buffer.go:54)."""

    def __init__(self, frames: List[Frame]):
        self.frames: Optional[List[Frame]] = list(frames)
        self.i = 0
        self.cur: Optional[Reader] = None

    def read(self, n=None):
        if self.frames is None:
            raise RuntimeError("TaskBuffer reader: read after close")
        while True:
            if self.cur is None:
                if self.i >= len(self.frames):
                    return None
                self.cur = FrameReader(self.frames[self.i])
                self.i += 1
            f = self.cur.read(n)
            if f is not None:
                return f
            self.cur = None

    def close(self):
        self.frames = None


class _StartGate:
    """Every task runs its first few milliseconds alone.

    A task over a handful of rows is a few hundred short GIL-bound Python/torch
    calls.  Several such tasks in flight hand the GIL (and the OS thread) back and
    forth on every call, and the evaluation runs several times *slower* than on one
    thread (measured, 400 shards x 3 shuffle phases on 8 procs: 1.8 s serialised vs
    14-18 s).  So a task starts under this gate; a watchdog opens the gate for the
    next task once the holder has run for ``SLICE_S`` — tiny tasks therefore run
    back to back, while tasks that compute, sleep or wait for I/O overlap exactly as
    before after their first few milliseconds.  The holder never blocks others for
    longer than the slice, so the gate cannot deadlock."""

    SLICE_S = 0.004
    HEAVY_RUN = 8        # this many tasks in a row outlived their slice: stop gating
    TINY_ROWS = 4096     # an ungated task this small switches gating back on

    def __init__(self):
        self.expired_run = 0
        self.waiters = 0
        self.lock = threading.Lock()      # the gate itself (released by holder or watchdog)
        self.meta = threading.Lock()      # guards holder/since
        self.holder: Optional[object] = None
        self.since = 0.0
        self.wake = threading.Event()
        self.closed = False
        self.thread: Optional[threading.Thread] = None

    def enter(self) -> Optional[object]:
        if self.expired_run >= self.HEAVY_RUN:
            return None  # heavy phase: tasks overlap from the start (see observe())
        if not self.lock.acquire(blocking=False):
            # contended: only now is the watchdog needed (it opens the gate once the
            # current holder has had its slice)
            if self.thread is None:
                with self.meta:
                    if self.thread is None:
                        self.thread = threading.Thread(target=self._watch, daemon=True,
                                                       name="bigslice-start-gate")
                        self.thread.start()
            with self.meta:
                self.waiters += 1
            self.wake.set()
            self.lock.acquire()
            with self.meta:
                self.waiters -= 1
        tok = object()
        with self.meta:
            self.holder, self.since = tok, time.perf_counter()
        return tok

    def leave(self, tok: object) -> None:
        with self.meta:
            if self.holder is tok:
                self.holder = None
                self.expired_run = 0  # finished inside its slice: tasks are tiny
                self.lock.release()

    def observe(self, rows: int) -> None:
        """An ungated task finished having processed ``rows`` rows."""
        if rows < self.TINY_ROWS:
            self.expired_run = 0

    def _watch(self) -> None:
        while not self.closed:
            self.wake.wait()
            if self.closed:
                return
            with self.meta:
                if self.waiters == 0:
                    self.wake.clear()
                    continue
                age = time.perf_counter() - self.since if self.holder is not None else None
                if age is not None and age >= self.SLICE_S:
                    self.holder = None        # the holder keeps running, ungated
                    self.expired_run += 1
                    self.lock.release()
                    age = None
            # a holder inside its slice: look again when the slice ends; otherwise the
            # gate is changing hands, check back shortly
            time.sleep(max(self.SLICE_S - age, 0.0002) if age is not None else 0.0005)

    def close(self) -> None:
        self.closed = True
        self.wake.set()


class _Limiter:
    """Counting semaphore with multi-unit acquire (procs)."""

    def __init__(self, n: int):
        self.n = n
        self.avail = n
        self.wide_waiters = 0
        self.cond = threading.Condition()

    def acquire(self, k: int) -> None:
        k = min(k, self.n)
        with self.cond:
            if k > 1:
                self.wide_waiters += 1
            try:
                self.cond.wait_for(lambda: self.avail >= k)
            finally:
                if k > 1:
                    self.wide_waiters -= 1
            self.avail -= k

    def release(self, k: int) -> None:
        k = min(k, self.n)
        with self.cond:
            self.avail += k
            # Waking every waiter on every release is a thundering herd when hundreds
            # of runnable tasks queue behind a few procs; single-unit waiters only need
            # one wake-up per freed unit.  Multi-unit waiters may need several releases
            # to add up, so their presence falls back to waking everyone.
            if self.wide_waiters:
                self.cond.notify_all()
            else:
                self.cond.notify(k)


def split_by_partition(f: Frame, ids: torch.Tensor, npart: int) -> List[Frame]:
    """Split rows of ``f`` into ``npart`` frames by destination id, keeping
    the relative row order inside each partition (stable)."""
    if f.device.type == "cuda":
        from ..ops import gpu
        if gpu.available():
            return gpu.partition_frame(f, ids, npart)
    ids = ids.to(torch.int64).cpu() if ids.device.type != "cpu" else ids.to(torch.int64)
    order = counts = None
    if ids.numel() >= 2048:
        # C++ stable counting sort, GIL released (csrc/native.cpp: host_partition_perm)
        from .. import _native
        nat = _native.lib()
        if nat is not None and hasattr(nat, "host_partition_perm"):
            ids = ids.contiguous()
            perm = torch.empty_like(ids)
            cnt = torch.empty(npart, dtype=torch.int64)
            if nat.host_partition_perm(ids.data_ptr(), ids.numel(), npart, perm.data_ptr(),
                                       cnt.data_ptr()) < 0:
                raise errors.fatal("partitioner returned a shard outside [0, %d)" % npart)
            order, counts = perm, cnt.tolist()
    if order is None:
        order = torch.sort(ids, stable=True)[1]
        counts = torch.bincount(ids, minlength=npart).tolist()
    g = f.take(order)
    out, off = [], 0
    empty = None
    for c in counts:
        if c == 0:  # one shared empty frame: with many shards most partitions of a tile are empty
            if empty is None:
                empty = g.slice(0, 0)
            out.append(empty)
            continue
        out.append(g.slice(off, off + c))
        off += c
    return out


class _StoredBuffer:
    """Task output parked in a Store (files) instead of memory."""

    def __init__(self, store, task: Task, buf: TaskBuffer, device):
        self.store, self.name, self.typ, self.device = store, task.name, task.type.typ, device
        self.npart = len(buf.parts)
        self.nrows = buf.nrows
        for p, frames in enumerate(buf.parts):
            w = store.create(task.name, p)
            try:
                for f in frames:
                    w.write(f.to("cpu") if f.device.type != "cpu" else f)
                w.commit()
            except BaseException:
                w.discard()
                raise
        self.parts = _SparseParts(self.npart)  # nothing resident

    def reader(self, partition: int) -> Reader:
        from ..sliceio import RetryReader
        return RetryReader(lambda off: self.store.open(self.name, partition, self.typ, off, self.device))

    def discard(self):
        for p in range(self.npart):
            self.store.discard(self.name, p)


class LocalExecutor(Executor):
    def __init__(self, parallelism: int = 1, device=None, store=None):
        self.store = store
        self.parallelism = max(1, int(parallelism))
        self.device = torch.device(device) if device is not None else None
        if self.device is not None and self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.limiter = _Limiter(self.parallelism)
        # host executors with several procs start tasks through the gate (GIL-bound
        # tiny tasks); device executors are bound by the GPU, not by the GIL
        self.gate: Optional[_StartGate] = (
            _StartGate() if self.parallelism > 1 and self.device is None
            and not os.environ.get("BIGSLICE_NO_START_GATE") else None)
        self.lock = threading.Lock()
        self.buffers: Dict[int, TaskBuffer] = {}
        self._pindex: Dict[int, tuple] = {}
        self.session = None
        self.tracer = None

    def name(self) -> str:
        return "local"

    def start(self, session):
        self.session = session
        self.tracer = getattr(session, "tracer", None)
        return None

    def max_workers(self) -> int:
        return max(8, self.parallelism * 4)

    def eventer(self):
        return getattr(self.session, "eventer_", None)

    # -- running -------------------------------------------------------------
    def run(self, task: Task) -> None:
        procs = task.pragma.procs()
        if task.pragma.exclusive():
            procs = self.parallelism
        if self.device is not None and self.device.type == "cuda":
            torch.cuda.set_device(self.device)  # CUDA's current device is per thread
        self.limiter.acquire(procs)
        if self.tracer is not None:
            self.tracer.event(self._pid(), task, "B")
        tok = self.gate.enter() if self.gate is not None else None
        try:
            self._run(task)
        finally:
            if tok is not None:
                self.gate.leave(tok)
            elif self.gate is not None:
                self.gate.observe(int(task.stats.get("inrecords", 0)) if task.stats else 0)
            self.limiter.release(procs)
            if self.tracer is not None:
                st = task.state()
                args = {}
                if st == TaskState.ERR:
                    args = {"error": str(task.err()), "error_type": "fatal"}
                elif st == TaskState.LOST:
                    args = {"error_type": "lost"}
                args.update({k: v for k, v in task.stats.items()})
                self.tracer.event(self._pid(), task, "E", **args)

    def _pid(self):
        return "local"

    def _run(self, task: Task, device=None) -> None:
        with span("task/run"):
            self._run_task(task, device)

    def _run_task(self, task: Task, device=None) -> None:
        device = device if device is not None else self.device
        try:
            ins = self.dep_readers(task)
        except errors.TaskLostError:
            task.set(TaskState.LOST)
            return
        except BaseException as e:
            if errors.is_fatal(e):
                task.error(e)
            else:
                task.set(TaskState.LOST)
            return
        task.set(TaskState.RUNNING)
        task.scope.reset()
        ctx = slicefunc.Context(task.scope, task.name.shard, device)
        try:
            with _context.use(ctx):
                out = task.do(ins)
                buf = self.buffer_output(task, out, ctx)
        except errors.TaskLostError:
            task.set(TaskState.LOST)
            return
        except errors.Error as e:
            if errors.is_fatal(e):
                task.error(e)
            else:
                task.set(TaskState.LOST)
            return
        except BaseException as e:
            if self.is_device_error(e):
                self.last_device_error = e
                task.set(TaskState.LOST)
                return
            # A panic in user code is an application error: fatal.
            tb = traceback.format_exc()
            task.error(errors.fatal("panic while evaluating slice: %s\n%s" % (e, tb), e))
            return
        finally:
            for r in ins:
                try:
                    r.close()
                except Exception:
                    pass
        if self.store is not None and task.num_out() > 0:
            try:
                buf = _StoredBuffer(self.store, task, buf, self.device)
            except BaseException as e:  # storage trouble is not an application error
                task.set(TaskState.LOST)
                return
        # Single dict operations are atomic under the GIL; this map is hit by every
        # task and a Python-level lock around it turns into a convoy (a holder that
        # loses the GIL keeps every other task waiting for a whole switch interval).
        self.buffers[id(task)] = buf
        task.set(TaskState.OK)

    last_device_error: Optional[BaseException] = None

    @staticmethod
    def is_device_error(e: BaseException) -> bool:
        """Failures of the device rather than of the program: the task is LOST
        (retried, possibly elsewhere), not failed."""
        if isinstance(e, torch.cuda.OutOfMemoryError):
            return True
        msg = str(e)
        return isinstance(e, RuntimeError) and ("CUDA error" in msg or "cudaError" in msg
                                                or "device-side assert" in msg)

    def dep_readers(self, task: Task) -> List[Reader]:
        """Readers for every dependency (expanded deps contribute one reader
        per producer task), failing with TaskLost if any output is gone."""
        ins: List[Reader] = []
        plain = type(self).reader is LocalExecutor.reader  # not overridden by a subclass
        for dep in task.deps:
            readers = []
            dts = dep.tasks()
            if plain and len(dts) > 8:
                # wide shuffle: one lock acquisition for all producers, and only the
                # producers that have rows for this partition contribute a reader
                for dt in dts:
                    if dt.state() != TaskState.OK:
                        raise errors.TaskLostError("dependency %s is not available" % dt.name)
                get = self.buffers.get
                bufs = [get(id(dt)) for dt in dts]
                for dt, b in zip(dts, bufs):
                    if b is None:
                        raise errors.TaskLostError("output of %s is not available" % dt.name)
                idx = self._partition_index(dts, bufs)
                if idx is not None:
                    # producers x consumers pairs are mostly empty: the group's output is
                    # indexed once (partition -> producers with rows), not scanned per consumer
                    readers = [b.reader(dep.partition) for b in idx.get(dep.partition, ())]
                else:
                    for b in bufs:
                        if isinstance(b, TaskBuffer) and not b.parts[dep.partition]:
                            continue
                        readers.append(b.reader(dep.partition))
                if dep.expand:
                    ins.extend(readers)
                else:
                    ins.append(MultiReader(readers) if readers else EmptyReader())
                continue
            for dt in dts:
                if dt.state() != TaskState.OK:
                    raise errors.TaskLostError("dependency %s is not available" % dt.name)
                readers.append(self.reader(dt, dep.partition))
            if dep.expand:
                ins.extend(readers)
            else:
                ins.append(MultiReader(readers))
        return ins

    def _partition_index(self, dts, bufs):
        """partition -> [buffers that hold rows for it], for a wide group whose
        outputs are all in memory; cached per group until any of its buffers is
        replaced (a recomputed task gets a new buffer object)."""
        if not all(isinstance(b, TaskBuffer) for b in bufs):
            return None
        key = id(dts[0])
        stamp = hash(tuple(map(id, bufs)))
        hit = self._pindex.get(key)
        if hit is not None and hit[0] == stamp and hit[1] == len(bufs):
            return hit[2]
        idx: Dict[int, List[TaskBuffer]] = {}
        for b in bufs:  # producer order is kept inside every partition
            for p in b.parts.d:
                idx.setdefault(p, []).append(b)
        if len(self._pindex) > 64:
            self._pindex.clear()
        self._pindex[key] = (stamp, len(bufs), idx)  # racing builders store equal indexes
        return idx

    def buffer_output(self, task: Task, out: Reader, ctx) -> TaskBuffer:
        """Drain the task's pipeline, partitioning (and map-side combining)
        its output (``bufferOutput`` + inline combine)."""
        npart = task.num_partition
        buf = TaskBuffer(npart)
        # Map-side combining happens *before* partitioning: one (spilling) combiner
        # per task instead of one per partition, and only the combined rows are
        # split.  The combiner's output is sorted by key and the split is stable, so
        # every partition still receives its rows combined and in key order — the
        # contract of the reference's per-partition combiners (worker.runCombine).
        combiner: Optional[Combiner] = None
        if task.combiner is not None:
            combiner = Combiner(task.type.typ if hasattr(task.type, "typ") else task.type,
                                task.combiner, task.name.op, ctx=ctx)
        t_read = t_write = 0.0
        nrec = 0

        from ..cogroup import broadcast_partitioner
        replicate = task.partitioner is broadcast_partitioner

        def emit(f: Frame) -> None:
            if replicate:       # a replicated dependency: every consumer shard reads every row
                for p in range(npart):
                    buf.append(p, f)
            elif npart > 1:
                parts = split_by_partition(f, task.partitioner(f, npart), npart)
                # the device partition kernel does not keep the row order inside a
                # partition; combined output must stay sorted by key for the consumers
                resort = combiner is not None and f.device.type == "cuda"
                for p, g in enumerate(parts):
                    if len(g):
                        buf.append(p, g.sorted() if resort else g)
            else:
                buf.append(0, f)
        try:
            while True:
                t0 = time.perf_counter()
                f = out.read(None)
                t1 = time.perf_counter()
                t_read += t1 - t0
                if f is None:
                    break
                if len(f) == 0:
                    continue
                nrec += len(f)
                if combiner is not None:
                    combiner.combine(f)
                else:
                    emit(f)
                t_write += time.perf_counter() - t1
            if combiner is not None:
                t1 = time.perf_counter()
                r = combiner.reader(self.device)
                for g in r:
                    emit(g)
                r.close()
                t_write += time.perf_counter() - t1
        finally:
            out.close()
            if combiner is not None:
                combiner.discard()
        task.stats = {"inrecords": nrec, "readDuration": int(t_read * 1e6),
                      "writeDuration": int(t_write * 1e6), "write": buf.nrows}
        return buf

    # -- output access ---------------------------------------------------------
    def reader(self, task: Task, partition: int) -> Reader:
        buf = self.buffers.get(id(task))
        if buf is None:
            raise errors.TaskLostError("output of %s is not available" % task.name)
        return buf.reader(partition)

    def discard(self, task: Task) -> None:
        if task.combine_key:
            return  # shared-combiner tasks are not individually discardable
        buf = self.buffers.pop(id(task), None)
        if isinstance(buf, _StoredBuffer):
            buf.discard()
        if task.state() == TaskState.OK:
            task.set(TaskState.LOST)

    def stored_bytes(self) -> int:
        bufs = list(self.buffers.values())
        return sum(f.nbytes() for b in bufs for part in b.parts for f in part)

    Run, Reader, Discard = run, reader, discard
