"""Compiler: slice DAG -> task DAG.

Chains of slices joined by single non-shuffle dependencies are *pipelined*
into one task per shard (their readers are composed, so a tile flows through
Map→Filter→… while resident in HBM); a shuffle dependency cuts the pipeline:
producer tasks partition their output into ``num_partition`` buckets and
consumer shard ``p`` reads bucket ``p`` of every producer.

Compilation is memoised by ``(slice, num_partition)`` so branches re-use
tasks; ``Result`` arguments re-use the tasks of earlier invocations (with
``*_shuffle`` shim tasks when a re-partition is needed); cached shards drop
their dependencies and read the cache instead.

Capability parity: reference ``exec/compile.go:20-398`` (pipeline, memoKey,
partitioner, compile, taskNamer, CompileEnv).  Task names and graph shapes
match the reference's golden files (``exec/testdata/*.graph``).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

from .. import errors, slices as _slices
from ..internal import slicecache
from ..sliceio import LabelReader, Reader
from ..slices import Pragmas, Slice, Unwrap, default_partitioner
from .task import Task, TaskDep, TaskName


class CompileEnv:
    """Compilation decisions that must be identical on every rank.

    The driver compiles with a *writable* environment which records, per task
    and pipelined-op index, whether the cache was used; other ranks compile
    with the frozen copy so they make the same choices even if the cache
    state changed in between (reference ``exec/compile.go:133-184``).
    """

    def __init__(self, writable: bool = True, cached: Optional[Dict[Tuple[str, int], bool]] = None):
        self._writable = writable
        self._cached: Dict[Tuple[str, int], bool] = dict(cached or {})

    @staticmethod
    def make() -> "CompileEnv":
        return CompileEnv(True)

    def is_writable(self) -> bool:
        return self._writable

    def freeze(self) -> None:
        self._writable = False

    def mark_cached(self, name: TaskName, op_idx: int) -> None:
        if not self._writable:
            raise RuntimeError("cannot mark tasks cached in a frozen CompileEnv")
        self._cached[(str(name), op_idx)] = True

    def is_cached(self, name: TaskName, op_idx: int) -> bool:
        return self._cached.get((str(name), op_idx), False)

    def __getstate__(self):
        return {"cached": self._cached}

    def __setstate__(self, st):
        self._writable = False
        self._cached = st["cached"]

    IsWritable, Freeze, MarkCached, IsCached = is_writable, freeze, mark_cached, is_cached


class ExecInvocation:
    """An invocation plus its compile environment (``execInvocation``)."""

    def __init__(self, inv, env: Optional[CompileEnv] = None):
        self.inv = inv
        self.index = inv.index
        self.location = inv.location
        self.exclusive = inv.exclusive
        self.env = env or CompileEnv.make()

    def invoke(self) -> Slice:
        """The invocation's slice graph (the Func body runs once per ExecInvocation)."""
        if getattr(self, "_slice", None) is None:
            self._slice = self.inv.invoke()
        return self._slice

    def uses_cache(self) -> bool:
        """Whether any slice of the graph consults a shard cache at compile time —
        the only thing a CompileEnv records, so only such invocations need their
        environment shipped to the other ranks."""
        from .session import Result
        seen = set()
        stack = [self.invoke()]
        while stack:
            sl = stack.pop()
            if id(sl) in seen:
                continue
            seen.add(id(sl))
            inner = Unwrap(sl)
            if isinstance(sl, slicecache.Cacheable) or isinstance(inner, slicecache.Cacheable):
                return True
            if isinstance(inner, Result):
                continue
            for i in range(sl.num_dep()):
                stack.append(sl.dep(i).slice)
        return False


def pipeline(slice_: Slice) -> List[Slice]:
    """The slices pipelined with ``slice_``: slices[0] depends on slices[1]…"""
    from .session import Result
    out: List[Slice] = []
    while True:
        if isinstance(Unwrap(slice_), Result):
            return out
        out.append(slice_)
        if slice_.num_dep() != 1:
            return out
        dep = slice_.dep(0)
        if dep.shuffle:
            return out
        if dep.slice.materialize():
            return out
        slice_ = dep.slice


class _Partitioner:
    __slots__ = ("num_partition", "partitioner", "combiner", "combine_key")

    def __init__(self, num_partition=0, partitioner=None, combiner=None, combine_key=""):
        self.num_partition = num_partition
        self.partitioner = partitioner
        self.combiner = combiner
        self.combine_key = combine_key

    def is_shuffle(self) -> bool:
        return self.num_partition != 0

    def n(self) -> int:
        return self.num_partition if self.num_partition else 1


class TaskNamer(dict):
    def new(self, name: str) -> str:
        c = self.get(name, 0)
        self[name] = c + 1
        return name if c == 0 else "%s%d" % (name, c)


class Compiler:
    def __init__(self, inv: ExecInvocation, machine_combiners: bool = False):
        self.namer = TaskNamer()
        self.inv = inv
        self.machine_combiners = machine_combiners
        self.memo: Dict[Tuple[int, int], List[Task]] = {}
        self._keep: List[Slice] = []  # keep memo keys (ids) alive

    def compile(self, slice_: Slice, part: _Partitioner) -> List[Task]:
        from .session import Result
        memo_key = None
        # Combiner / custom-partitioner tasks are never reused: equivalent
        # functions cannot be identified.
        if part.combiner is None and part.partitioner is None:
            memo_key = (id(slice_), part.num_partition)
            if memo_key in self.memo:
                return self.memo[memo_key]
        tasks = self._compile(slice_, part)
        if part.is_shuffle():
            for t in tasks:
                t.group = tasks
        if memo_key is not None:
            self.memo[memo_key] = tasks
            self._keep.append(slice_)
        return tasks

    def _compile(self, slice_: Slice, part: _Partitioner) -> List[Task]:
        from .session import Result
        inner = Unwrap(slice_)
        if isinstance(inner, Result):
            for task in inner.tasks:
                if task.combiner is not None:
                    raise errors.Error("cannot reuse task %s with combine key %s"
                                       % (task, task.combine_key))
            if not part.is_shuffle():
                return inner.tasks
            # Shim tasks whose only purpose is to (re-)shuffle stored output.
            op = self.namer.new("%s_shuffle" % inner.tasks[0].name.op)
            tasks = []
            for shard, task in enumerate(inner.tasks):
                tasks.append(Task(
                    slice_, TaskName(self.inv.index, op, shard, len(inner.tasks)),
                    invocation=self.inv, do=lambda readers: readers[0],
                    deps=[TaskDep(task, 0, False, "")], pragma=task.pragma,
                    slices=task.slices, num_partition=part.n(),
                    partitioner=part.partitioner or default_partitioner))
            return tasks
        slices_ = pipeline(slice_)
        pragmas = Pragmas()
        ops = ["inv%d" % self.inv.index]
        for s in reversed(slices_):
            ops.append(s.name().op)
            pragmas.append(s)
        op_name = self.namer.new("_".join(ops))
        nshard = slice_.num_shard()
        tasks = [Task(slices_[0], TaskName(self.inv.index, op_name, i, nshard),
                      invocation=self.inv, pragma=pragmas, num_partition=part.n(),
                      partitioner=part.partitioner or default_partitioner,
                      combiner=part.combiner, combine_key=part.combine_key,
                      slices=slices_)
                 for i in range(nshard)]
        last = slices_[-1]
        for i in range(last.num_dep()):
            dep = last.dep(i)
            if not dep.shuffle:
                dep_tasks = self.compile(dep.slice, _Partitioner())
                if len(dep_tasks) != len(tasks):
                    raise errors.fatal("compile: tasks:%d deptasks:%d" % (len(tasks), len(dep_tasks)))
                for shard, t in enumerate(tasks):
                    t.deps.append(TaskDep(dep_tasks[shard], 0, dep.expand, ""))
                continue
            combine_key = ""
            if last.combiner() is not None and self.machine_combiners:
                combine_key = op_name
            dep_part = _Partitioner(slice_.num_shard(), dep.partitioner, last.combiner(), combine_key)
            dep_tasks = self.compile(dep.slice, dep_part)
            # Every shard reads its own partition from all producer tasks.
            for partition, t in enumerate(tasks):
                t.deps.append(TaskDep(dep_tasks[0], partition, dep.expand, combine_key))
        # Compose readers, deepest op first, wiring the cache where configured.
        env = self.inv.env
        for op_idx in range(len(slices_) - 1, -1, -1):
            s = slices_[op_idx]
            label = "%s(%s)" % (s.name(), self.inv.location)
            shard_cache = slicecache.Empty
            inner_s = Unwrap(s)
            if isinstance(inner_s, slicecache.Cacheable):
                shard_cache = inner_s.cache()
            if env.is_writable():
                for shard, task in enumerate(tasks):
                    if shard_cache.is_cached(shard):
                        env.mark_cached(task.name, op_idx)
            for shard, task in enumerate(tasks):
                task.do = _compose(s, shard, task, op_idx, shard_cache, env, label)
        return tasks


def _compose(s: Slice, shard: int, task: Task, op_idx: int, shard_cache, env: CompileEnv,
             label: str) -> Callable[[List[Reader]], Reader]:
    prev = task.do
    if env.is_cached(task.name, op_idx):
        # Forget dependencies for cached shards: we read from the cache file.
        task.deps = []

        def do_cached(readers: List[Reader]) -> Reader:
            return LabelReader(shard_cache.cache_reader(shard), label)
        return do_cached
    if prev is None:
        def do_first(readers: List[Reader]) -> Reader:
            r = s.reader(shard, readers)
            r = shard_cache.writethrough_reader(shard, r)
            return LabelReader(r, label)
        return do_first

    def do_next(readers: List[Reader]) -> Reader:
        r = s.reader(shard, [prev(readers)])
        r = shard_cache.writethrough_reader(shard, r)
        return LabelReader(r, label)
    return do_next


def compile_slice(inv: ExecInvocation, slice_: Slice, machine_combiners: bool = False) -> List[Task]:
    """Compile ``slice_`` into its root tasks (``exec.compile``)."""
    c = Compiler(inv, machine_combiners)
    return c.compile(slice_, _Partitioner())


compile = compile_slice
