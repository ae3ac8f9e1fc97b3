"""Planner: slice DAG -> stage plan -> task DAG.

Compilation is two explicit passes over an intermediate representation, so that
the *shape* of a computation (what is fused, where rows cross shards, how each
crossing is carried out) exists as data before any task does:

1. :func:`plan_slice` walks the slice graph once (iteratively, pre-order) and
   emits a :class:`Plan`: :class:`Stage` nodes — maximal chains of slices joined
   by single narrow dependencies, fused into one task per shard so a tile flows
   through ``Map -> Filter -> ...`` while resident in HBM — connected by
   :class:`Edge` records, *narrow* (shard ``i`` reads shard ``i``) or *wide*
   (shard ``p`` reads partition ``p`` of every producer shard).  A wide edge
   carries its :class:`Exchange` description: partition count, partitioner,
   combiner, and the *lowering* the executors pick their code path from
   (``"reduce-table"``: device hash-aggregation + fused NVLink table exchange,
   ``csrc/reduce_exchange.cu``; ``"frames"``: generic partition/exchange of row
   frames).  Stages are shared between consumers that need the same
   partitioning; ``Result`` arguments become *imported* stages that re-use the
   tasks of an earlier invocation (with a ``*_shuffle`` re-partitioning stage in
   front when a consumer needs a different partitioning).
2. :func:`instantiate` turns every stage into its tasks, wires ``TaskDep``s from
   the edges and attaches one :class:`StageProgram` per stage — an interpreter of
   the fused op list that knows which shards start from a shard cache instead of
   their dependencies.

Behavioural parity with the reference compiler (``exec/compile.go:20-398``): task
names and graph shapes equal the golden files of ``exec/testdata/*.graph``
(``tests/test_compile.py``), memoisation by (slice, partition count), no re-use
across combiners / custom partitioners, cache-aware pipelines, ``CompileEnv``
for rank-consistent cache decisions (``exec/compile.go:133-184``).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

from .. import errors
from ..internal import slicecache
from ..sliceio import LabelReader, Reader
from ..slices import Pragmas, Slice, Unwrap, default_partitioner
from .task import Task, TaskDep, TaskName


class CompileEnv:
    """Cache decisions that must be identical on every rank.

    Rank 0 plans with a *writable* environment, which records for every (task,
    fused-op index) whether a shard cache was hit; the other ranks plan against
    the frozen copy, so that a cache file appearing in between cannot make
    their task graphs differ."""

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
    """An invocation plus its compile environment."""

    def __init__(self, inv, env: Optional[CompileEnv] = None):
        self.inv = inv
        self.index = inv.index
        self.location = inv.location
        self.exclusive = inv.exclusive
        self.env = env or CompileEnv.make()
        self.plan: Optional["Plan"] = None

    def invoke(self) -> Slice:
        """The invocation's slice graph (the Func body runs once per ExecInvocation)."""
        if getattr(self, "_slice", None) is None:
            self._slice = self.inv.invoke()
        return self._slice

    def uses_cache(self) -> bool:
        """Whether any slice of the graph consults a shard cache at plan time —
        the only thing a CompileEnv records, so only such invocations need their
        environment shipped to the other ranks."""
        todo, seen = [self.invoke()], set()
        while todo:
            sl = todo.pop()
            if id(sl) in seen:
                continue
            seen.add(id(sl))
            inner = Unwrap(sl)
            if isinstance(sl, slicecache.Cacheable) or isinstance(inner, slicecache.Cacheable):
                return True
            if _is_result(inner):
                continue
            todo.extend(sl.dep(i).slice for i in range(sl.num_dep()))
        return False


def _is_result(s: Slice) -> bool:
    from .session import Result
    return isinstance(s, Result)


# --------------------------------------------------------------------------
# IR

class Exchange:
    """How a stage's output crosses shards (absent: it does not)."""

    __slots__ = ("nparts", "partitioner", "combiner", "combine_key", "lowering")

    def __init__(self, nparts: int, partitioner=None, combiner=None, combine_key: str = ""):
        self.nparts = nparts
        self.partitioner = partitioner       # None: murmur3 of the key prefix
        self.combiner = combiner             # map-side combine before the exchange
        self.combine_key = combine_key
        self.lowering = "frames"

    def shareable(self) -> bool:
        """Stages may be shared by consumers only when their exchange is fully
        described by the partition count (functions cannot be compared)."""
        return self.combiner is None and self.partitioner is None


class Edge:
    __slots__ = ("producer", "wide", "expand", "combine_key")

    def __init__(self, producer: "Stage", wide: bool, expand: bool, combine_key: str = ""):
        self.producer, self.wide, self.expand, self.combine_key = producer, wide, expand, combine_key


class Stage:
    """A fused chain of slices, executed as one task per shard."""

    __slots__ = ("ops", "nshard", "exchange", "inputs", "op_name", "pragma", "imported", "tasks",
                 "reshuffle_of")

    def __init__(self, ops: List[Slice], nshard: int, exchange: Optional[Exchange]):
        self.ops = ops                      # ops[0] is the outermost slice, ops[-1] the source side
        self.nshard = nshard
        self.exchange = exchange
        self.inputs: List[Edge] = []
        self.op_name = ""
        self.pragma = None
        self.imported: Optional[List[Task]] = None   # tasks of an earlier invocation (Result)
        self.reshuffle_of: Optional[List[Task]] = None
        self.tasks: Optional[List[Task]] = None

    @property
    def head(self) -> Slice:
        return self.ops[0]

    @property
    def tail(self) -> Slice:
        return self.ops[-1]


class Plan:
    """Stages in creation (pre-)order; ``root`` produces the invocation's result."""

    def __init__(self, root: Stage, stages: List[Stage]):
        self.root, self.stages = root, stages

    def describe(self) -> str:
        """One line per stage: fused ops, shards, exchange and its lowering."""
        idx = {id(s): i for i, s in enumerate(self.stages)}
        lines = []
        for i, s in enumerate(self.stages):
            what = "import(%d tasks)" % len(s.imported) if s.imported is not None else s.op_name
            x = ""
            if s.exchange is not None:
                x = " => %d partitions [%s%s]" % (s.exchange.nparts, s.exchange.lowering,
                                                   ", combine" if s.exchange.combiner is not None else "")
            ins = ", ".join("%s s%d" % ("wide" if e.wide else "narrow", idx[id(e.producer)]) for e in s.inputs)
            lines.append("s%d %s x%d%s%s" % (i, what, s.nshard, x, (" <- " + ins) if ins else ""))
        return "\n".join(lines)


def fused_chain(slice_: Slice) -> List[Slice]:
    """The slices executed together with ``slice_``, outermost first: follow single
    narrow dependencies until a shuffle, a fan-in, a materialisation point or a
    ``Result`` of an earlier invocation."""
    chain: List[Slice] = []
    cur = slice_
    while not _is_result(Unwrap(cur)):
        chain.append(cur)
        if cur.num_dep() != 1:
            break
        d = cur.dep(0)
        if d.shuffle or d.slice.materialize():
            break
        cur = d.slice
    return chain


pipeline = fused_chain  # the reference's name for it


def _table_lowering(stage: Stage) -> str:
    """``reduce-table`` when the exchange can run as device hash aggregation plus
    the fused table exchange: default partitioner, a builtin combiner, a key prefix
    that packs into 64 bits (one scalar column, or two of at most 4 bytes each;
    ``ops/keypack.py``) and one scalar numeric value."""
    from ..ops import keypack
    x = stage.exchange
    if x is None or x.combiner is None or x.partitioner is not None:
        return "frames"
    if getattr(x.combiner, "builtin", None) not in ("add", "min", "max"):
        return "frames"
    typ = stage.head.typ
    if typ.num_out() != typ.prefix() + 1 or keypack.layout_for(typ) is None:
        return "frames"
    v = typ.out(typ.prefix())
    if not getattr(v, "is_tensor", False) or getattr(v, "kind", "") in ("vec", "list", "bool"):
        return "frames"
    return "reduce-table"


class _Names(dict):
    """Deterministic, collision-free task op names: ``x``, ``x1``, ``x2``…"""

    def fresh(self, base: str) -> str:
        n = self.get(base, 0)
        self[base] = n + 1
        return base if n == 0 else "%s%d" % (base, n)


class Planner:
    """Builds the :class:`Plan` of one invocation."""

    def __init__(self, inv: ExecInvocation, machine_combiners: bool = False):
        self.inv = inv
        self.machine_combiners = machine_combiners
        self.names = _Names()
        self.shared: Dict[Tuple[int, int], Stage] = {}
        self.stages: List[Stage] = []
        self._pins: List[Slice] = []  # keep ids of shared slices alive

    def plan(self, root: Slice) -> Plan:
        # Explicit work list instead of recursion: (slice, exchange, attach) where
        # `attach(stage)` links the finished stage into its consumer.  Stages are
        # named when first visited and inputs are visited in dependency order,
        # which fixes the naming order (…, x, x1 for repeated shapes).
        out: List[Stage] = []
        work: List[Tuple[Slice, Optional[Exchange], Callable[[Stage], None]]] = [(root, None, out.append)]
        while work:
            sl, xch, attach = work.pop()
            key = None
            if xch is None or xch.shareable():
                key = (id(sl), xch.nparts if xch is not None else 0)
                hit = self.shared.get(key)
                if hit is not None:
                    attach(hit)
                    continue
            stage, pending = self._open_stage(sl, xch)
            if key is not None:
                self.shared[key] = stage
                self._pins.append(sl)
            attach(stage)
            # depth-first, first dependency first: push in reverse
            work.extend(reversed(pending))
        return Plan(out[0], self.stages)

    def _open_stage(self, sl: Slice, xch: Optional[Exchange]):
        inner = Unwrap(sl)
        if _is_result(inner):
            return self._import_result(sl, inner, xch), []
        ops = fused_chain(sl)
        stage = Stage(ops, sl.num_shard(), xch)
        prag = Pragmas()
        for s in reversed(ops):
            prag.append(s)
        stage.pragma = prag
        stage.op_name = self.names.fresh("_".join(["inv%d" % self.inv.index] + [s.name().op for s in reversed(ops)]))
        if xch is not None:
            xch.lowering = _table_lowering(stage)
        self.stages.append(stage)
        pending = []
        tail = stage.tail
        for i in range(tail.num_dep()):
            dep = tail.dep(i)
            if not dep.shuffle:
                pending.append((dep.slice, None, self._linker(stage, i, False, dep.expand, "")))
                continue
            ckey = stage.op_name if (tail.combiner() is not None and self.machine_combiners) else ""
            dx = Exchange(sl.num_shard(), dep.partitioner, tail.combiner(), ckey)
            pending.append((dep.slice, dx, self._linker(stage, i, True, dep.expand, ckey)))
        stage.inputs = [None] * tail.num_dep()  # filled by the linkers, in dependency order
        return stage, pending

    @staticmethod
    def _linker(consumer: Stage, slot: int, wide: bool, expand: bool, ckey: str):
        def link(producer: Stage) -> None:
            if not wide and producer.nshard != consumer.nshard:
                raise errors.fatal("plan: a narrow edge joins %d shards to %d" % (producer.nshard, consumer.nshard))
            consumer.inputs[slot] = Edge(producer, wide, expand, ckey)
        return link

    def _import_result(self, sl: Slice, result, xch: Optional[Exchange]) -> Stage:
        """A ``Result`` argument: its tasks are taken as they are; a consumer that
        needs them partitioned gets a re-partitioning stage in front."""
        for t in result.tasks:
            if t.combiner is not None:
                raise errors.Error("a result whose tasks carry the shared combiner %r cannot be reused"
                                   % (t.combine_key,))
        if xch is None:
            st = Stage([sl], len(result.tasks), None)
            st.imported = list(result.tasks)
            self.stages.append(st)
            return st
        st = Stage([sl], len(result.tasks), xch)
        st.reshuffle_of = list(result.tasks)
        st.op_name = self.names.fresh("%s_shuffle" % result.tasks[0].name.op)
        self.stages.append(st)
        return st


# --------------------------------------------------------------------------
# programs: what a stage's task runs

class StageProgram:
    """Interpreter of a stage's fused op list.

    ``ops`` run source-side first.  A shard whose op ``i`` is served by a shard
    cache starts there — the ops below it and the task's dependencies are not
    needed — and every op writes through to its cache when it has one."""

    __slots__ = ("ops", "labels", "caches", "start_at")

    def __init__(self, ops: List[Slice], labels: List[str], caches: list):
        self.ops, self.labels, self.caches = ops, labels, caches
        self.start_at: Dict[int, int] = {}   # shard -> index of the outermost cached op

    def bind(self, shard: int) -> Callable[[List[Reader]], Reader]:
        return _BoundProgram(self, shard)


class _BoundProgram:
    __slots__ = ("prog", "shard")

    def __init__(self, prog: StageProgram, shard: int):
        self.prog, self.shard = prog, shard

    def __call__(self, readers: List[Reader]) -> Reader:
        p, shard = self.prog, self.shard
        n = len(p.ops)
        start = p.start_at.get(shard)
        if start is not None:
            cur = LabelReader(p.caches[start].cache_reader(shard), p.labels[start])
            first = start - 1
        else:
            cur, first = None, n - 1
        for i in range(first, -1, -1):      # ops[n-1] is the source side
            r = p.ops[i].reader(shard, readers if cur is None else [cur])
            r = p.caches[i].writethrough_reader(shard, r)
            cur = LabelReader(r, p.labels[i])
        return cur


def _passthrough(readers: List[Reader]) -> Reader:
    return readers[0]


# --------------------------------------------------------------------------
# instantiation

def instantiate(plan: Plan, inv: ExecInvocation) -> List[Task]:
    """Create the tasks of every stage of ``plan`` and wire their dependencies;
    returns the root stage's tasks."""
    env = inv.env
    for st in plan.stages:
        if st.tasks is not None:
            continue
        if st.imported is not None:
            st.tasks = st.imported
            continue
        x = st.exchange
        nparts = x.nparts if x is not None else 1
        part_fn = (x.partitioner if x is not None else None) or default_partitioner
        if st.reshuffle_of is not None:
            st.tasks = [Task(st.head, TaskName(inv.index, st.op_name, i, len(st.reshuffle_of)),
                             invocation=inv, do=_passthrough, deps=[TaskDep(t, 0, False, "")],
                             pragma=t.pragma, slices=t.slices, num_partition=nparts, partitioner=part_fn)
                        for i, t in enumerate(st.reshuffle_of)]
        else:
            st.tasks = [Task(st.head, TaskName(inv.index, st.op_name, i, st.nshard), invocation=inv,
                             pragma=st.pragma, num_partition=nparts, partitioner=part_fn,
                             combiner=x.combiner if x is not None else None,
                             combine_key=x.combine_key if x is not None else "", slices=st.ops)
                        for i in range(st.nshard)]
        if x is not None:
            for t in st.tasks:
                t.group = st.tasks
                t.lowering = x.lowering
    # edges: producers exist now, whatever the stage order
    for st in plan.stages:
        if st.imported is not None or st.reshuffle_of is not None:
            continue
        for e in st.inputs:
            prod = e.producer.tasks
            if e.wide:
                for p, t in enumerate(st.tasks):   # shard p reads partition p of every producer
                    t.deps.append(TaskDep(prod[0], p, e.expand, e.combine_key))
            else:
                for i, t in enumerate(st.tasks):
                    t.deps.append(TaskDep(prod[i], 0, e.expand, ""))
        _attach_program(st, inv, env)
    return plan.root.tasks


def _attach_program(st: Stage, inv: ExecInvocation, env: CompileEnv) -> None:
    n = len(st.ops)
    labels = ["%s(%s)" % (s.name(), inv.location) for s in st.ops]
    caches = []
    for s in st.ops:
        inner = Unwrap(s)
        caches.append(inner.cache() if isinstance(inner, slicecache.Cacheable) else slicecache.Empty)
    prog = StageProgram(st.ops, labels, caches)
    for shard, task in enumerate(st.tasks):
        if env.is_writable():
            for i in range(n):
                if caches[i].is_cached(shard):
                    env.mark_cached(task.name, i)
        hit = [i for i in range(n) if env.is_cached(task.name, i)]
        if hit:
            prog.start_at[shard] = min(hit)   # outermost cached op: nothing below it runs
            task.deps = []                    # the shard is read from its cache file
        task.do = prog.bind(shard)


def compile_slice(inv: ExecInvocation, slice_: Slice, machine_combiners: bool = False) -> List[Task]:
    """Plan ``slice_`` and instantiate its tasks; returns the root tasks."""
    plan = Planner(inv, machine_combiners).plan(slice_)
    inv.plan = plan
    return instantiate(plan, inv)


compile = compile_slice
