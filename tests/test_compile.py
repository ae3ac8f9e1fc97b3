"""Compiler golden tests: the 7 DAG shapes of the reference
(exec/compile_test.go:23-134, goldens exec/testdata/*.graph) must produce the
same task names and edges; plus CompileEnv and pipelined-cache behaviour."""
import os

import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200.exec.compile import CompileEnv, ExecInvocation, compile_slice
from bigslice_b200.exec.task import make_graph
from bigslice_b200.func import Invocation

ADD = bs.ops.add
ID = lambda i: i  # noqa: E731


def ints(*cols):
    return [torch.zeros(0, dtype=torch.int64) if c == "i" else torch.zeros(0, dtype=torch.float64)
            for c in cols]


def compile_(build):
    inv = ExecInvocation(Invocation(1, 0, (), False, "<unknown>"))
    return compile_slice(inv, build())


def graph(nodes, edges):
    return "\n".join(sorted(nodes) + sorted(edges)) + "\n"


def names(op, n=3):
    return ["inv1_%s@%d:%d" % (op, n, i) for i in range(n)]


def all_to_all(a, b, na=3, nb=3):
    return ["%s -> %s" % (x, y) for x in names(a, na) for y in names(b, nb)]


def one_to_one(a, b, n=3):
    return ["%s -> %s" % (x, y) for x, y in zip(names(a, n), names(b, n))]


def test_trivial():
    tasks = compile_(lambda: bs.Const(3, *ints("i")))
    assert make_graph(tasks) == graph(names("const"), [])
    assert all(t.pragma is not None for t in tasks)


def test_shuffle():
    tasks = compile_(lambda: bs.Reduce(bs.Const(3, *ints("i", "f")), ADD))
    assert make_graph(tasks) == graph(names("const") + names("reduce"),
                                      all_to_all("reduce", "const"))
    prod = tasks[0].deps[0].head
    assert prod.combiner is not None and prod.num_partition == 3 and prod.group is not None


def test_branch():
    def build():
        s = bs.Map(bs.Const(3, *ints("i")), ID, out=[bs.int64])
        return bs.Cogroup(bs.Map(s, ID, out=[bs.int64]), bs.Map(s, ID, out=[bs.int64]))
    want = graph(names("cogroup") + names("const_map_map") + names("const_map_map1"),
                 all_to_all("cogroup", "const_map_map") + all_to_all("cogroup", "const_map_map1"))
    assert make_graph(compile_(build)) == want


def test_branch_materialize():
    def build():
        s = bs.Map(bs.Const(3, *ints("i")), ID, bs.ExperimentalMaterialize, out=[bs.int64])
        return bs.Cogroup(bs.Map(s, ID, out=[bs.int64]), bs.Map(s, ID, out=[bs.int64]))
    want = graph(names("cogroup") + names("const_map") + names("map") + names("map1"),
                 all_to_all("cogroup", "map") + all_to_all("cogroup", "map1")
                 + one_to_one("map", "const_map") + one_to_one("map1", "const_map"))
    assert make_graph(compile_(build)) == want


def test_branch_shuffle():
    def build():
        s = bs.Const(3, *ints("i", "f"))
        return bs.Cogroup(s, bs.Reduce(s, ADD))
    want = graph(names("cogroup") + names("const") + names("const1") + names("reduce"),
                 all_to_all("cogroup", "const") + all_to_all("cogroup", "reduce")
                 + all_to_all("reduce", "const1"))
    assert make_graph(compile_(build)) == want


def test_branch_same_partitions_reuses_tasks():
    def build():
        s = bs.Map(bs.Const(3, *ints("i")), ID, out=[bs.int64])
        return bs.Cogroup(bs.Reshard(s, 2), bs.Reshard(s, 2))
    tasks = compile_(build)
    g = make_graph(tasks)
    assert "const_map1" not in g
    want_nodes = names("cogroup", 2) + names("const_map") + names("reshard", 2) + names("reshard1", 2)
    want_edges = (all_to_all("cogroup", "reshard", 2, 2) + all_to_all("cogroup", "reshard1", 2, 2)
                  + all_to_all("reshard", "const_map", 2, 3) + all_to_all("reshard1", "const_map", 2, 3))
    assert g == graph(want_nodes, want_edges)


def test_branch_different_partitions_not_reused():
    def build():
        s = bs.Map(bs.Const(3, *ints("i")), ID, out=[bs.int64])
        return bs.Cogroup(bs.Reshard(s, 1), bs.Reshard(s, 2))
    g = make_graph(compile_(build))
    assert "inv1_const_map@3:0" in g and "inv1_const_map1@3:0" in g


def test_task_graph_rendering():
    tasks = compile_(lambda: bs.Reduce(bs.Const(2, *ints("i", "f")), ADD))
    s = tasks[0].graph_string()
    assert "tasks:" in s and "dependencies:" in s
    assert "inv1_reduce@2:0" in s and "inv1_const@2:1[0]" in s and "[INIT]" in s


def test_compile_env_writable_vs_frozen(tmp_path):
    prefix = str(tmp_path / "c")

    def build():
        return bs.Cache(bs.Map(bs.Const(2, torch.arange(4)), ID, out=[bs.int64]), prefix)
    # Nothing cached: writable env records nothing, tasks keep their pipeline.
    inv = ExecInvocation(Invocation(1, 0, (), False, "x"))
    tasks = compile_slice(inv, build())
    assert not inv.env.is_cached(tasks[0].name, 0)
    # Populate the cache, then a frozen env from the earlier compile must
    # still compile *uncached* (consistency across ranks).
    res = bs.slicetest.run(build())
    assert sorted(res.rows()) == [(0,), (1,), (2,), (3,)]
    assert all(os.path.exists("%s-%04d-of-0002" % (prefix, i)) for i in range(2))
    inv.env.freeze()
    with pytest.raises(RuntimeError):
        inv.env.mark_cached(tasks[0].name, 0)
    frozen = ExecInvocation(Invocation(1, 0, (), False, "x"), inv.env)
    t2 = compile_slice(frozen, build())
    assert not frozen.env.is_cached(t2[0].name, 0)
    # A fresh writable env now sees the cache and drops dependencies.
    fresh = ExecInvocation(Invocation(1, 0, (), False, "x"))
    t3 = compile_slice(fresh, bs.Reduce(bs.Prefixed(bs.Map(build(), lambda x: (x, x)), 1), ADD))
    prod = t3[0].deps[0].head
    assert fresh.env.is_cached(prod.name, 1) and prod.deps == []


def test_plan_ir_and_lowering():
    """The planner's IR: stages, edges and the lowering chosen for each exchange."""
    from bigslice_b200.exec.compile import Planner
    inv = ExecInvocation(Invocation(1, 0, (), False, "<unknown>"))

    def build():
        s = bs.Map(bs.Const(3, *ints("i", "i")), lambda a, b: (a, b), out=[bs.int64, bs.int64])
        red = bs.Reduce(s, ADD)                                   # int64 key, int64 value, builtin add
        return bs.Cogroup(red, bs.Reshard(bs.Const(3, *ints("i", "f")), 3))
    plan = Planner(inv).plan(build())
    text = plan.describe()
    assert "reduce-table" in text and "frames" in text
    by_name = {s.op_name: s for s in plan.stages}
    prod = by_name["inv1_const_map"]
    assert prod.exchange is not None and prod.exchange.lowering == "reduce-table"
    assert [e.wide for e in by_name["inv1_reduce"].inputs] == [True]
    assert len(plan.root.inputs) == 2 and all(e.wide for e in plan.root.inputs)
    tasks = compile_slice(ExecInvocation(Invocation(1, 0, (), False, "<unknown>")), build())
    lows = {t.name.op: t.lowering for t in tasks[0].all() if t.group is not None}
    assert lows["inv1_const_map"] == "reduce-table"
    assert lows["inv1_reduce"] == "frames" and lows["inv1_const"] == "frames"
