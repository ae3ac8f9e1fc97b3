"""Combinator integration tests on the local executor (model: reference
slice_test.go:306-1036, reduce_test.go, cogroup_test.go, reshuffle_test.go,
reshard_test.go, scan_test.go, func_test.go)."""
import io
import logging
import os
import random
import threading

import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import errors, metrics, slicetest, typecheck
from bigslice_b200.slicetest import run, run_and_scan, sorted_rows

THIS = os.path.abspath(__file__)


def expect_type_error(msg, fn):
    """The error message and the caller's file must match (reference
    slice_test.go:257-286)."""
    with pytest.raises(typecheck.TypecheckError) as ei:
        fn()
    assert msg in ei.value.err, ei.value.err
    assert os.path.abspath(ei.value.file) == THIS
    return ei.value


# ---------------------------------------------------------------- Const

@pytest.mark.parametrize("n,nshard", [(0, 1), (1, 3), (10, 3), (100, 7), (5, 5)])
def test_const_shards_even_and_ordered(n, nshard):
    s = bs.Const(nshard, torch.arange(n), np.array(["s%d" % i for i in range(n)], dtype=object))
    sizes = [bs.slices.const_shard(n, nshard, i) for i in range(nshard)]
    lens = [e - b for b, e in sizes]
    assert sum(lens) == n and max(lens) - min(lens) <= 1
    res = run(s)
    assert res.rows() == [(i, "s%d" % i) for i in range(n)]


def test_const_errors():
    expect_type_error("const: must have at least one column", lambda: bs.Const(1))
    expect_type_error("const: shard must be >= 1", lambda: bs.Const(0, [1]))
    expect_type_error("const: invalid slice inputs", lambda: bs.Const(1, [1, 2], [1]))


# ---------------------------------------------------------------- ReaderFunc

def test_reader_func_state_zeroing_and_eof():
    calls = []

    def read(shard, state, ints: bs.int64, strs: bs.string):
        assert ints.tolist() == [0] * len(ints) and list(strs) == [""] * len(strs)
        if not hasattr(state, "n"):
            state.n = 0
        calls.append(shard)
        k = min(len(ints), 10 - state.n)
        for i in range(k):
            ints[i] = shard * 100 + state.n + i
            strs[i] = "x"
        state.n += k
        return (k, bs.EOF) if state.n >= 10 else k

    s = bs.ReaderFunc(3, read, batch=4)
    rows = run(s).rows()
    assert sorted(r[0] for r in rows) == sorted(s_ * 100 + i for s_ in range(3) for i in range(10))
    assert all(r[1] == "x" for r in rows)


def test_reader_func_raise_eof_and_generator():
    def read(shard, state, v):
        raise EOFError
    assert run(bs.ReaderFunc(2, read, out=[bs.int64])).rows() == []

    def gen(shard):
        for i in range(3):
            yield (torch.tensor([shard * 10 + i]),)
    assert sorted(run(bs.ReaderGen(2, gen, [bs.int64])).rows()) == \
        [(0,), (1,), (2,), (10,), (11,), (12,)]


def test_reader_func_errors(caplog):
    expect_type_error("readerfunc: invalid reader function type", lambda: bs.ReaderFunc(1, 123))
    expect_type_error("is not vectorized", lambda: bs.ReaderFunc(1, lambda shard, state: 0))

    def bad(shard, state, v: bs.int64):
        return 0, ValueError("disk on fire")
    res, err = slicetest.run_err(bs.ReaderFunc(1, bad))
    assert err is not None and errors.is_fatal(err) and "disk on fire" in str(err)


def test_reader_func_forgotten_eof_warns(caplog):
    def read(shard, state, v: bs.int64):
        state.n = getattr(state, "n", 0) + 1
        if state.n > 20:
            return 0, bs.EOF
        return 0
    with caplog.at_level(logging.WARNING, logger="bigslice"):
        run(bs.ReaderFunc(1, read))
    assert any("consecutive empty reads" in r.message for r in caplog.records)


# ---------------------------------------------------------------- WriterFunc

def test_writer_func_sees_all_rows_and_eof():
    lock = threading.Lock()
    seen, eofs = [], []

    def write(shard, state, err, ints, strs):
        with lock:
            seen.extend(zip(ints.tolist(), list(strs)))
            if err is bs.EOF:
                eofs.append(shard)
        return None

    s = bs.Const(3, torch.arange(9), np.array(list("abcdefghi"), dtype=object))
    w = bs.WriterFunc(s, write)
    assert run(w).rows() == [(i, "abcdefghi"[i]) for i in range(9)]
    assert sorted(seen) == [(i, "abcdefghi"[i]) for i in range(9)] and sorted(eofs) == [0, 1, 2]


def test_writer_func_error_is_fatal_and_typecheck():
    s = bs.Const(1, torch.arange(3))
    res, err = slicetest.run_err(bs.WriterFunc(s, lambda shard, state, err, v: ValueError("nope")))
    assert err is not None and errors.is_fatal(err)
    expect_type_error("writerfunc: invalid writer function type",
                      lambda: bs.WriterFunc(s, lambda shard, state: None))


# ---------------------------------------------------------------- Map / Filter / Flatmap

def test_map_vectorized_rowwise_and_types():
    s = bs.Const(2, torch.arange(6))
    m = bs.Map(s, lambda x: (x * 2, x.to(torch.float32) / 2))
    assert [t.name for t in m.typ] == ["int64", "float32"]
    assert run(m).rows() == [(i * 2, i / 2) for i in range(6)]
    r = bs.Map(s, bs.rowwise(lambda x: ("n%d" % x, x % 2 == 0)))
    assert [t.name for t in r.typ] == ["string", "bool"]
    assert run(r).rows() == [("n%d" % i, i % 2 == 0) for i in range(6)]

    def annotated(x: bs.int64) -> tuple[bs.int32]:
        return (x.to(torch.int32),)
    assert bs.Map(s, annotated).typ.out(0) is bs.int32
    assert bs.Map(s, lambda x: x + 1, out=[bs.float64]).typ.out(0) is bs.float64
    assert bs.Map(bs.Prefixed(bs.Map(s, lambda x: (x, x)), 2), lambda a, b: a).prefix() == 1


def test_map_errors():
    s = bs.Const(1, torch.arange(3))
    expect_type_error("map: invalid map function", lambda: bs.Map(s, 3))
    expect_type_error("does not match input slice type slice[1]<int64>",
                      lambda: bs.Map(s, lambda a, b: a))

    def wrong(x: bs.string) -> bs.string:
        return x
    expect_type_error("does not match input slice type", lambda: bs.Map(s, wrong))
    expect_type_error("map: need at least one output column", lambda: bs.Map(s, lambda x: None, out=[]))


def test_filter():
    s = bs.Const(3, torch.arange(20))
    assert run(bs.Filter(s, lambda x: x % 2 == 0)).rows() == [(i,) for i in range(0, 20, 2)]
    assert run(bs.Filter(s, lambda x: x < 0)).rows() == []
    assert run(bs.Filter(s, bs.rowwise(lambda x: x > 17))).rows() == [(18,), (19,)]
    expect_type_error("filter: predicate must return a single boolean value",
                      lambda: bs.Filter(s, lambda x: x))
    expect_type_error("filter: invalid predicate function", lambda: bs.Filter(s, None))
    expect_type_error("does not match input slice type", lambda: bs.Filter(s, lambda a, b: a == b))


def test_flatmap_rowwise_vectorized_and_overflow():
    s = bs.Const(2, np.array(["a b", "c", ""], dtype=object))

    @bs.rowwise
    def split(line):
        w = line.split()
        return w, [len(x) for x in w]
    f = bs.Flatmap(s, split, out=[bs.string, bs.int64])
    assert run(f).rows() == [("a", 1), ("b", 1), ("c", 1)]
    v = bs.Flatmap(bs.Const(2, torch.arange(4)),
                   lambda x: (torch.repeat_interleave(x, x),))
    assert run(v).rows() == [(1,), (2,), (2,), (3,), (3,), (3,)]
    # one row expanding beyond the read batch must carry the remainder
    big = bs.Flatmap(bs.Const(1, torch.tensor([1025])),
                     bs.rowwise(lambda n: (list(range(n)),)), out=[bs.int64])
    h = bs.Head(big, 2000)
    assert len(run(h).rows()) == 1025
    expect_type_error("flatmap: invalid flatmap function", lambda: bs.Flatmap(s, "x"))
    expect_type_error("does not match input slice type", lambda: bs.Flatmap(s, lambda a, b: (a,)))


# ---------------------------------------------------------------- Fold / Reduce

def test_fold_ordered_noncommutative():
    keys = np.array(["a", "b", "a", "a", "b"], dtype=object)
    vals = torch.tensor([1, 2, 3, 4, 5])
    s = bs.Const(2, keys, vals)
    f = bs.Fold(s, bs.rowwise(lambda acc, v: acc * 10 + v), acc=bs.int64)
    assert sorted(run(f).rows()) == [("a", 134), ("b", 25)]
    g = bs.Fold(bs.Const(3, torch.tensor([1, 2, 1, 2, 1]), torch.tensor([1., 2., 3., 4., 5.])),
                lambda acc, v: acc + v)
    assert sorted(run(g).rows()) == [(1, 9.0), (2, 6.0)]
    expect_type_error("Fold can be applied only for slices with at least two columns",
                      lambda: bs.Fold(bs.Const(1, [1]), lambda a: a))
    expect_type_error("fold: expected func(acc, t2, t3, ..., tn)",
                      lambda: bs.Fold(s, lambda a: a))


@pytest.mark.parametrize("fn", [bs.ops.add, lambda a, b: a + b, bs.rowwise(lambda a, b: a + b)])
@pytest.mark.parametrize("nshard", [1, 4])
def test_reduce_sum(fn, nshard):
    rng = random.Random(4)
    ks = [rng.randrange(50) for _ in range(2000)]
    s = bs.Const(nshard, torch.tensor(ks), torch.ones(2000, dtype=torch.int64))
    want = {}
    for k in ks:
        want[k] = want.get(k, 0) + 1
    assert dict(run(bs.Reduce(s, fn)).rows()) == want


def test_reduce_string_keys_prefix2_and_minmax():
    words = np.array(["x", "y", "x", "z", "y", "x"], dtype=object)
    s = bs.Const(3, words, torch.ones(6, dtype=torch.int64))
    assert dict(run(bs.Reduce(s, bs.ops.add)).rows()) == {"x": 3, "y": 2, "z": 1}
    p = bs.Prefixed(bs.Const(2, torch.tensor([1, 1, 2, 1]), words[:4], torch.tensor([5, 6, 7, 8])), 2)
    got = {(a, b): c for a, b, c in run(bs.Reduce(p, bs.ops.max)).rows()}
    assert got == {(1, "x"): 5, (1, "y"): 6, (2, "x"): 7, (1, "z"): 8}


def test_reduce_errors():
    s3 = bs.Const(1, [1], [2], [3])
    expect_type_error("the slice must only have one 1 residual column; has 2",
                      lambda: bs.Reduce(s3, bs.ops.add))
    s = bs.Const(1, [1], [2])
    expect_type_error("reduce: invalid reduce function", lambda: bs.Reduce(s, 1))
    expect_type_error("expected func(int64, int64) int64", lambda: bs.Reduce(s, lambda a: a))
    so = bs.Const(1, np.array([{"a": 1}], dtype=object), [1])
    expect_type_error("cannot be hashed", lambda: bs.Reduce(so, bs.ops.add))


# ---------------------------------------------------------------- Head / Scan / Prefixed

def test_head_is_per_shard():
    s = bs.Const(2, torch.arange(12))
    assert run(bs.Head(s, 2)).rows() == [(0,), (1,), (6,), (7,)]


def test_scan_runs_once_per_shard():
    lock = threading.Lock()
    got, calls = [], []

    def scan(shard, scanner):
        with lock:
            calls.append(shard)
        for (x,) in scanner:
            with lock:
                got.append(x)
    s = bs.Scan(bs.Const(4, torch.arange(10)), scan)
    assert s.num_out() == 0
    res = run(s)
    assert res.rows() == [] and sorted(got) == list(range(10)) and sorted(calls) == [0, 1, 2, 3]


def test_prefixed():
    s = bs.Const(1, [1], [2], [3])
    assert bs.Prefixed(s, 2).prefix() == 2 and bs.Unwrap(bs.Prefixed(s, 2)) is s
    expect_type_error("prefixed: prefix must include at least one column", lambda: bs.Prefixed(s, 0))
    expect_type_error("prefixed: prefix 4 is greater than number of columns 3",
                      lambda: bs.Prefixed(s, 4))


# ---------------------------------------------------------------- shuffles

def test_reshard_golden_placement():
    # ExampleReshard in the reference: keys 1..6 over 4 shards -> {0:2,1:1,2:1,3:2}
    s = bs.Reshard(bs.Const(2, torch.arange(1, 7)), 4)
    assert s.num_shard() == 4
    res = run(s)
    per_shard = [len(bs.sliceio.read_all(res.reader(i, []), res.typ)) for i in range(4)]
    assert per_shard == [2, 1, 1, 2]
    assert sorted(res.rows()) == [(i,) for i in range(1, 7)]
    same = bs.Const(3, torch.arange(3))
    assert bs.Reshard(same, 3) is same


def test_reshuffle_groups_keys_and_repartition():
    ks = torch.tensor([0, 1, 2, 3, 0, 1, 2, 3] * 5)
    res = run(bs.Reshuffle(bs.Const(4, ks, torch.arange(40))))
    home = {}
    for shard in range(4):
        f = bs.sliceio.read_all(res.reader(shard, []), res.typ)
        for k in set(f.col(0).tolist()):
            assert home.setdefault(k, shard) == shard
    assert len(res.rows()) == 40
    rep = bs.Repartition(bs.Const(4, ks, torch.arange(40)), lambda n, k, v: (k + 1) % n)
    res = run(rep)
    for shard in range(4):
        f = bs.sliceio.read_all(res.reader(shard, []), res.typ)
        assert all((k + 1) % 4 == shard for k in f.col(0).tolist())
    rowrep = bs.Repartition(bs.Const(2, ks, torch.arange(40)), bs.rowwise(lambda n, k, v: 0))
    res = run(rowrep)
    assert len(bs.sliceio.read_all(res.reader(0, []), res.typ)) == 40
    bad = bs.Repartition(bs.Const(2, ks, torch.arange(40)), lambda n, k, v: k + 100)
    _, err = slicetest.run_err(bad)
    assert err is not None and errors.is_fatal(err)
    expect_type_error("repartition: not a function", lambda: bs.Repartition(rep, 5))


# ---------------------------------------------------------------- Cogroup

def test_cogroup_two_inputs():
    a = bs.Const(2, torch.tensor([1, 2, 1, 3]), np.array(["a", "b", "c", "d"], dtype=object))
    b = bs.Const(3, torch.tensor([2, 4, 2]), torch.tensor([20., 40., 21.]))
    c = bs.Cogroup(a, b)
    assert c.num_shard() == 3 and c.prefix() == 1
    assert [t.name for t in c.typ] == ["int64", "[]string", "[]float32"]
    rows = sorted_rows(run(c))
    assert rows == [(1, ("a", "c"), ()), (2, ("b",), (20.0, 21.0)), (3, ("d",), ()), (4, (), (40.0,))]


def test_cogroup_single_and_prefix2_sorted_within_shard():
    s = bs.Const(2, torch.tensor([5, 3, 5, 1]), torch.tensor([1, 2, 3, 4]))
    res = run(bs.Cogroup(s))
    assert sorted_rows(res) == [(1, (4,)), (3, (2,)), (5, (1, 3))]
    for shard in range(res.num_shard()):
        ks = bs.sliceio.read_all(res.reader(shard, []), res.typ).col(0).tolist()
        assert ks == sorted(ks)
    p = bs.Prefixed(bs.Const(2, torch.tensor([1, 1, 1]), np.array(["x", "y", "x"], dtype=object),
                             torch.tensor([1, 2, 3])), 2)
    assert sorted_rows(run(bs.Cogroup(p, p))) == [(1, "x", (1, 3), (1, 3)), (1, "y", (2,), (2,))]


def test_cogroup_errors():
    a = bs.Const(1, [1], [2])
    expect_type_error("cogroup: expected at least one slice", lambda: bs.Cogroup())
    expect_type_error("cogroup: prefix mismatch: expected 1 but got 2",
                      lambda: bs.Cogroup(a, bs.Prefixed(a, 2)))
    expect_type_error("cogroup: key column type mismatch: expected int64 but got string",
                      lambda: bs.Cogroup(a, bs.Const(1, ["x"])))


# ---------------------------------------------------------------- ScanReader / misc

def test_scan_reader_stripes_lines():
    text = "".join("line%d\n" % i for i in range(10))
    s = bs.ScanReader(3, lambda: io.StringIO(text))
    res = run(s)
    assert sorted(res.rows()) == sorted(("line%d" % i,) for i in range(10))
    f0 = bs.sliceio.read_all(res.reader(0, []), res.typ)
    assert list(f0.col(0)) == ["line0", "line3", "line6", "line9"]


def test_panic_in_user_function_is_fatal():
    def boom(x):
        raise ValueError("kaboom")
    res, err = slicetest.run_err(bs.Map(bs.Const(1, torch.arange(3)), boom, out=[bs.int64]))
    assert err is not None and errors.is_fatal(err)
    assert "panic while evaluating slice" in str(err) and "kaboom" in str(err)


counter = metrics.new_counter("rows")


def test_metrics_visible_in_result_scope():
    def f(ctx, x):
        counter.incr(ctx, len(x))
        return x
    res = run(bs.Map(bs.Const(3, torch.arange(10)), f, out=[bs.int64]))
    assert counter.value(res.scope()) == 10


def test_names_and_helper():
    s = bs.Const(1, [1])
    assert str(s.name()).startswith("const@test_slices.py:")

    @bs.helper
    def make():
        return bs.Const(1, [1])
    a = make(); line = a.name().line  # noqa: E702
    assert os.path.abspath(a.name().file) == THIS
    import inspect
    assert line == inspect.currentframe().f_lineno - 3
    names = [bs.Const(1, [1]).name() for _ in range(2)]
    assert [n.index for n in names] == [0, 1]
    assert "const@" in bs.String(s) and "<int64>" in bs.String(s)


def test_pragmas_compose():
    s = bs.Const(1, [1])
    m = bs.Map(s, lambda x: x, bs.Procs(2), out=[bs.int64])
    m2 = bs.Map(m, lambda x: x, bs.Procs(1), bs.Exclusive, out=[bs.int64])
    assert m.procs() == 2 and not m.exclusive() and m2.exclusive()
    assert bs.Pragmas([m, m2]).procs() == 2 and bs.Pragmas([m, m2]).exclusive()
    assert bs.Map(s, lambda x: x, bs.ExperimentalMaterialize, out=[bs.int64]).materialize()


# ---- the custom-Slice protocol (exec/compile_test.go:432-458) ----
class _Doubling(bs.Slice):
    """User-defined slice: wraps another slice (one pipelined dependency) and
    overrides ``reader``."""

    calls = 0

    def __init__(self, inner: bs.Slice):
        super().__init__(bs.make_name("doubling"), inner.typ, inner.num_shard(),
                         deps=[bs.Dep(inner)])

    def reader(self, shard, deps):
        from bigslice_b200.frame import Frame
        from bigslice_b200.sliceio import FuncReader
        src = deps[0]
        typ = self.typ

        def read(n):
            f = src.read(n)
            if f is None:
                return None
            _Doubling.calls += 1
            return Frame([f.col(0), f.col(1) * 2], typ)
        return FuncReader(read)


class _MemoryShardCache:
    """A user-provided shard cache (any object with the ShardCache methods)."""

    def __init__(self):
        self.frames = {}

    def is_cached(self, shard):
        return shard in self.frames

    def writethrough_reader(self, shard, reader):
        from bigslice_b200.sliceio import FuncReader
        got = []

        def read(n):
            f = reader.read(n)
            if f is None:
                self.frames[shard] = got
                return None
            got.append(f)
            return f
        return FuncReader(read)

    def cache_reader(self, shard):
        from bigslice_b200.sliceio import FrameReader, MultiReader
        return MultiReader([FrameReader(f) for f in self.frames[shard]])


from bigslice_b200.internal import slicecache as _slicecache  # noqa: E402


class _CachedDoubling(_Doubling, _slicecache.Cacheable):
    """Cacheable is discovered by isinstance on the (unwrapped) user slice."""

    store = _MemoryShardCache()

    def cache(self):
        return type(self).store


@bs.Func
def _custom_pipeline(n: int, cached: bool):
    base = bs.Const(3, torch.arange(n), torch.arange(n) * 10)
    mid = (_CachedDoubling if cached else _Doubling)(base)
    return bs.Reduce(bs.Map(mid, lambda k, v: (k % 4, v)), bs.ops.add)


def test_user_defined_slice_and_cacheable_protocol():
    n = 100
    want = {}
    for i in range(n):
        want[i % 4] = want.get(i % 4, 0) + i * 20
    with bs.start(bs.Local, bs.Parallelism(2)) as sess:
        _Doubling.calls = 0
        assert dict(sess.run(_custom_pipeline, n, False).rows()) == want
        assert _Doubling.calls == 3  # pipelined with Const and Map: one tile per shard
        # the same user slice, now also Cacheable
        _CachedDoubling.store.frames.clear()
        _Doubling.calls = 0
        assert dict(sess.run(_custom_pipeline, n, True).rows()) == want
        assert sorted(_CachedDoubling.store.frames) == [0, 1, 2] and _Doubling.calls == 3
        res = sess.run(_custom_pipeline, n, True)  # served from the user's cache
        assert dict(res.rows()) == want and _Doubling.calls == 3


def test_fold_expr_non_associative(local_session):
    """FoldExpr: ordered, non-associative fold (polynomial hash of each key's rows
    in row order) vs a plain python loop."""
    import torch
    n = 5000
    keys = torch.arange(n) % 17
    vals = (torch.arange(n) * 7919) % 1000

    @bs.Func
    def job():
        return bs.FoldExpr(bs.Const(4, keys, vals), "(acc * 31 + c1) % 1000003", bs.int64)
    got = dict(local_session.run(job).rows())
    want = {}
    for k, v in zip(keys.tolist(), vals.tolist()):
        want[k] = (want.get(k, 0) * 31 + v) % 1000003
    # rows of a key are folded in shard order: Const splits rows contiguously and
    # the shuffle keeps producer order, so the global row order is preserved
    assert got == want


def test_fold_python_rounds_are_capped(local_session, monkeypatch):
    import torch
    from bigslice_b200 import slices
    monkeypatch.setattr(slices, "MAX_FOLD_ROUNDS", 8)
    keys = torch.zeros(100, dtype=torch.int64)
    vals = torch.ones(100, dtype=torch.int64)

    @bs.Func
    def job():
        return bs.Fold(bs.Const(1, keys, vals), lambda acc, v: acc * 2 + v, acc=bs.int64)
    with pytest.raises(bs.errors.Error) as ei:
        local_session.run(job)
    assert "FoldExpr" in str(ei.value)


def test_broadcast_join_left_semantics(local_session):
    """BroadcastJoin: small side replicated, large side stays where it is; one row
    per key of every large shard, small-only keys dropped."""
    import torch
    lk = torch.tensor([5, 1, 5, 9, 2, 2, 7])
    lv = torch.arange(7, dtype=torch.float32)
    sk = torch.tensor([2, 5, 100])
    sv = torch.tensor([20, 50, 1000])

    @bs.Func
    def job():
        large = bs.Reshuffle(bs.Const(3, lk, lv))        # partitioned by key
        return bs.BroadcastJoin(large, bs.Const(2, sk, sv))
    rows = {k: (sorted(a), sorted(b)) for k, a, b in local_session.run(job).rows()}
    assert rows == {1: ([1.0], []), 2: ([4.0, 5.0], [20]), 5: ([0.0, 2.0], [50]), 7: ([6.0], []),
                    9: ([3.0], [])}
