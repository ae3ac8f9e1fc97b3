"""Session / Result / Func / Cache tests (model: reference
exec/session_test.go:37-279, func_test.go:45-160, cache_test.go:28-351)."""
import os
import threading

import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import errors, slicetest, typecheck
from bigslice_b200.exec import TaskState
from bigslice_b200.exec.graph import DebugMux
from bigslice_b200.status import Status

THIS = os.path.abspath(__file__)


@bs.Func
def f_const(n: int):
    return bs.Const(n, torch.arange(n * 10))


@bs.Func
def f_inc(s: bs.Slice):
    return bs.Map(s, lambda x: x + 1)


@bs.Func
def f_reduce(s: bs.Slice):
    return bs.Reduce(bs.Map(s, lambda x: (x % 3, x)), bs.ops.add)


@bs.Func
def f_reshuffle(s: bs.Slice):
    return bs.Reshuffle(s)


map_calls = []
map_lock = threading.Lock()


@bs.Func
def f_counted(n: int):
    def fn(x):
        with map_lock:
            map_calls.append(len(x))
        return x
    return bs.Map(bs.Const(n, torch.arange(n * 4)), fn, out=[bs.int64])


@bs.Func
def f_panics():
    raise ValueError("func body panics")


def test_iterative_result_reuse(local_session):
    res = local_session.run(f_const, 4)
    for _ in range(5):
        res = local_session.run(f_inc, res)
    assert res.rows() == [(i + 5,) for i in range(40)]
    # a Result feeding a shuffle gets *_shuffle shim tasks
    red = local_session.run(f_reduce, res)
    want = {}
    for i in range(40):
        want[(i + 5) % 3] = want.get((i + 5) % 3, 0) + i + 5
    assert dict(red.rows()) == want
    sh = local_session.run(f_reshuffle, res)
    assert sorted(sh.rows()) == [(i + 5,) for i in range(40)]
    assert any("_shuffle" in str(d.head.name) for t in sh.tasks for d in t.deps)


def test_concurrent_reuse_runs_map_exactly_once_per_shard(local_session):
    del map_calls[:]
    base = local_session.run(f_counted, 6)
    assert len(map_calls) == 6
    results, errs = [], []

    def use():
        try:
            results.append(local_session.run(f_inc, base).rows())
        except BaseException as e:
            errs.append(e)
    ths = [threading.Thread(target=use) for _ in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs and all(r == [(i + 1,) for i in range(24)] for r in results)
    assert len(map_calls) == 6  # never recomputed


def test_func_panic_propagates_and_session_survives(local_session):
    with pytest.raises(ValueError):
        local_session.run(f_panics)
    assert local_session.run(f_const, 1).rows() == [(i,) for i in range(10)]


def test_discard_marks_tasks_lost_and_recomputes(local_session):
    del map_calls[:]
    res = local_session.run(f_counted, 3)
    assert len(map_calls) == 3
    res.discard()
    assert all(t.state() == TaskState.LOST for t in res.tasks)
    assert res.rows() == [(i,) for i in range(12)]  # scan re-evaluates lost tasks
    assert len(map_calls) == 6
    local_session.discard()
    assert all(t.state() == TaskState.LOST for t in res.tasks)
    assert local_session.run(f_inc, res).rows() == [(i + 1,) for i in range(12)]


def test_func_argument_typechecks(local_session):
    def at(msg, fn):
        with pytest.raises(typecheck.TypecheckError) as ei:
            fn()
        assert msg in ei.value.err and os.path.abspath(ei.value.file) == THIS
    at("wrong number of arguments: function takes 1 arguments, got 2",
       lambda: local_session.run(f_const, 1, 2))
    at("wrong type for argument 0: expected int, got str", lambda: local_session.run(f_const, "x"))
    at("cannot use nil as type int", lambda: local_session.run(f_const, None))
    at("does not implement interface Slice", lambda: local_session.run(f_inc, 3))
    at("is not a bigslice.Func", lambda: local_session.run(lambda: None))

    @bs.Func
    def nilable(xs: list, s: bs.Slice = None):
        return bs.Const(1, [1])
    assert local_session.run(nilable, None).rows() == [(1,)]
    at("func must return a single bigslice.Slice", lambda: local_session.run(bs.Func(lambda: 3)))
    with pytest.raises(typecheck.TypecheckError):
        bs.Func(3)


def test_type_errors_inside_func_point_at_run_call(local_session):
    @bs.Func
    def bad():
        return bs.Map(bs.Const(1, [1]), lambda a, b: a)
    with pytest.raises(typecheck.TypecheckError) as ei:
        local_session.run(bad)
    assert os.path.abspath(ei.value.file) == THIS and "map: function" in ei.value.err


def test_exclusive_and_registry_locations():
    ex = f_const.exclusive()
    assert ex.invocation("x", 1).exclusive and not f_const.invocation("x", 1).exclusive
    locs = bs.func.func_locations()
    assert all(":" in l for l in locs) and len(locs) >= 5
    assert bs.func.func_locations_diff(locs, locs) == []
    assert bs.func.func_locations_diff(locs, locs[:-1] + ["other:1"]) != []
    i1 = f_const.invocation("x", 1)
    i2 = f_const.invocation("x", 1)
    assert i2.index > i1.index


def test_session_options_status_trace_and_debug(tmp_path):
    st = Status()
    ev_ = bs.exec.ListEventer()
    path = str(tmp_path / "trace.json")
    sess = bs.start(bs.Local, bs.Parallelism(3), bs.exec.Status(st), bs.exec.Eventer(ev_),
                    bs.TracePath(path), bs.MaxLoad(0.5))
    assert sess.parallelism() == 3 and sess.max_load() == 0.5 and sess.status() is st
    sess.run(f_reduce, sess.run(f_const, 2))
    mux = DebugMux()
    sess.handle_debug(mux)
    ct, body = mux.get("/debug/tasks/graph")
    import json
    g = json.loads(body)
    assert g["nodes"] and all(n["state"] == "OK" for n in g["nodes"]) and g["links"]
    assert "traceEvents" in mux.get("/debug/trace")[1]
    assert "text/html" == mux.get("/debug/tasks")[0]
    srv = mux.serve()
    import urllib.request
    port = srv.server_address[1]
    assert b"/debug/trace" in urllib.request.urlopen("http://127.0.0.1:%d/debug" % port).read()
    srv.shutdown()
    sess.shutdown()
    tr = json.load(open(path))
    xs = [e for e in tr["traceEvents"] if e["ph"] == "X"]
    assert xs and all("inv" in e["name"] for e in xs)
    types = [t for t, _ in ev_.events]
    assert "bigslice:sessionStart" in types and "bigslice:taskComplete" in types
    with pytest.raises(ValueError):
        bs.Parallelism(0)


def test_procs_and_exclusive_limit_concurrency():
    running, peak = [0], [0]
    lock = threading.Lock()

    def slow(x):
        import time
        with lock:
            running[0] += 1
            peak[0] = max(peak[0], running[0])
        time.sleep(0.05)
        with lock:
            running[0] -= 1
        return x
    sess = bs.start(bs.Local, bs.Parallelism(4))
    s = bs.Map(bs.Const(8, torch.arange(8)), slow, bs.Procs(2), out=[bs.int64])
    slicetest.run(s, session=sess)
    assert peak[0] == 2  # 4 procs / 2 per task
    peak[0] = 0
    e = bs.Map(bs.Const(4, torch.arange(4)), slow, bs.Exclusive, out=[bs.int64])
    slicetest.run(e, session=sess)
    assert peak[0] == 1


# ------------------------------------------------------------------ cache

def _cache_files(prefix, n):
    return ["%s-%04d-of-%04d" % (prefix, i, n) for i in range(n)]


def test_cache_all_or_nothing_and_partial(tmp_path):
    computed = []
    lock = threading.Lock()

    def make(kind, prefix):
        def fn(x):
            with lock:
                computed.append(int(x[0]) if len(x) else -1)
            return x * 10
        s = bs.Map(bs.Const(4, torch.arange(8)), fn, out=[bs.int64])
        return (bs.Cache if kind == "all" else bs.CachePartial)(s, prefix)

    want = [(i * 10,) for i in range(8)]
    for kind in ("all", "partial"):
        prefix = str(tmp_path / kind)
        del computed[:]
        assert slicetest.run(make(kind, prefix)).rows() == want
        assert len(computed) == 4
        assert all(os.path.exists(p) for p in _cache_files(prefix, 4))
        del computed[:]
        assert slicetest.run(make(kind, prefix)).rows() == want
        assert computed == []  # fully served from cache
        os.unlink(_cache_files(prefix, 4)[1])
        assert slicetest.run(make(kind, prefix)).rows() == want
        assert len(computed) == (4 if kind == "all" else 1)
        assert all(os.path.exists(p) for p in _cache_files(prefix, 4))


def test_cache_error_leaves_no_valid_cache_and_readcache(tmp_path):
    prefix = str(tmp_path / "e")
    fail = [True]

    def fn(x):
        if fail[0]:
            raise ValueError("injected")
        return x
    build = lambda: bs.Cache(bs.Map(bs.Const(2, torch.arange(4)), fn, out=[bs.int64]), prefix)  # noqa: E731
    _, err = slicetest.run_err(build())
    assert err is not None
    assert not any(os.path.exists(p) for p in _cache_files(prefix, 2))
    fail[0] = False
    assert slicetest.run(build()).rows() == [(i,) for i in range(4)]
    rc = bs.ReadCache([bs.int64], 2, prefix)
    assert slicetest.run(rc).rows() == [(i,) for i in range(4)]
    _, err = slicetest.run_err(bs.ReadCache([bs.int64], 2, str(tmp_path / "missing")))
    assert err is not None


def test_cache_with_strings_and_shuffle_dep_dropped(tmp_path):
    prefix = str(tmp_path / "s")
    words = np.array(["a", "b", "a", "c"], dtype=object)

    def build():
        r = bs.Reduce(bs.Const(2, words, torch.ones(4, dtype=torch.int64)), bs.ops.add)
        return bs.Cache(r, prefix)
    assert dict(slicetest.run(build()).rows()) == {"a": 2, "b": 1, "c": 1}
    res = slicetest.run(build())
    assert dict(res.rows()) == {"a": 2, "b": 1, "c": 1}
    assert all(t.deps == [] for t in res.tasks)  # cached: no upstream tasks at all


# ---- invocation transport (exec/invocation.go, bigmachine_inv_disk_cache.go) ----
@bs.Func
def _inv_codec_func(n: int, label: str, prev=None):
    return bs.Const(2, torch.arange(n))


def test_invocation_codec_roundtrip_and_refs():
    from bigslice_b200.exec.compile import ExecInvocation
    from bigslice_b200.exec import invocation as I
    with bs.start(bs.Local) as sess:
        r = sess.run(_inv_codec_func, 4, "a")
        inv = _inv_codec_func.invocation("here.py:1", 8, "b", r)
        einv = ExecInvocation(inv)
        data = I.encode_invocation(einv)
        # without a resolver the Result travels as a reference
        d0 = I.decode_invocation(data)
        assert d0.inv.args[:2] == [8, "b"] and d0.inv.args[2] == I.InvocationRef(r.inv_index)
        assert (d0.index, d0.inv.func, d0.location) == (inv.index, inv.func, "here.py:1")
        assert d0.env.__class__ is einv.env.__class__
        d1 = I.decode_invocation(data, resolve=lambda i: r if i == r.inv_index else None)
        assert d1.inv.args[2] is r
        # fingerprints: equal for equal invocations, sensitive to args and to Result identity
        same = _inv_codec_func.invocation("elsewhere.py:9", 8, "b", r)
        same.index = inv.index
        assert I.fingerprint(same) == I.fingerprint(inv)
        other = _inv_codec_func.invocation("here.py:1", 9, "b", r)
        other.index = inv.index
        assert I.fingerprint(other) != I.fingerprint(inv)
        g = I.InvocationGraph()
        first = ExecInvocation(_inv_codec_func.invocation("x", 4, "a"))
        first.index = first.inv.index = r.inv_index
        g.add(first)
        g.add(einv)
        assert g.deps(einv.index) == {r.inv_index}
        assert g.closure(einv.index) == [r.inv_index, einv.index]


def test_inv_disk_cache(tmp_path):
    from bigslice_b200.exec.invocation import InvDiskCache
    c = InvDiskCache()
    calls = []

    def create(w):
        calls.append(1)
        w.write(b"x" * 100000)
        w.write(b"tail")
    for _ in range(3):
        rd = c.get_or_create(7, create)
        data = rd.read()
        rd.close()
        assert data == b"x" * 100000 + b"tail"
    assert len(calls) == 1 and 7 in c

    def bad(w):
        w.write(b"partial")
        raise ValueError("boom")
    with pytest.raises(bs.errors.Error):
        c.get_or_create(8, bad)
    assert 8 not in c
    rd = c.get_or_create(8, lambda w: w.write(b"ok"))
    assert rd.read(1) == b"o" and rd.read() == b"k"
    rd.close()
    d = c._dir
    c.close()
    import os
    assert not os.path.exists(d)
    with pytest.raises(RuntimeError):
        c.get_or_create(1, create)


@bs.Func
def _wide_shuffle(nshard: int, phases: int):
    s = bs.Const(nshard, torch.arange(nshard * 4), torch.ones(nshard * 4, dtype=torch.int64))
    for _ in range(phases):
        s = bs.Reduce(bs.Map(s, lambda k, v: (k + 1, v)), bs.ops.add)
    return s


def test_wide_shuffles_stay_cheap():
    """Hundreds of shards x several shuffle phases (the reference's
    TestMultiPhaseEval shape, with real data): per-(task, partition) overheads must
    stay small — this took minutes before combiners ran ahead of the partitioner,
    empty partitions stopped costing objects and tiny tasks stopped trading the GIL."""
    import time
    with bs.start(bs.Local, bs.Parallelism(4)) as sess:
        t0 = time.time()
        res = sess.run(_wide_shuffle, 200, 3)
        rows = dict(res.rows())
        dt = time.time() - t0
    assert rows == {k + 3: 1 for k in range(800)}
    assert dt < 30, "200 shards x 3 phases took %.1fs" % dt


def test_start_gate_never_blocks_dependent_work():
    """A task that waits for another *running* task must not be starved by the
    start gate (tiny tasks run serialised, but only for a few milliseconds)."""
    import time
    started = threading.Event()
    release = threading.Event()

    def first(x):
        started.set()
        assert release.wait(10), "the second task never got through the gate"
        return x

    def second(x):
        assert started.wait(10)
        release.set()
        return x

    @bs.Func
    def two():
        a = bs.Map(bs.Const(1, torch.arange(4)), first, out=[bs.int64])
        b = bs.Map(bs.Const(1, torch.arange(4)), second, out=[bs.int64])
        return bs.Cogroup(a, b)

    t0 = time.time()
    with bs.start(bs.Local, bs.Parallelism(2)) as sess:
        rows = sess.run(two).rows()
    assert len(rows) == 4 and time.time() - t0 < 10


def test_store_and_journal_session_options(tmp_path):
    """StoreOpt: task outputs live in files of a FileStore; Journal: every
    invocation is recorded (and an unencodable argument fails the run)."""
    import torch
    import bigslice_b200 as bs
    from bigslice_b200.exec.invocation import decode_invocation

    @bs.Func
    def job(n: int, tag):
        return bs.Reduce(bs.Const(2, torch.arange(n) % 5, torch.ones(n, dtype=torch.int64)), bs.ops.add)
    store_dir, jdir = str(tmp_path / "store"), str(tmp_path / "journal")
    with bs.start(bs.Local, bs.StoreOpt(store_dir), bs.Journal(jdir)) as sess:
        r = sess.run(job, 50, "x")
        assert dict(r.rows()) == {k: 10 for k in range(5)}
        files = [os.path.join(d, f) for d, _, fs in os.walk(store_dir) for f in fs]
        assert files, "no task output was written to the store"
        assert r.inv_index in sess.journal_
        with sess.journal_.get_or_create(r.inv_index, lambda w: None) as fh:
            inv = decode_invocation(fh.read())
        assert tuple(inv.inv.args) == (50, "x")
        with pytest.raises(bs.errors.Error):
            sess.run(job, 5, lambda: 0)   # a lambda cannot be encoded
