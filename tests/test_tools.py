"""Tests for the auxiliary components: metrics/stats, tracer + slicetrace,
status, sliceconfig, walker, tarslice, accumulators, topn, the static
analyzer, the slicer/urls/badfuncs programs and the example (model: reference
metrics/*_test.go, stats/stats_test.go, cmd/slicetrace/quartile_test.go,
internal/walker/walker_test.go, archive/tarslice/tarslice_test.go,
exec/topn_test.go, exec/slicestatus_test.go, example/max_test.go)."""
import io
import json
import os
import pickle
import random
import subprocess
import sys
import tarfile

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import bigslice_b200 as bs
from bigslice_b200 import accum, metrics, sliceconfig, slicetest, stats
from bigslice_b200.analysis import typecheck as analyzer
from bigslice_b200.archive import tarslice
from bigslice_b200.cmd import slicetrace
from bigslice_b200.example.max import int_max
from bigslice_b200.exec.topn import top_n
from bigslice_b200.exec.tracer import Tracer
from bigslice_b200.internal.walker import Walker
from bigslice_b200.status import Status

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_metrics_counters_merge_and_pickle():
    c1, c2 = metrics.new_counter("a"), metrics.new_counter("b")
    s1, s2 = metrics.Scope(), metrics.Scope()
    c1.incr(s1, 3)
    c1.incr(s2, 4)
    c2.incr(s2)
    s1.merge(s2)
    assert c1.value(s1) == 7 and c2.value(s1) == 1
    s3 = pickle.loads(pickle.dumps(s1))
    assert c1.value(s3) == 7
    s1.reset()
    assert c1.value(s1) == 0
    ctx = metrics.scoped_context(bs.Context(), s3)
    c2.incr(ctx, 5)
    assert c2.value(metrics.context_scope(ctx)) == 6


def test_stats_map():
    m = stats.Map()
    m.int("read").add(10)
    m.int("read").add(5)
    m.int("write").set(3)
    v = m.values()
    assert v == {"read": 15, "write": 3} and str(v) == "read:15 write:3"
    assert v.copy() is not v


def test_tracer_coalesces_and_slicetrace_summarises(tmp_path):
    tr = Tracer()

    class T:
        def __init__(self, n):
            self.name = n
    tr.event("m1", T("inv1_const_map@2:0"), "B", ts_us=0)
    tr.event("m1", T("inv1_const_map@2:1"), "B", ts_us=10)
    tr.event("m1", T("inv1_const_map@2:0"), "E", ts_us=1000, readDuration=100, writeDuration=50)
    tr.event("m1", T("inv1_const_map@2:1"), "E", ts_us=3010)
    tr.event("m2", T("inv1_reduce@2:0"), "B", ts_us=4000)
    tr.event("m2", T("inv1_reduce@2:0"), "E", ts_us=4500, error_type="lost")
    tr.event("m2", T("orphan@1:0"), "E", ts_us=1)  # dropped
    tr.complete("m2", "hash_combine", 4100, 200)
    data = json.loads(tr.marshal())
    xs = [e for e in data["traceEvents"] if e["ph"] == "X"]
    assert len(xs) == 4
    tids = {e["name"]: e["tid"] for e in xs if e["pid"] == 1}
    assert tids["inv1_const_map@2:0"] != tids["inv1_const_map@2:1"]  # concurrent -> own rows
    p = tmp_path / "t.json"
    p.write_text(json.dumps(data))
    summ = slicetrace.summarize(slicetrace.parse_trace(data))
    ops = {r["op"]: r for r in summ[1]}
    assert ops["const_map"]["tasks"] == 2 and abs(ops["const_map"]["read"] - 1e-4) < 1e-9
    assert ops["reduce"]["lost"] == 1 and ops["reduce"]["start"] == pytest.approx(0.004)
    out = io.StringIO()
    slicetrace.render(summ, out)
    assert "const_map" in out.getvalue() and slicetrace.main([str(p)]) == 0


def test_quartiles_fuzz():
    assert slicetrace.quartiles([1, 2, 3, 4, 5]) == (1, 2, 3, 4, 5)
    assert slicetrace.quartiles([]) == (0, 0, 0, 0, 0)
    rng = random.Random(0)
    for _ in range(200):
        xs = [rng.random() for _ in range(rng.randrange(1, 40))]
        mn, q1, q2, q3, mx = slicetrace.quartiles(xs)
        assert mn <= q1 <= q2 <= q3 <= mx and mn == min(xs) and mx == max(xs)


def test_status_and_slice_status():
    st_ = Status()
    sess = bs.start(bs.Local, bs.Parallelism(2), bs.exec.Status(st_))
    seen = []

    def spy(x):
        seen.append(st_.marshal())
        return x
    slicetest.run(bs.Reduce(bs.Map(bs.Const(2, torch.arange(8), torch.ones(8, dtype=torch.int64)),
                                   lambda k, v: (spy(k) % 2, v)), bs.ops.add), session=sess)
    text = "\n".join(seen)
    assert "tasks idle/running/done" in text and "slices" in text and "const@" in text
    assert st_.groups() == []  # groups removed when the run finishes


def test_sliceconfig_parse_profile_and_flags(tmp_path, monkeypatch):
    prof = tmp_path / "config"
    prof.write_text('# comment\nbigslice.parallelism = 3\nbigslice.max-load = 0.5\n')
    sess, rest = sliceconfig.parse(["-local", "prog", "--x"], profile=str(prof))
    assert rest == ["prog", "--x"] and sess.parallelism() == 3 and sess.max_load() == 0.5
    assert sess.executor.name() == "local"
    sess2, _ = sliceconfig.parse(["-local", "-set", "bigslice.parallelism=5", "-status",
                                  "-bigslice-internal-default-chunk-rows", "777"],
                                 profile=str(prof))
    from bigslice_b200.internal import defaultsize
    assert sess2.parallelism() == 5 and sess2.status() is not None and defaultsize.Chunk == 777
    defaultsize.set_chunk(1 << 16)
    sess3, _ = sliceconfig.parse(["-local", "-http", "127.0.0.1:0"], profile=str(tmp_path / "none"))
    import urllib.request
    port = sess3.debug_server.server_address[1]
    assert b"/debug/status" in urllib.request.urlopen("http://127.0.0.1:%d/debug" % port).read()
    sess3.debug_server.shutdown()
    with pytest.raises(ValueError):
        sliceconfig.options_from({"system": "mars"})


def test_walker(tmp_path):
    (tmp_path / "a" / "b").mkdir(parents=True)
    (tmp_path / "a" / "b" / "f1").write_text("x")
    (tmp_path / "a" / "f0").write_text("y")
    (tmp_path / "z").write_text("z")
    w = Walker(str(tmp_path))
    paths = [os.path.relpath(p, tmp_path) for p in w]
    assert paths == [".", "a", "a/b", "a/b/f1", "a/f0", "z"]
    assert w.err() is None
    w2 = Walker(str(tmp_path / "missing"))
    assert not w2.scan() and w2.err() is not None


def test_tarslice(tmp_path):
    path = tmp_path / "t.tar"
    with tarfile.open(path, "w") as tf:
        for i in range(10):
            data = ("file %d" % i).encode()
            ti = tarfile.TarInfo("f%d.txt" % i)
            ti.size = len(data)
            tf.addfile(ti, io.BytesIO(data))
    s = tarslice.Reader(3, lambda: open(path, "rb"))
    res = slicetest.run(s)
    got = sorted((e.name, e.body) for (e,) in res.rows())
    assert got == sorted(("f%d.txt" % i, ("file %d" % i).encode()) for i in range(10))
    shard0 = bs.sliceio.read_all(res.reader(0, []), res.typ)
    assert [e.name for e in shard0.col(0)] == ["f0.txt", "f3.txt", "f6.txt", "f9.txt"]


def test_accumulator():
    a = accum.make_accumulator(bs.string, bs.int64)
    f = bs.Frame.of(np.array(["x", "y", "x"], dtype=object), torch.tensor([1, 2, 3]))
    a.accumulate(f, lambda acc, v: acc * 10 + v)
    a.accumulate(f, lambda acc, v: acc * 10 + v)
    ks, vs = a.read()
    assert dict(zip(ks, vs)) == {"x": 1313, "y": 22} and a.read() is None
    with pytest.raises(TypeError):
        accum.make_accumulator(bs.obj, bs.int64)


def test_topn():
    assert top_n([5, 1, 9, 9, 3], 3) == [2, 3, 0]
    assert top_n(torch.tensor([5., 1., 9., 3.]), 2) == [2, 0]
    assert top_n([], 3) == [] and top_n([1], 0) == []


def test_static_analyzer():
    src = '''
import bigslice_b200 as bs

@bs.Func
def f(n: int, name: str):
    return bs.Const(n, [1])

g = bs.Func(lambda a, b=2: bs.Const(1, [a]))

def main(sess):
    sess.run(f, 1, "x")          # ok
    sess.run(f, 1)               # too few
    sess.must(f, "one", "x")     # wrong literal type
    sess.run(f, 1, lambda: 3)    # unserialisable
    sess.run(g, 1)               # ok (default)
    sess.run(g, 1, 2, 3)         # too many
    sess.run(other, 1, 2, 3)     # unknown: ignored
'''
    d = analyzer.check_source(src, "prog.py")
    assert len(d) == 4
    assert "prog.py:12: bigslice.Func f takes 2 arguments, but 1 given" in d[0]
    assert "expected int, got str" in d[1] and "cannot be a lambda" in d[2]
    assert "g takes 2 arguments, but 3 given" in d[3]


def run_module(mod, *args, nproc=1, timeout=300):
    if nproc == 1:
        cmd = [sys.executable, "-m", mod] + list(args)
    else:
        import socket
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()  # noqa: E702
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % nproc,
               "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", mod] + list(args)
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)


def test_slicer_and_urls_programs(tmp_path):
    r = run_module("bigslice_b200.cmd.slicer", "-local", "-parallelism", "4", "cogroup",
                   "--nshard", "4", "--nkey", "500")
    assert r.returncode == 0 and "0 bad groups, 0 missing" in r.stdout, r.stderr[-2000:]
    r = run_module("bigslice_b200.cmd.slicer", "-local", "reduce", "--nshard", "3", "--rows", "2000")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
    f = tmp_path / "urls.txt"
    f.write_text("\n".join(["http://a.com/x", "https://b.org/y,https://a.com/z", "junk", "a.com/q"]))
    r = run_module("bigslice_b200.cmd.urls", "-local", "--input", str(f), "--nshard", "2", "--top", "2")
    assert r.returncode == 0 and r.stdout.splitlines()[0] == "a.com\t3", r.stdout + r.stderr[-2000:]
    r = run_module("bigslice_b200.cmd.urls", "-local", "--synthetic", "5000", "--top", "3")
    assert r.returncode == 0 and r.stdout.splitlines()[0].startswith("site0.example.com\t")


def test_badfuncs_registry_determinism():
    r = run_module("bigslice_b200.cmd.badfuncs", "ok", nproc=2)
    assert r.returncode == 0 and r.stdout.count("ok, 16 rows") == 2, r.stderr[-2000:]
    r = run_module("bigslice_b200.cmd.badfuncs", "toolate", nproc=2)
    assert r.returncode == 0 and "late Func rejected" in r.stdout, r.stdout + r.stderr[-2000:]
    r = run_module("bigslice_b200.cmd.badfuncs", "random", nproc=2)
    assert r.returncode == 0 and "registry differs" in r.stdout, r.stdout + r.stderr[-2000:]


def test_bigslice_cli_run(tmp_path):
    prog = tmp_path / "prog.py"
    prog.write_text('''
import torch, bigslice_b200 as bs

@bs.Func
def f(n: int):
    return bs.Reduce(bs.Const(n, torch.arange(10) % 3, torch.ones(10, dtype=torch.int64)), bs.ops.add)

def bigslice_main(sess, args):
    print("RESULT", sorted(sess.run(f, 2).rows()), args)
    return 0
''')
    r = run_module("bigslice_b200.cmd.bigslice", "run", str(prog), "-local", "extra")
    assert r.returncode == 0 and "RESULT [(0, 4), (1, 3), (2, 3)] ['extra']" in r.stdout, r.stderr[-2000:]


# ---- example: property-based test of user slice code (example/max_test.go) ----

@settings(max_examples=25, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 20), st.integers(-1000, 1000)), max_size=60),
       st.integers(1, 4))
def test_example_int_max_property(rows, nshard):
    keys = torch.tensor([k for k, _ in rows], dtype=torch.int64)
    vals = torch.tensor([v for _, v in rows], dtype=torch.int64)
    want = {}
    for k, v in rows:
        want[k] = max(want.get(k, v), v)
    got = dict(slicetest.run(int_max(bs.Const(nshard, keys, vals))).rows())
    assert got == want
