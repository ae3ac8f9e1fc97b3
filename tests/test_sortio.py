"""sortio tests (model: reference sortio/sort_test.go:67-267): external sort
with forced spills, k-way merge, merge-reduce with 1- and 2-column keys, and
the spilling combiner (reference exec/combiner_test.go:38-163)."""
import random

import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import slicefunc, sliceio, slicetype, sortio
from bigslice_b200.exec.combiner import Combiner, CombiningFrame
from bigslice_b200.frame import Frame
from bigslice_b200.internal import defaultsize


def chunks(f, n):
    class R(sliceio.Reader):
        def __init__(self):
            self.i = 0

        def read(self, k=None):
            if self.i >= len(f):
                return None
            j = min(len(f), self.i + n)
            out = f.slice(self.i, j)
            self.i = j
            return out
    return R()


def test_sort_reader_spills(monkeypatch):
    monkeypatch.setattr(defaultsize, "SortCanary", 64)
    n = 20000
    rng = random.Random(3)
    f = Frame.of(torch.tensor([rng.randrange(1 << 30) for _ in range(n)]),
                 torch.arange(n))
    import bigslice_b200.sliceio.spiller as sp
    made = []
    orig = sp.Spiller.spill

    def spy(self, frame):
        made.append(len(frame))
        return orig(self, frame)
    monkeypatch.setattr(sp.Spiller, "spill", spy)
    r = sortio.sort_reader(chunks(f, 777), 1 << 15, f.typ)
    out = sliceio.read_all(r, f.typ)
    assert len(made) > 3  # really spilled multiple runs
    keys = out.col(0).tolist()
    assert keys == sorted(keys) and len(out) == n
    assert sorted(out.rows()) == sorted(f.rows())


def test_sort_reader_small_and_empty():
    f = Frame.of(torch.tensor([3, 1, 2]), np.array(["c", "a", "b"], dtype=object))
    out = sliceio.read_all(sortio.sort_reader(sliceio.FrameReader(f), 1 << 20, f.typ), f.typ)
    assert out.rows() == [(1, "a"), (2, "b"), (3, "c")]
    e = Frame.empty(f.typ)
    assert len(sliceio.read_all(sortio.sort_reader(sliceio.FrameReader(e), 1 << 20, f.typ), f.typ)) == 0


def test_merge_reader_many_streams():
    rng = random.Random(5)
    streams, allrows = [], []
    for s in range(50):
        ks = sorted(rng.randrange(10000) for _ in range(400))
        f = Frame.of(torch.tensor(ks), torch.full((400,), s))
        allrows += f.rows()
        streams.append(chunks(f, 37))
    typ = slicetype.New(bs.int64, bs.int64)
    out = sliceio.read_all(sortio.merge_reader(typ, streams), typ)
    keys = out.col(0).tolist()
    assert keys == sorted(keys) and sorted(out.rows()) == sorted(allrows)


def test_merge_reader_string_keys():
    a = Frame.of(np.array(["a", "c", "e"], dtype=object), torch.tensor([1, 2, 3]))
    b = Frame.of(np.array(["b", "c", "d"], dtype=object), torch.tensor([4, 5, 6]))
    out = sliceio.read_all(sortio.merge_reader(a.typ, [chunks(a, 1), chunks(b, 2)]), a.typ)
    assert [r[0] for r in out.rows()] == ["a", "b", "c", "c", "d", "e"]


@pytest.mark.parametrize("fn", [bs.ops.add, lambda a, b: a + b, bs.rowwise(lambda a, b: a + b)])
def test_reduce_reader_two_column_prefix(fn):
    rng = random.Random(11)
    typ = slicetype.New(bs.int64, bs.string, bs.int64, prefix=2)
    want = {}
    streams = []
    for s in range(8):
        rows = {}
        for _ in range(300):
            k = (rng.randrange(20), "k%d" % rng.randrange(5))
            rows[k] = rows.get(k, 0) + rng.randrange(10)
        for k, v in rows.items():
            want[k] = want.get(k, 0) + v
        ks = sorted(rows)
        f = Frame.of(torch.tensor([k[0] for k in ks]), np.array([k[1] for k in ks], dtype=object),
                     torch.tensor([rows[k] for k in ks]), typ=typ)
        streams.append(chunks(f, 13))
    out = sliceio.read_all(sortio.reduce_reader(typ, streams, slicefunc.of(fn)), typ)
    got = {(a, b): c for a, b, c in out.rows()}
    assert got == want
    assert [(a, b) for a, b, _ in out.rows()] == sorted(want)


def test_reduce_reader_duplicate_keys_within_stream():
    typ = slicetype.New(bs.int64, bs.int64)
    f = Frame.of(torch.tensor([1, 1, 1, 2, 2, 3]), torch.tensor([1, 1, 1, 1, 1, 1]))
    out = sliceio.read_all(sortio.reduce_reader(typ, [chunks(f, 2), chunks(f, 1)],
                                                slicefunc.of(bs.ops.add)), typ)
    assert out.rows() == [(1, 6), (2, 4), (3, 2)]


def test_combining_frame_fuzz():
    rng = random.Random(1)
    typ = slicetype.New(bs.int64, bs.int64)
    cf = CombiningFrame(typ, slicefunc.of(bs.ops.add))
    want = {}
    for _ in range(50):
        ks = [rng.randrange(5000) for _ in range(2000)]
        vs = [rng.randrange(100) for _ in range(2000)]
        for k, v in zip(ks, vs):
            want[k] = want.get(k, 0) + v
        cf.combine(Frame.of(torch.tensor(ks), torch.tensor(vs)))
    assert dict(cf.compact().rows()) == want and len(cf) == len(want)


@pytest.mark.parametrize("fn", [bs.ops.add, bs.ops.max, lambda a, b: a + b])
def test_combiner_spills_with_tiny_target(fn):
    typ = slicetype.New(bs.string, bs.int64)
    c = Combiner(typ, slicefunc.of(fn), target_size=2)
    rng = random.Random(2)
    want = {}
    name = getattr(fn, "__bigslice_builtin__", "add")
    for _ in range(30):
        ks = ["k%d" % rng.randrange(40) for _ in range(20)]
        vs = [rng.randrange(50) for _ in range(20)]
        for k, v in zip(ks, vs):
            want[k] = max(want.get(k, v), v) if name == "max" else want.get(k, 0) + v
        c.combine(Frame.of(np.array(ks, dtype=object), torch.tensor(vs)))
    assert c.spiller is not None and c.spiller.num_files() > 3
    r = c.reader()
    out = sliceio.read_all(r, typ)
    assert dict(out.rows()) == want
    assert [k for k, _ in out.rows()] == sorted(want)
