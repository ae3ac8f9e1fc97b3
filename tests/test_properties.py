"""Property-based tests (hypothesis) against plain-Python models — the
counterpart of the reference's gofuzz / testing/quick tiers:
frames, codec, combiner, merge/reduce streams, placement, and the combinators
end to end on the Local executor."""
import io
from collections import defaultdict

import torch
from hypothesis import HealthCheck, given, settings, strategies as st

import bigslice_b200 as bs
from bigslice_b200 import slicetype
from bigslice_b200.exec.combiner import Combiner, CombiningFrame
from bigslice_b200.frame import Frame, concat
from bigslice_b200.sliceio import Decoder, Encoder, FrameReader, read_all
from bigslice_b200.sortio import merge_reader, reduce_reader

SET = settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow])
KV = slicetype.Type([torch.int64, torch.int64], 1)

rows_st = st.lists(st.tuples(st.integers(-50, 50), st.integers(-1000, 1000)), max_size=200)


def frame_of(rows):
    if not rows:
        return Frame.empty(KV)
    k, v = zip(*rows)
    return Frame([torch.tensor(k, dtype=torch.int64), torch.tensor(v, dtype=torch.int64)], KV)


def rows_of(f):
    return list(zip(f.col(0).tolist(), f.col(1).tolist())) if len(f) else []


@SET
@given(rows_st, st.integers(0, 220), st.integers(0, 220))
def test_frame_slice_concat_take_sort_model(rows, a, b):
    f = frame_of(rows)
    lo, hi = sorted((min(a, len(rows)), min(b, len(rows))))
    assert rows_of(f.slice(lo, hi)) == rows[lo:hi]
    assert rows_of(concat([f.slice(0, lo), f.slice(lo, len(f))], KV)) == rows
    # sorted() is stable on the key prefix
    want = sorted(rows, key=lambda r: r[0])
    assert rows_of(f.sorted()) == want
    if rows:
        idx = torch.tensor([(i * 7) % len(rows) for i in range(len(rows))])
        assert rows_of(f.take(idx)) == [rows[i] for i in idx.tolist()]
        # hashing is a pure function of the key column
        h = f.hash(0)
        by_key = {}
        for (k, _), hv in zip(rows, h.tolist()):
            assert by_key.setdefault(k, hv) == hv


@SET
@given(rows_st, st.integers(1, 64))
def test_codec_roundtrip_any_chunking(rows, chunk):
    f = frame_of(rows)
    buf = io.BytesIO()
    enc = Encoder(buf)
    for i in range(0, len(f), chunk):
        enc.encode(f.slice(i, min(i + chunk, len(f))))
    buf.seek(0)
    dec = Decoder(buf, KV)
    got = []
    while True:
        g = dec.read(None)
        if g is None:
            break
        got.extend(rows_of(g))
    assert got == rows


@SET
@given(st.lists(rows_st, min_size=1, max_size=5), st.sampled_from(["add", "min", "max"]))
def test_combiner_matches_dict_model(batches, opname):
    op = {"add": bs.ops.add, "min": bs.ops.min, "max": bs.ops.max}[opname]
    py = {"add": lambda a, b: a + b, "min": min, "max": max}[opname]
    from bigslice_b200 import slicefunc
    fn = slicefunc.of(op)
    model = {}
    cf = CombiningFrame(KV, fn)
    comb = Combiner(KV, fn, target_size=16)  # tiny target: forces spills
    for rows in batches:
        for k, v in rows:
            model[k] = py(model[k], v) if k in model else v
        cf.combine(frame_of(rows))
        comb.combine(frame_of(rows))
    assert dict(rows_of(cf.compact())) == model
    out = read_all(comb.reader(), KV)
    got = rows_of(out)
    assert dict(got) == model and [k for k, _ in got] == sorted(model)  # one row per key, sorted
    comb.discard()


@SET
@given(st.lists(rows_st, min_size=1, max_size=6))
def test_merge_and_reduce_readers_model(streams):
    from bigslice_b200 import slicefunc
    sorted_streams = [sorted(s, key=lambda r: r[0]) for s in streams]
    merged = read_all(merge_reader(KV, [FrameReader(frame_of(s)) for s in sorted_streams]), KV)
    want = sorted((r for s in sorted_streams for r in s), key=lambda r: r[0])
    got = rows_of(merged)
    assert [k for k, _ in got] == [k for k, _ in want] and sorted(got) == sorted(want)
    model = defaultdict(int)
    for s in streams:
        for k, v in s:
            model[k] += v
    red = read_all(reduce_reader(KV, [FrameReader(frame_of(s)) for s in sorted_streams],
                                 slicefunc.of(bs.ops.add)), KV)
    assert rows_of(red) == sorted(model.items())


@bs.Func
def _prop_reduce(nshard: int, keys, vals):
    return bs.Reduce(bs.Const(nshard, keys, vals), bs.ops.add)


@bs.Func
def _prop_cogroup(nshard: int, ka, va, kb, vb):
    return bs.Cogroup(bs.Const(nshard, ka, va), bs.Const(max(nshard - 1, 1), kb, vb))


@bs.Func
def _prop_reshard(nshard: int, mshard: int, keys, vals):
    return bs.Reshard(bs.Const(nshard, keys, vals), mshard)


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(rows_st, rows_st, st.integers(1, 5), st.integers(1, 5))
def test_combinators_end_to_end_model(a, b, nshard, mshard):
    ka, va = frame_of(a).col(0), frame_of(a).col(1)
    kb, vb = frame_of(b).col(0), frame_of(b).col(1)
    with bs.start(bs.Local, bs.Parallelism(2)) as sess:
        model = defaultdict(int)
        for k, v in a:
            model[k] += v
        assert dict(sess.run(_prop_reduce, nshard, ka, va).rows()) == dict(model)
        # Reshard: a permutation of the rows, every key in exactly one shard, at the golden placement
        res = sess.run(_prop_reshard, nshard, mshard, ka, va)
        seen = []
        for shard, t in enumerate(res.tasks):
            for f in sess.executor.reader(t, 0):
                for k, v in rows_of(f):
                    seen.append((k, v))
                    assert int(frame_of([(k, v)]).hash(0)[0]) % mshard == shard or nshard == mshard
        assert sorted(seen) == sorted(a)
        # Cogroup: per key the multisets of both sides
        ga, gb = defaultdict(list), defaultdict(list)
        for k, v in a:
            ga[k].append(v)
        for k, v in b:
            gb[k].append(v)
        got = {}
        for k, la, lb in sess.run(_prop_cogroup, nshard, ka, va, kb, vb).rows():
            got[k] = (sorted(la), sorted(lb))
        assert got == {k: (sorted(ga.get(k, [])), sorted(gb.get(k, []))) for k in set(ga) | set(gb)}
