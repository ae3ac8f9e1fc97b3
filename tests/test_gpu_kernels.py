"""Numerics tests of the sm_100a kernels against plain PyTorch references.
All marked gpu: they run on the B200 box (pytest -m gpu)."""
import random

import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200.frame import Frame, hashing

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from bigslice_b200.ops import gpu as g
    assert g.available()
    return g


def dev(t):
    return t.cuda()


@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8,
                                   torch.float32, torch.float64])
def test_hash_column_matches_reference(gpu, dtype):
    g = torch.Generator().manual_seed(1)
    if dtype.is_floating_point:
        x = torch.randn(100003, generator=g).to(dtype)
    else:
        info = torch.iinfo(dtype)
        x = torch.randint(max(info.min, -2**62), min(info.max, 2**62), (100003,), generator=g).to(dtype)
    for seed in (0, 0x9acb0442):
        want = hashing.hash_tensor(x, seed)
        got = gpu.hash_column(dev(x), seed).cpu()
        assert torch.equal(got, want)
    a = gpu.hash_column(dev(x), 0)
    b = gpu.hash_column(dev(x), 5, out=a.clone(), xor_into=True).cpu()
    assert torch.equal(b, hashing.hash_tensor(x, 0) ^ hashing.hash_tensor(x, 5))


def test_golden_placement_on_device(gpu):
    f = Frame.of(torch.arange(7).cuda())
    assert (f.hash() % 4).tolist() == [0, 0, 0, 3, 2, 3, 1]


@pytest.mark.parametrize("npart", [1, 2, 7, 16, 300])
def test_partition_perm_and_frame(gpu, npart):
    n = 200001
    g = torch.Generator().manual_seed(2)
    keys = torch.randint(0, 1 << 40, (n,), generator=g)
    vals = torch.arange(n, dtype=torch.float32)
    ids = hashing.hash64(keys, 0) % npart
    perm, counts = gpu.partition_perm(None, dev(keys), n, npart)
    assert torch.equal(counts.cpu(), torch.bincount(ids, minlength=npart))
    assert torch.equal(torch.sort(perm.cpu())[0], torch.arange(n))
    off = 0
    for d, c in enumerate(counts.tolist()):
        assert bool((ids[perm[off:off + c].cpu()] == d).all())
        off += c
    parts = gpu.partition_frame(Frame.of(dev(keys), dev(vals)), dev(ids), npart)
    assert sum(len(p) for p in parts) == n
    for d, p in enumerate(parts):
        k = p.col(0).cpu()
        assert bool((hashing.hash64(k, 0) % npart == d).all())
        assert torch.equal(keys[p.col(1).cpu().long()], k)


def zipf_keys(n, vocab, seed):
    from bigslice_b200.models.wordcount import zipf_keys as zk
    return zk(n, vocab, 1.0, seed)


@pytest.mark.parametrize("vdtype", [torch.int64, torch.int32, torch.float32, torch.float64])
@pytest.mark.parametrize("op", ["add", "min", "max"])
def test_hash_combine_matches_torch(gpu, vdtype, op):
    n = 300007
    keys = zipf_keys(n, 5000, 3)
    g = torch.Generator().manual_seed(4)
    vals = torch.randint(-50, 50, (n,), generator=g).to(vdtype)
    uk, inv = torch.unique(keys, return_inverse=True)
    if op == "add":
        want = torch.zeros(len(uk), dtype=torch.float64).index_add_(0, inv, vals.double())
    else:
        want = torch.empty(len(uk), dtype=torch.float64).scatter_reduce_(
            0, inv, vals.double(), "amin" if op == "min" else "amax", include_self=False)
    for simple, warm in ((False, None), (True, None), (False, 1024), (True, 256)):
        t = gpu.HashTable(4 * len(uk), vdtype, op, "cuda", warm=warm)
        t.combine(dev(keys), dev(vals), force_simple=simple)
        k, v = t.compact()
        order = torch.sort(k.cpu())[1]
        assert torch.equal(k.cpu()[order], uk)
        assert torch.allclose(v.cpu()[order].double(), want, rtol=1e-5, atol=1e-3)


def test_hash_combine_special_key_unaligned_and_growth(gpu):
    EMPTY = -(1 << 63)
    keys = torch.tensor([EMPTY, 5, EMPTY, 7, 5] * 5000)
    vals = torch.ones(len(keys), dtype=torch.int64)
    k, v = gpu.hash_combine(dev(keys)[1:], dev(vals)[1:], "add")  # 8-byte aligned only
    got = dict(zip(k.tolist(), v.tolist()))
    assert got == {EMPTY: 9999, 5: 10000, 7: 5000}
    many = torch.arange(50000)
    k, v = gpu.hash_combine(dev(many), dev(torch.ones_like(many)), "add", distinct_hint=100)
    assert len(k) == 50000 and int(v.sum()) == 50000  # table grew after TableFull


@pytest.mark.parametrize("n", [0, 1, 5, 4096, 100003, 1 << 20])
def test_radix_sort_stable(gpu, n):
    g = torch.Generator().manual_seed(n)
    keys = torch.randint(-(1 << 62), 1 << 62, (n,), generator=g)
    if n > 10:
        keys[: n // 2] = keys[: n // 2] % 100 - 50  # duplicates + negatives
    sk, perm = gpu.sort_pairs_i64(dev(keys), None)
    wk, wperm = torch.sort(keys, stable=True)
    assert torch.equal(sk.cpu(), wk) and torch.equal(perm.cpu(), wperm)
    f = Frame.of(dev(keys), dev(torch.arange(n)))
    assert f.sorted().col(1).cpu().tolist() == wperm.tolist()


def test_compact_segreduce_expand_crc(gpu):
    g = torch.Generator().manual_seed(9)
    m = torch.rand(100001, generator=g) < 0.3
    assert torch.equal(gpu.compact_indices(dev(m)).cpu(), torch.nonzero(m).flatten())
    assert gpu.compact_indices(dev(torch.zeros(10, dtype=torch.bool))).numel() == 0
    keys = torch.sort(torch.randint(0, 2000, (50000,), generator=g))[0]
    starts = torch.ones(50000, dtype=torch.bool)
    starts[1:] = keys[1:] != keys[:-1]
    nseg = int(starts.sum())
    seg = torch.cumsum(starts.long(), 0) - 1
    for dt in (torch.int64, torch.float32, torch.float64, torch.int32):
        vals = torch.randint(-9, 9, (50000,), generator=g).to(dt)
        for op, red in (("add", "sum"), ("min", "amin"), ("max", "amax")):
            want = torch.zeros(nseg, dtype=dt).scatter_reduce_(0, seg, vals, red, include_self=False)
            got = gpu.segmented_reduce(dev(vals), dev(starts), nseg, op).cpu()
            assert torch.allclose(got.double(), want.double())
    counts = torch.randint(0, 5, (1000,), generator=g)
    offs = torch.zeros(1001, dtype=torch.int64)
    offs[1:] = torch.cumsum(counts, 0)
    assert torch.equal(gpu.expand_indices(dev(offs)).cpu(),
                       torch.repeat_interleave(torch.arange(1000), counts))
    import zlib
    data = torch.randint(0, 256, (100001,), generator=g, dtype=torch.uint8)
    assert gpu.crc32(dev(data)) == zlib.crc32(data.numpy().tobytes()) & 0xFFFFFFFF


def test_local_device_session_end_to_end(gpu):
    """All combinators on the single-GPU executor give the CPU answers."""
    n = 50000
    g = torch.Generator().manual_seed(5)
    keys = torch.randint(0, 500, (n,), generator=g)
    vals = torch.ones(n, dtype=torch.int64)

    @bs.Func
    def pipeline(nshard: int):
        s = bs.Const(nshard, keys, vals)
        s = bs.Filter(s, lambda k, v: k % 2 == 0)
        s = bs.Map(s, lambda k, v: (k // 2, v * 3))
        return bs.Reduce(s, bs.ops.add)

    want = {}
    for k in keys.tolist():
        if k % 2 == 0:
            want[k // 2] = want.get(k // 2, 0) + 3
    before = gpu.launches()
    with bs.start(bs.LocalDevice("cuda"), bs.Parallelism(2)) as sess:
        res = sess.run(pipeline, 4)
        assert all(f.device.type == "cuda" for f in res.frames())
        assert dict(res.rows()) == want
    assert gpu.launches() > before
    with bs.start(bs.LocalDevice("cuda")) as sess:
        a = bs.Const(3, torch.tensor([1, 2, 1, 3]), torch.tensor([10., 20., 30., 40.]))
        b = bs.Const(2, torch.tensor([2, 4]), torch.tensor([5, 6]))
        res = bs.slicetest.run(bs.Cogroup(a, b), session=sess)
        assert bs.slicetest.sorted_rows(res) == [(1, (10.0, 30.0), ()), (2, (20.0,), (5,)),
                                                 (3, (40.0,), ()), (4, (), (6,))]
        f = bs.Fold(bs.Const(2, torch.tensor([1, 2, 1]), torch.tensor([1., 2., 3.])),
                    lambda acc, v: acc + v)
        assert sorted(bs.slicetest.run(f, session=sess).rows()) == [(1, 4.0), (2, 2.0)]


def test_codec_device_crc_and_ragged_take(gpu):
    """Large device columns are checksummed by the GPU CRC kernel (stitched with
    crc32_combine) and still decode/verify on the host; ListColumn gathers use
    the expansion kernel."""
    import io
    from bigslice_b200 import sliceio, slicetype
    from bigslice_b200.frame import columns as C
    n = 300000
    f = Frame.of(torch.arange(n).cuda(), torch.rand(n, dtype=torch.float64).cuda(),
                 np.array(["s%d" % (i % 7) for i in range(n)], dtype=object))
    before = gpu.launches()
    data = sliceio.encode_frame(f)
    assert gpu.launches() > before
    g = sliceio.read_all(sliceio.Decoder(io.BytesIO(data), f.typ), f.typ)
    assert torch.equal(g.col(0), f.col(0).cpu()) and torch.equal(g.col(1), f.col(1).cpu())
    bad = bytearray(data)
    bad[len(bad) // 2] ^= 4
    with pytest.raises(Exception):
        sliceio.read_all(sliceio.Decoder(io.BytesIO(bytes(bad)), f.typ), f.typ)
    lc = C.col_from_list([[1, 2], [], [3], [4, 5, 6]], slicetype.List(bs.int64), "cuda")
    assert C.col_take(lc, torch.tensor([3, 0, 1])).tolist() == [[4, 5, 6], [1, 2], []]


@pytest.mark.parametrize("warm", [0, 1 << 12, 1 << 18])
def test_two_level_table_at_scale(gpu, warm):
    """8M Zipf rows through the TMA kernel into a warm+main table: exact counts,
    also when the same table is fed in several chunks (streaming combine)."""
    from bigslice_b200.models.wordcount import zipf_keys as zk
    n = 8 << 20
    keys = zk(n, 1 << 20, 1.0, seed=5, device="cuda")
    vals = torch.randint(1, 4, (n,), device="cuda")
    uk, inv = torch.unique(keys, return_inverse=True)
    want = torch.zeros(len(uk), dtype=torch.int64, device="cuda").index_add_(0, inv, vals)
    for chunks in (1, 3):
        t = gpu.HashTable(4 * len(uk), torch.int64, "add", "cuda", warm=warm)
        step = (n + chunks - 1) // chunks
        for i in range(0, n, step):
            t.combine(keys[i:i + step], vals[i:i + step])
        k, v = t.compact()
        order = torch.sort(k)[1]
        assert torch.equal(k[order], uk) and torch.equal(v[order], want)


def test_combine_on_arrival_shape(gpu):
    """Destination-side shape: every key occurs once or twice with large partial
    counts (two ranks' tables meeting); exact totals per key, through the TMA
    kernel and the two-level table, odd row count."""
    g = torch.Generator(device="cuda").manual_seed(11)
    nk = 5_000_001
    base = torch.randperm(1 << 23, device="cuda", generator=g)[:nk] * 0x9E3779B97F4A7C15 & 0x7FFFFFFFFFFFFFFF
    dup = base[torch.rand(nk, device="cuda", generator=g) < 0.6]
    keys = torch.cat([base, dup])
    keys = keys[torch.randperm(keys.numel(), device="cuda", generator=g)]
    vals = torch.randint(1, 9_000_000, (keys.numel(),), device="cuda", generator=g)
    uk, inv = torch.unique(keys, return_inverse=True)
    want = torch.zeros(len(uk), dtype=torch.int64, device="cuda").index_add_(0, inv, vals)
    for warm in (None, 1 << 16, 0):
        for rep in range(2):
            t = gpu.HashTable(2 * len(uk), torch.int64, "add", "cuda", warm=warm)
            t.combine(keys, vals)
            k, v = t.compact()
            order = torch.sort(k)[1]
            assert torch.equal(k[order], uk)
            assert torch.equal(v[order], want), (warm, rep, int((v[order] - want).abs().sum()))


@pytest.mark.gpu
def test_device_to_host_frame_copy_is_complete():
    """Frame.to('cpu') must not hand back memory a pending async copy is still filling."""
    from bigslice_b200.frame import Frame
    from bigslice_b200 import slicetype
    typ = slicetype.Type([torch.int64, torch.int64], 1)
    for it in range(20):
        n = 1 << 22
        a = torch.arange(n, device="cuda") + it
        # queue enough work ahead of the copy that an asynchronous D2H would lag the host
        b = a
        for _ in range(8):
            b = b * 3 + 1
        host = Frame([a, b], typ).to("cpu")
        want = a.cpu()
        for _ in range(8):
            want = want * 3 + 1
        assert torch.equal(host.col(0), a.cpu())
        assert torch.equal(host.col(1), want)


@pytest.mark.gpu
def test_scan_and_clear_compaction_and_table_pool(gpu):
    """compact(clear=True) returns the same rows and leaves every slot <EMPTY, identity>,
    so a pooled table is reusable after a 32-byte status reset."""
    g = torch.Generator(device="cuda").manual_seed(3)
    for vdtype, op in ((torch.int64, "add"), (torch.float32, "max"), (torch.int32, "min")):
        keys = torch.randint(-5000, 5000, (300000,), device="cuda", generator=g)
        keys[::1000] = -(1 << 63)  # the reserved key takes the special path
        vals = torch.randint(-100, 100, (300000,), device="cuda", generator=g).to(vdtype)
        t = gpu.HashTable(1 << 15, vdtype, op, "cuda", warm=1 << 10)
        ident = t.identity_bits()
        ref = None
        for rep in range(3):
            t.combine(keys, vals)
            k, v = t.compact(clear=True)
            order = torch.argsort(k)
            cur = (k[order].cpu(), v[order].cpu())
            if ref is None:
                ref = cur
                uk = torch.unique(keys)
                assert torch.equal(cur[0], uk.cpu())
            assert torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1])
            assert bool((t.slots[:, 0] == -(1 << 63)).all())
            want_bits = ident if ident < (1 << 63) else ident - (1 << 64)
            assert bool((t.slots[:, 1] == want_bits).all())
            t.reset_status()
    pool = gpu.table_pool()
    h0 = pool.hits
    a = gpu.hash_combine(keys, vals.to(torch.int64), "add")
    b = gpu.hash_combine(keys, vals.to(torch.int64), "add")
    assert pool.hits > h0
    oa, ob = torch.argsort(a[0]), torch.argsort(b[0])
    assert torch.equal(a[0][oa], b[0][ob]) and torch.equal(a[1][oa], b[1][ob])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32, torch.float32, torch.float64])
def test_topk_radix_select_matches_torch(gpu, dtype):
    g = torch.Generator(device="cuda").manual_seed(9)
    for n, k, span in ((1, 1, 10), (1000, 7, 5), (100000, 100, 50), (3000000, 1000, 1 << 40),
                       (65537, 65537, 3), (5000, 9000, 100)):
        if dtype.is_floating_point:
            v = (torch.randn(n, device="cuda", generator=g) * span).to(dtype)
            v[::7] = 0.0  # ties, and both zeros
            v[1::97] = -0.0
        else:
            hi = min(span, 1 << 30) if dtype == torch.int32 else span
            v = torch.randint(-hi, hi, (n,), device="cuda", generator=g).to(dtype)
        vals, idx = gpu.topk(v, k)
        kk = min(k, n)
        want = torch.sort(v, descending=True, stable=True)
        assert vals.shape[0] == kk and torch.equal(vals, want.values[:kk])
        assert torch.equal(v[idx], vals) and len(set(idx.tolist())) == kk
        # equal values are reported in index order (which ties make the cut is unspecified)
        same = vals[1:] == vals[:-1]
        assert bool((idx[1:][same] > idx[:-1][same]).all())
    from bigslice_b200.exec.topn import top_n
    v = torch.tensor([3, 9, 9, 1, 7], device="cuda")
    assert top_n(v, 3) == [1, 2, 4]


def test_fold_device_tiers(gpu):
    """Fold on the device: builtin functor (hash-aggregation kernels, a key with
    > 10^6 rows), FoldExpr (NVRTC segmented fold, non-associative) — both against
    plain python/torch references."""
    import time
    from bigslice_b200.models import wordcount as wc
    n = 1 << 24
    keys = wc.zipf_keys(n, 1 << 20, 1.0, seed=3, device="cuda")
    vals = torch.randint(-5, 6, (n,), device="cuda")
    uk, inv, cnt = torch.unique(keys, return_inverse=True, return_counts=True)
    assert int(cnt.max()) >= 10**6          # the head key owns > 1M rows
    with bs.start(bs.LocalDevice("cuda")) as sess:
        for op, ref in ((bs.ops.add, "sum"), (bs.ops.min, "amin"), (bs.ops.max, "amax")):
            sl = bs.Fold(bs.Const(1, keys, vals), op, acc=bs.int64)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            f = bs.slicetest.run(sl, session=sess).to_frame()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            want = torch.zeros(len(uk), dtype=torch.int64, device="cuda")
            if ref == "sum":
                want.index_add_(0, inv, vals)
            else:
                want.scatter_reduce_(0, inv, vals, ref, include_self=True)  # accumulator starts at 0
            assert torch.equal(f.col(0), uk) and torch.equal(f.col(1), want), op
            assert dt < 2.0  # the whole run (H2D of Const, shuffle, fold); the kernel itself is ~ms
        # non-associative expression, small enough for a python reference
        m = 200000
        k2 = (torch.arange(m) * 2654435761) % 1000
        v2 = torch.arange(m) % 97
        sl = bs.FoldExpr(bs.Const(1, k2, v2), "(acc * 31 + c1) % 1000003", bs.int64)
        got = dict(bs.slicetest.run(sl, session=sess).rows())
        want = {}
        for a, b in zip(k2.tolist(), v2.tolist()):
            want[a] = (want.get(a, 0) * 31 + b) % 1000003
        assert got == want


def test_cogroup_sort_merge_kernels(gpu):
    """Device Cogroup (radix sort + run detection + binary-search union + CSR
    sizes) against a python dict reference, including keys present on one side
    only, duplicates on both sides and empty inputs."""
    from bigslice_b200.cogroup import cogroup_frames
    g = torch.Generator().manual_seed(9)
    na, nb = 200003, 50021
    ka = torch.randint(0, 60000, (na,), generator=g)
    va = torch.rand(na, generator=g)
    kb = torch.randint(30000, 90000, (nb,), generator=g)
    vb = torch.randint(0, 1000, (nb,), generator=g)
    ta = bs.Type([bs.int64, bs.float32], 1)
    tb = bs.Type([bs.int64, bs.int64], 1)
    out_t = bs.Type([bs.int64, bs.List(bs.float32), bs.List(bs.int64)], 1)
    before = gpu.launches()
    f = cogroup_frames([Frame([ka.cuda(), va.cuda()], ta), Frame([kb.cuda(), vb.cuda()], tb)], 1, out_t)
    assert gpu.launches() - before >= 6
    want = {}
    for k, v in zip(ka.tolist(), va.tolist()):
        want.setdefault(k, ([], []))[0].append(v)
    for k, v in zip(kb.tolist(), vb.tolist()):
        want.setdefault(k, ([], []))[1].append(v)
    keys = f.col(0).cpu().tolist()
    assert keys == sorted(want)
    la, lb = f.col(1), f.col(2)
    oa, ob = la.offsets.cpu().tolist(), lb.offsets.cpu().tolist()
    xa, xb = la.values.cpu().tolist(), lb.values.cpu().tolist()
    for i in (0, 1, 2, len(keys) // 2, len(keys) - 1):
        k = keys[i]
        assert sorted(xa[oa[i]:oa[i + 1]]) == pytest.approx(sorted(want[k][0]))
        assert sorted(xb[ob[i]:ob[i + 1]]) == sorted(want[k][1])
    assert oa[-1] == na and ob[-1] == nb
    # one side empty
    e = Frame([torch.empty(0, dtype=torch.int64).cuda(), torch.empty(0, dtype=torch.int64).cuda()], tb)
    f2 = cogroup_frames([Frame([ka.cuda(), va.cuda()], ta), e], 1, out_t)
    assert f2.col(0).cpu().tolist() == sorted(set(ka.tolist()))
    assert int(f2.col(2).offsets[-1]) == 0


def test_broadcast_join_device_left_join(gpu):
    from bigslice_b200.cogroup import _device_join2
    g = torch.Generator().manual_seed(4)
    ka = torch.randint(0, 5000, (60000,), generator=g)
    va = torch.rand(60000, 4, generator=g)
    kb = torch.randint(2500, 7500, (9000,), generator=g)
    vb = torch.randint(0, 99, (9000,), generator=g)
    ta = bs.Type([bs.int64, bs.Vec(bs.float32, 4)], 1)
    tb = bs.Type([bs.int64, bs.int64], 1)
    out_t = bs.Type([bs.int64, bs.List(bs.Vec(bs.float32, 4)), bs.List(bs.int64)], 1)
    f = _device_join2([Frame([ka.cuda(), va.cuda()], ta), Frame([kb.cuda(), vb.cuda()], tb)], out_t, left=True)
    keys = f.col(0).cpu().tolist()
    assert keys == sorted(set(ka.tolist()))
    want_b = {}
    for k, v in zip(kb.tolist(), vb.tolist()):
        want_b.setdefault(k, []).append(v)
    ob = f.col(2).offsets.cpu().tolist()
    xb = f.col(2).values.cpu().tolist()
    for i, k in enumerate(keys):
        assert sorted(xb[ob[i]:ob[i + 1]]) == sorted(want_b.get(k, [])), k
    assert int(f.col(1).offsets[-1]) == 60000


def test_merge_path_kernel(gpu):
    g = torch.Generator().manual_seed(11)
    for na, nb in ((0, 5), (7, 0), (1000, 1), (100003, 250007), (2048, 2048), (5, 100000)):
        a = torch.sort(torch.randint(-50000, 50000, (na,), generator=g))[0].cuda()
        b = torch.sort(torch.randint(-50000, 50000, (nb,), generator=g))[0].cuda()
        keys, src = gpu.merge_sorted_i64(a, b)
        both = torch.cat([a, b])
        want, order = torch.sort(both, stable=True)      # stable: rows of a first among equals
        assert torch.equal(keys, want), (na, nb)
        assert torch.equal(src, order), (na, nb)


def test_tiered_external_sort_on_device(gpu, monkeypatch):
    """Device runs are parked in pinned host memory, then (host budget exhausted) in
    spill files, and merged back with the merge-path kernel."""
    from bigslice_b200 import sortio
    from bigslice_b200.sliceio import FrameReader, tiers
    n = 400000
    g = torch.Generator().manual_seed(2)
    k = torch.randint(0, 1 << 40, (n,), generator=g)
    v = torch.arange(n, dtype=torch.float32)
    typ = bs.Type([bs.int64, bs.float32], 1)
    f = Frame([k.cuda(), v.cuda()], typ)
    made = []
    real = tiers.TieredSpiller

    class Spy(real):
        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            made.append(self)
    monkeypatch.setattr(tiers, "TieredSpiller", Spy)
    monkeypatch.setattr(tiers, "HOST_BUDGET", 2 << 20)       # ~2 runs fit in the host tier
    monkeypatch.setattr(sortio.sort.MergeReader, "EAGER_ROWS", 1000)   # merge run by run, not sort-it-all
    before = gpu.launches()
    r = sortio.sort_reader(FrameReader(f), 1 << 20, typ, device="cuda")   # 1 MiB runs -> several runs
    out = []
    for x in r:
        assert x.device.type == "cuda"
        out.append(x)
    got = torch.cat([x.col(0) for x in out]).cpu()
    gv = torch.cat([x.col(1) for x in out]).cpu()
    want, order = torch.sort(k, stable=True)
    assert torch.equal(got, want)
    assert torch.equal(torch.sort(gv)[0], v)                 # every row exactly once
    assert torch.equal(k[gv.long()], got)                    # payload travelled with its key
    assert made and made[0].stats["host_runs"] >= 1 and made[0].stats["file_runs"] >= 1, made[0].stats
    assert gpu.launches() > before
