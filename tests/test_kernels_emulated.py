"""Hand-written CUDA kernels, run on the CPU through the SIMT emulator (tests/emu).

``csrc/{strings,merge,cogroup,misc,hash_partition,sort,hash_combine}.cu`` are compiled *unmodified* (kernels and
host wrappers; launches rewritten, inline PTX dropped) against ``tests/emu/simt.h``
and loaded in place of the CUDA library, so the framework's own python wrappers
(``ops/gpu.py``, ``ops/strings.py``) drive the real kernels over CPU tensors:
indexing, tails, shared memory, ballots/shuffles, atomics and the host-side glue
between launches are all exercised by the CPU test suite.  A second pass runs the
same calls with AddressSanitizer: a memcheck of the device code without a GPU.
(The GPU suite compares the same kernels with PyTorch references on a B200.)
"""
import ctypes
import os
import shutil
import subprocess
import sys
import zlib

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import emu  # noqa: E402

from bigslice_b200.frame import hashing  # noqa: E402
from bigslice_b200.ops import gpu, strings  # noqa: E402

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
FILES = ["strings", "merge", "cogroup", "misc", "hash_partition", "sort", "hash_combine"]


class EmuLib:
    """Looks a ``bs_*`` symbol up in whichever emulated object exports it."""

    def __init__(self, asan=False):
        self.libs = [ctypes.CDLL(emu.build(f, asan=asan)) for f in FILES]

    def __getattr__(self, name):
        for lib in self.libs:
            try:
                return getattr(lib, name)
            except AttributeError:
                continue
        raise AttributeError(name)


@pytest.fixture
def emulated(monkeypatch):
    lib = EmuLib(asan=bool(os.environ.get("BS_EMU_ASAN")))
    monkeypatch.setattr(gpu, "require", lambda: lib)
    monkeypatch.setattr(gpu, "_lib", lib)
    monkeypatch.setattr(gpu, "_pool", gpu.TablePool())
    monkeypatch.setattr(gpu, "_stream", lambda: ctypes.c_void_p(0))
    monkeypatch.setattr(strings, "_cuda", lambda col: True)
    return lib


G = torch.Generator().manual_seed(3)


def test_compaction_segmented_reduce_and_expand(emulated):
    for n in (1, 31, 2049, 10007):
        mask = torch.rand(n, generator=G) < 0.37
        assert torch.equal(gpu.compact_indices(mask), torch.nonzero(mask).flatten())
    assert gpu.compact_indices(torch.zeros(5000, dtype=torch.bool)).numel() == 0
    # segmented reduce: every op, integer and float values, segments that straddle warps and CTAs
    n = 6000
    starts = torch.rand(n, generator=G) < 0.02
    starts[0] = True
    seg = torch.cumsum(starts.to(torch.int64), 0) - 1
    nseg = int(seg[-1]) + 1
    for dt in (torch.int64, torch.int32, torch.float32, torch.float64):
        vals = torch.randint(-50, 50, (n,), generator=G).to(dt)
        for op, red in (("add", "sum"), ("min", "amin"), ("max", "amax")):
            want = torch.zeros(nseg, dtype=dt).scatter_reduce_(0, seg, vals, red, include_self=False)
            assert torch.equal(gpu.segmented_reduce(vals, starts, nseg, op), want), (dt, op)
    counts = torch.randint(0, 5, (777,), generator=G)
    offs = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(counts, 0)])
    assert torch.equal(gpu.expand_indices(offs), torch.repeat_interleave(torch.arange(777), counts))


def test_hashing_partitioning_gather_and_radix_sort(emulated):
    # bit-exact murmur3 of every key width, seeds, XOR accumulation across key columns
    for dt in (torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8, torch.float32, torch.float64):
        c = torch.randint(-100, 100, (3001,), generator=G).to(dt)
        for seed in (0, 77):
            assert torch.equal(gpu.hash_column(c, seed), hashing.hash_tensor(c, seed)), (dt, seed)
    a, b = torch.randint(-9, 9, (500,), generator=G), torch.randint(0, 5, (500,), generator=G).to(torch.int32)
    h = gpu.hash_column(a)
    gpu.hash_column(b, out=h, xor_into=True)
    assert torch.equal(h, hashing.hash_tensor(a, 0) ^ hashing.hash_tensor(b, 0))
    # partition permutation: keys hashed in the kernel, and explicit ids
    keys = torch.randint(-1000, 1000, (9001,), generator=G)
    for npart in (1, 7, 64):
        dest = hashing.hash_tensor(keys, 0) % npart
        for ids, ks in ((None, keys), (dest, None)):
            perm, counts = gpu.partition_perm(ids, ks, keys.numel(), npart)
            assert torch.equal(counts, torch.bincount(dest, minlength=npart))
            assert torch.equal(torch.sort(perm).values, torch.arange(keys.numel()))   # a permutation
            grouped = dest[perm]
            assert torch.equal(grouped, torch.sort(grouped).values)                      # grouped by destination
    # gather of wide and narrow rows
    wide = torch.randn(700, 4, generator=G)
    idx = torch.randint(0, 700, (1234,), generator=G)
    assert torch.equal(gpu.gather_rows(wide, idx), wide[idx])
    assert torch.equal(gpu.gather_rows(keys, idx), keys[idx])
    # stable LSD radix sort of (key, payload) pairs
    for n in (2, 300, 5000):
        k = torch.randint(-50, 50, (n,), generator=G) * (1 << 33) + torch.randint(0, 3, (n,), generator=G)
        sk, perm = gpu.sort_pairs_i64(k, None)
        want = torch.sort(k, stable=True)
        assert torch.equal(sk, want.values) and torch.equal(perm, want.indices)


def _reduce_reference(k, v, op):
    uk, inv = torch.unique(k, return_inverse=True)
    red = {"add": "sum", "min": "amin", "max": "amax"}[op]
    return uk, torch.zeros(uk.numel(), dtype=v.dtype).scatter_reduce_(0, inv, v, red, include_self=False)


def test_flagship_hash_combine_kernel_with_its_tma_pipeline(emulated):
    """``hash_combine_tma_kernel``: the producer warp's TMA loads and the mbarrier
    full/empty pipeline run under the emulator's mbarrier model (tests/emu/ptx_emu.h);
    shared-memory hot table, warm level, main level, the ragged tail kernel, the
    reserved EMPTY key, destination counting and overflow reporting."""
    n = 40000                                    # > 4 tiles: the TMA kernel takes the aligned part
    k = (torch.randint(0, 3000, (n,), generator=G) ** 2) * 7919 - 12345
    k[::97] = -(1 << 63)                         # the table's EMPTY marker as a real key
    for dt in (torch.int64, torch.int32, torch.float32, torch.float64):
        v = torch.randint(-9, 9, (n,), generator=G).to(dt)
        for op in ("add", "min", "max"):
            uk, acc = gpu.hash_combine(k, v, op)
            wk, wv = _reduce_reference(k, v, op)
            o = torch.sort(uk).indices
            assert torch.equal(uk[o], wk) and torch.equal(acc[o], wv), (dt, op)
    # a table with a warm (L2) level in front, fed tile by tile, then scan-and-clear
    v = torch.ones(n, dtype=torch.int64)
    t = gpu.HashTable(8192, torch.int64, "add", "cpu", warm=1024)
    t.track(8)                                   # count first-time insertions per destination shard
    for part in torch.chunk(torch.arange(n), 3):
        t.combine(k[part], v[part])
    uk, acc = t.compact(clear=True)
    wk, wv = _reduce_reference(k, v, "add")
    o = torch.sort(uk).indices
    assert torch.equal(uk[o], wk) and torch.equal(acc[o], wv)
    dest = hashing.hash_tensor(wk[wk != -(1 << 63)], 0) % 8        # the EMPTY key is routed out of band
    assert torch.equal(t.dest_counts[:8], torch.bincount(dest, minlength=8))
    assert int((t.slots[:, 0] != -(1 << 63)).sum()) == 0           # the table is empty again
    # overflow is reported, not silent: 3000 distinct keys cannot fit 1024 + 256 slots
    small = gpu.HashTable(1024, torch.int64, "add", "cpu", warm=256)
    small.combine(k, v)
    assert small.failed() > 0
    with pytest.raises(gpu.TableFull):
        small.compact()
    # compaction into a buffer that is too small keeps writing inside it and reports the need
    t2 = gpu.HashTable(8192, torch.int64, "add", "cpu", warm=0)
    t2.combine(k, v)
    ok, ov = torch.full((100,), -1), torch.full((100,), -1)
    t2.compact_into(ok, ov, clear=False)
    assert int(t2.status[3]) == wk.numel() and int((ok != -1).sum()) == 100


def test_topk_and_crc32(emulated):
    for dt in (torch.int64, torch.float32, torch.int32):
        v = torch.randint(-1000, 1000, (5003,), generator=G).to(dt)
        for k in (1, 10, 300):
            vals, idx = gpu.topk(v, k)
            assert torch.equal(vals, torch.topk(v, k).values) and torch.equal(v[idx], vals)
    data = torch.randint(0, 256, (100003,), generator=G).to(torch.uint8)
    assert gpu.crc32(data) == zlib.crc32(data.numpy().tobytes())


def test_merge_path_is_a_stable_merge(emulated):
    for na, nb in ((0, 5), (7, 0), (100, 3), (2048, 2048), (5000, 3001)):
        a = torch.sort(torch.randint(0, 200, (na,), generator=G)).values
        b = torch.sort(torch.randint(0, 200, (nb,), generator=G)).values
        keys, src = gpu.merge_sorted_i64(a, b)
        both = torch.cat([a, b])
        want = torch.sort(both, stable=True)      # a-rows first among equal keys == stable sort of [a | b]
        assert torch.equal(keys, want.values) and torch.equal(src, want.indices)


def test_cogroup_union_kernels(emulated):
    a = torch.sort(torch.randint(0, 300, (4000,), generator=G)).values
    b = torch.sort(torch.randint(100, 500, (1500,), generator=G)).values
    sa, sb = gpu.run_starts(a), gpu.run_starts(b)
    ua, ca = torch.unique_consecutive(a, return_counts=True)
    ub, cb = torch.unique_consecutive(b, return_counts=True)
    assert torch.equal(a[sa], ua) and torch.equal(b[sb], ub)
    r, m = gpu.cogroup_search(ua, ub)
    assert torch.equal(r, torch.searchsorted(ub, ua)) and torch.equal(m.bool(), torch.isin(ua, ub))
    uk, za, zb = gpu.cogroup_union2(ua, sa, a.numel(), ub, sb, b.numel())
    want = torch.unique(torch.cat([ua, ub]))
    assert torch.equal(uk, want)
    da, db = dict(zip(ua.tolist(), ca.tolist())), dict(zip(ub.tolist(), cb.tolist()))
    assert za.tolist() == [da.get(k, 0) for k in want.tolist()]
    assert zb.tolist() == [db.get(k, 0) for k in want.tolist()]


def test_string_kernels(emulated):
    words = ["", "a", "hello", "naïve café", "x" * 37, "The quick  brown\tfox", "https://User@Example.COM:8080/a?b#c",
             "b.org/y z", "ftp://HOST", "no-host-here?", "tail "] * 9
    col = strings.from_objects(words)
    for seed in (0, 0x9747b28c):
        assert torch.equal(strings.hash32(col, seed),
                           hashing.hash_objects(np.array(words, dtype=object), seed))
    ids = strings.key64(col)
    lo, hi = strings.hash32(col, 0), strings.hash32(col, 0x9747b28c)
    assert torch.equal(ids, (lo | (hi << 32)).to(torch.int64))      # (wraps to the signed image)
    toks, parent = strings.split(col)
    want_t, want_p = [], []
    for i, w in enumerate(words):
        for t in w.split():
            want_t.append(t)
            want_p.append(i)
    assert toks.tolist() == want_t and parent.tolist() == want_p
    assert strings.url_domain(col).tolist()[:11] == ["", "a", "hello", "naïve café", "x" * 37, "the quick  brown\tfox",
                                                     "example.com:8080", "b.org", "host", "no-host-here", "tail "]
    idx = torch.tensor([10, 2, 2, 0, 6])
    assert strings.take(col, idx).tolist() == [words[i] for i in idx.tolist()]
    d_ids, d_strs = strings.dictionary(col, ids)
    assert strings.verify_dictionary(col, ids, d_ids, d_strs)
    forged = ids.clone()
    forged[1] = ids[2]                                               # "a" claims the id of "hello"
    assert not strings.verify_dictionary(col, forged, d_ids, d_strs)


@pytest.mark.skipif(not os.path.exists(emu.asan_preload().split(":")[0]), reason="needs libasan")
def test_the_same_calls_are_clean_under_address_sanitizer():
    """Every kernel above once more, in a subprocess with the emulated objects built
    with -fsanitize=address: an out-of-bounds load or store anywhere in the device code
    (or in the host glue that sizes the buffers) aborts the run."""
    env = dict(os.environ, BS_EMU_ASAN="1", LD_PRELOAD=emu.asan_preload(),
               ASAN_OPTIONS="detect_leaks=0:abort_on_error=1:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", __file__,
                        "-k", "not address_sanitizer and not order_threads_run and not random_shapes"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "passed" in r.stdout and "AddressSanitizer" not in r.stderr


def test_hash_combine_random_shapes(emulated):
    """Randomised tables through the combine kernels: sizes around the TMA kernel's
    tile threshold, warm levels, high load factors, colliding key patterns, the reserved
    key, every op and value type.  Either the result is exact or the table says it
    overflowed — never a silently wrong aggregate."""
    import random
    rnd = random.Random(2024)
    patterns = ("uniform", "strided", "few", "negative")
    for trial in range(40):
        n = rnd.choice([1, 17, 511, 4096, 16384, 20001, 33000])
        pat = rnd.choice(patterns)
        nk = max(1, rnd.choice([1, 7, n // 50 + 1, n // 3 + 1, n]))
        g = torch.Generator().manual_seed(trial)
        base = torch.randint(0, nk, (n,), generator=g)
        if pat == "strided":
            k = base * 4096                        # many keys share low bits
        elif pat == "few":
            k = base % 5
        elif pat == "negative":
            k = -base * 7919 - 1
        else:
            k = base * 2654435761 % (1 << 40)
        if rnd.random() < 0.3:
            k[:: max(1, n // 9)] = -(1 << 63)
        dt = rnd.choice([torch.int64, torch.int32, torch.float32, torch.float64])
        op = rnd.choice(["add", "min", "max"])
        v = torch.randint(-20, 20, (n,), generator=g).to(dt)
        distinct = int(torch.unique(k).numel())
        cap = max(1024, int(distinct * rnd.choice([0.5, 1.2, 2.0, 4.0])))
        warm = rnd.choice([0, 0, 256, 1024])
        t = gpu.HashTable(cap, dt, op, "cpu", warm=warm)
        for part in torch.chunk(torch.arange(n), rnd.choice([1, 2, 5])):
            if part.numel():
                t.combine(k[part], v[part], force_simple=rnd.random() < 0.2)
        wk, wv = _reduce_reference(k, v, op)
        if t.failed():
            assert distinct > 0.5 * t.total, (trial, distinct, t.total)   # only a really full table may give up
            continue
        uk, acc = t.compact()
        o = torch.sort(uk).indices
        assert torch.equal(uk[o], wk) and torch.equal(acc[o], wv), (trial, n, pat, nk, dt, op, cap, warm)


def test_kernels_do_not_depend_on_the_order_threads_run_in():
    """Schedule fuzzing: the emulator visits CTAs and, on every pass, runnable threads in
    a random order (``BS_EMU_SCHED=<seed>``).  A kernel that is only right because thread
    i happens to run before thread i+1 between two barriers — how the shared verdict
    variable of ``xchg_plan_kernel`` went wrong — fails here.  Kernels, generated kernels
    and the multi-rank exchanges are all re-run under one fuzzed schedule."""
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, BS_EMU_SCHED="20260921")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider",
                        os.path.join(here, "test_kernels_emulated.py"), os.path.join(here, "test_exchange_emulated.py"),
                        os.path.join(here, "test_jit_emulated.py"),
                        "-k", "not address_sanitizer and not order_threads_run"],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "passed" in r.stdout


def test_n_way_cogroup_composed_from_the_union_kernels(emulated, monkeypatch):
    """Cogroup of one, three and four inputs with int64 / int32 / uint8 keys on the
    device path (``cogroup._device_join_n``: sort, run detection, pairwise key unions,
    one search per input) equals the generic host merge, group for group."""
    import bigslice_b200 as bs
    from bigslice_b200 import cogroup
    from bigslice_b200.frame import Frame

    def frames(kdt, n_inputs):
        out = []
        for i in range(n_inputs):
            n = [700, 0, 333, 41][i]
            k = torch.randint(0, 60 + 20 * i, (n,), generator=G).to(kdt)
            v = torch.randn(n, generator=G)
            w = torch.randint(0, 9, (n,), generator=G)
            cols, typ = ([k, v], [kdt, torch.float32]) if i % 2 == 0 else ([k, v, w], [kdt, torch.float32, torch.int64])
            out.append(Frame(cols, bs.Type([bs.slicetype.dtype_of(d) for d in typ], 1)))
        return out

    def rows(f):
        return sorted((r[0], tuple(sorted(x) if isinstance(x, (list, tuple)) else x for x in r[1:])) for r in f.rows())
    for kdt in (torch.int64, torch.int32, torch.uint8):
        for n_inputs in (1, 3, 4):
            fs = frames(kdt, n_inputs)
            vals = []
            for f in fs:
                vals += [bs.List(c) for c in f.typ.cols[1:]]
            out_typ = bs.Type([bs.slicetype.dtype_of(kdt)] + vals, 1)
            monkeypatch.setattr(cogroup, "_on_device", lambda t: False)
            want = cogroup.cogroup_frames(fs, 1, out_typ)
            monkeypatch.setattr(cogroup, "_on_device", lambda t: isinstance(t, torch.Tensor))
            got = cogroup._device_join_n(fs, out_typ)
            assert got is not None and got.col(0).dtype == kdt
            assert rows(got) == rows(want), (kdt, n_inputs)
            assert torch.equal(got.col(0), torch.sort(got.col(0)).values)       # keys come out sorted
    # a two-column key of narrow integers: packed into one order-preserving word
    def frames2(n_inputs):
        out = []
        for i in range(n_inputs):
            n = [500, 300, 77][i]
            a = torch.randint(-6, 6, (n,), generator=G).to(torch.int32)
            b = torch.randint(-3, 3, (n,), generator=G).to(torch.int16)
            out.append(Frame([a, b, torch.randn(n, generator=G)], bs.Type([bs.int32, bs.int16, bs.float32], 2)))
        return out
    for n_inputs in (2, 3):
        fs = frames2(n_inputs)
        out_typ = bs.Type([bs.int32, bs.int16] + [bs.List(bs.float32)] * n_inputs, 2)
        monkeypatch.setattr(cogroup, "_on_device", lambda t: False)
        want = cogroup.cogroup_frames(fs, 2, out_typ)
        monkeypatch.setattr(cogroup, "_on_device", lambda t: isinstance(t, torch.Tensor))
        got = cogroup.cogroup_frames(fs, 2, out_typ)
        key = lambda f: [(r[0], r[1], tuple(tuple(sorted(x)) for x in r[2:])) for r in f.rows()]   # noqa: E731
        assert key(got) == key(want)                                 # same groups, same (lexicographic) order
        assert got.col(0).dtype == torch.int32 and got.col(1).dtype == torch.int16
    # not covered: float keys, a 64-bit column in a two-column key -> the generic path
    monkeypatch.setattr(cogroup, "_on_device", lambda t: isinstance(t, torch.Tensor))
    f = Frame([torch.rand(5), torch.rand(5)], bs.Type([bs.float32, bs.float32], 1))
    assert cogroup._device_join_n([f], bs.Type([bs.float32, bs.List(bs.float32)], 1)) is None
    f = Frame([torch.arange(5), torch.arange(5).to(torch.int32), torch.rand(5)], bs.Type([bs.int64, bs.int32, bs.float32], 2))
    assert cogroup._device_join_n([f], bs.Type([bs.int64, bs.int32, bs.List(bs.float32)], 2), 2) is None


def test_hash_combine_of_frames_with_packed_key_prefixes(emulated):
    """``gpu.hash_combine_frame`` (the in-process GPU executor's Reduce): narrow, float
    and two-column key prefixes are packed into the table's 64-bit key and unpacked."""
    import bigslice_b200 as bs
    from bigslice_b200.frame import Frame
    n = 20000
    cases = {
        "int32": ([torch.randint(-300, 300, (n,), generator=G).to(torch.int32)], [bs.int32]),
        "float64": ([torch.randint(-50, 50, (n,), generator=G).to(torch.float64) * 0.5], [bs.float64]),
        "int16+uint8": ([torch.randint(-20, 20, (n,), generator=G).to(torch.int16),
                         torch.randint(0, 9, (n,), generator=G).to(torch.uint8)], [bs.int16, bs.uint8]),
        "float32+int32": ([torch.randint(-4, 4, (n,), generator=G).to(torch.float32) * 0.25,
                           torch.randint(-9, 9, (n,), generator=G).to(torch.int32)], [bs.float32, bs.int32]),
    }
    v = torch.randint(1, 4, (n,), generator=G)
    for name, (keys, kt) in cases.items():
        f = Frame(keys + [v], bs.Type(kt + [bs.int64], len(keys)))
        assert gpu.can_hash_combine(f, "add"), name
        out = gpu.hash_combine_frame(f, "add")
        want = {}
        for row in zip(*[k.tolist() for k in keys], v.tolist()):
            want[row[:-1]] = want.get(row[:-1], 0) + row[-1]
        got = {tuple(r[:-1]): r[-1] for r in out.rows()}
        assert got == want, name
        assert [c.dtype for c in out.cols()[:-1]] == [k.dtype for k in keys]
    # not packable: three key columns, a 64-bit column in a two-column key, a string key
    k = torch.arange(4)
    assert not gpu.can_hash_combine(Frame([k, k.to(torch.int32), v[:4]], bs.Type([bs.int64, bs.int32, bs.int64], 2)), "add")
    assert not gpu.can_hash_combine(Frame([k, k, k, v[:4]], bs.Type([bs.int64] * 4, 3)), "add")


def test_merge_reader_device_step_merges_sorted_streams_with_the_merge_path_kernel(emulated, monkeypatch):
    """``sortio.MergeReader`` over several key-sorted streams: the device step (emit bound
    and per-stream cuts by searchsorted, pairwise stable merges with ``csrc/merge.cu``)
    yields the same rows, in the same stable order, as sorting everything at once."""
    import bigslice_b200 as bs
    from bigslice_b200 import sortio
    from bigslice_b200.frame import Frame
    from bigslice_b200.sliceio import FrameReader
    from bigslice_b200.sortio import sort as sortmod
    monkeypatch.setattr(sortmod, "_device_key", lambda k: isinstance(k, torch.Tensor) and k.dtype == torch.int64)
    monkeypatch.setattr(sortmod.MergeReader, "EAGER_ROWS", 100)     # merge step by step, do not sort it all
    typ = bs.Type([bs.int64, bs.int64, bs.float32], 1)
    streams, allk, allid = [], [], []
    base = 0
    for s in range(5):
        n = [4000, 1, 2500, 0, 3333][s]
        k = torch.sort(torch.randint(0, 500, (n,), generator=G)).values
        ident = torch.arange(base, base + n)
        base += n
        allk.append(k)
        allid.append(ident)
        streams.append(FrameReader(Frame([k, ident, torch.randn(n, generator=G)], typ)))
    calls = []
    real = gpu.merge_sorted_i64
    monkeypatch.setattr(gpu, "merge_sorted_i64", lambda a, b: (calls.append(1), real(a, b))[1])
    out = list(sortio.merge_reader(typ, streams, tile=1024))
    assert calls, "the merge-path kernel was not used"
    gk = torch.cat([f.col(0) for f in out])
    gid = torch.cat([f.col(1) for f in out])
    k, ident = torch.cat(allk), torch.cat(allid)
    order = torch.sort(k, stable=True).indices                         # streams in order = stable across streams
    assert torch.equal(gk, k[order])
    assert torch.equal(torch.sort(gid).values, ident)                  # every row exactly once
    assert torch.equal(k[gid], gk)                                     # payloads travelled with their keys


def test_text_pipeline_runs_the_string_kernels(emulated):
    """lines -> Tokenize -> StringIds -> Reduce(add) and urls -> UrlDomains: the text
    combinators drive the tokeniser / id / URL-host / span kernels (emulated here) and
    verify every batch's dictionary."""
    import bigslice_b200 as bs
    from bigslice_b200 import slicetest, text
    words = ["alpha", "beta", "gamma", "δelta", "e"]
    lines = np.array([" ".join(words[(i + j) % 5] for j in range(1 + i % 4)) for i in range(400)], dtype=object)
    d = text.IdDictionary()
    s = bs.Reduce(text.StringIds(text.Tokenize(bs.Const(3, lines)), d), bs.ops.add)
    rows = slicetest.run(s).rows()
    names = d.lookup([k for k, _ in rows])
    want = {}
    for ln in lines.tolist():
        for w in ln.split():
            want[w] = want.get(w, 0) + 1
    assert {names[k]: v for k, v in rows} == want
    urls = np.array(["https://A.example.com/x?y", "http://b.org", "mailto:nobody", "c.net/p", ""] * 30, dtype=object)
    doms = [r[0] for r in slicetest.run(text.UrlDomains(bs.Const(2, urls))).rows()]
    # (same answers as the host implementation: no scheme -> everything up to the first '/')
    assert sorted(set(doms)) == ["a.example.com", "b.org", "c.net", "mailto:nobody"]
    assert doms.count("b.org") == 30 and len(doms) == 120            # rows without a host are dropped
