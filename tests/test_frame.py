"""Frame / slicetype / hashing unit tests (model: reference frame/frame_test.go,
frame/ops_test.go, slicetype/slicetype_test.go; golden vectors SURVEY §8.3)."""
import io
import random

import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import frame as fr
from bigslice_b200 import slicetype
from bigslice_b200.frame import Frame, hashing


def fuzz_frame(n, seed=0):
    rng = random.Random(seed)
    ints = torch.tensor([rng.randrange(-1000, 1000) for _ in range(n)], dtype=torch.int64)
    floats = torch.tensor([rng.random() for _ in range(n)], dtype=torch.float64)
    strs = np.empty(n, dtype=object)
    strs[:] = ["s%d" % rng.randrange(100) for _ in range(n)]
    return Frame.of(ints, floats, strs)


def test_slicetype_basics():
    t = slicetype.New(bs.int64, bs.string, bs.float32, prefix=2)
    assert t.num_out() == 3 and t.prefix() == 2
    assert str(t) == "slice[2]<int64, string, float32>"
    assert slicetype.Assignable(t, slicetype.New(bs.int64, bs.string, bs.obj))
    assert not slicetype.Assignable(t, slicetype.New(bs.int64, bs.string))
    c = slicetype.Concat(t, slicetype.New(bs.bool_))
    assert [d.name for d in c] == ["int64", "string", "float32", "bool"]
    assert slicetype.Slice(t, 1, 3).prefix() == 1
    assert slicetype.Signature(slicetype.New(bs.int64), slicetype.New(bs.int64, bs.string)) \
        == "func(int64) (int64, string)"
    assert slicetype.dtype_of(int) is bs.int64 and slicetype.dtype_of("float32") is bs.float32
    assert slicetype.List(bs.int64).name == "[]int64"


def test_make_slice_copy():
    f = Frame.make(slicetype.New(bs.int64, bs.string), 4, 10)
    assert len(f) == 4 and f.cap() == 10
    assert f.col(0).tolist() == [0, 0, 0, 0] and list(f.col(1)) == [""] * 4
    g = fuzz_frame(6)
    h = Frame.make(g.typ, 6)
    assert fr.copy(h, g) == 6
    assert h.rows() == g.rows()
    s = g.slice(2, 5)
    assert len(s) == 3 and s.rows() == g.rows()[2:5]
    # slices share storage (zero copy)
    s.col(0)[0] = 12345
    assert g.col(0)[2].item() == 12345
    with pytest.raises(IndexError):
        g.slice(3, 99)


def test_grow_append_policy():
    f = Frame.make(slicetype.New(bs.int64), 0, 0)
    caps = set()
    for i in range(2000):
        f = f.grow(1)
        f.col(0)[i] = i
        caps.add(f.cap())
    assert f.col(0).tolist() == list(range(2000))
    assert len(caps) < 40  # amortised growth
    a, b = fuzz_frame(5, 1), fuzz_frame(7, 2)
    c = a.append_frame(b)
    assert c.rows() == a.rows() + b.rows()


def test_swap_zero_less():
    f = Frame.of(torch.tensor([3, 1, 2]), np.array(["c", "a", "b"], dtype=object))
    assert f.less(1, 0) and not f.less(0, 1) and not f.less(1, 1)
    f.swap(0, 1)
    assert f.rows() == [(1, "a"), (3, "c"), (2, "b")]
    f.zero()
    assert f.rows() == [(0, "")] * 3


def test_less_multi_column_prefix():
    f = Frame.of(torch.tensor([1, 1, 0]), torch.tensor([5, 2, 9]), prefix=2)
    assert f.less(1, 0) and f.less(2, 0) and not f.less(0, 1)
    assert f.sorted().rows() == [(0, 9), (1, 2), (1, 5)]


@pytest.mark.parametrize("n", [0, 1, 2, 100, 5000])
def test_sort_fuzz(n):
    f = fuzz_frame(n, n)
    s = f.sorted()
    keys = s.col(0).tolist()
    assert keys == sorted(keys)
    assert sorted(s.rows()) == sorted(f.rows())
    # stability: equal keys keep input order
    g = Frame.of(torch.tensor([1, 0, 1, 0]), torch.tensor([0, 1, 2, 3]))
    assert g.sorted().rows() == [(0, 1), (0, 3), (1, 0), (1, 2)]


def test_sort_strings_and_inplace():
    f = Frame.of(np.array(["b", "a", "c"], dtype=object), torch.tensor([1, 2, 3]))
    f.sort()
    assert f.rows() == [("a", 2), ("b", 1), ("c", 3)]


def test_golden_murmur3_vectors():
    want = {0: 0x63852AFC, 1: 0x53075D44, 2: 0xC61FA2F4, 3: 0xA33B57B3,
            4: 0xD0220D26, 5: 0x67C25EF7, 6: 0x6B0175D1}
    f = Frame.of(torch.arange(7))
    assert f.hash().tolist() == [want[i] for i in range(7)]
    assert [h % 4 for h in f.hash().tolist()] == [0, 0, 0, 3, 2, 3, 1]
    assert hashing.murmur3_bytes(b"a") == 0x3C2569B2
    s = Frame.of(np.array(["a"], dtype=object))
    assert s.hash().tolist() == [0x3C2569B2]


def test_hash_types_and_xor():
    # 32-bit ints hash their value widened to 4 LE bytes
    v = torch.tensor([7, -3], dtype=torch.int32)
    h = hashing.hash_tensor(v, 0).tolist()
    assert h[0] == hashing.murmur3_bytes((7).to_bytes(4, "little"))
    assert h[1] == hashing.murmur3_bytes((-3 & 0xFFFFFFFF).to_bytes(4, "little"))
    x = torch.tensor([1.5], dtype=torch.float64)
    assert hashing.hash_tensor(x, 9).tolist()[0] == \
        hashing.murmur3_bytes(np.float64(1.5).tobytes(), 9)
    x32 = torch.tensor([1.5], dtype=torch.float32)
    assert hashing.hash_tensor(x32, 0).tolist()[0] == \
        hashing.murmur3_bytes(np.float32(1.5).tobytes(), 0)
    # multi-column key = XOR of column hashes; seed changes the hash
    f = Frame.of(torch.tensor([4]), torch.tensor([5]), prefix=2)
    assert f.hash().tolist()[0] == 0xD0220D26 ^ 0x67C25EF7
    assert f.hash(1).tolist() != f.hash(0).tolist()
    b = Frame.of(torch.tensor([True, False]))
    assert b.hash(10).tolist() == [11, 10]


def test_register_ops_custom_type():
    class Pt:
        def __init__(self, x):
            self.x = x

    fr.register_ops(Pt, fr.Ops(less=lambda a, b: a.x < b.x,
                               hash_with_seed=lambda v, seed: (v.x * 31 + seed) & 0xFFFFFFFF))
    try:
        with pytest.raises(RuntimeError):
            fr.register_ops(Pt, fr.Ops())
        col = np.empty(3, dtype=object)
        col[:] = [Pt(3), Pt(1), Pt(2)]
        f = Frame.of(col, torch.tensor([0, 1, 2]))
        assert [p.x for p in f.sorted().col(0)] == [1, 2, 3]
        assert f.hash(5).tolist() == [3 * 31 + 5, 31 + 5, 2 * 31 + 5]
    finally:
        fr.unregister_ops(Pt)


def test_list_column_take_slice_concat():
    lc = fr.col_from_list([[1, 2], [], [3]], slicetype.List(bs.int64))
    assert lc.tolist() == [[1, 2], [], [3]]
    assert fr.col_take(lc, torch.tensor([2, 0])).tolist() == [[3], [1, 2]]
    assert fr.col_slice(lc, 1, 3).tolist() == [[], [3]]
    assert fr.col_concat([lc, lc]).tolist() == [[1, 2], [], [3]] * 2


def test_write_tab_and_repr():
    f = Frame.of(torch.tensor([1, 2]), np.array(["x", "y"], dtype=object))
    b = io.StringIO()
    f.write_tab(b)
    assert b.getvalue() == "1\tx\n2\ty\n"
    assert "frame[2]" in repr(f)


def test_vec_column():
    v = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    f = Frame.of(torch.tensor([2, 0, 1]), v)
    assert f.typ.out(1).name == "[4]float32"
    s = f.sorted()
    assert s.col(1)[0].tolist() == [4.0, 5.0, 6.0, 7.0]


# ---- native host data-path kernels (csrc/native.cpp) vs the torch reference ----
def _without_native(fn):
    import os
    from bigslice_b200 import _native
    os.environ["BIGSLICE_NO_NATIVE"] = "1"
    _native.reset_for_tests()
    try:
        return fn()
    finally:
        del os.environ["BIGSLICE_NO_NATIVE"]
        _native.reset_for_tests()


def test_native_host_hash_partition_combine_match_reference():
    from bigslice_b200 import _native, slicefunc
    from bigslice_b200.exec.combiner import combine_frame
    from bigslice_b200.exec.local import split_by_partition
    from bigslice_b200.frame import hashing
    if _native.lib() is None:
        pytest.skip("native runtime not built")
    g = torch.Generator().manual_seed(4)
    n = 50000
    for dt in (torch.int64, torch.int32, torch.float64, torch.float32, torch.int16, torch.uint8):
        col = (torch.randint(-(1 << 40), 1 << 40, (n,), generator=g) % 251 - 100).to(dt) \
            if dt != torch.int64 else torch.randint(-(1 << 62), 1 << 62, (n,), generator=g)
        for seed in (0, 0x9ACB0442):
            want = _without_native(lambda: hashing.hash_tensor(col, seed))
            assert torch.equal(hashing.hash_tensor(col, seed), want), (dt, seed)
    keys = torch.randint(0, 3000, (n,), generator=g)
    typ = slicetype.Type([torch.int64, torch.int64], 1)
    f = Frame([keys, torch.arange(n)], typ)
    ids = keys % 7
    parts = split_by_partition(f, ids, 7)
    ref = _without_native(lambda: split_by_partition(f, ids, 7))
    for a, b in zip(parts, ref):  # stable: identical row order inside every partition
        assert torch.equal(a.col(0), b.col(0)) and torch.equal(a.col(1), b.col(1))
    with pytest.raises(bs.errors.Error):
        split_by_partition(f, ids + 1, 7)  # shard 7 is out of range
    for vdt in (torch.int64, torch.int32, torch.float32, torch.float64):
        for op in (bs.ops.add, bs.ops.min, bs.ops.max):
            vals = torch.randint(-50, 50, (n,), generator=g).to(vdt)
            fr = Frame([keys, vals], slicetype.Type([torch.int64, vdt], 1))
            fn = slicefunc.of(op)
            got = combine_frame(fr, fn)
            want = _without_native(lambda: combine_frame(fr, fn))
            assert torch.equal(got.col(0), want.col(0)) and torch.equal(got.col(1), want.col(1)), (vdt, op)
    # integer adds wrap like the device atomics
    big = Frame([torch.zeros(4096, dtype=torch.int64), torch.full((4096,), (1 << 62) + 5)], typ)
    out = combine_frame(big, slicefunc.of(bs.ops.add))
    assert out.col(1).tolist() == [int(torch.full((4096,), (1 << 62) + 5).sum())]  # torch wraps too


def test_native_object_hash_and_group_match_reference():
    from bigslice_b200 import _native, slicefunc
    from bigslice_b200.exec.combiner import combine_frame
    if _native.lib() is None:
        pytest.skip("native runtime not built")
    rng = np.random.default_rng(2)
    words = np.array(["w%d" % i for i in rng.integers(0, 3000, 20000)] + ["", "héllo", None], dtype=object)
    for seed in (0, 77):
        want = _without_native(lambda: hashing.hash_objects(words, seed))
        assert torch.equal(hashing.hash_objects(words, seed), want)
    raw = np.array([b"ab", bytearray(b"cd"), b""], dtype=object)
    assert hashing.hash_objects(raw, 1).tolist() == [hashing.murmur3_bytes(b, 1) for b in (b"ab", b"cd", b"")]
    with pytest.raises(TypeError):
        hashing.hash_objects(np.array([1.5, "a"], dtype=object), 0)
    keys = words[:20000]
    for vdt in (torch.int64, torch.float32):
        vals = torch.from_numpy(rng.integers(-9, 9, 20000)).to(vdt)
        f = Frame([keys, vals], slicetype.Type([slicetype.string, vdt], 1))
        for op in (bs.ops.add, bs.ops.min, bs.ops.max):
            fn = slicefunc.of(op)
            got = combine_frame(f, fn)
            want = _without_native(lambda: combine_frame(f, fn))
            assert list(got.col(0)) == list(want.col(0)) and torch.equal(got.col(1), want.col(1))


def test_two_narrow_integer_key_columns_sort_with_one_packed_sort():
    import bigslice_b200 as bs
    g = torch.Generator().manual_seed(9)
    for dts in ((torch.int32, torch.int32), (torch.int16, torch.uint8), (torch.int8, torch.int32), (torch.uint8, torch.int16)):
        lo = [0 if d == torch.uint8 else -120 for d in dts]
        a = torch.randint(lo[0], 120, (3000,), generator=g).to(dts[0])
        b = torch.randint(lo[1], 120, (3000,), generator=g).to(dts[1])
        f = Frame([a, b, torch.arange(3000)], bs.Type([bs.slicetype.dtype_of(d) for d in dts] + [bs.int64], 2))
        want = sorted(range(3000), key=lambda i: (int(a[i]), int(b[i])))   # stable lexicographic order
        assert f.sort_indices().tolist() == want
        s = f.sorted()
        assert s.col(2).tolist() == want
