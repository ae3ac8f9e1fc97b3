"""Key packing for the device tables: round trip, and the partition hash the
kernels compute from the packed word equals Frame.hash (the reference placement)."""
import torch

from bigslice_b200 import slicetype as st
from bigslice_b200.frame import Frame, hashing
from bigslice_b200.ops import keypack


def _bits(x):
    if x.dtype == torch.float32:
        return x.view(torch.int32)
    if x.dtype == torch.float64:
        return x.view(torch.int64)
    return x


def _device_mode_hash(packed, mode):
    """What csrc/reduce_exchange.cu: rx_dest computes, with the host murmur3."""
    if mode == 0:
        return hashing.hash64(packed, 0)
    lo = (packed & 0xFFFFFFFF)
    lo = torch.where(lo >= 0x80000000, lo - (1 << 32), lo).to(torch.int32)
    h = hashing.hash_tensor(lo, 0)
    if mode == 2:
        hi = ((packed >> 32) & 0xFFFFFFFF)
        hi = torch.where(hi >= 0x80000000, hi - (1 << 32), hi).to(torch.int32)
        h = h ^ hashing.hash_tensor(hi, 0)
    return h


def test_pack_roundtrip_and_placement():
    g = torch.Generator().manual_seed(0)
    n = 5000
    i32 = torch.randint(-2**31, 2**31 - 1, (n,), generator=g).to(torch.int32)
    f32 = torch.randn(n, generator=g)
    i8 = torch.randint(-128, 127, (n,), generator=g).to(torch.int8)
    u8 = torch.randint(0, 255, (n,), generator=g).to(torch.uint8)
    i16 = torch.randint(-2**15, 2**15 - 1, (n,), generator=g).to(torch.int16)
    f64 = torch.randn(n, generator=g, dtype=torch.float64)
    i64 = torch.randint(-2**62, 2**62, (n,), generator=g)
    v = torch.zeros(n, dtype=torch.int64)
    cases = [([i64], [st.int64]), ([f64], [st.float64]), ([i32], [st.int32]), ([f32], [st.float32]),
             ([i8], [st.int8]), ([u8], [st.uint8]), ([i16], [st.int16]),
             ([i32, f32], [st.int32, st.float32]), ([i8, i16], [st.int8, st.int16]), ([u8, i32], [st.uint8, st.int32])]
    for cols, dts in cases:
        typ = st.Type(dts + [st.int64], len(cols))
        lay = keypack.layout_for(typ)
        assert lay is not None, dts
        p = keypack.pack(cols, lay)
        assert p.dtype == torch.int64
        back = keypack.unpack(p, lay)
        for a, b in zip(cols, back):
            assert a.dtype == b.dtype and torch.equal(_bits(a), _bits(b)), dts
        want = Frame(list(cols) + [v], typ).hash(0)
        assert torch.equal(_device_mode_hash(p, lay.mode), want), dts


def test_layout_rejects_what_does_not_fit():
    assert keypack.layout_for(st.Type([st.int64, st.int32, st.int64], 2)) is None      # 12 bytes
    assert keypack.layout_for(st.Type([st.string, st.int64], 1)) is None
    assert keypack.layout_for(st.Type([st.int32, st.int32, st.int32, st.int64], 3)) is None
    assert keypack.layout_for(st.Type([st.int64, st.int64], 1)).plain_int64
