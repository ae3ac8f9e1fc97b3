"""sliceio tests (model: reference sliceio/codec_test.go, reader_test.go,
spiller_test.go): codec round trips over chunk sizes, custom codecs with
session state, single-bit corruption must error, reader adapters, scanner."""
import io
import json
import random

import numpy as np
import pytest
import torch

import bigslice_b200 as bs
from bigslice_b200 import errors, frame as fr, sliceio, slicetype
from bigslice_b200.frame import Frame


def mixed_frame(n, seed=1):
    rng = random.Random(seed)
    return Frame.of(
        torch.tensor([rng.randrange(1 << 40) for _ in range(n)], dtype=torch.int64),
        torch.tensor([rng.random() for _ in range(n)], dtype=torch.float32),
        np.array(["w%d" % rng.randrange(50) for _ in range(n)] or [], dtype=object),
        torch.tensor([rng.random() > 0.5 for _ in range(n)], dtype=torch.bool),
        typ=slicetype.New(bs.int64, bs.float32, bs.string, bs.bool_))


def roundtrip(f, chunk):
    b = io.BytesIO()
    enc = sliceio.Encoder(b)
    for i in range(0, max(len(f), 1), chunk):
        enc.write(f.slice(i, min(len(f), i + chunk)))
    b.seek(0)
    return sliceio.read_all(sliceio.Decoder(b, f.typ), f.typ)


@pytest.mark.parametrize("chunk", [1, 33, 50, 100, 200])
def test_codec_roundtrip_chunks(chunk):
    f = mixed_frame(100)
    g = roundtrip(f, chunk)
    assert g.rows() == f.rows()


def test_codec_types():
    cols = [torch.tensor([1, -2], dtype=dt) for dt in
            (torch.int8, torch.int16, torch.int32, torch.int64, torch.uint8,
             torch.float16, torch.bfloat16, torch.float64)]
    cols.append(np.array([b"ab", b""], dtype=object))
    cols.append(np.array([{"a": 1}, None], dtype=object))
    cols.append(torch.arange(8, dtype=torch.float32).reshape(2, 4))
    cols.append(fr.col_from_list([[1], [2, 3]], slicetype.List(bs.int64)))
    f = Frame.of(*cols)
    g = roundtrip(f, 2)
    assert g.typ == f.typ
    for a, b in zip(f.cols(), g.cols()):
        assert fr.col_tolist(a) == fr.col_tolist(b)


def test_codec_empty_and_zero_values():
    f = mixed_frame(0)
    assert len(roundtrip(f, 1)) == 0
    z = Frame.make(slicetype.New(bs.int64, bs.string), 3)
    assert roundtrip(z, 3).rows() == [(0, "")] * 3


def test_codec_single_bit_corruption_is_detected():
    f = mixed_frame(20)
    b = io.BytesIO()
    sliceio.Encoder(b).write(f)
    data = bytearray(b.getvalue())
    rng = random.Random(7)
    integrity = 0
    for _ in range(60):
        pos = rng.randrange(len(data))
        bit = 1 << rng.randrange(8)
        data[pos] ^= bit
        with pytest.raises(Exception) as ei:
            sliceio.read_all(sliceio.Decoder(io.BytesIO(bytes(data)), f.typ), f.typ)
        if errors.is_kind(ei.value, errors.Kind.Integrity):
            integrity += 1
        data[pos] ^= bit
    assert integrity >= 55  # nearly all flips surface as Integrity errors
    # truncated stream
    with pytest.raises(errors.IntegrityError):
        sliceio.read_all(sliceio.Decoder(io.BytesIO(bytes(data[:-3])), f.typ), f.typ)


class JSONThing:
    def __init__(self, v):
        self.v = v

    def __eq__(self, o):
        return isinstance(o, JSONThing) and o.v == self.v


def test_custom_codec_with_session_state():
    key = fr.Key("count")
    calls = {"enc": 0, "dec": 0}

    def enc(session, vals):
        st = session.state(key, lambda: {"n": 0})
        st["n"] += 1
        calls["enc"] = st["n"]
        return json.dumps([v.v for v in vals]).encode()

    def dec(session, data, n):
        st = session.state(key, lambda: {"n": 0})
        st["n"] += 1
        calls["dec"] = st["n"]
        return [JSONThing(v) for v in json.loads(data)]

    fr.register_ops(JSONThing, fr.Ops(encode=enc, decode=dec))
    try:
        col = np.empty(4, dtype=object)
        col[:] = [JSONThing(i) for i in range(4)]
        f = Frame.of(col, torch.arange(4))
        g = roundtrip(f, 2)
        assert list(g.col(0)) == list(col)
        assert calls == {"enc": 2, "dec": 2}  # state persisted across batches
    finally:
        fr.unregister_ops(JSONThing)


def test_unencodable_object_is_fatal():
    col = np.empty(1, dtype=object)
    col[0] = lambda: 1
    with pytest.raises(errors.Error) as ei:
        sliceio.encode_frame(Frame.of(col))
    assert errors.is_fatal(ei.value)


def test_frame_and_multi_reader():
    f = mixed_frame(10)
    r = sliceio.FrameReader(f)
    a = r.read(4)
    b = r.read(None)
    assert len(a) == 4 and len(b) == 6 and r.read(1) is None
    closed = []
    rs = [sliceio.ReaderWithCloseFunc(sliceio.FrameReader(f.slice(i, i + 5)),
                                      lambda i=i: closed.append(i)) for i in (0, 5)]
    m = sliceio.MultiReader(rs)
    got = sliceio.read_all(m, f.typ)
    assert got.rows() == f.rows() and closed == [0, 5]


def test_multireader_sticky_error_and_err_reader():
    boom = RuntimeError("boom")
    m = sliceio.MultiReader([sliceio.FrameReader(mixed_frame(2)), sliceio.ErrReader(boom)])
    assert len(m.read()) == 2
    for _ in range(2):
        with pytest.raises(RuntimeError):
            m.read()
    assert sliceio.EmptyReader().read() is None


def test_read_full_and_closing_reader():
    f = mixed_frame(10)
    r = sliceio.FrameReader(f)
    assert len(sliceio.read_full(r, 7, f.typ)) == 7
    assert len(sliceio.read_full(r, 7, f.typ)) == 3
    assert sliceio.read_full(r, 7, f.typ) is None
    closed = []
    c = sliceio.ClosingReader(sliceio.ReaderWithCloseFunc(sliceio.FrameReader(f), lambda: closed.append(1)))
    while c.read(3) is not None:
        pass
    assert closed == [1]


def test_label_reader_sets_label():
    seen = []

    class R(sliceio.Reader):
        def read(self, n=None):
            seen.append(sliceio.current_label())
            return None

    sliceio.LabelReader(R(), "map@x.py:3").read()
    assert seen == ["map@x.py:3"] and sliceio.current_label() is None


def test_scanner_rows_and_vectors():
    f = mixed_frame(7)
    sc = sliceio.Scanner(sliceio.FrameReader(f), f.typ)
    rows = []
    while sc.scan():
        rows.append(sc.row())
    assert rows == f.rows() and sc.err() is None
    sc = sliceio.Scanner(sliceio.FrameReader(f), f.typ)
    cols = sc.scanv(4)
    assert len(cols) == 4 and len(cols[0]) == 4
    assert list(sliceio.Scanner(sliceio.FrameReader(f), f.typ)) == f.rows()
    sc = sliceio.Scanner(sliceio.FrameReader(f), f.typ)
    assert not sc.scan(nargs=2)
    assert isinstance(sc.err(), bs.typecheck.TypecheckError)
    sc = sliceio.Scanner(sliceio.ErrReader(ValueError("x")), f.typ)
    assert not sc.scan() and isinstance(sc.err(), ValueError)


def test_spiller():
    sp = sliceio.Spiller("t")
    frames = [mixed_frame(n, n) for n in (5, 0, 300)]
    for f in frames:
        assert sp.spill(f) > 0
    got = [sliceio.read_all(r, frames[0].typ) for r in sp.closing_readers()]
    assert [g.rows() for g in got] == [f.rows() for f in frames]
    import os
    d = sp.dir
    sp.cleanup()
    assert not os.path.exists(d)


def test_crc32_matches_zlib():
    import zlib
    data = bytes(range(256)) * 37
    assert sliceio.crc32(data) == zlib.crc32(data) & 0xFFFFFFFF
    assert sliceio.crc32(data[100:], sliceio.crc32(data[:100])) == zlib.crc32(data) & 0xFFFFFFFF


def test_object_columns_refuse_unknown_classes():
    """Object columns are restricted pickles: data types load, anything else (here a
    class whose unpickling would run code) is refused with an IntegrityError unless
    the application registered it."""
    import io
    import pickle
    import numpy as np
    from bigslice_b200 import errors, slicetype
    from bigslice_b200.frame import Frame
    from bigslice_b200 import sliceio
    from bigslice_b200.sliceio import codec

    class Evil:
        def __reduce__(self):
            return (print, ("pwned",))
    typ = slicetype.Type([slicetype.dtype_of(object)], 1)
    good = np.empty(3, dtype=object)
    good[:] = [{"a": 1}, (1, 2.5, "x"), [b"b", None]]
    data = codec.encode_frame(Frame([good], typ))
    back = sliceio.read_all(codec.Decoder(io.BytesIO(data), typ), typ)
    assert back.col(0).tolist() == good.tolist()
    bad = np.empty(1, dtype=object)
    bad[0] = Evil()
    data = codec.encode_frame(Frame([bad], typ))
    with pytest.raises(errors.IntegrityError):
        sliceio.read_all(codec.Decoder(io.BytesIO(data), typ), typ)
