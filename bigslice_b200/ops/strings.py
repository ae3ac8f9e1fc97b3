"""Device string columns: CSR ``(offsets, bytes)`` + the kernels of ``csrc/strings.cu``.

A :class:`StringColumn` holds ``n`` variable-length byte strings as
``offsets[n+1]`` (int64) and ``data`` (uint8) on any device.  It is the device
form of a ``string``/``bytes`` frame column (host frames keep numpy object
arrays); :func:`from_objects` / :meth:`StringColumn.to_objects` convert, packing
with the native C++ codec helpers.

What runs on the GPU (one kernel each, no per-row host work):

* :func:`hash32` — seeded murmur3 of every string, bit-exact with the host
  hashing and therefore with the reference's shard placement of string keys
  (``frame/ops_builtin.go:13-20``);
* :func:`key64` — a 64-bit dictionary id per string (two differently seeded
  murmur3 hashes): string keys ride the int64 hash-aggregation / fused-exchange
  path as ids, and :func:`dictionary` keeps one representative string per id;
  :func:`verify_dictionary` proves that equal ids are equal strings (a collision
  would otherwise merge two keys silently);
* :func:`split` — tokeniser: every row is split on delimiter bytes into a new
  string column plus each token's parent row (the Flatmap of the tutorial
  wordcount, ``docs/index.md:164-278``);
* :func:`url_domain` — the host part of every URL, lower-cased (``cmd/urls``);
* :func:`take` — gather rows.

CPU tensors run the same operations through numpy/torch (tests, small data).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

c_void_p, c_int, c_ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong

WHITESPACE = b" \t\r\n\x0b\x0c"


from ..frame.columns import StringColumn  # noqa: E402  (the storage class lives with the other column types)


def from_objects(values: Sequence, device=None, is_str: Optional[bool] = None) -> StringColumn:
    """Pack python strings/bytes (list or object array) into a StringColumn."""
    vals = values.tolist() if isinstance(values, np.ndarray) else list(values)
    if is_str is None:
        is_str = not (vals and isinstance(vals[0], (bytes, bytearray)))
    from ..sliceio.codec import _encode_varlen
    blob = _encode_varlen(vals, is_str)
    n = len(vals)
    offs = torch.from_numpy(np.frombuffer(blob, dtype="<i8", count=n + 1).copy())
    data = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8, offset=(n + 1) * 8).copy())
    col = StringColumn(offs, data, is_str)
    return col.to(device) if device is not None else col


def _lib():
    from . import gpu
    return gpu.require()


def _stream():
    from . import gpu
    return gpu._stream()


def _p(t: torch.Tensor) -> c_void_p:
    return c_void_p(t.data_ptr())


def _chk(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError("bigslice_b200 string kernel %s failed: cuda error %d" % (what, rc))


def _cuda(col: StringColumn) -> bool:
    return col.offsets.is_cuda


def hash32(col: StringColumn, seed: int = 0) -> torch.Tensor:
    """Seeded murmur3 of every string: int64 tensor of uint32 values."""
    n = len(col)
    if _cuda(col):
        out = torch.empty(n, dtype=torch.int64, device=col.device)
        _chk(_lib().bs_str_hash(_p(col.offsets), _p(col.data), c_ll(n), ctypes.c_uint(seed & 0xFFFFFFFF),
                                _p(out), _stream()), "hash")
        return out
    from ..frame import hashing
    return hashing.hash_objects(col.to_objects() if col.is_str else
                                np.array(col.tolist(), dtype=object), seed)


def key64(col: StringColumn) -> torch.Tensor:
    """64-bit dictionary id of every string (int64; never the table's reserved key)."""
    n = len(col)
    if _cuda(col):
        out = torch.empty(n, dtype=torch.int64, device=col.device)
        _chk(_lib().bs_str_key64(_p(col.offsets), _p(col.data), c_ll(n), _p(out), _stream()), "key64")
        return out
    lo, hi = hash32(col, 0), hash32(col, 0x9747b28c)
    k = lo | (hi << 32)
    k = torch.where(k >= (1 << 63), k - (1 << 64), k) if k.dtype != torch.int64 else k
    k = k.to(torch.int64)
    return torch.where(k == -(1 << 63), torch.full_like(k, (1 << 63) - 1), k)


def _spans_to_column(src_data: torch.Tensor, start: torch.Tensor, length: torch.Tensor, is_str: bool
                     ) -> StringColumn:
    """Materialise spans ``src_data[start[i] : start[i]+length[i]]`` as a new column."""
    n = start.numel()
    dev = src_data.device
    offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(length, 0, out=offs[1:])
    total = int(offs[-1]) if n else 0
    out = torch.empty(total, dtype=torch.uint8, device=dev)
    if total:
        if src_data.is_cuda:
            _chk(_lib().bs_str_copy_spans(_p(src_data), _p(start.contiguous()), _p(offs), c_ll(n), _p(out),
                                          _stream()), "copy_spans")
        else:
            row = torch.repeat_interleave(torch.arange(n), length)
            pos = torch.arange(total) - offs[:-1][row] + start[row]
            out = src_data[pos]
    return StringColumn(offs, out, is_str)


def take(col: StringColumn, idx: torch.Tensor) -> StringColumn:
    idx = idx.to(device=col.device, dtype=torch.int64)
    return _spans_to_column(col.data, col.offsets[:-1][idx].contiguous(), col.lengths()[idx].contiguous(),
                            col.is_str)


def split(col: StringColumn, delims: bytes = WHITESPACE) -> Tuple[StringColumn, torch.Tensor]:
    """Split every row on ``delims``: ``(tokens, parent)`` where ``parent[t]`` is the
    row token ``t`` came from (tokens appear in row order, left to right)."""
    n = len(col)
    dev = col.device
    if _cuda(col):
        lib = _lib()
        d = torch.tensor(list(delims), dtype=torch.uint8, device=dev)
        counts = torch.empty(n, dtype=torch.int64, device=dev)
        _chk(lib.bs_str_split_count(_p(col.offsets), _p(col.data), c_ll(n), _p(d), c_int(len(delims)), _p(counts),
                                    _stream()), "split_count")
        base = torch.cumsum(counts, 0) - counts
        ntok = int(counts.sum()) if n else 0
        start = torch.empty(ntok, dtype=torch.int64, device=dev)
        length = torch.empty(ntok, dtype=torch.int64, device=dev)
        parent = torch.empty(ntok, dtype=torch.int64, device=dev)
        if ntok:
            _chk(lib.bs_str_split_fill(_p(col.offsets), _p(col.data), c_ll(n), _p(d), c_int(len(delims)), _p(base),
                                       _p(start), _p(length), _p(parent), _stream()), "split_fill")
        return _spans_to_column(col.data, start, length, col.is_str), parent
    # host: plain python split with the same delimiter semantics
    raw = col.data.numpy().tobytes()
    offs = col.offsets.numpy()
    dset = set(delims)
    toks, parent = [], []
    for i in range(n):
        cur = bytearray()
        for b in raw[offs[i]:offs[i + 1]]:
            if b in dset:
                if cur:
                    toks.append(bytes(cur))
                    parent.append(i)
                    cur = bytearray()
            else:
                cur.append(b)
        if cur:
            toks.append(bytes(cur))
            parent.append(i)
    out = from_objects(toks, is_str=False)
    out.is_str = col.is_str
    return out, torch.tensor(parent, dtype=torch.int64)


def url_domain(col: StringColumn) -> StringColumn:
    """The host part of every URL (``scheme://[user@]host[:port]/...`` or
    ``host/...``), lower-cased; an empty string where there is none."""
    n = len(col)
    if _cuda(col):
        data = col.data.clone()  # lower-casing happens in place
        start = torch.empty(n, dtype=torch.int64, device=col.device)
        length = torch.empty(n, dtype=torch.int64, device=col.device)
        _chk(_lib().bs_str_domain(_p(col.offsets), _p(data), c_ll(n), _p(start), _p(length), _stream()), "domain")
        return _spans_to_column(data, start, length, col.is_str)
    out = []
    for s in col.to_objects().tolist():
        s = s if isinstance(s, str) else s.decode("utf-8", "replace")
        i = s.find("://")
        rest = s[i + 3:] if i >= 0 else s
        for sep in "/?#":
            j = rest.find(sep)
            if j >= 0:
                rest = rest[:j]
        if "@" in rest:
            rest = rest.rsplit("@", 1)[1]
        out.append(rest.lower())
    return from_objects(out, is_str=col.is_str)


def dictionary(col: StringColumn, ids: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, StringColumn]:
    """``(distinct ids ascending, one representative string per id)``."""
    ids = key64(col) if ids is None else ids
    srt, order = torch.sort(ids, stable=True)
    first = torch.ones(srt.numel(), dtype=torch.bool, device=srt.device)
    if srt.numel() > 1:
        first[1:] = srt[1:] != srt[:-1]
    rep = order[first]
    return srt[first], take(col, rep)


def verify_dictionary(col: StringColumn, ids: torch.Tensor, dict_ids: torch.Tensor, dict_strs: StringColumn
                      ) -> bool:
    """True iff every row's string equals the dictionary entry of its id (no two
    different strings share an id)."""
    n = len(col)
    if n == 0:
        return True
    slot = torch.searchsorted(dict_ids, ids)
    if _cuda(col):
        eq = torch.empty(n, dtype=torch.uint8, device=col.device)
        a_start, a_len = col.offsets[:-1].contiguous(), col.lengths().contiguous()
        b_start = dict_strs.offsets[:-1][slot].contiguous()
        b_len = dict_strs.lengths()[slot].contiguous()
        _chk(_lib().bs_str_equal_spans(_p(col.data), _p(a_start), _p(a_len), _p(dict_strs.data), _p(b_start),
                                       _p(b_len), c_ll(n), _p(eq), _stream()), "equal_spans")
        return bool(eq.all())
    a, b = col.to_objects(), dict_strs.to_objects()
    return all(a[i] == b[int(slot[i])] for i in range(n))
