"""Columnar frames: the unit of data every layer exchanges.

A :class:`Frame` is N equal-length typed columns plus a key *prefix*.  Columns
are torch tensors (CPU or B200 HBM), host object arrays or ragged list columns
(:mod:`.columns`).  Frames have Go-slice-like ``off/len/cap`` semantics so
``slice`` is zero-copy and ``grow``/``append_frame`` are amortised.

Capability parity: reference ``frame/frame.go:82-484`` — Make/Slices/Values/
Copy/AppendFrame/Slice/Grow/Ensure/Swap/Less/Hash/HashWithSeed/Zero and the
``sort.Interface`` — re-designed around whole-batch tensor ops: ``sorted()``
and ``hash()`` are vectorised (or sm_100a kernels for device frames) instead
of per-row calls.
"""
from __future__ import annotations

import io
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from .. import slicetype
from ..slicetype import DType, Type
from . import columns as C
from . import hashing
from .ops import ops_for_value


import os as _os
_OWN_SORT = bool(_os.environ.get("BIGSLICE_OWN_SORT"))


def _grow_cap(need: int, cap: int) -> int:
    """Growth policy: 2x below 1024 rows, 1.25x above (frame.go:458-484)."""
    newcap = cap
    doublecap = newcap + newcap
    if need > doublecap:
        return need
    if cap < 1024:
        newcap = max(doublecap, 1)
    else:
        while 0 < newcap < need:
            newcap += newcap // 4
        if newcap <= 0:
            newcap = need
    return max(newcap, need)


_NARROW_INTS = (torch.int32, torch.int16, torch.int8, torch.uint8)


class Frame:
    """N typed equal-length columns with a key prefix."""

    __slots__ = ("_data", "_off", "_len", "_cap", "typ")

    def __init__(self, data: Sequence, typ: Type, off: int = 0,
                 length: Optional[int] = None, cap: Optional[int] = None):
        self._data = list(data)
        self.typ = typ
        full = C.col_len(self._data[0]) if self._data else 0
        self._off = off
        self._cap = (full - off) if cap is None else cap
        self._len = self._cap if length is None else length

    # -- constructors --------------------------------------------------------
    @staticmethod
    def make(typ, length: int, cap: Optional[int] = None, device=None) -> "Frame":
        """Zeroed frame of ``length`` rows and capacity ``cap`` (frame.Make)."""
        typ = slicetype.as_type(typ)
        cap = length if cap is None else cap
        if length < 0 or length > cap:
            raise ValueError("frame.make: invalid len, cap")
        return Frame([C.col_zeros(t, cap, device) for t in typ.cols], typ, 0, length, cap)

    @staticmethod
    def of(*cols, prefix: int = 1, typ: Optional[Type] = None, device=None) -> "Frame":
        """Frame over user-provided columns (frame.Slices / frame.Values)."""
        if len(cols) == 1 and isinstance(cols[0], (list, tuple)) and cols[0] \
                and C.is_column(cols[0][0]):
            cols = tuple(cols[0])
        dts = typ.cols if typ is not None else [None] * len(cols)
        data = [C.as_column(c, dt, device) for c, dt in zip(cols, dts)]
        n = C.col_len(data[0]) if data else 0
        for c in data:
            if C.col_len(c) != n:
                raise ValueError("frame: columns have unequal lengths")
        if typ is None:
            typ = Type([slicetype.dtype_of_value(c) for c in data], prefix)
        return Frame(data, typ)

    slices = of

    @staticmethod
    def empty(typ, device=None) -> "Frame":
        return Frame.make(typ, 0, 0, device)

    # -- basic accessors -----------------------------------------------------
    def __len__(self) -> int:
        return self._len

    def len(self) -> int:
        return self._len

    def cap(self) -> int:
        return self._cap

    def num_out(self) -> int:
        return len(self._data)

    def out(self, i: int) -> DType:
        return self.typ.cols[i]

    def prefix(self) -> int:
        return self.typ.prefix()

    NumOut, Out, Prefix, Len, Cap = num_out, out, prefix, len, cap

    def prefixed(self, prefix: int) -> "Frame":
        return Frame(self._data, Type(self.typ.cols, prefix), self._off, self._len, self._cap)

    Prefixed = prefixed

    def is_zero(self) -> bool:
        return not self._data

    def col(self, i: int):
        """Column ``i`` as a view over rows [0, len)."""
        c = self._data[i]
        if isinstance(c, (C.ListColumn, C.StringColumn)):
            if self._off == 0 and self._len == len(c):
                return c
            return C.col_slice(c, self._off, self._off + self._len)
        return c[self._off:self._off + self._len]

    def cols(self) -> List[Any]:
        return [self.col(i) for i in range(len(self._data))]

    columns = cols

    @property
    def device(self) -> torch.device:
        for c in self._data:
            if isinstance(c, (torch.Tensor, C.ListColumn, C.StringColumn)):
                return C.col_device(c)
        return torch.device("cpu")

    def to(self, device) -> "Frame":
        if device is None:
            return self
        return Frame([C.col_to(c, device) for c in self.cols()], self.typ)

    def nbytes(self) -> int:
        return sum(C.col_nbytes(c) for c in self.cols())

    # -- slicing / growth ------------------------------------------------------
    def slice(self, i: int, j: int) -> "Frame":
        """Zero-copy rows [i, j); like Go, j may extend into capacity."""
        if i < 0 or j < i or j > self._cap:
            raise IndexError("frame.slice: slice index out of bounds")
        return Frame(self._data, self.typ, self._off + i, j - i, self._cap - i)

    Slice = slice

    def __getitem__(self, s):
        if isinstance(s, slice):
            i, j, st = s.indices(self._len)
            assert st == 1
            return self.slice(i, max(i, j))
        raise TypeError("frame index must be a slice")

    def grow(self, n: int) -> "Frame":
        """Frame with ``n`` more rows, reallocating per the growth policy."""
        if self._len + n <= self._cap:
            return Frame(self._data, self.typ, self._off, self._len + n, self._cap)
        newcap = _grow_cap(self._len + n, self._cap)
        dev = self.device
        data = []
        for c, t in zip(self.cols(), self.typ.cols):
            nc = C.col_zeros(t, newcap, dev)
            C.col_copy(nc, 0, c, 0, self._len)
            data.append(nc)
        return Frame(data, self.typ, 0, self._len + n, newcap)

    Grow = grow

    def ensure(self, n: int) -> "Frame":
        """Frame of exactly ``n`` rows, reusing capacity when possible."""
        if n <= self._cap:
            return Frame(self._data, self.typ, self._off, n, self._cap)
        return self.grow(n - self._len)

    Ensure = ensure

    def append_frame(self, g: "Frame") -> "Frame":
        """Append rows of ``g`` (frame.AppendFrame)."""
        if self.is_zero():
            return g.copy()
        if any(isinstance(c, C.ListColumn) for c in self._data):
            return concat([self, g])
        n0 = self._len
        f = self.grow(len(g))
        copy(f.slice(n0, n0 + len(g)), g)
        return f

    AppendFrame = append_frame

    def copy(self) -> "Frame":
        return Frame([C.col_clone(c) for c in self.cols()], self.typ)

    def take(self, idx) -> "Frame":
        """Gather rows by index tensor (permutation / selection)."""
        return Frame([C.col_take(c, idx) for c in self.cols()], self.typ)

    def mask(self, m: torch.Tensor) -> "Frame":
        idx = torch.nonzero(m, as_tuple=False).flatten()
        return self.take(idx)

    # -- element ops -----------------------------------------------------------
    def swap(self, i: int, j: int) -> None:
        for c in self.cols():
            if isinstance(c, torch.Tensor):
                tmp = c[i].clone()
                c[i] = c[j]
                c[j] = tmp
            elif isinstance(c, np.ndarray):
                c[i], c[j] = c[j], c[i]
            else:
                raise TypeError("cannot swap rows of a list column in place")

    Swap = swap

    def zero(self) -> None:
        """Zero rows [0, len) in place (frame.Zero)."""
        for i, c in enumerate(self._data):
            if isinstance(c, C.ListColumn):
                continue
            C.col_zero_(c, self._off, self._off + self._len)

    Zero = zero

    def row(self, i: int) -> tuple:
        out = []
        for c in self.cols():
            if isinstance(c, torch.Tensor):
                v = c[i]
                out.append(v.item() if v.dim() == 0 else v.tolist())
            elif isinstance(c, C.ListColumn):
                out.append(C.col_tolist(c.row(i)))
            else:
                out.append(c[i])
        return tuple(out)

    def rows(self) -> List[tuple]:
        lists = [C.col_tolist(c) for c in self.cols()]
        return list(zip(*lists)) if lists else []

    def less(self, i: int, j: int) -> bool:
        """Lexicographic row order over the prefix columns (frame.Less)."""
        for k in range(self.prefix()):
            c = self.col(k)
            a, b = _elem(c, i), _elem(c, j)
            if _lt(a, b):
                return True
            if _lt(b, a):
                return False
        return False

    Less = less

    def hash(self, seed: int = 0) -> torch.Tensor:
        """Per-row key hash: XOR of per-prefix-column murmur3 (frame.Hash).

        Returns an int64 tensor holding uint32 values on the frame's device.
        Device frames use the sm_100a kernel; this is the reference path.
        """
        h = None
        for k in range(self.prefix()):
            hk = hash_column(self.col(k), seed)
            h = hk if h is None else (h ^ hk.to(h.device))
        if h is None:
            h = torch.zeros(self._len, dtype=torch.int64)
        return h

    def hash_with_seed(self, seed: int) -> torch.Tensor:
        return self.hash(seed)

    Hash, HashWithSeed = hash, hash_with_seed

    def hash_row(self, i: int, seed: int = 0) -> int:
        return int(self.slice(i, i + 1).hash(seed)[0])

    # -- sorting ---------------------------------------------------------------
    def sort_indices(self) -> torch.Tensor:
        """Stable permutation ordering rows by the prefix columns."""
        n = self._len
        keys = [self.col(k) for k in range(self.prefix())]
        if len(keys) == 1 and isinstance(keys[0], torch.Tensor) and keys[0].dim() == 1:
            k0 = keys[0]
            # Device sorts go to the library (CUB onesweep behind torch.sort): the
            # in-tree LSD radix sort measured 0.53x of it on B200
            # (profiles/r2_kernels_vs_torch.jsonl) and is kept only as an opt-in
            # reference (BIGSLICE_OWN_SORT=1).
            if k0.is_cuda and k0.dtype == torch.int64 and _OWN_SORT:
                from ..ops import gpu
                if gpu.available():
                    return gpu.sort_pairs_i64(k0, None)[1]
            return torch.sort(k0, stable=True)[1]
        if len(keys) == 2 and all(isinstance(c, torch.Tensor) and c.dim() == 1 and c.dtype in _NARROW_INTS
                                  for c in keys) and keys[0].device == keys[1].device:
            # two integer columns of <= 32 bits: ONE sort of an order-preserving 64-bit word
            # (first column in the high half, second biased into the low half) instead of
            # one stable sort + gather per column
            word = (keys[0].to(torch.int64) << 32) | ((keys[1].to(torch.int64) + (1 << 31)) & 0xFFFFFFFF)
            return torch.sort(word, stable=True)[1]
        perm = None
        for c in reversed(keys):
            if isinstance(c, torch.Tensor):
                cc = c if perm is None else c.index_select(0, perm.to(c.device))
                if cc.dim() == 2:
                    p = torch.arange(n, device=cc.device)
                    for j in reversed(range(cc.shape[1])):
                        p = p[torch.sort(cc[p, j], stable=True)[1]]
                else:
                    p = torch.sort(cc, stable=True)[1]
            else:
                cc = c if perm is None else c[perm.cpu().numpy()]
                p = torch.from_numpy(_argsort_objects(cc))
            perm = p if perm is None else perm.to(p.device)[p]
        if perm is None:
            perm = torch.arange(n)
        return perm

    def sorted(self) -> "Frame":
        """Rows sorted by prefix columns (stable)."""
        if self._len <= 1:
            return self
        return self.take(self.sort_indices())

    def sort(self) -> None:
        """In-place sort (``sort.Sort(frame)`` in the reference)."""
        s = self.sorted()
        copy(self, s)

    # -- debug -----------------------------------------------------------------
    def write_tab(self, w, i: Optional[int] = None) -> None:
        rows = self.rows() if i is None else [self.row(i)]
        for r in rows:
            w.write("\t".join(str(v) for v in r) + "\n")

    WriteTab = write_tab

    def __repr__(self) -> str:
        b = io.StringIO()
        b.write("frame[%d]%s{" % (self._len, self.typ))
        rows = self.slice(0, min(self._len, 8)).rows()
        b.write("; ".join(",".join(str(v) for v in r) for r in rows))
        if self._len > 8:
            b.write("; ...")
        b.write("}")
        return b.getvalue()


def _elem(c, i):
    if isinstance(c, torch.Tensor):
        v = c[i]
        return v.item() if v.dim() == 0 else tuple(v.tolist())
    return c[i]


def _lt(a, b) -> bool:
    o = ops_for_value(a)
    if o is not None and o.less is not None:
        return bool(o.less(a, b))
    if isinstance(a, bool) and isinstance(b, bool):
        return (not a) and b
    return a < b


class _Key:
    __slots__ = ("v",)

    def __init__(self, v):
        self.v = v

    def __lt__(self, other):
        return _lt(self.v, other.v)


def _argsort_objects(c: np.ndarray) -> np.ndarray:
    if len(c) and ops_for_value(c[0]) is not None:
        idx = sorted(range(len(c)), key=lambda i: _Key(c[i]))
        return np.asarray(idx, dtype=np.int64)
    try:
        return np.argsort(c, kind="stable").astype(np.int64)
    except TypeError:
        idx = sorted(range(len(c)), key=lambda i: _Key(c[i]))
        return np.asarray(idx, dtype=np.int64)


def hash_column(c, seed: int = 0) -> torch.Tensor:
    """Per-element seeded murmur3 of one column -> int64 tensor of uint32."""
    if isinstance(c, C.StringColumn):
        from ..ops import strings
        return strings.hash32(c, seed)
    if isinstance(c, torch.Tensor):
        if c.is_cuda:
            from ..ops import gpu
            if gpu.available() and c.dim() == 1 and c.dtype in gpu.HASHABLE:
                return gpu.hash_column(c, seed)
        return hashing.hash_tensor(c, seed)
    if isinstance(c, np.ndarray):
        if len(c):
            o = ops_for_value(c[0])
            if o is not None and o.hash_with_seed is not None:
                return torch.tensor([o.hash_with_seed(v, seed) & 0xFFFFFFFF for v in c],
                                    dtype=torch.int64)
        return hashing.hash_objects(c, seed)
    raise TypeError("unhashable column type %r" % (type(c),))


def copy(dst: Frame, src: Frame) -> int:
    """Copy min(len) rows from src into dst; returns rows copied (frame.Copy)."""
    n = min(len(dst), len(src))
    for i in range(dst.num_out()):
        C.col_copy(dst._data[i], dst._off, src._data[i], src._off, n) \
            if not isinstance(src._data[i], C.ListColumn) else None
    return n


Copy = copy


def concat(frames: Sequence[Frame], typ: Optional[Type] = None, device=None) -> Frame:
    """Concatenate frames row-wise into one batch."""
    frames = [f for f in frames if f is not None and not f.is_zero()]
    if not frames:
        if typ is None:
            raise ValueError("concat of no frames needs a type")
        return Frame.empty(typ, device)
    if len(frames) == 1:
        return frames[0]
    typ = typ or frames[0].typ
    data = []
    for i, t in enumerate(typ.cols):
        data.append(C.col_concat([f.col(i) for f in frames], t))
    return Frame(data, typ)


def make(typ, length: int, cap: Optional[int] = None, device=None) -> Frame:
    return Frame.make(typ, length, cap, device)


def slices(*cols, prefix: int = 1) -> Frame:
    return Frame.of(*cols, prefix=prefix)


Make, Slices, Values = make, slices, slices


def compatible(f: Frame, g: Frame) -> bool:
    return f.typ.cols == g.typ.cols
