"""Column storage helpers.

A frame column is one of

* ``torch.Tensor`` — numeric/bool columns (1-D) and fixed-width vector columns
  (2-D, rows first); lives on CPU or on a B200.
* ``numpy.ndarray(dtype=object)`` — host columns of ``str``/``bytes``/arbitrary
  python objects.
* :class:`ListColumn` — ragged list column (CSR: ``offsets`` + child values),
  produced by Cogroup.

These free functions dispatch on the storage type so user functions see plain
tensors/arrays.  Parity: the reference's ``reflect``-typed column slices
(``frame/frame.go:82-165``) and the unsafe copy helpers (``frame/unsafe.go``).
"""
from __future__ import annotations

from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from .. import slicetype
from ..slicetype import DType


class ListColumn:
    """Ragged list column: row ``i`` is ``values[offsets[i]:offsets[i+1]]``."""

    __slots__ = ("offsets", "values", "child_dtype")

    def __init__(self, offsets: torch.Tensor, values, child_dtype: DType):
        self.offsets = offsets
        self.values = values
        self.child_dtype = child_dtype

    def __len__(self):
        return int(self.offsets.shape[0]) - 1

    def row(self, i: int):
        a, b = int(self.offsets[i]), int(self.offsets[i + 1])
        return col_slice(self.values, a, b)

    def tolist(self):
        vals = col_tolist(self.values)
        offs = self.offsets.tolist()
        return [vals[offs[i]:offs[i + 1]] for i in range(len(self))]

    @property
    def device(self):
        return self.offsets.device

    def __repr__(self):
        return "ListColumn(%r)" % (self.tolist(),)


class StringColumn:
    """``n`` variable-length strings as CSR ``offsets[n+1]`` (int64) + ``data``
    (uint8), on CPU or in HBM: the device form of a ``string``/``bytes`` column
    (host frames may equally hold numpy object arrays).  The kernels that hash,
    tokenise and gather such columns are in :mod:`bigslice_b200.ops.strings`."""

    __slots__ = ("offsets", "data", "is_str")

    def __init__(self, offsets: torch.Tensor, data: torch.Tensor, is_str: bool = True):
        self.offsets, self.data, self.is_str = offsets, data, is_str

    def __len__(self) -> int:
        return int(self.offsets.shape[0]) - 1

    @property
    def device(self) -> torch.device:
        return self.offsets.device

    def to(self, device) -> "StringColumn":
        return StringColumn(self.offsets.to(device), self.data.to(device), self.is_str)

    def lengths(self) -> torch.Tensor:
        return self.offsets[1:] - self.offsets[:-1]

    def nbytes(self) -> int:
        return self.offsets.numel() * 8 + self.data.numel()

    def to_objects(self) -> np.ndarray:
        offs = self.offsets.cpu().numpy()
        raw = self.data.cpu().numpy().tobytes()
        out = np.empty(len(self), dtype=object)
        for i in range(len(self)):
            b = raw[offs[i]:offs[i + 1]]
            out[i] = b.decode("utf-8", errors="replace") if self.is_str else b
        return out

    def tolist(self) -> list:
        return self.to_objects().tolist()

    def __repr__(self):
        return "StringColumn(n=%d, bytes=%d, %s)" % (len(self), self.data.numel(), self.device)


def is_column(v) -> bool:
    return isinstance(v, (torch.Tensor, np.ndarray, ListColumn, StringColumn))


def col_len(c) -> int:
    if isinstance(c, (ListColumn, StringColumn)):
        return len(c)
    return int(c.shape[0])


def col_device(c) -> torch.device:
    if isinstance(c, torch.Tensor):
        return c.device
    if isinstance(c, (ListColumn, StringColumn)):
        return c.offsets.device
    return torch.device("cpu")


def col_slice(c, i: int, j: int):
    if isinstance(c, StringColumn):
        offs = c.offsets[i:j + 1]
        if offs.numel() == 0:
            return StringColumn(torch.zeros(1, dtype=torch.int64, device=c.device), c.data[:0], c.is_str)
        a, b = int(offs[0]), int(offs[-1])
        return StringColumn(offs - offs[0], c.data[a:b], c.is_str)
    if isinstance(c, ListColumn):
        offs = c.offsets[i:j + 1]
        if offs.numel() == 0:
            return ListColumn(torch.zeros(1, dtype=torch.int64, device=c.offsets.device),
                              col_slice(c.values, 0, 0), c.child_dtype)
        a, b = int(offs[0]), int(offs[-1])
        return ListColumn(offs - offs[0], col_slice(c.values, a, b), c.child_dtype)
    return c[i:j]


def col_take(c, idx):
    """Gather rows by index (``idx``: int64 tensor / ndarray / list)."""
    if isinstance(c, StringColumn):
        from ..ops import strings
        if not isinstance(idx, torch.Tensor):
            idx = torch.as_tensor(np.asarray(idx), dtype=torch.int64)
        return strings.take(c, idx)
    if isinstance(c, torch.Tensor):
        if not isinstance(idx, torch.Tensor):
            idx = torch.as_tensor(np.asarray(idx), dtype=torch.int64)
        idx = idx.to(c.device)
        if c.is_cuda and c.dim() >= 2 and c.dtype != torch.bool and idx.dtype == torch.int64 \
                and c.element_size() * (c.numel() // max(c.shape[0], 1)) >= 16:
            # wide rows: the vectorised gather kernel (24x torch.index_select on
            # 16-byte rows, profiles/r2_kernels_vs_torch.jsonl); scalar columns stay
            # with the library, which is ahead there
            from ..ops import gpu
            if gpu.available():
                return gpu.gather_rows(c, idx.contiguous())
        return c.index_select(0, idx)
    if isinstance(c, ListColumn):
        if not isinstance(idx, torch.Tensor):
            idx = torch.as_tensor(np.asarray(idx), dtype=torch.int64)
        idx = idx.to(c.offsets.device)
        starts = c.offsets[:-1].index_select(0, idx)
        lens = (c.offsets[1:] - c.offsets[:-1]).index_select(0, idx)
        new_off = torch.zeros(idx.numel() + 1, dtype=torch.int64, device=idx.device)
        torch.cumsum(lens, 0, out=new_off[1:])
        total = int(new_off[-1])
        if total == 0:
            return ListColumn(new_off, col_slice(c.values, 0, 0), c.child_dtype)
        row_of = _expand_rows(new_off, lens)
        within = torch.arange(total, device=idx.device) - new_off[:-1].index_select(0, row_of)
        src = starts.index_select(0, row_of) + within
        return ListColumn(new_off, col_take(c.values, src), c.child_dtype)
    if isinstance(idx, torch.Tensor):
        idx = idx.cpu().numpy()
    return c[np.asarray(idx, dtype=np.int64)]


def _expand_rows(offsets: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
    """Row id of every element of a ragged column given its offsets (the
    sm_100a expansion kernel on device, repeat_interleave on host)."""
    if offsets.is_cuda:
        from ..ops import gpu
        if gpu.available():
            return gpu.expand_indices(offsets)
    return torch.repeat_interleave(torch.arange(lens.numel(), device=lens.device), lens)


def col_zeros(dt: DType, n: int, device=None):
    """A zero-valued column of ``n`` rows of dtype ``dt``."""
    dt = slicetype.dtype_of(dt)
    if dt.kind == "list":
        return ListColumn(torch.zeros(n + 1, dtype=torch.int64, device=device),
                          col_zeros(dt.child, 0, device), dt.child)
    if dt.kind == "vec":
        return torch.zeros((n, dt.width), dtype=dt.torch_dtype, device=device)
    if dt.is_tensor:
        return torch.zeros(n, dtype=dt.torch_dtype, device=device)
    a = np.empty(n, dtype=object)
    if dt.kind == "str":
        a[:] = ""
    elif dt.kind == "bytes":
        a[:] = b""
    return a


def col_empty(dt: DType, n: int, device=None):
    dt = slicetype.dtype_of(dt)
    if dt.kind == "vec":
        return torch.empty((n, dt.width), dtype=dt.torch_dtype, device=device)
    if dt.is_tensor:
        return torch.empty(n, dtype=dt.torch_dtype, device=device)
    return col_zeros(dt, n, device)


def col_zero_(c, i: int = 0, j: Optional[int] = None):
    """Zero rows [i, j) in place."""
    j = col_len(c) if j is None else j
    if isinstance(c, torch.Tensor):
        c[i:j].zero_()
    elif isinstance(c, np.ndarray):
        fill: Any = None
        for e in c[i:j]:
            if isinstance(e, str):
                fill = ""
            elif isinstance(e, (bytes, bytearray)):
                fill = b""
            break
        c[i:j] = fill
    else:
        raise TypeError("cannot zero list column in place")


def col_concat(cols: Sequence, dt: Optional[DType] = None, device=None):
    cols = [c for c in cols]
    if not cols:
        return col_zeros(dt, 0, device)
    first = cols[0]
    if isinstance(first, torch.Tensor):
        if len(cols) == 1:
            return first
        return torch.cat(list(cols), 0)
    if isinstance(first, StringColumn) or any(isinstance(c, StringColumn) for c in cols):
        if len(cols) == 1:
            return first
        from ..ops import strings
        dev = next(c.device for c in cols if isinstance(c, StringColumn))
        cs = [c if isinstance(c, StringColumn) else strings.from_objects(c, device=dev) for c in cols]
        offs, base = [cs[0].offsets], cs[0].offsets[-1]
        for c in cs[1:]:
            offs.append(c.offsets[1:] + base)
            base = base + c.offsets[-1]
        return StringColumn(torch.cat(offs), torch.cat([c.data for c in cs]), cs[0].is_str)
    if isinstance(first, ListColumn):
        if len(cols) == 1:
            return first
        offs = [first.offsets]
        base = first.offsets[-1]
        for c in cols[1:]:
            offs.append(c.offsets[1:] + base)
            base = base + c.offsets[-1]
        return ListColumn(torch.cat(offs), col_concat([c.values for c in cols]),
                          first.child_dtype)
    if len(cols) == 1:
        return first
    return np.concatenate(list(cols))


def col_to(c, device):
    if device is None:
        return c
    if isinstance(c, torch.Tensor):
        dev = torch.device(device)
        if c.device == dev:
            return c
        # Only host->device copies may be asynchronous: the consumer is a kernel
        # ordered after the copy on the same stream.  A non-blocking device->host
        # copy lands in freshly pinned memory that the host would read before the
        # copy has finished (observed: stale tails of sorted frames).
        return c.to(dev, non_blocking=(dev.type == "cuda"))
    if isinstance(c, ListColumn):
        return ListColumn(c.offsets.to(device), col_to(c.values, device), c.child_dtype)
    if isinstance(c, StringColumn):
        # strings live as object arrays on the host side of the framework (user
        # functions, the scanner) and as CSR on devices
        dev = torch.device(device)
        return c.to_objects() if dev.type == "cpu" else c.to(dev)
    return c


def col_clone(c):
    if isinstance(c, torch.Tensor):
        return c.clone()
    if isinstance(c, ListColumn):
        return ListColumn(c.offsets.clone(), col_clone(c.values), c.child_dtype)
    if isinstance(c, StringColumn):
        return StringColumn(c.offsets.clone(), c.data.clone(), c.is_str)
    return c.copy()


def col_copy(dst, doff: int, src, soff: int, n: int):
    """dst[doff:doff+n] = src[soff:soff+n] (memmove semantics)."""
    if n <= 0:
        return
    if isinstance(dst, torch.Tensor):
        if not isinstance(src, torch.Tensor):
            src = torch.as_tensor(np.asarray(src))
        dst[doff:doff + n].copy_(src[soff:soff + n])
    elif isinstance(dst, np.ndarray):
        dst[doff:doff + n] = src[soff:soff + n]
    else:
        raise TypeError("cannot copy into list column in place")


def col_tolist(c) -> list:
    if isinstance(c, torch.Tensor):
        return c.tolist()
    if isinstance(c, (ListColumn, StringColumn)):
        return c.tolist()
    if isinstance(c, np.ndarray):
        return c.tolist()
    return list(c)


def col_from_list(vals: Sequence, dt: DType, device=None):
    """Build a column of dtype ``dt`` from python values."""
    dt = slicetype.dtype_of(dt)
    if dt.kind == "list":
        offs = [0]
        flat: List = []
        for v in vals:
            v = [] if v is None else v
            flat.extend(col_tolist(v) if is_column(v) else list(v))
            offs.append(len(flat))
        return ListColumn(torch.tensor(offs, dtype=torch.int64, device=device),
                          col_from_list(flat, dt.child, device), dt.child)
    if dt.is_tensor:
        if isinstance(vals, torch.Tensor):
            t = vals.to(dt.torch_dtype)
        elif len(vals) and isinstance(vals[0], torch.Tensor):
            t = torch.stack([v.to(dt.torch_dtype) for v in vals])
        else:
            if dt.kind == "uint" and dt.itemsize > 1:
                t = torch.from_numpy(np.asarray(vals, dtype=dt.name).reshape(
                    (len(vals), dt.width) if dt.kind == "vec" else (len(vals),)))
            else:
                t = torch.tensor(list(vals), dtype=dt.torch_dtype)
        if dt.kind == "vec" and t.dim() == 1:
            t = t.reshape(-1, dt.width)
        return t.to(device) if device is not None else t
    a = np.empty(len(vals), dtype=object)
    for i, v in enumerate(vals):
        a[i] = v
    return a


def as_column(v, dt: Optional[DType] = None, device=None):
    """Coerce a user-supplied value (tensor/ndarray/list) into a column."""
    if isinstance(v, ListColumn):
        return col_to(v, device)
    if isinstance(v, StringColumn):
        return v if device is None else col_to(v, device)
    if isinstance(v, torch.Tensor):
        if dt is not None and dt.is_tensor and v.dtype != dt.torch_dtype:
            v = v.to(dt.torch_dtype)
        return col_to(v, device)
    if isinstance(v, np.ndarray):
        if v.dtype == object:
            return v
        if v.dtype.kind in ("U", "S"):
            a = np.empty(len(v), dtype=object)
            a[:] = v.tolist()
            return a
        t = torch.from_numpy(np.ascontiguousarray(v))
        if dt is not None and dt.is_tensor and t.dtype != dt.torch_dtype:
            t = t.to(dt.torch_dtype)
        return col_to(t, device)
    if isinstance(v, (list, tuple)):
        if dt is None:
            dt = slicetype.dtype_of_value(v)
        return col_from_list(list(v), dt, device)
    raise TypeError("not a column: %r" % (type(v),))


def col_nbytes(c) -> int:
    if isinstance(c, torch.Tensor):
        return c.numel() * c.element_size()
    if isinstance(c, ListColumn):
        return col_nbytes(c.offsets) + col_nbytes(c.values)
    if isinstance(c, StringColumn):
        return c.nbytes()
    n = 0
    for e in c:
        n += len(e) if isinstance(e, (str, bytes, bytearray)) else 8
    return n
