"""Map-side combiners: per-key pre-aggregation in front of a shuffle.

:class:`CombiningFrame` keeps one row per distinct key and folds new batches
into it; :class:`Combiner` adds spilling: when the number of live keys reaches
``target_size`` the table is sorted and written out as a sorted run, and
``reader()`` merge-reduces all runs plus the final table.

The reference implements this as an open-addressing hash table built on a
frame with one reflective combine per probe (``exec/combiner.go:36-357``).
Here a batch is combined with whole-batch operations: on B200 frames with a
builtin functor use the sm_100a hash-aggregation kernel
(``csrc/hash_combine.cu``: shared-memory hot-key table + HBM open-addressing
table with atomic combine); otherwise sort + segmented reduce.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .. import stats
from ..frame import Frame, concat
from ..internal import defaultsize
from ..sliceio import EmptyReader, FrameReader, Reader, Spiller
from ..slicetype import Type
from ..sortio import combine_sorted, reduce_reader

# expvar-like global counters (reference exec/combiner.go:24-29)
combiner_stats = stats.Map()
_keys = combiner_stats.int("combinerKeys")
_records = combiner_stats.int("combinerRecords")
_spills = combiner_stats.int("combinerSpills")


def _gpu_combinable(f: Frame, func) -> bool:
    if f.device.type != "cuda" or func.builtin is None:
        return False
    from ..ops import gpu
    return gpu.available() and gpu.can_hash_combine(f, func.builtin)


def combine_frame(f: Frame, func, ctx=None) -> Frame:
    """One row per key (unordered on the GPU hash path, sorted otherwise)."""
    if len(f) <= 1:
        return f
    if _gpu_combinable(f, func):
        from ..ops import gpu
        return gpu.hash_combine_frame(f, func.builtin)
    g = _host_hash_combine(f, func)
    if g is not None:
        return g
    return combine_sorted(f.sorted(), func, ctx)


def _host_group_combine(f: Frame, k, v, op: str) -> Optional[Frame]:
    """str / bytes keys: C++ grouping (``host_group_objects``: one hash-map pass
    over UTF-8 views, no Python loop), a torch scatter for the values, and a sort
    of the distinct keys only."""
    if not (isinstance(v, torch.Tensor) and v.device.type == "cpu" and v.dim() == 1
            and v.dtype in _HOST_VTYPE):
        return None
    from .. import _native
    nat = _native.lib()
    if nat is None or not hasattr(nat, "host_group_objects"):
        return None
    try:
        firsts, group = nat.host_group_objects(k.tolist())
    except TypeError:
        return None  # mixed / exotic key objects: the generic sorted path handles them
    gid = torch.from_numpy(group)
    m = len(firsts)
    if op == "add":
        out = torch.zeros(m, dtype=v.dtype).index_add_(0, gid, v)
    else:
        out = torch.empty(m, dtype=v.dtype).scatter_reduce_(0, gid, v, "amin" if op == "min" else "amax",
                                                            include_self=False)
    keys = k[firsts]
    g = Frame([keys, out], f.typ)
    return g.sorted()


_HOST_VTYPE = {torch.int32: 0, torch.int64: 1, torch.float32: 2, torch.float64: 3}
_HOST_OP = {"add": 0, "min": 1, "max": 2}


def _host_hash_combine(f: Frame, func) -> Optional[Frame]:
    """CPU frames of <int64 key, numeric value> with a builtin functor: C++
    open-addressing aggregation with the GIL released (``csrc/native.cpp:
    host_hash_combine``), then a sort of the *distinct* keys only — instead of
    sorting every row.  Returns None when it does not apply."""
    op = getattr(func, "builtin", None)
    if op not in _HOST_OP or f.num_out() != 2 or f.prefix() != 1 or len(f) < 2048:
        return None
    k, v = f.col(0), f.col(1)
    if isinstance(k, np.ndarray) and k.dtype == object:
        return _host_group_combine(f, k, v, op)
    if not (isinstance(k, torch.Tensor) and isinstance(v, torch.Tensor) and k.device.type == "cpu"
            and k.dtype == torch.int64 and k.dim() == 1 and v.dim() == 1 and v.dtype in _HOST_VTYPE):
        return None
    from .. import _native
    nat = _native.lib()
    if nat is None or not hasattr(nat, "host_hash_combine"):
        return None
    k, v = k.contiguous(), v.contiguous()
    ok, ov = torch.empty_like(k), torch.empty_like(v)
    m = nat.host_hash_combine(k.data_ptr(), v.data_ptr(), k.numel(), _HOST_VTYPE[v.dtype], _HOST_OP[op],
                              ok.data_ptr(), ov.data_ptr())
    if m < 0:
        return None
    ok, ov = ok[:m], ov[:m]
    order = torch.sort(ok, stable=True)[1]
    return Frame([ok[order], ov[order]], f.typ)


class CombiningFrame:
    """Distinct-key table with batch combine (``combiningFrame``)."""

    def __init__(self, typ: Type, func, ctx=None):
        self.typ = typ
        self.func = func
        self.ctx = ctx
        self.table: Optional[Frame] = None
        self.pending: List[Frame] = []
        self.pending_rows = 0

    def combine(self, f: Frame) -> None:
        """Fold the rows of ``f`` into the table."""
        if len(f) == 0:
            return
        _records.add(len(f))
        self.pending.append(f)
        self.pending_rows += len(f)
        # Amortise: compact when the pending rows rival the table size.
        if self.pending_rows >= max(defaultsize.Chunk, self.len_table()):
            self.compact()

    Combine = combine

    def len_table(self) -> int:
        return len(self.table) if self.table is not None else 0

    def compact(self) -> Frame:
        """Fold pending rows; returns the (unsorted-on-GPU) distinct table."""
        if self.pending:
            parts = ([self.table] if self.table is not None else []) + self.pending
            self.pending, self.pending_rows = [], 0
            self.table = combine_frame(concat(parts, self.typ), self.func, self.ctx)
        if self.table is None:
            self.table = Frame.empty(self.typ)
        return self.table

    Compact = compact

    def __len__(self):
        return len(self.compact())

    def sorted(self) -> Frame:
        return self.compact().sorted()

    def reset(self) -> None:
        self.table, self.pending, self.pending_rows = None, [], 0


class Combiner:
    """Spilling combiner (``exec.combiner``)."""

    def __init__(self, typ: Type, func, name: str = "combiner",
                 target_size: Optional[int] = None, ctx=None):
        self.typ = typ
        self.func = func
        self.name = name
        self.ctx = ctx
        self.target_size = target_size if target_size is not None else defaultsize.Chunk * 100
        self.frame = CombiningFrame(typ, func, ctx)
        self.spiller: Optional[Spiller] = None
        self.total = 0

    def combine(self, f: Frame) -> None:
        self.frame.combine(f)
        if self.frame.pending_rows + self.frame.len_table() >= self.target_size:
            if len(self.frame) >= self.target_size:
                self.spill()

    Combine = combine

    def spill(self) -> None:
        """Write the current table as a sorted run and start afresh."""
        run = self.frame.sorted()
        if len(run) == 0:
            return
        if self.spiller is None:
            self.spiller = Spiller(self.name)
        self.spiller.spill(run)
        _spills.add(1)
        self.total += len(run)
        self.frame.reset()

    def reader(self, device=None, sort: bool = True) -> Reader:
        """Stream of the combined rows: one row per key, sorted by key."""
        if self.spiller is None:
            tbl = self.frame.sorted() if sort else self.frame.compact()
            _keys.add(len(tbl))
            return FrameReader(tbl) if len(tbl) else EmptyReader()
        final = self.frame.sorted()
        readers = self.spiller.closing_readers(device)
        if len(final):
            readers.append(FrameReader(final))
        spiller = self.spiller
        r = reduce_reader(self.typ, readers, self.func, self.ctx)
        from ..sliceio import ReaderWithCloseFunc
        return ReaderWithCloseFunc(r, spiller.cleanup)

    Reader = reader

    def discard(self) -> None:
        if self.spiller is not None:
            self.spiller.cleanup()
            self.spiller = None
        self.frame.reset()

    Discard = discard
