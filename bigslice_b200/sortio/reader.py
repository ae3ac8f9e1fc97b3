"""Merge-and-combine of key-sorted streams.

``reduce_reader`` merges N sorted streams and folds rows with equal keys using
the combiner, emitting one row per key.  Unlike the reference (which assumes
keys are unique within each stream and pops one row per heap op,
``sortio/reader.go:36-130``) this version tolerates duplicate keys inside a
stream: the trailing key group of every merged batch is carried over until the
next batch proves the key complete, and each batch is combined with one
segmented reduction.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from ..frame import Frame, concat
from ..frame import columns as C
from ..ops.segmented import group_starts, segmented_reduce
from ..sliceio import Reader
from ..slicetype import Type
from .sort import merge_reader


def combine_sorted(f: Frame, func, ctx=None) -> Frame:
    """Collapse equal-key runs of a sorted frame with ``func`` (last column)."""
    if len(f) <= 1:
        return f
    nk = f.prefix()
    starts = group_starts(f, nk)
    if bool(starts.all()):
        return f
    vals = f.col(f.num_out() - 1)
    red = segmented_reduce(vals, starts, func.fn, func.rowwise, ctx, func)
    idx = torch.nonzero(starts).flatten()
    cols = [C.col_take(f.col(k), idx) for k in range(f.num_out() - 1)]
    if isinstance(red, torch.Tensor) and isinstance(vals, torch.Tensor):
        red = red.to(vals.dtype)
    cols.append(red)
    return Frame(cols, f.typ)


class ReduceReader(Reader):
    def __init__(self, typ: Type, readers: Sequence[Reader], func, ctx=None):
        self.typ = typ
        self.func = func
        self.ctx = ctx
        self.merged = merge_reader(typ, readers)
        self.carry: Optional[Frame] = None
        self.done = False

    def read(self, n=None):
        while not self.done:
            f = self.merged.read(None)
            if f is None:
                self.done = True
                out, self.carry = self.carry, None
                if out is not None and len(out):
                    return combine_sorted(out, self.func, self.ctx)
                return None
            if self.carry is not None and len(self.carry):
                f = concat([self.carry, f], self.typ)
                self.carry = None
            if len(f) == 0:
                continue
            f = combine_sorted(f, self.func, self.ctx)
            # hold back the last key group: more rows for it may follow
            if len(f) == 1:
                self.carry = f
                continue
            self.carry = f.slice(len(f) - 1, len(f))
            return f.slice(0, len(f) - 1)
        return None

    def close(self):
        self.merged.close()


def reduce_reader(typ: Type, readers: Sequence[Reader], func, ctx=None) -> Reader:
    """``sortio.Reduce``: merge sorted readers, combining equal keys."""
    return ReduceReader(typ, list(readers), func, ctx)


Reduce = reduce_reader
