"""Reshuffle / Repartition / Reshard: pure shuffles.

* ``Reshuffle(s)``   — hash-partition rows by key prefix, same shard count.
* ``Repartition(s, fn)`` — partition with a user function
  ``fn(nshard, *cols) -> shard ids``.
* ``Reshard(s, n)``  — shuffle to a different shard count (identity if equal).

These are the framework's all-to-all: on the B200 executor the producer side
runs the fused hash→partition→exchange kernel that stores every row straight
into the destination GPU's receive buffer over NVLink (``csrc/exchange.cu``),
with NCCL all-to-all kept only as the measured baseline.

Capability parity: reference ``reshuffle.go:37-88``, ``reshard.go:15-45``.
"""
from __future__ import annotations

from typing import Callable, List

import numpy as np
import torch

from . import context as _context
from . import slicefunc, slicetype, typecheck
from .frame import Frame
from .frame import ops as frame_ops
from .sliceio import MultiReader, Reader
from .slices import Dep, Slice, _check_slice, _fn_name, make_name
from .slicetype import Type


def _can_partition(typ: Type) -> str:
    for i in range(typ.prefix()):
        if not frame_ops.can_hash(typ.out(i)):
            return "key column(%d) type %s cannot be hashed" % (i, typ.out(i))
    return ""


class _Reshuffle(Slice):
    def __init__(self, s: Slice, op: str, nshard: int, partitioner=None):
        super().__init__(make_name(op), s.typ, nshard,
                         [Dep(s, shuffle=True, partitioner=partitioner)], ())

    def reader(self, shard, deps: List[Reader]) -> Reader:
        return deps[0]


def Reshuffle(slice_: Slice) -> Slice:
    """Shuffle rows to shards by ``murmur3(key prefix) % nshard``."""
    _check_slice("reshuffle", slice_)
    why = _can_partition(slice_.typ)
    if why:
        typecheck.panic(why)
    return _Reshuffle(slice_, "reshuffle", slice_.num_shard())


def Repartition(slice_: Slice, partition: Callable) -> Slice:
    """Shuffle rows with a custom partitioner ``partition(nshard, *cols)``
    returning the destination shard of every row (vectorised; wrap with
    ``rowwise`` for a per-row function)."""
    _check_slice("repartition", slice_)
    if not callable(partition):
        typecheck.panicf("repartition: not a function: %s", type(partition).__name__)
    func = slicefunc.Func(partition)
    if not func.accepts(slice_.num_out() + 1):
        expect_arg = Type([slicetype.int64] + list(slice_.typ.cols))
        typecheck.panicf("repartition: expected %s, got %s",
                         slicetype.Signature(expect_arg, Type([slicetype.int64])),
                         _fn_name(partition))

    def partitioner(frame: Frame, nshard: int) -> torch.Tensor:
        ctx = _context.current()
        if func.rowwise:
            rows = frame.rows()
            ids = [int(func.call_row(ctx, (nshard,) + tuple(r))) for r in rows]
            out = torch.tensor(ids, dtype=torch.int64)
        else:
            (out,) = func.call(ctx, [nshard] + frame.cols())
            if not isinstance(out, torch.Tensor):
                out = torch.as_tensor(np.asarray(out))
            out = out.to(torch.int64)
        if len(out) and (int(out.min()) < 0 or int(out.max()) >= nshard):
            from . import errors
            raise errors.fatal("repartition: partition function returned shard outside [0, %d)" % nshard)
        return out

    return _Reshuffle(slice_, "repartition", slice_.num_shard(), partitioner)


def Reshard(slice_: Slice, nshard: int) -> Slice:
    """Shuffle into ``nshard`` shards (the slice itself if already so)."""
    _check_slice("reshard", slice_)
    why = _can_partition(slice_.typ)
    if why:
        typecheck.panic(why)
    if nshard < 1:
        typecheck.panic("reshard: shard must be >= 1")
    if slice_.num_shard() == nshard:
        return slice_
    return _Reshuffle(slice_, "reshard", nshard)
