"""Keyed accumulators: per-key state for ordered folds on the host path.

``Fold`` evaluates with sort + segmented fold (``ops/segmented.py``); these
accumulators provide the equivalent incremental map-based form used when rows
arrive in batches that cannot be materialised at once, and are part of the
public surface for custom slices.

Capability parity: reference ``accum.go:20-186`` (``Accumulator`` with
``Accumulate(in, fn)`` / ``Read(keys, vals)``, specialised for string, int and
int64 keys; any hashable python key works here).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Tuple

from . import slicetype
from .frame import Frame
from .frame import columns as C
from .slicetype import DType


def can_make_accumulator_for_key(dt: DType) -> bool:
    """Keys must be hashable scalars (``canMakeAccumulatorForKey``)."""
    return dt.kind in ("int", "uint", "str", "bytes", "bool", "float")


class Accumulator:
    """Accumulates values per key; values start at the zero of ``acc_type``."""

    def __init__(self, key_type: DType, acc_type: DType):
        self.key_type = slicetype.dtype_of(key_type)
        self.acc_type = slicetype.dtype_of(acc_type)
        self.state: Dict[Any, Any] = {}
        self._order: List[Any] = []

    def _zero(self):
        return C.col_tolist(C.col_zeros(self.acc_type, 1))[0]

    def accumulate(self, frame: Frame, fn: Callable, ctx=None) -> None:
        """Fold every row of ``frame`` (key in column 0) with ``fn(acc, *vals)``."""
        keys = C.col_tolist(frame.col(0))
        vals = [C.col_tolist(frame.col(i)) for i in range(1, frame.num_out())]
        for i, k in enumerate(keys):
            if k not in self.state:
                self.state[k] = self._zero()
                self._order.append(k)
            self.state[k] = fn(self.state[k], *[v[i] for v in vals])

    def read(self, n: Optional[int] = None) -> Optional[Tuple[list, list]]:
        """Drain up to ``n`` accumulated ``(keys, values)``; None when empty."""
        if not self._order:
            return None
        take = self._order if n is None else self._order[:n]
        self._order = [] if n is None else self._order[n:]
        ks = list(take)
        vs = [self.state.pop(k) for k in ks]
        return ks, vs

    def __len__(self):
        return len(self._order)

    Accumulate, Read = accumulate, read


def make_accumulator(key_type, acc_type) -> Accumulator:
    kt = slicetype.dtype_of(key_type)
    if not can_make_accumulator_for_key(kt):
        raise TypeError("key type %s cannot be accumulated" % kt)
    return Accumulator(kt, acc_type)
