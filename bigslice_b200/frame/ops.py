"""Per-element-type operations registry.

Built-in dtypes (ints, floats, bool, string, bytes) have hashing / ordering /
codec support out of the box.  ``object`` columns holding user types can be
made hashable, sortable or custom-encodable with :func:`register_ops`,
keyed by python class.

Capability parity: reference ``frame/ops.go:31-134`` (``Ops{Less,
HashWithSeed, Encode, Decode}``, ``RegisterOps`` with double-registration
panic, ``CanCompare``/``CanHash``) and the generated builtins
(``frame/ops_builtin.go``).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, Optional

from ..slicetype import DType


@dataclass
class Ops:
    """Operations on elements of a user type stored in ``object`` columns.

    ``less(a, b) -> bool``; ``hash_with_seed(v, seed) -> uint32``;
    ``encode(session, values) -> bytes`` / ``decode(session, data, n) ->
    list`` implement a custom column codec with per-stream ``session`` state
    (see :mod:`bigslice_b200.frame.codec`).
    """

    less: Optional[Callable[[Any, Any], bool]] = None
    hash_with_seed: Optional[Callable[[Any, int], int]] = None
    encode: Optional[Callable[[Any, list], bytes]] = None
    decode: Optional[Callable[[Any, bytes, int], list]] = None


_registry: Dict[type, Ops] = {}


def register_ops(cls: type, ops: Ops) -> None:
    """Register ops for a python class; registering twice is an error."""
    if cls in _registry:
        raise RuntimeError("frame.register_ops: ops already registered for type %s" % cls.__name__)
    _registry[cls] = ops


RegisterOps = register_ops


def unregister_ops(cls: type) -> None:  # test helper
    _registry.pop(cls, None)


def ops_for(cls: type) -> Optional[Ops]:
    for k in cls.__mro__:
        if k in _registry:
            return _registry[k]
    return None


def ops_for_value(v) -> Optional[Ops]:
    return ops_for(type(v)) if v is not None else None


def can_compare(dt: DType) -> bool:
    """Whether columns of ``dt`` can be ordered (``frame.CanCompare``)."""
    if dt.comparable:
        return True
    return False


def can_hash(dt: DType) -> bool:
    """Whether columns of ``dt`` can be hashed (``frame.CanHash``)."""
    if dt.hashable:
        return True
    return False


def can_compare_value(col) -> bool:
    import numpy as np
    if isinstance(col, np.ndarray) and col.dtype == object and len(col):
        o = ops_for_value(col[0])
        return o is not None and o.less is not None
    return False


CanCompare, CanHash = can_compare, can_hash
