"""Column-type tuples for slices and frames.

A :class:`Type` is an ordered tuple of column :class:`DType` descriptors plus
a *prefix*: the number of leading columns that form the row key (used by
hashing, sorting, Reduce, Cogroup, ...).

Capability parity: reference ``slicetype/slicetype.go:17-161`` (``Type``,
``New``, ``Assignable``, ``Columns``, ``Concat``, ``Append``, ``Slice``,
``String``, ``Signature``).  The design is new: dtypes are first-class
descriptors that map to torch/numpy storage rather than ``reflect.Type``.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Optional, Tuple

import numpy as np
import torch


@dataclass(frozen=True)
class DType:
    """Descriptor of one column's element type.

    ``kind`` is one of ``int``, ``uint``, ``float``, ``bool`` (tensor-backed),
    ``str``, ``bytes``, ``object`` (host object columns), ``list`` (ragged list
    column: offsets + child values) or ``vec`` (fixed-width vector per row,
    stored as a 2-D tensor).
    """

    name: str
    kind: str
    torch_dtype: Optional[torch.dtype] = None
    itemsize: int = 0
    child: Optional["DType"] = None
    width: int = 0

    def __repr__(self) -> str:  # Go-like spelling in messages
        return self.name

    @property
    def is_tensor(self) -> bool:
        return self.torch_dtype is not None and self.kind != "list"

    @property
    def is_list(self) -> bool:
        return self.kind == "list"

    @property
    def hashable(self) -> bool:
        return self.kind in ("int", "uint", "float", "bool", "str", "bytes")

    @property
    def comparable(self) -> bool:
        return self.kind in ("int", "uint", "float", "bool", "str", "bytes")


int8 = DType("int8", "int", torch.int8, 1)
int16 = DType("int16", "int", torch.int16, 2)
int32 = DType("int32", "int", torch.int32, 4)
int64 = DType("int64", "int", torch.int64, 8)
uint8 = DType("uint8", "uint", torch.uint8, 1)
uint16 = DType("uint16", "uint", torch.uint16, 2)
uint32 = DType("uint32", "uint", torch.uint32, 4)
uint64 = DType("uint64", "uint", torch.uint64, 8)
float16 = DType("float16", "float", torch.float16, 2)
bfloat16 = DType("bfloat16", "float", torch.bfloat16, 2)
float32 = DType("float32", "float", torch.float32, 4)
float64 = DType("float64", "float", torch.float64, 8)
bool_ = DType("bool", "bool", torch.bool, 1)
string = DType("string", "str")
bytes_ = DType("bytes", "bytes")
obj = DType("object", "object")

_ALL = [int8, int16, int32, int64, uint8, uint16, uint32, uint64, float16,
        bfloat16, float32, float64, bool_, string, bytes_, obj]
_BY_NAME = {d.name: d for d in _ALL}
_BY_TORCH = {d.torch_dtype: d for d in _ALL if d.torch_dtype is not None}


def List(child: "DType | type | str") -> DType:
    """Ragged list-of-``child`` column type (Cogroup value columns)."""
    c = dtype_of(child)
    return DType("[]" + c.name, "list", c.torch_dtype, c.itemsize, c)


def Vec(child: "DType | type | str", width: int) -> DType:
    """Fixed-width vector column: ``width`` elements of ``child`` per row."""
    c = dtype_of(child)
    if not c.is_tensor:
        raise TypeError("Vec child must be a tensor dtype")
    return DType("[%d]%s" % (width, c.name), "vec", c.torch_dtype,
                 c.itemsize * width, c, width)


_PY = {int: int64, float: float64, bool: bool_, str: string, bytes: bytes_,
       object: obj}


def dtype_of(x) -> DType:
    """Coerce a python type, torch/numpy dtype, name or DType into a DType."""
    if isinstance(x, DType):
        return x
    if isinstance(x, str):
        if x in _BY_NAME:
            return _BY_NAME[x]
        if x == "str":
            return string
        if x == "int":
            return int64
        if x == "float":
            return float64
        raise TypeError("unknown dtype name %r" % (x,))
    if isinstance(x, torch.dtype):
        if x in _BY_TORCH:
            return _BY_TORCH[x]
        raise TypeError("unsupported torch dtype %r" % (x,))
    if isinstance(x, type) and x in _PY:
        return _PY[x]
    origin = getattr(x, "__origin__", None)
    if origin in (list,):
        (a,) = x.__args__
        return List(dtype_of(a))
    try:
        nd = np.dtype(x)
    except Exception:
        raise TypeError("cannot interpret %r as a column dtype" % (x,))
    if nd == np.dtype(object):
        return obj
    if nd.kind in ("U", "S"):
        return string if nd.kind == "U" else bytes_
    name = nd.name
    if name in _BY_NAME:
        return _BY_NAME[name]
    raise TypeError("cannot interpret %r as a column dtype" % (x,))


def dtype_of_value(v) -> DType:
    """Infer the DType of a column *value* (tensor, ndarray, list...)."""
    from .frame.columns import ListColumn
    if isinstance(v, ListColumn):
        return List(v.child_dtype)
    if isinstance(v, torch.Tensor):
        d = dtype_of(v.dtype)
        if v.dim() == 2:
            return Vec(d, v.shape[1])
        return d
    if isinstance(v, np.ndarray):
        if v.dtype == object:
            for e in v:
                if e is None:
                    continue
                if isinstance(e, str):
                    return string
                if isinstance(e, (bytes, bytearray)):
                    return bytes_
                return obj
            return obj
        if v.dtype.kind == "U":
            return string
        if v.dtype.kind == "S":
            return bytes_
        return dtype_of(v.dtype)
    if isinstance(v, (list, tuple)):
        if len(v) == 0:
            return int64  # Go's default in tests is []int
        e = v[0]
        if isinstance(e, bool):
            return bool_
        if isinstance(e, int):
            return int64
        if isinstance(e, float):
            return float64
        if isinstance(e, str):
            return string
        if isinstance(e, (bytes, bytearray)):
            return bytes_
        return obj
    raise TypeError("cannot infer column dtype of %r" % (type(v),))


class Type:
    """An ordered tuple of column dtypes and a key prefix.

    Mirrors the ``slicetype.Type`` interface (NumOut/Out/Prefix).
    """

    __slots__ = ("cols", "_prefix")

    def __init__(self, cols: Iterable, prefix: int = 1):
        self.cols: Tuple[DType, ...] = tuple(dtype_of(c) for c in cols)
        self._prefix = int(prefix)

    def num_out(self) -> int:
        return len(self.cols)

    def out(self, i: int) -> DType:
        return self.cols[i]

    def prefix(self) -> int:
        return self._prefix

    NumOut, Out, Prefix = num_out, out, prefix

    def __len__(self):
        return len(self.cols)

    def __iter__(self):
        return iter(self.cols)

    def __getitem__(self, i):
        return self.cols[i]

    def __eq__(self, other):
        return isinstance(other, Type) and self.cols == other.cols

    def __hash__(self):
        return hash(self.cols)

    def __repr__(self):
        return string_of(self)


def New(*cols, prefix: int = 1) -> Type:
    """Construct a Type from column dtypes (``slicetype.New``)."""
    if len(cols) == 1 and isinstance(cols[0], (list, tuple)):
        cols = tuple(cols[0])
    return Type(cols, prefix)


def as_type(t, prefix: Optional[int] = None) -> Type:
    if isinstance(t, Type):
        if prefix is None or prefix == t.prefix():
            return t
        return Type(t.cols, prefix)
    if hasattr(t, "num_out") and hasattr(t, "out"):
        return Type([t.out(i) for i in range(t.num_out())],
                    t.prefix() if prefix is None else prefix)
    return Type(t, 1 if prefix is None else prefix)


def Assignable(in_: Type, to: Type) -> bool:
    """Whether column types ``in_`` can be assigned to ``to``.

    Assignable if same arity and each column is identical, or the target
    column is ``object`` (analogue of interface assignability,
    ``slicetype/slicetype.go:54-66``).
    """
    if in_.num_out() != to.num_out():
        return False
    for a, b in zip(in_.cols, to.cols):
        if a != b and b.kind != "object":
            return False
    return True


def Columns(t: Type) -> Tuple[DType, ...]:
    return tuple(t.cols)


def Concat(*types: Type) -> Type:
    """Concatenate column lists; prefix is taken from the first type."""
    cols = []
    for t in types:
        cols.extend(t.cols)
    prefix = types[0].prefix() if types else 1
    return Type(cols, prefix)


def Append(t: Type, *extra) -> Type:
    return Type(tuple(t.cols) + tuple(dtype_of(e) for e in extra), t.prefix())


def Slice(t: Type, i: int, j: int) -> Type:
    """Columns [i,j) of t; prefix shrinks accordingly (never below 0)."""
    prefix = max(0, min(t.prefix(), j) - i)
    return Type(t.cols[i:j], prefix)


def string_of(t: Type) -> str:
    """``slice[prefix]<a, b, c>`` rendering (``slicetype.String``)."""
    inner = ", ".join(c.name for c in t.cols)
    return "slice[%d]<%s>" % (t.prefix(), inner)


String = string_of


def Signature(arg: Type, ret: Type) -> str:
    """``func(a, b) (c, d)`` rendering used in typecheck messages."""
    a = ", ".join(c.name for c in arg.cols)
    r = ", ".join(c.name for c in ret.cols)
    if ret.num_out() == 0:
        return "func(%s)" % a
    if ret.num_out() == 1:
        return "func(%s) %s" % (a, r)
    return "func(%s) (%s)" % (a, r)
