"""Symbolic tracing of python row functions into fused-kernel expressions.

``Map(s, lambda k, v: (k % 13, v * 2))`` is written against whole column
batches (torch tensors).  Evaluated eagerly, every operator is its own PyTorch
kernel with its own temporary; a Map → Filter → Map pipeline over a 2^27-row
tile is a dozen launches and a dozen passes over HBM.  This module calls the
function **once with symbolic columns** instead.  Every operator on a
:class:`Sym` records a typed C expression (torch's type-promotion rules are
applied while tracing, python's ``%`` and ``//`` keep their floor semantics
through helper templates in the kernel prelude), so a traceable function becomes
a :class:`bigslice_b200.ops.jit.Stage` and is compiled with NVRTC into the same
single fused sm_100a kernel as its traceable neighbours (``ops/jit.py``).

Anything the tracer does not understand — data-dependent control flow
(``if k > 0``), ``len(col)``, indexing a lookup table, unknown torch functions,
string/list columns, a ``ctx`` parameter — raises :class:`TraceError` and the
stage simply runs eagerly as before.  Traced stages keep the original function
(``Stage.eager``) for host frames, and the first device tile of every traced
chain is cross-checked against the eager result (``slices._ExprReader``), so an
impure or mis-traced function falls back instead of producing different rows.

:func:`recognise_combiner` uses the same machinery on two-argument combiners:
``lambda a, b: a + b`` (the reference's canonical Reduce function,
``reduce.go:30-41``) is recognised as the builtin ``add`` and lowers to the
fused hash-combine/exchange kernels instead of a sort + segmented tree reduce.

This is what moves tier (c) of the user-function model (vectorised python
callables) onto tier (b)'s kernels; the reference calls the user
function reflectively once per row (``slice.go:621-632,700-720``).
"""
from __future__ import annotations

import math
import os
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from .. import slicetype
from ..slicetype import DType
from . import jit

__all__ = ["TraceError", "Sym", "trace_map", "trace_filter", "trace_fold", "recognise_combiner", "enabled"]

#: dtypes a traced expression may carry (uint16/32/64 have no torch arithmetic to compare against)
_OK = (torch.int8, torch.int16, torch.int32, torch.int64, torch.uint8, torch.float32, torch.float64, torch.bool)
_CT = {torch.int8: "signed char", torch.int16: "short", torch.int32: "int", torch.int64: "long long",
       torch.uint8: "unsigned char", torch.float32: "float", torch.float64: "double", torch.bool: "bool"}
_RANGE = {torch.int8: (-128, 127), torch.int16: (-32768, 32767), torch.int32: (-2 ** 31, 2 ** 31 - 1),
          torch.int64: (-2 ** 63, 2 ** 63 - 1), torch.uint8: (0, 255)}


def enabled() -> bool:
    return os.environ.get("BIGSLICE_TRACE", "1") != "0"


class TraceError(Exception):
    """The function cannot be expressed as a fused expression; run it eagerly."""


def _is_float(dt) -> bool:
    return dt in (torch.float32, torch.float64)


def _is_int(dt) -> bool:
    return dt in _RANGE


def _literal(v, dt) -> str:
    """C literal of python scalar ``v`` converted to ``dt`` the way torch converts
    a scalar operand."""
    if dt == torch.bool:
        return "true" if bool(v) else "false"
    if _is_int(dt):
        if isinstance(v, float):
            if not math.isfinite(v):
                raise TraceError("non-finite scalar in an integer expression")
            v = int(v)
        v = int(v)
        lo, hi = _RANGE[dt]
        if not lo <= v <= hi:
            raise TraceError("scalar %d does not fit %s" % (v, dt))
        if v == -2 ** 63:
            return "((long long)(-9223372036854775807LL - 1))"
        return "((%s)(%dLL))" % (_CT[dt], v)
    v = float(v)
    if v != v:
        return "((%s)(0.0 / 0.0))" % _CT[dt]
    if math.isinf(v):
        return "((%s)(%s1.0 / 0.0))" % (_CT[dt], "-" if v < 0 else "")
    return "((%s)(%r))" % (_CT[dt], v)


def _scalar(x):
    """Python scalar of a scalar-like operand, or None."""
    if isinstance(x, (bool, int, float)):
        return x
    if isinstance(x, np.generic):
        return x.item()
    if isinstance(x, torch.Tensor) and x.dim() == 0 and x.dtype in _OK:
        return x.item()
    return None


def _dummy(x):
    return torch.empty(1, dtype=x.dtype) if isinstance(x, Sym) else x


class Sym:
    """A symbolic column: a C expression over ``c0, c1, ...`` and its torch dtype."""

    __slots__ = ("expr", "dtype")
    __array_priority__ = 1000.0   # numpy scalars defer to our reflected operators
    __array_ufunc__ = None

    def __init__(self, expr: str, dtype: torch.dtype):
        if dtype not in _OK:
            raise TraceError("dtype %s is not traceable" % (dtype,))
        self.expr, self.dtype = expr, dtype

    # -- helpers -------------------------------------------------------------
    @staticmethod
    def lift(x, dt) -> str:
        """Expression of operand ``x`` converted to ``dt``."""
        if isinstance(x, Sym):
            return x.expr if x.dtype == dt else "((%s)%s)" % (_CT[dt], x.expr)
        v = _scalar(x)
        if v is None:
            raise TraceError("operand of type %s is not traceable" % type(x).__name__)
        return _literal(v, dt)

    @staticmethod
    def common(a, b) -> torch.dtype:
        for x in (a, b):
            if not isinstance(x, Sym) and _scalar(x) is None:
                raise TraceError("operand of type %s is not traceable" % type(x).__name__)
        a = a if isinstance(a, Sym) else _scalar(a)
        b = b if isinstance(b, Sym) else _scalar(b)
        dt = torch.result_type(_dummy(a), _dummy(b))
        if dt not in _OK:
            raise TraceError("dtype %s is not traceable" % (dt,))
        return dt

    def _arith(self, op: str, a, b) -> "Sym":
        dt = Sym.common(a, b)
        if dt == torch.bool:
            raise TraceError("arithmetic on bool columns")
        return Sym("((%s)(%s %s %s))" % (_CT[dt], Sym.lift(a, dt), op, Sym.lift(b, dt)), dt)

    def _cmp(self, op: str, a, b) -> "Sym":
        dt = Sym.common(a, b)
        return Sym("(%s %s %s)" % (Sym.lift(a, dt), op, Sym.lift(b, dt)), torch.bool)

    def _bits(self, op: str, a, b) -> "Sym":
        dt = Sym.common(a, b)
        if _is_float(dt):
            raise TraceError("bitwise operator on a float column")
        return Sym("((%s)(%s %s %s))" % (_CT[dt], Sym.lift(a, dt), op, Sym.lift(b, dt)), dt)

    def _shift(self, op: str, a, b) -> "Sym":
        dt = Sym.common(a, b)
        if not _is_int(dt):
            raise TraceError("shift of a non-integer column")
        return Sym("((%s)(%s %s %s))" % (_CT[dt], Sym.lift(a, dt), op, Sym.lift(b, dt)), dt)

    def _truediv(self, a, b) -> "Sym":
        dt = Sym.common(a, b)
        if not _is_float(dt):
            dt = torch.get_default_dtype()
        return Sym("(%s / %s)" % (Sym.lift(a, dt), Sym.lift(b, dt)), dt)

    def _floordiv(self, a, b) -> "Sym":
        dt = Sym.common(a, b)
        if not _is_int(dt):
            raise TraceError("floor division of non-integer columns")
        return Sym("bs_floordiv(%s, %s)" % (Sym.lift(a, dt), Sym.lift(b, dt)), dt)

    def _mod(self, a, b) -> "Sym":
        dt = Sym.common(a, b)
        if dt == torch.bool:
            raise TraceError("arithmetic on bool columns")
        return Sym("bs_pymod(%s, %s)" % (Sym.lift(a, dt), Sym.lift(b, dt)), dt)

    def _pow(self, a, b) -> "Sym":
        e = _scalar(b)
        if isinstance(a, Sym) and isinstance(e, int) and not isinstance(e, bool) and 1 <= e <= 4 \
                and a.dtype != torch.bool:
            x = a.expr
            return Sym("((%s)(%s))" % (_CT[a.dtype], " * ".join([x] * e)), a.dtype)
        dt = Sym.common(a, b)
        if not _is_float(dt):
            raise TraceError("integer power")
        return Sym("pow(%s, %s)" % (Sym.lift(a, dt), Sym.lift(b, dt)), dt)

    # -- operators -----------------------------------------------------------
    def __add__(self, o): return self._arith("+", self, o)
    def __radd__(self, o): return self._arith("+", o, self)
    def __sub__(self, o): return self._arith("-", self, o)
    def __rsub__(self, o): return self._arith("-", o, self)
    def __mul__(self, o): return self._arith("*", self, o)
    def __rmul__(self, o): return self._arith("*", o, self)
    def __truediv__(self, o): return self._truediv(self, o)
    def __rtruediv__(self, o): return self._truediv(o, self)
    def __floordiv__(self, o): return self._floordiv(self, o)
    def __rfloordiv__(self, o): return self._floordiv(o, self)
    def __mod__(self, o): return self._mod(self, o)
    def __rmod__(self, o): return self._mod(o, self)
    def __pow__(self, o): return self._pow(self, o)
    def __rpow__(self, o): return self._pow(o, self)
    def __and__(self, o): return self._bits("&", self, o)
    def __rand__(self, o): return self._bits("&", o, self)
    def __or__(self, o): return self._bits("|", self, o)
    def __ror__(self, o): return self._bits("|", o, self)
    def __xor__(self, o): return self._bits("^", self, o)
    def __rxor__(self, o): return self._bits("^", o, self)
    def __lshift__(self, o): return self._shift("<<", self, o)
    def __rshift__(self, o): return self._shift(">>", self, o)
    def __lt__(self, o): return self._cmp("<", self, o)
    def __le__(self, o): return self._cmp("<=", self, o)
    def __gt__(self, o): return self._cmp(">", self, o)
    def __ge__(self, o): return self._cmp(">=", self, o)
    def __eq__(self, o): return self._cmp("==", self, o)   # noqa: E704
    def __ne__(self, o): return self._cmp("!=", self, o)
    __hash__ = None

    def __neg__(self):
        if self.dtype == torch.bool:
            raise TraceError("negation of a bool column")
        return Sym("((%s)(-%s))" % (_CT[self.dtype], self.expr), self.dtype)

    def __pos__(self):
        return self

    def __invert__(self):
        if self.dtype == torch.bool:
            return Sym("(!%s)" % self.expr, torch.bool)
        if not _is_int(self.dtype):
            raise TraceError("~ of a float column")
        return Sym("((%s)(~%s))" % (_CT[self.dtype], self.expr), self.dtype)

    def __abs__(self):
        return _abs(self)

    # anything that needs a concrete value ends the trace
    def __bool__(self):
        raise TraceError("data-dependent control flow")

    def __len__(self):
        raise TraceError("len() of a column")

    def __iter__(self):
        raise TraceError("iteration over a column")

    def __index__(self):
        raise TraceError("column used as an index")

    def __int__(self):
        raise TraceError("column converted to int")

    def __float__(self):
        raise TraceError("column converted to float")

    def __getitem__(self, _):
        raise TraceError("indexing a column")

    def __getattr__(self, name):
        raise TraceError("column attribute %r is not traceable" % name)

    # -- tensor-method surface -------------------------------------------------
    def to(self, *args, **kwargs):
        dt = kwargs.get("dtype")
        for a in args:
            if isinstance(a, torch.dtype):
                dt = a
            elif isinstance(a, (str, torch.device)):
                raise TraceError("device move inside a row function")
        if dt is None:
            return self
        return _cast(self, dt)

    def type(self, dt=None):
        if dt is None:
            raise TraceError("type() query")
        return _cast(self, dt)

    def float(self): return _cast(self, torch.float32)     # noqa: E704
    def double(self): return _cast(self, torch.float64)    # noqa: E704
    def long(self): return _cast(self, torch.int64)        # noqa: E704
    def int(self): return _cast(self, torch.int32)         # noqa: E704
    def short(self): return _cast(self, torch.int16)       # noqa: E704
    def bool(self): return _cast(self, torch.bool)         # noqa: E704
    def abs(self): return _abs(self)                       # noqa: E704
    def neg(self): return -self                            # noqa: E704
    def sqrt(self): return _math1("sqrt", self)            # noqa: E704
    def exp(self): return _math1("exp", self)              # noqa: E704
    def log(self): return _math1("log", self)              # noqa: E704
    def floor(self): return _round("floor", self)          # noqa: E704
    def ceil(self): return _round("ceil", self)            # noqa: E704
    def square(self): return self * self                   # noqa: E704
    def clone(self): return self                           # noqa: E704
    def contiguous(self): return self                      # noqa: E704
    def logical_not(self): return ~_cast(self, torch.bool)  # noqa: E704
    def logical_and(self, o): return _cast(self, torch.bool) & _as_bool(o)   # noqa: E704
    def logical_or(self, o): return _cast(self, torch.bool) | _as_bool(o)    # noqa: E704
    def minimum(self, o): return _minmax("bs_min", self, o)   # noqa: E704
    def maximum(self, o): return _minmax("bs_max", self, o)   # noqa: E704
    def remainder(self, o): return self % o                # noqa: E704
    def fmod(self, o): return _fmod(self, o)               # noqa: E704

    def clamp(self, min=None, max=None):  # noqa: A002
        return _clamp(self, min, max)

    clip = clamp

    def where(self, cond, other):
        raise TraceError("Tensor.where")

    # torch.* functions applied to symbolic columns
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        h = _TORCH.get(func)
        if h is None:
            raise TraceError("torch function %s is not traceable" % getattr(func, "__name__", func))
        return h(*args, **kwargs)

    def __repr__(self):
        return "Sym(%s: %s)" % (self.expr, self.dtype)


def _cast(x: Sym, dt) -> Sym:
    if not isinstance(dt, torch.dtype):
        try:
            dt = slicetype.dtype_of(dt).torch_dtype
        except Exception:
            raise TraceError("cast to %r" % (dt,))
    if dt not in _OK:
        raise TraceError("cast to %s" % (dt,))
    if dt == x.dtype:
        return x
    if dt == torch.bool:
        return Sym("(%s != %s)" % (x.expr, _literal(0, x.dtype)), torch.bool)
    return Sym("((%s)%s)" % (_CT[dt], x.expr), dt)


def _as_sym(x, like: Optional[Sym] = None) -> Sym:
    if isinstance(x, Sym):
        return x
    v = _scalar(x)
    if v is None:
        raise TraceError("operand of type %s is not traceable" % type(x).__name__)
    dt = torch.result_type(v, v) if like is None else Sym.common(like, v)
    if dt not in _OK:
        raise TraceError("dtype %s" % (dt,))
    return Sym(_literal(v, dt), dt)


def _as_bool(x) -> Sym:
    return _cast(_as_sym(x), torch.bool)


def _abs(x) -> Sym:
    x = _as_sym(x)
    if x.dtype in (torch.bool, torch.uint8):
        return x
    return Sym("bs_abs(%s)" % x.expr, x.dtype)


def _math1(fn: str, x) -> Sym:
    x = _as_sym(x)
    dt = x.dtype if _is_float(x.dtype) else torch.get_default_dtype()
    return Sym("%s(%s)" % (fn, Sym.lift(x, dt)), dt)


def _round(fn: str, x) -> Sym:
    x = _as_sym(x)
    if not _is_float(x.dtype):
        return x
    return Sym("%s(%s)" % (fn, x.expr), x.dtype)


def _minmax(fn: str, a, b) -> Sym:
    if not isinstance(a, Sym) and not isinstance(b, Sym):
        raise TraceError("min/max of constants")
    dt = Sym.common(a, b)
    return Sym("%s(%s, %s)" % (fn, Sym.lift(a, dt), Sym.lift(b, dt)), dt)


def _fmod(a, b) -> Sym:
    dt = Sym.common(a, b)
    if dt == torch.bool:
        raise TraceError("arithmetic on bool columns")
    if _is_float(dt):
        return Sym("fmod(%s, %s)" % (Sym.lift(a, dt), Sym.lift(b, dt)), dt)
    return Sym("((%s)(%s %% %s))" % (_CT[dt], Sym.lift(a, dt), Sym.lift(b, dt)), dt)


def _clamp(x, min=None, max=None) -> Sym:  # noqa: A002
    x = _as_sym(x)
    if min is None and max is None:
        raise TraceError("clamp without bounds")
    for b in (min, max):
        if b is not None and not isinstance(b, Sym) and _scalar(b) is None:
            raise TraceError("clamp bound")
    dt = x.dtype
    for b in (min, max):
        if b is not None:
            dt = torch.promote_types(dt, Sym.common(x, b))
    if dt == torch.bool:
        raise TraceError("clamp of a bool column")
    e = Sym.lift(x, dt)
    if min is not None:
        e = "bs_max(%s, %s)" % (e, Sym.lift(min, dt))
    if max is not None:
        e = "bs_min(%s, %s)" % (e, Sym.lift(max, dt))
    return Sym(e, dt)


def _where(cond, a=None, b=None) -> Sym:
    if a is None or b is None:
        raise TraceError("torch.where(cond) index form")
    c = _as_bool(cond)
    if not isinstance(a, Sym) and not isinstance(b, Sym):
        sa, sb = _scalar(a), _scalar(b)
        if sa is None or sb is None:
            raise TraceError("torch.where operands")
        dt = torch.result_type(sa, sb)
    else:
        dt = Sym.common(a, b)
    if dt not in _OK:
        raise TraceError("dtype %s" % (dt,))
    return Sym("(%s ? %s : %s)" % (c.expr, Sym.lift(a, dt), Sym.lift(b, dt)), dt)


def _const_like(value):
    def f(x, dtype=None, **kw):
        if set(kw) - {"device", "layout", "memory_format", "requires_grad", "pin_memory"}:
            raise TraceError("unsupported keyword")
        dt = dtype or _as_sym(x).dtype
        return Sym(_literal(value, dt), dt)
    return f


def _full_like(x, fill_value, dtype=None, **kw):
    v = _scalar(fill_value)
    if v is None:
        raise TraceError("full_like fill value")
    dt = dtype or _as_sym(x).dtype
    return Sym(_literal(v, dt), dt)


def _div(a, b, rounding_mode=None):
    s = a if isinstance(a, Sym) else b
    if rounding_mode is None:
        return s._truediv(a, b)
    if rounding_mode == "floor":
        return s._floordiv(a, b)
    if rounding_mode == "trunc":
        dt = Sym.common(a, b)
        if not _is_int(dt):
            raise TraceError("trunc division of float columns")
        return Sym("((%s)(%s / %s))" % (_CT[dt], Sym.lift(a, dt), Sym.lift(b, dt)), dt)
    raise TraceError("rounding mode %r" % (rounding_mode,))


def _binop(name):
    def f(a, b, **kw):
        if kw:
            raise TraceError("keyword arguments to torch.%s" % name)
        s = a if isinstance(a, Sym) else b
        if not isinstance(s, Sym):
            raise TraceError("no symbolic operand")
        swapped = s is not a
        m = getattr(Sym, ("__r%s__" if swapped else "__%s__") % name)
        return m(s, a if swapped else b)
    return f


_TORCH = {
    torch.where: _where,
    torch.abs: _abs, torch.absolute: _abs,
    torch.minimum: lambda a, b: _minmax("bs_min", a, b),
    torch.maximum: lambda a, b: _minmax("bs_max", a, b),
    torch.sqrt: lambda x: _math1("sqrt", x), torch.exp: lambda x: _math1("exp", x),
    torch.log: lambda x: _math1("log", x),
    torch.floor: lambda x: _round("floor", x), torch.ceil: lambda x: _round("ceil", x),
    torch.clamp: _clamp, torch.clip: _clamp,
    torch.logical_and: lambda a, b: _as_bool(a) & _as_bool(b),
    torch.logical_or: lambda a, b: _as_bool(a) | _as_bool(b),
    torch.logical_xor: lambda a, b: _as_bool(a) ^ _as_bool(b),
    torch.logical_not: lambda a: ~_as_bool(a),
    torch.remainder: lambda a, b: (a % b) if isinstance(a, Sym) else b.__rmod__(a),
    torch.fmod: _fmod,
    torch.div: _div, torch.true_divide: lambda a, b: _div(a, b),
    torch.floor_divide: lambda a, b: _div(a, b, "floor"),
    torch.add: _binop("add"), torch.sub: _binop("sub"), torch.mul: _binop("mul"),
    torch.eq: lambda a, b: _as_sym(a, b if isinstance(b, Sym) else None)._cmp("==", a, b),
    torch.ne: lambda a, b: _as_sym(a, b if isinstance(b, Sym) else None)._cmp("!=", a, b),
    torch.lt: lambda a, b: _as_sym(a, b if isinstance(b, Sym) else None)._cmp("<", a, b),
    torch.le: lambda a, b: _as_sym(a, b if isinstance(b, Sym) else None)._cmp("<=", a, b),
    torch.gt: lambda a, b: _as_sym(a, b if isinstance(b, Sym) else None)._cmp(">", a, b),
    torch.ge: lambda a, b: _as_sym(a, b if isinstance(b, Sym) else None)._cmp(">=", a, b),
    torch.bitwise_and: _binop("and"), torch.bitwise_or: _binop("or"), torch.bitwise_xor: _binop("xor"),
    torch.bitwise_not: lambda a: ~_as_sym(a),
    torch.neg: lambda a: -_as_sym(a), torch.negative: lambda a: -_as_sym(a),
    torch.square: lambda a: _as_sym(a) * _as_sym(a),
    torch.pow: lambda a, b: (a ** b) if isinstance(a, Sym) else b.__rpow__(a),
    torch.zeros_like: _const_like(0), torch.ones_like: _const_like(1), torch.full_like: _full_like,
    torch.clone: lambda a, **kw: a,
}


# ---------------------------------------------------------------------------
# entry points

def _traceable_types(dts: Sequence[DType]) -> bool:
    for d in dts:
        if d.kind == "vec" or not d.is_tensor or d.torch_dtype not in _OK:
            return False
    return True


def _call(func, in_types: Sequence[DType]) -> list:
    """Run the user function on symbolic columns; returns its outputs as a list."""
    if func.fn is None or func.rowwise or func.takes_ctx:
        raise TraceError("not a plain vectorised function")
    if not _traceable_types(in_types):
        raise TraceError("input column types")
    syms = [Sym("c%d" % j, t.torch_dtype) for j, t in enumerate(in_types)]
    try:
        r = func.fn(*syms)
    except TraceError:
        raise
    except Exception as e:  # the function did something a Sym cannot take part in
        raise TraceError("%s: %s" % (type(e).__name__, e))
    if r is None:
        return []
    return list(r) if isinstance(r, (tuple, list)) else [r]


def infer_out(func, in_types: Sequence[DType]) -> Optional[List[DType]]:
    """Output column types of a traceable function, from the trace alone (used when
    probing the function with a zero row fails, e.g. ``lambda k: 100 // k``)."""
    try:
        outs = _call(func, in_types)
        tys = []
        for o in outs:
            if isinstance(o, Sym):
                tys.append(slicetype.dtype_of(o.dtype))
            else:
                v = _scalar(o)
                if v is None:
                    return None
                tys.append(slicetype.dtype_of(torch.result_type(v, v)))
        return tys or None
    except TraceError:
        return None


def trace_map(func, in_types: Sequence[DType], out_types: Sequence[DType]) -> Optional[jit.Stage]:
    """A fused Map stage equivalent to ``func`` over columns of ``in_types``, or
    None when the function is not traceable."""
    if not enabled() or not _traceable_types(out_types):
        return None
    try:
        outs = _call(func, in_types)
        if len(outs) != len(out_types):
            return None
        exprs = []
        for o, t in zip(outs, out_types):
            if isinstance(o, Sym):
                exprs.append(o.expr)
            else:
                v = _scalar(o)
                if v is None:
                    return None
                exprs.append(_literal(v, t.torch_dtype))
    except TraceError:
        return None
    return jit.Stage(exprs=exprs, out=list(out_types))


def trace_filter(func, in_types: Sequence[DType]) -> Optional[jit.Stage]:
    """A fused Filter stage equivalent to predicate ``func``, or None."""
    if not enabled():
        return None
    try:
        outs = _call(func, in_types)
        if len(outs) != 1 or not isinstance(outs[0], Sym):
            return None
        return jit.Stage(pred=_cast(outs[0], torch.bool).expr)
    except TraceError:
        return None


def trace_fold(func, acc_type: DType, val_types: Sequence[DType]) -> Optional[str]:
    """The fold function ``fold(acc, v1, .., vn)`` as a C expression over ``acc`` and
    ``c1 .. cn`` (the form ``FoldExpr`` takes), or None."""
    if not enabled() or func.fn is None or func.rowwise or func.takes_ctx:
        return None
    if not _traceable_types([acc_type] + list(val_types)):
        return None
    syms = [Sym("acc", acc_type.torch_dtype)] + \
           [Sym("c%d" % (j + 1), t.torch_dtype) for j, t in enumerate(val_types)]
    try:
        r = func.fn(*syms)
    except Exception:
        return None
    if isinstance(r, (tuple, list)):
        if len(r) != 1:
            return None
        r = r[0]
    if isinstance(r, Sym):
        return r.expr
    v = _scalar(r)
    try:
        return None if v is None else _literal(v, acc_type.torch_dtype)
    except TraceError:
        return None


def recognise_combiner(fn: Callable, dt: DType) -> Optional[str]:
    """Name of the builtin combiner a two-argument function is equivalent to
    (``lambda a, b: a + b`` → ``"add"``), or None."""
    if not enabled() or fn is None or not _traceable_types([dt]):
        return None
    if getattr(fn, "__bigslice_rowwise__", False):
        return None
    t = dt.torch_dtype
    a, b = Sym("A", t), Sym("B", t)
    try:
        r = fn(a, b)
    except Exception:
        return None
    if isinstance(r, (tuple, list)):
        if len(r) != 1:
            return None
        r = r[0]
    if not isinstance(r, Sym) or r.dtype != t:
        return None
    C = _CT[t]
    forms = {}
    if t != torch.bool:
        forms["add"] = ["((%s)(A + B))" % C, "((%s)(B + A))" % C]
        forms["mul"] = ["((%s)(A * B))" % C, "((%s)(B * A))" % C]
        forms["min"] = ["bs_min(A, B)", "bs_min(B, A)"]
        forms["max"] = ["bs_max(A, B)", "bs_max(B, A)"]
    if not _is_float(t):
        forms["and"] = ["((%s)(A & B))" % C, "((%s)(B & A))" % C]
        forms["or"] = ["((%s)(A | B))" % C, "((%s)(B | A))" % C]
        forms["xor"] = ["((%s)(A ^ B))" % C, "((%s)(B ^ A))" % C]
    forms["first"] = ["A"]
    for name, fs in forms.items():
        if r.expr in fs:
            return name
    # min/max written with torch.where: where(a < b, a, b) and friends
    w = {"((A < B) ? A : B)": "min", "((A <= B) ? A : B)": "min",
         "((B < A) ? B : A)": "min", "((B <= A) ? B : A)": "min",
         "((A > B) ? A : B)": "max", "((A >= B) ? A : B)": "max",
         "((B > A) ? B : A)": "max", "((B >= A) ? B : A)": "max"}
    if r.expr in w and not _is_float(t):   # NaN handling differs for floats
        return w[r.expr]
    return None


# ---------------------------------------------------------------------------
# verdicts of the first-tile cross-check (slices._ExprReader): functions whose
# fused result disagreed with their eager result are never traced again.

_rejected = set()
_verified = set()


def _key(fn):
    code = getattr(fn, "__code__", None)
    return code if code is not None else id(fn)


def rejected(fn) -> bool:
    return _key(fn) in _rejected


def reject(fn) -> None:
    _rejected.add(_key(fn))


def verified(fn) -> bool:
    return _key(fn) in _verified


def mark_verified(fn) -> None:
    _verified.add(_key(fn))
