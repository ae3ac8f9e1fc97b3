"""User-function descriptors.

A :class:`Func` wraps a user callable used by Map/Filter/Flatmap/Fold/Reduce
etc.  Unlike the reference (one reflective call per row), functions here are
**vectorised by default**: they receive whole column batches (torch tensors
for numeric columns, numpy object arrays for str/bytes/object columns) and
return column batches, so they run at tensor speed on the GPU.  Per-row
python functions are supported through :func:`rowwise`.

A function whose first parameter is named ``ctx`` (or annotated with
:class:`Context`) receives the task context (metrics scope), mirroring the
optional leading ``context.Context`` argument in the reference.

Capability parity: reference ``slicefunc/func.go:40-104`` (Func{In,Out,
IsVariadic}, Of, Call, IsNil).
"""
from __future__ import annotations

import inspect
import typing
from typing import Any, Callable, List, Optional, Sequence

import numpy as np
import torch

from . import slicetype
from .slicetype import DType, Type


class Context:
    """Per-task call context handed to user functions that ask for it."""

    def __init__(self, scope=None, shard: int = 0, device=None, values=None):
        self.scope = scope
        self.shard = shard
        self.device = device
        self.values = values or {}

    def metrics_scope(self):
        return self.scope


_BACKGROUND = Context()


def background() -> Context:
    return _BACKGROUND


def rowwise(fn: Callable) -> Callable:
    """Mark ``fn`` as a per-row function (called once per row with scalars)."""
    try:
        fn.__bigslice_rowwise__ = True
        return fn
    except AttributeError:  # builtins / bound methods
        def wrapper(*a, **k):
            return fn(*a, **k)
        wrapper.__bigslice_rowwise__ = True
        wrapper.__wrapped__ = fn
        return wrapper


def vectorized(fn: Callable) -> Callable:
    """Explicitly mark ``fn`` as operating on whole column batches."""
    fn.__bigslice_rowwise__ = False
    return fn


def _ann_to_types(ann) -> Optional[List[DType]]:
    if ann is inspect.Signature.empty or ann is None:
        return None
    origin = typing.get_origin(ann)
    try:
        if origin in (tuple, typing.Tuple):
            return [slicetype.dtype_of(a) for a in typing.get_args(ann)]
        return [slicetype.dtype_of(ann)]
    except TypeError:
        return None


class Func:
    """Descriptor of a user function: parameter/result types and calling mode."""

    def __init__(self, fn: Optional[Callable], in_: Optional[Type] = None,
                 out: Optional[Type] = None, rowwise_: Optional[bool] = None,
                 builtin: Optional[str] = None):
        self.fn = fn
        self.In = in_
        self.Out = out
        self.is_variadic = False
        self.takes_ctx = False
        self.nparams = 0
        self.builtin = builtin or getattr(fn, "__bigslice_builtin__", None)
        if rowwise_ is None:
            rowwise_ = bool(getattr(fn, "__bigslice_rowwise__", False))
        self.rowwise = rowwise_
        if fn is not None:
            self._inspect()

    def _inspect(self):
        try:
            sig = inspect.signature(self.fn)
        except (TypeError, ValueError):
            self.is_variadic = True
            return
        params = list(sig.parameters.values())
        if params and (params[0].name == "ctx" or params[0].annotation is Context):
            self.takes_ctx = True
            params = params[1:]
        self.is_variadic = any(p.kind == p.VAR_POSITIONAL for p in params)
        self.nparams = len([p for p in params if p.kind in
                            (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)])
        try:
            hints = typing.get_type_hints(self.fn)
        except Exception:
            hints = {}
        if self.In is None:
            tys = []
            for p in params:
                a = hints.get(p.name, p.annotation)
                t = _ann_to_types(a)
                if t is None or len(t) != 1:
                    tys = None
                    break
                tys.append(t[0])
            if tys is not None and params and not self.is_variadic:
                self.In = Type(tys)
        if self.Out is None:
            t = _ann_to_types(hints.get("return", sig.return_annotation))
            if t is not None:
                self.Out = Type(t)

    def is_nil(self) -> bool:
        return self.fn is None and self.builtin is None

    IsNil = is_nil

    def accepts(self, n: int) -> bool:
        """Whether the function can be called with ``n`` column arguments."""
        if self.fn is None:
            return True
        if self.is_variadic:
            return n >= self.nparams
        return n == self.nparams

    # -- calling -----------------------------------------------------------
    def call(self, ctx: Optional[Context], cols: Sequence[Any]) -> List[Any]:
        """Apply to a batch of columns; returns a list of output columns."""
        if self.rowwise:
            return self._call_rows(ctx, cols)
        if self.takes_ctx:
            r = self.fn(ctx or _BACKGROUND, *cols)
        else:
            r = self.fn(*cols)
        if r is None:
            return []
        if isinstance(r, (tuple, list)):
            return list(r)
        return [r]

    Call = call

    def call_row(self, ctx: Optional[Context], vals: Sequence[Any]):
        if self.takes_ctx:
            return self.fn(ctx or _BACKGROUND, *vals)
        return self.fn(*vals)

    def _call_rows(self, ctx, cols):
        from .frame import columns as C
        n = C.col_len(cols[0]) if cols else 0
        lists = [C.col_tolist(c) for c in cols]
        outs: Optional[List[list]] = None
        for i in range(n):
            r = self.call_row(ctx, [l[i] for l in lists])
            if not isinstance(r, tuple):
                r = (r,)
            if outs is None:
                outs = [[] for _ in r]
            for o, v in zip(outs, r):
                o.append(v)
        if outs is None:
            if self.Out is None:
                return []
            outs = [[] for _ in range(self.Out.num_out())]
        device = None
        for c in cols:
            if isinstance(c, torch.Tensor):
                device = c.device
                break
        if self.Out is not None:
            return [C.col_from_list(o, t, device) for o, t in zip(outs, self.Out.cols)]
        return [C.col_from_list(o, slicetype.dtype_of_value(o), device) for o in outs]

    # -- type inference ----------------------------------------------------
    def infer_out(self, arg: Type, flat: bool = False) -> Optional[Type]:
        """Infer output column types by probing with a one-row zero batch.

        ``flat`` is for Flatmap functions whose per-row results are lists.
        Returns None if the probe fails (callers ask for an explicit ``out``).
        """
        if self.Out is not None:
            return self.Out
        from .frame import columns as C
        try:
            probe = [C.col_zeros(t, 1, None) for t in arg.cols]
            if self.rowwise:
                vals = [C.col_tolist(c)[0] for c in probe]
                r = self.call_row(_BACKGROUND, vals)
                if not isinstance(r, tuple):
                    r = (r,)
                if flat:
                    tys = []
                    for v in r:
                        t = slicetype.dtype_of_value(v if len(v) else [])
                        tys.append(t)
                    out = Type(tys)
                else:
                    out = Type([_scalar_dtype(v) for v in r])
            else:
                r = self.call(_BACKGROUND, probe)
                out = Type([slicetype.dtype_of_value(v) for v in r])
        except Exception:
            return None
        self.Out = out
        return out


def _scalar_dtype(v) -> DType:
    if isinstance(v, bool):
        return slicetype.bool_
    if isinstance(v, (int, np.integer)):
        return slicetype.int64
    if isinstance(v, (float, np.floating)):
        return slicetype.float64
    if isinstance(v, str):
        return slicetype.string
    if isinstance(v, (bytes, bytearray)):
        return slicetype.bytes_
    if isinstance(v, torch.Tensor) and v.dim() == 0:
        return slicetype.dtype_of(v.dtype)
    return slicetype.obj


def of(fn, in_=None, out=None) -> Func:
    """Build a Func from a callable (``slicefunc.Of``)."""
    if isinstance(fn, Func):
        return fn
    if fn is not None and not callable(fn):
        raise TypeError("slicefunc.of: not a function: %r" % (fn,))
    if in_ is not None and not isinstance(in_, Type):
        in_ = Type(in_)
    if out is not None and not isinstance(out, Type):
        if not isinstance(out, (list, tuple)):
            out = [out]
        out = Type(out)
    return Func(fn, in_, out)


Of = of
Nil = Func(None)
