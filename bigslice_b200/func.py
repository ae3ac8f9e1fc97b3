"""Func registry and invocations.

A :class:`FuncValue` wraps a python function that *builds* a slice graph.
Funcs are registered globally in call order so that, in multi-process (one
rank per GPU) sessions, every rank can resolve ``(func index, args)`` to the
same graph deterministically — each rank re-invokes the Func and compiles the
identical task DAG, which is how shuffles line up across GPUs without
shipping closures.  An :class:`Invocation` is a Func applied to arguments and
gets a unique, monotonically increasing index used in task names (``inv3_...``).

Capability parity: reference ``func.go:19-258`` (Func, FuncValue.{Invocation,
Apply,Exclusive,NumIn,In}, Invocation, registration race guard, nil-argument
and arity/type checks).
"""
from __future__ import annotations

import inspect
import itertools
import threading
import typing
from typing import Any, Callable, List, Sequence

from . import typecheck
from .slices import Slice

_funcs: List["FuncValue"] = []
_funcs_lock = threading.Lock()
_funcs_busy = 0  # registration race detector (func.go:26-28,175-182)
_inv_index = itertools.count(1)
_inv_lock = threading.Lock()


class Invocation:
    """A Func applied to arguments (``bigslice.Invocation``)."""

    __slots__ = ("index", "func", "args", "exclusive", "location")

    def __init__(self, index: int, func: int, args: Sequence[Any], exclusive: bool,
                 location: str):
        self.index = index
        self.func = func
        self.args = list(args)
        self.exclusive = exclusive
        self.location = location

    def invoke(self) -> Slice:
        """Re-run the Func body to build the slice graph."""
        return _funcs[self.func].apply(*self.args)

    Invoke = invoke

    def __repr__(self):
        return "Invocation(inv%d, func %d, %s)" % (self.index, self.func, self.location)


_NILABLE = (list, dict, tuple, object, typing.Any)


def _check_arg(i: int, ann, val) -> None:
    if ann is inspect.Parameter.empty or ann is typing.Any:
        return
    origin = typing.get_origin(ann)
    if origin is typing.Union:
        args = typing.get_args(ann)
        if val is None and type(None) in args:
            return
        for a in args:
            try:
                _check_arg(i, a, val)
                return
            except typecheck.TypecheckError:
                continue
        typecheck.panicf("wrong type for argument %d: expected %s, got %s", i, ann,
                         type(val).__name__)
    base = origin or ann
    if val is None:
        if base in _NILABLE or base is Slice or (isinstance(base, type) and issubclass(base, Slice)):
            return
        typecheck.panicf("cannot use nil as type %s in argument to function",
                         getattr(base, "__name__", base))
    if not isinstance(base, type):
        return
    if base is float and isinstance(val, int) and not isinstance(val, bool):
        return
    if base is Slice or issubclass(base, Slice):
        if not isinstance(val, Slice):
            typecheck.panicf("wrong type for argument %d: type %s does not implement interface %s",
                             i, type(val).__name__, base.__name__)
        return
    if not isinstance(val, base) or (base is int and isinstance(val, bool)):
        typecheck.panicf("wrong type for argument %d: expected %s, got %s", i,
                         base.__name__, type(val).__name__)


class FuncValue:
    """A registered slice-building function (``bigslice.FuncValue``)."""

    def __init__(self, fn: Callable, index: int, exclusive: bool = False):
        self.fn = fn
        self.index = index
        self._exclusive = exclusive
        self.sig = inspect.signature(fn)
        try:
            self.hints = typing.get_type_hints(fn)
        except Exception:
            self.hints = {}
        self.file, self.line = typecheck.caller_location()
        self.__name__ = getattr(fn, "__name__", "func")
        self.__doc__ = fn.__doc__

    def num_in(self) -> int:
        return len(self.sig.parameters)

    def in_(self, i: int):
        p = list(self.sig.parameters.values())[i]
        return self.hints.get(p.name, p.annotation)

    NumIn, In = num_in, in_

    def exclusive(self) -> "FuncValue":
        """A copy whose invocations run on dedicated resources
        (``FuncValue.Exclusive``)."""
        v = FuncValue.__new__(FuncValue)
        v.__dict__.update(self.__dict__)
        v._exclusive = True
        return v

    Exclusive = exclusive

    def _typecheck(self, args: Sequence[Any]) -> None:
        params = list(self.sig.parameters.values())
        variadic = any(p.kind == p.VAR_POSITIONAL for p in params)
        fixed = [p for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        required = [p for p in fixed if p.default is p.empty]
        if (not variadic and len(args) > len(fixed)) or len(args) < len(required):
            typecheck.panicf("wrong number of arguments: function takes %d arguments, got %d",
                             len(fixed), len(args))
        for i, (p, a) in enumerate(zip(fixed, args)):
            _check_arg(i, self.hints.get(p.name, p.annotation), a)

    def invocation(self, location: str, *args) -> Invocation:
        """Typecheck ``args`` and mint a new invocation (``FuncValue.
        Invocation``)."""
        self._typecheck(args)
        with _inv_lock:
            idx = next(_inv_index)
        return Invocation(idx, self.index, args, self._exclusive, location)

    Invocation = invocation

    def apply(self, *args) -> Slice:
        """Call the function body, returning the slice it builds."""
        s = self.fn(*args)
        if not isinstance(s, Slice):
            typecheck.panicf("bigslice.Func: func must return a single bigslice.Slice")
        return s

    Apply = apply

    def __call__(self, *args) -> Slice:
        return self.apply(*args)

    def __repr__(self):
        return "bigslice.Func(%s)" % self.__name__


def Func(fn: Callable) -> FuncValue:
    """Register a slice-building function.

    Must be called at import/module-initialisation time, in the same order in
    every process of a session (multi-rank executors verify the registry
    matches across ranks)."""
    global _funcs_busy
    if not callable(fn):
        typecheck.panicf("bigslice.Func: argument to func is a %s, not a func", type(fn).__name__)
    _funcs_busy += 1
    if _funcs_busy != 1:
        _funcs_busy -= 1
        raise RuntimeError("bigslice.Func: data race")
    try:
        with _funcs_lock:
            v = FuncValue(fn, len(_funcs))
            _funcs.append(v)
    finally:
        _funcs_busy -= 1
    return v





def func_by_index(i: int) -> FuncValue:
    return _funcs[i]


def func_locations() -> List[str]:
    """``file:line`` of every registered Func, in registration order; ranks
    compare these at session start (``exec/slicemachine.go:689-702``)."""
    return ["%s:%d" % (f.file, f.line) for f in _funcs]


FuncLocations = func_locations


def func_locations_diff(lhs: Sequence[str], rhs: Sequence[str]) -> List[str]:
    """Human-readable diff of two Func registries (``FuncLocationsDiff``)."""
    out = []
    for i in range(max(len(lhs), len(rhs))):
        a = lhs[i] if i < len(lhs) else None
        b = rhs[i] if i < len(rhs) else None
        if a == b:
            out.append(a)
        else:
            if a is not None:
                out.append("- " + a)
            if b is not None:
                out.append("+ " + b)
    return out if any(x.startswith(("- ", "+ ")) for x in out if x) else []
