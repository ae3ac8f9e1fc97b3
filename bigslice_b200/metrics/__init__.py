"""User-defined metrics: counters aggregated in scopes.

Counters are declared at module initialisation (deterministic order, like
Funcs); user functions increment them through the task context
(``ctx.scope``) and the per-task scopes are merged into
``Result.scope()``.

Capability parity: reference ``metrics/metrics.go:33-126`` (NewCounter,
Counter.{Incr,Value}), ``metrics/scope.go:16-156`` (Scope.{Merge,Reset},
context plumbing, serialisable for transport between ranks).
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional

_lock = threading.Lock()
_counters: List["Counter"] = []


class Scope:
    """A set of metric instances (one int64 per registered counter)."""

    def __init__(self, values: Optional[Dict[int, int]] = None):
        self._lock = threading.Lock()
        self._values: Dict[int, int] = dict(values or {})

    def _incr(self, idx: int, n: int) -> None:
        with self._lock:
            self._values[idx] = self._values.get(idx, 0) + n

    def _value(self, idx: int) -> int:
        with self._lock:
            return self._values.get(idx, 0)

    def merge(self, other: "Scope") -> None:
        """Add ``other``'s values into this scope (``Scope.Merge``)."""
        with other._lock:
            items = list(other._values.items())
        with self._lock:
            for k, v in items:
                self._values[k] = self._values.get(k, 0) + v

    def reset(self, other: Optional["Scope"] = None) -> None:
        """Reset to ``other``'s values, or to zero (``Scope.Reset``)."""
        vals = {}
        if other is not None:
            with other._lock:
                vals = dict(other._values)
        with self._lock:
            self._values = vals

    def to_dict(self) -> Dict[int, int]:
        with self._lock:
            return dict(self._values)

    # picklable for transport between ranks (gob in the reference)
    def __getstate__(self):
        return self.to_dict()

    def __setstate__(self, st):
        self._lock = threading.Lock()
        self._values = dict(st)

    Merge, Reset = merge, reset


class Counter:
    """A simple integer counter metric."""

    def __init__(self, index: int, name: str = ""):
        self.index = index
        self.name = name

    def incr(self, scope_or_ctx, n: int = 1) -> None:
        scope = _scope_of(scope_or_ctx)
        if scope is not None:
            scope._incr(self.index, int(n))

    def value(self, scope_or_ctx) -> int:
        scope = _scope_of(scope_or_ctx)
        return scope._value(self.index) if scope is not None else 0

    Incr, Value = incr, value


def _scope_of(x) -> Optional[Scope]:
    if isinstance(x, Scope):
        return x
    return getattr(x, "scope", None)


def new_counter(name: str = "") -> Counter:
    """Register a new counter; call at module initialisation time."""
    with _lock:
        c = Counter(len(_counters), name)
        _counters.append(c)
    return c


NewCounter = new_counter


def scoped_context(ctx, scope: Scope):
    """A copy of ``ctx`` carrying ``scope`` (``metrics.ScopedContext``)."""
    from ..slicefunc import Context
    return Context(scope, getattr(ctx, "shard", 0), getattr(ctx, "device", None),
                   dict(getattr(ctx, "values", {}) or {}))


def context_scope(ctx) -> Optional[Scope]:
    return getattr(ctx, "scope", None)


ScopedContext, ContextScope = scoped_context, context_scope
