"""Built-in associative combiners.

``Reduce``/``Fold`` accept arbitrary callables, but these named functors are
recognised by the engine and lowered to template-specialised sm_100a kernels
(hash-combine with ``atomicAdd``/``atomicMin``/... and the fused
combine+exchange path) instead of generic segmented tree reductions.

Each functor is also an ordinary vectorised python function so the same graph
runs on the CPU executor.

The reference calls the user's combiner reflectively per row
(``exec/combiner.go:149-174``, ``sortio/reader.go:97-105``); naming the common
associative operations is what lets the engine pick a kernel instead.
"""
from __future__ import annotations

import torch


def _tag(name):
    def deco(fn):
        fn.__bigslice_builtin__ = name
        fn.__bigslice_rowwise__ = False
        return fn
    return deco


@_tag("add")
def add(a, b):
    return a + b


@_tag("mul")
def mul(a, b):
    return a * b


@_tag("min")
def min(a, b):  # noqa: A001 - deliberate name
    if isinstance(a, torch.Tensor):
        return torch.minimum(a, b)
    return a if a <= b else b


@_tag("max")
def max(a, b):  # noqa: A001
    if isinstance(a, torch.Tensor):
        return torch.maximum(a, b)
    return a if a >= b else b


@_tag("and")
def band(a, b):
    return a & b


@_tag("or")
def bor(a, b):
    return a | b


@_tag("xor")
def bxor(a, b):
    return a ^ b


@_tag("first")
def first(a, b):
    return a


BUILTINS = {f.__bigslice_builtin__: f for f in (add, mul, min, max, band, bor, bxor, first)}

#: torch.scatter_reduce names for builtin combiners
SCATTER = {"add": "sum", "mul": "prod", "min": "amin", "max": "amax"}


def builtin_name(fn):
    return getattr(fn, "__bigslice_builtin__", None)
