"""Operators: builtin combiner functors, segmented ops and the sm_100a kernel
bindings (:mod:`.gpu`)."""
from . import functors
from .functors import add, mul, band, bor, bxor, first
from .functors import min, max  # noqa: A004

__all__ = ["functors", "add", "mul", "min", "max", "band", "bor", "bxor", "first"]
