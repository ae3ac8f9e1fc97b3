"""Runtime type-checking helpers with source-located errors.

Combinators validate their arguments at graph-construction time and raise
:class:`TypecheckError` carrying the *user's* file:line (the first stack frame
outside this package), re-attributable to the ``Session.run`` call site with
:func:`location`.

Capability parity: reference ``typecheck/typecheck.go:16-57`` (Equal, Slices,
Devectorize), ``typecheck/func.go:13-43`` (CanApply), ``typecheck/error.go``
(Error, Panic/Panicf, Location, TestCalldepth).
"""
from __future__ import annotations

import contextlib
import os
import sys
from typing import Optional, Sequence, Tuple

from . import slicetype
from .slicetype import DType, Type

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))

#: Extra frames to skip when attributing errors (tests may tweak this, the
#: analogue of ``typecheck.TestCalldepth``).
TestCalldepth = 0

#: Code objects of user functions marked with ``@bigslice.helper``; their
#: frames are skipped when attributing names and errors (``bigslice.Helper``).
_helper_codes = set()


def mark_helper(fn):
    code = getattr(fn, "__code__", None)
    if code is not None:
        _helper_codes.add(code)
    return fn


class TypecheckError(TypeError):
    """A located type error (``typecheck.Error{Err,File,Line}``)."""

    def __init__(self, msg: str, file: str = "", line: int = 0):
        self.err = msg
        self.file = file
        self.line = line
        super().__init__(self._render())

    def _render(self) -> str:
        if self.file:
            return "%s:%d: %s" % (self.file, self.line, self.err)
        return self.err

    def relocate(self, file: str, line: int) -> "TypecheckError":
        return TypecheckError(self.err, file, line)


Error = TypecheckError


_inside_cache: dict = {}


def caller_location(skip_helpers: int = 0) -> Tuple[str, int]:
    """File and line of the innermost frame outside this package."""
    f = sys._getframe(1)
    skipped = 0
    while f is not None:
        code = f.f_code
        inside = _inside_cache.get(code)
        if inside is None:  # classified once per code object: this runs on every Session.run
            fn = os.path.abspath(code.co_filename)
            inside = fn.startswith(_PKG_DIR + os.sep) and \
                not fn.startswith(os.path.join(_PKG_DIR, "example") + os.sep) and \
                not fn.startswith(os.path.join(_PKG_DIR, "cmd") + os.sep) and \
                not fn.startswith(os.path.join(_PKG_DIR, "models") + os.sep)
            if len(_inside_cache) < 4096:
                _inside_cache[code] = inside
        helper = code in _helper_codes
        if not inside and not helper:
            if skipped >= skip_helpers + TestCalldepth:
                return f.f_code.co_filename, f.f_lineno
            skipped += 1
        f = f.f_back
    return "<unknown>", 0


def panicf(fmt: str, *args) -> None:
    """Raise a TypecheckError located at the user's call site."""
    msg = fmt % args if args else fmt
    file, line = caller_location()
    raise TypecheckError(msg, file, line)


def panic(msg: str) -> None:
    panicf("%s", msg)


Panicf, Panic = panicf, panic


@contextlib.contextmanager
def location(file: str, line: int):
    """Re-attribute TypecheckErrors raised inside to (file, line).

    Used by ``Session.run`` so errors in a Func body point at the run call
    (reference ``typecheck/error.go:60-79``, ``exec/session.go:276-280``).
    """
    try:
        yield
    except TypecheckError as e:
        raise e.relocate(file, line) from None


Location = location


def equal(a: Type, b: Type) -> bool:
    """Column-wise type equality (prefix ignored), ``typecheck.Equal``."""
    if a.num_out() != b.num_out():
        return False
    return all(x == y for x, y in zip(a.cols, b.cols))


Equal = equal


def slices(*values) -> Optional[Type]:
    """Type of a set of column values, or None if any is not a column.

    ``typecheck.Slices``: verifies every argument is a column and returns the
    element types.
    """
    cols = []
    for v in values:
        try:
            cols.append(slicetype.dtype_of_value(v))
        except TypeError:
            return None
    return Type(cols, 1)


Slices = slices


def devectorize(t: Type) -> Optional[Type]:
    """Strip one level of list from every column (``typecheck.Devectorize``)."""
    out = []
    for c in t.cols:
        if not c.is_list:
            return None
        out.append(c.child)
    return Type(out, t.prefix())


Devectorize = devectorize


def can_apply(fn_in: Sequence[DType], arg: Type, variadic: bool = False) -> bool:
    """Whether a function with parameter types ``fn_in`` accepts ``arg``.

    ``typecheck.CanApply`` including variadic final parameters: a parameter of
    kind ``object`` accepts anything.
    """
    fn_in = list(fn_in)
    n = arg.num_out()
    if variadic:
        if not fn_in:
            return False
        fixed, var = fn_in[:-1], fn_in[-1]
        if n < len(fixed):
            return False
        fn_in = fixed + [var] * (n - len(fixed))
    if len(fn_in) != n:
        return False
    for p, a in zip(fn_in, arg.cols):
        if p != a and p.kind != "object":
            return False
    return True


CanApply = can_apply
