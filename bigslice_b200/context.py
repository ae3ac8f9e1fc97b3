"""Thread-local execution context.

Executors install a :class:`~bigslice_b200.slicefunc.Context` (metrics scope,
shard, target device) around every task run; combinator readers pick it up
here and pass it to user functions that ask for ``ctx`` (the analogue of the
``context.Context`` threaded through ``Reader.Read`` in the reference,
``exec/local.go:80-82``).
"""
from __future__ import annotations

import contextlib
import threading

from .slicefunc import Context, background

_tls = threading.local()


def current() -> Context:
    return getattr(_tls, "ctx", None) or background()


@contextlib.contextmanager
def use(ctx: Context):
    prev = getattr(_tls, "ctx", None)
    _tls.ctx = ctx
    try:
        yield ctx
    finally:
        _tls.ctx = prev


def current_device():
    return current().device
