"""Streaming frame readers.

The reader contract is batch-producing rather than buffer-filling:
``Reader.read(n)`` returns the next :class:`~bigslice_b200.frame.Frame` of at
most ``n`` rows (``n=None`` lets the reader pick its natural tile, typically a
whole shard resident in HBM) or ``None`` at end of stream.  Readers hand over
ownership of the returned frame (no reuse), are not safe for concurrent use,
and may return short or empty frames before EOF.

Capability parity: reference ``sliceio/reader.go:29-267`` (Reader contract,
EOF sentinel, MultiReader, FrameReader, ReadAll, ReadFull, ErrReader,
ReaderWithCloseFunc, ClosingReader, EmptyReader, PprofReader).
"""
from __future__ import annotations

import threading
from typing import Callable, Iterable, List, Optional, Sequence

from ..frame import Frame, concat
from ..slicetype import Type


class _EOF(Exception):
    """End-of-stream sentinel (``sliceio.EOF``); also usable as an exception."""

    def __repr__(self):
        return "EOF"


EOF = _EOF("EOF")


class Reader:
    """Base class / protocol for frame streams."""

    def read(self, n: Optional[int] = None) -> Optional[Frame]:
        raise NotImplementedError

    def close(self) -> None:
        pass

    def __iter__(self):
        while True:
            f = self.read(None)
            if f is None:
                return
            yield f

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class FuncReader(Reader):
    """Adapts ``fn(n) -> Frame|None`` (plus optional close) into a Reader."""

    def __init__(self, fn: Callable[[Optional[int]], Optional[Frame]],
                 close: Optional[Callable[[], None]] = None):
        self._fn = fn
        self._close = close

    def read(self, n=None):
        return self._fn(n)

    def close(self):
        if self._close is not None:
            c, self._close = self._close, None
            c()


class EmptyReader(Reader):
    """A reader that is immediately at EOF."""

    def read(self, n=None):
        return None


class ErrReader(Reader):
    """A reader that fails with ``err`` on every read (``sliceio.ErrReader``)."""

    def __init__(self, err: BaseException):
        self.err = err

    def read(self, n=None):
        raise self.err


class FrameReader(Reader):
    """Streams the rows of one in-memory frame (``sliceio.FrameReader``)."""

    def __init__(self, frame: Frame):
        self._frame = frame
        self._pos = 0

    def read(self, n=None):
        if self._frame is None:
            return None
        total = len(self._frame)
        if self._pos >= total:
            self._frame = None
            return None
        j = total if n is None else min(total, self._pos + max(int(n), 1))
        out = self._frame.slice(self._pos, j)
        self._pos = j
        return out


class MultiReader(Reader):
    """Concatenation of readers, closing each as it is exhausted.

    ``sliceio.MultiReader`` (reader.go:73-121): errors are sticky.
    """

    def __init__(self, readers: Sequence[Reader]):
        self._q: List[Reader] = list(readers)
        self._err: Optional[BaseException] = None

    def read(self, n=None):
        if self._err is not None:
            raise self._err
        while self._q:
            try:
                f = self._q[0].read(n)
            except BaseException as e:
                self._err = e
                raise
            if f is not None:
                return f
            self._q.pop(0).close()
        return None

    def close(self):
        q, self._q = self._q, []
        for r in q:
            r.close()


class ReaderWithCloseFunc(Reader):
    """Reader with an attached close callback (``ReaderWithCloseFunc``)."""

    def __init__(self, reader: Reader, close_func: Callable[[], None]):
        self.reader = reader
        self.close_func = close_func

    def read(self, n=None):
        return self.reader.read(n)

    def close(self):
        f, self.close_func = self.close_func, None
        try:
            self.reader.close()
        finally:
            if f is not None:
                f()


class ClosingReader(Reader):
    """Closes the underlying reader as soon as it reports EOF or an error."""

    def __init__(self, reader: Reader):
        self._r: Optional[Reader] = reader

    def read(self, n=None):
        if self._r is None:
            return None
        try:
            f = self._r.read(n)
        except BaseException:
            self.close()
            raise
        if f is None:
            self.close()
        return f

    def close(self):
        r, self._r = self._r, None
        if r is not None:
            r.close()


_label = threading.local()


def current_label() -> Optional[str]:
    """Profiling label of the op currently reading on this thread."""
    return getattr(_label, "name", None)


class LabelReader(Reader):
    """Tags reads with the op's name for profilers (analogue of PprofReader).

    The label is visible through :func:`current_label` and, on CUDA, as an
    NVTX range so ncu/nsys attribute kernels to ``<op>@file:line``.
    """

    def __init__(self, reader: Reader, label: str, nvtx: bool = False):
        self.reader = reader
        self.label = label
        self.nvtx = nvtx

    def read(self, n=None):
        prev = getattr(_label, "name", None)
        _label.name = self.label
        pushed = False
        if self.nvtx:
            try:
                import torch
                torch.cuda.nvtx.range_push(self.label)
                pushed = True
            except Exception:
                pushed = False
        try:
            return self.reader.read(n)
        finally:
            if pushed:
                import torch
                torch.cuda.nvtx.range_pop()
            _label.name = prev

    def close(self):
        self.reader.close()


PprofReader = LabelReader


def read_all(reader: Reader, typ: Optional[Type] = None, device=None) -> Frame:
    """Drain a reader into one frame (``sliceio.ReadAll``)."""
    frames = []
    while True:
        f = reader.read(None)
        if f is None:
            break
        if len(f):
            frames.append(f)
    reader.close()
    if not frames:
        if typ is None:
            raise ValueError("read_all: empty stream needs a type")
        return Frame.empty(typ, device)
    return concat(frames, typ)


def read_full(reader: Reader, n: int, typ: Optional[Type] = None) -> Optional[Frame]:
    """Read exactly ``n`` rows unless EOF intervenes (``sliceio.ReadFull``).

    Returns None if the stream was already at EOF.
    """
    frames, got = [], 0
    while got < n:
        f = reader.read(n - got)
        if f is None:
            break
        if len(f) > n - got:  # reader ignored the hint; keep the remainder out
            raise ValueError("read_full: reader returned more rows than requested")
        frames.append(f)
        got += len(f)
    if not frames:
        return None
    return concat(frames, typ)


class Rebatch(Reader):
    """Re-chunks a stream so every read returns at most ``n`` rows and no
    rows are lost when the producer ignores the size hint."""

    def __init__(self, reader: Reader):
        self.reader = reader
        self._pending: Optional[Frame] = None

    def read(self, n=None):
        f = self._pending
        self._pending = None
        if f is None:
            f = self.reader.read(n)
            if f is None:
                return None
        if n is not None and len(f) > n:
            self._pending = f.slice(n, len(f))
            f = f.slice(0, n)
        return f

    def close(self):
        self.reader.close()


class RetryReader(Reader):
    """Re-opens the stream at the current row offset when a read fails.

    ``opener(offset_rows) -> Reader``.  Failures are retried with exponential
    backoff up to ``retries`` times per incident (reference ``retryReader``,
    ``exec/bigmachine.go:1332-1402``: 5 tries, backoff x2).  Fatal errors and
    integrity errors are not retried."""

    def __init__(self, opener: Callable[[int], Reader], retries: int = 5,
                 backoff: float = 0.01, max_backoff: float = 1.0):
        self.opener = opener
        self.retries, self.backoff, self.max_backoff = retries, backoff, max_backoff
        self.offset = 0
        self.r: Optional[Reader] = None
        self.reopens = 0

    def read(self, n=None):
        import time
        from .. import errors
        tries, delay = 0, self.backoff
        while True:
            try:
                if self.r is None:
                    self.r = self.opener(self.offset)
                f = self.r.read(n)
                if f is not None:
                    self.offset += len(f)
                return f
            except BaseException as e:
                if errors.is_fatal(e) or errors.is_kind(e, errors.Kind.Integrity) \
                        or errors.is_kind(e, errors.Kind.NotExist) or isinstance(e, (KeyboardInterrupt,)):
                    raise
                tries += 1
                if tries > self.retries:
                    raise errors.Error("read failed after %d retries" % self.retries,
                                       errors.Kind.TooManyTries, cause=e)
                if self.r is not None:
                    try:
                        self.r.close()
                    except Exception:
                        pass
                    self.r = None
                self.reopens += 1
                time.sleep(delay)
                delay = min(delay * 2, self.max_backoff)

    def close(self):
        if self.r is not None:
            self.r.close()
            self.r = None


ReadAll, ReadFull = read_all, read_full
