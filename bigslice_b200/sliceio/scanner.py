"""Row/vector scanner over a frame stream — the user-facing result iterator.

``Scanner.scan()`` advances one row (``scanner.row()`` / the values returned
by iteration), ``Scanner.scanv()`` returns whole column batches.  Errors are
sticky and reported by ``err()`` like the reference; iteration raises them.

Capability parity: reference ``sliceio/scanner.go:27-141`` (Scan, Scanv, Err,
Close; arity misuse is a typecheck error).
"""
from __future__ import annotations

from typing import Iterator, List, Optional

from .. import typecheck
from ..frame import Frame
from ..frame import columns as C
from ..internal import defaultsize
from ..slicetype import Type
from .reader import Reader


class Scanner:
    def __init__(self, reader: Reader, typ: Type):
        self.reader: Optional[Reader] = reader
        self.typ = typ
        self._rows: List[tuple] = []
        self._pos = 0
        self._cur: Optional[tuple] = None
        self._err: Optional[BaseException] = None
        self._done = False

    # -- row API -------------------------------------------------------------
    def _fill(self) -> bool:
        while self._pos >= len(self._rows):
            if self._done or self._err is not None:
                return False
            try:
                f = self.reader.read(defaultsize.Chunk)
            except BaseException as e:
                self._err = e
                return False
            if f is None:
                self._done = True
                return False
            self._rows = f.to("cpu").rows() if f.device.type != "cpu" else f.rows()
            self._pos = 0
        return True

    def scan(self, nargs: Optional[int] = None) -> bool:
        """Advance to the next row; False at end of stream or on error.

        ``nargs`` lets callers assert the arity they expect (the reference
        checks the number/types of the pointers passed to ``Scan``).
        """
        if nargs is not None and nargs != self.typ.num_out():
            self._err = typecheck.TypecheckError(
                "wrong arity: expected %d columns, got %d" % (self.typ.num_out(), nargs),
                *typecheck.caller_location())
            return False
        if not self._fill():
            return False
        self._cur = self._rows[self._pos]
        self._pos += 1
        return True

    def row(self) -> tuple:
        return self._cur

    values = row

    def scanv(self, n: Optional[int] = None) -> Optional[list]:
        """Next batch as a list of columns (None at EOF / error)."""
        if self._err is not None or self._done:
            return None
        if self._pos < len(self._rows):
            raise RuntimeError("scanner: cannot mix scan and scanv mid-batch")
        try:
            f = self.reader.read(n)
        except BaseException as e:
            self._err = e
            return None
        if f is None:
            self._done = True
            return None
        return f.cols()

    def scan_frame(self, n: Optional[int] = None) -> Optional[Frame]:
        if self._err is not None or self._done:
            return None
        try:
            f = self.reader.read(n)
        except BaseException as e:
            self._err = e
            return None
        if f is None:
            self._done = True
        return f

    def err(self) -> Optional[BaseException]:
        return self._err

    def close(self) -> None:
        r, self.reader = self.reader, None
        if r is not None:
            r.close()

    Scan, Scanv, Err, Close = scan, scanv, err, close

    # -- python protocol -----------------------------------------------------
    def __iter__(self) -> Iterator[tuple]:
        while self.scan():
            yield self._cur
        if self._err is not None:
            raise self._err

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
