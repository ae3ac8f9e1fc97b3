"""ScanReader: a slice of text lines read from a stream.

Lines are striped across shards: every shard opens the stream and keeps lines
``shard, shard+nshard, ...`` (capability parity with reference
``scan.go:22-69``).  The column is a host string column; the tokenising /
hashing work downstream runs vectorised.
"""
from __future__ import annotations

import io
from typing import Callable

import numpy as np

from . import slicetype, typecheck
from .frame import Frame
from .internal import defaultsize
from .sliceio import FuncReader
from .slices import Slice, make_name
from .slicetype import Type


class _ScanReader(Slice):
    def __init__(self, nshard: int, opener: Callable):
        super().__init__(make_name("reader"), Type([slicetype.string]), nshard)
        self.opener = opener

    def reader(self, shard, deps):
        nshard = self._num_shard
        state = {"fh": None, "it": None, "line": 0, "done": False}

        def read(n):
            if state["done"]:
                return None
            if state["fh"] is None:
                fh = self.opener()
                if isinstance(fh, (bytes, str)):
                    fh = io.StringIO(fh if isinstance(fh, str) else fh.decode())
                state["fh"] = fh
                state["it"] = iter(fh)
            rows = n or defaultsize.Chunk
            out = []
            for raw in state["it"]:
                i = state["line"]
                state["line"] = i + 1
                if i % nshard != shard:
                    continue
                if isinstance(raw, bytes):
                    raw = raw.decode("utf-8", "replace")
                out.append(raw.rstrip("\n").rstrip("\r"))
                if len(out) >= rows:
                    break
            else:
                state["done"] = True
                close()
            if not out:
                return None
            a = np.empty(len(out), dtype=object)
            a[:] = out
            return Frame([a], self.typ)

        def close():
            fh, state["fh"] = state["fh"], None
            if fh is not None and hasattr(fh, "close"):
                fh.close()

        return FuncReader(read, close)


def ScanReader(nshard: int, reader: Callable) -> Slice:
    """A ``<string>`` slice of the lines of the stream returned by ``reader()``
    (a file-like object), striped over ``nshard`` shards."""
    if not callable(reader):
        typecheck.panicf("scanreader: invalid reader function %s", type(reader).__name__)
    if nshard < 1:
        typecheck.panic("scanreader: shard must be >= 1")
    return _ScanReader(nshard, reader)
