"""Slices over tar archives.

``Reader(nshard, opener)`` returns a ``Slice<Entry>`` where every entry of the
archive opened by ``opener()`` becomes one row holding the header and the
file's bytes; entries are striped across shards (each shard scans the archive
and keeps entries ``shard, shard+nshard, ...``).

Capability parity: reference ``archive/tarslice/tarslice.go:17-84``.
"""
from __future__ import annotations

import tarfile
from dataclasses import dataclass
from typing import Callable

import numpy as np

from .. import slicetype, typecheck
from ..frame import Frame
from ..internal import defaultsize
from ..sliceio import FuncReader
from ..slices import Slice, make_name
from ..slicetype import Type


@dataclass
class Entry:
    """One archive member: its tar header and contents."""

    header: tarfile.TarInfo
    body: bytes

    @property
    def name(self) -> str:
        return self.header.name


class _TarSlice(Slice):
    def __init__(self, nshard: int, opener: Callable):
        super().__init__(make_name("tarslice"), Type([slicetype.obj]), nshard)
        self.opener = opener

    def reader(self, shard, deps):
        nshard = self._num_shard
        st = {"fh": None, "tar": None, "it": None, "i": 0, "done": False}

        def close():
            for k in ("tar", "fh"):
                o, st[k] = st[k], None
                if o is not None and hasattr(o, "close"):
                    o.close()

        def read(n):
            if st["done"]:
                return None
            if st["tar"] is None:
                st["fh"] = self.opener()
                st["tar"] = tarfile.open(fileobj=st["fh"], mode="r|*")
                st["it"] = iter(st["tar"])
            rows = n or min(defaultsize.Chunk, 1024)
            out = []
            for info in st["it"]:
                i = st["i"]
                st["i"] = i + 1
                if i % nshard != shard:
                    continue
                body = b""
                if info.isfile():
                    f = st["tar"].extractfile(info)
                    body = f.read() if f is not None else b""
                out.append(Entry(info, body))
                if len(out) >= rows:
                    break
            else:
                st["done"] = True
                close()
            if not out:
                return None
            a = np.empty(len(out), dtype=object)
            for j, e in enumerate(out):
                a[j] = e
            return Frame([a], self.typ)

        return FuncReader(read, close)


def Reader(nshard: int, opener: Callable) -> Slice:
    """A slice of :class:`Entry` over the tar stream returned by ``opener()``."""
    if not callable(opener):
        typecheck.panicf("tarslice: invalid opener %s", type(opener).__name__)
    if nshard < 1:
        typecheck.panic("tarslice: shard must be >= 1")
    return _TarSlice(nshard, opener)
