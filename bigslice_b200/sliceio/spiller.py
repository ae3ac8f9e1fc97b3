"""Spill frames to temporary files and stream them back.

Used by the external sorter and the spilling combiner when a working set
outgrows its HBM/host budget.  Each ``spill`` call writes one file (one codec
stream, batched in ``SpillBatchSize`` rows) under a private temp directory.

Capability parity: reference ``sliceio/spiller.go:24-127`` (NewSpiller, Spill,
Readers, ClosingReaders, Cleanup; 3-level random fan-out directories).
"""
from __future__ import annotations

import os
import random
import shutil
import tempfile
from typing import List

from ..frame import Frame
from ..internal import defaultsize
from .codec import Decoder, Encoder
from .reader import ClosingReader, Reader

#: Rows per encoded batch inside a spill file; 0 = use ``defaultsize.Chunk``.
SpillBatchSize = 0


class Spiller:
    def __init__(self, name: str = "spill"):
        self.dir = tempfile.mkdtemp(prefix="bigslice-%s-" % name)
        self._files: List[str] = []
        self._typ = None

    def spill(self, frame: Frame) -> int:
        """Write ``frame`` as a new spill file; returns bytes written."""
        self._typ = frame.typ
        sub = os.path.join(self.dir, *("%02x" % random.randrange(256) for _ in range(3)))
        os.makedirs(sub, exist_ok=True)
        path = os.path.join(sub, "%06d" % len(self._files))
        batch = SpillBatchSize or defaultsize.Chunk
        with open(path, "wb") as fh:
            enc = Encoder(fh)
            n = len(frame)
            if n == 0:
                enc.write(frame)
            for i in range(0, n, batch):
                enc.write(frame.slice(i, min(n, i + batch)))
            nbytes = enc.bytes_written
        self._files.append(path)
        return nbytes

    def _open(self, path: str, device=None) -> Decoder:
        fh = open(path, "rb")
        return Decoder(fh, self._typ, device, closer=fh.close)

    def readers(self, device=None) -> List[Reader]:
        """One reader per spill file, in spill order."""
        return [self._open(p, device) for p in self._files]

    def closing_readers(self, device=None) -> List[Reader]:
        """Like readers(), each closing its file at EOF/error."""
        return [ClosingReader(self._open(p, device)) for p in self._files]

    def num_files(self) -> int:
        return len(self._files)

    def cleanup(self) -> None:
        shutil.rmtree(self.dir, ignore_errors=True)
        self._files = []

    Spill, Readers, ClosingReaders, Cleanup = spill, readers, closing_readers, cleanup


def NewSpiller(name: str = "spill") -> Spiller:
    return Spiller(name)
