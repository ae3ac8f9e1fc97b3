"""Tiered spilling for device-resident runs: HBM -> pinned host -> NVMe.

A B200 holds 180 GB, so most shards never leave HBM; when a sort (or a Cogroup
input) exceeds its budget, its key-sorted runs are *parked* rather than dropped
to the slow path: first in pinned host memory (tier 1: a D2H copy at PCIe speed,
read back with asynchronous H2D copies, tile by tile), and only when the host
budget is used up in spill files on disk through the frame codec (tier 2, the
reference's only tier: ``sliceio/spiller.go:24-127``).

``TieredSpiller`` has the :class:`~bigslice_b200.sliceio.Spiller` interface
(``spill`` / ``closing_readers`` / ``cleanup``), so the external sort uses either.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch

from ..frame import Frame
from ..frame import columns as C
from .reader import Reader
from .spiller import Spiller

#: pinned host memory all tiered spillers of this process may hold
HOST_BUDGET = int(os.environ.get("BIGSLICE_SPILL_HOST_BYTES", 16 << 30))
_host_used = 0


class _HostRun:
    """One run in pinned host memory."""

    def __init__(self, f: Frame):
        self.typ = f.typ
        self.n = len(f)
        self.cols = []
        self.nbytes = 0
        for c in f.cols():
            if isinstance(c, torch.Tensor):
                h = torch.empty(c.shape, dtype=c.dtype, pin_memory=torch.cuda.is_available())
                h.copy_(c, non_blocking=False)
                self.cols.append(h)
                self.nbytes += h.numel() * h.element_size()
            else:
                self.cols.append(C.col_to(c, "cpu"))
                self.nbytes += C.col_nbytes(c)


class _HostRunReader(Reader):
    """Streams a pinned run back to ``device`` in tiles (async H2D on the current
    stream: the consumer's kernels are ordered after the copy)."""

    def __init__(self, run: _HostRun, device, tile: int = 1 << 22):
        self.run, self.device, self.tile = run, device, tile
        self.pos = 0

    def read(self, n=None):
        if self.run is None or self.pos >= self.run.n:
            self.run = None
            return None
        j = min(self.run.n, self.pos + (int(n) if n else self.tile))
        cols = []
        for c in self.run.cols:
            part = c[self.pos:j]
            if isinstance(part, torch.Tensor) and self.device is not None:
                part = part.to(self.device, non_blocking=True)
            cols.append(part)
        self.pos = j
        return Frame(cols, self.run.typ)

    def close(self):
        self.run = None


class TieredSpiller:
    def __init__(self, name: str = "spill", device=None, host_budget: Optional[int] = None):
        self.name, self.device = name, device
        self.host_budget = HOST_BUDGET if host_budget is None else host_budget
        self.runs: List[object] = []       # _HostRun | index into the file spiller
        self.files: Optional[Spiller] = None
        self.stats = {"host_runs": 0, "host_bytes": 0, "file_runs": 0, "file_bytes": 0}
        self._typ = None

    def spill(self, frame: Frame) -> int:
        """Park ``frame`` (a sorted run); returns the bytes it occupies."""
        global _host_used
        self._typ = frame.typ
        need = frame.nbytes()
        if _host_used + need <= self.host_budget:
            run = _HostRun(frame)
            _host_used += run.nbytes
            self.runs.append(run)
            self.stats["host_runs"] += 1
            self.stats["host_bytes"] += run.nbytes
            return run.nbytes
        if self.files is None:
            self.files = Spiller(self.name)
        nb = self.files.spill(frame.to("cpu") if frame.device.type != "cpu" else frame)
        self.runs.append(self.files.num_files() - 1)
        self.stats["file_runs"] += 1
        self.stats["file_bytes"] += nb
        return nb

    def closing_readers(self, device=None) -> List[Reader]:
        device = device if device is not None else self.device
        file_readers = self.files.closing_readers(device) if self.files is not None else []
        out = []
        for r in self.runs:
            out.append(_HostRunReader(r, device) if isinstance(r, _HostRun) else file_readers[r])
        return out

    readers = closing_readers

    def num_files(self) -> int:
        return len(self.runs)

    def cleanup(self) -> None:
        global _host_used
        for r in self.runs:
            if isinstance(r, _HostRun):
                _host_used -= r.nbytes
        self.runs = []
        if self.files is not None:
            self.files.cleanup()
            self.files = None

    Spill, Readers, ClosingReaders, Cleanup = spill, closing_readers, closing_readers, cleanup
