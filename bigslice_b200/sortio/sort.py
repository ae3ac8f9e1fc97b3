"""Sorting and merging of frame streams.

* :func:`sort_reader` — external sort: reads a canary batch to estimate row
  size, sorts runs of ~``spill_target`` bytes (radix sort kernel on device
  frames), spills runs through :class:`~bigslice_b200.sliceio.Spiller` when
  more than one run is needed, then k-way merges the runs.
* :func:`merge_reader` — k-way merge of key-sorted streams.  Instead of one
  heap operation per row the merge is *batched*: every step computes the
  largest key bound that is safe to emit (the smallest last-key among the
  streams' buffered tiles), slices the ``<= bound`` prefix off every tile with
  a vectorised lexicographic compare and emits them as one sorted batch.

Capability parity: reference ``sortio/sort.go:31-222`` (SortReader,
FrameBuffer/FrameBufferHeap, NewMergeReader).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from ..frame import Frame, concat
from ..frame import columns as C
from ..internal import defaultsize
from ..sliceio import EmptyReader, FrameReader, Reader, Rebatch, Spiller, read_full
from ..slicetype import Type


def sort_frame(f: Frame) -> Frame:
    """Stable sort of rows by the prefix (key) columns."""
    return f.sorted()


def _le_bound(f: Frame, bound: Sequence, nkeys: int, strict: bool = False) -> torch.Tensor:
    """mask[i] = key(f[i]) <= bound (lexicographic; ``<`` when strict)."""
    n = len(f)
    res: Optional[np.ndarray | torch.Tensor] = None
    # Evaluate from the last key column to the first:
    #   le_k = lt_k | (eq_k & le_{k+1})
    for k in reversed(range(nkeys)):
        c = f.col(k)
        b = bound[k]
        if isinstance(c, torch.Tensor):
            if c.dim() == 2:
                bt = torch.as_tensor(b, dtype=c.dtype, device=c.device)
                lt = torch.zeros(n, dtype=torch.bool, device=c.device)
                eq = torch.ones(n, dtype=torch.bool, device=c.device)
                for j in range(c.shape[1]):
                    lt = lt | (eq & (c[:, j] < bt[j]))
                    eq = eq & (c[:, j] == bt[j])
            else:
                lt, eq = c < b, c == b
            lt, eq = lt.cpu(), eq.cpu()
        else:
            lt = torch.from_numpy(np.asarray(c < b, dtype=bool))
            eq = torch.from_numpy(np.asarray(c == b, dtype=bool))
        if res is None:
            res = lt if strict else (lt | eq)
        else:
            res = lt | (eq & res)
    if res is None:
        res = torch.ones(n, dtype=torch.bool)
    return res


def _key_at(f: Frame, i: int, nkeys: int) -> tuple:
    out = []
    for k in range(nkeys):
        c = f.col(k)
        if isinstance(c, torch.Tensor):
            v = c[i]
            out.append(v.item() if v.dim() == 0 else tuple(v.tolist()))
        else:
            out.append(c[i])
    return tuple(out)


class FrameBuffer:
    """A buffered tile of one sorted stream (``sortio.FrameBuffer``)."""

    def __init__(self, reader: Reader, index: int = 0, tile: Optional[int] = None):
        self.reader = reader
        self.index = index
        self.tile = tile
        self.frame: Optional[Frame] = None
        self.eof = False

    def fill(self) -> bool:
        """Ensure a non-empty tile is buffered; False once exhausted."""
        while (self.frame is None or len(self.frame) == 0) and not self.eof:
            f = self.reader.read(self.tile)
            if f is None:
                self.eof = True
                self.frame = None
                self.reader.close()
                return False
            self.frame = f
        return self.frame is not None and len(self.frame) > 0

    Fill = fill


def _device_key(k) -> bool:
    """A key column the merge-path kernel takes: int64, on a GPU with the kernel library
    (tests run the kernel under the CPU emulator and replace this predicate)."""
    if not (isinstance(k, torch.Tensor) and k.is_cuda and k.dtype == torch.int64 and k.dim() == 1):
        return False
    from ..ops import gpu
    return gpu.available()


class MergeReader(Reader):
    """Batched k-way merge of sorted streams."""

    def __init__(self, typ: Type, readers: Sequence[Reader], tile: Optional[int] = None):
        self.typ = typ
        self.nkeys = typ.prefix()
        self.bufs: List[FrameBuffer] = [FrameBuffer(r, i, tile) for i, r in enumerate(readers)]
        self._err: Optional[BaseException] = None
        self._first = True

    def read(self, n=None):
        if self._err is not None:
            raise self._err
        try:
            return self._step()
        except BaseException as e:
            self._err = e
            raise

    #: Streams are drained eagerly up to this many rows in total: a consumer of a
    #: wide shuffle merges hundreds of streams of a few rows each, and one sort of
    #: everything beats hundreds of incremental merge steps.
    EAGER_ROWS = 1 << 22

    def _drain_small(self) -> Optional[Frame]:
        """If all streams together hold at most EAGER_ROWS rows, read them
        completely and return their sorted union (the streams are then exhausted);
        otherwise leave what was read buffered and return None."""
        total, complete = 0, True
        for b in self.bufs:
            frames = [b.frame] if b.frame is not None and len(b.frame) else []
            total += sum(len(f) for f in frames)
            while not b.eof and total <= self.EAGER_ROWS:
                f = b.reader.read(b.tile)
                if f is None:
                    b.eof = True
                    b.reader.close()
                    break
                if len(f):
                    frames.append(f)
                    total += len(f)
            b.frame = (frames[0] if len(frames) == 1 else concat(frames, self.typ)) if frames else None
            complete = complete and b.eof
        if not complete:
            return None
        parts = [b.frame for b in self.bufs if b.frame is not None]
        self.bufs = []
        if not parts:
            return Frame.empty(self.typ)
        out = parts[0] if len(parts) == 1 else concat(parts, self.typ)
        return out.sorted() if len(parts) > 1 else out

    def _step(self) -> Optional[Frame]:
        if self._first:
            self._first = False
            if len(self.bufs) > 2:
                out = self._drain_small()
                if out is not None:
                    return out if len(out) else None
        live = [b for b in self.bufs if b.fill()]
        self.bufs = live
        if not live:
            return None
        if len(live) == 1:
            b = live[0]
            f, b.frame = b.frame, None
            return f
        # Safe bound: the smallest buffered last-key (Python tuple compare is
        # lexicographic); every unbuffered row of every stream is >= it.
        dev_out = self._device_step(live)
        if dev_out is not None:
            return dev_out
        bound = min(_key_at(b.frame, len(b.frame) - 1, self.nkeys) for b in live)
        parts = []
        for b in live:
            f = b.frame
            m = _le_bound(f, bound, self.nkeys)
            cnt = int(m.sum())
            if cnt:
                parts.append(f.slice(0, cnt))
                b.frame = f.slice(cnt, len(f)) if cnt < len(f) else None
        out = concat(parts, self.typ)
        return out.sorted()

    def _device_step(self, live) -> Optional[Frame]:
        """Device tiles keyed by one int64 column: the emit bound and the per-stream
        cuts are computed on the device (one searchsorted per stream instead of
        host-side masks) and the cuts are merged with the merge-path kernel
        (``csrc/merge.cu``) instead of being concatenated and sorted again."""
        if self.nkeys != 1:
            return None
        for b in live:
            if not _device_key(b.frame.col(0)):
                return None
        from ..ops import gpu
        lasts = torch.stack([b.frame.col(0)[-1] for b in live])
        bound = lasts.min()
        cuts = [int(torch.searchsorted(b.frame.col(0), bound, right=True)) for b in live]
        parts = []
        for b, cnt in zip(live, cuts):
            if cnt:
                parts.append(b.frame.slice(0, cnt))
                b.frame = b.frame.slice(cnt, len(b.frame)) if cnt < len(b.frame) else None
        while len(parts) > 1:   # pairwise tree of stable merges
            nxt = [gpu.merge_frames(parts[i], parts[i + 1]) for i in range(0, len(parts) - 1, 2)]
            if len(parts) % 2:
                nxt.append(parts[-1])
            parts = nxt
        return parts[0]

    def close(self):
        for b in self.bufs:
            b.reader.close()
        self.bufs = []


def merge_reader(typ: Type, readers: Sequence[Reader], tile: Optional[int] = None) -> Reader:
    """K-way merge of key-sorted readers (``sortio.NewMergeReader``)."""
    readers = list(readers)
    if not readers:
        return EmptyReader()
    return MergeReader(typ, readers, tile)


NewMergeReader = merge_reader


class _SortReader(Reader):
    def __init__(self, reader: Reader, spill_target: int, typ: Type, device=None):
        self.src = Rebatch(reader)
        self.spill_target = spill_target
        self.typ = typ
        self.device = device
        self._out: Optional[Reader] = None
        self._spiller: Optional[Spiller] = None
        self._err: Optional[BaseException] = None

    def _start(self):
        canary = read_full(self.src, defaultsize.SortCanary, self.typ)
        if canary is None:
            self._out = EmptyReader()
            return
        runs: List[Frame] = []
        nbytes = max(canary.nbytes(), 1)
        per_row = max(nbytes / max(len(canary), 1), 1e-9)
        run_rows = max(int(self.spill_target / per_row), defaultsize.SortCanary, 1)
        pending = [canary]
        pending_rows = len(canary)
        eof = False
        spiller: Optional[Spiller] = None
        first_run: Optional[Frame] = None
        first_dev = canary.device if canary.device.type == "cuda" else None
        nruns = 0
        while not eof or pending:
            while not eof and pending_rows < run_rows:
                f = self.src.read(run_rows - pending_rows)
                if f is None:
                    eof = True
                    break
                pending.append(f)
                pending_rows += len(f)
            if not pending:
                break
            run = concat(pending, self.typ).sorted()
            pending, pending_rows = [], 0
            nruns += 1
            if eof and nruns == 1:
                first_run = run
                break
            if spiller is None:
                # device runs are parked in pinned host memory first, files second
                if run.device.type == "cuda":
                    from ..sliceio.tiers import TieredSpiller
                    spiller = TieredSpiller("sort", run.device)
                else:
                    spiller = Spiller("sort")
            spiller.spill(run)
            # Re-estimate the run size from what we have observed (the
            # reference adjusts within +-5% of the spill target).
            per_row = max(run.nbytes() / max(len(run), 1), 1e-9)
            run_rows = max(int(self.spill_target / per_row), 1)
        self.src.close()
        if spiller is None:
            self._out = FrameReader(first_run) if first_run is not None else EmptyReader()
            return
        self._spiller = spiller
        dev = self.device if self.device is not None else (first_dev if first_dev is not None else None)
        self._out = merge_reader(self.typ, spiller.closing_readers(dev))

    def read(self, n=None):
        if self._err is not None:
            raise self._err
        try:
            if self._out is None:
                self._start()
            f = self._out.read(n)
        except BaseException as e:
            self._err = e
            self._cleanup()
            raise
        if f is None:
            self._cleanup()
        return f

    def _cleanup(self):
        if self._spiller is not None:
            self._spiller.cleanup()
            self._spiller = None

    def close(self):
        if self._out is not None:
            self._out.close()
        self._cleanup()


def sort_reader(reader: Reader, spill_target: int, typ: Type, device=None) -> Reader:
    """Externally sort a stream by its key prefix (``sortio.SortReader``)."""
    return _SortReader(reader, spill_target, typ, device)


SortReader = sort_reader
