"""Flagship workload: Reduce-by-key "wordcount" over <int64 key, int64 count>
rows with Zipf-distributed keys (BASELINE.json config 2; the shape of the
reference's tutorial wordcount, docs/index.md:164-278, and cmd/slicer reduce).

Provides the synthetic data generator, two source slices (rows resident in HBM;
rows streamed from pinned host memory with double-buffered H2D copies) and the
registered ``Func`` the benchmark runs through ``Session.run``.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .. import context as _context
from .. import slicetype
from ..frame import Frame
from ..func import Func
from ..ops import functors
from ..reduce import Reduce
from ..sliceio import FuncReader
from ..slices import Slice, make_name
from ..slicetype import Type

ROW_TYPE = Type([slicetype.int64, slicetype.int64], 1)


def zipf_keys(n: int, vocab: int, s: float = 1.0, seed: int = 0, device="cpu") -> torch.Tensor:
    """``n`` int64 keys drawn from Zipf(s) over ``vocab`` distinct words.

    Ranks are mapped to scattered 63-bit key values (multiplication by an odd
    constant) so keys are not small consecutive integers."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    w = torch.arange(1, vocab + 1, dtype=torch.float64, device=device).pow_(-s)
    cdf = torch.cumsum(w, 0)
    cdf /= cdf[-1].clone()
    out = torch.empty(n, dtype=torch.int64, device=device)
    step = 1 << 24
    for i in range(0, n, step):
        m = min(step, n - i)
        u = torch.rand(m, dtype=torch.float64, device=device, generator=g)
        r = torch.searchsorted(cdf, u).clamp_(max=vocab - 1)
        out[i:i + m] = (r * 0x9E3779B97F4A7C15) & 0x7FFFFFFFFFFFFFFF
    return out


class DeviceRows(Slice):
    """Source slice over per-shard column tensors already resident in HBM."""

    def __init__(self, shards: List[Tuple[torch.Tensor, ...]], typ: Type = ROW_TYPE):
        super().__init__(make_name("devicerows"), typ, len(shards))
        self.shards = shards

    def rows_hint(self, shard: int) -> int:
        return int(self.shards[shard][0].shape[0])

    def reader(self, shard, deps):
        done = [False]

        def read(n):
            if done[0]:
                return None
            done[0] = True
            return Frame(list(self.shards[shard]), self.typ)
        return FuncReader(read)


class PinnedRows(Slice):
    """Source slice over per-shard columns in *pinned host memory*.

    Every evaluation copies the shard host->device in ``chunk_rows`` pieces on a
    dedicated copy stream into a small ring of persistent device staging buffers:
    the H2D copy of chunk i+1 overlaps the consumer's kernels on chunk i, nothing
    is allocated per step, and the consumer waits on CUDA events, never on the
    host.  A staging buffer is reused once the consumer's work on it has been
    *enqueued* and finished (an event recorded on the consumer's stream when it
    asks for the next chunk)."""

    RING = 3

    def __init__(self, shards: List[Tuple[torch.Tensor, ...]], typ: Type = ROW_TYPE,
                 chunk_rows: int = 1 << 24):
        super().__init__(make_name("pinnedrows"), typ, len(shards))
        self.shards = shards
        self.chunk_rows = chunk_rows
        self.h2d_bytes = 0
        self._state = {}   # (shard, device) -> [copy stream, staging buffers, consumed events]

    def rows_hint(self, shard: int) -> int:
        return int(self.shards[shard][0].shape[0])

    def _staging(self, shard: int, dev, cols):
        key = (shard, str(dev))
        st = self._state.get(key)
        if st is None:
            stream = torch.cuda.Stream(device=dev)
            bufs = [[torch.empty((self.chunk_rows,) + tuple(c.shape[1:]), dtype=c.dtype, device=dev)
                     for c in cols] for _ in range(self.RING)]
            st = self._state[key] = [stream, bufs, [None] * self.RING, 0]
        return st

    def reader(self, shard, deps):
        cols = self.shards[shard]
        n = int(cols[0].shape[0]) if cols else 0
        pos = [0]
        dev = _context.current_device()
        on_gpu = dev is not None and dev.type == "cuda" and n > 0
        st = self._staging(shard, dev, cols) if on_gpu else None
        last_slot = [None]

        def read(k):
            if on_gpu and last_slot[0] is not None:
                # the consumer's kernels on the previous chunk are enqueued by now
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                st[2][last_slot[0]] = ev
                last_slot[0] = None
            if pos[0] >= n:
                return None
            j = min(n, pos[0] + min(k or self.chunk_rows, self.chunk_rows))
            if not on_gpu:
                out = [c[pos[0]:j] for c in cols]
            else:
                stream, bufs, consumed, turn = st
                slot = turn % self.RING
                st[3] = turn + 1
                with torch.cuda.stream(stream):
                    if consumed[slot] is not None:
                        stream.wait_event(consumed[slot])
                    out = []
                    for c, b in zip(cols, bufs[slot]):
                        d = b[:j - pos[0]]
                        d.copy_(c[pos[0]:j], non_blocking=True)
                        out.append(d)
                    ev = torch.cuda.Event()
                    ev.record(stream)
                torch.cuda.current_stream(dev).wait_event(ev)
                last_slot[0] = slot
            self.h2d_bytes += sum(t.numel() * t.element_size() for t in out)
            pos[0] = j
            return Frame(out, self.typ)
        return FuncReader(read)


_SOURCES = {}


def register_source(name: str, s: Slice) -> None:
    """Sources are process-local objects; Funcs look them up by name so the
    invocation arguments stay small and identical on every rank."""
    _SOURCES[name] = s


@Func
def wordcount(source: str):
    """Count occurrences per key: ``Reduce(rows, add)``."""
    return Reduce(_SOURCES[source], functors.add)


@Func
def wordcount_top(source: str, n: int):
    """Reduce then keep the ``n`` most frequent keys of every shard (the
    cmd/urls top-N shape)."""
    red = Reduce(_SOURCES[source], functors.add)

    def top(k, v):
        kk = min(n, v.numel())
        if v.is_cuda:
            from ..ops import gpu
            if gpu.available():
                idx = gpu.topk(v, kk)[1]  # radix select (csrc/misc.cu)
                return k[idx], v[idx]
        idx = torch.topk(v, kk).indices
        return k[idx], v[idx]
    from ..slices import Flatmap
    return Flatmap(red, top, out=[slicetype.int64, slicetype.int64])
