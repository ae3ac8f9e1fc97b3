"""Text combinators on string columns: tokenising, URL hosts, dictionary ids.

These are ordinary ``Flatmap``/``Map`` stages whose batch functions run the
string kernels (``ops/strings.py``, ``csrc/strings.cu``) when the column is a
device :class:`~bigslice_b200.frame.StringColumn`, and plain python on host
object arrays — so the tutorial wordcount (``docs/index.md:164-278``) and the
``cmd/urls`` domain count (``cmd/urls/urls.go:37-126``) run with no per-row host
work on a B200:

    lines  --Tokenize-->  words  --StringIds-->  <id, 1>  --Reduce(add)-->  counts
                                        \\--> IdDictionary (id -> word, per rank, in HBM)

String *keys* travel as 64-bit dictionary ids (two seeded murmur3 hashes), which
is what lets them take the int64 hash-aggregation / fused NVLink exchange path;
every batch is verified (equal ids <=> equal strings), so a hash collision is a
loud error rather than two merged keys.
"""
from __future__ import annotations

import threading
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import context as _context
from . import errors, slicetype
from .frame.columns import StringColumn
from .ops import strings as S
from .slices import Flatmap, Map, Slice


def _as_strings(c, device=None) -> StringColumn:
    if isinstance(c, StringColumn):
        return c
    return S.from_objects(c, device=device)


def _like_input(out: StringColumn, src):
    """Host pipelines keep object arrays; device pipelines keep CSR columns."""
    return out if isinstance(src, StringColumn) and src.device.type != "cpu" else out.to_objects()


def ToDevice(slice_: Slice) -> Slice:
    """Move the (single) string column of ``slice_`` into HBM as a CSR column when
    the executor runs on a device (identity on CPU executors)."""
    def move(lines):
        dev = _context.current_device()
        if dev is None or dev.type == "cpu":
            return (lines,)
        return (_as_strings(lines).to(dev),)
    return Map(slice_, move, out=[slice_.out(0)])


def Tokenize(slice_: Slice, delims: bytes = S.WHITESPACE) -> Slice:
    """``<line>`` -> one row per token (split on ``delims``)."""
    def split(lines):
        col = _as_strings(lines)
        toks, _ = S.split(col, delims)
        return (_like_input(toks, lines),)
    return Flatmap(slice_, split, out=[slice_.out(0)])


def UrlDomains(slice_: Slice) -> Slice:
    """``<url>`` -> ``<host>`` (rows without a host are dropped)."""
    def hosts(urls):
        col = _as_strings(urls)
        dom = S.url_domain(col)
        keep = torch.nonzero(dom.lengths() > 0).flatten()
        if keep.numel() != len(dom):
            dom = S.take(dom, keep)
        return (_like_input(dom, urls),)
    return Flatmap(slice_, hosts, out=[slice_.out(0)])


class IdDictionary:
    """id -> string, accumulated per process while a pipeline runs (device
    resident when the strings are)."""

    def __init__(self):
        self._lock = threading.Lock()
        self._ids: List[torch.Tensor] = []
        self._strs: List[StringColumn] = []
        self._merged = None

    def add(self, ids: torch.Tensor, strs: StringColumn) -> None:
        with self._lock:
            self._ids.append(ids)
            self._strs.append(strs)
            self._merged = None

    def _merge(self):
        with self._lock:
            if self._merged is None and self._ids:
                from .frame import columns as C
                dev = self._strs[0].device
                ids = torch.cat([i.to(dev) for i in self._ids])
                strs = C.col_concat([s.to(dev) for s in self._strs])
                self._merged = S.dictionary(strs, ids)
                self._ids, self._strs = [self._merged[0]], [self._merged[1]]
            return self._merged

    def __len__(self) -> int:
        m = self._merge()
        return 0 if m is None else int(m[0].numel())

    def lookup(self, ids: Sequence[int]) -> Dict[int, Optional[str]]:
        """Strings of ``ids`` known to this process (others map to None)."""
        m = self._merge()
        out: Dict[int, Optional[str]] = {int(i): None for i in ids}
        if m is None or not out:
            return out
        d_ids, d_strs = m
        q = torch.tensor(list(out), dtype=torch.int64, device=d_ids.device)
        pos = torch.searchsorted(d_ids, q).clamp(max=d_ids.numel() - 1)
        hit = d_ids[pos] == q
        got = S.take(d_strs, pos[hit]).to_objects().tolist()
        for i, s in zip(q[hit].tolist(), got):
            out[i] = s
        return out

    def lookup_global(self, ids: Sequence[int], group=None) -> Dict[int, Optional[str]]:
        """Like :meth:`lookup`, asking every rank (collective; for the handful of
        ids of a top-N report)."""
        import torch.distributed as dist
        mine = self.lookup(ids)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return mine
        everyone: List = [None] * dist.get_world_size(group)
        dist.all_gather_object(everyone, {k: v for k, v in mine.items() if v is not None}, group=group)
        for d in everyone:
            for k, v in d.items():
                if mine.get(k) is None:
                    mine[k] = v
        return mine


def StringIds(slice_: Slice, dictionary: IdDictionary, verify: bool = True) -> Slice:
    """``<string>`` -> ``<id int64, 1 int64>``: the key becomes its 64-bit
    dictionary id, ``dictionary`` learns one representative string per id."""
    def encode(strs):
        col = _as_strings(strs)
        ids = S.key64(col)
        d_ids, d_strs = S.dictionary(col, ids)
        if verify and not S.verify_dictionary(col, ids, d_ids, d_strs):
            raise errors.fatal("text.StringIds: two different strings share a 64-bit id (hash collision)")
        dictionary.add(d_ids, d_strs)
        return ids, torch.ones_like(ids)
    return Map(slice_, encode, out=[slicetype.int64, slicetype.int64])
