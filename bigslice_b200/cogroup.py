"""Cogroup: n-ary generalized join / group-by.

For every distinct key (the shared key prefix) across the inputs the result
has one row ``<key..., [values of input 1]..., [values of input 2]...>``;
inputs that lack the key contribute empty lists.  The output is sorted by key
within each shard and its prefix is the key width.

Design (instead of the reference's per-row heap merge over externally sorted
inputs, ``cogroup.go:123-265``): every input shard is sorted by key (radix
sort on device), run boundaries give per-input unique keys and group sizes, a
tagged merge of the unique keys yields the key union, and each value column
becomes a CSR :class:`~bigslice_b200.frame.ListColumn` whose values are simply
the sorted input column — no per-row work, no copying of values beyond the
sort gather.

Capability parity: reference ``cogroup.go:46-112`` (typechecks, numShard =
max of inputs, shuffle deps) and ``:191-260`` (merge semantics).
"""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import context as _context
from . import slicetype, typecheck
from .frame import Frame, ListColumn, concat
from .frame import columns as C
from .frame import ops as frame_ops
from .ops.segmented import group_starts
from .sliceio import FuncReader, Reader, Rebatch, read_all
from .slices import Dep, Slice, _check_slice, make_name
from .slicetype import Type


def _device_join2(frames: Sequence[Frame], out_typ: Type, left: bool = False):
    """Two inputs, one int64 key, rows in HBM: radix sort + the sort-merge union
    kernels of ``csrc/cogroup.cu`` (run detection, binary-search merge, CSR
    sizes).  ``left``: only the keys of the first input (BroadcastJoin).  Returns
    None when the shape is not covered."""
    if len(frames) != 2 or out_typ.prefix() != 1:
        return None
    from .ops import gpu
    for f in frames:
        k = f.col(0) if f.num_out() else None
        if not (isinstance(k, torch.Tensor) and k.is_cuda and k.dtype == torch.int64 and k.dim() == 1):
            return None
    if not gpu.available():
        return None
    dev = frames[0].device
    srt, uniq, starts = [], [], []
    for f in frames:
        g = f.prefixed(1).sorted() if len(f) > 1 else f.prefixed(1)
        srt.append(g)
        k = g.col(0).contiguous()
        st = gpu.run_starts(k) if len(g) else torch.zeros(0, dtype=torch.int64, device=dev)
        starts.append(st)
        uniq.append(gpu.gather_rows(k, st) if len(g) else k)
    if left:
        return _left_join(srt, uniq, starts, out_typ, dev)
    uk, za, zb = gpu.cogroup_union2(uniq[0], starts[0], len(srt[0]), uniq[1], starts[1], len(srt[1]))
    if uk.numel() == 0:
        return Frame.empty(out_typ, dev)
    cols: List = [uk]
    for g, sizes in zip(srt, (za, zb)):
        offs = torch.zeros(uk.numel() + 1, dtype=torch.int64, device=dev)
        torch.cumsum(sizes, 0, out=offs[1:])
        for j in range(1, g.num_out()):
            cols.append(ListColumn(offs, g.col(j), g.out(j)))   # the sorted input column, zero copy
    return Frame(cols, out_typ)


_INT_KEYS = (torch.int64, torch.int32, torch.int16, torch.int8, torch.uint8)


def _on_device(t) -> bool:
    """A key column the cogroup kernels can take (tests run the kernels under the CPU
    emulator and replace this predicate)."""
    from .ops import gpu
    return isinstance(t, torch.Tensor) and t.is_cuda and gpu.available()


_NARROW_INT = (torch.int32, torch.int16, torch.int8, torch.uint8)


def _key_word(f: Frame, nkeys: int):
    """The key prefix of ``f`` as one int64 per row whose signed order is the
    lexicographic order of the key columns: one integer column of any width, or two
    integer columns of at most 32 bits (first column in the high half, second biased
    into the low half).  None when the prefix is not covered or not on the device."""
    ks = [f.col(i) for i in range(nkeys)] if f.num_out() >= nkeys else []
    if len(ks) != nkeys or not all(_on_device(k) and k.dim() == 1 for k in ks):
        return None
    if nkeys == 1 and ks[0].dtype in _INT_KEYS:
        return ks[0].to(torch.int64).contiguous()
    if nkeys == 2 and all(k.dtype in _NARROW_INT for k in ks):
        return ((ks[0].to(torch.int64) << 32) | ((ks[1].to(torch.int64) + (1 << 31)) & 0xFFFFFFFF)).contiguous()
    return None


def _key_columns(word: torch.Tensor, out_typ: Type, nkeys: int) -> List[torch.Tensor]:
    if nkeys == 1:
        return [word.to(out_typ.out(0).torch_dtype)]
    return [(word >> 32).to(out_typ.out(0).torch_dtype),
            ((word & 0xFFFFFFFF) - (1 << 31)).to(out_typ.out(1).torch_dtype)]


def _device_join_n(frames: Sequence[Frame], out_typ: Type, nkeys: int = 1):
    """Any number of inputs; a key prefix that maps to one order-preserving int64 word
    (:func:`_key_word`): the same kernels as :func:`_device_join2`, composed.  Every
    input is sorted by its key word (one radix sort) and its runs detected; the
    distinct-key lists are united pairwise (``cogroup_union2`` on keys only); each
    input's CSR group sizes come from one binary-search pass of the union against its
    distinct keys (``cogroup_search``).  Returns None when the shape is not covered."""
    if not frames or out_typ.prefix() != nkeys:
        return None
    from .ops import gpu
    words = [_key_word(f, nkeys) for f in frames]
    if any(w is None for w in words):
        return None
    dev = frames[0].device
    i64 = dict(dtype=torch.int64, device=dev)
    srt, uniq, starts = [], [], []
    for f, w in zip(frames, words):
        if len(f) > 1:
            w, perm = torch.sort(w, stable=True)
            f = Frame([C.col_take(c, perm) for c in f.cols()], f.typ)
        srt.append(f)
        st = gpu.run_starts(w) if len(f) else torch.zeros(0, **i64)
        starts.append(st)
        uniq.append(w[st] if len(f) else w)
    union = uniq[0]
    for u in uniq[1:]:
        if u.numel() == 0:
            continue
        if union.numel() == 0:
            union = u
            continue
        union, _, _ = gpu.cogroup_union2(union, torch.arange(union.numel(), **i64), union.numel(),
                                         u, torch.arange(u.numel(), **i64), u.numel())
    if union.numel() == 0:
        return Frame.empty(out_typ, dev)
    cols: List = _key_columns(union, out_typ, nkeys)
    for g, u, st in zip(srt, uniq, starts):
        if u.numel():
            rank, match = gpu.cogroup_search(union, u)
            run = torch.diff(torch.cat([st, torch.tensor([len(g)], **i64)]))
            sizes = torch.where(match.bool(), run[rank.clamp(max=u.numel() - 1)], torch.zeros_like(rank))
        else:
            sizes = torch.zeros(union.numel(), **i64)
        offs = torch.zeros(union.numel() + 1, **i64)
        torch.cumsum(sizes, 0, out=offs[1:])
        for j in range(nkeys, g.num_out()):
            cols.append(ListColumn(offs, g.col(j), g.out(j)))   # the sorted input column, zero copy
    return Frame(cols, out_typ)


def _left_join(srt, uniq, starts, out_typ: Type, dev) -> Frame:
    """Rows for the distinct keys of input 0 only.  The large side's value columns
    are its sorted columns (zero copy, CSR offsets = its run starts); of the small
    side only the rows whose key occurs on the large side are gathered."""
    from .ops import gpu
    ka, kb = uniq
    na, nb = len(srt[0]), len(srt[1])
    if ka.numel() == 0:
        return Frame.empty(out_typ, dev)
    ra, ma = gpu.cogroup_search(ka, kb)
    cols: List = [ka]
    offs_a = torch.cat([starts[0], torch.tensor([na], dtype=torch.int64, device=dev)])
    for j in range(1, srt[0].num_out()):
        cols.append(ListColumn(offs_a, srt[0].col(j), srt[0].out(j)))
    if kb.numel():
        run_b = torch.diff(torch.cat([starts[1], torch.tensor([nb], dtype=torch.int64, device=dev)]))
        sizes_b = torch.where(ma.bool(), run_b[ra.clamp(max=kb.numel() - 1)], torch.zeros_like(ra))
        _, mb = gpu.cogroup_search(kb, ka)
        keep = torch.nonzero(torch.repeat_interleave(mb.bool(), run_b)).flatten()
    else:
        sizes_b = torch.zeros(ka.numel(), dtype=torch.int64, device=dev)
        keep = torch.zeros(0, dtype=torch.int64, device=dev)
    offs_b = torch.zeros(ka.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(sizes_b, 0, out=offs_b[1:])
    for j in range(1, srt[1].num_out()):
        cols.append(ListColumn(offs_b, C.col_take(srt[1].col(j), keep), srt[1].out(j)))
    return Frame(cols, out_typ)


def cogroup_frames(frames: Sequence[Frame], nkeys: int, out_typ: Type) -> Frame:
    """Merge-join whole (unsorted) input shards into the cogrouped frame."""
    out = _device_join2(frames, out_typ) if nkeys == 1 else None
    if out is None:
        # n inputs / narrow or two-column integer keys: the same kernels, composed
        out = _device_join_n(frames, out_typ, nkeys)
    if out is not None:
        return out
    dev = None
    for f in frames:
        if len(f) and f.device.type != "cpu":
            dev = f.device
    sorted_in: List[Frame] = []
    uniq: List[Frame] = []
    counts: List[torch.Tensor] = []
    for f in frames:
        f = f.prefixed(nkeys).sorted()
        sorted_in.append(f)
        starts = group_starts(f, nkeys)
        idx = torch.nonzero(starts).flatten()
        key_typ = Type(f.typ.cols[:nkeys], nkeys)
        uniq.append(Frame([C.col_take(f.col(k), idx) for k in range(nkeys)], key_typ))
        n = len(f)
        if idx.numel():
            ends = torch.cat([idx[1:], torch.tensor([n], device=idx.device, dtype=idx.dtype)])
            counts.append(ends - idx)
        else:
            counts.append(torch.zeros(0, dtype=torch.int64, device=idx.device))
    key_typ = Type(out_typ.cols[:nkeys], nkeys)
    allk = concat([u for u in uniq], key_typ)
    nall = len(allk)
    if nall == 0:
        return Frame.empty(out_typ, dev)
    perm = allk.sort_indices()
    sk = allk.take(perm)
    gstarts = group_starts(sk, nkeys)
    gid_sorted = torch.cumsum(gstarts.to(torch.int64), 0) - 1
    nunion = int(gid_sorted[-1]) + 1
    gid = torch.empty(nall, dtype=torch.int64, device=gid_sorted.device)
    gid[perm.to(gid.device)] = gid_sorted
    uidx = torch.nonzero(gstarts).flatten()
    cols = [C.col_take(sk.col(k), uidx) for k in range(nkeys)]
    base = 0
    for f, u, cnt in zip(sorted_in, uniq, counts):
        g = gid[base:base + len(u)]
        base += len(u)
        cdev = cnt.device
        sizes = torch.zeros(nunion, dtype=torch.int64, device=cdev)
        if len(u):
            sizes[g.to(cdev)] = cnt
        offs = torch.zeros(nunion + 1, dtype=torch.int64, device=cdev)
        torch.cumsum(sizes, 0, out=offs[1:])
        for j in range(nkeys, f.num_out()):
            cols.append(ListColumn(offs, f.col(j), f.out(j)))
    return Frame(cols, out_typ)


def broadcast_partitioner(frame: Frame, nshard: int) -> torch.Tensor:
    """Marker partitioner of a *replicating* dependency: every row goes to every
    shard.  Executors recognise the function itself (a partitioner can only name
    one shard per row); calling it is an error."""
    raise RuntimeError("broadcast_partitioner marks a replicated dependency; it is not called")


class _BroadcastJoin(Slice):
    def __init__(self, large: Slice, small: Slice, typ: Type):
        super().__init__(make_name("broadcastjoin"), typ, large.num_shard(),
                         [Dep(large, shuffle=False), Dep(small, shuffle=True, partitioner=broadcast_partitioner)],
                         ())
        self.in_typs = [large.typ, small.typ]

    def reader(self, shard, deps: List[Reader]) -> Reader:
        done = [False]

        def read(n):
            if done[0]:
                return None
            done[0] = True
            dev = _context.current_device()
            frames = [read_all(r, t, dev) for r, t in zip(deps, self.in_typs)]
            if len(frames[0]) == 0:
                return None
            out = _device_join2(frames, self.typ, left=True)
            if out is not None:
                return out if len(out) else None
            out = cogroup_frames(frames, 1, self.typ)
            # left join: keep the keys that occur in this shard of the large side
            la = out.col(1)
            have = (la.offsets[1:] - la.offsets[:-1]) > 0
            if not bool(have.all()):
                out = out.take(torch.nonzero(have).flatten())
            return out if len(out) else None

        def close():
            for r in deps:
                r.close()
        return Rebatch(FuncReader(read, close))


def BroadcastJoin(large: Slice, small: Slice) -> Slice:
    """Broadcast-shaped join (an extension; the reference has only the shuffled
    ``Cogroup``): the *small* input is replicated on every shard — on B200 through
    NVSwitch multicast, one ``multimem.st`` per 16 bytes — and joined against the
    *large* input **where it already is** (no shuffle of the large side).

    Output, per shard of ``large``: one row ``<key, [large values], [small values]>``
    for every key of that shard (left join; ``[]`` when the small side lacks the
    key), sorted by key.  When ``large`` is partitioned by key (the output of a
    Reduce, Reshuffle, Cogroup…) this equals the rows of ``Cogroup(large, small)``
    that have a large side; otherwise a key that occurs in several shards of
    ``large`` yields one row per shard."""
    for i, s in enumerate((large, small)):
        _check_slice("broadcastjoin", s)
        if s.prefix() != 1 or s.num_out() < 1:
            typecheck.panicf("broadcastjoin: slice %d must have exactly one key column", i)
    if large.out(0) != small.out(0):
        typecheck.panicf("broadcastjoin: key column type mismatch: expected %s but got %s",
                         large.out(0), small.out(0))
    if not frame_ops.can_hash(large.out(0)) or not frame_ops.can_compare(large.out(0)):
        typecheck.panicf("broadcastjoin: key type %s cannot be hashed and sorted", large.out(0))
    cols = [large.out(0)]
    for s in (large, small):
        for j in range(1, s.num_out()):
            cols.append(slicetype.List(s.out(j)))
    return _BroadcastJoin(large, small, Type(cols, 1))


class _Cogroup(Slice):
    def __init__(self, slices_: Sequence[Slice], typ: Type, nshard: int, nkeys: int):
        super().__init__(make_name("cogroup"), typ, nshard,
                         [Dep(s, shuffle=True) for s in slices_], ())
        self.nkeys = nkeys
        self.in_typs = [s.typ for s in slices_]

    def reader(self, shard, deps: List[Reader]) -> Reader:
        done = [False]

        def read(n):
            if done[0]:
                return None
            done[0] = True
            dev = _context.current_device()
            frames = [read_all(r, t, dev) for r, t in zip(deps, self.in_typs)]
            out = cogroup_frames(frames, self.nkeys, self.typ)
            return out if len(out) else None

        def close():
            for r in deps:
                r.close()

        return Rebatch(FuncReader(read, close))


def Cogroup(*slices_: Slice) -> Slice:
    """Group the rows of all inputs by their (shared) key prefix."""
    if len(slices_) == 0:
        typecheck.panic("cogroup: expected at least one slice")
    key_types = None
    prefix = 0
    for i, s in enumerate(slices_):
        _check_slice("cogroup", s)
        if s.num_out() == 0:
            typecheck.panicf("cogroup: slice %d has no columns", i)
        if key_types is None:
            prefix = s.prefix()
            key_types = [s.out(k) for k in range(prefix)]
        else:
            if s.prefix() != prefix:
                typecheck.panicf("cogroup: prefix mismatch: expected %d but got %d",
                                 prefix, s.prefix())
            for k in range(prefix):
                if s.out(k) != key_types[k]:
                    typecheck.panicf("cogroup: key column type mismatch: expected %s but got %s",
                                     key_types[k], s.out(k))
    for i, t in enumerate(key_types):
        if not frame_ops.can_hash(t):
            typecheck.panicf("cogroup: key column(%d) type %s cannot be hashed", i, t)
        if not frame_ops.can_compare(t):
            typecheck.panicf("cogroup: key column(%d) type %s cannot be sorted", i, t)
    cols = list(key_types)
    for s in slices_:
        for j in range(prefix, s.num_out()):
            cols.append(slicetype.List(s.out(j)))
    nshard = max(s.num_shard() for s in slices_)
    return _Cogroup(slices_, Type(cols, prefix), nshard, prefix)
