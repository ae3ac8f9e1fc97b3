"""Key-segment utilities over sorted frames: boundaries, segmented reduce/fold.

These are the device-agnostic (torch) forms; on B200 the builtin combiners are
served by the sm_100a segmented-reduce kernel in ``csrc/segreduce.cu`` and the
hash-combine kernels, selected in :mod:`bigslice_b200.ops.gpu`.

Replaces the reference's per-row heap pops + reflective combine
(``sortio/reader.go:88-127``) and ``map[key]acc`` fold (``slice.go:920-933``)
with whole-batch segmented operations.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from ..frame import Frame
from ..frame import columns as C
from .functors import SCATTER, builtin_name


import os as _os
_OWN_SEGREDUCE = bool(_os.environ.get("BIGSLICE_OWN_SEGREDUCE"))


def _neq_prev(c) -> torch.Tensor:
    """mask[i] = (c[i] != c[i-1]) for i>=1, mask[0] = True (on col's device)."""
    n = C.col_len(c)
    if isinstance(c, torch.Tensor):
        m = torch.ones(n, dtype=torch.bool, device=c.device)
        if n > 1:
            d = c[1:] != c[:-1]
            if d.dim() == 2:
                d = d.any(1)
            m[1:] = d
        return m
    m = np.ones(n, dtype=bool)
    if n > 1:
        m[1:] = c[1:] != c[:-1]
    return torch.from_numpy(m)


def group_starts(sorted_frame: Frame, nkeys: Optional[int] = None) -> torch.Tensor:
    """Boolean mask marking the first row of every key group."""
    nkeys = sorted_frame.prefix() if nkeys is None else nkeys
    n = len(sorted_frame)
    dev = sorted_frame.device
    m = None
    for k in range(nkeys):
        mk = _neq_prev(sorted_frame.col(k)).to(dev)
        m = mk if m is None else (m | mk)
    if m is None:
        m = torch.zeros(n, dtype=torch.bool, device=dev)
        if n:
            m[0] = True
    return m


def segmented_reduce(values, starts: torch.Tensor, fn: Callable,
                     rowwise: bool = False, ctx=None, func=None):
    """Reduce ``values`` within segments delimited by ``starts``.

    ``fn(a, b) -> c`` must be associative and commutative (Reduce's contract,
    reference ``reduce.go:19-21``).  Returns one value per segment.
    """
    n = C.col_len(values)
    if n == 0:
        return values
    seg = torch.cumsum(starts.to(torch.int64), 0) - 1
    nseg = int(seg[-1]) + 1
    name = builtin_name(fn) or (func.builtin if func is not None else None)
    if isinstance(values, torch.Tensor) and not rowwise:
        seg = seg.to(values.device)
        if name in SCATTER and values.dtype not in (torch.bool,) and values.dim() == 1:
            if values.is_cuda and _OWN_SEGREDUCE:
                # opt-in: the in-tree kernel measured 0.56x of index_add_ on B200
                # (profiles/r2_kernels_vs_torch.jsonl), so the library is the default
                from . import gpu
                if gpu.available() and values.dtype in gpu.SEGREDUCE_DTYPES:
                    return gpu.segmented_reduce(values, starts.to(values.device), nseg, name)
            if name == "add":
                out = torch.zeros(nseg, dtype=values.dtype, device=values.device)
                return out.index_add_(0, seg, values)
            out = torch.empty(nseg, dtype=values.dtype, device=values.device)
            return out.scatter_reduce_(0, seg, values, SCATTER[name], include_self=False)
        if name == "first":
            return values[starts.to(values.device)]
        return _tree_reduce(values, seg, nseg, fn, ctx, func)
    # host object values or per-row function: sequential fold per segment
    vals = C.col_tolist(values)
    segl = seg.tolist()
    out: List = [None] * nseg
    seen = [False] * nseg
    call = (lambda a, b: func.call_row(ctx, (a, b))) if func is not None else fn
    for v, s in zip(vals, segl):
        if seen[s]:
            out[s] = call(out[s], v)
        else:
            out[s] = v
            seen[s] = True
    if isinstance(values, torch.Tensor):
        return torch.tensor(out, dtype=values.dtype, device=values.device) if values.dim() == 1 \
            else torch.stack([torch.as_tensor(o) for o in out]).to(values.device)
    a = np.empty(nseg, dtype=object)
    for i, v in enumerate(out):
        a[i] = v
    return a


def _tree_reduce(values: torch.Tensor, seg: torch.Tensor, nseg: int, fn, ctx, func):
    """Pairwise in-segment tree reduction with an arbitrary vectorised fn."""
    vals, segs = values, seg
    call = (lambda a, b: func.call(ctx, [a, b])[0]) if func is not None else fn
    while vals.shape[0] > nseg:
        n = vals.shape[0]
        first = torch.ones(n, dtype=torch.bool, device=vals.device)
        first[1:] = segs[1:] != segs[:-1]
        idx = torch.arange(n, device=vals.device)
        start_idx = torch.cummax(torch.where(first, idx, torch.zeros_like(idx)), 0)[0]
        pos = idx - start_idx
        even = (pos % 2) == 0
        has_next = torch.zeros(n, dtype=torch.bool, device=vals.device)
        has_next[:-1] = segs[1:] == segs[:-1]
        pair = even & has_next
        pidx = torch.nonzero(pair).flatten()
        merged = call(vals[pidx], vals[pidx + 1])
        vals = vals.clone()
        vals[pidx] = merged.to(vals.dtype)
        keep = even
        vals, segs = vals[keep], segs[keep]
    return vals


def segmented_fold(acc_dtype, cols: Sequence, starts: torch.Tensor, func, ctx=None,
                   init=None):
    """Ordered per-segment fold: ``acc = fn(acc, v...)`` in row order.

    The accumulator starts at the zero value (reference ``slice.go:858-861``,
    ``accum.go``).  Vectorised functions are applied in "rounds": round ``r``
    folds the r-th row of every segment at once, so the cost is
    ``max_segment_len`` vectorised calls; per-row functions loop on the host.
    """
    n = C.col_len(cols[0]) if cols else 0
    dev = C.col_device(cols[0]) if cols else None
    seg = torch.cumsum(starts.to(torch.int64), 0) - 1
    nseg = (int(seg[-1]) + 1) if n else 0
    acc = C.col_zeros(acc_dtype, nseg, dev if acc_dtype.is_tensor else None)
    if n == 0:
        return acc
    if func.rowwise or not isinstance(acc, torch.Tensor):
        accl = C.col_tolist(acc)
        lists = [C.col_tolist(c) for c in cols]
        for i, s in enumerate(seg.tolist()):
            accl[s] = func.call_row(ctx, [accl[s]] + [l[i] for l in lists])
        return C.col_from_list(accl, acc_dtype, dev if acc_dtype.is_tensor else None)
    name = func.builtin
    if name in ("add",) and len(cols) == 1 and isinstance(cols[0], torch.Tensor) \
            and cols[0].dtype == acc.dtype:
        return acc.index_add_(0, seg.to(acc.device), cols[0])
    idx = torch.arange(n, device=seg.device)
    first_idx = torch.nonzero(starts.to(seg.device)).flatten()
    pos = idx - first_idx[seg]
    maxlen = int(pos.max()) + 1
    order = torch.sort(pos, stable=True)[1]
    counts = torch.bincount(pos, minlength=maxlen).tolist()
    off = 0
    for r in range(maxlen):
        rows = order[off:off + counts[r]]
        off += counts[r]
        s = seg[rows].to(acc.device)
        args = [acc[s]] + [C.col_take(c, rows) for c in cols]
        acc[s] = func.call(ctx, args)[0].to(acc.dtype)
    return acc
