#!/usr/bin/env python
"""Hand-written utility kernels vs their PyTorch/CUB equivalents (one GPU).

For every pair: same inputs, results compared, CUDA-event timed (best of 5 after
2 warm-ups; inputs far larger than the 126 MB L2).  The output (one JSON line per
pair) is committed under profiles/ and decides which implementation the
framework uses by default — a kernel that loses to the library is not kept on
the hot path.

    python benchmarks/kernels_vs_torch.py > gpurun_out/kernels_vs_torch.jsonl
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bigslice_b200.ops import gpu  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out


def report(name, n, ours_ms, lib_ms, ok, bytes_moved, note=""):
    print(json.dumps({"op": name, "n": n, "ours_ms": round(ours_ms, 4), "torch_ms": round(lib_ms, 4),
                      "speedup_vs_torch": round(lib_ms / ours_ms, 3), "equal": bool(ok),
                      "ours_GBps": round(bytes_moved / ours_ms / 1e6, 1), "note": note}), flush=True)


def main():
    assert gpu.available()
    dev = "cuda"
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    n = 1 << 26
    keys = torch.randint(-(1 << 62), 1 << 62, (n,), device=dev, generator=g)

    # radix sort of int64 keys with the sorting permutation as payload
    t_o, (ks, perm) = timed(lambda: gpu.sort_pairs_i64(keys, None))
    t_l, (ks2, perm2) = timed(lambda: torch.sort(keys, stable=True))
    report("sort_pairs_i64 (keys + permutation)", n, t_o, t_l, torch.equal(ks, ks2) and torch.equal(perm, perm2),
           n * 16 * 2 * 8, "torch.sort = CUB onesweep")

    # stream compaction: indices of set flags
    mask = torch.rand(n, device=dev, generator=g) < 0.3
    t_o, a = timed(lambda: gpu.compact_indices(mask))
    t_l, b = timed(lambda: torch.nonzero(mask).flatten())
    report("compact_indices (30% set)", n, t_o, t_l, torch.equal(a, b), n + a.numel() * 8, "torch.nonzero = CUB select")

    # gather rows by index (16-byte rows)
    rows = torch.randn(n // 4, 4, device=dev, generator=g)
    idx = torch.randint(0, n // 4, (n // 4,), device=dev, generator=g)
    t_o, a = timed(lambda: gpu.gather_rows(rows, idx))
    t_l, b = timed(lambda: rows.index_select(0, idx))
    report("gather_rows (16 B rows, random)", n // 4, t_o, t_l, torch.equal(a, b), (n // 4) * (8 + 32))
    col = torch.randint(0, 1 << 40, (n,), device=dev, generator=g)
    idx = torch.randint(0, n, (n,), device=dev, generator=g)
    t_o, a = timed(lambda: gpu.gather_rows(col, idx))
    t_l, b = timed(lambda: col.index_select(0, idx))
    report("gather_rows (8 B rows, random)", n, t_o, t_l, torch.equal(a, b), n * 24)

    # top-k (k = 1000) of int64 counts
    vals = torch.randint(0, 1 << 30, (n,), device=dev, generator=g)
    t_o, (v1, i1) = timed(lambda: gpu.topk(vals, 1000))
    t_l, (v2, i2) = timed(lambda: torch.topk(vals, 1000))
    report("topk k=1000", n, t_o, t_l, torch.equal(v1, v2), n * 8, "torch.topk = radix select + sort")

    # seeded murmur3 of an int64 column
    from bigslice_b200.frame import hashing
    t_o, a = timed(lambda: gpu.hash_column(keys))
    t_l, b = timed(lambda: hashing.hash_tensor(keys, 0))
    report("murmur3 hash_column", n, t_o, t_l, torch.equal(a, b), n * 16, "reference = the same hash in ~20 torch ops")

    # segmented sum over sorted keys
    k = torch.sort(torch.randint(0, 1 << 22, (n,), device=dev, generator=g))[0]
    v = torch.ones(n, dtype=torch.int64, device=dev)
    starts = torch.ones(n, dtype=torch.bool, device=dev)
    starts[1:] = k[1:] != k[:-1]
    nseg = int(starts.sum())
    seg = torch.cumsum(starts.to(torch.int64), 0) - 1
    t_o, a = timed(lambda: gpu.segmented_reduce(v, starts, nseg, "add"))
    t_l, b = timed(lambda: torch.zeros(nseg, dtype=torch.int64, device=dev).index_add_(0, seg, v))
    report("segmented_reduce add", n, t_o, t_l, torch.equal(a, b), n * 16, "torch index_add_ (atomics)")

    # hash aggregation vs sort-based unique+index_add
    from bigslice_b200.models import wordcount as wc
    zk = wc.zipf_keys(n, 1 << 23, 1.0, seed=5, device=dev)
    ones = torch.ones(n, dtype=torch.int64, device=dev)
    t_o, (uk, uv) = timed(lambda: gpu.hash_combine(zk, ones, "add", distinct_hint=1 << 23))

    def lib():
        u, inv = torch.unique(zk, return_inverse=True)
        return u, torch.zeros(u.numel(), dtype=torch.int64, device=dev).index_add_(0, inv, ones)
    t_l, (lk, lv) = timed(lib, reps=3, warm=1)
    o = torch.sort(uk)[1]
    report("hash_combine (Zipf, 2^26 rows)", n, t_o, t_l, torch.equal(uk[o], lk) and torch.equal(uv[o], lv),
           n * 16, "torch.unique(return_inverse) + index_add_")


if __name__ == "__main__":
    main()
