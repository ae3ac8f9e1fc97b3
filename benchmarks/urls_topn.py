#!/usr/bin/env python
"""BASELINE config 5: the cmd/urls shape — Flatmap -> Reduce -> top-N — through the
public API on device-resident rows.

Per GPU ``--rows`` synthetic URL rows (int64 url id = domain * PAGES + page, Zipf
domains).  One step = Session.run of

    Flatmap(url -> <domain, 1>)  ->  Reduce(add)  ->  per-shard top-N (radix select)

i.e. the per-rank streaming hash aggregation, the fused table exchange over NVLink,
combine on arrival and ``gpu.topk``.  Device-timed with CUDA events (perf_counter on
CPU), max over ranks; rows/s counts input URL rows.

    torchrun --nproc-per-node N benchmarks/urls_topn.py --rows 134217728 --top 100
    python benchmarks/urls_topn.py --device cpu --rows 200000        # plumbing check
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bigslice_b200 as bs  # noqa: E402
from bigslice_b200.exec.spmd import init_process_group  # noqa: E402
from bigslice_b200.models import wordcount as wc  # noqa: E402

PAGES = 4096
TOP = [100]


@bs.Func
def urls_topn():
    def to_domain(url):
        return url // PAGES, torch.ones_like(url)
    counts = bs.Reduce(bs.Flatmap(wc._SOURCES["urls"], to_domain, out=[bs.int64, bs.int64]), bs.ops.add)

    def top(dom, cnt):
        k = min(TOP[0], cnt.numel())
        if cnt.is_cuda:
            from bigslice_b200.ops import gpu
            idx = gpu.topk(cnt, k)[1]
        else:
            idx = torch.topk(cnt, k).indices
        return dom[idx], cnt[idx]
    return bs.Flatmap(counts, top, out=[bs.int64, bs.int64])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 27)
    ap.add_argument("--domains", type=int, default=1 << 22)
    ap.add_argument("--top", type=int, default=100)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--out", default="gpurun_out/urls_topn.json")
    a = ap.parse_args()
    TOP[0] = a.top
    cuda = a.device == "cuda"
    init_process_group("cuda" if cuda else "cpu")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    dom = wc.zipf_keys(a.rows, a.domains, 1.0, seed=11 + rank, device=dev) % a.domains  # ids in [0, domains)
    g = torch.Generator(device=dev)
    g.manual_seed(100 + rank)
    urls = dom * PAGES + torch.randint(0, PAGES, (a.rows,), device=dev, generator=g)
    typ = bs.Type([bs.int64], 1)
    empty = (torch.empty(0, dtype=torch.int64, device=dev),)
    wc.register_source("urls", wc.DeviceRows([(urls,) if r == rank else empty for r in range(world)], typ))
    sess = bs.start(bs.Spmd(device="cuda" if cuda else None, arena_bytes=2 << 30), bs.Parallelism(2))
    want_total = a.rows * world
    times = []
    for it in range(a.steps + 3):
        if cuda:
            torch.cuda.synchronize()
        dist.barrier()
        if cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        res = sess.run(urls_topn)
        if cuda:
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        else:
            ms = (time.perf_counter() - t0) * 1e3
        rows = res.rows()
        res.discard()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it >= 3:
            times.append(float(t[0]))
    # correctness of the last step: the global top domain's count, against a direct count
    best = max((c for _, c in rows), default=0)
    tb = torch.tensor([best], dtype=torch.int64, device=dev)
    dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    local = torch.bincount(dom, minlength=a.domains)
    dist.all_reduce(local)
    ok = int(tb[0]) == int(local.max())
    ms = sorted(times)[len(times) // 2]
    out = {"metric": "urls top-N rows/s (Flatmap -> Reduce -> topN)", "world": world, "rows_per_gpu": a.rows,
           "domains": a.domains, "top": a.top, "ms_per_step": ms, "rows_per_s": want_total / (ms / 1e3),
           "top_count_ok": ok, "device": a.device,
           "modes": sorted(set(x["mode"] for x in sess.executor.exchange_log))}
    if rank == 0:
        print(json.dumps(out))
        os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
        json.dump(out, open(a.out, "w"))
    sess.shutdown()
    dist.barrier()
    assert ok, "top count mismatch"


if __name__ == "__main__":
    main()
