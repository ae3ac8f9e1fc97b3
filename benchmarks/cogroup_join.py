#!/usr/bin/env python
"""BASELINE config 3: Cogroup two-table join through the public API.

Per GPU: a "large" table of --large rows and a "small" table of --small rows,
both (int64 key, 4 x float32 payload); keys of the small table are a subset of
the large table's key space.  One step = Session.run(Cogroup(large, small)):
two fused NVLink exchanges, per-shard radix sorts, run detection and CSR
assembly.  Device-timed (CUDA events, max over ranks); rows/s counts input rows
of both tables.

    torchrun --nproc-per-node N benchmarks/cogroup_join.py --large 67108864 --small 4194304
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bigslice_b200 as bs  # noqa: E402
from bigslice_b200.exec.spmd import init_process_group  # noqa: E402
from bigslice_b200.models import wordcount as wc  # noqa: E402

T = bs.Type([bs.int64, bs.Vec(bs.float32, 4)], 1)


@bs.Func
def join():
    return bs.Cogroup(wc._SOURCES["large"], wc._SOURCES["small"])


@bs.Func
def bjoin():
    """The broadcast-shaped variant: the small side is replicated through the
    switch (multimem.st), the large side is joined where it is."""
    return bs.BroadcastJoin(wc._SOURCES["large"], wc._SOURCES["small"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--large", type=int, default=1 << 26)
    ap.add_argument("--small", type=int, default=1 << 22)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--out", default="gpurun_out/cogroup.json")
    ap.add_argument("--broadcast", action="store_true", help="BroadcastJoin instead of Cogroup")
    a = ap.parse_args()
    init_process_group("cuda")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(rank)
    keyspace = a.large * world // 4
    lk = torch.randint(0, keyspace, (a.large,), device=dev, generator=g)
    lv = torch.rand(a.large, 4, device=dev, generator=g)
    sk = torch.randint(0, keyspace, (a.small,), device=dev, generator=g)
    sv = torch.rand(a.small, 4, device=dev, generator=g)
    e = lambda: (torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, 4, device=dev))  # noqa: E731
    wc.register_source("large", wc.DeviceRows([(lk, lv) if r == rank else e() for r in range(world)], T))
    wc.register_source("small", wc.DeviceRows([(sk, sv) if r == rank else e() for r in range(world)], T))
    arena = (a.large + a.small) * 24 * 2 + (256 << 20)
    sess = bs.start(bs.Spmd(device="cuda", shuffle="fused", arena_bytes=arena), bs.Parallelism(2))
    times = []
    nkeys = 0
    for it in range(a.steps + 1):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = sess.run(bjoin if a.broadcast else join)
        e1.record()
        torch.cuda.synchronize()
        nkeys = sum(len(f) for f in res.frames())
        res.discard()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if it >= 1:
            times.append(float(t[0]))
    ms = sorted(times)[len(times) // 2]
    tot = torch.tensor([nkeys], device=dev)
    dist.all_reduce(tot)
    out = {"metric": "BroadcastJoin rows/s" if a.broadcast else "Cogroup join rows/s", "world": world,
           "rows_large_per_gpu": a.large,
           "rows_small_per_gpu": a.small, "ms_per_step": ms,
           "rows_per_s": (a.large + a.small) * world / (ms / 1e3), "distinct_keys": int(tot[0]),
           "modes": sorted(set(x["mode"] for x in sess.executor.exchange_log))}
    if rank == 0:
        print(json.dumps(out))
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(out, open(a.out, "w"))
    sess.shutdown()
    dist.barrier()


if __name__ == "__main__":
    main()
