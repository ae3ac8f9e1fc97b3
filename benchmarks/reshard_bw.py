#!/usr/bin/env python
"""BASELINE config 4: Reshard shuffle-bandwidth sweep.

An 8 x float32 frame (32 B/row) is resharded W -> 2W partitions; per-GPU message
sizes sweep 1 MB .. 1 GB.  Times the fused NVLink exchange (csrc/exchange.cu)
and the NCCL all_to_all baseline with CUDA events (max over ranks) and reports
per-GPU egress bandwidth against the 900 GB/s/dir NVLink 5 nominal and the
770 GB/s measured peer-copy figure (B200_PROFILING.md).

    torchrun --nproc-per-node N benchmarks/reshard_bw.py [--max-mb 1024]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bigslice_b200.exec.spmd import init_process_group  # noqa: E402
from bigslice_b200.parallel.peer import PeerGroup, collective_exchange_tensors  # noqa: E402
from bigslice_b200.frame import hashing  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-mb", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/reshard_bw.jsonl")
    a = ap.parse_args()
    init_process_group("cuda")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    ctrl = dist.new_group(backend="gloo")
    peers = PeerGroup(ctrl, arena_bytes=int(a.max_mb * 1.35 * (1 << 20)) + (64 << 20), device=dev)
    nparts = 2 * world
    results = []
    mb = 1
    while mb <= a.max_mb:
        n = mb * (1 << 20) // 32
        g = torch.Generator(device=dev)
        g.manual_seed(rank * 1000 + mb)
        cols = [torch.rand(n, device=dev, generator=g) for _ in range(8)]
        row = {"mb_per_gpu": mb, "rows_per_gpu": n, "world": world, "nparts": nparts}
        # pre-partitioned copy for the "library all_to_all only" lower bound of the baseline
        from bigslice_b200.ops import gpu as _g
        _h = (_g.hash_column(cols[0], 0) % nparts) % world
        _order = torch.sort(_h, stable=True)[1]
        _cnt = torch.bincount(_h, minlength=world)
        _rc = torch.empty_like(_cnt)
        dist.all_to_all_single(_rc, _cnt)
        _sorted = [c.index_select(0, _order).contiguous() for c in cols]
        _in, _out = _cnt.tolist(), _rc.tolist()
        _recv = [torch.empty(int(_rc.sum()), dtype=c.dtype, device=dev) for c in cols]
        for mode in ("fused", "fused_cols", "nccl", "a2a_only"):
            times = []
            for it in range(a.iters + 2):
                torch.cuda.synchronize()
                dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                if mode == "fused":        # row records + TMA bulk stores (+ unpack on arrival)
                    got = peers.exchange_rows(cols, nparts, nkeys=1)
                    nrecv = sum(c[0].shape[0] for c in got.values())
                elif mode == "fused_cols":  # round-1 column mover (zero-copy views of the arena)
                    got = peers.exchange(cols, nparts, nkeys=1, zero_copy=True)
                    nrecv = sum(c[0].shape[0] for c in got.values())
                elif mode == "a2a_only":
                    for r_, s_ in zip(_recv, _sorted):
                        dist.all_to_all_single(r_, s_, _out, _in)
                    nrecv = _recv[0].shape[0]
                else:
                    from bigslice_b200.ops import gpu
                    h = gpu.hash_column(cols[0], 0) % nparts
                    got = collective_exchange_tensors(cols + [h], h % world, None)
                    nrecv = got[0].shape[0]
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if it >= 2:
                    times.append(float(t[0]))
            ms = sorted(times)[len(times) // 2]
            egress = n * 32 * (world - 1) / world
            row[mode + "_ms"] = ms
            row[mode + "_gbps_per_gpu"] = egress / (ms / 1e3) / 1e9 if world > 1 else n * 32 / (ms / 1e3) / 1e9
        row["fused_frac_of_770"] = row["fused_gbps_per_gpu"] / 770.0
        tot = torch.tensor([nrecv], device=dev)
        dist.all_reduce(tot)
        row["rows_ok"] = int(tot[0]) == n * world
        results.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        mb *= 4
    if rank == 0:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        with open(a.out, "w") as fh:
            for r in results:
                fh.write(json.dumps(r) + "\n")
    dist.barrier()


if __name__ == "__main__":
    main()
