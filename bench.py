#!/usr/bin/env python
"""Headline benchmark: Reduce-by-key wordcount (BASELINE.json config 2).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One *step* is one complete ``Session.run`` of ``Reduce(rows, add)`` over
``rows_per_gpu * N`` synthetic <int64 key, int64 count> rows with Zipf(1.0)
keys over a 2^24-word vocabulary (weak scaling: 2^27 rows per GPU, 1.07 G rows
at N=8): per-rank TMA hash-aggregation, fused hash->partition->NVLink exchange
of the distinct keys, combine-on-arrival and compaction into the result frames.

* ``value``  rows/s with the input already resident in HBM, device-timed with
  CUDA events over exactly K steps, max over ranks.
* ``e2e``    the same job through the same public API, with every step's input
  copied from pinned host memory (double-buffered H2D) and a device->host read
  of the result summary (distinct keys, total count).

Rank 0 prints one JSON line.  ``--impl reference`` reports that the Go
reference cannot be built in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--rows-per-gpu", type=int, default=1 << 27)
    ap.add_argument("--vocab", type=int, default=1 << 24)
    ap.add_argument("--zipf", type=float, default=1.0)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--shuffle", default="fused",
                    help="fused (peer-memory kernels) | nccl-table (same combine kernels, rows through "
                         "the library all-to-all: the like-for-like baseline) | nccl (generic frames "
                         "through the library path)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--shards-per-gpu", type=int, default=1,
                    help="split every GPU's rows into this many shards (tasks); >1 also exercises "
                         "the fused exchange on a single GPU")
    return ap.parse_args()


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed
    region (profiling recipe's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


class NvmlSampler:
    """Polls SM clock and clock-event (throttle) reasons through NVML every few
    milliseconds from a thread, so that even a timed region of a few tens of
    milliseconds contains samples; only samples taken inside [t0, t1] count."""

    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
               ("sw_power_cap", 0x4))

    def __init__(self, index: int = 0, period_s: float = 0.004):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        except Exception:
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.period = period_s
        self.samples = []
        self._stop = threading.Event()
        self.t0 = self.t1 = None
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))

    def start(self):
        threading.Thread(target=self._run, daemon=True).start()

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), float(mhz), int(rs)))
            except Exception:
                pass
            self._stop.wait(self.period)

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        self._stop.set()
        inside = [s for s in self.samples if self.t0 is not None and self.t0 <= s[0] <= self.t1]
        note = "timed region"
        if not inside:  # cannot happen with a >4 ms region, but never report nothing
            inside, note = self.samples[-3:], "nearest samples (region shorter than the sampling period)"
        reasons = set()
        for _, _, rs in inside:
            for nm, bit in self.REASONS:
                if rs & bit:
                    reasons.add(nm)
        sm = [s[1] for s in inside]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvml", "window": note}


def make_sampler(index: int):
    try:
        return NvmlSampler(index)
    except Exception:
        return ClockSampler(index)


def reference_arm(args):
    why = ("grailbio/bigslice is a pure-Go module: no setup.py/pyproject.toml for pip to install "
           "(pip: 'not installable'), no Go toolchain or module cache in this image, no network")
    print(json.dumps({"impl": "reference", "unavailable": why}))


def main():
    args = parse()
    if args.impl == "reference":
        reference_arm(args)
        return 0
    # stdout carries exactly one JSON line: anything libraries print to fd 1
    # (e.g. NCCL's version banner) is sent to stderr instead.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        return run(args, os.fdopen(json_fd, "w"))
    finally:
        sys.stdout.flush()


def run(args, json_out):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        json_out.write(json.dumps({"error": "bench.py needs a CUDA device", "n_gpus": args.gpus}) + "\n")
        json_out.flush()
        return 1
    import bigslice_b200 as bs
    from bigslice_b200.exec.spmd import init_process_group
    from bigslice_b200.models import wordcount as wc
    from bigslice_b200.ops import gpu

    init_process_group("cuda")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    n = args.rows_per_gpu
    keys = wc.zipf_keys(n, args.vocab, args.zipf, seed=args.seed + rank, device=dev)
    vals = torch.ones(n, dtype=torch.int64, device=dev)
    # every rank registers the same source *names*; shard r holds rank r's rows
    empty = (torch.empty(0, dtype=torch.int64, device=dev),) * 2
    S = max(args.shards_per_gpu, 1)
    per = (n + S - 1) // S

    def split(cols, nothing):
        """Shard i lives on rank i % world: rank r's j-th piece is shard r + j*world."""
        out = []
        for i in range(S * world):
            j = i // world
            out.append(tuple(c[j * per:(j + 1) * per] for c in cols) if i % world == rank else nothing)
        return out
    shards_dev = split((keys, vals), empty)
    wc.register_source("rows_hbm", wc.DeviceRows(shards_dev))
    sess = bs.start(bs.Spmd(device="cuda", shuffle=args.shuffle, arena_bytes=2 << 30),
                    bs.Parallelism(2))
    ex = sess.executor

    def step(source):
        res = sess.run(wc.wordcount, source)
        return res

    def summarize(res):
        """Device->host read of the step's result: distinct keys, total count."""
        d, t = 0, 0
        for f in res.frames():
            d += len(f)
            t += int(f.col(1).sum().item())
        return d, t

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    def timed(source, steps, with_d2h):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = gpu.launches()
        ev0.record()
        last = None
        for _ in range(steps):
            res = step(source)
            if with_d2h:
                last = summarize(res)
            res.discard()
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), gpu.launches() - l0, last

    # ---- device-resident arm ----
    for _ in range(max(args.warmup, 3)):
        res = step("rows_hbm")
        chk = summarize(res)
        res.discard()
    tot = torch.tensor([chk[1]], dtype=torch.int64, device=dev)
    dist.all_reduce(tot)
    assert int(tot[0]) == n * world, "wordcount lost rows: %d != %d" % (int(tot[0]), n * world)
    sampler = make_sampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
        time.sleep(0.02)
    from bigslice_b200.utils import spans
    m0 = spans.mark()
    tw0 = time.perf_counter()
    ms, launches, _ = timed("rows_hbm", args.steps, False)
    if hasattr(sampler, "window"):
        sampler.window(tw0, time.perf_counter())
    if spans.enabled() and rank == 0:
        sys.stderr.write("host spans per step (rank 0):\n%s\n" % spans.format_summary(m0, args.steps))
        raw = spans.raw(m0)
        # timeline of the last step (ms relative to its Session.run start)
        starts = [t0 for (nm, th, t0, t1) in raw if nm == "run/compile"]
        if starts:
            z = starts[-1]
            sys.stderr.write("timeline of the last step (ms from run start):\n")
            for nm, th, t0, t1 in sorted(raw, key=lambda r: r[2]):
                if t0 >= z:
                    sys.stderr.write("  %7.3f .. %7.3f  %-26s %s\n" % ((t0 - z) * 1e3, (t1 - z) * 1e3, nm, th))
    clocks = sampler.stop() if rank == 0 else None
    total_rows = n * world
    value = total_rows * args.steps / (ms / 1e3)

    # ---- end-to-end arm: pinned host -> H2D each step, D2H of the result ----
    e2e = None
    if not args.no_e2e:
        hk = torch.empty(n, dtype=torch.int64).pin_memory()
        hv = torch.empty(n, dtype=torch.int64).pin_memory()
        hk.copy_(keys)
        hv.copy_(vals)
        hempty = (torch.empty(0, dtype=torch.int64),) * 2
        src = wc.PinnedRows(split((hk, hv), hempty))
        wc.register_source("rows_pinned", src)
        for _ in range(3):
            r = step("rows_pinned")
            summarize(r)
            r.discard()
        ksteps = args.e2e_steps or min(args.steps, 5)
        src.h2d_bytes = 0
        ms2, _, last = timed("rows_pinned", ksteps, True)
        e2e = {"value": total_rows * ksteps / (ms2 / 1e3), "unit": "rows/s",
               "h2d_bytes_per_step": src.h2d_bytes // ksteps, "d2h_bytes_per_step": 8 * 2,
               "steps": ksteps, "ms_per_step": ms2 / ksteps, "distinct_keys_rank0": last[0]}

    if rank == 0:
        out = {
            "metric": "rows/sec (whole box, device-timed, max over ranks) for Reduce-by-key wordcount",
            "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int64 keys/values (non-neural workload; no reduced precision)",
            "data": "synthetic: Zipf(s=%.2f) keys over %d-word vocabulary, counts=1, random seed per rank"
                    % (args.zipf, args.vocab),
            "config": {"model": "Reduce-by-key wordcount (int64,int64) rows, Zipf keys",
                       "global_rows": total_rows, "global_batch": total_rows, "rows_per_gpu": n,
                       "row_bytes": 16,
                       "parallelism": "dp%d (%d shard%s per GPU)" % (world, S, "" if S == 1 else "s"),
                       "shuffle": args.shuffle,
                       "l2_policy": "inputs (%.1f GB/GPU) larger than the 126 MB L2" % (n * 16 / 1e9)},
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": e2e,
            "exchange_modes": sorted(set(e["mode"] for e in ex.exchange_log)),
        }
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    sess.shutdown()
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
