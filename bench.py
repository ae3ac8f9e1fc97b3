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
    ap.add_argument("--no-baseline", action="store_true", help="skip the library (NCCL / torch) arm")
    ap.add_argument("--no-proxy", action="store_true", help="skip the CPU Local-executor proxy")
    ap.add_argument("--baseline-steps", type=int, default=3)
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


def perkey_check(res, keys, world, rank, S, dev):
    """Exact per-key comparison of one step's result with a reference made of
    library pieces only (see the call site)."""
    import torch
    import torch.distributed as dist
    from bigslice_b200.frame import hashing
    nparts = S * world
    uk, cnt = torch.unique(keys, return_counts=True)
    dest = (hashing.hash_tensor(uk, 0) % nparts) % world          # pure-torch murmur3
    order = torch.sort(dest, stable=True)[1]
    uk, cnt, dest = uk[order], cnt[order], dest[order]
    send = torch.bincount(dest, minlength=world)
    recv = torch.empty_like(send)
    if world > 1:
        dist.all_to_all_single(recv, send)
    else:
        recv = send.clone()
    rk = torch.empty(int(recv.sum()), dtype=uk.dtype, device=dev)
    rc = torch.empty(int(recv.sum()), dtype=cnt.dtype, device=dev)
    if world > 1:
        dist.all_to_all_single(rk, uk, recv.tolist(), send.tolist())
        dist.all_to_all_single(rc, cnt, recv.tolist(), send.tolist())
    else:
        rk, rc = uk, cnt
    wk, inv = torch.unique(rk, return_inverse=True)
    wc_ = torch.zeros(wk.numel(), dtype=torch.int64, device=dev).index_add_(0, inv, rc)
    f = res.to_frame()
    gk, gc = f.col(0), f.col(1)
    o = torch.sort(gk)[1]
    ok = gk.numel() == wk.numel() and torch.equal(gk[o], wk) and torch.equal(gc[o], wc_)
    res.discard()
    t = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t[0]))


def library_baseline(args, keys, vals, world, rank, dev, dist):
    """The same Reduce step built from library calls only, device-timed the same
    way.  N>1: the framework's own combine kernels with the rows shuffled by NCCL
    ``all_to_all_single`` (what `--shuffle nccl-table` runs: compaction, hashing,
    local partition, three collectives, combine on arrival).  N=1 (no shuffle):
    ``torch.unique(return_inverse) + index_add_`` (CUB radix sort underneath)."""
    import torch
    from bigslice_b200.ops import gpu
    from bigslice_b200.parallel.peer import collective_exchange_tensors
    n = keys.numel()

    def one():
        if world == 1:
            uk, inv = torch.unique(keys, return_inverse=True)
            return uk, torch.zeros(uk.numel(), dtype=vals.dtype, device=dev).index_add_(0, inv, vals)
        k, v = gpu.hash_combine(keys, vals, "add", distinct_hint=1 << 24)
        dest = gpu.hash_column(k) % world
        k, v, _ = collective_exchange_tensors([k, v, dest], dest, None)
        return gpu.hash_combine(k, v, "add", distinct_hint=k.numel())
    for _ in range(2):
        one()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.baseline_steps):
        one()
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms[0]) / args.baseline_steps
    what = ("torch.unique(return_inverse)+index_add_ (CUB sort) on one GPU" if world == 1 else
            "hash_combine kernels + NCCL all_to_all_single shuffle of the combined rows (nccl-table)")
    return {"value": n * world / (ms / 1e3), "unit": "rows/s", "ms_per_step": ms,
            "steps": args.baseline_steps, "what": what}


def cpu_proxy(args, rows=1 << 21):
    """NOT the reference: the repo's own CPU `Local` executor (C++ host combiner,
    8 shards) on a 2^21-row sample of the same distribution — a labelled CPU
    anchor, as BASELINE.md section 2 permits when the Go reference cannot run."""
    import time as _t
    import torch
    import bigslice_b200 as bs
    from bigslice_b200.models import wordcount as wc
    try:
        k = wc.zipf_keys(rows, args.vocab, args.zipf, seed=99)
        v = torch.ones(rows, dtype=torch.int64)
        nsh = 8

        @bs.Func
        def job():
            return bs.Reduce(bs.Const(nsh, k, v), bs.ops.add)
        with bs.start(bs.Local, bs.Parallelism(8)) as sess:
            sess.run(job).discard()
            t0 = _t.perf_counter()
            sess.run(job).discard()
            dt = _t.perf_counter() - t0
        return {"rows": rows, "rows_per_s": rows / dt, "seconds": dt,
                "what": "bigslice_b200 Local executor on CPU (8 shards, 8 procs), labelled proxy - "
                        "not the Go reference"}
    except Exception as e:  # the proxy must never break the benchmark
        return {"error": str(e)[:200]}


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
        """Distinct keys and total count of the step's result (device-side sums,
        two scalars read back): used by the warm-up checks."""
        d, t = 0, 0
        for f in res.frames():
            d += len(f)
            t += int(f.col(1).sum().item())
        return d, t

    host_out = {}

    def fetch(res):
        """End-to-end read of the step's result: every result frame is copied
        device->host into pinned buffers (the call a user makes to consume the
        counts).  Returns (distinct keys, total count, bytes copied)."""
        d, t, nbytes = 0, 0, 0
        for i, f in enumerate(res.frames()):
            n = len(f)
            for c, col in enumerate(f.cols()):
                key = (i, c)
                hb = host_out.get(key)
                if hb is None or hb.numel() < n:
                    hb = host_out[key] = torch.empty(max(n, 1) * 5 // 4, dtype=col.dtype).pin_memory()
                hb[:n].copy_(col, non_blocking=True)
                nbytes += n * col.element_size()
            d += n
        torch.cuda.current_stream().synchronize()  # the rows are on the host now
        return d, t, nbytes

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
        host = []
        for _ in range(steps):
            h0 = time.perf_counter()
            res = step(source)
            if with_d2h:
                last = fetch(res)
            res.discard()
            host.append((time.perf_counter() - h0) * 1e3)
        ev1.record()
        if os.environ.get("BIGSLICE_SPANS") and rank == 0:
            sys.stderr.write("host ms per step (rank 0, %s): %s\n" % (source, " ".join("%.2f" % h for h in host)))
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
    # Per-key correctness at every N: the result must equal, key by key, an
    # independent reference built from library pieces only — torch.unique of the
    # rank-local rows, the framework-independent murmur3 placement (pure torch),
    # an NCCL all_to_all of the per-rank partials and a second torch.unique.
    perkey_ok = perkey_check(step("rows_hbm"), keys, world, rank, S, dev)
    assert perkey_ok, "per-key check failed on rank %d" % rank
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
            fetch(r)  # also allocates the pinned result buffers outside the timed region
            r.discard()
        ksteps = args.e2e_steps or min(args.steps, 5)
        src.h2d_bytes = 0
        ms2, _, last = timed("rows_pinned", ksteps, True)
        e2e = {"value": total_rows * ksteps / (ms2 / 1e3), "unit": "rows/s",
               "h2d_bytes_per_step": src.h2d_bytes // ksteps, "d2h_bytes_per_step": last[2],
               "d2h": "every result frame (keys + counts) copied to pinned host memory each step",
               "steps": ksteps, "ms_per_step": ms2 / ksteps, "distinct_keys_rank0": last[0]}

    # ---- library baseline: same job, shuffle through NCCL (--shuffle nccl-table) ----
    lib_base = None
    if args.shuffle == "fused" and not args.no_baseline:
        lib_base = library_baseline(args, keys, vals, world, rank, dev, dist)
    proxy = cpu_proxy(args) if (rank == 0 and not args.no_proxy) else None

    if rank == 0:
        out = {
            "metric": "rows/sec (whole box, device-timed, max over ranks) for Reduce-by-key wordcount",
            "value": value, "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": (value / lib_base["value"]) if lib_base else None,
            "baseline_def": ("BASELINE.md publishes no number and the Go reference cannot be built here; "
                             "vs_baseline = value / the library arm below (same rows, same step, measured in "
                             "this process): " + (lib_base["what"] if lib_base else "n/a")),
            "library_baseline": lib_base,
            "proxy_cpu_local": proxy,
            "perkey_check": "passed (torch.unique + NCCL all_to_all reference, every key and count)",
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
            "graphs": dict(ex.rx.stats) if getattr(ex, "rx", None) is not None else None,
        }
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    sess.shutdown()
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
