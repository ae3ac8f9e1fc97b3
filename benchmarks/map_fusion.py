#!/usr/bin/env python
"""Python Map → Filter → Map pipeline: traced + fused into one kernel vs eager PyTorch.

The same three lambdas are evaluated over one resident tile (2^26 rows, two
columns — inputs 4× the 126 MB L2) through the framework's chain reader:

  * ``fused``  — the functions are traced (``ops/trace.py``) and the whole chain is one
    NVRTC-compiled count kernel + one write kernel that emits only the kept rows;
  * ``eager``  — ``BIGSLICE_TRACE=0``: every operator of every lambda is its own PyTorch
    kernel with its own temporary (what round 1 did).

CUDA-event timed, best of 5 after 2 warm-ups; results compared row for row.
One JSON line:

    python benchmarks/map_fusion.py [--rows N] > gpurun_out/map_fusion.json
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bigslice_b200 as bs  # noqa: E402
from bigslice_b200.frame import Frame  # noqa: E402
from bigslice_b200.ops import jit  # noqa: E402
from bigslice_b200.sliceio import FrameReader  # noqa: E402


def pipeline(src):
    s = bs.Map(src, lambda k, v: (k % 1009, v * 2.0 + 1.0, k))
    s = bs.Filter(s, lambda k, v, o: (o % 3 != 0) & (v > 0.0))
    return bs.Map(s, lambda k, v, o: (k, v + o))


def run_chain(frame):
    """Instantiate the three stages' readers over ``frame`` and drain them."""
    src = bs.Const(1, frame.col(0)[:1], frame.col(1)[:1])   # types only; data comes from the reader
    stages = []
    s = pipeline(src)
    while s is not src:
        stages.append(s)
        s = s.dep(0).slice
    rd = FrameReader(frame)
    for st in reversed(stages):
        rd = st.reader(0, [rd])
    out = []
    while True:
        f = rd.read(None)
        if f is None:
            break
        out.append(f)
    return out[0] if out else None


def timed(fn, dev, reps=5, warm=2):
    for _ in range(warm):
        out = fn()
    best = 1e30
    for _ in range(reps):
        if dev == "cuda":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        else:
            import time
            t = time.perf_counter()
            out = fn()
            best = min(best, (time.perf_counter() - t) * 1e3)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 26)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    a = ap.parse_args()
    dev = a.device
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    k = torch.randint(-(1 << 40), 1 << 40, (a.rows,), device=dev, generator=g)
    v = torch.randn(a.rows, device=dev, generator=g)
    frame = Frame([k, v], bs.slicetype.Type([bs.int64, bs.float32], 1))

    os.environ["BIGSLICE_TRACE"] = "1"
    before = jit.launch_count
    t_f, out_f = timed(lambda: run_chain(frame), dev)
    fused_launches = (jit.launch_count - before) // 7
    os.environ["BIGSLICE_TRACE"] = "0"
    t_e, out_e = timed(lambda: run_chain(frame), dev)
    same = len(out_f) == len(out_e) and torch.equal(out_f.col(0), out_e.col(0)) and \
        torch.allclose(out_f.col(1), out_e.col(1), rtol=1e-5)
    in_bytes = a.rows * 12
    out_bytes = len(out_f) * 12
    print(json.dumps({"bench": "python Map->Filter->Map over one tile", "rows": a.rows, "device": dev,
                      "kept_rows": len(out_f), "fused_ms": round(t_f, 3), "eager_ms": round(t_e, 3),
                      "speedup": round(t_e / t_f, 2), "equal": bool(same),
                      "fused_kernel_launches_per_tile": fused_launches,
                      "fused_effective_GBps": round((in_bytes + out_bytes) / t_f / 1e6, 1)}))


if __name__ == "__main__":
    main()
