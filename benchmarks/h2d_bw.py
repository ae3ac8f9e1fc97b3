#!/usr/bin/env python
"""Pinned host->device copy bandwidth (the floor of the end-to-end benchmark arm)."""
import json
import sys
import torch

n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for nstream in (1, 2, 4):
    streams = [torch.cuda.Stream() for _ in range(nstream)]
    chunk = n // nstream
    best = 0.0
    for it in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i, s in enumerate(streams):
            s.wait_event(e0)
            with torch.cuda.stream(s):
                d[i * chunk:(i + 1) * chunk].copy_(h[i * chunk:(i + 1) * chunk], non_blocking=True)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, n / (e0.elapsed_time(e1) / 1e3) / 1e9)
    out["h2d_gbps_%dstream" % nstream] = best
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
h.copy_(d, non_blocking=True)
e1.record()
torch.cuda.synchronize()
out["d2h_gbps"] = n / (e0.elapsed_time(e1) / 1e3) / 1e9
print(json.dumps(out))
