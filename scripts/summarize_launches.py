#!/usr/bin/env python
"""Per-kernel time table from an `ncu --metrics gpu__time_duration.sum --csv` launch log.

    python scripts/summarize_launches.py gpurun_out/launches.csv [last_n_launches]

Durations under ncu are serialised per kernel (no overlap), so the table shows where
device time goes, not the step time."""
import csv
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    last = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path, newline="") as fh:
        lines = [ln for ln in fh if not ln.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        rows.append((r["Kernel Name"].split("(")[0][:70], v * scale))
    if last:
        rows = rows[-last:]
    agg = OrderedDict()
    for k, us in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | % |\n|---|---|---|---|")
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f |" % (k, c, us, 100 * us / max(tot, 1e-9)))
    print("| **all** | %d | %.1f | 100 |" % (len(rows), tot))


if __name__ == "__main__":
    main()
