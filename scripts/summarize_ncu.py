#!/usr/bin/env python
"""Summarise an .ncu-rep (read here without a GPU) into a markdown table under profiles/.

    python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/name.md "title" ["note"]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    note = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as fh:
        fh.write("# %s\n\nSource: `%s` (`ncu --set full --clock-control none --import-source on`).\n\n" % (title, rep))
        for vals in rows[2:]:
            d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
            fh.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % d.get("Kernel Name", ("?",))[0][:100])
            for k in KEYS:
                if k in d:
                    fh.write("| %s | %s | %s |\n" % (k, d[k][0], d[k][1]))
            fh.write("\n")
        if note:
            fh.write(note + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
