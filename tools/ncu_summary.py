#!/usr/bin/env python
"""Summarises one kernel of an .ncu-rep (ncu --set full) as JSON: the metrics DESIGN.md quotes.
usage: tools/ncu_summary.py report.ncu-rep "<command that produced it>" > profiles/<name>.json"""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active"]


def main():
    rep, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        m = {k: {"unit": d[k][0], "value": d[k][1]} for k in WANT if k in d}
        m.update({k: {"unit": d[k][0], "value": d[k][1]} for k in d
                  if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")})
        out.append({"command": cmd, "kernel": d.get("Kernel Name", ("", ""))[1], "grid": d.get("Grid Size", ("", ""))[1],
                    "block": d.get("Block Size", ("", ""))[1], "metrics": m})
    json.dump(out[0] if len(out) == 1 else out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
