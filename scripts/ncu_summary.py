#!/usr/bin/env python
"""Summaries of ncu output for profiles/: per-kernel time shares from a launch list CSV, and key
metrics per kernel from a .ncu-rep (via `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
]


def short(name):
    n = name.split("(")[0]
    return n.replace("void ", "").replace("kyd::", "")


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10 and r[0].isdigit()]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        agg[short(r[4])][0] += 1
        agg[short(r[4])][1] += float(r[-1])
    total = sum(v[1] for v in agg.values())
    print(f"{'kernel':40s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:40s} {n:8d} {ns / 1e6:10.3f} {100 * ns / total:6.1f}%")
    print(f"{'TOTAL':40s} {sum(v[0] for v in agg.values()):8d} {total / 1e6:10.3f}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    name_i = idx["Kernel Name"]
    for r in rows[2:]:
        print(f"--- {r[idx['ID']]} {short(r[name_i])}")
        for k in KEYS:
            if k in idx:
                print(f"    {k:70s} {r[idx[k]]:>16s} {units[idx[k]]}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2])
