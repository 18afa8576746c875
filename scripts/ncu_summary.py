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


LOBES = {"0": "lambert", "1": "mirror", "2": "fresnel", "3": "phong"}


def kernel_class(name):
    """k_shade<0, 1, 1, 1> -> k_shade<lambert>; k_intersect<1> -> k_intersect; others by their bare name"""
    n = short(name)
    base = n.split("<")[0].strip()
    if base == "k_shade" and "<" in n:
        return f"k_shade<{LOBES.get(n.split('<')[1].split(',')[0].strip(), '?')}>"
    return base


def to_json(path, out_path, workload):
    """profiles/rNN_ncu_summary.json: what bench.py quotes (DRAM traffic, issue-slot utilisation per kernel class; the
    launch with the longest duration stands for its class), tagged with the hash of the kernel sources it was captured from"""
    import json, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}

    def num(r, k):
        try:
            return float(r[idx[k]].replace(",", ""))
        except (KeyError, ValueError):
            return None
    unit_of = {h: rows[1][i] for h, i in idx.items()}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    time_scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}   # -> us (older ncu prints "usecond" etc.)
    kernels = {}
    for r in rows[2:]:
        c = kernel_class(r[idx["Kernel Name"]])
        us = num(r, "gpu__time_duration.sum")
        tu = unit_of.get("gpu__time_duration.sum", "us")
        us = us * next((v for k, v in time_scale.items() if tu.startswith(k)), 1.0)
        if c in kernels and kernels[c]["duration_us"] >= us:
            continue
        rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
        kernels[c] = {
            "launch": short(r[idx["Kernel Name"]]), "duration_us": us,
            "registers": num(r, "launch__registers_per_thread"),
            "issue_active": (num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0) / 100,
            "warps_active": (num(r, "sm__warps_active.avg.pct_of_peak_sustained_active") or 0) / 100,
            "lanes_per_inst": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
            "pipe_fma": (num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") or 0) / 100,
            "pipe_fp64": (num(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") or 0) / 100,
            "dram_pct_of_peak": (num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") or 0) / 100,
            "dram_bytes": (rd or 0) * scale.get(unit_of.get("dram__bytes_read.sum", "byte"), 1.0)
                          + (wr or 0) * scale.get(unit_of.get("dram__bytes_write.sum", "byte"), 1.0),
            "warp_instructions": num(r, "smsp__inst_executed.sum"),
        }
    doc = {"capture": os.path.basename(path), "workload": workload, "source_hash": bench.source_hash(),
           "kernel_symbols": sorted({k.split("<")[0] for k in kernels}), "kernels": kernels,
           "how": "ncu --set full --clock-control none --import-source on; scripts/ncu_summary.py json"}
    json.dump(doc, open(out_path, "w"), indent=1)
    print(json.dumps(doc, indent=1))


def stalls(path):
    """warp-stall reasons per issued instruction, per captured launch"""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    keys = [h for h in hdr if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
    for r in rows[2:]:
        vals = sorted(((float(r[idx[k]].replace(",", "")) if r[idx[k]] else 0.0,
                        k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in keys), reverse=True)[:8]
        print(f"--- {r[idx['ID']]} {short(r[idx['Kernel Name']])}")
        print("    " + ", ".join(f"{k} {v:.2f}" for v, k in vals))


if __name__ == "__main__":
    if sys.argv[1] == "stalls":
        stalls(sys.argv[2])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2])
    elif sys.argv[1] == "json":
        to_json(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "bench.py --steps 1 --warmup 1 --spp-per-step 2")
    else:
        report(sys.argv[2])
