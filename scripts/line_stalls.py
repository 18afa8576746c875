#!/usr/bin/env python
"""Stall samples per CUDA source line: python scripts/line_stalls.py report kernel_regex [skip] [stall_col] [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
col = sys.argv[4] if len(sys.argv) > 4 else "stall_long_sb"
top = int(sys.argv[5]) if len(sys.argv) > 5 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; items = {}; total = 0; allsamples = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = {}
        for i, h in enumerate(r):
            hdr.setdefault(h, i)
        continue
    if hdr and len(r) > 8 and r[0].isdigit():
        v = r[hdr[col]]; sm = r[hdr["# Samples"]]
        key = (cur, int(r[0]))
        if key not in items and v.isdigit():
            items[key] = (int(v), int(sm) if sm.isdigit() else 0, r[1].strip()[:95])
for v, sm, _ in items.values():
    total += v; allsamples += sm
print("kernel", rx, col, total, "of", allsamples, "samples")
for (f, ln), (v, sm, src) in sorted(items.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v/max(1,total):5.1f}%  {f}:{ln}: {src}")
