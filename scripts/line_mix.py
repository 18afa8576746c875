#!/usr/bin/env python
"""Executed instructions per CUDA source line of one kernel from an ncu report:
python scripts/line_mix.py report.ncu-rep kernel_regex [launch_skip] [top]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; items = {}; total = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = {}
        for i, h in enumerate(r):
            hdr.setdefault(h, i)
        continue
    if hdr and len(r) > 8 and r[0].isdigit():
        ie = r[hdr["Instructions Executed"]]
        if ie.isdigit() and int(ie) > 0:
            te = int(r[hdr["Thread Instructions Executed"]])
            key = (cur, int(r[0]))
            if key not in items:
                items[key] = (int(ie), te, r[1].strip()[:100])
for ie, te, src in items.values():
    total += ie
print("kernel", rx, "total warp-inst", total)
for (f, ln), (ie, te, src) in sorted(items.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*ie/total:5.1f}% lanes {te/ie:5.1f}  {f}:{ln}: {src}")
