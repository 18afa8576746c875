#!/usr/bin/env python
"""Warp-stall sampling summary of one kernel from an ncu report (SASS source page):
python scripts/stall_mix.py report.ncu-rep kernel_regex [launch_skip]"""
import csv, subprocess, sys
from collections import defaultdict
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {}
for i, h in enumerate(hdr):
    ix.setdefault(h, i)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = defaultdict(int); byop = defaultdict(lambda: defaultdict(int)); samples = 0
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix["# Samples"]].isdigit(): continue
    toks = r[ix["Source"]].split()
    op = (toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")).split(".")[0]
    samples += int(r[ix["# Samples"]])
    for s in stalls:
        v = r[ix[s]]
        if v.isdigit():
            tot[s] += int(v); byop[s][op] += int(v)
print(rows[0][1][:80], "samples", samples)
for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]:
    top = ", ".join(f"{o}:{100*c/v:.0f}%" for o, c in sorted(byop[s].items(), key=lambda kv: -kv[1])[:6])
    print(f"  {s:28s} {100*v/max(1,samples):5.1f}%   {top}")
