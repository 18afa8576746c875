#!/usr/bin/env python
"""Executed-instruction mix of one kernel from an ncu report (source page, SASS view):
python scripts/sass_mix.py report.ncu-rep kernel_regex [launch_skip]"""
import csv, subprocess, sys
from collections import defaultdict
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0, 0])
tot = [0, 0]
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix["Instructions Executed"]].isdigit(): continue
    src = r[ix["Source"]].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = op.split(".")[0]
    ie = int(r[ix["Instructions Executed"]]); te = int(r[ix["Thread Instructions Executed"]])
    agg[op][0] += ie; agg[op][1] += te; agg[op][2] += 1
    tot[0] += ie; tot[1] += te
print(rows[0][1][:90])
print(f"static SASS instructions {len(rows)-2}, executed warp-inst {tot[0]}, thread-inst {tot[1]}, avg active lanes {tot[1]/max(1,tot[0]):.1f}")
for op, (ie, te, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"  {op:10s} static {n:5d}  warp-inst {ie:10d} {100*ie/tot[0]:5.1f}%  lanes {te/max(1,ie):5.1f}")
