#!/usr/bin/env python
"""Static SASS instruction count per CUDA source line (code-size attribution): report kernel_regex [skip] [top]"""
import csv, subprocess, sys
from collections import defaultdict
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; key = None; cnt = defaultdict(int); src = {}; execd = defaultdict(int); hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = {}
        for i, h in enumerate(r): hdr.setdefault(h, i)
        continue
    if not hdr or len(r) < 9: continue
    if r[0].isdigit():
        key = (cur, int(r[0])); src[key] = r[1].strip()[:90]
    elif r[0] == "" and r[2].startswith("0x") and key:
        cnt[key] += 1
        v = r[hdr["Instructions Executed"]]
        if v.isdigit() and int(v) > 0: execd[key] += 1
tot = sum(cnt.values()); hot = sum(execd.values())
print(f"kernel {rx}: static SASS {tot}, of which executed at least once {hot}")
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v:6d} ({execd[k]:5d} hot)  {k[0]}:{k[1]}: {src[k]}")
