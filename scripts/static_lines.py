#!/usr/bin/env python
"""Static SASS instruction count per CUDA source line of one kernel, from `nvdisasm -gi` output of the cubin:
python scripts/static_lines.py all.sass kernel_symbol_substring [top] [depth]
depth 0 attributes an instruction to the innermost inlined line, depth -1 to the outermost (the line in the kernel body)."""
import re, sys
from collections import Counter
path, key = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
depth = int(sys.argv[4]) if len(sys.argv) > 4 else 0
inside = False; cur = []; fresh = True; counts = Counter(); total = 0
for ln in open(path, errors="replace"):
    if ln.startswith(".text."):
        inside = key in ln
        cur = []; fresh = True
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            cur = []; fresh = False
        cur.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        fresh = True
        total += 1
        if cur:
            counts[cur[min(depth, len(cur) - 1) if depth >= 0 else max(0, len(cur) + depth)]] += 1
print("static instructions", total)
src = {}
for (f, l), c in counts.most_common(top):
    if f not in src:
        try: src[f] = open("ky_b200/csrc/" + f).read().split("\n")
        except OSError: src[f] = []
    text = src[f][l - 1].strip()[:110] if l - 1 < len(src[f]) else ""
    print(f"{c:6d} {100*c/total:5.1f}%  {f}:{l}: {text}")
