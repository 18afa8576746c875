#!/usr/bin/env python
"""Times the reference's own CUDA kernel (smallpt2pbrt/smallpt_kernel.cu built for sm_100a, oracle/_ref) in THIS process and
prints one JSON line.  Run as a child process: the reference's CHECK_CUDA exits the process on any CUDA error.
usage: python scripts/ref_cuda_smallpt.py width height spp base_stack_bytes"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kyref

w, h, spp, stack = (int(v) for v in sys.argv[1:5])
kyref.smallpt_cuda(max(8, w // 8), max(8, h // 8), 2, stack)       # context, module load, clocks
film, sec = kyref.smallpt_cuda(w, h, spp, stack)
print(json.dumps({"msamples_per_s": w * h * spp / sec / 1e6, "seconds": sec, "mean": float(film.mean()), "base_stack_bytes": stack}))
