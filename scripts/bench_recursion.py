#!/usr/bin/env python
"""Throughput of the three recursive integrators (ky.cpp:4191-4514) in their wavefront form and in the per-pixel kernel:
Cornell default scene 1024x768, depth 5, both_mis.  python scripts/bench_recursion.py [spp]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ky_b200 as ky

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w, h = 1024, 768
dev = ky.Device(0)
scene = ky.Scene(ky.SCENE_CORNELL, w, h, ky.CB_DEFAULT)
dev.upload(scene)
film = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
for name, integ in (("simple_path_tracing_recursion", ky.INT_SIMPLE_PT_RECURSION), ("path_tracing_recursion", ky.INT_PT_RECURSION),
                    ("path_tracing_recursion_defered", ky.INT_PT_RECURSION_DEFERED), ("path_tracing_iteration", ky.INT_PT_ITERATION)):
    for org, fl in (("wavefront", 0), ("per-pixel kernel", ky.FLAG_FUSED)):
        d = ky.render_desc(w, h, spp, integrator=integ, max_depth=5, direct_sample=ky.DS_BOTH_MIS, flags=ky.FLAG_ACCUMULATE | fl)
        dev.render_device(d, film.data_ptr()); dev.stats()
        times = []
        for _ in range(3):
            film.zero_()
            dev.render_device(d, film.data_ptr())
            st = dev.stats()
            times.append(st.device_ms)
        ms = sorted(times)[1]
        print(f"{name:34s} {org:18s} {w * h * spp / ms / 1e3:8.1f} Msamples/s  {st.rays / ms / 1e3:8.1f} Mrays/s  {ms:8.2f} ms  rays/sample {st.rays / (w * h * spp):.2f}")
