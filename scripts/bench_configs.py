#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs (C1-C4) on one GPU: parity-test cases, not bench lines, but the
numbers belong in profiles/.  python scripts/bench_configs.py [spp_slice]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ky_b200 as ky

S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = ky.Device(0)
CONFIGS = [
    ("C1 smallpt scene 1024x768 PT d5 both_mis", ky.SCENE_SMALLPT, 0, 1024, 768, ky.INT_PT_ITERATION, 5, ky.DS_BOTH_MIS, 64),
    ("C2 cornell 1024x768 direct_lighting both_mis", ky.SCENE_CORNELL, ky.CB_DEFAULT, 1024, 768, ky.INT_DIRECT_LIGHTING, 0, ky.DS_BOTH_MIS, 256),
    ("C2 cornell 1024x768 direct_lighting bsdf", ky.SCENE_CORNELL, ky.CB_DEFAULT, 1024, 768, ky.INT_DIRECT_LIGHTING, 0, ky.DS_BSDF, 256),
    ("C2 cornell 1024x768 direct_lighting light", ky.SCENE_CORNELL, ky.CB_DEFAULT, 1024, 768, ky.INT_DIRECT_LIGHTING, 0, ky.DS_LIGHT, 256),
    ("C3 veach 1280x720 PT d5 both_mis", ky.SCENE_VEACH, 0, 1280, 720, ky.INT_PT_ITERATION, 5, ky.DS_BOTH_MIS, 1024),
    ("C3 veach 1280x720 PT d5 bsdf", ky.SCENE_VEACH, 0, 1280, 720, ky.INT_PT_ITERATION, 5, ky.DS_BSDF, 1024),
    ("C3 veach 1280x720 PT d5 light", ky.SCENE_VEACH, 0, 1280, 720, ky.INT_PT_ITERATION, 5, ky.DS_LIGHT, 1024),
    ("C4 cornell-area panel 480x360 PT d8 both_mis", ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_AREA, 480, 360, ky.INT_PT_ITERATION, 8, ky.DS_BOTH_MIS, 1024),
    ("C4 cornell-env panel 480x360 PT d8 both_mis", ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_ENVIRONMENT, 480, 360, ky.INT_PT_ITERATION, 8, ky.DS_BOTH_MIS, 1024),
    ("C4 cornell-point panel 480x360 PT d8 both_mis", ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_POINT, 480, 360, ky.INT_PT_ITERATION, 8, ky.DS_BOTH_MIS, 1024),
    ("-- pixel kernel: veach 1280x720 PT d5 both_mis", ky.SCENE_VEACH, 0, 1280, 720, ky.INT_PT_ITERATION, 5, ky.DS_BOTH_MIS, -1024),
    ("-- pixel kernel: cornell 1024x768 pt_recursion d5", ky.SCENE_CORNELL, ky.CB_DEFAULT, 1024, 768, ky.INT_PT_RECURSION, 5, ky.DS_BOTH_MIS, -64),
]
ONLY = os.environ.get("KYD_BENCH_ONLY")
for name, sid, flags, w, h, integ, depth, ds, spp in CONFIGS:
    if ONLY and ONLY not in name:
        continue
    fl = ky.FLAG_FUSED if spp < 0 else int(os.environ.get("KYD_BENCH_FLAGS", "0"))
    spp = abs(spp)
    scene = ky.Scene(sid, w, h, flags)
    dev.upload(scene)
    film = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    n = min(S * max(1, (3840 * 2160) // (w * h)), spp)
    d = ky.render_desc(w, h, spp, integrator=integ, max_depth=depth, direct_sample=ds, sample_begin=0, sample_end=n, flags=ky.FLAG_ACCUMULATE | fl)
    dev.render_device(d, film.data_ptr()); dev.stats()
    times = []
    for _ in range(3):
        film.zero_()
        dev.render_device(d, film.data_ptr())
        st = dev.stats()
        times.append(st.device_ms)
    ms = sorted(times)[1]
    samples = w * h * n
    if os.environ.get("KYD_STAGE_TIMING") == "1":
        print("    stage ms:", {k: round(st.stage_ms[j], 1) for j, k in enumerate(["raygen", "intersect", "shade", "light_sample", "shadow", "-", "accumulate", "pixel"]) if st.stage_ms[j] > 0})
    print(f"{name:52s} {n:5d} spp slice  {samples / ms / 1e3:8.1f} Msamples/s  {st.rays / ms / 1e3:8.1f} Mrays/s (ref-equivalent)  {st.rays / samples:5.2f} rays/sample  {ms:8.2f} ms")
