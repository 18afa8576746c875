# round 2 parity soaks against the C oracle (all host cores): the new traversal, fused shade, k_nee early-outs, recursion wavefront
mkdir -p gpurun_out
bash scripts/gpu_soak.sh > /dev/null 2>&1; cp gpurun_out/soak.txt gpurun_out/r02_soak.txt
timeout 1500 python scripts/soak_c5_1g.py >> gpurun_out/r02_soak.txt 2>&1
timeout 1500 python - <<'PY' >> gpurun_out/r02_soak.txt 2>&1
import sys, time; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, ky_b200 as ky, kyo
d = ky.Device(0)
jobs = [("C3 veach 1280x720 PT d5 both_mis", ky.SCENE_VEACH, 0, 1280, 720, 256, 5, ky.INT_PT_ITERATION),
        ("cornell 1024x768 path_tracing_recursion d5", ky.SCENE_CORNELL, ky.CB_DEFAULT, 1024, 768, 32, 5, ky.INT_PT_RECURSION),
        ("cornell 1024x768 path_tracing_recursion_defered d5", ky.SCENE_CORNELL, ky.CB_DEFAULT, 1024, 768, 32, 5, ky.INT_PT_RECURSION_DEFERED),
        ("cornell 1024x768 simple_path_tracing_recursion d5", ky.SCENE_CORNELL, ky.CB_DEFAULT, 1024, 768, 32, 5, ky.INT_SIMPLE_PT_RECURSION),
        ("C4 cornell+environment 480x360 PT d8 both_mis", ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_ENVIRONMENT, 480, 360, 256, 8, ky.INT_PT_ITERATION),
        ("C4 cornell+point 480x360 PT d8 both_mis", ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_POINT, 480, 360, 256, 8, ky.INT_PT_ITERATION)]
for name, sid, fl, w, h, spp, depth, integ in jobs:
    scene = ky.Scene(sid, w, h, fl); d.upload(scene)
    desc = ky.render_desc(w, h, spp, integrator=integ, max_depth=depth, flags=0)
    got = d.render(desc); st = d.stats()
    t = time.time(); want, rays = kyo.render(scene, desc); dt = time.time() - t
    diff = (got.view(np.uint32) != want.view(np.uint32)).any(axis=-1)
    print(f"{name} @ {spp} spp: {w*h*spp/1e6:.1f} Msamples, {rays/1e6:.0f} Mrays; pixels differing {int(diff.sum())} of {w*h}; rays device {st.rays} oracle {rays}; oracle {dt:.1f} s")
PY
cat gpurun_out/r02_soak.txt
