mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; cut -c1-300 gpurun_out/bench_now.json; grep -c . gpurun_out/bench_now.json
KYD_STAGE_TIMING=1 timeout 300 python scripts/bench_configs.py 16 > gpurun_out/configs_now.txt 2>&1; grep -v "stage ms" gpurun_out/configs_now.txt | cut -c1-112
timeout 300 python - <<'PY'
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, ky_b200 as ky, cases, torch
d = ky.Device(0)
base = cases.big_scene(extra=360)
w, h, spp = 1024, 768, 16
import ctypes as C
scene = cases.big_scene(extra=360)
scene.desc.camera = ky.Scene(ky.SCENE_CORNELL, w, h).desc.camera
d.upload(scene)
film = torch.zeros((h, w, 3), device="cuda")
desc = ky.render_desc(w, h, spp, flags=ky.FLAG_ACCUMULATE)
d.render_device(desc, film.data_ptr()); d.stats(); film.zero_()
d.render_device(desc, film.data_ptr()); st = d.stats()
print(f"BVH scene ({scene.desc.surface_count} surfaces) {w}x{h}@{spp} PT d5 both_mis: {w*h*spp/st.device_ms/1e3:.1f} Msamples/s, {st.rays/st.device_ms/1e3:.0f} Mrays/s, {st.device_ms:.1f} ms")
PY
