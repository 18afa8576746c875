# round 2: everything that needs two GPUs -- the multi-GPU context behind the C ABI, the NCCL path, bench.py under torchrun
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_distributed.py -q > gpurun_out/r02_multi_tests.log 2>&1; tail -5 gpurun_out/r02_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -5 gpurun_out/r02_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n2.json"))
print("N=2 C5", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d.get("multi_gpu_parity"), d.get("multi_gpu_parity_detail"), "reduce_ms", d["reduce_ms"])
for c, v in d.get("configs", {}).items():
    print("   ", c, round(v["msamples_per_s"], 1), "x", v["n_gpus"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_ref_n2.json 2> gpurun_out/r02_bench_ref_n2.err; cut -c1-300 gpurun_out/r02_bench_ref_n2.json
# the in-library multi-GPU context on the headline job: 2 GPUs behind one kyd_render call
timeout 600 python - <<'PY'
import sys, time; sys.path.insert(0, '.')
import numpy as np, ky_b200 as ky
w, h, spp = 3840, 2160, 64
scene = ky.Scene(ky.SCENE_CORNELL, w, h, ky.CB_DEFAULT)
film = np.empty((h, w, 3), np.float32)
for devices in ([0], [0, 1]):
    dev = ky.Device(devices); dev.upload(scene)
    d = ky.render_desc(w, h, 16384, sample_begin=0, sample_end=spp)
    dev.render(d, film)
    t0 = time.perf_counter(); dev.render(d, film); dt = time.perf_counter() - t0
    st = dev.stats()
    print("kyd_render on devices", devices, f"{w*h*spp/dt/1e6:.0f} Msamples/s end to end (host film), device interval {st.device_ms:.1f} ms, rays {st.rays}")
    dev.close()
PY
