# round 2: the 8-GPU record -- default line (C5, with C1-C4 in `configs` and the parity flag) and the C3 line, under torchrun
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; tail -3 gpurun_out/r02_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config C3 --no-configs --steps 8 --warmup 3 > gpurun_out/r02_bench_n8_c3.json 2> gpurun_out/r02_bench_n8_c3.err; tail -3 gpurun_out/r02_bench_n8_c3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n8.json"))
print("N=8 C5", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d.get("multi_gpu_parity"), "reduce_ms", round(d["reduce_ms"], 2))
for c, v in d.get("configs", {}).items():
    print("   ", c, round(v["msamples_per_s"], 1), "x", v["n_gpus"])
d = json.load(open("gpurun_out/r02_bench_n8_c3.json"))
print("N=8 C3 line", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["stage_ms_per_step"])
PY
