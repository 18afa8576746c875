# round 2: the 4-GPU default line (C5, with C1-C4 in `configs` and the parity flag), under torchrun
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 8 --warmup 3 > gpurun_out/r02_bench_n4.json 2> gpurun_out/r02_bench_n4.err; tail -2 gpurun_out/r02_bench_n4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_n4.json"))
print("N=4 C5", round(d["value"]), "e2e", round(d["e2e"]["value"]), "parity", d.get("multi_gpu_parity"), "reduce_ms", round(d["reduce_ms"], 2))
for c, v in d.get("configs", {}).items():
    print("   ", c, round(v["msamples_per_s"], 1), "x", v["n_gpus"])
PY
