# round 2, session C: two-phase traversal v2 (aligned classifiers) -- parity, self-test, timing; A/B of occupancy variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02c_gpu_tests.log 2>&1; tail -3 gpurun_out/r02c_gpu_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; tail -3 gpurun_out/r02c_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c_bench.json"))
print("C5", round(d["value"]), d["stage_ms_per_step"], "e2e", round(d["e2e"]["value"]))
for c, v in d.get("configs", {}).items():
    print(c, round(v["msamples_per_s"], 1), v["stage_ms_rank0"])
PY
bash scripts/gpu_ab.sh base mb5 mb6 imb4 2>&1 | tail -8
