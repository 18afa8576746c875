# round 2, session B: two-phase traversal -- parity, self-test, timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02b_gpu_tests.log 2>&1; tail -15 gpurun_out/r02b_gpu_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; cut -c1-200 gpurun_out/r02b_bench.json; tail -3 gpurun_out/r02b_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02b_bench.json"))
print("C5", round(d["value"]), d["stage_ms_per_step"], "e2e", round(d["e2e"]["value"]))
for c, v in d.get("configs", {}).items():
    print(c, round(v["msamples_per_s"], 1), v["stage_ms_rank0"])
PY
