# round 2, session F: full GPU suite + bench + profile of the compact one-light shade (no record prefetch)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02f_gpu_tests.log 2>&1; tail -5 gpurun_out/r02f_gpu_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -3 gpurun_out/r02f_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02f_bench.json"))
print("C5", round(d["value"]), d["stage_ms_per_step"], "e2e", round(d["e2e"]["value"]))
for c, v in d.get("configs", {}).items():
    print(c, round(v["msamples_per_s"], 1), v["stage_ms_rank0"])
PY
bash scripts/gpu_prof.sh r02f C3 > /dev/null 2>&1
ls gpurun_out/r02f_*
