# round 2, session I: k_nee pair kernel for multi-light headline scenes: parity, bench, profiles
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02i_gpu_tests.log 2>&1; tail -4 gpurun_out/r02i_gpu_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02i_gpu_tests.log | head -20
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -3 gpurun_out/r02i_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02i_bench.json"))
print("C5", round(d["value"]), {k: round(v, 1) for k, v in d["stage_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]))
for c, v in d.get("configs", {}).items():
    print("   ", c, round(v["msamples_per_s"], 1), {k: round(x, 1) for k, x in v["stage_ms_rank0"].items()})
PY
bash scripts/gpu_prof.sh r02i C3 > /dev/null 2>&1
ls gpurun_out/r02i_* | head -30
