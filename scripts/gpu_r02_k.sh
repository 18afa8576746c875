# round 2, session K: candidate final build -- full GPU suite, smoke, bench, profiles (C5 + C3) summarised on the box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02k_gpu_tests.log 2>&1; tail -4 gpurun_out/r02k_gpu_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02k_gpu_tests.log | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -3 gpurun_out/r02k_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02k_bench.json"))
print("C5", round(d["value"]), {k: round(v, 1) for k, v in d["stage_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]), "cpu", d.get("cpu_baseline", {}).get("value"))
for c, v in d.get("configs", {}).items():
    print("   ", c, round(v["msamples_per_s"], 1), {k: round(x, 1) for k, x in v["stage_ms_rank0"].items()}, v.get("reference_gpu"))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err; cut -c1-200 gpurun_out/r02k_bench_ref.json
bash scripts/gpu_prof.sh r02k C3 > /dev/null 2>&1
ls gpurun_out/r02k_* | head -30
