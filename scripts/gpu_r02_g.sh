# round 2, session G: fused intersect-in-shade: parity (full suite), A/B fused vs not (same binary, KYD_FUSE_INTERSECT)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02g_gpu_tests.log 2>&1; tail -8 gpurun_out/r02g_gpu_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02g_gpu_tests.log | head -20
for fuse in 1 0; do
KYD_FUSE_INTERSECT=$fuse timeout 600 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/r02g_bench_fuse$fuse.json 2> gpurun_out/r02g_bench_fuse$fuse.err
python - <<PY
import json
d = json.load(open("gpurun_out/r02g_bench_fuse$fuse.json"))
print("fuse=$fuse C5", round(d["value"]), {k: round(v, 1) for k, v in d["stage_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]))
for c, v in d.get("configs", {}).items():
    print("   ", c, round(v["msamples_per_s"], 1), {k: round(x, 1) for k, x in v["stage_ms_rank0"].items()})
PY
done
