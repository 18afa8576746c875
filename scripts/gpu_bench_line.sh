# the default bench line (as the driver runs it), the reference arm, and the C3 line.   usage: bash scripts/gpu_bench_line.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("C5", round(d["value"]), {k: round(v, 1) for k, v in d["stage_ms_per_step"].items()}, "e2e", round(d["e2e"]["value"]), "cpu", d.get("cpu_baseline", {}).get("value"), "launches", d["gpu_launches"])
print("roofline", d["roofline"]["frac"], "fp32", d["roofline_fp32_issue"]["frac"], d["roofline_fp32_issue"]["frac_reference_equivalent_rays"], "ncu same sources", (d.get("ncu") or {}).get("same_sources"))
for c, v in d.get("configs", {}).items():
    print("   ", c, round(v["msamples_per_s"], 1), {k: round(x, 1) for k, x in v["stage_ms_rank0"].items()}, v.get("reference_gpu"))
PY
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; cut -c1-160 gpurun_out/${tag}_bench_ref.json
timeout 600 python bench.py --config C3 --no-configs > gpurun_out/${tag}_bench_c3.json 2> gpurun_out/${tag}_bench_c3.err; cut -c1-120 gpurun_out/${tag}_bench_c3.json
