# the default bench line and the C3 line of the committed build with the committed ncu summary beside them (ncu.same_sources: true)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02y_bench.json 2> gpurun_out/r02y_bench.err; tail -2 gpurun_out/r02y_bench.err
timeout 600 python bench.py --config C3 --no-configs > gpurun_out/r02y_bench_c3.json 2> gpurun_out/r02y_bench_c3.err
python - <<'PY'
import json
for f in ("gpurun_out/r02y_bench.json", "gpurun_out/r02y_bench_c3.json"):
    d = json.load(open(f))
    print(d["config"]["workload"][:20], round(d["value"]), "e2e", round(d["e2e"]["value"]), "roofline", round(d["roofline"]["frac"], 3), "ncu same sources", d["ncu"]["same_sources"], d["ncu"]["file"], "clocks", d["clocks"])
PY
