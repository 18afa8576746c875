# round 2, session N: LCG jump table -- full GPU suite + both headline lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02n_gpu_tests.log 2>&1; tail -3 gpurun_out/r02n_gpu_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02n_gpu_tests.log | head
for c in C5 C3; do
python bench.py --config $c --no-cpu-baseline --no-configs --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/r02n_$c.json 2> gpurun_out/r02n_$c.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02n_$c.json"))
print("$c", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
done
