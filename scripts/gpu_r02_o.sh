# round 2, session O: k_nee with deferred BSDF-sampled queries + analytic rsqrt for almost-unit squares -- full GPU suite, both headline lines, A/B on C3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02o_gpu_tests.log 2>&1; tail -3 gpurun_out/r02o_gpu_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02o_gpu_tests.log | head
for c in C5 C3; do
timeout 600 python bench.py --config $c --no-cpu-baseline --no-configs --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/r02o_$c.json 2> gpurun_out/r02o_$c.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02o_$c.json"))
print("$c", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
done
for v in "$@"; do
  export KYD_LIB=$PWD/ky_b200/lib/ab/libkyd_$v.so
  for c in C3 C5; do
  timeout 600 python bench.py --config $c --no-cpu-baseline --no-configs --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/r02o_${v}_$c.json 2> gpurun_out/r02o_${v}_$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r02o_${v}_$c.json"))
print("$v $c", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
  done
done
