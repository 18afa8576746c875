# usage: bash scripts/gpu_ab.sh variant...   (A/B of ky_b200/lib/ab/libkyd_<variant>.so on bench.py; first variant also runs the parity tests)
mkdir -p gpurun_out
first=1
for v in "$@"; do
  export KYD_LIB=$PWD/ky_b200/lib/ab/libkyd_$v.so
  if [ $first = 1 ] && [ "$v" != base ]; then
    python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/ab_parity_$v.log 2>&1; tail -2 gpurun_out/ab_parity_$v.log
  fi
  first=0
  python bench.py --no-cpu-baseline --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v.json"))
print("$v", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
done
