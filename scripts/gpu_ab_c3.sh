# usage: bash scripts/gpu_ab_c3.sh variant...   (A/B of ky_b200/lib/ab/libkyd_<variant>.so on the C3 (Veach) bench line; first variant also runs the parity tests)
mkdir -p gpurun_out
first=1
for v in "$@"; do
  export KYD_LIB=$PWD/ky_b200/lib/ab/libkyd_$v.so
  if [ $first = 1 ]; then
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_kat.py tests/test_gpu_full_size.py -x -q > gpurun_out/abc3_parity_$v.log 2>&1; tail -2 gpurun_out/abc3_parity_$v.log
  fi
  first=0
  python bench.py --config C3 --no-cpu-baseline --no-configs --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/abc3_$v.json 2> gpurun_out/abc3_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/abc3_$v.json"))
print("$v", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
done
