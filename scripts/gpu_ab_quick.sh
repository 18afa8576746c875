# usage: bash scripts/gpu_ab_quick.sh <config> variant...   (bench line of ky_b200/lib/ab/libkyd_<variant>.so builds on one configuration, no tests)
mkdir -p gpurun_out
c=$1; shift
for v in "$@"; do
  export KYD_LIB=$PWD/ky_b200/lib/ab/libkyd_$v.so
  timeout 600 python bench.py --config $c --no-cpu-baseline --no-configs --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/abq_${v}_$c.json 2> gpurun_out/abq_${v}_$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/abq_${v}_$c.json"))
print("$v $c", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
done
