# usage: bash scripts/gpu_ab_both.sh variant...  (C5 and C3 bench lines of ky_b200/lib/ab/libkyd_<variant>.so)
mkdir -p gpurun_out
for v in "$@"; do
  export KYD_LIB=$PWD/ky_b200/lib/ab/libkyd_$v.so
  for c in C5 C3; do
  python bench.py --config $c --no-cpu-baseline --no-configs --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/abb_${v}_$c.json 2> gpurun_out/abb_${v}_$c.err
  python - <<PY
import json
d=json.load(open("gpurun_out/abb_${v}_$c.json"))
print("$v $c", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
  done
done
