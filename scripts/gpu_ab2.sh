# usage: bash scripts/gpu_ab2.sh variant   (parity + C5 bench + C1/C3/C4 configs for one variant)
mkdir -p gpurun_out
v=$1
export KYD_LIB=$PWD/ky_b200/lib/ab/libkyd_$v.so
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/ab_parity_$v.log 2>&1; tail -2 gpurun_out/ab_parity_$v.log
timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
python - <<PY
import json
d=json.load(open("gpurun_out/ab_$v.json"))
print("$v", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
KYD_STAGE_TIMING=1 timeout 300 python scripts/bench_configs.py 16 > gpurun_out/configs_$v.txt 2>&1
grep -B1 "C1 \|C3 veach 1280x720 PT d5 both_mis\|C4 " gpurun_out/configs_$v.txt | cut -c1-140
