# grid-size sweep of the 128-thread kernels (shade, k_nee) on the lockstep build: KYD_GRID128 = blocks per SM
mkdir -p gpurun_out
for g in 4 8 12 16 24 32 48; do for c in C5 C3; do
KYD_GRID128=$g timeout 600 python bench.py --config $c --no-cpu-baseline --no-configs --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/grid_${g}_$c.json 2> gpurun_out/grid_${g}_$c.err
python - <<PY
import json
d=json.load(open("gpurun_out/grid_${g}_$c.json"))
print("grid128=$g $c", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
done; done
