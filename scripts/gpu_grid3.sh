# grid-size sweeps on the lockstep build: KYD_GRID_NEE (k_nee blocks per SM, shade at its default 4) on C3; KYD_GRID_SHADE 2..6 on C5 and C3
mkdir -p gpurun_out
run() { # tag config
timeout 600 python bench.py --config $2 --no-cpu-baseline --no-configs --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/grid3_$1_$2.json 2> gpurun_out/grid3_$1_$2.err
python - <<PY
import json
d=json.load(open("gpurun_out/grid3_$1_$2.json"))
print("$1 $2", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
}
for g in 6 12 18 24 36 48; do KYD_GRID_NEE=$g run nee$g C3; done
for g in 2 3 4 5 6 8; do KYD_GRID_SHADE=$g run shade$g C5; KYD_GRID_SHADE=$g run shade$g C3; done
KYD_GRID256=4 run i256_4 C5; KYD_GRID256=8 run i256_8 C5; KYD_GRID256=32 run i256_32 C5
