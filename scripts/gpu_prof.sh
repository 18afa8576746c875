# ncu evidence of the current build: launch list + --set full capture of one bounce (C5), and of C3 when asked
# usage: bash scripts/gpu_prof.sh <tag> [C3]
tag=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_shadow|k_accumulate" -s 5 -c 9 -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu_full.log 2>&1
if [ "$2" = "C3" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_c3.csv python bench.py --config C3 --steps 1 --warmup 1 --spp-per-step 16 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_shadow|k_accumulate|k_nee" -s 7 -c 8 -o gpurun_out/${tag}_prof_c3 python bench.py --config C3 --steps 1 --warmup 1 --spp-per-step 16 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu_c3_full.log 2>&1
fi
ls -la gpurun_out/${tag}_*
