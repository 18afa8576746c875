mkdir -p gpurun_out
./tools/_build/membench > gpurun_out/membench2.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_shadow|k_accumulate" -s 5 -c 9 -o gpurun_out/prof_final python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_final_full.log 2>&1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
cat gpurun_out/bench_final.json | cut -c1-1500
KYD_STAGE_TIMING=1 python scripts/bench_configs.py 16 > gpurun_out/configs_final.txt 2>&1
