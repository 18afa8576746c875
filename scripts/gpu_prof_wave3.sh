mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_wave3.csv python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_wave3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_shadow" -s 18 -c 6 -o gpurun_out/prof_wave3 python bench.py --steps 1 --warmup 1 --spp-per-step 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_wave3_full.log 2>&1
