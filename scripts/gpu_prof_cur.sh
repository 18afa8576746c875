mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_shadow|k_raygen" -s 7 -c 8 -o gpurun_out/prof_cur python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_cur_full.log 2>&1
