mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 8 -c 2 -o gpurun_out/prof_shade2 python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_shade2.log 2>&1
