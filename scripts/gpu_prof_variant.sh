# usage: bash scripts/gpu_prof_variant.sh variant   -> gpurun_out/prof_<variant>.ncu-rep (k_shade launches of bounce 1)
mkdir -p gpurun_out
export KYD_LIB=$PWD/ky_b200/lib/ab/libkyd_$1.so
ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 8 -c 2 -o gpurun_out/prof_$1 python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_$1.log 2>&1
