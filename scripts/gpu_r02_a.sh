# round 2, session A: the library fixes + multi context + new bench line on one GPU; Veach (C3) profile of the round-1 kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02a_gpu_tests.log 2>&1; tail -4 gpurun_out/r02a_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; cut -c1-400 gpurun_out/r02a_bench.json; tail -3 gpurun_out/r02a_bench.err
timeout 600 python bench.py --config C3 --no-configs --no-cpu-baseline > gpurun_out/r02a_bench_c3.json 2> gpurun_out/r02a_bench_c3.err; cut -c1-300 gpurun_out/r02a_bench_c3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02a_launches_c3.csv python bench.py --config C3 --steps 1 --warmup 1 --spp-per-step 16 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/r02a_ncu_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_shadow|k_accumulate" -s 7 -c 8 -o gpurun_out/r02a_prof_c3 python bench.py --config C3 --steps 1 --warmup 1 --spp-per-step 16 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/r02a_ncu_c3_full.log 2>&1
ls -la gpurun_out/r02a_*
