mkdir -p gpurun_out
KYD_BENCH_ONLY="C3 veach 1280x720 PT d5 both_mis" ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_shadow" -s 8 -c 8 -o gpurun_out/prof_veach python scripts/bench_configs.py 2 > gpurun_out/ncu_veach.log 2>&1
