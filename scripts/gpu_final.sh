# round 2, final evidence of the committed build: full GPU suite, smoke, bench lines (default, reference arm, C3), ncu summaries
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02final_gpu_tests.log 2>&1; tail -3 gpurun_out/r02final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/gpu_prof.sh r02final C3 > /dev/null 2>&1
for f in kernels lines stalls; do cp gpurun_out/r02final_prof_$f.txt profiles/r02_c5_$f.txt; cp gpurun_out/r02final_prof_c3_$f.txt profiles/r02_c3_$f.txt; done
cp gpurun_out/r02final_prof_ncu_summary.json profiles/r02_ncu_summary.json
bash scripts/gpu_bench_line.sh r02final
ls gpurun_out/r02final_* | wc -l
