# round 2, final evidence of the committed build (second session): full GPU suite, smoke, ncu summaries of C5 and C3, bench lines
# (default, reference arm, C3).  Outputs gpurun_out/r02z_*; copied into profiles/ by hand afterwards.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02z_gpu_tests.log 2>&1; tail -3 gpurun_out/r02z_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/gpu_prof.sh r02z C3 > /dev/null 2>&1
bash scripts/gpu_bench_line.sh r02z
ls gpurun_out/r02z_* | wc -l
