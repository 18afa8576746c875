# round 2, session D: KAT harness + A/B of traversal variants + profile of the current build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02d_gpu_tests.log 2>&1; tail -5 gpurun_out/r02d_gpu_tests.log
bash scripts/gpu_ab.sh onephase cur noinline noinline_mb6 2>&1 | grep -v Traceback | tail -6
bash scripts/gpu_prof.sh r02d > /dev/null 2>&1
ls gpurun_out/r02d_*
