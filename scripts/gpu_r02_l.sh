# round 2, session L: wavefront form of the recursive integrators: parity (full suite) + throughput
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02l_gpu_tests.log 2>&1; tail -4 gpurun_out/r02l_gpu_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02l_gpu_tests.log | head -20
timeout 600 python scripts/bench_recursion.py 64 > gpurun_out/r02l_recursion.txt 2>&1; cat gpurun_out/r02l_recursion.txt
