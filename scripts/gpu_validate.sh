mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; cut -c1-300 gpurun_out/bench_now.json; grep -c . gpurun_out/bench_now.json
KYD_STAGE_TIMING=1 timeout 300 python scripts/bench_configs.py 16 > gpurun_out/configs_now.txt 2>&1; grep -v "stage ms" gpurun_out/configs_now.txt | cut -c1-112
