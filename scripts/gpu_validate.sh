mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; cut -c1-1200 gpurun_out/bench_now.json
KYD_STAGE_TIMING=1 python scripts/bench_configs.py 16 > gpurun_out/configs_now.txt 2>&1; grep -v "stage ms" gpurun_out/configs_now.txt | cut -c1-112
