mkdir -p gpurun_out
python -m pytest tests/test_gpu_film_stage.py -x -q > gpurun_out/film_tests.log 2>&1; tail -5 gpurun_out/film_tests.log
python scripts/bench_film_stage.py > gpurun_out/film_bench.txt 2>&1; cat gpurun_out/film_bench.txt
