set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_wave.json 2> gpurun_out/bench_wave.err; tail -3 gpurun_out/bench_wave.err; cat gpurun_out/bench_wave.json
