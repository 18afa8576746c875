python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 4 --warmup 3 > gpurun_out/bench_line.json 2> gpurun_out/bench_line.err; tail -2 gpurun_out/bench_line.err; cat gpurun_out/bench_line.json
