mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for mode in wavefront wavefront-split; do
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --mode $mode 2> gpurun_out/bench_$mode.err | tee gpurun_out/bench_$mode.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$mode', round(d['value'],1), 'Msamples/s', 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'frac', round(d['roofline']['frac'],3))"
done
