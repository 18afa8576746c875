python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Msamples/s', 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'frac', round(d['roofline']['frac'],3))"
