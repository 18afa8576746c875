mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Msamples/s')"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cur.csv python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_cur.log 2>&1
