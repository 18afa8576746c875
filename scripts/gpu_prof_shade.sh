mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), 'Msamples/s')"
ncu --set full --clock-control none --import-source on -k regex:"k_shade" -s 8 -c 4 -o gpurun_out/prof_shade python bench.py --steps 1 --warmup 1 --spp-per-step 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_shade.log 2>&1
