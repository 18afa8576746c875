mkdir -p gpurun_out
for w in 262144 524288 1048576 2097152 4194304 8294400; do
  python bench.py --steps 3 --warmup 2 --spp-per-step 16 --e2e-steps 1 --no-cpu-baseline --wave-paths $w 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($w, round(d['value'],1), 'Msamples/s', d['gpu_launches'])"
done
