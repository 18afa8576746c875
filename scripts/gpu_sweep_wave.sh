for w in 262144 524288 1048576 2097152 4194304 16777216; do
  KYD_STAGE_TIMING=1 python bench.py --steps 3 --warmup 2 --spp-per-step 16 --e2e-steps 1 --no-cpu-baseline --wave-paths $w 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); st=d['stage_ms_per_step']; print($w, round(d['value'],1), 'Msamples/s  ms/step', round(d['ms_per_step'],1), 'stage sum', round(sum(st.values()),1), {k: round(v,1) for k,v in st.items()}, 'launches', d['gpu_launches'])"
done
