for v in base traits1 traits1_inline; do
  KYD_LIB=$PWD/ky_b200/lib/ab/libkyd_$v.so KYD_STAGE_TIMING=1 python bench.py --steps 3 --warmup 2 --e2e-steps 1 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), 'Msamples/s', {k: round(v,1) for k,v in d['stage_ms_per_step'].items()})"
done
