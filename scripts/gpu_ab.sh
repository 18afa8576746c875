for v in base noprefetch inline slowrsqrt noprefetch_inline noprefetch_inline_mb3 noprefetch_inline_mb5; do
  KYD_LIB=$PWD/ky_b200/lib/ab/libkyd_$v.so python bench.py --steps 4 --warmup 2 --e2e-steps 1 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value'],1), 'Msamples/s')"
done
