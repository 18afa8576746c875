for ds in idle light_mis bsdf_mis both_mis; do
KYD_STAGE_TIMING=1 python bench.py --steps 3 --warmup 2 --e2e-steps 1 --no-cpu-baseline --direct-sample $ds 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$ds', round(d['value'],1), 'Msamples/s', 'ms/step', round(d['ms_per_step'],2), {k: round(v,1) for k,v in d['stage_ms_per_step'].items()}, 'rays/sample', round(d['rays_per_sample'],2))"
done
