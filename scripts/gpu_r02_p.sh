# round 2, session P: intersect fused into the multi-light shade kernels -- full GPU suite, both headline lines, run-time A/B on C3
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02p_gpu_tests.log 2>&1; tail -3 gpurun_out/r02p_gpu_tests.log; grep -E "^(FAILED|ERROR)" gpurun_out/r02p_gpu_tests.log | head
run() {  # tag config
timeout 600 python bench.py --config $2 --no-cpu-baseline --no-configs --steps 4 --warmup 3 --e2e-steps 1 > gpurun_out/r02p_$1_$2.json 2> gpurun_out/r02p_$1_$2.err
python - <<PY
import json
d=json.load(open("gpurun_out/r02p_$1_$2.json"))
print("$1 $2", round(d["value"],1), {k:round(x,1) for k,x in d["stage_ms_per_step"].items()})
PY
}
run cur C5; run cur C3
# (session P only) KYD_FUSE_INTERSECT_MANY=0 run nofusemany C3
