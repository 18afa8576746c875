# round 2, session E: code-size-oriented shade + unified boolean query: parity and A/B
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02e_gpu_tests.log 2>&1; tail -5 gpurun_out/r02e_gpu_tests.log
bash scripts/gpu_ab.sh cur nopf nopf_mb5 nopf_mb6 mb3 u1_mb5 2>&1 | grep -v Traceback | tail -8
