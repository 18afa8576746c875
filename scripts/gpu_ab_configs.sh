# usage: bash scripts/gpu_ab_configs.sh variant...   (bench_configs.py per variant)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = current ]; then unset KYD_LIB; else export KYD_LIB=$PWD/ky_b200/lib/ab/libkyd_$v.so; fi
  KYD_STAGE_TIMING=1 python scripts/bench_configs.py 16 > gpurun_out/configs_$v.txt 2>&1
  echo "== $v"; grep -v "stage ms" gpurun_out/configs_$v.txt | cut -c1-110
done
