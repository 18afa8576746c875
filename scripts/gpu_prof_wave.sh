set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/launches_wave.csv python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_wave.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_light_sample|k_shadow|k_raygen|k_accumulate" -s 60 -c 14 -o gpurun_out/prof_wave python bench.py --steps 1 --warmup 1 --spp-per-step 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/ncu_wave_full.log 2>&1
ls -la gpurun_out
