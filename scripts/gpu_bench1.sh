set -x
mkdir -p gpurun_out
python bench.py --steps 4 --warmup 3 --mode pixel > gpurun_out/bench_pixel.json 2> gpurun_out/bench_pixel.err; tail -3 gpurun_out/bench_pixel.err; cat gpurun_out/bench_pixel.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_pixel.csv python bench.py --steps 1 --warmup 1 --spp-per-step 4 --e2e-steps 1 --no-cpu-baseline --mode pixel > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/launches_pixel.csv
ncu --set full --clock-control none --import-source on -k regex:k_render_pixels -s 1 -c 1 -o gpurun_out/prof_pixel python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline --mode pixel > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
