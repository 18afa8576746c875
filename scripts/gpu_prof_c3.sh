# C3 (Veach) only: --set full capture of one bounce, summarised on the box.   usage: bash scripts/gpu_prof_c3.sh <tag>
tag=${1:-r02}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_shadow|k_accumulate|k_nee" -s 7 -c 9 -o gpurun_out/${tag}_prof_c3 python bench.py --config C3 --steps 1 --warmup 1 --spp-per-step 16 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu_c3_full.log 2>&1
python scripts/ncu_summary.py report gpurun_out/${tag}_prof_c3.ncu-rep > gpurun_out/${tag}_prof_c3_kernels.txt 2>&1
python scripts/ncu_summary.py stalls gpurun_out/${tag}_prof_c3.ncu-rep > gpurun_out/${tag}_prof_c3_stalls.txt 2>&1
python scripts/line_mix.py gpurun_out/${tag}_prof_c3.ncu-rep k_nee 0 70 > gpurun_out/${tag}_prof_c3_lines.txt 2>&1
python scripts/stall_mix.py gpurun_out/${tag}_prof_c3.ncu-rep k_nee 0 > gpurun_out/${tag}_prof_c3_stall_mix.txt 2>&1
rm -f gpurun_out/${tag}_prof_c3.ncu-rep
