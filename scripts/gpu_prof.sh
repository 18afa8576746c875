# ncu evidence of the current build: launch list + --set full capture of one bounce (C5), and of C3 when asked.  The reports
# are summarised on the box (gpurun_out/ is merged back only below 64 MiB): per-kernel metrics, launch shares, stall reasons,
# the JSON bench.py quotes.   usage: bash scripts/gpu_prof.sh <tag> [C3]
tag=${1:-r02}
mkdir -p gpurun_out
summarise() { # report-prefix workload-string
  python scripts/ncu_summary.py report gpurun_out/$1.ncu-rep > gpurun_out/$1_kernels.txt 2>&1
  python scripts/ncu_summary.py json gpurun_out/$1.ncu-rep gpurun_out/$1_ncu_summary.json "$2" > /dev/null 2>&1
  python scripts/ncu_summary.py stalls gpurun_out/$1.ncu-rep > gpurun_out/$1_stalls.txt 2>&1
  for k in k_shade k_intersect k_nee k_shadow; do python scripts/line_mix.py gpurun_out/$1.ncu-rep $k 0 60 >> gpurun_out/$1_lines.txt 2>&1; done
  for k in k_shade k_nee; do for c in stall_wait stall_short_sb stall_no_inst stall_long_sb stall_branch_resolving; do
    echo "== $k $c" >> gpurun_out/$1_line_stalls.txt; python scripts/line_stalls.py gpurun_out/$1.ncu-rep $k 0 $c 14 >> gpurun_out/$1_line_stalls.txt 2>&1; done
    python scripts/stall_mix.py gpurun_out/$1.ncu-rep $k 0 >> gpurun_out/$1_stall_mix.txt 2>&1; done
  rm -f gpurun_out/$1.ncu-rep
}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launch_shares.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_shadow|k_accumulate|k_nee" -s 5 -c 9 -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 1 --spp-per-step 2 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu_full.log 2>&1
summarise ${tag}_prof "bench.py --steps 1 --warmup 1 --spp-per-step 2 (C5), launches 5..13"
if [ "$2" = "C3" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${tag}_launches_c3.csv python bench.py --config C3 --steps 1 --warmup 1 --spp-per-step 16 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu_c3.log 2>&1
python scripts/ncu_summary.py launches gpurun_out/${tag}_launches_c3.csv > gpurun_out/${tag}_launch_shares_c3.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_intersect|k_shade|k_shadow|k_accumulate|k_nee" -s 7 -c 9 -o gpurun_out/${tag}_prof_c3 python bench.py --config C3 --steps 1 --warmup 1 --spp-per-step 16 --e2e-steps 1 --no-cpu-baseline --no-configs > gpurun_out/${tag}_ncu_c3_full.log 2>&1
summarise ${tag}_prof_c3 "bench.py --config C3 --steps 1 --warmup 1 --spp-per-step 16 (C3 Veach), launches 7..15"
fi
ls -la gpurun_out/${tag}_*
