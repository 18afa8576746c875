# usage: bash scripts/gpu_ab_both_quick.sh variant...   (bench lines of ky_b200/lib/ab/libkyd_<variant>.so builds on C5 and C3, no tests)
for v in "$@"; do bash scripts/gpu_ab_quick.sh C5 $v; bash scripts/gpu_ab_quick.sh C3 $v; done
