mkdir -p gpurun_out
cat > /tmp/dbg.py <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, ky_b200 as ky, cases
d = ky.Device(0)
for name,w,h,seed,scale in cases.STAGE_FILMS:
    f = cases.stage_film(w,h,seed,scale)
    for fmt in (0,1,2):
        try:
            d.film_encode(f, fmt); print(name, fmt, 'ok')
        except Exception as e:
            print(name, fmt, 'ERR', str(e)[-60:])
PY
python /tmp/dbg.py > gpurun_out/dbg.txt 2>&1
compute-sanitizer --print-limit 5 python /tmp/dbg.py > gpurun_out/dbg_san.txt 2>&1
tail -30 gpurun_out/dbg.txt; grep -v "^$" gpurun_out/dbg_san.txt | head -40
