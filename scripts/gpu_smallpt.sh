mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_smallpt_f64.py -x -q 2>&1 | tail -8
timeout 300 python - <<'PY'
import sys, time; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, ky_b200 as ky, kyo
d = ky.Device(0)
for w, h, spp in ((1024, 768, 64), (1024, 768, 256)):
    d.render_smallpt_f64(w, h, 4)
    f = d.render_smallpt_f64(w, h, spp); st = d.stats()
    print(f"smallpt f64 {w}x{h}@{spp}: {st.device_ms:.1f} ms  {w*h*spp/st.device_ms/1e3:.1f} Msamples/s  mean {f.mean():.6f}")
t = time.time(); want = kyo.smallpt_f64(1024, 768, 8); dt = time.time() - t
print(f"oracle (C port, all host cores) 1024x768@8: {dt:.2f} s  {1024*768*8/dt/1e6:.2f} Msamples/s")
got = d.render_smallpt_f64(1024, 768, 8)
err = np.abs(got - want).max(axis=-1)
print("1024x768@8 vs oracle: pixels > 1e-9:", int((err > 1e-9).sum()), "max", float(err.max()), "median", float(np.median(err)))
PY
