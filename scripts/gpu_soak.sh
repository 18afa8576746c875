# large-sample parity soak: how often does the libm contract (double function rounded once, two implementations) flip a float?
mkdir -p gpurun_out
timeout 1500 python - <<'PY' 2>&1 | tee gpurun_out/soak.txt
import sys, time; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, ky_b200 as ky, kyo
d = ky.Device(0)
for name, sid, fl, w, h, spp, depth in (("C5 cornell 3840x2160", ky.SCENE_CORNELL, ky.CB_DEFAULT, 3840, 2160, 16, 5),
                                        ("C3 veach 1280x720", ky.SCENE_VEACH, 0, 1280, 720, 64, 5),
                                        ("C1 smallpt 1024x768", ky.SCENE_SMALLPT, 0, 1024, 768, 64, 5)):
    scene = ky.Scene(sid, w, h, fl); d.upload(scene)
    desc = ky.render_desc(w, h, spp, max_depth=depth, flags=0)
    got = d.render(desc); st = d.stats()
    t = time.time(); want, rays = kyo.render(scene, desc); dt = time.time() - t
    diff = (got.view(np.uint32) != want.view(np.uint32)).any(axis=-1)
    err = np.abs(got - want).max() if diff.any() else 0.0
    print(f"{name} @ {spp} spp: {w*h*spp/1e6:.1f} Msamples, {rays/1e6:.0f} Mrays; pixels differing {int(diff.sum())} of {w*h}; max abs {err:.3g}; rays device {st.rays} oracle {rays}; oracle {dt:.1f} s")
PY
