import sys, time; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, ky_b200 as ky, kyo
d = ky.Device(0)
w, h, spp = 3840, 2160, 128
scene = ky.Scene(ky.SCENE_CORNELL, w, h, ky.CB_DEFAULT); d.upload(scene)
tot_diff = 0; tot_rays = 0; nan_px = 0
for b in range(0, spp, 32):
    desc = ky.render_desc(w, h, spp, max_depth=5, flags=0, sample_begin=b, sample_end=b + 32)
    got = d.render(desc); st = d.stats()
    want, rays = kyo.render(scene, desc)
    diff = (got.view(np.uint32) != want.view(np.uint32)).any(axis=-1)
    tot_diff += int(diff.sum()); tot_rays += rays; nan_px += int(np.isnan(want).any(axis=-1).sum())
    assert st.rays == rays
    if diff.any():
        print("range", b, "differing pixels", np.argwhere(diff)[:4], got[diff][:2], want[diff][:2])
print(f"C5 cornell 3840x2160, samples {b+32} x 4K = {w*h*spp/1e6:.0f} Msamples, {tot_rays/1e9:.2f} Grays: pixels differing (summed over four 32-spp partial films) {tot_diff}; NaN pixels in the reference films {nan_px}; ray counts equal")
