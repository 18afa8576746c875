"""The whole C3 job (Veach 1280x720 @ 1024 spp, PT depth 5 both_mis) against the C oracle, as four 256-spp sample ranges:
every partial film compared bit for bit, ray counts equal.  ~4 minutes of oracle time on 16 host cores."""
import sys, time; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, ky_b200 as ky, kyo
d = ky.Device(0)
w, h, spp = 1280, 720, 1024
scene = ky.Scene(ky.SCENE_VEACH, w, h, 0); d.upload(scene)
tot_diff = 0; tot_rays = 0; t0 = time.time()
for b in range(0, spp, 256):
    desc = ky.render_desc(w, h, spp, max_depth=5, flags=0, sample_begin=b, sample_end=b + 256)
    got = d.render(desc); st = d.stats()
    want, rays = kyo.render(scene, desc)
    diff = (got.view(np.uint32) != want.view(np.uint32)).any(axis=-1)
    tot_diff += int(diff.sum()); tot_rays += rays
    assert st.rays == rays, (st.rays, rays)
    if diff.any():
        print("range", b, "differing pixels", np.argwhere(diff)[:4], got[diff][:2], want[diff][:2])
print(f"C3 veach 1280x720, the whole job: {w*h*spp/1e6:.0f} Msamples, {tot_rays/1e9:.2f} Grays: pixels differing (summed over four 256-spp partial films) {tot_diff}; ray counts equal; {time.time()-t0:.0f} s")
