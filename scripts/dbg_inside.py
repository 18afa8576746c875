import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, ky_b200 as ky, cases, kyo
dev = ky.Device(0)
scene = cases.inside_sphere_light_scene(True)
dev.upload(scene)
W, H = 24, 16
import ctypes
scene2 = cases.CustomScene(ky.Scene(ky.SCENE_CORNELL, W, H), scene._keep[1][:], scene._keep[2][:], scene._keep[3][:1], scene._keep[4][:], -1)
dev.upload(scene2)
for integ, depth in ((ky.INT_DIRECT_LIGHTING, 0), (ky.INT_PT_ITERATION, 1)):
    desc = ky.render_desc(W, H, 1, integrator=integ, max_depth=depth, direct_sample=ky.DS_LIGHT, flags=ky.FLAG_CLAMP)
    want, rays = kyo.render(scene2, desc)
    desc.flags = ky.FLAG_CLAMP
    got = dev.render(desc)
    st = dev.stats()
    print("wavefront: rays", st.rays, "oracle", rays, "traced", st.rays_traced, "lines", st.shade_light_lines, "vertices", st.shade_vertices)
    desc.flags = ky.FLAG_CLAMP | ky.FLAG_FUSED
    gp = dev.render(desc); stp = dev.stats()
    print("pixel: bad", int((gp.view(np.uint32) != want.view(np.uint32)).any(axis=-1).sum()), "traced", stp.rays_traced)
    bad = (got.view(np.uint32) != want.view(np.uint32)).any(axis=-1)
    print("integrator", integ, "bad", int(bad.sum()), "lit pixels want", int((want.sum(-1) > 0).sum()), "got", int((got.sum(-1) > 0).sum()))
    ys, xs = np.nonzero(bad)
    for y, x in list(zip(ys, xs))[:6]:
        print("  pixel", x, y, "want", want[y, x], "got", got[y, x])
    print("".join("".join("X" if bad[y, x] else ("o" if want[y, x].sum() > 0 else ".") for x in range(W)) + "\n" for y in range(H)))
