"""Shared list of (scene, integrator, strategy) cases used by the golden generator and the parity tests."""
import ky_b200 as ky

W, H, SPP = 48, 32, 4

SCENES = {
    "cornell": (ky.SCENE_CORNELL, ky.CB_DEFAULT),
    "cornell_point": (ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_POINT),
    "cornell_direction": (ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_DIRECTION),
    "cornell_environment": (ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_ENVIRONMENT),
    "cornell_large_glass": (ky.SCENE_CORNELL, ky.CB_LARGE_GLASS | ky.CB_LIGHT_AREA),
    "cornell_large_mirror_all_lights": (ky.SCENE_CORNELL, ky.CB_LARGE_MIRROR | 15),
    "veach": (ky.SCENE_VEACH, 0),
    "smallpt": (ky.SCENE_SMALLPT, 0),
    "shapes": (ky.SCENE_SHAPES, 0),
}

STRATEGIES = {"idle": ky.DS_IDLE, "bsdf": ky.DS_BSDF, "light": ky.DS_LIGHT, "bsdf_mis": ky.DS_BSDF_MIS,
              "light_mis": ky.DS_LIGHT_MIS, "both_mis": ky.DS_BOTH_MIS}

INTEGRATORS = {"position": ky.INT_POSITION, "normal": ky.INT_NORMAL, "basecolor": ky.INT_BASECOLOR,
               "direct_lighting": ky.INT_DIRECT_LIGHTING, "simple_pt_recursion": ky.INT_SIMPLE_PT_RECURSION,
               "pt_recursion": ky.INT_PT_RECURSION, "pt_recursion_defered": ky.INT_PT_RECURSION_DEFERED,
               "pt_iteration": ky.INT_PT_ITERATION}


def film_cases():
    """(name, scene_key, integrator, direct_sample, max_depth, spp)"""
    out = []
    for sk in SCENES:
        for ik in ("position", "normal", "basecolor"):
            out.append((f"{sk}/{ik}", sk, INTEGRATORS[ik], ky.DS_IDLE, 0, 1))
        for dk, ds in STRATEGIES.items():
            out.append((f"{sk}/pt_iteration/{dk}", sk, ky.INT_PT_ITERATION, ds, 5, SPP))
        for ik in ("direct_lighting", "simple_pt_recursion", "pt_recursion", "pt_recursion_defered"):
            out.append((f"{sk}/{ik}/both_mis", sk, INTEGRATORS[ik], ky.DS_BOTH_MIS, 5, SPP))
    # BASELINE config 4 uses depth 8; config 2 compares strategies under direct lighting
    out.append(("cornell/pt_iteration/both_mis/depth8", "cornell", ky.INT_PT_ITERATION, ky.DS_BOTH_MIS, 8, SPP))
    out.append(("cornell/direct_lighting/bsdf", "cornell", ky.INT_DIRECT_LIGHTING, ky.DS_BSDF, 0, SPP))
    out.append(("cornell/direct_lighting/light", "cornell", ky.INT_DIRECT_LIGHTING, ky.DS_LIGHT, 0, SPP))
    out.append(("veach/pt_iteration/both_mis/depth1", "veach", ky.INT_PT_ITERATION, ky.DS_BOTH_MIS, 1, SPP))
    return out


def make_scene(scene_key, w=W, h=H):
    sid, flags = SCENES[scene_key]
    return ky.Scene(sid, w, h, flags)


# ---- film output stage -------------------------------------------------------------------------------------------
FILM_EDGE_VALUES = [0.0, -0.0, 1.0, 1.5, -3.7, float("nan"), float("inf"), float("-inf"), 1e-33, 1e-32, 9.9e-33, 1e-20,
                    300.0, 1e30, -1e30, 3e9, -3e9, 0.5, 2.0 ** -126, 1e-45, 0.0031308, 0.999999, 1.0000001, 255.0, 256.0]


def stage_film(width, height, seed, scale=False):
    """A film for the output stage: uniform values, the edge values above sprinkled in, optionally wild magnitudes."""
    import numpy as np
    rng = np.random.default_rng(seed)
    f = rng.random((height, width, 3)).astype(np.float32)
    if scale:
        f *= rng.choice(np.array([1, 10, 1e-3, 1e5, -1, 1e-30, 1e25], np.float32), size=f.shape)
    flat = f.reshape(-1)
    edge = np.array(FILM_EDGE_VALUES, np.float32)
    for k in range(0, flat.size - edge.size, max(edge.size + 7, flat.size // 5)):
        flat[k:k + edge.size] = edge if (k // 7) % 2 == 0 else edge[::-1]
    return f


# (name, width, height, seed, scale): odd sizes so that 3*w is not a multiple of 4 and words straddle bmp lines
STAGE_FILMS = [("37x21", 37, 21, 1, False), ("37x21s", 37, 21, 2, True), ("1x1", 1, 1, 3, False), ("5x3", 5, 3, 4, True),
               ("64x48", 64, 48, 5, False), ("3x7", 3, 7, 6, True)]


# ---- hand-assembled scenes (kyd_scene_desc built in Python) ------------------------------------------------------
class CustomScene:
    """A kyd_scene_desc assembled from ctypes arrays; quacks like ky_b200.Scene for Device.upload and kyo.render."""

    def __init__(self, base, shapes, materials, lights, surfaces, environment_light=-1):
        import ctypes as C
        import ky_b200 as ky
        self._keep = (base,
                      (ky.Shape * len(shapes))(*shapes), (ky.Material * len(materials))(*materials),
                      (ky.Light * max(1, len(lights)))(*lights), (ky.Surface * len(surfaces))(*surfaces))
        d = ky.SceneDesc()
        d.camera = base.desc.camera
        d.shape_count, d.shapes = len(shapes), self._keep[1]
        d.material_count, d.materials = len(materials), self._keep[2]
        d.light_count, d.lights = len(lights), self._keep[3]
        d.surface_count, d.surfaces = len(surfaces), self._keep[4]
        d.environment_light = environment_light
        self.desc = d
        self.desc_ptr = C.pointer(d)
        self.width, self.height = base.width, base.height


def coplanar_tie_scene(order):
    """Cornell box + a disk and a triangle lying exactly in the plane of the back wall (axis-aligned, so all three report
    bit-identical hit distances where they overlap), each with its own matte colour.  order = "first": the two come
    BEFORE the Cornell surfaces in the list (they must win the ties), "last": after (the wall must win).  The reference
    resolves equal distances by list order (strict `t < tmax`, ky.cpp:3172-3184); the device visits surfaces grouped by
    kind, so this pins its tie rule."""
    import ky_b200 as ky
    base = make_scene("cornell")
    shapes, materials, lights, surfaces = base.shapes, base.materials, base.lights, base.surfaces
    wall = None
    for i, sf in enumerate(surfaces):
        s = shapes[sf.shape]
        zs = [s.p0[2], s.p1[2], s.p2[2], s.p3[2]]
        if s.kind == ky.SHAPE_RECTANGLE and abs(s.normal[2]) == 1.0 and max(zs) == min(zs) and sf.area_light < 0:
            if wall is None or zs[0] > shapes[surfaces[wall].shape].p0[2]:
                wall = i
    w = shapes[surfaces[wall].shape]
    cx = (w.p0[0] + w.p1[0] + w.p2[0] + w.p3[0]) / 4
    cy = (w.p0[1] + w.p1[1] + w.p2[1] + w.p3[1]) / 4
    disk = ky.Shape()
    disk.kind = ky.SHAPE_DISK
    disk.p0[:] = [cx, cy, w.p0[2]]
    disk.normal[:] = list(w.normal)
    disk.radius, disk.radius_sq, disk.area = 160.0, 160.0 * 160.0, 3.14159274 * 160.0 * 160.0
    tri = ky.Shape()
    tri.kind = ky.SHAPE_TRIANGLE
    tri.p0[:], tri.p1[:], tri.p2[:] = list(w.p0), list(w.p1), list(w.p2)
    tri.normal[:] = list(w.normal)
    tri.area = w.area / 2

    def matte(r, g, b):
        m = ky.Material()
        m.kind = ky.MAT_MATTE
        m.diffuse[:] = [r, g, b]
        return m

    n_s, n_m = len(shapes), len(materials)
    shapes = shapes + [disk, tri]
    materials = materials + [matte(0.9, 0.1, 0.1), matte(0.1, 0.9, 0.1)]
    extra = []
    for k in range(2):
        sf = ky.Surface()
        sf.shape, sf.material, sf.area_light = n_s + k, n_m + k, -1
        extra.append(sf)
    surfaces = extra + surfaces if order == "first" else surfaces + extra
    return CustomScene(base, shapes, materials, lights, surfaces, base.desc.environment_light)


def light_tie_scene(order):
    """Veach scene + a matte copy of the first sphere light's surface (same shape: bit-identical hit distances), listed
    before ("first": the copy hides the light from every BSDF-sampled ray) or after ("last": the light wins) the others.
    Pins the tie rule of the occlusion form of BSDF-sampled light queries (scene_blocked_before)."""
    import ky_b200 as ky
    base = make_scene("veach")
    shapes, materials, lights, surfaces = base.shapes, base.materials, base.lights, base.surfaces
    lit = next(i for i, sf in enumerate(surfaces) if sf.area_light >= 0 and shapes[sf.shape].kind == ky.SHAPE_SPHERE)
    m = ky.Material()
    m.kind = ky.MAT_MATTE
    m.diffuse[:] = [0.2, 0.3, 0.9]
    copy = ky.Surface()
    copy.shape, copy.material, copy.area_light = surfaces[lit].shape, len(materials), -1
    surfaces = [copy] + surfaces if order == "first" else surfaces + [copy]
    return CustomScene(base, shapes, materials + [m], lights, surfaces, base.desc.environment_light)


# ---- FP64 smallpt validation mode: (width, height, samples per pixel) ----------------------------------------------
SMALLPT_F64_CASES = [(96, 72, 16), (64, 48, 40), (33, 17, 5)]


# ---- trapezoidal (tent-filter, 2x2 sub-pixel) sampler: (name, scene key, integrator, direct sample, depth, spp) -------
def trapezoidal_cases():
    import ky_b200 as ky
    return [("trap/cornell/pt", "cornell", ky.INT_PT_ITERATION, ky.DS_BOTH_MIS, 5, 8),
            ("trap/veach/pt", "veach", ky.INT_PT_ITERATION, ky.DS_BOTH_MIS, 3, 4),
            ("trap/smallpt/pt", "smallpt", ky.INT_PT_ITERATION, ky.DS_BOTH_MIS, 5, 4),
            ("trap/cornell/direct", "cornell", ky.INT_DIRECT_LIGHTING, ky.DS_LIGHT_MIS, 0, 4),
            ("trap/shapes/normal", "shapes", ky.INT_NORMAL, ky.DS_BOTH_MIS, 0, 4),
            ("trap/cornell/recursion", "cornell", ky.INT_PT_RECURSION, ky.DS_BOTH_MIS, 3, 4)]


def big_scene(extra=360, seed=3):
    """Cornell box + `extra` small spheres / triangles / rectangles / disks scattered inside it (matte, mirror, glass and
    plastic): more surfaces than constant memory holds (KYD_MAX_SURFACES), so the device walks its bounding-volume
    hierarchy while the oracle walks the list.  Shapes are built by the host classes' constructors (ky.describe_shape)."""
    import numpy as np
    import ky_b200 as ky
    base = make_scene("cornell")
    shapes, materials, lights, surfaces = base.shapes, base.materials, base.lights, base.surfaces
    rng = np.random.default_rng(seed)

    def unit(v):
        return v / np.linalg.norm(v)

    n_mat = len(materials)
    for kind, params in ((0, [0.7, 0.6, 0.2, 0, 0, 0, 0]), (1, [0.9, 0.9, 0.9, 0, 0, 0, 0]), (2, [1, 1, 1, 1, 1, 1, 1.5]),
                         (3, [0.2, 0.3, 0.1, 0.6, 0.6, 0.6, 90])):
        materials = materials + [ky.describe_material(kind, params)]
    for k in range(extra):
        c = rng.uniform([40, 30, 40], [510, 500, 520])
        size = rng.uniform(6, 28)
        kind = int(rng.integers(0, 4))
        if kind == ky.SHAPE_SPHERE:
            params = [*c, size]
        elif kind == ky.SHAPE_DISK:
            params = [*c, *unit(rng.normal(size=3)), size]
        else:
            u = unit(rng.normal(size=3)) * size
            v = unit(np.cross(u, rng.normal(size=3))) * size * rng.uniform(0.5, 1.5)
            if kind == ky.SHAPE_TRIANGLE:
                params = [*c, *(c + u), *(c + v), float(k % 2)]
            else:
                params = [*c, *(c + u), *(c + u + v), *(c + v), float(k % 2)]
        sf = ky.Surface()
        sf.shape, sf.material, sf.area_light = len(shapes), n_mat + int(rng.integers(0, 4)), -1
        shapes = shapes + [ky.describe_shape(kind, params)]
        surfaces = surfaces + [sf]
    return CustomScene(base, shapes, materials, lights, surfaces, base.desc.environment_light)


def inside_sphere_light_scene(only_light=False):
    """Cornell box inside a sphere area light of radius 6 around the origin (the camera sits inside it too): every shading
    point takes the inside branch of sphere_t::sample_direction -- area sampling with the SHADING point's normal in the
    pdf (ky.cpp:1422-1444, a reference quirk) -- and of sphere_t::pdf_direction, which falls through to the generic
    re-intersection (ky.cpp:1503-1513, 1055-1090).  No reference scene reaches them.  only_light: the sphere is the scene's
    single light (the specialised single-sphere-light kernels); else it is a second light next to the ceiling rectangle."""
    import ky_b200 as ky
    base = make_scene("cornell")
    shapes, materials, lights, surfaces = base.shapes, base.materials, base.lights, base.surfaces
    ball = ky.describe_shape(ky.SHAPE_SPHERE, [0.0, 0.0, 0.0, 6.0])
    glow = ky.Light()
    glow.kind = ky.LIGHT_AREA
    glow.color[:] = [0.35, 0.3, 0.25]
    glow.shape = len(shapes)
    shell = ky.Surface()
    shell.shape, shell.material = len(shapes), surfaces[2].material
    if only_light:
        kept = []
        for sf in surfaces:
            c = ky.Surface()
            c.shape, c.material, c.area_light = sf.shape, sf.material, -1
            kept.append(c)
        shell.area_light = 0
        return CustomScene(base, shapes + [ball], materials, [glow], kept + [shell], -1)
    shell.area_light = len(lights)
    return CustomScene(base, shapes + [ball], materials, lights + [glow], surfaces + [shell], base.desc.environment_light)
