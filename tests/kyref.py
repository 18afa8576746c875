"""ctypes binding of oracle/_ref (the REFERENCE compiled by oracle/ref/build_ref.sh).

Test infrastructure only.  Nothing in the product imports this module.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

# integrator_enum_t (ky.cpp:3625-3654)
POSITION, NORMAL, BASECOLOR = 0, 1, 2
DIRECT_LIGHTING = 6
SIMPLE_PT_RECURSION, PT_RECURSION, PT_RECURSION_DEFERED, PT_ITERATION = 8, 9, 10, 11
# direct_sample_enum_t (ky.cpp:3608-3623)
IDLE, BSDF, LIGHT, BSDF_MIS, LIGHT_MIS, BOTH_MIS = 0, 4, 8, 16, 32, 48
# cornell_box_enum_t (ky.cpp:3121-3144)
LIGHT_AREA, LIGHT_DIRECTION, LIGHT_POINT, LIGHT_ENVIRONMENT = 1, 2, 4, 8
LARGE_MIRROR, LARGE_GLASS, SMALL_MIRROR, SMALL_GLASS = 16, 32, 64, 128
BOTH_SMALL = SMALL_MIRROR | SMALL_GLASS
DEFAULT_SCENE = BOTH_SMALL | LIGHT_AREA
# scenes of ref_addon.cpp
CORNELL, VEACH, SMALLPT, SHAPES = 0, 1, 2, 3
# samplers of ref_addon.cpp
RANDOM_SAMPLER, LCG48_SAMPLER, DEBUG_SAMPLER, TRAPEZOIDAL_SAMPLER = 0, 1, 2, 3


class RenderDesc(C.Structure):
    _fields_ = [
        ("scene", C.c_int), ("scene_flags", C.c_int),
        ("width", C.c_int), ("height", C.c_int),
        ("spp", C.c_int), ("sample_offset", C.c_int),
        ("integrator", C.c_int), ("max_depth", C.c_int), ("direct_sample", C.c_int),
        ("sampler", C.c_int), ("threads", C.c_int), ("clamp", C.c_int),
        ("seed", C.c_ulonglong),
    ]


def available(kind="det"):
    return os.path.exists(os.path.join(REF_DIR, f"libky_ref_{kind}.so"))


_libs = {}


def lib(kind="det"):
    if kind not in _libs:
        l = C.CDLL(os.path.join(REF_DIR, f"libky_ref_{kind}.so"))
        l.kyref_render.restype = C.c_int
        l.kyref_shape_area.restype = C.c_float
        l.kyref_plastic_random_of.restype = C.c_float
        _libs[kind] = l
    return _libs[kind]


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def render(scene, width, height, spp, integrator=PT_ITERATION, max_depth=5, direct_sample=BOTH_MIS,
           scene_flags=DEFAULT_SCENE, sampler=LCG48_SAMPLER, seed=1234, sample_offset=0, threads=0,
           kind="det"):
    """Returns (film[h,w,3] float32 -- clamped per pixel as the reference stores it, seconds, rays)."""
    d = RenderDesc(scene, scene_flags, width, height, spp, sample_offset, integrator, max_depth,
                   direct_sample, sampler, threads, 1, seed)
    film = np.zeros((height, width, 3), np.float32)
    sec = C.c_double(0)
    rays = C.c_ulonglong(0)
    rc = lib(kind).kyref_render(C.byref(d), _fp(film), C.byref(sec), C.byref(rays))
    if rc != 0:
        raise RuntimeError("kyref_render failed")
    return film, sec.value, rays.value


def sampler_floats(kind_sampler, seed, x, y, sample_index, n, kind="det"):
    out = np.zeros(n, np.float32)
    lib(kind).kyref_sampler_floats(C.c_int(kind_sampler), C.c_ulonglong(seed), x, y, sample_index, n, _fp(out))
    return out


def plastic_random(position, wo, kind="det"):
    p = np.ascontiguousarray(position, np.float32)
    w = np.ascontiguousarray(wo, np.float32)
    return float(lib(kind).kyref_plastic_random_of(_fp(p), _fp(w)))


def shape_intersect(shape_kind, params, rays, kind="det"):
    params = np.ascontiguousarray(params, np.float32)
    rays = np.ascontiguousarray(rays, np.float32)
    out = np.zeros((rays.shape[0], 8), np.float32)
    lib(kind).kyref_shape_intersect(shape_kind, _fp(params), rays.shape[0], _fp(rays), _fp(out))
    return out


def shape_area(shape_kind, params, kind="det"):
    params = np.ascontiguousarray(params, np.float32)
    return np.float32(lib(kind).kyref_shape_area(shape_kind, _fp(params)))


def shape_sample_direction(shape_kind, params, inp, kind="det"):
    params = np.ascontiguousarray(params, np.float32)
    inp = np.ascontiguousarray(inp, np.float32)
    out = np.zeros((inp.shape[0], 7), np.float32)
    lib(kind).kyref_shape_sample_direction(shape_kind, _fp(params), inp.shape[0], _fp(inp), _fp(out))
    return out


def shape_pdf_direction(shape_kind, params, inp, kind="det"):
    params = np.ascontiguousarray(params, np.float32)
    inp = np.ascontiguousarray(inp, np.float32)
    out = np.zeros(inp.shape[0], np.float32)
    lib(kind).kyref_shape_pdf_direction(shape_kind, _fp(params), inp.shape[0], _fp(inp), _fp(out))
    return out


def material_bsdf(mat_kind, params, inp, kind="det"):
    params = np.ascontiguousarray(params, np.float32)
    inp = np.ascontiguousarray(inp, np.float32)
    out = np.zeros((inp.shape[0], 13), np.float32)
    lib(kind).kyref_material_bsdf(mat_kind, _fp(params), inp.shape[0], _fp(inp), _fp(out))
    return out


def camera_rays(scene, scene_flags, width, height, pts, kind="det"):
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.zeros((pts.shape[0], 6), np.float32)
    lib(kind).kyref_camera_rays(scene, scene_flags, width, height, pts.shape[0], _fp(pts), _fp(out))
    return out


def light_sample(scene, scene_flags, light_index, inp, kind="det"):
    inp = np.ascontiguousarray(inp, np.float32)
    out = np.zeros((inp.shape[0], 11), np.float32)
    rc = lib(kind).kyref_light_sample(scene, scene_flags, light_index, inp.shape[0], _fp(inp), _fp(out))
    if rc != 0:
        raise IndexError("light index")
    return out


def store_film(fmt, width, height, floats, kind="verbatim"):
    """The reference's own store_ppm_impl (0) / store_bmp_impl (1) / store_hdr_impl (2): returns the file's bytes."""
    import tempfile
    floats = np.ascontiguousarray(floats, np.float32)
    assert floats.size == width * height * 3
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "film.out")
        rc = lib(kind).kyref_store_film(C.c_int(fmt), path.encode(), C.c_int(width), C.c_int(height), floats.ctypes.data_as(C.c_void_p))
        assert rc == 0
        with open(path, "rb") as f:
            return f.read()


def gamma_encoding(x, kind="verbatim"):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.size, np.uint8)
    lib(kind).kyref_gamma_encoding(C.c_longlong(x.size), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out.reshape(x.shape)


_smallpt = None


def smallpt_available():
    return os.path.exists(os.path.join(REF_DIR, "libsmallpt_kernel_ref.so"))


def smallpt_f64(width, height, samples_per_pixel):
    """The reference's own smallpt_kernel.cpp (CPU_RENDER), compiled by oracle/ref/build_ref.sh."""
    global _smallpt
    if _smallpt is None:
        _smallpt = C.CDLL(os.path.join(REF_DIR, "libsmallpt_kernel_ref.so"))
    out = np.zeros((height, width, 3), np.float64)
    assert _smallpt.smallpt_ref_render(C.c_int(width), C.c_int(height), C.c_int(samples_per_pixel), out.ctypes.data_as(C.c_void_p)) == 0
    return out


_smallpt_cuda = None


def smallpt_cuda_available():
    return os.path.exists(os.path.join(REF_DIR, "libsmallpt_kernel_cuda_ref.so"))


def smallpt_cuda(width, height, samples_per_pixel, base_stack_bytes=0):
    """The reference's own CUDA kernel (smallpt_kernel.cu, built for sm_100a by oracle/ref/build_ref.sh): returns
    (film[h, w, 3] float64 rows bottom-up, wall-clock seconds of its Device::Render).  Needs a GPU.  The reference's error
    check exits the PROCESS on a CUDA error (and as written the kernel overflows its stack on B200): call it from a child
    process (scripts/ref_cuda_smallpt.py)."""
    global _smallpt_cuda
    if _smallpt_cuda is None:
        _smallpt_cuda = C.CDLL(os.path.join(REF_DIR, "libsmallpt_kernel_cuda_ref.so"))
    out = np.zeros((height, width, 3), np.float64)
    sec = C.c_double(0)
    rc = _smallpt_cuda.smallpt_ref_cuda_render(C.c_int(width), C.c_int(height), C.c_int(samples_per_pixel), out.ctypes.data_as(C.c_void_p), C.byref(sec),
                                               C.c_int(base_stack_bytes))
    assert rc == 0
    return out, sec.value
