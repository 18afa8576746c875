"""ky_b200 -- Python binding of the ky-b200 rendering core.

The product is native: ``lib/libkyd.so`` (sm_100a CUDA kernels behind the C ABI of ``include/kyd.h``)
and ``lib/libky_host.so`` (the C++20 host class surface of ``include/ky.hpp`` that mirrors the
reference's scene / integrator classes and entry points).  This module only maps those two libraries
into Python with ctypes so that tests and ``bench.py`` can drive them; it contains no rendering code
and no fallback: if the CUDA library is missing or no GPU is present, calls raise.
"""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")

# ---- enums of include/kyd.h ----------------------------------------------------------------------
SHAPE_SPHERE, SHAPE_RECTANGLE, SHAPE_TRIANGLE, SHAPE_DISK = 0, 1, 2, 3
MAT_MATTE, MAT_MIRROR, MAT_GLASS, MAT_PLASTIC = 0, 1, 2, 3
LIGHT_POINT, LIGHT_DIRECTION, LIGHT_AREA, LIGHT_ENVIRONMENT = 0, 1, 2, 3

INT_POSITION, INT_NORMAL, INT_BASECOLOR = 0, 1, 2
INT_DIRECT_LIGHTING = 6
INT_SIMPLE_PT_RECURSION, INT_PT_RECURSION, INT_PT_RECURSION_DEFERED, INT_PT_ITERATION = 8, 9, 10, 11

DS_IDLE, DS_BSDF, DS_LIGHT, DS_BSDF_MIS, DS_LIGHT_MIS, DS_BOTH_MIS = 0, 4, 8, 16, 32, 48
LIGHTING_EMIT, LIGHTING_DIRECT, LIGHTING_INDIRECT, LIGHTING_ALL = 1, 2, 4, 31
SAMPLER_LCG48, SAMPLER_DEBUG, SAMPLER_TRAPEZOIDAL = 0, 1, 2
FLAG_CLAMP, FLAG_FUSED, FLAG_ACCUMULATE, FLAG_SPLIT_LIGHT_SAMPLE = 1, 2, 4, 8

# scenes of ky_host_scene_create
SCENE_CORNELL, SCENE_VEACH, SCENE_SMALLPT, SCENE_SHAPES = 0, 1, 2, 3
# cornell_box_enum_t (reference ky.cpp:3121-3144)
CB_LIGHT_AREA, CB_LIGHT_DIRECTION, CB_LIGHT_POINT, CB_LIGHT_ENVIRONMENT = 1, 2, 4, 8
CB_LARGE_MIRROR, CB_LARGE_GLASS, CB_SMALL_MIRROR, CB_SMALL_GLASS = 16, 32, 64, 128
CB_BOTH_SMALL = CB_SMALL_MIRROR | CB_SMALL_GLASS
CB_DEFAULT = CB_BOTH_SMALL | CB_LIGHT_AREA

_f3 = C.c_float * 3


class Shape(C.Structure):
    _fields_ = [("kind", C.c_int32), ("p0", _f3), ("p1", _f3), ("p2", _f3), ("p3", _f3), ("normal", _f3),
                ("radius", C.c_float), ("radius_sq", C.c_float), ("area", C.c_float)]


class Material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("diffuse", _f3), ("specular", _f3), ("transmission", _f3),
                ("eta", C.c_float), ("exponent", C.c_float),
                ("diffuse_probability", C.c_float), ("specular_probability", C.c_float)]


class Light(C.Structure):
    _fields_ = [("kind", C.c_int32), ("color", _f3), ("position", _f3), ("direction", _f3),
                ("world_radius", C.c_float), ("shape", C.c_int32)]


class Surface(C.Structure):
    _fields_ = [("shape", C.c_int32), ("material", C.c_int32), ("area_light", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("position", _f3), ("front", _f3), ("right", _f3), ("up", _f3),
                ("resolution", C.c_float * 2), ("origin_push", C.c_float)]


class SceneDesc(C.Structure):
    _fields_ = [("camera", Camera),
                ("shape_count", C.c_int32), ("shapes", C.POINTER(Shape)),
                ("material_count", C.c_int32), ("materials", C.POINTER(Material)),
                ("light_count", C.c_int32), ("lights", C.POINTER(Light)),
                ("surface_count", C.c_int32), ("surfaces", C.POINTER(Surface)),
                ("environment_light", C.c_int32)]


class RenderDesc(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32),
                ("sample_begin", C.c_int32), ("sample_end", C.c_int32),
                ("integrator", C.c_int32), ("max_depth", C.c_int32), ("direct_sample", C.c_int32),
                ("lighting", C.c_int32), ("sampler", C.c_int32),
                ("seed", C.c_uint64), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("rays", C.c_uint64), ("rays_traced", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("device_ms", C.c_double), ("stage_ms", C.c_double * 8),
                ("shade_vertices", C.c_uint64), ("shade_light_lines", C.c_uint64), ("intersect_rays", C.c_uint64)]


class EntryParams(C.Structure):
    _fields_ = [("sub_width", C.c_int), ("sub_height", C.c_int), ("spp", C.c_int), ("depth", C.c_int)]


KYD_SYMBOLS = ["kyd_create", "kyd_create_multi", "kyd_device_count", "kyd_destroy", "kyd_last_error", "kyd_upload_scene", "kyd_render",
               "kyd_render_device", "kyd_clamp_device", "kyd_get_stats", "kyd_set_wave_paths", "kyd_selftest", "kyd_kat",
               "kyd_film_body_bytes", "kyd_film_header", "kyd_film_encode", "kyd_film_encode_device", "kyd_render_smallpt_f64"]

# kyd_kat_kind
KAT_SHAPE_INTERSECT, KAT_SHAPE_SAMPLE_DIRECTION, KAT_SHAPE_PDF_DIRECTION, KAT_MATERIAL_BSDF = 0, 1, 2, 3
KAT_CAMERA_RAYS, KAT_LIGHT_SAMPLE, KAT_SAMPLER = 4, 5, 6
TRAITS_ANY, TRAITS_AREA_RECTANGLE, TRAITS_AREA_SPHERE = 0, 1, 2

# kyd_film_format
FILM_GAMMA8, FILM_BMP24, FILM_RGBE = 0, 1, 2

_kyd = None
_host = None


def kyd_path():
    # KYD_LIB: an alternative build of the CUDA library (A/B measurements of kernel variants)
    return os.environ.get("KYD_LIB") or os.path.join(LIB_DIR, "libkyd.so")


def host_path():
    return os.path.join(LIB_DIR, "libky_host.so")


def kyd():
    """libkyd.so (the CUDA library).  Raises if it has not been built: there is no fallback."""
    global _kyd
    if _kyd is None:
        if not os.path.exists(kyd_path()):
            raise RuntimeError(f"{kyd_path()} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        l = C.CDLL(kyd_path(), mode=C.RTLD_GLOBAL)
        l.kyd_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        l.kyd_create_multi.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int]
        l.kyd_device_count.argtypes = [C.c_void_p]
        l.kyd_destroy.argtypes = [C.c_void_p]
        l.kyd_destroy.restype = None
        l.kyd_last_error.argtypes = [C.c_void_p]
        l.kyd_last_error.restype = C.c_char_p
        l.kyd_upload_scene.argtypes = [C.c_void_p, C.POINTER(SceneDesc)]
        l.kyd_render.argtypes = [C.c_void_p, C.POINTER(RenderDesc), C.c_void_p]
        l.kyd_render_device.argtypes = [C.c_void_p, C.POINTER(RenderDesc), C.c_void_p, C.c_void_p]
        l.kyd_clamp_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        l.kyd_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
        l.kyd_set_wave_paths.argtypes = [C.c_void_p, C.c_int64]
        l.kyd_selftest.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)]
        l.kyd_kat.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        l.kyd_render_smallpt_f64.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.kyd_film_body_bytes.argtypes = [C.c_int, C.c_int, C.c_int]
        l.kyd_film_body_bytes.restype = C.c_int64
        l.kyd_film_header.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        l.kyd_film_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        l.kyd_film_encode_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _kyd = l
    return _kyd


def host():
    """libky_host.so (C++ host surface).  Loads libkyd.so first: the host surface links against it."""
    global _host
    if _host is None:
        kyd()
        l = C.CDLL(host_path())
        l.ky_host_scene_create.restype = C.c_void_p
        l.ky_host_scene_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int]
        l.ky_host_scene_desc.restype = C.POINTER(SceneDesc)
        l.ky_host_scene_desc.argtypes = [C.c_void_p]
        l.ky_host_scene_destroy.restype = None
        l.ky_host_scene_destroy.argtypes = [C.c_void_p]
        l.ky_host_last_error.restype = C.c_char_p
        l.ky_host_render_entry.argtypes = [C.c_char_p, C.POINTER(EntryParams), C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _host = l
    return _host


class Scene:
    """A flattened scene produced by the C++ host surface (scene_t::flatten of include/ky.hpp)."""

    def __init__(self, scene_id, width, height, flags=CB_DEFAULT):
        self._h = host().ky_host_scene_create(scene_id, flags, width, height)
        if not self._h:
            raise RuntimeError(host().ky_host_last_error().decode())
        self.desc_ptr = host().ky_host_scene_desc(self._h)
        self.desc = self.desc_ptr.contents
        self.width, self.height = width, height

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                host().ky_host_scene_destroy(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass

    @property
    def shapes(self):
        return [self.desc.shapes[i] for i in range(self.desc.shape_count)]

    @property
    def materials(self):
        return [self.desc.materials[i] for i in range(self.desc.material_count)]

    @property
    def lights(self):
        return [self.desc.lights[i] for i in range(self.desc.light_count)]

    @property
    def surfaces(self):
        return [self.desc.surfaces[i] for i in range(self.desc.surface_count)]


def render_desc(width, height, spp, integrator=INT_PT_ITERATION, max_depth=5, direct_sample=DS_BOTH_MIS,
                sample_begin=0, sample_end=None, lighting=LIGHTING_ALL, sampler=SAMPLER_LCG48, seed=1234,
                flags=FLAG_CLAMP):
    if sample_end is None:
        sample_end = max(1, spp)
    return RenderDesc(width, height, spp, sample_begin, sample_end, integrator, max_depth, direct_sample,
                      lighting, sampler, seed, flags, 0)


def describe_shape(kind, params):
    """kyd_shape of a shape built by the host classes' constructor (include/ky.hpp)."""
    p = np.ascontiguousarray(params, np.float32)
    out = Shape()
    if host().ky_host_shape_describe(C.c_int(kind), p.ctypes.data_as(C.c_void_p), C.byref(out)) != 0:
        raise RuntimeError(host().ky_host_last_error().decode())
    return out


def describe_material(kind, params):
    p = np.ascontiguousarray(params, np.float32)
    out = Material()
    if host().ky_host_material_describe(C.c_int(kind), p.ctypes.data_as(C.c_void_p), C.byref(out)) != 0:
        raise RuntimeError(host().ky_host_last_error().decode())
    return out


def host_sampler_floats(seed, x, y, sample_index, n):
    out = np.zeros(n, np.float32)
    host().ky_host_sampler_floats(C.c_uint64(seed), C.c_int(x), C.c_int(y), C.c_int(sample_index), C.c_int(n), out.ctypes.data_as(C.c_void_p))
    return out


class Device:
    """One kyd context: one CUDA device and one stream, or -- `device` a list of ordinals -- a multi-GPU context
    (kyd_create_multi) whose renders split the sample range over the devices and sum the partial films on the first."""

    def __init__(self, device=0):
        self._ctx = C.c_void_p()
        if isinstance(device, (list, tuple)):
            arr = (C.c_int * len(device))(*device)
            rc = kyd().kyd_create_multi(C.byref(self._ctx), arr, len(device))
        else:
            rc = kyd().kyd_create(C.byref(self._ctx), device)
        if rc != 0:
            msg = kyd().kyd_last_error(None).decode()
            self._ctx = None
            raise RuntimeError(f"kyd_create failed ({rc}): {msg}")

    @property
    def device_count(self):
        return kyd().kyd_device_count(self._ctx)

    def close(self):
        if self._ctx:
            kyd().kyd_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"kyd error {rc}: {kyd().kyd_last_error(self._ctx).decode()}")

    def upload(self, scene):
        self._scene = scene  # keep the arrays alive
        self._check(kyd().kyd_upload_scene(self._ctx, scene.desc_ptr))

    def render(self, desc, out=None):
        """Host-buffer render (the call integrator_t::render makes). Returns film[h, w, 3] float32."""
        if out is None:
            out = np.empty((desc.height, desc.width, 3), np.float32)
        self._check(kyd().kyd_render(self._ctx, C.byref(desc), out.ctypes.data_as(C.c_void_p)))
        return out

    def render_device(self, desc, device_ptr, stream=None):
        self._check(kyd().kyd_render_device(self._ctx, C.byref(desc), C.c_void_p(device_ptr), C.c_void_p(stream or 0)))

    def clamp_device(self, device_ptr, n, stream=None):
        self._check(kyd().kyd_clamp_device(self._ctx, C.c_void_p(device_ptr), n, C.c_void_p(stream or 0)))

    def film_encode(self, film, fmt):
        """Film output stage: host film[h, w, 3] float32 -> body bytes (uint8) of the reference's ppm / bmp / hdr file."""
        film = np.ascontiguousarray(film, np.float32)
        h, w = film.shape[:2]
        n = kyd().kyd_film_body_bytes(fmt, w, h)
        if n < 0:
            raise ValueError("unknown film format or empty film")
        out = np.empty(n, np.uint8)
        self._check(kyd().kyd_film_encode(self._ctx, film.ctypes.data_as(C.c_void_p), w, h, fmt, out.ctypes.data_as(C.c_void_p)))
        return out

    def film_encode_device(self, film_ptr, width, height, fmt, out_ptr, stream=None):
        self._check(kyd().kyd_film_encode_device(self._ctx, C.c_void_p(film_ptr), width, height, fmt, C.c_void_p(out_ptr), C.c_void_p(stream or 0)))

    def render_smallpt_f64(self, width, height, samples_per_pixel):
        """FP64 validation mode: the reference's double-precision smallpt.  Returns film[h, w, 3] float64, rows bottom-up."""
        out = np.empty((height, width, 3), np.float64)
        self._check(kyd().kyd_render_smallpt_f64(self._ctx, width, height, samples_per_pixel, out.ctypes.data_as(C.c_void_p)))
        return out

    def set_wave_paths(self, paths):
        self._check(kyd().kyd_set_wave_paths(self._ctx, paths))

    def selftest(self, which, first, count):
        out = (C.c_uint64 * 2)()
        self._check(kyd().kyd_selftest(self._ctx, which, first, count, out))
        return int(out[0]), int(out[1])

    def kat(self, which, inputs, obj=None, index=0, traits=0):
        """Device known-answer harness (kyd_kat): one device function of the path on `inputs` [n, in_floats]."""
        out_floats = {KAT_SHAPE_INTERSECT: 8, KAT_SHAPE_SAMPLE_DIRECTION: 7, KAT_SHAPE_PDF_DIRECTION: 1, KAT_MATERIAL_BSDF: 13,
                      KAT_CAMERA_RAYS: 6, KAT_LIGHT_SAMPLE: 11, KAT_SAMPLER: 16}[which]
        inputs = np.ascontiguousarray(inputs, np.float32)
        n = inputs.shape[0]
        out = np.zeros((n, out_floats) if out_floats > 1 else (n,), np.float32)
        self._check(kyd().kyd_kat(self._ctx, which, C.byref(obj) if obj is not None else None, index, traits, n,
                                  inputs.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out

    def stats(self):
        s = Stats()
        self._check(kyd().kyd_get_stats(self._ctx, C.byref(s)))
        return s


def film_header(fmt, width, height):
    """The bytes the reference writes in front of the body (pure host function of libkyd)."""
    buf = (C.c_uint8 * 128)()
    n = kyd().kyd_film_header(fmt, width, height, buf, 128)
    if n < 0:
        raise ValueError("unknown film format or empty film")
    return bytes(buf[:n])


def render_entry(name, sub_width=0, sub_height=0, spp=0, depth=0, render=True):
    """Runs one of the reference-named entry points (include/ky_entry.hpp); returns the film."""
    p = EntryParams(sub_width, sub_height, spp, depth)
    w, h = C.c_int(0), C.c_int(0)
    if host().ky_host_render_entry(name.encode(), C.byref(p), None, C.byref(w), C.byref(h)) != 0:
        raise RuntimeError(host().ky_host_last_error().decode())
    if not render:
        return np.zeros((h.value, w.value, 3), np.float32)
    film = np.zeros((h.value, w.value, 3), np.float32)
    if host().ky_host_render_entry(name.encode(), C.byref(p), film.ctypes.data_as(C.c_void_p), C.byref(w), C.byref(h)) != 0:
        raise RuntimeError(host().ky_host_last_error().decode())
    return film
