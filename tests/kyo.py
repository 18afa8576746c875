"""ctypes binding of oracle/kyo.c (the C restatement oracle).  Test infrastructure only."""
import ctypes as C
import os

import numpy as np

import ky_b200 as ky

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_build", "libkyo.so")

_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(LIB)
        l.kyo_render.argtypes = [C.POINTER(ky.SceneDesc), C.POINTER(ky.RenderDesc), C.c_void_p, C.POINTER(C.c_uint64)]
        l.kyo_plastic_random.restype = C.c_float
        _lib = l
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def render(scene, desc):
    """-> (film[h,w,3] float32, rays)"""
    film = np.zeros((desc.height, desc.width, 3), np.float32)
    rays = C.c_uint64(0)
    lib().kyo_render(scene.desc_ptr, C.byref(desc), film.ctypes.data_as(C.c_void_p), C.byref(rays))
    return film, rays.value


def sampler_floats(kind, seed, x, y, sample_index, n):
    out = np.zeros(n, np.float32)
    lib().kyo_sampler_floats(C.c_int(kind), C.c_uint64(seed), C.c_int(x), C.c_int(y), C.c_int(sample_index), C.c_int(n), _fp(out))
    return out


def plastic_random(position, wo):
    p = np.ascontiguousarray(position, np.float32)
    w = np.ascontiguousarray(wo, np.float32)
    return float(lib().kyo_plastic_random(_fp(p), _fp(w)))


def shape_intersect(shape, rays):
    rays = np.ascontiguousarray(rays, np.float32)
    out = np.zeros((rays.shape[0], 8), np.float32)
    lib().kyo_shape_intersect(C.byref(shape), C.c_int(rays.shape[0]), _fp(rays), _fp(out))
    return out


def shape_sample_direction(shape, inp):
    inp = np.ascontiguousarray(inp, np.float32)
    out = np.zeros((inp.shape[0], 7), np.float32)
    lib().kyo_shape_sample_direction(C.byref(shape), C.c_int(inp.shape[0]), _fp(inp), _fp(out))
    return out


def shape_pdf_direction(shape, inp):
    inp = np.ascontiguousarray(inp, np.float32)
    out = np.zeros(inp.shape[0], np.float32)
    lib().kyo_shape_pdf_direction(C.byref(shape), C.c_int(inp.shape[0]), _fp(inp), _fp(out))
    return out


def material_bsdf(material, inp):
    inp = np.ascontiguousarray(inp, np.float32)
    out = np.zeros((inp.shape[0], 13), np.float32)
    lib().kyo_material_bsdf(C.byref(material), C.c_int(inp.shape[0]), _fp(inp), _fp(out))
    return out


def camera_rays(scene, pts):
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.zeros((pts.shape[0], 6), np.float32)
    lib().kyo_camera_rays(C.byref(scene.desc.camera), C.c_int(pts.shape[0]), _fp(pts), _fp(out))
    return out


def light_sample(scene, light_index, inp):
    inp = np.ascontiguousarray(inp, np.float32)
    out = np.zeros((inp.shape[0], 11), np.float32)
    lib().kyo_light_sample(scene.desc_ptr, C.c_int(light_index), C.c_int(inp.shape[0]), _fp(inp), _fp(out))
    return out


def film_encode(fmt, film):
    """Body bytes of the reference's ppm (values) / bmp / hdr file for film[h, w, 3]."""
    film = np.ascontiguousarray(film, np.float32)
    h, w = film.shape[:2]
    lib().kyo_film_body_bytes.restype = C.c_int64
    n = lib().kyo_film_body_bytes(C.c_int(fmt), C.c_int(w), C.c_int(h))
    assert n >= 0
    out = np.zeros(n, np.uint8)
    rc = lib().kyo_film_encode(C.c_int(fmt), C.c_int(w), C.c_int(h), film.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def gamma_encoding(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(x.size, np.uint8)
    lib().kyo_gamma_encoding(C.c_int64(x.size), x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out.reshape(x.shape)


def smallpt_f64(width, height, samples_per_pixel):
    """C restatement of the reference's FP64 smallpt (oracle/smallpt_f64.c): film[h, w, 3] float64, rows bottom-up."""
    out = np.zeros((height, width, 3), np.float64)
    rc = lib().kyo_smallpt_f64(C.c_int(width), C.c_int(height), C.c_int(samples_per_pixel), out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out
