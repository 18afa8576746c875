"""Generates tests/golden/*.npz from the REFERENCE itself (oracle/_ref/libky_ref_det.so, built by
oracle/ref/build_ref.sh from /root/reference/ky.cpp).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures pin the C oracle (and through it the device) where /root/reference does not exist.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import ky_b200 as ky  # noqa: E402
import kyref  # noqa: E402

REF_SCENE = {ky.SCENE_CORNELL: kyref.CORNELL, ky.SCENE_VEACH: kyref.VEACH, ky.SCENE_SMALLPT: kyref.SMALLPT, ky.SCENE_SHAPES: kyref.SHAPES}


def films():
    out = {}
    for name, sk, integ, ds, depth, spp in cases.film_cases():
        sid, flags = cases.SCENES[sk]
        film, _, rays = kyref.render(REF_SCENE[sid], cases.W, cases.H, spp, integrator=integ, max_depth=depth,
                                     direct_sample=ds, scene_flags=flags, sampler=kyref.LCG48_SAMPLER, seed=1234)
        out[name] = film
        out[name + "#rays"] = np.array([rays], np.uint64)
    return out


def film_stage():
    """Files written by the reference's own store_ppm_impl / store_bmp_impl / store_hdr_impl for cases.STAGE_FILMS."""
    out = {}
    for name, w, h, seed, scale in cases.STAGE_FILMS:
        f = cases.stage_film(w, h, seed, scale)
        for fmt, ext in ((0, "ppm"), (1, "bmp"), (2, "hdr")):
            out[f"{name}.{ext}"] = np.frombuffer(kyref.store_film(fmt, w, h, f), np.uint8)
    return out


def trapezoidal():
    """Films of the reference driven by the trapezoidal sampler plug-in (oracle/ref/ref_addon.cpp)."""
    out = {}
    for name, sk, integ, ds, depth, spp in cases.trapezoidal_cases():
        sid, flags = cases.SCENES[sk]
        film, _, rays = kyref.render(REF_SCENE[sid], cases.W, cases.H, spp, integrator=integ, max_depth=depth, direct_sample=ds,
                                     scene_flags=flags, sampler=kyref.TRAPEZOIDAL_SAMPLER, seed=1234)
        out[name] = film
        out[name + "#rays"] = np.array([rays], np.uint64)
    return out


def smallpt_f64():
    """Films of the reference's FP64 smallpt (smallpt2pbrt/smallpt_kernel.cpp compiled by oracle/ref/build_ref.sh)."""
    return {f"{w}x{h}@{spp}": kyref.smallpt_f64(w, h, spp) for w, h, spp in cases.SMALLPT_F64_CASES}


def rng(seed):
    return np.random.default_rng(seed)


def unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


SHAPE_PARAMS = {
    "sphere": (0, [0.3, -0.2, 0.5, 0.7]),
    "rectangle": (1, [-1, -1, 0.2, 1, -1, 0.1, 1, 1, 0.3, -1, 1, 0.4, 0]),
    "rectangle_flipped": (1, [-1, -1, 0.2, 1, -1, 0.1, 1, 1, 0.3, -1, 1, 0.4, 1]),
    "triangle": (2, [-1, -0.8, 0.1, 1.1, -0.9, 0.2, 0.1, 1.2, 0.4, 0]),
    "disk": (3, [0.1, 0.2, 0.3, 0.2, -0.3, 1.0, 0.9]),
}

MATERIAL_PARAMS = {
    "matte": (0, [0.8, 0.5, 0.3, 0, 0, 0, 0]),
    "mirror": (1, [0.9, 0.95, 1.0, 0, 0, 0, 0]),
    "glass": (2, [1, 0.9, 0.8, 0.7, 0.8, 0.9, 1.6]),
    "plastic90": (3, [0.1, 0.1, 0.1, 0.7, 0.7, 0.7, 90]),
    "plastic5000": (3, [0.07, 0.09, 0.13, 1, 1, 1, 5000]),
}


def kats():
    out = {}
    # sampler streams: reference mt19937_64 anchors (verbatim build) and the LCG48 contract
    out["sampler/mt19937_seed1234"] = kyref.sampler_floats(0, 1234, 0, 0, 0, 6, kind="verbatim")
    for (x, y, s) in [(0, 0, 0), (3, 5, 7), (1023, 767, 63), (3839, 2159, 16383)]:
        out[f"sampler/lcg48/{x}_{y}_{s}"] = kyref.sampler_floats(1, 1234, x, y, s, 16)
    out["sampler/lcg48/seed99"] = kyref.sampler_floats(1, 99, 10, 20, 30, 16)

    g = rng(7)
    n = 1500
    # rays aimed at the unit-ish neighbourhood of the shapes, plus grazing / degenerate ones
    o = g.uniform(-3, 3, (n, 3)).astype(np.float32)
    target = g.uniform(-1.2, 1.2, (n, 3)).astype(np.float32)
    d = unit(target - o).astype(np.float32)
    tmax = np.where(g.uniform(size=n) < 0.7, np.inf, g.uniform(0.5, 6, n)).astype(np.float32)
    rays = np.concatenate([o, d, tmax[:, None]], axis=1).astype(np.float32)
    rays[:50, 3:6] = np.array([1, 0, 0], np.float32)  # axis aligned
    rays[50:100, 3:6] = np.array([0, 0, -1], np.float32)
    out["shape/rays"] = rays
    p = g.uniform(-2.5, 2.5, (n, 3)).astype(np.float32)
    p[:200] *= 0.1  # inside the sphere
    nrm = unit(g.normal(size=(n, 3))).astype(np.float32)
    u = g.uniform(size=(n, 2)).astype(np.float32)
    u[:20] = 0.0
    u[20:40, 0] = 0.5
    u[20:40, 1] = 0.5
    wi = unit(target - p).astype(np.float32)
    out["shape/sample_in"] = np.concatenate([p, nrm, u], axis=1).astype(np.float32)
    out["shape/pdf_in"] = np.concatenate([p, nrm, wi], axis=1).astype(np.float32)
    for name, (kind, params) in SHAPE_PARAMS.items():
        out[f"shape/{name}/intersect"] = kyref.shape_intersect(kind, params, rays)
        out[f"shape/{name}/sample_direction"] = kyref.shape_sample_direction(kind, params, out["shape/sample_in"])
        out[f"shape/{name}/pdf_direction"] = kyref.shape_pdf_direction(kind, params, out["shape/pdf_in"])
        out[f"shape/{name}/area"] = np.array([kyref.shape_area(kind, params)], np.float32)

    # BSDFs: {p, n, wo, wi, u}
    wo = unit(g.normal(size=(n, 3))).astype(np.float32)
    wi2 = unit(g.normal(size=(n, 3))).astype(np.float32)
    wi2[: n // 2] = unit(wi2[: n // 2] + 3 * (2 * np.sum(wo[: n // 2] * nrm[: n // 2], axis=1, keepdims=True) * nrm[: n // 2] - wo[: n // 2])).astype(np.float32)
    out["bsdf/in"] = np.concatenate([p, nrm, wo, wi2, u], axis=1).astype(np.float32)
    for name, (kind, params) in MATERIAL_PARAMS.items():
        out[f"bsdf/{name}"] = kyref.material_bsdf(kind, params, out["bsdf/in"])

    # camera rays of every scene
    pts = g.uniform(0, 1, (500, 2)).astype(np.float32) * np.array([cases.W, cases.H], np.float32)
    out["camera/in"] = pts
    for sk, (sid, flags) in cases.SCENES.items():
        out[f"camera/{sk}"] = kyref.camera_rays(REF_SCENE[sid], flags, cases.W, cases.H, pts)

    # lights of the scenes that cover every light kind / shape kind
    lin = np.concatenate([p * 0.5, nrm, u, wi], axis=1).astype(np.float32)[:800]
    out["light/in"] = lin
    for sk in ("cornell", "cornell_large_mirror_all_lights", "veach", "shapes", "smallpt"):
        sid, flags = cases.SCENES[sk]
        i = 0
        while True:
            try:
                out[f"light/{sk}/{i}"] = kyref.light_sample(REF_SCENE[sid], flags, i, lin)
            except IndexError:
                break
            i += 1
    return out


if __name__ == "__main__":
    assert kyref.available("det") and kyref.available("verbatim"), "build oracle/_ref first (oracle/ref/build_ref.sh)"
    if "--film-stage-only" not in sys.argv:
        np.savez_compressed(os.path.join(HERE, "golden_films.npz"), **films())
        np.savez_compressed(os.path.join(HERE, "golden_kat.npz"), **kats())
    np.savez_compressed(os.path.join(HERE, "golden_film_stage.npz"), **film_stage())
    np.savez_compressed(os.path.join(HERE, "golden_smallpt_f64.npz"), **smallpt_f64())
    np.savez_compressed(os.path.join(HERE, "golden_trapezoidal.npz"), **trapezoidal())
    print("wrote", os.listdir(HERE))
