"""GPU: the reference-named entry points of the C++ host surface (include/ky_entry.hpp) -- every panel of every grid
equals the oracle's render of that panel's (scene, integrator, strategy), bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import ky_b200 as ky
import kyo

pytestmark = pytest.mark.gpu

SW, SH = 40, 24
LIGHTS = {"point": ky.CB_LIGHT_POINT, "direction": ky.CB_LIGHT_DIRECTION, "area": ky.CB_LIGHT_AREA, "environment": ky.CB_LIGHT_ENVIRONMENT}


def panel(film, row, col):
    return film[row * SH:(row + 1) * SH, col * SW:(col + 1) * SW]


def oracle(scene_id, flags, spp, **kw):
    scene = ky.Scene(scene_id, SW, SH, flags)
    film, _ = kyo.render(scene, ky.render_desc(SW, SH, spp, **kw))
    return film


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_render_debug():
    film = ky.render_entry("render_debug", SW, SH, 2)
    for col, integ in enumerate([ky.INT_POSITION, ky.INT_NORMAL, ky.INT_BASECOLOR]):
        assert same(panel(film, 0, col), oracle(ky.SCENE_VEACH, 0, 2, integrator=integ))


def test_render_mis_scene():
    film = ky.render_entry("render_mis_scene", SW, SH, 2)
    order = [ky.DS_BSDF, ky.DS_LIGHT, ky.DS_IDLE, ky.DS_BSDF_MIS, ky.DS_LIGHT_MIS, ky.DS_BOTH_MIS]  # ky.cpp:4885-4893
    for k, ds in enumerate(order):
        assert same(panel(film, k // 3, k % 3), oracle(ky.SCENE_VEACH, 0, 2, direct_sample=ds)), k


def test_render_multiple_scene():
    film = ky.render_entry("render_multiple_scene", SW, SH, 2)
    for row, ds in enumerate([ky.DS_BSDF, ky.DS_LIGHT, ky.DS_BOTH_MIS]):  # ky.cpp:4829-4836
        for col, light in enumerate(["point", "direction", "area", "environment"]):  # ky.cpp:4821-4827
            assert same(panel(film, row, col), oracle(ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | LIGHTS[light], 2, direct_sample=ds)), (row, col)


def test_render_direct_sample_enum():
    film = ky.render_entry("render_direct_sample_enum", SW, SH, 2)
    for row, light in enumerate(["point", "direction", "area", "environment"]):
        for col, ds in enumerate([ky.DS_BSDF, ky.DS_LIGHT, ky.DS_BSDF_MIS, ky.DS_LIGHT_MIS, ky.DS_BOTH_MIS]):
            assert same(panel(film, row, col), oracle(ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | LIGHTS[light], 2, direct_sample=ds)), (row, col)


def test_render_multiple_integrator():
    film = ky.render_entry("render_multiple_integrator", SW, SH, 2)
    integrators = [ky.INT_DIRECT_LIGHTING, ky.INT_SIMPLE_PT_RECURSION, ky.INT_PT_RECURSION, ky.INT_PT_RECURSION_DEFERED, ky.INT_PT_ITERATION]
    for row, light in enumerate(["point", "direction", "area", "environment"]):
        for col, integ in enumerate(integrators):
            assert same(panel(film, row, col), oracle(ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | LIGHTS[light], 2, integrator=integ)), (row, col)


def test_render_lighting_enum_panels_and_their_sum():
    film = ky.render_entry("render_lighting_enum", SW, SH, 4)
    parts = []
    for col, le in enumerate([ky.LIGHTING_EMIT, ky.LIGHTING_DIRECT, ky.LIGHTING_INDIRECT, ky.LIGHTING_ALL]):
        want = oracle(ky.SCENE_CORNELL, ky.CB_DEFAULT, 4, integrator=ky.INT_PT_RECURSION_DEFERED, max_depth=10, lighting=le)
        assert same(panel(film, 0, col), want), col
        scene = ky.Scene(ky.SCENE_CORNELL, SW, SH, ky.CB_DEFAULT)
        raw, _ = kyo.render(scene, ky.render_desc(SW, SH, 4, integrator=ky.INT_PT_RECURSION_DEFERED, max_depth=10, lighting=le, flags=0))
        parts.append(raw)
    # emit + direct + indirect is the unfiltered estimator, up to FP32 re-association
    total = parts[0] + parts[1] + parts[2]
    assert np.allclose(total, parts[3], rtol=2e-5, atol=1e-6)
    # the emit panel is black except where the camera sees the light; the direct panel has black specular spheres
    assert (parts[0] > 0).any() and (parts[0] == 0).mean() > 0.9


def test_render_single_scene_and_cli(tmp_path):
    film = ky.render_entry("render_single_scene", SW, SH, 2)
    want = oracle(ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_ENVIRONMENT, 2)
    assert same(film, want)
    exe = os.path.join(ky.LIB_DIR, "ky")
    out = subprocess.run([exe, "render_single_scene", "2", str(SW), str(SH)], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    bmp = (tmp_path / "render_single_scene.bmp").read_bytes()
    assert bmp[:2] == b"BM" and len(bmp) == 54 + SW * SH * 3
    # first stored row is the film's bottom row, BGR, gamma 1/2.2 (ky.cpp:1548, 1719-1733)
    px = np.frombuffer(bmp[54:54 + 3], np.uint8)
    expect = [int(np.float64(np.clip(want[SH - 1, 0, c], 0, 1)) ** (1 / 2.2) * 255 + .5) for c in (2, 1, 0)]
    assert list(px) == expect
