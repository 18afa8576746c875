"""GPU: a scene with more surfaces than constant memory holds (372 > KYD_MAX_SURFACES): the device answers the three scene
queries through its bounding-volume hierarchy (the accelerator the reference leaves as an empty hook, ky.cpp:3097-3115), the
oracle walks the surface list like the reference.  Films and ray counts must be identical bit for bit."""
import ctypes as C

import numpy as np
import pytest

import cases
import ky_b200 as ky
import kyo

pytestmark = pytest.mark.gpu

W, H = 96, 64


@pytest.fixture(scope="module")
def scene():
    return cases.big_scene()


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("integrator,ds,depth,flags", [
    (ky.INT_PT_ITERATION, ky.DS_BOTH_MIS, 5, 0),
    (ky.INT_PT_ITERATION, ky.DS_BOTH_MIS, 5, ky.FLAG_FUSED),
    (ky.INT_PT_ITERATION, ky.DS_BSDF_MIS, 4, 0),
    (ky.INT_PT_ITERATION, ky.DS_LIGHT, 4, 0),
    (ky.INT_PT_ITERATION, ky.DS_BOTH_MIS, 4, ky.FLAG_SPLIT_LIGHT_SAMPLE),
    (ky.INT_DIRECT_LIGHTING, ky.DS_BOTH_MIS, 0, 0),
    (ky.INT_PT_RECURSION_DEFERED, ky.DS_BOTH_MIS, 3, 0),
    (ky.INT_NORMAL, ky.DS_BOTH_MIS, 0, 0),
    (ky.INT_BASECOLOR, ky.DS_BOTH_MIS, 0, 0),
], ids=["pt", "pt-pixel", "pt-bsdf_mis", "pt-light", "pt-split", "direct", "recursion", "normal", "basecolor"])
def test_bvh_scene_matches_the_linear_walk(device, scene, integrator, ds, depth, flags):
    assert scene.desc.surface_count > 64
    device.upload(scene)
    desc = ky.render_desc(W, H, 4, integrator=integrator, max_depth=depth, direct_sample=ds, flags=flags | ky.FLAG_CLAMP)
    got = device.render(desc)
    want, rays = kyo.render(scene, desc)
    differing = int((_bits(got) != _bits(want)).any(axis=-1).sum())
    assert differing == 0, f"{differing} of {W * H} pixels differ, max abs {np.nanmax(np.abs(got - want))}"
    assert device.stats().rays == rays


def test_small_scene_after_a_large_one(device, scene):
    """The two builds of the kernels keep separate scene bindings: switching back and forth must not mix them up."""
    small = cases.make_scene("cornell")
    desc = ky.render_desc(cases.W, cases.H, 4)
    for s in (scene, small, scene, small):
        device.upload(s)
        got = device.render(desc)
        want, rays = kyo.render(s, desc)
        assert np.array_equal(_bits(got), _bits(want)) and device.stats().rays == rays


def test_surface_limit(device, scene):
    d = scene.desc
    too_many = ky.SceneDesc()
    C.memmove(C.byref(too_many), C.byref(d), C.sizeof(ky.SceneDesc))
    surfaces = (ky.Surface * 4001)(*([d.surfaces[0]] * 4001))
    too_many.surface_count, too_many.surfaces = 4001, surfaces
    rc = ky.kyd().kyd_upload_scene(device._ctx, C.byref(too_many))
    assert rc == 4  # KYD_ERR_LIMIT
