"""GPU: the device functions of the path, one by one, against fixtures generated from the REFERENCE build
(tests/golden/golden_kat.npz, tests/golden/make_golden.py) -- the same vectors that pin the C oracle in
tests/test_oracle_golden.py, run through kyd_kat on the device.  Bit-exact bar.  Reaches branches no film reaches: shading
points inside a sphere (the first 200 sample / pdf inputs), total internal reflection, the disk's parallel-ray reject,
the inf -> 0 pdf guards.  Plus films of a scene whose every shading point lies inside a sphere light."""
import os

import numpy as np
import pytest

import cases
import ky_b200 as ky
import kyo
from golden.make_golden import MATERIAL_PARAMS, SHAPE_PARAMS

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KAT = np.load(os.path.join(HERE, "golden", "golden_kat.npz"))


def mismatch(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    assert a.shape == b.shape
    return int((a != b).sum())


@pytest.mark.parametrize("name", list(SHAPE_PARAMS))
def test_device_shape_functions(device, name):
    kind, params = SHAPE_PARAMS[name]
    shape = ky.describe_shape(kind, params)
    got = device.kat(ky.KAT_SHAPE_INTERSECT, KAT["shape/rays"], shape)
    want = KAT[f"shape/{name}/intersect"]
    hit = want[:, 0] == 1
    assert mismatch(got[:, 0], want[:, 0]) == 0
    assert mismatch(got[hit], want[hit]) == 0 and hit.sum() > 50
    # the specialised instantiations compute the same function for the shape kind they are compiled for
    traits = [ky.TRAITS_ANY] + ([ky.TRAITS_AREA_SPHERE] if kind == ky.SHAPE_SPHERE else []) + ([ky.TRAITS_AREA_RECTANGLE] if kind == ky.SHAPE_RECTANGLE else [])
    for t in traits:
        assert mismatch(device.kat(ky.KAT_SHAPE_SAMPLE_DIRECTION, KAT["shape/sample_in"], shape, traits=t), KAT[f"shape/{name}/sample_direction"]) == 0
        assert mismatch(device.kat(ky.KAT_SHAPE_PDF_DIRECTION, KAT["shape/pdf_in"], shape, traits=t), KAT[f"shape/{name}/pdf_direction"]) == 0


def test_inputs_reach_the_inside_sphere_branches():
    kind, params = SHAPE_PARAMS["sphere"]
    p = KAT["shape/sample_in"][:, :3]
    inside = ((p - np.array(params[:3], np.float32)) ** 2).sum(axis=1) <= params[3] ** 2
    assert inside.sum() >= 100 and (~inside).sum() >= 1000


@pytest.mark.parametrize("name", list(MATERIAL_PARAMS))
def test_device_bsdf_functions(device, name):
    kind, params = MATERIAL_PARAMS[name]
    material = ky.describe_material(kind, params)
    got = device.kat(ky.KAT_MATERIAL_BSDF, KAT["bsdf/in"], material)
    want = KAT[f"bsdf/{name}"]
    assert mismatch(got, want) == 0
    if name == "glass":   # both outcomes of the dielectric occur, and total internal reflection among the refraction attempts
        types = set(want[:, 7].astype(int).tolist())
        assert {17, 18} <= types


@pytest.mark.parametrize("sk", list(cases.SCENES))
def test_device_camera_rays(device, sk):
    device.upload(cases.make_scene(sk))
    assert mismatch(device.kat(ky.KAT_CAMERA_RAYS, KAT["camera/in"]), KAT[f"camera/{sk}"]) == 0


@pytest.mark.parametrize("key", [k for k in KAT.files if k.startswith("light/") and k != "light/in"])
def test_device_light_sampling(device, key):
    _, sk, idx = key.split("/")
    scene = cases.make_scene(sk)
    device.upload(scene)
    traits = [ky.TRAITS_ANY]
    kinds = {(l.kind, scene.shapes[l.shape].kind if l.kind == ky.LIGHT_AREA else -1) for l in scene.lights}
    if kinds == {(ky.LIGHT_AREA, ky.SHAPE_SPHERE)}:
        traits.append(ky.TRAITS_AREA_SPHERE)
    if kinds == {(ky.LIGHT_AREA, ky.SHAPE_RECTANGLE)} and len(scene.lights) == 1:
        traits.append(ky.TRAITS_AREA_RECTANGLE)
    for t in traits:
        assert mismatch(device.kat(ky.KAT_LIGHT_SAMPLE, KAT["light/in"], index=int(idx), traits=t), KAT[key]) == 0


@pytest.mark.parametrize("key", [k for k in KAT.files if k.startswith("sampler/lcg48/")])
def test_device_sampler_stream(device, key):
    want = KAT[key].copy()
    if key.endswith("seed99"):
        seed, x, y, s = 99, 10, 20, 30
    else:
        seed = 1234
        x, y, s = (int(v) for v in key.split("/")[-1].split("_"))
    got = device.kat(ky.KAT_SAMPLER, np.array([[x, y, s, 0]], np.float32), index=seed)[0]
    # the fixture's first two floats went through get_camera_sample: (pixel + u) - pixel (ky.cpp:971-974)
    got[0] = (np.float32(x) + got[0]) - np.float32(x)
    got[1] = (np.float32(y) + got[1]) - np.float32(y)
    assert mismatch(got, want) == 0


@pytest.mark.parametrize("only_light", [False, True], ids=["second_light", "single_light"])
@pytest.mark.parametrize("ds", ["bsdf", "light", "both_mis"])
@pytest.mark.parametrize("flags", [0, ky.FLAG_FUSED], ids=["wavefront", "pixel"])
def test_film_with_every_shading_point_inside_a_sphere_light(device, only_light, ds, flags):
    scene = cases.inside_sphere_light_scene(only_light)
    device.upload(scene)
    desc = ky.render_desc(cases.W, cases.H, cases.SPP, direct_sample=cases.STRATEGIES[ds], flags=ky.FLAG_CLAMP | flags)
    got = device.render(desc)
    want, rays = kyo.render(scene, desc)
    assert mismatch(got, want) == 0
    assert device.stats().rays == rays
    if ds != "bsdf":
        assert float(want.mean()) > 0.02   # the shell lights itself and the box (BSDF-sampled rays only see its dark inside)
