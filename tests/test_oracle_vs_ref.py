"""CPU, build container only: the C oracle against the live REFERENCE build (oracle/_ref), at other
sizes / seeds / sample offsets than the committed fixtures, plus the statistical link between the
deterministic reference configuration and the verbatim one."""
import numpy as np
import pytest

import cases
import ky_b200 as ky
import kyo
import kyref

pytestmark = pytest.mark.skipif(not (kyref.available("det") and kyref.available("verbatim")),
                                reason="oracle/_ref not built (no /root/reference here)")

REF_SCENE = {ky.SCENE_CORNELL: kyref.CORNELL, ky.SCENE_VEACH: kyref.VEACH, ky.SCENE_SMALLPT: kyref.SMALLPT, ky.SCENE_SHAPES: kyref.SHAPES}
LUM = np.array([0.212671, 0.715160, 0.072169], np.float32)


@pytest.mark.parametrize("sk", list(cases.SCENES))
@pytest.mark.parametrize("seed,offset", [(1234, 0), (77, 5)])
def test_oracle_equals_reference(sk, seed, offset):
    w, h, spp = 80, 45, 3
    sid, flags = cases.SCENES[sk]
    scene = ky.Scene(sid, w, h, flags)
    desc = ky.render_desc(w, h, spp, seed=seed, sample_begin=offset, sample_end=offset + spp)
    got, rays = kyo.render(scene, desc)
    # the reference sampler counts samples from `offset`; weights are 1/spp in both
    want, _, ref_rays = kyref.render(REF_SCENE[sid], w, h, spp, scene_flags=flags, seed=seed, sample_offset=offset)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert rays == ref_rays


@pytest.mark.parametrize("sk", ["cornell", "veach", "shapes"])
def test_oracle_equals_reference_with_the_trapezoidal_sampler(sk):
    w, h, spp = 70, 40, 8
    sid, flags = cases.SCENES[sk]
    scene = ky.Scene(sid, w, h, flags)
    desc = ky.render_desc(w, h, spp, seed=99, sampler=ky.SAMPLER_TRAPEZOIDAL)
    got, rays = kyo.render(scene, desc)
    want, _, ref_rays = kyref.render(REF_SCENE[sid], w, h, spp, scene_flags=flags, seed=99, sampler=kyref.TRAPEZOIDAL_SAMPLER)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert rays == ref_rays
    # and it is a different image from the box-filtered one
    plain, _ = kyo.render(scene, ky.render_desc(w, h, spp, seed=99))
    assert not np.array_equal(plain.view(np.uint32), got.view(np.uint32))


def test_crlibm_shim_is_linked_into_the_deterministic_build():
    # with glibc's own sinf/cosf a few percent of Lambert samples differ in the last bit
    from golden.make_golden import MATERIAL_PARAMS
    import os
    kat = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_kat.npz"))
    kind, params = MATERIAL_PARAMS["matte"]
    det = kyref.material_bsdf(kind, params, kat["bsdf/in"], kind="det")
    verbatim = kyref.material_bsdf(kind, params, kat["bsdf/in"], kind="verbatim")
    differing = (det.view(np.uint32) != verbatim.view(np.uint32)).any(axis=1).mean()
    assert 0.001 < differing < 0.2
    assert np.nanmax(np.abs(det - verbatim)) < 1e-6


@pytest.mark.parametrize("scene,flags,w,h", [(kyref.CORNELL, kyref.DEFAULT_SCENE, 128, 96), (kyref.VEACH, 0, 160, 90)])
def test_deterministic_configuration_is_statistically_the_reference(scene, flags, w, h):
    """Different unbiased estimators of the same integral: mean luminance of the clamped films agrees
    within Monte-Carlo noise."""
    spp = 32
    det, _, rays_det = kyref.render(scene, w, h, spp, scene_flags=flags, kind="det")
    ver, _, rays_ver = kyref.render(scene, w, h, spp, scene_flags=flags, kind="verbatim", sampler=kyref.RANDOM_SAMPLER, threads=1)
    ld, lv = float((det @ LUM).mean()), float((ver @ LUM).mean())
    assert abs(ld - lv) / lv < 0.03, (ld, lv)
    assert abs(rays_det - rays_ver) / rays_ver < 0.02
