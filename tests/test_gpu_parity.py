"""GPU parity tests: the sm_100a path (through the C ABI, host buffers) against the C oracle and the
golden films.  Bar: bit-exact films (byte identical float32) -- the only admitted deviation is the
documented ~1e-8-per-call double-libm rounding difference, none of which occurs in these cases."""
import numpy as np
import pytest

import cases
import ky_b200 as ky
import kyo

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _assert_same(got, want, what):
    diff = (_bits(got) != _bits(want)).any(axis=-1)
    assert not diff.any(), f"{what}: {int(diff.sum())} of {diff.size} pixels differ, max abs {np.nanmax(np.abs(got - want))}"


@pytest.mark.parametrize("case", cases.film_cases(), ids=lambda c: c[0])
def test_film_matches_oracle(device, case):
    name, sk, integ, ds, depth, spp = case
    scene = cases.make_scene(sk)
    desc = ky.render_desc(cases.W, cases.H, spp, integrator=integ, max_depth=depth, direct_sample=ds)
    device.upload(scene)
    got = device.render(desc)
    want, rays = kyo.render(scene, desc)
    _assert_same(got, want, name)
    st = device.stats()
    assert st.samples == cases.W * cases.H * max(1, spp)
    assert st.rays == rays, f"{name}: reference-equivalent ray count {st.rays} != oracle {rays}"
    assert st.rays_traced <= st.rays
