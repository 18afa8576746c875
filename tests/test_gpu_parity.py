"""GPU parity tests: the sm_100a path (through the C ABI, host buffers) against the C oracle and the
golden films.  Bar: bit-exact films (byte identical float32) -- the only admitted deviation is the
documented ~1e-8-per-call double-libm rounding difference, none of which occurs in these cases."""
import os

import numpy as np
import pytest

import cases
import ky_b200 as ky
import kyo

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
FILMS = np.load(os.path.join(HERE, "golden", "golden_films.npz"))


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _assert_same(got, want, what):
    diff = (_bits(got) != _bits(want)).any(axis=-1)
    assert not diff.any(), f"{what}: {int(diff.sum())} of {diff.size} pixels differ, max abs {np.nanmax(np.abs(got - want))}"


# kernel organisations: the split wavefront (default), the same with waves far smaller than the film
# (several tiles, or several samples per wave), and the per-pixel kernel
MODES = {"wavefront": (0, 0), "wavefront_small_waves": (0, 1000), "wavefront_multi_spp": (0, cases.W * cases.H * 3),
         "wavefront_split_light_sample": (ky.FLAG_SPLIT_LIGHT_SAMPLE, 0), "pixel": (ky.FLAG_FUSED, 0)}


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("case", cases.film_cases(), ids=lambda c: c[0])
def test_film_matches_oracle_and_reference_golden(device, case, mode):
    name, sk, integ, ds, depth, spp = case
    flags, wave = MODES[mode]
    if mode != "wavefront" and integ not in (ky.INT_PT_ITERATION, ky.INT_DIRECT_LIGHTING):
        pytest.skip("only one kernel organisation exists for this integrator")
    scene = cases.make_scene(sk)
    desc = ky.render_desc(cases.W, cases.H, spp, integrator=integ, max_depth=depth, direct_sample=ds, flags=ky.FLAG_CLAMP | flags)
    device.set_wave_paths(wave)
    device.upload(scene)
    got = device.render(desc)
    device.set_wave_paths(0)
    want, rays = kyo.render(scene, desc)
    _assert_same(got, want, name + " vs oracle")
    _assert_same(got, FILMS[name], name + " vs reference golden")
    st = device.stats()
    assert st.samples == cases.W * cases.H * max(1, spp)
    assert st.rays == rays == int(FILMS[name + "#rays"][0]), f"{name}: reference-equivalent ray count {st.rays} != oracle {rays}"
    assert st.rays_traced <= st.rays
    assert st.kernel_launches >= 1


@pytest.mark.parametrize("flags", [0, ky.FLAG_FUSED], ids=["wavefront", "pixel"])
def test_sample_ranges_and_device_film_accumulation(device, flags):
    """A job cut into sample ranges (what a multi-GPU split hands out) accumulated into one device-side film
    equals the one-shot render bit for bit when the ranges are rendered in order."""
    import torch
    w, h, spp = 64, 40, 6
    scene = ky.Scene(ky.SCENE_CORNELL, w, h)
    device.upload(scene)
    whole = device.render(ky.render_desc(w, h, spp, flags=ky.FLAG_CLAMP | flags))
    film = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    for b, e in [(0, 1), (1, 4), (4, 6)]:
        d = ky.render_desc(w, h, spp, sample_begin=b, sample_end=e, flags=ky.FLAG_ACCUMULATE | flags)
        device.render_device(d, film.data_ptr())
    device.clamp_device(film.data_ptr(), film.numel())
    torch.cuda.synchronize()
    _assert_same(film.cpu().numpy(), whole, "range-split film")
    want, _ = kyo.render(scene, ky.render_desc(w, h, spp))
    _assert_same(whole, want, "one-shot film")


def test_fast_rsqrt_is_exact_for_every_float(device):
    """vec3 normalize's (float)(1.0 / sqrt((double)s)) (ky.cpp:314): the FP32 fast path equals the FP64 definition
    for all 2^32 float bit patterns; only ~6e-5 of the in-range inputs need the FP64 slow path."""
    bad, slow = device.selftest(0, 0, 1 << 32)
    assert bad == 0
    in_range = (int(np.float32(2.0 ** 60).view(np.uint32)) - int(np.float32(2.0 ** -60).view(np.uint32)))
    outside = (1 << 32) - in_range
    assert slow - outside < 2e-4 * in_range, (slow, outside)


def test_fast_pow_matches_its_definition(device):
    """powf contract value (float)pow((double)x, (double)y): the exp2/log2 fast path with its rounding-interval check
    agrees with the definition on 2^31 argument pairs (Phong exponents, random exponents, bases dense near 1)."""
    bad, slow = device.selftest(1, 12345, 1 << 31)
    assert bad == 0
    assert slow < 0.2 * (1 << 31)


@pytest.mark.parametrize("order", ["first", "last"])
@pytest.mark.parametrize("integrator,flags", [(ky.INT_BASECOLOR, 0), (ky.INT_PT_ITERATION, 0), (ky.INT_PT_ITERATION, ky.FLAG_FUSED),
                                              (ky.INT_DIRECT_LIGHTING, 0)])
def test_equal_hit_distances_resolve_in_list_order(device, order, integrator, flags):
    """A disk and a triangle exactly in the plane of a Cornell wall: three shape kinds report bit-identical distances.  The
    reference keeps the first surface in list order; the device walks surfaces grouped by kind and must still do so."""
    scene = cases.coplanar_tie_scene(order)
    desc = ky.render_desc(cases.W, cases.H, 4, integrator=integrator, max_depth=3, flags=flags)
    device.upload(scene)
    got = device.render(desc)
    want, rays = kyo.render(scene, desc)
    _assert_same(got, want, f"tie scene {order}")
    assert device.stats().rays == rays
    if integrator == ky.INT_BASECOLOR:
        # the tie really happens and really matters: the two orders give different images
        other = kyo.render(cases.coplanar_tie_scene("last" if order == "first" else "first"), desc)[0]
        assert (_bits(other) != _bits(want)).any()


@pytest.mark.parametrize("order", ["first", "last"])
@pytest.mark.parametrize("ds,flags", [(ky.DS_BOTH_MIS, 0), (ky.DS_BSDF_MIS, 0), (ky.DS_BSDF, 0), (ky.DS_BOTH_MIS, ky.FLAG_FUSED)])
def test_light_surface_tied_with_another_surface(device, order, ds, flags):
    """A matte copy of a sphere light's surface: the BSDF-sampled query must see the light only if the light's surface
    comes first in list order (the occlusion form of the query has its own tie rule)."""
    scene = cases.light_tie_scene(order)
    desc = ky.render_desc(cases.W, cases.H, 4, integrator=ky.INT_PT_ITERATION, max_depth=3, direct_sample=ds, flags=flags)
    device.upload(scene)
    got = device.render(desc)
    want, rays = kyo.render(scene, desc)
    _assert_same(got, want, f"light tie scene {order}")
    assert device.stats().rays == rays


TRAP = np.load(os.path.join(HERE, "golden", "golden_trapezoidal.npz"))


@pytest.mark.parametrize("flags", [0, ky.FLAG_FUSED], ids=["wavefront", "pixel"])
@pytest.mark.parametrize("case", cases.trapezoidal_cases(), ids=lambda c: c[0])
def test_trapezoidal_sampler(device, case, flags):
    """Tent-filtered 2x2 sub-pixel camera samples (KYD_SAMPLER_TRAPEZOIDAL) against the oracle and the reference golden,
    one-shot and as two sample ranges accumulated on the device."""
    import torch
    name, sk, integ, ds, depth, spp = case
    if flags == 0 and integ not in (ky.INT_PT_ITERATION, ky.INT_DIRECT_LIGHTING):
        pytest.skip("only one kernel organisation exists for this integrator")
    scene = cases.make_scene(sk)
    device.upload(scene)
    desc = ky.render_desc(cases.W, cases.H, spp, integrator=integ, max_depth=depth, direct_sample=ds, sampler=ky.SAMPLER_TRAPEZOIDAL,
                          flags=flags | ky.FLAG_CLAMP)
    got = device.render(desc)
    want, rays = kyo.render(scene, desc)
    _assert_same(got, want, name)
    _assert_same(got, TRAP[name], name + " (golden)")
    assert device.stats().rays == rays == int(TRAP[name + "#rays"][0])
    film = torch.zeros((cases.H, cases.W, 3), dtype=torch.float32, device="cuda")
    for b, e in ((0, spp // 2 + 1), (spp // 2 + 1, spp)):
        part = ky.render_desc(cases.W, cases.H, spp, integrator=integ, max_depth=depth, direct_sample=ds, sampler=ky.SAMPLER_TRAPEZOIDAL,
                              sample_begin=b, sample_end=e, flags=flags | ky.FLAG_ACCUMULATE)
        device.render_device(part, film.data_ptr())
    device.clamp_device(film.data_ptr(), film.numel())
    torch.cuda.synchronize()
    _assert_same(film.cpu().numpy(), want, name + " (ranges)")


def test_debug_sampler(device):
    scene = cases.make_scene("cornell")
    desc = ky.render_desc(cases.W, cases.H, 2, sampler=ky.SAMPLER_DEBUG)
    device.upload(scene)
    for flags in (0, ky.FLAG_FUSED):
        desc.flags = ky.FLAG_CLAMP | flags
        got = device.render(desc)
        want, _ = kyo.render(scene, desc)
        _assert_same(got, want, "debug_sampler_t")


def test_errors_are_reported_not_swallowed(device):
    scene = cases.make_scene("cornell")
    device.upload(scene)
    bad = ky.render_desc(8, 8, 1, integrator=5)
    with pytest.raises(RuntimeError, match="unsupported integrator"):
        device.render(bad)
    with pytest.raises(RuntimeError, match="spp must be positive"):
        device.render(ky.render_desc(8, 8, 0, sample_end=1))
    fresh = ky.Device(0)
    with pytest.raises(RuntimeError, match="before kyd_upload_scene"):
        fresh.render(ky.render_desc(8, 8, 1))
    fresh.close()


_SWITCH_CHILD = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.environ["KY_TESTS"]); sys.path.insert(0, os.path.dirname(os.environ["KY_TESTS"]))
import cases, ky_b200 as ky
films = np.load(os.path.join(os.environ["KY_TESTS"], "golden", "golden_films.npz"))
dev = ky.Device(0)
bad = []
for name, sk, integ, ds, depth, spp in cases.film_cases():
    if integ != ky.INT_PT_ITERATION or ds != ky.DS_BOTH_MIS:
        continue
    for wave in (0, 1000):
        dev.set_wave_paths(wave)
        dev.upload(cases.make_scene(sk))
        got = dev.render(ky.render_desc(cases.W, cases.H, spp, integrator=integ, max_depth=depth, direct_sample=ds, flags=ky.FLAG_CLAMP))
        if (np.ascontiguousarray(got).view(np.uint32) != films[name].view(np.uint32)).any() or dev.stats().rays != int(films[name + "#rays"][0]):
            bad.append((name, wave))
print("BAD", bad)
sys.exit(1 if bad else 0)
"""


@pytest.mark.parametrize("env", [{"KYD_FUSE_INTERSECT": "0"}, {"KYD_FUSE_INTERSECT_MANY": "0"},
                                 {"KYD_GRID_SHADE": "7", "KYD_GRID_NEE": "5", "KYD_GRID256": "3"},
                                 {"KYD_GRID_SHADE": "1", "KYD_GRID_NEE": "1", "KYD_GRID256": "1"}],
                         ids=["unfused", "unfused_many", "odd_grids", "one_block_per_sm"])
def test_run_time_switches_do_not_change_films(env):
    """The library's run-time knobs (intersect stage fused into shade or not -- for one light and for several --, grid sizes of
    the persistent kernels) are read once per process: a child process renders every headline-configuration film with them
    set, at two wave sizes, and compares with the reference-generated golden films bit for bit."""
    import subprocess, sys
    e = dict(os.environ, KY_TESTS=HERE, **env)
    r = subprocess.run([sys.executable, "-c", _SWITCH_CHILD], env=e, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
