"""GPU: the BASELINE configurations at their real film sizes (reduced sample counts; sample-count independence is the
range-accumulation property).  Every kernel organisation must give the same bits, the oracle must agree with them, and
sample ranges accumulated on the device must equal the one-shot film."""
import numpy as np
import pytest

import ky_b200 as ky
import kyo

pytestmark = pytest.mark.gpu

# (name, scene, scene flags, width, height, integrator, depth, direct sample, spp)
CONFIGS = [
    ("C1 smallpt 1024x768", ky.SCENE_SMALLPT, 0, 1024, 768, ky.INT_PT_ITERATION, 5, ky.DS_BOTH_MIS, 4),
    ("C2 cornell direct 1024x768", ky.SCENE_CORNELL, ky.CB_DEFAULT, 1024, 768, ky.INT_DIRECT_LIGHTING, 0, ky.DS_BOTH_MIS, 4),
    ("C3 veach 1280x720", ky.SCENE_VEACH, 0, 1280, 720, ky.INT_PT_ITERATION, 5, ky.DS_BOTH_MIS, 4),
    ("C4 cornell env panel 480x360 d8", ky.SCENE_CORNELL, ky.CB_BOTH_SMALL | ky.CB_LIGHT_ENVIRONMENT, 480, 360, ky.INT_PT_ITERATION, 8, ky.DS_BOTH_MIS, 8),
    ("C5 cornell 3840x2160", ky.SCENE_CORNELL, ky.CB_DEFAULT, 3840, 2160, ky.INT_PT_ITERATION, 5, ky.DS_BOTH_MIS, 2),
]


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: c[0])
def test_baseline_config_at_full_resolution(device, cfg):
    import torch
    name, sid, sflags, w, h, integ, depth, ds, spp = cfg
    scene = ky.Scene(sid, w, h, sflags)
    device.upload(scene)

    def desc(flags=ky.FLAG_CLAMP, **kw):
        return ky.render_desc(w, h, spp, integrator=integ, max_depth=depth, direct_sample=ds, flags=flags, **kw)

    device.set_wave_paths(0)
    one_shot = device.render(desc())
    rays = device.stats().rays
    samples = device.stats().samples
    assert samples == w * h * spp

    # the oracle (all host cores), bit for bit, and the reference's ray count
    want, want_rays = kyo.render(scene, desc())
    differing = int((_bits(one_shot) != _bits(want)).any(axis=-1).sum())
    assert differing == 0, f"{name}: {differing} of {w * h} pixels differ from the oracle"
    assert rays == want_rays

    # organisation independence: many small waves / tiles, and the per-pixel kernel
    device.set_wave_paths(1 << 19)
    small_waves = device.render(desc())
    assert np.array_equal(_bits(small_waves), _bits(one_shot)) and device.stats().rays == rays
    device.set_wave_paths(0)
    pixel = device.render(desc(flags=ky.FLAG_CLAMP | ky.FLAG_FUSED))
    assert np.array_equal(_bits(pixel), _bits(one_shot)) and device.stats().rays == rays

    # sample ranges accumulated on the device (what a multi-step or multi-GPU job does), then clamped
    film = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda")
    cut = spp // 2
    for b, e in ((0, cut), (cut, spp)):
        device.render_device(desc(flags=ky.FLAG_ACCUMULATE, sample_begin=b, sample_end=e), film.data_ptr())
    device.clamp_device(film.data_ptr(), film.numel())
    torch.cuda.synchronize()
    assert np.array_equal(_bits(film.cpu().numpy()), _bits(one_shot))


def test_nan_sample_keeps_the_reference_bit_pattern(device):
    """Sample 13 of pixel (1816, 1777) of the C5 job has a NaN radiance in the reference (0 * inf in a throughput).  The film
    must carry it with the bit pattern the reference's x86-64 build produces (SSE default NaN, 0xffc00000), in both kernel
    organisations, clamped or not."""
    w, h, spp = 3840, 2160, 16
    scene = ky.Scene(ky.SCENE_CORNELL, w, h, ky.CB_DEFAULT)
    device.upload(scene)
    device.set_wave_paths(0)
    for flags in (0, ky.FLAG_CLAMP, ky.FLAG_FUSED):
        desc = ky.render_desc(w, h, spp, max_depth=5, sample_begin=13, sample_end=14, flags=flags)
        got = device.render(desc)
        want, rays = kyo.render(scene, desc)
        assert np.isnan(want[1777, 1816]).all() and int(np.isnan(want).any(axis=-1).sum()) == 1
        assert (_bits(want[1777, 1816]) == 0xFFC00000).all()
        assert np.array_equal(_bits(got), _bits(want))
        assert device.stats().rays == rays
