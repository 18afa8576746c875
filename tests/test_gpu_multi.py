"""GPU: the multi-GPU context behind the drop-in (kyd_create_multi) and the context-sharing rules of one device.

A multi context whose device list repeats ordinal 0 exercises the whole multi path -- sample split, one host thread per
rank, partial films, rank-order sum kernel, clamp after the sum -- on a single-GPU box; with two GPUs present the same
cases run across real peers."""
import os
import subprocess
import threading

import numpy as np
import pytest

import cases
import ky_b200 as ky
import kyo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H = 64, 36


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _shares(begin, end, n):
    total = end - begin
    base, extra = divmod(total, n)
    out = []
    for r in range(n):
        b = begin + r * base + min(r, extra)
        out.append((b, b + base + (1 if r < extra else 0)))
    return out


def _device_lists():
    import torch
    lists = [[0, 0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        lists += [[0, 1], [1, 0, 1]]
    return lists


@pytest.mark.parametrize("scene_id,flags,spp", [(ky.SCENE_CORNELL, 0, 7), (ky.SCENE_VEACH, 0, 5), (ky.SCENE_CORNELL, ky.FLAG_FUSED, 2)])
def test_multi_context_equals_the_rank_order_sum_of_its_shares(device, scene_id, flags, spp):
    scene = ky.Scene(scene_id, W, H)
    device.upload(scene)
    whole_desc = ky.render_desc(W, H, spp, flags=ky.FLAG_CLAMP | flags)
    single = device.render(whole_desc)
    want, rays = kyo.render(scene, ky.render_desc(W, H, spp))
    assert np.array_equal(_bits(single), _bits(want))
    for devices in _device_lists():
        n = len(devices)
        # what the ranks render, summed on the host in rank order with float32 additions, clamped after the sum
        acc = None
        for b, e in _shares(0, spp, n):
            if e == b and acc is not None:
                continue
            part = device.render(ky.render_desc(W, H, spp, sample_begin=b, sample_end=e, flags=flags))
            acc = part if acc is None else (acc + part).astype(np.float32)
        expect = np.clip(acc, 0.0, 1.0).astype(np.float32)
        multi = ky.Device(devices)
        assert multi.device_count == n
        multi.upload(scene)
        got = multi.render(whole_desc)
        st = multi.stats()
        multi.close()
        assert np.array_equal(_bits(got), _bits(expect)), f"devices {devices}"
        assert st.rays == rays and st.samples == W * H * spp
        assert np.all(np.abs(got - want) <= 2e-6 * np.maximum(1.0, np.abs(want)))


def test_multi_context_device_film_and_accumulate(device):
    import torch
    spp = 8
    scene = ky.Scene(ky.SCENE_CORNELL, W, H)
    multi = ky.Device([0, 0])
    multi.upload(scene)
    film = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda:0")
    for b, e in ((0, 3), (3, 8)):
        multi.render_device(ky.render_desc(W, H, spp, sample_begin=b, sample_end=e, flags=ky.FLAG_ACCUMULATE), film.data_ptr())
    multi.clamp_device(film.data_ptr(), film.numel())
    torch.cuda.synchronize()
    multi.close()
    want, _ = kyo.render(scene, ky.render_desc(W, H, spp))
    got = film.cpu().numpy()
    assert np.all(np.abs(got - want) <= 2e-6 * np.maximum(1.0, np.abs(want)))


def test_two_contexts_on_one_device_from_two_threads():
    """The constant-memory scene is one symbol per device: renders of different contexts take turns (kyd_api.cu, SceneSlot)."""
    jobs = [(ky.SCENE_CORNELL, 48, 32, 3), (ky.SCENE_VEACH, 40, 24, 2)]
    wants = []
    for sid, w, h, spp in jobs:
        sc = ky.Scene(sid, w, h)
        wants.append(kyo.render(sc, ky.render_desc(w, h, spp))[0])
    errors = []

    def worker(k):
        try:
            sid, w, h, spp = jobs[k]
            dev = ky.Device(0)
            dev.set_wave_paths(1024)   # many launches per render: plenty of room to interleave if nothing prevented it
            sc = ky.Scene(sid, w, h)
            dev.upload(sc)
            for _ in range(12):
                got = dev.render(ky.render_desc(w, h, spp))
                if not np.array_equal(_bits(got), _bits(wants[k])):
                    errors.append(f"context {k}: film differs")
                    break
            dev.close()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_render_is_ordered_after_work_on_the_default_stream(device):
    """film.zero_() / film.fill_() on torch's default stream followed by an accumulating render on the context's own stream."""
    import torch
    w, h, spp = 512, 512, 1
    scene = ky.Scene(ky.SCENE_CORNELL, w, h)
    device.upload(scene)
    want = device.render(ky.render_desc(w, h, spp, flags=0))
    big = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
    film = torch.empty((h, w, 3), dtype=torch.float32, device="cuda")
    for _ in range(3):
        film.fill_(7.0)
        big.zero_()          # keeps the default stream busy so that an unordered render would start before zero_()
        film.zero_()
        device.render_device(ky.render_desc(w, h, spp, flags=ky.FLAG_ACCUMULATE), film.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(_bits(film.cpu().numpy()), _bits(want))


def test_failed_upload_keeps_the_previous_scene(device):
    scene = cases.make_scene("cornell")
    device.upload(scene)
    desc = ky.render_desc(cases.W, cases.H, 2)
    before = device.render(desc)
    bad = cases.make_scene("veach")
    keep = bad.desc.lights[bad.desc.light_count - 1].kind
    bad.desc.lights[bad.desc.light_count - 1].kind = 99     # found late: after surfaces, materials and the other lights
    with pytest.raises(RuntimeError, match="unknown light kind"):
        device.upload(bad)
    bad.desc.lights[bad.desc.light_count - 1].kind = keep
    device._scene = scene
    after = device.render(desc)
    assert np.array_equal(_bits(before), _bits(after))
    with pytest.raises(RuntimeError, match="sample_end exceeds spp"):
        device.render(ky.render_desc(8, 8, 4, sample_begin=2, sample_end=5))
    # statistics asked for before any render: no stale CUDA error may leak into the next call
    fresh = ky.Device(0)
    fresh.stats()
    fresh.upload(scene)
    assert np.array_equal(_bits(fresh.render(desc)), _bits(before))
    fresh.close()


def test_cli_entry_point_uses_every_listed_device(tmp_path):
    """`ky render_single_scene` (the reference's main(), ky.cpp:4937-4949) through the C++ surface: KY_CUDA_DEVICES makes
    integrator_t::render a multi-GPU call; the image equals the single-device one up to the last rounding of the sum."""
    import torch
    exe = os.path.join(ROOT, "ky_b200", "lib", "ky")
    outs = []
    devices = "0,1" if torch.cuda.device_count() >= 2 else "0,0"
    for env_devices in (None, devices):
        d = tmp_path / ("multi" if env_devices else "single")
        d.mkdir()
        env = dict(os.environ)
        env.pop("KY_CUDA_DEVICES", None)
        if env_devices:
            env["KY_CUDA_DEVICES"] = env_devices
        subprocess.run([exe, "render_single_scene", "8", "96", "96"], cwd=d, env=env, check=True, capture_output=True, timeout=300)
        outs.append(np.frombuffer((d / "render_single_scene.bmp").read_bytes(), np.uint8).astype(np.int32))
    assert outs[0].shape == outs[1].shape
    diff = np.abs(outs[0] - outs[1])
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
