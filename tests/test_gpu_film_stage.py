"""GPU: film output stage through the C ABI (kyd_film_encode / kyd_film_encode_device) against the C oracle, the
reference-written golden files, and - for gamma_encoding - every float of [0, 1]."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cases
import ky_b200 as ky
import kyo
from test_film_stage_cpu import FORMATS, GOLD, split_file, table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    d = ky.Device(0)
    yield d
    d.close()


@pytest.mark.parametrize("film", cases.STAGE_FILMS, ids=lambda f: f[0])
@pytest.mark.parametrize("fmt,ext", FORMATS)
def test_device_matches_oracle_and_reference_files(dev, film, fmt, ext):
    name, w, h, seed, scale = film
    f = cases.stage_film(w, h, seed, scale)
    got = dev.film_encode(f, fmt)
    assert np.array_equal(got, kyo.film_encode(fmt, f))
    _, body = split_file(fmt, GOLD[f"{name}.{ext}"], w, h)
    assert np.array_equal(got, body)


@pytest.mark.parametrize("fmt,ext", FORMATS)
def test_device_pointers_full_size(dev, fmt, ext):
    """C5's film size, device-resident in and out; checked against the oracle in full."""
    import torch
    w, h = 3840, 2160
    f = cases.stage_film(w, h, 9, fmt == ky.FILM_RGBE)
    film = torch.from_numpy(f).cuda()
    body = torch.zeros(ky.kyd().kyd_film_body_bytes(fmt, w, h), dtype=torch.uint8, device="cuda")
    dev.film_encode_device(film.data_ptr(), w, h, fmt, body.data_ptr())
    torch.cuda.synchronize()
    got = body.cpu().numpy()
    assert np.array_equal(got, kyo.film_encode(fmt, f))
    if fmt == ky.FILM_BMP24:  # size-independent property: the bmp body is the gamma8 body, lines reversed, BGR
        g8 = dev.film_encode(f, ky.FILM_GAMMA8).reshape(h, w, 3)
        assert np.array_equal(got.reshape(h, w, 3), g8[::-1, :, ::-1])


def test_gamma_encoding_every_float_of_unit_interval(dev):
    """All 2^30 - 2^23 + 1 floats of [0, 1] (plus their negatives' clamp): the device bytes equal the counting rule of
    the generated table, which scripts/make_gamma_table.py checked against the reference for the same set."""
    import torch
    bits = torch.from_numpy(table().astype(np.int64)).cuda()
    chunk = 3 * (1 << 22)
    body = torch.zeros(chunk, dtype=torch.uint8, device="cuda")
    bad = 0
    first = 0
    while first <= 0x3F800000:
        n = min(chunk, 0x3F800000 + 1 - first)
        n3 = (n + 2) // 3 * 3
        x = torch.arange(first, first + n3, dtype=torch.int64, device="cuda").clamp_(max=0x3F800000)
        film = x.to(torch.int32).view(torch.float32).contiguous()
        torch.cuda.synchronize()  # the context's own stream does not wait for torch's
        dev.film_encode_device(film.data_ptr(), n3 // 3, 1, ky.FILM_GAMMA8, body.data_ptr())
        torch.cuda.synchronize()
        want = torch.searchsorted(bits, x, right=True) - 1
        bad += int((body[:n3].to(torch.int64) != want).sum())
        first += n
    assert bad == 0


def test_film_writers_through_the_host_classes(dev, tmp_path):
    """film_t::store_image / store_device (device stage) write the same files as the host-side writers and the golden."""
    name, w, h, seed, scale = cases.STAGE_FILMS[1]
    f = cases.stage_film(w, h, seed, scale)
    host = ky.host()
    host.ky_host_film_store.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    for kind, ext in ((1, "ppm"), (2, "bmp"), (3, "hdr")):
        p_dev, p_host = tmp_path / f"dev.{ext}", tmp_path / f"host.{ext}"
        assert host.ky_host_film_store(kind, str(p_dev).encode(), w, h, f.ctypes.data_as(C.c_void_p)) == 0
        assert host.ky_host_film_store(kind + 10, str(p_host).encode(), w, h, f.ctypes.data_as(C.c_void_p)) == 0
        assert p_dev.read_bytes() == p_host.read_bytes() == bytes(GOLD[f"{name}.{ext}"])
    stem = tmp_path / "image"
    assert host.ky_host_film_store(0, str(stem).encode(), w, h, f.ctypes.data_as(C.c_void_p)) == 0
    assert (tmp_path / "image.bmp").read_bytes() == bytes(GOLD[f"{name}.bmp"])


def test_bad_arguments(dev):
    with pytest.raises(ValueError):
        dev.film_encode(np.zeros((2, 2, 3), np.float32), 7)
    import torch
    film = torch.zeros(64, device="cuda")
    body = torch.zeros(64, dtype=torch.uint8, device="cuda")
    with pytest.raises(RuntimeError):
        dev.film_encode_device(film.data_ptr() + 4, 2, 2, ky.FILM_GAMMA8, body.data_ptr())
    with pytest.raises(RuntimeError):
        dev.film_encode_device(film.data_ptr(), 2, 2, ky.FILM_GAMMA8, body.data_ptr() + 1)
    with pytest.raises(RuntimeError):
        dev.film_encode_device(0, 2, 2, ky.FILM_GAMMA8, body.data_ptr())
