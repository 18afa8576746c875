"""CPU: film output stage (gamma / BMP / RGBE, ky.cpp:1548, 1661-1782).  The C oracle against the files the
REFERENCE's own writers produced (tests/golden/golden_film_stage.npz, and oracle/_ref live where it exists), the
generated gamma threshold table against the oracle, and the pure-host header function of libkyd."""
import os
import re

import numpy as np
import pytest

import cases
import ky_b200 as ky
import kyo
import kyref

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "golden_film_stage.npz"))
FORMATS = [(ky.FILM_GAMMA8, "ppm"), (ky.FILM_BMP24, "bmp"), (ky.FILM_RGBE, "hdr")]


def split_file(fmt, data, width, height):
    """(header bytes, body as uint8 array) of a file written by the reference; ppm numbers are parsed back to bytes"""
    data = bytes(data)
    if fmt == ky.FILM_GAMMA8:
        m = re.match(rb"P3\n\d+ \d+\n255\n", data)
        body = np.array([int(t) for t in data[m.end():].split()], np.uint8)
        return data[:m.end()], body
    n = width * height * (4 if fmt == ky.FILM_RGBE else 3)
    return data[:len(data) - n], np.frombuffer(data[len(data) - n:], np.uint8)


def table():
    src = open(os.path.join(HERE, "..", "ky_b200", "csrc", "kyd_gamma_table.h")).read()
    bits = np.array([int(t, 16) for t in re.findall(r"0x([0-9a-f]{8})u", src)], np.uint32)
    assert bits.size == 256
    return bits


@pytest.mark.parametrize("film", cases.STAGE_FILMS, ids=lambda f: f[0])
@pytest.mark.parametrize("fmt,ext", FORMATS)
def test_oracle_matches_reference_files(film, fmt, ext):
    name, w, h, seed, scale = film
    f = cases.stage_film(w, h, seed, scale)
    header, body = split_file(fmt, GOLD[f"{name}.{ext}"], w, h)
    assert np.array_equal(kyo.film_encode(fmt, f), body)
    assert ky.film_header(fmt, w, h) == header


@pytest.mark.skipif(not kyref.available("verbatim"), reason="oracle/_ref not built")
@pytest.mark.parametrize("fmt,ext", FORMATS)
def test_oracle_matches_reference_live(fmt, ext):
    w, h = 61, 33
    f = cases.stage_film(w, h, 77, True)
    header, body = split_file(fmt, kyref.store_film(fmt, w, h, f), w, h)
    assert np.array_equal(kyo.film_encode(fmt, f), body)
    assert ky.film_header(fmt, w, h) == header
    x = np.random.default_rng(5).random(200000).astype(np.float32) ** 3
    assert np.array_equal(kyref.gamma_encoding(x), kyo.gamma_encoding(x))


def test_gamma_table_is_the_oracles_step_function():
    bits = table()
    assert bits[0] == 0 and bits[255] <= 0x3F800000 and (np.diff(bits.astype(np.int64)) > 0).all()
    at = kyo.gamma_encoding(bits.view(np.float32))
    below = kyo.gamma_encoding((bits[1:] - 1).view(np.float32))
    assert np.array_equal(at, np.arange(256, dtype=np.uint8))
    assert np.array_equal(below, np.arange(255, dtype=np.uint8))
    # and in between: 2 M random floats of [0, 1] + the edge values through the table's counting rule
    rng = np.random.default_rng(11)
    x = np.concatenate([rng.integers(0, 0x3F800001, 2_000_000, dtype=np.uint32).view(np.float32),
                        np.array(cases.FILM_EDGE_VALUES, np.float32)])
    xc = np.where(x < 0, np.float32(0), np.where(x > 1, np.float32(1), x))
    count = np.searchsorted(bits.view(np.float32), xc, side="right") - 1
    count = np.where(np.isnan(x), 0, count)
    assert np.array_equal(count.astype(np.uint8), kyo.gamma_encoding(x))


def test_body_sizes_and_bad_arguments():
    l = ky.kyd()
    assert l.kyd_film_body_bytes(ky.FILM_GAMMA8, 7, 5) == 105
    assert l.kyd_film_body_bytes(ky.FILM_BMP24, 7, 5) == 105
    assert l.kyd_film_body_bytes(ky.FILM_RGBE, 7, 5) == 140
    assert l.kyd_film_body_bytes(3, 7, 5) == -1
    assert l.kyd_film_body_bytes(ky.FILM_RGBE, 0, 5) == -1
    with pytest.raises(ValueError):
        ky.film_header(9, 4, 4)
