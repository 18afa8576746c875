"""FP64 smallpt validation mode: the C restatement oracle/smallpt_f64.c against the reference file
itself (oracle/_ref/libsmallpt_kernel_ref.so, live and as golden fixture) on the CPU; the device kernel against the oracle
on the GPU."""
import os

import numpy as np
import pytest

import cases
import ky_b200 as ky
import kyo
import kyref

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "golden_smallpt_f64.npz"))


@pytest.mark.parametrize("case", cases.SMALLPT_F64_CASES, ids=lambda c: f"{c[0]}x{c[1]}@{c[2]}")
def test_oracle_matches_reference_golden(case):
    w, h, spp = case
    got = kyo.smallpt_f64(w, h, spp)
    assert np.array_equal(got.view(np.uint64), GOLD[f"{w}x{h}@{spp}"].view(np.uint64))


@pytest.mark.skipif(not kyref.smallpt_available(), reason="oracle/_ref not built")
def test_oracle_matches_reference_live():
    w, h, spp = 80, 60, 12
    assert np.array_equal(kyo.smallpt_f64(w, h, spp).view(np.uint64), kyref.smallpt_f64(w, h, spp).view(np.uint64))


@pytest.mark.gpu
@pytest.mark.parametrize("case", cases.SMALLPT_F64_CASES + [(256, 192, 32)], ids=lambda c: f"{c[0]}x{c[1]}@{c[2]}")
def test_device_matches_oracle(device, case):
    """Same arithmetic in IEEE double on both sides except sin / cos (CUDA's vs glibc's double functions, both within
    an ulp or two): per-pixel agreement to 1e-9, apart from pixels where that last-bit difference flips a branch of a path
    (allowed: at most 1 in 2000 pixels)."""
    w, h, spp = case
    got = device.render_smallpt_f64(w, h, spp)
    want = kyo.smallpt_f64(w, h, spp)
    err = np.abs(got - want).max(axis=-1)
    assert (err > 1e-9).sum() <= max(1, w * h // 2000), (int((err > 1e-9).sum()), float(err.max()))
    assert np.median(err) < 1e-13
    st = device.stats()
    assert st.samples == w * h * spp


@pytest.mark.gpu
def test_device_rejects_bad_arguments(device):
    with pytest.raises(RuntimeError):
        device.render_smallpt_f64(0, 4, 1)
    with pytest.raises(RuntimeError):
        device.render_smallpt_f64(4, 4, 0)
