"""The drop-in proof: the reference's OWN entry points and main() (ky.cpp from `void render_single_scene(` to the end of the
file, extracted verbatim by oracle/ref/build_ref.sh) compile against include/ky.hpp, link with libkyd.so, and -- on the GPU
box -- write the same images as this repository's parameterised copies of those entry points (include/ky_entry.hpp)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "ky_ref_entries")
CLI = os.path.join(ROOT, "ky_b200", "lib", "ky")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(not os.path.exists("/root/reference/ky.cpp"), reason="needs /root/reference (build container)")
def test_reference_entry_points_compile_against_the_host_surface():
    inc = os.path.join(ROOT, "oracle", "_ref", "ref_entries.inc")
    assert os.path.exists(EXE) and os.path.exists(inc)
    text = open(inc).read()
    ref = open("/root/reference/ky.cpp").read()
    assert text in ref                                   # verbatim: a contiguous piece of the reference file
    for name in ("render_single_scene", "render_debug", "render_multiple_integrator", "render_direct_sample_enum",
                 "render_multiple_scene", "render_mis_scene", "int main("):
        assert name in text


@pytest.mark.skipif(_has_gpu() or not os.path.exists(EXE), reason="a GPU is present / binary not built")
def test_reference_entry_points_fail_loudly_without_a_gpu(tmp_path):
    out = subprocess.run([EXE, "render_debug"], cwd=tmp_path, capture_output=True, text=True)
    assert out.returncode != 0 and "no CUDA device" in out.stderr
    assert not list(tmp_path.iterdir())


# entry point -> (image the reference's code writes, arguments of this repository's CLI for the same job)
ENTRIES = {
    "main": ("single.bmp", ["render_single_scene", "4"]),                 # main(argc=2, "16") -> 16 / 4 spp (ky.cpp:4690)
    "render_debug": ("render_debug.bmp", ["render_debug"]),
    "render_multiple_integrator": ("direct_sample.bmp", ["render_multiple_integrator"]),
    "render_direct_sample_enum": ("direct_sample.bmp", ["render_direct_sample_enum"]),
    "render_multiple_scene": ("light_mis.bmp", ["render_multiple_scene"]),
    "render_mis_scene": ("veach_mis.bmp", ["render_mis_scene"]),
}


@pytest.mark.gpu
@pytest.mark.parametrize("entry", list(ENTRIES))
def test_reference_entry_points_render_the_same_images(entry, tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/ky_ref_entries not built (no /root/reference at build time)")
    image, cli_args = ENTRIES[entry]
    a, b = tmp_path / "reference_code", tmp_path / "cli"
    a.mkdir()
    b.mkdir()
    subprocess.run([EXE, entry] + (["16"] if entry == "main" else []), cwd=a, check=True, capture_output=True, timeout=600)
    subprocess.run([CLI] + cli_args, cwd=b, check=True, capture_output=True, timeout=600)
    got = (a / image).read_bytes()
    want = (b / (cli_args[0] + ".bmp")).read_bytes()
    assert len(got) == len(want) and len(got) > 54
    assert np.array_equal(np.frombuffer(got, np.uint8), np.frombuffer(want, np.uint8))
