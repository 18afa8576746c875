"""CPU: the C-ABI library loads, exports every symbol include/kyd.h declares, and refuses to work
without a GPU instead of falling back."""
import ctypes as C
import os
import re

import pytest

import ky_b200 as ky

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kyd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kyd_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_documented_entry_points():
    assert declared_symbols() == sorted(ky.KYD_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ky.kyd()
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_struct_layouts_match_the_header():
    # sizes the C compiler gives the PODs (kyd.h is plain C): checked through a tiny compiled probe
    import subprocess, tempfile
    src = r'''
    #include <stdio.h>
    #include "kyd.h"
    int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(kyd_shape), sizeof(kyd_material), sizeof(kyd_light),
        sizeof(kyd_surface), sizeof(kyd_camera), sizeof(kyd_scene_desc), sizeof(kyd_render_desc), sizeof(kyd_stats)); return 0; }
    '''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "probe.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "probe")
        subprocess.run(["gcc", "-I" + os.path.join(ROOT, "include"), c, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()]
    want = [C.sizeof(t) for t in (ky.Shape, ky.Material, ky.Light, ky.Surface, ky.Camera, ky.SceneDesc, ky.RenderDesc, ky.Stats)]
    assert sizes == want


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="a GPU is present")
def test_no_gpu_means_an_error_not_a_fallback():
    with pytest.raises(RuntimeError, match="no CUDA device|kyd_create failed"):
        ky.Device(0)
    # and the host class surface propagates it as an exception (LOG_ERROR throws in the reference)
    with pytest.raises(RuntimeError):
        ky.render_entry("render_debug", 16, 16, 1)
