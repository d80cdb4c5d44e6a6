"""CPU tests of the drop-in boundary: libacfm_b200.so loads without a GPU and exports exactly the
symbols include/acfm_b200.h declares; argument validation returns status codes (no compute)."""
import ctypes
import os
import re

import pytest
import torch

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "acfm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(acfm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from acfm_video_3d_reconstruction_b200 import _lib
    lib = _lib.lib()
    names = _declared()
    assert "acfm_raster_fwd" in names and "acfm_project_fwd" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/acfm_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    assert lib.acfm_version() >= 100


def test_bad_arguments_return_status_not_crash():
    from acfm_video_3d_reconstruction_b200 import _lib
    lib = _lib.lib()
    # K = 0 and K > 64 are rejected before any CUDA call
    st = lib.acfm_raster_fwd(None, None, 1, 0, 1, 3, 1, 8, 8, 0, 0.0, 0, 0, 0.0, None, None, None, None, None, None, None, 0, None)
    assert st == 1 and b"faces_per_pixel" in lib.acfm_last_error_string()
    st = lib.acfm_raster_fwd(None, None, 1, 0, 1, 3, 1, 8, 8, 65, 0.0, 0, 0, 0.0, None, None, None, None, None, None, None, 0, None)
    assert st == 2
    with pytest.raises(ValueError):
        _lib.check(st, "acfm_raster_fwd")
    # N not a multiple of NB
    st = lib.acfm_project_fwd(ctypes.c_void_p(16), ctypes.c_void_p(16), 3, 2, 4, 0.0, 1.0, 1.0, 0.0, ctypes.c_void_p(16), None)
    assert st == 1
    smem, ctas, thr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.acfm_raster_fwd_launch_info(512, 642, 1280, 256, 256, 20, smem, ctas, thr) == 0
    assert ctas.value == 512 * 64 and thr.value in (128, 256) and 0 < smem.value <= 227 * 1024


def test_cpu_tensors_are_refused():
    from acfm_video_3d_reconstruction_b200 import NeuralRenderer
    r = NeuralRenderer(32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        r(torch.zeros(1, 3, 3), torch.zeros(1, 1, 3, dtype=torch.int64), torch.zeros(1, 7))
