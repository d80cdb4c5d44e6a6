"""ctypes binding of libacfm_b200.so (C ABI: include/acfm_b200.h).

The CUDA library is the product; there is no CPU or eager-PyTorch fallback.  Importing this
module without the built library raises, and every call on a non-CUDA tensor raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libacfm_b200.so")

_c_int, _c_f, _c_vp, _c_i64 = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int64
_pi = ctypes.POINTER(ctypes.c_int)

# name -> argtypes; must list every symbol include/acfm_b200.h declares (tests/test_abi.py checks)
SIGNATURES = {
    "acfm_version": [],
    "acfm_last_error_string": [],
    "acfm_set_raster_epsilon": [_c_f],
    "acfm_set_raster_bwd_headroom_bits": [_c_int],
    "acfm_get_raster_epsilon": [],
    "acfm_project_fwd": [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_f, _c_f, _c_f, _c_f, _c_vp, _c_vp],
    "acfm_project_bwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_f, _c_f, _c_vp, _c_vp, _c_vp],
    "acfm_skin_project_fwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_f, _c_f, _c_f, _c_f, _c_vp,
                              _c_vp, _c_vp],
    "acfm_skin_bwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp],
    "acfm_handle_solve_workspace_bytes": [_c_int, _c_int],
    "acfm_handle_solve_fwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, ctypes.c_double, _c_vp, _c_vp, _c_i64, _c_vp],
    "acfm_handle_solve_bwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp, _c_i64, _c_vp],
    "acfm_handle_solve_singular": [_c_vp, _c_int, _c_int, _c_vp],
    "acfm_softmax_cols_fwd": [_c_vp, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_softmax_cols_bwd": [_c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_raster_fwd": [_c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_int,
                        _c_int, _c_f, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_i64, _c_vp],
    "acfm_raster_fwd_workspace_bytes": [_c_int, _c_int, _c_int],
    "acfm_raster_loss_workspace_bytes": [_c_int, _c_int, _c_int, _c_int],
    "acfm_raster_fwd_train": [_c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_f, _c_vp, _c_vp,
                              _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp],
    "acfm_raster_soft_bwd_train": [_c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_vp, _c_vp,
                                   _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp],
    "acfm_raster_lean_workspace_bytes": [_c_int, _c_int, _c_int, _c_int],
    "acfm_raster_fwd_lean": [_c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_f, _c_vp, _c_vp,
                             _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp, _c_i64, _c_vp],
    "acfm_raster_soft_bwd_lean": [_c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _c_vp, _c_vp,
                                  _c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp],
    "acfm_raster_soft_bwd": [_c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f,
                             _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp, _c_vp],
    "acfm_raster_dists_bwd": [_c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp,
                              _c_vp, _c_vp, _c_vp],
    "acfm_shade_fwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_int, _c_int, _c_int,
                       _c_int, _c_vp, _c_int, _c_i64, _c_f, _c_f, _c_f, _c_f, _c_vp, _c_vp],
    "acfm_shade_bwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_int, _c_int, _c_int,
                       _c_int, _c_vp, _c_int, _c_i64, _c_f, _c_f, _c_f, _c_f, _c_vp, _c_vp, _c_i64, _c_vp, _c_vp],
    "acfm_camera_assemble_fwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_f, _c_vp, _c_vp],
    "acfm_camera_assemble_bwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_f, _c_vp, _c_vp],
    "acfm_uv_sample_fwd": [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_uv_sample_bwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_mask_sums_fwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_mask_sums_bwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_mask_loss_combine_fwd": [_c_vp, _c_int, _c_int, _c_f, _c_f, _c_f, _c_vp, _c_vp],
    "acfm_mask_loss_combine_bwd": [_c_vp, _c_vp, _c_int, _c_int, _c_f, _c_f, _c_f, _c_vp, _c_vp],
    "acfm_visible_verts": [_c_vp, _c_i64, _c_vp, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_bds_loss_fwd": [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp],
    "acfm_bds_loss_bwd": [_c_vp, _c_int, _c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_of_loss_fwd": [_c_vp, _c_int, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp,
                         _c_vp, _c_vp],
    "acfm_of_loss_bwd": [_c_vp, _c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_kp_loss_fwd": [_c_vp, _c_int, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_kp_loss_bwd": [_c_vp, _c_int, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_hypothesis_weight_fwd": [_c_vp, _c_int, _c_int, _c_vp, _c_vp, _c_vp],
    "acfm_hypothesis_weight_bwd": [_c_vp, _c_vp, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_laplacian_fwd": [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_edt_fwd": [_c_vp, _c_int, _c_int, _c_int, _c_f, _c_int, _c_vp, _c_vp, _c_vp, _c_vp],
    "acfm_boundaries_count": [_c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp],
    "acfm_boundaries_write": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_laplacian_smoothing_fwd": [_c_vp, _c_vp, _c_int, _c_i64, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp, _c_vp],
    "acfm_laplacian_smoothing_bwd": [_c_vp, _c_vp, _c_int, _c_i64, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_edge_rigidity_fwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_edge_rigidity_bwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp, _c_vp],
    "acfm_raster_fwd_launch_info": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _pi, _pi, _pi],
    "acfm_correlation_out_shape": [_c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _pi, _pi, _pi],
    "acfm_correlation_fwd": [_c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp, _c_vp],
    "acfm_correlation_bwd": [_c_vp, _c_vp, _c_vp, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_vp,
                             _c_vp, _c_vp],
}

_lib = None


def lib():
    """Load (once) and return the CDLL.  Raises if the library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or acfm_video_3d_reconstruction_b200/csrc/build.sh). There is no CPU fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)
            fn.argtypes = argtypes
            fn.restype = (ctypes.c_char_p if name == "acfm_last_error_string" else
                          ctypes.c_int64 if name in ("acfm_raster_fwd_workspace_bytes", "acfm_handle_solve_workspace_bytes",
                                                    "acfm_raster_loss_workspace_bytes", "acfm_raster_lean_workspace_bytes") else
                          ctypes.c_float if name == "acfm_get_raster_epsilon" else ctypes.c_int)
        if os.environ.get("ACFM_NVTX", "0") not in ("", "0"):
            l = _NvtxLib(l)   # every C-ABI call becomes an NVTX range named after the entry point (nsys / ncu --nvtx)
        _lib = l
    return _lib


class _NvtxLib:
    """Wraps the CDLL so that each entry point runs inside an NVTX range (ACFM_NVTX=1): the kernels of one call group under
    its name in a timeline, e.g. acfm_raster_fwd = prep + padding + rasterizer.  Off by default: two extra Python calls per launch."""

    def __init__(self, cdll):
        self._cdll = cdll

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)
        if not name.startswith("acfm_") or name in ("acfm_last_error_string", "acfm_version"):
            return fn

        def ranged(*args):
            torch.cuda.nvtx.range_push(name)
            try:
                return fn(*args)
            finally:
                torch.cuda.nvtx.range_pop()
        setattr(self, name, ranged)
        return ranged


def check(status, what):
    """Map an acfm_status to the exception the reference would raise (SURVEY.md §8b Errors)."""
    if status == 0:
        return
    msg = lib().acfm_last_error_string().decode("utf-8", "replace")
    if status in (1, 2):
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("acfm_video_3d_reconstruction_b200 runs on CUDA tensors only (no CPU fallback); "
                               f"got a tensor on {t.device}")


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def stream_of(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


# optional callable(name, phase) invoked right before (0) / after (1) a dominant kernel's launch; bench.py records
# CUDA events in it to time that kernel on its own stream inside the timed region
event_hook = None

# count of kernels launched through the C ABI by this process (bench.py reports it)
launches = 0


def count(n=1):
    global launches
    launches += n
