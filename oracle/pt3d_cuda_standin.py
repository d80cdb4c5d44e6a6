"""oracle/pt3d_cuda_standin.py — TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

GPU stand-in for "the reference's PyTorch3D 0.3.0 CUDA path" of the mask render fwd + bwd (BASELINE.json's ">= 50x" target
has no other denominator: PyTorch3D is not in /root/reference and cannot be installed).  The rasterizer kernels are the
plain restatement in pt3d_cuda_standin.cu (coarse binning, one-thread-per-pixel fine kernel with a local top-K array,
per-fragment global-atomic backward); everything around them is the chain of torch ops PyTorch3D runs for
NeuralRenderer.forward's mask branch (/root/reference/multiframe/nnutils/nmr.py:143-172):

    Meshes packing            verts_packed[faces_packed]            (index_select; backward = index_add with atomics)
    SoftSilhouetteShader      texels = ones_like(bary)              (N,H,W,K,3) — allocated and discarded, as upstream does
    sigmoid_alpha_blend       mask = pix_to_face >= 0; prob = sigmoid(-dists / sigma) * mask; alpha = prod(1 - prob, -1);
                              image[..., 3] = 1 - alpha

Only tests/ and bench.py's gpu_standin leg import this module.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libacfm_pt3d_standin.so")
_lib = None
last_bin_max = 0   # fullest bin of the last render_mask(check_overflow=True) call


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def available():
    return os.path.exists(_LIB_PATH)


def default_bin_size(S):
    """rasterize_meshes(bin_size=None) on CUDA."""
    import math
    return 8 if S <= 64 else int(2 ** max(math.ceil(math.log2(S)) - 4, 4))


def default_max_faces_per_bin(num_verts_packed):
    return int(max(10000, num_verts_packed / 5))


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Rasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, face_verts, N, F, S, blur, K, bin_size, M, check_overflow):
        dev = face_verts.device
        B = 1 + (S - 1) // bin_size
        fpb = torch.empty((N, B, B), dtype=torch.int32, device=dev)
        bin_faces = torch.empty((N, B, B, M), dtype=torch.int32, device=dev)
        p2f = torch.full((N, S, S, K), -1, dtype=torch.int64, device=dev)
        zbuf = torch.full((N, S, S, K), -1.0, dtype=torch.float32, device=dev)
        dists = torch.full((N, S, S, K), -1.0, dtype=torch.float32, device=dev)
        bary = torch.full((N, S, S, K, 3), -1.0, dtype=torch.float32, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = lib().acfm_standin_rasterize(_p(face_verts), N, F, S, S, ctypes.c_float(blur), K, bin_size, M, _p(fpb), _p(bin_faces),
                                          _p(p2f), _p(zbuf), _p(dists), _p(bary), st)
        if rc:
            raise RuntimeError(f"acfm_standin_rasterize: cuda error {rc}")
        if check_overflow:   # PyTorch3D prints "Bin size was too small ..." and drops faces; a baseline must not
            global last_bin_max
            last_bin_max = int(fpb.max())
            if last_bin_max > M:
                raise RuntimeError(f"max_faces_per_bin={M} overflows: a bin holds {last_bin_max} faces")
        ctx.save_for_backward(face_verts, p2f)
        ctx.mark_non_differentiable(p2f, zbuf, bary)
        return p2f, zbuf, bary, dists

    @staticmethod
    def backward(ctx, _g0, _g1, _g2, grad_dists):
        face_verts, p2f = ctx.saved_tensors
        N, S, _, K = p2f.shape
        g = torch.zeros_like(face_verts)
        st = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
        rc = lib().acfm_standin_rasterize_backward(_p(face_verts), _p(p2f), _p(grad_dists.contiguous()), N, S, S, K, _p(g), st)
        if rc:
            raise RuntimeError(f"acfm_standin_rasterize_backward: cuda error {rc}")
        return g, None, None, None, None, None, None, None, None


def render_mask(ndc, faces, S, blur, K, sigma, max_faces_per_bin=None, check_overflow=False):
    """ndc (N,V,3) rasterizer-space vertices (requires_grad ok), faces (F,3) int64 shared topology -> mask (N,S,S), pix_to_face.
    max_faces_per_bin=None: PyTorch3D's default, max(10000, V_packed / 5)."""
    N, V, _ = ndc.shape
    F = faces.shape[0]
    verts_packed = ndc.reshape(N * V, 3)
    faces_packed = (faces[None] + (torch.arange(N, device=ndc.device) * V)[:, None, None]).reshape(N * F, 3)
    face_verts = verts_packed[faces_packed]                                   # (N*F,3,3)
    M = default_max_faces_per_bin(N * V) if max_faces_per_bin is None else int(max_faces_per_bin)
    p2f, zbuf, bary, dists = _Rasterize.apply(face_verts.contiguous(), N, F, S, float(blur), K, default_bin_size(S), M, check_overflow)
    texels = torch.ones_like(bary)                                            # SoftSilhouetteShader: colors = ones_like(bary)
    del texels
    m = (p2f >= 0).float()                                                    # sigmoid_alpha_blend
    prob = torch.sigmoid(-dists / sigma) * m
    alpha = torch.prod(1.0 - prob, dim=-1)
    return 1.0 - alpha, p2f
