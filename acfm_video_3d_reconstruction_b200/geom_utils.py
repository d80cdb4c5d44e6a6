"""Drop-in for the projection half of the reference's nnutils/geom_utils.py
(/root/reference/multiframe/nnutils/geom_utils.py:48-153), backed by acfm_project_fwd/bwd.

Same names, argument meaning and output shapes as the reference; inputs must be CUDA float32.
"""
import torch

from . import _lib
from . import functional as F_


def orthographic_proj_withz(X, cam, offset_z=0.):
    """X: B x N x 3, cam: B x 7 [sc, tx, ty, quaternions] -> B x N x 3 (geom_utils.py:62-79)."""
    return F_.project(X, cam, offset_z=offset_z)


def orthographic_proj(X, cam):
    """X: B x N x 3, cam: B x 7 -> B x N x 2 (geom_utils.py:48-59)."""
    return F_.project(X, cam, offset_z=0.0)[:, :, :2]


def quat_rotate(X, q):
    """Rotate points by (un-normalised) quaternions: X B x N x 3, q B x 4 (geom_utils.py:131-153)."""
    import torch
    cam = torch.cat([torch.ones_like(q[:, :1]), torch.zeros_like(q[:, :2]), q], dim=1)
    return F_.project(X, cam, offset_z=0.0)


def mesh_laplacian(meshes, method="uniform", faces=None):
    """Dense (V,V) Laplacian of ONE mesh (geom_utils.py:159-255).  `meshes` is either a (V,3) vertex tensor with
    `faces` (F,3), or any object with verts_packed() / faces_packed() (the reference passes a PyTorch3D Meshes
    holding the template: monocular/main.py:124, multiframe/main.py:600-601).  method 'uniform' | 'cot'.
    No gradient, as in the reference."""
    if faces is None:
        verts, faces = meshes.verts_packed(), meshes.faces_packed()
    else:
        verts = meshes
    if method not in ("uniform", "cot"):
        raise ValueError(f"mesh_laplacian: unknown method {method!r}")
    _lib.require_cuda(verts, faces)
    with torch.no_grad():
        verts = F_._f32c(verts.detach())
        fa, i64, _, F = F_._faces_arg(faces, 1)
        V = verts.shape[0]
        L = torch.empty((V, V), dtype=torch.float32, device=verts.device)
        with torch.cuda.device(verts.device):
            st = _lib.lib().acfm_laplacian_fwd(_lib.ptr(verts), _lib.ptr(fa), i64, V, F, 0 if method == "uniform" else 1,
                                               _lib.ptr(L), _lib.stream_of(verts))
        _lib.check(st, "acfm_laplacian_fwd")
        _lib.count(3)
    return L
