"""Drop-in for the projection half of the reference's nnutils/geom_utils.py
(/root/reference/multiframe/nnutils/geom_utils.py:48-153), backed by acfm_project_fwd/bwd.

Same names, argument meaning and output shapes as the reference; inputs must be CUDA float32.
"""
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
