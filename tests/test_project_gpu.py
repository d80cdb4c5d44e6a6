"""GPU parity: acfm_project_fwd/bwd (through the geom_utils / NeuralRenderer mirrors) vs the golden
vectors generated from the reference's geom_utils.py and vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import pt3d_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


def test_projection_bit_exact_vs_reference_golden():
    from acfm_video_3d_reconstruction_b200 import geom_utils
    g = util.golden("projection.npz")
    X, cam = torch.from_numpy(g["X"]).cuda(), torch.from_numpy(g["cam"]).cuda()
    for oz in (0.0, 5.0):
        out = geom_utils.orthographic_proj_withz(X, cam, offset_z=oz).cpu().numpy()
        assert np.array_equal(out, g[f"withz_{int(oz)}"])
    assert np.array_equal(geom_utils.orthographic_proj(X, cam).cpu().numpy(), g["proj"])
    assert np.array_equal(geom_utils.quat_rotate(X, cam[:, 3:]).cpu().numpy(), g["quat_rotate"])


def test_projection_backward_vs_reference_golden():
    from acfm_video_3d_reconstruction_b200 import geom_utils
    g = util.golden("projection.npz")
    X = torch.from_numpy(g["X"]).cuda().requires_grad_(True)
    cam = torch.from_numpy(g["cam"]).cuda().requires_grad_(True)
    w = torch.from_numpy(g["grad_w"]).float().cuda()
    (geom_utils.orthographic_proj_withz(X, cam, offset_z=5.0) * w).sum().backward()
    # tolerance: vertex/camera gradients within 1e-3 relative (BASELINE.json north_star); fp32 vs fp64 truth
    assert util.rel_err(X.grad.cpu().numpy(), g["grad_X"]) < 1e-5
    assert util.rel_err(cam.grad.cpu().numpy(), g["grad_cam"]) < 1e-4


def test_multiplex_broadcast_and_ndc_flags():
    """verts (NB,V,3) shared by G hypotheses (n = g*NB + b) == the reference's pred_v.repeat(G,1,1)."""
    from acfm_video_3d_reconstruction_b200 import functional as F_
    v, _ = util.template("horse")
    NB, G = 3, 4
    X = util.synth_verts(v, NB, seed=5)
    cam = util.synth_cams(NB * G, seed=6)
    ref = orc.view(orc.project(np.tile(X, (G, 1, 1)), cam, 0.0), yflip=True)
    Xc = torch.from_numpy(X).cuda().requires_grad_(True)
    cc = torch.from_numpy(cam).cuda().requires_grad_(True)
    out = F_.project(Xc, cc, offset_z=0.0, sx=-1.0, sy=-1.0, z_add=F_.EYE_Z)
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    # backward: grad_verts sums over the G renders of each mesh
    w = torch.randn_like(out)
    (out * w).sum().backward()
    Xr = torch.from_numpy(np.tile(X, (G, 1, 1))).cuda().requires_grad_(True)
    cr = torch.from_numpy(cam).cuda().requires_grad_(True)
    (F_.project(Xr, cr, offset_z=0.0, sx=-1.0, sy=-1.0, z_add=F_.EYE_Z) * w).sum().backward()
    assert util.rel_err(Xc.grad.cpu().numpy(), Xr.grad.view(G, NB, -1, 3).sum(0).cpu().numpy()) < 1e-5
    assert util.rel_err(cc.grad.cpu().numpy(), cr.grad.cpu().numpy()) < 1e-5


def test_empty_batch():
    from acfm_video_3d_reconstruction_b200 import geom_utils
    out = geom_utils.orthographic_proj_withz(torch.zeros(0, 5, 3).cuda(), torch.zeros(0, 7).cuda())
    assert out.shape == (0, 5, 3)
