"""GPU parity of the texture branch (hard raster + atlas / vertex-colour shading + softmax_rgb_blend) against
the oracle restatement (oracle/pt3d_oracle.py: hard_raster, atlas_shade, blend_texels)."""
import numpy as np
import pytest
import torch

from oracle import pt3d_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


def _setup(N=3, S=128, T=6, seed=0):
    v, f = util.template("horse")
    X, cam = util.synth_verts(v, N, seed=seed), util.synth_cams(N, seed=seed + 1)
    faces = np.repeat(f[None], N, 0)
    rng = np.random.default_rng(seed)
    atlas = rng.random((N, f.shape[0], T, T, 3)).astype(np.float32)
    return v, f, X, cam, faces, atlas, rng


def test_atlas_render_vs_oracle():
    from acfm_video_3d_reconstruction_b200 import NeuralRenderer
    v, f, X, cam, faces, atlas, rng = _setup()
    S = 128
    fr = orc.hard_raster(X, faces, cam, img_size=S, offset_z=0.0)
    imgs_ref, sil_ref = orc.atlas_shade(fr, atlas)
    r = NeuralRenderer(S)
    tex = torch.from_numpy(atlas).cuda().requires_grad_(True)
    imgs, sil, p2f = r(torch.from_numpy(X).cuda(), torch.from_numpy(faces).cuda(), torch.from_numpy(cam).cuda(), textures=tex)
    assert imgs.shape == (3, 3, S, S) and sil.shape == (3, S, S) and p2f.shape == (3, S, S, 1)
    assert not r.mask_only
    assert np.array_equal(p2f.cpu().numpy(), fr["pix_to_face"])
    # tolerance: 1e-5 of the unit colour range (silhouettes / images within 1e-5, BASELINE.json north_star)
    assert np.abs(imgs.detach().cpu().numpy() - imgs_ref).max() <= 1e-5
    assert np.abs(sil.detach().cpu().numpy() - sil_ref).max() <= 1e-5
    cov = fr["pix_to_face"][..., 0] >= 0
    assert cov.mean() > 0.03 and (imgs_ref.transpose(0, 2, 3, 1)[~cov] == 0).all()
    # gradient w.r.t. the atlas: rgb is linear in the texels => scatter of g * wnum/denom
    g = rng.standard_normal(imgs_ref.shape).astype(np.float32)
    (imgs * torch.from_numpy(g).cuda()).sum().backward()
    prob = 1.0 / (1.0 + np.exp(fr["dists"][..., 0].astype(np.float64) / 1e-4))
    w = np.where(cov, prob / (prob + 1e-10), 0.0)
    R = atlas.shape[2]
    b = fr["bary"][..., 0, :2]
    wxy = np.floor(b * np.float32(R)).astype(np.int64)
    below = (b.sum(-1) * np.float32(R) - wxy.astype(np.float32).sum(-1)) <= 1.0
    wx = np.clip(np.where(below, wxy[..., 0], R - 1 - wxy[..., 0]), 0, R - 1)
    wy = np.clip(np.where(below, wxy[..., 1], R - 1 - wxy[..., 1]), 0, R - 1)
    gref = np.zeros((atlas.size // 3, 3))
    idx = (fr["pix_to_face"][..., 0] * R + wy) * R + wx
    np.add.at(gref, idx[cov], (g.transpose(0, 2, 3, 1) * w[..., None])[cov])
    assert util.rel_err(tex.grad.cpu().numpy().reshape(-1, 3), gref) < 1e-3


def test_vertex_colour_render_and_vertex_gradient():
    from acfm_video_3d_reconstruction_b200 import NeuralRenderer
    v, f, X, cam, faces, _, rng = _setup(N=2, S=96, seed=5)
    S = 96
    colors = rng.random((1, v.shape[0], 3)).astype(np.float32)
    fr = orc.hard_raster(X, faces, cam, img_size=S, offset_z=0.0)
    imgs_ref, sil_ref = orc.blend_texels(fr, orc.vertex_colors_as_texels(fr, colors, faces))
    r = NeuralRenderer(S)
    Xc = torch.from_numpy(X).cuda().requires_grad_(True)
    col = torch.from_numpy(colors[0]).cuda().requires_grad_(True)  # bird_vis.py passes a (V,3) tensor, int32 faces
    imgs, sil, p2f = r(Xc, torch.from_numpy(faces).int().cuda(), torch.from_numpy(cam).cuda(), textures=col, atlas=False)
    assert np.array_equal(p2f.cpu().numpy(), fr["pix_to_face"])
    assert np.abs(imgs.detach().cpu().numpy() - imgs_ref).max() <= 1e-5
    assert np.abs(sil.detach().cpu().numpy() - sil_ref).max() <= 1e-5
    # silhouette gradient to the vertices (through dists) vs the oracle backward
    gs = rng.standard_normal(sil_ref.shape).astype(np.float32)
    (sil * torch.from_numpy(gs).cuda()).sum().backward()
    gd = orc.sigmoid_alpha_blend_backward(fr["dists"], fr["pix_to_face"], gs)
    g_ndc = orc.scatter_face_grads(orc.rasterize_backward(fr["face_verts"], fr["pix_to_face"], grad_dists=gd), faces, v.shape[0])
    from oracle import torch_ref
    Xd = torch.from_numpy(X).double().requires_grad_(True)
    (torch_ref.to_ndc(Xd, torch.from_numpy(cam).double(), 0.0) * torch.from_numpy(g_ndc).double()).sum().backward()
    assert util.rel_err(Xc.grad.cpu().numpy(), Xd.grad.numpy()) < 1e-3
    assert col.grad is not None and torch.isfinite(col.grad).all()
