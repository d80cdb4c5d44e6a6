"""GPU parity of the reprojection-loss kernels (visible vertices, bds_loss, optical_flow_loss, kp_l2_loss, hypothesis
weighting) and the Laplacian kernels, through the package's loss_utils / geom_utils mirrors (C ABI underneath), against
  * tests/golden/reproj.npz + losses.npz — values and fp64 gradients produced by the REFERENCE's own code, and
  * oracle/torch_ref.py on fresh seeded inputs.
Bars: losses within 1e-5 relative, gradients within 1e-3 relative (BASELINE.json north_star); index-like outputs
(visibility bitmaps, sampled flow) exact."""
import numpy as np
import pytest
import torch

from oracle import pt3d_oracle as orc
from oracle import torch_ref
from tests import util

pytestmark = pytest.mark.gpu


def _horse_faces():
    _, f = util.template("horse")
    return torch.from_numpy(f)


def test_bds_loss_vs_reference_golden():
    from acfm_video_3d_reconstruction_b200 import loss_utils
    g = util.golden("reproj.npz")
    hf = _horse_faces()
    N = g["bds_proj"].shape[0]
    faces = hf[None].repeat(N, 1, 1).cuda()
    p2f = torch.from_numpy(g["bds_p2f"]).long().cuda()
    proj = torch.from_numpy(g["bds_proj"]).cuda().requires_grad_(True)
    bds = torch.from_numpy(g["bds_pts"]).cuda()                      # NB = 2 rows for N = 4 renders: no G-fold repeat
    vis = loss_utils.visible_vertices(p2f, faces, proj.shape[1])
    ref_vis = torch_ref.visible_vertices_v(p2f.cpu(), faces.cpu(), proj.shape[1])
    assert torch.equal(vis.cpu() != 0, ref_vis) and 0.2 < ref_vis.float().mean() < 0.9
    torch.manual_seed(int(g["bds_seed"]))                            # same draw as the reference made
    loss = loss_utils.bds_loss(proj, bds, faces, p2f, reduce=False, n_samples=200)
    assert np.allclose(loss.detach().cpu().numpy(), g["bds_loss"], rtol=2e-5, atol=0)
    (loss * torch.from_numpy(g["bds_w"]).float().cuda()).sum().backward()
    assert util.rel_err(proj.grad.cpu().numpy(), g["bds_grad"]) < 1e-3
    # shared faces (stride 0), int32 faces, reduce=True, 3-component verts (xy used in place)
    torch.manual_seed(5)
    a = loss_utils.bds_loss(proj.detach(), bds, faces[:1].expand(N, -1, -1), p2f, n_samples=1000)
    torch.manual_seed(5)
    p3 = torch.cat([proj.detach(), torch.ones_like(proj[..., :1])], -1)
    b = loss_utils.bds_loss(p3, bds.repeat(2, 1, 1), faces.int(), p2f, reduce=False, n_samples=1000)
    assert torch.allclose(a, b.mean(), rtol=1e-6)


def test_bds_all_invisible_and_all_visible():
    from acfm_video_3d_reconstruction_b200 import loss_utils
    hf = _horse_faces().cuda()
    V = 642
    verts = (torch.rand(1, V, 2, device="cuda") - 0.5).requires_grad_(True)
    bds = torch.cat([torch.rand(1, 50, 2, device="cuda"), torch.ones(1, 50, 1, device="cuda")], -1)
    empty = torch.full((1, 16, 16, 1), -1, dtype=torch.int64, device="cuda")
    l = loss_utils.bds_loss(verts, bds, hf[None], empty, reduce=False)
    assert float(l.detach()) == 50 * 1000.0                                    # nothing visible: every distance is 1000
    l.sum().backward()
    assert float(verts.grad.abs().sum()) == 0.0
    allf = torch.arange(16 * 16 * 5, dtype=torch.int64, device="cuda").reshape(1, 16, 16, 5) % hf.shape[0]
    allf[..., 0] = torch.arange(256, device="cuda").reshape(1, 16, 16) * 5 % hf.shape[0]
    ref = torch_ref.bds_loss(verts.detach().cpu(), bds.cpu(), hf[None].cpu(), allf.cpu(), torch.arange(50))
    torch.manual_seed(0)
    l2 = loss_utils.bds_loss(verts, bds, hf[None], allf, reduce=False)
    assert torch.allclose(l2.cpu(), ref, rtol=1e-5)


def test_optical_flow_loss_vs_reference_golden():
    from acfm_video_3d_reconstruction_b200 import OF_NeuralRenderer, loss_utils
    g = util.golden("reproj.npz")
    hf = _horse_faces()
    M = torch.from_numpy(g["of_meshes"]).cuda().requires_grad_(True)
    cams = torch.from_numpy(g["of_cams"]).cuda().requires_grad_(True)
    flows = torch.from_numpy(g["of_flows"]).cuda()                   # (1,T,H,W,2): shared by both sequences
    B, T, V, _ = M.shape
    faces_of = hf[None, None].repeat(B, T, 1, 1).cuda()
    r = OF_NeuralRenderer(flows.shape[2])
    loss, of_pred, vis, pts, smp = loss_utils.optical_flow_loss(M, faces_of, cams, flows, r, None, reduce=False)
    assert np.array_equal(vis.cpu().numpy(), g["of_vis"])
    assert np.array_equal(smp.cpu().numpy(), g["of_samples"])
    assert np.allclose(of_pred.detach().cpu().numpy(), g["of_pred"], rtol=0, atol=1e-5)
    assert np.allclose(pts.detach().cpu().numpy(), g["of_pts"], rtol=0, atol=1e-6)
    assert np.allclose(loss.detach().cpu().numpy(), g["of_loss"], rtol=1e-5, atol=0)
    (loss * torch.from_numpy(g["of_w"]).float().cuda()).sum().backward()
    assert util.rel_err(M.grad.cpu().numpy(), g["of_grad_meshes"]) < 1e-3
    assert util.rel_err(cams.grad.cpu().numpy(), g["of_grad_cams"]) < 1e-3
    # reduce=True and an explicit pix_to_face (the reference's other branch); flows given per sequence
    p2f = r(r.proj_fn(M.detach().reshape(B * T, V, 3), cams.detach()), faces_of.reshape(B * T, -1, 3))
    l2 = loss_utils.optical_flow_loss(M.detach(), faces_of, cams.detach(), flows.repeat(B, 1, 1, 1, 1), r, p2f.repeat(1, 1, 1, 3))[0]
    assert torch.allclose(l2, loss.detach().sum(), rtol=1e-6)


def test_kp_loss_vs_reference_golden():
    from acfm_video_3d_reconstruction_b200 import loss_utils
    g = util.golden("losses.npz")
    kp_pred = torch.from_numpy(g["kp_pred"]).cuda().requires_grad_(True)
    kp_gt = torch.from_numpy(g["kp_gt"]).cuda()
    l = loss_utils.kp_l2_loss(kp_pred, kp_gt, reduction="none")
    assert np.allclose(l.detach().cpu().numpy(), g["kp"], rtol=1e-6, atol=0)
    assert np.allclose(float(loss_utils.kp_l2_loss(kp_pred, kp_gt).detach()), g["kp_mean"], rtol=1e-6)
    w = torch.rand(kp_pred.shape[0], device="cuda")
    (l * w).sum().backward()
    pd = torch.from_numpy(g["kp_pred"]).double().requires_grad_(True)
    (torch_ref.kp_l2_loss(pd, torch.from_numpy(g["kp_gt"]).double()) * w.cpu().double()).sum().backward()
    assert util.rel_err(kp_pred.grad.cpu().numpy(), pd.grad.numpy()) < 1e-5
    # G-fold broadcast of the targets (N = 2 NB), strided prediction (xy of a projected (N,Kp,3) tensor)
    p3 = torch.cat([kp_pred.detach().repeat(2, 1, 1), torch.zeros(10, 15, 1, device="cuda")], -1)
    l3 = loss_utils.kp_l2_loss(p3, kp_gt, reduction="none")
    assert torch.allclose(l3, l.detach().repeat(2))


def test_hypothesis_weighting_vs_reference_golden():
    from acfm_video_3d_reconstruction_b200 import loss_utils
    g = util.golden("reproj.npz")
    tl = torch.from_numpy(g["hyp_loss"]).float().cuda().requires_grad_(True)
    tot, probs = loss_utils.hypothesis_weighting(tl)
    assert np.allclose(float(tot), g["hyp_total"], rtol=1e-6) and np.allclose(probs.cpu().numpy(), g["hyp_probs"], rtol=1e-5)
    (3.0 * tot).backward()
    assert np.allclose(tl.grad.cpu().numpy(), 3.0 * g["hyp_grad"], rtol=1e-5)


@pytest.mark.parametrize("name", ["horse", "bird"])
def test_laplacians(name):
    from acfm_video_3d_reconstruction_b200 import geom_utils
    v, f = util.template(name)
    vt, ft = torch.from_numpy(v), torch.from_numpy(f)
    Lc = geom_utils.mesh_laplacian(vt.cuda(), "cot", faces=ft.cuda())
    ref = torch_ref.laplacian_cot(vt.double(), ft).numpy()
    assert util.rel_err(Lc.cpu().numpy(), ref) < 1e-4
    if name == "horse":
        assert util.rel_err(Lc.cpu().numpy(), util.golden("reproj.npz")["lap_cot"]) < 1e-5   # the reference's own output
    Lu = geom_utils.mesh_laplacian(vt.cuda(), "uniform", faces=ft.int().cuda())
    assert torch.equal(Lu.cpu(), torch_ref.laplacian_uniform(v.shape[0], ft))

    class M:  # duck-typed Meshes, as the reference passes it
        def verts_packed(self): return vt.cuda()
        def faces_packed(self): return ft.cuda()
    assert torch.equal(geom_utils.mesh_laplacian(M(), "uniform"), Lu)
