"""CPU tests: the oracle against the golden vectors generated from the reference, against closed-form
cases, and its backward against fp64 finite differences (SURVEY.md §4 'test pyramid')."""
import numpy as np
import pytest
import torch

from oracle import pt3d_oracle as orc
from oracle import torch_ref
from tests import util


def test_projection_bit_exact_vs_reference_golden():
    g = util.golden("projection.npz")
    for oz in (0.0, 5.0):
        out = orc.project(g["X"], g["cam"], oz)
        assert np.array_equal(out, g[f"withz_{int(oz)}"]), "C oracle projection differs from reference geom_utils"
    # torch restatement, fp32: same op order => bit-exact too
    t = torch_ref.orthographic_proj_withz(torch.from_numpy(g["X"]), torch.from_numpy(g["cam"]), 5.0).numpy()
    assert np.array_equal(t, g["withz_5"])
    qr = torch_ref.quat_rotate(torch.from_numpy(g["X"]), torch.from_numpy(g["cam"][:, 3:])).numpy()
    assert np.array_equal(qr, g["quat_rotate"])


def test_projection_gradient_restatement_vs_reference_golden():
    g = util.golden("projection.npz")
    X = torch.from_numpy(g["X"]).double().requires_grad_(True)
    cam = torch.from_numpy(g["cam"]).double().requires_grad_(True)
    (torch_ref.orthographic_proj_withz(X, cam, 5.0) * torch.from_numpy(g["grad_w"])).sum().backward()
    assert util.rel_err(X.grad.numpy(), g["grad_X"]) < 1e-12
    assert util.rel_err(cam.grad.numpy(), g["grad_cam"]) < 1e-12


def test_losses_restatement_vs_reference_golden():
    g = util.golden("losses.npz")
    pred, targ, edt = (torch.from_numpy(g[k]) for k in ("pred", "targ", "edt"))
    assert np.allclose(torch_ref.l1_loss(pred, targ).numpy(), g["l1"], rtol=1e-6, atol=0)
    assert np.allclose(torch_ref.iou_loss(pred, targ).numpy(), g["iou"], rtol=1e-6, atol=0)
    assert np.allclose(torch_ref.edt_loss(pred, edt).numpy(), g["edt_l"], rtol=1e-6, atol=0)
    assert np.allclose(torch_ref.kp_l2_loss(torch.from_numpy(g["kp_pred"]), torch.from_numpy(g["kp_gt"])).numpy(),
                       g["kp"], rtol=1e-6, atol=0)


def test_raster_regression_pin():
    g = util.golden("raster_small.npz")
    N = g["X"].shape[0]
    faces = np.repeat(g["faces"][None], N, 0)
    fr = orc.neural_renderer_mask(g["X"], faces, g["cam"], img_size=64, offset_z=5.0)
    assert np.array_equal(fr["ndc"], g["ndc"])
    assert np.array_equal(fr["pix_to_face"], g["pix_to_face"].astype(np.int64))
    assert np.array_equal(fr["zbuf"], g["zbuf"]) and np.array_equal(fr["dists"], g["dists"])
    assert np.allclose(fr["mask"], g["mask"], atol=1e-6)
    gn = orc.neural_renderer_mask_backward(fr, faces, g["grad_mask"])
    assert util.rel_err(gn, g["grad_ndc"]) < 1e-5


def _one_triangle():
    # NDC triangle, z = 1; PyTorch3D NDC: +x left, +y up
    v = np.array([[[-0.5, -0.5, 1.0], [0.5, -0.5, 1.0], [0.0, 0.5, 1.0]]], np.float32)
    f = np.array([[[0, 1, 2]]], np.int64)
    return v, f


def test_single_triangle_closed_form():
    v, f = _one_triangle()
    S, K = 32, 4
    blur = 0.01
    fr = orc.rasterize(v, f, S, blur, K)
    p2f, d, z = fr["pix_to_face"], fr["dists"], fr["zbuf"]
    assert (p2f[..., 1:] == -1).all() and (d[..., 1:] == -1).all()
    for yi in range(S):
        for xi in range(S):
            x = -1 + (2 * (S - 1 - xi) + 1) / S
            y = -1 + (2 * (S - 1 - yi) + 1) / S
            # inside test via half-planes
            inside = (y > -0.5) and (y < 0.5 - 2 * abs(x)) if abs(x) < 0.5 else False
            # exact squared distance to the three segments
            def seg(ax, ay, bx, by):
                t = ((x - ax) * (bx - ax) + (y - ay) * (by - ay)) / ((bx - ax) ** 2 + (by - ay) ** 2)
                t = min(max(t, 0.0), 1.0)
                return (x - ax - t * (bx - ax)) ** 2 + (y - ay - t * (by - ay)) ** 2
            dist = min(seg(-0.5, -0.5, 0.5, -0.5), seg(-0.5, -0.5, 0.0, 0.5), seg(0.5, -0.5, 0.0, 0.5))
            if inside:
                assert p2f[0, yi, xi, 0] == 0 and abs(d[0, yi, xi, 0] + dist) < 1e-6 and abs(z[0, yi, xi, 0] - 1) < 1e-5
            elif dist < blur * 0.999:
                assert p2f[0, yi, xi, 0] == 0 and abs(d[0, yi, xi, 0] - dist) < 1e-6
            elif dist > blur * 1.001:
                assert p2f[0, yi, xi, 0] == -1
    m = orc.sigmoid_alpha_blend(d, p2f, 1e-3)
    k = p2f[..., 0] >= 0
    assert np.allclose(m[k], 1 / (1 + np.exp(d[..., 0][k] / 1e-3)), atol=1e-6) and (m[~k] == 0).all()


def test_topk_by_depth_and_packed_ids():
    # three stacked copies of the triangle at z = 3, 1, 2 in two meshes: K=2 keeps the two nearest, sorted
    v, _ = _one_triangle()
    vs = np.concatenate([v + [0, 0, 2.0], v, v + [0, 0, 1.0]], 1).astype(np.float32)
    vs = np.concatenate([vs, vs], 0)
    f = np.array([[0, 1, 2], [3, 4, 5], [6, 7, 8]], np.int64)[None].repeat(2, 0)
    fr = orc.rasterize(vs, f, 16, 0.0, 2)
    c = fr["pix_to_face"][0, 8, 8]
    assert list(c) == [1, 2] and np.allclose(fr["zbuf"][0, 8, 8], [1, 2])
    assert list(fr["pix_to_face"][1, 8, 8]) == [3 + 1, 3 + 2]  # packed ids n*F+f
    # behind-camera faces are dropped
    vb = v.copy(); vb[..., 2] = -1
    assert (orc.rasterize(vb, f[:1, :1], 16, 0.0, 1)["pix_to_face"] == -1).all()


def test_hard_raster_clip_and_of_no_yflip():
    v, f = _one_triangle()
    fr = orc.rasterize(v, f, 16, 0.0, 1, clip_bary=True)
    k = fr["pix_to_face"][..., 0] >= 0
    b = fr["bary"][..., 0, :][k]
    assert np.allclose(b.sum(-1), 1, atol=1e-5) and (b >= 0).all()
    # OF renderer: no y flip => image is the vertical mirror of the textured/soft render of the same verts
    verts = np.array([[[-0.5, -0.6, 1.0], [0.5, -0.6, 1.0], [0.0, 0.3, 1.0]]], np.float32)
    of = orc.of_renderer(verts, f, 16)["pix_to_face"][0, ..., 0] >= 0
    ndc_flip = orc.view(verts, yflip=True)
    fl = orc.rasterize(ndc_flip, f, 16, 0.0, 1)["pix_to_face"][0, ..., 0] >= 0
    assert np.array_equal(of, fl[::-1]) and of.any()


def test_backward_vs_fp64_finite_differences():
    """d(sum w*mask)/d ndc from the restated backward == central differences of the fp64 forward."""
    rng = np.random.default_rng(0)
    v, f = util.icosphere(1)  # 42 v / 80 f
    N, S, K = 1, 24, 8
    X = (v[None] * [1.0, 0.6, 0.5]).astype(np.float64)
    cam = util.synth_cams(N, 3).astype(np.float64)
    faces = np.repeat(f[None], N, 0)
    sigma, blur = 1e-2, float(np.log(1 / 1e-4 - 1) * 1e-2)  # wide blur => smooth enough for FD
    w = rng.standard_normal((N, S, S))

    def fwd(ndc):
        fr = orc.rasterize(ndc, faces, S, blur, K, want_bary=False)
        return fr, orc.sigmoid_alpha_blend(fr["dists"], fr["pix_to_face"], sigma)

    ndc = orc.view(orc.project(X, cam, 0.0, np.float64))
    fr, m = fwd(ndc)
    gd = orc.sigmoid_alpha_blend_backward(fr["dists"], fr["pix_to_face"], w, sigma)
    g = orc.scatter_face_grads(orc.rasterize_backward(fr["face_verts"], fr["pix_to_face"], grad_dists=gd), faces, v.shape[0])
    assert np.abs(g[..., 2]).max() == 0
    h = 1e-6
    checked = 0
    for vi in rng.choice(v.shape[0], 12, replace=False):
        for c in range(2):
            a, b = ndc.copy(), ndc.copy()
            a[0, vi, c] += h; b[0, vi, c] -= h
            fa, ma = fwd(a); fb, mb = fwd(b)
            if not (np.array_equal(fa["pix_to_face"], fr["pix_to_face"]) and np.array_equal(fb["pix_to_face"], fr["pix_to_face"])):
                continue  # fragment set changed inside the stencil: the loss is discontinuous there
            fd = ((ma - mb) * w).sum() / (2 * h)
            assert abs(fd - g[0, vi, c]) <= 2e-4 * max(1.0, abs(fd)), (vi, c, fd, g[0, vi, c])
            checked += 1
    assert checked >= 8


def test_reprojection_losses_restatement_vs_reference_golden():
    """torch_ref.bds_loss / optical_flow_loss / laplacian_cot / hypothesis_weighting against values and fp64 gradients
    produced by the reference's own loss_utils / geom_utils / main.py lines (tests/golden/make_golden.py reproj)."""
    g = util.golden("reproj.npz")
    _, hf = util.template("horse")
    hf = torch.from_numpy(hf)
    # bds_loss
    N = g["bds_proj"].shape[0]
    faces = hf[None].repeat(N, 1, 1)
    bds = torch.from_numpy(g["bds_pts"]).repeat(N // g["bds_pts"].shape[0], 1, 1)
    p2f = torch.from_numpy(g["bds_p2f"]).long()
    sel = torch.from_numpy(g["bds_sel"])
    l = torch_ref.bds_loss(torch.from_numpy(g["bds_proj"]), bds, faces, p2f, sel)
    assert np.allclose(l.numpy(), g["bds_loss"], rtol=2e-5, atol=0)   # the reference's cdist goes through a matmul
    pd = torch.from_numpy(g["bds_proj"]).double().requires_grad_(True)
    (torch_ref.bds_loss(pd, bds.double(), faces, p2f, sel) * torch.from_numpy(g["bds_w"])).sum().backward()
    assert util.rel_err(pd.grad.numpy(), g["bds_grad"]) < 1e-6
    # optical_flow_loss: visibility render by the C oracle, as in the golden script
    M, cams, flows = (torch.from_numpy(g[k]) for k in ("of_meshes", "of_cams", "of_flows"))
    B, T, V, _ = M.shape
    faces_of = hf[None, None].repeat(B, T, 1, 1)
    proj = orc.project(M.reshape(B * T, V, 3).numpy(), cams.numpy(), 0.0)
    p2f_of = torch.from_numpy(orc.of_renderer(proj, faces_of.reshape(B * T, -1, 3).numpy(), img_size=flows.shape[2])["pix_to_face"])
    fl = flows.repeat(B, 1, 1, 1, 1)
    loss, pr, m, gt = torch_ref.optical_flow_loss(M, faces_of, cams, fl, p2f_of)
    assert np.array_equal(m.numpy(), g["of_vis"]) and m.sum() > 50
    assert np.allclose(loss.numpy(), g["of_loss"], rtol=1e-5, atol=0)
    assert np.allclose(pr.numpy(), g["of_pred"], rtol=0, atol=1e-5) and np.array_equal(gt.numpy(), g["of_samples"])
    Md, cd = M.double().requires_grad_(True), cams.double().requires_grad_(True)
    (torch_ref.optical_flow_loss(Md, faces_of, cd, fl.double(), p2f_of)[0] * torch.from_numpy(g["of_w"])).sum().backward()
    assert util.rel_err(Md.grad.numpy(), g["of_grad_meshes"]) < 1e-9
    assert util.rel_err(cd.grad.numpy(), g["of_grad_cams"]) < 1e-9
    # cotangent Laplacian, hypothesis weighting
    hv, _ = util.template("horse")
    assert util.rel_err(torch_ref.laplacian_cot(torch.from_numpy(hv), hf).numpy(), g["lap_cot"]) < 1e-5
    tl = torch.from_numpy(g["hyp_loss"]).requires_grad_(True)
    tot, probs = torch_ref.hypothesis_weighting(tl)
    tot.backward()
    assert np.allclose(tot.detach().numpy(), g["hyp_total"]) and np.allclose(probs.numpy(), g["hyp_probs"])
    assert np.allclose(tl.grad.numpy(), g["hyp_grad"])


def test_uniform_laplacian_closed_form():
    v, f = util.icosphere(1)
    L = torch_ref.laplacian_uniform(v.shape[0], torch.from_numpy(f))
    assert torch.allclose(L.sum(1), torch.zeros(v.shape[0]), atol=1e-6)          # rows: -1 + deg * 1/deg
    assert set(np.unique((L > 0).sum(1).numpy())) <= {5, 6}                         # icosphere vertex degrees
    assert torch.equal(torch.diagonal(L), -torch.ones(v.shape[0]))


def test_target_maps_restatement_vs_reference_golden():
    """oracle/targets_ref.py against utils/image.py outputs (tests/golden/make_golden.py targets); the EDT additionally
    against exhaustive search, and the fp64-sqrt-then-fp32 rounding the CUDA kernel relies on against correctly rounded
    fp32 square roots of every squared distance a 1024^2 map can produce."""
    from oracle import targets_ref as tr
    g = util.golden("targets.npz")
    for sfx, k in (("", 50), ("_b", 20)):
        m = g["masks" + sfx]
        assert np.array_equal(np.stack([tr.compute_dt(x, norm=False) for x in m]), g["dt_raw" + sfx])
        assert np.array_equal(np.stack([tr.compute_dt(x) for x in m]), g["dt_norm" + sfx])
        assert np.array_equal(np.stack([tr.compute_dt_barrier(x, k=k) for x in m]), g["barrier" + sfx])
        assert np.array_equal(tr.compute_boundaries(m), g["boundaries" + sfx])
    m = g["masks"][0]
    assert np.array_equal(tr.edt_bruteforce(m == 1), np.rint(g["dt_raw"][0] ** 2).astype(np.int64))
    n = np.arange(0, 2 * 1024 * 1024 + 1, dtype=np.float64)
    assert np.array_equal(np.sqrt(n).astype(np.float32), np.sqrt(n.astype(np.float32)))


def test_prior_losses_restatement_vs_reference_golden():
    g = util.golden("priors.npz")
    hv, hf = util.template("horse")
    hv, hf = torch.from_numpy(hv), torch.from_numpy(hf)
    X = torch.from_numpy(g["X"])
    assert np.allclose(torch_ref.locally_rigid(X, hv, hf).numpy(), g["rigid"], rtol=1e-5)
    assert np.allclose(torch_ref.laplacian_smoothing_cot(X, hf).numpy(), g["smooth"], rtol=1e-5)
    Xd, td = X.double().requires_grad_(True), hv.double().requires_grad_(True)
    torch_ref.locally_rigid(Xd, td, hf).backward()
    assert util.rel_err(Xd.grad.numpy(), g["rigid_grad_X"]) < 1e-9 and util.rel_err(td.grad.numpy(), g["rigid_grad_t"]) < 1e-9
    Xd = X.double().requires_grad_(True)
    torch_ref.laplacian_smoothing_cot(Xd, hf).backward()
    assert util.rel_err(Xd.grad.numpy(), g["smooth_grad_X"]) < 1e-4   # the constant weights are fp32 on both sides (summation order differs)


def test_correlation_restatement_closed_form():
    """oracle/correlation_ref.py on cases with a known answer (CPU): output shape formula of correlation_cuda.cc:24-32, the
    centre displacement of identical inputs is the per-pixel mean square, a shifted copy peaks at its shift, zero padding."""
    import torch
    from oracle import correlation_ref as cref
    assert cref.out_shape(64, 64, 4, 1, 4, 1, 1) == (81, 64, 64)
    assert cref.out_shape(24, 24, 20, 1, 20, 2, 2) == (441, 12, 12)
    assert cref.out_shape(18, 22, 3, 3, 2, 1, 1) == (25, 18, 22)
    x = torch.randn(2, 5, 10, 12, generator=torch.Generator().manual_seed(0), dtype=torch.float64)
    c = cref.correlation(x, x, 4, 1, 4, 1, 1)
    assert c.shape == (2, 81, 10, 12)
    assert torch.allclose(c[:, 40], (x * x).mean(1))
    assert torch.allclose(c[:, 0, :4, :], torch.zeros(2, 4, 12, dtype=torch.float64))        # displacement (-4,-4): rows 0..3 read padding
    y = torch.roll(x, shifts=(1, 2), dims=(2, 3))
    c = cref.correlation(x, y, 4, 1, 4, 1, 1)[:, :, 4:6, 4:8].mean(dim=(0, 2, 3))
    assert int(c.argmax()) == (1 + 4) * 9 + (2 + 4)


@pytest.mark.parametrize("tag", ["plain", "affine", "mirror"])
def test_camera_assembly_restatement_vs_reference_lines(tag):
    """oracle/torch_ref.assemble_cameras against cameras.npz, which make_golden.py produced by executing the reference's own
    mirror_cameras / transform_cameras / assembly lines (multiframe/main.py:113-138, :573-582)."""
    from oracle import torch_ref
    g = util.golden("cameras.npz")
    out = torch_ref.assemble_cameras(torch.from_numpy(g["raw"]).double(), torch.from_numpy(g[f"{tag}_mirror"]).double(),
                                     torch.from_numpy(g[f"{tag}_transforms"]).double(), float(g["scale_lr_decay"]))
    assert np.abs(out.numpy() - g[f"{tag}_cam_pred64"]).max() < 1e-12
