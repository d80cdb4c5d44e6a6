"""GPU parity of the rasterizer path (C ABI -> NeuralRenderer / OF_NeuralRenderer mirrors) against the
CPU oracle on identical inputs.

Bars (BASELINE.json north_star): pix_to_face bit-exact away from depth ties — here the kernel uses the
oracle's (z, face) order and strict IEEE arithmetic, so we demand full equality of pix_to_face, zbuf and
dists; silhouettes within 1e-5 of the unit range (absolute; a pure relative bound is unreachable for
1-prod(1-p) near 0 in fp32) and per-render sums within 1e-5 relative; gradients within
1e-3 relative (max-norm)."""
import numpy as np
import pytest
import torch

from oracle import pt3d_oracle as orc
from oracle import torch_ref
from tests import util

pytestmark = pytest.mark.gpu


def _gpu_mask_render(X, faces, cam, S, offset_z, faces_dtype=torch.int64, shared=False, K=20):
    from acfm_video_3d_reconstruction_b200 import NeuralRenderer
    from acfm_video_3d_reconstruction_b200 import functional as F_
    r = NeuralRenderer(S, offset_z=offset_z)
    r.faces_per_pixel = K
    Xc, cc = torch.from_numpy(X).cuda(), torch.from_numpy(cam).cuda()
    fc = torch.from_numpy(faces).to(faces_dtype).cuda()
    if shared:
        fc = fc[:1].expand(X.shape[0], -1, -1)
    ndc = r.to_ndc(Xc, cc)
    mask, p2f, zbuf, dists = F_.soft_silhouette(ndc, fc, S, r.blur_radius, K, r.sigma)
    return dict(ndc=ndc, mask=mask, pix_to_face=p2f, zbuf=zbuf, dists=dists)


def _assert_fragments_equal(gpu, ref):
    p_g, p_r = gpu["pix_to_face"].cpu().numpy(), ref["pix_to_face"]
    assert p_g.dtype == np.int64
    neq = (p_g != p_r)
    assert not neq.any(), f"pix_to_face differs at {neq.sum()} of {neq.size} entries"
    assert np.array_equal(gpu["zbuf"].cpu().numpy(), ref["zbuf"])
    assert np.array_equal(gpu["dists"].cpu().numpy(), ref["dists"])
    if "mask" in ref and gpu.get("mask") is not None:
        m_g, m_r = gpu["mask"].cpu().numpy(), ref["mask"]
        assert np.abs(m_g - m_r).max() <= 1e-5
        s_g, s_r = m_g.reshape(len(m_g), -1).sum(1), m_r.reshape(len(m_r), -1).sum(1)
        assert np.all(np.abs(s_g - s_r) <= 1e-5 * np.maximum(s_r, 1e-30))


def test_raster_epsilon_setting_vs_oracle_variant():
    """kEpsilon is a run-time setting of the rasterizer (acfm_set_raster_epsilon): with 1e-30, the value of early PyTorch3D
    releases, the kernel reproduces the oracle's 1e-30 variant bit for bit — and the two settings do differ
    (tests/test_oracle_variants.py says where)."""
    from acfm_video_3d_reconstruction_b200 import functional as F_
    v, f = util.template("bird")
    N, S = 3, 128
    X, cam = util.synth_verts(v, N, seed=51), util.synth_cams(N, seed=52)
    faces = np.repeat(f[None], N, 0)
    base = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=5.0)
    old = F_.set_raster_epsilon(1e-30)
    try:
        assert old == np.float32(1e-8)
        orc.set_variant(k_eps=1e-30)
        ref = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=5.0)
        out = _gpu_mask_render(X, faces, cam, S, 5.0)
    finally:
        orc.set_variant()
        F_.set_raster_epsilon(old)
    _assert_fragments_equal(out, ref)
    assert (ref["pix_to_face"] != base["pix_to_face"]).any()
    _assert_fragments_equal(_gpu_mask_render(X, faces, cam, S, 5.0), base)   # and back


def test_golden_small():
    g = util.golden("raster_small.npz")
    N = g["X"].shape[0]
    faces = np.repeat(g["faces"][None], N, 0)
    out = _gpu_mask_render(g["X"], faces, g["cam"], 64, 5.0)
    assert np.array_equal(out["ndc"].cpu().numpy(), g["ndc"])
    ref = dict(pix_to_face=g["pix_to_face"].astype(np.int64), zbuf=g["zbuf"], dists=g["dists"], mask=g["mask"])
    _assert_fragments_equal(out, ref)


@pytest.mark.parametrize("name,S,offset_z,fdt,shared", [
    ("bird", 256, 5.0, torch.int64, False),   # monocular config (C1/C2 shapes)
    ("horse", 256, 0.0, torch.int64, True),   # multiframe config, shared topology (stride-0 faces)
    ("bird", 128, 5.0, torch.int32, False),   # int32 faces (bird_vis.py caller)
    ("horse", 100, 0.0, torch.int64, False),  # size not a multiple of the tile/region
])
def test_soft_silhouette_vs_oracle(name, S, offset_z, fdt, shared):
    v, f = util.template(name)
    N = 4
    X = util.synth_verts(v, N, seed=S)
    cam = util.synth_cams(N, seed=S + 1)
    faces = np.repeat(f[None], N, 0)
    ref = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=offset_z)
    out = _gpu_mask_render(X, faces, cam, S, offset_z, fdt, shared)
    assert np.array_equal(out["ndc"].cpu().numpy(), ref["ndc"])
    _assert_fragments_equal(out, ref)
    cov = (ref["pix_to_face"][..., 0] >= 0).mean()
    full = (ref["pix_to_face"][..., -1] >= 0).mean()
    assert cov > 0.03 and full > 0.005, "test must exercise the K-truncation"


def test_highres_k50_vs_oracle():
    """C4 shape: 2562 v / 5120 f, 512^2, K = 50 (one render; the oracle is O(pixels x faces))."""
    v, f = util.icosphere(4)
    v = (v * np.array([0.8, 0.48, 0.4], np.float32)).astype(np.float32)
    X, cam, faces = v[None], util.synth_cams(1, seed=11), f[None]
    ref = orc.neural_renderer_mask(X, faces, cam, img_size=512, offset_z=0.0, K=50)
    out = _gpu_mask_render(X, faces, cam, 512, 0.0, K=50)
    _assert_fragments_equal(out, ref)


def test_hard_raster_and_of_renderer_vs_oracle():
    from acfm_video_3d_reconstruction_b200 import OF_NeuralRenderer
    from acfm_video_3d_reconstruction_b200 import functional as F_
    v, f = util.template("horse")
    N = 3
    X, cam = util.synth_verts(v, N, seed=3), util.synth_cams(N, seed=4)
    faces = np.repeat(f[None], N, 0)
    # texture-branch raster: blur 0, K = 1, clipped barycentrics
    ref = orc.hard_raster(X, faces, cam, img_size=128, offset_z=0.0)
    fr = F_.rasterize(torch.from_numpy(ref["ndc"]).cuda(), torch.from_numpy(faces).cuda(), 128, 0.0, 1,
                      clip_barycentric_coords=True, want_bary=True)
    _assert_fragments_equal(fr, ref)
    assert np.array_equal(fr["bary"].cpu().numpy(), ref["bary"])
    # OF_NeuralRenderer: already-projected verts, no y flip
    proj = orc.project(X, cam, 0.0)
    ref_of = orc.of_renderer(proj, faces, img_size=128)
    p2f = OF_NeuralRenderer(128)(torch.from_numpy(proj).cuda(), torch.from_numpy(faces).cuda())
    assert p2f.shape == (N, 128, 128, 1) and np.array_equal(p2f.cpu().numpy(), ref_of["pix_to_face"])


def test_edge_cases():
    from acfm_video_3d_reconstruction_b200 import functional as F_
    v, f = util.icosphere(1)
    v = v * 0.5
    # (a) mesh completely off screen, (b) behind the camera (z < 0), (c) degenerate faces, (d) duplicated depth (ties)
    off = v + np.array([5.0, 0, 2.0], np.float32)
    behind = v + np.array([0, 0, -3.0], np.float32)
    ok = v + np.array([0, 0, 2.0], np.float32)
    ndc = np.stack([off, behind, ok, ok]).astype(np.float32)
    faces = np.repeat(f[None], 4, 0).copy()
    faces[2, :10] = faces[2, :10, :1]            # zero-area faces
    faces[3, 40:] = faces[3, :40]                # duplicate faces => exact z ties, broken by face id
    ref = orc.rasterize(ndc, faces, 64, F_.BLUR_SOFT * 10, 6, want_bary=False)
    ref["mask"] = orc.sigmoid_alpha_blend(ref["dists"], ref["pix_to_face"], F_.SIGMA)
    fr = F_.rasterize(torch.from_numpy(ndc).cuda(), torch.from_numpy(faces).cuda(), 64, F_.BLUR_SOFT * 10, 6,
                      sigma=F_.SIGMA, want_mask=True)
    _assert_fragments_equal(fr, ref)
    assert (ref["pix_to_face"][:2] == -1).all() and (ref["pix_to_face"][2:, ..., 0] >= 0).any()
    # (e) empty batch
    e = F_.rasterize(torch.zeros(0, 12, 3).cuda(), torch.zeros(0, 20, 3, dtype=torch.int64).cuda(), 32, 0.0, 2)
    assert e["pix_to_face"].shape == (0, 32, 32, 2)
    # (f) K out of range -> ValueError like the reference's PyTorch3D checks
    with pytest.raises(ValueError):
        F_.rasterize(torch.from_numpy(ndc).cuda(), torch.from_numpy(faces).cuda(), 64, 0.0, 65)
    # (g) cull_backfaces
    refc = orc.rasterize(ndc[2:], faces[2:], 64, 0.0, 2, cull_backfaces=True, want_bary=False)
    frc = F_.rasterize(torch.from_numpy(ndc[2:]).cuda(), torch.from_numpy(faces[2:]).cuda(), 64, 0.0, 2, cull_backfaces=True)
    _assert_fragments_equal(frc, refc)


@pytest.mark.parametrize("upstream", ["ones", "normal"])
def test_backward_second_pass_of_the_fixed_point_accumulation(upstream):
    """The backward adds every contribution to ONE int32 per vertex component and verifies the headroom afterwards: a region
    whose summed magnitudes reach 2^bits is accumulated again with a smaller scale.  With bits lowered from 30 to 20 nearly
    every region of an ordinary render takes the second pass: same gradient (to the coarser quantum), still the oracle's."""
    from acfm_video_3d_reconstruction_b200 import _lib
    from acfm_video_3d_reconstruction_b200 import functional as F_
    v, f = util.template("bird")
    N, S = 3, 128
    X, cam = util.synth_verts(v, N, seed=61), util.synth_cams(N, seed=62)
    faces = np.repeat(f[None], N, 0)
    ref = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=5.0)
    gm = np.ones((N, S, S), np.float32) if upstream == "ones" else np.random.default_rng(3).standard_normal((N, S, S)).astype(np.float32)
    g_ref = orc.neural_renderer_mask_backward(ref, faces, gm)
    ndc0 = torch.from_numpy(ref["ndc"]).cuda()

    def grad():
        x = ndc0.clone().requires_grad_(True)
        m, _, _, _ = F_.soft_silhouette(x, torch.from_numpy(faces).cuda(), S)
        return torch.autograd.grad((m * torch.from_numpy(gm).cuda()).sum(), x)[0].cpu().numpy()

    g30 = grad()
    assert util.rel_err(g30, g_ref) < 1e-3
    try:
        _lib.check(_lib.lib().acfm_set_raster_bwd_headroom_bits(20), "acfm_set_raster_bwd_headroom_bits")
        g24 = grad()
    finally:
        _lib.check(_lib.lib().acfm_set_raster_bwd_headroom_bits(30), "acfm_set_raster_bwd_headroom_bits")
    assert not np.array_equal(g24, g30)                       # the second pass ran (another quantum)
    assert util.rel_err(g24, g30) < 3e-4 and util.rel_err(g24, g_ref) < 1e-3
    with pytest.raises(ValueError):
        _lib.check(_lib.lib().acfm_set_raster_bwd_headroom_bits(31), "acfm_set_raster_bwd_headroom_bits")


def test_backward_vs_oracle():
    """grad wrt screen vertices, then end to end through NeuralRenderer to vertices and cameras."""
    from acfm_video_3d_reconstruction_b200 import NeuralRenderer
    v, f = util.template("bird")
    N, S = 3, 128
    X, cam = util.synth_verts(v, N, seed=21), util.synth_cams(N, seed=22)
    faces = np.repeat(f[None], N, 0)
    ref = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=5.0)
    rng = np.random.default_rng(1)
    gm = rng.standard_normal((N, S, S)).astype(np.float32)
    g_ndc_ref = orc.neural_renderer_mask_backward(ref, faces, gm)

    r = NeuralRenderer(S, offset_z=5.0)
    Xc = torch.from_numpy(X).cuda().requires_grad_(True)
    cc = torch.from_numpy(cam).cuda().requires_grad_(True)
    ndc = r.to_ndc(Xc, cc)
    ndc.retain_grad()
    mask, p2f = r(Xc, torch.from_numpy(faces).cuda(), cc)  # public API
    assert np.array_equal(p2f.cpu().numpy(), ref["pix_to_face"])
    from acfm_video_3d_reconstruction_b200 import functional as F_
    m2, _, _, _ = F_.soft_silhouette(ndc, torch.from_numpy(faces).cuda(), S)
    (m2 * torch.from_numpy(gm).cuda()).sum().backward()
    assert util.rel_err(ndc.grad.cpu().numpy(), g_ndc_ref) < 1e-3
    gX1, gc1 = Xc.grad.clone(), cc.grad.clone()
    Xc.grad = None; cc.grad = None
    (mask * torch.from_numpy(gm).cuda()).sum().backward()
    assert torch.equal(gX1 != 0, Xc.grad != 0)
    # chain the oracle's d/d ndc through the (pinned) fp64 projection restatement
    Xd = torch.from_numpy(X).double().requires_grad_(True)
    cd = torch.from_numpy(cam).double().requires_grad_(True)
    (torch_ref.to_ndc(Xd, cd, 5.0) * torch.from_numpy(g_ndc_ref).double()).sum().backward()
    assert util.rel_err(Xc.grad.cpu().numpy(), Xd.grad.numpy()) < 1e-3
    assert util.rel_err(cc.grad.cpu().numpy(), cd.grad.numpy()) < 1e-3


def test_many_random_poses_partly_off_screen():
    """24 views with scales 0.15-1.3 and translations up to +-0.7 (partly off screen, tiny and huge on screen): exercises the
    conservative edge/bbox culling at every scale and the record-overflow path (tiny meshes put all faces in one region)."""
    v, f = util.template("horse")
    N, S = 24, 96
    gen = torch.Generator().manual_seed(77)
    cam = util.synth_cams(N, seed=5)
    cam[:, 0] = (torch.rand(N, generator=gen) * 1.15 + 0.15).numpy()
    cam[:, 1:3] = (torch.rand(N, 2, generator=gen) * 1.4 - 0.7).numpy()
    X = util.synth_verts(v, N, seed=9)
    faces = np.repeat(f[None], N, 0)
    ref = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=0.0)
    out = _gpu_mask_render(X, faces, cam, S, 0.0)
    _assert_fragments_equal(out, ref)
    assert (ref["pix_to_face"][..., -1] >= 0).any(axis=(1, 2)).sum() >= N // 2


def _launch_threads(N, V, F, S, K):
    import ctypes
    from acfm_video_3d_reconstruction_b200 import _lib
    th = ctypes.c_int(0)
    _lib.check(_lib.lib().acfm_raster_fwd_launch_info(N, V, F, S, S, K, None, None, ctypes.byref(th)), "launch_info")
    return th.value


@pytest.mark.parametrize("K", [24, 50, 51, 60, 63, 64])
def test_generic_k(K):
    """K = 20 runs the straight-line specialisation of the per-pixel sets; every other K the loop version.  K = 51 / 60 / 63
    are values for which a 16-bit magic division of the scalar output pass was wrong (round-1 advisor finding): every slot
    of every pixel must be written."""
    v, f = util.template("bird")
    N, S = 2, 128
    assert _launch_threads(N, v.shape[0], f.shape[0], S, K) in (128, 256)
    X, cam = util.synth_verts(v, N, seed=31), util.synth_cams(N, seed=32)
    cam[:, 0] *= 0.6   # smaller on screen: more faces per pixel, the deep lists fill up
    faces = np.repeat(f[None], N, 0)
    ref = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=5.0, K=K)
    out = _gpu_mask_render(X, faces, cam, S, 5.0, K=K)
    _assert_fragments_equal(out, ref)
    assert (ref["pix_to_face"][..., -1] >= 0).any()


def test_four_warp_ctas_on_a_large_mesh():
    """A mesh with 8192 faces at K = 64 leaves no room for eight warps' sets beside the region face list: the library falls
    back to 4-warp CTAs (the only shape besides 8 warps that is built).  Same results."""
    n = 64
    gy, gx = np.meshgrid(np.linspace(-0.8, 0.8, n + 1), np.linspace(-0.8, 0.8, n + 1), indexing="ij")
    rng = np.random.default_rng(3)
    z = 3.0 + 0.3 * np.sin(3 * gx) * np.cos(2 * gy) + 0.01 * rng.standard_normal(gx.shape)
    ndc = np.stack([gx, gy, z], -1).reshape(1, -1, 3).astype(np.float32)
    ndc[..., :2] += (0.004 * rng.standard_normal(ndc[..., :2].shape)).astype(np.float32)
    idx = np.arange((n + 1) * (n + 1)).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    faces = np.concatenate([np.stack([a, b, c], 1), np.stack([b, d, c], 1)])[None].astype(np.int64)
    S, K, blur = 64, 64, 1.5e-2
    assert faces.shape[1] == 8192 and _launch_threads(1, ndc.shape[1], 8192, S, K) == 128
    from acfm_video_3d_reconstruction_b200 import functional as F_
    ref = orc.rasterize(ndc, faces, S, blur, K, want_bary=False)
    mask, p2f, zb, di = F_.soft_silhouette(torch.from_numpy(ndc).cuda(), torch.from_numpy(faces).cuda(), S, blur, K, 1e-3)
    _assert_fragments_equal(dict(pix_to_face=p2f, zbuf=zb, dists=di), ref)
    assert (ref["pix_to_face"][..., -1] >= 0).mean() > 0.05   # the 64-deep sets do fill up


def test_misaligned_outputs_take_the_scalar_path():
    """Output tensors that are not 16-byte aligned (a caller's view into a larger buffer) go through the scalar output pass."""
    from acfm_video_3d_reconstruction_b200 import _lib
    from acfm_video_3d_reconstruction_b200 import functional as F_
    v, f = util.template("bird")
    N, S, K = 2, 64, 20
    X, cam = util.synth_verts(v, N, seed=41), util.synth_cams(N, seed=42)
    faces = np.repeat(f[None], N, 0)
    ref = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=5.0, K=K)
    ndc = torch.from_numpy(ref["ndc"]).cuda()
    fc = torch.from_numpy(faces).cuda()
    n = N * S * S * K
    p2f = torch.full((n + 1,), 7, dtype=torch.int64, device="cuda")[1:]          # 8-byte aligned only
    zb = torch.full((n + 1,), 7.0, dtype=torch.float32, device="cuda")[1:]       # 4-byte aligned only
    di = torch.full((n + 1,), 7.0, dtype=torch.float32, device="cuda")[1:]
    mask = torch.empty((N, S, S), dtype=torch.float32, device="cuda")
    assert p2f.data_ptr() % 16 != 0 and zb.data_ptr() % 16 != 0
    st = _lib.lib().acfm_raster_fwd(_lib.ptr(ndc), _lib.ptr(fc), 1, fc.shape[1] * 3, N, ndc.shape[1], fc.shape[1], S, S, K,
                                    float(F_.BLUR_SOFT), 0, 0, float(F_.SIGMA), _lib.ptr(p2f), _lib.ptr(zb), _lib.ptr(di), None,
                                    _lib.ptr(mask), None, None, 0, _lib.stream_of(ndc))
    _lib.check(st, "acfm_raster_fwd")
    out = dict(pix_to_face=p2f.view(N, S, S, K), zbuf=zb.view(N, S, S, K), dists=di.view(N, S, S, K), mask=mask)
    _assert_fragments_equal(out, ref)


@pytest.mark.parametrize("scale,blur,K", [(0.05, 1e-3, 3), (0.5, 0.0, 8), (0.5, 0.05, 8), (5.0, 1e-3, 4), (500.0, 1e-2, 6), (0.5, 1e-3, 64)])
def test_triangle_soups_forward_and_backward(scale, blur, K):
    """Random triangle soups — overlapping, sliver, off-screen, behind-the-camera and screen-filling faces (coordinates up to
    ~1e3 NDC): the conservative culling margins of the forward and the out-of-range (global atomic) branch of the fixed-point
    backward must hold for arbitrary inputs, not just for the templates."""
    from acfm_video_3d_reconstruction_b200 import functional as F_
    gen = torch.Generator().manual_seed(int(scale * 1000) + K)
    N, F, S = 5, 40, 48
    V = 3 * F
    xy = scale * torch.randn(N, V, 2, generator=gen)
    xy[:, ::7] *= 0.02                                   # some slivers / tiny faces
    z = torch.rand(N, V, 1, generator=gen) * 2.5 + 0.5
    z[:, ::11] -= 2.0                                    # some vertices behind the camera
    ndc = torch.cat([xy, z], -1).numpy().astype(np.float32)
    faces = np.repeat(np.arange(V, dtype=np.int64).reshape(1, F, 3), N, 0)
    ref = orc.rasterize(ndc, faces, S, blur, K, want_bary=False)
    nd = torch.from_numpy(ndc).cuda().requires_grad_(True)
    if blur > 0:
        mask, p2f, zb, d = F_.soft_silhouette(nd, torch.from_numpy(faces).cuda(), S, blur, K, 1e-2)
        out = dict(pix_to_face=p2f, zbuf=zb, dists=d)
    else:
        out = F_.rasterize(nd.detach(), torch.from_numpy(faces).cuda(), S, blur, K)
    _assert_fragments_equal(out, ref)
    assert (ref["pix_to_face"] >= 0).mean() > 0.01
    if blur > 0:
        ref["ndc"] = ndc
        gm = torch.randn(N, S, S, generator=gen).numpy().astype(np.float32)
        gd = orc.sigmoid_alpha_blend_backward(ref["dists"], ref["pix_to_face"], gm, 1e-2)
        g_ref = orc.scatter_face_grads(orc.rasterize_backward(ref["face_verts"], ref["pix_to_face"], grad_dists=gd), faces, V)
        (mask * torch.from_numpy(gm).cuda()).sum().backward()
        assert util.rel_err(nd.grad.cpu().numpy(), g_ref) < 1e-3


def test_torch_dense_standin_matches_oracle():
    """The pure-PyTorch dense renderer bench.py times as the "PyTorch3D CUDA path" stand-in (BASELINE.md B-GPU-torch) computes
    the same silhouettes and gradients as the C oracle (fragments may differ at ties: it is not bit-exact by construction)."""
    from oracle import torch_dense
    v, f = util.template("bird")
    N, S = 2, 64
    X, cam = util.synth_verts(v, N, seed=41), util.synth_cams(N, seed=42)
    faces = np.repeat(f[None], N, 0)
    ref = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=5.0)
    gm = np.random.default_rng(3).standard_normal((N, S, S)).astype(np.float32)
    g_ref = orc.neural_renderer_mask_backward(ref, faces, gm)
    nd = torch.from_numpy(ref["ndc"]).cuda().requires_grad_(True)
    mask, p2f = torch_dense.soft_silhouette(nd, torch.from_numpy(f).cuda(), S, orc.BLUR_SOFT, orc.K_SOFT, orc.SIGMA)
    assert np.abs(mask.detach().cpu().numpy() - ref["mask"]).max() < 1e-4
    assert (p2f.cpu().numpy() != ref["pix_to_face"]).mean() < 1e-3
    (mask * torch.from_numpy(gm).cuda()).sum().backward()
    assert util.rel_err(nd.grad.cpu().numpy(), g_ref) < 1e-3


@pytest.mark.parametrize("S,K,hard", [(256, 20, False), (100, 5, False), (90, 3, False), (256, 1, True), (72, 1, True), (512, 50, False)])
def test_split_path_equals_single_kernel(S, K, hard, monkeypatch):
    """acfm_raster_fwd with a workspace (region classification + rasterizer on the live regions + concurrent TMA padding
    kernel, runs of empty regions) writes byte for byte what the single-kernel path writes: image sizes that are not
    multiples of the region (and, at S = 90 / K = 3, rows that are not 16-byte multiples, which take the plain-store
    fallback of the fill kernel), K = 1 with barycentrics, the largest K of the reference's configs; one render has no
    vertex on screen (all regions empty), one fills the image (no empty region)."""
    from acfm_video_3d_reconstruction_b200 import functional as F_
    v, f = util.template("bird")
    N = 6
    X, cam = util.synth_verts(v, N, seed=5), util.synth_cams(N, seed=6)
    cam[1, 1:3] = (3.0, -3.0)      # off screen
    cam[2, 0] = 3.0                # larger than the image
    cam[3, 1:3] = (0.8, 0.0)       # half outside
    ndc = F_.project(torch.from_numpy(X).cuda(), torch.from_numpy(cam).cuda(), 5.0, -1.0, -1.0, F_.EYE_Z)
    faces = torch.from_numpy(f)[None].cuda()

    def render():
        if hard:
            return F_.rasterize(ndc, faces, S, 0.0, 1, clip_barycentric_coords=True, want_bary=True)
        return F_.rasterize(ndc, faces, S, F_.BLUR_SOFT, K, sigma=F_.SIGMA, want_mask=True)

    monkeypatch.setattr(F_, "SPLIT_FILL", True)
    a = render()
    monkeypatch.setattr(F_, "SPLIT_FILL", False)
    b = render()
    torch.cuda.synchronize()
    for key in ("pix_to_face", "zbuf", "dists", "bary", "mask"):
        if a[key] is None:
            assert b[key] is None
            continue
        assert torch.equal(a[key], b[key]), key
    assert (a["pix_to_face"][1] == -1).all() and (a["pix_to_face"][0] >= 0).any()


@pytest.mark.parametrize("split", [True, False])
def test_fused_visible_vertices_equal_the_pix_to_face_route(split, monkeypatch):
    """The (N,V) visible-vertex map written by the render (vertices of every pixel's nearest face) equals what
    acfm_visible_verts derives from pix_to_face[..., 0] — the reference's fi_maps/unique/scatter_ block — for the soft
    K = 20 render (NeuralRenderer.forward_with_visibility) and the hard K = 1 render (OF_NeuralRenderer); the map is
    returned explicitly and handed to the losses as `visible=` (same loss either way)."""
    from acfm_video_3d_reconstruction_b200 import NeuralRenderer, OF_NeuralRenderer, loss_utils
    from acfm_video_3d_reconstruction_b200 import functional as F_
    monkeypatch.setattr(F_, "SPLIT_FILL", split)
    v, f = util.template("horse")
    N = 5
    X, cam = util.synth_verts(v, N, seed=11), util.synth_cams(N, seed=12)
    cam[1, 1:3] = (3.0, 3.0)   # off screen: nothing visible
    Xc, cc = torch.from_numpy(X).cuda(), torch.from_numpy(cam).cuda()
    faces = torch.from_numpy(f)[None].cuda().expand(N, -1, -1)
    r = NeuralRenderer(128, offset_z=5.0)
    mask, p2f, fused = r.forward_with_visibility(Xc, faces, cc)
    m0, p0 = r(Xc, faces, cc)
    assert torch.equal(mask, m0) and torch.equal(p2f, p0)
    plain = loss_utils.visible_vertices(p2f, faces, v.shape[0])
    assert torch.equal(fused, plain) and fused[1].sum() == 0 and 0 < fused[0].sum() < v.shape[0]
    proj = F_.project(Xc, cc, 5.0)
    ofr = OF_NeuralRenderer(128)
    p1, vis1 = ofr.forward_with_visibility(proj, faces)
    assert torch.equal(p1, ofr(proj, faces))
    assert torch.equal(vis1, loss_utils.visible_vertices(p1, faces, v.shape[0]))
    gen = torch.Generator().manual_seed(3)
    bds = torch.cat([torch.rand(N, 50, 2, generator=gen) * 1.6 - 0.8, torch.ones(N, 50, 1)], -1).cuda()
    sel = torch.arange(50).cuda()
    a = loss_utils.bds_loss(proj, bds, faces, p2f, reduce=False, indices=sel, visible=fused)
    b = loss_utils.bds_loss(proj, bds, faces, p2f, reduce=False, indices=sel)
    assert torch.equal(a, b)


def test_rasterize_of_and_tree_defaults():
    """NeuralRenderer.rasterize_of (multiframe/nnutils/nmr.py:131-141): the hard K = 1 render under an explicit (R, T) view,
    returned as Fragments — equals the oracle on the transformed vertices.  The two trees' modules differ in offset_z only."""
    from acfm_video_3d_reconstruction_b200 import monocular, multiframe
    assert monocular.NeuralRenderer(64).offset_z == 5.0 and multiframe.NeuralRenderer(64).offset_z == 0.0
    v, f = util.template("bird")
    N, S = 2, 64
    X = util.synth_verts(v, N, seed=61) * 0.6
    faces = np.repeat(f[None], N, 0)
    R = np.repeat(np.diag([-1.0, 1.0, 1.0]).astype(np.float32)[None], N, 0)
    T = np.repeat(np.array([[0.0, 0.0, 2.732]], np.float32), N, 0)
    fr = multiframe.NeuralRenderer(S).rasterize_of(torch.from_numpy(X).cuda(), torch.from_numpy(faces).cuda(),
                                                   torch.from_numpy(R).cuda(), torch.from_numpy(T).cuda())
    view = (X @ R[0] + T[:, None, :]).astype(np.float32)
    ref = orc.rasterize(view, faces, S, 0.0, 1, clip_bary=False, want_bary=True)
    assert np.array_equal(fr.pix_to_face.cpu().numpy(), ref["pix_to_face"])
    assert np.array_equal(fr.zbuf.cpu().numpy(), ref["zbuf"]) and np.array_equal(fr.dists.cpu().numpy(), ref["dists"])
    assert np.array_equal(fr.bary_coords.cpu().numpy(), ref["bary"])
    assert (ref["pix_to_face"] >= 0).mean() > 0.05


@pytest.mark.parametrize("with_edt,NBdiv", [(True, 1), (True, 4), (False, 2)])
def test_fused_mask_losses_equal_the_two_pass_route(with_edt, NBdiv):
    """acfm_raster_fwd_train / acfm_raster_soft_bwd_train (the mask-loss sums accumulated in the render's epilogue, their
    backward formed inside the rasterizer backward) against the unfused route (render, then acfm_mask_sums_fwd / _bwd on the
    mask): same mask and fragments, sums within 1e-5 relative, vertex / camera gradients within 1e-4; an extra explicit
    gradient on the mask adds up with the one from the sums; the forward is bit-identical from call to call (ordered reductions)."""
    from acfm_video_3d_reconstruction_b200 import NeuralRenderer, loss_utils
    v, f = util.template("bird")
    N, S = 8, 128
    NB = N // NBdiv
    X, cam = util.synth_verts(v, N, seed=71), util.synth_cams(N, seed=72)
    cam[3, 1:3] = (3.0, 3.0)     # one render off screen: its sums are those of the bare target
    gen = torch.Generator().manual_seed(5)
    tgt = (torch.rand(NB, S, S, generator=gen) > 0.6).float().cuda()
    edt = (torch.rand(NB, S, S, generator=gen) * 3).cuda() if with_edt else None
    wsum = torch.randn(N, 4, generator=gen).cuda()
    wmask = (0.01 * torch.randn(N, S, S, generator=gen)).cuda()
    faces = torch.from_numpy(f)[None].cuda().expand(N, -1, -1)
    r = NeuralRenderer(S, offset_z=5.0)
    res = []
    for fused in (True, True, False):
        Xc = torch.from_numpy(X).cuda().requires_grad_(True)
        cc = torch.from_numpy(cam).cuda().requires_grad_(True)
        if fused:
            mask, p2f, sums = r.forward_with_losses(Xc, faces, cc, tgt, edt)
        else:
            mask, p2f = r(Xc, faces, cc)
            sums = loss_utils.mask_sums(mask, tgt, edt)
        ((sums * wsum).sum() + (mask * wmask).sum()).backward()
        res.append((mask.detach(), p2f, sums.detach(), Xc.grad.clone(), cc.grad.clone()))
    for a, b in zip(res[0][:3], res[1][:3]):   # mask, fragments, sums: ordered reductions, bit-identical from call to call
        assert torch.equal(a, b)               # (the gradients go through float atomics across CTAs: 1e-6, not bits)
    assert util.rel_err(res[0][3].cpu().numpy(), res[1][3].cpu().numpy()) < 1e-5
    assert torch.equal(res[0][0], res[2][0]) and torch.equal(res[0][1], res[2][1])
    s_f, s_u = res[0][2].cpu().numpy(), res[2][2].cpu().numpy()
    assert np.all(np.abs(s_f - s_u) <= 1e-5 * np.maximum(np.abs(s_u), 1.0)), np.abs(s_f - s_u).max()
    if not with_edt:
        assert (s_f[:, 3] == 0).all()
    assert np.allclose(s_f[3, 0], float(tgt[3 % NB].abs().sum()), rtol=1e-6) and s_f[3, 1] == 0
    assert util.rel_err(res[0][3].cpu().numpy(), res[2][3].cpu().numpy()) < 1e-4
    assert util.rel_err(res[0][4].cpu().numpy(), res[2][4].cpu().numpy()) < 1e-4
    ls = loss_utils.losses_from_sums(res[0][2], S * S, with_edt)
    ref = loss_utils.mask_losses(res[2][0], tgt, edt)
    for k in ref:
        assert torch.allclose(ls[k], ref[k], rtol=1e-5, atol=1e-7)


def test_backward_with_forward_work_lists_equals_full_grid(monkeypatch):
    """acfm_raster_soft_bwd visiting only the regions the forward found live (its work lists, heaviest first) gives the
    gradient of the full-grid launch: the skipped regions hold no fragment.  One render is off screen, one fills the image."""
    from acfm_video_3d_reconstruction_b200 import functional as F_
    v, f = util.template("bird")
    N, S = 6, 160
    X, cam = util.synth_verts(v, N, seed=21), util.synth_cams(N, seed=22)
    cam[1, 1:3] = (3.0, -3.0)
    cam[2, 0] = 3.0
    ndc = F_.project(torch.from_numpy(X).cuda(), torch.from_numpy(cam).cuda(), 5.0, -1.0, -1.0, F_.EYE_Z).detach()
    faces = torch.from_numpy(f)[None].cuda()
    g = torch.randn(N, S, S, generator=torch.Generator().manual_seed(1)).cuda()
    grads = []
    for split in (True, False):
        monkeypatch.setattr(F_, "SPLIT_FILL", split)
        x = ndc.clone().requires_grad_(True)
        mask, _, _, _ = F_.soft_silhouette(x, faces, S)
        grads.append(torch.autograd.grad((mask * g).sum(), x)[0])
    assert grads[0][1].abs().sum() == 0 and grads[0][0].abs().sum() > 0
    assert util.rel_err(grads[0].cpu().numpy(), grads[1].cpu().numpy()) < 1e-5


def test_non_square_image_split_equals_single_kernel():
    """The C ABI takes H != W (the Python mirror only renders squares): split and single-kernel paths agree byte for byte."""
    from acfm_video_3d_reconstruction_b200 import _lib
    from acfm_video_3d_reconstruction_b200 import functional as F_
    v, f = util.template("horse")
    N, H, W, K = 3, 96, 200, 8
    X, cam = util.synth_verts(v, N, seed=41), util.synth_cams(N, seed=42)
    ndc = F_.project(torch.from_numpy(X).cuda(), torch.from_numpy(cam).cuda(), 5.0, -1.0, -1.0, F_.EYE_Z).detach().contiguous()
    faces = torch.from_numpy(f).cuda().contiguous()
    outs = []
    for use_ws in (True, False):
        p2f = torch.empty(N, H, W, K, dtype=torch.int64, device="cuda")
        zbuf, dists = torch.empty(N, H, W, K, device="cuda"), torch.empty(N, H, W, K, device="cuda")
        mask = torch.empty(N, H, W, device="cuda")
        nws = int(_lib.lib().acfm_raster_fwd_workspace_bytes(N, H, W)) if use_ws else 0
        ws = torch.empty(nws, dtype=torch.uint8, device="cuda") if use_ws else None
        st = _lib.lib().acfm_raster_fwd(_lib.ptr(ndc), _lib.ptr(faces), 1, 0, N, v.shape[0], f.shape[0], H, W, K, F_.BLUR_SOFT, 0, 0, F_.SIGMA,
                                        _lib.ptr(p2f), _lib.ptr(zbuf), _lib.ptr(dists), None, _lib.ptr(mask), None, _lib.ptr(ws), nws,
                                        _lib.stream_of(ndc))
        _lib.check(st, "acfm_raster_fwd")
        torch.cuda.synchronize()
        outs.append((p2f, zbuf, dists, mask))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert (outs[0][0] >= 0).any()


@pytest.mark.parametrize("S,K", [(256, 20), (100, 5), (90, 3), (72, 1)])
def test_outputs_are_written_inside_their_bounds_only(S, K):
    """Guard bands around every output of acfm_raster_fwd (split path: TMA bulk stores of region rows, runs of regions,
    plain-store fallbacks) and around the backward's gradient stay untouched; every output element is written."""
    import ctypes
    from acfm_video_3d_reconstruction_b200 import _lib
    from acfm_video_3d_reconstruction_b200 import functional as F_
    v, f = util.template("bird")
    N, G = 3, 64                                       # G guard elements on both sides (keeps 16-byte alignment)
    X, cam = util.synth_verts(v, N, seed=31), util.synth_cams(N, seed=32)
    cam[1, 1:3] = (3.0, 3.0)
    ndc = F_.project(torch.from_numpy(X).cuda(), torch.from_numpy(cam).cuda(), 5.0, -1.0, -1.0, F_.EYE_Z).detach().contiguous()
    faces = torch.from_numpy(f).cuda().contiguous()
    V, Fn = v.shape[0], f.shape[0]

    def guarded(numel, dtype, poison):
        buf = torch.full((numel + 2 * G,), poison, dtype=dtype, device="cuda")
        return buf, buf[G:G + numel]

    frag = N * S * S * K
    bp, p2f = guarded(frag, torch.int64, -7)
    bz, zbuf = guarded(frag, torch.float32, 123.0)
    bd, dists = guarded(frag, torch.float32, 123.0)
    bm, mask = guarded(N * S * S, torch.float32, 123.0)
    bv, vis = guarded(N * V, torch.float32, 123.0)
    nws = int(_lib.lib().acfm_raster_fwd_workspace_bytes(N, S, S))
    bw, ws = guarded(nws, torch.uint8, 77)
    soft = K > 1
    st = _lib.lib().acfm_raster_fwd(_lib.ptr(ndc), _lib.ptr(faces), 1, 0, N, V, Fn, S, S, K, F_.BLUR_SOFT if soft else 0.0, 0, 0,
                                    F_.SIGMA if soft else 0.0, _lib.ptr(p2f), _lib.ptr(zbuf), _lib.ptr(dists), None,
                                    _lib.ptr(mask) if soft else None, _lib.ptr(vis), _lib.ptr(ws), nws, _lib.stream_of(ndc))
    _lib.check(st, "acfm_raster_fwd")
    torch.cuda.synchronize()
    for buf, view, poison in ((bp, p2f, -7), (bz, zbuf, 123.0), (bd, dists, 123.0), (bv, vis, 123.0), (bw, ws, 77)) + (((bm, mask, 123.0),) if soft else ()):
        assert (buf[:G] == poison).all() and (buf[-G:] == poison).all(), "guard band overwritten"
    assert (p2f != -7).all() and (zbuf != 123.0).all() and (dists != 123.0).all() and (vis != 123.0).all()
    assert (p2f >= 0).any() and (p2f.view(N, -1)[1] == -1).all()
    if soft:
        assert (mask != 123.0).all()
        bg, g = guarded(N * V * 3, torch.float32, 123.0)
        gm = torch.ones(N * S * S, device="cuda")
        st = _lib.lib().acfm_raster_soft_bwd(_lib.ptr(ndc), _lib.ptr(faces), 1, 0, N, V, Fn, S, S, K, F_.SIGMA, _lib.ptr(p2f), _lib.ptr(dists),
                                             _lib.ptr(mask), _lib.ptr(gm), _lib.ptr(g), _lib.ptr(ws), _lib.stream_of(ndc))
        _lib.check(st, "acfm_raster_soft_bwd")
        torch.cuda.synchronize()
        assert (bg[:G] == 123.0).all() and (bg[-G:] == 123.0).all() and (g != 123.0).all() and g.view(N, -1)[1].abs().sum() == 0


@pytest.mark.parametrize("clip,blur", [(False, 0.0), (True, 0.0), (False, 2e-3)])
def test_hard_render_depth_culls_are_exact(clip, blur):
    """K = 1 keeps the nearest fragment in registers and drops faces whose nearest vertex lies behind what a pixel (or a whole
    tile) already has.  The bound behind that cull — fragment depth >= (sum of the barycentric weights) x (nearest vertex depth)
    — is stressed here: large faces in the back, thousands of tiny faces barely above the degenerate-face threshold in front of
    and behind them (their weights add up to as little as 1/2, so their fragments are NEARER than their nearest vertex), faces
    crossing z = 0, and a blur band (where the cull must switch itself off).  Bit-exact against the oracle."""
    from acfm_video_3d_reconstruction_b200 import functional as F_
    gen = torch.Generator().manual_seed(17 + int(clip) + int(blur * 1e4))
    N, S = 3, 96
    nbig, ntiny = 60, 1500
    F = nbig + ntiny
    V = 3 * F
    c_big = torch.rand(N, nbig, 1, 2, generator=gen) * 2 - 1
    big = c_big + 0.5 * torch.randn(N, nbig, 3, 2, generator=gen)
    # tiny faces AROUND pixel centres (x = 1 - (2 i + 1) / S), so that they do own pixels: circumradius 6e-5 .. 2e-3, i.e. areas of
    # 5e-9 .. 5e-6 against kEpsilon = 1e-8
    c_tiny = 1.0 - (2.0 * torch.randint(0, S, (N, ntiny, 1, 2), generator=gen).float() + 1.0) / S
    size = 10 ** (torch.rand(N, ntiny, 1, 1, generator=gen) * 1.5 - 4.2)
    ang = torch.rand(N, ntiny, 1, generator=gen) * 6.2832 + torch.tensor([0.0, 2.0944, 4.1888])
    tiny = c_tiny + size * torch.stack([torch.cos(ang), torch.sin(ang)], -1) * (1.0 + 0.3 * torch.rand(N, ntiny, 3, 1, generator=gen))
    xy = torch.cat([big, tiny], 1).reshape(N, V, 2)
    z_big = torch.rand(N, nbig, 3, 1, generator=gen) * 2.0 + 0.5
    z_big[:, ::9] -= 1.5                                                          # some cross z = 0
    z_tiny = torch.rand(N, ntiny, 1, 1, generator=gen) * 3.0 + 0.05 + 0.3 * torch.rand(N, ntiny, 3, 1, generator=gen)
    z = torch.cat([z_big, z_tiny], 1).reshape(N, V, 1)
    ndc = torch.cat([xy, z], -1).numpy().astype(np.float32)
    faces = np.repeat(np.arange(V, dtype=np.int64).reshape(1, F, 3), N, 0)
    ref = orc.rasterize(ndc, faces, S, blur, 1, clip_bary=clip, want_bary=True)
    out = F_.rasterize(torch.from_numpy(ndc).cuda(), torch.from_numpy(faces).cuda(), S, blur, 1, clip_barycentric_coords=clip, want_bary=True)
    _assert_fragments_equal(out, ref)
    assert np.array_equal(out["bary"].cpu().numpy(), ref["bary"])
    hit = ref["pix_to_face"][..., 0] % F
    assert (ref["pix_to_face"] >= 0).mean() > 0.5 and ((hit >= nbig) & (ref["pix_to_face"][..., 0] >= 0)).sum() > 50   # tiny faces do win pixels


@pytest.mark.parametrize("with_target", [True, False])
def test_lean_mode_equals_the_parity_route(with_target):
    """acfm_raster_fwd_lean / _soft_bwd_lean (no fragment tensors; compact fragments of the live regions between forward and
    backward): silhouette, fused loss sums and visible vertices bit-identical to the API-parity render's, same gradient."""
    from acfm_video_3d_reconstruction_b200 import functional as F_
    from acfm_video_3d_reconstruction_b200 import NeuralRenderer
    v, f = util.template("bird")
    N, S, NB = 6, 160, 3
    X, cam = util.synth_verts(v, N, seed=71), util.synth_cams(N, seed=72)
    cam[1, 1] += 1.2                                          # one render mostly off screen
    ndc = NeuralRenderer(S, offset_z=5.0).to_ndc(torch.from_numpy(X).cuda(), torch.from_numpy(cam).cuda()).detach()
    faces = torch.from_numpy(f).cuda()[None]
    gen = torch.Generator().manual_seed(9)
    target = (torch.rand(NB, S, S, generator=gen) > 0.6).float().cuda() if with_target else None
    edt = torch.rand(NB, S, S, generator=gen).cuda() if with_target else None
    gm = torch.randn(N, S, S, generator=gen).cuda()
    gs = (torch.randn(N, 4, generator=gen) * 1e-3).cuda()

    def run(lean):
        x = ndc.clone().requires_grad_(True)
        if lean:
            out = F_.soft_silhouette_lean(x, faces, S, target, edt, want_vis=True)
            mask, sums, vis = (out[0], out[1], out[2]) if with_target else (out[0], None, out[1])
        elif with_target:
            mask, _, _, _, sums, vis = F_.soft_silhouette_losses(x, faces, S, target, edt, want_vis=True)
        else:
            mask, _, _, _, vis = F_.soft_silhouette(x, faces, S, want_vis=True)
            sums = None
        loss = (mask * gm).sum() + ((sums * gs).sum() if sums is not None else 0.0)
        g, = torch.autograd.grad(loss, x)
        return mask.detach(), sums, vis, g

    m0, s0, v0, g0 = run(False)
    m1, s1, v1, g1 = run(True)
    assert torch.equal(m0, m1) and torch.equal(v0, v1) and float(m0.sum()) > 100
    if with_target:
        assert torch.equal(s0, s1)
    assert util.rel_err(g1.cpu().numpy(), g0.cpu().numpy()) < 1e-4 and float(g0.abs().sum()) > 0
    with pytest.raises(ValueError):                          # built for the reference's K = 20 only (unsupported size)
        F_.soft_silhouette_lean(ndc, faces, S, faces_per_pixel=8)
