"""Parity at BASELINE.json's FULL sizes, where the CPU oracle cannot rasterize the whole batch in seconds:
size-independent properties of the fragment tensors over the entire batch, the oracle itself on a random handful of
renders taken out of the full batch, batch invariance (a render does not depend on its neighbours or on its index), and
linearity of the backward.  Configurations: C2 (512 renders of 256 x 256, K = 20: bench.py's workload), C4 (2562-vertex
template, 512 x 512, K = 50) and a batch whose fragment tensors pass 2^31 elements (64-bit addressing in the rasterizer,
the TMA padding kernel and the backward)."""
import numpy as np
import pytest
import torch

from oracle import pt3d_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


def _render(name, frames, G, S, K, seed=0, offset_z=5.0):
    from acfm_video_3d_reconstruction_b200 import functional as F_
    from acfm_video_3d_reconstruction_b200 import synthetic
    wl = synthetic.Workload(name, frames, G, 8, S, seed=seed, offset_z=offset_z)
    gen = torch.Generator().manual_seed(seed + 5)
    X = (wl.mean_v[None] + 0.02 * torch.randn(frames, wl.V, 3, generator=gen)).cuda()
    cams = wl.cams.cuda()
    ndc = F_.project(X, cams, offset_z, -1.0, -1.0, F_.EYE_Z).detach()
    faces = wl.faces[None].cuda()
    fr = F_.rasterize(ndc, faces, S, F_.BLUR_SOFT, K, sigma=F_.SIGMA, want_mask=True)
    return wl, ndc, faces, fr


def _check_render_properties(fr, n, F, K, blur):
    """Everything PyTorch3D's fragment layout promises, for render n (one render at a time: no batch-sized temporaries)."""
    p2f, z, d, m = fr["pix_to_face"][n], fr["zbuf"][n], fr["dists"][n], fr["mask"][n]
    valid = p2f >= 0
    assert torch.equal(valid, z >= 0) and torch.equal(valid, d != -1.0)       # -1 padding is consistent across tensors
    assert not (valid[..., 1:] & ~valid[..., :-1]).any()                       # front-packed
    assert ((p2f[valid] >= n * F) & (p2f[valid] < (n + 1) * F)).all()          # packed ids of THIS render
    both = valid[..., 1:]
    dz = z[..., 1:] - z[..., :-1]
    assert (dz[both] >= 0).all()                                               # depth-ascending
    tie = both & (dz == 0)
    assert (p2f[..., 1:][tie] > p2f[..., :-1][tie]).all()                      # depth ties in face order
    assert (d[valid] < blur).all()                                            # inside (negative) or within the blur band
    prob = torch.sigmoid(-d.double() / 1e-4) * valid
    want = 1.0 - torch.prod(1.0 - prob, dim=-1)
    assert (m.double() - want).abs().max() <= 1e-5
    assert torch.equal(m == 0, ~valid[..., 0])                                 # the backward's "no fragments" test


@pytest.mark.parametrize("name,frames,G,S,K,offset_z", [
    ("bird", 64, 8, 256, 20, 5.0),      # C2: bench.py's workload
    ("ico4", 4, 2, 512, 50, 0.0),       # C4: high-res stress
])
def test_full_size_batch(name, frames, G, S, K, offset_z):
    from acfm_video_3d_reconstruction_b200 import functional as F_
    wl, ndc, faces, fr = _render(name, frames, G, S, K, offset_z=offset_z)
    N = ndc.shape[0]
    assert N == frames * G
    for n in range(N):
        _check_render_properties(fr, n, wl.F, K, F_.BLUR_SOFT)
    # the oracle on a random handful of renders of the batch: bit-exact fragments
    rng = np.random.default_rng(3)
    pick = sorted(rng.choice(N, size=min(N, 6 if S <= 256 else 2), replace=False).tolist())
    ref = orc.rasterize(ndc[pick].cpu().numpy(), np.repeat(wl.faces.numpy()[None], len(pick), 0), S, orc.BLUR_SOFT, K, want_bary=False)
    for i, n in enumerate(pick):
        got = fr["pix_to_face"][n].cpu().numpy()
        want = np.where(ref["pix_to_face"][i] >= 0, ref["pix_to_face"][i] - i * wl.F + n * wl.F, -1)
        assert np.array_equal(got, want), f"pix_to_face of render {n}"
        assert np.array_equal(fr["zbuf"][n].cpu().numpy(), ref["zbuf"][i])
        assert np.array_equal(fr["dists"][n].cpu().numpy(), ref["dists"][i])
    # depth ties (SURVEY.md section 9.9 i): how many covered pixels of the whole batch hold one, and — on the handful — that
    # the reference's OTHER backend rule (PyTorch3D's CUDA kernels keep ties in traversal order, its CPU rasterizer orders
    # them by face id like this kernel and the oracle) changes nothing away from those pixels
    ties = cov = 0
    for n in range(N):
        z, v = fr["zbuf"][n], fr["pix_to_face"][n] >= 0
        ties += int(((z[..., 1:] == z[..., :-1]) & v[..., 1:]).any(-1).sum())
        cov += int(v[..., 0].sum())
    print(f"{name} {S}x{S} K={K}: {ties} of {cov} covered pixels of the {N} renders hold an exact depth tie ({100.0 * ties / cov:.2f} %)")
    assert ties / cov < 0.10
    sub_ndc, sub_faces = ndc[pick].cpu().numpy(), np.repeat(wl.faces.numpy()[None], len(pick), 0)
    orc.set_variant(tie_cuda=True)
    try:
        ref_cuda = orc.rasterize(sub_ndc, sub_faces, S, orc.BLUR_SOFT, K, want_bary=False)
    finally:
        orc.set_variant()
    deeper = orc.rasterize(sub_ndc, sub_faces, S, orc.BLUR_SOFT, K + 1, want_bary=False)   # a tie can straddle the K-th slot
    zd, vd = deeper["zbuf"], deeper["pix_to_face"] >= 0
    tie_px = ((zd[..., 1:] == zd[..., :-1]) & vd[..., 1:]).any(-1)
    differs = (ref_cuda["pix_to_face"] != ref["pix_to_face"]).any(-1) | (ref_cuda["dists"] != ref["dists"]).any(-1)
    assert not (differs & ~tie_px).any()
    # batch invariance: the same renders on their own, in another order
    sub = list(reversed(pick))
    alone = F_.rasterize(ndc[sub].contiguous(), faces, S, F_.BLUR_SOFT, K, sigma=F_.SIGMA, want_mask=True)
    for i, n in enumerate(sub):
        a = alone["pix_to_face"][i]
        assert torch.equal(torch.where(a >= 0, a + (n - i) * wl.F, a), fr["pix_to_face"][n])
        assert torch.equal(alone["zbuf"][i], fr["zbuf"][n]) and torch.equal(alone["dists"][i], fr["dists"][n])
        assert torch.equal(alone["mask"][i], fr["mask"][n])


def test_full_size_backward_is_linear_and_batch_invariant():
    """d loss / d ndc at C2's full size: linear in the upstream gradient (fixed-point accumulation included), equal to the
    gradient of the same renders taken alone, zero for renders whose upstream gradient is zero."""
    from acfm_video_3d_reconstruction_b200 import functional as F_
    wl, ndc, faces, _ = _render("bird", 64, 8, 256, 20)
    N, S = ndc.shape[0], 256
    gen = torch.Generator(device="cuda").manual_seed(1)
    g1 = torch.randn(N, S, S, device="cuda", generator=gen)
    g2 = torch.randn(N, S, S, device="cuda", generator=gen)
    g2[5] = 0.0

    def grad(x, g):
        x = x.clone().requires_grad_(True)
        mask, _, _, _ = F_.soft_silhouette(x, faces, S)
        return torch.autograd.grad((mask * g).sum(), x)[0]

    a, b, ab = grad(ndc, g1), grad(ndc, g2), grad(ndc, 0.5 * g1 - 2.0 * g2)
    scale = ab.abs().amax(dim=(1, 2), keepdim=True).clamp_min(1e-30)
    assert ((0.5 * a - 2.0 * b - ab).abs() / scale).max() < 1e-3
    assert (b[5] == 0).all() and (a[..., 2] == 0).all()
    sub = [511, 0, 300]
    alone = grad(ndc[sub].contiguous(), g1[sub].contiguous())
    for i, n in enumerate(sub):
        assert util.rel_err(alone[i].cpu().numpy(), a[n].cpu().numpy()) < 1e-4     # atomics order differs between launches


def test_fragment_tensors_beyond_2_31_elements():
    """104 renders of 1024 x 1024 with K = 20: 2.18e9 fragment slots per tensor (35 GB in all).  The last renders sit past
    the 32-bit element range: their properties hold, they equal the same renders taken alone, and so does their gradient."""
    from acfm_video_3d_reconstruction_b200 import functional as F_
    S, K = 1024, 20
    wl, ndc, faces, fr = _render("bird", 13, 8, S, K)
    N = ndc.shape[0]
    assert N * S * S * K > 2 ** 31
    for n in (0, N // 2, N - 2, N - 1):
        _check_render_properties(fr, n, wl.F, K, F_.BLUR_SOFT)
    sub = [N - 1, N - 2]
    alone = F_.rasterize(ndc[sub].contiguous(), faces, S, F_.BLUR_SOFT, K, sigma=F_.SIGMA, want_mask=True)
    for i, n in enumerate(sub):
        a = alone["pix_to_face"][i]
        assert torch.equal(torch.where(a >= 0, a + (n - i) * wl.F, a), fr["pix_to_face"][n])
        assert torch.equal(alone["dists"][i], fr["dists"][n]) and torch.equal(alone["mask"][i], fr["mask"][n])
    del alone, fr
    g = torch.zeros(N, S, S, device="cuda")
    g[N - 1] = 1.0
    x = ndc.clone().requires_grad_(True)
    mask, _, _, _ = F_.soft_silhouette(x, faces, S)
    gx = torch.autograd.grad((mask * g).sum(), x)[0]
    assert (gx[: N - 1] == 0).all() and gx[N - 1].abs().sum() > 0
    x1 = ndc[N - 1:].clone().requires_grad_(True)
    m1, _, _, _ = F_.soft_silhouette(x1, faces, S)
    g1 = torch.autograd.grad(m1.sum(), x1)[0]
    assert util.rel_err(g1[0].cpu().numpy(), gx[N - 1].cpu().numpy()) < 1e-4
