"""Correlation cost volume: the product kernels against
  (1) the reference's OWN extension, compiled from /root/reference into oracle/_ref/correlation_cuda.so (sm_100a) — the pin;
  (2) the plain-torch restatement oracle/correlation_ref.py (fp64 autograd = gradient truth), which is itself checked
      against (1).
Tolerance: 1e-5 relative (max-norm) for the forward — fp32 sums in a different order than the reference's 32-lane strided
partials + shuffle tree —, 1e-4 for the backward (the reference accumulates 32 partials serially)."""
import numpy as np
import pytest
import torch

from oracle import correlation_ref as cref
from tests import util

pytestmark = pytest.mark.gpu

# (B, C, H, W, pad, k, md, s1, s2): MaskFlowNet's call at its pyramid levels for a 256 x 256 input, then the general cases
NET = [(2, 196, 4, 4, 4, 1, 4, 1, 1), (2, 128, 8, 8, 4, 1, 4, 1, 1), (3, 96, 16, 16, 4, 1, 4, 1, 1), (2, 64, 32, 32, 4, 1, 4, 1, 1),
       (2, 32, 64, 64, 4, 1, 4, 1, 1), (1, 20, 37, 50, 4, 1, 4, 1, 1), (1, 9, 11, 70, 4, 1, 4, 1, 1)]
GENERAL = [(2, 8, 20, 24, 4, 1, 4, 1, 2), (1, 6, 18, 22, 3, 3, 2, 1, 1), (2, 5, 21, 19, 6, 1, 4, 2, 2), (1, 4, 24, 24, 20, 1, 20, 2, 2),
           (1, 7, 16, 20, 2, 1, 4, 1, 1), (1, 3, 30, 26, 6, 3, 4, 1, 1)]


def _inputs(cfg, seed=0, dtype=torch.float32):
    B, C, H, W = cfg[:4]
    gen = torch.Generator().manual_seed(seed)
    a = torch.randn(B, C, H, W, generator=gen, dtype=dtype).cuda()
    b = torch.randn(B, C, H, W, generator=gen, dtype=dtype).cuda()
    return a, b


@pytest.fixture(scope="module")
def ref_ext():
    mod = cref.load_reference_extension()
    if mod is None:
        pytest.skip("oracle/_ref/correlation_cuda.so was not built (no /root/reference at build time)")
    return mod


@pytest.mark.parametrize("cfg", NET + GENERAL)
def test_forward_vs_restatement(cfg):
    from acfm_video_3d_reconstruction_b200.correlation import Correlation, out_shape
    a, b = _inputs(cfg)
    args = cfg[4:]
    out = Correlation(*args)(a, b)
    want = cref.correlation(a.double(), b.double(), *args)
    assert out.shape == want.shape == (cfg[0],) + cref.out_shape(cfg[2], cfg[3], *args) == (cfg[0],) + out_shape(cfg[2], cfg[3], *args)
    assert util.rel_err(out.cpu().numpy(), want.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("cfg", NET + GENERAL)
def test_forward_vs_reference_extension(cfg, ref_ext):
    """The pin: the reference's own kernel on the same inputs, and the restatement against it."""
    from acfm_video_3d_reconstruction_b200.correlation import Correlation
    a, b = _inputs(cfg, seed=1)
    args = cfg[4:]
    ref = cref.reference_forward(ref_ext, a, b, *args)
    torch.cuda.synchronize()
    assert util.rel_err(Correlation(*args)(a, b).cpu().numpy(), ref.cpu().numpy()) < 1e-5
    assert util.rel_err(cref.correlation(a, b, *args).cpu().numpy(), ref.cpu().numpy()) < 1e-5


def test_tiled_kernel_with_pad_larger_than_max_displacement():
    """pad_size > max_displacement adds a ring of border outputs around the pad == md result (bit-identical interior)."""
    from acfm_video_3d_reconstruction_b200.correlation import Correlation
    a, b = _inputs((2, 40, 45, 67))
    tiled = Correlation(4, 1, 4, 1, 1)(a, b)                       # k = 1, s = 1, md = 4: the register-tiled kernel
    # the same displacements through the generic kernel: stride2 = 1 with md = 4 but pad 5 keeps D = 9 and shifts the window
    wide = Correlation(5, 1, 4, 1, 1)(a, b)                        # tiled kernel again, pad != md
    assert torch.equal(wide[:, :, 1:-1, 1:-1], tiled)              # pad + 1 only adds a border ring of outputs
    want = cref.correlation(a.double(), b.double(), 5, 1, 4, 1, 1)
    assert util.rel_err(wide.cpu().numpy(), want.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("cfg", [NET[2], NET[5], GENERAL[0], GENERAL[1], GENERAL[2], GENERAL[5]])
def test_backward_vs_fp64_autograd(cfg):
    from acfm_video_3d_reconstruction_b200.correlation import Correlation
    a, b = _inputs(cfg, seed=2)
    args = cfg[4:]
    a.requires_grad_(True); b.requires_grad_(True)
    out = Correlation(*args)(a, b)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(3)).cuda()
    ga, gb = torch.autograd.grad((out * g).sum(), (a, b))
    a64, b64 = a.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    wa, wb = torch.autograd.grad((cref.correlation(a64, b64, *args) * g.double()).sum(), (a64, b64))
    assert util.rel_err(ga.cpu().numpy(), wa.cpu().numpy()) < 1e-5
    assert util.rel_err(gb.cpu().numpy(), wb.cpu().numpy()) < 1e-5
    # only one gradient requested
    a2 = a.detach().requires_grad_(True)
    (g2,) = torch.autograd.grad((Correlation(*args)(a2, b.detach()) * g).sum(), (a2,))
    assert torch.equal(g2, ga)


@pytest.mark.parametrize("cfg", [NET[1], NET[3], NET[5]])
def test_backward_vs_reference_extension(cfg, ref_ext):
    from acfm_video_3d_reconstruction_b200.correlation import Correlation
    a, b = _inputs(cfg, seed=4)
    args = cfg[4:]
    a.requires_grad_(True); b.requires_grad_(True)
    out = Correlation(*args)(a, b)
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(5)).cuda()
    ga, gb = torch.autograd.grad((out * g).sum(), (a, b))
    ra, rb = cref.reference_backward(ref_ext, a.detach(), b.detach(), g, *args)
    torch.cuda.synchronize()
    assert util.rel_err(ga.cpu().numpy(), ra.cpu().numpy()) < 1e-4
    assert util.rel_err(gb.cpu().numpy(), rb.cpu().numpy()) < 1e-4


def test_errors_and_edge_cases():
    from acfm_video_3d_reconstruction_b200.correlation import Correlation
    a, b = _inputs((1, 4, 8, 8))
    with pytest.raises(RuntimeError):
        Correlation(4, 1, 4, 1, 1)(a.cpu(), b.cpu())                     # no CPU fallback
    with pytest.raises(ValueError):
        Correlation(4, 1, 4, 1, 1)(a.half(), b.half())
    with pytest.raises(ValueError):
        Correlation(0, 1, 4, 1, 1)(a, b)                                 # padded input smaller than the border
    with pytest.raises(ValueError):
        Correlation(4, 2, 4, 1, 1)(a, b)                                 # even kernel size
    empty = Correlation(4, 1, 4, 1, 1)(a[:0], b[:0])
    assert empty.shape == (0, 81, 8, 8)
    # shifted copy: the cost volume peaks at the displacement of the shift
    x = torch.randn(1, 16, 24, 24, generator=torch.Generator().manual_seed(7)).cuda()
    y = torch.roll(x, shifts=(2, -3), dims=(2, 3))                       # y[p + (2,-3)] = x[p]
    cost = Correlation(4, 1, 4, 1, 1)(x, y)[0, :, 8:16, 8:16].mean(dim=(1, 2))
    assert int(cost.argmax()) == (2 + 4) * 9 + (-3 + 4)


@pytest.mark.parametrize("cfg", [NET[5], NET[6], NET[0], GENERAL[0]])
def test_outputs_are_written_inside_their_bounds_only(cfg):
    """Guard bands around the cost volume and the input gradients stay untouched (odd sizes: scalar-store and 4-byte
    staging fallbacks of the tiled kernel; the generic kernel), every element is written."""
    from acfm_video_3d_reconstruction_b200 import _lib
    from acfm_video_3d_reconstruction_b200.correlation import out_shape
    a, b = _inputs(cfg, seed=9)
    B, C, H, W = cfg[:4]
    args = cfg[4:]
    ch, oh, ow = out_shape(H, W, *args)
    G = 64

    def guarded(numel):
        buf = torch.full((numel + 2 * G,), 123.0, device="cuda")
        return buf, buf[G:G + numel]

    bo, out = guarded(B * ch * oh * ow)
    st = _lib.lib().acfm_correlation_fwd(_lib.ptr(a), _lib.ptr(b), B, C, H, W, *args, _lib.ptr(out), _lib.stream_of(a))
    _lib.check(st, "acfm_correlation_fwd")
    g = torch.randn(B * ch * oh * ow, device="cuda")
    b1, g1 = guarded(a.numel())
    b2, g2 = guarded(a.numel())
    st = _lib.lib().acfm_correlation_bwd(_lib.ptr(a), _lib.ptr(b), _lib.ptr(g), B, C, H, W, *args, _lib.ptr(g1), _lib.ptr(g2), _lib.stream_of(a))
    _lib.check(st, "acfm_correlation_bwd")
    torch.cuda.synchronize()
    for buf, view in ((bo, out), (b1, g1), (b2, g2)):
        assert (buf[:G] == 123.0).all() and (buf[-G:] == 123.0).all(), "guard band overwritten"
        assert (view != 123.0).all()
    want = cref.correlation(a.double(), b.double(), *args)
    assert util.rel_err(out.view(want.shape).cpu().numpy(), want.cpu().numpy()) < 1e-5
