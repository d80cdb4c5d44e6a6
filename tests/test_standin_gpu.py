"""The GPU stand-in for PyTorch3D's CUDA rasterizer (oracle/pt3d_cuda_standin.{cu,py}: bench.py's `gpu_standin` denominator)
against the C oracle: same fragments, same silhouette, same gradients — so that what bench.py times beside our kernels is a
correct renderer, not a strawman.  Its arithmetic is contracted into FMAs by nvcc as PyTorch3D's own CUDA build is, so the
comparison allows the few pixels whose inside / blur decision or depth order sits within an ulp (SURVEY.md section 9.9 ii)."""
import numpy as np
import pytest
import torch

from oracle import pt3d_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M", [None, 2048])   # (at 128^2 the mesh sits in ~3 x 3 bins of 16 px: several hundred faces each)
def test_standin_matches_oracle(M):
    from oracle import pt3d_cuda_standin as sd
    assert sd.available(), "oracle/_build/libacfm_pt3d_standin.so missing: run `make -C oracle`"
    v, f = util.template("bird")
    N, S, K = 3, 128, 20
    X, cam = util.synth_verts(v, N, seed=81), util.synth_cams(N, seed=82)
    faces = np.repeat(f[None], N, 0)
    ref = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=5.0, K=K)
    gm = np.random.default_rng(0).standard_normal((N, S, S)).astype(np.float32)
    g_ref = orc.neural_renderer_mask_backward(ref, faces, gm)
    nd = torch.from_numpy(ref["ndc"]).cuda().requires_grad_(True)
    mask, p2f = sd.render_mask(nd, torch.from_numpy(f).cuda(), S, orc.BLUR_SOFT, K, orc.SIGMA, max_faces_per_bin=M, check_overflow=True)
    (mask * torch.from_numpy(gm).cuda()).sum().backward()
    same = (p2f.cpu().numpy() == ref["pix_to_face"]).all(-1)
    assert same.mean() > 0.995, same.mean()
    assert np.abs(mask.detach().cpu().numpy() - ref["mask"])[same].max() < 1e-5
    assert util.rel_err(nd.grad.cpu().numpy(), g_ref) < 2e-3
