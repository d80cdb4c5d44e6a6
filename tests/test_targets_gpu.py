"""GPU parity of the target-map kernels (compute_dt / compute_dt_barrier / compute_boundaries)
against tests/golden/targets.npz — outputs of the reference's own utils/image.py — and against scipy on larger maps.
Bars: distance transforms and boundary lists exact (fp32 of the reference's fp64 values); barrier map 1e-6 relative
(exp in fp64 on both sides, last-ulp differences of the two libm's allowed)."""
import numpy as np
import pytest
import torch

from oracle import targets_ref as tr
from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sfx,k", [("", 50), ("_b", 20)])
def test_target_maps_vs_reference_golden(sfx, k):
    from acfm_video_3d_reconstruction_b200 import image_utils
    g = util.golden("targets.npz")
    m = torch.from_numpy(g["masks" + sfx]).cuda()
    dt_raw = image_utils.compute_dt(m, norm=False)
    assert np.array_equal(dt_raw.cpu().numpy(), g["dt_raw" + sfx].astype(np.float32))
    assert np.array_equal(image_utils.compute_dt(m).cpu().numpy(), g["dt_norm" + sfx].astype(np.float32))
    bar = image_utils.compute_dt_barrier(m, k=k)
    assert np.allclose(bar.cpu().numpy(), g["barrier" + sfx].astype(np.float32), rtol=1e-6, atol=1e-30)
    e2, b2 = image_utils.compute_dt_both(m, k=k)
    assert torch.equal(e2, dt_raw) and torch.equal(b2, bar)
    assert np.array_equal(image_utils.compute_boundaries(m).cpu().numpy(), g["boundaries" + sfx])
    # single (H,W) mask, as the reference calls compute_dt
    assert torch.equal(image_utils.compute_dt(m[0], norm=False), dt_raw[0])


def test_target_maps_256_vs_scipy():
    """C2-sized masks (our own renders, thresholded): 16 maps of 256^2."""
    from acfm_video_3d_reconstruction_b200 import NeuralRenderer, image_utils, synthetic
    v, f = util.template("bird")
    n = 16
    cams = synthetic.cameras(n, 1, seed=3).cuda()
    with torch.no_grad():
        mk, _ = NeuralRenderer(256, offset_z=5.0)(torch.from_numpy(v)[None].repeat(n, 1, 1).cuda(),
                                                  torch.from_numpy(f)[None].repeat(n, 1, 1).cuda(), cams)
    m = (mk > 0.5).float()
    mn = m.cpu().numpy()
    e, b = image_utils.compute_dt_both(m)
    assert np.array_equal(e.cpu().numpy(), np.stack([tr.compute_dt(x, norm=False) for x in mn]).astype(np.float32))
    assert np.allclose(b.cpu().numpy(), np.stack([tr.compute_dt_barrier(x) for x in mn]).astype(np.float32), rtol=1e-6, atol=1e-30)
    assert np.array_equal(image_utils.compute_boundaries(m).cpu().numpy(), tr.compute_boundaries(mn))
    assert 0.03 < float(m.mean()) < 0.5
