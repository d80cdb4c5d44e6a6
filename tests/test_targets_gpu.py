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


def test_boundaries_with_capacity_is_capturable():
    """compute_boundaries(max_bd=cap): no host sync — the call records into a CUDA graph; a larger cap pads with (-1,-1,0),
    a smaller one keeps the first cap points of the raster scan; the true counts come back on the device."""
    from acfm_video_3d_reconstruction_b200 import image_utils
    g = util.golden("targets.npz")
    m = torch.from_numpy(g["masks"]).cuda()
    want = g["boundaries"]
    nb, L = want.shape[0], want.shape[1]
    image_utils.compute_boundaries(m, max_bd=L + 7)            # warm-up outside the capture
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.cuda.graph(gr, stream=s):
        big, cnt = image_utils.compute_boundaries(m, max_bd=L + 7, return_counts=True)
        small = image_utils.compute_boundaries(m, max_bd=max(L - 5, 1))
    gr.replay()
    torch.cuda.synchronize()
    big, small, cnt = big.cpu().numpy(), small.cpu().numpy(), cnt.cpu().numpy()
    assert np.array_equal(big[:, :L], want)
    assert np.array_equal(big[:, L:], np.broadcast_to(np.float32([-1, -1, 0]), (nb, 7, 3)))
    assert np.array_equal(small, want[:, :small.shape[1]])
    assert np.array_equal(cnt, (want[..., 2] > 0).sum(1))
