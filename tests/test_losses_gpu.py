"""GPU parity of the fused loss kernels vs golden vectors generated from the reference's loss_utils.py."""
import numpy as np
import pytest
import torch

from oracle import torch_ref
from tests import util

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # "losses within 1e-5 relative" (BASELINE.json north_star)


def test_mask_losses_vs_reference_golden():
    from acfm_video_3d_reconstruction_b200 import loss_utils
    g = util.golden("losses.npz")
    pred, targ, edt = (torch.from_numpy(g[k]).cuda() for k in ("pred", "targ", "edt"))
    np.testing.assert_allclose(loss_utils.l1_loss(pred, targ, reduce=False).cpu().numpy(), g["l1"], rtol=RTOL)
    np.testing.assert_allclose(loss_utils.l1_loss(pred, targ).cpu().numpy(), g["l1_mean"], rtol=RTOL)
    np.testing.assert_allclose(loss_utils.iou_loss(pred, targ, reduce=False).cpu().numpy(), g["iou"], rtol=RTOL)
    np.testing.assert_allclose(loss_utils.iou_loss(pred, targ).cpu().numpy(), g["iou_mean"], rtol=RTOL)
    np.testing.assert_allclose(loss_utils.edt_loss(pred, edt, reduce=False).cpu().numpy(), g["edt_l"], rtol=RTOL)
    np.testing.assert_allclose(loss_utils.edt_loss(pred, edt).cpu().numpy(), g["edt_mean"], rtol=RTOL)
    kp = loss_utils.kp_l2_loss(torch.from_numpy(g["kp_pred"]).cuda(), torch.from_numpy(g["kp_gt"]).cuda(), reduction="none")
    np.testing.assert_allclose(kp.cpu().numpy(), g["kp"], rtol=RTOL)
    allm = loss_utils.mask_losses(pred, targ, edt[:, 0])
    np.testing.assert_allclose(allm["l1"].cpu().numpy(), g["l1"], rtol=RTOL)
    np.testing.assert_allclose(allm["iou_loss"].cpu().numpy(), g["iou"], rtol=RTOL)
    np.testing.assert_allclose(allm["edt"].cpu().numpy(), g["edt_l"], rtol=RTOL)


def test_mask_losses_hypothesis_broadcast_and_grad():
    """target with NB rows against N = G*NB renders == the reference's target.repeat(G,1,1); grads vs fp64."""
    from acfm_video_3d_reconstruction_b200 import loss_utils
    gen = torch.Generator().manual_seed(0)
    NB, G, S = 3, 4, 64
    pred = torch.rand(NB * G, S, S, generator=gen)
    targ = (torch.rand(NB, S, S, generator=gen) > 0.5).float()
    edt = torch.rand(NB, S, S, generator=gen) * 4
    w = torch.randn(3, NB * G, generator=gen)
    pc = pred.cuda().requires_grad_(True)
    out = loss_utils.mask_losses(pc, targ.cuda(), edt.cuda())
    (out["l1"] * w[0].cuda() + out["iou_loss"] * w[1].cuda() + out["edt"] * w[2].cuda()).sum().backward()
    pd = pred.double().requires_grad_(True)
    td, ed = targ.double().repeat(G, 1, 1), edt.double().repeat(G, 1, 1)
    l1, io, el = torch_ref.l1_loss(pd, td), torch_ref.iou_loss(pd, td), torch_ref.edt_loss(pd, ed[:, None])
    (l1 * w[0].double() + io * w[1].double() + el * w[2].double()).sum().backward()
    np.testing.assert_allclose(out["l1"].detach().cpu().numpy(), l1.detach().numpy(), rtol=RTOL)
    np.testing.assert_allclose(out["iou_loss"].detach().cpu().numpy(), io.detach().numpy(), rtol=RTOL)
    np.testing.assert_allclose(out["edt"].detach().cpu().numpy(), el.detach().numpy(), rtol=RTOL)
    assert util.rel_err(pc.grad.cpu().numpy(), pd.grad.numpy()) < 1e-4


@pytest.mark.parametrize("w", [(1.0, 0.0, 0.1), (0.5, 2.0, 0.0), (0.0, 1.0, 3.0)])
def test_combined_mask_loss_equals_the_three_losses(w):
    """combine_mask_losses = w_l1 * l1_loss + w_iou * iou_loss + w_edt * edt_loss (reduce=False) formed from the four sums, values
    and gradient to the sums against the same combination written with torch ops in fp64."""
    from acfm_video_3d_reconstruction_b200 import loss_utils
    gen = torch.Generator().manual_seed(7)
    N, HW = 37, 96 * 96
    sums = (torch.rand(N, 4, generator=gen) * HW * 0.3 + 1.0).cuda().requires_grad_(True)
    gp = torch.randn(N, generator=gen).cuda()
    per = loss_utils.combine_mask_losses(sums, HW, *w)
    g, = torch.autograd.grad((per * gp).sum(), sums)
    s64 = sums.detach().double().requires_grad_(True)
    ref = w[0] * s64[:, 0] / HW + w[1] * (1 - s64[:, 1] / (s64[:, 2] + 1e-6)) + w[2] * s64[:, 3] / HW
    g64, = torch.autograd.grad((ref * gp.double()).sum(), s64)
    assert torch.allclose(per.double(), ref, rtol=1e-6, atol=1e-7)
    assert torch.allclose(g.double(), g64, rtol=1e-5, atol=1e-12)
    ls = loss_utils.losses_from_sums(sums.detach(), HW)
    assert torch.allclose(per, w[0] * ls["l1"] + w[1] * ls["iou_loss"] + w[2] * ls["edt"], rtol=1e-6, atol=1e-7)
    assert loss_utils.combine_mask_losses(sums[:0], HW).shape == (0,)
