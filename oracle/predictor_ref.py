"""oracle/predictor_ref.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the post-processing optimisation loop of MeshPredictor.forward
(/root/reference/multiframe/nnutils/predictor.py:287-349): per iteration the reference's batched handle solve, the
(restated) PyTorch3D soft-silhouette render through oracle/raster_oracle.c with its backward, l1 / edt / boundary losses
(torch_ref, pinned to reference-generated goldens), Adam(lr=5e-3).  The predictor module itself cannot be imported
(mesh_net needs gdist / kornia / pytorch3d), so the loop is restated line by line; PARITY UNPINNED for the renderer.
"""
import numpy as np
import torch

from oracle import pt3d_oracle as orc
from oracle import torch_ref


class _SoftRender(torch.autograd.Function):
    """NeuralRenderer.forward mask branch on CPU: ndc (N,V,3) -> mask; backward through the C oracle."""

    @staticmethod
    def forward(ctx, ndc, faces, img_size):
        fr = orc.rasterize(ndc.detach().numpy().astype(np.float32), faces, img_size, orc.BLUR_SOFT, orc.K_SOFT, want_bary=False)
        fr["ndc"] = ndc.detach().numpy().astype(np.float32)
        mask = orc.sigmoid_alpha_blend(fr["dists"], fr["pix_to_face"])
        ctx.fr, ctx.faces = fr, faces
        ctx.p2f = torch.from_numpy(fr["pix_to_face"])
        return torch.from_numpy(mask).to(ndc.dtype), ctx.p2f

    @staticmethod
    def backward(ctx, g, _):
        gn = orc.neural_renderer_mask_backward(ctx.fr, ctx.faces, g.numpy().astype(np.float32))
        return torch.from_numpy(gn).to(g.dtype), None, None


def post_optimize(mean_v, lbs, L, delta_v_res, cam_pred, masks, edts_barrier, boundaries, faces, sample_indices, img_size,
                  offset_z=0.0, lr=5e-3, mask_loss_wt=1.0, boundaries_reg_wt=1.0, edt_reg_wt=0.1, bdt_reg_wt=0.1,
                  optimize_camera=False, solve_dtype=torch.float32):
    """All tensors torch CPU (fp32); faces (NB,F,3) int64.  Returns dict(losses, pred_v, cam_pred, delta_v_res, mask_pred, grad0).
    solve_dtype=torch.float64 runs the handle solve in double (the reference's fp32 batched Cholesky carries ~3e-5 absolute
    vertex error and ~2e-3 relative gradient error at cond ~1e3): the truth for gradient checks."""
    f32 = torch.float32
    mean_v, lbs, L = mean_v.to(solve_dtype), lbs.to(solve_dtype), L.to(solve_dtype)
    NB = delta_v_res.shape[0]
    A = lbs.t()[None].repeat(NB, 1, 1)                                   # self.lbs (predictor.py:257-258)
    mean = mean_v[None].repeat(NB, 1, 1)
    delta_v_ms = A.bmm(mean)
    Lb = L[None].repeat(NB, 1, 1)
    A_augm = Lb.permute(0, 2, 1).matmul(Lb) + A.permute(0, 2, 1).matmul(A)
    delta = torch.bmm(Lb, mean)
    dres = delta_v_res.clone().requires_grad_(True)
    params = [dres]
    scale, trans, quat = cam_pred[:, :1].clone(), cam_pred[:, 1:3].clone(), cam_pred[:, 3:].clone()
    if optimize_camera:
        scale, trans, quat = (t.requires_grad_(True) for t in (scale, trans, quat))
        params += [scale, trans, quat]
    opt = torch.optim.Adam(params, lr=lr)
    fn = faces.numpy()
    losses = []

    def objective(sel):
        cam = torch.cat([scale, trans, torch.nn.functional.normalize(quat, dim=-1)], 1) if optimize_camera else cam_pred
        delta_v = delta_v_ms + dres.to(solve_dtype)
        b = Lb.permute(0, 2, 1) @ delta + A.permute(0, 2, 1) @ delta_v
        pred_v = torch.cholesky_solve(b, torch.linalg.cholesky(A_augm)).to(f32)
        mask_pred, p2f = _SoftRender.apply(torch_ref.to_ndc(pred_v, cam, offset_z), fn, img_size)
        mask_loss = torch_ref.l1_loss(mask_pred, masks).mean()
        pred_proj = torch_ref.orthographic_proj_withz(pred_v, cam, 0.0)[..., :2]
        edt_loss = torch_ref.edt_loss(mask_pred, edts_barrier.reshape(NB, 1, *masks.shape[1:])).mean()
        bdt_loss = torch_ref.bds_loss(pred_proj, boundaries, faces, p2f, sel).mean()
        total = mask_loss_wt * mask_loss + boundaries_reg_wt * (bdt_reg_wt * edt_loss + edt_reg_wt * bdt_loss)
        return total, pred_v, cam, mask_pred

    grad0 = None
    for it in range(sample_indices.shape[0]):
        total, pred_v, cam, mask_pred = objective(sample_indices[it])
        opt.zero_grad()
        total.backward()
        if grad0 is None:   # the first iteration's gradient, for parity checks (an Adam trajectory hides its scale)
            grad0 = dict(delta=dres.grad.clone())
            if optimize_camera:
                grad0.update(scale=scale.grad.clone(), trans=trans.grad.clone(), quat=quat.grad.clone())
        opt.step()
        losses.append(float(total.detach()))
    # like the reference (predictor.py:309-349), what is left when the loop ends is the LAST iteration's forward pass
    # (self.pred_v / self.cam_pred / mask_pred before the last Adam step); the parameters themselves are one step further
    return dict(losses=np.asarray(losses), pred_v=pred_v.detach(), cam_pred=cam.detach(), delta_v_res=dres.detach(),
                mask_pred=mask_pred.detach(), grad0=grad0)
