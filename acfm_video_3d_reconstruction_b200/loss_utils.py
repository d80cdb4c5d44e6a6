"""Drop-in for the rendered-output losses of the reference's nnutils/loss_utils.py
(/root/reference/multiframe/nnutils/loss_utils.py; identical file in monocular/), backed by the fused
sm_100a kernels.  Same names / argument meaning / `reduce` behaviour as the reference.  `target`-like
arguments may carry NB <= N entries (N % NB == 0): render n uses entry n % NB, which is what the callers'
`.repeat(num_guesses, 1, 1)` expresses (multiframe/main.py:644,716) without materialising it.
"""
import torch

from . import _lib
from . import functional as F_


class _MaskSums(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask, target, edt):
        _lib.require_cuda(mask, target, edt)
        mask, target = F_._f32c(mask), F_._f32c(target)
        edt = F_._f32c(edt) if edt is not None else None
        N = mask.shape[0]
        HW = mask[0].numel() if N else 1
        NB = target.shape[0]
        if N and (NB == 0 or N % NB or target[0].numel() != HW or (edt is not None and edt.numel() != target.numel())):
            raise ValueError(f"mask {tuple(mask.shape)} vs target {tuple(target.shape)}: batch must divide, pixels must match")
        sums = torch.empty((N, 4), dtype=torch.float32, device=mask.device)
        with torch.cuda.device(mask.device):
            st = _lib.lib().acfm_mask_sums_fwd(_lib.ptr(mask), _lib.ptr(target), _lib.ptr(edt), N, max(NB, 1), HW,
                                               _lib.ptr(sums), _lib.stream_of(mask))
        _lib.check(st, "acfm_mask_sums_fwd")
        _lib.count(2)
        ctx.save_for_backward(mask, target, edt)
        return sums

    @staticmethod
    def backward(ctx, grad_sums):
        mask, target, edt = ctx.saved_tensors
        N = mask.shape[0]
        HW = mask[0].numel() if N else 1
        g = torch.empty_like(mask)
        with torch.cuda.device(mask.device):
            st = _lib.lib().acfm_mask_sums_bwd(_lib.ptr(mask), _lib.ptr(target), _lib.ptr(edt), _lib.ptr(F_._f32c(grad_sums)),
                                               N, max(target.shape[0], 1), HW, _lib.ptr(g), _lib.stream_of(mask))
        _lib.check(st, "acfm_mask_sums_bwd")
        _lib.count()
        return g, None, None


def mask_sums(mask, target, edt=None):
    """(N,4): sum|m-t|, sum m t, sum (m+t-mt), sum edt m per render — one fused pass."""
    return _MaskSums.apply(mask, target, edt)


def mask_losses(mask, target, edt=None):
    """All per-render silhouette losses in one pass: dict(l1, iou_loss, edt) each (N,)."""
    s = mask_sums(mask, target, edt)
    hw = float(mask[0].numel()) if mask.shape[0] else 1.0
    out = dict(l1=s[:, 0] / hw, iou_loss=1 - s[:, 1] / (s[:, 2] + 1e-6))
    if edt is not None:
        out["edt"] = s[:, 3] / hw
    return out


def _reduce(per_render, reduce):
    return per_render.mean() if reduce else per_render


def l1_loss(predict, target, reduce=True):
    """loss_utils.py:72-77"""
    return _reduce(mask_sums(predict, target)[:, 0] / float(predict[0].numel()), reduce)


def iou(predict, target, eps=1e-6, reduce=True):
    """loss_utils.py:18-29"""
    s = mask_sums(predict, target)
    v = s[:, 1] / (s[:, 2] + eps)
    return v.sum() / v.nelement() if reduce else v


def iou_loss(predict, target, reduce=True):
    """loss_utils.py:31-32"""
    return 1 - iou(predict, target, reduce=reduce)


def edt_loss(mask_rendered, edt, reduce=True):
    """loss_utils.py:245-253; edt (NB,1,H,W) or (NB,H,W)"""
    s = mask_sums(mask_rendered, torch.zeros_like(edt).reshape(edt.shape[0], -1), edt.reshape(edt.shape[0], -1))
    return _reduce(s[:, 3] / float(mask_rendered[0].numel()), reduce)


def kp_l2_loss(kp_pred, kp_gt, reduction='mean'):
    """loss_utils.py:341-356 (an L1 over visible keypoints, despite the name).  kp_gt may carry NB <= N rows."""
    N, NB = kp_pred.shape[0], kp_gt.shape[0]
    if N != NB:
        kp_gt = kp_gt.repeat(N // NB, 1, 1)
    vis = (kp_gt[:, :, 2] > 0).float()
    loss = (kp_pred - kp_gt[:, :, :2]).abs().sum(-1) * vis
    loss = loss.mean(-1) / (vis.mean(-1) + 1e-4)
    return loss.mean() if reduction == 'mean' else loss
