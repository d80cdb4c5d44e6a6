"""Camera-multiplex assembly — drop-in for the elementwise block of ShapeTrainer.forward that turns the G
per-frame camera embeddings into `cam_pred` (/root/reference/multiframe/main.py:551-584) including
`mirror_cameras` (:113-125) and `transform_cameras` (:128-138), as one fused sm_100a kernel (fwd + bwd).
"""
import torch

from . import _lib
from . import functional as F_


class _Assemble(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, mirror_flag, transforms, scale_lr_decay):
        _lib.require_cuda(raw, mirror_flag, transforms)
        raw = F_._f32c(raw)
        if raw.dim() != 2 or raw.shape[1] != 7:
            raise ValueError(f"cameras must be (G*NB,7), got {tuple(raw.shape)}")
        N = raw.shape[0]
        NB = N
        mf = tf = None
        if mirror_flag is not None:
            mf = mirror_flag.reshape(-1).float().contiguous()
            NB = mf.shape[0]
        if transforms is not None:
            tf = F_._f32c(transforms.reshape(-1, 4))
            NB = tf.shape[0]
        if (mf is not None and tf is not None and mf.shape[0] != tf.shape[0]) or NB == 0 or N % NB:
            raise ValueError("mirror_flag (NB) / transforms (NB,4) must share NB with N % NB == 0")
        out = torch.empty_like(raw)
        with torch.cuda.device(raw.device):
            st = _lib.lib().acfm_camera_assemble_fwd(_lib.ptr(raw), _lib.ptr(mf), _lib.ptr(tf), N, NB, scale_lr_decay,
                                                     _lib.ptr(out), _lib.stream_of(raw))
        _lib.check(st, "acfm_camera_assemble_fwd")
        _lib.count()
        ctx.save_for_backward(raw, mf, tf)
        ctx.cfg = (N, NB, scale_lr_decay)
        return out

    @staticmethod
    def backward(ctx, g):
        raw, mf, tf = ctx.saved_tensors
        N, NB, lam = ctx.cfg
        graw = torch.empty_like(raw)
        with torch.cuda.device(raw.device):
            st = _lib.lib().acfm_camera_assemble_bwd(_lib.ptr(raw), _lib.ptr(mf), _lib.ptr(tf), _lib.ptr(F_._f32c(g)), N, NB,
                                                     lam, _lib.ptr(graw), _lib.stream_of(raw))
        _lib.check(st, "acfm_camera_assemble_bwd")
        _lib.count()
        return graw, None, None, None


def assemble_cameras(cameras, mirror_flag=None, transforms=None, scale_lr_decay=0.05):
    """cameras: (G, NB, 7) stacked embedding rows (or (G*NB,7) hypothesis-major); mirror_flag (NB,) 0/1;
    transforms (NB,4) = [scale, tx, ty, flag].  Returns cam_pred (G*NB, 7) = main.py:573-582."""
    raw = cameras.reshape(-1, 7)
    return _Assemble.apply(raw, mirror_flag, transforms, float(scale_lr_decay))
