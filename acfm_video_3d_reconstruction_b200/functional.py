"""Autograd-level wrappers of the C-ABI kernels (include/acfm_b200.h).

Each function validates its tensors, calls libacfm_b200.so through ctypes on the tensor's device
and current stream, and (where differentiable) registers the CUDA backward.  Outputs are
allocated with torch's caching allocator so autograd / DataParallel gather keep working
(SURVEY.md §8b "Ownership").
"""
import math
import os

import torch

from . import _lib

EYE_Z = 2.732  # look_at_view_transform(eye=(0,0,-2.732)): /root/reference/multiframe/nnutils/nmr.py:144
SIGMA = 1e-4   # BlendParams(sigma=1e-4): nmr.py:153
BLUR_SOFT = math.log(1.0 / 1e-4 - 1.0) * 1e-4  # RasterizationSettings.blur_radius: nmr.py:157
K_SOFT = 20    # faces_per_pixel: nmr.py:158


def _f32c(t):
    if t.dtype != torch.float32:
        raise ValueError(f"expected a float32 tensor, got {t.dtype}")
    return t.contiguous()


def _faces_arg(faces, N):
    """faces (N,F,3) or (1,F,3)/(F,3) int64|int32 -> (tensor kept alive, is_i64, batch stride, F)."""
    if faces.dtype not in (torch.int64, torch.int32):
        raise ValueError(f"faces must be int64 or int32, got {faces.dtype}")
    if faces.dim() == 2:
        faces = faces[None]
    if faces.dim() != 3 or faces.shape[-1] != 3:
        raise ValueError(f"faces must have shape (N,F,3), got {tuple(faces.shape)}")
    F = faces.shape[1]
    if faces.shape[0] == 1 or (faces.shape[0] == N and faces.stride(0) == 0):
        f0 = faces[0].contiguous()
        return f0, int(faces.dtype == torch.int64), 0, F
    if faces.shape[0] != N:
        raise ValueError(f"faces batch {faces.shape[0]} does not match {N} renders")
    return faces.contiguous(), int(faces.dtype == torch.int64), F * 3, F


# -------------------------------------------------------------------------------------------------
# projection
# -------------------------------------------------------------------------------------------------
class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, cams, offset_z, sx, sy, z_add):
        _lib.require_cuda(verts, cams)
        verts, cams = _f32c(verts), _f32c(cams)
        if verts.dim() != 3 or verts.shape[-1] != 3 or cams.dim() != 2 or cams.shape[-1] != 7:
            raise ValueError(f"expected verts (NB,V,3) and cams (N,7), got {tuple(verts.shape)}, {tuple(cams.shape)}")
        NB, V, _ = verts.shape
        N = cams.shape[0]
        out = torch.empty((N, V, 3), dtype=torch.float32, device=verts.device)
        with torch.cuda.device(verts.device):
            st = _lib.lib().acfm_project_fwd(_lib.ptr(verts), _lib.ptr(cams), N, NB, V, offset_z, sx, sy, z_add,
                                             _lib.ptr(out), _lib.stream_of(verts))
        _lib.check(st, "acfm_project_fwd")
        _lib.count()
        ctx.save_for_backward(verts, cams)
        ctx.sx, ctx.sy = sx, sy
        return out

    @staticmethod
    def backward(ctx, grad_out):
        verts, cams = ctx.saved_tensors
        NB, V, _ = verts.shape
        N = cams.shape[0]
        grad_out = _f32c(grad_out)
        gv = torch.empty_like(verts) if ctx.needs_input_grad[0] else None
        gc = torch.empty_like(cams) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(verts.device):
            st = _lib.lib().acfm_project_bwd(_lib.ptr(verts), _lib.ptr(cams), _lib.ptr(grad_out), N, NB, V, ctx.sx,
                                             ctx.sy, _lib.ptr(gv), _lib.ptr(gc), _lib.stream_of(verts))
        _lib.check(st, "acfm_project_bwd")
        _lib.count()
        return gv, gc, None, None, None, None


def project(verts, cams, offset_z=0.0, sx=1.0, sy=1.0, z_add=0.0):
    """out = (sx*p.x, sy*p.y, p.z + z_add), p = orthographic_proj_withz(verts[n % NB], cams[n], offset_z)."""
    return _Project.apply(verts, cams, float(offset_z), float(sx), float(sy), float(z_add))


# -------------------------------------------------------------------------------------------------
# rasterization
# -------------------------------------------------------------------------------------------------
# True: acfm_raster_fwd gets a scratch buffer and runs its split path (region classification + rasterizer + concurrent fill
# kernel, see include/acfm_b200.h); False: the single-kernel path.  Same results; a switch for tests and A/B timing.
SPLIT_FILL = os.environ.get("ACFM_SPLIT_FILL", "1") != "0"


def set_raster_epsilon(eps=1e-8):
    """kEpsilon of the rasterizer (PyTorch3D geometry_utils: barycentric denominator + eps, degenerate face / edge tests).
    1e-8 reproduces the release the reference pins (0.3.0); 1e-30 the releases before 0.2.  Process-wide; returns the old value."""
    old = float(_lib.lib().acfm_get_raster_epsilon())
    _lib.check(_lib.lib().acfm_set_raster_epsilon(float(eps)), "acfm_set_raster_epsilon")
    return old


def rasterize(ndc, faces, image_size, blur_radius, faces_per_pixel, clip_barycentric_coords=False,
              cull_backfaces=False, sigma=0.0, want_bary=False, want_mask=False, want_vis=False):
    """rasterize_meshes on screen-space verts (no autograd).  Returns dict of pix_to_face / zbuf / dists
    [/ bary / mask / vis].  vis (N,V): vertices of the faces nearest at some pixel, a by-product of the render that
    loss_utils.bds_loss / optical_flow_loss otherwise recompute from pix_to_face[..., 0] (a strided 1 GB read at C3)."""
    _lib.require_cuda(ndc, faces)
    ndc = _f32c(ndc)
    N, V, _ = ndc.shape
    H = W = int(image_size)
    K = int(faces_per_pixel)
    fa, i64, fstride, F = _faces_arg(faces, N)
    dev = ndc.device
    p2f = torch.empty((N, H, W, K), dtype=torch.int64, device=dev)
    zbuf = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
    dists = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
    bary = torch.empty((N, H, W, K, 3), dtype=torch.float32, device=dev) if want_bary else None
    mask = torch.empty((N, H, W), dtype=torch.float32, device=dev) if want_mask else None
    vis = torch.empty((N, V), dtype=torch.float32, device=dev) if want_vis else None
    # scratch of the split path (region work lists): the empty regions are padded by a second, concurrent kernel
    ws_bytes = int(_lib.lib().acfm_raster_fwd_workspace_bytes(N, H, W)) if SPLIT_FILL else 0
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
    if _lib.event_hook is not None:
        _lib.event_hook("raster_fwd", 0)
    with torch.cuda.device(dev):
        st = _lib.lib().acfm_raster_fwd(_lib.ptr(ndc), _lib.ptr(fa), i64, fstride, N, V, F, H, W, K,
                                        float(blur_radius), int(clip_barycentric_coords), int(cull_backfaces),
                                        float(sigma), _lib.ptr(p2f), _lib.ptr(zbuf), _lib.ptr(dists), _lib.ptr(bary),
                                        _lib.ptr(mask), _lib.ptr(vis), _lib.ptr(ws), ws_bytes, _lib.stream_of(ndc))
    _lib.check(st, "acfm_raster_fwd")
    _lib.count((2 if ws is not None else 1) + (1 if vis is not None else 0))
    if _lib.event_hook is not None:
        _lib.event_hook("raster_fwd", 1)
    return dict(pix_to_face=p2f, zbuf=zbuf, dists=dists, bary=bary, mask=mask, vis=vis, work=ws)


def _train_render(ndc, faces, image_size, blur_radius, K, sigma, want_vis, target=None, edt=None):
    """acfm_raster_fwd_train: the render a training step differentiates — fragments, mask, optionally the visible-vertex map
    and the fused mask-loss sums.  Returns a dict."""
    _lib.require_cuda(ndc, faces, target, edt)
    ndc = _f32c(ndc)
    N, V, _ = ndc.shape
    H = W = int(image_size)
    fa, i64, fstride, F = _faces_arg(faces, N)
    dev = ndc.device
    NB = 0
    if target is not None:
        target = _f32c(target)
        edt = _f32c(edt) if edt is not None else None
        NB = target.shape[0]
        if N and (NB == 0 or N % NB or target.numel() != NB * H * W or (edt is not None and edt.numel() != target.numel())):
            raise ValueError(f"renders {N} vs target {tuple(target.shape)}: the batch must divide, the pixels must match")
    p2f = torch.empty((N, H, W, K), dtype=torch.int64, device=dev)
    zbuf = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
    dists = torch.empty((N, H, W, K), dtype=torch.float32, device=dev)
    mask = torch.empty((N, H, W), dtype=torch.float32, device=dev)
    sums = torch.empty((N, 4), dtype=torch.float32, device=dev) if target is not None else None
    vis = torch.empty((N, V), dtype=torch.float32, device=dev) if want_vis else None
    L = _lib.lib()
    ws_bytes = int(L.acfm_raster_fwd_workspace_bytes(N, H, W)) if SPLIT_FILL else 0
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) if ws_bytes else None
    lw_bytes = int(L.acfm_raster_loss_workspace_bytes(N, max(NB, 1), H, W)) if target is not None else 0
    lw = torch.empty((max(lw_bytes, 16),), dtype=torch.uint8, device=dev) if target is not None else None
    if _lib.event_hook is not None:
        _lib.event_hook("raster_fwd", 0)
    if N:
        with torch.cuda.device(dev):
            st = L.acfm_raster_fwd_train(_lib.ptr(ndc), _lib.ptr(fa), i64, fstride, N, V, F, H, W, K, float(blur_radius), float(sigma),
                                         _lib.ptr(p2f), _lib.ptr(zbuf), _lib.ptr(dists), _lib.ptr(mask), _lib.ptr(vis),
                                         _lib.ptr(target), _lib.ptr(edt), NB, _lib.ptr(sums), _lib.ptr(lw), lw_bytes, _lib.ptr(ws),
                                         ws_bytes, _lib.stream_of(ndc))
        _lib.check(st, "acfm_raster_fwd_train")
        # prep + padding + rasterizer (or the one kernel), + target base and reduce for the fused sums, + the visibility memset
        _lib.count((3 if ws is not None else 1) + (2 if sums is not None else 0) + (1 if vis is not None else 0))
    if _lib.event_hook is not None:
        _lib.event_hook("raster_fwd", 1)
    return dict(ndc=ndc, pix_to_face=p2f, zbuf=zbuf, dists=dists, mask=mask, sums=sums, vis=vis, work=ws, target=target, edt=edt)


def _train_render_bwd(saved, cfg, work, grad_mask, grad_sums):
    ndc, faces, p2f, dists, mask, target, edt = saved
    S, K, sigma = cfg
    N, V, _ = ndc.shape
    fa, i64, fstride, F = _faces_arg(faces, N)
    grad_mask = _f32c(grad_mask) if grad_mask is not None else None
    grad_sums = _f32c(grad_sums) if grad_sums is not None else None
    g = torch.empty_like(ndc)
    if _lib.event_hook is not None:
        _lib.event_hook("raster_bwd", 0)
    with torch.cuda.device(ndc.device):
        st = _lib.lib().acfm_raster_soft_bwd_train(_lib.ptr(ndc), _lib.ptr(fa), i64, fstride, N, V, F, S, S, K, sigma, _lib.ptr(p2f),
                                                   _lib.ptr(dists), _lib.ptr(mask), _lib.ptr(grad_mask),
                                                   _lib.ptr(grad_sums), _lib.ptr(target), _lib.ptr(edt),
                                                   target.shape[0] if target is not None else 1, _lib.ptr(g), _lib.ptr(work),
                                                   _lib.stream_of(ndc))
    _lib.check(st, "acfm_raster_soft_bwd_train")
    _lib.count(2)  # memset + kernel
    if _lib.event_hook is not None:
        _lib.event_hook("raster_bwd", 1)
    return g


class _SoftSilhouette(torch.autograd.Function):
    """ndc (N,V,3) -> mask (N,H,W), pix_to_face, zbuf, dists [, vis]; differentiable in ndc through dists."""

    @staticmethod
    def forward(ctx, ndc, faces, image_size, blur_radius, K, sigma, want_vis=False):
        fr = _train_render(ndc, faces, image_size, blur_radius, K, sigma, want_vis)
        ctx.save_for_backward(fr["ndc"], faces, fr["pix_to_face"], fr["dists"], fr["mask"])
        ctx.cfg = (int(image_size), int(K), float(sigma))
        ctx.work = fr["work"]   # the forward's region work lists: the backward visits the live regions only, heaviest first
        ctx.mark_non_differentiable(fr["pix_to_face"], fr["zbuf"], fr["dists"])
        ctx.set_materialize_grads(False)
        if want_vis:
            ctx.mark_non_differentiable(fr["vis"])
            return fr["mask"], fr["pix_to_face"], fr["zbuf"], fr["dists"], fr["vis"]
        return fr["mask"], fr["pix_to_face"], fr["zbuf"], fr["dists"]

    @staticmethod
    def backward(ctx, grad_mask, _g1, _g2, _g3, _g4=None):
        if grad_mask is None or ctx.saved_tensors[0].shape[0] == 0:
            return None, None, None, None, None, None, None
        g = _train_render_bwd(ctx.saved_tensors + (None, None), ctx.cfg, ctx.work, grad_mask, None)
        return g, None, None, None, None, None, None


def soft_silhouette(ndc, faces, image_size, blur_radius=BLUR_SOFT, faces_per_pixel=K_SOFT, sigma=SIGMA, want_vis=False):
    """-> mask, pix_to_face, zbuf, dists [, vis].  want_vis: the render also marks the visible vertices — (N,V) 0/1 floats,
    what bds_loss / optical_flow_loss derive from pix_to_face[..., 0] — returned as a fifth tensor (pass it to the losses as
    `visible=`)."""
    return _SoftSilhouette.apply(ndc, faces, int(image_size), float(blur_radius), int(faces_per_pixel), float(sigma), bool(want_vis))


class _SoftSilhouetteLosses(torch.autograd.Function):
    """The soft-silhouette render with the per-render mask-loss sums fused in (acfm_raster_fwd_train / _soft_bwd_train):
    ndc (N,V,3), target (NB,H,W), edt (NB,H,W) or None -> mask, pix_to_face, zbuf, dists, sums (N,4) [, vis].
    Differentiable in ndc through BOTH the mask and the sums; d loss / d mask of the sums is formed inside the rasterizer
    backward and never materialised."""

    @staticmethod
    def forward(ctx, ndc, faces, target, edt, image_size, blur_radius, K, sigma, want_vis):
        fr = _train_render(ndc, faces, image_size, blur_radius, K, sigma, want_vis, target, edt)
        ctx.save_for_backward(fr["ndc"], faces, fr["pix_to_face"], fr["dists"], fr["mask"], fr["target"], fr["edt"])
        ctx.cfg = (int(image_size), int(K), float(sigma))
        ctx.work = fr["work"]
        ctx.mark_non_differentiable(fr["pix_to_face"], fr["zbuf"], fr["dists"])
        ctx.set_materialize_grads(False)
        if want_vis:
            ctx.mark_non_differentiable(fr["vis"])
            return fr["mask"], fr["pix_to_face"], fr["zbuf"], fr["dists"], fr["sums"], fr["vis"]
        return fr["mask"], fr["pix_to_face"], fr["zbuf"], fr["dists"], fr["sums"]

    @staticmethod
    def backward(ctx, grad_mask, _g1, _g2, _g3, grad_sums, _g5=None):
        none = (None,) * 9
        if (grad_mask is None and grad_sums is None) or ctx.saved_tensors[0].shape[0] == 0:
            return none
        g = _train_render_bwd(ctx.saved_tensors, ctx.cfg, ctx.work, grad_mask, grad_sums)
        return (g,) + none[1:]


def soft_silhouette_losses(ndc, faces, image_size, target, edt=None, blur_radius=BLUR_SOFT, faces_per_pixel=K_SOFT, sigma=SIGMA,
                           want_vis=False):
    """-> mask, pix_to_face, zbuf, dists, sums (N,4) [, vis]; sums = {sum|m-t|, sum m t, sum (m+t-mt), sum edt m} per render
    (loss_utils.losses_from_sums turns them into l1 / iou / edt losses)."""
    return _SoftSilhouetteLosses.apply(ndc, faces, target, edt, int(image_size), float(blur_radius), int(faces_per_pixel),
                                       float(sigma), bool(want_vis))


# ---- lean training mode (not API parity; bench.py reports it separately) -----------------------------------------------------
class _SoftSilhouetteLean(torch.autograd.Function):
    """acfm_raster_fwd_lean / acfm_raster_soft_bwd_lean: mask, the fused loss sums and the visible-vertex map WITHOUT the (N,H,W,K)
    fragment tensors — the fragments of the regions the mesh touches stay in a compact scratch between forward and backward."""

    @staticmethod
    def forward(ctx, ndc, faces, target, edt, image_size, blur_radius, K, sigma, want_vis):
        _lib.require_cuda(ndc, faces, target, edt)
        ndc = _f32c(ndc)
        N, V, _ = ndc.shape
        H = W = int(image_size)
        fa, i64, fstride, F = _faces_arg(faces, N)
        dev = ndc.device
        NB = 0
        if target is not None:
            target = _f32c(target)
            edt = _f32c(edt) if edt is not None else None
            NB = target.shape[0]
            if N and (NB == 0 or N % NB or target.numel() != NB * H * W or (edt is not None and edt.numel() != target.numel())):
                raise ValueError(f"renders {N} vs target {tuple(target.shape)}: the batch must divide, the pixels must match")
        L = _lib.lib()
        mask = torch.empty((N, H, W), dtype=torch.float32, device=dev)
        sums = torch.empty((N, 4), dtype=torch.float32, device=dev) if target is not None else None
        vis = torch.empty((N, V), dtype=torch.float32, device=dev) if want_vis else None
        ws = torch.empty((max(int(L.acfm_raster_fwd_workspace_bytes(N, H, W)), 16),), dtype=torch.uint8, device=dev)
        lean = torch.empty((max(int(L.acfm_raster_lean_workspace_bytes(N, H, W, K)), 16),), dtype=torch.uint8, device=dev)
        lw_bytes = int(L.acfm_raster_loss_workspace_bytes(N, max(NB, 1), H, W)) if target is not None else 0
        lw = torch.empty((max(lw_bytes, 16),), dtype=torch.uint8, device=dev) if target is not None else None
        if N:
            with torch.cuda.device(dev):
                st = L.acfm_raster_fwd_lean(_lib.ptr(ndc), _lib.ptr(fa), i64, fstride, N, V, F, H, W, K, float(blur_radius), float(sigma),
                                            _lib.ptr(mask), _lib.ptr(vis), _lib.ptr(target), _lib.ptr(edt), NB, _lib.ptr(sums), _lib.ptr(lw),
                                            lw_bytes, _lib.ptr(lean), lean.numel(), _lib.ptr(ws), ws.numel(), _lib.stream_of(ndc))
            _lib.check(st, "acfm_raster_fwd_lean")
            _lib.count(3 + (2 if sums is not None else 0) + (1 if vis is not None else 0))   # mask memset, prep, rasterizer, ...
        ctx.save_for_backward(ndc, faces, mask, target, edt)
        ctx.cfg = (H, int(K), float(sigma))
        ctx.scratch = (lean, ws)
        ctx.set_materialize_grads(False)
        outs = (mask,) + ((sums,) if sums is not None else ()) + ((vis,) if vis is not None else ())
        if vis is not None:
            ctx.mark_non_differentiable(vis)
        ctx.has_sums = sums is not None
        return outs if len(outs) > 1 else mask

    @staticmethod
    def backward(ctx, grad_mask, *rest):
        ndc, faces, mask, target, edt = ctx.saved_tensors
        grad_sums = rest[0] if (ctx.has_sums and rest) else None
        none = (None,) * 9
        N, V, _ = ndc.shape
        if (grad_mask is None and grad_sums is None) or N == 0:
            return none
        S, K, sigma = ctx.cfg
        lean, ws = ctx.scratch
        fa, i64, fstride, F = _faces_arg(faces, N)
        grad_mask = _f32c(grad_mask) if grad_mask is not None else None
        grad_sums = _f32c(grad_sums) if grad_sums is not None else None
        g = torch.empty_like(ndc)
        with torch.cuda.device(ndc.device):
            st = _lib.lib().acfm_raster_soft_bwd_lean(_lib.ptr(ndc), _lib.ptr(fa), i64, fstride, N, V, F, S, S, K, sigma, _lib.ptr(mask),
                                                      _lib.ptr(grad_mask), _lib.ptr(grad_sums), _lib.ptr(target), _lib.ptr(edt),
                                                      target.shape[0] if target is not None else 1, _lib.ptr(g), _lib.ptr(lean), _lib.ptr(ws),
                                                      _lib.stream_of(ndc))
        _lib.check(st, "acfm_raster_soft_bwd_lean")
        _lib.count(2)
        return (g,) + none[1:]


def soft_silhouette_lean(ndc, faces, image_size, target=None, edt=None, blur_radius=BLUR_SOFT, faces_per_pixel=K_SOFT, sigma=SIGMA,
                         want_vis=False):
    """Lean training render: -> mask [, sums (N,4) if a target is given] [, vis].  No pix_to_face / zbuf / dists: use it where the
    step consumes only the silhouette, its losses and the visible vertices (the reference's monocular step); same values and
    gradients as soft_silhouette_losses.  K = 20 only."""
    return _SoftSilhouetteLean.apply(ndc, faces, target, edt, int(image_size), float(blur_radius), int(faces_per_pixel), float(sigma),
                                     bool(want_vis))
