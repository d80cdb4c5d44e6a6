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


def losses_from_sums(s, pixels, with_edt=True):
    """(N,4) sums of mask_sums() or of NeuralRenderer.forward_with_losses -> dict(l1, iou_loss[, edt]) each (N,):
    l1_loss / iou_loss / edt_loss with reduce=False (loss_utils.py:18-32,72-77,245-253)."""
    hw = float(pixels)
    out = dict(l1=s[:, 0] / hw, iou_loss=1 - s[:, 1] / (s[:, 2] + 1e-6))
    if with_edt:
        out["edt"] = s[:, 3] / hw
    return out


class _CombineMaskLosses(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sums, pixels, w_l1, w_iou, w_edt):
        _lib.require_cuda(sums)
        sums = F_._f32c(sums)
        N = sums.shape[0]
        per = torch.empty((N,), dtype=torch.float32, device=sums.device)
        with torch.cuda.device(sums.device):
            st = _lib.lib().acfm_mask_loss_combine_fwd(_lib.ptr(sums), N, pixels, w_l1, w_iou, w_edt, _lib.ptr(per), _lib.stream_of(sums))
        _lib.check(st, "acfm_mask_loss_combine_fwd")
        _lib.count()
        ctx.save_for_backward(sums)
        ctx.cfg = (pixels, w_l1, w_iou, w_edt)
        return per

    @staticmethod
    def backward(ctx, grad_per):
        sums, = ctx.saved_tensors
        g = torch.empty_like(sums)
        with torch.cuda.device(sums.device):
            st = _lib.lib().acfm_mask_loss_combine_bwd(_lib.ptr(sums), _lib.ptr(F_._f32c(grad_per)), sums.shape[0], *ctx.cfg, _lib.ptr(g),
                                                       _lib.stream_of(sums))
        _lib.check(st, "acfm_mask_loss_combine_bwd")
        _lib.count()
        return g, None, None, None, None


def combine_mask_losses(sums, pixels, w_l1=1.0, w_iou=0.0, w_edt=0.0):
    """(N,4) sums -> (N,) w_l1 * l1_loss + w_iou * iou_loss + w_edt * edt_loss per render (reduce=False), the weighted sum the
    callers form from the three losses (multiframe/main.py:644-645,715-716), in one kernel each way instead of a dozen
    element-wise torch launches on N numbers."""
    return _CombineMaskLosses.apply(sums, int(pixels), float(w_l1), float(w_iou), float(w_edt))


def mask_losses(mask, target, edt=None):
    """All per-render silhouette losses in one pass over a rendered mask: dict(l1, iou_loss, edt) each (N,)."""
    return losses_from_sums(mask_sums(mask, target, edt), mask[0].numel() if mask.shape[0] else 1, edt is not None)


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


class _KpLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kp_pred, kp_gt):
        _lib.require_cuda(kp_pred, kp_gt)
        kp_pred, kp_gt = F_._f32c(kp_pred), F_._f32c(kp_gt)
        N, Kp, ps = kp_pred.shape
        NB = kp_gt.shape[0]
        if kp_gt.shape[1:] != (Kp, 3) or ps < 2 or (N and (NB == 0 or N % NB)):
            raise ValueError(f"kp_pred {tuple(kp_pred.shape)} vs kp_gt {tuple(kp_gt.shape)}")
        loss = torch.empty((N,), dtype=torch.float32, device=kp_pred.device)
        with torch.cuda.device(kp_pred.device):
            st = _lib.lib().acfm_kp_loss_fwd(_lib.ptr(kp_pred), ps, _lib.ptr(kp_gt), N, max(NB, 1), Kp, _lib.ptr(loss),
                                             _lib.stream_of(kp_pred))
        _lib.check(st, "acfm_kp_loss_fwd")
        _lib.count()
        ctx.save_for_backward(kp_pred, kp_gt)
        return loss

    @staticmethod
    def backward(ctx, g):
        kp_pred, kp_gt = ctx.saved_tensors
        N, Kp, ps = kp_pred.shape
        out = torch.empty_like(kp_pred)
        with torch.cuda.device(kp_pred.device):
            st = _lib.lib().acfm_kp_loss_bwd(_lib.ptr(kp_pred), ps, _lib.ptr(kp_gt), _lib.ptr(F_._f32c(g)), N,
                                             max(kp_gt.shape[0], 1), Kp, _lib.ptr(out), _lib.stream_of(kp_pred))
        _lib.check(st, "acfm_kp_loss_bwd")
        _lib.count()
        return out, None


def kp_l2_loss(kp_pred, kp_gt, reduction='mean'):
    """loss_utils.py:341-356 (an L1 over visible keypoints, despite the name).  kp_gt may carry NB <= N rows."""
    loss = _KpLoss.apply(kp_pred, kp_gt)
    return loss.mean() if reduction == 'mean' else loss


def visible_vertices(pix_to_face, faces, num_verts):
    """(N,V) 0/1 floats: vertices of the faces that are nearest at some pixel — the fi_maps/unique/scatter_ block of
    bds_loss (loss_utils.py:213-223) and optical_flow_loss (:432-441).  pix_to_face (N,H,W,K) int64 packed ids."""
    _lib.require_cuda(pix_to_face, faces)
    if pix_to_face.dtype != torch.int64:
        pix_to_face = pix_to_face.long()
    pix_to_face = pix_to_face.contiguous()
    N, H, W, K = pix_to_face.shape
    fa, i64, fstride, F = F_._faces_arg(faces, N)
    vis = torch.empty((N, num_verts), dtype=torch.float32, device=pix_to_face.device)
    with torch.cuda.device(vis.device):
        st = _lib.lib().acfm_visible_verts(_lib.ptr(pix_to_face), K, _lib.ptr(fa), i64, fstride, N, num_verts, F, H * W,
                                           _lib.ptr(vis), _lib.stream_of(vis))
    _lib.check(st, "acfm_visible_verts")
    _lib.count(2)
    return vis


class _BdsLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, vis, bds, sel):
        verts, bds = F_._f32c(verts), F_._f32c(bds)
        N, V, vs = verts.shape
        NB, P, _ = bds.shape
        S = sel.numel()
        loss = torch.empty((N,), dtype=torch.float32, device=verts.device)
        argmin = torch.empty((N, S), dtype=torch.int32, device=verts.device)
        with torch.cuda.device(verts.device):
            st = _lib.lib().acfm_bds_loss_fwd(_lib.ptr(verts), vs, _lib.ptr(vis), _lib.ptr(bds), _lib.ptr(sel), N, max(NB, 1), V,
                                              P, S, _lib.ptr(loss), _lib.ptr(argmin), _lib.stream_of(verts))
        _lib.check(st, "acfm_bds_loss_fwd")
        _lib.count(2)
        ctx.save_for_backward(verts, bds, sel, argmin)
        return loss

    @staticmethod
    def backward(ctx, g):
        verts, bds, sel, argmin = ctx.saved_tensors
        N, V, vs = verts.shape
        NB, P, _ = bds.shape
        gv = torch.empty_like(verts)
        with torch.cuda.device(verts.device):
            st = _lib.lib().acfm_bds_loss_bwd(_lib.ptr(verts), vs, _lib.ptr(bds), _lib.ptr(sel), _lib.ptr(argmin),
                                              _lib.ptr(F_._f32c(g)), N, max(NB, 1), V, P, sel.numel(), _lib.ptr(gv),
                                              _lib.stream_of(verts))
        _lib.check(st, "acfm_bds_loss_bwd")
        _lib.count(2)
        return gv, None, None, None


def bds_loss(verts, bds, faces, pix_to_face, reduce=True, n_samples=1000, k=1, indices=None, visible=None):
    """loss_utils.py:204-237.  verts (N,V,2|3) projected vertices, bds (NB|N,P,3) boundary points [x,y,mask],
    faces (N,F,3), pix_to_face (N,H,W,K).  Like the reference, up to n_samples boundary points are drawn with
    torch.randperm from the default (CPU) generator on every call — unless `indices` (int64, on the device) is given,
    which keeps the call free of host work (CUDA-graph capture, predictor.PostOptimizer).
    visible (N,V): the visible-vertex map when the render already produced it (NeuralRenderer.forward_with_visibility /
    forward_with_losses); otherwise it is derived from pix_to_face[..., 0] as the reference does (same values)."""
    if k != 1:
        raise ValueError("bds_loss: only k=1 (the reference's only call) is implemented")
    _lib.require_cuda(verts, bds, faces, pix_to_face)
    if indices is None:
        indices = torch.randperm(bds.shape[1])[:n_samples].to(verts.device)
    vis = visible if visible is not None else visible_vertices(pix_to_face, faces, verts.shape[1])
    if vis.shape != (verts.shape[0], verts.shape[1]):
        raise ValueError(f"visible must be (N,V) = {(verts.shape[0], verts.shape[1])}, got {tuple(vis.shape)}")
    loss = _BdsLoss.apply(verts, F_._f32c(vis), bds, indices.contiguous())
    return loss.mean() if reduce else loss


class Boundaries_Loss(torch.nn.Module):
    """loss_utils.py:240-242"""

    def forward(self, verts, bds, faces, pix_to_face, reduce=True, n_samples=1000, visible=None):
        return bds_loss(verts, bds, faces, pix_to_face, reduce=reduce, n_samples=n_samples, visible=visible)


class _OfLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proj, vis, flows, B, T):
        proj, flows = F_._f32c(proj), F_._f32c(flows)
        BT, V, ps = proj.shape
        NBf, H, W = flows.shape[0] // T, flows.shape[1], flows.shape[2]
        dev = proj.device
        loss = torch.zeros((B, T - 1), dtype=torch.float32, device=dev)
        of_pred = torch.empty((B, T - 1, V, 2), dtype=torch.float32, device=dev)
        samples = torch.empty((B, T - 1, V, 2), dtype=torch.float32, device=dev)
        vis_out = torch.empty((B, T - 1, V), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.lib().acfm_of_loss_fwd(_lib.ptr(proj), ps, _lib.ptr(vis), _lib.ptr(flows), B, max(NBf, 1), T, V, H, W,
                                             _lib.ptr(loss), _lib.ptr(of_pred), _lib.ptr(vis_out), _lib.ptr(samples),
                                             _lib.stream_of(proj))
        _lib.check(st, "acfm_of_loss_fwd")
        _lib.count()
        ctx.save_for_backward(vis_out, of_pred, samples)
        ctx.cfg = (ps, B, T, V, H, W)
        ctx.mark_non_differentiable(vis_out, samples)
        return loss, of_pred, vis_out, samples

    @staticmethod
    def backward(ctx, g, g_pred, _g2, _g3):
        vis_out, of_pred, samples = ctx.saved_tensors
        ps, B, T, V, H, W = ctx.cfg
        gp = torch.empty((B * T, V, ps), dtype=torch.float32, device=vis_out.device)
        g = F_._f32c(g) if g is not None else torch.zeros((B, T - 1), dtype=torch.float32, device=vis_out.device)
        with torch.cuda.device(vis_out.device):
            st = _lib.lib().acfm_of_loss_bwd(_lib.ptr(vis_out), _lib.ptr(of_pred), _lib.ptr(samples), _lib.ptr(g), ps, B, T, V,
                                             H, W, _lib.ptr(gp), _lib.stream_of(vis_out))
        _lib.check(st, "acfm_of_loss_bwd")
        _lib.count(2)
        if g_pred is not None:  # of_pred = W/2 * vis * (p_t - p_{t+1}) is also returned; rarely differentiated
            w = 0.5 * W * vis_out[..., None] * g_pred
            gp4 = gp.view(B, T, V, ps)
            gp4[:, :-1, :, :2] += w
            gp4[:, 1:, :, :2] -= w
        return gp, None, None, None, None


def optical_flow_loss(meshes, faces, cams, flows, renderer, pix_to_face, reduce=True, visible=None):
    """loss_utils.py:419-474.  meshes (B,T,V,3), faces (B,T,F,3), cams (B*T,7), flows (B|B/G,T,H,W,2);
    renderer: an OF_NeuralRenderer (its proj_fn projects, its forward gives the K=1 visibility render when
    pix_to_face is None).  visible (B*T,V): the visible-vertex map, if a render already produced it.
    Returns (loss, of_pred, visible_vertices, predicted_points, samples_ofs_gt) as the reference does."""
    _lib.require_cuda(meshes, faces, cams, flows)
    b, t, nv, _ = meshes.shape
    bt = b * t
    predicted_points = renderer.proj_fn(meshes.reshape(bt, nv, -1), cams.reshape(bt, -1))
    faces_bt = faces.reshape(bt, faces.shape[2], 3)
    with torch.no_grad():
        if visible is not None:
            vis = F_._f32c(visible).reshape(bt, nv)
        elif pix_to_face is None and hasattr(renderer, "forward_with_visibility"):
            _, vis = renderer.forward_with_visibility(predicted_points.reshape(bt, nv, 3), faces_bt)   # the K = 1 render marks them itself
        else:
            if pix_to_face is None:
                pix_to_face = renderer(predicted_points.reshape(bt, nv, 3), faces_bt)
            vis = visible_vertices(pix_to_face[..., :1], faces_bt, nv)
    flows_bt = flows.reshape(-1, flows.shape[2], flows.shape[3], flows.shape[4])
    loss, of_pred, vis_out, samples = _OfLoss.apply(predicted_points, vis, flows_bt, b, t)
    if reduce:
        loss = loss.sum()
    return loss, of_pred, vis_out, predicted_points[:, :, :2].reshape(b, t, nv, 2), samples


class Optical_Flow_Loss(torch.nn.Module):
    """loss_utils.py:477-479"""

    def forward(self, meshes, faces, cams, flows, renderer, pix_to_face, reduce=True, visible=None):
        return optical_flow_loss(meshes, faces, cams, flows, renderer, pix_to_face, reduce=reduce, visible=visible)


class _HypWeight(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loss):
        _lib.require_cuda(loss)
        loss = F_._f32c(loss)
        G, M = loss.shape
        probs = torch.empty_like(loss)
        total = torch.empty((1,), dtype=torch.float32, device=loss.device)
        with torch.cuda.device(loss.device):
            st = _lib.lib().acfm_hypothesis_weight_fwd(_lib.ptr(loss), G, M, _lib.ptr(probs), _lib.ptr(total), _lib.stream_of(loss))
        _lib.check(st, "acfm_hypothesis_weight_fwd")
        _lib.count(2)
        ctx.save_for_backward(probs)
        ctx.mark_non_differentiable(probs)
        return total[0], probs

    @staticmethod
    def backward(ctx, g, _gp):
        probs, = ctx.saved_tensors
        G, M = probs.shape
        out = torch.empty_like(probs)
        g = F_._f32c(g.reshape(1))
        with torch.cuda.device(probs.device):
            st = _lib.lib().acfm_hypothesis_weight_bwd(_lib.ptr(probs), _lib.ptr(g), G, M, _lib.ptr(out), _lib.stream_of(probs))
        _lib.check(st, "acfm_hypothesis_weight_bwd")
        _lib.count()
        return out


def hypothesis_weighting(total_loss):
    """total_loss (G, B*T) per hypothesis and frame -> (scalar, probs): probs = softmax(-total_loss, dim=0).detach();
    scalar = (total_loss * probs).sum(0).mean()  (multiframe/main.py:735-746)."""
    return _HypWeight.apply(total_loss)


# -------------------------------------------------------------------------------------------------
# shape priors on the deformed meshes
# -------------------------------------------------------------------------------------------------
def mesh_edges(faces):
    """(E,2) int64 unique undirected edges of one mesh's faces (F,3) — Meshes.edges_packed() for a single topology."""
    f = faces[0] if faces.dim() == 3 else faces
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], 0).long()
    e = torch.sort(e, dim=1)[0]
    return torch.unique(e, dim=0)


def _verts_faces(meshes, faces):
    if faces is None and hasattr(meshes, "verts_padded"):
        return meshes.verts_padded(), meshes.faces_padded()
    return meshes, faces


class _LapSmooth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, faces):
        _lib.require_cuda(verts, faces)
        verts = F_._f32c(verts)
        N, V, _ = verts.shape
        fa, i64, fstride, F = F_._faces_arg(faces, N)
        dev = verts.device
        loss = torch.empty((N,), dtype=torch.float32, device=dev)
        unit = torch.empty((N, V, 4), dtype=torch.float32, device=dev)
        ws = torch.empty((N, V, 4), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.lib().acfm_laplacian_smoothing_fwd(_lib.ptr(verts), _lib.ptr(fa), i64, fstride, N, V, F, _lib.ptr(loss),
                                                         _lib.ptr(unit), _lib.ptr(ws), _lib.stream_of(verts))
        _lib.check(st, "acfm_laplacian_smoothing_fwd")
        _lib.count(4)
        ctx.save_for_backward(verts, fa, unit)
        ctx.cfg = (i64, fstride, F)
        return loss

    @staticmethod
    def backward(ctx, g):
        verts, fa, unit = ctx.saved_tensors
        i64, fstride, F = ctx.cfg
        N, V, _ = verts.shape
        gv = torch.empty_like(verts)
        with torch.cuda.device(verts.device):
            st = _lib.lib().acfm_laplacian_smoothing_bwd(_lib.ptr(verts), _lib.ptr(fa), i64, fstride, _lib.ptr(unit),
                                                         _lib.ptr(F_._f32c(g)), N, V, F, _lib.ptr(gv), _lib.stream_of(verts))
        _lib.check(st, "acfm_laplacian_smoothing_bwd")
        _lib.count(2)
        return gv, None


def mesh_laplacian_smoothing(meshes, faces=None, method="cot"):
    """pytorch3d.loss.mesh_laplacian_smoothing(meshes, method="cot") as the trainer calls it (multiframe/main.py:699-704).
    `meshes`: (N,V,3) vertices with `faces` (N|1,F,3), or an object with verts_padded() / faces_padded()."""
    if method != "cot":
        raise ValueError("mesh_laplacian_smoothing: only method='cot' (the reference's call) is implemented")
    verts, faces = _verts_faces(meshes, faces)
    per_mesh = _LapSmooth.apply(verts, faces)
    return per_mesh.sum() / max(verts.shape[0], 1)


class _Rigid(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, tmpl, edges):
        _lib.require_cuda(verts, tmpl, edges)
        verts, tmpl = F_._f32c(verts), F_._f32c(tmpl)
        if edges.dtype not in (torch.int64, torch.int32):
            raise ValueError("edges must be int64 or int32")
        edges = edges.contiguous()
        N, V, _ = verts.shape
        NT = tmpl.shape[0]
        if tmpl.shape[1:] != (V, 3) or NT == 0 or N % NT:
            raise ValueError(f"template {tuple(tmpl.shape)} does not broadcast over verts {tuple(verts.shape)}")
        loss = torch.empty((N,), dtype=torch.float32, device=verts.device)
        with torch.cuda.device(verts.device):
            st = _lib.lib().acfm_edge_rigidity_fwd(_lib.ptr(verts), _lib.ptr(tmpl), _lib.ptr(edges), int(edges.dtype == torch.int64),
                                                   N, NT, V, edges.shape[0], _lib.ptr(loss), _lib.stream_of(verts))
        _lib.check(st, "acfm_edge_rigidity_fwd")
        _lib.count(2)
        ctx.save_for_backward(verts, tmpl, edges)
        return loss

    @staticmethod
    def backward(ctx, g):
        verts, tmpl, edges = ctx.saved_tensors
        N, V, _ = verts.shape
        gv = torch.empty_like(verts)
        gt = torch.empty_like(tmpl) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(verts.device):
            st = _lib.lib().acfm_edge_rigidity_bwd(_lib.ptr(verts), _lib.ptr(tmpl), _lib.ptr(edges), int(edges.dtype == torch.int64),
                                                   _lib.ptr(F_._f32c(g)), N, tmpl.shape[0], V, edges.shape[0], _lib.ptr(gv),
                                                   _lib.ptr(gt), _lib.stream_of(verts))
        _lib.check(st, "acfm_edge_rigidity_bwd")
        _lib.count(3)
        return gv, gt, None


def locally_rigid_fn(meshes, mesh_template, edges=None):
    """loss_utils.py:150-164: sum over edges of (deformed length - template length)^2, / number of meshes.
    meshes / mesh_template: (N,V,3) / (NT|V,3) vertex tensors with `edges` (E,2) (see mesh_edges), or objects with
    verts_padded() and edges_packed() over ONE shared topology."""
    if hasattr(meshes, "verts_padded"):
        verts, tmpl = meshes.verts_padded(), mesh_template.verts_padded()
        if edges is None:
            edges = mesh_edges(meshes.faces_padded()[:1])
    else:
        verts, tmpl = meshes, mesh_template
    if tmpl.dim() == 2:
        tmpl = tmpl[None]
    return _Rigid.apply(verts, tmpl, edges).sum() / max(verts.shape[0], 1)


class Locally_Rigid(torch.nn.Module):
    """loss_utils.py:167-169"""

    def forward(self, meshes, mesh_template, edges=None):
        return locally_rigid_fn(meshes, mesh_template, edges)


def deform_l2reg(V):
    """loss_utils.py:322-327 (one torch reduction; kept for API completeness)"""
    return torch.mean(torch.norm(V.reshape(-1, V.shape[-1]), p=2, dim=1))
