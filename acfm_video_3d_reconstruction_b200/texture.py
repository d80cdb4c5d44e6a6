"""Texture branch of NeuralRenderer.forward and the UV texture sampler.

render_textured  — hard rasterization (blur 0, K=1, clip_barycentric_coords) + TexturesAtlas / vertex-colour
                   shading + softmax_rgb_blend (/root/reference/multiframe/nnutils/nmr.py:173-200).
uv_sample        — sampling step of TexturePredictorUV.forward
                   (/root/reference/multiframe/nnutils/mesh_net.py:166-179): bilinear grid_sample
                   (align_corners=True, zero padding) of the predicted UV image at the fixed per-face
                   T x T grid, (tanh + 1) / 2.
"""
import torch

from . import _lib
from . import functional as F_

# BlendParams(background_color=0) defaults (nmr.py:81) and softmax_rgb_blend's znear / zfar defaults
SIGMA_TEX, GAMMA_TEX, ZNEAR, ZFAR = 1e-4, 1e-4, 1.0, 100.0


class _Textured(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ndc, faces, tex, image_size, mode):
        fr = F_.rasterize(ndc, faces, image_size, 0.0, 1, clip_barycentric_coords=True, want_bary=True)
        ndc_c = F_._f32c(ndc)
        tex = F_._f32c(tex)
        N, V, _ = ndc_c.shape
        S = int(image_size)
        fa, i64, fstride, F = F_._faces_arg(faces, N)
        if mode == 0:
            if tex.dim() != 5 or tex.shape[0] != N or tex.shape[1] != F or tex.shape[2] != tex.shape[3] or tex.shape[4] != 3:
                raise ValueError(f"atlas textures must be (N,F,R,R,3) with N={N}, F={F}; got {tuple(tex.shape)}")
            R, NC = tex.shape[2], 1
        else:
            if tex.dim() != 3 or tex.shape[1] != V or tex.shape[2] != 3 or N % tex.shape[0]:
                raise ValueError(f"vertex textures must be (N|1,V,3) with V={V}; got {tuple(tex.shape)}")
            R, NC = 1, tex.shape[0]
        rgba = torch.empty((N, S, S, 4), dtype=torch.float32, device=ndc_c.device)
        with torch.cuda.device(ndc_c.device):
            st = _lib.lib().acfm_shade_fwd(_lib.ptr(fr["pix_to_face"]), _lib.ptr(fr["bary"]), _lib.ptr(fr["dists"]),
                                           _lib.ptr(fr["zbuf"]), N, S, S, 1, mode, _lib.ptr(tex), R, V, F, NC, _lib.ptr(fa),
                                           i64, fstride, SIGMA_TEX, GAMMA_TEX, ZNEAR, ZFAR, _lib.ptr(rgba),
                                           _lib.stream_of(ndc_c))
        _lib.check(st, "acfm_shade_fwd")
        _lib.count()
        ctx.save_for_backward(ndc_c, faces, tex, fr["pix_to_face"], fr["bary"], fr["dists"], fr["zbuf"])
        ctx.cfg = (S, mode, R, NC)
        ctx.mark_non_differentiable(fr["pix_to_face"])
        ctx.set_materialize_grads(False)
        return rgba, fr["pix_to_face"]

    @staticmethod
    def backward(ctx, g_rgba, _g):
        if g_rgba is None:
            return None, None, None, None, None
        ndc, faces, tex, p2f, bary, dists, zbuf = ctx.saved_tensors
        S, mode, R, NC = ctx.cfg
        N, V, _ = ndc.shape
        fa, i64, fstride, F = F_._faces_arg(faces, N)
        g_rgba = F_._f32c(g_rgba)
        g_tex = torch.empty_like(tex)
        need_v = ctx.needs_input_grad[0]
        g_d = torch.empty_like(dists) if need_v else None
        L = _lib.lib()
        with torch.cuda.device(ndc.device):
            st = L.acfm_shade_bwd(_lib.ptr(p2f), _lib.ptr(bary), _lib.ptr(dists), _lib.ptr(zbuf), N, S, S, 1, mode,
                                  _lib.ptr(tex), R, V, F, NC, _lib.ptr(fa), i64, fstride, SIGMA_TEX, GAMMA_TEX, ZNEAR, ZFAR,
                                  _lib.ptr(g_rgba), _lib.ptr(g_tex), g_tex.numel(), _lib.ptr(g_d), _lib.stream_of(ndc))
            _lib.check(st, "acfm_shade_bwd")
            _lib.count(2)
            g_ndc = None
            if need_v:
                g_ndc = torch.empty_like(ndc)
                st = L.acfm_raster_dists_bwd(_lib.ptr(ndc), _lib.ptr(fa), i64, fstride, N, V, F, S, S, 1, _lib.ptr(p2f),
                                             _lib.ptr(dists), _lib.ptr(g_d), _lib.ptr(g_ndc), _lib.stream_of(ndc))
                _lib.check(st, "acfm_raster_dists_bwd")
                _lib.count(2)
        return g_ndc, None, g_tex, None, None


def render_textured(ndc, faces, textures, image_size, atlas=True):
    """-> (imgs (N,3,H,W), sil (N,H,W), pix_to_face (N,H,W,1)) as NeuralRenderer.forward(textures=...)."""
    _lib.require_cuda(ndc, faces, textures)
    if atlas:
        mode, tex = 0, textures.to(ndc.device)
    else:
        mode, tex = 1, (textures[None] if textures.dim() == 2 else textures)
        if ndc.requires_grad and torch.is_grad_enabled():
            # vertex-colour mode interpolates with the barycentrics: PyTorch3D would also send grad_bary / grad_zbuf of the IMAGE
            # back to the vertices; here the vertices only receive the gradient that flows through dists (complete for `sil`,
            # partial for `imgs`).  The reference never differentiates this mode (bird_vis.VisRenderer: visualisation only), so
            # say so instead of being silently partial.
            import warnings
            warnings.warn("render_textured(atlas=False): d imgs / d vertices through the barycentric interpolation is not "
                          "propagated (the silhouette's gradient is complete); detach the vertices for visualisation", stacklevel=2)
    rgba, p2f = _Textured.apply(ndc, faces, tex, int(image_size), mode)
    imgs = rgba[..., :3].permute(0, 3, 1, 2)
    return imgs, rgba[..., 3], p2f


class _UVSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, uvimage, grid, apply_tanh):
        _lib.require_cuda(uvimage, grid)
        uvimage, grid = F_._f32c(uvimage), F_._f32c(grid)
        B, C, Hu, Wu = uvimage.shape
        P = grid.numel() // 2
        out = torch.empty((B, P, C), dtype=torch.float32, device=uvimage.device)
        with torch.cuda.device(uvimage.device):
            st = _lib.lib().acfm_uv_sample_fwd(_lib.ptr(uvimage), _lib.ptr(grid), B, C, Hu, Wu, P, int(apply_tanh),
                                               _lib.ptr(out), _lib.stream_of(uvimage))
        _lib.check(st, "acfm_uv_sample_fwd")
        _lib.count()
        ctx.save_for_backward(out, grid)
        ctx.cfg = (B, C, Hu, Wu, P, int(apply_tanh))
        return out

    @staticmethod
    def backward(ctx, g):
        out, grid = ctx.saved_tensors
        B, C, Hu, Wu, P, at = ctx.cfg
        gi = torch.empty((B, C, Hu, Wu), dtype=torch.float32, device=out.device)
        with torch.cuda.device(out.device):
            st = _lib.lib().acfm_uv_sample_bwd(_lib.ptr(out), _lib.ptr(grid), _lib.ptr(F_._f32c(g)), B, C, Hu, Wu, P, at,
                                               _lib.ptr(gi), _lib.stream_of(out))
        _lib.check(st, "acfm_uv_sample_bwd")
        _lib.count(2)
        return gi, None, None


def uv_sample(uvimage, uv_sampler, tex_size=None, apply_tanh=True):
    """uvimage (B,3,Hu,Wu), uv_sampler (1|., F, T*T, 2) or (F,T,T,2) in [-1,1]  ->  atlas (B,F,T,T,3) in [0,1]:
    the `tex_pred` of TexturePredictorUV.forward (mesh_net.py:169-172)."""
    g = uv_sampler
    if g.dim() == 4 and g.shape[0] == 1 and g.shape[-1] == 2 and tex_size is None:  # (1,F,T*T,2)
        F, TT = g.shape[1], g.shape[2]
        T = int(round(TT ** 0.5))
    else:
        F, T = g.shape[-4] if g.dim() >= 4 else g.shape[0], g.shape[-2]
        if g.dim() == 4 and g.shape[-1] == 2 and g.shape[1] == g.shape[2]:  # (F,T,T,2)
            F, T = g.shape[0], g.shape[1]
    out = _UVSample.apply(uvimage, g.reshape(-1, 2), apply_tanh)
    return out.view(uvimage.shape[0], F, T, T, uvimage.shape[1])
