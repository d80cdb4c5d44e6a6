"""Drop-in for the reference's nnutils/nmr.py renderer wrappers
(/root/reference/multiframe/nnutils/nmr.py:54-238, /root/reference/monocular/nnutils/nmr.py:56-290).

`NeuralRenderer` / `OF_NeuralRenderer` keep the reference's constructor, attributes and forward
signatures; the PyTorch3D 0.3.0 objects they used to build per call are replaced by the sm_100a
kernels of libacfm_b200.so (projection -> tile-binned rasterizer -> fused blend, and their backward).
Modules hold no parameters or buffers and are re-entrant (DataParallel replicas, main.py:184-193).
"""
import collections

import torch

from . import functional as F_
from . import geom_utils

# what PyTorch3D's MeshRasterizer returns (pytorch3d.renderer.mesh.rasterizer.Fragments): rasterize_of() hands it back
Fragments = collections.namedtuple("Fragments", ["pix_to_face", "zbuf", "bary_coords", "dists"])


class NeuralRenderer(torch.nn.Module):
    """forward(vertices[N,V,3], faces[N,F,3], cams[N,7], textures=None, atlas=True)
         -> (masks[N,H,W], pix_to_face[N,H,W,20])                      if textures is None
         -> (imgs[N,3,H,W], sil[N,H,W], pix_to_face[N,H,W,1])          otherwise
    offset_z: 0.0 is the multiframe default (multiframe/nnutils/nmr.py:119); the monocular tree
    uses 5.0 (monocular/nnutils/nmr.py:164): `acfm_video_3d_reconstruction_b200.monocular.NeuralRenderer` has that
    default, `...multiframe.NeuralRenderer` this one, so that each tree's `from nnutils.nmr import NeuralRenderer`
    can be pointed at its own module without touching a call site.

    Beyond the reference's interface (opt-in, same kernels): forward_with_visibility() also returns the visible-vertex map
    that bds_loss / optical_flow_loss need, forward_with_losses() also returns the per-render mask-loss sums fused into the
    render.  K = faces_per_pixel <= 64 (PyTorch3D allows 150; the reference uses 20 and 1)."""

    def __init__(self, img_size=256, offset_z=0.):
        super(NeuralRenderer, self).__init__()
        self.img_size = img_size
        self.proj_fn = geom_utils.orthographic_proj_withz
        self.offset_z = offset_z
        self.mask_only = True
        # hard-coded by the reference (nmr.py:144-159)
        self.sigma = F_.SIGMA
        self.blur_radius = F_.BLUR_SOFT
        self.faces_per_pixel = F_.K_SOFT

    def ambient_light_only(self):
        return

    def set_bgcolor(self, color):
        return

    def project_points(self, verts, cams):
        proj = self.proj_fn(verts, cams)
        return proj[:, :, :2]

    def to_ndc(self, vertices, cams):
        """proj_fn + `vs[:, :, 1] *= -1` + R=diag(-1,1,1), T=(0,0,2.732) (nmr.py:144-149)."""
        return F_.project(vertices, cams, offset_z=self.offset_z, sx=-1.0, sy=-1.0, z_add=F_.EYE_Z)

    def rasterize_of(self, verts, faces, R, T):
        """nmr.py:131-141 (defined by the reference, called by none of its paths): hard K = 1 rasterization of `verts` under
        the view X R + T (PyTorch3D's row-vector convention, identity SfM-orthographic projection) -> Fragments."""
        with torch.no_grad():
            view = torch.matmul(verts, R.to(verts.dtype)) + T.to(verts.dtype)[:, None, :]
            fr = F_.rasterize(view, faces, self.img_size, 0.0, 1, want_bary=True)
        return Fragments(fr["pix_to_face"], fr["zbuf"], fr["bary"], fr["dists"])

    def forward_with_visibility(self, vertices, faces, cams):
        """-> (masks, pix_to_face, visible[N,V]): the mask render that also marks the vertices of every pixel's nearest face —
        pass `visible=` to loss_utils.bds_loss / optical_flow_loss, which otherwise re-read pix_to_face[..., 0]."""
        self.mask_only = True
        masks, pix_to_face, _, _, vis = F_.soft_silhouette(self.to_ndc(vertices, cams), faces, self.img_size, self.blur_radius,
                                                           self.faces_per_pixel, self.sigma, want_vis=True)
        return masks, pix_to_face, vis

    def forward_with_losses(self, vertices, faces, cams, mask_gt, edt=None, want_visibility=False):
        """-> (masks, pix_to_face, sums[N,4][, visible]).  sums = {sum|m-t|, sum m t, sum (m+t-mt), sum edt m} per render,
        accumulated in the render's epilogue (loss_utils.losses_from_sums -> l1 / iou / edt losses); their backward runs
        inside the rasterizer backward, so neither a second pass over the mask nor grad_mask exists.  mask_gt / edt
        (NB,H,W) with NB | N: render n uses entry n % NB (the callers' .repeat(num_guesses, 1, 1), main.py:644,716)."""
        self.mask_only = True
        out = F_.soft_silhouette_losses(self.to_ndc(vertices, cams), faces, self.img_size,
                                        mask_gt.reshape(mask_gt.shape[0], self.img_size, self.img_size),
                                        None if edt is None else edt.reshape(edt.shape[0], self.img_size, self.img_size),
                                        self.blur_radius, self.faces_per_pixel, self.sigma, want_vis=want_visibility)
        return (out[0], out[1], out[4]) + ((out[5],) if want_visibility else ())

    def forward(self, vertices, faces, cams, textures=None, atlas=True):
        ndc = self.to_ndc(vertices, cams)
        if textures is None:
            self.mask_only = True
            masks, pix_to_face, _, _ = F_.soft_silhouette(ndc, faces, self.img_size, self.blur_radius,
                                                          self.faces_per_pixel, self.sigma)
            return masks, pix_to_face
        self.mask_only = False
        from . import texture
        return texture.render_textured(ndc, faces, textures, self.img_size, atlas=atlas)


class OF_NeuralRenderer(torch.nn.Module):
    """forward(verts[N,V,3] already projected, faces[N,F,3]) -> pix_to_face[N,H,W,1]
    (hard K=1 raster, no y-flip: multiframe/nnutils/nmr.py:208-238)."""

    def __init__(self, img_size=256):
        super(OF_NeuralRenderer, self).__init__()
        self.img_size = img_size
        self.proj_fn = geom_utils.orthographic_proj_withz
        self.offset_z = 5.

    def project_points(self, verts, cams):
        proj = self.proj_fn(verts, cams)
        return proj[:, :, :2]

    def _render(self, verts, faces, want_vis):
        with torch.no_grad():
            # R = diag(-1,1,1), T = (0,0,2.732): exact sign flip, one rounding on z
            # no host-side constants here (keeps the call CUDA-graph capturable)
            ndc = torch.stack([-verts[..., 0], verts[..., 1], verts[..., 2] + F_.EYE_Z], dim=-1)
            return F_.rasterize(ndc, faces, self.img_size, 0.0, 1, want_vis=want_vis)

    def forward(self, verts, faces):
        return self._render(verts, faces, False)["pix_to_face"]

    def forward_with_visibility(self, verts, faces):
        """-> (pix_to_face, visible[N,V]): the render's only consumer, optical_flow_loss, needs the visible vertices; the
        render marks them itself instead of leaving a strided re-read of pix_to_face to the loss."""
        fr = self._render(verts, faces, True)
        return fr["pix_to_face"], fr["vis"]
