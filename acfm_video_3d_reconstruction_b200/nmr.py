"""Drop-in for the reference's nnutils/nmr.py renderer wrappers
(/root/reference/multiframe/nnutils/nmr.py:54-238, /root/reference/monocular/nnutils/nmr.py:56-290).

`NeuralRenderer` / `OF_NeuralRenderer` keep the reference's constructor, attributes and forward
signatures; the PyTorch3D 0.3.0 objects they used to build per call are replaced by the sm_100a
kernels of libacfm_b200.so (projection -> tile-binned rasterizer -> fused blend, and their backward).
Modules hold no parameters or buffers and are re-entrant (DataParallel replicas, main.py:184-193).
"""
import torch

from . import functional as F_
from . import geom_utils


class NeuralRenderer(torch.nn.Module):
    """forward(vertices[N,V,3], faces[N,F,3], cams[N,7], textures=None, atlas=True)
         -> (masks[N,H,W], pix_to_face[N,H,W,20])                      if textures is None
         -> (imgs[N,3,H,W], sil[N,H,W], pix_to_face[N,H,W,1])          otherwise
    offset_z: 0.0 is the multiframe default (multiframe/nnutils/nmr.py:119); the monocular tree
    uses 5.0 (monocular/nnutils/nmr.py:164) — pass offset_z=5. or set the attribute."""

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
        # True: the mask render also marks the visible vertices (rides on the returned pix_to_face; bds_loss uses it
        # instead of re-reading pix_to_face[..., 0]).  Set it where the boundary loss follows the render (multiframe).
        self.emit_visibility = False

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

    def forward(self, vertices, faces, cams, textures=None, atlas=True):
        ndc = self.to_ndc(vertices, cams)
        if textures is None:
            self.mask_only = True
            masks, pix_to_face, _, _ = F_.soft_silhouette(ndc, faces, self.img_size, self.blur_radius,
                                                          self.faces_per_pixel, self.sigma, want_vis=self.emit_visibility)
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

    def forward(self, verts, faces):
        with torch.no_grad():
            # R = diag(-1,1,1), T = (0,0,2.732): exact sign flip, one rounding on z
            # no host-side constants here (keeps the call CUDA-graph capturable)
            ndc = torch.stack([-verts[..., 0], verts[..., 1], verts[..., 2] + F_.EYE_Z], dim=-1)
            fr = F_.rasterize(ndc, faces, self.img_size, 0.0, 1, want_vis=True)   # its only consumer needs the visibility
        return fr["pix_to_face"]
