"""oracle/pt3d_oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python face of the CPU oracle (oracle/raster_oracle.c/.inc): numpy in, numpy out.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.

PARITY UNPINNED for the rasterizer (PyTorch3D 0.3.0 is not in /root/reference and
not installable; algorithm restated per SURVEY.md §9).  The projection and the
losses ARE pinned against vectors generated from the reference's own
nnutils/geom_utils.py and nnutils/loss_utils.py (tests/golden/make_golden.py).

Reference call path restated here:
  NeuralRenderer.forward     /root/reference/multiframe/nnutils/nmr.py:143-200
                             /root/reference/monocular/nnutils/nmr.py:192-252
  OF_NeuralRenderer.forward  /root/reference/multiframe/nnutils/nmr.py:224-238
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libacfm_oracle.so")
_lib = None

# Constants hard-coded by the reference (multiframe/nnutils/nmr.py:144-159).
SIGMA = 1e-4
BLUR_SOFT = float(np.log(1.0 / 1e-4 - 1.0) * 1e-4)
K_SOFT = 20
EYE_Z = 2.732


def build(force=False):
    """Compile oracle/_build/libacfm_oracle.so with the committed Makefile."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("raster_oracle.c", "raster_oracle.inc", "Makefile"))
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < src_m:
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.acfm_oracle_max_threads.restype = ctypes.c_int
    return _lib


def set_variant(k_eps=1e-8, zmax_keps=False, tie_cuda=False):
    """Select a variant of the restated rasterizer (the SURVEY.md §9 items that could not be checked against the upstream
    source; see raster_oracle.inc).  Process-wide; tests that change it restore the default ()."""
    lib().acfm_oracle_set_variant(ctypes.c_double(k_eps), ctypes.c_int(int(zmax_keps)), ctypes.c_int(int(tie_cuda)))


def max_threads():
    return int(lib().acfm_oracle_max_threads())


def _sfx(dtype):
    return "_f32" if np.dtype(dtype) == np.float32 else "_f64"


def _ct(dtype):
    return ctypes.c_float if np.dtype(dtype) == np.float32 else ctypes.c_double


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def project(X, cam, offset_z=0.0, dtype=np.float32):
    """geom_utils.orthographic_proj_withz (geom_utils.py:62-79)."""
    X = np.ascontiguousarray(X, dtype=dtype)
    cam = np.ascontiguousarray(cam, dtype=dtype)
    N, V, _ = X.shape
    out = np.empty_like(X)
    fn = getattr(lib(), "acfm_oracle_project" + _sfx(dtype))
    fn(_p(X), _p(cam), _ct(dtype)(offset_z), ctypes.c_int(N), ctypes.c_int(V), _p(out))
    return out


def view(proj, yflip=True):
    """y-flip + R=diag(-1,1,1), T=(0,0,2.732): SURVEY.md §9.1."""
    proj = np.ascontiguousarray(proj)
    out = np.empty_like(proj)
    fn = getattr(lib(), "acfm_oracle_view" + _sfx(proj.dtype))
    fn(_p(proj), ctypes.c_int64(proj.size // 3), ctypes.c_int(1 if yflip else 0), _p(out))
    return out


def gather_faces(ndc, faces):
    ndc = np.ascontiguousarray(ndc)
    faces = np.ascontiguousarray(faces, dtype=np.int64)
    N, V, _ = ndc.shape
    F = faces.shape[1]
    fv = np.empty((N * F, 3, 3), dtype=ndc.dtype)
    fn = getattr(lib(), "acfm_oracle_gather_faces" + _sfx(ndc.dtype))
    fn(_p(ndc), _p(faces), ctypes.c_int(N), ctypes.c_int(V), ctypes.c_int(F), _p(fv))
    return fv


def rasterize(ndc, faces, image_size, blur_radius, K, clip_bary=False, cull_backfaces=False, threads=0,
              want_bary=True):
    """rasterize_meshes (naive CPU) on NDC verts (N,V,3) and faces (N,F,3).
    Returns pix_to_face (N,H,W,K) int64 packed ids n*F+f, zbuf, bary, dists."""
    ndc = np.ascontiguousarray(ndc)
    dtype = ndc.dtype
    N = ndc.shape[0]
    F = faces.shape[1]
    H = W = int(image_size)
    fv = gather_faces(ndc, faces)
    first = (np.arange(N, dtype=np.int64) * F)
    count = np.full((N,), F, dtype=np.int64)
    p2f = np.empty((N, H, W, K), dtype=np.int64)
    zbuf = np.empty((N, H, W, K), dtype=dtype)
    dists = np.empty((N, H, W, K), dtype=dtype)
    bary = np.empty((N, H, W, K, 3), dtype=dtype) if want_bary else None
    fn = getattr(lib(), "acfm_oracle_rasterize" + _sfx(dtype))
    fn(_p(fv), _p(first), _p(count), ctypes.c_int(N), ctypes.c_int(H), ctypes.c_int(W),
       _ct(dtype)(blur_radius), ctypes.c_int(K), ctypes.c_int(int(clip_bary)), ctypes.c_int(int(cull_backfaces)),
       _p(p2f), _p(zbuf), _p(bary), _p(dists), ctypes.c_int(threads))
    return dict(pix_to_face=p2f, zbuf=zbuf, bary=bary, dists=dists, face_verts=fv)


def sigmoid_alpha_blend(dists, pix_to_face, sigma=SIGMA):
    dists = np.ascontiguousarray(dists)
    K = dists.shape[-1]
    mask = np.empty(dists.shape[:-1], dtype=dists.dtype)
    fn = getattr(lib(), "acfm_oracle_sigmoid_alpha_blend" + _sfx(dists.dtype))
    fn(_p(dists), _p(np.ascontiguousarray(pix_to_face)), _ct(dists.dtype)(sigma), ctypes.c_int64(mask.size),
       ctypes.c_int(K), _p(mask))
    return mask


def sigmoid_alpha_blend_backward(dists, pix_to_face, grad_mask, sigma=SIGMA):
    dists = np.ascontiguousarray(dists)
    K = dists.shape[-1]
    gd = np.empty_like(dists)
    fn = getattr(lib(), "acfm_oracle_sigmoid_alpha_blend_backward" + _sfx(dists.dtype))
    fn(_p(dists), _p(np.ascontiguousarray(pix_to_face)), _ct(dists.dtype)(sigma),
       _p(np.ascontiguousarray(grad_mask, dtype=dists.dtype)), ctypes.c_int64(grad_mask.size), ctypes.c_int(K), _p(gd))
    return gd


def rasterize_backward(face_verts, pix_to_face, grad_dists=None, grad_zbuf=None):
    """grad wrt packed face_verts (Ftot,3,3): SURVEY.md §9.6."""
    face_verts = np.ascontiguousarray(face_verts)
    dtype = face_verts.dtype
    N, H, W, K = pix_to_face.shape
    g = np.zeros_like(face_verts)
    gd = None if grad_dists is None else np.ascontiguousarray(grad_dists, dtype=dtype)
    gz = None if grad_zbuf is None else np.ascontiguousarray(grad_zbuf, dtype=dtype)
    fn = getattr(lib(), "acfm_oracle_rasterize_backward" + _sfx(dtype))
    fn(_p(face_verts), _p(np.ascontiguousarray(pix_to_face)), _p(gz), _p(gd), ctypes.c_int(N), ctypes.c_int(H),
       ctypes.c_int(W), ctypes.c_int(K), _p(g))
    return g


def scatter_face_grads(grad_face_verts, faces, V):
    """Backward of verts_packed[faces_packed]: (N*F,3,3) -> (N,V,3)."""
    N, F, _ = faces.shape
    g = np.zeros((N, V, 3), dtype=grad_face_verts.dtype)
    gf = grad_face_verts.reshape(N, F, 3, 3)
    for n in range(N):
        np.add.at(g[n], faces[n].reshape(-1), gf[n].reshape(-1, 3))
    return g


# ---------------------------------------------------------------------------------------------
# Whole-path restatements (what the reference's renderer objects return)
# ---------------------------------------------------------------------------------------------
def neural_renderer_mask(vertices, faces, cams, img_size=256, offset_z=0.0, K=K_SOFT, dtype=np.float32, threads=0):
    """NeuralRenderer.forward(vertices, faces, cams) mask branch -> dict(mask, pix_to_face, zbuf, dists, ndc)."""
    proj = project(vertices, cams, offset_z, dtype)
    ndc = view(proj, yflip=True)
    fr = rasterize(ndc, faces, img_size, BLUR_SOFT, K, clip_bary=False, threads=threads, want_bary=False)
    fr["mask"] = sigmoid_alpha_blend(fr["dists"], fr["pix_to_face"])
    fr["ndc"] = ndc
    fr["proj"] = proj
    return fr


def neural_renderer_mask_backward(fr, faces, grad_mask):
    """d loss / d ndc verts (N,V,3) for a mask-branch render `fr` (output of neural_renderer_mask)."""
    gd = sigmoid_alpha_blend_backward(fr["dists"], fr["pix_to_face"], grad_mask)
    gfv = rasterize_backward(fr["face_verts"], fr["pix_to_face"], grad_dists=gd)
    return scatter_face_grads(gfv, np.asarray(faces), fr["ndc"].shape[1])


def of_renderer(verts, faces, img_size=256, dtype=np.float32, threads=0):
    """OF_NeuralRenderer.forward(verts, faces) -> pix_to_face (N,H,W,1). No y-flip (nmr.py:224-238)."""
    ndc = view(np.ascontiguousarray(verts, dtype=dtype), yflip=False)
    return rasterize(ndc, faces, img_size, 0.0, 1, clip_bary=False, threads=threads, want_bary=False)


def hard_raster(vertices, faces, cams, img_size=256, offset_z=0.0, dtype=np.float32, threads=0):
    """Texture-branch rasterization: blur 0, K=1, clip_barycentric_coords=True (nmr.py:85-87,183-195)."""
    proj = project(vertices, cams, offset_z, dtype)
    ndc = view(proj, yflip=True)
    fr = rasterize(ndc, faces, img_size, 0.0, 1, clip_bary=True, threads=threads, want_bary=True)
    fr["ndc"] = ndc
    return fr


def vertex_colors_as_texels(fr, colors, faces):
    """Textures(verts_rgb) interpolation with the (clipped) barycentrics: texel (N,H,W,K,3)."""
    p2f, bary = fr["pix_to_face"], fr["bary"]
    N, F = faces.shape[0], faces.shape[1]
    m = p2f >= 0
    fi = np.where(m, p2f, 0)
    n_idx, f_idx = fi // F, fi % F
    cols = np.broadcast_to(colors, (N,) + colors.shape[-2:])
    vid = faces[n_idx, f_idx]                                      # (N,H,W,K,3)
    c = cols[n_idx[..., None], vid]                                # (N,H,W,K,3 verts,3 rgb)
    return (bary[..., None] * c).sum(-2) * m[..., None]


def atlas_shade(fr, atlas, sigma=SIGMA, gamma=1e-4, znear=1.0, zfar=100.0, eps=1e-10, bg_texel="zero"):
    """TexturesAtlas.sample_textures + ambient-only Phong + softmax_rgb_blend: SURVEY.md §9.7.
    atlas (N,F,T,T,3).  Returns imgs (N,3,H,W), sil (N,H,W).
    bg_texel: what a padding fragment (pix_to_face = -1) samples — "zero" (texels masked, SURVEY §9.7's reading) or "wrap"
    (python's atlas_packed[-1]: the last face's texel, unmasked; the other reading of the 0.3.0 source).  The blend gives
    such fragments weight 0 either way; tests/test_oracle_variants.py asserts the images are identical."""
    p2f, bary, dists, zbuf = fr["pix_to_face"], fr["bary"], fr["dists"], fr["zbuf"]
    N, H, W, K = p2f.shape
    dt = dists.dtype
    R = atlas.shape[2]
    ap = np.ascontiguousarray(atlas, dtype=dt).reshape(-1, R, R, 3)
    m = p2f >= 0
    b = bary[..., :2].astype(dt)
    wxy = np.floor(b * dt.type(R)).astype(np.int64)
    # TexturesAtlas.sample_textures: below_diag = (bary_w01.sum(-1) * R - w_xy.float().sum(-1)) <= 1.0
    below = (b.sum(-1) * dt.type(R) - wxy.astype(dt).sum(-1)) <= 1.0
    wx = np.where(below, wxy[..., 0], R - 1 - wxy[..., 0])
    wy = np.where(below, wxy[..., 1], R - 1 - wxy[..., 1])
    fi = np.where(m, p2f, 0 if bg_texel == "zero" else ap.shape[0] - 1)
    wx = np.clip(np.where(m, wx, 0), 0, R - 1)
    wy = np.clip(np.where(m, wy, 0), 0, R - 1)
    texel = ap[fi, wy, wx]                                                  # (N,H,W,K,3)
    if bg_texel == "zero":
        texel = texel * m[..., None]
    return blend_texels(fr, texel, sigma, gamma, znear, zfar, eps)


def blend_texels(fr, texel, sigma=SIGMA, gamma=1e-4, znear=1.0, zfar=100.0, eps=1e-10):
    """softmax_rgb_blend (background 0) of per-fragment colours texel (N,H,W,K,3): SURVEY.md §9.7."""
    p2f, dists, zbuf = fr["pix_to_face"], fr["dists"], fr["zbuf"]
    dt = dists.dtype
    m = p2f >= 0
    prob = (1.0 / (1.0 + np.exp(dists.astype(np.float64) / sigma))).astype(dt) * m
    alpha = np.prod(1.0 - prob, axis=-1)
    zinv = ((zfar - zbuf) / (zfar - znear)).astype(dt) * m
    zmax = np.maximum(zinv.max(-1, keepdims=True), eps).astype(dt)
    wnum = prob * np.exp(((zinv - zmax) / gamma).astype(np.float64)).astype(dt)
    delta = np.maximum(np.exp(((eps - zmax) / gamma).astype(np.float64)), eps).astype(dt)
    denom = wnum.sum(-1, keepdims=True) + delta
    rgb = (wnum[..., None] * texel).sum(-2) / denom                         # bg = 0
    return np.transpose(rgb, (0, 3, 1, 2)).astype(dt), (1.0 - alpha).astype(dt)
