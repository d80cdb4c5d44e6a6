"""oracle/targets_ref.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/scipy restatement of the reference's target-map preprocessing (/root/reference/multiframe/utils/image.py:94-146):
compute_dt, compute_dt_barrier, compute_boundaries.  The reference calls scipy.ndimage.distance_transform_edt and
skimage.segmentation.find_boundaries; scipy is present, skimage is not, so find_boundaries is restated from its
published implementation (mode 'thick': grey_dilation != grey_erosion over the connectivity-1 footprint).
Pinned by tests/golden/targets.npz, produced by the reference's own functions (tests/golden/make_golden.py targets).
"""
import numpy as np
import scipy.ndimage as ndi


def compute_dt(mask, norm=True):
    """utils/image.py:94-102"""
    dist = ndi.distance_transform_edt(1 - mask)
    return dist / max(mask.shape) if norm else dist


def compute_dt_barrier(mask, k=50):
    """utils/image.py:105-116"""
    diff = (ndi.distance_transform_edt(1 - mask) - ndi.distance_transform_edt(mask)) / max(mask.shape)
    return 1.0 / (1 + np.exp(k * -diff))


def find_boundaries(m):
    fp = ndi.generate_binary_structure(m.ndim, 1)
    return ndi.grey_dilation(m, footprint=fp) != ndi.grey_erosion(m, footprint=fp)


def compute_boundaries(masks):
    """utils/image.py:122-146: (NB, max_bd, 3) float32 [x, y, valid], padded with the normalised zero (-1, -1, 0)."""
    pts = [np.transpose(find_boundaries(m).nonzero()) for m in masks]
    n = max(p.shape[0] for p in pts)
    out = np.zeros((len(pts), n, 3))
    for i, p in enumerate(pts):
        out[i, :p.shape[0], :2] = p
        out[i, :p.shape[0], 2] = 1
    y = (out[..., 0] / masks.shape[1] - 0.5) * 2
    x = (out[..., 1] / masks.shape[2] - 0.5) * 2
    return np.stack([x, y, out[..., 2]], -1).astype(np.float32)


def edt_bruteforce(mask_is_feature):
    """Exact squared EDT by exhaustive search (small maps): independent check of scipy and of the CUDA kernel."""
    H, W = mask_is_feature.shape
    ys, xs = np.nonzero(mask_is_feature)
    yy, xx = np.mgrid[0:H, 0:W]
    d2 = (yy[..., None] - ys) ** 2 + (xx[..., None] - xs) ** 2
    return d2.min(-1)
