"""Drop-in for the target-map half of the reference's utils/image.py
(/root/reference/multiframe/utils/image.py:94-146; same file in monocular/): compute_dt, compute_dt_barrier,
compute_boundaries — on the GPU, batched, exact.  The reference runs them on the CPU for every mask of every step
(scipy EDT twice per mask + skimage find_boundaries, multiframe/main.py:364-377) and copies the results to the device.

Inputs are CUDA float tensors (H,W) or (NB,H,W) holding the ground-truth masks; outputs are float32 CUDA tensors with the
values the reference obtains after its `torch.tensor(...).float()` (the reference's float64 intermediates are computed in
fp64 here too).
"""
import torch

from . import _lib
from . import functional as F_


def _batched(mask):
    _lib.require_cuda(mask)
    m = mask.float() if mask.dtype != torch.float32 else mask
    if m.dim() == 2:
        return m[None].contiguous(), True
    if m.dim() != 3:
        raise ValueError(f"mask must be (H,W) or (NB,H,W), got {tuple(mask.shape)}")
    return m.contiguous(), False


def _edt(mask, k, norm, want_edt, want_barrier):
    m, squeeze = _batched(mask)
    NB, H, W = m.shape
    dev = m.device
    edt = torch.empty_like(m) if want_edt else None
    bar = torch.empty_like(m) if want_barrier else None
    ws = torch.empty((2, NB, H, W), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().acfm_edt_fwd(_lib.ptr(m), NB, H, W, float(k), int(norm), _lib.ptr(edt), _lib.ptr(bar), _lib.ptr(ws),
                                     _lib.stream_of(m))
    _lib.check(st, "acfm_edt_fwd")
    _lib.count(2)
    outs = [t[0] if squeeze else t for t in (edt, bar) if t is not None]
    return outs[0] if len(outs) == 1 else outs


def compute_dt(mask, norm=True):
    """distance_transform_edt(1 - mask) [/ max(mask.shape)] (utils/image.py:94-102)."""
    return _edt(mask, 50.0, norm, True, False)


def compute_dt_barrier(mask, k=50):
    """1 / (1 + exp(-k (edt(1 - mask) - edt(mask)) / max(mask.shape))) (utils/image.py:105-116)."""
    return _edt(mask, k, False, False, True)


def compute_dt_both(mask, k=50):
    """(compute_dt(mask, norm=False), compute_dt_barrier(mask, k)) from one pass — what set_input needs
    (multiframe/main.py:368-369)."""
    return _edt(mask, k, False, True, True)


def compute_boundaries(masks, max_bd=None, return_counts=False):
    """(NB, max_bd, 3) float32 [x, y, valid] boundary points of each mask, padded to the longest list
    (utils/image.py:122-146).  max_bd=None: one host sync, the output length depends on the data (the reference's
    behaviour).  max_bd=int: the caller's capacity — no host sync, so the call can be captured in a CUDA graph with the
    rest of the step; lists longer than max_bd are truncated in raster-scan order (their tail is dropped) and shorter ones
    padded with (-1, -1, 0) as usual.  return_counts=True also returns the (NB,) int32 device tensor of true list lengths,
    which the caller can check against max_bd after the step."""
    m, _ = _batched(masks)
    NB, H, W = m.shape
    dev = m.device
    offs = torch.empty((NB, H), dtype=torch.int32, device=dev)
    tot = torch.empty((NB,), dtype=torch.int32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        st = L.acfm_boundaries_count(_lib.ptr(m), NB, H, W, _lib.ptr(offs), _lib.ptr(tot), _lib.stream_of(m))
        _lib.check(st, "acfm_boundaries_count")
        if max_bd is None:
            max_bd = int(tot.max()) if NB else 0
        max_bd = int(max_bd)
        out = torch.empty((NB, max_bd, 3), dtype=torch.float32, device=dev)
        st = L.acfm_boundaries_write(_lib.ptr(m), _lib.ptr(offs), _lib.ptr(tot), NB, H, W, max_bd, _lib.ptr(out), _lib.stream_of(m))
        _lib.check(st, "acfm_boundaries_write")
    _lib.count(3)
    return (out, tot) if return_counts else out
