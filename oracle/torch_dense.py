"""oracle/torch_dense.py — TEST INFRASTRUCTURE / BASELINE, NOT PRODUCT CODE.

Dense pure-PyTorch soft-silhouette renderer (every pixel x every face, top-K by depth, sigmoid blend) with autograd for
the backward: the "B-GPU-torch" stand-in of BASELINE.md §2 for "PyTorch3D's CUDA path", which cannot be installed here.
It runs on whatever device its inputs are on; bench.py times it on the same B200 on a bounded sample (`gpu_standin`),
tests/test_raster_gpu.py checks it against the C oracle.  It follows SURVEY.md §9.4-9.5 in fp32 with torch's own operator
fusion / FMA behaviour, so it is NOT bit-exact (fragments can differ at ties); it is a speed baseline with checked values.
Not a faithful model of PyTorch3D's speed either: PyTorch3D bins faces coarse-to-fine, this does O(pixels x faces) work
with ~60 elementwise kernels — quoted ratios against it are labelled "stand-in" everywhere.
"""
import torch

K_EPS = 1e-8


def _edge(px, py, ax, ay, bx, by):
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax)


def _seg_dist(px, py, ax, ay, bx, by):
    bax, bay = bx - ax, by - ay
    l2 = bax * bax + bay * bay
    t = ((px - ax) * bax + (py - ay) * bay) / l2.clamp(min=K_EPS)
    t = t.clamp(0, 1)
    dx, dy = px - (ax + t * bax), py - (ay + t * bay)
    d = dx * dx + dy * dy
    return torch.where(l2 <= K_EPS, (px - bx) ** 2 + (py - by) ** 2, d)


def soft_silhouette(ndc, faces, image_size, blur_radius, K, sigma):
    """ndc (N,V,3) screen-space vertices (as acfm_project_fwd emits them), faces (F,3) shared topology.
    Returns mask (N,S,S), pix_to_face (N,S,S,K) packed ids (-1 pad).  Differentiable w.r.t. ndc through the distances."""
    N, V, _ = ndc.shape
    F = faces.shape[0]
    S = image_size
    dev = ndc.device
    i = torch.arange(S, device=dev, dtype=torch.float32)
    c = -1 + (2 * (S - 1 - i) + 1) / S                     # PixToNdc of the flipped index
    py, px = torch.meshgrid(c, c, indexing="ij")
    px, py = px.reshape(-1, 1), py.reshape(-1, 1)          # (P,1)
    masks, p2fs = [], []
    sq = blur_radius ** 0.5
    for n in range(N):
        fv = ndc[n][faces]                                  # (F,3,3)
        x0, y0, z0 = fv[:, 0, 0][None], fv[:, 0, 1][None], fv[:, 0, 2][None]
        x1, y1, z1 = fv[:, 1, 0][None], fv[:, 1, 1][None], fv[:, 1, 2][None]
        x2, y2, z2 = fv[:, 2, 0][None], fv[:, 2, 1][None], fv[:, 2, 2][None]
        area = _edge(x0, y0, x1, y1, x2, y2)
        den = _edge(x2, y2, x0, y0, x1, y1) + K_EPS
        w0 = _edge(px, py, x1, y1, x2, y2) / den
        w1 = _edge(px, py, x2, y2, x0, y0) / den
        w2 = _edge(px, py, x0, y0, x1, y1) / den
        pz = w0 * z0 + w1 * z1 + w2 * z2
        dist = torch.minimum(torch.minimum(_seg_dist(px, py, x0, y0, x1, y1), _seg_dist(px, py, x0, y0, x2, y2)),
                             _seg_dist(px, py, x1, y1, x2, y2))
        inside = (w0 > 0) & (w1 > 0) & (w2 > 0)
        xmin, xmax = torch.minimum(torch.minimum(x0, x1), x2) - sq, torch.maximum(torch.maximum(x0, x1), x2) + sq
        ymin, ymax = torch.minimum(torch.minimum(y0, y1), y2) - sq, torch.maximum(torch.maximum(y0, y1), y2) + sq
        ok = (torch.maximum(torch.maximum(z0, z1), z2) >= 0) & (area.abs() > K_EPS)
        cand = ok & ~((px > xmax) | (px < xmin) | (py > ymax) | (py < ymin)) & (pz >= 0) & (inside | (dist < blur_radius))
        zsel = torch.where(cand, pz.detach(), torch.full_like(pz, float("inf")))
        zk, idx = torch.topk(zsel, min(K, F), dim=1, largest=False, sorted=True)
        valid = torch.isfinite(zk)
        sd = torch.where(inside, -dist, dist).gather(1, idx)
        prob = torch.sigmoid(-sd / sigma) * valid
        masks.append((1 - torch.prod(1 - prob, dim=1)).view(S, S))
        p2f = torch.where(valid, idx + n * F, torch.full_like(idx, -1))
        if p2f.shape[1] < K:
            p2f = torch.cat([p2f, torch.full((p2f.shape[0], K - p2f.shape[1]), -1, device=dev, dtype=p2f.dtype)], 1)
        p2fs.append(p2f.view(S, S, K))
    return torch.stack(masks), torch.stack(p2fs)
