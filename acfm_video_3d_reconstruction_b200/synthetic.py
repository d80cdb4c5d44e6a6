"""Synthetic inputs of the reference's shapes for parity tests and bench.py: there is no
network for datasets or checkpoints, so templates come from the committed fixture
(tests/golden/templates.npz = the reference's OBJ templates) or an icosphere, and handle weights,
cameras and targets are drawn from seeded generators that mirror the reference's initialisers.
Init-time host code (numpy / torch CPU), not on the hot path.
"""
import os

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def icosphere(subdiv=3):
    """642 v / 1280 f at 3 subdivisions, 2562 / 5120 at 4 (utils/mesh.py:13-17 of the reference)."""
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
         (8, 6, 7), (9, 8, 1)]
    v = [np.asarray(x, np.float64) / np.linalg.norm(x) for x in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.asarray(v, np.float32), np.asarray(f, np.int64)


def template(name="bird"):
    """'bird' / 'horse': the reference's 642 v / 1280 f templates (fixture); 'ico3' / 'ico4': ellipsoids.
    Normalised to max |coordinate| = 1 except the icospheres (radius 0.8, scaled (1, .6, .5))."""
    if name in ("bird", "horse"):
        t = np.load(os.path.join(_ROOT, "tests", "golden", "templates.npz"))
        v, f = t[f"{name}_v"].astype(np.float32), t[f"{name}_f"].astype(np.int64)
        return (v / np.abs(v).max()).astype(np.float32), f
    if name in ("ico3", "ico4"):
        v, f = icosphere(int(name[-1]))
        return (v * 0.8 * np.array([1.0, 0.6, 0.5], np.float32)).astype(np.float32), f
    raise ValueError(name)


def uniform_laplacian(V, faces):
    """Meshes.laplacian_packed() as used by mesh_laplacian(.., 'uniform') (geom_utils.py:247-248;
    SURVEY.md §9.8): L[i,j] = 1/deg(i) on edges, L[i,i] = -1."""
    A = np.zeros((V, V), np.float32)
    for a, b in ((0, 1), (1, 2), (2, 0)):
        A[faces[:, a], faces[:, b]] = 1
        A[faces[:, b], faces[:, a]] = 1
    deg = A.sum(1, keepdims=True)
    L = A / np.maximum(deg, 1)
    L[np.arange(V), np.arange(V)] = -1
    return L.astype(np.float32)


def handle_weights(verts, num_handles, pp=16):
    """Farthest-point handles + lbs = ln(clamp(1/d^pp)) as MeshNet.__init__ (mesh_net.py:540-560) with
    Euclidean instead of geodesic distance (gdist is not installed).  Returns the raw (V,Kh) parameter;
    the model applies softmax over VERTICES (get_lbs, mesh_net.py:597-599)."""
    V = verts.shape[0]
    idx = [int(np.argmin(verts[:, 1]))]
    d = np.linalg.norm(verts - verts[idx[0]], axis=1)
    for _ in range(num_handles - 1):
        idx.append(int(np.argmax(d)))
        d = np.minimum(d, np.linalg.norm(verts - verts[idx[-1]], axis=1))
    idx = np.sort(np.asarray(idx))
    dist = np.linalg.norm(verts[:, None] - verts[None, idx], axis=-1)
    with np.errstate(divide="ignore"):
        lbs = 1.0 / dist ** pp
    lbs[np.isinf(lbs)] = 0
    mx = lbs.max(0)
    lbs[idx, np.arange(num_handles)] = mx
    return np.log(np.clip(lbs, 1e-10, None)).astype(np.float32), idx


def cameras(num_frames, G, seed=0):
    """(G*num_frames, 7) hypothesis-major [s,tx,ty,q]: s~U(.55,.85), t~U(-.1,.1)^2, q=normalize(N(0,I)),
    hypothesis g adds a yaw of 2 pi g / G (mirrors the embedding init, mesh_net.py:424-442)."""
    gen = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(num_frames, 4, generator=gen), dim=-1)
    s = torch.rand(num_frames, 1, generator=gen) * 0.3 + 0.55
    t = torch.rand(num_frames, 2, generator=gen) * 0.2 - 0.1
    out = []
    for g in range(G):
        a = 2 * np.pi * g / G
        qy = torch.tensor([np.cos(a / 2), 0.0, np.sin(a / 2), 0.0], dtype=torch.float32)
        # q_g = q_yaw (x) q  (Hamilton product)
        w0, x0, y0, z0 = qy
        w1, x1, y1, z1 = q.unbind(-1)
        qq = torch.stack([w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1, w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1,
                          w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1, w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1], -1)
        out.append(torch.cat([s, t, qq], 1))
    return torch.cat(out, 0).float()


def uv_sampler(verts, faces, tex_size=6):
    """compute_uvsampler + get_spherical_coords (utils/mesh.py:197-238 of the reference): per-face T x T
    barycentric sample points pushed to spherical UV in [-1,1].  Returns (F, T, T, 2) float32."""
    a = np.arange(tex_size, dtype=np.float64) / (tex_size - 1)
    coords = np.stack([(x, y) for x in a for y in a])                       # itertools.product(alpha, beta)
    vs = verts[faces].astype(np.float64)
    v2, v0v2, v1v2 = vs[:, 2], vs[:, 0] - vs[:, 2], vs[:, 1] - vs[:, 2]
    samples = np.dstack([v0v2, v1v2]).dot(coords.T) + v2.reshape(-1, 3, 1)  # F x 3 x T*T
    X = np.transpose(samples, (0, 2, 1)).reshape(-1, 3)
    rad = np.linalg.norm(X, axis=1)
    theta = np.arccos(np.clip(X[:, 2] / np.maximum(rad, 1e-12), -1, 1))
    phi = np.arctan2(X[:, 1], X[:, 0])
    uv = np.stack([((phi + np.pi) / (2 * np.pi)) * 2 - 1, (theta / np.pi) * 2 - 1], 1)
    return uv.reshape(-1, tex_size, tex_size, 2).astype(np.float32)


class Workload:
    """One data-parallel shard of a training step's hot-path inputs (host tensors)."""

    def __init__(self, template_name="bird", frames=64, G=8, handles=32, img_size=256, seed=0, offset_z=5.0):
        v, f = template(template_name)
        self.V, self.F = v.shape[0], f.shape[0]
        self.frames, self.G, self.Kh, self.img_size, self.offset_z = frames, G, handles, img_size, offset_z
        self.mean_v = torch.from_numpy(v)
        self.faces = torch.from_numpy(f)
        lbs, self.handle_idx = handle_weights(v, handles)
        self.lbs_param = torch.from_numpy(lbs)
        self.L = torch.from_numpy(uniform_laplacian(self.V, f))
        gen = torch.Generator().manual_seed(seed + 17)
        self.delta = 0.05 * torch.randn(frames, handles, 3, generator=gen)
        self.cams = cameras(frames, G, seed)
        self.seed = seed

    @property
    def renders(self):
        return self.frames * self.G
