"""Shared helpers for the tests: fixtures, synthetic inputs, comparisons."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def template(name="bird"):
    t = golden("templates.npz")
    return t[f"{name}_v"].copy(), t[f"{name}_f"].copy()


def synth_cams(n, seed=0, G=1):
    """s~U(.55,.85), t~U(-.1,.1)^2, q=normalize(N(0,I))."""
    gen = torch.Generator().manual_seed(seed)
    q = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1)
    s = torch.rand(n, 1, generator=gen) * 0.3 + 0.55
    t = torch.rand(n, 2, generator=gen) * 0.2 - 0.1
    return torch.cat([s, t, q], 1).numpy().astype(np.float32)


def synth_verts(v, n, seed=0, noise=0.02):
    gen = torch.Generator().manual_seed(seed + 1000)
    return (torch.from_numpy(v)[None].repeat(n, 1, 1) + noise * torch.randn(n, v.shape[0], 3, generator=gen)).numpy()


def icosphere(subdiv=3):
    """642 v / 1280 f at 3 subdivisions, 2562 / 5120 at 4 (the reference's template sizes)."""
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


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
