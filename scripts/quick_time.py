"""Scratch timing of the raster kernels (CUDA events); bench.py is the judged harness."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acfm_video_3d_reconstruction_b200 import NeuralRenderer, functional as F_
from tests import util

N = int(sys.argv[1]) if len(sys.argv) > 1 else 512
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
v, f = util.template("bird")
X = torch.from_numpy(util.synth_verts(v, N, 1)).cuda().requires_grad_(True)
cam = torch.from_numpy(util.synth_cams(N, 2)).cuda().requires_grad_(True)
faces = torch.from_numpy(f)[None].cuda().expand(N, -1, -1)
r = NeuralRenderer(S, offset_z=5.0); r.faces_per_pixel = K
gm = torch.randn(N, S, S, device="cuda")
def ev(): return torch.cuda.Event(enable_timing=True)
for it in range(3):
    ndc = r.to_ndc(X, cam)
    e = [ev() for _ in range(4)]
    e[0].record()
    mask, p2f, zb, d = F_.soft_silhouette(ndc, faces, S, r.blur_radius, K, r.sigma)
    e[1].record()
    loss = (mask * gm).sum()
    e[2].record()
    g, = torch.autograd.grad(loss, ndc)
    e[3].record()
    torch.cuda.synchronize()
    fwd, bwd = e[0].elapsed_time(e[1]), e[2].elapsed_time(e[3])
    fb = S * S * (16 * K + 4) * N; bb = S * S * (12 * K + 8) * N
    print(f"iter {it}: fwd {fwd:.3f} ms ({fb / fwd / 1e6:.0f} GB/s alg)  bwd(+loss bwd) {bwd:.3f} ms ({bb / bwd / 1e6:.0f} GB/s alg)  "
          f"{N / (fwd + bwd) * 1e3:.0f} renders/s; cov {(p2f[..., 0] >= 0).float().mean().item():.3f}")
