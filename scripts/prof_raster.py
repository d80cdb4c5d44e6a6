"""Launches the raster fwd/bwd kernels a few times on the C2 shapes (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acfm_video_3d_reconstruction_b200 import NeuralRenderer, functional as F_, synthetic

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
tmpl = sys.argv[4] if len(sys.argv) > 4 else "bird"
v, f = synthetic.template(tmpl)
G = 8
X = torch.from_numpy(v)[None].repeat(N // G, 1, 1).cuda()
cam = synthetic.cameras(N // G, G, seed=0).cuda()
faces = torch.from_numpy(f)[None].cuda()
r = NeuralRenderer(S, offset_z=5.0)
ndc = F_.project(X, cam, 5.0, -1.0, -1.0, F_.EYE_Z).requires_grad_(True)
gm = torch.randn(N, S, S, device="cuda")
for it in range(2):
    mask, p2f, zb, d = F_.soft_silhouette(ndc, faces, S, r.blur_radius, K, r.sigma)
    g, = torch.autograd.grad((mask * gm).sum(), ndc)
torch.cuda.synchronize()
print("ok", float(mask.mean()), float(g.abs().sum()))
