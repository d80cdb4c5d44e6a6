"""Launches the K=1 hard rasterizer (OF_NeuralRenderer path) a few times on C3-like shapes (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acfm_video_3d_reconstruction_b200 import functional as F_, synthetic
v, f = synthetic.template("horse")
G, NB = 8, 32
X = torch.from_numpy(v)[None].repeat(NB, 1, 1).cuda()
cam = synthetic.cameras(NB, G, seed=0).cuda()
faces = torch.from_numpy(f)[None].cuda()
ndc = F_.project(X, cam, 0.0, -1.0, -1.0, F_.EYE_Z)
for it in range(3):
    fr = F_.rasterize(ndc, faces, 256, 0.0, 1)
torch.cuda.synchronize()
print("ok", float((fr["pix_to_face"] >= 0).float().mean()))
