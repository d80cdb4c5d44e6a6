"""Scratch timing of the lean-mode render (forward op and backward) on the C2 workload.  usage: [LIB=...] time_lean.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acfm_video_3d_reconstruction_b200 import _lib, functional as F_, synthetic
if os.environ.get("LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["LIB"])
wl = synthetic.Workload("bird", 64, 8, 32, 256, seed=0)
X = wl.mean_v[None].repeat(64, 1, 1).cuda()
ndc = F_.project(X, wl.cams.cuda(), 5.0, -1.0, -1.0, F_.EYE_Z)
faces = wl.faces[None].cuda()
tgt = (torch.rand(64, 256, 256, device="cuda") > 0.7).float(); edt = torch.rand(64, 256, 256, device="cuda")
gs = torch.randn(512, 4, device="cuda") * 1e-3
for lean in (False, True):
    tf, tb = [], []
    for it in range(10):
        x = ndc.clone().requires_grad_(True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record()
        if lean: mask, sums = F_.soft_silhouette_lean(x, faces, 256, tgt, edt)
        else: mask, _, _, _, sums = F_.soft_silhouette_losses(x, faces, 256, tgt, edt)
        e[1].record()
        loss = (sums * gs).sum()
        e[2].record()
        g, = torch.autograd.grad(loss, x)
        e[3].record(); torch.cuda.synchronize()
        tf.append(e[0].elapsed_time(e[1])); tb.append(e[2].elapsed_time(e[3]))
    print(os.path.basename(_lib.LIB_PATH), "lean" if lean else "parity", "fwd %.3f ms  bwd(+loss bwd) %.3f ms" % (sorted(tf[3:])[3], sorted(tb[3:])[3]), flush=True)
