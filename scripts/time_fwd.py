"""Scratch timing of the forward rasterizer alone on the bench workloads (CUDA events on the launching stream).
usage: time_fwd.py [bird|horse] [frames] [img] [K]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acfm_video_3d_reconstruction_b200 import functional as F_, synthetic

name = sys.argv[1] if len(sys.argv) > 1 else "bird"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
K = int(sys.argv[4]) if len(sys.argv) > 4 else 20
wl = synthetic.Workload(name, frames, 8, 32, S, seed=0)
X = wl.mean_v[None].repeat(frames, 1, 1).cuda()
ndc = F_.project(X, wl.cams.cuda(), 5.0, -1.0, -1.0, F_.EYE_Z)
faces = wl.faces[None].cuda()
N = ndc.shape[0]
ts = []
for it in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if K == 1:  # the hard renders of the path (OF visibility / texture branch)
        fr = F_.rasterize(ndc, faces, S, 0.0, 1, clip_barycentric_coords=True, want_bary=bool(os.environ.get("BARY")))
    else:
        fr = F_.rasterize(ndc, faces, S, F_.BLUR_SOFT, K, sigma=F_.SIGMA, want_mask=True)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
    del fr
t = sorted(ts[2:])[len(ts[2:]) // 2]
fb = S * S * (16 * K + 4) * N
print(f"{name} N={N} {S}^2 K={K} split={F_.SPLIT_FILL} only={os.environ.get('ACFM_FWD_ONLY')} per_sm={os.environ.get('ACFM_FILL_PER_SM')}: "
      f"fwd {t:.3f} ms (min {min(ts):.3f})  {fb / t / 1e6:.0f} GB/s alg")
