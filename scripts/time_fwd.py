"""Scratch timing of the forward rasterizer alone on the bench workloads (CUDA events on the launching stream).
usage: [LIB=path/to/variant.so] time_fwd.py [template:frames:img:K ...]     (default bird:64:256:20; 8 hypotheses per frame)
LIB loads another build of the library (A/B runs of kernel variants in one gpurun call); bench.py is the judged harness."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acfm_video_3d_reconstruction_b200 import _lib, functional as F_, synthetic

if os.environ.get("LIB"):
    import ctypes
    _lib.LIB_PATH = os.path.abspath(os.environ["LIB"])
    probe = ctypes.CDLL(_lib.LIB_PATH)
    for name in list(_lib.SIGNATURES):  # older builds lack the newer entry points
        if not hasattr(probe, name):
            del _lib.SIGNATURES[name]
tag = os.path.basename(_lib.LIB_PATH)
for spec in (sys.argv[1:] or ["bird:64:256:20"]):
    name, frames, S, K = spec.split(":")
    frames, S, K = int(frames), int(S), int(K)
    wl = synthetic.Workload(name, frames, 8, 32, S, seed=0)
    X = wl.mean_v[None].repeat(frames, 1, 1).cuda()
    ndc = F_.project(X, wl.cams.cuda(), 5.0, -1.0, -1.0, F_.EYE_Z)
    faces = wl.faces[None].cuda()
    N = ndc.shape[0]
    ts = []
    for it in range(10):
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
    t = sorted(ts[3:])[len(ts[3:]) // 2]
    fb = S * S * (16 * K + 4) * N
    print(f"[{tag}] {name} N={N} {S}^2 K={K} split={F_.SPLIT_FILL}: fwd {t:.3f} ms (min {min(ts):.3f})  {fb / t / 1e6:.0f} GB/s alg"
          + ("  all: " + " ".join(f"{x:.2f}" for x in ts) if os.environ.get("ALL") else ""), flush=True)
    del ndc, X
    torch.cuda.empty_cache()
