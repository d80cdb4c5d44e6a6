"""Scratch timing of the backward rasterizer alone on the bench workloads (CUDA events on the launching stream): the forward
of the product build renders once, then the backward of the build under test (LIB) runs 10 times on those fragments, with the
upstream gradient formed from grad_sums (fused losses, as bench.py's step) or from an explicit grad_mask (MODE=mask).
usage: [LIB=path/to/variant.so] [MODE=sums|mask] time_bwd.py [template:frames:img:K ...]     (default bird:64:256:20)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acfm_video_3d_reconstruction_b200 import _lib, functional as F_, synthetic

if os.environ.get("LIB"):
    import ctypes
    _lib.LIB_PATH = os.path.abspath(os.environ["LIB"])
    probe = ctypes.CDLL(_lib.LIB_PATH)
    for name in list(_lib.SIGNATURES):
        if not hasattr(probe, name):
            del _lib.SIGNATURES[name]
tag = os.path.basename(_lib.LIB_PATH)
mode = os.environ.get("MODE", "sums")
for spec in (sys.argv[1:] or ["bird:64:256:20"]):
    name, frames, S, K = spec.split(":")
    frames, S, K = int(frames), int(S), int(K)
    wl = synthetic.Workload(name, frames, 8, 32, S, seed=0)
    X = wl.mean_v[None].repeat(frames, 1, 1).cuda()
    ndc = F_.project(X, wl.cams.cuda(), 5.0, -1.0, -1.0, F_.EYE_Z)
    faces = wl.faces[None].cuda()
    N = ndc.shape[0]
    gen = torch.Generator(device="cuda").manual_seed(1)
    with torch.no_grad():
        tgt = F_.rasterize(ndc[:frames], faces, S, F_.BLUR_SOFT, K, sigma=F_.SIGMA, want_mask=True)["mask"]
        target = (tgt.roll(3, 2) > 0.5).float()
    edt = torch.rand(frames, S, S, device="cuda", generator=gen)
    fr = F_._train_render(ndc, faces, S, F_.BLUR_SOFT, K, F_.SIGMA, False, target, edt)
    gs = torch.randn(N, 4, device="cuda", generator=gen) * 1e-3
    gm = torch.randn(N, S, S, device="cuda", generator=gen) if mode == "mask" else None
    saved = (fr["ndc"], faces, fr["pix_to_face"], fr["dists"], fr["mask"], target if mode != "mask" else None, edt if mode != "mask" else None)
    ts = []
    for it in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g = F_._train_render_bwd(saved, (S, K, F_.SIGMA), fr["work"], gm, gs if mode != "mask" else None)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts[3:])[len(ts[3:]) // 2]
    print(f"[{tag}] {name} N={N} {S}^2 K={K} {mode}: bwd {t:.3f} ms (min {min(ts):.3f})  |g| {float(g.abs().sum()):.6e}", flush=True)
    del fr, saved, g
    torch.cuda.empty_cache()
