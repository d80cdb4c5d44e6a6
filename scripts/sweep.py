"""Batch / resolution sweep of the soft-silhouette render fwd+bwd (BASELINE.json configs[4]) on one GPU.
  python scripts/sweep.py [--out profiles/sweep_r01.md]
For every (renders N, image size S) it times NeuralRenderer-level fwd (project + raster + blend) and bwd (mask-loss
gradient -> raster bwd -> projection bwd) with CUDA events, renders processed in chunks that fit in memory, and prints
renders/s and the fraction of the API-parity HBM roofline.  Template: bird 642 v / 1280 f, K = 20."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from acfm_video_3d_reconstruction_b200 import functional as F_  # noqa: E402
from acfm_video_3d_reconstruction_b200 import synthetic  # noqa: E402


def run(N, S, K, v, f, peak, budget_gb=40.0):
    G = 8
    chunk = max(G, min(N, int(budget_gb * 1e9 / (S * S * (16 * K + 8))) // G * G))
    X = torch.from_numpy(v)[None].repeat(chunk // G, 1, 1).cuda()
    cam = synthetic.cameras(chunk // G, G, seed=0).cuda().requires_grad_(True)
    faces = torch.from_numpy(f)[None].cuda()
    gm = torch.randn(chunk, S, S, device="cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]

    def one():
        ndc = F_.project(X, cam, 5.0, -1.0, -1.0, F_.EYE_Z)
        mask, _, _, _ = F_.soft_silhouette(ndc, faces, S, F_.BLUR_SOFT, K, F_.SIGMA)
        (mask * gm).sum().backward()

    for _ in range(2):
        one()
    reps = max(1, -(-N // chunk))
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        one()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    done = reps * chunk
    alg = S * S * (16 * K + 4) + S * S * (12 * K + 8) + 36 * v.shape[0] + 48 * f.shape[0]
    rps = done / (ms * 1e-3)
    return dict(N=N, S=S, chunk=chunk, renders_per_s=rps, ms_per_render=ms / done, roofline_frac=rps * alg / 1e9 / peak)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--K", type=int, default=20)
    args = ap.parse_args()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pk = os.path.join(root, "MEASURED_PEAKS.json")
    peak = float(json.load(open(pk))["hbm_gbs"]) if os.path.exists(pk) else 6650.0
    v, f = synthetic.template("bird")
    rows = []
    for S in (128, 256, 512, 1024):
        for N in (64, 512, 4096):
            if S == 1024 and N == 4096:
                N = 1024      # 4096 x 1024^2 x K=20 fragments = 1.4 TB per pass; time 1024 renders and say so
            r = run(N, S, args.K, v, f, peak)
            rows.append(r)
            print(json.dumps(r), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            fh.write("# Batch / resolution sweep, soft-silhouette render fwd+bwd, 1 x B200 (bird 642v/1280f, K=%d)\n\n" % args.K)
            fh.write("`python scripts/sweep.py` — project + raster fwd + blend, mask-gradient -> raster bwd -> projection bwd; CUDA events;\n"
                     "renders processed in chunks of `chunk`; roofline = API-parity algorithmic bytes / measured HBM peak (%.1f GB/s).\n\n" % peak)
            fh.write("| renders | image | chunk | renders/s | us/render | fraction of HBM roofline |\n|---|---|---|---|---|---|\n")
            for r in rows:
                fh.write(f"| {r['N']} | {r['S']}^2 | {r['chunk']} | {r['renders_per_s']:.0f} | {r['ms_per_render'] * 1e3:.2f} | {r['roofline_frac']:.3f} |\n")


if __name__ == "__main__":
    main()
