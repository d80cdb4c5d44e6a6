"""Batch / resolution sweep of the soft-silhouette render fwd+bwd (BASELINE.json configs[4]) on 1 / 2 / 4 / 8 GPUs, with the CPU
restatement timed beside it.
  python scripts/sweep.py [--out profiles/sweep_r05.md] [--no-cpu]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep.py [--out ...]
Under torchrun the N renders of a row are split evenly over the ranks (no collective on the data path: every render is
independent), every rank times its share, the row reports all renders / the slowest rank's time.
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


def cpu_baseline(S, K, v, f, renders=8):
    """renders/s of the CPU restatement (oracle/: PyTorch3D 0.3.0's naive CPU rasterizer + blend + backward, all host threads)
    at this resolution, on a bounded sample — the "vs CPU baseline" column of BASELINE config 5."""
    import time
    import numpy as np
    from oracle import pt3d_oracle as orc
    X = np.repeat(v[None], renders, 0)
    cam = synthetic.cameras(renders // 8 or 1, 8, seed=0).numpy()[:renders]
    faces = np.repeat(f[None], renders, 0)
    gm = np.random.default_rng(0).standard_normal((renders, S, S)).astype(np.float32)
    t0 = time.time()
    fr = orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=5.0, K=K)
    orc.neural_renderer_mask_backward(fr, faces, gm)
    return renders / (time.time() - t0), orc.max_threads()


def run(N, S, K, v, f, peak, world=1, budget_gb=40.0):
    G = 8
    N_total, N = N, max(G, N // world)
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
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t[0])
    done = reps * chunk * world
    alg = S * S * (16 * K + 4) + S * S * (12 * K + 8) + 36 * v.shape[0] + 48 * f.shape[0]
    rps = done / (ms * 1e-3)
    return dict(N=N_total, S=S, chunk=chunk, gpus=world, renders_per_s=rps, ms_per_render=ms / done,
                roofline_frac=rps * alg / 1e9 / (peak * world))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--K", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pk = os.path.join(root, "MEASURED_PEAKS.json")
    peak = float(json.load(open(pk))["hbm_gbs"]) if os.path.exists(pk) else 6650.0
    v, f = synthetic.template("bird")
    rows = []
    for S in (128, 256, 512, 1024):
        cpu = (None, None)
        if rank == 0 and not args.no_cpu:
            cpu = cpu_baseline(S, args.K, v, f)
        for N in (64, 512, 4096):
            if S == 1024 and N == 4096:
                N = 1024      # 4096 x 1024^2 x K=20 fragments = 1.4 TB per pass; time 1024 renders and say so
            r = run(N, S, args.K, v, f, peak, world)
            r["cpu_renders_per_s"], r["cpu_threads"] = cpu
            rows.append(r)
            if rank == 0:
                print(json.dumps(r), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()
    if args.out and rank == 0:
        with open(args.out, "w") as fh:
            fh.write("# Batch / resolution sweep, soft-silhouette render fwd+bwd, %d x B200 (bird 642v/1280f, K=%d)\n\n" % (world, args.K))
            fh.write("`python scripts/sweep.py` — project + raster fwd + blend, mask-gradient -> raster bwd -> projection bwd; CUDA events;\n"
                     "renders processed in chunks of `chunk`; roofline = API-parity algorithmic bytes / measured HBM peak (%.1f GB/s).\n\n" % peak)
            fh.write("CPU column: oracle/ (restated PyTorch3D 0.3.0 naive CPU rasterizer + blend + backward), all host threads, 8 renders per resolution.\n\n")
            fh.write("| renders (all GPUs) | image | GPUs | chunk / GPU | renders/s | us/render | fraction of HBM roofline (per GPU) | CPU renders/s (threads) | GPU / CPU |\n|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                c = r["cpu_renders_per_s"]
                fh.write(f"| {r['N']} | {r['S']}^2 | {r['gpus']} | {r['chunk']} | {r['renders_per_s']:.0f} | {r['ms_per_render'] * 1e3:.2f} | {r['roofline_frac']:.3f} | "
                         + (f"{c:.2f} ({r['cpu_threads']}) | {r['renders_per_s'] / c:.0f} |\n" if c else "- | - |\n"))


if __name__ == "__main__":
    main()
