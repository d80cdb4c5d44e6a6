#!/usr/bin/env python
"""bench.py — render fwd+bwd frames/s of the ACFM hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload C2|C4]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic input on every rank:
  skinning matrix (Woodbury update of a once-inverted Laplacian term, fp64) -> fused handle deformation + G-hypothesis projection -> tile-binned soft
  rasterizer (fragments pix_to_face/zbuf/dists materialised, fused silhouette blend) -> fused mask losses
  (l1 + edt) -> hypothesis softmax weighting -> backward to handle offsets, cameras and handle weights ->
  (N > 1) NCCL all-reduce of the shared-parameter gradients.
Workload at N=1 is BASELINE.json configs[1] (C2): bird template 642 v / 1280 f, batch 64 frames x 8 camera
hypotheses = 512 renders of 256 x 256, K = 20, 32 handles.  Weak scaling: every rank owns its own 64 frames.
`value` = renders of all ranks / max-over-ranks device time, inputs resident in HBM.  `e2e` = the same through the
public API with host (pinned) inputs copied in and the loss + gradients copied out inside the timed region (step replayed
from a CUDA graph, the next step's inputs prefetched on a copy stream; `e2e.serial_value` = without the prefetch).
Side measurements in the same line (N = 1): `target_maps`, `post_optimize`, `correlation` (vs the reference's own extension
compiled for sm_100a), `gpu_standin`.
`--impl reference` times the CPU oracle (restated PyTorch3D 0.3.0 CPU algorithm; the real wheel is not installable)
on the host cores for the same metric.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (template, frames/rank, G, handles, img, K, offset_z)
    "C2": dict(template="bird", frames=64, G=8, handles=32, img=256, K=20, offset_z=5.0,
               desc="C2 monocular bird step: 642v/1280f, batch 64 x 8 camera hypotheses = 512 renders/GPU, 256x256, K=20, 32 handles"),
    "C3": dict(template="horse", frames=64, G=8, handles=16, img=256, K=20, offset_z=0.0, full=True, clip_frames=16,
               desc="C3 multiframe quadruped step, one GPU's share: 4 clips x 16 frames x 8 cameras = 512 renders/GPU, 256x256, K=20, "
                    "16 handles; full loss set (camera assembly, mask l1 + edt, boundary, optical-flow, keypoint, hypothesis weighting)"),
    "C4": dict(template="ico4", frames=8, G=8, handles=32, img=512, K=50, offset_z=0.0,
               desc="C4 high-res stress: 2562v/5120f subdivided template, 8 frames x 8 hypotheses = 64 renders/GPU, 512x512, K=50"),
}
W_EDT = 0.1  # edt_reg_wt-like weight on the edt term (any fixed weight exercises the same kernels)


def alg_bytes(img, K, V, F):
    """Algorithmic HBM bytes per render of the rasterizer, BASELINE.md §3 (API-parity mode)."""
    fwd = img * img * (16 * K + 4) + 12 * V + 24 * F
    bwd = img * img * (12 * K + 8) + 24 * V + 24 * F
    return fwd, bwd


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler
# ---------------------------------------------------------------------------------------------------------------
class ClocksNvml:
    """SM clock and clock-event reasons sampled through NVML every ~4 ms from a thread (nvidia-smi's fastest loop, 50 ms, puts
    one or two samples into a 65 ms timed region); same result keys as Clocks below, which is the fallback."""

    def __init__(self, index):
        import pynvml
        self.nv = pynvml
        pynvml.nvmlInit()
        uuid = None
        try:   # honour CUDA_VISIBLE_DEVICES: NVML enumerates every GPU of the box
            uuid = torch.cuda.get_device_properties(index).uuid
            self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
        except Exception:
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
        self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)   # fail here, not in the thread
        self.rows, self.stop_flag = [], False
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.time(), float(sm), int(r)))
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self, t0, t1):
        nv = self.nv
        self.stop_flag = True
        self.th.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1]
        window = "timed"
        if len(rows) < 3:
            rows, window = self.rows, "warmup+timed"
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        reasons = sorted({n for _, _, r in rows for n, bit in names if r & bit})
        sm = [r[1] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "reasons": reasons, "samples": len(sm),
                "window": window, "source": "nvml, 4 ms"}


def make_clocks(index):
    try:
        return ClocksNvml(index)
    except Exception:
        return Clocks(index)


class Clocks:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for r in self.rows if t0 <= r[0] <= t1 + 0.1]
        window = "timed"
        if len(rows) < 3:
            rows, window = self.rows, "warmup+timed"
        sm, mx, reasons = [], [], set()
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------------------------------
# the step
# ---------------------------------------------------------------------------------------------------------------
class HotPath:
    """Device-resident constants of one rank's shard + the step itself (public API of the package only)."""

    def __init__(self, cfg, rank, device):
        from acfm_video_3d_reconstruction_b200 import synthetic
        self.cfg, self.device = cfg, device
        wl = synthetic.Workload(cfg["template"], cfg["frames"], cfg["G"], cfg["handles"], cfg["img"], seed=rank,
                                offset_z=cfg["offset_z"])
        self.wl = wl
        self.mean_v = wl.mean_v.to(device)
        self.lbs_param = wl.lbs_param.to(device).requires_grad_(True)
        self.L = wl.L.to(device)
        from acfm_video_3d_reconstruction_b200 import deform
        self.solver = deform.HandleSolver(self.L)                  # the Laplacian is fixed at init (monocular/main.py:124)
        self.faces = wl.faces.to(device)[None]                     # shared topology, (1,F,3) int64
        # host-side (pinned) per-step inputs
        self.h_delta = wl.delta.pin_memory()
        self.h_cams = wl.cams.pin_memory()
        tgt, edt = self._targets(rank)
        self.h_target, self.h_edt = tgt.pin_memory(), edt.pin_memory()
        self.h_loss = torch.empty((), dtype=torch.float32).pin_memory()
        self.h_gdelta = torch.empty_like(wl.delta).pin_memory()
        self.h_gcams = torch.empty_like(wl.cams).pin_memory()
        if cfg.get("full"):
            self._full_inputs(rank)
        self.bucket, self.side = None, None

    def make_bucket(self, mbytes):
        """A synthetic flat bucket standing for the gradients of the parameters shared by all frames (encoder, heads, texture
        net: 12-15 M fp32 = 50-60 MB) — the model itself is out of scope, its all-reduce is not: the
        bucket is summed over the ranks every step on a side stream, started when the backward starts."""
        self.bucket = torch.zeros(int(mbytes * (1 << 20)) // 4, dtype=torch.float32, device=self.device) if mbytes > 0 else None
        self.side = torch.cuda.Stream(device=self.device) if mbytes > 0 else None

    def kick_bucket(self):
        if self.bucket is None:
            return
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            torch.distributed.all_reduce(self.bucket)

    def join_bucket(self):
        if self.bucket is not None:
            torch.cuda.current_stream().wait_stream(self.side)

    def _full_inputs(self, rank):
        """Device-resident synthetic targets of the full multiframe loss set: boundary points of the target
        masks (our GPU compute_boundaries), noisy projected keypoints, masked random flow, mirror / affine augmentation flags."""
        from acfm_video_3d_reconstruction_b200 import image_utils
        cfg, d = self.cfg, self.device
        gen = torch.Generator().manual_seed(100 + rank)
        NB, T = cfg["frames"], cfg["clip_frames"]
        tgt = self.h_target.to(d)
        self.boundaries = image_utils.compute_boundaries(tgt)
        self.bds_sel = torch.randperm(self.boundaries.shape[1], generator=gen)[:1000].to(d)
        Kp = 16
        self.vert2kp = torch.randn(Kp, self.wl.V, generator=gen).mul(4.0).to(d).requires_grad_(True)
        kp_xy = torch.rand(NB, Kp, 2, generator=gen) * 1.2 - 0.6
        self.kps = torch.cat([kp_xy, (torch.rand(NB, Kp, 1, generator=gen) > 0.2).float()], -1).to(d)
        flows = 2.0 * torch.randn(NB // T, T, cfg["img"], cfg["img"], 2, generator=gen)
        self.flows = flows.to(d) * tgt.reshape(NB // T, T, cfg["img"], cfg["img"], 1)
        self.mirror = (torch.rand(NB, generator=gen) > 0.5).float().to(d)
        tr = torch.cat([torch.rand(NB, 1, generator=gen) * 0.2 + 0.9, torch.rand(NB, 2, generator=gen) * 0.1 - 0.05,
                        (torch.rand(NB, 1, generator=gen) > 0.5).float()], 1)
        self.transforms = tr.to(d)
        from acfm_video_3d_reconstruction_b200 import OF_NeuralRenderer
        self.of_renderer = OF_NeuralRenderer(cfg["img"])

    def _targets(self, rank):
        """mask_gt = our own render of an independently drawn pose, thresholded; edt = scipy EDT of it (setup only)."""
        from scipy.ndimage import distance_transform_edt
        from acfm_video_3d_reconstruction_b200 import NeuralRenderer, synthetic
        cfg = self.cfg
        with torch.no_grad():
            cams = synthetic.cameras(cfg["frames"], 1, seed=1000 + rank).to(self.device)
            r = NeuralRenderer(cfg["img"], offset_z=cfg["offset_z"])
            r.faces_per_pixel = cfg["K"]
            m, _ = r(self.mean_v[None].repeat(cfg["frames"], 1, 1), self.faces.expand(cfg["frames"], -1, -1), cams)
            tgt = (m > 0.5).float().cpu()
        edt = torch.from_numpy(np.stack([distance_transform_edt(1 - t.numpy()) for t in tgt]).astype(np.float32))
        return tgt, edt

    def h2d(self):
        d = self.device
        return (self.h_delta.to(d, non_blocking=True), self.h_cams.to(d, non_blocking=True),
                self.h_target.to(d, non_blocking=True), self.h_edt.to(d, non_blocking=True))

    def h2d_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.h_delta, self.h_cams, self.h_target, self.h_edt))

    def d2h_bytes(self):
        return sum(t.numel() * t.element_size() for t in (self.h_loss, self.h_gdelta, self.h_gcams))

    def step(self, delta, cams, target, edt, world=1):
        from acfm_video_3d_reconstruction_b200 import deform, loss_utils, parallel
        from acfm_video_3d_reconstruction_b200 import functional as F_
        cfg = self.cfg
        delta = delta.detach().requires_grad_(True)
        cams = cams.detach().requires_grad_(True)
        self.lbs_param.grad = None
        lbs = deform.get_lbs(self.lbs_param)                       # MeshNet.get_lbs: softmax over vertices
        W = deform.skinning_matrix(lbs, self.L, solver=self.solver)
        full = cfg.get("full", False)
        cam_pred = cams
        if full:   # `cams` are the raw per-frame camera embeddings: multiframe/main.py:573-582
            from acfm_video_3d_reconstruction_b200 import camera
            cam_pred = camera.assemble_cameras(cams, self.mirror, self.transforms, 0.05)
        pred_v, ndc = deform.deform_and_project(self.mean_v, W, delta, cam_pred, offset_z=cfg["offset_z"])
        # the render with the mask-loss sums fused into its epilogue (NeuralRenderer.forward_with_losses); the full loss set
        # also takes the visible-vertex map from the render (the boundary loss below)
        if getattr(self, "lean", False) and not full:
            # lean training mode (side measurement `lean`): no fragment tensors, compact fragments between forward and backward
            mask, sums = F_.soft_silhouette_lean(ndc, self.faces, cfg["img"], target, edt, F_.BLUR_SOFT, cfg["K"], F_.SIGMA)
            out, p2f = None, None
        else:
            out = F_.soft_silhouette_losses(ndc, self.faces, cfg["img"], target, edt, F_.BLUR_SOFT, cfg["K"], F_.SIGMA, want_vis=full)
            mask, p2f, sums = out[0], out[1], out[4]
        per = loss_utils.combine_mask_losses(sums, cfg["img"] * cfg["img"], w_l1=1.0, w_edt=W_EDT)   # l1 + W_EDT * edt per render
        if full:
            G, NB, T = cfg["G"], cfg["frames"], cfg["clip_frames"]
            self.vert2kp.grad = None
            pred_proj = F_.project(pred_v, cam_pred, 0.0)          # renderer.project_points (main.py:715), xy used in place
            per = per + W_EDT * loss_utils.bds_loss(pred_proj, self.boundaries, self.faces, p2f, reduce=False, indices=self.bds_sel, visible=out[5])
            kp_verts = torch.softmax(self.vert2kp, dim=1).matmul(pred_v)                       # main.py:691-692
            per = per + loss_utils.kp_l2_loss(F_.project(kp_verts, cam_pred, 0.0), self.kps, reduction='none')
            of = loss_utils.optical_flow_loss(pred_v.repeat(G, 1, 1).reshape(G * NB // T, T, -1, 3),
                                              self.faces.expand(G * NB, -1, -1).reshape(G * NB // T, T, -1, 3), cam_pred, self.flows,
                                              self.of_renderer, None, reduce=False)[0]         # (G*B, T-1)
            per = per + 0.1 * of.mean(1).repeat_interleave(T)
        total, _ = loss_utils.hypothesis_weighting(per.view(cfg["G"], delta.shape[0]))         # multiframe/main.py:735-746
        if world > 1:
            self.kick_bucket()             # the shared-parameter bucket travels while the backward runs
        total.backward()
        if world > 1:
            self.allreduce()
            self.join_bucket()
        return total.detach(), delta.grad, cams.grad

    def allreduce(self):
        from acfm_video_3d_reconstruction_b200 import parallel
        parallel.allreduce_shared_grads([self.lbs_param])          # shared-parameter gradient


def run_ours(args):
    from acfm_video_3d_reconstruction_b200 import _lib
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nnodes=1 --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=device)
    cfg = WORKLOADS[args.workload]
    hp = HotPath(cfg, rank, device)
    if world > 1:
        hp.make_bucket(args.shared_grad_mb)
    N_r = cfg["frames"] * cfg["G"]
    fwd_b, bwd_b = alg_bytes(cfg["img"], cfg["K"], hp.wl.V, hp.wl.F)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # per-kernel events for the roofline of the dominant kernel (the forward rasterizer)
    kev = []
    _lib.event_hook = lambda name, phase: kev.append((name, phase, _rec()))

    def _rec():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    dev_in = hp.h2d()
    torch.cuda.synchronize()
    clocks = make_clocks(local) if rank == 0 else None
    for _ in range(args.warmup):
        hp.step(*dev_in, world=world)
    kev.clear()
    # ---- timed: device-resident inputs ------------------------------------------------------------------------
    barrier()
    launches0 = _lib.launches
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        hp.step(*dev_in, world=world)
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = _lib.launches - launches0
    _lib.event_hook = None
    k_ms = [a[2].elapsed_time(b[2]) for a, b in zip(kev[0::2], kev[1::2]) if a[0] == "raster_fwd"]
    kb_ms = [a[2].elapsed_time(b[2]) for a, b in zip(kev[0::2], kev[1::2]) if a[0] == "raster_bwd"]
    clk = clocks.stop(t0, t1) if clocks else None
    ms_nobucket, ar_check = None, None
    if world > 1:
        # the same K steps without the synthetic bucket (only the 82 KB lbs gradient is all-reduced), then the sums are checked
        bucket, hp.bucket = hp.bucket, None
        barrier()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0.record()
        for _ in range(args.steps):
            hp.step(*dev_in, world=world)
        n1.record()
        barrier()
        ms_nobucket = n0.elapsed_time(n1)
        hp.bucket = bucket
        ar_check = allreduce_check(hp, dev_in, rank, world, device)
    # ---- timed: end to end (pinned host -> device -> host) ----------------------------------------------------
    # The caller reads the loss every step, so the ~80 host-side launches of a step would sit on the critical path after each
    # synchronisation: the step is captured in a CUDA graph once (graphs.CapturedStep) and replayed; every step copies its
    # inputs from pinned host memory into the graph's static buffers and copies loss + gradients back, inside the timed region.
    e2e_mode = "cuda graph replay" + (" + NCCL all-reduce of the shared-parameter gradient after each replay" if world > 1 else "")
    step_fn = None
    captured = None
    try:
        from acfm_video_3d_reconstruction_b200 import graphs
        captured = graphs.CapturedStep(lambda d_, c_, t_, e_: hp.step(d_, c_, t_, e_, world=1), dev_in)

        def step_fn(*host):
            if world > 1:
                hp.kick_bucket()
            out = captured(*host)
            if world > 1:
                hp.allreduce()                                      # lbs_param.grad lives in the graph's static memory
                hp.join_bucket()
            return out
    except Exception as exc:  # capture is an optimisation, not a requirement
        e2e_mode = f"eager launches (graph capture failed: {type(exc).__name__})"
        step_fn = None
    if world > 1:   # all ranks must take the same path (the collective count per step must match)
        ok = torch.tensor([1.0 if step_fn is not None else 0.0], device=device)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        if float(ok) == 0.0:
            step_fn, e2e_mode = None, "eager launches (graph capture failed on a rank)"
    if step_fn is None:
        def step_fn(*host):
            return hp.step(*[h.to(device, non_blocking=True) for h in host], world=world)
    host_in = (hp.h_delta, hp.h_cams, hp.h_target, hp.h_edt)
    for _ in range(2):
        step_fn(*host_in)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        loss, gd, gc = step_fn(*host_in)
        hp.h_loss.copy_(loss, non_blocking=True)
        hp.h_gdelta.copy_(gd, non_blocking=True)
        hp.h_gcams.copy_(gc, non_blocking=True)
        torch.cuda.current_stream().synchronize()                  # the caller reads the loss every step
    f1.record()
    barrier()
    ms_e2e_serial = f0.elapsed_time(f1)
    ms_e2e = ms_e2e_serial
    # Same measurement with the input pipeline a training loop uses (graphs.PrefetchedStep): the pinned-host inputs of step
    # i+1 are copied in on a copy stream while step i replays.  Every timed step still performs one full H2D of its inputs
    # and the D2H read of loss + gradients, and the host still waits for the loss every step.
    if captured is not None and step_fn is not None and not e2e_mode.startswith("eager"):
        pipe = graphs.PrefetchedStep(captured)
        for _ in range(2):
            pipe.prefetch(*host_in)
            pipe.run()
            if world > 1:
                hp.allreduce()
        barrier()
        pipe.prefetch(*host_in)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            if world > 1:
                hp.kick_bucket()
            loss, gd, gc = pipe.run()
            if world > 1:
                hp.allreduce()
                hp.join_bucket()
            pipe.prefetch(*host_in)                                    # next step's inputs: overlaps this step's kernels
            hp.h_loss.copy_(loss, non_blocking=True)
            hp.h_gdelta.copy_(gd, non_blocking=True)
            hp.h_gcams.copy_(gc, non_blocking=True)
            torch.cuda.current_stream().synchronize()                  # the caller reads the loss every step
        pipe.copy_stream.synchronize()                                 # the last prefetch belongs to the timed region too
        g1.record()
        barrier()
        ms_e2e = g0.elapsed_time(g1)
        e2e_mode += "; inputs of step i+1 prefetched (pinned host -> device on a copy stream) during step i"
    c3 = c3_side(args, rank, world, device) if args.workload == "C2" and not args.no_c3 else None
    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_e2e_serial, ms_nobucket], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms, ms_e2e, ms_e2e_serial, ms_nobucket = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    peak, peak_src = peaks()
    k_avg = float(np.mean(k_ms)) if k_ms else None
    achieved = fwd_b * N_r / (k_avg * 1e-3) / 1e9 if k_avg else None
    kb_avg = float(np.mean(kb_ms)) if kb_ms else None
    achieved_b = bwd_b * N_r / (kb_avg * 1e-3) / 1e9 if kb_avg else None
    out = {
        "metric": "render fwd+bwd frames/sec (x camera hyps)", "value": N_r * world * args.steps / (ms * 1e-3),
        "unit": "renders/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "renders_per_gpu": N_r, "fragments": "materialised (API-parity mode)",
                   "l2": "per-step working set %.1f GB >> 126 MB L2 (no explicit flush)" % ((fwd_b + bwd_b) * N_r / 1e9),
                   "parallelism": f"dp{world} (frames sharded, NCCL all-reduce of shared-parameter grads)"},
        "e2e": {"value": N_r * world * args.steps / (ms_e2e * 1e-3), "unit": "renders/s",
                "h2d_bytes_per_step": hp.h2d_bytes(), "d2h_bytes_per_step": hp.d2h_bytes(), "mode": e2e_mode,
                "serial_value": N_r * world * args.steps / (ms_e2e_serial * 1e-3),
                "serial_note": "same loop without the prefetch: H2D, replay, D2H strictly one after the other"},
        "gpu_launches": launches,
        "roofline": {"kernel": "acfm_raster_fwd_losses = raster_prep_kernel, then raster_fill_kernel (TMA padding of the empty regions) with "
                               "raster_fwd_kernel (live regions, fused blend + mask-loss sums) dispatched beside it by programmatic "
                               "dependent launch, + the two loss reductions; one stream, timed around the whole call",
                     "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if achieved else None, "traffic": None, "peak_source": peak_src,
                     "alg_bytes_per_launch": fwd_b * N_r, "avg_launch_ms": k_avg, "launches_timed": len(k_ms)},
        "roofline_bwd": {"kernel": "raster_soft_bwd_kernel (+ memset of grad_ndc)", "bound": "latency (see note)", "peak": peak, "unit": "GB/s",
                         "avg_launch_ms": kb_avg, "launches_timed": len(kb_ms),
                         "frac": None, "achieved": None, "traffic": None,
                         "alg_bytes_per_launch": bwd_b * N_r, "alg_achieved": achieved_b,
                         "alg_frac": achieved_b / peak if achieved_b else None,
                         "note": "headline = `frac` (DRAM bytes of the ncu capture / launch time / measured HBM peak) and `sm_throughput_pct` "
                                 "(ncu): the kernel legitimately never reads the fragment lists of pixels with mask == 0 or zero upstream "
                                 "gradient (87 % of them), so the all-fragments denominator of SURVEY.md 8d (`alg_*`, > 1 of peak) no longer "
                                 "describes it; it is bound by shared-memory / L2 latency, not by HBM"},
        "clocks": clk,
    }
    if world > 1:
        out["comm"] = {"per_step": f"NCCL all-reduce of the lbs gradient (82 KB) after the backward + a {args.shared_grad_mb} MB synthetic "
                                   "shared-parameter bucket (SURVEY.md section 5: encoder / head gradients) all-reduced on a side stream "
                                   "from the start of the backward",
                       "shared_grad_mb": args.shared_grad_mb,
                       "value_without_bucket": N_r * world * args.steps / (ms_nobucket * 1e-3),
                       "ms_per_step_without_bucket": ms_nobucket / args.steps}
        out["allreduce_check"] = ar_check
    if c3 is not None:
        out["c3"] = c3
    traffic = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic):
        try:
            tj = json.load(open(traffic)).get(args.workload, {})
            out["roofline"]["traffic"] = tj.get("acfm_raster_fwd", tj.get("raster_fwd_kernel"))
            out["roofline_bwd"]["traffic"] = tj.get("raster_soft_bwd_kernel")
            out["roofline_bwd"]["sm_throughput_pct"] = tj.get("raster_soft_bwd_kernel_sm_throughput_pct")
            if out["roofline_bwd"]["traffic"] and kb_avg:
                out["roofline_bwd"]["achieved"] = out["roofline_bwd"]["traffic"] / (kb_avg * 1e-3) / 1e9
                out["roofline_bwd"]["frac"] = out["roofline_bwd"]["achieved"] / peak
        except Exception:
            pass
    def lean_bench():
        """Side measurement (NOT the headline, not API parity): the same C2 step with the render in lean training mode —
        mask + fused loss sums, no (N,H,W,K) fragment tensors, no padding kernel; the fragments of the live regions stay in a
        compact scratch for the backward (SURVEY.md 8d / BASELINE.md section 3 "lean mode")."""
        if cfg.get("full") or cfg["K"] != 20:
            return {"skipped": "lean mode is built for the K = 20 mask step"}
        hp.lean = True
        try:
            ref_loss = float(hp.step(*dev_in, world=1)[0])
            for _ in range(3):
                hp.step(*dev_in, world=1)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(args.steps):
                hp.step(*dev_in, world=1)
            b.record()
            torch.cuda.synchronize()
            ms_l = a.elapsed_time(b) / args.steps
        finally:
            hp.lean = False
        par_loss = float(hp.step(*dev_in, world=1)[0])
        return {"value": N_r / (ms_l * 1e-3), "unit": "renders/s (this GPU)", "ms_per_step": ms_l,
                "loss_equals_parity_mode": bool(ref_loss == par_loss),
                "note": "lean training mode: silhouette, loss sums and gradients identical to the API-parity step, but pix_to_face / zbuf / "
                        "dists are not materialised (no padding is written); reported separately, never in `value`"}

    def side(key, fn):
        # side measurements must never cost the headline line: a failure is reported in place of the number
        try:
            out[key] = fn()
        except Exception as e:  # noqa: BLE001
            out[key] = {"error": f"{type(e).__name__}: {e}"[:300]}

    side("lean", lean_bench)
    if world == 1:
        side("target_maps", lambda: target_maps_bench(hp, cfg, peak, cpu=not args.no_cpu_baseline))
        side("post_optimize", lambda: post_optimize_bench(hp, cfg, cpu=not args.no_cpu_baseline))
        side("correlation", lambda: correlation_bench(hp, peak))
        if not args.no_cpu_baseline:
            side("gpu_standin", lambda: gpu_standin_bench(hp, cfg))
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(cfg, hp.wl, hp.h_target, hp.h_edt, max_seconds=20.0)
    print(json.dumps(out), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


def allreduce_check(hp, dev_in, rank, world, device):
    """On-GPU correctness of the data-parallel step (SURVEY.md section 4 iv): (1) the all-reduced lbs gradient equals the sum
    of the all-gathered per-rank gradients; (2) the bucket all-reduce sums a known pattern exactly; (3) for 2 ranks, the sharded
    step equals ONE rank's step on the same global batch (the per-rank loss is a mean over the rank's frames, so the summed
    gradient is `world` times the global-batch gradient)."""
    dist = torch.distributed
    hp.step(*dev_in, world=1)
    g_local = hp.lbs_param.grad.detach().clone()
    parts = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(parts, g_local)
    ref = torch.zeros_like(g_local)
    for p_ in parts:
        ref += p_
    hp.step(*dev_in, world=world)
    g_ar = hp.lbs_param.grad.detach().clone()
    out = {"world": world, "grad_vs_gathered_sum_rel": float((g_ar - ref).abs().max() / ref.abs().max())}
    if hp.bucket is not None:
        hp.bucket.fill_(float(rank + 1))
        dist.all_reduce(hp.bucket)
        out["bucket_sum_exact"] = bool((hp.bucket == world * (world + 1) / 2).all())
        hp.bucket.zero_()
    if world == 2:
        G, NB = hp.cfg["G"], hp.cfg["frames"]
        gathered = []
        for t in dev_in:
            buf = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(buf, t.contiguous())
            gathered.append(buf)
        if rank == 0:
            delta = torch.cat(gathered[0])
            cams = torch.cat([torch.cat([c[g * NB:(g + 1) * NB] for c in gathered[1]]) for g in range(G)])   # hypothesis-major
            hp.step(delta, cams, torch.cat(gathered[2]), torch.cat(gathered[3]), world=1)
            g_one = hp.lbs_param.grad.detach()
            out["sharded_vs_single_rank_rel"] = float((g_ar / world - g_one).abs().max() / g_one.abs().max())
        dist.barrier()
    return out


def c3_side(args, rank, world, device):
    """Side measurement at every N: BASELINE config 3's per-GPU share (full multiframe loss set), inputs resident, same
    timing rules as the main line (barrier + synchronize around K steps, max over ranks)."""
    cfg = WORKLOADS["C3"]
    hp3 = HotPath(cfg, rank, device)
    if world > 1:
        hp3.make_bucket(args.shared_grad_mb)
    dev_in = hp3.h2d()
    for _ in range(max(3, args.warmup)):
        hp3.step(*dev_in, world=world)
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        hp3.step(*dev_in, world=world)
    e1.record()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t[0])
    N_r = cfg["frames"] * cfg["G"]
    return {"workload": cfg["desc"], "value": N_r * world * args.steps / (ms * 1e-3), "unit": "renders/s", "ms_per_step": ms / args.steps,
            "n_gpus": world}


def target_maps_bench(hp, cfg, peak, cpu=True, iters=10):
    """Side measurement (not part of `value`): the per-step target maps of set_input for this
    rank's frames — compute_dt(norm=False) + compute_dt_barrier + compute_boundaries — on the GPU vs the reference's
    CPU route (scipy EDT twice per mask + find_boundaries, serial in the main thread) on a bounded sample."""
    from acfm_video_3d_reconstruction_b200 import image_utils
    m = hp.h_target.to(hp.device)
    nb, H, W = m.shape
    for _ in range(3):
        image_utils.compute_dt_both(m)
        image_utils.compute_boundaries(m)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        image_utils.compute_dt_both(m)
        bd = image_utils.compute_boundaries(m)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    alg = nb * (H * W * (4 + 4 + 4) + bd.shape[1] * 12)          # mask in, edt + barrier out, boundary list out
    out = {"value": nb / (ms * 1e-3), "unit": "masks/s", "masks": nb, "size": [H, W], "ms": ms,
           "alg_bytes": alg, "achieved_gbs": alg / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak,
           "note": "includes the one host sync compute_boundaries needs for its data-dependent length"}
    if cpu:
        from oracle import targets_ref as tr
        mn = hp.h_target[:8].numpy()
        t = time.time()
        for x in mn:
            tr.compute_dt(x, norm=False)
            tr.compute_dt_barrier(x)
        tr.compute_boundaries(mn)
        dt = time.time() - t
        out["cpu_baseline"] = {"value": len(mn) / dt, "unit": "masks/s", "cores": 1, "kind": "port",
                               "sample": f"{len(mn)} masks, scipy.ndimage EDT x3 + restated find_boundaries, serial as in set_input"}
    return out


def correlation_bench(hp, peak, pairs=60, iters=20):
    """Side measurement: the five correlation cost volumes of one MaskFlowNet forward (pyramid levels 6..2 of
    a 256 x 256 frame pair: C = 196/128/96/64/32 at 4^2 .. 64^2, md = 4: MaskFlownet.py:97-116,266-355) for one GPU's share of a
    C3 step (4 clips x 15 adjacent-frame pairs), ours against the reference's OWN extension compiled for sm_100a
    (oracle/_ref/correlation_cuda.so), both timed with CUDA events on the same inputs."""
    from acfm_video_3d_reconstruction_b200.correlation import Correlation
    from oracle import correlation_ref as cref
    levels = [(196, 4), (128, 8), (96, 16), (64, 32), (32, 64)]
    gen = torch.Generator().manual_seed(0)
    data = [(torch.randn(pairs, C, S, S, generator=gen).to(hp.device), torch.randn(pairs, C, S, S, generator=gen).to(hp.device)) for C, S in levels]
    corr = Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)
    ref = cref.load_reference_extension()

    def timed(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    with torch.no_grad():
        ours = timed(lambda: [corr(a, b) for a, b in data])
        # the same five launches replayed from a CUDA graph: device time without the ~5 x 30 us of Python / ctypes per call
        try:
            side = torch.cuda.Stream(device=hp.device)
            side.wait_stream(torch.cuda.current_stream(hp.device))
            with torch.cuda.stream(side):
                [corr(a, b) for a, b in data]
            torch.cuda.current_stream(hp.device).wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                keep = [corr(a, b) for a, b in data]
            ours_graph = timed(g.replay)
            del keep
        except Exception:
            ours_graph = None
        per_level = [timed(lambda a=a, b=b: corr(a, b)) for a, b in data]
        top = per_level[-1]
        C, S = levels[-1]
        alg = pairs * (2 * C + 81) * S * S * 4
        out = {"unit": "ms per flow forward (5 cost volumes)", "pairs": pairs, "ms": ours, "ms_cuda_graph": ours_graph,
               "pairs_per_s": pairs / ((ours_graph or ours) * 1e-3),
               "ms_per_level_eager": {f"C{c}_{s}x{s}": t for (c, s), t in zip(levels, per_level)},
               "largest_level": {"shape": [pairs, C, S, S], "ms": top, "alg_bytes": alg, "achieved_gbs": alg / (top * 1e-3) / 1e9,
                                 "frac_of_hbm_peak": alg / (top * 1e-3) / 1e9 / peak,
                                 "tflops": 2.0 * pairs * 81 * C * S * S / (top * 1e-3) / 1e12}}
        if ref is not None:
            r = timed(lambda: [cref.reference_forward(ref, a, b, 4, 1, 4, 1, 1) for a, b in data])
            rl = [timed(lambda a=a, b=b: cref.reference_forward(ref, a, b, 4, 1, 4, 1, 1)) for a, b in data]
            out["reference_extension"] = {"ms": r, "ms_per_level_eager": {f"C{c}_{s}x{s}": t for (c, s), t in zip(levels, rl)},
                                          "kind": "reference (its own correlation_cuda_kernel.cu, compiled for sm_100a into oracle/_ref), eager",
                                          "ours_speedup_eager": r / ours}
    return out


def gpu_standin_bench(hp, cfg):
    """Side measurement: a STAND-IN for "the reference's PyTorch3D 0.3.0 CUDA path" on this same GPU, at the workload's FULL
    size (512 renders at C2), raster fwd + bwd of the mask render: oracle/pt3d_cuda_standin.{cu,py} restates PyTorch3D's coarse
    binning / one-thread-per-pixel fine / per-fragment-atomic backward kernels and runs the torch shader chain around them
    (checked against the C oracle in tests/test_standin_gpu.py).  PyTorch3D itself cannot be installed here, so this — not
    PyTorch3D — is the denominator that can be measured; two figures: max_faces_per_bin at the library default (the fine
    kernel reads all 10000 slots of a bin per pixel) and tightened to what the mesh needs (what a careful user would pass).
    The render half of OUR step (project + fused render fwd/bwd through the public API) is timed beside it."""
    from acfm_video_3d_reconstruction_b200 import functional as F_
    from oracle import pt3d_cuda_standin as sd
    if not sd.available():
        return {"unavailable": "oracle/_build/libacfm_pt3d_standin.so was not built"}
    N = cfg["frames"] * cfg["G"]
    S, K = cfg["img"], cfg["K"]
    with torch.no_grad():
        X = hp.mean_v[None].repeat(cfg["frames"], 1, 1)
        ndc0 = F_.project(X, hp.h_cams.to(hp.device), cfg["offset_z"], -1.0, -1.0, F_.EYE_Z)
    gm = torch.randn(N, S, S, device=hp.device)
    faces1 = hp.faces[0]

    def timed(fn, chunk, iters):
        fn(0, chunk)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            for k in range(0, N, chunk):
                fn(k, chunk)
        e1.record()
        torch.cuda.synchronize()
        return N * iters / (e0.elapsed_time(e1) * 1e-3)

    def standin(M):
        def fn(k, c):
            nd = ndc0[k:k + c].clone().requires_grad_(True)
            mask, _ = sd.render_mask(nd, faces1, S, F_.BLUR_SOFT, K, F_.SIGMA, max_faces_per_bin=M)
            (mask * gm[k:k + c]).sum().backward()
        return fn

    def ours(k, c):
        nd = ndc0[k:k + c].clone().requires_grad_(True)
        mask = F_.soft_silhouette(nd, hp.faces, S, F_.BLUR_SOFT, K, F_.SIGMA)[0]
        (mask * gm[k:k + c]).sum().backward()

    out = {"unit": "renders/s", "kind": "stand-in: PyTorch3D 0.3.0's CUDA rasterizer restated (oracle/pt3d_cuda_standin.cu, sm_100a -O3) + "
                                         "its torch shader chain; raster fwd+bwd only",
           "sample": f"all {N} renders of the workload"}
    # the library default: max_faces_per_bin = max(10000, V_packed / 5); bin_faces alone is N*16*16*M*4 B -> chunks of 64 renders
    out["default_max_faces_per_bin"] = timed(standin(None), 64, 1)
    # a capacity fitted to THIS workload (the fullest bin of any render, rounded up to 64): the stand-in at its best, which
    # no real caller could know in advance
    fullest = 0
    with torch.no_grad():
        for k in range(0, N, 64):
            sd.render_mask(ndc0[k:k + 64], faces1, S, F_.BLUR_SOFT, K, F_.SIGMA, max_faces_per_bin=None, check_overflow=True)
            fullest = max(fullest, sd.last_bin_max)
    tight = max(64, (fullest + 63) // 64 * 64)
    out["tight_capacity"] = tight
    out["tight_max_faces_per_bin"] = timed(standin(tight), 128, 2)
    out["ours_same_call"] = timed(ours, N, 3)
    out["value"] = out["tight_max_faces_per_bin"]
    out["ours_over_standin"] = out["ours_same_call"] / out["value"]
    return out


def post_optimize_bench(hp, cfg, cpu=True, frames=12, iters=20):
    """Side measurement (not part of `value`): the test-time post-optimisation loop (SURVEY.md 8f rank 2,
    predictor.py:287-349) at the reference's eval batch of 12 frames, 20 Adam iterations, CUDA-graph replay vs eager
    launches vs the CPU restatement (bounded to 2 iterations)."""
    from acfm_video_3d_reconstruction_b200 import image_utils
    from acfm_video_3d_reconstruction_b200.predictor import PostOptimizer
    d = hp.device
    frames = min(frames, cfg["frames"])          # cams are hypothesis-major: the first cfg["frames"] rows are hypothesis 0
    masks = hp.h_target[:frames].to(d)
    edts = image_utils.compute_dt_barrier(masks)
    bds = image_utils.compute_boundaries(masks)
    lbs = torch.softmax(hp.lbs_param.detach(), dim=0)
    delta = hp.h_delta[:frames].to(d)
    cams = hp.h_cams[:frames].to(d)
    gen = torch.Generator().manual_seed(0)
    sel = torch.stack([torch.randperm(bds.shape[1], generator=gen)[:1000] for _ in range(iters)]).to(d)
    out = {"frames": frames, "iters": iters, "unit": "ms per iteration"}
    for name, graph in (("eager", False), ("cuda_graph", True)):
        po = PostOptimizer(img_size=cfg["img"], offset_z=cfg["offset_z"], num_optim_iter=iters, of_loss_wt=0.0, use_cuda_graph=graph)
        po.renderer.faces_per_pixel = cfg["K"]
        po.run(hp.mean_v, lbs, hp.L, delta, cams, masks, edts, bds, hp.faces, sample_indices=sel)       # warm-up
        torch.cuda.synchronize()
        t = time.time()
        r = po.run(hp.mean_v, lbs, hp.L, delta, cams, masks, edts, bds, hp.faces, sample_indices=sel)
        torch.cuda.synchronize()
        out[name] = (time.time() - t) * 1e3 / iters
        out["loss_first_last"] = [float(r["losses"][0]), float(r["losses"][-1])]
    if cpu:
        from oracle import predictor_ref
        t = time.time()
        predictor_ref.post_optimize(hp.mean_v.cpu(), lbs.cpu(), hp.L.cpu(), delta.cpu(), cams.cpu(), masks.cpu(), edts.cpu(), bds.cpu(),
                                    hp.faces.cpu().expand(frames, -1, -1).contiguous(), sel[:2].cpu(), cfg["img"], offset_z=cfg["offset_z"])
        out["cpu_baseline"] = {"value": (time.time() - t) * 1e3 / 2, "unit": "ms per iteration", "kind": "port",
                               "sample": "2 iterations, oracle/predictor_ref.py (reference's batched solve + restated CPU rasterizer)"}
    return out


# ---------------------------------------------------------------------------------------------------------------
# CPU arms (the oracle as checker-turned-baseline: the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------------------
def _cpu_targets(wl, cfg, frames, threads):
    """mask_gt / edt of the sampled frames for the CPU arm (set-up, untimed): an oracle render of an independently drawn pose."""
    from scipy.ndimage import distance_transform_edt
    from acfm_video_3d_reconstruction_b200 import synthetic
    from oracle import pt3d_oracle as orc
    cams = synthetic.cameras(cfg["frames"], 1, seed=1000).numpy()[frames]
    X = wl.mean_v.numpy()[None].repeat(len(frames), 0)
    faces = wl.faces.numpy()[None].repeat(len(frames), 0)
    m = orc.neural_renderer_mask(X, faces, cams, img_size=cfg["img"], offset_z=cfg["offset_z"], K=cfg["K"], threads=threads)["mask"]
    tgt = (m > 0.5).astype(np.float32)
    edt = np.stack([distance_transform_edt(1 - t) for t in tgt]).astype(np.float32)
    return torch.from_numpy(tgt), torch.from_numpy(edt)


def _oracle_step(wl, cfg, frames, threads, target, edt):
    """The reference's CPU implementation of one hot-path step over `frames` (all G hypotheses of each): the reference's
    own deformation block (multiframe/main.py:586-609: per-frame batched Cholesky, torch CPU), projection, the restated
    PyTorch3D 0.3.0 naive CPU rasterizer + blend (oracle/, OpenMP), mask losses, hypothesis weighting, and the backward of
    all of it down to the handle offsets, cameras and handle weights."""
    from oracle import pt3d_oracle as orc
    from oracle import torch_ref
    G, nb, FT = cfg["G"], len(frames), cfg["frames"]
    lbs_param = wl.lbs_param.clone().requires_grad_(True)
    delta = wl.delta[frames].clone().requires_grad_(True)
    rows = torch.cat([g * FT + torch.as_tensor(frames) for g in range(G)])
    cams = wl.cams[rows].clone().requires_grad_(True)
    # ---- deformation exactly as the reference batches it (B*T identical V x V systems) ----
    lbs = torch.softmax(lbs_param, dim=0).t()[None].repeat(nb, 1, 1)          # (nb,Kh,V)
    mean = wl.mean_v[None].repeat(nb, 1, 1)
    delta_v = lbs.bmm(mean) + delta
    Lb = wl.L[None].repeat(nb, 1, 1)
    A_augm = Lb.permute(0, 2, 1).matmul(Lb) + lbs.permute(0, 2, 1).matmul(lbs)
    rhs = Lb.permute(0, 2, 1) @ torch.bmm(Lb, mean) + lbs.permute(0, 2, 1) @ delta_v
    pred_v = torch.cholesky_solve(rhs, torch.linalg.cholesky(A_augm))
    # ---- render (C oracle) ----
    ndc_t = torch_ref.to_ndc(pred_v.repeat(G, 1, 1), cams, cfg["offset_z"])
    faces = wl.faces.numpy()[None].repeat(G * nb, 0)
    fr = orc.rasterize(ndc_t.detach().numpy(), faces, cfg["img"], orc.BLUR_SOFT, cfg["K"], threads=threads, want_bary=False)
    fr["ndc"] = ndc_t.detach().numpy()
    mask = torch.from_numpy(orc.sigmoid_alpha_blend(fr["dists"], fr["pix_to_face"])).requires_grad_(True)
    # ---- losses + hypothesis weighting (torch CPU) ----
    per = (torch_ref.l1_loss(mask, target.repeat(G, 1, 1)) + W_EDT * torch_ref.edt_loss(mask, edt.repeat(G, 1, 1)[:, None])).view(G, nb)
    total, _ = torch_ref.hypothesis_weighting(per)
    total.backward()
    # ---- backward: blend + rasterizer (C oracle), then projection / deformation (torch autograd) ----
    g_ndc = orc.neural_renderer_mask_backward(fr, faces, mask.grad.numpy())
    ndc_t.backward(torch.from_numpy(g_ndc))
    return float(total.detach()), delta.grad, cams.grad, lbs_param.grad


def cpu_baseline(cfg, wl, target, edt, max_seconds=20.0):
    from oracle import pt3d_oracle as orc
    threads = orc.max_threads()
    torch.set_num_threads(threads)
    G = cfg["G"]
    t = time.time()
    _oracle_step(wl, cfg, [0], threads, target[:1], edt[:1])  # page-in / thread pool warm-up + cost estimate
    est = time.time() - t
    nf = int(max(1, min(cfg["frames"], max_seconds / max(est, 1e-6))))
    frames = list(range(nf))
    t = time.time()
    _oracle_step(wl, cfg, frames, threads, target[:nf], edt[:nf])
    dt = time.time() - t
    return {"value": nf * G / dt, "unit": "renders/s", "cores": threads, "kind": "port",
            "sample": f"{nf} frames x {G} hypotheses = {nf * G} renders of the same workload: reference deformation block (batched "
                      f"Cholesky, torch CPU) + oracle/ restatement of PyTorch3D 0.3.0's naive CPU rasterizer (OpenMP over render,row) "
                      f"+ blend + mask losses + hypothesis weighting, fwd+bwd, {dt:.1f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    from acfm_video_3d_reconstruction_b200 import synthetic
    from oracle import pt3d_oracle as orc
    orc.build()
    cfg = WORKLOADS[args.workload]
    wl = synthetic.Workload(cfg["template"], cfg["frames"], cfg["G"], cfg["handles"], cfg["img"], seed=0, offset_z=cfg["offset_z"])
    # rank 0 runs alone and may use every host core (torchrun exports OMP_NUM_THREADS=1 to its workers: override it)
    threads = max(orc.max_threads(), os.cpu_count() or 1)
    torch.set_num_threads(threads)
    G = cfg["G"]
    nf = max(1, min(cfg["frames"], max(1, threads // G) if cfg["img"] <= 256 else 1))
    frames = list(range(nf))
    target, edt = _cpu_targets(wl, cfg, frames, threads)
    for _ in range(args.warmup):
        _oracle_step(wl, cfg, frames[:1], threads, target[:1], edt[:1])
    t = time.time()
    for _ in range(args.steps):
        _oracle_step(wl, cfg, frames, threads, target, edt)
    dt = time.time() - t
    n = nf * G
    val = n * args.steps / dt
    sample = (f"each step = {nf} frames x {G} hypotheses = {n} of the workload's {wl.renders} renders/GPU (bounded sample): reference "
              f"deformation block (torch CPU) + restated PyTorch3D 0.3.0 CPU rasterizer (oracle/, OpenMP {threads} threads) + losses, "
              f"fwd+bwd; the real PyTorch3D wheel is not installable offline")
    print(json.dumps({
        "impl": "reference", "metric": "render fwd+bwd frames/sec (x camera hyps)", "value": val, "unit": "renders/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg["desc"], "renders_per_step_sample": n},
        "cpu_baseline": {"value": val, "unit": "renders/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "renders/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--shared-grad-mb", type=float, default=60.0,
                    help="N>1: size of the synthetic shared-parameter gradient bucket all-reduced every step (0 = none)")
    ap.add_argument("--no-c3", action="store_true", help="skip the C3 (full multiframe loss set) side measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
