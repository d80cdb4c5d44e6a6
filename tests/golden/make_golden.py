"""Generates the committed golden fixtures in tests/golden/ by IMPORTING THE REFERENCE
(/root/reference, python) in the authoring container.  /root/reference does not exist on the GPU
box, so tests only ever read the .npz files written here.

  python tests/golden/make_golden.py

What is pinned to the reference's own code:
  templates.npz   verts/faces of the reference's template meshes (monocular/meshes/bird_aligned.obj,
                  multiframe/meshes/horse.obj), parsed from the OBJ text
  projection.npz  nnutils/geom_utils.py: orthographic_proj_withz / orthographic_proj / quat_rotate
  losses.npz      nnutils/loss_utils.py: l1_loss, iou_loss, edt_loss, kp_l2_loss, bds_loss,
                  optical_flow_loss (torch CPU, seeded inputs)
What is NOT pinned upstream (PyTorch3D 0.3.0 is not installable: parity unpinned):
  raster_small.npz  fragments / masks / gradients from oracle/ itself — a regression pin of the
                    restated algorithm only.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(REF, "multiframe"))
sys.modules.setdefault("lpips", types.ModuleType("lpips"))  # loss_utils imports it at module scope only
from nnutils import geom_utils as ref_geom  # noqa: E402
from nnutils import loss_utils as ref_loss  # noqa: E402

from oracle import pt3d_oracle as orc  # noqa: E402


def load_obj(path):
    v, f = [], []
    for line in open(path):
        t = line.split()
        if not t:
            continue
        if t[0] == "v":
            v.append([float(x) for x in t[1:4]])
        elif t[0] == "f":
            f.append([int(x.split("/")[0]) - 1 for x in t[1:4]])
    return np.asarray(v, np.float32), np.asarray(f, np.int64)


def synth_cams(n, gen):
    q = torch.nn.functional.normalize(torch.randn(n, 4, generator=gen), dim=-1)
    s = torch.rand(n, 1, generator=gen) * 0.3 + 0.55
    t = torch.rand(n, 2, generator=gen) * 0.2 - 0.1
    return torch.cat([s, t, q], 1)


def main():
    gen = torch.Generator().manual_seed(0)
    bird_v, bird_f = load_obj(os.path.join(REF, "monocular/meshes/bird_aligned.obj"))
    horse_v, horse_f = load_obj(os.path.join(REF, "multiframe/meshes/horse.obj"))
    np.savez_compressed(os.path.join(HERE, "templates.npz"), bird_v=bird_v, bird_f=bird_f, horse_v=horse_v,
                        horse_f=horse_f)

    # ---- projection (pinned) ----
    N = 6
    X = torch.from_numpy(bird_v)[None].repeat(N, 1, 1) + 0.02 * torch.randn(N, bird_v.shape[0], 3, generator=gen)
    cam = synth_cams(N, gen)
    cam[1, 3:] *= 1.7  # un-normalised quaternion: the reference does not normalise inside quat_rotate
    out = {"X": X.numpy(), "cam": cam.numpy()}
    for oz in (0.0, 5.0):
        out[f"withz_{int(oz)}"] = ref_geom.orthographic_proj_withz(X, cam, offset_z=oz).numpy()
    out["proj"] = ref_geom.orthographic_proj(X, cam).numpy()
    out["quat_rotate"] = ref_geom.quat_rotate(X, cam[:, 3:]).numpy()
    # gradients through the reference projection (fp64 for truth)
    Xd = X.double().requires_grad_(True)
    cd = cam.double().requires_grad_(True)
    w = torch.randn(N, bird_v.shape[0], 3, generator=gen).double()
    (ref_geom.orthographic_proj_withz(Xd, cd, offset_z=5.0) * w).sum().backward()
    out.update(grad_w=w.numpy(), grad_X=Xd.grad.numpy(), grad_cam=cd.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "projection.npz"), **out)

    # ---- simple per-render losses (pinned) ----
    B, S = 5, 32
    pred = torch.rand(B, S, S, generator=gen)
    targ = (torch.rand(B, S, S, generator=gen) > 0.5).float()
    edt = torch.rand(B, 1, S, S, generator=gen) * 3
    kp_pred = torch.rand(B, 15, 2, generator=gen) * 2 - 1
    kp_gt = torch.cat([torch.rand(B, 15, 2, generator=gen) * 2 - 1, (torch.rand(B, 15, 1, generator=gen) > 0.3).float()], -1)
    np.savez_compressed(
        os.path.join(HERE, "losses.npz"), pred=pred.numpy(), targ=targ.numpy(), edt=edt.numpy(), kp_pred=kp_pred.numpy(),
        kp_gt=kp_gt.numpy(),
        l1=ref_loss.l1_loss(pred, targ, reduce=False).numpy(), l1_mean=ref_loss.l1_loss(pred, targ).numpy(),
        iou=ref_loss.iou_loss(pred, targ, reduce=False).numpy(), iou_mean=ref_loss.iou_loss(pred, targ).numpy(),
        edt_l=ref_loss.edt_loss(pred, edt, reduce=False).numpy(), edt_mean=ref_loss.edt_loss(pred, edt).numpy(),
        kp=ref_loss.kp_l2_loss(kp_pred, kp_gt, reduction="none").numpy(), kp_mean=ref_loss.kp_l2_loss(kp_pred, kp_gt).numpy())

    # ---- rasterizer regression pin (oracle self-golden; parity unpinned upstream) ----
    N = 2
    Xr = (torch.from_numpy(bird_v)[None].repeat(N, 1, 1)).numpy()
    camr = synth_cams(N, gen).numpy()
    faces = np.repeat(bird_f[None], N, 0)
    fr = orc.neural_renderer_mask(Xr, faces, camr, img_size=64, offset_z=5.0)
    gm = torch.randn(N, 64, 64, generator=gen).numpy().astype(np.float32)
    g_ndc = orc.neural_renderer_mask_backward(fr, faces, gm)
    np.savez_compressed(os.path.join(HERE, "raster_small.npz"), X=Xr, cam=camr, faces=bird_f, ndc=fr["ndc"],
                        pix_to_face=fr["pix_to_face"].astype(np.int32), zbuf=fr["zbuf"], dists=fr["dists"], mask=fr["mask"],
                        grad_mask=gm, grad_ndc=g_ndc)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
