"""Generates the committed golden fixtures in tests/golden/ by IMPORTING THE REFERENCE
(/root/reference, python) in the authoring container.  /root/reference does not exist on the GPU
box, so tests only ever read the .npz files written here.

  python tests/golden/make_golden.py            # everything
  python tests/golden/make_golden.py reproj      # only the named sections (base | reproj | targets | priors | cameras)

What is pinned to the reference's own code:
  templates.npz   verts/faces of the reference's template meshes (monocular/meshes/bird_aligned.obj,
                  multiframe/meshes/horse.obj), parsed from the OBJ text
  projection.npz  nnutils/geom_utils.py: orthographic_proj_withz / orthographic_proj / quat_rotate
  losses.npz      nnutils/loss_utils.py: l1_loss, iou_loss, edt_loss, kp_l2_loss (torch CPU, seeded inputs)
  reproj.npz      nnutils/loss_utils.py: bds_loss, optical_flow_loss (values + fp64 gradients), run by the reference's
                  own code on visibility maps rendered by oracle/; nnutils/geom_utils.py: mesh_laplacian(.., 'cot')
                  through a duck-typed Meshes; the hypothesis weighting lines of multiframe/main.py:735-746
  targets.npz     utils/image.py: compute_dt, compute_dt_barrier, compute_boundaries run by the reference's own code
                  (scipy present; cv2 stubbed — unused on this path; skimage absent: find_boundaries is restated from its
                  published implementation, grey_dilation != grey_erosion over the connectivity-1 footprint)
  priors.npz      nnutils/loss_utils.py: locally_rigid_fn run by the reference's own code on a duck-typed packed Meshes;
                  mesh_laplacian_smoothing(method="cot"): PyTorch3D 0.3.0's few lines restated around the reference's own
                  geom_utils.laplacian_cot (values + fp64 gradients)
  cameras.npz     multiframe/main.py: mirror_cameras, transform_cameras and the camera-assembly block of ShapeTrainer.forward
                  (:573-582) EXECUTED from the reference's source text (values + fp64 gradients); the three
                  pytorch3d.transforms functions the mirror branch calls are restated from the published v0.3.0 code
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


class _Meshes:
    """The three accessors geom_utils.mesh_laplacian / laplacian_cot use on a PyTorch3D Meshes."""

    def __init__(self, v, f):
        self.v, self.f, self.device = v, f, v.device

    def isempty(self):
        return False

    def verts_packed(self):
        return self.v

    def faces_packed(self):
        return self.f


class _RefOFRenderer:
    """What optical_flow_loss needs from OF_NeuralRenderer: proj_fn (the reference's own) and a K=1 visibility
    render of already-projected points (oracle/: PyTorch3D itself is not installable)."""

    def __init__(self, img_size):
        self.img_size = img_size
        self.proj_fn = ref_geom.orthographic_proj_withz

    def __call__(self, verts, faces):
        fr = orc.of_renderer(verts.detach().float().numpy(), faces.numpy(), img_size=self.img_size)
        return torch.from_numpy(fr["pix_to_face"])


def reproj(horse_v, horse_f):
    gen = torch.Generator().manual_seed(7)
    out = {}
    # ---- bds_loss: G=2 hypotheses x 2 frames, soft K=20 render for pix_to_face (oracle), boundary points random
    S, N, NB, P = 64, 4, 2, 300
    V, F = horse_v.shape[0], horse_f.shape[0]
    X = torch.from_numpy(horse_v)[None].repeat(N, 1, 1) + 0.02 * torch.randn(N, V, 3, generator=gen)
    cam = synth_cams(N, gen)
    faces = torch.from_numpy(horse_f)[None].repeat(N, 1, 1)
    fr = orc.neural_renderer_mask(X.numpy(), faces.numpy(), cam.numpy(), img_size=S, offset_z=0.0)
    p2f = torch.from_numpy(fr["pix_to_face"])
    bds = torch.cat([torch.rand(NB, P, 2, generator=gen) * 1.6 - 0.8, (torch.rand(NB, P, 1, generator=gen) > 0.2).float()], -1)
    proj = ref_geom.orthographic_proj_withz(X, cam, 0.0)[:, :, :2]
    torch.manual_seed(123)
    sel = torch.randperm(P)[:200]
    torch.manual_seed(123)
    loss = ref_loss.bds_loss(proj, bds.repeat(N // NB, 1, 1), faces, p2f, reduce=False, n_samples=200)
    pd = proj.double().requires_grad_(True)
    torch.manual_seed(123)
    w = torch.rand(N, generator=gen).double()
    (ref_loss.bds_loss(pd, bds.repeat(N // NB, 1, 1).double(), faces, p2f, reduce=False, n_samples=200) * w).sum().backward()
    out.update(bds_X=X.numpy(), bds_cam=cam.numpy(), bds_p2f=fr["pix_to_face"][..., :2].astype(np.int32), bds_proj=proj.numpy(),
               bds_pts=bds.numpy(), bds_sel=sel.numpy(), bds_loss=loss.numpy(), bds_w=w.numpy(), bds_grad=pd.grad.numpy(), bds_seed=123)
    # ---- optical_flow_loss: B' = 2 sequences (one flow field shared: the G-fold repeat), T = 3 frames
    B, T, H = 2, 3, 64
    M = torch.from_numpy(horse_v)[None, None].repeat(B, T, 1, 1) + 0.03 * torch.randn(B, T, V, 3, generator=gen)
    cams = synth_cams(B, gen)[:, None].repeat(1, T, 1)
    cams[:, :, 1:3] += 0.05 * torch.randn(B, T, 2, generator=gen)
    cams = cams.reshape(B * T, 7)
    flows = 3.0 * torch.randn(1, T, H, H, 2, generator=gen)
    flows = flows * (torch.rand(1, T, H, H, 1, generator=gen) > 0.3).float()   # zero flow = "no measurement"
    faces_of = torch.from_numpy(horse_f)[None, None].repeat(B, T, 1, 1)
    r = _RefOFRenderer(H)
    l, of_pred, visv, pts, smp = ref_loss.optical_flow_loss(M, faces_of, cams, flows.repeat(B, 1, 1, 1, 1), r, None, reduce=False)
    Md = M.double().requires_grad_(True)
    cd = cams.double().requires_grad_(True)
    w2 = torch.rand(B, T - 1, generator=gen).double()
    l64 = ref_loss.optical_flow_loss(Md, faces_of, cd, flows.repeat(B, 1, 1, 1, 1).double(), r, None, reduce=False)[0]
    (l64 * w2).sum().backward()
    out.update(of_meshes=M.numpy(), of_cams=cams.numpy(), of_flows=flows.numpy(), of_loss=l.numpy(), of_pred=of_pred.numpy(),
               of_vis=visv.numpy(), of_pts=pts.numpy(), of_samples=smp.numpy(), of_w=w2.numpy(), of_grad_meshes=Md.grad.numpy(),
               of_grad_cams=cd.grad.numpy())
    # ---- cotangent Laplacian of the template (reference code, duck-typed Meshes)
    hv = torch.from_numpy(horse_v)
    out["lap_cot"] = ref_geom.mesh_laplacian(_Meshes(hv, torch.from_numpy(horse_f)), "cot").numpy()
    # ---- hypothesis weighting (multiframe/main.py:735-746)
    tl = torch.rand(4, 6, generator=gen).double().requires_grad_(True)
    probs = torch.softmax(-tl, dim=0).detach()
    tot = (tl * probs).sum(0).mean()
    tot.backward()
    out.update(hyp_loss=tl.detach().numpy(), hyp_probs=probs.numpy(), hyp_total=tot.detach().numpy(), hyp_grad=tl.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "reproj.npz"), **out)


def targets():
    import scipy.ndimage as ndi
    sys.modules.setdefault("cv2", types.ModuleType("cv2"))
    sk, seg = types.ModuleType("skimage"), types.ModuleType("skimage.segmentation")

    def find_boundaries(label_img, connectivity=1, mode="thick", background=0):
        """skimage.segmentation.find_boundaries, mode='thick' (skimage/segmentation/boundaries.py)."""
        if label_img.dtype == "bool":
            label_img = label_img.astype(np.uint8)
        fp = ndi.generate_binary_structure(label_img.ndim, connectivity)
        return ndi.grey_dilation(label_img, footprint=fp) != ndi.grey_erosion(label_img, footprint=fp)

    seg.find_boundaries = find_boundaries
    sk.segmentation = seg
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.segmentation", seg)
    sys.path.insert(0, os.path.join(REF, "multiframe"))
    from utils import image as ref_image
    t = np.load(os.path.join(HERE, "templates.npz"))
    gen = torch.Generator().manual_seed(11)
    N, S = 3, 64
    X = torch.from_numpy(t["horse_v"])[None].repeat(N, 1, 1).numpy()
    cam = synth_cams(N, gen).numpy()
    faces = np.repeat(t["horse_f"][None], N, 0)
    m = (orc.neural_renderer_mask(X, faces, cam, img_size=S, offset_z=0.0)["mask"] > 0.5).astype(np.float32)
    m[0, 20:24, 30:34] = 0          # a hole
    m[1, 0, :] = 1                  # touches the border
    out = {"masks": m}
    out["dt_raw"] = np.stack([ref_image.compute_dt(x, norm=False) for x in m])
    out["dt_norm"] = np.stack([ref_image.compute_dt(x) for x in m])
    out["barrier"] = np.stack([ref_image.compute_dt_barrier(x) for x in m])
    out["boundaries"] = ref_image.compute_boundaries(m)
    # non-square, non-power-of-two map; an empty and a full mask (scipy's degenerate-input behaviour)
    r = (torch.rand(3, 40, 72, generator=gen) > 0.97).float().numpy()
    r[1] = 0
    r[2] = 1
    out["masks_b"] = r
    out["dt_raw_b"] = np.stack([ref_image.compute_dt(x, norm=False) for x in r])
    out["dt_norm_b"] = np.stack([ref_image.compute_dt(x) for x in r])
    out["barrier_b"] = np.stack([ref_image.compute_dt_barrier(x, k=20) for x in r])
    out["boundaries_b"] = ref_image.compute_boundaries(r)
    np.savez_compressed(os.path.join(HERE, "targets.npz"), **out)


class _PackedMeshes:
    """Batch of N meshes sharing one topology, with the packed accessors locally_rigid_fn / laplacian_cot use."""

    def __init__(self, verts, faces):
        self.N, self.V = verts.shape[0], verts.shape[1]
        self.v, self.f, self.device = verts, faces, verts.device
        e = torch.cat([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]], 0)
        self.e = torch.unique(torch.sort(e, dim=1)[0], dim=0)

    def __len__(self):
        return self.N

    def isempty(self):
        return False

    def verts_packed(self):
        return self.v.reshape(-1, 3)

    def faces_packed(self):
        return (self.f[None] + torch.arange(self.N)[:, None, None] * self.V).reshape(-1, 3)

    def edges_packed(self):
        return (self.e[None] + torch.arange(self.N)[:, None, None] * self.V).reshape(-1, 2)


def _laplacian_smoothing_cot(meshes):
    """pytorch3d.loss.mesh_laplacian_smoothing(meshes, method="cot"), v0.3.0, around the reference's laplacian_cot."""
    N = len(meshes)
    verts_packed = meshes.verts_packed()
    weights = torch.full((verts_packed.shape[0],), 1.0 / meshes.V, dtype=verts_packed.dtype)
    with torch.no_grad():   # the reference's laplacian_cot is float32-only; the weights are constants of the loss
        L, _ = ref_geom.laplacian_cot(_PackedMeshes(meshes.v.detach().float(), meshes.f))
        norm_w = torch.sparse.sum(L, dim=1).to_dense().view(-1, 1)
        idx = norm_w > 0
        norm_w[idx] = 1.0 / norm_w[idx]
        L, norm_w = L.to(verts_packed.dtype), norm_w.to(verts_packed.dtype)
    loss = L.mm(verts_packed) * norm_w - verts_packed
    return (loss.norm(dim=1) * weights).sum() / N


def priors():
    t = np.load(os.path.join(HERE, "templates.npz"))
    gen = torch.Generator().manual_seed(21)
    hv, hf = torch.from_numpy(t["horse_v"]), torch.from_numpy(t["horse_f"])
    N = 3
    X = hv[None].repeat(N, 1, 1) + 0.03 * torch.randn(N, hv.shape[0], 3, generator=gen)
    out = {"X": X.numpy()}
    rigid = ref_loss.locally_rigid_fn(_PackedMeshes(X, hf), _PackedMeshes(hv[None].repeat(N, 1, 1), hf))
    out["rigid"] = rigid.numpy()
    Xd, td = X.double().requires_grad_(True), hv.double().requires_grad_(True)
    ref_loss.locally_rigid_fn(_PackedMeshes(Xd, hf), _PackedMeshes(td[None].repeat(N, 1, 1), hf)).backward()
    out.update(rigid_grad_X=Xd.grad.numpy(), rigid_grad_t=td.grad.numpy())
    out["smooth"] = _laplacian_smoothing_cot(_PackedMeshes(X, hf)).numpy()
    Xd = X.double().requires_grad_(True)
    sm = _laplacian_smoothing_cot(_PackedMeshes(Xd, hf))
    sm.backward()
    out.update(smooth64=sm.detach().numpy(), smooth_grad_X=Xd.grad.numpy())
    np.savez_compressed(os.path.join(HERE, "priors.npz"), **out)


# ---- camera multiplex assembly ---------------------------------------------------------------------
def _pt3d_transforms():
    """pytorch3d.transforms pieces that mirror_cameras star-imports (multiframe/main.py:37), restated from the published
    v0.3.0 rotation_conversions.py.  Everything else in this section is the reference's own source."""
    def standardize_quaternion(q):
        return torch.where(q[..., 0:1] < 0, -q, q)

    def quaternion_raw_multiply(a, b):
        aw, ax, ay, az = torch.unbind(a, -1)
        bw, bx, by, bz = torch.unbind(b, -1)
        return torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                            aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)

    def quaternion_multiply(a, b):
        return standardize_quaternion(quaternion_raw_multiply(a, b))

    def _copysign(a, b):
        return torch.where((a < 0) != (b < 0), -a, a)

    def _sqrt_positive_part(x):
        ret = torch.zeros_like(x)
        ret[x > 0] = torch.sqrt(x[x > 0])
        return ret

    def matrix_to_quaternion(m):
        m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
        o0 = 0.5 * _sqrt_positive_part(1 + m00 + m11 + m22)
        x = 0.5 * _sqrt_positive_part(1 + m00 - m11 - m22)
        y = 0.5 * _sqrt_positive_part(1 - m00 + m11 - m22)
        z = 0.5 * _sqrt_positive_part(1 - m00 - m11 + m22)
        return torch.stack((o0, _copysign(x, m[..., 2, 1] - m[..., 1, 2]), _copysign(y, m[..., 0, 2] - m[..., 2, 0]),
                            _copysign(z, m[..., 1, 0] - m[..., 0, 1])), -1)

    return dict(standardize_quaternion=standardize_quaternion, quaternion_multiply=quaternion_multiply,
                matrix_to_quaternion=matrix_to_quaternion)


def cameras():
    """cameras.npz: cam_pred produced by EXECUTING the reference's own lines — the functions mirror_cameras and
    transform_cameras (multiframe/main.py:113-138, taken out of the module with ast because main.py itself imports absl
    flags, PyTorch3D, visdom ...) and the assembly block of ShapeTrainer.forward (main.py:573-582, run against stand-in
    `self` / `opts` objects) — values and fp64 gradients."""
    import ast
    import textwrap
    src = open(os.path.join(REF, "multiframe", "main.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "F": torch.nn.functional}
    ns.update(_pt3d_transforms())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("mirror_cameras", "transform_cameras"):
            exec(compile(ast.Module([node], []), "multiframe/main.py", "exec"), ns)
    lines = src.splitlines()
    # the `else:` body (plain 7-vector embeddings; every documented command runs without --az_el_cam) and the lines after it
    q = max(i for i, l in enumerate(lines) if l.strip() == "quats = cameras[..., 3:]")   # the one in forward() (:573)
    assert lines[q - 1].strip() == "cameras = cameras.reshape(opts.num_guesses, -1, 7)"
    cat = max(i for i, l in enumerate(lines) if l.strip() == "self.cam_pred = torch.cat([scales, translations, quats_n], dim=2)")
    last = max(i for i, l in enumerate(lines) if "self.transforms.repeat(opts.num_guesses, 1))" in l)
    assert q + 4 == cat and last == cat + 5
    block = textwrap.dedent("\n".join(lines[q - 1:cat])) + "\n" + textwrap.dedent("\n".join(lines[cat:last + 1]))
    assert "F.relu(opts.scale_lr_decay" in block and "mirror_cameras(" in block and "transform_cameras(" in block
    gen = torch.Generator().manual_seed(11)
    G, NB = 8, 6
    raw = torch.randn(G, NB, 7, generator=gen)
    raw[0, 0, 0] = -40.0                    # relu clamps the scale
    raw[1, 2, 3] = -abs(raw[1, 2, 3])       # negative real part: standardize_quaternion flips it under the mirror
    import hashlib
    out = {"raw": raw.numpy(), "scale_lr_decay": np.float64(0.05),
           "block_sha1": np.array(hashlib.sha1(block.encode()).hexdigest())}   # which lines ran (not the lines themselves)
    w = torch.randn(G * NB, 7, generator=gen)
    for tag, mirror, tf in (
            ("plain", torch.zeros(NB), torch.cat([torch.ones(NB, 1), torch.zeros(NB, 3)], 1)),
            ("affine", torch.zeros(NB), torch.cat([torch.rand(NB, 1, generator=gen) + 0.5, torch.randn(NB, 2, generator=gen) * 0.1,
                                                  (torch.rand(NB, 1, generator=gen) > 0.4).float()], 1)),
            ("mirror", (torch.rand(NB, generator=gen) > 0.4).float(),
             torch.cat([torch.rand(NB, 1, generator=gen) + 0.5, torch.randn(NB, 2, generator=gen) * 0.1,
                        (torch.rand(NB, 1, generator=gen) > 0.4).float()], 1))):
        res = {}
        for dt in (torch.float32, torch.float64):
            cams = raw.to(dt).clone().requires_grad_(True)
            me = types.SimpleNamespace(input_imgs=torch.zeros(NB, 3, 8, 8), mirror_flag=mirror.to(dt), transforms=tf.to(dt))
            opts = types.SimpleNamespace(num_guesses=G, scale_lr_decay=0.05)
            env = dict(ns, self=me, opts=opts, cameras=cams)
            exec(block, env)
            res[dt] = (me.cam_pred, cams)
        (res[torch.float64][0] * w.double()).sum().backward()
        out.update({f"{tag}_mirror": mirror.numpy(), f"{tag}_transforms": tf.numpy(),
                    f"{tag}_cam_pred": res[torch.float32][0].detach().numpy(),
                    f"{tag}_cam_pred64": res[torch.float64][0].detach().numpy(),
                    f"{tag}_grad_raw": res[torch.float64][1].grad.numpy()})
    out["grad_w"] = w.numpy()
    np.savez_compressed(os.path.join(HERE, "cameras.npz"), **out)


def main():
    sections = set(sys.argv[1:]) or {"base", "reproj", "targets", "priors", "cameras"}
    if "cameras" in sections:
        cameras()
    if "priors" in sections:
        priors()
    if "targets" in sections:
        targets()
    if "reproj" in sections:
        t = np.load(os.path.join(HERE, "templates.npz")) if "base" not in sections else None
        if t is not None:
            reproj(t["horse_v"], t["horse_f"])
    if "base" not in sections:
        print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))
        return
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
    if "reproj" in sections:
        reproj(horse_v, horse_f)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
