import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acfm_video_3d_reconstruction_b200 import _lib, functional as F_, synthetic
wl = synthetic.Workload("bird", 16, 8, 32, 256, seed=0)
X = wl.mean_v[None].repeat(16, 1, 1).cuda()
ndc = F_.project(X, wl.cams.cuda(), 5.0, -1.0, -1.0, F_.EYE_Z)
faces = wl.faces[None].cuda()
N = ndc.shape[0]
gen = torch.Generator(device="cuda").manual_seed(1)
fr = F_._train_render(ndc, faces, 256, F_.BLUR_SOFT, 20, F_.SIGMA, False)
gm = torch.randn(N, 256, 256, device="cuda", generator=gen)
saved = (fr["ndc"], faces, fr["pix_to_face"], fr["dists"], fr["mask"], None, None)
outs = {}
for name in sys.argv[1:]:
    L = ctypes.CDLL(os.path.abspath(f".variants/lib_{name}.so"))
    g = torch.empty_like(ndc)
    fn = L.acfm_raster_soft_bwd
    fn.argtypes = _lib.SIGNATURES["acfm_raster_soft_bwd"]
    st = fn(_lib.ptr(ndc), _lib.ptr(faces), 1, 0, N, ndc.shape[1], faces.shape[1], 256, 256, 20, float(F_.SIGMA), _lib.ptr(fr["pix_to_face"]),
            _lib.ptr(fr["dists"]), _lib.ptr(fr["mask"]), _lib.ptr(gm), _lib.ptr(g), None, None)
    torch.cuda.synchronize()
    assert st == 0
    outs[name] = g.clone()
ref = outs[sys.argv[1]]
for k, v in outs.items():
    d = (v - ref).abs()
    print(k, "max abs", float(d.max()), "ref max", float(ref.abs().max()), "rel", float(d.max() / ref.abs().max()), "nan", int(torch.isnan(v).sum()))
