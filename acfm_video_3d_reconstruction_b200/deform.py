"""Handle-based deformation of the template ("LBS") — drop-in for the block the reference inlines in
ShapeTrainer.forward (/root/reference/multiframe/main.py:586-609, monocular/main.py:203-218) and
MeshPredictor.forward (multiframe/nnutils/predictor.py:257-276).

The reference factorises B*T identical V x V systems per step.  Because
  L^T L m + A^T (A m + D) = (L^T L + A^T A) m + A^T D,
the solve collapses to pred_v = mean_v + W D with W = (L^T L + A^T A)^-1 A^T: one
V x V Cholesky per step (torch, differentiable w.r.t. the handle weights), then a fused sm_100a kernel
for the per-frame contraction and the camera-multiplex projection.
"""
import torch

from . import _lib
from . import functional as F_


class _SkinningMatrix(torch.autograd.Function):
    """W = (L^T L + lbs lbs^T)^-1 lbs with a closed-form backward: with Z = M^-1 gW (one more solve against the same
    factor),  g_lbs = Z - Z (W^T lbs) - W (Z^T lbs).  torch's autograd through cholesky + cholesky_solve reaches the
    same values through ~100 V x V triangular-solve launches per step."""

    @staticmethod
    def forward(ctx, lbs, LtL):
        M = torch.addmm(LtL, lbs, lbs.t())              # A_augm, main.py:605  (A = lbs^T)
        u = torch.linalg.cholesky(M)
        W = torch.cholesky_solve(lbs, u)                # (V,Kh)
        ctx.save_for_backward(lbs, u, W)
        return W

    @staticmethod
    def backward(ctx, gW):
        lbs, u, W = ctx.saved_tensors
        Z = torch.cholesky_solve(gW.contiguous(), u)
        g = Z - Z.matmul(W.t().matmul(lbs)) - W.matmul(Z.t().matmul(lbs))
        return g, None


class _GetLbs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lbs_param):
        _lib.require_cuda(lbs_param)
        x = F_._f32c(lbs_param)
        V, K = x.shape
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            st = _lib.lib().acfm_softmax_cols_fwd(_lib.ptr(x), V, K, _lib.ptr(y), _lib.stream_of(x))
        _lib.check(st, "acfm_softmax_cols_fwd")
        _lib.count()
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        y, = ctx.saved_tensors
        V, K = y.shape
        gx = torch.empty_like(y)
        with torch.cuda.device(y.device):
            st = _lib.lib().acfm_softmax_cols_bwd(_lib.ptr(y), _lib.ptr(F_._f32c(g)), V, K, _lib.ptr(gx), _lib.stream_of(y))
        _lib.check(st, "acfm_softmax_cols_bwd")
        _lib.count()
        return gx


def get_lbs(lbs_param):
    """MeshNet.get_lbs (mesh_net.py:597-599): softmax of the (V,Kh) handle-weight parameter over the VERTICES (dim 0)."""
    return _GetLbs.apply(lbs_param)


class HandleSolver:
    """Per-step solve for the skinning matrix W = (L^T L + lbs lbs^T)^-1 lbs when the Laplacian L is constant across steps
    (monocular: built once at init, monocular/main.py:124; multiframe: rebuilt every forward from a template that does not
    change unless --symmetric, multiframe/main.py:600-601).

    L^T L is then a constant PSD matrix whose null space is the constant vector (L 1 = 0 on a connected mesh), and the only
    per-step part, lbs lbs^T, has rank K_h.  With P~ = L^T L + (c/V) 1 1^T (full rank; inverted ONCE, here, in fp64) and
    U~ = [lbs, 1], S = diag(I, -c/V):   M = P~ + U~ S U~^T  and Woodbury gives  W = (P~^-1 U~) C^-1 [:, :K_h],
    C = S^-1 + U~^T P~^-1 U~.  The per-step part runs in libacfm_b200 (csrc/handle_solve.cu: acfm_handle_solve_fwd / _bwd,
    fp64, 4 + 7 small kernels, deterministic) instead of a V x V Cholesky factorisation and two V x V triangular solves per FRAME in
    the reference.  fp64 keeps the result at the accuracy of the direct fp64 solve (cond(P~) ~ 1e4-1e5); W is returned in
    fp32.  Differentiable w.r.t. lbs (closed-form backward)."""

    def __init__(self, L):
        Ld = L.detach().double()
        V = Ld.shape[0]
        P = Ld.t().matmul(Ld)
        self.c = float(torch.trace(P)) / V
        self.V = V
        self.ok = False
        ones = torch.ones((V, 1), dtype=torch.float64, device=L.device)
        Pt = P + (self.c / V) * ones.matmul(ones.t())
        null_resid = float(Ld.matmul(ones).abs().max())
        try:
            self.Pinv = torch.linalg.inv(Pt).contiguous()
            self.Pinv_ones = self.Pinv.sum(1).contiguous()          # Pinv 1: the constant last column of Pinv [lbs, 1]
            resid = float((self.Pinv.matmul(Pt) - torch.eye(V, dtype=torch.float64, device=L.device)).abs().max())
            self.ok = bool(torch.isfinite(self.Pinv).all()) and resid < 1e-6 and null_resid < 1e-5 and self.c > 0 and L.is_cuda
        except RuntimeError:
            self.ok = False

    def workspace(self, Kh):
        n = int(_lib.lib().acfm_handle_solve_workspace_bytes(self.V, Kh))
        return torch.empty((n,), dtype=torch.uint8, device=self.Pinv.device)

    def __call__(self, lbs):
        return _SolverMatrix.apply(lbs, self)


class _SolverMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lbs, solver):
        _lib.require_cuda(lbs)
        lbs = F_._f32c(lbs.detach())
        V, Kh = lbs.shape
        if V != solver.V:
            raise ValueError(f"lbs has {V} rows, the solver was built for {solver.V} vertices")
        ws = solver.workspace(Kh)
        W = torch.empty((V, Kh), dtype=torch.float32, device=lbs.device)
        with torch.cuda.device(lbs.device):
            st = _lib.lib().acfm_handle_solve_fwd(_lib.ptr(solver.Pinv), _lib.ptr(solver.Pinv_ones), _lib.ptr(lbs), V, Kh, solver.c / V,
                                                  _lib.ptr(W), _lib.ptr(ws), ws.numel(), _lib.stream_of(lbs))
        _lib.check(st, "acfm_handle_solve_fwd")
        _lib.count(4)
        ctx.solver, ctx.ws = solver, ws
        ctx.save_for_backward(lbs)
        return W

    @staticmethod
    def backward(ctx, gW):
        lbs, = ctx.saved_tensors
        V, Kh = lbs.shape
        gW = F_._f32c(gW)
        g = torch.empty_like(lbs)
        with torch.cuda.device(lbs.device):
            st = _lib.lib().acfm_handle_solve_bwd(_lib.ptr(ctx.solver.Pinv), _lib.ptr(lbs), _lib.ptr(gW), V, Kh, _lib.ptr(g),
                                                  _lib.ptr(ctx.ws), ctx.ws.numel(), _lib.stream_of(lbs))
        _lib.check(st, "acfm_handle_solve_bwd")
        _lib.count(7)
        return g, None


def skinning_matrix(lbs, L, LtL=None, solver=None):
    """lbs (V,Kh): softmax-over-vertices handle weights (MeshNet.get_lbs, mesh_net.py:597-599);
    L (V,V): dense mesh Laplacian (geom_utils.mesh_laplacian; a constant: the reference builds it under no_grad).
    LtL: optional precomputed L^T L (it only changes when the template does).
    solver: optional HandleSolver(L) built once for a constant L — replaces the per-step factorisation by two GEMMs.
    Returns W (V,Kh), differentiable w.r.t. lbs."""
    if solver is not None and solver.ok:
        return solver(lbs)
    if LtL is None:
        LtL = L.detach().t().matmul(L.detach())
    return _SkinningMatrix.apply(lbs, LtL)


class _SkinProject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mean_v, W, delta, cams, offset_z, sx, sy, z_add):
        _lib.require_cuda(mean_v, W, delta, cams)
        mean_v, W, delta = F_._f32c(mean_v), F_._f32c(W), F_._f32c(delta)
        V, Kh = W.shape
        NB = delta.shape[0]
        if mean_v.shape != (V, 3) or delta.shape != (NB, Kh, 3):
            raise ValueError(f"expected mean_v (V,3), W (V,Kh), delta (NB,Kh,3); got {tuple(mean_v.shape)}, "
                             f"{tuple(W.shape)}, {tuple(delta.shape)}")
        dev = mean_v.device
        pred_v = torch.empty((NB, V, 3), dtype=torch.float32, device=dev)
        ndc, G = None, 0
        if cams is not None:
            cams = F_._f32c(cams)
            if cams.dim() != 2 or cams.shape[1] != 7 or NB == 0 or cams.shape[0] % NB:
                raise ValueError(f"cams must be (G*NB,7), got {tuple(cams.shape)} for NB={NB}")
            G = cams.shape[0] // NB
            ndc = torch.empty((G * NB, V, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = _lib.lib().acfm_skin_project_fwd(_lib.ptr(mean_v), _lib.ptr(W), _lib.ptr(delta), _lib.ptr(cams), NB, G,
                                                  V, Kh, offset_z, sx, sy, z_add, _lib.ptr(pred_v), _lib.ptr(ndc),
                                                  _lib.stream_of(mean_v))
        _lib.check(st, "acfm_skin_project_fwd")
        _lib.count()
        ctx.save_for_backward(W, delta, cams, pred_v)
        ctx.sx, ctx.sy = sx, sy
        ctx.set_materialize_grads(False)
        if ndc is None:
            return pred_v
        return pred_v, ndc

    @staticmethod
    def backward(ctx, g_pred, g_ndc=None):
        W, delta, cams, pred_v = ctx.saved_tensors
        NB, V, _ = pred_v.shape
        Kh = W.shape[1]
        dev = pred_v.device
        L = _lib.lib()
        gp = None
        g_cams = None
        with torch.cuda.device(dev):
            if cams is not None and g_ndc is not None:
                gp = torch.empty_like(pred_v)
                g_cams = torch.empty_like(cams) if ctx.needs_input_grad[3] else None
                st = L.acfm_project_bwd(_lib.ptr(pred_v), _lib.ptr(cams), _lib.ptr(F_._f32c(g_ndc)), cams.shape[0], NB, V,
                                        ctx.sx, ctx.sy, _lib.ptr(gp), _lib.ptr(g_cams), _lib.stream_of(pred_v))
                _lib.check(st, "acfm_project_bwd")
                _lib.count()
            if g_pred is not None:
                gp = F_._f32c(g_pred) if gp is None else gp + g_pred
            if gp is None:
                return None, None, None, g_cams, None, None, None, None
            g_mean = torch.empty((V, 3), dtype=torch.float32, device=dev) if ctx.needs_input_grad[0] else None
            g_W = torch.empty_like(W) if ctx.needs_input_grad[1] else None
            g_delta = torch.empty_like(delta) if ctx.needs_input_grad[2] else None
            st = L.acfm_skin_bwd(_lib.ptr(W), _lib.ptr(delta), _lib.ptr(gp), NB, V, Kh, _lib.ptr(g_delta), _lib.ptr(g_W),
                                 _lib.ptr(g_mean), _lib.stream_of(pred_v))
            _lib.check(st, "acfm_skin_bwd")
            _lib.count(2)
        return g_mean, g_W, g_delta, g_cams, None, None, None, None


def deform(mean_v, W, delta_v):
    """pred_v (NB,V,3) = mean_v + W @ delta_v[b]  — the result of the reference's cholesky_solve (main.py:607-608)."""
    return _SkinProject.apply(mean_v, W, delta_v, None, 0.0, 1.0, 1.0, 0.0)


def deform_and_project(mean_v, W, delta_v, cams, offset_z=0.0, sx=-1.0, sy=-1.0, z_add=F_.EYE_Z):
    """Fused deformation + G-hypothesis projection.  cams (G*NB,7) hypothesis-major as in main.py:578.
    Returns pred_v (NB,V,3) and the rasterizer-space vertices ndc (G*NB,V,3) that
    NeuralRenderer.to_ndc(pred_v.repeat(G,1,1), cams) would produce."""
    return _SkinProject.apply(mean_v, W, delta_v, cams, float(offset_z), float(sx), float(sy), float(z_add))


def handle_deform_reference_form(mean_v, lbs, L, delta_v_res):
    """Same inputs as the reference block: mean_v (V,3), lbs (V,Kh) = model.get_lbs(), L (V,V),
    delta_v_res (NB,Kh,3) -> pred_v (NB,V,3)."""
    return deform(mean_v, skinning_matrix(lbs, L), delta_v_res)
