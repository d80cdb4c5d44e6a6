"""GPU parity of the handle deformation + multiplex projection against the reference's own formulation
(batched solve of (L^T L + A^T A) X = L^T L m + A^T (A m + D), /root/reference/multiframe/main.py:586-609)
evaluated in fp64 on the CPU."""
import numpy as np
import pytest
import torch

from oracle import pt3d_oracle as orc
from tests import util

pytestmark = pytest.mark.gpu


def _reference_block_fp64(mean_v, lbs, L, delta_res):
    """Literal restatement of main.py:586-608 (fp64, CPU)."""
    NB = delta_res.shape[0]
    lbs_ = lbs.t()[None].repeat(NB, 1, 1)
    mean = mean_v[None].repeat(NB, 1, 1)
    delta_v = lbs_.bmm(mean) + delta_res
    Lb = L[None].repeat(NB, 1, 1)
    delta = torch.bmm(Lb, mean)
    A_augm = Lb.permute(0, 2, 1).matmul(Lb) + lbs_.permute(0, 2, 1).matmul(lbs_)
    b = Lb.permute(0, 2, 1) @ delta + lbs_.permute(0, 2, 1) @ delta_v
    u = torch.linalg.cholesky(A_augm)
    return torch.cholesky_solve(b, u)


def test_deform_and_project_vs_reference_block():
    from acfm_video_3d_reconstruction_b200 import deform, synthetic
    wl = synthetic.Workload("bird", frames=5, G=3, handles=16, seed=3)
    lbs = torch.softmax(wl.lbs_param, dim=0)
    ref = _reference_block_fp64(wl.mean_v.double(), lbs.double(), wl.L.double(), wl.delta.double())

    mean_v = wl.mean_v.cuda().requires_grad_(True)
    lbs_c = lbs.cuda().requires_grad_(True)
    delta = wl.delta.cuda().requires_grad_(True)
    cams = wl.cams.cuda().requires_grad_(True)
    W = deform.skinning_matrix(lbs_c, wl.L.cuda())
    pred_v, ndc = deform.deform_and_project(mean_v, W, delta, cams, offset_z=5.0)
    # reference's own fp32 batched Cholesky is only good to ~4e-5 absolute: compare to fp64 truth
    assert np.abs(pred_v.detach().cpu().numpy() - ref.numpy()).max() < 5e-5
    # the projection of OUR pred_v is bit-exact w.r.t. the reference projection code
    ndc_ref = orc.view(orc.project(np.tile(pred_v.detach().cpu().numpy(), (3, 1, 1)), wl.cams.numpy(), 5.0), yflip=True)
    assert np.array_equal(ndc.detach().cpu().numpy(), ndc_ref)
    assert np.array_equal(deform.deform(mean_v, W, delta).detach().cpu().numpy(), pred_v.detach().cpu().numpy())

    # gradients vs fp64 autograd through the reference block + projection restatement
    from oracle import torch_ref
    gen = torch.Generator().manual_seed(0)
    w_ndc = torch.randn(ndc.shape, generator=gen)
    w_pv = torch.randn(pred_v.shape, generator=gen)
    ((ndc * w_ndc.cuda()).sum() + (pred_v * w_pv.cuda()).sum()).backward()
    md = wl.mean_v.double().requires_grad_(True)
    ld = lbs.double().requires_grad_(True)
    dd = wl.delta.double().requires_grad_(True)
    cd = wl.cams.double().requires_grad_(True)
    pv = _reference_block_fp64(md, ld, wl.L.double(), dd)
    nd = torch_ref.to_ndc(pv.repeat(3, 1, 1), cd, 5.0)
    ((nd * w_ndc.double()).sum() + (pv * w_pv.double()).sum()).backward()
    assert util.rel_err(delta.grad.cpu().numpy(), dd.grad.numpy()) < 1e-3
    assert util.rel_err(cams.grad.cpu().numpy(), cd.grad.numpy()) < 1e-3
    assert util.rel_err(mean_v.grad.cpu().numpy(), md.grad.numpy()) < 1e-3
    assert util.rel_err(lbs_c.grad.cpu().numpy(), ld.grad.numpy()) < 1e-3


def test_handle_solver_matches_direct_solve():
    """HandleSolver (constant Laplacian: one fp64 inverse at init, Woodbury update per step) against the fp64 direct solve,
    values and gradient w.r.t. the handle weights; it is also more accurate than the fp32 Cholesky route."""
    from acfm_video_3d_reconstruction_b200 import deform, synthetic
    wl = synthetic.Workload("horse", frames=2, G=1, handles=32, seed=5)
    lbs = torch.softmax(wl.lbs_param, dim=0)
    L = wl.L.cuda()
    solver = deform.HandleSolver(L)
    assert solver.ok
    lc = lbs.cuda().requires_grad_(True)
    W = deform.skinning_matrix(lc, L, solver=solver)
    g = torch.randn(W.shape, generator=torch.Generator().manual_seed(0))
    (W * g.cuda()).sum().backward()
    ld = lbs.double().requires_grad_(True)
    M = wl.L.double().t() @ wl.L.double() + ld @ ld.t()
    W64 = torch.cholesky_solve(ld, torch.linalg.cholesky(M))
    (W64 * g.double()).sum().backward()
    assert util.rel_err(W.detach().cpu().numpy(), W64.detach().numpy()) < 1e-6
    assert util.rel_err(lc.grad.cpu().numpy(), ld.grad.numpy()) < 1e-5
    W32 = deform.skinning_matrix(lbs.cuda(), L)
    assert util.rel_err(W.detach().cpu().numpy(), W64.detach().numpy()) <= util.rel_err(W32.cpu().numpy(), W64.detach().numpy())
    # a Laplacian whose null space is not the constants (here: not a Laplacian at all) falls back to the direct route
    bad = deform.HandleSolver(torch.eye(L.shape[0], device="cuda"))
    assert not bad.ok
    assert torch.allclose(deform.skinning_matrix(lbs.cuda(), torch.eye(L.shape[0], device="cuda"), solver=bad),
                          deform.skinning_matrix(lbs.cuda(), torch.eye(L.shape[0], device="cuda")))


@pytest.mark.parametrize("name,Kh", [("bird", 16), ("horse", 64), ("ico4", 8)])
def test_handle_solve_kernels_other_sizes(name, Kh):
    """acfm_handle_solve_fwd / _bwd (csrc/handle_solve.cu) at the reference's other handle counts (16 / 64) and on the
    2562-vertex template: against the fp64 direct solve; two calls give bit-identical results (deterministic reductions);
    the pivot check reports a regular system."""
    from acfm_video_3d_reconstruction_b200 import _lib, deform, synthetic
    wl = synthetic.Workload(name, frames=1, G=1, handles=Kh, seed=3)
    lbs = torch.softmax(wl.lbs_param, dim=0)
    L = wl.L.cuda()
    solver = deform.HandleSolver(L)
    assert solver.ok
    g = torch.randn(lbs.shape, generator=torch.Generator().manual_seed(1))
    outs = []
    for _ in range(2):
        lc = lbs.cuda().requires_grad_(True)
        W = solver(lc)
        (W * g.cuda()).sum().backward()
        outs.append((W.detach().clone(), lc.grad.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ld = lbs.double().requires_grad_(True)
    M = wl.L.double().t() @ wl.L.double() + ld @ ld.t()
    W64 = torch.cholesky_solve(ld, torch.linalg.cholesky(M))
    (W64 * g.double()).sum().backward()
    assert util.rel_err(outs[0][0].cpu().numpy(), W64.detach().numpy()) < 1e-6
    assert util.rel_err(outs[0][1].cpu().numpy(), ld.grad.numpy()) < 1e-5
    ws = solver.workspace(Kh)
    Wt = torch.empty(lbs.shape, device="cuda")
    lc = lbs.cuda().contiguous()
    _lib.check(_lib.lib().acfm_handle_solve_fwd(_lib.ptr(solver.Pinv), _lib.ptr(solver.Pinv_ones), _lib.ptr(lc), lc.shape[0], Kh,
                                                solver.c / solver.V, _lib.ptr(Wt), _lib.ptr(ws), ws.numel(), _lib.stream_of(lc)),
               "acfm_handle_solve_fwd")
    assert _lib.lib().acfm_handle_solve_singular(_lib.ptr(ws), lc.shape[0], Kh, _lib.stream_of(lc)) == 0
    assert torch.equal(Wt, outs[0][0])


def test_get_lbs_softmax_over_vertices():
    from acfm_video_3d_reconstruction_b200 import deform
    gen = torch.Generator().manual_seed(2)
    for V, K in ((642, 32), (2562, 16), (7, 1)):
        x = (3 * torch.randn(V, K, generator=gen)).requires_grad_(True)
        w = torch.randn(V, K, generator=gen)
        (torch.softmax(x.double(), dim=0) * w.double()).sum().backward()
        xc = x.detach().cuda().requires_grad_(True)
        y = deform.get_lbs(xc)
        (y * w.cuda()).sum().backward()
        assert util.rel_err(y.detach().cpu().numpy(), torch.softmax(x.detach().double(), 0).numpy()) < 1e-6
        assert util.rel_err(xc.grad.cpu().numpy(), x.grad.numpy()) < 1e-5
        assert torch.allclose(y.sum(0), torch.ones(K, device="cuda"), atol=1e-5)
