"""Scratch timing of the handle-solve kernels (acfm_handle_solve_fwd / _bwd) with CUDA events: us per forward and per
forward + backward, eager launches and CUDA-graph replay.  usage: time_solver.py [template] [handles]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acfm_video_3d_reconstruction_b200 import deform, synthetic

name = sys.argv[1] if len(sys.argv) > 1 else "bird"
Kh = int(sys.argv[2]) if len(sys.argv) > 2 else 32
wl = synthetic.Workload(name, 1, 1, Kh, 64, seed=0)
solver = deform.HandleSolver(wl.L.cuda())
lbs = torch.softmax(wl.lbs_param, 0).cuda()
g = torch.randn_like(lbs)


def fwd():
    return solver(lbs.detach().requires_grad_(True))


def both():
    x = lbs.detach().requires_grad_(True)
    (solver(x) * g).sum().backward()
    return x.grad


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print(f"{name} V={lbs.shape[0]} Kh={Kh}: eager fwd {timed(fwd):.1f} us, fwd+bwd {timed(both):.1f} us", flush=True)
gr = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    both()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr, stream=s):
        both()
print(f"  graph replay fwd+bwd {timed(gr.replay):.1f} us")
