"""Launches the correlation cost volume at the flow network's largest level a few times (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acfm_video_3d_reconstruction_b200.correlation import Correlation
C, S, B = (int(x) for x in (sys.argv[1:4] + ["32", "64", "60"][len(sys.argv) - 1:]))
a, b = torch.randn(B, C, S, S, device="cuda"), torch.randn(B, C, S, S, device="cuda")
corr = Correlation(4, 1, 4, 1, 1)
with torch.no_grad():
    for _ in range(5):
        out = corr(a, b)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
