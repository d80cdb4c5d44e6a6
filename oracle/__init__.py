"""oracle/ — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's render-and-reproject path (C: raster_oracle.c/.inc; numpy/torch:
pt3d_oracle.py, torch_ref.py).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import anything from here.  Rasterizer parity is UNPINNED (PyTorch3D 0.3.0
is not installable); projection and losses are pinned by tests/golden/ (generated from the reference).
"""
