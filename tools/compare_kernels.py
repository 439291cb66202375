"""Times the brick-batched stiffness kernel against the simple cell-coloured kernel (64^3, P4, fp64)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wave_fenics_b200 as wfx  # noqa: E402

P, N = 4, int(sys.argv[1]) if len(sys.argv) > 1 else 64
mesh = wfx.create_box_hex(N, P, (0.1,) * 3, perturb=0.15)
geo = wfx.Geometry(mesh, P)
x = torch.randn(mesh.ndofs, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
for name, mode in (("brick", wfx.capi.STIFF_AUTO), ("cell-colour", wfx.capi.STIFF_CELL_COLOUR)):
    op = wfx.StiffnessOperator(mesh, P, geometry=geo, mode=mode)
    for _ in range(5):
        op.apply(x, y, beta=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        op.apply(x, y, beta=0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name:12s} {ms:.3f} ms/apply  {mesh.ndofs / ms / 1e6:.2f} GDoF/s  (y = K x, beta = 0, {op.info()['nlaunches']} launches)")
