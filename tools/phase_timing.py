"""Per-phase cycle breakdown of stiff_brick_kernel (needs the -DWFX_TIMING build):
WFX_LIB=$PWD/wave-fenics_b200/libwavefx_timing.so python tools/phase_timing.py"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wave_fenics_b200 as wfx

P, N = 4, 64
perturb = float(sys.argv[1]) if len(sys.argv) > 1 else 0.15   # 0: affine fast path
mesh = wfx.create_box_hex(N, P, (0.1,) * 3, perturb=perturb)
geo = wfx.Geometry(mesh, P)
stiff = wfx.StiffnessOperator(mesh, P, geometry=geo)
mass = wfx.MassOperator(mesh, P, geometry=geo)
x = torch.randn(mesh.ndofs, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
nb, W = 4096, 8
buf = torch.zeros(nb * W * 12, dtype=torch.int64, device="cuda")
f = wfx.capi.lib.wfx_debug_set_timing_buffer
f.argtypes = [C.c_void_p]
f.restype = C.c_int
assert f(C.c_void_p(buf.data_ptr())) == 0
for _ in range(3):
    buf.zero_()
    stiff.apply_scaled(x, mass.inverse_diagonal_ptr(), y)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(nb * W, 12).astype(np.float64)
names = ["0 staging", "1 gather+sync", "2 transform1+sync", "3 Gmult(+G wait)", "4 G prefetch issue",
         "5 sync+transform2+sync", "6 combine+scatter", "7 round barrier", "8 pdl wait", "9 writeback", "10 staging: issue", "11 staging: copies landed"]
t = t[t.sum(axis=1) > 0]          # the single-launch kernel fills one row per resident warp only
tot = t[:, :12].sum(axis=1)
print(f"perturb {perturb}  kernel {stiff.kernel_info()}")
print(f"per-warp total cycles: mean {tot.mean():.0f}")
for i, n in enumerate(names):
    print(f"  {n:26s} mean {t[:, i].mean():9.0f} cyc  {100 * t[:, i].mean() / tot.mean():5.1f}%   per cell {t[:, i].mean() / 8:7.0f}")
