"""RK4 stepping rate on a small mesh (BASELINE configs[0] scale), where the step is launch-bound:
compare WFX_WAVE_GRAPH=1 (CUDA-graph replay of the step, default) with WFX_WAVE_GRAPH=0.
  python tools/rk4_small_mesh.py [cells_per_axis] [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wave_fenics_b200 as wfx

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 400
P = 4
mesh = wfx.create_box_hex(N, P, (0.1,) * 3)
eqn = wfx.LinearGLLOpt(mesh, None, P, 1500.0, 0.5e6, 6e4)
eqn.init()
dt = wfx.cfl_timestep(mesh.h_min, 1500.0, P, 0.5e6)
eqn.rk4(0.0, 1.0, dt, max_steps=20)
torch.cuda.synchronize()
t0 = time.perf_counter()
eqn.rk4(20 * dt, 1.0, dt, max_steps=steps)
torch.cuda.synchronize()
el = time.perf_counter() - t0
u, _ = eqn.get_state()
print(f"graph={os.environ.get('WFX_WAVE_GRAPH', '1')} cells {N}^3 dofs {mesh.ndofs}: {el / steps * 1e6:.1f} us/step, "
      f"|u|max {abs(u).max():.6e}")
