import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
import wave_fenics_b200 as wfx
def t(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for P, N in ((2, 128), (4, 64)):
    mesh = wfx.create_box_hex(N, P, (0.1,)*3, perturb=0.15)
    for dt, tdt in ((np.float32, torch.float32),):
        geo = wfx.Geometry(mesh, P, dt)
        op = wfx.StiffnessOperator(mesh, P, dtype=dt, geometry=geo)
        mass = wfx.MassOperator(mesh, P, dtype=dt, geometry=geo)
        x = torch.randn(mesh.ndofs, dtype=tdt, device="cuda"); y = torch.empty_like(x)
        ms = t(lambda: op.apply_scaled(x, mass.inverse_diagonal_ptr(), y))
        ms2 = t(lambda: op.apply(x, y, beta=0))
        try: ki = op.kernel_info()
        except Exception as e: ki = str(e)
        print(P, 'f32', 'scaled %.4f' % ms, 'plain %.4f' % ms2, ki, op.info()['ncolours'], flush=True)
