"""Small-mesh workloads for tools/sanitize.sh (one case per invocation, results checked against
the per-cell kernel so that a sanitizer-clean run is also a correct one)."""
import copy
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wave_fenics_b200 as wfx  # noqa: E402

L = 0.1


def check(mesh, P, dtype=np.float64, **kw):
    geo = wfx.Geometry(mesh, P, dtype=dtype)
    op = wfx.StiffnessOperator(mesh, P, dtype=dtype, geometry=geo, **kw)
    ref = wfx.StiffnessOperator(mesh, P, dtype=dtype, geometry=geo, mode=wfx.capi.STIFF_CELL_COLOUR)
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    x = torch.randn(mesh.ndofs, dtype=tdt, device="cuda")
    y, yr = torch.empty_like(x), torch.empty_like(x)
    op.apply(x, y, beta=0)
    ref.apply(x, yr, beta=0)
    err = float((y - yr).norm() / yr.norm())
    assert err < (1e-13 if dtype == np.float64 else 1e-5), err
    return op, geo, x, y


def main(case):
    if case in ("regular", "generic"):
        op, *_ = check(wfx.create_box_hex(5, 4, (L,) * 3, perturb=0.15), 4)
        print(case, op.kernel_info())
    elif case == "affine":
        op, *_ = check(wfx.create_box_hex(5, 4, (L,) * 3), 4)
        assert op.kernel_info()["affine"]
    elif case == "mixed":
        base = wfx.create_box_hex(8, 4, (L,) * 3, perturb=0.15)
        mesh = copy.copy(base)
        dup = np.array([3, 200])
        mesh.xdofs = np.concatenate([base.xdofs, base.xdofs[dup]])
        mesh.dofmap = np.concatenate([base.dofmap, base.ndofs + np.arange(2 * 125, dtype=np.int32).reshape(2, 125)])
        mesh.ndofs = base.ndofs + 250
        op, *_ = check(mesh, 4)
        assert op.kernel_info()["mixed"]
    elif case == "parts":
        mesh = copy.copy(wfx.create_box_hex(6, 4, (L,) * 3, perturb=0.15))
        M = 25
        mesh.halo = {"send_indices": (np.arange(M * M) + 12 * M * M).astype(np.int32), "recv_indices": np.zeros(0, dtype=np.int32)}
        op, geo, x, y = check(mesh, 4)
        y2 = torch.empty_like(y)
        op.apply_part(x, y2, 0)
        op.apply_part(x, y2, 1)
        assert torch.equal(y, y2)
    elif case == "cell":
        check(wfx.create_box_hex(4, 3, (L,) * 3, perturb=0.15), 3, mode=wfx.capi.STIFF_CELL_COLOUR)
    elif case == "p2":
        check(wfx.create_box_hex(9, 2, (L,) * 3, perturb=0.15), 2)
    elif case == "p6":
        check(wfx.create_box_hex(3, 6, (L,) * 3, perturb=0.15), 6)
    elif case == "fp32":
        check(wfx.create_box_hex(5, 4, (L,) * 3, perturb=0.15), 4, dtype=np.float32)
    elif case in ("rk4", "rk4_nograph"):
        mesh = wfx.create_box_hex(4, 4, (L,) * 3, perturb=0.15)
        eqn = wfx.LinearGLLOpt(mesh, None, 4, 1500.0, 0.5e6, 6e4)
        eqn.init()
        eqn.rk4(0.0, 1.0, wfx.cfl_timestep(mesh.h_min, 1500.0, 4, 0.5e6), max_steps=4)
        u, v = eqn.get_state()
        assert np.isfinite(u).all() and np.abs(u).max() > 0
    elif case == "setup":
        mesh = wfx.create_box_hex(4, 4, (L,) * 3, perturb=0.15)
        geo = wfx.Geometry(mesh, 4)
        wfx.MassOperator(mesh, 4, geometry=geo).diagonal()
        wfx.BoundaryOperator(mesh, 4).facet_masses()
        geo.scale_cells(np.full(mesh.ncells, 1.5))
    else:
        raise SystemExit(f"unknown case {case}")
    torch.cuda.synchronize()
    print("case ok:", case)


if __name__ == "__main__":
    main(sys.argv[1])
