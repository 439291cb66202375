"""Generates tests/golden/reference_*.npz: outputs of the REFERENCE's OWN CODE run in this container.

The reference's cell kernels, operator call loops and RK4 time stepper (common/operators.hpp,
common/LinearGLL.hpp) are cut out of its headers and compiled into oracle/_ref/libwfref_cpu.so by
oracle/build_ref.py (container stand-ins: oracle/ref_cpu_shim.cpp); this script runs that library on small
seeded problems and stores what it returns: y = K x, m = M 1 and the state (u, v) after an RK4 run from rest
and from a random state.  Inputs are regenerated from the seeds below; the geometric factors, the basis table,
the permutation and the facet masses handed to the reference code come from the oracle (the reference gets
them from Basix / DOLFINx / FFCx, absent here).  The fixtures travel where /root/reference does not.

    python tests/golden/make_reference_golden.py        # rewrites the fixtures (needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

L, C0, F0, P0 = 0.1, 1500.0, 0.5e6, 6e4
# name -> (degree, cells per axis, vertex perturbation, dof renumbering seed or None, RK4 steps, last-step fraction)
CASES = {
    "reference_p4_n2_perturbed": (4, (2, 2, 2), 0.15, None, 10, 0.0),
    "reference_p2_n3_renumbered": (2, (3, 3, 2), 0.15, 9, 14, 0.3),
    "reference_p3_ragged_affine": (3, (3, 2, 1), 0.0, None, 9, 0.0),
}


def setup(wfx, orc, name):
    P, shape, perturb, renumber, steps, frac = CASES[name]
    mesh = wfx.create_box_hex(shape, P, tuple(L * s / 8 for s in shape), perturb=perturb, renumber=renumber)
    G, detJ = orc.precompute_geometric_data(mesh, P)
    m1, m2 = orc.boundary_facet_mass(mesh, P)
    rng = np.random.default_rng(77)
    x = rng.standard_normal(mesh.ndofs)
    u0, v0 = rng.standard_normal(mesh.ndofs), 1e3 * rng.standard_normal(mesh.ndofs)
    dt = wfx.cfl_timestep(mesh.h_min, C0, P, F0)
    return P, mesh, G, detJ, m1, m2, x, u0, v0, dt, (steps + frac) * dt


def compute(wfx, orc, name, apply_k, apply_m, run_rk4):
    """The stored quantities of one case through the given implementations (reference code or oracle)."""
    P, mesh, G, detJ, m1, m2, x, u0, v0, dt, tf = setup(wfx, orc, name)
    kx = np.zeros(mesh.ndofs)
    apply_k(mesh, P, G, x, kx)
    m = np.zeros(mesh.ndofs)
    apply_m(mesh, P, detJ, np.ones(mesh.ndofs), m)
    ua, va = np.zeros(mesh.ndofs), np.zeros(mesh.ndofs)
    run_rk4(mesh, P, G, m, m1, m2, C0, F0, P0, 0.0, tf, dt, ua, va)
    ub, vb = u0.copy(), v0.copy()
    run_rk4(mesh, P, G, m, m1, m2, C0, F0, P0, 0.0, tf, dt, ub, vb)
    return {"stiffness_of_x": kx, "lumped_mass": m, "rk4_from_rest_u": ua, "rk4_from_rest_v": va,
            "rk4_from_state_u": ub, "rk4_from_state_v": vb, "rk4_dt": np.float64(dt), "rk4_tf": np.float64(tf)}


def compute_reference(wfx, orc, name):
    return compute(wfx, orc, name, orc.reference_stiffness_apply, orc.reference_mass_apply, orc.reference_rk4)


def compute_oracle(wfx, orc, name):
    def rk4(mesh, P, G, m, m1, m2, c0, f0, p0, t0, tf, dt, u, v):
        orc.rk4(mesh, P, G, m, m1, m2, c0, f0, p0, t0, tf, dt, u, v, sumfact=False, nthreads=1)
    return compute(wfx, orc, name, lambda mesh, P, G, x, y: orc.stiffness_apply(mesh, P, G, x, y, dense=True),
                   orc.mass_apply, rk4)


if __name__ == "__main__":
    import wave_fenics_b200 as wfx
    from oracle import oracle as orc
    if orc.ref_cpu() is None:
        raise SystemExit("oracle/_ref/libwfref_cpu.so cannot be built: /root/reference is not present")
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        out = compute_reference(wfx, orc, name)
        np.savez_compressed(os.path.join(here, name + ".npz"), **out)
        print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
