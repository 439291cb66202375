"""Generates tests/golden/*.npz.

IMPORTANT: the reference (Excalibur-SLE/wave-fenics) ships no golden vectors for this path and cannot
be built or imported as a whole in this environment (what of it does compile produces
tests/golden/reference_*.npz, see make_reference_golden.py), so these fixtures are outputs of THIS
repository's CPU oracle (oracle/wave_oracle.c, strict build: -O2 -ffp-contract=off, one thread), not of
the reference.  They (1) pin the oracle against drift -- tests/test_golden.py re-runs it and compares
-- and (2) give the GPU box stored values to check the CUDA path against without running the oracle.
Inputs are not stored: the meshes and vectors are regenerated from the seeds below.

    python tests/golden/make_golden.py        # rewrites the fixtures
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

L, C0, F0, P0 = 0.1, 1500.0, 0.5e6, 6e4
# name -> (degree, cells per axis, vertex perturbation, dof renumbering seed or None, RK4 steps)
CASES = {
    "p2_n3_perturbed": (2, (3, 3, 3), 0.15, None, 12),
    "p4_n2_perturbed_renumbered": (4, (2, 2, 2), 0.15, 5, 8),
    "p4_ragged_affine": (4, (3, 2, 1), 0.0, None, 8),
    "p5_n2_perturbed": (5, (2, 2, 2), 0.15, None, 0),
}


def make_mesh(wfx, name):
    P, shape, perturb, renumber, _ = CASES[name]
    lengths = tuple(L * s / 8 for s in shape)
    return P, wfx.create_box_hex(shape, P, lengths, perturb=perturb, renumber=renumber)


def inputs(ndofs):
    rng = np.random.default_rng(42)
    return rng.standard_normal(ndofs)


def compute(wfx, orc, name):
    """All golden quantities of one case from the oracle (one thread: bitwise repeatable)."""
    P, mesh = make_mesh(wfx, name)
    steps = CASES[name][4]
    G, detJ = orc.precompute_geometric_data(mesh, P)
    m = np.zeros(mesh.ndofs)
    orc.mass_apply(mesh, P, detJ, np.ones(mesh.ndofs), m)
    x = inputs(mesh.ndofs)
    kx = np.zeros(mesh.ndofs)
    orc.stiffness_apply(mesh, P, G, x, kx, dense=True, nthreads=1)
    m1, m2 = orc.boundary_facet_mass(mesh, P)
    out = {"perm": orc.perm(P).astype(np.int32), "gll_points": orc.gll(P)[0], "gll_weights": orc.gll(P)[1],
           "deriv_1d": orc.deriv_1d(P, clamp=True), "detJ": detJ, "G_upper": G[:, :, *np.triu_indices(3)],
           "lumped_mass": m, "stiffness_of_x": kx, "facet_mass_1": m1, "facet_mass_2": m2}
    if steps:
        dt = wfx.cfl_timestep(mesh.h_min, C0, P, F0)
        tf = L / C0 + 8.0 / F0
        u, v = np.zeros(mesh.ndofs), np.zeros(mesh.ndofs)
        s, t_end = orc.rk4(mesh, P, G, m, m1, m2, C0, F0, P0, 0.0, tf, dt, u, v, max_steps=steps,
                           sumfact=False, nthreads=1)
        out.update({"rk4_u": u, "rk4_v": v, "rk4_steps": np.int64(s), "rk4_t_end": np.float64(t_end),
                    "rk4_dt": np.float64(dt)})
    return out


def main():
    import wave_fenics_b200 as wfx
    from oracle import oracle as orc
    here = os.path.dirname(os.path.abspath(__file__))
    for name in CASES:
        data = compute(wfx, orc, name)
        path = os.path.join(here, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: {os.path.getsize(path) / 1024:.1f} KB, {len(data)} arrays")


if __name__ == "__main__":
    main()
