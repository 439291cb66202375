#!/usr/bin/env python
"""The cpu_planar3d demo (demo/cpu_planar3d/main.cpp:14-98) on the B200 path.

Same physical set-up and control flow as the reference driver: speed of sound 1500 m/s, source
0.5 MHz, amplitude 60 kPa, domain length 0.1 m, degree 4, CFL time step snapped to whole steps
per period, final time L/c0 + 8/f0, LinearGLLOpt.init() then .rk4(t0, tf, dt), "Solve time".
The reference reads `../mesh.xdmf` (not in its repository); here the mesh is a structured
n x m x m hexahedral box with tag 1 on x = 0 (source) and tag 2 on x = L (absorbing).

    python demo/planar3d.py [--nx 32] [--nyz 4] [--steps N] [--check]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wave_fenics_b200 as wfx  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=32, help="cells along the propagation direction")
    ap.add_argument("--nyz", type=int, default=4, help="cells across")
    ap.add_argument("--steps", type=int, default=0, help="stop after this many steps (0: run to the final time)")
    ap.add_argument("--check", action="store_true", help="compare with the CPU oracle (small meshes)")
    args = ap.parse_args()

    speedOfSound, sourceFrequency, pressureAmplitude = 1500.0, 0.5e6, 60000.0   # main.cpp:25-30
    period = 1.0 / sourceFrequency
    domainLength = 0.1                                                           # :33
    degreeOfBasis = 4                                                            # :36
    side = domainLength / args.nx
    mesh = wfx.create_box_hex((args.nx, args.nyz, args.nyz), degreeOfBasis,
                              (domainLength, side * args.nyz, side * args.nyz))
    timeStepSize = wfx.cfl_timestep(mesh.h_min, speedOfSound, degreeOfBasis, sourceFrequency)  # :61-66
    startTime, finalTime = 0.0, domainLength / speedOfSound + 8.0 / sourceFrequency            # :63-64
    print(f"Number of step per period: {int(round(period / timeStepSize))}")
    print(f"dt = {timeStepSize:.15g}")
    nstep = int((finalTime - startTime) / timeStepSize + 1)
    eqn = wfx.LinearGLLOpt(mesh, None, degreeOfBasis, speedOfSound, sourceFrequency, pressureAmplitude)  # :75
    print(f"Number of steps: {nstep}")
    print(f"Degrees of freedom: {mesh.ndofs_global}")
    eqn.init()                                                                   # :83
    eqn.ctx.synchronize()
    t0 = time.perf_counter()
    steps, t_end = eqn.rk4(startTime, finalTime, timeStepSize, max_steps=args.steps)  # :88
    eqn.ctx.synchronize()
    print(f"Solve time: {time.perf_counter() - t0:.6f}  ({steps} steps, t = {t_end:.6e})")   # :92
    u, v = eqn.get_state()
    X = wfx.dof_coordinates(mesh)
    print(f"max |p| = {np.abs(u).max():.6e} at x = {X[np.abs(u).argmax(), 0]:.4f}")
    if args.check:
        from oracle import oracle
        G, detJ = oracle.precompute_geometric_data(mesh, degreeOfBasis)
        m = np.zeros(mesh.ndofs)
        oracle.mass_apply(mesh, degreeOfBasis, detJ, np.ones(mesh.ndofs), m)
        m1, m2 = oracle.boundary_facet_mass(mesh, degreeOfBasis)
        uo, vo = np.zeros(mesh.ndofs), np.zeros(mesh.ndofs)
        oracle.rk4(mesh, degreeOfBasis, G, m, m1, m2, speedOfSound, sourceFrequency, pressureAmplitude,
                   startTime, finalTime, timeStepSize, uo, vo, max_steps=args.steps, sumfact=True,
                   nthreads=oracle.max_threads())
        print(f"relative L2 difference to the CPU oracle: u {np.linalg.norm(u - uo) / np.linalg.norm(uo):.3e}"
              f"  v {np.linalg.norm(v - vo) / np.linalg.norm(vo):.3e}")


if __name__ == "__main__":
    main()
