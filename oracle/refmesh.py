"""Structured hexahedral box mesh in DOLFINx layout, numpy only.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): the mesh generator the CPU
reference arm of bench.py and the oracle-side tests use, so that they never import the
product package (the product's own generator, wave-fenics_b200/mesh.py, takes its tensor
permutation from libwavefx.so; this one takes it from the oracle's `wo_perm`).  Same
conventions: vertex v = ix + 2 iy + 4 iz, cells c = (cx*n1 + cy)*n2 + cz,
cell-local dofs in DOLFINx order, lexicographic global dof numbering, perturbation drawn
from numpy.random.default_rng(seed) on the global vertex grid (interior vertices only).
Stands in for mesh::create_box (demo/gpu_operator/main.cpp:62-63).
"""
from types import SimpleNamespace

import numpy as np

from . import oracle


def lattice_pos(P):
    a = np.arange(P + 1)
    return np.where(a == 0, 0, np.where(a == 1, P, a - 1))


def box(n, P, lengths=(1.0, 1.0, 1.0), perturb=0.0, seed=1234):
    if isinstance(n, int):
        n = (n, n, n)
    n = tuple(int(v) for v in n)
    h = [lengths[a] / n[a] for a in range(3)]
    ix = [np.arange(n[a] + 1) for a in range(3)]
    X, Y, Z = np.meshgrid(*ix, indexing="ij")
    pts = np.stack([X * h[0], Y * h[1], Z * h[2]], axis=-1).astype(np.float64)
    if perturb > 0.0:
        d = np.random.default_rng(seed).uniform(-1.0, 1.0, size=pts.shape)
        for a in range(3):
            idx = [slice(None)] * 3
            for end in (0, n[a]):
                idx[a] = end
                d[tuple(idx)] = 0.0
                idx[a] = slice(None)
        pts = pts + perturb * np.array(h) * d
    x = pts.reshape(-1, 3)
    cx, cy, cz = [v.reshape(-1) for v in np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]),
                                                      indexing="ij")]
    v = np.arange(8)
    vx, vy, vz = v & 1, (v >> 1) & 1, (v >> 2) & 1
    xdofs = (((cx[:, None] + vx) * (n[1] + 1) + (cy[:, None] + vy)) * (n[2] + 1)
             + (cz[:, None] + vz)).astype(np.int32)
    np1 = P + 1
    perm = oracle.perm(P).astype(np.int64)
    pos = lattice_pos(P)
    ta, tb, tc = [v.reshape(-1) for v in np.meshgrid(np.arange(np1), np.arange(np1), np.arange(np1),
                                                      indexing="ij")]
    M = [P * n[a] + 1 for a in range(3)]
    ndofs = M[0] * M[1] * M[2]
    assert ndofs < 2 ** 31
    dofmap = np.empty((len(cx), np1 ** 3), dtype=np.int32)
    gid = (((cx[:, None] * P + pos[ta]) * M[1] + (cy[:, None] * P + pos[tb])) * M[2]
           + (cz[:, None] * P + pos[tc]))
    dofmap[:, perm] = gid.astype(np.int32)
    return SimpleNamespace(P=P, shape=n, x=x, xdofs=xdofs, dofmap=dofmap, ncells=len(cx), ndofs=ndofs,
                           size_local=ndofs, lengths=tuple(lengths))
