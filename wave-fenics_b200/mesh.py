"""Synthetic structured hexahedral meshes in DOLFINx layout.

Stands in for `mesh::create_box` / XDMF ingest of the reference drivers
(demo/gpu_operator/main.cpp:62-63, demo/cpu_planar3d/main.cpp:39-45) and for the cartesian
partitioner of demo/gpu_cg/mesh.hpp:37-48,252-328: it produces exactly the arrays the
operators consume -- geometry.x(), geometry.dofmap(), V->dofmap()->list() for a degree-P
GLL Lagrange space, and (cell, local facet, tag) boundary triplets.

Conventions: hexahedron vertex v = ix + 2 iy + 4 iz; cell-local dofs in
DOLFINx order (vertices, edges, faces, interior); `perm` maps tensor index
t = ix*n^2 + iy*n + iz (1-D order [0, 1, interior]) to the local dof.
"""
from dataclasses import dataclass, field

import numpy as np

from . import capi


@dataclass
class HexMesh:
    P: int
    shape: tuple            # local cells per axis
    x: np.ndarray           # [npts, 3] float64
    xdofs: np.ndarray       # [ncells, 8] int32
    dofmap: np.ndarray      # [ncells, (P+1)^3] int32, DOLFINx local order
    ndofs: int              # local vector length: owned + ghosts
    size_local: int         # owned entries come first
    facet_cells: np.ndarray
    facet_local: np.ndarray
    facet_tags: np.ndarray
    h_min: float            # mesh::h minimum (largest vertex distance of the smallest cell)
    lengths: tuple
    ndofs_global: int = 0
    # halo description (empty on one rank); see partition.py
    halo: dict = field(default_factory=dict)
    # global dof id of every local entry (for checks / gathers)
    global_dofs: np.ndarray = None

    @property
    def ncells(self):
        return self.xdofs.shape[0]

    @property
    def nd(self):
        return (self.P + 1) ** 3


def lattice_pos(P):
    """ascending lattice position of the 1-D dofs in [0, 1, interior] order"""
    a = np.arange(P + 1)
    return np.where(a == 0, 0, np.where(a == 1, P, a - 1))


def _vertex_coords(n, lengths, origin, perturb, seed, global_n=None, offset=(0, 0, 0)):
    """Vertex coordinates of an (n0,n1,n2)-cell block that is part of a global_n grid.
    The perturbation is a function of the GLOBAL vertex index so that partitions agree."""
    gn = n if global_n is None else global_n
    h = [lengths[a] / gn[a] for a in range(3)]
    ix = [np.arange(n[a] + 1) + offset[a] for a in range(3)]
    X, Y, Z = np.meshgrid(ix[0], ix[1], ix[2], indexing="ij")
    pts = np.stack([origin[0] + X * h[0], origin[1] + Y * h[1], origin[2] + Z * h[2]], axis=-1)
    pts = pts.astype(np.float64)
    if perturb > 0.0:
        # full global perturbation field, deterministic in the seed; interior vertices only
        rng = np.random.default_rng(seed)
        d = rng.uniform(-1.0, 1.0, size=(gn[0] + 1, gn[1] + 1, gn[2] + 1, 3))
        for a in range(3):
            idx = [slice(None)] * 3
            for end in (0, gn[a]):
                idx[a] = end
                d[tuple(idx) + (slice(None),)] = 0.0
                idx[a] = slice(None)
        sub = d[ix[0][0]:ix[0][-1] + 1, ix[1][0]:ix[1][-1] + 1, ix[2][0]:ix[2][-1] + 1]
        pts = pts + perturb * np.array(h) * sub
    return pts.reshape(-1, 3)


def cell_h(x, xdofs):
    """dolfinx mesh::h for hexahedra: largest vertex-vertex distance."""
    v = x[xdofs]  # [nc, 8, 3]
    best = np.zeros(len(xdofs))
    for a in range(8):
        for b in range(a + 1, 8):
            d = v[:, a, :] - v[:, b, :]
            np.maximum(best, np.einsum("ij,ij->i", d, d), out=best)
    return np.sqrt(best)


def create_box_hex(n, P, lengths=(1.0, 1.0, 1.0), origin=(0.0, 0.0, 0.0), perturb=0.0, seed=1234,
                   renumber=None, tags=((0, 0, 1), (0, 1, 2))):
    """Structured n[0] x n[1] x n[2] hexahedral mesh of a box with a degree-P GLL space.

    perturb : interior vertices are moved by perturb*h*U(-1,1) (non-affine cells, full 3x3 G)
    renumber: None = lexicographic global dof numbering; an int seeds a random permutation
    tags    : (axis, side, tag) triplets naming tagged boundary faces; default tag 1 on x=0
              (source) and tag 2 on x=L (absorbing), as assumed for cpu_planar3d.
    """
    if isinstance(n, int):
        n = (n, n, n)
    n = tuple(int(v) for v in n)
    nc = n[0] * n[1] * n[2]
    np1 = P + 1
    nd = np1 ** 3
    x = _vertex_coords(n, lengths, origin, perturb, seed)

    cx, cy, cz = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij")
    cx, cy, cz = cx.reshape(-1), cy.reshape(-1), cz.reshape(-1)  # c = (cx*n1 + cy)*n2 + cz
    v = np.arange(8)
    vx, vy, vz = v & 1, (v >> 1) & 1, (v >> 2) & 1
    xdofs = (((cx[:, None] + vx) * (n[1] + 1) + (cy[:, None] + vy)) * (n[2] + 1)
             + (cz[:, None] + vz)).astype(np.int32)

    perm = capi.compute_permutations(P).astype(np.int64)
    pos = lattice_pos(P)
    ta, tb, tc = np.meshgrid(np.arange(np1), np.arange(np1), np.arange(np1), indexing="ij")
    pa, pb, pc = pos[ta.reshape(-1)], pos[tb.reshape(-1)], pos[tc.reshape(-1)]
    M = [P * n[a] + 1 for a in range(3)]
    ndofs = M[0] * M[1] * M[2]
    if ndofs >= 2 ** 31:
        raise ValueError("more than 2^31 dofs on one rank")
    # global lattice id of point t of cell c = id of the cell's origin corner + a per-point offset
    base = ((cx * P) * M[1] + cy * P) * M[2] + cz * P
    off = np.empty(nd, dtype=np.int64)
    off[perm] = (pa * M[1] + pb) * M[2] + pc          # column perm[t] of the dofmap is tensor point t
    global_dofs = np.arange(ndofs, dtype=np.int64)
    if renumber is not None:
        new_of_old = np.random.default_rng(renumber).permutation(ndofs)
        dofmap = new_of_old[base[:, None] + off[None, :]].astype(np.int32)
    else:
        dofmap = base.astype(np.int32)[:, None] + off.astype(np.int32)[None, :]

    fc, fl, ft = [], [], []
    # local facet on the low / high side of each axis
    low, high = {0: 2, 1: 1, 2: 0}, {0: 3, 1: 4, 2: 5}
    cc = [cx, cy, cz]
    for axis, side, tag in tags:
        sel = np.nonzero(cc[axis] == (0 if side == 0 else n[axis] - 1))[0]
        fc.append(sel.astype(np.int32))
        fl.append(np.full(len(sel), low[axis] if side == 0 else high[axis], dtype=np.int32))
        ft.append(np.full(len(sel), tag, dtype=np.int32))
    cat = lambda L: np.concatenate(L) if L else np.zeros(0, dtype=np.int32)
    return HexMesh(P=P, shape=n, x=x, xdofs=xdofs, dofmap=dofmap, ndofs=ndofs, size_local=ndofs,
                   facet_cells=cat(fc), facet_local=cat(fl), facet_tags=cat(ft),
                   h_min=float(cell_h(x, xdofs).min()), lengths=tuple(lengths), ndofs_global=ndofs,
                   global_dofs=global_dofs)


def dof_coordinates(mesh):
    """Physical coordinates of every local dof (trilinear map of the GLL nodes)."""
    P, np1 = mesh.P, mesh.P + 1
    pts, _ = capi.gll(P)
    perm = capi.compute_permutations(P)
    ta, tb, tc = np.meshgrid(np.arange(np1), np.arange(np1), np.arange(np1), indexing="ij")
    X = np.stack([pts[ta.reshape(-1)], pts[tb.reshape(-1)], pts[tc.reshape(-1)]], axis=-1)  # [nd,3]
    v = np.arange(8)
    bits = np.stack([v & 1, (v >> 1) & 1, (v >> 2) & 1], axis=-1)  # [8,3]
    # phi_v(X) = prod_a (X_a if bit else 1-X_a)
    phi = np.prod(np.where(bits[None, :, :] == 1, X[:, None, :], 1.0 - X[:, None, :]), axis=-1)
    out = np.zeros((mesh.ndofs, 3))
    xc = mesh.x[mesh.xdofs]  # [nc,8,3]
    coords = np.einsum("tv,cva->cta", phi, xc)  # tensor order
    out[mesh.dofmap[:, perm].reshape(-1)] = coords.reshape(-1, 3)
    return out


def cfl_timestep(h_min, c0, P, f0, cfl=0.5):
    """Time step of demo/cpu_planar3d/main.cpp:61-66 (snapped to whole steps per period)."""
    dt = cfl * h_min / (c0 * P ** 2)
    period = 1.0 / f0
    steps_per_period = int(period / dt + 1)
    return period / steps_per_period
