"""Cartesian partition of the structured hex mesh over the GPUs of one box, and the ghost-dof
halo description in the form the reference's VectorUpdater reads from the DOLFINx IndexMap
(demo/gpu_scatter_mpi/VectorUpdater.hpp:31-59; rank grid as decompose3d,
demo/gpu_cg/mesh.hpp:37-48).  GhostMode::none semantics: a rank holds only its own cells;
dofs on inter-rank faces/edges/vertices are owned by exactly one rank and are ghosts on
the others.  Local vector layout = DOLFINx la::Vector: [owned | ghosts].
"""
import ctypes as C

import numpy as np

from . import capi
from .mesh import HexMesh, _vertex_coords, cell_h, lattice_pos


def rank_grid(world):
    """2^x ranks -> (2^x0, 2^x1, 2^x2) with x = x0+x1+x2, as decompose3d (demo/gpu_cg/mesh.hpp:37-48,
    which only handles powers of two); other sizes: the factorisation a >= b >= c with the smallest
    a + b + c, i.e. the least interface area per rank."""
    if world < 1:
        raise ValueError("world size must be positive")
    if world & (world - 1):
        best = (world, 1, 1)
        for c in range(1, int(round(world ** (1 / 3))) + 2):
            if world % c:
                continue
            for b in range(c, int((world // c) ** 0.5) + 2):
                if (world // c) % b == 0 and (world // c) // b >= b:
                    cand = ((world // c) // b, b, c)
                    if sum(cand) < sum(best):
                        best = cand
        return best
    x = world.bit_length() - 1
    q, r = divmod(x, 3)
    return tuple(2 ** (q + (1 if i < r else 0)) for i in range(3))


def _block(n, parts, idx):
    """cells [lo, hi) of block idx when n cells are split into `parts` near-equal blocks"""
    base, rem = divmod(n, parts)
    lo = idx * base + min(idx, rem)
    return lo, lo + base + (1 if idx < rem else 0)


def rank_coords(grid, rank):
    return (rank // (grid[1] * grid[2]), (rank // grid[2]) % grid[1], rank % grid[2])


def _owner_block_1d(gl, bounds, P):
    """owner block index along one axis of global lattice coordinate(s) gl: block b owns
    lattice points [lo_b*P, hi_b*P), the last block also its upper end."""
    starts = np.array([b[0] * P for b in bounds])
    idx = np.searchsorted(starts, gl, side="right") - 1
    return np.minimum(idx, len(bounds) - 1)


def create_box_hex_partition(gshape, P, lengths, grid, rank, origin=(0.0, 0.0, 0.0), perturb=0.0,
                             seed=1234, tags=((0, 0, 1), (0, 1, 2))):
    """This rank's part of the gshape-cell global box mesh (see mesh.create_box_hex)."""
    gshape = tuple(int(v) for v in gshape)
    rc = rank_coords(grid, rank)
    bounds = [[_block(gshape[a], grid[a], i) for i in range(grid[a])] for a in range(3)]
    lo = [bounds[a][rc[a]][0] for a in range(3)]
    hi = [bounds[a][rc[a]][1] for a in range(3)]
    n = tuple(hi[a] - lo[a] for a in range(3))
    nc = n[0] * n[1] * n[2]
    np1, nd = P + 1, (P + 1) ** 3
    x = _vertex_coords(n, lengths, origin, perturb, seed, global_n=gshape, offset=tuple(lo))

    cx, cy, cz = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij")
    cx, cy, cz = cx.reshape(-1), cy.reshape(-1), cz.reshape(-1)
    v = np.arange(8)
    vx, vy, vz = v & 1, (v >> 1) & 1, (v >> 2) & 1
    xdofs = (((cx[:, None] + vx) * (n[1] + 1) + (cy[:, None] + vy)) * (n[2] + 1) + (cz[:, None] + vz)).astype(np.int32)

    # local lattice of dofs and its global ids / owners
    M = [P * n[a] + 1 for a in range(3)]
    GM = [P * gshape[a] + 1 for a in range(3)]
    gl = [np.arange(M[a]) + lo[a] * P for a in range(3)]               # global lattice coords per axis
    ob = [_owner_block_1d(gl[a], bounds[a], P) for a in range(3)]      # owner block per axis
    OX, OY, OZ = np.meshgrid(ob[0], ob[1], ob[2], indexing="ij")
    owner = ((OX * grid[1] + OY) * grid[2] + OZ).reshape(-1)
    GX, GY, GZ = np.meshgrid(gl[0], gl[1], gl[2], indexing="ij")
    gid = ((GX.astype(np.int64) * GM[1] + GY) * GM[2] + GZ).reshape(-1)
    nl = gid.size
    owned = np.nonzero(owner == rank)[0]
    ghost = np.nonzero(owner != rank)[0]
    ghost = ghost[np.lexsort((gid[ghost], owner[ghost]))]              # by owner, then global id
    order = np.concatenate([owned, ghost])                              # lattice index of local entry
    local_of_lattice = np.empty(nl, dtype=np.int64)
    local_of_lattice[order] = np.arange(nl)
    size_local = len(owned)

    perm = capi.compute_permutations(P).astype(np.int64)
    pos = lattice_pos(P)
    ta, tb, tc = np.meshgrid(np.arange(np1), np.arange(np1), np.arange(np1), indexing="ij")
    pa, pb, pc = pos[ta.reshape(-1)], pos[tb.reshape(-1)], pos[tc.reshape(-1)]
    # lattice index of point t of cell c = index of the cell's origin corner + a per-point offset
    base = ((cx * P) * M[1] + cy * P) * M[2] + cz * P
    off = np.empty(nd, dtype=np.int64)
    off[perm] = (pa * M[1] + pb) * M[2] + pc          # column perm[t] of the dofmap is tensor point t
    lut = local_of_lattice.astype(np.int32)
    dofmap = lut[base[:, None] + off[None, :]]

    # boundary facets: only where this block touches the global boundary
    fc, fl, ft = [], [], []
    low, high = {0: 2, 1: 1, 2: 0}, {0: 3, 1: 4, 2: 5}
    cc = [cx, cy, cz]
    for axis, side, tag in tags:
        if side == 0 and lo[axis] != 0:
            continue
        if side == 1 and hi[axis] != gshape[axis]:
            continue
        sel = np.nonzero(cc[axis] == (0 if side == 0 else n[axis] - 1))[0]
        fc.append(sel.astype(np.int32))
        fl.append(np.full(len(sel), low[axis] if side == 0 else high[axis], dtype=np.int32))
        ft.append(np.full(len(sel), tag, dtype=np.int32))
    cat = lambda L: np.concatenate(L) if L else np.zeros(0, dtype=np.int32)

    # halo lists.  forward direction = owner -> ghost.  recv: my ghosts grouped by owner (already
    # sorted by owner, global id).  send: for every other rank that ghosts my owned dofs, the
    # owned local indices sorted by global id -- computed locally from the geometry: rank s holds
    # lattice point g iff g lies in s's closed block.
    halo = {"recv_ranks": [], "recv_offsets": [0], "recv_indices": [], "send_ranks": [], "send_offsets": [0],
            "send_indices": []}
    gown = owner[ghost]
    for r in np.unique(gown):
        sel = np.nonzero(gown == r)[0]
        halo["recv_ranks"].append(int(r))
        halo["recv_indices"].append((size_local + sel).astype(np.int32))
        halo["recv_offsets"].append(halo["recv_offsets"][-1] + len(sel))
    gx, gy, gz = GX.reshape(-1)[owned], GY.reshape(-1)[owned], GZ.reshape(-1)[owned]
    for s in range(grid[0] * grid[1] * grid[2]):
        if s == rank:
            continue
        sc = rank_coords(grid, s)
        inside = np.ones(len(owned), dtype=bool)
        for a, g in enumerate((gx, gy, gz)):
            b0, b1 = bounds[a][sc[a]]
            inside &= (g >= b0 * P) & (g <= b1 * P)
        sel = np.nonzero(inside)[0]
        if len(sel) == 0:
            continue
        sel = sel[np.argsort(gid[owned][sel], kind="stable")]
        halo["send_ranks"].append(s)
        halo["send_indices"].append(sel.astype(np.int32))  # owned entries come first: local index = position
        halo["send_offsets"].append(halo["send_offsets"][-1] + len(sel))
    for k in ("recv_indices", "send_indices"):
        halo[k] = cat(halo[k])
    for k in ("recv_ranks", "recv_offsets", "send_ranks", "send_offsets"):
        halo[k] = np.asarray(halo[k], dtype=np.int32)

    return HexMesh(P=P, shape=n, x=x, xdofs=xdofs, dofmap=dofmap, ndofs=nl, size_local=size_local,
                   facet_cells=cat(fc), facet_local=cat(fl), facet_tags=cat(ft),
                   h_min=float(cell_h(x, xdofs).min()), lengths=tuple(lengths),
                   ndofs_global=GM[0] * GM[1] * GM[2], halo=halo, global_dofs=gid[order])


class Halo:
    """NCCL ghost update for vectors on this partition (wfx_halo_*).  `group` is the
    torch.distributed process group used once to hand out the NCCL unique id."""

    def __init__(self, mesh, ctx, dtype=np.float64, group=None, comm=None):
        """comm: the wfx_comm handle of another Halo on the same ranks (shared, not owned)."""
        import torch.distributed as dist
        self.ctx = ctx
        self.owns_comm = comm is None
        if comm is None:
            rank, world = dist.get_rank(group), dist.get_world_size(group)
            uid = [None]
            if rank == 0:
                buf = C.create_string_buffer(128)
                capi.call("wfx_comm_unique_id", buf)
                uid = [buf.raw]
            dist.broadcast_object_list(uid, src=0, group=group)
            comm = C.c_void_p()
            capi.call("wfx_comm_create", ctx.handle, uid[0], world, rank, C.byref(comm))
        self.comm = comm
        h = mesh.halo
        self.handle = C.c_void_p()
        capi.call("wfx_halo_create", ctx.handle, self.comm, capi.dtype_code(dtype),
                  int(mesh.size_local), int(mesh.ndofs - mesh.size_local),
                  len(h["send_ranks"]), capi.i32p(h["send_ranks"]), capi.i32p(h["send_offsets"]),
                  capi.i32p(h["send_indices"]), len(h["recv_ranks"]), capi.i32p(h["recv_ranks"]),
                  capi.i32p(h["recv_offsets"]), capi.i32p(h["recv_indices"]), C.byref(self.handle))

    @property
    def transport(self):
        """'p2p' (NVLink peer memory, one fused kernel) or 'nccl' for the fused ghost reduction"""
        t = C.c_int()
        capi.call("wfx_halo_transport", self.handle, C.byref(t))
        return "p2p" if t.value else "nccl"

    def _run(self, name, x):
        import torch
        capi.call(name, self.handle, C.c_void_p(x.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream))

    def update_fwd(self, x):
        self._run("wfx_halo_update_fwd", x)

    def update_rev(self, x):
        self._run("wfx_halo_update_rev", x)

    def update_rev_fwd(self, x):
        self._run("wfx_halo_update_rev_fwd", x)

    def update_rev_fwd_scaled(self, x, scale_ptr, stream=None):
        """ghost -> owner add, owner multiplies the sum by scale (1/m), owner -> ghost copy"""
        import torch
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        capi.call("wfx_halo_update_rev_fwd_scaled", self.handle, C.c_void_p(x.data_ptr()),
                  C.c_void_p(scale_ptr), C.c_void_p(st))

    def __del__(self):
        if getattr(self, "handle", None) and getattr(capi, "lib", None) is not None:
            capi.lib.wfx_halo_destroy(self.handle)
            if self.owns_comm:
                capi.lib.wfx_comm_destroy(self.comm)
            self.handle = None


def make_halo(mesh, ctx, dtype=np.float64):
    return Halo(mesh, ctx, dtype)


def exchange_rev_fwd_host(mesh, vec, group=None):
    """Host emulation of update_rev followed by update_fwd with torch.distributed point-to-point
    messages (any backend; used by the gloo CPU tests): ghost -> owner add in neighbour order, then
    owner -> ghost copy."""
    import torch
    import torch.distributed as dist
    h = mesh.halo
    # reverse: send my ghost values to their owners, receive contributions for my owned dofs
    reqs, rbufs = [], []
    for i, r in enumerate(h["send_ranks"]):
        buf = torch.empty(int(h["send_offsets"][i + 1] - h["send_offsets"][i]), dtype=torch.float64)
        rbufs.append(buf)
        reqs.append(dist.irecv(buf, src=int(r), group=group))
    for i, r in enumerate(h["recv_ranks"]):
        idx = h["recv_indices"][h["recv_offsets"][i]:h["recv_offsets"][i + 1]]
        reqs.append(dist.isend(torch.from_numpy(vec[idx].copy()), dst=int(r), group=group))
    for q in reqs:
        q.wait()
    for i, r in enumerate(h["send_ranks"]):
        idx = h["send_indices"][h["send_offsets"][i]:h["send_offsets"][i + 1]]
        vec[idx] += rbufs[i].numpy()  # indices are unique within one neighbour's list
    # forward
    reqs, rbufs = [], []
    for i, r in enumerate(h["recv_ranks"]):
        buf = torch.empty(int(h["recv_offsets"][i + 1] - h["recv_offsets"][i]), dtype=torch.float64)
        rbufs.append(buf)
        reqs.append(dist.irecv(buf, src=int(r), group=group))
    for i, r in enumerate(h["send_ranks"]):
        idx = h["send_indices"][h["send_offsets"][i]:h["send_offsets"][i + 1]]
        reqs.append(dist.isend(torch.from_numpy(vec[idx].copy()), dst=int(r), group=group))
    for q in reqs:
        q.wait()
    for i, r in enumerate(h["recv_ranks"]):
        idx = h["recv_indices"][h["recv_offsets"][i]:h["recv_offsets"][i + 1]]
        vec[idx] = rbufs[i].numpy()
    return vec
