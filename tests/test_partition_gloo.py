"""world_size > 1 host logic on CPU (gloo): the cartesian partition, its owner/ghost index
data in the VectorUpdater contract, and the reverse+forward ghost update, checked by
assembling the oracle operator rank by rank and comparing with the one-rank result."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

P, L = 3, 0.1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, gshape, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import wave_fenics_b200 as wfx
        from wave_fenics_b200 import partition
        from oracle import oracle
        grid = partition.rank_grid(world)
        mesh = partition.create_box_hex_partition(gshape, P, (L, L, L), grid, rank, perturb=0.15)
        G, detJ = oracle.precompute_geometric_data(mesh, P)
        # the same global vector on every rank, restricted to the local entries
        ndg = mesh.ndofs_global
        xg = np.random.default_rng(42).standard_normal(ndg)
        x = xg[mesh.global_dofs].copy()
        y = np.zeros(mesh.ndofs)
        oracle.stiffness_apply(mesh, P, G, x, y, dense=False)
        partition.exchange_rev_fwd_host(mesh, y)
        m = np.zeros(mesh.ndofs)
        oracle.mass_apply(mesh, P, detJ, np.ones(mesh.ndofs), m)
        partition.exchange_rev_fwd_host(mesh, m)
        m1, m2 = oracle.boundary_facet_mass(mesh, P)
        partition.exchange_rev_fwd_host(mesh, m1)
        q.put((rank, mesh.global_dofs, mesh.size_local, y, m, m1, mesh.ncells))
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.parametrize("world,gshape", [(2, (4, 3, 2)), (4, (4, 4, 3)), (6, (6, 4, 2))])
def test_partitioned_operator_equals_global(world, gshape, wfx, orc):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, gshape, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # one-rank reference
    mesh = wfx.create_box_hex(gshape, P, (L, L, L), perturb=0.15)
    G, detJ = orc.precompute_geometric_data(mesh, P)
    xg = np.random.default_rng(42).standard_normal(mesh.ndofs)
    yg = np.zeros(mesh.ndofs)
    orc.stiffness_apply(mesh, P, G, xg, yg, dense=False)
    mg = np.zeros(mesh.ndofs)
    orc.mass_apply(mesh, P, detJ, np.ones(mesh.ndofs), mg)
    m1g, _ = orc.boundary_facet_mass(mesh, P)
    owned_seen = np.zeros(mesh.ndofs, dtype=int)
    assert sum(r[6] for r in results) == mesh.ncells
    for rank, gd, size_local, y, m, m1, _ in results:
        owned_seen[gd[:size_local]] += 1
        # every local copy (owned and ghost) holds the assembled value
        assert np.linalg.norm(y - yg[gd]) < 1e-13 * np.linalg.norm(yg)
        np.testing.assert_allclose(m, mg[gd], rtol=1e-14)
        np.testing.assert_allclose(m1, m1g[gd], rtol=1e-13, atol=1e-20)
    assert (owned_seen == 1).all()  # every dof has exactly one owner


def test_rank_grid_matches_reference_decompose3d(wfx):
    from wave_fenics_b200 import partition
    assert [partition.rank_grid(w) for w in (1, 2, 4, 8, 16)] == [(1, 1, 1), (2, 1, 1), (2, 2, 1), (2, 2, 2), (4, 2, 2)]
    # not a power of two (the reference has no rule): most cubic factorisation
    assert [partition.rank_grid(w) for w in (3, 6, 12, 27)] == [(3, 1, 1), (3, 2, 1), (3, 2, 2), (3, 3, 3)]
