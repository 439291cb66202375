"""CPU-side checks of the product: host tables against the oracle, the C-ABI library loads
and exports every symbol include/wavefx.h declares, mesh generator invariants, and the
atomic-free scatter plans (built and verified on the host, no GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(wfx):
    hdr = open(os.path.join(ROOT, "include", "wavefx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(wfx_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 40
    lib = ctypes.CDLL(wfx.capi.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.wfx_version() == 100
    # every declared symbol is bound in the Python layer too
    bound = set(wfx.capi._SIG) | {"wfx_last_error", "wfx_version"}
    assert not [n for n in names if n not in bound]


def test_no_gpu_fails_loudly(wfx):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(wfx.WfxError, match="no CPU fallback|no CUDA device"):
        wfx.Context(0)


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 6, 7])
def test_tables_match_oracle(wfx, orc, P):
    assert (wfx.capi.compute_permutations(P) == orc.perm(P)).all()
    p, w = wfx.capi.gll(P)
    po, wo = orc.gll(P)
    np.testing.assert_allclose(p, po, rtol=0, atol=2e-16)
    np.testing.assert_allclose(w, wo, rtol=4e-16, atol=0)
    D, Do = wfx.capi.deriv_1d(P), orc.deriv_1d(P)
    np.testing.assert_allclose(D, Do, rtol=1e-14, atol=1e-14)
    assert ((D == 0) == (Do == 0)).all()


def test_reorder_dofmap_contract(wfx, orc):
    # common/permute.hpp:22-26: out[c*nd+t] = in[c*nd+perm[t]]
    P = 3
    mesh = wfx.create_box_hex(2, P, renumber=3)
    out = wfx.capi.reorder_dofmap(mesh.dofmap, P)
    perm = wfx.capi.compute_permutations(P)
    assert (out == mesh.dofmap[:, perm]).all()
    assert (out == orc.reorder_dofmap(mesh.dofmap, P)).all()


def test_tabulate_1d_matches_oracle(wfx, orc):
    for P, q in [(2, 4), (4, 8), (5, 10)]:
        for d in (0, 1):
            np.testing.assert_allclose(wfx.capi.tabulate_1d(P, q, d), orc.tabulate_1d(P, q, d),
                                       rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("P", [2, 4])
def test_mesh_generator(wfx, P):
    N = 3
    mesh = wfx.create_box_hex(N, P, (0.1, 0.1, 0.1))
    assert mesh.ndofs == (P * N + 1) ** 3 and mesh.dofmap.shape == (N ** 3, (P + 1) ** 3)
    # every dof referenced, shared dofs coincide geometrically
    assert np.array_equal(np.unique(mesh.dofmap), np.arange(mesh.ndofs))
    X = wfx.dof_coordinates(mesh)
    pts, _ = wfx.capi.gll(P)
    perm = wfx.capi.compute_permutations(P)
    n = P + 1
    h = 0.1 / N
    for c in (0, 13, 26):
        cz, cy, cx = c % N, (c // N) % N, c // (N * N)
        for t in (0, 7, n ** 3 - 1, n ** 3 // 2):
            i, j, k = t // (n * n), (t // n) % n, t % n
            want = np.array([(cx + pts[i]) * h, (cy + pts[j]) * h, (cz + pts[k]) * h])
            np.testing.assert_allclose(X[mesh.dofmap[c, perm[t]]], want, atol=1e-15)
    # vertex dofs 0..7 sit on the geometry vertices in the same order
    np.testing.assert_allclose(X[mesh.dofmap[:, :8]], mesh.x[mesh.xdofs], atol=1e-15)
    # tags: 1 on x=0 (local facet 2), 2 on x=L (local facet 3)
    assert set(mesh.facet_local[mesh.facet_tags == 1]) == {2}
    assert set(mesh.facet_local[mesh.facet_tags == 2]) == {3}
    assert (mesh.facet_tags == 1).sum() == N * N


def _centroids(mesh):
    return mesh.x[mesh.xdofs].mean(axis=1)


@pytest.mark.parametrize("P,N,be,W", [(4, 8, 4, 8), (4, 6, 4, 8), (2, 8, 8, 16), (3, 5, 4, 8),
                                      (5, 4, 2, 1), (7, 2, 2, 1)])
def test_brick_plan_structured(wfx, P, N, be, W):
    mesh = wfx.create_box_hex(N, P, perturb=0.15)
    s = wfx.capi.debug_plan_stats(P, mesh.dofmap, mesh.ndofs, _centroids(mesh), be, W)
    nb_axis = -(-N // be)
    assert s["batches"] == nb_axis ** 3
    assert s["batch_colours"] == min(8, s["batches"])
    assert s["cell_colours"] == 8
    # regular bricks are placed on a padded lattice (bank-conflict-free strides): a few holes
    full = (P * min(be, N) + 1) ** 3
    assert full <= s["nloc_max"] <= 1.07 * (P * be + 1) ** 3
    assert s["regular_batches"] == s["batches"]
    assert s["untouched"] == 0
    if N % be == 0 and be ** 3 // 8 >= W:
        assert s["padded_slots"] == 0


def test_brick_plan_random_numbering_and_no_geometry(wfx):
    P, N = 4, 4
    mesh = wfx.create_box_hex(N, P, renumber=11)
    rng = np.random.default_rng(5)
    shuffled = mesh.dofmap[rng.permutation(mesh.ncells)]
    # without centroids the plan batches cells in the given (here random) order: still valid
    s = wfx.capi.debug_plan_stats(P, shuffled, mesh.ndofs, None, 4, 8)
    assert s["batches"] >= 1 and s["untouched"] == 0 and s["regular_batches"] == 0
    # a tight shared-memory capacity forces batch splitting
    s2 = wfx.capi.debug_plan_stats(P, mesh.dofmap, mesh.ndofs, _centroids(mesh), 4, 8, nloc_cap=2000)
    assert s2["nloc_max"] <= 2000 and s2["batches"] > 1
    # vector entries no cell references (ghost-only slots) are reported
    s3 = wfx.capi.debug_plan_stats(P, mesh.dofmap, mesh.ndofs + 5, _centroids(mesh), 4, 8)
    assert s3["untouched"] == 5


def test_plan_rejects_bad_dofmap(wfx):
    mesh = wfx.create_box_hex(2, 2)
    bad = mesh.dofmap.copy()
    bad[0, 0] = mesh.ndofs
    with pytest.raises(wfx.WfxError, match="out of range"):
        wfx.capi.debug_plan_stats(2, bad, mesh.ndofs, None, 4, 8)


def test_cfl_timestep_matches_reference_formula(wfx):
    # demo/cpu_planar3d/main.cpp:61-66 with mesh::h = sqrt(3) * side (App. A.7)
    side = 0.1 / 16
    mesh = wfx.create_box_hex(2, 4, (2 * side,) * 3)
    assert abs(mesh.h_min - np.sqrt(3) * side) < 1e-15
    dt = wfx.cfl_timestep(mesh.h_min, 1500.0, 4, 0.5e6)
    raw = 0.5 * np.sqrt(3) * side / (1500.0 * 16)
    spp = int(2e-6 / raw + 1)
    assert dt == 2e-6 / spp and dt <= raw


@pytest.mark.parametrize("P", [1, 2, 4, 6])
def test_tabulate_basis_and_permutation_matches_oracle(wfx, orc, P):
    # a2: the dense table of common/operators.hpp:13-32 (exported for callers of that function)
    table, perm = wfx.capi.tabulate_basis_and_permutation(P)
    nd = (P + 1) ** 3
    assert np.array_equal(perm, orc.perm(P))
    np.testing.assert_allclose(table[1:], orc.tabulate_dphi(P).reshape(3, nd, nd), rtol=1e-13, atol=1e-13)
    # collocation: phi_i(x_q) = delta in DOLFINx dof order
    ident = np.zeros((nd, nd))
    ident[np.arange(nd), perm] = 1.0
    assert np.array_equal(table[0], ident)


@pytest.mark.parametrize("P,shape,W,mesh_n", [(5, (4, 2, 2), 2, 4), (4, (4, 4, 2), 4, 8), (6, (2, 2, 4), 2, 4), (2, (8, 4, 4), 16, 8)])
def test_brick_plan_non_cubic_bricks(wfx, P, shape, W, mesh_n):
    # bricks need not be cubes: rounds of W conflict-free cells exist as soon as one axis has >= 2W/4 cells
    mesh = wfx.create_box_hex(mesh_n, P, perturb=0.15)
    code = shape[0] | (shape[1] << 8) | (shape[2] << 16)
    s = wfx.capi.debug_plan_stats(P, mesh.dofmap, mesh.ndofs, _centroids(mesh), code, W)
    nb = 1
    for a in range(3):
        nb *= -(-mesh_n // shape[a])
    assert s["batches"] == nb
    assert s["regular_batches"] == s["batches"]
    dense = (P * shape[0] + 1) * (P * shape[1] + 1) * (P * shape[2] + 1)
    assert dense <= s["nloc_max"] <= 1.05 * dense
    # every round is full: cells per brick / 8 parity classes >= W in these cases
    assert s["padded_slots"] == 0


def test_brick_plan_invariants_on_random_connectivity(wfx):
    """Property test (hypothesis): for ARBITRARY cell -> dof connectivity -- nothing like a mesh --
    the planner must still produce plans that satisfy every invariant the kernels rely on
    (wfx_debug_plan_stats runs verify_brick_plan / verify_cell_colour_plan and fails otherwise)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.sampled_from([1, 2]), st.integers(1, 14), st.sampled_from([1, 2, 4, 8]),
           st.sampled_from([1, 2, 3, 4]), st.booleans(), st.integers(0, 3))
    def run(seed, P, ncells, W, be, with_centroids, extra):
        rng = np.random.default_rng(seed)
        nd = (P + 1) ** 3
        ndofs = int(rng.integers(nd, nd * ncells + 1)) + extra
        dofmap = np.stack([rng.choice(ndofs - extra, size=nd, replace=False) for _ in range(ncells)]).astype(np.int32)
        cen = rng.uniform(0, 1, size=(ncells, 3)).astype(np.float32) if with_centroids else None
        cap = int(rng.integers(nd, 4 * nd + 1))
        s = wfx.capi.debug_plan_stats(P, dofmap, ndofs, cen, be, W, nloc_cap=cap)
        assert s["nloc_max"] <= max(cap, nd)
        assert s["batches"] >= 1 and s["rounds"] >= s["batches"]
        assert s["untouched"] == ndofs - len(np.unique(dofmap))

    run()


@pytest.mark.parametrize("P,shape,brick,W", [(2, (5, 4, 3), (8, 8, 8), 16), (3, (3, 5, 2), (8, 8, 4), 16),
                                             (4, (5, 6, 9), (4, 4, 4), 8), (5, (3, 2, 4), (4, 4, 2), 4)])
def test_stream_plan_invariants_on_meshes(wfx, P, shape, brick, W):
    """The streamed-cell kernel's plan (FIRST / LAST flags in the per-point dofmap, relabelled axes, both cell
    orders) on ragged structured meshes: wfx_debug_stream_plan_check replays the execution order and fails on
    any violated invariant.  Lexicographic numbering puts the contiguous (z) axis on the fast lane index;
    renumbered dofs keep the mesh's axes."""
    capi = wfx.capi
    mesh = wfx.create_box_hex(shape, P, (1.0, 0.7, 1.3), perturb=0.1)
    cen = mesh.x[mesh.xdofs].mean(axis=1)
    for order in (True, False):
        s = capi.debug_stream_plan_check(P, mesh.dofmap, mesh.ndofs, cen, order, brick, W)
        assert s["axis_perm"] == [1, 2, 0] and s["untouched"] == 0
        assert s["colours"] == 8 or order  # cell colours: the 8 parity classes; batch colours: up to 8
        s = capi.debug_stream_plan_check(P, mesh.dofmap, mesh.ndofs + 3, cen, order, brick, W, relabel_axes=False)
        assert s["axis_perm"] == [0, 1, 2] and s["untouched"] == 3
    ren = wfx.create_box_hex(shape, P, (1.0, 0.7, 1.3), perturb=0.1, renumber=5)
    assert capi.debug_stream_plan_check(P, ren.dofmap, ren.ndofs, cen, True, brick, W)["axis_perm"] == [0, 1, 2]
    # rank-shared dofs (a lattice plane): interface batches first, never LAST; brick order only
    M = [P * n + 1 for n in shape]
    plane = (np.arange(M[1] * M[2]) + (M[0] // 2) * M[1] * M[2]).astype(np.int32)
    s = capi.debug_stream_plan_check(P, mesh.dofmap, mesh.ndofs, cen, True, brick, W, shared=plane)
    assert s["part_split"] > 0
    with pytest.raises(wfx.WfxError, match="partitioned"):
        capi.debug_stream_plan_check(P, mesh.dofmap, mesh.ndofs, cen, False, brick, W, shared=plane)


def test_stream_plan_invariants_on_random_connectivity(wfx):
    """Property test: arbitrary cell -> dof connectivity, both cell orders, with and without shared dofs."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.integers(0, 2 ** 31 - 1), st.sampled_from([2, 3]), st.integers(1, 12), st.sampled_from([1, 2, 4, 8]),
           st.sampled_from([1, 2, 4]), st.booleans(), st.booleans(), st.integers(0, 3))
    def run(seed, P, ncells, W, be, brick_order, with_shared, extra):
        rng = np.random.default_rng(seed)
        nd = (P + 1) ** 3
        ndofs = int(rng.integers(nd, nd * ncells + 1)) + extra
        dofmap = np.stack([rng.choice(ndofs - extra, size=nd, replace=False) for _ in range(ncells)]).astype(np.int32)
        cen = rng.uniform(0, 1, size=(ncells, 3)).astype(np.float32)
        shared = rng.choice(ndofs, size=max(1, ndofs // 7), replace=False) if (with_shared and brick_order) else None
        s = wfx.capi.debug_stream_plan_check(P, dofmap, ndofs, cen, brick_order, (be, be, be), W, shared=shared)
        assert s["untouched"] == ndofs - len(np.unique(dofmap))

    run()


def test_structured_coordinates_from_connectivity(wfx):
    """The planner's integer cell coordinates come from the mesh connectivity, not the geometry:
    exact for any cell order, independent of shear / grading; rejected for meshes that are not one
    consistently oriented block."""
    capi = wfx.capi
    mesh = wfx.create_box_hex((5, 3, 4), 2, (1.0, 0.2, 3.0), perturb=0.2)
    ok, ijk = capi.debug_structured_coords(mesh.xdofs, mesh.x.shape[0])
    cx, cy, cz = np.meshgrid(np.arange(5), np.arange(3), np.arange(4), indexing="ij")
    want = np.stack([cx.reshape(-1), cy.reshape(-1), cz.reshape(-1)], axis=-1)
    assert ok and np.array_equal(ijk, want)
    # any cell order
    perm = np.random.default_rng(0).permutation(mesh.ncells)
    ok, ijk = capi.debug_structured_coords(mesh.xdofs[perm], mesh.x.shape[0])
    assert ok and np.array_equal(ijk, want[perm])
    # a cell with mirrored local vertex order: not a consistently oriented block
    bad = mesh.xdofs.copy()
    bad[7] = bad[7][[1, 0, 3, 2, 5, 4, 7, 6]]
    ok, _ = capi.debug_structured_coords(bad, mesh.x.shape[0])
    assert not ok
    # two disconnected blocks
    two = np.concatenate([mesh.xdofs, mesh.xdofs + mesh.x.shape[0]])
    ok, _ = capi.debug_structured_coords(two, 2 * mesh.x.shape[0])
    assert not ok
    # a large block stays fast (the neighbour search runs on the host's threads)
    import time
    big = wfx.create_box_hex(48, 1, (1.0,) * 3)
    t0 = time.perf_counter()
    ok, ijk = capi.debug_structured_coords(big.xdofs, big.x.shape[0])
    assert ok and ijk.max() == 47 and time.perf_counter() - t0 < 20
