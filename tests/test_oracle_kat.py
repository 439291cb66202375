"""Known-answer tests that pin the CPU oracle.

The reference ships no golden vectors for this path and cannot be built here, so the oracle
(oracle/wave_oracle.c) is pinned by analytic identities that follow from the reference's own
formulas, plus the one literal vector the survey records (the P2 tensor permutation).
"""
import numpy as np
import pytest

L = 0.1
C0 = 1500.0


def _mesh(wfx, N, P, perturb=0.0, **kw):
    return wfx.create_box_hex(N, P, (L, L, L), perturb=perturb, **kw)


def test_perm_p2_literal(orc):
    # tensor -> DOLFINx local dof for P2
    want = [0, 4, 10, 2, 6, 14, 9, 17, 22, 1, 5, 12, 3, 7, 15, 11, 18, 23, 8, 16, 21, 13, 19, 24, 20, 25, 26]
    assert orc.perm(2).tolist() == want


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 6, 7])
def test_perm_is_bijection(orc, P):
    p = orc.perm(P)
    assert sorted(p.tolist()) == list(range((P + 1) ** 3))


def test_gll5_literal(orc):
    # GLL-5 on [0,1] in [0, 1, interior] order
    pts, wts = orc.gll(4)
    np.testing.assert_allclose(pts, [0, 1, 0.1726731646, 0.5, 0.8273268354], atol=1e-10)
    np.testing.assert_allclose(wts, [0.05, 0.05, 0.2722222222, 0.3555555556, 0.2722222222], atol=1e-10)


@pytest.mark.parametrize("P", [2, 3, 4, 5, 6, 7])
def test_gll_exactness_and_derivative_matrix(orc, P):
    pts, wts = orc.gll(P)
    assert abs(wts.sum() - 1.0) < 1e-15
    for k in range(2 * P):  # GLL with P+1 points is exact to degree 2P-1
        assert abs((wts * pts ** k).sum() - 1.0 / (k + 1)) < 1e-14
    D = orc.deriv_1d(P, clamp=False)
    # D differentiates polynomials of degree <= P exactly at the nodes
    for k in range(P + 1):
        want = k * pts ** (k - 1) if k else np.zeros_like(pts)
        np.testing.assert_allclose(D @ pts ** k, want, atol=5e-13)
    # App. A.5: D[0][0] = -P(P+1)/2, corner-to-corner entry = +-1 exactly after the clamp
    Dc = orc.deriv_1d(P, clamp=True)
    assert abs(Dc[0, 0] + P * (P + 1) / 2) < 1e-12
    assert abs(Dc[0, 1]) == 1.0 and abs(Dc[1, 0]) == 1.0
    inter = np.arange(2, P + 1)
    assert np.all(Dc[inter, inter] == 0.0)


def test_clamp_rule(orc):
    c = orc.lib().wo_clamp_value
    assert c(5e-9) == 0.0 and c(-9.9e-9) == 0.0 and c(1.1e-8) == 1.1e-8
    assert c(1.0 + 5e-6) == 1.0 and c(-1.0 + 9e-6) == -1.0 and c(1.0 + 2e-5) == 1.0 + 2e-5


@pytest.mark.parametrize("P", [2, 4])
def test_dense_tables_are_tensor_products(orc, P):
    n, nd = P + 1, (P + 1) ** 3
    T = orc.tabulate_dphi(P)
    D, perm = orc.deriv_1d(P), orc.perm(P)
    # d/dy table at point q=(a,b,c) and dof (a,j,c) equals D[b][j]; zero when a or c differ
    rng = np.random.default_rng(0)
    for _ in range(50):
        a, b, c, j = rng.integers(0, n, 4)
        q = (a * n + b) * n + c
        assert T[1, q, perm[(a * n + j) * n + c]] == D[b, j]
        a2 = (a + 1) % n
        assert T[1, q, perm[(a2 * n + j) * n + c]] == 0.0
    assert np.count_nonzero(T) <= 3 * nd * n


@pytest.mark.parametrize("P,perturb", [(2, 0.0), (4, 0.0), (4, 0.15), (3, 0.15)])
def test_mass_sums_to_volume(wfx, orc, P, perturb):
    mesh = _mesh(wfx, 3, P, perturb)
    _, detJ = orc.precompute_geometric_data(mesh, P)
    m = np.zeros(mesh.ndofs)
    orc.mass_apply(mesh, P, detJ, np.ones(mesh.ndofs), m)
    assert abs(m.sum() - L ** 3) < 1e-15 * 10
    assert m.min() > 0


def test_mass_interior_node_value(wfx, orc):
    # cube cells of side h: m at a cell-interior node = h^3 w_a w_b w_c
    P, N = 4, 2
    mesh = _mesh(wfx, N, P)
    _, detJ = orc.precompute_geometric_data(mesh, P)
    m = np.zeros(mesh.ndofs)
    orc.mass_apply(mesh, P, detJ, np.ones(mesh.ndofs), m)
    _, w = orc.gll(P)
    h = L / N
    dof = mesh.dofmap[0, orc.perm(P)[(2 * 5 + 3) * 5 + 4]]
    want = h ** 3 * w[2] * w[3] * w[4]
    assert abs(m[dof] - want) < 1e-14 * want


def test_geometry_cube_values(wfx, orc):
    # App. A.6: cube of side h => G = h w_q I, detJ w = h^3 w_q
    P, N = 3, 2
    mesh = _mesh(wfx, N, P)
    G, detJ = orc.precompute_geometric_data(mesh, P)
    _, w = orc.gll(P)
    wq = np.einsum("i,j,k->ijk", w, w, w).reshape(-1)
    h = L / N
    np.testing.assert_allclose(detJ[0], h ** 3 * wq, rtol=1e-13)
    for a in range(3):
        np.testing.assert_allclose(G[1, :, a, a], h * wq, rtol=1e-12)
    assert np.all(G[:, :, 0, 1] == 0.0) and np.all(G[:, :, 1, 2] == 0.0)


@pytest.mark.parametrize("P", [2, 4])
def test_stiffness_kills_constants_and_linear_energy(wfx, orc, P):
    mesh = _mesh(wfx, 3, P)
    G, _ = orc.precompute_geometric_data(mesh, P)
    y = np.zeros(mesh.ndofs)
    orc.stiffness_apply(mesh, P, G, np.ones(mesh.ndofs), y)
    scale = C0 ** 2 * (L / 3)  # magnitude of one diagonal contribution
    assert np.abs(y).max() < 1e-9 * scale
    # x^T K x for x = a.r equals -c0^2 |a|^2 vol (exact on affine cells)
    a = np.array([1.0, -2.0, 0.5])
    x = wfx.dof_coordinates(mesh) @ a
    y[:] = 0
    orc.stiffness_apply(mesh, P, G, x, y)
    want = -C0 ** 2 * (a @ a) * L ** 3
    assert abs(x @ y - want) < 1e-12 * abs(want)


@pytest.mark.parametrize("P,perturb", [(2, 0.15), (3, 0.15), (4, 0.0), (4, 0.15), (5, 0.15)])
def test_stiffness_symmetric_and_sumfact_equals_dense(wfx, orc, P, perturb):
    mesh = _mesh(wfx, 2 if P > 4 else 3, P, perturb, renumber=7)
    G, _ = orc.precompute_geometric_data(mesh, P)
    rng = np.random.default_rng(42)
    x1, x2 = rng.standard_normal(mesh.ndofs), rng.standard_normal(mesh.ndofs)
    y1, y2, y1s = np.zeros(mesh.ndofs), np.zeros(mesh.ndofs), np.zeros(mesh.ndofs)
    orc.stiffness_apply(mesh, P, G, x1, y1)
    orc.stiffness_apply(mesh, P, G, x2, y2)
    orc.stiffness_apply(mesh, P, G, x1, y1s, dense=False)
    # G is symmetric to roundoff only ((K d) K^T), hence 1e-12 not 1e-16
    assert abs(x2 @ y1 - x1 @ y2) < 1e-12 * abs(x2 @ y1)
    assert np.linalg.norm(y1 - y1s) < 1e-14 * np.linalg.norm(y1)
    # accumulation semantics: y += A x
    y3 = y1.copy()
    orc.stiffness_apply(mesh, P, G, x1, y3)
    np.testing.assert_allclose(y3, 2 * y1, rtol=1e-14)


def test_openmp_path_matches_serial(wfx, orc):
    mesh = _mesh(wfx, 3, 4, 0.15)
    G, _ = orc.precompute_geometric_data(mesh, 4)
    x = np.random.default_rng(1).standard_normal(mesh.ndofs)
    y1, y2 = np.zeros(mesh.ndofs), np.zeros(mesh.ndofs)
    orc.stiffness_apply(mesh, 4, G, x, y1, nthreads=1)
    orc.stiffness_apply(mesh, 4, G, x, y2, nthreads=3)
    assert np.linalg.norm(y1 - y2) < 1e-14 * np.linalg.norm(y1)


def test_boundary_masses_sum_to_areas(wfx, orc):
    mesh = _mesh(wfx, 3, 4, 0.15)
    m1, m2 = orc.boundary_facet_mass(mesh, 4)
    assert abs(m1.sum() - L * L) < 1e-15 and abs(m2.sum() - L * L) < 1e-15
    X = wfx.dof_coordinates(mesh)
    assert np.all(X[m1 > 0, 0] < 1e-14) and np.all(np.abs(X[m2 > 0, 0] - L) < 1e-14)


def test_general_point_building_blocks(wfx, orc):
    # common/precompute.hpp variants: K = J^-1, G = K K^T detJ w, no fabs / clamp
    mesh = _mesh(wfx, 2, 2, 0.15)
    pts, wts = orc.gauss_legendre(3)
    P3 = np.array([[a, b, c] for a in pts for b in pts for c in pts])
    W3 = np.array([a * b * c for a in wts for b in wts for c in wts])
    d = orc.jacobian_data(mesh, P3, W3)
    I = np.einsum("cqij,cqjk->cqik", d["J"], d["K"])
    np.testing.assert_allclose(I, np.broadcast_to(np.eye(3), I.shape), atol=1e-12)
    np.testing.assert_allclose(d["detJ"], np.linalg.det(d["J"]), rtol=1e-12)
    want = np.einsum("cqik,cqjk->cqij", d["K"], d["K"]) * (d["detJ"] * W3)[..., None, None]
    np.testing.assert_allclose(d["G"], want, rtol=1e-12, atol=1e-18)
    assert abs((d["detJ"] * W3).sum() - L ** 3) < 1e-15


def test_tabulate_1d_partition_of_unity(orc):
    t0, t1 = orc.tabulate_1d(4, 8, 0), orc.tabulate_1d(4, 8, 1)
    np.testing.assert_allclose(t0.sum(axis=1), 1.0, atol=1e-14)
    np.testing.assert_allclose(t1.sum(axis=1), 0.0, atol=1e-12)


def test_rk4_plane_wave_sanity(wfx, orc):
    """cpu_planar3d in miniature: a few periods of the windowed source launch a wave from
    x=0 that travels at c0; before it arrives the far field is still zero and the solution
    only depends on x."""
    P, c0, f0, p0 = 4, 1500.0, 0.5e6, 6e4
    Lx = 0.012  # 4 wavelengths (lambda = 3 mm)
    mesh = wfx.create_box_hex((16, 1, 1), P, (Lx, Lx / 16, Lx / 16))
    G, detJ = orc.precompute_geometric_data(mesh, P)
    m = np.zeros(mesh.ndofs)
    orc.mass_apply(mesh, P, detJ, np.ones(mesh.ndofs), m)
    m1, m2 = orc.boundary_facet_mass(mesh, P)
    dt = wfx.cfl_timestep(mesh.h_min, c0, P, f0)
    u, v = np.zeros(mesh.ndofs), np.zeros(mesh.ndofs)
    tf = 2.0 / f0
    steps, t_end = orc.rk4(mesh, P, G, m, m1, m2, c0, f0, p0, 0.0, tf, dt, u, v, sumfact=True)
    assert steps > 10 and abs(t_end - tf) < 1e-15
    X = wfx.dof_coordinates(mesh)
    assert np.isfinite(u).all() and np.abs(u).max() > 0
    # the front has travelled c0*tf = 6 mm: nothing beyond ~7 mm yet
    assert np.abs(u[X[:, 0] > 0.008]).max() < 1e-3 * np.abs(u).max()
    # planar: u depends on x only
    key = np.round(X[:, 0] / Lx * 1e6).astype(np.int64)
    for k in np.unique(key)[::7]:
        vals = u[key == k]
        assert np.ptp(vals) <= 1e-9 * np.abs(u).max()


@pytest.mark.parametrize("P", [2, 3, 4])
def test_stiffness_against_first_principles_numpy(wfx, orc, P):
    """Independent derivation, sharing no code with the oracle or the library tables: GLL nodes from the
    roots of P_n' (numpy Legendre), weights 2/(n(n+1) P_n(x)^2), Lagrange derivatives by barycentric
    formulas, the trilinear Jacobian differentiated by hand, and
        v^T K u = -c0^2 sum_cells sum_q grad_X u(q) . G_q grad_X v(q),  G_q = clamp(J_q^-1 (w_q |det J_q|) J_q^-T)
    on a perturbed (non-affine) mesh with random fields.  Only the dof layout (perm, dofmap) is taken
    from the code under test -- it is pinned separately by the literal P2 permutation."""
    from numpy.polynomial import legendre as Lg
    n = P + 1
    # GLL on [-1,1]: endpoints and roots of P_P'; map to [0,1]
    cP = np.zeros(P + 1)
    cP[P] = 1.0
    xi = np.concatenate([[-1.0], np.sort(Lg.legroots(Lg.legder(cP))), [1.0]])
    w = 2.0 / (P * (P + 1) * Lg.legval(xi, cP) ** 2)
    x01, w01 = (xi + 1) / 2, w / 2
    # barycentric derivative matrix on ascending nodes: Dasc[q,i] = l_i'(x_q)
    bw = np.array([1.0 / np.prod([x01[i] - x01[j] for j in range(n) if j != i]) for i in range(n)])
    Dasc = np.zeros((n, n))
    for q in range(n):
        for i in range(n):
            if i != q:
                Dasc[q, i] = bw[i] / bw[q] / (x01[q] - x01[i])
        Dasc[q, q] = -Dasc[q].sum()
    # code ordering of 1-D nodes: [0, 1, interior...]  (ascending position of code index a)
    pos = np.array([0, P] + list(range(1, P)))
    pts = x01[pos]
    wts = w01[pos]
    D = Dasc[np.ix_(pos, pos)]
    mesh = _mesh(wfx, 2, P, perturb=0.2)
    perm = orc.perm(P)
    rng = np.random.default_rng(11)
    u, v = rng.standard_normal(mesh.ndofs), rng.standard_normal(mesh.ndofs)
    total = 0.0
    A, B, Cc = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    for c in range(mesh.ncells):
        dofs = mesh.dofmap[c][perm]                      # tensor order t = (a*n + b)*n + c
        U, V = u[dofs].reshape(n, n, n), v[dofs].reshape(n, n, n)
        gU = np.stack([np.einsum("qa,abc->qbc", D, U), np.einsum("qb,abc->aqc", D, U), np.einsum("qc,abc->abq", D, U)], -1)
        gV = np.stack([np.einsum("qa,abc->qbc", D, V), np.einsum("qb,abc->aqc", D, V), np.einsum("qc,abc->abq", D, V)], -1)
        xv = mesh.x[mesh.xdofs[c]]                       # vertex v = ix + 2 iy + 4 iz
        X = np.stack([pts[A], pts[B], pts[Cc]], -1)      # reference point of every node
        J = np.zeros((n, n, n, 3, 3))
        for vtx in range(8):
            bits = [(vtx >> a) & 1 for a in range(3)]
            f = [X[..., a] if bits[a] else 1.0 - X[..., a] for a in range(3)]
            df = [1.0 if bits[a] else -1.0 for a in range(3)]
            for a in range(3):                           # d phi_v / d X_a
                g = df[a] * f[(a + 1) % 3] * f[(a + 2) % 3]
                J[..., :, a] += xv[vtx][:, None, None, None].transpose(1, 2, 3, 0) * g[..., None]
        detJ = np.abs(np.linalg.det(J))
        Jinv = np.linalg.inv(J)
        wq = wts[A] * wts[B] * wts[Cc]
        Gq = np.einsum("...ia,...ja->...ij", Jinv, Jinv) * (wq * detJ)[..., None, None]   # J^-1 (detJ w) J^-T
        # the reference's clamp of G (common/precomputation.hpp:105-107, xt::isclose defaults): with an
        # ABSOLUTE tolerance of 1e-8 it zeroes real entries on a mesh this small -- a 1e-7 relative effect
        # on v^T K u that belongs to the reference's definition of the operator
        for target in (-1.0, 0.0, 1.0):
            Gq = np.where(np.abs(Gq - target) <= 1e-8 + 1e-5 * abs(target), target, Gq)
        total += float(np.einsum("...i,...ij,...j->...", gU, Gq, gV).sum())
    want = -C0 * C0 * total
    G, _ = orc.precompute_geometric_data(mesh, P)
    ku = np.zeros(mesh.ndofs)
    orc.stiffness_apply(mesh, P, G, u, ku, dense=True)
    got = float(v @ ku)
    assert abs(got - want) <= 1e-11 * abs(want), (got, want)


def test_rk4_against_textbook_runge_kutta(wfx, orc):
    """The oracle's restatement of LinearGLLOpt::rk4 (stage bookkeeping of LinearGLL.hpp:244-270) equals
    the textbook classical RK4 applied to y' = F(t, y), y = (u, v),
        F = (v,  M^-1 (-c0^2 K u + c0^2 g(t) m1 - c0 m2 v)),
    written here directly in numpy (it shares the operator pieces with the oracle, not the time stepper)."""
    P, c0, f0, p0 = 3, C0, 0.5e6, 6e4
    mesh = wfx.create_box_hex((3, 2, 2), P, (L * 3 / 8, L * 2 / 8, L * 2 / 8), perturb=0.1)
    G, detJ = orc.precompute_geometric_data(mesh, P)
    m = np.zeros(mesh.ndofs)
    orc.mass_apply(mesh, P, detJ, np.ones(mesh.ndofs), m)
    m1, m2 = orc.boundary_facet_mass(mesh, P)
    dt = wfx.cfl_timestep(mesh.h_min, c0, P, f0)
    w0, T, alpha = 2.0 * np.pi * f0, 1.0 / f0, 4.0

    def F(t, u, v):
        window = 0.5 * (1.0 - np.cos(f0 * np.pi * t / alpha)) if t < T * alpha else 1.0
        g = window * p0 * w0 / c0 * np.cos(w0 * t)
        b = np.zeros(mesh.ndofs)
        orc.stiffness_apply(mesh, P, G, u, b, dense=True)
        return v.copy(), (b + c0 * c0 * g * m1 - c0 * m2 * v) / m

    steps = 7
    u, v, t = np.zeros(mesh.ndofs), np.zeros(mesh.ndofs), 0.0
    for _ in range(steps):
        k1u, k1v = F(t, u, v)
        k2u, k2v = F(t + dt / 2, u + dt / 2 * k1u, v + dt / 2 * k1v)
        k3u, k3v = F(t + dt / 2, u + dt / 2 * k2u, v + dt / 2 * k2v)
        k4u, k4v = F(t + dt, u + dt * k3u, v + dt * k3v)
        u = u + dt / 6 * (k1u + 2 * k2u + 2 * k3u + k4u)
        v = v + dt / 6 * (k1v + 2 * k2v + 2 * k3v + k4v)
        t += dt
    uo, vo = np.zeros(mesh.ndofs), np.zeros(mesh.ndofs)
    s, t_end = orc.rk4(mesh, P, G, m, m1, m2, c0, f0, p0, 0.0, 1.0, dt, uo, vo, max_steps=steps)
    assert s == steps and abs(t_end - t) < 1e-18
    assert np.abs(u).max() > 0
    assert np.linalg.norm(uo - u) <= 1e-12 * np.linalg.norm(u)
    assert np.linalg.norm(vo - v) <= 1e-12 * np.linalg.norm(v)


@pytest.mark.parametrize("P", [2, 4])
def test_lumped_and_facet_masses_against_first_principles(wfx, orc, P):
    """m_i = sum over cells of w_q |det J_q| at node i (collocated GLL), and the facet masses
    m_Gamma,i = sum over tagged facets of w_f |J_f| with |J_f| the area element of the face map, both
    evaluated here with an independent GLL rule and hand-differentiated trilinear map."""
    from numpy.polynomial import legendre as Lg
    n = P + 1
    cP = np.zeros(P + 1)
    cP[P] = 1.0
    xi = np.concatenate([[-1.0], np.sort(Lg.legroots(Lg.legder(cP))), [1.0]])
    w = 2.0 / (P * (P + 1) * Lg.legval(xi, cP) ** 2)
    pos = np.array([0, P] + list(range(1, P)))
    pts, wts = ((xi + 1) / 2)[pos], (w / 2)[pos]
    mesh = _mesh(wfx, 2, P, perturb=0.2)
    perm = orc.perm(P)
    A, B, Cc = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    X = np.stack([pts[A], pts[B], pts[Cc]], -1)

    def jacobian(xv):
        J = np.zeros((n, n, n, 3, 3))
        for vtx in range(8):
            bits = [(vtx >> a) & 1 for a in range(3)]
            f = [X[..., a] if bits[a] else 1.0 - X[..., a] for a in range(3)]
            for a in range(3):
                g = (1.0 if bits[a] else -1.0) * f[(a + 1) % 3] * f[(a + 2) % 3]
                J[..., :, a] += xv[vtx][None, None, None, :] * g[..., None]
        return J

    m = np.zeros(mesh.ndofs)
    m1, m2 = np.zeros(mesh.ndofs), np.zeros(mesh.ndofs)
    wq = wts[A] * wts[B] * wts[Cc]
    facets = {(int(c), int(f)): int(t) for c, f, t in zip(mesh.facet_cells, mesh.facet_local, mesh.facet_tags)}
    # hexahedron facets: 0 z-, 1 y-, 2 x-, 3 x+, 4 y+, 5 z+  -> (normal axis, fixed tensor index)
    face_of = {0: (2, 0), 1: (1, 0), 2: (0, 0), 3: (0, 1), 4: (1, 1), 5: (2, 1)}
    for c in range(mesh.ncells):
        dofs = mesh.dofmap[c][perm].reshape(n, n, n)
        J = jacobian(mesh.x[mesh.xdofs[c]])
        np.add.at(m, dofs, wq * np.abs(np.linalg.det(J)))
        for lf, (axis, side) in face_of.items():
            tag = facets.get((c, lf))
            if tag is None:
                continue
            t1, t2 = [a for a in range(3) if a != axis]
            area = np.linalg.norm(np.cross(J[..., :, t1], J[..., :, t2]), axis=-1)   # |dx/dX_t1 x dx/dX_t2|
            wf = (wts[[A, B, Cc][t1]] * wts[[A, B, Cc][t2]]) * area
            sel = [slice(None)] * 3
            sel[axis] = side                                                         # code index 0 / 1 = the two ends
            np.add.at(m1 if tag == 1 else m2, dofs[tuple(sel)], wf[tuple(sel)])
    _, detJ = orc.precompute_geometric_data(mesh, P)
    mo = np.zeros(mesh.ndofs)
    orc.mass_apply(mesh, P, detJ, np.ones(mesh.ndofs), mo)
    m1o, m2o = orc.boundary_facet_mass(mesh, P)
    np.testing.assert_allclose(mo, m, rtol=1e-12)
    assert m1.sum() > 0 and m2.sum() > 0
    np.testing.assert_allclose(m1o, m1, rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(m2o, m2, rtol=1e-12, atol=1e-18)
