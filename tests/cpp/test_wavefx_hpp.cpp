// Exercises the C++ wrappers (wave-fenics_b200/hpp/wavefx.hpp) the way the reference's
// drivers use its classes (demo/gpu_operator_monolithic/main.cpp:94-121, demo/cpu_planar3d/main.cpp:75-90):
// build V, construct the operators, apply them to host vectors, run the RK4 model.
// Self-checking; prints "hpp ok" on success.  Needs a GPU to run, only a compiler to build.
#include "wavefx.hpp"

#include <cmath>
#include <cstdio>
#include <numeric>
#include <random>

struct BoxMesh
{
  int P, N;
  std::vector<double> x;
  std::vector<std::int32_t> xdofs, dofmap, fcell, flocal, ftag;
  std::int64_t ndofs;
  wavefx::SpaceView view() const
  {
    wavefx::SpaceView V;
    V.degree = P;
    V.ncells = (std::int64_t)N * N * N;
    V.npoints = (std::int64_t)x.size() / 3;
    V.x = x.data();
    V.xdofs = xdofs.data();
    V.ndofs = V.size_local = ndofs;
    V.dofmap = dofmap.data();
    V.nfacets = (std::int64_t)fcell.size();
    V.facet_cell = fcell.data();
    V.facet_local = flocal.data();
    V.facet_tag = ftag.data();
    return V;
  }
};

static BoxMesh make_box(int N, int P, double L)
{
  BoxMesh m;
  m.P = P;
  m.N = N;
  const int n = P + 1, nd = n * n * n, M = P * N + 1;
  m.ndofs = (std::int64_t)M * M * M;
  for (int i = 0; i <= N; ++i)
    for (int j = 0; j <= N; ++j)
      for (int k = 0; k <= N; ++k)
      {
        m.x.push_back(L * i / N);
        m.x.push_back(L * j / N);
        m.x.push_back(L * k / N);
      }
  std::vector<std::int32_t> perm(nd);
  wavefx::check(wfx_compute_permutations(P, perm.data()));
  auto pos = [&](int a) { return a == 0 ? 0 : (a == 1 ? P : a - 1); };
  for (int cx = 0; cx < N; ++cx)
    for (int cy = 0; cy < N; ++cy)
      for (int cz = 0; cz < N; ++cz)
      {
        const std::int32_t c = (cx * N + cy) * N + cz;
        for (int v = 0; v < 8; ++v)
          m.xdofs.push_back(((cx + (v & 1)) * (N + 1) + (cy + ((v >> 1) & 1))) * (N + 1) + (cz + ((v >> 2) & 1)));
        std::vector<std::int32_t> cd(nd);
        for (int a = 0; a < n; ++a)
          for (int b = 0; b < n; ++b)
            for (int d = 0; d < n; ++d)
              cd[perm[(a * n + b) * n + d]] = ((cx * P + pos(a)) * M + (cy * P + pos(b))) * M + (cz * P + pos(d));
        m.dofmap.insert(m.dofmap.end(), cd.begin(), cd.end());
        if (cx == 0) { m.fcell.push_back(c); m.flocal.push_back(2); m.ftag.push_back(1); }
        if (cx == N - 1) { m.fcell.push_back(c); m.flocal.push_back(3); m.ftag.push_back(2); }
      }
  return m;
}

#define REQUIRE(cond)                                                              \
  do {                                                                             \
    if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } \
  } while (0)

int main()
{
  const int P = 4, N = 4;
  const double L = 0.1;
  BoxMesh mesh = make_box(N, P, L);
  auto V = mesh.view();
  auto ctx = std::make_shared<wavefx::Context>(0);
  auto geom = std::make_shared<wavefx::Geometry<double>>(ctx, V);
  std::map<std::string, double> params{{"c0", 1500.0}};
  wavefx::StiffnessOperator<double> stiff(geom, V, P, params);
  wavefx::MassOperatorCPU<double> mass(geom, V, P);
  REQUIRE(stiff.num_cells() == (std::size_t)N * N * N && stiff.num_dofs() == 125);

  std::vector<double> one(mesh.ndofs, 1.0), m(mesh.ndofs, 0.0), y(mesh.ndofs, 0.0);
  mass(one, m); // m = M.1 (LinearGLL.hpp:102-110)
  REQUIRE(std::fabs(std::accumulate(m.begin(), m.end(), 0.0) - L * L * L) < 1e-12 * L * L * L);
  stiff(one, y); // K annihilates constants
  double ymax = 0;
  for (double v : y) ymax = std::max(ymax, std::fabs(v));
  REQUIRE(ymax < 1e-9 * 1500.0 * 1500.0 * L / N);

  std::mt19937 rng(42);
  std::normal_distribution<double> nrm;
  std::vector<double> a(mesh.ndofs), b(mesh.ndofs), Ka(mesh.ndofs, 0.0), Kb(mesh.ndofs, 0.0);
  for (auto& v : a) v = nrm(rng);
  for (auto& v : b) v = nrm(rng);
  stiff(a, Ka);
  stiff(b, Kb);
  const double s1 = std::inner_product(b.begin(), b.end(), Ka.begin(), 0.0);
  const double s2 = std::inner_product(a.begin(), a.end(), Kb.begin(), 0.0);
  REQUIRE(std::fabs(s1 - s2) < 1e-11 * std::fabs(s1)); // symmetry
  std::vector<double> Ka2 = Ka;
  stiff(a, Ka2); // accumulation semantics: y += A x
  for (std::size_t i = 0; i < Ka.size(); ++i) REQUIRE(std::fabs(Ka2[i] - 2 * Ka[i]) <= 1e-12 * std::fabs(Ka[i]) + 1e-300);

  bool threw = false;
  try { std::vector<double> shorty(3); stiff(a, shorty); } catch (const std::runtime_error&) { threw = true; }
  REQUIRE(threw);

  int degree = P;
  double c0 = 1500.0, f0 = 0.5e6, p0 = 6e4;
  wavefx::LinearGLLOpt eqn(ctx, V, degree, c0, f0, p0);
  eqn.init();
  double t0 = 0.0, dt = 0.5 * std::sqrt(3.0) * (L / N) / (c0 * P * P), tf = 20.5 * dt;
  REQUIRE(eqn.rk4(t0, tf, dt) == 21);
  std::vector<double> u, v;
  eqn.solution(u, v);
  double umax = 0;
  for (double w : u) { REQUIRE(std::isfinite(w)); umax = std::max(umax, std::fabs(w)); }
  REQUIRE(umax > 0);
  // structured fast path on the (affine) box mesh; batched host apply equals the single-vector path
  REQUIRE(geom->num_affine_cells() == (std::int64_t)N * N * N && stiff.affine_fast_path());
  {
    std::vector<double> y1(mesh.ndofs, 0.0), y2(mesh.ndofs, 0.0), r1(mesh.ndofs), r2(mesh.ndofs);
    stiff(a, y1);
    stiff(b, y2);
    wavefx::stiffness_mass_apply_batch<double>(stiff, mass, {a.data(), b.data()}, {r1.data(), r2.data()});
    for (std::size_t i = 0; i < y1.size(); ++i)
    {
      REQUIRE(std::fabs(r1[i] - y1[i] / m[i]) <= 1e-12 * std::fabs(y1[i] / m[i]) + 1e-300);
      REQUIRE(std::fabs(r2[i] - y2[i] / m[i]) <= 1e-12 * std::fabs(y2[i] / m[i]) + 1e-300);
    }
  }
  // probes, snapshots and a restart from a snapshot
  {
    wavefx::LinearGLLOpt run(ctx, V, degree, c0, f0, p0);
    run.init();
    run.set_probes({0, (std::int32_t)(mesh.ndofs / 2)}, 64);
    std::vector<double> su, sv;
    std::int64_t sstep = -1;
    double st = 0;
    run.set_snapshot(10, [&](std::int64_t step, double t, const double* uu, const double* vv) {
      if (step == 10) { sstep = step; st = t; su.assign(uu, uu + mesh.ndofs); sv.assign(vv, vv + mesh.ndofs); }
    });
    double a0 = 0.0;
    REQUIRE(run.rk4(a0, tf, dt) == 21);
    std::vector<double> pt, pv, u1, v1;
    run.probe_series(pt, pv);
    REQUIRE(pt.size() == 21 && pv.size() == 42 && sstep == 10);
    run.solution(u1, v1);
    for (std::size_t i = 0; i < u.size(); ++i) REQUIRE(u1[i] == u[i]); // same run as above, bitwise
    wavefx::LinearGLLOpt again(ctx, V, degree, c0, f0, p0);
    again.set_state(su, sv);
    REQUIRE(again.rk4(st, tf, dt) == 11);
    std::vector<double> u2, v2;
    again.solution(u2, v2);
    for (std::size_t i = 0; i < u.size(); ++i) REQUIRE(u2[i] == u[i] && v2[i] == v[i]);
  }
  std::printf("hpp ok: |u|max %.3e after 21 steps\n", umax);
  return 0;
}
