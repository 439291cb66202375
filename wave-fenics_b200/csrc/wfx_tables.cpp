// Host-side element tables: GLL rule, 1-D derivative matrix, tensor->DOLFINx dof
// permutation, 1-D tabulation at Gauss points.  These replace the Basix calls of the
// reference (common/operators.hpp:13-32, common/permute.hpp:12-17,
// common/precompute.hpp:179-199); Basix itself is not a dependency.
#include "wfx_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <array>

namespace wfx
{
static thread_local std::string g_error;

void set_error(const char* fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

void fail(const char* fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error(buf);
}

const char* last_error() { return g_error.c_str(); }

double clamp_m101(double v)
{
  // xt::isclose(v, t): |v - t| <= atol + rtol |t|, atol = 1e-8, rtol = 1e-5, applied
  // for t = -1, 0, 1 in that order (common/precomputation.hpp:105-107).
  auto close = [](double a, double t) { return std::fabs(a - t) <= 1e-8 + 1e-5 * std::fabs(t); };
  if (close(v, -1.0)) v = -1.0;
  if (close(v, 0.0)) v = 0.0;
  if (close(v, 1.0)) v = 1.0;
  return v;
}

namespace
{
using ld = long double;

// Legendre polynomial L_N and its first derivative by the three-term recurrence.
void legendre_pair(int N, ld x, ld& L, ld& dL)
{
  ld Lm1 = 1, Lc = x, dLm1 = 0, dLc = 1;
  if (N == 0) { L = 1; dL = 0; return; }
  for (int k = 1; k < N; ++k)
  {
    ld Ln = ((2 * k + 1) * x * Lc - k * Lm1) / (k + 1);
    ld dLn = dLm1 + (2 * k + 1) * Lc;
    Lm1 = Lc; Lc = Ln; dLm1 = dLc; dLc = dLn;
  }
  L = Lc;
  dL = dLc;
}

// Nodes of the (P+1)-point Gauss-Lobatto-Legendre rule on [0,1], ascending, with weights.
void gll_sorted(int P, std::vector<ld>& x, std::vector<ld>& w)
{
  const int n = P + 1;
  const ld pi = acosl(-1.0L);
  x.assign(n, 0);
  w.assign(n, 0);
  x[0] = -1;
  x[P] = 1;
  // interior nodes: zeros of L_P'.  Newton on L_P' with L_P'' from Legendre's ODE.
  for (int i = 1; i <= P / 2; ++i)
  {
    ld xi = -cosl(pi * i / P);
    for (int it = 0; it < 200; ++it)
    {
      ld L, dL;
      legendre_pair(P, xi, L, dL);
      ld d2L = (2 * xi * dL - ld(P) * (P + 1) * L) / (1 - xi * xi);
      ld step = dL / d2L;
      xi -= step;
      if (fabsl(step) < 1e-20L) break;
    }
    x[i] = xi;
    x[P - i] = -xi;
  }
  if (P % 2 == 0) x[P / 2] = 0;
  for (int i = 0; i < n; ++i)
  {
    ld L, dL;
    legendre_pair(P, x[i], L, dL);
    w[i] = 2 / (ld(P) * (P + 1) * L * L);
  }
  for (int i = 0; i < n; ++i)
  {
    x[i] = (x[i] + 1) / 2;
    w[i] = w[i] / 2;
  }
}

// position (0..P, ascending) of 1-D dof a in the interval element's [0, 1, interior] order
inline int lattice_pos(int a, int P) { return a == 0 ? 0 : (a == 1 ? P : a - 1); }
} // namespace

void gll_points_weights(int P, double* pts, double* wts)
{
  if (P < 1 || P + 1 > WFX_MAXN) fail("degree %d out of range [1,%d]", P, WFX_MAXN - 1);
  std::vector<ld> x, w;
  gll_sorted(P, x, w);
  for (int a = 0; a <= P; ++a)
  {
    pts[a] = (double)x[lattice_pos(a, P)];
    wts[a] = (double)w[lattice_pos(a, P)];
  }
}

void deriv_1d(int P, double* D, bool clamp)
{
  if (P < 1 || P + 1 > WFX_MAXN) fail("degree %d out of range [1,%d]", P, WFX_MAXN - 1);
  const int n = P + 1;
  std::vector<ld> x, w;
  gll_sorted(P, x, w);
  // barycentric weights
  std::vector<ld> lam(n, 1);
  for (int i = 0; i < n; ++i)
    for (int m = 0; m < n; ++m)
      if (m != i) lam[i] *= (x[i] - x[m]);
  for (int q = 0; q < n; ++q)
    for (int i = 0; i < n; ++i)
    {
      const int qs = lattice_pos(q, P), is = lattice_pos(i, P);
      ld v = 0;
      if (qs != is) v = lam[qs] / (lam[is] * (x[qs] - x[is]));
      else
        for (int m = 0; m < n; ++m)
          if (m != is) v += 1 / (x[is] - x[m]);
      D[q * n + i] = clamp ? clamp_m101((double)v) : (double)v;
    }
}

void tensor_perm(int P, int32_t* perm)
{
  if (P < 1 || P + 1 > WFX_MAXN) fail("degree %d out of range [1,%d]", P, WFX_MAXN - 1);
  const int n = P + 1, ni = P - 1;
  // DOLFINx hexahedron: vertex v at (v&1, (v>>1)&1, (v>>2)&1); sub-entity vertex lists.
  static const int E[12][2] = {{0, 1}, {0, 2}, {0, 4}, {1, 3}, {1, 5}, {2, 3},
                               {2, 6}, {3, 7}, {4, 5}, {4, 6}, {5, 7}, {6, 7}};
  static const int F[6][4] = {{0, 1, 2, 3}, {0, 1, 4, 5}, {0, 2, 4, 6},
                              {1, 3, 5, 7}, {2, 3, 6, 7}, {4, 5, 6, 7}};
  auto vpos = [&](int v) { return std::array<int, 3>{(v & 1) * P, ((v >> 1) & 1) * P, ((v >> 2) & 1) * P}; };
  std::map<std::array<int, 3>, int> dof_at; // lattice position -> DOLFINx local dof
  int dof = 0;
  for (int v = 0; v < 8; ++v) dof_at[vpos(v)] = dof++;
  for (int e = 0; e < 12; ++e)
  {
    auto a = vpos(E[e][0]), b = vpos(E[e][1]);
    for (int i = 1; i <= ni; ++i)
      dof_at[{a[0] + (b[0] - a[0]) / P * i, a[1] + (b[1] - a[1]) / P * i, a[2] + (b[2] - a[2]) / P * i}] = dof++;
  }
  for (int f = 0; f < 6; ++f)
  {
    auto o = vpos(F[f][0]), a = vpos(F[f][1]), b = vpos(F[f][2]);
    for (int j = 1; j <= ni; ++j)   // second face axis slow
      for (int i = 1; i <= ni; ++i) // first face axis fast
      {
        std::array<int, 3> p;
        for (int d = 0; d < 3; ++d) p[d] = o[d] + (a[d] - o[d]) / P * i + (b[d] - o[d]) / P * j;
        dof_at[p] = dof++;
      }
  }
  for (int k = 1; k <= ni; ++k)
    for (int j = 1; j <= ni; ++j)
      for (int i = 1; i <= ni; ++i) dof_at[{i, j, k}] = dof++;
  if (dof != n * n * n || (int)dof_at.size() != n * n * n) fail("internal: dof lattice incomplete");
  for (int ix = 0; ix < n; ++ix)
    for (int iy = 0; iy < n; ++iy)
      for (int iz = 0; iz < n; ++iz)
        perm[(ix * n + iy) * n + iz]
            = dof_at.at({lattice_pos(ix, P), lattice_pos(iy, P), lattice_pos(iz, P)});
}
} // namespace wfx

using namespace wfx;

extern "C" const char* wfx_last_error(void) { return wfx::last_error(); }
extern "C" int wfx_version(void) { return WFX_VERSION; }

extern "C" int wfx_gll(int P, double* pts, double* wts)
{
  WFX_API_BEGIN
  gll_points_weights(P, pts, wts);
  WFX_API_END
}

extern "C" int wfx_deriv_1d(int P, double* D)
{
  WFX_API_BEGIN
  deriv_1d(P, D, true);
  WFX_API_END
}

extern "C" int wfx_compute_permutations(int P, int32_t* perm)
{
  WFX_API_BEGIN
  tensor_perm(P, perm);
  WFX_API_END
}

// tabulate_basis_and_permutation (common/operators.hpp:13-32): table[4][nq][nd], the basis and its
// three reference derivatives at the GLL points (= the nodes: collocated), dof axis in DOLFINx order,
// point axis in the tensor order of the quadrature, every entry clamped as at :26-29.
extern "C" int wfx_tabulate_basis_and_permutation(int P, double* table, int32_t* perm_out)
{
  WFX_API_BEGIN
  const int n = P + 1, nd = n * n * n;
  std::vector<int32_t> perm(nd);
  tensor_perm(P, perm.data());
  if (perm_out) std::copy(perm.begin(), perm.end(), perm_out);
  if (table)
  {
    double D[WFX_MAXN * WFX_MAXN];
    deriv_1d(P, D, true);
    std::fill(table, table + (size_t)4 * nd * nd, 0.0);
    auto at = [&](int a, size_t q, int dof) -> double& { return table[((size_t)a * nd + q) * nd + dof]; };
    for (int qa = 0; qa < n; ++qa)
      for (int qb = 0; qb < n; ++qb)
        for (int qc = 0; qc < n; ++qc)
        {
          const size_t q = ((size_t)qa * n + qb) * n + qc;
          at(0, q, perm[q]) = 1.0; // phi_i(x_q) = delta
          for (int i = 0; i < n; ++i)
          {
            at(1, q, perm[(i * n + qb) * n + qc]) = D[qa * n + i];
            at(2, q, perm[(qa * n + i) * n + qc]) = D[qb * n + i];
            at(3, q, perm[(qa * n + qb) * n + i]) = D[qc * n + i];
          }
        }
  }
  WFX_API_END
}

extern "C" int wfx_reorder_dofmap(int P, int64_t ncells, const int32_t* in, int32_t* out)
{
  WFX_API_BEGIN
  const int nd = (P + 1) * (P + 1) * (P + 1);
  std::vector<int32_t> perm(nd);
  tensor_perm(P, perm.data());
  for (int64_t c = 0; c < ncells; ++c)
    for (int t = 0; t < nd; ++t) out[c * nd + t] = in[c * nd + perm[t]];
  WFX_API_END
}

extern "C" int wfx_tabulate_1d(int P, int q, int derivative, double* table, int* npoints)
{
  WFX_API_BEGIN
  if (P < 1 || P + 1 > WFX_MAXN) fail("degree %d out of range", P);
  if (derivative < 0 || derivative > 1) fail("derivative must be 0 or 1");
  if (q < 0) fail("quadrature degree must be >= 0");
  const int m = (q + 2) / 2, n = P + 1;
  if (npoints) *npoints = m;
  if (!table) return 0;
  // Gauss-Legendre nodes on [0,1] (Gauss-Jacobi with alpha = beta = 0)
  const ld pi = acosl(-1.0L);
  std::vector<ld> g(m);
  for (int i = 0; i < m; ++i)
  {
    ld xi = -cosl(pi * (4 * i + 3) / (4 * m + 2));
    for (int it = 0; it < 200; ++it)
    {
      ld L, dL;
      legendre_pair(m, xi, L, dL);
      ld step = L / dL;
      xi -= step;
      if (fabsl(step) < 1e-20L) break;
    }
    g[i] = (xi + 1) / 2;
  }
  std::vector<ld> x, w;
  gll_sorted(P, x, w);
  for (int qi = 0; qi < m; ++qi)
    for (int a = 0; a < n; ++a)
    {
      const int is = lattice_pos(a, P);
      ld val = 0;
      if (derivative == 0)
      {
        val = 1;
        for (int r = 0; r < n; ++r)
          if (r != is) val *= (g[qi] - x[r]) / (x[is] - x[r]);
      }
      else
      {
        for (int s = 0; s < n; ++s)
        {
          if (s == is) continue;
          ld term = 1 / (x[is] - x[s]);
          for (int r = 0; r < n; ++r)
            if (r != is && r != s) term *= (g[qi] - x[r]) / (x[is] - x[r]);
          val += term;
        }
      }
      table[qi * n + a] = (double)val;
    }
  WFX_API_END
}
