/*
 * wave_oracle.c -- CPU restatement of the waveFEniCS matrix-free hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under wave-fenics_b200/ may call, link or
 * import this file.  It is used by tests/, by __graft_entry__.smoke() and by the
 * cpu_baseline / --impl reference legs of bench.py, always as the checker or the
 * timed CPU baseline, never as the product path.
 *
 * PARITY, what is pinned and what is not.  The reference (/root/reference) has no
 * tests, golden vectors or fixtures for this path, and its headers cannot be built
 * as a whole (DOLFINx, Basix, xtensor, FFCx, MPI are absent and un-vendored).
 *  PINNED to the reference's own code: the cell kernels skernel / mkernel, the call
 *   loops of StiffnessOperator / MassOperatorCPU, kernels::copy / axpy and
 *   LinearGLLOpt::init / f0 / f1 / rk4 are cut out of the reference headers at build
 *   time and compiled against container stand-ins (oracle/build_ref.py,
 *   oracle/ref_cpu_shim.cpp -> oracle/_ref/libwfref_cpu.so); the restatements below
 *   reproduce them BIT FOR BIT -- operators, right-hand side, whole RK4 trajectories
 *   (tests/test_reference_pins.py).  Likewise reorder_dofmap's loop, precompute.hpp's
 *   dot, the demo's CFL time step and the partition arithmetic decompose3d /
 *   compute_cartesian_indices (oracle/_ref/libwfref_mesh.so), and gather / scatter /
 *   transform1 against the reference's CUDA kernels (oracle/_ref/libwfref_cuda.so).
 *  UNPINNED (third-party arithmetic the reference calls, absent here): the Basix /
 *   DOLFINx pieces -- GLL quadrature, gll_warped Lagrange tabulation, tensor-product
 *   permutation, cmap tabulation, math::det / math::inv inside
 *   precompute_geometric_data, the FFCx facet kernel -- are restated from their
 *   published algorithms; each such function says [recalled].  Their pins are the
 *   analytic known answers in tests/test_oracle_kat.py.
 *
 * Every function cites the reference file:line it follows (relative to
 * /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define WO_MAXN 12 /* max points per direction (P <= 11) */

/* xt::isclose(a, b) with xtensor defaults rtol=1e-5, atol=1e-8
 * (common/precomputation.hpp:56-58,105-107; common/operators.hpp:27-29). */
static inline int wo_isclose(double v, double t) { return fabs(v - t) <= 1e-8 + 1e-5 * fabs(t); }

/* The three xt::filtration(...) = value statements, in the reference's order. */
static inline double wo_clamp(double v)
{
  if (wo_isclose(v, -1.0)) v = -1.0;
  if (wo_isclose(v, 0.0)) v = 0.0;
  if (wo_isclose(v, 1.0)) v = 1.0;
  return v;
}
double wo_clamp_value(double v) { return wo_clamp(v); }

/* ------------------------------------------------------------------------- */
/* 1-D GLL points / weights on [0,1], Basix ordering [0, 1, interior ascending]
 * [recalled: basix::quadrature::make_quadrature(gll, ...) as used at
 * common/precomputation.hpp:48-51 and common/operators.hpp:16-19]. */
static void legendre(int n, long double x, long double* p, long double* dp)
{
  long double p0 = 1.0L, p1 = x;
  if (n == 0) { *p = 1.0L; *dp = 0.0L; return; }
  for (int k = 2; k <= n; ++k)
  {
    long double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k;
    p0 = p1;
    p1 = pk;
  }
  *p = p1;
  *dp = n * (x * p1 - p0) / (x * x - 1.0L); /* valid for |x| != 1 */
}

static void gll_ascending(int P, long double* x, long double* w)
{
  const int n = P + 1;
  const long double pi = 3.14159265358979323846264338327950288L;
  x[0] = -1.0L;
  x[P] = 1.0L;
  for (int i = 1; i < P; ++i)
  {
    long double xi = -cosl(pi * i / P);
    for (int it = 0; it < 100; ++it)
    {
      long double p, dp;
      legendre(P, xi, &p, &dp);
      long double ddp = (2.0L * xi * dp - (long double)P * (P + 1) * p) / (1.0L - xi * xi);
      long double dx = dp / ddp;
      xi -= dx;
      if (fabsl(dx) < 1e-19L) break;
    }
    x[i] = xi;
  }
  for (int i = 0; i < n / 2; ++i) /* symmetrise */
  {
    long double a = 0.5L * (x[P - i] - x[i]);
    x[i] = -a;
    x[P - i] = a;
  }
  if (n % 2 == 1) x[P / 2] = 0.0L;
  for (int i = 0; i < n; ++i)
  {
    long double p, dp;
    if (i == 0 || i == P) p = (i == 0 && (P % 2)) ? -1.0L : 1.0L;
    else legendre(P, x[i], &p, &dp);
    w[i] = 2.0L / ((long double)P * (P + 1) * p * p);
  }
  for (int i = 0; i < n; ++i) /* map [-1,1] -> [0,1] */
  {
    x[i] = 0.5L * (x[i] + 1.0L);
    w[i] = 0.5L * w[i];
  }
}

/* ascending index of 1-D dof a in the [0, 1, interior] ordering */
static inline int asc_of(int a, int P) { return a == 0 ? 0 : (a == 1 ? P : a - 1); }

int wo_gll(int P, double* pts, double* wts)
{
  long double x[WO_MAXN], w[WO_MAXN];
  if (P < 1 || P + 1 > WO_MAXN) return -1;
  gll_ascending(P, x, w);
  for (int a = 0; a <= P; ++a)
  {
    pts[a] = (double)x[asc_of(a, P)];
    wts[a] = (double)w[asc_of(a, P)];
  }
  return 0;
}

/* 1-D derivative matrix D[q][i] = l_i'(x_q) on [0,1] in [0,1,interior] ordering,
 * clamped like the reference clamps its tables (common/operators.hpp:26-29).
 * [recalled: basix element.tabulate(1, pts) restricted to one direction]. */
int wo_deriv_1d(int P, double* D, int clamp)
{
  long double x[WO_MAXN], w[WO_MAXN], bw[WO_MAXN], Da[WO_MAXN][WO_MAXN];
  const int n = P + 1;
  if (P < 1 || n > WO_MAXN) return -1;
  gll_ascending(P, x, w);
  for (int i = 0; i < n; ++i)
  {
    bw[i] = 1.0L;
    for (int m = 0; m < n; ++m)
      if (m != i) bw[i] /= (x[i] - x[m]);
  }
  for (int q = 0; q < n; ++q)
    for (int i = 0; i < n; ++i)
    {
      if (q != i) Da[q][i] = (bw[i] / bw[q]) / (x[q] - x[i]);
      else
      {
        long double s = 0.0L;
        for (int m = 0; m < n; ++m)
          if (m != i) s += 1.0L / (x[i] - x[m]);
        Da[q][i] = s;
      }
    }
  for (int q = 0; q < n; ++q)
    for (int i = 0; i < n; ++i)
    {
      double v = (double)Da[asc_of(q, P)][asc_of(i, P)];
      D[q * n + i] = clamp ? wo_clamp(v) : v;
    }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Tensor index t = ix*n^2 + iy*n + iz  ->  DOLFINx local dof.
 * [recalled: element.get_tensor_product_representation()[0] second member, used
 * at common/operators.hpp:24, common/permute.hpp:17, common/precompute.hpp:197].
 * Built from the DOLFINx hexahedron entity numbering. */
static const int HEX_EDGES[12][2] = {{0, 1}, {0, 2}, {0, 4}, {1, 3}, {1, 5}, {2, 3},
                                     {2, 6}, {3, 7}, {4, 5}, {4, 6}, {5, 7}, {6, 7}};
static const int HEX_FACES[6][4] = {{0, 1, 2, 3}, {0, 1, 4, 5}, {0, 2, 4, 6},
                                    {1, 3, 5, 7}, {2, 3, 6, 7}, {4, 5, 6, 7}};

int wo_perm(int P, int* perm)
{
  const int n = P + 1, ni = P - 1;
  if (P < 1 || n > WO_MAXN) return -1;
  for (int ix = 0; ix < n; ++ix)
    for (int iy = 0; iy < n; ++iy)
      for (int iz = 0; iz < n; ++iz)
      {
        const int a[3] = {ix, iy, iz};
        int nint = 0;
        for (int d = 0; d < 3; ++d) nint += (a[d] >= 2);
        int dof = -1;
        if (nint == 0) dof = a[0] + 2 * a[1] + 4 * a[2];
        else if (nint == 1)
        {
          int v0 = 0, v1 = 0, pos = 0;
          for (int d = 0; d < 3; ++d)
          {
            if (a[d] >= 2) { v1 += (1 << d); pos = a[d] - 2; }
            else { v0 += a[d] << d; v1 += a[d] << d; }
          }
          for (int e = 0; e < 12; ++e)
            if (HEX_EDGES[e][0] == v0 && HEX_EDGES[e][1] == v1) dof = 8 + e * ni + pos;
        }
        else if (nint == 2)
        {
          int fixed = 0, side = 0, p[2], k = 0;
          for (int d = 0; d < 3; ++d)
          {
            if (a[d] >= 2) p[k++] = a[d] - 2;
            else { fixed = d; side = a[d]; }
          }
          /* lowest vertex of the face identifies it */
          int vmin = side << fixed, f = -1;
          for (int g = 0; g < 6; ++g)
          {
            int ok = 1;
            for (int m = 0; m < 4; ++m)
              if (((HEX_FACES[g][m] >> fixed) & 1) != side) ok = 0;
            if (ok && HEX_FACES[g][0] == vmin) f = g;
          }
          dof = 8 + 12 * ni + f * ni * ni + p[0] + ni * p[1];
        }
        else
          dof = 8 + 12 * ni + 6 * ni * ni + (a[0] - 2) + ni * (a[1] - 2) + ni * ni * (a[2] - 2);
        if (dof < 0) return -2;
        perm[(ix * n + iy) * n + iz] = dof;
      }
  return 0;
}

/* Dense derivative tables dphi[a][q][dof] in DOLFINx dof order, clamped:
 * tabulate_basis_and_permutation (common/operators.hpp:13-32), _dphi slice at :178. */
int wo_tabulate_dphi(int P, double* dphi)
{
  const int n = P + 1, nd = n * n * n;
  double D[WO_MAXN * WO_MAXN];
  int* perm = (int*)malloc(sizeof(int) * nd);
  if (wo_deriv_1d(P, D, 0) || wo_perm(P, perm)) { free(perm); return -1; }
  memset(dphi, 0, sizeof(double) * 3 * (size_t)nd * nd);
  for (int qa = 0; qa < n; ++qa)
    for (int qb = 0; qb < n; ++qb)
      for (int qc = 0; qc < n; ++qc)
      {
        const size_t q = (qa * n + qb) * n + qc;
        for (int i = 0; i < n; ++i)
        {
          /* d/dx: varies along first tensor index, delta in the others */
          dphi[(0 * (size_t)nd + q) * nd + perm[(i * n + qb) * n + qc]] = wo_clamp(D[qa * n + i]);
          dphi[(1 * (size_t)nd + q) * nd + perm[(qa * n + i) * n + qc]] = wo_clamp(D[qb * n + i]);
          dphi[(2 * (size_t)nd + q) * nd + perm[(qa * n + qb) * n + i]] = wo_clamp(D[qc * n + i]);
        }
      }
  free(perm);
  return 0;
}

/* reorder_dofmap (common/permute.hpp:10-28): out[c*nd+t] = in[c*nd+perm[t]] */
void wo_reorder_dofmap(int P, int64_t ncells, const int32_t* in, int32_t* out)
{
  const int n = P + 1, nd = n * n * n;
  int* perm = (int*)malloc(sizeof(int) * nd);
  wo_perm(P, perm);
  for (int64_t c = 0; c < ncells; ++c)
    for (int t = 0; t < nd; ++t) out[c * nd + t] = in[c * nd + perm[t]];
  free(perm);
}

/* ------------------------------------------------------------------------- */
/* dolfinx::math helpers [recalled: dolfinx/common/math.h, early-2022 main]. */
static inline double diffprod(double a, double b, double c, double d)
{
  double w = b * c;
  double err = fma(-b, c, w);
  double diff = fma(a, d, -w);
  return diff + err;
}
static inline double det3(const double A[3][3])
{
  double w0 = diffprod(A[1][1], A[1][2], A[2][1], A[2][2]);
  double w1 = diffprod(A[1][0], A[1][2], A[2][0], A[2][2]);
  double w2 = diffprod(A[1][0], A[1][1], A[2][0], A[2][1]);
  double w3 = diffprod(A[0][0], A[0][1], w1, w0);
  return fma(A[0][2], w2, w3);
}
static inline void inv3(const double A[3][3], double B[3][3])
{
  double w0 = diffprod(A[1][1], A[1][2], A[2][1], A[2][2]);
  double w1 = diffprod(A[1][0], A[1][2], A[2][0], A[2][2]);
  double w2 = diffprod(A[1][0], A[1][1], A[2][0], A[2][1]);
  double det = diffprod(A[0][0], A[0][1], w1, w0);
  det = fma(A[0][2], w2, det);
  det = 1.0 / det;
  B[0][0] = w0 * det;
  B[1][0] = -w1 * det;
  B[2][0] = w2 * det;
  B[0][1] = diffprod(A[0][2], A[0][1], A[2][2], A[2][1]) * det;
  B[0][2] = diffprod(A[0][1], A[0][2], A[1][1], A[1][2]) * det;
  B[1][1] = diffprod(A[0][0], A[0][2], A[2][0], A[2][2]) * det;
  B[1][2] = diffprod(A[1][0], A[0][0], A[1][2], A[0][2]) * det;
  B[2][1] = diffprod(A[2][0], A[0][0], A[2][1], A[0][1]) * det;
  B[2][2] = diffprod(A[0][0], A[1][0], A[0][1], A[1][1]) * det;
}

/* P1 hexahedron coordinate-element derivatives at reference point X, vertex
 * v = vx + 2 vy + 4 vz  [recalled: cmap.tabulate(1, points), precomputation.hpp:54-59] */
static inline void cmap_dphi(const double X[3], double d[3][8], int clamp)
{
  for (int v = 0; v < 8; ++v)
  {
    double f[3], s[3];
    for (int a = 0; a < 3; ++a)
    {
      int b = (v >> a) & 1;
      f[a] = b ? X[a] : 1.0 - X[a];
      s[a] = b ? 1.0 : -1.0;
    }
    d[0][v] = s[0] * f[1] * f[2];
    d[1][v] = f[0] * s[1] * f[2];
    d[2][v] = f[0] * f[1] * s[2];
    if (clamp)
      for (int a = 0; a < 3; ++a) d[a][v] = wo_clamp(d[a][v]);
  }
}

/* precompute_geometric_data (common/precomputation.hpp:18-110).
 * x: [npts][3], xdofs: [ncells][8].  G: [ncells][nq][3][3], detJ: [ncells][nq]. */
int wo_precompute_geometric_data(int P, int64_t ncells, const double* x, const int32_t* xdofs,
                                 double* G, double* detJ)
{
  const int n = P + 1, nq = n * n * n;
  double pts[WO_MAXN], wts[WO_MAXN];
  if (wo_gll(P, pts, wts)) return -1;
  double(*dphi)[3][8] = malloc(sizeof(double[3][8]) * nq);
  double* w = malloc(sizeof(double) * nq);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      for (int k = 0; k < n; ++k)
      {
        const int q = (i * n + j) * n + k;
        const double X[3] = {pts[i], pts[j], pts[k]};
        cmap_dphi(X, dphi[q], 1); /* :54-59 tabulate + clamp */
        w[q] = wts[i] * wts[j] * wts[k];
      }
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncells; ++c)
  {
    double coords[8][3];
    for (int v = 0; v < 8; ++v)
      for (int a = 0; a < 3; ++a) coords[v][a] = x[3 * (int64_t)xdofs[8 * c + v] + a];
    for (int q = 0; q < nq; ++q)
    {
      double J[3][3], K[3][3];
      for (int i = 0; i < 3; ++i) /* :83-91 */
        for (int j = 0; j < 3; ++j)
        {
          double s = 0.0;
          for (int v = 0; v < 8; ++v) s += coords[v][i] * dphi[q][j][v];
          J[i][j] = s;
        }
      const double dj = fabs(det3(J)) * w[q]; /* :95 */
      detJ[c * nq + q] = dj;
      inv3(J, K); /* :96 */
      double* g = G + (c * nq + q) * 9;
      for (int i = 0; i < 3; ++i) /* :99-100: dot(J_inv*detJ, J_inv^T, G) accumulating from 0 */
        for (int j = 0; j < 3; ++j)
        {
          double s = 0.0;
          for (int k = 0; k < 3; ++k) s += (K[i][k] * dj) * K[j][k];
          g[3 * i + j] = wo_clamp(s); /* :105-107 */
        }
    }
  }
  free(dphi);
  free(w);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* General-point variants (common/precompute.hpp): no fabs, no clamp, weights
 * applied by the caller.  points: [nq][3].  J out: [ncells][nq][3][3]. */
void wo_compute_jacobian(int64_t ncells, int nq, const double* points, const double* x,
                         const int32_t* xdofs, double* J) /* precompute.hpp:49-96 */
{
#pragma omp parallel for schedule(static)
  for (int64_t c = 0; c < ncells; ++c)
  {
    double coords[8][3];
    for (int v = 0; v < 8; ++v)
      for (int a = 0; a < 3; ++a) coords[v][a] = x[3 * (int64_t)xdofs[8 * c + v] + a];
    for (int q = 0; q < nq; ++q)
    {
      double d[3][8];
      cmap_dphi(points + 3 * q, d, 0);
      double* Jq = J + (c * nq + q) * 9;
      for (int i = 0; i < 3; ++i) /* dot(coords, dphi_q, J, transpose=true), :18-41 */
        for (int j = 0; j < 3; ++j)
        {
          double s = 0.0;
          for (int k = 0; k < 8; ++k) s += coords[k][i] * d[j][k];
          Jq[3 * i + j] = s;
        }
    }
  }
}
void wo_compute_jacobian_determinant(int64_t n, const double* J, double* detJ) /* :102-116 */
{
  for (int64_t i = 0; i < n; ++i) detJ[i] = det3((const double(*)[3])(J + 9 * i));
}
void wo_compute_jacobian_inverse(int64_t n, const double* J, double* K) /* :122-143 */
{
  for (int64_t i = 0; i < n; ++i) inv3((const double(*)[3])(J + 9 * i), (double(*)[3])(K + 9 * i));
}
/* compute_geometrical_factor (:148-176): G = K K^T * (detJ * w_q) */
void wo_compute_geometrical_factor(int64_t ncells, int nq, const double* J, const double* detJ,
                                   const double* weights, double* G)
{
  for (int64_t c = 0; c < ncells; ++c)
    for (int q = 0; q < nq; ++q)
    {
      double K[3][3];
      const double dj = detJ[c * nq + q] * weights[q];
      inv3((const double(*)[3])(J + (c * nq + q) * 9), K);
      double* g = G + (c * nq + q) * 9;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
        {
          double s = 0.0;
          for (int k = 0; k < 3; ++k) s += K[i][k] * K[j][k]; /* dot(K, KT, G) */
          g[3 * i + j] = s * dj;
        }
    }
}

/* Gauss-Legendre points on [0,1], m points  [recalled: basix gauss_jacobi with
 * alpha=0 on the interval, m=(q+2)/2; used by tabulate_1d precompute.hpp:179-189] */
int wo_gauss_legendre(int m, double* pts, double* wts)
{
  const long double pi = 3.14159265358979323846264338327950288L;
  if (m < 1 || m > 64) return -1;
  for (int i = 0; i < m; ++i)
  {
    long double xi = -cosl(pi * (i + 0.75L) / (m + 0.5L)), p, dp;
    for (int it = 0; it < 100; ++it)
    {
      legendre(m, xi, &p, &dp);
      long double dx = p / dp;
      xi -= dx;
      if (fabsl(dx) < 1e-19L) break;
    }
    legendre(m, xi, &p, &dp);
    pts[i] = (double)(0.5L * (xi + 1.0L));
    wts[i] = (double)(1.0L / ((1.0L - xi * xi) * dp * dp));
  }
  return 0;
}

/* tabulate_1d (precompute.hpp:179-189): table[q][i] = d^k/dx^k l_i(x_q), k=0,1,
 * l_i the GLL-warped Lagrange basis in [0,1,interior] ordering, at nq given points. */
int wo_tabulate_1d(int P, int nq, const double* points, int derivative, double* table)
{
  long double x[WO_MAXN], w[WO_MAXN];
  const int n = P + 1;
  if (P < 1 || n > WO_MAXN || derivative < 0 || derivative > 1) return -1;
  gll_ascending(P, x, w);
  for (int q = 0; q < nq; ++q)
    for (int a = 0; a < n; ++a)
    {
      const int i = asc_of(a, P);
      const long double xq = points[q];
      long double val = 0.0L;
      if (derivative == 0)
      {
        val = 1.0L;
        for (int m = 0; m < n; ++m)
          if (m != i) val *= (xq - x[m]) / (x[i] - x[m]);
      }
      else
      {
        for (int r = 0; r < n; ++r)
        {
          if (r == i) continue;
          long double t = 1.0L / (x[i] - x[r]);
          for (int m = 0; m < n; ++m)
            if (m != i && m != r) t *= (xq - x[m]) / (x[i] - x[m]);
          val += t;
        }
      }
      table[q * n + a] = (double)val;
    }
  return 0;
}

/* ------------------------------------------------------------------------- */
/* MassOperatorCPU::operator() + mkernel (common/operators.hpp:36-40,86-108):
 *   y[dof(c,perm[t])] += x[dof(c,perm[t])] * detJ[c,t]   */
void wo_mass_apply(int P, int64_t ncells, const int32_t* dofmap, const double* detJ,
                   const double* x, double* y)
{
  const int n = P + 1, nd = n * n * n;
  int* perm = (int*)malloc(sizeof(int) * nd);
  double *_x = malloc(sizeof(double) * nd), *_y = malloc(sizeof(double) * nd);
  wo_perm(P, perm);
  for (int64_t c = 0; c < ncells; ++c)
  {
    const int32_t* cell_dofs = dofmap + c * nd;
    for (int t = 0; t < nd; ++t) _x[t] = x[cell_dofs[perm[t]]];
    for (int t = 0; t < nd; ++t) _y[t] = 0.0;
    for (int iq = 0; iq < nd; ++iq) _y[iq] = _x[iq] * detJ[c * nd + iq]; /* mkernel */
    for (int t = 0; t < nd; ++t) y[cell_dofs[perm[t]]] += _y[t];
  }
  free(perm);
  free(_x);
  free(_y);
}

/* skernel (common/operators.hpp:113-133): dense, c0 = 1500 hard-coded. */
static inline void skernel(double* A, const double* w, const double* G, const double* dphi,
                           int nq, int nd)
{
  const double c0 = 1500.0;
  const double coeff = -1.0 * c0 * c0;
  const double *d0 = dphi, *d1 = dphi + (size_t)nq * nd, *d2 = dphi + 2 * (size_t)nq * nd;
  for (int iq = 0; iq < nq; iq++)
  {
    const double* _G = G + iq * 9;
    double w0 = 0.0, w1 = 0.0, w2 = 0.0;
    for (int ic = 0; ic < nd; ic++)
    {
      w0 += w[ic] * d0[(size_t)iq * nd + ic];
      w1 += w[ic] * d1[(size_t)iq * nd + ic];
      w2 += w[ic] * d2[(size_t)iq * nd + ic];
    }
    const double fw0 = coeff * (_G[0] * w0 + _G[1] * w1 + _G[2] * w2);
    const double fw1 = coeff * (_G[3] * w0 + _G[4] * w1 + _G[5] * w2);
    const double fw2 = coeff * (_G[6] * w0 + _G[7] * w1 + _G[8] * w2);
    for (int i = 0; i < nd; i++)
      A[i] += fw0 * d0[(size_t)iq * nd + i] + fw1 * d1[(size_t)iq * nd + i]
              + fw2 * d2[(size_t)iq * nd + i];
  }
}

/* StiffnessOperator::operator() (common/operators.hpp:183-200): y += -c0^2 K x.
 * nthreads <= 1: the reference's serial cell loop.  nthreads > 1: OpenMP over
 * cells with thread-private y (a stand-in for one MPI rank per core). */
void wo_stiffness_apply_dense(int P, int64_t ncells, int64_t ndofs, const int32_t* dofmap,
                              const double* G, const double* x, double* y, int nthreads)
{
  const int n = P + 1, nd = n * n * n, nq = nd;
  double* dphi = malloc(sizeof(double) * 3 * (size_t)nq * nd);
  wo_tabulate_dphi(P, dphi);
  if (nthreads <= 1)
  {
    double *_x = malloc(sizeof(double) * nd), *_y = malloc(sizeof(double) * nd);
    for (int64_t c = 0; c < ncells; ++c)
    {
      const int32_t* cell_dofs = dofmap + c * nd;
      for (int i = 0; i < nd; ++i) _x[i] = x[cell_dofs[i]];
      for (int i = 0; i < nd; ++i) _y[i] = 0.0;
      skernel(_y, _x, G + c * nq * 9, dphi, nq, nd);
      for (int i = 0; i < nd; ++i) y[cell_dofs[i]] += _y[i];
    }
    free(_x);
    free(_y);
  }
  else
  {
#ifdef _OPENMP
    omp_set_num_threads(nthreads);
#endif
    double* ypriv = calloc((size_t)nthreads * ndofs, sizeof(double));
#pragma omp parallel
    {
      int tid = 0;
#ifdef _OPENMP
      tid = omp_get_thread_num();
#endif
      double* yt = ypriv + (size_t)tid * ndofs;
      double *_x = malloc(sizeof(double) * nd), *_y = malloc(sizeof(double) * nd);
#pragma omp for schedule(static)
      for (int64_t c = 0; c < ncells; ++c)
      {
        const int32_t* cell_dofs = dofmap + c * nd;
        for (int i = 0; i < nd; ++i) _x[i] = x[cell_dofs[i]];
        for (int i = 0; i < nd; ++i) _y[i] = 0.0;
        skernel(_y, _x, G + c * nq * 9, dphi, nq, nd);
        for (int i = 0; i < nd; ++i) yt[cell_dofs[i]] += _y[i];
      }
      free(_x);
      free(_y);
#pragma omp for schedule(static)
      for (int64_t i = 0; i < ndofs; ++i)
      {
        double s = y[i];
        for (int t = 0; t < nthreads; ++t) s += ypriv[(size_t)t * ndofs + i];
        y[i] = s;
      }
    }
    free(ypriv);
  }
  free(dphi);
}

/* The same operator in sum-factorised form -- NOT how the
 * reference computes it; equal to skernel up to summation order.  Used to check
 * the oracle against itself and as a fast checker at sizes where the dense
 * form takes minutes.  c0 = 1500 as in skernel. */
void wo_stiffness_apply_sumfact(int P, int64_t ncells, int64_t ndofs, const int32_t* dofmap,
                                const double* G, const double* x, double* y, int nthreads)
{
  const int n = P + 1, nd = n * n * n;
  double D[WO_MAXN * WO_MAXN];
  int* perm = (int*)malloc(sizeof(int) * nd);
  wo_deriv_1d(P, D, 1);
  wo_perm(P, perm);
  if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
  omp_set_num_threads(nthreads);
#endif
  double* ypriv = nthreads > 1 ? calloc((size_t)nthreads * ndofs, sizeof(double)) : NULL;
#pragma omp parallel
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    double* yt = nthreads > 1 ? ypriv + (size_t)tid * ndofs : y;
    double* xt = malloc(sizeof(double) * nd * 5);
    double *f0 = xt + nd, *f1 = xt + 2 * nd, *f2 = xt + 3 * nd, *yl = xt + 4 * nd;
    const double coeff = -1.0 * 1500.0 * 1500.0;
#pragma omp for schedule(static)
    for (int64_t c = 0; c < ncells; ++c)
    {
      const int32_t* cd = dofmap + c * nd;
      for (int t = 0; t < nd; ++t) xt[t] = x[cd[perm[t]]];
      for (int a = 0; a < n; ++a)
        for (int b = 0; b < n; ++b)
          for (int k = 0; k < n; ++k)
          {
            const int q = (a * n + b) * n + k;
            double w0 = 0, w1 = 0, w2 = 0;
            for (int m = 0; m < n; ++m)
            {
              w0 += D[a * n + m] * xt[(m * n + b) * n + k];
              w1 += D[b * n + m] * xt[(a * n + m) * n + k];
              w2 += D[k * n + m] * xt[(a * n + b) * n + m];
            }
            const double* g = G + (c * nd + q) * 9;
            f0[q] = coeff * (g[0] * w0 + g[1] * w1 + g[2] * w2);
            f1[q] = coeff * (g[3] * w0 + g[4] * w1 + g[5] * w2);
            f2[q] = coeff * (g[6] * w0 + g[7] * w1 + g[8] * w2);
          }
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
          for (int k = 0; k < n; ++k)
          {
            double s = 0;
            for (int m = 0; m < n; ++m)
              s += D[m * n + i] * f0[(m * n + j) * n + k] + D[m * n + j] * f1[(i * n + m) * n + k]
                   + D[m * n + k] * f2[(i * n + j) * n + m];
            yl[(i * n + j) * n + k] = s;
          }
      for (int t = 0; t < nd; ++t) yt[cd[perm[t]]] += yl[t];
    }
    free(xt);
    if (nthreads > 1)
    {
#pragma omp for schedule(static)
      for (int64_t i = 0; i < ndofs; ++i)
      {
        double s = y[i];
        for (int t = 0; t < nthreads; ++t) s += ypriv[(size_t)t * ndofs + i];
        y[i] = s;
      }
    }
  }
  free(ypriv);
  free(perm);
}

/* ------------------------------------------------------------------------- */
/* Boundary linear form L (demo/cpu_planar3d/forms.ufl:21-24, assembled at
 * common/LinearGLL.hpp:175).  With the GLL facet rule (degree 6 at P4 => the
 * facet's own nodes) the form is diagonal:
 *    b_i += c0^2 * g * m1_i  -  c0 * m2_i * v_n[i]
 * with m_tag,i = sum over facets with that tag containing node i of
 * w_a w_b |dx/ds x dx/dt|.  [recalled: FFCx-generated exterior-facet kernel;
 * SURVEY.md App. A.11].  This routine accumulates m1 (tag 1) and m2 (tag 2). */
int wo_boundary_facet_mass(int P, int64_t nfacets, const int32_t* fcell, const int32_t* flocal,
                           const int32_t* ftag, const double* x, const int32_t* xdofs,
                           const int32_t* dofmap, double* m1, double* m2)
{
  const int n = P + 1, nd = n * n * n;
  double pts[WO_MAXN], wts[WO_MAXN];
  static const int F_AXIS[6] = {2, 1, 0, 0, 1, 2};
  static const int F_SIDE[6] = {0, 0, 0, 1, 1, 1};
  int* perm = (int*)malloc(sizeof(int) * nd);
  if (wo_gll(P, pts, wts) || wo_perm(P, perm)) { free(perm); return -1; }
  for (int64_t f = 0; f < nfacets; ++f)
  {
    const int64_t c = fcell[f];
    const int lf = flocal[f], tag = ftag[f];
    double* m = tag == 1 ? m1 : (tag == 2 ? m2 : NULL);
    if (!m) continue;
    const int ax = F_AXIS[lf], side = F_SIDE[lf];
    const int t1 = ax == 0 ? 1 : 0, t2 = ax == 2 ? 1 : 2; /* tangential axes, ascending */
    double coords[8][3];
    for (int v = 0; v < 8; ++v)
      for (int a = 0; a < 3; ++a) coords[v][a] = x[3 * (int64_t)xdofs[8 * c + v] + a];
    for (int ia = 0; ia < n; ++ia)
      for (int ib = 0; ib < n; ++ib)
      {
        int idx[3];
        idx[ax] = side; /* 1-D index 0 = lower end, 1 = upper end */
        idx[t1] = ia;
        idx[t2] = ib;
        const double X[3] = {pts[idx[0]], pts[idx[1]], pts[idx[2]]};
        double d[3][8], J[3][3];
        cmap_dphi(X, d, 1);
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j)
          {
            double s = 0.0;
            for (int v = 0; v < 8; ++v) s += coords[v][i] * d[j][v];
            J[i][j] = s;
          }
        const double a0 = J[0][t1], a1 = J[1][t1], a2 = J[2][t1];
        const double b0 = J[0][t2], b1 = J[1][t2], b2 = J[2][t2];
        const double cx = a1 * b2 - a2 * b1, cy = a2 * b0 - a0 * b2, cz = a0 * b1 - a1 * b0;
        const double scale = sqrt(cx * cx + cy * cy + cz * cz);
        const int t = (idx[0] * n + idx[1]) * n + idx[2];
        m[dofmap[c * nd + perm[t]]] += wts[ia] * wts[ib] * scale;
      }
  }
  free(perm);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* LinearGLLOpt::f1 and ::rk4 (common/LinearGLL.hpp:151-192,198-287), serial
 * (one rank: no ghosts, scatter_fwd / scatter_rev are no-ops).
 *   m    lumped mass = MassOperatorCPU(1)  (:102-110)
 *   m1,m2 boundary facet masses for tags 1 and 2 (forms.ufl:21-24)
 * u_n, v_n: in = initial state (init(): zeros, :131-134), out = final state.
 * sumfact != 0 swaps the dense skernel for the sum-factorised form (fast checker).
 * Returns the number of steps taken; *t_end receives the final time. */
int64_t wo_rk4(int P, int64_t ncells, int64_t ndofs, const int32_t* dofmap, const double* G,
               const double* m, const double* m1, const double* m2, double c0, double freq0,
               double p0, double t0, double tf, double dt, int64_t max_steps, double* u_n,
               double* v_n, int sumfact, int nthreads, double* t_end)
{
  const double w0 = 2.0 * M_PI * freq0, T = 1.0 / freq0, alpha = 4.0; /* :96-99 */
  const double a_runge[4] = {0.0, 0.5, 0.5, 1.0};
  const double b_runge[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
  const double c_runge[4] = {0.0, 0.5, 0.5, 1.0};
  const size_t nb = sizeof(double) * ndofs;
  double *u_ = malloc(nb), *v_ = malloc(nb), *un = malloc(nb), *vn = malloc(nb);
  double *u0 = malloc(nb), *v0 = malloc(nb), *ku = malloc(nb), *kv = malloc(nb), *b = malloc(nb);
  memcpy(u_, u_n, nb); /* :213-214 */
  memcpy(v_, v_n, nb);
  memcpy(ku, u_, nb); /* :229-230 */
  memcpy(kv, v_, nb);
  double t = t0;
  int64_t step = 0;
  while (t < tf) /* :241 */
  {
    if (max_steps > 0 && step >= max_steps) break;
    dt = dt < tf - t ? dt : tf - t; /* :242 */
    memcpy(u0, u_, nb);
    memcpy(v0, v_, nb);
    for (int i = 0; i < 4; ++i)
    {
      memcpy(un, u0, nb); /* :250-251 */
      memcpy(vn, v0, nb);
      const double adt = dt * a_runge[i];
      for (int64_t k = 0; k < ndofs; ++k) un[k] = ku[k] * adt + un[k]; /* :253 axpy */
      for (int64_t k = 0; k < ndofs; ++k) vn[k] = kv[k] * adt + vn[k]; /* :254 */
      const double tn = t + c_runge[i] * dt;                           /* :257 */
      memcpy(ku, vn, nb);                                              /* f0 :141-144 */
      /* f1 :151-192 */
      double window = tn < T * alpha ? 0.5 * (1.0 - cos(freq0 * M_PI * tn / alpha)) : 1.0;
      const double g = window * p0 * w0 / c0 * cos(w0 * tn); /* :162 */
      memset(b, 0, nb);                                       /* :173 */
      if (sumfact) wo_stiffness_apply_sumfact(P, ncells, ndofs, dofmap, G, un, b, nthreads);
      else wo_stiffness_apply_dense(P, ncells, ndofs, dofmap, G, un, b, nthreads); /* :174 */
      for (int64_t k = 0; k < ndofs; ++k) /* :175 assemble_vector(b, L) */
        b[k] += c0 * c0 * g * m1[k] - c0 * m2[k] * vn[k];
      for (int64_t k = 0; k < ndofs; ++k) kv[k] = b[k] / m[k]; /* :188-191 */
      const double bdt = dt * b_runge[i];
      for (int64_t k = 0; k < ndofs; ++k) u_[k] = ku[k] * bdt + u_[k]; /* :264 */
      for (int64_t k = 0; k < ndofs; ++k) v_[k] = kv[k] * bdt + v_[k]; /* :265 */
    }
    t += dt;
    step += 1;
  }
  memcpy(u_n, u_, nb); /* :282-283 */
  memcpy(v_n, v_, nb);
  if (t_end) *t_end = t;
  free(u_); free(v_); free(un); free(vn); free(u0); free(v0); free(ku); free(kv); free(b);
  return step;
}

int wo_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
