// C entry points over the REFERENCE's own CPU code of the hot path, for oracle/_ref/libwfref_cpu.so.
// TEST INFRASTRUCTURE ONLY.
//
// The reference's headers cannot be compiled as a whole (they include DOLFINx, Basix, xtensor and the
// FFCx-generated forms.h, all absent here).  What needs none of their arithmetic is cut out of the
// headers WHERE THEY LIE under /root/reference by oracle/build_ref.py at build time (into
// oracle/_ref/*.inc, deleted again after the compile; nothing of it is committed) and compiled here
// against small stand-ins for the container types it touches:
//   common/operators.hpp   mkernel :36-40, skernel :113-133 (the cell kernels),
//                          MassOperatorCPU::operator() :85-108, StiffnessOperator::operator() :182-200
//   common/LinearGLL.hpp   kernels::copy / axpy :15-35, LinearGLLOpt::init / f0 / f1 / rk4 :130-287
// so the gather / kernel / scatter loops, the right-hand side f1 (window, source amplitude, order of
// operations, b / m) and the whole RK4 loop (tableau, stage algebra, axpy on owned entries, final
// copies) that the tests compare the oracle with ARE the reference's.  Supplied from outside, because
// they come from the un-vendored dependencies: the geometric factors G / detJ, the basis table dphi,
// the permutation, and the boundary form `fem::assemble_vector(_b, *L)` (FFCx kernel), which is
// restated below as the diagonal GLL facet form of forms.ufl:21-24.  One rank: the scatters are no-ops.
// This file itself contains no reference code.
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <iostream>
#include <iterator>
#include <map>
#include <memory>
#include <string>
#include <vector>

// ---- stand-ins for the container types the cut code touches ------------------------------------
namespace xtl
{
template <typename T>
class span
{
public:
  span() = default;
  span(T* p, std::size_t n) : _p(p), _n(n) {}
  template <typename U>
  span(const span<U>& o) : _p(o.data()), _n(o.size()) {}
  T* begin() const { return _p; }
  T* end() const { return _p + _n; }
  const T* cbegin() const { return _p; }
  const T* cend() const { return _p + _n; }
  T& operator[](std::size_t i) const { return _p[i]; }
  T* data() const { return _p; }
  std::size_t size() const { return _n; }

private:
  T* _p = nullptr;
  std::size_t _n = 0;
};
} // namespace xtl
namespace tcb
{
using xtl::span;
}

namespace xt
{
// row-major array view: data(), shape(i), (a, q, i) call operator -- what the cut code uses of xt::xtensor
template <typename T, std::size_t R>
class xtensor
{
public:
  xtensor() = default;
  xtensor(T* data, std::initializer_list<std::size_t> shape) : _d(data)
  {
    std::size_t k = 0;
    for (std::size_t s : shape) _s[k++] = s;
  }
  T* data() { return _d; }
  const T* data() const { return _d; }
  std::size_t shape(std::size_t i) const { return _s[i]; }
  const T& operator()(std::size_t a, std::size_t q, std::size_t i) const { return _d[(a * _s[1] + q) * _s[2] + i]; }

private:
  T* _d = nullptr;
  std::size_t _s[4] = {0, 0, 0, 0};
};
template <typename T>
class xarray
{
public:
  xarray(std::initializer_list<T> v) : _v(v) {}
  const T& operator()(std::size_t i) const { return _v[i]; }

private:
  std::vector<T> _v;
};
} // namespace xt

namespace standin
{
namespace common
{
class IndexMap
{
public:
  enum class Mode { insert, add };
  explicit IndexMap(std::int32_t n) : _n(n) {}
  std::int32_t size_local() const { return _n; }
  std::int32_t num_ghosts() const { return 0; }

private:
  std::int32_t _n;
};
} // namespace common
namespace la
{
template <typename T, typename Alloc = std::allocator<T>>
class Vector
{
public:
  Vector(std::shared_ptr<const common::IndexMap> map, int bs)
      : _map(map), _x((std::size_t)map->size_local() * bs, T(0)), _p(_x.data()), _n(_x.size())
  {
  }
  // a view of the caller's array (no copy)
  Vector(std::shared_ptr<const common::IndexMap> map, int bs, T* external)
      : _map(map), _p(external), _n((std::size_t)map->size_local() * bs)
  {
  }
  Vector(const Vector&) = delete;
  Vector& operator=(const Vector&) = delete;
  xtl::span<const T> array() const { return xtl::span<const T>(_p, _n); }
  xtl::span<T> mutable_array() { return xtl::span<T>(_p, _n); }
  std::shared_ptr<const common::IndexMap> map() const { return _map; }
  void set(T v) { std::fill(_p, _p + _n, v); }
  void scatter_fwd() {}                     // one rank
  void scatter_rev(common::IndexMap::Mode) {} // one rank

private:
  std::shared_ptr<const common::IndexMap> _map;
  std::vector<T> _x;
  T* _p = nullptr;
  std::size_t _n = 0;
};
} // namespace la
namespace graph
{
template <typename T>
class AdjacencyList
{
public:
  AdjacencyList() = default;
  AdjacencyList(const T* data, std::int32_t nlinks) : _d(data), _k(nlinks) {}
  xtl::span<const T> links(std::int32_t node) const { return xtl::span<const T>(_d + (std::int64_t)node * _k, (std::size_t)_k); }

private:
  const T* _d = nullptr;
  std::int32_t _k = 0;
};
} // namespace graph
namespace fem
{
template <typename T>
class Function
{
public:
  explicit Function(std::shared_ptr<const common::IndexMap> map) : _x(std::make_shared<la::Vector<T>>(map, 1)) {}
  std::shared_ptr<la::Vector<T>> x() { return _x; }

private:
  std::shared_ptr<la::Vector<T>> _x;
};
// The boundary linear form L of demo/cpu_planar3d/forms.ufl:21-24 under its GLL facet rule (degree 6 at
// P4: collocated, hence diagonal):  b_i += c0^2 g_i m1_i - c0 vn_i m2_i  with the facet masses m1 (tag 1)
// and m2 (tag 2).  FFCx generates this kernel in the reference: restated, like in the oracle.
template <typename T>
struct Form
{
  const double* m1 = nullptr;
  const double* m2 = nullptr;
  double c0 = 0;
  std::shared_ptr<Function<T>> g, v_n;
};
template <typename T>
void assemble_vector(xtl::span<T> b, const Form<T>& L)
{
  const xtl::span<const T> g = L.g->x()->array(), vn = L.v_n->x()->array();
  for (std::size_t i = 0; i < b.size(); ++i)
    b[i] += L.c0 * L.c0 * g[i] * L.m1[i] - L.c0 * L.m2[i] * vn[i]; // (association as in oracle/wave_oracle.c)
}
} // namespace fem
} // namespace standin

// ---- the reference's code, in the surroundings it expects ----------------------------------------
namespace reference
{
using namespace standin;

#include "ref_cell_kernels.inc" // mkernel, skernel

template <typename T>
class MassOperatorCPU
{
private:
  std::vector<T> _x, _y;
  std::int32_t _ncells, _ndofs;
  graph::AdjacencyList<std::int32_t> _dofmap;
  xt::xtensor<double, 2> _detJ, _phi;
  std::vector<int> _perm;

public:
  MassOperatorCPU(std::int32_t ncells, int nd, const std::int32_t* dofmap, double* detJ, const int* perm)
      : _x(nd), _y(nd), _ncells(ncells), _ndofs(nd), _dofmap(dofmap, nd),
        _detJ(detJ, {(std::size_t)ncells, (std::size_t)nd}), _perm(perm, perm + nd)
  {
  }
#include "ref_mass_call.inc" // MassOperatorCPU::operator()
};

template <typename T>
class StiffnessOperator
{
private:
  std::vector<T> _x, _y;
  std::int32_t _ncells, _ndofs;
  graph::AdjacencyList<std::int32_t> _dofmap;
  xt::xtensor<double, 4> G;
  xt::xtensor<double, 2> _detJ;
  xt::xtensor<double, 3> _dphi;
  std::map<std::string, double> _params;

public:
  StiffnessOperator(std::int32_t ncells, int nd, const std::int32_t* dofmap, double* G9, double* dphi)
      : _x(nd), _y(nd), _ncells(ncells), _ndofs(nd), _dofmap(dofmap, nd),
        G(G9, {(std::size_t)ncells, (std::size_t)nd, 3, 3}), _detJ(nullptr, {(std::size_t)ncells, (std::size_t)nd}),
        _dphi(dphi, {3, (std::size_t)nd, (std::size_t)nd})
  {
  }
#include "ref_stiffness_call.inc" // StiffnessOperator::operator()
};

#include "ref_wave_kernels.inc" // namespace kernels { copy, axpy }

class LinearGLLOpt
{
private:
  int rank = 0, size = 1;

protected:
  double c0_, freq0_, p0_, w0_, T_, alpha_, window_ = 0;
  std::shared_ptr<fem::Form<double>> L;
  std::shared_ptr<fem::Function<double>> g, u_n, v_n;
  std::shared_ptr<la::Vector<double>> m, b;
  xtl::span<double> _g, out;
  xtl::span<const double> m_, b_;
  tcb::span<double> _b;
  std::shared_ptr<const common::IndexMap> index_map;
  int bs = 1;
  std::shared_ptr<StiffnessOperator<double>> stiff_op;

public:
  // what the reference's constructor (LinearGLL.hpp:69-128) sets up, from arrays
  LinearGLLOpt(std::int32_t ncells, std::int32_t ndofs, int nd, const std::int32_t* dofmap, double* G9, double* dphi,
               const double* mvec, const double* m1, const double* m2, double speedOfSound, double sourceFrequency,
               double pressureAmplitude)
  {
    index_map = std::make_shared<const common::IndexMap>(ndofs);
    g = std::make_shared<fem::Function<double>>(index_map);
    u_n = std::make_shared<fem::Function<double>>(index_map);
    v_n = std::make_shared<fem::Function<double>>(index_map);
    _g = g->x()->mutable_array();
    c0_ = speedOfSound;
    freq0_ = sourceFrequency;
    p0_ = pressureAmplitude;
    w0_ = 2.0 * M_PI * freq0_;
    T_ = 1.0 / freq0_;
    alpha_ = 4.0;
    m = std::make_shared<la::Vector<double>>(index_map, bs);
    std::copy(mvec, mvec + ndofs, m->mutable_array().begin());
    L = std::make_shared<fem::Form<double>>();
    L->m1 = m1, L->m2 = m2, L->c0 = c0_, L->g = g, L->v_n = v_n;
    stiff_op = std::make_shared<StiffnessOperator<double>>(ncells, nd, dofmap, G9, dphi);
    b = std::make_shared<la::Vector<double>>(index_map, bs);
    _b = b->mutable_array();
  }
  void set_state(const double* u, const double* v)
  {
    std::copy(u, u + index_map->size_local(), u_n->x()->mutable_array().begin());
    std::copy(v, v + index_map->size_local(), v_n->x()->mutable_array().begin());
  }
  void get_state(double* u, double* v)
  {
    const auto a = u_n->x()->array(), c = v_n->x()->array();
    std::copy(a.begin(), a.end(), u);
    std::copy(c.begin(), c.end(), v);
  }
  std::shared_ptr<la::Vector<double>> make_vector() { return std::make_shared<la::Vector<double>>(index_map, bs); }
#include "ref_wave_methods.inc" // init, f0, f1, rk4
};
} // namespace reference

extern "C" {
// A[nd] += skernel(w[nd]; G[nq][3][3], dphi[3][nq][nd])   -- the call of operators.hpp:195
void wfref_skernel(double* A, const double* w, const double* G, const double* dphi, int nq, int nd)
{
  std::map<std::string, double> params; // ignored by the reference (c0 = 1500 is hard-coded, :114)
  const xt::xtensor<double, 3> t(const_cast<double*>(dphi), {3, (std::size_t)nq, (std::size_t)nd});
  reference::skernel<double>(A, w, params, G, t, nq, nd);
}
// A[nq] = mkernel(w[nq]; detJ[nq])                         -- the call of operators.hpp:101
void wfref_mkernel(double* A, const double* w, const double* detJ, int nq, int nd)
{
  reference::mkernel<double>(A, w, nullptr, detJ, nullptr, nq, nd);
}
// y += A x through the reference's StiffnessOperator::operator() / MassOperatorCPU::operator()
void wfref_stiffness_apply(int ncells, int ndofs, int nd, const std::int32_t* dofmap, double* G9, double* dphi,
                           const double* x, double* y)
{
  using namespace standin;
  auto map = std::make_shared<const common::IndexMap>(ndofs);
  la::Vector<double> vx(map, 1, const_cast<double*>(x)), vy(map, 1, y); // views: the operator reads x, adds into y
  reference::StiffnessOperator<double> op(ncells, nd, dofmap, G9, dphi);
  op(vx, vy);
}
void wfref_mass_apply(int ncells, int ndofs, int nd, const std::int32_t* dofmap, double* detJ, const int* perm,
                      const double* x, double* y)
{
  using namespace standin;
  auto map = std::make_shared<const common::IndexMap>(ndofs);
  la::Vector<double> vx(map, 1, const_cast<double*>(x)), vy(map, 1, y);
  reference::MassOperatorCPU<double> op(ncells, nd, dofmap, detJ, perm);
  op(vx, vy);
}
// the reference's LinearGLLOpt::rk4 from (u, v) at t0; dv/dt at (t, u, v) through its f1
void wfref_rk4(int ncells, int ndofs, int nd, const std::int32_t* dofmap, double* G9, double* dphi, const double* m,
               const double* m1, const double* m2, double c0, double f0, double p0, double t0, double tf, double dt,
               double* u, double* v)
{
  reference::LinearGLLOpt eq(ncells, ndofs, nd, dofmap, G9, dphi, m, m1, m2, c0, f0, p0);
  eq.set_state(u, v);
  eq.rk4(t0, tf, dt);
  eq.get_state(u, v);
}
void wfref_f1(int ncells, int ndofs, int nd, const std::int32_t* dofmap, double* G9, double* dphi, const double* m,
              const double* m1, const double* m2, double c0, double f0, double p0, double t, const double* u,
              const double* v, double* result)
{
  reference::LinearGLLOpt eq(ncells, ndofs, nd, dofmap, G9, dphi, m, m1, m2, c0, f0, p0);
  auto vu = eq.make_vector(), vv = eq.make_vector(), vr = eq.make_vector();
  std::copy(u, u + ndofs, vu->mutable_array().begin());
  std::copy(v, v + ndofs, vv->mutable_array().begin());
  eq.f1(t, vu, vv, vr);
  std::copy(vr->array().begin(), vr->array().end(), result);
}
int wfref_cpu_version() { return 2; }
}
