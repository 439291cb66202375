// C entry points over the REFERENCE's own partition arithmetic -- decompose3d and
// compute_cartesian_indices (demo/gpu_cg/mesh.hpp:37-62), which partition.py follows -- and its
// reorder_dofmap (common/permute.hpp:10-28), for
// oracle/_ref/libwfref_cpu.so's sibling libwfref_mesh.so.  TEST INFRASTRUCTURE ONLY.  The two functions
// are cut out of the header where it lies under /root/reference at build time (oracle/build_ref.py,
// oracle/_ref/ref_mesh_functions.inc, deleted again after the compile); this file supplies the
// xt::xtensor stand-in they need and contains no reference code.
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <functional>
#include <initializer_list>
#include <numeric>
#include <vector>

namespace xt
{
template <typename T>
struct zeros_expr
{
  std::vector<std::size_t> shape;
};
template <typename T, typename S>
zeros_expr<T> zeros(std::initializer_list<S> shape)
{
  zeros_expr<T> z;
  for (S s : shape) z.shape.push_back((std::size_t)s);
  return z;
}
// owning row-major array of rank R: [i], (i, j), begin / end -- what the cut functions use
template <typename T, std::size_t R>
class xtensor
{
public:
  xtensor() = default;
  template <typename U>
  xtensor(const zeros_expr<U>& z) : _shape(z.shape)
  {
    std::size_t n = 1;
    for (std::size_t s : _shape) n *= s;
    _d.assign(n, T(0));
  }
  T& operator[](std::size_t i) { return _d[i]; }
  const T& operator[](std::size_t i) const { return _d[i]; }
  T& operator()(std::size_t i, std::size_t j) { return _d[i * _shape[1] + j]; }
  const T& operator()(std::size_t i, std::size_t j) const { return _d[i * _shape[1] + j]; }
  typename std::vector<T>::iterator begin() { return _d.begin(); }
  typename std::vector<T>::iterator end() { return _d.end(); }
  std::size_t shape(std::size_t i) const { return _shape[i]; }

private:
  std::vector<std::size_t> _shape;
  std::vector<T> _d;
};
} // namespace xt

// stand-in for the Basix calls of reorder_dofmap (common/permute.hpp:10-28): the element only has to hand
// out the tensor-product permutation, which the caller supplies (it is Basix's arithmetic, absent here)
#include <tuple>
namespace basix
{
namespace element
{
enum class family { P };
enum class lagrange_variant { gll_warped };
} // namespace element
namespace cell
{
enum class type { hexahedron };
}
inline std::vector<int>& supplied_perm()
{
  static std::vector<int> p;
  return p;
}
struct FiniteElement
{
  std::vector<std::tuple<std::vector<int>, std::vector<int>>> get_tensor_product_representation() const
  {
    return {std::make_tuple(std::vector<int>(), supplied_perm())};
  }
};
inline FiniteElement create_element(element::family, cell::type, int, element::lagrange_variant) { return FiniteElement(); }
} // namespace basix

namespace reference
{
#include "ref_mesh_functions.inc" // decompose3d, compute_cartesian_indices, reorder_dofmap
} // namespace reference

extern "C" {
// 2^x ranks -> (2^x0, 2^x1, 2^x2)
void wfref_decompose3d(int x, int* out)
{
  const xt::xtensor<int, 1> n = reference::decompose3d(x);
  for (int a = 0; a < 3; ++a) out[a] = n[a];
}
// out = reorder_dofmap(in) with the given tensor-product permutation (common/permute.hpp:10-28)
void wfref_reorder_dofmap(int p, int nd, int ncells, const int* perm, const int* in, int* out)
{
  basix::supplied_perm().assign(perm, perm + nd);
  std::vector<int> vin(in, in + (std::size_t)ncells * nd), vout((std::size_t)ncells * nd, -1);
  reference::reorder_dofmap(vout, vin, p);
  for (std::size_t i = 0; i < vout.size(); ++i) out[i] = vout[i];
}
// the temporal parameters of the reference demo (demo/cpu_planar3d/main.cpp:59-66: CFL time step snapped to
// whole steps per period, final time), computed by the demo's own statements from the smallest mesh size
void wfref_demo_time_parameters(double meshSize, double speedOfSound, double sourceFrequency, double domainLength,
                                int degreeOfBasis, double* dt, double* tf, int* steps_per_period)
{
  double period = 1 / sourceFrequency; // (main.cpp:30)
  using std::pow;
#include "ref_demo_params.inc"
  (void)startTime;
  *dt = timeStepSize;
  *tf = finalTime;
  *steps_per_period = stepPerPeriod;
}
// C (3x3, accumulated into) = op(A) op(B) through the reference's dot (common/precompute.hpp:17-41), the
// small matrix product of compute_jacobian (transpose: A [nodes][3], B [3][nodes]) and of
// compute_geometrical_factor (A, B 3x3)
struct Mat
{
  double* d;
  int r, c;
  int shape(int i) const { return i == 0 ? r : c; }
  double& operator()(int i, int j) { return d[i * c + j]; }
  const double& operator()(int i, int j) const { return d[i * c + j]; }
};
void wfref_dot(double* A, int ar, int ac, double* B, int br, int bc, double* C, int transpose)
{
  Mat a{A, ar, ac}, b{B, br, bc}, c{C, 3, 3};
  reference::dot(a, b, c, transpose != 0);
}
// rank -> (Ix, Iy, Iz) for all ranks of a procs[0] x procs[1] x procs[2] grid; out [size][3]
void wfref_cartesian_indices(const int* procs, long long* out)
{
  xt::xtensor<int, 1> p = xt::zeros<int>({3});
  for (int a = 0; a < 3; ++a) p[a] = procs[a];
  const int size = procs[0] * procs[1] * procs[2];
  const xt::xtensor<std::size_t, 2> idx = reference::compute_cartesian_indices(p);
  for (int i = 0; i < size; ++i)
    for (int a = 0; a < 3; ++a) out[3 * i + a] = (long long)idx(i, a);
}
}
