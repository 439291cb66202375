// C entry points over the REFERENCE's own CUDA primitives, compiled from the sources where they lie
// (/root/reference/common/cuda/{scatter,transform,mass_kernel}.cu) into oracle/_ref/libwfref_cuda.so
// by oracle/build_ref.py.  TEST INFRASTRUCTURE ONLY: the reference's gather / atomic scatter /
// transform1 kernels are the part of the hot path's prior art that compiles without DOLFINx, Basix or
// xtensor; the GPU tests use this library as the checker of wfx_gather, wfx_scatter_add and the
// diagonal mass apply (SpectralMassOperator::apply, common/cuda/spectral_mass.hpp:84-89).
// This file contains no reference code: it declares the reference's templates (scatter.hpp:7-14,
// transform.hpp:7-8) through their own headers and forwards to them.
#include "scatter.hpp"
#include "transform.hpp"

#include <cstdint>

extern "C" {
void ref_gather_f64(std::int32_t n, const std::int32_t* idx, const double* in, double* out) { gather<double>(n, idx, in, out, 512); }
void ref_gather_f32(std::int32_t n, const std::int32_t* idx, const float* in, float* out) { gather<float>(n, idx, in, out, 512); }
void ref_scatter_f64(std::int32_t n, const std::int32_t* idx, const double* in, double* out) { scatter<double>(n, idx, in, out, 512); }
void ref_scatter_f32(std::int32_t n, const std::int32_t* idx, const float* in, float* out) { scatter<float>(n, idx, in, out, 512); }
void ref_transform1_f64(std::int32_t n, const double* in, double* detJ, double* out) { transform1<double>(n, in, detJ, out, 512); }
void ref_transform1_f32(std::int32_t n, const float* in, float* detJ, float* out) { transform1<float>(n, in, detJ, out, 512); }
int ref_cuda_version() { return 1; }
}
