// Micro-benchmark: how fast can the stiffness kernel's G stream (6000 B per cell, one warp
// per cell, 8 cells per warp, 2 CTAs x 8 warps per SM) be pulled from HBM, depending on
// how the loads are issued.  Not part of the product; used to choose the kernel design.
//   mode 0: LDG.128 by 25 lanes, consume right after the load (no prefetch)
//   mode 1: LDG.128 by 25 lanes, next cell requested before the current one is consumed
//   mode 2: as mode 1 with all 32 lanes (6144 B per cell)
//   mode 3: cp.async.bulk (TMA) 6000 B into shared memory, double-buffered per warp
// argv: mode burn scattered
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ double2 ld_stream(const double2* p)
{
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int CELL_B = 6000, CELLS = 64, W = 8;

__device__ __forceinline__ int cell_of(int batch, int r, int w, int scattered)
{
  if (!scattered) return batch * CELLS + r * W + w;
  // 64^3 lexicographic mesh, 4x4x4 bricks, colour-parity rounds like the real plan
  const int bx = batch / 256, by = (batch / 16) % 16, bz = batch % 16;
  const int lx = (r & 1) + 2 * (w & 1), ly = ((r >> 1) & 1) + 2 * ((w >> 1) & 1), lz = ((r >> 2) & 1) + 2 * ((w >> 2) & 1);
  return ((bx * 4 + lx) * 64 + (by * 4 + ly)) * 64 + bz * 4 + lz;
}

template <int LANES, bool PREFETCH>
__global__ void __launch_bounds__(256, 2) k_ldg(const double* __restrict__ G, double* out, int burn, int scattered)
{
  extern __shared__ double sm[];
  constexpr int NV = (LANES == 25) ? 15 : 12;
  const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
  const bool ok = lane < LANES;
  double2 g[NV];
  double acc = 0;
  auto load = [&](int cell) {
    const double2* p = reinterpret_cast<const double2*>(G + (size_t)cell * (CELL_B / 8)) + lane;
#pragma unroll
    for (int v = 0; v < NV; ++v) g[v] = ld_stream(p + v * LANES);
  };
  if (PREFETCH && ok) load(cell_of(blockIdx.x, 0, w, scattered));
  for (int r = 0; r < 8; ++r)
  {
    if (!PREFETCH && ok) load(cell_of(blockIdx.x, r, w, scattered));
    double s = 0;
    if (ok)
    {
#pragma unroll
      for (int v = 0; v < NV; ++v) s += g[v].x + g[v].y;
    }
    if (PREFETCH && ok && r + 1 < 8) load(cell_of(blockIdx.x, r + 1, w, scattered));
    for (int it = 0; it < burn; ++it) s = fma(s, 1.0000001, 1e-9);
    acc += s;
    __syncthreads();
  }
  if (acc == 123.456) out[0] = acc + sm[0];
}

__global__ void __launch_bounds__(256, 2) k_tma(const double* __restrict__ G, double* out, int burn, int scattered)
{
  extern __shared__ __align__(128) unsigned char smraw[];
  // per warp: 2 stages x 6000 B (padded to 6016), then 2 mbarriers per warp at the end
  constexpr int STG = 6016;
  const int w = threadIdx.x / 32, lane = threadIdx.x % 32;
  unsigned char* buf = smraw + w * 2 * STG;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smraw + W * 2 * STG) + w * 2;
  if (lane == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)));
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  auto issue = [&](int cell, int st) {
    if (lane == 0)
    {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar + st)), "r"(CELL_B) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(buf + st * STG)),
                   "l"(G + (size_t)cell * (CELL_B / 8)), "r"(CELL_B), "r"(smem_u32(bar + st))
                   : "memory");
    }
  };
  double acc = 0;
  issue(cell_of(blockIdx.x, 0, w, scattered), 0);
  for (int r = 0; r < 8; ++r)
  {
    const int st = r & 1;
    if (r + 1 < 8) issue(cell_of(blockIdx.x, r + 1, w, scattered), st ^ 1);
    const uint32_t parity = (r >> 1) & 1;
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(done) : "r"(smem_u32(bar + st)), "r"(parity) : "memory");
    double s = 0;
    if (lane < 25)
    {
      const double2* p = reinterpret_cast<const double2*>(buf + st * STG) + lane;
#pragma unroll
      for (int v = 0; v < 15; ++v) { double2 q = p[v * 25]; s += q.x + q.y; }
    }
    for (int it = 0; it < burn; ++it) s = fma(s, 1.0000001, 1e-9);
    acc += s;
    __syncwarp();
    __syncthreads();
  }
  if (acc == 123.456) out[0] = acc;
}

int main(int argc, char** argv)
{
  const int mode = argc > 1 ? atoi(argv[1]) : 0, burn = argc > 2 ? atoi(argv[2]) : 0, scattered = argc > 3 ? atoi(argv[3]) : 0;
  const int nb = 4096;
  const size_t bytes = (size_t)nb * CELLS * 6144;
  double *G, *out;
  CK(cudaMalloc(&G, bytes));
  CK(cudaMalloc(&out, 8));
  CK(cudaMemset(G, 0, bytes));
  const size_t smem = 100 * 1024;
  auto launch = [&]() {
    if (mode == 0) k_ldg<25, false><<<nb, 256, smem>>>(G, out, burn, scattered);
    else if (mode == 1) k_ldg<25, true><<<nb, 256, smem>>>(G, out, burn, scattered);
    else if (mode == 2) k_ldg<32, true><<<nb, 256, smem>>>(G, out, burn, scattered);
    else k_tma<<<nb, 256, smem>>>(G, out, burn, scattered);
  };
  CK(cudaFuncSetAttribute(k_ldg<25, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_ldg<25, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_ldg<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int i = 0; i < 3; ++i) launch();
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int i = 0; i < reps; ++i) launch();
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  const double moved = (double)nb * CELLS * (mode == 2 ? 6144 : 6000);
  printf("mode %d burn %5d scattered %d : %.3f ms  %.0f GB/s\n", mode, burn, scattered, ms, moved / ms / 1e6);
  return 0;
}
