// Device context and raw buffers of the C ABI.
#include "wfx_internal.h"

#include <nvtx3/nvToolsExt.h>

#include <chrono>
#include <cstdlib>

using namespace wfx;

namespace
{
double wall_now()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // namespace
wfx::SetupTimer::SetupTimer(const char* w) : what(w), last(wall_now())
{
  const char* e = std::getenv("WFX_VERBOSE");
  on = e && std::atoi(e) != 0;
}
void wfx::SetupTimer::lap(const char* phase)
{
  if (!on) return;
  cudaDeviceSynchronize();
  const double t = wall_now();
  std::fprintf(stderr, "[wfx setup] %-22s %-28s %8.3f s\n", what, phase, t - last);
  last = t;
}
wfx::NvtxRange::NvtxRange(const char* name) { nvtxRangePushA(name); }
wfx::NvtxRange::~NvtxRange() { nvtxRangePop(); }

extern "C" int wfx_ctx_create(int device, wfx_ctx** out)
{
  WFX_API_BEGIN
  if (!out) fail("ctx output pointer is NULL");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    fail("no CUDA device available (%s): libwavefx has no CPU fallback",
         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
  if (device < 0 || device >= count) fail("device %d out of range (have %d)", device, count);
  WFX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  WFX_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    fail("device %d is sm_%d%d; libwavefx is built for sm_100a only", device, prop.major, prop.minor);
  auto* ctx = new wfx_ctx;
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  *out = ctx;
  WFX_API_END
}

extern "C" int wfx_ctx_destroy(wfx_ctx* ctx)
{
  WFX_API_BEGIN
  delete ctx;
  WFX_API_END
}

extern "C" int wfx_ctx_sync(wfx_ctx* ctx)
{
  WFX_API_BEGIN
  if (!ctx) fail("ctx is NULL");
  ScopedDevice sd(ctx->device);
  WFX_CUDA(cudaDeviceSynchronize());
  WFX_API_END
}

extern "C" int wfx_malloc(wfx_ctx* ctx, int64_t nbytes, void** ptr)
{
  WFX_API_BEGIN
  if (!ctx || !ptr) fail("NULL argument");
  ScopedDevice sd(ctx->device);
  *ptr = nullptr;
  if (nbytes > 0) WFX_CUDA(cudaMalloc(ptr, (size_t)nbytes));
  WFX_API_END
}

extern "C" int wfx_free(wfx_ctx* ctx, void* ptr)
{
  WFX_API_BEGIN
  if (!ctx) fail("ctx is NULL");
  ScopedDevice sd(ctx->device);
  if (ptr) WFX_CUDA(cudaFree(ptr));
  WFX_API_END
}

extern "C" int wfx_memcpy_h2d(wfx_ctx* ctx, void* dst, const void* src, int64_t nbytes)
{
  WFX_API_BEGIN
  if (!ctx) fail("ctx is NULL");
  ScopedDevice sd(ctx->device);
  if (nbytes > 0) WFX_CUDA(cudaMemcpy(dst, src, (size_t)nbytes, cudaMemcpyHostToDevice));
  WFX_API_END
}

extern "C" int wfx_memcpy_d2h(wfx_ctx* ctx, void* dst, const void* src, int64_t nbytes)
{
  WFX_API_BEGIN
  if (!ctx) fail("ctx is NULL");
  ScopedDevice sd(ctx->device);
  if (nbytes > 0) WFX_CUDA(cudaMemcpy(dst, src, (size_t)nbytes, cudaMemcpyDeviceToHost));
  WFX_API_END
}
