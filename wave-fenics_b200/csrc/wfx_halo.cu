// Ghost-dof halo exchange over NCCL (NVLink 5 / NVSwitch inside one box).
// Replaces VectorUpdater (demo/gpu_scatter_mpi/VectorUpdater.hpp:21-230): CUDA-aware
// MPI_Irecv/MPI_Send per neighbour become one ncclGroup of ncclSend/ncclRecv on the
// caller's stream; the atomicAdd unpack of update_rev (:197-198, common/cuda/scatter.cu:38-45)
// becomes a segmented reduction that adds the neighbours' contributions in neighbour order.
#include "wfx_internal.h"

#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <dlfcn.h>

using namespace wfx;

// NCCL is bound at first use with dlopen, not at link time: a process that also hosts
// PyTorch must end up with ONE libnccl.so.2 (torch bundles a newer one than the system's;
// whichever is already loaded is reused by soname).
namespace
{
struct NcclApi
{
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi& nccl()
{
  static NcclApi api;
  static bool ready = false;
  if (!ready)
  {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) fail("cannot load libnccl.so.2: %s", dlerror());
    auto sym = [&](const char* name) {
      void* p = dlsym(h, name);
      if (!p) fail("libnccl.so.2 lacks %s", name);
      return p;
    };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    ready = true;
  }
  return api;
}
} // namespace

#define WFX_NCCL(call)                                                                           \
  do                                                                                             \
  {                                                                                              \
    ncclResult_t r_ = (call);                                                                    \
    if (r_ != ncclSuccess)                                                                       \
      wfx::fail("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, nccl().GetErrorString(r_));      \
  } while (0)

struct wfx_comm
{
  wfx_ctx* ctx = nullptr;
  ncclComm_t comm = nullptr;
  int nranks = 0, rank = 0;
};

struct wfx_halo
{
  wfx_ctx* ctx = nullptr;
  wfx_comm* comm = nullptr;
  int dtype = WFX_F64;
  std::vector<int32_t> send_ranks, send_off, recv_ranks, recv_off;
  int64_t nsend = 0, nrecv = 0;
  DevBuf<int32_t> d_send_idx, d_recv_idx;
  DevBuf<unsigned char> d_send_buf, d_recv_buf;
  // reverse accumulate: unique owned targets and their buffer positions
  int64_t nuniq = 0;
  DevBuf<int32_t> d_uniq, d_useg_src;
  DevBuf<int64_t> d_useg_off;
};

namespace
{
template <typename T>
__global__ void pack_kernel(int64_t n, const int32_t* __restrict__ idx, const T* __restrict__ x,
                            T* __restrict__ buf)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) buf[i] = x[idx[i]];
}
template <typename T>
__global__ void unpack_copy_kernel(int64_t n, const int32_t* __restrict__ idx,
                                   const T* __restrict__ buf, T* __restrict__ x)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) x[idx[i]] = buf[i];
}
template <typename T>
__global__ void unpack_add_kernel(int64_t nuniq, const int32_t* __restrict__ uniq,
                                  const int64_t* __restrict__ off, const int32_t* __restrict__ src,
                                  const T* __restrict__ buf, T* __restrict__ x,
                                  const T* __restrict__ scale)
{
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= nuniq) return;
  T s = x[uniq[j]];
  for (int64_t p = off[j]; p < off[j + 1]; ++p) s += buf[src[p]];
  if (scale) s *= scale[uniq[j]];
  x[uniq[j]] = s;
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + 255) / 256); }

template <typename T>
void exchange(wfx_halo* h, bool forward, T* x, cudaStream_t st, const T* scale = nullptr)
{
  const ncclDataType_t dt = sizeof(T) == 8 ? ncclDouble : ncclFloat;
  T* sbuf = (T*)h->d_send_buf.p;
  T* rbuf = (T*)h->d_recv_buf.p;
  if (forward)
  {
    if (h->nsend) pack_kernel<T><<<grid_for(h->nsend), 256, 0, st>>>(h->nsend, h->d_send_idx.p, x, sbuf);
    WFX_NCCL(nccl().GroupStart());
    for (size_t i = 0; i < h->recv_ranks.size(); ++i)
      WFX_NCCL(nccl().Recv(rbuf + h->recv_off[i], h->recv_off[i + 1] - h->recv_off[i], dt, h->recv_ranks[i], h->comm->comm, st));
    for (size_t i = 0; i < h->send_ranks.size(); ++i)
      WFX_NCCL(nccl().Send(sbuf + h->send_off[i], h->send_off[i + 1] - h->send_off[i], dt, h->send_ranks[i], h->comm->comm, st));
    WFX_NCCL(nccl().GroupEnd());
    if (h->nrecv) unpack_copy_kernel<T><<<grid_for(h->nrecv), 256, 0, st>>>(h->nrecv, h->d_recv_idx.p, rbuf, x);
  }
  else
  {
    if (h->nrecv) pack_kernel<T><<<grid_for(h->nrecv), 256, 0, st>>>(h->nrecv, h->d_recv_idx.p, x, rbuf);
    WFX_NCCL(nccl().GroupStart());
    for (size_t i = 0; i < h->send_ranks.size(); ++i)
      WFX_NCCL(nccl().Recv(sbuf + h->send_off[i], h->send_off[i + 1] - h->send_off[i], dt, h->send_ranks[i], h->comm->comm, st));
    for (size_t i = 0; i < h->recv_ranks.size(); ++i)
      WFX_NCCL(nccl().Send(rbuf + h->recv_off[i], h->recv_off[i + 1] - h->recv_off[i], dt, h->recv_ranks[i], h->comm->comm, st));
    WFX_NCCL(nccl().GroupEnd());
    if (h->nuniq)
      unpack_add_kernel<T><<<grid_for(h->nuniq), 256, 0, st>>>(h->nuniq, h->d_uniq.p, h->d_useg_off.p, h->d_useg_src.p, sbuf, x, scale);
  }
  WFX_CUDA(cudaGetLastError());
}

void run(wfx_halo* h, int what, void* x, void* stream, const void* scale = nullptr)
{
  if (!h) fail("halo is NULL");
  if (!x) fail("halo: NULL vector");
  ScopedDevice sd(h->ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (h->dtype == WFX_F64)
  {
    if (what & 1) exchange<double>(h, false, (double*)x, st, (const double*)scale);
    if (what & 2) exchange<double>(h, true, (double*)x, st);
  }
  else
  {
    if (what & 1) exchange<float>(h, false, (float*)x, st, (const float*)scale);
    if (what & 2) exchange<float>(h, true, (float*)x, st);
  }
}
} // namespace

namespace wfx
{
int halo_dtype(const wfx_halo* h) { return h->dtype; }
} // namespace wfx

extern "C" int wfx_comm_unique_id(char id[128])
{
  WFX_API_BEGIN
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId uid;
  WFX_NCCL(nccl().GetUniqueId(&uid));
  memcpy(id, &uid, 128);
  WFX_API_END
}

extern "C" int wfx_comm_create(wfx_ctx* ctx, const char id[128], int nranks, int rank, wfx_comm** out)
{
  WFX_API_BEGIN
  if (!ctx || !id || !out) fail("NULL argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) fail("bad rank %d of %d", rank, nranks);
  ScopedDevice sd(ctx->device);
  auto c = std::make_unique<wfx_comm>();
  c->ctx = ctx;
  c->nranks = nranks;
  c->rank = rank;
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  WFX_NCCL(nccl().CommInitRank(&c->comm, nranks, uid, rank));
  *out = c.release();
  WFX_API_END
}

extern "C" int wfx_comm_destroy(wfx_comm* c)
{
  WFX_API_BEGIN
  if (c)
  {
    ScopedDevice sd(c->ctx->device);
    if (c->comm) nccl().CommDestroy(c->comm);
    delete c;
  }
  WFX_API_END
}

extern "C" int wfx_halo_create(wfx_ctx* ctx, wfx_comm* comm, int dtype, int n_send_nbr,
                               const int32_t* send_ranks, const int32_t* send_offsets,
                               const int32_t* send_indices, int n_recv_nbr,
                               const int32_t* recv_ranks, const int32_t* recv_offsets,
                               const int32_t* recv_indices, wfx_halo** out)
{
  WFX_API_BEGIN
  if (!ctx || !comm || !out) fail("NULL argument");
  if (dtype != WFX_F64 && dtype != WFX_F32) fail("unknown dtype %d", dtype);
  if (n_send_nbr < 0 || n_recv_nbr < 0) fail("negative neighbour count");
  ScopedDevice sd(ctx->device);
  auto h = std::make_unique<wfx_halo>();
  h->ctx = ctx;
  h->comm = comm;
  h->dtype = dtype;
  h->send_ranks.assign(send_ranks, send_ranks + n_send_nbr);
  h->recv_ranks.assign(recv_ranks, recv_ranks + n_recv_nbr);
  if (n_send_nbr) h->send_off.assign(send_offsets, send_offsets + n_send_nbr + 1);
  else h->send_off.assign(1, 0);
  if (n_recv_nbr) h->recv_off.assign(recv_offsets, recv_offsets + n_recv_nbr + 1);
  else h->recv_off.assign(1, 0);
  for (int r : h->send_ranks)
    if (r < 0 || r >= comm->nranks || r == comm->rank) fail("halo: bad destination rank %d", r);
  for (int r : h->recv_ranks)
    if (r < 0 || r >= comm->nranks || r == comm->rank) fail("halo: bad source rank %d", r);
  h->nsend = h->send_off.back();
  h->nrecv = h->recv_off.back();
  const size_t esz = dtype == WFX_F64 ? 8 : 4;
  if (h->nsend)
  {
    h->d_send_idx.upload(send_indices, (size_t)h->nsend);
    h->d_send_buf.alloc((size_t)h->nsend * esz);
    // group buffer positions by owned target index, ascending position (neighbour order)
    std::vector<std::pair<int32_t, int32_t>> pr((size_t)h->nsend);
    for (int64_t p = 0; p < h->nsend; ++p) pr[p] = {send_indices[p], (int32_t)p};
    std::sort(pr.begin(), pr.end());
    std::vector<int32_t> uniq, src((size_t)h->nsend);
    std::vector<int64_t> off;
    for (int64_t p = 0; p < h->nsend; ++p)
    {
      if (p == 0 || pr[p].first != pr[p - 1].first)
      {
        uniq.push_back(pr[p].first);
        off.push_back(p);
      }
      src[p] = pr[p].second;
    }
    off.push_back(h->nsend);
    h->nuniq = (int64_t)uniq.size();
    h->d_uniq.upload(uniq);
    h->d_useg_off.upload(off);
    h->d_useg_src.upload(src);
  }
  if (h->nrecv)
  {
    h->d_recv_idx.upload(recv_indices, (size_t)h->nrecv);
    h->d_recv_buf.alloc((size_t)h->nrecv * esz);
  }
  *out = h.release();
  WFX_API_END
}

extern "C" int wfx_halo_update_fwd(wfx_halo* h, void* x, void* stream)
{
  WFX_API_BEGIN
  run(h, 2, x, stream);
  WFX_API_END
}

extern "C" int wfx_halo_update_rev(wfx_halo* h, void* x, void* stream)
{
  WFX_API_BEGIN
  run(h, 1, x, stream);
  WFX_API_END
}

extern "C" int wfx_halo_update_rev_fwd(wfx_halo* h, void* x, void* stream)
{
  WFX_API_BEGIN
  run(h, 3, x, stream);
  WFX_API_END
}

extern "C" int wfx_halo_update_rev_fwd_scaled(wfx_halo* h, void* x, const void* scale, void* stream)
{
  WFX_API_BEGIN
  if (!scale) fail("halo: scale vector is NULL");
  run(h, 3, x, stream, scale);
  WFX_API_END
}

extern "C" int wfx_halo_destroy(wfx_halo* h)
{
  WFX_API_BEGIN
  if (h)
  {
    ScopedDevice sd(h->ctx->device);
    delete h;
  }
  WFX_API_END
}
