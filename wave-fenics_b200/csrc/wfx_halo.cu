// Ghost-dof halo exchange inside one NVLink 5 / NVSwitch box.
// Replaces VectorUpdater (demo/gpu_scatter_mpi/VectorUpdater.hpp:21-230).
//
// Two transports:
//  * peer memory (default when every rank of the communicator is a process on this host whose
//    GPU is peer-accessible): the fused ghost reduction  update_rev_fwd[_scaled]  -- the one
//    exchange a time-step stage needs -- is ONE kernel.  Every rank writes its ghost partial sums
//    straight into the owner's receive buffer over NVLink (buffers exported with cudaIpc at
//    create time), releases a per-CTA flag there, the owner adds the contributions in neighbour
//    order (atomic-free segmented reduction, optional 1/m scaling), writes the finished value
//    straight into every ghost holder's receive buffer, releases a flag, and the holders unpack.
//    No pack kernels, no NCCL groups, no host involvement: two NVLink one-way trips.
//  * NCCL (fallback, and the checker of the first in the multi-GPU tests): CUDA-aware
//    MPI_Irecv/MPI_Send per neighbour (:114-129,171-186) become one ncclGroup of
//    ncclSend/ncclRecv per direction on the caller's stream; the atomicAdd unpack of update_rev
//    (:197-198, common/cuda/scatter.cu:38-45) becomes the same segmented reduction.
// update_fwd / update_rev on their own (set-up, tests) always take the NCCL path.
#include "wfx_internal.h"

#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <dlfcn.h>
#include <unistd.h>

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
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi load_nccl()
{
  NcclApi api;
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
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  return api;
}

const NcclApi& nccl()
{
  static const NcclApi api = load_nccl(); // function-local static: initialised once, thread-safe
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

namespace
{
constexpr int P2P_CTAS = 64;     // CTAs of the fused exchange kernel = flags per neighbour and direction
constexpr int P2P_THREADS = 256;
constexpr int P2P_MAXR = 64;     // largest communicator the peer-memory transport handles

// what a rank publishes to the others at create time (all-gathered as bytes)
struct P2PPub
{
  cudaIpcMemHandle_t handle;      // of the rank's pool
  uint64_t host_hash;
  int32_t ok, device, pid, pad;
  int64_t rbuf_off[P2P_MAXR];     // byte offset in the pool where rank r's ghost contributions land (-1: none)
  int64_t fbuf_off[P2P_MAXR];     // ... where owner r's finished values land
  int64_t rflag_off[P2P_MAXR];    // byte offset of the flag row rank r releases after its reverse push
  int64_t fflag_off[P2P_MAXR];    // ... after its forward push
  int32_t rcount[P2P_MAXR];       // entries expected from rank r in the reverse / forward direction
  int32_t fcount[P2P_MAXR];
};

uint64_t host_hash()
{
  char name[256] = {0};
  gethostname(name, sizeof(name) - 1);
  uint64_t h = 1469598103934665603ull;
  for (const char* p = name; *p; ++p) h = (h ^ (unsigned char)*p) * 1099511628211ull;
  // processes in different containers of one machine may share a host name but not /dev/shm:
  // the boot id is the same, the IPC namespace is not -- good enough here, the open is checked anyway
  return h;
}
} // namespace

struct wfx_halo
{
  wfx_ctx* ctx = nullptr;
  wfx_comm* comm = nullptr;
  int dtype = WFX_F64;
  std::vector<int32_t> send_ranks, send_off, recv_ranks, recv_off;
  int64_t nsend = 0, nrecv = 0;
  int64_t n = 0; // size_local + num_ghosts
  DevBuf<int32_t> d_send_idx, d_recv_idx;
  DevBuf<unsigned char> d_send_buf, d_recv_buf;
  // reverse accumulate: unique owned targets and their buffer positions
  int64_t nuniq = 0;
  DevBuf<int32_t> d_uniq, d_useg_src;
  DevBuf<int64_t> d_useg_off;
  // peer-memory transport
  bool p2p = false;
  uint32_t epoch = 0;
  unsigned char* pool = nullptr;      // rbuf | fbuf | rflags | fflags, exported with cudaIpc
  size_t rbuf_o = 0, fbuf_o = 0, rflag_o = 0, fflag_o = 0, pool_bytes = 0;
  std::vector<void*> peer_base;       // opened pools, indexed by rank (nullptr: not a neighbour)
  DevBuf<void*> d_rbuf_dst, d_fbuf_dst;          // per recv / send neighbour: where my values go
  DevBuf<uint32_t*> d_rflag_dst, d_fflag_dst;    // per recv / send neighbour: the flag row I release
  DevBuf<int32_t> d_send_off, d_recv_off;
  DevBuf<uint8_t> d_slot_nbr;                    // send slot -> send neighbour
  ~wfx_halo()
  {
    for (void* p : peer_base)
      if (p) cudaIpcCloseMemHandle(p);
    if (pool) cudaFree(pool);
  }
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

// ---- peer-memory transport ------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// data another GPU wrote into this GPU's memory: read at L2, never from a stale L1 line
template <typename T> __device__ __forceinline__ T ld_peer_written(const T* p) { return __ldcg(p); }

#ifdef WFX_CHECKED
#define WFX_DEV_ASSERT(cond)                                                                          \
  do                                                                                                  \
  {                                                                                                   \
    if (!(cond))                                                                                      \
    {                                                                                                 \
      printf("wfx check failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__,         \
             (int)blockIdx.x, (int)threadIdx.x);                                                      \
      __trap();                                                                                       \
    }                                                                                                 \
  } while (0)
#else
#define WFX_DEV_ASSERT(cond) ((void)0)
#endif

template <typename T>
struct P2PArgs
{
  int64_t n;      // vector length (checked builds)
  int64_t nsend;  // entries of the reverse receive buffer
  T* x;
  const T* scale; // nullable
  int n_recv_nbr, n_send_nbr;
  const int32_t* recv_off;
  const int32_t* recv_idx;
  T* const* rbuf_dst;
  uint32_t* const* rflag_dst;
  int64_t nuniq;
  const int32_t* uniq;
  const int64_t* useg_off;
  const int32_t* useg_src;
  const T* rbuf;
  const uint32_t* rflags;
  const uint8_t* slot_nbr;
  const int32_t* send_off;
  T* const* fbuf_dst;
  uint32_t* const* fflag_dst;
  int64_t nrecv;
  const T* fbuf;
  const uint32_t* fflags;
  uint32_t epoch;
};

// The fused ghost reduction.  No CTA waits for another CTA of its own grid, only for flags that
// remote grids release, and no remote CTA needs more than this rank's flags of the same phase:
// the CTAs need not be co-resident, so the kernel may share the GPU with the interior batches.
template <typename T>
__global__ void __launch_bounds__(P2P_THREADS)
halo_p2p_kernel(const P2PArgs<T> a)
{
  const int cta = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
  constexpr int U = 8; // independent elements per thread and pass: the index -> value -> store chains
                       // are latency-bound, so all loads of a pass are issued before the first store
  const int stride = G * P2P_THREADS;
  // 1. ghost partial sums -> their owners' receive buffers
  for (int n = 0; n < a.n_recv_nbr; ++n)
  {
    const int beg = a.recv_off[n], end = a.recv_off[n + 1];
    T* dst = a.rbuf_dst[n];
    for (int i0 = beg + cta * P2P_THREADS + tid; i0 < end; i0 += U * stride)
    {
      int32_t idx[U];
      T v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) idx[u] = i0 + u * stride < end ? a.recv_idx[i0 + u * stride] : -1;
#pragma unroll
      for (int u = 0; u < U; ++u)
      {
        WFX_DEV_ASSERT(idx[u] < a.n);
        v[u] = idx[u] >= 0 ? a.x[idx[u]] : T(0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (idx[u] >= 0) dst[i0 + u * stride - beg] = v[u];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid < a.n_recv_nbr) st_release_sys(a.rflag_dst[tid] + cta, a.epoch);
  // 2. all contributions to my owned dofs have arrived
  for (int f = tid; f < a.n_send_nbr * G; f += P2P_THREADS)
    while ((int32_t)(ld_acquire_sys(a.rflags + f) - a.epoch) < 0) __nanosleep(32);
  __syncthreads();
  // 3. owner: add in neighbour order, scale, keep, and hand the finished value to every holder
  for (int64_t j0 = cta * P2P_THREADS + tid; j0 < a.nuniq; j0 += (int64_t)U * stride)
  {
    int32_t d[U];
    int64_t p0[U], p1[U];
    T s[U], sc[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int64_t j = j0 + (int64_t)u * stride;
      const bool ok = j < a.nuniq;
      d[u] = ok ? a.uniq[j] : -1;
      p0[u] = ok ? a.useg_off[j] : 0;
      p1[u] = ok ? a.useg_off[j + 1] : 0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      WFX_DEV_ASSERT(d[u] < a.n);
      s[u] = d[u] >= 0 ? a.x[d[u]] : T(0);
      sc[u] = (d[u] >= 0 && a.scale) ? a.scale[d[u]] : T(1);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      for (int64_t p = p0[u]; p < p1[u]; ++p)
      {
        WFX_DEV_ASSERT(a.useg_src[p] >= 0 && a.useg_src[p] < a.nsend);
        s[u] += ld_peer_written(a.rbuf + a.useg_src[p]); // neighbour order: bitwise equal to the NCCL path
      }
      s[u] *= sc[u];
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      if (d[u] < 0) continue;
      a.x[d[u]] = s[u];
      for (int64_t p = p0[u]; p < p1[u]; ++p)
      {
        const int32_t q = a.useg_src[p];
        const int n = a.slot_nbr[q];
        a.fbuf_dst[n][q - a.send_off[n]] = s[u];
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid < a.n_send_nbr) st_release_sys(a.fflag_dst[tid] + cta, a.epoch);
  // 4. the owners' values of my ghosts have arrived
  for (int f = tid; f < a.n_recv_nbr * G; f += P2P_THREADS)
    while ((int32_t)(ld_acquire_sys(a.fflags + f) - a.epoch) < 0) __nanosleep(32);
  __syncthreads();
  // 5. unpack
  for (int64_t i0 = cta * P2P_THREADS + tid; i0 < a.nrecv; i0 += (int64_t)U * stride)
  {
    int32_t idx[U];
    T v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      const int64_t i = i0 + (int64_t)u * stride;
      idx[u] = i < a.nrecv ? a.recv_idx[i] : -1;
      v[u] = i < a.nrecv ? ld_peer_written(a.fbuf + i) : T(0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (idx[u] >= 0) a.x[idx[u]] = v[u];
  }
}

inline unsigned grid_for(int64_t n) { return (unsigned)((n + 255) / 256); }

template <typename T>
void exchange_p2p(wfx_halo* h, T* x, cudaStream_t st, const T* scale)
{
  P2PArgs<T> a;
  a.n = h->n;
  a.nsend = h->nsend;
  a.x = x;
  a.scale = scale;
  a.n_recv_nbr = (int)h->recv_ranks.size();
  a.n_send_nbr = (int)h->send_ranks.size();
  a.recv_off = h->d_recv_off.p;
  a.recv_idx = h->d_recv_idx.p;
  a.rbuf_dst = (T* const*)h->d_rbuf_dst.p;
  a.rflag_dst = h->d_rflag_dst.p;
  a.nuniq = h->nuniq;
  a.uniq = h->d_uniq.p;
  a.useg_off = h->d_useg_off.p;
  a.useg_src = h->d_useg_src.p;
  a.rbuf = (const T*)(h->pool + h->rbuf_o);
  a.rflags = (const uint32_t*)(h->pool + h->rflag_o);
  a.slot_nbr = h->d_slot_nbr.p;
  a.send_off = h->d_send_off.p;
  a.fbuf_dst = (T* const*)h->d_fbuf_dst.p;
  a.fflag_dst = h->d_fflag_dst.p;
  a.nrecv = h->nrecv;
  a.fbuf = (const T*)(h->pool + h->fbuf_o);
  a.fflags = (const uint32_t*)(h->pool + h->fflag_o);
  a.epoch = ++h->epoch;
  halo_p2p_kernel<T><<<P2P_CTAS, P2P_THREADS, 0, st>>>(a);
  WFX_CUDA(cudaGetLastError());
}

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
  if (what == 3 && h->p2p)
  {
    if (h->dtype == WFX_F64) exchange_p2p<double>(h, (double*)x, st, (const double*)scale);
    else exchange_p2p<float>(h, (float*)x, st, (const float*)scale);
    return;
  }
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

// all-gather of one fixed-size record per rank through NCCL (create time only)
template <typename R>
void allgather_records(wfx_comm* c, const R& mine, std::vector<R>& all)
{
  DevBuf<unsigned char> d_in(sizeof(R)), d_out(sizeof(R) * (size_t)c->nranks);
  WFX_CUDA(cudaMemcpy(d_in.p, &mine, sizeof(R), cudaMemcpyHostToDevice));
  WFX_NCCL(nccl().AllGather(d_in.p, d_out.p, sizeof(R), ncclChar, c->comm, nullptr));
  WFX_CUDA(cudaStreamSynchronize(nullptr));
  all.resize((size_t)c->nranks);
  WFX_CUDA(cudaMemcpy(all.data(), d_out.p, sizeof(R) * (size_t)c->nranks, cudaMemcpyDeviceToHost));
}

// Sets up the peer-memory transport; returns false (after every rank has agreed) when it is not
// available.  Collective over the communicator.
bool setup_p2p(wfx_halo* h, size_t esz)
{
  wfx_comm* c = h->comm;
  const int nr = c->nranks, me = c->rank;
  auto align = [](size_t v) { return (v + 255) & ~(size_t)255; };
  P2PPub pub;
  memset(&pub, 0, sizeof(pub));
  pub.ok = nr <= P2P_MAXR && h->send_ranks.size() < 256 && h->recv_ranks.size() < 256 ? 1 : 0;
  pub.host_hash = host_hash();
  pub.device = h->ctx->device;
  pub.pid = (int32_t)getpid();
  for (int r = 0; r < P2P_MAXR; ++r) pub.rbuf_off[r] = pub.fbuf_off[r] = pub.rflag_off[r] = pub.fflag_off[r] = -1;
  // pool: rbuf (contributions to my owned dofs, send-list layout) | fbuf (owners' values of my
  // ghosts, recv-list layout) | one flag row of P2P_CTAS words per neighbour and direction
  h->rbuf_o = 0;
  h->fbuf_o = align(h->rbuf_o + (size_t)h->nsend * esz);
  h->rflag_o = align(h->fbuf_o + (size_t)h->nrecv * esz);
  h->fflag_o = align(h->rflag_o + h->send_ranks.size() * P2P_CTAS * 4);
  h->pool_bytes = align(h->fflag_o + h->recv_ranks.size() * P2P_CTAS * 4) + 256;
  if (pub.ok)
  {
    if (cudaMalloc((void**)&h->pool, h->pool_bytes) != cudaSuccess || cudaMemset(h->pool, 0, h->pool_bytes) != cudaSuccess
        || cudaDeviceSynchronize() != cudaSuccess || cudaIpcGetMemHandle(&pub.handle, h->pool) != cudaSuccess)
    {
      cudaGetLastError();
      pub.ok = 0;
    }
  }
  if (pub.ok)
  {
    for (size_t i = 0; i < h->send_ranks.size(); ++i)
    {
      const int r = h->send_ranks[i];
      pub.rbuf_off[r] = (int64_t)(h->rbuf_o + (size_t)h->send_off[i] * esz);
      pub.rflag_off[r] = (int64_t)(h->rflag_o + i * P2P_CTAS * 4);
      pub.rcount[r] = h->send_off[i + 1] - h->send_off[i];
    }
    for (size_t j = 0; j < h->recv_ranks.size(); ++j)
    {
      const int r = h->recv_ranks[j];
      pub.fbuf_off[r] = (int64_t)(h->fbuf_o + (size_t)h->recv_off[j] * esz);
      pub.fflag_off[r] = (int64_t)(h->fflag_o + j * P2P_CTAS * 4);
      pub.fcount[r] = h->recv_off[j + 1] - h->recv_off[j];
    }
  }
  std::vector<P2PPub> all;
  allgather_records(c, pub, all);
  bool ok = true;
  for (int r = 0; r < nr; ++r) ok = ok && all[r].ok && all[r].host_hash == pub.host_hash;
  // the two ends of every edge must agree on the message sizes (guards against inconsistent input)
  if (ok)
  {
    for (size_t j = 0; j < h->recv_ranks.size() && ok; ++j)
    {
      const int o = h->recv_ranks[j];
      ok = all[o].rbuf_off[me] >= 0 && all[o].rcount[me] == h->recv_off[j + 1] - h->recv_off[j];
    }
    for (size_t i = 0; i < h->send_ranks.size() && ok; ++i)
    {
      const int g = h->send_ranks[i];
      ok = all[g].fbuf_off[me] >= 0 && all[g].fcount[me] == h->send_off[i + 1] - h->send_off[i];
    }
  }
  // open the neighbours' pools
  h->peer_base.assign((size_t)nr, nullptr);
  if (ok)
  {
    std::vector<int> nbrs(h->send_ranks.begin(), h->send_ranks.end());
    nbrs.insert(nbrs.end(), h->recv_ranks.begin(), h->recv_ranks.end());
    std::sort(nbrs.begin(), nbrs.end());
    nbrs.erase(std::unique(nbrs.begin(), nbrs.end()), nbrs.end());
    for (int r : nbrs)
    {
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
      {
        cudaGetLastError();
        ok = false;
        break;
      }
      h->peer_base[r] = p;
    }
  }
  // agree: one rank that cannot open a pool sends everybody to the NCCL path
  struct Vote { int32_t ok; };
  std::vector<Vote> votes;
  allgather_records(c, Vote{ok ? 1 : 0}, votes);
  for (const Vote& v : votes) ok = ok && v.ok;
  if (!ok)
  {
    for (void*& p : h->peer_base)
    {
      if (p) cudaIpcCloseMemHandle(p);
      p = nullptr;
    }
    if (h->pool) cudaFree(h->pool);
    h->pool = nullptr;
    return false;
  }
  std::vector<void*> rdst, fdst;
  std::vector<uint32_t*> rfl, ffl;
  for (size_t j = 0; j < h->recv_ranks.size(); ++j)
  {
    const int o = h->recv_ranks[j];
    rdst.push_back((unsigned char*)h->peer_base[o] + all[o].rbuf_off[me]);
    rfl.push_back((uint32_t*)((unsigned char*)h->peer_base[o] + all[o].rflag_off[me]));
  }
  std::vector<uint8_t> slot_nbr((size_t)h->nsend);
  for (size_t i = 0; i < h->send_ranks.size(); ++i)
  {
    const int g = h->send_ranks[i];
    fdst.push_back((unsigned char*)h->peer_base[g] + all[g].fbuf_off[me]);
    ffl.push_back((uint32_t*)((unsigned char*)h->peer_base[g] + all[g].fflag_off[me]));
    for (int q = h->send_off[i]; q < h->send_off[i + 1]; ++q) slot_nbr[q] = (uint8_t)i;
  }
  h->d_rbuf_dst.upload(rdst);
  h->d_rflag_dst.upload(rfl);
  h->d_fbuf_dst.upload(fdst);
  h->d_fflag_dst.upload(ffl);
  h->d_slot_nbr.upload(slot_nbr);
  h->d_send_off.upload(h->send_off);
  h->d_recv_off.upload(h->recv_off);
  return true;
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

namespace
{
// offsets start at 0 and never decrease, neighbour ranks are valid and distinct, indices lie in the
// stated part of the vector: bad input must fail here, not as out-of-bounds device writes
void validate_lists(const char* what, int nnbr, const int32_t* ranks, const int32_t* offsets,
                    const int32_t* indices, int nranks, int me, int64_t lo, int64_t hi)
{
  if (nnbr == 0) return;
  if (!ranks || !offsets) fail("halo: %s ranks / offsets are NULL", what);
  if (offsets[0] != 0) fail("halo: %s offsets do not start at 0", what);
  std::vector<uint8_t> seen((size_t)nranks, 0);
  for (int i = 0; i < nnbr; ++i)
  {
    if (offsets[i + 1] < offsets[i]) fail("halo: %s offsets decrease at neighbour %d", what, i);
    const int r = ranks[i];
    if (r < 0 || r >= nranks || r == me) fail("halo: bad %s rank %d", what, r);
    if (seen[r]) fail("halo: %s rank %d listed twice", what, r);
    seen[r] = 1;
  }
  const int64_t n = offsets[nnbr];
  if (n > 0 && !indices) fail("halo: %s indices are NULL", what);
  for (int64_t p = 0; p < n; ++p)
    if (indices[p] < lo || indices[p] >= hi)
      fail("halo: %s index %d outside [%lld, %lld)", what, indices[p], (long long)lo, (long long)hi);
}
} // namespace

extern "C" int wfx_halo_create(wfx_ctx* ctx, wfx_comm* comm, int dtype, int64_t size_local,
                               int64_t num_ghosts, int n_send_nbr, const int32_t* send_ranks,
                               const int32_t* send_offsets, const int32_t* send_indices,
                               int n_recv_nbr, const int32_t* recv_ranks,
                               const int32_t* recv_offsets, const int32_t* recv_indices,
                               wfx_halo** out)
{
  WFX_API_BEGIN
  if (!ctx || !comm || !out) fail("NULL argument");
  if (dtype != WFX_F64 && dtype != WFX_F32) fail("unknown dtype %d", dtype);
  if (n_send_nbr < 0 || n_recv_nbr < 0) fail("negative neighbour count");
  if (size_local < 0 || num_ghosts < 0) fail("halo: negative vector size");
  validate_lists("send", n_send_nbr, send_ranks, send_offsets, send_indices, comm->nranks, comm->rank, 0, size_local);
  validate_lists("receive", n_recv_nbr, recv_ranks, recv_offsets, recv_indices, comm->nranks, comm->rank, size_local,
                 size_local + num_ghosts);
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
  h->nsend = h->send_off.back();
  h->nrecv = h->recv_off.back();
  h->n = size_local + num_ghosts;
  {
    // a ghost slot is filled by exactly one owner
    std::vector<uint8_t> seen((size_t)num_ghosts, 0);
    for (int64_t p = 0; p < h->nrecv; ++p)
    {
      uint8_t& s = seen[recv_indices[p] - size_local];
      if (s) fail("halo: ghost slot %d listed twice", recv_indices[p]);
      s = 1;
    }
  }
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
  // transport: peer memory unless told otherwise or unavailable (collective decision)
  int want = -1; // -1 auto, 0 NCCL, 1 peer memory required
  if (const char* e = std::getenv("WFX_HALO_TRANSPORT"))
  {
    if (!strcmp(e, "nccl")) want = 0;
    else if (!strcmp(e, "p2p")) want = 1;
    else if (strcmp(e, "auto")) fail("WFX_HALO_TRANSPORT must be auto, nccl or p2p");
  }
  if (want != 0 && comm->nranks > 1)
  {
    h->p2p = setup_p2p(h.get(), esz);
    if (want == 1 && !h->p2p) fail("halo: peer-memory transport requested but not available");
  }
  *out = h.release();
  WFX_API_END
}

extern "C" int wfx_halo_transport(wfx_halo* h, int* transport)
{
  WFX_API_BEGIN
  if (!h || !transport) fail("NULL argument");
  *transport = h->p2p ? 1 : 0;
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
