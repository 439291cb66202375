// Wave model and classical RK4 time integrator on the device.
// Replaces LinearGLLOpt::{init,f0,f1,rk4} (common/LinearGLL.hpp:131-287).  The reference
// makes ~17 full-vector passes per stage (copies, axpys, zero, divide) plus three host
// MPI scatters; here a stage is: stiffness apply (b = -c0^2 K un), boundary term on the
// facet dofs, one optional halo reduction of b, and ONE fused pointwise kernel that does
// kv = b/m, both solution updates and the next stage state.
#include "wfx_internal.h"

#include <cmath>
#include <cstdlib>

using namespace wfx;

struct wfx_stiffness;
struct wfx_mass;
struct wfx_boundary;
struct wfx_halo;

struct wfx_wave
{
  wfx_ctx* ctx = nullptr;
  wfx_stiffness* stiff = nullptr;
  wfx_mass* mass = nullptr;
  wfx_boundary* bnd = nullptr;
  wfx_halo* halo = nullptr;
  int dtype = WFX_F64;
  int64_t n = 0, size_local = 0;
  double c0 = 0, f0 = 0, p0 = 0;
  const void* minv = nullptr;
  // u_, v_: solution; u0, v0: start of step; un, vn: stage state; b: right-hand side
  DevBuf<unsigned char> u_, v_, u0, v0, un, vn, b;
  // distributed runs: the ghost reduction runs on its own stream, ordered by events
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_iface = nullptr, ev_halo = nullptr;
  // one-rank runs: a full time step (4 stages, ~40 launches) is captured once per step size
  // into a CUDA graph and replayed; the four source amplitudes of the step live in g_dev
  bool use_graph = true;
  cudaGraphExec_t step_graph = nullptr;
  double graph_dt = 0;
  DevBuf<double> g_dev;
  cudaStream_t own_stream = nullptr; // capture needs a non-legacy stream
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  // time-stepping position (reset by init / set_state): steps completed and the time reached
  int64_t steps_done = 0;
  double t_now = 0;
  // probes: u at selected dofs after every completed step, kept on the device until asked for
  int64_t nprobes = 0, probe_cap = 0, probe_rec = 0;
  DevBuf<int32_t> d_probe_idx;
  DevBuf<unsigned char> d_probe_val;
  std::vector<double> probe_t;
  // snapshots: every `snap_every` steps u, v are copied device -> device on the work stream, then
  // device -> pinned host on a copy stream while the time stepping goes on; the callback runs on the
  // calling thread once the copy has landed (at the next snapshot or at the end of rk4)
  int64_t snap_every = 0;
  wfx_snapshot_fn snap_fn = nullptr;
  void* snap_user = nullptr;
  DevBuf<unsigned char> snap_u, snap_v;
  void *snap_hu = nullptr, *snap_hv = nullptr; // pinned
  cudaStream_t snap_stream = nullptr;
  cudaEvent_t ev_snap_ready = nullptr, ev_snap_done = nullptr;
  bool snap_pending = false;
  int64_t snap_step = 0;
  double snap_t = 0;
  ~wfx_wave()
  {
    if (snap_hu) cudaFreeHost(snap_hu);
    if (snap_hv) cudaFreeHost(snap_hv);
    if (snap_stream) cudaStreamDestroy(snap_stream);
    if (ev_snap_ready) cudaEventDestroy(ev_snap_ready);
    if (ev_snap_done) cudaEventDestroy(ev_snap_done);
    if (step_graph) cudaGraphExecDestroy(step_graph);
    if (own_stream) cudaStreamDestroy(own_stream);
    if (ev_in) cudaEventDestroy(ev_in);
    if (ev_out) cudaEventDestroy(ev_out);
    if (ev_iface) cudaEventDestroy(ev_iface);
    if (ev_halo) cudaEventDestroy(ev_halo);
    if (comm_stream) cudaStreamDestroy(comm_stream);
  }
};

namespace
{
// Stage update, fused.  STAGE 0 reads the solution as stage state
// and saves it as u0/v0; STAGE 3 only finishes the solution.
//   kv = b / m                          (LinearGLL.hpp:188-191; m^-1 precomputed, :179-181)
//   u_ += b_i dt ku,  v_ += b_i dt kv   (:264-265), ku = vn (f0, :141-144)
//   un' = u0 + a_{i+1} dt ku, vn' = v0 + a_{i+1} dt kv   (:250-254 of the next stage)
template <typename T, int STAGE>
__global__ void __launch_bounds__(256)
rk_stage_kernel(int64_t n, const T* __restrict__ b, const T* __restrict__ minv, T* __restrict__ u_,
                T* __restrict__ v_, T* __restrict__ u0, T* __restrict__ v0, T* __restrict__ un,
                T* __restrict__ vn, T bdt, T adt_next)
{
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  const T kv = b[k] * minv[k];
  T us = u_[k], vs = v_[k];
  T ku, ub, vb;
  if (STAGE == 0)
  {
    ku = vs;
    ub = us;
    vb = vs;
    u0[k] = ub;
    v0[k] = vb;
  }
  else
  {
    ku = vn[k];
    if (STAGE < 3)
    {
      ub = u0[k];
      vb = v0[k];
    }
  }
  u_[k] = ku * bdt + us;
  v_[k] = kv * bdt + vs;
  if (STAGE < 3)
  {
    un[k] = ku * adt_next + ub;
    vn[k] = kv * adt_next + vb;
  }
}

template <typename T>
void launch_stage(int stage, int64_t n, const void* b, const void* minv, void* u_, void* v_,
                  void* u0, void* v0, void* un, void* vn, double bdt, double adt_next,
                  cudaStream_t st)
{
  const unsigned grid = (unsigned)((n + 255) / 256);
#define WFX_STAGE(S)                                                                             \
  rk_stage_kernel<T, S><<<grid, 256, 0, st>>>(n, (const T*)b, (const T*)minv, (T*)u_, (T*)v_,   \
                                               (T*)u0, (T*)v0, (T*)un, (T*)vn, (T)bdt, (T)adt_next)
  switch (stage)
  {
  case 0: WFX_STAGE(0); break;
  case 1: WFX_STAGE(1); break;
  case 2: WFX_STAGE(2); break;
  default: WFX_STAGE(3); break;
  }
#undef WFX_STAGE
  WFX_CUDA(cudaGetLastError());
}
template <typename T>
__global__ void probe_kernel(int64_t n, const int32_t* __restrict__ idx, const T* __restrict__ u, T* __restrict__ out)
{
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = u[idx[i]];
}

// the pending snapshot's copy has landed: hand it to the callback
void flush_snapshot(wfx_wave* w)
{
  if (!w->snap_pending) return;
  WFX_CUDA(cudaEventSynchronize(w->ev_snap_done));
  w->snap_pending = false;
  if (w->snap_fn) w->snap_fn(w->snap_user, w->snap_step, w->snap_t, w->snap_hu, w->snap_hv);
}

// after a completed step: probe record and, every snap_every steps, a snapshot
void after_step(wfx_wave* w, cudaStream_t ws)
{
  const size_t esz = w->dtype == WFX_F64 ? 8 : 4;
  if (w->nprobes && w->probe_rec < w->probe_cap)
  {
    const unsigned grid = (unsigned)((w->nprobes + 255) / 256);
    unsigned char* out = w->d_probe_val.p + (size_t)w->probe_rec * w->nprobes * esz;
    if (w->dtype == WFX_F64)
      probe_kernel<double><<<grid, 256, 0, ws>>>(w->nprobes, w->d_probe_idx.p, (const double*)w->u_.p, (double*)out);
    else
      probe_kernel<float><<<grid, 256, 0, ws>>>(w->nprobes, w->d_probe_idx.p, (const float*)w->u_.p, (float*)out);
    WFX_CUDA(cudaGetLastError());
    w->probe_t.push_back(w->t_now);
    w->probe_rec += 1;
  }
  if (w->snap_every > 0 && w->steps_done % w->snap_every == 0)
  {
    flush_snapshot(w); // the staging buffers are free again (usually long since)
    const size_t nb = (size_t)w->n * esz;
    WFX_CUDA(cudaMemcpyAsync(w->snap_u.p, w->u_.p, nb, cudaMemcpyDeviceToDevice, ws));
    WFX_CUDA(cudaMemcpyAsync(w->snap_v.p, w->v_.p, nb, cudaMemcpyDeviceToDevice, ws));
    WFX_CUDA(cudaEventRecord(w->ev_snap_ready, ws));
    WFX_CUDA(cudaStreamWaitEvent(w->snap_stream, w->ev_snap_ready, 0));
    WFX_CUDA(cudaMemcpyAsync(w->snap_hu, w->snap_u.p, nb, cudaMemcpyDeviceToHost, w->snap_stream));
    WFX_CUDA(cudaMemcpyAsync(w->snap_hv, w->snap_v.p, nb, cudaMemcpyDeviceToHost, w->snap_stream));
    WFX_CUDA(cudaEventRecord(w->ev_snap_done, w->snap_stream));
    // the next snapshot's device copy must not overtake this one's read: ws waits for it then
    w->snap_pending = true;
    w->snap_step = w->steps_done;
    w->snap_t = w->t_now;
  }
}

__global__ void set_source_kernel(double* g_dev, double g0, double g1, double g2, double g3)
{
  g_dev[0] = g0, g_dev[1] = g1, g_dev[2] = g2, g_dev[3] = g3;
}

// One time step of size dt on stream st: the four stages of LinearGLL.hpp:244-270.  g[i] is the
// source amplitude of stage i; with g_from_dev the boundary kernels read it from w->g_dev
// instead (graph capture).
void enqueue_step(wfx_wave* w, double dt, const double* g, bool g_from_dev, cudaStream_t st)
{
  const double a_runge[5] = {0.0, 0.5, 0.5, 1.0, 0.0};
  const double b_runge[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
  auto boundary = [&](int i, const void* vn) {
    if (!w->bnd) return;
    if (g_from_dev) boundary_apply_dev(w->bnd, w->c0, w->g_dev.p + i, vn, w->b.p, st);
    else if (wfx_boundary_apply(w->bnd, w->c0, g[i], vn, w->b.p, st)) fail("%s", wfx_last_error()); // :175
  };
  static const char* const stage_name[4] = {"wfx rk4 stage 0", "wfx rk4 stage 1", "wfx rk4 stage 2", "wfx rk4 stage 3"};
  for (int i = 0; i < 4; ++i)
  {
    NvtxRange stage_range(stage_name[i]);
    // stage state: the solution itself at stage 0 (a_0 = 0), else un/vn
    const void* un = i == 0 ? w->u_.p : w->un.p;
    const void* vn = i == 0 ? w->v_.p : w->vn.p;
    // f1 (:151-192)
    if (!w->halo)
    {
      {
        NvtxRange r("wfx stiffness apply");
        if (wfx_stiffness_apply(w->stiff, un, w->b.p, 0, st)) fail("%s", wfx_last_error()); // :173-174
      }
      boundary(i, vn);
    }
    else
    {
      // distributed: interface cells first, their ghost reduction (:176, which also stands
      // for the scatter_fwd of the next stage, :164,167) on the comm stream while the interior
      // cells run; the boundary term (facet masses assembled over the ranks) is added by
      // every copy of a dof after the reduction.
      static const bool overlap = [] { const char* e = std::getenv("WFX_WAVE_OVERLAP"); return !e || std::atoi(e) != 0; }();
      const bool split = overlap && stiffness_has_split(w->stiff);
      {
        NvtxRange r("wfx stiffness interface part");
        if (wfx_stiffness_apply_part(w->stiff, un, nullptr, w->b.p, 0, split ? 0 : -1, st)) fail("%s", wfx_last_error());
      }
      WFX_CUDA(cudaEventRecord(w->ev_iface, st));
      WFX_CUDA(cudaStreamWaitEvent(w->comm_stream, w->ev_iface, 0));
      {
        NvtxRange r("wfx ghost reduction");
        if (wfx_halo_update_rev_fwd(w->halo, w->b.p, w->comm_stream)) fail("%s", wfx_last_error());
      }
      WFX_CUDA(cudaEventRecord(w->ev_halo, w->comm_stream));
      {
        NvtxRange r("wfx stiffness interior part");
        if (split && wfx_stiffness_apply_part(w->stiff, un, nullptr, w->b.p, 0, 1, st)) fail("%s", wfx_last_error());
      }
      WFX_CUDA(cudaStreamWaitEvent(st, w->ev_halo, 0));
      boundary(i, vn);
    }
    NvtxRange update_range("wfx fused stage update");
    if (w->dtype == WFX_F64)
      launch_stage<double>(i, w->n, w->b.p, w->minv, w->u_.p, w->v_.p, w->u0.p, w->v0.p, w->un.p,
                           w->vn.p, dt * b_runge[i], dt * a_runge[i + 1], st);
    else
      launch_stage<float>(i, w->n, w->b.p, w->minv, w->u_.p, w->v_.p, w->u0.p, w->v0.p, w->un.p,
                          w->vn.p, dt * b_runge[i], dt * a_runge[i + 1], st);
  }
}
} // namespace

extern "C" int wfx_wave_create(wfx_ctx* ctx, wfx_stiffness* stiff, wfx_mass* mass,
                               wfx_boundary* bnd, wfx_halo* halo, int64_t size_local, double c0,
                               double f0, double p0, wfx_wave** out)
{
  WFX_API_BEGIN
  if (!ctx || !stiff || !mass || !out) fail("NULL argument");
  ScopedDevice sd(ctx->device);
  int64_t ndofs = 0;
  if (wfx_stiffness_info(stiff, nullptr, nullptr, &ndofs, nullptr, nullptr, nullptr, nullptr))
    fail("%s", wfx_last_error());
  auto w = std::make_unique<wfx_wave>();
  w->ctx = ctx;
  w->stiff = stiff;
  w->mass = mass;
  w->bnd = bnd;
  w->halo = halo;
  w->n = ndofs;
  w->size_local = size_local;
  w->c0 = c0;
  w->f0 = f0;
  w->p0 = p0;
  w->dtype = stiffness_dtype(stiff);
  if (mass_dtype(mass) != w->dtype) fail("wave: mass operator dtype differs from the stiffness operator's");
  if (bnd && boundary_dtype(bnd) != w->dtype) fail("wave: boundary operator dtype differs from the stiffness operator's");
  if (halo)
  {
    // every stage exchanges the model's right-hand side through this halo: same scalar type or
    // the buffers are reinterpreted
    if (halo_dtype(halo) != w->dtype) fail("wave: halo dtype differs from the model dtype");
    // m.scatter_rev(add) (LinearGLL.hpp:110) and the same for the facet masses (both idempotent).
    // They are summed in fp64: an fp64 model's halo serves, an fp32 model must have assembled them
    // beforehand with a separate fp64 halo (wfx_mass_assemble / wfx_boundary_assemble).
    if (w->dtype == WFX_F64)
    {
      if (wfx_mass_assemble(mass, halo)) fail("%s", wfx_last_error());
      if (bnd && wfx_boundary_assemble(bnd, halo)) fail("%s", wfx_last_error());
    }
    if (!mass_assembled(mass))
      fail("wave: distributed fp32 model needs the mass assembled first (wfx_mass_assemble with an fp64 halo)");
    if (bnd && !boundary_assembled(bnd))
      fail("wave: distributed fp32 model needs the facet masses assembled first (wfx_boundary_assemble with an fp64 halo)");
    // highest priority: the small pack / NCCL / unpack kernels must not queue behind the interior batches
    int prio_lo = 0, prio_hi = 0;
    WFX_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    WFX_CUDA(cudaStreamCreateWithPriority(&w->comm_stream, cudaStreamNonBlocking, prio_hi));
    WFX_CUDA(cudaEventCreateWithFlags(&w->ev_iface, cudaEventDisableTiming));
    WFX_CUDA(cudaEventCreateWithFlags(&w->ev_halo, cudaEventDisableTiming));
  }
  if (const char* e = std::getenv("WFX_WAVE_GRAPH")) w->use_graph = std::atoi(e) != 0;
  if (!halo)
  {
    w->g_dev.alloc(4);
    WFX_CUDA(cudaStreamCreateWithFlags(&w->own_stream, cudaStreamNonBlocking));
    WFX_CUDA(cudaEventCreateWithFlags(&w->ev_in, cudaEventDisableTiming));
    WFX_CUDA(cudaEventCreateWithFlags(&w->ev_out, cudaEventDisableTiming));
  }
  if (wfx_mass_inverse_diagonal(mass, &w->minv)) fail("%s", wfx_last_error());
  const size_t nb = (size_t)ndofs * (w->dtype == WFX_F64 ? 8 : 4);
  for (auto* v : {&w->u_, &w->v_, &w->u0, &w->v0, &w->un, &w->vn, &w->b})
  {
    v->alloc(nb);
    if (nb) WFX_CUDA(cudaMemset(v->p, 0, nb));
  }
  *out = w.release();
  WFX_API_END
}

extern "C" int wfx_wave_init(wfx_wave* w)
{
  WFX_API_BEGIN
  if (!w) fail("wave model is NULL");
  ScopedDevice sd(w->ctx->device);
  if (w->u_.n)
  {
    WFX_CUDA(cudaMemset(w->u_.p, 0, w->u_.n));
    WFX_CUDA(cudaMemset(w->v_.p, 0, w->v_.n));
  }
  w->steps_done = 0;
  w->t_now = 0;
  w->probe_rec = 0;
  w->probe_t.clear();
  WFX_API_END
}

extern "C" int wfx_wave_set_state(wfx_wave* w, const void* u, const void* v)
{
  WFX_API_BEGIN
  if (!w) fail("wave model is NULL");
  ScopedDevice sd(w->ctx->device);
  if (u && w->u_.n) WFX_CUDA(cudaMemcpy(w->u_.p, u, w->u_.n, cudaMemcpyHostToDevice));
  if (v && w->v_.n) WFX_CUDA(cudaMemcpy(w->v_.p, v, w->v_.n, cudaMemcpyHostToDevice));
  w->steps_done = 0;
  w->t_now = 0;
  w->probe_rec = 0;
  w->probe_t.clear();
  // u->scatter_fwd(), v->scatter_fwd() (LinearGLL.hpp:164,167): the stage kernels update ghost
  // entries locally from then on, so the copies are made consistent once, here
  if (w->halo && w->n)
  {
    if (u && wfx_halo_update_fwd(w->halo, w->u_.p, nullptr)) fail("%s", wfx_last_error());
    if (v && wfx_halo_update_fwd(w->halo, w->v_.p, nullptr)) fail("%s", wfx_last_error());
    WFX_CUDA(cudaDeviceSynchronize());
  }
  WFX_API_END
}

extern "C" int wfx_wave_get_state(wfx_wave* w, void* u, void* v)
{
  WFX_API_BEGIN
  if (!w) fail("wave model is NULL");
  ScopedDevice sd(w->ctx->device);
  if (u && w->u_.n) WFX_CUDA(cudaMemcpy(u, w->u_.p, w->u_.n, cudaMemcpyDeviceToHost));
  if (v && w->v_.n) WFX_CUDA(cudaMemcpy(v, w->v_.p, w->v_.n, cudaMemcpyDeviceToHost));
  WFX_API_END
}

extern "C" int wfx_wave_state_ptrs(wfx_wave* w, void** u, void** v)
{
  WFX_API_BEGIN
  if (!w) fail("wave model is NULL");
  if (u) *u = w->u_.p;
  if (v) *v = w->v_.p;
  WFX_API_END
}

namespace
{
// g(t) of LinearGLL.hpp:155-162: Hann ramp over the first alpha periods
double source_amplitude(const wfx_wave* w, double tn)
{
  const double w0 = 2.0 * M_PI * w->f0, T = 1.0 / w->f0, alpha = 4.0; // :96-99
  const double window = tn < T * alpha ? 0.5 * (1.0 - std::cos(w->f0 * M_PI * tn / alpha)) : 1.0;
  return window * w->p0 * w0 / w->c0 * std::cos(w0 * tn);
}
} // namespace

extern "C" int wfx_wave_f0(wfx_wave* w, double, const void* u, const void* v, void* result, void* stream)
{
  WFX_API_BEGIN
  if (!w) fail("wave model is NULL");
  if (!u || !v || !result) fail("f0: NULL vector");
  ScopedDevice sd(w->ctx->device);
  const size_t nb = (size_t)w->n * (w->dtype == WFX_F64 ? 8 : 4);
  if (result != v) WFX_CUDA(cudaMemcpyAsync(result, v, nb, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  WFX_API_END
}

extern "C" int wfx_wave_f1(wfx_wave* w, double t, const void* u, const void* v, void* result, void* stream)
{
  WFX_API_BEGIN
  if (!w) fail("wave model is NULL");
  if (!u || !v || !result) fail("f1: NULL vector");
  if (result == u || result == v) fail("f1: result must not alias u or v");
  ScopedDevice sd(w->ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const double g = source_amplitude(w, t);
  if (wfx_stiffness_apply(w->stiff, u, w->b.p, 0, st)) fail("%s", wfx_last_error());                     // :173-174
  if (w->halo && wfx_halo_update_rev_fwd(w->halo, w->b.p, st)) fail("%s", wfx_last_error());             // :176
  if (w->bnd && wfx_boundary_apply(w->bnd, w->c0, g, v, w->b.p, st)) fail("%s", wfx_last_error());       // :175
  if (wfx_mass_apply_inverse(w->mass, w->b.p, result, st)) fail("%s", wfx_last_error());                 // :188-191
  WFX_API_END
}

extern "C" int wfx_wave_rk4(wfx_wave* w, double t0, double tf, double dt, int64_t max_steps,
                            int64_t* steps_out, double* t_end, void* stream)
{
  WFX_API_BEGIN
  if (!w) fail("wave model is NULL");
  ScopedDevice sd(w->ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const double c_runge[4] = {0.0, 0.5, 0.5, 1.0};
  // Graph replay (one rank only: the distributed step spans two streams and NCCL).  A capture
  // cannot run on the legacy default stream: such callers are moved to an internal stream that
  // is ordered after and before the caller's stream by events.
  const bool graph_ok = w->use_graph && !w->halo && stiffness_graph_safe(w->stiff);
  cudaStream_t ws = st;
  if (graph_ok && (st == nullptr || st == cudaStreamLegacy))
  {
    ws = w->own_stream;
    WFX_CUDA(cudaEventRecord(w->ev_in, st));
    WFX_CUDA(cudaStreamWaitEvent(ws, w->ev_in, 0));
  }
  double t = t0;
  int64_t step = 0;
  const double dt_nominal = dt;
  while (t < tf) // :241
  {
    if (max_steps > 0 && step >= max_steps) break;
    NvtxRange step_range("wfx rk4 step");
    dt = std::min(dt, tf - t); // :242
    double g[4];
    for (int i = 0; i < 4; ++i)
    {
      const double tn = t + c_runge[i] * dt; // :257
      g[i] = source_amplitude(w, tn); // :155-162
    }
    if (graph_ok && dt == dt_nominal)
    {
      if (!w->step_graph || w->graph_dt != dt)
      {
        if (w->step_graph) WFX_CUDA(cudaGraphExecDestroy(w->step_graph));
        w->step_graph = nullptr;
        cudaGraph_t graph = nullptr;
        WFX_CUDA(cudaStreamBeginCapture(ws, cudaStreamCaptureModeThreadLocal));
        try
        {
          enqueue_step(w, dt, nullptr, true, ws);
        }
        catch (...)
        {
          cudaStreamEndCapture(ws, &graph);
          if (graph) cudaGraphDestroy(graph);
          throw;
        }
        WFX_CUDA(cudaStreamEndCapture(ws, &graph));
        const cudaError_t ie = cudaGraphInstantiate(&w->step_graph, graph, 0);
        cudaGraphDestroy(graph);
        WFX_CUDA(ie);
        w->graph_dt = dt;
      }
      set_source_kernel<<<1, 1, 0, ws>>>(w->g_dev.p, g[0], g[1], g[2], g[3]);
      WFX_CUDA(cudaGraphLaunch(w->step_graph, ws));
    }
    else enqueue_step(w, dt, g, false, ws);
    t += dt;
    step += 1;
    w->steps_done += 1;
    w->t_now = t;
    if (w->nprobes || w->snap_every > 0)
    {
      if (w->snap_pending && w->snap_every > 0 && (w->steps_done % w->snap_every) == 0)
        WFX_CUDA(cudaStreamWaitEvent(ws, w->ev_snap_done, 0));
      after_step(w, ws);
    }
  }
  if (w->snap_pending)
  {
    // deliver the last snapshot before returning (the callback may read solver state)
    flush_snapshot(w);
  }
  if (ws != st)
  {
    WFX_CUDA(cudaEventRecord(w->ev_out, ws));
    WFX_CUDA(cudaStreamWaitEvent(st, w->ev_out, 0));
  }
  if (steps_out) *steps_out = step;
  if (t_end) *t_end = t;
  WFX_API_END
}

extern "C" int wfx_wave_set_probes(wfx_wave* w, int64_t nprobes, const int32_t* dofs_host, int64_t max_records)
{
  WFX_API_BEGIN
  if (!w) fail("wave model is NULL");
  if (nprobes < 0 || max_records < 0 || (nprobes > 0 && !dofs_host)) fail("bad probe list");
  for (int64_t i = 0; i < nprobes; ++i)
    if (dofs_host[i] < 0 || dofs_host[i] >= w->n) fail("probe dof %d out of range", dofs_host[i]);
  ScopedDevice sd(w->ctx->device);
  w->nprobes = nprobes;
  w->probe_cap = nprobes ? max_records : 0;
  w->probe_rec = 0;
  w->probe_t.clear();
  if (nprobes)
  {
    w->d_probe_idx.upload(dofs_host, (size_t)nprobes);
    w->d_probe_val.alloc((size_t)nprobes * (size_t)max_records * (w->dtype == WFX_F64 ? 8 : 4));
  }
  WFX_API_END
}

extern "C" int wfx_wave_get_probe_series(wfx_wave* w, int64_t* nrecords, double* t_host, void* values_host)
{
  WFX_API_BEGIN
  if (!w) fail("wave model is NULL");
  ScopedDevice sd(w->ctx->device);
  if (nrecords) *nrecords = w->probe_rec;
  if (t_host)
    for (int64_t r = 0; r < w->probe_rec; ++r) t_host[r] = w->probe_t[r];
  if (values_host && w->probe_rec)
  {
    WFX_CUDA(cudaDeviceSynchronize());
    WFX_CUDA(cudaMemcpy(values_host, w->d_probe_val.p,
                        (size_t)w->probe_rec * w->nprobes * (w->dtype == WFX_F64 ? 8 : 4), cudaMemcpyDeviceToHost));
  }
  WFX_API_END
}

extern "C" int wfx_wave_set_snapshot(wfx_wave* w, int64_t every, wfx_snapshot_fn fn, void* user)
{
  WFX_API_BEGIN
  if (!w) fail("wave model is NULL");
  if (every < 0) fail("snapshot interval must not be negative");
  ScopedDevice sd(w->ctx->device);
  flush_snapshot(w);
  w->snap_every = fn ? every : 0;
  w->snap_fn = fn;
  w->snap_user = user;
  if (w->snap_every > 0 && !w->snap_stream)
  {
    const size_t nb = (size_t)w->n * (w->dtype == WFX_F64 ? 8 : 4);
    w->snap_u.alloc(nb);
    w->snap_v.alloc(nb);
    WFX_CUDA(cudaMallocHost(&w->snap_hu, nb ? nb : 1));
    WFX_CUDA(cudaMallocHost(&w->snap_hv, nb ? nb : 1));
    WFX_CUDA(cudaStreamCreateWithFlags(&w->snap_stream, cudaStreamNonBlocking));
    WFX_CUDA(cudaEventCreateWithFlags(&w->ev_snap_ready, cudaEventDisableTiming));
    WFX_CUDA(cudaEventCreateWithFlags(&w->ev_snap_done, cudaEventDisableTiming));
  }
  WFX_API_END
}

extern "C" int wfx_wave_destroy(wfx_wave* w)
{
  WFX_API_BEGIN
  if (w)
  {
    ScopedDevice sd(w->ctx->device);
    delete w;
  }
  WFX_API_END
}
