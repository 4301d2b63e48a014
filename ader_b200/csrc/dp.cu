// Data-parallel optimiser step over NVLink peer memory (the reference is single-device:
// main.py:96,120,143, so everything here is new).
//
// One process per GPU; every rank holds a full replica of theta and its own gradient of the step
// (rows of the batch are sharded, loss denominators are the GLOBAL counts, so the wanted gradient is
// the plain SUM over ranks).  Instead of  all-reduce(grad) -> Adam on every rank  (2 x 7/8 of the
// gradient over the links AND the full 28 B/param optimiser traffic on every GPU), ONE kernel does
//
//     reduce-scatter (peer loads)  ->  TF1 Adam on the owned slice  ->  all-gather (peer stores)
//
// Rank r owns a contiguous slice of the update index space [item-table rows 1..V | dense parameters]
// (the same index space k_adam walks).  For its slice it loads the gradient of every rank over NVLink
// (ld.global.cv: peer lines must not be served from a stale L1), adds them in rank order 0..N-1
// (fixed order: run-to-run deterministic, and every replica receives the same bits), applies
// adam_update_elem (the very expression the single-GPU kernels evaluate) with its LOCAL m / v --
// the optimiser state of a slice is only ever touched by its owner -- and stores the new theta
// into every replica.  Link traffic per GPU and step: (N-1)/N of the gradient in, (N-1)/N of theta out.
//
// Synchronisation (no host, graph-capturable): a flag block per rank in peer-mapped memory,
//   flags[src] (src < N): arrival counter written by rank `src` (monotonic epoch numbers),
//   EPOCH: this rank's epoch (2 per step), DONE: CTA completion counter, ERR: timeout marker.
// k_dp_arrive (one warp, directly in front of k_dp_adam): publishes "my gradient is complete" (epoch e+1) to every
// rank and waits until all ranks have published (A).  k_dp_adam works, and the last CTA to finish publishes
// e+2 = "I have read your gradients and written your theta" (B).  ader_dp_wait (first kernel of the next step) waits
// for B from every rank before anything overwrites the gradient or reads theta.  Only single-warp kernels ever spin
// (they leave the SMs to whatever the other ranks still have to run -- also when several emulated ranks share one
// GPU in the tests); spins time out (~20 s) and mark ERR instead of hanging the GPU.
#include "common.cuh"
#include <string.h>

namespace ader {
namespace dp {

constexpr int F_EPOCH = 32, F_DONE = 33, F_ERR = 34;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float2 ld_cv2(const float* p) {
  float2 v;
  asm volatile("ld.global.cv.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// wait until flags[src] >= want (signed distance: epochs wrap after 2^31 steps); false on timeout.  The poll itself is a
// relaxed load of LOCAL memory (the peer wrote the flag over the link); one acquire fence follows the successful poll
// (an acquire load per poll and a 64 ns sleep made each cross-GPU barrier cost ~10 us, measured with the peer traffic off).
__device__ __forceinline__ bool wait_flag(const uint32_t* f, uint32_t want, uint32_t* err) {
  if ((int32_t)(ld_relaxed_sys(f) - want) < 0) {
    const unsigned long long t0 = gtime();
    unsigned spins = 0;
    while ((int32_t)(ld_relaxed_sys(f) - want) < 0) {
      if ((++spins & 1023u) == 0 && gtime() - t0 > 20000000000ull) { atomicExch(err, 1u); return false; }
    }
  }
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  return true;
}

struct DpFlags {
  int rank, world;
  uint32_t* flags[ADER_DP_MAX_RANKS];
};

struct DpArgs {
  int rank, world;
  float* theta[ADER_DP_MAX_RANKS];
  const float* grad[ADER_DP_MAX_RANKS];
  uint32_t* flags[ADER_DP_MAX_RANKS];
  float *m, *v;
  int* state;
  // update index space = two element ranges (item-table rows 1..V, dense parameters), walked in QUADS of four
  // floats so that peer traffic moves in 16-byte accesses.  A range starts on an even element; when its byte offset is
  // 8 mod 16 the first quad holds only its upper float2 (ph = 1), and a range may end with a half quad.
  long long e0[2];            // first element of the range
  long long n2[2];            // float2 units in the range
  int ph[2];                  // 1: unit 0 sits in the upper half of quad 0
  long long q0[2];            // first global quad of the range (q0[0] = 0), q_total = all quads
  long long q_lo, q_hi;       // quads owned by this rank
  float lr, beta1, beta2, eps, ewc_lambda;
  const float *fisher, *theta_star;
  int fused_arrive;
};

__device__ __forceinline__ float4 ld_cg4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_cg2(const float* p) {
  float2 v;
  asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

// W = upper bound of the world size (array extents in registers), U = quads per thread and trip: U * W peer loads of
// 16 bytes are in flight per thread before the first add (NVLink round trips are ~2-3 us: the link only fills with
// megabytes outstanding, and only with 16-byte accesses -- 8-byte ones reached ~150 GB/s per direction).
// Peer lines are never in this SM's L1 at kernel start (L1 is invalidated at launch boundaries) and .cg keeps them out.
template <int W, int U>
struct DpTrip {
  long long el[U];           // element of the quad's first float (16-byte aligned)
  int msk[U];                // bit 0: lower float2 valid, bit 1: upper float2 valid; 0: no quad
  float4 gr[U][W];
};

template <int W, int U>
__device__ __forceinline__ void dp_load(const DpArgs& a, long long base, long long stride, DpTrip<W, U>& t) {
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const long long q = base + (long long)u * stride;
    t.el[u] = 0; t.msk[u] = 0;
    if (q < a.q_hi) {
      const int s = q >= a.q0[1] ? 1 : 0;
      const long long u0 = 2 * (q - a.q0[s]) - a.ph[s];                 // unit of the quad's lower half
      t.el[u] = a.e0[s] + 2 * u0;
      t.msk[u] = ((u0 >= 0) ? 1 : 0) | ((u0 + 1 < a.n2[s]) ? 2 : 0);
    }
#pragma unroll
    for (int r = 0; r < W; ++r) {
      if (r < a.world && t.msk[u]) {
        if (t.msk[u] == 3) t.gr[u][r] = ld_cg4(a.grad[r] + t.el[u]);
        else {
          const float2 h = ld_cg2(a.grad[r] + t.el[u] + (t.msk[u] == 2 ? 2 : 0));
          t.gr[u][r] = (t.msk[u] == 2) ? make_float4(0.f, 0.f, h.x, h.y) : make_float4(h.x, h.y, 0.f, 0.f);
        }
      }
    }
  }
}

template <int W, int U>
__device__ __forceinline__ void dp_apply(const DpArgs& a, const DpTrip<W, U>& t, float lr_t, bool ewc) {
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!t.msk[u]) continue;
    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < W; ++r)
      if (r < a.world) {                                                    // rank order: deterministic
        g[0] = __fadd_rn(g[0], t.gr[u][r].x); g[1] = __fadd_rn(g[1], t.gr[u][r].y);
        g[2] = __fadd_rn(g[2], t.gr[u][r].z); g[3] = __fadd_rn(g[3], t.gr[u][r].w);
      }
    const int lo = (t.msk[u] & 1) ? 0 : 2, hi = (t.msk[u] & 2) ? 4 : 2;
    float th[4], m[4], v[4];
#pragma unroll
    for (int h = 0; h < 4; h += 2) {
      if (h < lo || h >= hi) continue;
      const long long x = t.el[u] + h;
      const float2 t2 = *reinterpret_cast<const float2*>(a.theta[a.rank] + x);
      const float2 m2 = *reinterpret_cast<const float2*>(a.m + x);
      const float2 v2 = *reinterpret_cast<const float2*>(a.v + x);
      float2 f = make_float2(0.f, 0.f), ts = make_float2(0.f, 0.f);
      if (ewc) {
        f = *reinterpret_cast<const float2*>(a.fisher + x);
        ts = *reinterpret_cast<const float2*>(a.theta_star + x);
      }
      th[h] = t2.x; th[h + 1] = t2.y; m[h] = m2.x; m[h + 1] = m2.y; v[h] = v2.x; v[h + 1] = v2.y;
      adam_update_elem(g[h], th[h], m[h], v[h], lr_t, a.beta1, a.beta2, a.eps, a.ewc_lambda, f.x, ts.x);
      adam_update_elem(g[h + 1], th[h + 1], m[h + 1], v[h + 1], lr_t, a.beta1, a.beta2, a.eps, a.ewc_lambda, f.y, ts.y);
      *reinterpret_cast<float2*>(a.m + x) = make_float2(m[h], m[h + 1]);
      *reinterpret_cast<float2*>(a.v + x) = make_float2(v[h], v[h + 1]);
    }
#pragma unroll
    for (int r = 0; r < W; ++r) {
      if (r >= a.world) continue;
      if (t.msk[u] == 3) *reinterpret_cast<float4*>(a.theta[r] + t.el[u]) = make_float4(th[0], th[1], th[2], th[3]);
      else *reinterpret_cast<float2*>(a.theta[r] + t.el[u] + lo) = make_float2(th[lo], th[lo + 1]);
    }
  }
}

// W = upper bound of the world size (array extents in registers), U = quads per thread and trip.  Software pipeline over
// the trips of a thread: the peer LOADS of trip k+1 are issued before trip k is summed, updated and STORED into the
// peers, so the inbound direction of the link (gradient loads) and the outbound one (theta stores, and the replies to
// the peers' loads) are busy at the same time; the host sizes the grid for ~4 trips per thread with >= 2.5 MB of peer
// loads in flight.  Peer accesses are 16 bytes wide.  Peer lines are never in this SM's L1 at kernel start (L1 is
// invalidated at launch boundaries) and .cg keeps them out.
template <int W, int U>
__global__ void __launch_bounds__(256) k_dp_adam(DpArgs a) {
  __shared__ float s_lr;
  __shared__ uint32_t s_epoch;
  uint32_t* myf = a.flags[a.rank];
  if (threadIdx.x == 0) {
    s_epoch = ld_acquire_sys(myf + F_EPOCH);
    const int t = a.state[0] + 1;
    const double lr_t = (double)a.lr * sqrt(1.0 - pow((double)a.beta2, (double)t)) / (1.0 - pow((double)a.beta1, (double)t));
    s_lr = (float)lr_t;
  }
  __syncthreads();
  const uint32_t e = s_epoch;
  if (a.fused_arrive) {        // barrier A inside this kernel (one launch less on the chain): CTA 0 publishes, every CTA waits
    if (blockIdx.x == 0 && (int)threadIdx.x < a.world) st_release_sys(a.flags[threadIdx.x] + a.rank, e + 1);
    if ((int)threadIdx.x < a.world) wait_flag(myf + threadIdx.x, e + 1, myf + F_ERR);
    __syncthreads();
  }

  const float lr_t = s_lr;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long step = stride * U;
  const bool ewc = a.ewc_lambda != 0.f;
  long long base = a.q_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  DpTrip<W, U> t0, t1;
  if (base < a.q_hi) dp_load<W, U>(a, base, stride, t0);
  while (base < a.q_hi) {
    if (base + step < a.q_hi) dp_load<W, U>(a, base + step, stride, t1);
    dp_apply<W, U>(a, t0, lr_t, ewc);
    base += step;
    if (base >= a.q_hi) break;
    if (base + step < a.q_hi) dp_load<W, U>(a, base + step, stride, t0);
    dp_apply<W, U>(a, t1, lr_t, ewc);
    base += step;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                            // this CTA's peer stores are performed before its ticket
    const uint32_t done = atomicAdd(myf + F_DONE, 1u);
    if (done == gridDim.x - 1) {                       // last CTA: publish B and advance the local state
      myf[F_DONE] = 0u;
      a.state[0] = a.state[0] + 1;
      a.state[1] = __float_as_int(lr_t);
      st_release_sys(myf + F_EPOCH, e + 2);
      __threadfence_system();
      for (int r = 0; r < a.world; ++r) st_release_sys(a.flags[r] + a.rank, e + 2);
    }
  }
}

// A: stream order puts every gradient write of this rank before this kernel
__global__ void k_dp_arrive(DpFlags a) {
  uint32_t* myf = a.flags[a.rank];
  const uint32_t e = ld_acquire_sys(myf + F_EPOCH);
  if ((int)threadIdx.x < a.world) {
    st_release_sys(a.flags[threadIdx.x] + a.rank, e + 1);       // kernel boundary: the gradient is already in device memory
    wait_flag(myf + threadIdx.x, e + 1, myf + F_ERR);
  }
}

__global__ void k_dp_wait(int rank, int world, uint32_t* myf) {
  const uint32_t e = ld_acquire_sys(myf + F_EPOCH);
  if ((int)threadIdx.x < world) wait_flag(myf + threadIdx.x, e, myf + F_ERR);
}

}  // namespace dp
}  // namespace ader

using namespace ader;
using namespace ader::dp;

// CUDA loads kernels lazily, and a first-use load may wait for the context to drain: it must not happen while an
// arrive kernel is already spinning (several emulated ranks in one process would stall until the timeout), so
// every kernel of this file is loaded before the first launch of any of them.
static void dp_preload() {
  static bool done = false;
  if (done) return;
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, k_dp_adam<2, 2>);
  cudaFuncGetAttributes(&fa, k_dp_adam<4, 1>);
  cudaFuncGetAttributes(&fa, k_dp_adam<8, 1>);
  cudaFuncGetAttributes(&fa, k_dp_adam<16, 1>);
  cudaFuncGetAttributes(&fa, k_dp_arrive);
  cudaFuncGetAttributes(&fa, k_dp_wait);
  cudaGetLastError();
  done = true;
}

static int check_comm(const AderDpComm* c) {
  dp_preload();
  ADER_CHECK_ARG(c, "dp: comm is NULL");
  ADER_CHECK_ARG(c->world >= 1 && c->world <= ADER_DP_MAX_RANKS && c->rank >= 0 && c->rank < c->world, "dp: bad rank %d / world %d", c->rank, c->world);
  for (int r = 0; r < c->world; ++r) ADER_CHECK_ARG(c->theta[r] && c->grad[r] && c->flags[r], "dp: NULL pointer for rank %d", r);
  return 0;
}

extern "C" int32_t ader_dp_wait(const AderDpComm* c, void* stream) {
  if (int e = check_comm(c)) return e;
  k_dp_wait<<<1, 32, 0, (cudaStream_t)stream>>>(c->rank, c->world, c->flags[c->rank]);
  ADER_CHECK_LAUNCH("dp_wait");
  return 0;
}

extern "C" int32_t ader_dp_adam_step(const AderModel* m, const AderDpComm* c, float* adam_m, float* adam_v, int32_t* state,
                                     const AderAdamArgs* a, void* stream) {
  if (int e = check_model(m)) return e;
  if (int e = check_comm(c)) return e;
  ADER_CHECK_ARG(adam_m && adam_v && state && a, "dp_adam_step: NULL pointer");
  ADER_CHECK_ARG(a->V >= 1 && a->V < m->v_tab, "dp_adam_step: max_item %d outside table", a->V);
  ADER_CHECK_ARG(a->ewc_lambda == 0.f || (a->fisher && a->theta_star), "dp_adam_step: EWC needs fisher and theta_star");
  const Layout l = make_layout(m);
  DpArgs d;
  d.rank = c->rank; d.world = c->world;
  for (int r = 0; r < ADER_DP_MAX_RANKS; ++r) {
    d.theta[r] = r < c->world ? c->theta[r] : nullptr;
    d.grad[r] = r < c->world ? c->grad[r] : nullptr;
    d.flags[r] = r < c->world ? c->flags[r] : nullptr;
  }
  d.m = adam_m; d.v = adam_v; d.state = state;
  // two element ranges, both starting and ending on even elements (d is even); quads are 16-byte aligned float4s
  const long long starts[2] = {(long long)m->d, l.off_pos};
  const long long counts[2] = {(long long)a->V * m->d, l.dense_count()};
  long long q = 0;
  for (int s = 0; s < 2; ++s) {
    d.e0[s] = starts[s]; d.n2[s] = counts[s] / 2;
    d.ph[s] = (int)((starts[s] / 2) & 1);                    // byte offset 8 mod 16 <=> odd float2 index
    d.q0[s] = q;
    q += (d.n2[s] + d.ph[s] + 1) / 2;
  }
  const long long q_total = q;
  d.q_lo = q_total * c->rank / c->world;
  d.q_hi = q_total * (c->rank + 1) / c->world;
  d.lr = a->lr; d.beta1 = a->beta1; d.beta2 = a->beta2; d.eps = a->eps; d.ewc_lambda = a->ewc_lambda;
  d.fisher = a->fisher; d.theta_star = a->theta_star;
  ADER_CHECK_ARG(((uintptr_t)adam_m % 16) == 0 && ((uintptr_t)adam_v % 16) == 0, "dp_adam_step: optimiser slots must be 16-byte aligned");
  for (int r = 0; r < c->world; ++r)
    ADER_CHECK_ARG(((uintptr_t)c->theta[r] % 16) == 0 && ((uintptr_t)c->grad[r] % 16) == 0, "dp_adam_step: theta / grad of rank %d must be 16-byte aligned", r);
  const long long mine = d.q_hi - d.q_lo;
  const int U = c->world <= 2 ? 2 : 1;
  int grid = cdiv(mine > 0 ? mine : 1, 256 * U * 4);  // ~4 pipelined trips per thread
  const long long per_thread = (long long)U * (c->world > 1 ? c->world - 1 : 1) * 16;
  const int min_grid = cdiv(2500000, 256 * per_thread);     // >= 2.5 MB of peer loads in flight (link bandwidth x latency)
  if (grid < min_grid) grid = min_grid;
  if (grid > 148 * 2) grid = 148 * 2;
  if ((long long)grid * 256 * U > mine && mine > 0) grid = cdiv(mine, 256 * U);
  DpFlags fl;
  fl.rank = c->rank; fl.world = c->world;
  for (int r = 0; r < ADER_DP_MAX_RANKS; ++r) fl.flags[r] = d.flags[r];
  cudaStream_t st = (cudaStream_t)stream;
  // separate_arrive != 0: barrier A as its own single-warp kernel, so that no multi-CTA grid ever spins (several emulated
  // ranks sharing ONE GPU in the tests); real ranks fold it into the update kernel
  d.fused_arrive = c->separate_arrive ? 0 : 1;
  if (c->separate_arrive) k_dp_arrive<<<1, 32, 0, st>>>(fl);
  if (c->world <= 2) k_dp_adam<2, 2><<<grid, 256, 0, st>>>(d);
  else if (c->world <= 4) k_dp_adam<4, 1><<<grid, 256, 0, st>>>(d);
  else if (c->world <= 8) k_dp_adam<8, 1><<<grid, 256, 0, st>>>(d);
  else k_dp_adam<16, 1><<<grid, 256, 0, st>>>(d);
  ADER_CHECK_LAUNCH("dp_adam_step");
  return 0;
}

extern "C" int32_t ader_dp_status(const AderDpComm* c, int32_t* err_out, uint32_t* epoch_out) {
  if (int e = check_comm(c)) return e;
  uint32_t w[3] = {0, 0, 0};
  if (cudaMemcpy(w, c->flags[c->rank] + F_EPOCH, sizeof(w), cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail(-3, "dp_status: %s", cudaGetErrorString(cudaGetLastError()));
  if (epoch_out) *epoch_out = w[0];
  if (err_out) *err_out = (int32_t)w[2];
  return 0;
}

// ---- CUDA IPC plumbing: map another process's device allocation (one process per GPU) -------------------
typedef int (*cuMemGetAddressRange_t)(unsigned long long*, size_t*, unsigned long long);

extern "C" int32_t ader_ipc_export(const void* dev_ptr, void* handle64, int64_t* offset) {
  ADER_CHECK_ARG(dev_ptr && handle64 && offset, "ipc_export: NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)
    return fail(-3, "ipc_export: cuMemGetAddressRange unavailable");
  unsigned long long base = 0; size_t size = 0;
  if (((cuMemGetAddressRange_t)fn)(&base, &size, (unsigned long long)(uintptr_t)dev_ptr) != 0)
    return fail(-3, "ipc_export: not a device allocation");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(-3, "ipc_export: %s", cudaGetErrorString(e)); }
  memcpy(handle64, &h, 64);
  *offset = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
  return 0;
}

extern "C" int32_t ader_ipc_open(const void* handle64, void** base_ptr) {
  ADER_CHECK_ARG(handle64 && base_ptr, "ipc_open: NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(base_ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(-3, "ipc_open: %s", cudaGetErrorString(e)); }
  return 0;
}

extern "C" int32_t ader_ipc_close(void* base_ptr) {
  if (!base_ptr) return 0;
  cudaError_t e = cudaIpcCloseMemHandle(base_ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(-3, "ipc_close: %s", cudaGetErrorString(e)); }
  return 0;
}
