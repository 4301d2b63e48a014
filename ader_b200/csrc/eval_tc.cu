// Full-vocabulary evaluation on the 5th-gen tensor cores with the [R, V] score matrix never written
// (ADER.py:99-103 `pred_last = argsort(argsort(-test_logits))` + util.py:323-339: the metrics only read
// rank(gt) = #{j : s_j > s_gt} + #{j < gt : s_j == s_gt}).
//
// Filter and refine -- ranks come out IDENTICAL to the exact fp32 path (ader_eval_rank_topk):
//   1. exact anchor: s_gt[i] = the fp32 score of the ground-truth item, computed with the very fmaf chain of the exact
//      logits kernel (sgemm.cu: acc = fmaf(a_k, b_k, acc), k ascending), one thread per row;
//   2. tcgen05 pass: approximate scores S~ = rep . E^T from a two-term bf16 split of both operands
//      (x = x_hi + x_lo + r, |r| <= 2^-18 |x|) as three products hi.hi + hi.lo + lo.hi accumulated in one TMEM tile:
//      |S~_ij - s_ij| <= eps_i := C_ERR * ||rep_i|| * max_j ||E_j||  (split remainder 1.2e-5, fp32 accumulation of 480
//      terms and the exact path's own rounding together stay below 1e-4 of sum |a_k b_k| <= ||a|| ||b||; C_ERR = 2^-12);
//      the epilogue counts  S~_ij > s_gt + eps_i  (certainly above) and appends the few columns with
//      |S~_ij - s_gt| <= eps_i to a per-row candidate list;
//   3. refine: candidates are re-scored with the exact fmaf chain and compared exactly (ties -> lower index first).
// A row whose candidate list overflows (CAP entries) raises the overflow flag; the host then takes the exact path.
//
// Operands: hi | lo halves side by side in one row-major bf16 matrix [rows][320] (320 = 5 swizzle spans of 64: the
// span that holds hi[128..159] also holds lo[0..31], so a tile is 5 regions instead of 6).  Tiles reach shared memory by
// TMA (SWIZZLE_128B) and are K-major UMMA operands; a k-step of either half is addressed by its column.
//   X (stationary): 128 rows of rep      = 5 regions x 16 KB = 80 KB
//   Y (streamed)  :  64 rows of the table = 5 regions x  8 KB = 40 KB, ring of 3
//   S tile 128 x 64 fp32 in TMEM (2 buffers x 64 columns), 30 MMAs (M = 128, N = 64, K = 16) per tile.
// Work: CTA = (row tile, vocabulary chunk); counts are merged with integer atomics (order independent).
#include "tc_common.cuh"
#include <stdlib.h>

namespace ader {
namespace ev {
using namespace ader::tc;

constexpr int D2 = 320;                 // hi | lo columns of the operand matrices
constexpr int HALF = 160;               // column of lo[0]
constexpr int NREG = 5;
constexpr int XREG = 128 * 128, YREG = 64 * 128;
constexpr int XBYTES = NREG * XREG, YBYTES = NREG * YREG;
constexpr int NST = 3;
constexpr int TM = 128, TN = 64;
constexpr int NTHREADS = 320, NEPI = 256;
constexpr int SMEM_EVAL = 1024 + XBYTES + NST * YBYTES + 256;
constexpr float C_ERR = 1.0f / 4096.0f;

struct EvalArgs {
  int R, V, n_mtiles, n_vtiles, n_chunks;
  const float* lo_thr;       // [R] s_gt - eps
  const float* hi_thr;       // [R] s_gt + eps
  int* above;                // [R] certainly-above count (atomicAdd)
  int* cand_cnt;             // [R]
  int* cand;                 // [R][cap]
  int cap;
  int* flags;                // [0] candidate overflow, [1] pipeline timeout
  // top-k (modes MAX / BOTH)
  float* maxbuf;             // MAX: [n_chunks * 2][R] largest approximate score seen by each (chunk, column half) of a row
  const float* lo_topk;      // BOTH: [R] columns with S~ >= lo_topk[i] are top-k candidates
  int* tk_cnt;               // [R]
  int* tk_cand;              // [R][cap]
};
enum { EV_RANK = 0, EV_MAX = 1, EV_BOTH = 2 };

// fp32 rows -> [rows_pad][320] bf16 (hi | lo), optional row norms and their maximum (as int bits of a non-negative float)
__global__ void __launch_bounds__(256) k_pack_hilo(const float* __restrict__ src, long long ld, int n_rows, int d, int rows_pad,
                                                   __nv_bfloat16* __restrict__ out, float* __restrict__ norm, int* __restrict__ max_norm) {
  const int row = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= rows_pad) return;
  float ss = 0.f;
  for (int c = lane; c < HALF; c += 32) {
    float x = 0.f;
    if (row < n_rows && c < d) x = src[(long long)row * ld + c];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    out[(long long)row * D2 + c] = h;
    out[(long long)row * D2 + HALF + c] = l;
    ss = fmaf(x, x, ss);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0 && row < n_rows) {
    const float nrm = sqrtf(ss) * 1.0001f;            // rounding of the norm itself must not shrink the bound
    if (norm) norm[row] = nrm;
    if (max_norm) atomicMax(max_norm, __float_as_int(nrm));
  }
}

// exact score of the ground-truth item (the fmaf chain of sgemm.cu) and the certainty band around it
__global__ void __launch_bounds__(256) k_gt_band(const float* __restrict__ table1, const float* __restrict__ rep, const int* __restrict__ gt,
                                                 int R, int d, const float* __restrict__ rnorm, const int* __restrict__ max_norm,
                                                 float* __restrict__ sg, float* __restrict__ lo_thr, float* __restrict__ hi_thr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  const float* a = rep + (long long)i * d;
  const float* b = table1 + (long long)(gt[i] - 1) * d;
  float acc = 0.f;
  for (int k = 0; k < d; ++k) acc = fmaf(a[k], b[k], acc);
  const float eps = C_ERR * rnorm[i] * __int_as_float(*max_norm);
  sg[i] = acc; lo_thr[i] = acc - eps; hi_thr[i] = acc + eps;
}

// K-major SW128 descriptor of the 16-column k-step that starts at matrix column `col` of a tile whose regions are `reg` bytes
__device__ __forceinline__ uint64_t desc_col(uint32_t base, int col, int reg) {
  return make_desc_sw128(base + (col >> 6) * reg + (col & 63) * 2, 16, 1024);
}

template <int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) k_eval_tc(EvalArgs a, const __grid_constant__ CUtensorMap tm_x,
                                                         const __grid_constant__ CUtensorMap tm_y) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sX = smem;
  uint8_t* sY = smem + XBYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + XBYTES + NST * YBYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  const uint32_t bar0 = smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * i; };
  constexpr int B_XFULL = 0, B_YFULL = 1, B_YEMPTY = 4, B_TFULL = 7, B_TEMPTY = 9;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int* err = a.flags + 1;

  const int x_tile = blockIdx.x % a.n_mtiles, chunk = blockIdx.x / a.n_mtiles;
  const int y_lo = (int)((long long)chunk * a.n_vtiles / a.n_chunks);
  const int y_hi = (int)((long long)(chunk + 1) * a.n_vtiles / a.n_chunks);
  const int n_it = max(0, y_hi - y_lo);

  if (threadIdx.x == 0) {
    mbar_init(BAR(B_XFULL), 1);
    for (int s = 0; s < NST; ++s) { mbar_init(BAR(B_YFULL + s), 1); mbar_init(BAR(B_YEMPTY + s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(B_TFULL + s), 1); mbar_init(BAR(B_TEMPTY + s), NEPI); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_y) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0 && n_it > 0) {
      mbar_expect_tx(BAR(B_XFULL), XBYTES);
#pragma unroll
      for (int j = 0; j < NREG; ++j) tma_load_2d(smem_u32(sX) + j * XREG, &tm_x, 64 * j, x_tile * TM, BAR(B_XFULL));
      for (int it = 0; it < n_it; ++it) {
        const int ys = it % NST; const uint32_t yph = (it / NST) & 1;
        mbar_wait(BAR(B_YEMPTY + ys), yph ^ 1, err);
        mbar_expect_tx(BAR(B_YFULL + ys), YBYTES);
#pragma unroll
        for (int j = 0; j < NREG; ++j)
          tma_load_2d(smem_u32(sY + ys * YBYTES) + j * YREG, &tm_y, 64 * j, (y_lo + it) * TN, BAR(B_YFULL + ys));
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && n_it > 0) {
      constexpr uint32_t IDESC = make_idesc(TM, TN, 0, 0);
      const uint32_t xa = smem_u32(sX);
      mbar_wait(BAR(B_XFULL), 0, err);
      for (int it = 0; it < n_it; ++it) {
        const int s = it & 1; const uint32_t ph = (it >> 1) & 1;
        const int ys = it % NST; const uint32_t yph = (it / NST) & 1;
        mbar_wait(BAR(B_YFULL + ys), yph, err);
        mbar_wait(BAR(B_TEMPTY + s), ph ^ 1, err);
        tc_fence_after();
        const uint32_t ya = smem_u32(sY + ys * YBYTES);
        // hi.hi + hi.lo + lo.hi (lo.lo is below the error bound): 3 x 10 k-steps into one accumulator
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const int ac = (p == 2) ? HALF : 0, bc = (p == 1) ? HALF : 0;
#pragma unroll
          for (int k = 0; k < 10; ++k)
            umma_bf16(tmem + s * TN, desc_col(xa, ac + 16 * k, XREG), desc_col(ya, bc + 16 * k, YREG), IDESC, (p | k) != 0);
        }
        umma_commit(BAR(B_YEMPTY + ys));
        umma_commit(BAR(B_TFULL + s));
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;      // TMEM lane quarter, which 32 of the tile's 64 columns
    const int row = q * 32 + lane, gi = x_tile * TM + row;
    const uint32_t tlane = (uint32_t)(q * 32) << 16;
    const bool live = gi < a.R;
    const bool ranking = MODE != EV_MAX;
    const float lo = (live && ranking) ? a.lo_thr[gi] : INFINITY, hi = (live && ranking) ? a.hi_thr[gi] : INFINITY;
    const float lo_k = (live && MODE == EV_BOTH) ? a.lo_topk[gi] : INFINITY;
    float vmax = -INFINITY;
    int above = 0;
    for (int it = 0; it < n_it; ++it) {
      const int s = it & 1; const uint32_t ph = (it >> 1) & 1;
      mbar_wait(BAR(B_TFULL + s), ph, err);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem + tlane + s * TN + half * 32, r);
      tc_fence_before();
      mbar_arrive(BAR(B_TEMPTY + s));
      const int j0 = (y_lo + it) * TN + half * 32;
      if (!live) continue;
      const bool whole = j0 + 32 <= a.V;          // else: zero padding rows of the table matrix are not items
      if (MODE == EV_MAX) {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (whole || j0 + i < a.V) vmax = fmaxf(vmax, __uint_as_float(r[i]));
        continue;
      }
      unsigned unsure = 0u, topc = 0u;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float v = __uint_as_float(r[i]);
        const bool ok = whole || j0 + i < a.V;
        above += (ok && v > hi) ? 1 : 0;
        unsure |= (ok && v >= lo && v <= hi) ? (1u << i) : 0u;
        if (MODE == EV_BOTH) topc |= (ok && v >= lo_k) ? (1u << i) : 0u;
      }
      while (unsure) {                            // rare: a handful of columns per row in the whole pass
        const int i = __ffs(unsure) - 1; unsure &= unsure - 1;
        const int pos = atomicAdd(a.cand_cnt + gi, 1);
        if (pos < a.cap) a.cand[(long long)gi * a.cap + pos] = j0 + i; else atomicExch(a.flags, 1);
      }
      while (topc) {
        const int i = __ffs(topc) - 1; topc &= topc - 1;
        const int pos = atomicAdd(a.tk_cnt + gi, 1);
        if (pos < a.cap) a.tk_cand[(long long)gi * a.cap + pos] = j0 + i; else atomicExch(a.flags, 1);
      }
    }
    if (live && MODE == EV_MAX) a.maxbuf[(long long)(chunk * 2 + half) * a.R + gi] = vmax;
    if (live && ranking && above) atomicAdd(a.above + gi, above);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
  }
}

// exact re-score of the candidates of a row (warp per row, lane per candidate): rank = certainly-above + exactly-above
__global__ void __launch_bounds__(256) k_refine(const float* __restrict__ table1, const float* __restrict__ rep, const int* __restrict__ gt,
                                                int R, int d, const float* __restrict__ sg, const int* __restrict__ above,
                                                const int* __restrict__ cand_cnt, const int* __restrict__ cand, int cap,
                                                int* __restrict__ rank) {
  const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= R) return;
  const int n = min(cand_cnt[i], cap), g = gt[i] - 1;
  const float s0 = sg[i];
  const float* a = rep + (long long)i * d;
  int c = 0;
  for (int t = lane; t < n; t += 32) {
    const int j = cand[(long long)i * cap + t];
    if (j == g) continue;
    const float* b = table1 + (long long)j * d;
    float acc = 0.f;
    for (int k = 0; k < d; ++k) acc = fmaf(a[k], b[k], acc);
    c += (acc > s0) || (acc == s0 && j < g);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) rank[i] = above[i] + c;
}

// tau = the k-th largest of a row's local maxima (each belongs to a different item, so at least k items score >= tau - eps);
// every top-k item then has an approximate score >= tau - 2 eps.  Warp per row.
__global__ void __launch_bounds__(256) k_topk_threshold(const float* __restrict__ maxbuf, int n_slots, int R, int k,
                                                        const float* __restrict__ rnorm, const int* __restrict__ max_norm,
                                                        float* __restrict__ lo_topk) {
  const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= R) return;
  float v[8];                                     // n_slots <= 256
#pragma unroll
  for (int q = 0; q < 8; ++q) { const int sidx = lane + 32 * q; v[q] = sidx < n_slots ? maxbuf[(long long)sidx * R + i] : -INFINITY; }
  float kth = -INFINITY;
  for (int round = 0; round < k; ++round) {       // k rounds of "remove the current maximum"
    float best = -INFINITY; int bq = -1;
#pragma unroll
    for (int q = 0; q < 8; ++q) if (v[q] > best) { best = v[q]; bq = q; }
    float wb = best;
#pragma unroll
    for (int o = 16; o; o >>= 1) wb = fmaxf(wb, __shfl_xor_sync(0xffffffffu, wb, o));
    const unsigned who = __ballot_sync(0xffffffffu, best == wb && bq >= 0);
    if (who && lane == __ffs(who) - 1) {
#pragma unroll
      for (int q = 0; q < 8; ++q) if (q == bq) v[q] = -INFINITY;
    }
    kth = wb;
  }
  if (lane == 0) lo_topk[i] = kth - 2.f * C_ERR * rnorm[i] * __int_as_float(*max_norm);
}

// exact re-score of the top-k candidates of a row and selection of the k best (score descending, ties -> lower index):
// warp per row, up to 8 candidates per lane (cap = 256)
__global__ void __launch_bounds__(256) k_topk_refine(const float* __restrict__ table1, const float* __restrict__ rep, int R, int d,
                                                     const int* __restrict__ tk_cnt, const int* __restrict__ tk_cand, int cap, int k,
                                                     int* __restrict__ topk_item, float* __restrict__ topk_score) {
  const int i = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (i >= R) return;
  const int n = min(tk_cnt[i], cap);
  const float* a = rep + (long long)i * d;
  float sc[8]; int id[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int t = lane + 32 * q;
    sc[q] = -INFINITY; id[q] = 0x7fffffff;
    if (t < n) {
      const int j = tk_cand[(long long)i * cap + t];
      const float* b = table1 + (long long)j * d;
      float acc = 0.f;
      for (int c = 0; c < d; ++c) acc = fmaf(a[c], b[c], acc);
      sc[q] = acc; id[q] = j;
    }
  }
  for (int round = 0; round < k; ++round) {
    float bs = -INFINITY; int bi = 0x7fffffff, bq = -1;
#pragma unroll
    for (int q = 0; q < 8; ++q) if (id[q] != 0x7fffffff && (sc[q] > bs || (sc[q] == bs && id[q] < bi))) { bs = sc[q]; bi = id[q]; bq = q; }
    float ws = bs; int wi = bi;
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float os = __shfl_xor_sync(0xffffffffu, ws, o); const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (os > ws || (os == ws && oi < wi)) { ws = os; wi = oi; }
    }
    if (bq >= 0 && bi == wi) {
#pragma unroll
      for (int q = 0; q < 8; ++q) if (q == bq) id[q] = 0x7fffffff;
    }
    if (lane == 0) {
      const bool ok = wi != 0x7fffffff;
      topk_item[(long long)i * k + round] = ok ? wi + 1 : 0;
      topk_score[(long long)i * k + round] = ok ? ws : -INFINITY;
    }
  }
}

static int eval_chunks(int n_mtiles, int n_vtiles) {
  int best = 1; double best_eff = 0.0;
  const int cmax = n_vtiles / 8 > 1 ? (n_vtiles / 8 < 128 ? n_vtiles / 8 : 128) : 1;      // >= 8 streamed tiles per CTA, <= 256 top-k slots per row
  for (int c = 1; c <= cmax; ++c) {
    const long long ctas = (long long)n_mtiles * c;
    const long long waves = (ctas + 147) / 148;
    const double eff = (double)ctas / (148.0 * waves);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = c; }
  }
  return best;
}

struct EvalWs {
  __nv_bfloat16 *x16, *y16;
  float *rnorm, *sg, *lo, *hi;
  int *max_norm, *above, *cand_cnt, *cand, *flags, *tk_cnt, *tk_cand;
  float *lo_topk, *maxbuf;
  size_t bytes;
};
static EvalWs carve(int R, int V, int cap, char* base) {
  EvalWs w; size_t o = 0;
  auto take = [&](size_t n) { char* p = base ? base + o : nullptr; o += align_up(n); return p; };
  const int nm = cdiv(R, TM), nv = cdiv(V, TN);
  w.x16 = (__nv_bfloat16*)take((size_t)nm * TM * D2 * 2);
  w.y16 = (__nv_bfloat16*)take((size_t)nv * TN * D2 * 2);
  w.rnorm = (float*)take(sizeof(float) * R);
  w.sg = (float*)take(sizeof(float) * R);
  w.lo = (float*)take(sizeof(float) * R);
  w.hi = (float*)take(sizeof(float) * R);
  // zeroed per call in one memset: [max_norm(4 ints) | flags(4 ints) | above R | cand_cnt R]
  w.max_norm = (int*)take(sizeof(int) * (8 + 2 * (size_t)R));
  w.flags = w.max_norm + 4; w.above = w.max_norm + 8; w.cand_cnt = w.above + R;
  w.cand = (int*)take(sizeof(int) * (size_t)R * cap);
  w.tk_cnt = (int*)take(sizeof(int) * (size_t)R);
  w.tk_cand = (int*)take(sizeof(int) * (size_t)R * cap);
  w.lo_topk = (float*)take(sizeof(float) * (size_t)R);
  w.maxbuf = (float*)take(sizeof(float) * (size_t)R * 2 * 128);
  w.bytes = o;
  return w;
}

}  // namespace ev
}  // namespace ader

using namespace ader;
using namespace ader::ev;

extern "C" size_t ader_eval_rank_tc_ws_bytes(const AderModel* m, int32_t R, int32_t V) {
  if (check_model(m) || R <= 0 || V <= 0 || m->d > HALF) return 0;
  return carve(R, V, ADER_EVAL_CAND_CAP, nullptr).bytes;
}

static int eval_tc_run(const AderModel* m, const float* theta, const float* rep, const int32_t* gt, int R, int V, int k, void* ws,
                       int32_t* rank, int32_t* topk_item, float* topk_score, int32_t* overflow, cudaStream_t st) {
  const int d = m->d, cap = ADER_EVAL_CAND_CAP;
  EvalWs w = carve(R, V, cap, (char*)ws);
  const int nm = cdiv(R, TM), nv = cdiv(V, TN), nc = eval_chunks(nm, nv);
  ADER_CHECK_ARG(k == 0 || (2 * nc >= k && 2 * nc <= 256), "eval_rank_tc: top-%d needs %d..128 vocabulary chunks per row tile (have %d): use ader_eval_rank_topk", k, (k + 1) / 2, nc);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_eval_tc<EV_RANK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_EVAL);
    cudaFuncSetAttribute(k_eval_tc<EV_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_EVAL);
    cudaFuncSetAttribute(k_eval_tc<EV_BOTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_EVAL);
    attr = true;
  }
  CUtensorMap tx, ty;
  if (int e = make_map2d(&tx, w.x16, D2, (uint64_t)nm * TM, D2 * 2, TM)) return e;
  if (int e = make_map2d(&ty, w.y16, D2, (uint64_t)nv * TN, D2 * 2, TN)) return e;
  const float* table1 = theta + d;                       // item 1 = table row 1 (ADER.py:90-91)
  cudaMemsetAsync(w.max_norm, 0, sizeof(int) * (8 + 2 * (size_t)R), st);
  k_pack_hilo<<<cdiv((long long)nv * TN * 32, 256), 256, 0, st>>>(table1, d, V, d, nv * TN, w.y16, nullptr, w.max_norm);
  k_pack_hilo<<<cdiv((long long)nm * TM * 32, 256), 256, 0, st>>>(rep, d, R, d, nm * TM, w.x16, w.rnorm, nullptr);
  k_gt_band<<<cdiv(R, 256), 256, 0, st>>>(table1, rep, gt, R, d, w.rnorm, w.max_norm, w.sg, w.lo, w.hi);
  EvalArgs a;
  a.R = R; a.V = V; a.n_mtiles = nm; a.n_vtiles = nv; a.n_chunks = nc;
  a.lo_thr = w.lo; a.hi_thr = w.hi; a.above = w.above; a.cand_cnt = w.cand_cnt; a.cand = w.cand; a.cap = cap; a.flags = w.flags;
  a.maxbuf = w.maxbuf; a.lo_topk = w.lo_topk; a.tk_cnt = w.tk_cnt; a.tk_cand = w.tk_cand;
  if (k > 0) {
    cudaMemsetAsync(w.tk_cnt, 0, sizeof(int) * (size_t)R, st);
    k_eval_tc<EV_MAX><<<nm * nc, NTHREADS, SMEM_EVAL, st>>>(a, tx, ty);
    k_topk_threshold<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(w.maxbuf, 2 * nc, R, k, w.rnorm, w.max_norm, w.lo_topk);
    k_eval_tc<EV_BOTH><<<nm * nc, NTHREADS, SMEM_EVAL, st>>>(a, tx, ty);
    k_topk_refine<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(table1, rep, R, d, w.tk_cnt, w.tk_cand, cap, k, topk_item, topk_score);
  } else {
    k_eval_tc<EV_RANK><<<nm * nc, NTHREADS, SMEM_EVAL, st>>>(a, tx, ty);
  }
  k_refine<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(table1, rep, gt, R, d, w.sg, w.above, w.cand_cnt, w.cand, cap, rank);
  cudaMemcpyAsync(overflow, w.flags, sizeof(int), cudaMemcpyDeviceToDevice, st);
  ADER_CHECK_LAUNCH("eval_rank_tc");
  return 0;
}

extern "C" int32_t ader_eval_rank_tc(const AderModel* m, const float* theta, const float* rep, const int32_t* gt, int32_t R,
                                     int32_t V, void* ws, int32_t* rank, int32_t* overflow, void* stream) {
  if (int e = check_model(m)) return e;
  ADER_CHECK_ARG(theta && rep && gt && ws && rank && overflow, "eval_rank_tc: NULL pointer");
  ADER_CHECK_ARG(R > 0 && V >= 1 && V < m->v_tab, "eval_rank_tc: bad sizes");
  ADER_CHECK_ARG(m->d <= HALF, "eval_rank_tc: hidden_units must be <= %d", HALF);
  return eval_tc_run(m, theta, rep, gt, R, V, 0, ws, rank, nullptr, nullptr, overflow, (cudaStream_t)stream);
}

extern "C" int32_t ader_eval_topk_chunks(const AderModel* m, int32_t R, int32_t V) {
  if (check_model(m) || R <= 0 || V <= 0) return 0;
  return eval_chunks(cdiv(R, TM), cdiv(V, TN));
}

extern "C" int32_t ader_eval_rank_topk_tc(const AderModel* m, const float* theta, const float* rep, const int32_t* gt, int32_t R,
                                          int32_t V, int32_t k, void* ws, int32_t* rank, int32_t* topk_item, float* topk_score,
                                          int32_t* overflow, void* stream) {
  if (int e = check_model(m)) return e;
  ADER_CHECK_ARG(theta && rep && gt && ws && rank && overflow && topk_item && topk_score, "eval_rank_topk_tc: NULL pointer");
  ADER_CHECK_ARG(R > 0 && V >= 1 && V < m->v_tab && k >= 1 && k <= 32, "eval_rank_topk_tc: bad sizes");
  ADER_CHECK_ARG(m->d <= HALF, "eval_rank_topk_tc: hidden_units must be <= %d", HALF);
  return eval_tc_run(m, theta, rep, gt, R, V, k, ws, rank, topk_item, topk_score, overflow, (cudaStream_t)stream);
}
