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
};

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
    const float lo = live ? a.lo_thr[gi] : INFINITY, hi = live ? a.hi_thr[gi] : INFINITY;
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
      if (live) {
        if (j0 + 32 <= a.V) {
          int unsure = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = __uint_as_float(r[i]);
            above += (v > hi) ? 1 : 0;
            unsure |= (v >= lo && v <= hi) ? (1 << i) : 0;
          }
          while (unsure) {                                  // rare: a handful of columns per row in the whole pass
            const int i = __ffs(unsure) - 1; unsure &= unsure - 1;
            const int pos = atomicAdd(a.cand_cnt + gi, 1);
            if (pos < a.cap) a.cand[(long long)gi * a.cap + pos] = j0 + i; else atomicExch(a.flags, 1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            if (j0 + i >= a.V) continue;                    // zero padding rows of the table matrix are not items
            const float v = __uint_as_float(r[i]);
            if (v > hi) ++above;
            else if (v >= lo) {
              const int pos = atomicAdd(a.cand_cnt + gi, 1);
              if (pos < a.cap) a.cand[(long long)gi * a.cap + pos] = j0 + i; else atomicExch(a.flags, 1);
            }
          }
        }
      }
    }
    if (live && above) atomicAdd(a.above + gi, above);
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

static int eval_chunks(int n_mtiles, int n_vtiles) {
  int best = 1; double best_eff = 0.0;
  const int cmax = n_vtiles / 8 > 1 ? (n_vtiles / 8 < 148 ? n_vtiles / 8 : 148) : 1;      // >= 8 streamed tiles per CTA
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
  int *max_norm, *above, *cand_cnt, *cand, *flags;
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

extern "C" int32_t ader_eval_rank_tc(const AderModel* m, const float* theta, const float* rep, const int32_t* gt, int32_t R,
                                     int32_t V, void* ws, int32_t* rank, int32_t* overflow, void* stream) {
  if (int e = check_model(m)) return e;
  ADER_CHECK_ARG(theta && rep && gt && ws && rank && overflow, "eval_rank_tc: NULL pointer");
  ADER_CHECK_ARG(R > 0 && V >= 1 && V < m->v_tab, "eval_rank_tc: bad sizes");
  ADER_CHECK_ARG(m->d <= HALF, "eval_rank_tc: hidden_units must be <= %d", HALF);
  cudaStream_t st = (cudaStream_t)stream;
  const int d = m->d, cap = ADER_EVAL_CAND_CAP;
  EvalWs w = carve(R, V, cap, (char*)ws);
  const int nm = cdiv(R, TM), nv = cdiv(V, TN), nc = eval_chunks(nm, nv);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(k_eval_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_EVAL); attr = true; }
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
  k_eval_tc<<<nm * nc, NTHREADS, SMEM_EVAL, st>>>(a, tx, ty);
  k_refine<<<cdiv((long long)R * 32, 256), 256, 0, st>>>(table1, rep, gt, R, d, w.sg, w.above, w.cand_cnt, w.cand, cap, rank);
  cudaMemcpyAsync(overflow, w.flags, sizeof(int), cudaMemcpyDeviceToDevice, st);
  ADER_CHECK_LAUNCH("eval_rank_tc");
  return 0;
}
