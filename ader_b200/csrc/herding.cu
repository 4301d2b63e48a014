// Segmented herding exemplar selection (util.py:401-434).  One CTA per label item.
// All arithmetic is fp32 in the order NumPy uses where that order is defined:
//   norm_j = sqrt(pairwise_sum(x*x))   (np.linalg.norm axis reduce over the contiguous axis)
//   D_j    = x / norm_j ;  mu = (sequential sum over candidates) / n ;  w <- (w + mu) - D[:, i]
// The 150-term dot w.D_j is a BLAS call in the reference (summation order unspecified); here
// it is a lane-strided sum + butterfly.  Picks can therefore differ only at near-ties.
#include "common.cuh"

namespace ader {

// NumPy pairwise_sum for a contiguous float32 vector of squares (numpy/core/src/umath/loops_utils.h):
// n < 8: sequential; n <= 128: 8 accumulators; else split at n/2 rounded down to a multiple of 8.
__device__ float np_pairwise_sq(const float* __restrict__ x, int n) {
  if (n < 8) {
    float r = 0.f;
    for (int i = 0; i < n; ++i) r = __fadd_rn(r, __fmul_rn(x[i], x[i]));
    return r;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = __fmul_rn(x[k], x[k]);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = __fadd_rn(r[k], __fmul_rn(x[i + k], x[i + k]));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, __fmul_rn(x[i], x[i]));
    return res;
  }
  int n2 = n / 2; n2 -= n2 % 8;
  return __fadd_rn(np_pairwise_sq(x, n2), np_pairwise_sq(x + n2, n - n2));
}

__global__ void __launch_bounds__(256) k_herding(const float* __restrict__ rep, int d, const int* __restrict__ cand,
                                                 const int* __restrict__ seg_off, const int* __restrict__ quota,
                                                 const int* __restrict__ max_steps, float* __restrict__ Dn,
                                                 int* __restrict__ selected, int* __restrict__ picks,
                                                 int* __restrict__ n_picked) {
  __shared__ float w[256], mu[256];
  __shared__ float bestv[8]; __shared__ int bestj[8];
  __shared__ int pick_s, cnt_s;
  const int s = blockIdx.x;
  const int off = seg_off[s], n = seg_off[s + 1] - off;
  const int m = min(quota[s], n);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (n <= 0 || m <= 0) { if (tid == 0) n_picked[s] = 0; return; }

  // D = rep^T / ||rep^T||_2, one thread per candidate for the norm (NumPy's summation order)
  for (int j = tid; j < n; j += blockDim.x) {
    const float* x = rep + (long long)cand[off + j] * d;
    float nrm = sqrtf(np_pairwise_sq(x, d));
    float* o = Dn + (long long)(off + j) * d;
    for (int c = 0; c < d; ++c) o[c] = __fdiv_rn(x[c], nrm);
    selected[off + j] = 0;
  }
  __syncthreads();
  if (tid < d) {
    float acc = 0.f;
    for (int j = 0; j < n; ++j) acc = __fadd_rn(acc, Dn[(long long)(off + j) * d + tid]);
    float mval = __fdiv_rn(acc, (float)n);
    mu[tid] = mval; w[tid] = mval;
  }
  if (tid == 0) cnt_s = 0;
  __syncthreads();

  const int steps = max_steps[s];
  for (int step = 0; step < steps; ++step) {
    float bv = -INFINITY; int bj = 0x7fffffff;
    for (int j = warp; j < n; j += 8) {
      const float* D = Dn + (long long)(off + j) * d;
      float acc = 0.f;
      for (int c = lane; c < d; c += 32) acc = fmaf(w[c], D[c], acc);
#pragma unroll
      for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (acc > bv) { bv = acc; bj = j; }          // j ascending inside a warp: first max kept
    }
    if (lane == 0) { bestv[warp] = bv; bestj[warp] = bj; }
    __syncthreads();
    if (tid == 0) {
      float fv = bestv[0]; int fj = bestj[0];
      for (int q = 1; q < 8; ++q)
        if (bestv[q] > fv || (bestv[q] == fv && bestj[q] < fj)) { fv = bestv[q]; fj = bestj[q]; }
      // every dot product NaN (a zero-norm rep row: 0/0 in the normalisation): no comparison succeeded; np.argmax
      // returns index 0 for an all-NaN vector (util.py:426), and the index must stay inside the segment
      if (fj < 0 || fj >= n) fj = 0;
      pick_s = fj;
      if (!selected[off + fj]) { selected[off + fj] = 1; picks[off + cnt_s] = fj; cnt_s = cnt_s + 1; }
    }
    __syncthreads();
    const int pk = pick_s;
    if (tid < d) w[tid] = __fsub_rn(__fadd_rn(w[tid], mu[tid]), Dn[(long long)(off + pk) * d + tid]);
    const bool done = (cnt_s == m);
    __syncthreads();
    if (done) break;
  }
  if (tid == 0) n_picked[s] = cnt_s;
}

}  // namespace ader

using namespace ader;

extern "C" size_t ader_herding_ws_bytes(const AderModel* m, int32_t N) {
  if (check_model(m) || N <= 0) return 0;
  return align_up(sizeof(float) * (size_t)N * m->d) + align_up(sizeof(int) * (size_t)N);
}

extern "C" int32_t ader_herding_segmented(const AderModel* m, const float* rep, int32_t N, const int32_t* cand,
                                          const int32_t* seg_off, int32_t n_seg, const int32_t* quota,
                                          const int32_t* max_steps, void* ws, int32_t* picks, int32_t* n_picked,
                                          void* stream) {
  if (int e = check_model(m)) return e;
  ADER_CHECK_ARG(rep && cand && seg_off && quota && max_steps && ws && picks && n_picked, "herding: NULL pointer");
  ADER_CHECK_ARG(N > 0 && n_seg > 0, "herding: empty input");
  float* Dn = (float*)ws;
  int* selected = (int*)((char*)ws + align_up(sizeof(float) * (size_t)N * m->d));
  k_herding<<<n_seg, 256, 0, (cudaStream_t)stream>>>(rep, m->d, cand, seg_off, quota, max_steps, Dn, selected, picks, n_picked);
  ADER_CHECK_LAUNCH("herding");
  return 0;
}
