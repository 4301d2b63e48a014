// Generic strided fp32 GEMM with fused epilogues -- the exact-arithmetic workhorse of the
// parity path (encoder dense layers, weight gradients with deterministic split-K, materialised
// logits for the fp32 reference mode).  64x64x16 tiles, 256 threads, 4x4 register micro-tiles.
// Summation order over k is fixed (sequential inside a split, splits reduced in index order by
// reduce_partials) so results are run-to-run deterministic.
#include "common.cuh"

namespace ader {

void gemm_defaults(GemmArgs& g) {
  g.A = nullptr; g.B = nullptr; g.C = nullptr;
  g.a_rs = g.a_cs = g.b_rs = g.b_cs = g.c_rs = g.c_cs = 0;
  g.M = g.N = g.K = 0; g.dM = nullptr; g.dK = nullptr;
  g.bias = nullptr; g.resid = nullptr; g.resid_ld = 0; g.relu_mask = nullptr; g.mask_ld = 0;
  g.relu = 0; g.accumulate = 0; g.alpha = 1.0f; g.splits = 1; g.split_stride = 0; g.colsum = nullptr;
  g.drop_p = 0.f; g.drop_seed = 0; g.drop_site = 0;
}

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g) {
  const int M = g.dM ? *g.dM : g.M;
  const int K = g.dK ? *g.dK : g.K;
  const int N = g.N;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= M || n0 >= N) return;
  const int z = blockIdx.z;
  const int kchunk = ((K + g.splits - 1) / g.splits + BK - 1) / BK * BK;
  const int k_lo = z * kchunk;
  const int k_hi = min(K, k_lo + kchunk);

  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_colsum = (g.colsum != nullptr) && (blockIdx.y == 0) && (ty == 0);

  const bool a_kfast = (g.a_cs == 1);   // k contiguous in A
  const bool b_nfast = (g.b_cs == 1);   // n contiguous in B

  // register-staged double buffering: the global loads of tile k+1 are in flight while tile k is
  // multiplied out of shared memory.
  float ra[4], rb[4];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int mm, kk;
      if (a_kfast) { kk = idx & (BK - 1); mm = idx >> 4; } else { mm = idx & (BM - 1); kk = idx >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < M && gk < k_hi) ? g.A[(long long)gm * g.a_rs + (long long)gk * g.a_cs] : 0.f;
      int nn;
      if (b_nfast) { nn = idx & (BN - 1); kk = idx >> 6; } else { kk = idx & (BK - 1); nn = idx >> 4; }
      int gn = n0 + nn; gk = k0 + kk;
      rb[i] = (gn < N && gk < k_hi) ? g.B[(long long)gk * g.b_rs + (long long)gn * g.b_cs] : 0.f;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int mm, kk, nn;
      if (a_kfast) { kk = idx & (BK - 1); mm = idx >> 4; } else { mm = idx & (BM - 1); kk = idx >> 6; }
      As[kk][mm] = ra[i];
      if (b_nfast) { nn = idx & (BN - 1); kk = idx >> 6; } else { kk = idx & (BK - 1); nn = idx >> 4; }
      Bs[kk][nn] = rb[i];
    }
  };
  if (k_lo < k_hi) { load_tile(k_lo); store_tile(); }
  __syncthreads();
  for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
    const bool has_next = k0 + BK < k_hi;
    if (has_next) load_tile(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a.x, a.y, a.z, a.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (do_colsum) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bsum[j] += bv[j];
      }
    }
    __syncthreads();
    if (has_next) store_tile();
    __syncthreads();
  }

  float* C = g.C + (long long)z * g.split_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = g.alpha * acc[i][j];
      if (g.bias) v += g.bias[gn];
      if (g.relu) v = fmaxf(v, 0.f);
      if (g.relu_mask) v = (g.relu_mask[(long long)gm * g.mask_ld + gn] > 0.f) ? v : 0.f;
      if (g.drop_p > 0.f) v *= drop_scale(g.drop_seed, g.drop_site, (uint64_t)gm * N + gn, g.drop_p);
      if (g.resid) v += g.resid[(long long)gm * g.resid_ld + gn];
      long long o = (long long)gm * g.c_rs + (long long)gn * g.c_cs;
      if (g.accumulate) v += C[o];
      C[o] = v;
    }
  }
  if (do_colsum) {
    float* cs = g.colsum + (long long)z * g.split_stride;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn < N) cs[gn] = bsum[j];
    }
  }
}

int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0) return 0;
  dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), g.splits);
  if (grid.y > 65535u) return fail(-2, "sgemm: M=%d too large for grid.y", g.M);
  sgemm_kernel<<<grid, 256, 0, st>>>(g);
  ADER_CHECK_LAUNCH("sgemm_kernel");
  return 0;
}

}  // namespace ader
