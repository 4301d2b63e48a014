// The encoder as ONE forward kernel and ONE backward kernel ("chain collapse").
//
// Attention never crosses a session (modules.py:177-223: keys are the earlier tokens of the same row) and every other
// sub-layer is token-wise, so nothing in the SASRec blocks of ADER.py:25-85 needs a grid-wide dependency: a CTA that owns
// a contiguous token range can run LN1 + Q/K/V, attention + LN2, the FFN, the next block and the final LayerNorm back to
// back, separated by __syncthreads only - plus ONE neighbour handshake per block where a session straddles two ranges.  The
// per-sub-layer kernels of encoder_fused.cuh (7 launches forward, 7 backward: a dependent chain of ~180 us at the
// reference's batch sizes, every link paying launch latency and a cold weight load) become the phases of two persistent
// kernels built from the same tile bodies (qkv_fwd_tile, attn_fwd_group, ffn_fwd_tile, ffn_bwd_tile, attn_bwd_group,
// qkv_bwd_tile), so the arithmetic - and therefore every saved activation and gradient - is bit-identical to the
// per-sub-layer path (tests/test_gpu_parity.py compares the two).
//
// Partition.  CTA c of G owns tokens [ceil(cT/G), ceil((c+1)T/G)): equal shares (one 32-token tile per phase at the
// reference's batch sizes), whatever the session lengths.  A session cut by a range boundary needs its neighbours' rows in
// the attention phases only: forward, the K / V rows of EARLIER tokens (lower CTAs); backward, gY / D of LATER queries
// (higher CTAs).  Each CTA publishes "my rows of block b are written" (st.release.gpu after __threadfence) and polls the
// flags of the few CTAs its sessions reach into (ld.acquire.gpu, bounded spin).  All CTAs are co-resident (grid <= SMs, one
// CTA per SM) and a CTA only ever waits for a phase every CTA reaches without waiting, so the handshake cannot deadlock.
//
// Weights.  Three 53 760-byte shadow slots in shared memory.  The shadows of the NEXT phase are fetched by cp.async.bulk
// while the current phase computes (attention needs no weights at all, so both FFN matrices land behind it):
//   forward   Q/K/V: Wq->S2 Wk->S0 Wv->S1 | attention (K/V rows staged in S2 + the operand tiles) | FFN: W1->S0 W2->S1
//   backward  FFN: W2^T->S0 W1^T->S1 | attention (staging in S2 + tiles) | Q/K/V: Wk^T->S0 Wv^T->S1 first, Wq^T->S2 last
// one mbarrier per slot group, phase parity = block counter.
#pragma once

namespace ader {
namespace fz {

constexpr int CH_MAXB = 2;                       // blocks per chained kernel (the reference's num_blocks)

__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// token range [ta, tb) of CTA c of G, and the owner of token t under the same rule
__device__ __forceinline__ int chain_cut(int T, int G, int c) { return (int)(((long long)T * c + G - 1) / G); }
__device__ __forceinline__ int chain_owner(int T, int G, int t) {
  int c = (int)(((long long)t * G) / max(T, 1));
  c = min(max(c, 0), G - 1);
  while (c + 1 < G && chain_cut(T, G, c + 1) <= t) ++c;
  while (c > 0 && chain_cut(T, G, c) > t) --c;
  return c;
}
__device__ __forceinline__ void flag_publish(int* flag) {      // call behind a __syncthreads, one thread
  __threadfence();
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(1) : "memory");
}
__device__ __forceinline__ void flag_wait(const int* flag) {
  int v = 0;
  for (uint32_t spin = 0; !v; ++spin) {
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (!v && spin > (1u << 24)) __trap();     // a lost neighbour must not hang the GPU
  }
}

// ---- attention for the chained kernels: a TEAM of 8 lanes per query ------------------------------------------------
// The per-sub-layer kernel (attn_fwd_group) gives a whole warp to one query and reduces 8 key dots by an 11-shuffle
// reduce-scatter: fine with 24 warps per SM, a long dependent chain with the 8 warps of a chained CTA.  Sessions are short
// (median 3-5 tokens), so here every one of the <= 32 queries of the tile gets 8 lanes (19 features each, 8 x 19 = 152 >= d):
// one dot product = 19 FMAs + 3 shuffles, all 32 queries of the tile run at once, scores live in shared memory, the softmax
// is a plain two-pass one (no online rescaling).  K rows of [first session start, tile end) are staged once, then V rows
// into the same buffer.  Plain fp32 throughout; single head (the chained path requires num_heads == 1).
constexpr int TEAM = 8, TE = 19, ALD = 152;
__host__ __device__ constexpr int att_tile_rows(int L) { return TM + L - 1; }
__host__ __device__ constexpr int att_sld(int L) { return L + 2; }
__host__ __device__ constexpr size_t att_tile_smem(int L) { return sizeof(float) * ((size_t)att_tile_rows(L) * ALD + (size_t)TM * att_sld(L)); }

// rows [0, n) of a [*, d] fp32 matrix -> shared rows of stride ALD, padding columns zeroed (d even, 8 loads in flight)
__device__ __forceinline__ void stage_rows_ald(float* __restrict__ dst, const float* src, int n, int d) {
  const int h = d >> 1, total = n * (ALD / 2);
  for (int base = threadIdx.x; base < total; base += 8 * NTHR) {
    float2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * NTHR, r = idx / (ALD / 2), c2 = idx % (ALD / 2);
      v[u] = (idx < total && c2 < h) ? *reinterpret_cast<const float2*>(src + (long long)r * d + 2 * c2) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int idx = base + u * NTHR;
      if (idx < total) *reinterpret_cast<float2*>(dst + (idx / (ALD / 2)) * ALD + 2 * (idx % (ALD / 2))) = v[u];
    }
  }
}
__device__ __forceinline__ float team_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ float team_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One tile of <= TM query tokens [t0, min(t0 + TM, T)); every thread of the CTA calls it (two __syncthreads inside, the
// caller adds one before the staging buffer is reused).
__device__ __forceinline__ void attn_fwd_tile(const AttnFwdArgs& a, int t0, int T, float* __restrict__ att_sm, uint64_t seed_eff) {
  const int tid = threadIdx.x, team = tid >> 3, tl = tid & 7;
  const int d = a.d, L = a.L, sld = att_sld(L);
  const int hi = min(t0 + TM, T);
  const int lo = a.row_off[a.tok_row[t0]];
  float* KV = att_sm;
  float* Ssm = att_sm + att_tile_rows(L) * ALD + team * sld;
  const bool live = t0 + team < hi;
  const int tk = live ? t0 + team : hi - 1;
  const int off = a.row_off[a.tok_row[tk]];
  const int i = live ? tk - off : -1;             // keys 0..i of the query's own session
  const DropSite dsp = drop_site(seed_eff, a.site, a.drop_p);
  float q[TE], q1[TE];
#pragma unroll
  for (int u = 0; u < TE; ++u) {
    const int c = tl + TEAM * u;
    q[u] = (c < d) ? a.Q[(long long)tk * d + c] : 0.f;
    q1[u] = (c < d) ? a.Q1[(long long)tk * d + c] : 0.f;
  }
  FZ_TL(6, 0);
  stage_rows_ald(KV, a.K + (long long)lo * d, hi - lo, d);
  __syncthreads();
  FZ_TL(6, 1);
  const float inv_denom = 1.0f / sqrtf((float)d);
  const float* rows = KV + (off - lo) * ALD + tl;
  const int imax = warp_max_i(i);
  // ---- scores s_j = q . k_j / sqrt(d), running maximum
  float mx = -INFINITY;
  for (int j = 0; j <= imax; ++j) {
    const float* kr = rows + min(j, max(i, 0)) * ALD;
    float p = 0.f;
#pragma unroll
    for (int u = 0; u < TE; ++u) p = fmaf(q[u], kr[TEAM * u], p);
    p = team_sum(p) * inv_denom;
    if (j <= i) { mx = fmaxf(mx, p); if (tl == 0) Ssm[j] = p; }
  }
  __syncwarp();
  FZ_TL(6, 2);
  // ---- softmax over keys 0..i (lane tl takes keys tl, tl + 8, ...); probabilities to HBM, dropped ones stay in smem
  float lsum = 0.f;
  for (int j = tl; j <= i; j += TEAM) { const float e = expf(Ssm[j] - mx); Ssm[j] = e; lsum += e; }
  lsum = team_sum(lsum);
  const float inv_l = 1.0f / lsum;
  const long long po = (long long)tk * L;
  if (live) {
    for (int j = tl; j < L; j += TEAM) {
      float p = 0.f;
      if (j <= i) {
        p = Ssm[j] * inv_l;
        Ssm[j] = (a.drop_p > 0.f) ? p * drop_mul(dsp, (uint64_t)(po + j)) : p;
      }
      a.probs[po + j] = p;
    }
  }
  FZ_TL(6, 3);
  __syncthreads();                                  // every team is done with the K rows
  FZ_TL(6, 4);
  stage_rows_ald(KV, a.V + (long long)lo * d, hi - lo, d);
  __syncthreads();
  FZ_TL(6, 5);
  // ---- o = sum_j Pd_j v_j ; y = o + q1 (residual on the NORMALISED queries, modules.py:223) ; z = LN2(y)
  float y[TE];
#pragma unroll
  for (int u = 0; u < TE; ++u) y[u] = 0.f;
  for (int j = 0; j <= i; ++j) {                    // dead teams (i = -1) skip; no shuffles inside
    const float p = Ssm[j];
    const float* vr = rows + j * ALD;
#pragma unroll
    for (int u = 0; u < TE; ++u) y[u] = fmaf(p, vr[TEAM * u], y[u]);
  }
  __syncwarp();
  FZ_TL(6, 6);
  // LayerNorm parameters first: a load behind a store through an unrelated pointer is not hoisted (possible aliasing),
  // 19 serialised L2 round trips otherwise
  float lg[TE], lb[TE];
#pragma unroll
  for (int u = 0; u < TE; ++u) { const int c = tl + TEAM * u; lg[u] = (c < d) ? a.ln_g[c] : 0.f; lb[u] = (c < d) ? a.ln_b[c] : 0.f; }
  float s1 = 0.f;
#pragma unroll
  for (int u = 0; u < TE; ++u) {
    const int c = tl + TEAM * u;
    y[u] = (live && c < d) ? y[u] + q1[u] : 0.f;
    if (live && c < d) a.Y[(long long)tk * d + c] = y[u];
    s1 += y[u];
  }
  const float inv_d = 1.0f / (float)d;
  const float mean = team_sum(s1) * inv_d;
  float qq = 0.f;
#pragma unroll
  for (int u = 0; u < TE; ++u) { const int c = tl + TEAM * u; if (c < d) { const float w = y[u] - mean; qq += w * w; } }
  const float rstd = rsqrtf(team_sum(qq) * inv_d + 1e-8f);
  if (live) {
#pragma unroll
    for (int u = 0; u < TE; ++u) {
      const int c = tl + TEAM * u;
      if (c < d) a.Z[(long long)tk * d + c] = lg[u] * ((y[u] - mean) * rstd) + lb[u];
    }
    if (tl == 0) { a.mean2[tk] = mean; a.rstd2[tk] = rstd; }
  }
  FZ_TL(6, 7);
}

// the same tile as a kernel of its own (per-sub-layer path): CTA per 32 query tokens
__global__ void __launch_bounds__(NTHR) k_attn_ln_fwd_team(const __grid_constant__ AttnFwdArgs a) {
  extern __shared__ __align__(16) float att_sm[];
  const uint64_t seed_eff = eff_seed(a.seed, a.d_step);
  const int T = *a.dT;
  const int t0 = blockIdx.x * TM;
  if (t0 >= T) return;
  pdl_wait(); pdl_go();
  attn_fwd_tile(a, t0, T, att_sm, seed_eff);
}

struct ChainFwdArgs {
  QkvFwdArgs q[CH_MAXB]; AttnFwdArgs at[CH_MAXB]; FfnFwdArgs f[CH_MAXB];
  int nb, M;
  const float* xfinal; float *rep, *meanf, *rstdf; const float *lnf_b, *lnf_g;
  int* flags;                 // [CH_MAXB][stride] neighbour flags, zero at launch
  int flag_stride;
};
constexpr size_t CHAIN_FWD_SMEM = 3 * WMAT_BYTES + 2 * ATILE_BYTES + TM * 4 + 32;
constexpr size_t CHAIN_FWD_STAGE = WMAT_BYTES + 2 * ATILE_BYTES;          // attention staging: slot 2 + both operand tiles

__global__ void __launch_bounds__(NTHR, 1) k_chain_fwd(const __grid_constant__ ChainFwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  op_t* S0 = reinterpret_cast<op_t*>(smem);
  op_t* S1 = reinterpret_cast<op_t*>(smem + WMAT_BYTES);
  op_t* S2 = reinterpret_cast<op_t*>(smem + 2 * WMAT_BYTES);
  op_t* A1 = reinterpret_cast<op_t*>(smem + 3 * WMAT_BYTES);
  op_t* A2 = A1 + TM * LDS;
  float* rs = reinterpret_cast<float*>(smem + 3 * WMAT_BYTES + 2 * ATILE_BYTES);
  const uint32_t barQ = smem_u32(smem + 3 * WMAT_BYTES + 2 * ATILE_BYTES + TM * 4), barKV = barQ + 8, barF = barQ + 16;
  float* att_sm = reinterpret_cast<float*>(S2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ng = warp >> 1;
  if (tid == 0) {
    mbar_init(barQ, 1); mbar_init(barKV, 1); mbar_init(barF, 1);
    mbar_expect_tx(barQ, WMAT_BYTES);
    load_wmat(smem_u32(S2), a.q[0].Wq, barQ);
    mbar_expect_tx(barKV, 2 * WMAT_BYTES);
    load_wmat(smem_u32(S0), a.q[0].Wk, barKV);
    load_wmat(smem_u32(S1), a.q[0].Wv, barKV);
  }
  FZ_TL(2, 0);
  pdl_wait(); pdl_go();
  __syncthreads();
  const int T = *a.q[0].dT, d = a.q[0].d;
  const int G = gridDim.x, cta = blockIdx.x;
  const int ta = chain_cut(T, G, cta), tb = chain_cut(T, G, cta + 1);
  // lower CTAs whose K / V rows this CTA's first session reaches into
  const int c_lo = ta < tb ? chain_owner(T, G, a.q[0].row_off[a.q[0].tok_row[ta]]) : cta;
  FZ_TL(2, 1);
  FZ_TLV(2, 15, tb - ta);
  const bool work = ta < tb;
  const uint64_t seed_eff = eff_seed(a.q[0].seed, a.q[0].d_step);

  for (int b = 0; b < a.nb; ++b) {
    const uint32_t par = b & 1;
    {   // ---- LN1 + Q / K / V
      const QkvFwdArgs& q = a.q[b];
      float lg[NE], lb[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) { const int c = lane + 32 * e; lg[e] = (c < d) ? q.ln_g[c] : 0.f; lb[e] = (c < d) ? q.ln_b[c] : 0.f; }
      float2 bq[NT], bk[NT], bv[NT];
      load_bias_frag(bq, q.bq, d, ng, lane); load_bias_frag(bk, q.bk, d, ng, lane); load_bias_frag(bv, q.bv, d, ng, lane);
      for (int t0 = ta; t0 < tb; t0 += TM)
        qkv_fwd_tile(q, t0, tb, seed_eff, A1, A2, rs, S2, S0, S1, lg, lb, bq, bk, bv,
                     [&] { mbar_wait(barQ, par); }, [&] { mbar_wait(barKV, par); });
    }
    FZ_TL(2, 2 + 3 * b);
    if (tid == 0) {
      flag_publish(a.flags + b * a.flag_stride + cta);            // Q / K / V rows of this range are written
      if (work) {                                // both FFN shadows land while the attention phase runs
        fence_proxy_async();
        mbar_expect_tx(barF, 2 * WMAT_BYTES);
        load_wmat(smem_u32(S0), a.f[b].W1, barF);
        load_wmat(smem_u32(S1), a.f[b].W2, barF);
      }
    }
    for (int cc = c_lo + tid; cc < cta; cc += NTHR) flag_wait(a.flags + b * a.flag_stride + cc);
    __syncthreads();
    FZ_TL(7, b);
    // ---- attention + residual + LN2 (K, then V rows of the tile's sessions staged in slot 2 + the operand tiles)
    for (int t0 = ta; t0 < tb; t0 += TM) { attn_fwd_tile(a.at[b], t0, tb, att_sm, seed_eff); __syncthreads(); }
    for (int idx = tid; idx < TM * LDS; idx += NTHR) A2[idx] = to_op(0.f);      // padding columns of the hidden tile
    fence_proxy_async();
    __syncthreads();
    FZ_TL(2, 3 + 3 * b);
    if (work && tid == 0 && b + 1 < a.nb) {
      mbar_expect_tx(barQ, WMAT_BYTES);
      load_wmat(smem_u32(S2), a.q[b + 1].Wq, barQ);
    }
    {   // ---- FFN
      const FfnFwdArgs& f = a.f[b];
      float2 b1[NT], b2[NT];
      load_bias_frag(b1, f.b1, d, ng, lane); load_bias_frag(b2, f.b2, d, ng, lane);
      for (int t0 = ta; t0 < tb; t0 += TM)
        ffn_fwd_tile(f, t0, tb, seed_eff, A1, A2, S0, S1, b1, b2, [&] { mbar_wait(barF, par); });
    }
    FZ_TL(2, 4 + 3 * b);
    if (work && tid == 0 && b + 1 < a.nb) {
      fence_proxy_async();
      mbar_expect_tx(barKV, 2 * WMAT_BYTES);
      load_wmat(smem_u32(S0), a.q[b + 1].Wk, barKV);
      load_wmat(smem_u32(S1), a.q[b + 1].Wv, barKV);
    }
  }
  if (!work && tid == 0) { mbar_wait(barQ, 0); mbar_wait(barKV, 0); }          // never exit under an in-flight copy

  // ---- final LayerNorm on the last token of every session that ends in this range (ADER.py:82-85)
  for (int t = ta + warp; t < tb; t += NTHR / 32) {
    const int r = a.q[0].tok_row[t];
    if (t != a.q[0].row_off[r + 1] - 1) continue;
    float* out = a.rep + (long long)r * d;
    const float* x = a.xfinal + (long long)t * d;
    float v[NE], s = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e) { const int c = lane + 32 * e; v[e] = (c < d) ? x[c] : 0.f; s += v[e]; }
    const float mean = warp_sum(s) / (float)d;
    float qq = 0.f;
#pragma unroll
    for (int e = 0; e < NE; ++e) { const int c = lane + 32 * e; if (c < d) { const float u = v[e] - mean; qq += u * u; } }
    const float rstd = 1.0f / sqrtf(warp_sum(qq) / (float)d + 1e-8f);
#pragma unroll
    for (int e = 0; e < NE; ++e) { const int c = lane + 32 * e; if (c < d) out[c] = a.lnf_g[c] * ((v[e] - mean) * rstd) + a.lnf_b[c]; }
    if (lane == 0) { a.meanf[r] = mean; a.rstdf[r] = rstd; }
  }
  if (cta == 0) {                                // empty rows: zeros, carry no gradient
    for (int r = warp; r < a.M; r += NTHR / 32) {
      if (a.q[0].row_len[r] != 0) continue;
      for (int c = lane; c < d; c += 32) a.rep[(long long)r * d + c] = 0.f;
      if (lane == 0) { a.meanf[r] = 0.f; a.rstdf[r] = 0.f; }
    }
  }
  FZ_TL(2, 8);
}

// ---- backward ------------------------------------------------------------------------------------------------------
struct ChainBwdArgs {
  FfnBwdArgs f[CH_MAXB]; AttnBwdArgs at[CH_MAXB]; QkvBwdArgs q[CH_MAXB];
  int nb, M;
  const int *row_off, *tok_row;
  const float *d_rep, *xfinal, *meanf, *rstdf, *lnf_g; float* gx_top;       // final-LayerNorm backward -> gradient of the top block's output
  int* flags;                 // [CH_MAXB][stride] neighbour flags + completion counter at flags[done_at], zero at launch
  int flag_stride, done_at;
};
constexpr size_t CHAIN_BWD_SMEM = 3 * WMAT_BYTES + 3 * ATILE_BYTES + FTILE_BYTES + 2 * TM * 4 + 32;
constexpr size_t CHAIN_BWD_STAGE = WMAT_BYTES + 3 * ATILE_BYTES + FTILE_BYTES;

__global__ void __launch_bounds__(NTHR, 1) k_chain_bwd(const __grid_constant__ ChainBwdArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  op_t* S0 = reinterpret_cast<op_t*>(smem);
  op_t* S1 = reinterpret_cast<op_t*>(smem + WMAT_BYTES);
  op_t* S2 = reinterpret_cast<op_t*>(smem + 2 * WMAT_BYTES);
  op_t* A1 = reinterpret_cast<op_t*>(smem + 3 * WMAT_BYTES);
  op_t* A2 = A1 + TM * LDS;
  op_t* A3 = A2 + TM * LDS;
  float* Ft2 = reinterpret_cast<float*>(smem + 3 * WMAT_BYTES);                // aliases A1..A3 behind the products
  float* Ft = reinterpret_cast<float*>(smem + 3 * WMAT_BYTES + 3 * ATILE_BYTES);
  float* rsq = reinterpret_cast<float*>(smem + 3 * WMAT_BYTES + 3 * ATILE_BYTES + FTILE_BYTES);
  float* rskv = rsq + TM;
  const uint32_t barQ = smem_u32(rskv + TM), barKV = barQ + 8, barF = barQ + 16;
  float* att_sm = reinterpret_cast<float*>(S2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int top = a.nb - 1;
  if (tid == 0) {
    mbar_init(barQ, 1); mbar_init(barKV, 1); mbar_init(barF, 1);
    mbar_expect_tx(barF, 2 * WMAT_BYTES);
    load_wmat(smem_u32(S0), a.f[top].W2b, barF);
    load_wmat(smem_u32(S1), a.f[top].W1b, barF);
  }
  FZ_TL(3, 0);
  pdl_wait(); pdl_go();
  __syncthreads();
  FZ_TL(3, 1);
  const int T = *a.f[0].dT, d = a.f[0].d;
  const int G = gridDim.x, cta = blockIdx.x;
  const int ta = chain_cut(T, G, cta), tb = chain_cut(T, G, cta + 1);
  // higher CTAs whose gY / D rows this CTA's last session reaches into
  const int c_hi = ta < tb ? chain_owner(T, G, a.row_off[a.tok_row[tb - 1] + 1] - 1) : cta;
  const bool work = ta < tb;
  const uint64_t seed_eff = eff_seed(a.f[0].seed, a.f[0].d_step);
  const float inv_keep = a.f[0].drop_p > 0.f ? 1.f / (1.f - a.f[0].drop_p) : 1.f;

  // ---- final LayerNorm backward: the last token of a session takes LNf'(d_rep[row]), every other token zero
  for (int t = ta + warp; t < tb; t += NTHR / 32) {
    const int r = a.tok_row[t];
    float* gx = a.gx_top + (long long)t * d;
    if (t != a.row_off[r + 1] - 1) { for (int c = lane; c < d; c += 32) gx[c] = 0.f; continue; }
    ln_row_bwd(a.d_rep + (long long)r * d, a.xfinal + (long long)t * d, a.meanf[r], a.rstdf[r], a.lnf_g, gx, d, lane, false);
  }
  __syncthreads();
  FZ_TL(3, 2);
  FZ_TLV(3, 15, tb - ta);

  for (int it = 0; it < a.nb; ++it) {
    const int b = top - it;
    const uint32_t par = it & 1;
    {   // ---- FFN dgrad + LN2 backward
      const FfnBwdArgs& f = a.f[b];
      float gam[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) { const int c = lane + 32 * e; gam[e] = (c < d) ? f.ln_g[c] : 0.f; }
      for (int idx = tid; idx < TM * LDS; idx += NTHR) A2[idx] = to_op(0.f);
      __syncthreads();
      for (int t0 = ta; t0 < tb; t0 += TM)
        ffn_bwd_tile(f, t0, tb, seed_eff, A1, A2, Ft, rsq, S0, S1, gam, inv_keep, [&] { mbar_wait(barF, par); });
    }
    FZ_TL(3, 3 + 3 * it);
    if (tid == 0) {
      flag_publish(a.flags + b * a.flag_stride + cta);            // gY / D rows of this range are written
      if (work) {                                // Wk^T, Wv^T land while the attention phase runs
        fence_proxy_async();
        mbar_expect_tx(barKV, 2 * WMAT_BYTES);
        load_wmat(smem_u32(S0), a.q[b].Wkb, barKV);
        load_wmat(smem_u32(S1), a.q[b].Wvb, barKV);
      }
    }
    for (int cc = cta + 1 + tid; cc <= c_hi; cc += NTHR) flag_wait(a.flags + b * a.flag_stride + cc);
    __syncthreads();
    // ---- attention backward (staging in slot 2 + operand tiles + the fp32 tile)
    for (int t0 = ta; t0 < tb; t0 += ATT_TOK) { attn_bwd_group(a.at[b], t0, tb, att_sm, seed_eff); __syncthreads(); }
    fence_proxy_async();
    __syncthreads();
    FZ_TL(3, 4 + 3 * it);
    if (work && tid == 0) {
      mbar_expect_tx(barQ, WMAT_BYTES);
      load_wmat(smem_u32(S2), a.q[b].Wqb, barQ);
    }
    {   // ---- Q / K / V dgrad + LN1 backward
      const QkvBwdArgs& q = a.q[b];
      float gam[NE];
#pragma unroll
      for (int e = 0; e < NE; ++e) { const int c = lane + 32 * e; gam[e] = (c < d) ? q.ln_g[c] : 0.f; }
      for (int t0 = ta; t0 < tb; t0 += TM)
        qkv_bwd_tile(q, t0, tb, seed_eff, A1, A2, A3, Ft, Ft2, rsq, rskv, S2, S0, S1, gam,
                     [&] { mbar_wait(barKV, par); }, [&] { mbar_wait(barQ, par); });
    }
    FZ_TL(3, 5 + 3 * it);
    if (work && tid == 0 && it + 1 < a.nb) {
      fence_proxy_async();
      mbar_expect_tx(barF, 2 * WMAT_BYTES);
      load_wmat(smem_u32(S0), a.f[b - 1].W2b, barF);
      load_wmat(smem_u32(S1), a.f[b - 1].W1b, barF);
    }
  }
  if (!work && tid == 0) mbar_wait(barF, 0);
  // the last CTA to finish clears the flags: a second backward pass over the same forward starts clean
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    if (atomicAdd(a.flags + a.done_at, 1) == G - 1) {
      for (int b = 0; b < a.nb; ++b) for (int c = 0; c < G; ++c) a.flags[b * a.flag_stride + c] = 0;
      a.flags[a.done_at] = 0;
    }
  }
}

}  // namespace fz
}  // namespace ader
