// Microbenchmark: latency / throughput of the legacy warp-level mma.sync.m16n8k16 (fp16 -> fp32) on sm_100a.
// For NCH independent accumulator chains per warp, DEPTH dependent steps each, W warps per CTA, one CTA per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int NCH, int DEPTH, bool TF32>
__global__ void k_probe(long long* out, float* sink, uint32_t seed) {
  float acc[NCH][4];
#pragma unroll
  for (int j = 0; j < NCH; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
  uint32_t a0 = seed + threadIdx.x, a1 = seed * 3 + threadIdx.x, a2 = seed * 5, a3 = seed * 7, b0 = seed * 11 + threadIdx.x, b1 = seed * 13;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int rep = 0; rep < 16; ++rep) {
#pragma unroll
    for (int k = 0; k < DEPTH; ++k)
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        if (TF32) mma_tf32(acc[j], a0, a1, a2, a3, b0 + j, b1 + k);
        else mma_16816(acc[j], a0, a1, a2, a3, b0 + j, b1 + k);
      }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NCH; ++j) s += acc[j][0] + acc[j][1] + acc[j][2] + acc[j][3];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int NCH, int DEPTH, bool TF32>
void run(int warps, long long* d_out, float* d_sink) {
  k_probe<NCH, DEPTH, TF32><<<148, warps * 32>>>(d_out, d_sink, 1);
  k_probe<NCH, DEPTH, TF32><<<148, warps * 32>>>(d_out, d_sink, 1);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
  const double n = 16.0 * DEPTH * NCH;
  printf("%s chains %2d depth %2d warps %2d: %8lld cycles, %.1f cycles per mma per warp, %.2f mma/cycle/SM\n", TF32 ? "tf32 m16n8k8 " : "f16 m16n8k16", NCH, DEPTH,
         warps, c, c / n, n * warps / c);
}

int main() {
  long long* d_out; float* d_sink;
  cudaMalloc(&d_out, 8); cudaMalloc(&d_sink, 4);
  for (int w : {1, 4, 8, 16}) {
    run<1, 10, false>(w, d_out, d_sink);
    run<5, 10, false>(w, d_out, d_sink);
    run<10, 10, false>(w, d_out, d_sink);
    run<25, 4, false>(w, d_out, d_sink);
    run<25, 4, true>(w, d_out, d_sink);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
