// Microbenchmark: throughput of legacy warp-level mma.sync m16n8k8 tf32 (and FFMA2 for reference) on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void mma_k(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) c[i][q] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) s += c[i][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma2_k(float* out, int iters) {
  float2 c[32];
  for (int i = 0; i < 32; ++i) c[i] = make_float2(0.f, 0.f);
  float2 x = make_float2(threadIdx.x * 1e-3f, 1.f);
  float w = 1.0001f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = __ffma2_rn(x, make_float2(w, w), c[i]);
  }
  float s = 0; for (int i = 0; i < 32; ++i) s += c[i].x + c[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps : {4, 8, 16}) {
    int iters = 20000; float ms;
    mma_k<<<148, warps * 32>>>(out, 10);
    cudaEventRecord(e0); mma_k<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double macs = 148.0 * warps * iters * 8 * (16 * 8 * 8);
    printf("mma.sync tf32 m16n8k8: %2d warps/SM: %.3f ms, %.1f TMAC/s = %.0f MAC/clk/SM @1.965GHz\n", warps, ms, macs / ms / 1e9, macs / (ms * 1e-3) / 148 / 1.965e9);
    ffma2_k<<<148, warps * 32>>>(out, 10);
    cudaEventRecord(e0); ffma2_k<<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    double fmas = 148.0 * warps * 32 * iters * 32 * 2;
    printf("FFMA2                : %2d warps/SM: %.3f ms, %.1f TFMA/s = %.0f FMA/clk/SM\n", warps, ms, fmas / ms / 1e9, fmas / (ms * 1e-3) / 148 / 1.965e9);
  }
  return 0;
}
