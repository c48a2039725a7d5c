// tcgen05 issue-rate probe for the "lock-step >= 128-column beam pass" idea.
// NOT part of the product and NOT YET RUN ON HARDWARE (written at the end of round 1, after the GPU budget was
// spent); compile-checked only.  Run it under a timeout:   nvcc -gencode arch=compute_100a,code=sm_100a -o tc_probe
// tools/next/tc_probe.cu && timeout 30 ./tc_probe
//
// Question it answers: how many cycles does one  D[tmem 128 x N] += A[smem 128 x 8] * B[smem N x 8]^T  kind::tf32
// MMA take as a function of N, when A (the weights) comes from a DIFFERENT shared-memory tile every instruction
// (as in the weight-streaming pass) and B (the beam states) stays put?  If the cost is flat in N up to N ~ 64 the
// pass cost stops depending on the number of columns and more lanes per CTA become free; a 3xTF32 split
// (hi*hi + lo*hi + hi*lo) costs 3 such MMAs per k-step.
//
// Layout: canonical K-major, no swizzle: core matrices of 8 rows x 16 bytes; a 128 x 8 tf32 operand is
// 16 (row groups) x 2 (k halves) core matrices.  The numerical content is irrelevant here (random bits).
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor), SWIZZLE_NONE
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);            // start address, 16-byte units
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;  // leading-dimension byte offset (between k core matrices)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;  // stride-dimension byte offset (between 8-row groups)
  d |= (uint64_t)1 << 46;                            // descriptor version 1 (Blackwell)
  return d;                                          // layout type 0 = no swizzle, base offset 0
}
// instruction descriptor (InstrDescriptor): F32 accumulate, TF32 x TF32, both K-major, M = 128
__host__ __device__ inline uint32_t make_idesc(int N) {
  uint32_t d = 0;
  d |= 1u << 4;                  // c_format = F32
  d |= 2u << 7;                  // a_format = TF32
  d |= 2u << 10;                 // b_format = TF32
  d |= (uint32_t)(N >> 3) << 17; // n_dim
  d |= (uint32_t)(128 >> 4) << 24;  // m_dim
  return d;
}

constexpr int kATiles = 32;              // distinct A tiles (4 KB each) cycled through = 128 KB of "weights"
constexpr int kATileBytes = 128 * 8 * 4;
constexpr int kBBytes = 256 * 8 * 4;     // B operand for the largest N

__global__ void __launch_bounds__(128, 1) tc_probe_kernel(int N, int iters, int a_stride_tiles, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // fill the operands with something finite
  for (int i = tid; i < (kATiles * kATileBytes + kBBytes) / 4; i += 128)
    reinterpret_cast<float*>(smem)[i] = 1.0f + (float)(i & 1023) * 9.765625e-4f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {  // one warp allocates all 512 TMEM columns (accumulator: N fp32 columns x 128 lanes)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;

  if (tid == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + kATiles * kATileBytes);
    const uint32_t idesc = make_idesc(N);
    // A tile: [k half (2)][row group (16)][8 rows][16 B]  ->  LBO = 16 * 128 B, SBO = 128 B
    // B tile: [k half (2)][row group (N/8)][8 rows][16 B] ->  LBO = (N/8) * 128 B, SBO = 128 B
    const uint64_t bdesc = make_desc(b0, (uint32_t)(N / 8) * 128u, 128u);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t adesc = make_desc(a0 + (uint32_t)((i * a_stride_tiles) % kATiles) * kATileBytes, 16u * 128u, 128u);
      const uint32_t accumulate = i > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
          : "memory");
    }
    const long long t1 = clock64();  // issue done (the MMAs are asynchronous)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();  // all MMAs retired
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512));
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  const int smem = kATiles * kATileBytes + kBBytes;
  cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  printf("kind::tf32  M=128 K=8, %d MMAs per run; A tile stride 1 = a new 4 KB weight tile per MMA, 0 = same tile\n", iters);
  for (int stride : {1, 0})
    for (int N : {16, 32, 64, 128, 256}) {
      long long h[2] = {0, 0};
      for (int rep = 0; rep < 2; ++rep) {
        tc_probe_kernel<<<1, 128, smem>>>(N, iters, stride, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
      }
      const double cyc = (double)h[1] / iters;
      printf("A stride %d  N=%3d: %7.1f cycles/MMA (issue %5.1f)  -> %7.1f tf32 MAC/clk/SM, 3xTF32 equivalent %6.1f fp32-grade MAC/clk/SM\n",
             stride, N, cyc, (double)h[0] / iters, 128.0 * N * 8 / cyc, 128.0 * N * 8 / cyc / 3.0);
    }
  printf("reference points: packed FFMA2 = 126 FMA/clk/SM measured (tools/mma_bench.cu), beam kernel today ~ 50 %% of that\n");
  return 0;
}
