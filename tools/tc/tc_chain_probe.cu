// tcgen05 dependency-chain probe: is the ~146 cycles per 128 x N x 16 kind::f16 MMA that the beam kernel's pass sees
// (uis_beam_tc.cuh, N' = 64 or 96, a new 4 KB A tile every instruction) a THROUGHPUT limit of shared-memory-fed MMAs,
// or the LATENCY of back-to-back accumulations into the same TMEM accumulator?  Issues 4096 MMAs round-robin over
// `nacc` independent accumulators (different TMEM column ranges) and reports cycles per MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_chain_probe.bin tc_chain_probe.cu && timeout 60 ./tc_chain_probe.bin
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  return d;  // SWIZZLE_NONE
}
__host__ __device__ inline uint32_t make_idesc(int N) {  // F32 accumulate, F16 x F16, K-major, M = 128
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
constexpr int kATiles = 24;              // distinct 4 KB A tiles cycled through (a new one per MMA)
constexpr int kATileBytes = 128 * 16 * 2;
constexpr int kBBytes = 256 * 128;        // B operand atom (N <= 256 rows of 128 B)

__global__ void __launch_bounds__(128, 1) chain_kernel(int N, int iters, int nacc, int same_a, int commit_every, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar, bar2;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (kATiles * kATileBytes + kBBytes) / 2; i += 128)
    reinterpret_cast<unsigned short*>(smem)[i] = (unsigned short)(0x3c00u + (i & 255));  // fp16 values in [1, 1.25)
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_slot;
  if (warp == 1) {  // the whole warp runs the loop with uniform operands; one elected lane issues (as the beam kernel does)
    const uint32_t a16 = smem_u32(smem) >> 4, b16 = smem_u32(smem + kATiles * kATileBytes) >> 4;
    const uint32_t idesc = make_idesc(N);
    // SWIZZLE_128B K-major descriptors (the beam kernel's layout): rows of 128 B, 8-row groups 1024 B apart
    const uint64_t desc0 = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t bdesc = desc0 + b16;
    const long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      const uint64_t adesc = desc0 + (uint64_t)(a16 + (uint32_t)(same_a ? 0 : ((i / 4) % (kATiles / 4))) * (4 * kATileBytes >> 4));
      uint32_t pred;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      if (pred) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t d = tmem_d + (uint32_t)(((i + kk) % nacc) * N);
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(d), "l"(adesc + 2 * kk), "l"(bdesc + 2 * kk), "r"(idesc), "r"((uint32_t)(i + kk >= nacc))
              : "memory");
        }
        if (commit_every == 2)  // (a second commit in the middle of the group is emulated by two per group)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        if (commit_every && (i % (commit_every < 4 ? 4 : commit_every)) == 0)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 32) {
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
      const long long t2 = clock64();
      out[blockIdx.x * 2 + 0] = t1 - t0;
      out[blockIdx.x * 2 + 1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(512));
}

int main() {
  long long* out;
  cudaMalloc(&out, 148 * 16);
  const int smem = kATiles * kATileBytes + kBBytes;
  cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4096;
  printf("kind::f16 M=128 K=16, %d MMAs round-robin over nacc accumulators; a new 4 KB A tile per MMA unless same_a\n", iters);
  for (int grid : {148})
    for (int ce : {0, 2, 4, 8, 16})
      for (int same_a : {0, 1})
        for (int N : {96, 256})
          for (int nacc : {1}) {
            if (nacc * N > 512) continue;
            if (same_a && (ce || nacc > 1 || grid > 1)) continue;
            long long h[2] = {0, 0};
            for (int rep = 0; rep < 2; ++rep) {
              chain_kernel<<<grid, 128, smem>>>(N, iters, nacc, same_a, ce, out);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("N=%d nacc=%d: %s\n", N, nacc, cudaGetErrorString(e)); return 1; }
              cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            }
            printf("grid=%3d mmas_per_commit=%d same_a=%d N=%3d nacc=%d: %7.1f cycles/MMA (issue %6.1f)\n", grid, ce, same_a, N, nacc,
                   (double)h[1] / iters, (double)h[0] / iters);
          }
  return 0;
}
