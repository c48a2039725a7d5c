// tcgen05 weight-pass probe: the core of the tensor-core beam pass, stand-alone and self-checking.
//
//   OUT[r][n] = sum_k W[r][k] * Hs[n][k]      r < 2304 (= 3H + H + D rows of W_hh, W1, W2), k < 512, n < N columns
//
// computed the way the beam kernel's pass does it:
//   * W is split on the host into two fp16 planes  W*2^sw = Whi + Wlo  (22 significant bits, fp32-grade) and
//     streamed as the A operand through a shared-memory ring by tensor-map TMA (cp.async.bulk.tensor.2d,
//     128-byte swizzle, boxes of 128 rows x 64 k);
//   * the hidden columns are split by the CTA's threads (Hs*2^sh = Hhi + Hlo) into the canonical K-major
//     128B-swizzled layout and stay in shared memory as the B operand;
//   * one thread issues 3 x tcgen05.mma.kind::f16 (hi*hi + lo*hi + hi*lo) per 16-wide k step into a TMEM
//     accumulator (128 rows x N columns fp32, double-buffered over row tiles);
//   * four epilogue warps read the accumulator back with tcgen05.ld and write OUT.
// Prints max |error| against an fp64 reference and the cycles one pass takes (1 CTA per SM, all SMs busy, so
// L2 / shared-memory contention is the real one).  Every wait is bounded: a protocol bug traps instead of hanging.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_gemm_probe tc_gemm_probe.cu && timeout 120 ./tc_gemm_probe
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

constexpr int kRows = 2304, kK = 512, kMT = kRows / 128, kKA = kK / 64;  // 18 row tiles, 8 k-atoms of 64
constexpr int kStageBytes = 32 * 1024;                                   // hi box + lo box
constexpr int kThreads = 192;                                            // 4 epilogue warps + TMA warp + MMA warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// bounded wait: ~2 s at 2 GHz, then trap (a wrong barrier protocol must not hang the box)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)
        : "memory");
    if (!ok && clock64() - t0 > 4000000000ll) {
      printf("mbar_wait timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}

// shared-memory matrix descriptor, K-major, 128-byte swizzle (cute/arch/mma_sm100_desc.hpp: SmemDescriptor):
// rows of 128 B (64 fp16), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                  // leading byte offset (unused for swizzled K-major; CUTLASS writes 1)
  d |= (uint64_t)(1024 >> 4) << 32;        // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                  // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                  // layout type 2 = SWIZZLE_128B
  return d;
}
// instruction descriptor: fp32 accumulate, fp16 x fp16, both operands K-major, M = 128
__host__ __device__ inline uint32_t make_idesc_f16(int N) {
  uint32_t d = 0;
  d |= 1u << 4;                     // c_format = F32
  d |= 0u << 7;                     // a_format = F16
  d |= 0u << 10;                    // b_format = F16
  d |= (uint32_t)(N >> 3) << 17;    // n_dim
  d |= (uint32_t)(128 >> 4) << 24;  // m_dim
  return d;
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

struct Params {
  const float* hs;   // [N][512] fp32 columns
  float* out;        // [ctas][2304][N]   (only CTA 0's copy is checked)
  long long* cycles; // [ctas][4]
  int N, stages, mode;  // mode 0 = full; 1 = TMA only (no MMA); 2 = MMA only (ring filled once, no TMA in the loop)
  float sh, inv_scale;
  int write_all;
  int commit_every, wait_every;  // mode 2 only: issue-side overhead experiments (commit to a dummy barrier / re-test a completed barrier every n MMAs)
};

__global__ void __launch_bounds__(kThreads, 1) tc_pass_kernel(const __grid_constant__ CUtensorMap wmap, const Params p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // manual 1024-byte alignment (the swizzle pattern is a function of the shared-memory address bits)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int N = p.N, S = p.stages;
  unsigned char* ring = smem;                                   // S x 32 KB
  unsigned char* bhi = ring + (size_t)S * kStageBytes;          // [8 atoms][N rows][128 B]
  unsigned char* blo = bhi + (size_t)N * 1024;
  uint64_t* bars = reinterpret_cast<uint64_t*>(blo + (size_t)N * 1024);
  uint64_t* full = bars;            // [S]
  uint64_t* empty = bars + S;       // [S]
  uint64_t* tfull = bars + 2 * S;   // [2]
  uint64_t* tempty = tfull + 2;     // [2]
  uint64_t* dummy = tempty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dummy + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 128); }
    mbar_init(dummy, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {  // the MMA warp owns the TMEM allocation: 2 accumulators of N columns -> power of two >= 32
    uint32_t cols = 32;
    while (cols < 2u * N) cols <<= 1;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // B operand: split every hidden column into fp16 hi / lo planes, K-major 128B-swizzled:
  //   element (n, k) -> atom k/64, row n, 16-byte chunk ((k%64)/8) ^ (n%8), byte (k%8)*2
  for (int i = tid; i < N * kK; i += kThreads) {
    const int n = i / kK, k = i % kK;
    const float v = p.hs[(size_t)n * kK + k] * p.sh;
    const __half hi = __float2half_rn(v);
    const __half lo = __float2half_rn(v - __half2float(hi));
    const uint32_t off = (uint32_t)(k / 64) * (uint32_t)N * 128u + (uint32_t)n * 128u +
                         ((((uint32_t)(k % 64) / 8u) ^ ((uint32_t)n & 7u)) * 16u) + (uint32_t)(k % 8) * 2u;
    *reinterpret_cast<__half*>(bhi + off) = hi;
    *reinterpret_cast<__half*>(blo + off) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int total = kMT * kKA;  // ring tiles per pass
  if (warp == 4) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      const int ntiles = (p.mode == 2) ? S : total;
      for (int it = 0; it < ntiles; ++it) {
        const int s = it % S, ph = (it / S) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        const int mt = it / kKA, ka = it % kKA;
        mbar_expect_tx(&full[s], kStageBytes);
        tma_load_2d(ring + (size_t)s * kStageBytes, &wmap, ka * 64, mt * 128, &full[s]);                   // hi plane
        tma_load_2d(ring + (size_t)s * kStageBytes + 16384, &wmap, ka * 64, kRows + mt * 128, &full[s]);  // lo plane
      }
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(N);
      const long long t0 = clock64();
      for (int mt = 0; mt < kMT; ++mt) {
        const int buf = mt & 1;
        mbar_wait(&tempty[buf], ((mt >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * N);
        for (int ka = 0; ka < kKA; ++ka) {
          const int it = mt * kKA + ka;
          const int s = it % S, ph = (it / S) & 1;
          if (p.mode != 2 || it < S) mbar_wait(&full[s], (p.mode == 2) ? 0 : ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (p.mode != 1) {
            const uint32_t a_hi = smem_u32(ring + (size_t)s * kStageBytes), a_lo = a_hi + 16384;
            const uint32_t b_hi = smem_u32(bhi) + (uint32_t)ka * (uint32_t)N * 128u;
            const uint32_t b_lo = smem_u32(blo) + (uint32_t)ka * (uint32_t)N * 128u;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {  // 4 k-steps of 16 inside the 64-wide swizzle atom: +32 bytes each
              const uint64_t dah = make_desc_sw128(a_hi + kk * 32), dal = make_desc_sw128(a_lo + kk * 32);
              const uint64_t dbh = make_desc_sw128(b_hi + kk * 32), dbl = make_desc_sw128(b_lo + kk * 32);
              mma_f16(d_tmem, dah, dbh, idesc, (ka | kk) != 0);
              mma_f16(d_tmem, dal, dbh, idesc, 1);
              mma_f16(d_tmem, dah, dbl, idesc, 1);
              if (p.mode == 2) {
                const int nm = ((mt * kKA + ka) * 4 + kk + 1) * 3;  // MMAs issued so far
                if (p.commit_every && nm % p.commit_every == 0) mma_commit(dummy);
                if (p.wait_every && nm % p.wait_every == 0) {
                  mbar_wait(&full[0], 0);  // phase 0 completed long ago: the cost of a successful test
                  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
              }
            }
          }
          if (p.mode != 2) mma_commit(&empty[s]);  // frees the ring slot when the MMAs above have read it
        }
        mma_commit(&tfull[buf]);
      }
      p.cycles[blockIdx.x * 4 + 0] = clock64() - t0;  // issue time
    }
  } else {
    // ------------------------------------------------------------ epilogue warps 0..3: TMEM lanes 32*warp ..
    const long long t0 = clock64();
    for (int mt = 0; mt < kMT; ++mt) {
      const int buf = mt & 1;
      mbar_wait(&tfull[buf], (mt >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = mt * 128 + warp * 32 + lane;
      float* orow = p.out + ((size_t)blockIdx.x * kRows + row) * N;
      for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * N + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (p.write_all || blockIdx.x == 0)
#pragma unroll
          for (int q = 0; q < 16; ++q) orow[c0 + q] = __uint_as_float(v[q]) * p.inv_scale;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&tempty[buf]);
    }
    if (tid == 0) p.cycles[blockIdx.x * 4 + 1] = clock64() - t0;  // whole pass as seen by the epilogue
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) {
    uint32_t cols = 32;
    while (cols < 2u * N) cols <<= 1;
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols));
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);

  // weights in (-0.1, 0.1) like the trained fixture, hidden columns in (-1, 1)
  std::vector<float> W((size_t)kRows * kK), Hs((size_t)64 * kK);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (double)(st >> 11) / 9007199254740992.0; };
  for (auto& w : W) w = (float)((rnd() * 2 - 1) * 0.1);
  for (auto& h : Hs) h = (float)std::tanh((rnd() * 2 - 1) * 1.5);
  const float sw = 1024.f, sh = 256.f;
  std::vector<__half> planes((size_t)2 * kRows * kK);
  for (size_t i = 0; i < W.size(); ++i) {
    const float v = W[i] * sw;
    const __half hi = __float2half_rn(v);
    planes[i] = hi;
    planes[W.size() + i] = __float2half_rn(v - __half2float(hi));
  }
  __half* d_planes;
  float *d_hs, *d_out;
  long long* d_cyc;
  CK(cudaMalloc(&d_planes, planes.size() * 2));
  CK(cudaMemcpy(d_planes, planes.data(), planes.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_hs, Hs.size() * 4));
  CK(cudaMemcpy(d_hs, Hs.data(), Hs.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_out, (size_t)kRows * 64 * 4));
  CK(cudaMalloc(&d_cyc, (size_t)sms * 4 * 8));

  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode), cudaEnableDefault, &qres));
  if (!encode || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not available\n"); return 2; }
  alignas(64) CUtensorMap wmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)kK, (cuuint64_t)2 * kRows};
  const cuuint64_t gstr[1] = {(cuuint64_t)kK * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  CUresult cr = encode(&wmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_planes, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 2; }

  // fp64 reference of the fp32 inputs
  std::vector<double> ref((size_t)kRows * 64);
  for (int r = 0; r < kRows; ++r)
    for (int n = 0; n < 64; ++n) {
      double s = 0;
      for (int k = 0; k < kK; ++k) s += (double)W[(size_t)r * kK + k] * (double)Hs[(size_t)n * kK + k];
      ref[(size_t)r * 64 + n] = s;
    }
  // fp32 FMA chain in k order, for scale
  double fp32_err = 0;
  for (int r = 0; r < kRows; r += 7)
    for (int n = 0; n < 64; n += 5) {
      float s = 0;
      for (int k = 0; k < kK; ++k) s = fmaf(W[(size_t)r * kK + k], Hs[(size_t)n * kK + k], s);
      fp32_err = std::max(fp32_err, std::fabs((double)s - ref[(size_t)r * 64 + n]));
    }
  printf("reference: max |fp32 fma chain - fp64| = %.3e (sampled)\n", fp32_err);

  CK(cudaFuncSetAttribute(tc_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  const int only_n = argc > 1 ? atoi(argv[1]) : 0;
  for (int N : {16, 32, 48, 64}) {
    if (only_n && N != only_n) continue;
    for (int S : {3, 4}) {
      const size_t smem = 1024 + (size_t)S * kStageBytes + 2 * (size_t)N * 1024 + 256;
      if (smem > 227 * 1024) continue;
      for (int mode : {0, 1, 2}) {
        for (int grid : {1, sms}) {
          if (mode != 0 && grid == 1) continue;
          Params p{d_hs, d_out, d_cyc, N, S, mode, sh, 1.0f / (sw * sh), 0, 0, 0};
          CK(cudaMemset(d_out, 0, (size_t)kRows * 64 * 4));
          double best = 1e30, best_issue = 1e30;
          for (int rep = 0; rep < 3; ++rep) {
            tc_pass_kernel<<<grid, kThreads, smem>>>(wmap, p);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("N=%d S=%d mode=%d grid=%d: %s\n", N, S, mode, grid, cudaGetErrorString(e)); return 1; }
            std::vector<long long> cyc((size_t)grid * 4);
            CK(cudaMemcpy(cyc.data(), d_cyc, cyc.size() * 8, cudaMemcpyDeviceToHost));
            double mx = 0, mi = 0;
            for (int b = 0; b < grid; ++b) { mx = std::max(mx, (double)cyc[b * 4 + 1]); mi = std::max(mi, (double)cyc[b * 4 + 0]); }
            best = std::min(best, mx);
            best_issue = std::min(best_issue, mi);
          }
          double err = -1;
          if (mode == 0) {
            std::vector<float> out((size_t)kRows * N);
            CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
            err = 0;
            for (int r = 0; r < kRows; ++r)
              for (int n = 0; n < N; ++n) err = std::max(err, std::fabs((double)out[(size_t)r * N + n] - ref[(size_t)r * 64 + n]));
          }
          printf("N=%2d stages=%d mode=%d (%s) grid=%3d: pass %8.0f cycles (slowest CTA; issue %8.0f) = %6.1f us @1.965GHz, "
                 "%5.1f cycles per 128xNx16 MMA, max|err| %.3e\n",
                 N, S, mode, mode == 0 ? "tma+mma" : (mode == 1 ? "tma only" : "mma only"), grid, best, best_issue,
                 best / 1965.0, best / (kMT * kKA * 12.0), err);
        }
      }
    }
  }
  // issue-side overhead of the per-box protocol (MMA only, N = 48, one CTA per SM)
  for (int ce : {0, 12, 6, 3})
    for (int we : {0, 12, 6, 3}) {
      const int N = 48, S = 3;
      const size_t smem = 1024 + (size_t)S * kStageBytes + 2 * (size_t)N * 1024 + 256;
      Params p{d_hs, d_out, d_cyc, N, S, 2, sh, 1.0f / (sw * sh), 0, ce, we};
      double best = 1e30;
      for (int rep = 0; rep < 3; ++rep) {
        tc_pass_kernel<<<sms, kThreads, smem>>>(wmap, p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("overhead probe: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> cyc((size_t)sms * 4);
        CK(cudaMemcpy(cyc.data(), d_cyc, cyc.size() * 8, cudaMemcpyDeviceToHost));
        double mx = 0;
        for (int b = 0; b < sms; ++b) mx = std::max(mx, (double)cyc[b * 4 + 1]);
        best = std::min(best, mx);
      }
      printf("mma only N=48: commit every %2d MMAs, barrier test every %2d MMAs: %6.1f cycles per MMA\n", ce, we,
             best / (kMT * kKA * 12.0));
    }
  return 0;
}
