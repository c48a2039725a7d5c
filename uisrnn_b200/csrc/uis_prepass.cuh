// Pre-pass kernels (run once per uis_predict call, before the persistent beam kernel):
//   * cast_f64_f32:  the reference's `torch.from_numpy(seq).float()` (uisrnn.py:525-526), done on
//     the device so the host never touches the data (round-to-nearest-even = numpy/torch cast).
//   * input_proj_kernel:  gi[n][:] = W_ih x_n + b_ih for every DISTINCT frame of every utterance.
//     In the reference this product is recomputed inside nn.GRU for every (hypothesis, candidate)
//     and again for every winner (uisrnn.py:422-424, 448-450, 557-558); it depends on the frame
//     only, so it is one fp32 GEMM [rows x D] x [D x 3H] up front (test_iteration tiling reuses rows).
//   * init_state_kernel: (mean0, hidden0) = CoreRNN(zeros, rnn_init_hidden), uisrnn.py:435-439.
#pragma once
#include "uis_common.cuh"

namespace uis {

__global__ void cast_f64_f32_kernel(const double* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __double2float_rn(in[i]);
}

// Models smaller than the kernel shape run zero-padded (uis_model_create): rows of d values -> rows of dp >= d values.
__global__ void cast_pad_f64_f32_kernel(const double* __restrict__ in, float* __restrict__ out, size_t rows, int d, int dp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x, n = rows * (size_t)dp;
  for (; i < n; i += stride) {
    const size_t r = i / dp;
    const int c = (int)(i % dp);
    out[i] = c < d ? __double2float_rn(in[r * d + c]) : 0.f;
  }
}
__global__ void pad_rows_f32_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, int d, int dp) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x, n = rows * (size_t)dp;
  for (; i < n; i += stride) {
    const size_t r = i / dp;
    const int c = (int)(i % dp);
    out[i] = c < d ? in[r * d + c] : 0.f;
  }
}

// C[M][N] = A[M][K] * Bt[K][N] + bias[N]   (all fp32, row-major; K % 4 == 0, N % 4 == 0)
// 128x128 CTA tile, BK = 16, 256 threads, 8x8 register micro-tile, fp32 FMA (k ascending).
constexpr int PBM = 128, PBN = 128, PBK = 16;
__global__ void __launch_bounds__(256) input_proj_kernel(const float* __restrict__ A, const float* __restrict__ Bt,
                                                         const float* __restrict__ bias, float* __restrict__ C,
                                                         int M, int N, int K) {
  __shared__ __align__(16) float As[PBK][PBM + 4];
  __shared__ __align__(16) float Bs[PBK][PBN];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * PBM, n0 = blockIdx.x * PBN;
  const int tx = tid % 16, ty = tid / 16;  // micro-tile: rows ty*8.., cols tx*4 and 64+tx*4
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[i][q] = 0.f;

  for (int k0 = 0; k0 < K; k0 += PBK) {
    // A tile: 128 rows x 16 k = 512 float4, 2 per thread
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int f = tid + r * 256;
      const int row = f / 4, kq = (f % 4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + row < M && k0 + kq < K) v = *reinterpret_cast<const float4*>(A + (size_t)(m0 + row) * K + k0 + kq);
      As[kq + 0][row] = v.x; As[kq + 1][row] = v.y; As[kq + 2][row] = v.z; As[kq + 3][row] = v.w;
    }
    // B tile: 16 k x 128 cols = 512 float4, 2 per thread
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int f = tid + r * 256;
      const int kk = f / 32, nq = (f % 32) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < K && n0 + nq < N) v = *reinterpret_cast<const float4*>(Bt + (size_t)(k0 + kk) * N + n0 + nq);
      *reinterpret_cast<float4*>(&Bs[kk][nq]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < PBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[i][q] = fmaf(a[i], b[q], acc[i][q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + ty * 8 + i;
    if (row >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int coln = n0 + h * 64 + tx * 4;
      if (coln < N) {
        const float4 bv = *reinterpret_cast<const float4*>(bias + coln);
        float4 o;
        o.x = __fadd_rn(acc[i][h * 4 + 0], bv.x); o.y = __fadd_rn(acc[i][h * 4 + 1], bv.y);
        o.z = __fadd_rn(acc[i][h * 4 + 2], bv.z); o.w = __fadd_rn(acc[i][h * 4 + 3], bv.w);
        *reinterpret_cast<float4*>(C + (size_t)row * N + coln) = o;
      }
    }
  }
}

// One CTA of H threads.  Weights in the k-major layouts used by the beam kernel.
// whh_t: [depth][H][3H]; wih_up_t: [depth-1][H][3H] (layers >= 1); bih: [depth][3H]; bhh: [depth][3H]; h0: [depth][H].
__global__ void __launch_bounds__(1024) init_state_kernel(const float* __restrict__ whh_t, const float* __restrict__ wih_up_t,
                                  const float* __restrict__ w1_t, const float* __restrict__ w2_t,
                                  const float* __restrict__ bih, const float* __restrict__ bhh,
                                  const float* __restrict__ b1, const float* __restrict__ b2,
                                  const float* __restrict__ h0, int H, int D, int depth, float* __restrict__ mean0,
                                  float* __restrict__ hidden0) {
  extern __shared__ float sm[];  // [H] layer state, [H] layer input / output, [H] act
  float* shp = sm; float* shx = sm + H; float* sa = sm + 2 * H;
  const int j = threadIdx.x;
  for (int l = 0; l < depth; ++l) {
    if (j < H) shp[j] = h0[(size_t)l * H + j];
    __syncthreads();
    float hn = 0.f;
    if (j < H) {
      const float* wh = whh_t + (size_t)l * H * 3 * H;
      float ar = 0.f, az = 0.f, an = 0.f;
      for (int k = 0; k < H; ++k) {
        const float x = shp[k];
        ar = fmaf(wh[(size_t)k * 3 * H + j], x, ar);
        az = fmaf(wh[(size_t)k * 3 * H + H + j], x, az);
        an = fmaf(wh[(size_t)k * 3 * H + 2 * H + j], x, an);
      }
      // layer 0: x = 0  =>  W_ih x + b_ih = b_ih exactly; layer l >= 1: x = h'_{l-1}
      float ir = 0.f, iz = 0.f, in = 0.f;
      if (l > 0) {
        const float* wi = wih_up_t + (size_t)(l - 1) * H * 3 * H;
        for (int k = 0; k < H; ++k) {
          const float x = shx[k];
          ir = fmaf(wi[(size_t)k * 3 * H + j], x, ir);
          iz = fmaf(wi[(size_t)k * 3 * H + H + j], x, iz);
          in = fmaf(wi[(size_t)k * 3 * H + 2 * H + j], x, in);
        }
      }
      const float* bi = bih + (size_t)l * 3 * H;
      const float* bh = bhh + (size_t)l * 3 * H;
      const float r = sigmoid_f32(__fadd_rn(__fadd_rn(ir, bi[j]), __fadd_rn(ar, bh[j])));
      const float z = sigmoid_f32(__fadd_rn(__fadd_rn(iz, bi[H + j]), __fadd_rn(az, bh[H + j])));
      const float n = tanhf(__fadd_rn(__fadd_rn(in, bi[2 * H + j]), __fmul_rn(r, __fadd_rn(an, bh[2 * H + j]))));
      hn = __fadd_rn(__fmul_rn(__fsub_rn(shp[j], n), z), n);
      hidden0[(size_t)l * H + j] = hn;
    }
    __syncthreads();
    if (j < H) shx[j] = hn;
    __syncthreads();
  }
  if (j < H) {
    float a = 0.f;
    for (int k = 0; k < H; ++k) a = fmaf(w1_t[(size_t)k * H + j], shx[k], a);
    sa[j] = fmaxf(__fadd_rn(a, b1[j]), 0.f);
  }
  __syncthreads();
  if (j < D) {
    float a = 0.f;
    for (int k = 0; k < H; ++k) a = fmaf(w2_t[(size_t)k * D + j], sa[k], a);
    mean0[j] = __fadd_rn(a, b2[j]);
  }
}

}  // namespace uis
