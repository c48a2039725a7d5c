// Persistent beam-search kernel for UIS-RNN predict() on sm_100a  (look_ahead = 1, depth = 1).
//
// What it replaces (all under /root/reference/uisrnn/): the whole loop body of
// UISRNN.predict_single (uisrnn.py:529-561) -- _calculate_score (:455-477), the np.sort/argsort
// top-k (:546-552), _update_beam_state for the winners (:388-453), CoreRNN.forward (:45-52) and
// loss_func.weighted_mse_loss (loss_func.py:19-41) -- with BeamState (:55-77) kept on device.
//
// Mapping to the hardware
//   * one persistent CTA per SM; a CTA pulls utterances (longest first) from a global queue and
//     runs ALL test_iteration*N beam steps of an utterance without returning to the host.
//   * per step the only heavy work is  h' = GRU(x_t, h_src),  mean = W2 relu(W1 h' + b1) + b2  for
//     the M <= beam_size DISTINCT source states of the step's winners: three skinny
//     (rows x 512) x (512 x M) products.  The fp32 weights (4.7 MB, L2-resident) are streamed
//     through a 4-stage shared-memory ring by a producer warp with 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx); 512 consumer threads each own one hidden unit
//     (3 gate rows) and keep 3*M accumulators in registers, so each weight element fetched from
//     L2 feeds M FMAs and no cross-thread reduction is needed.
//   * hypothesis state is a slot pool in global memory (L2): slot = (mean[D], hidden[H]) written
//     once and never modified; a hypothesis is a table of (slot, block count, visit count) per
//     cluster held in shared memory.  A child differs from its parent in ONE table entry, so the
//     re-pack after the top-k is an index shuffle plus M slot writes (BeamState copies in the
//     reference are shallow too, uisrnn.py:66-69).
//   * scoring needs no GRU at look_ahead 1 (uisrnn.py:411-414 uses the pre-update mean;
//     :438-443 uses the per-model constant CoreRNN(0, h0)), so candidates are scored first
//     (one warp per candidate, fp64 log terms from host-built tables), ranked by counting, and
//     only the winners' distinct source states go through the GRU.
#pragma once
#include "uis_common.cuh"

namespace uis {

constexpr int kStages = 4;            // weight ring depth
constexpr int kStageBytes = 24 * 1024; // bytes per ring stage
constexpr int kInitSlot = 0;          // pool slot holding (mean0, hidden0)

struct TabEntry {
  int slot;    // index into the CTA's slot pool
  int blocks;  // block_counts[c]        (uisrnn.py:431-432, 451)
  int visits;  // #{trace == c}          (uisrnn.py:425-428)
  int pad;
};

struct BeamParams {
  // model (device pointers)
  const float* whh_t;    // [H][3H]   = gru.weight_hh_l0 transposed (k-major)
  const float* w1_t;     // [H][H]    = linear_mean1.weight transposed
  const float* w2_t;     // [H][D]    = linear_mean2.weight transposed
  const float* bhh;      // [3H]
  const float* b1;       // [H]
  const float* b2;       // [D]
  const float* wvec;     // [D]  1 / (2 sigma2)
  const float* mean0;    // [D]
  const float* hidden0;  // [H]
  double log_p0, log_1mp0, log_alpha;
  const double* logn;    // [>= maxTN + 2]  log(i)
  const double* logtot;  // [>= maxTN + 2]  log(i + crp_alpha)
  // inputs
  const float* x;           // [rows][D]
  const float* gi;          // [rows][3H]  W_ih x + b_ih
  const long long* row_off; // [U + 1]
  const int* order;         // [U] utterance ids, longest first
  int U, B, Kcap, T, P, maxN;
  // per-CTA workspace
  float* pool_mean;    // [ctas][P][D]
  float* pool_hidden;  // [ctas][P][H]
  unsigned* bp;        // [ctas][maxN][B]  (parent << 16) | cluster
  int* queue;          // [1] next position in `order`
  // outputs
  int* labels;  // [rows]
  int* status;  // [U]   0 ok, -4 overflow
  unsigned long long* stats;  // [8]
  // debug taps (device buffers, may be null)
  int trace_utt;
  int trace_capacity;
  int* dbg_win;            // [cap][2]
  float* dbg_score;        // [cap]
  long long* dbg_off;      // [steps + 1]
  float* dbg_final_scores; // [U][B]
  int* dbg_final_k;        // [U]
  float* dbg_best_mean;    // [Kcap][D]
  float* dbg_best_hidden;  // [Kcap][H]
  int* dbg_best_blocks;    // [Kcap]
};

struct SmemLayout {
  unsigned ring, xa, xb, xt, gi, wv, tabs, meta, candoff, keys, svals, wins, cols, used, bars,
      misc, total;
};

__host__ __device__ inline unsigned align_up(unsigned v, unsigned a) { return (v + a - 1) / a * a; }

template <int H, int D, int MP>
__host__ __device__ inline SmemLayout make_layout(int B, int Kcap) {
  SmemLayout L;
  unsigned o = 0;
  L.ring = o;    o += kStages * kStageBytes;
  L.xa = o;      o += H * MP * 4;
  L.xb = o;      o += H * MP * 4;
  L.xt = o;      o += 2 * D * 4;
  L.gi = o;      o += 2 * 3 * H * 4;
  L.wv = o;      o += D * 4;
  L.tabs = o;    o += 2u * B * Kcap * 16;
  L.meta = o;    o += 2u * 4 * B * 4;            // K,last,tot,nl  x2 generations
  L.candoff = o; o += align_up((B + 1) * 4, 16);
  const unsigned ne = (unsigned)B * (Kcap + 1);
  L.keys = o;    o += align_up(ne * 8, 16);
  L.svals = o;   o += align_up(ne * 4, 16);
  L.wins = o;    o += align_up(3u * B * 4, 16);
  L.cols = o;    o += align_up(5u * B * 4, 16);  // col, colsrc, colnew, colvis, colblk
  const unsigned pw = ((unsigned)B * Kcap + B + 1 + 31) / 32;
  L.used = o;    o += align_up(pw * 4, 16);
  L.bars = o;    o += 2 * kStages * 8;
  L.misc = o;    o += 64;
  L.total = o;
  return L;
}

// misc[] indices
enum { MI_PUBLISHED = 0, MI_DONE, MI_UIDX, MI_NFINITE, MI_M, MI_NWIN, MI_ERR, MI_KMAX };

template <int V> struct Pow2Floor { static constexpr int value = (V >= 2) ? 2 * Pow2Floor<V / 2>::value : 1; };
template <> struct Pow2Floor<1> { static constexpr int value = 1; };
template <> struct Pow2Floor<0> { static constexpr int value = 1; };

template <int H, int D>
struct Tiles {
  static constexpr int KT_HH = Pow2Floor<kStageBytes / (12 * H)>::value;  // k-rows of W_hh^T per stage
  static constexpr int KT_1 = Pow2Floor<kStageBytes / (4 * H)>::value > H ? H : Pow2Floor<kStageBytes / (4 * H)>::value;
  static constexpr int KT_2 = Pow2Floor<kStageBytes / (4 * D)>::value > H ? H : Pow2Floor<kStageBytes / (4 * D)>::value;
  static constexpr int N_HH = H / KT_HH, N_1 = H / KT_1, N_2 = H / KT_2;
  static constexpr int TILES_PER_PASS = N_HH + N_1 + N_2;
  static_assert(H % KT_HH == 0 && H % KT_1 == 0 && H % KT_2 == 0, "tile split");
};

// ------------------------------------------------------------------ producer (one thread)
template <int H, int D>
__device__ void producer_loop(const BeamParams& p, float* ring, uint64_t* full, uint64_t* empty,
                              volatile int* misc) {
  using TL = Tiles<H, D>;
  unsigned it = 0;
  int pass = 0;
  for (;;) {
    while (*(volatile int*)&misc[MI_PUBLISHED] <= pass) {
      if (*(volatile int*)&misc[MI_DONE]) return;
      __nanosleep(64);
    }
    __threadfence_block();
    for (int seg = 0; seg < 3; ++seg) {
      const float* src = seg == 0 ? p.whh_t : (seg == 1 ? p.w1_t : p.w2_t);
      const int ntiles = seg == 0 ? TL::N_HH : (seg == 1 ? TL::N_1 : TL::N_2);
      const unsigned bytes = seg == 0 ? TL::KT_HH * 3 * H * 4 : (seg == 1 ? TL::KT_1 * H * 4 : TL::KT_2 * D * 4);
      for (int t = 0; t < ntiles; ++t, ++it) {
        const unsigned s = it % kStages, ph = (it / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&full[s], bytes);
        tma_bulk_g2s(reinterpret_cast<char*>(ring) + (size_t)s * kStageBytes,
                     reinterpret_cast<const char*>(src) + (size_t)t * bytes, bytes, &full[s]);
      }
    }
    ++pass;
  }
}

// Consume one full weight pass without computing (used when a step has no winner, so that the
// producer, which was already told about the pass, never blocks on a full ring).
template <int H, int D>
__device__ __forceinline__ void drain_pass(uint64_t* full, uint64_t* empty, unsigned& it, int lane) {
  for (int t = 0; t < Tiles<H, D>::TILES_PER_PASS; ++t, ++it) {
    const unsigned s = it % kStages, ph = (it / kStages) & 1;
    mbar_wait(&full[s], ph);
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
}

// ------------------------------------------------------------------ consumer passes
// Thread j owns hidden unit j.  acc layout: [gate][column].
template <int H, int D, int MP, int NC>
__device__ __forceinline__ void gru_pass(const float* __restrict__ ring, uint64_t* full, uint64_t* empty,
                                         unsigned& it, const float* __restrict__ XA, float* __restrict__ XB,
                                         const float* __restrict__ gi, float bhr, float bhz, float bhn,
                                         int M, const int* __restrict__ colnew, float* __restrict__ pool_hidden,
                                         int j, int lane) {
  using TL = Tiles<H, D>;
  float ar[4 * NC], az[4 * NC], an[4 * NC];
#pragma unroll
  for (int i = 0; i < 4 * NC; ++i) ar[i] = az[i] = an[i] = 0.f;
  for (int tile = 0; tile < TL::N_HH; ++tile, ++it) {
    const unsigned s = it % kStages, ph = (it / kStages) & 1;
    mbar_wait(&full[s], ph);
    const float* wt = ring + (size_t)s * (kStageBytes / 4);
#pragma unroll
    for (int kk = 0; kk < TL::KT_HH; ++kk) {
      const float wr = wt[kk * 3 * H + j];
      const float wz = wt[kk * 3 * H + H + j];
      const float wn = wt[kk * 3 * H + 2 * H + j];
      const float4* xp = reinterpret_cast<const float4*>(XA + (size_t)(tile * TL::KT_HH + kk) * MP);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float4 x = xp[c];
        ar[4 * c + 0] = fmaf(wr, x.x, ar[4 * c + 0]); ar[4 * c + 1] = fmaf(wr, x.y, ar[4 * c + 1]);
        ar[4 * c + 2] = fmaf(wr, x.z, ar[4 * c + 2]); ar[4 * c + 3] = fmaf(wr, x.w, ar[4 * c + 3]);
        az[4 * c + 0] = fmaf(wz, x.x, az[4 * c + 0]); az[4 * c + 1] = fmaf(wz, x.y, az[4 * c + 1]);
        az[4 * c + 2] = fmaf(wz, x.z, az[4 * c + 2]); az[4 * c + 3] = fmaf(wz, x.w, az[4 * c + 3]);
        an[4 * c + 0] = fmaf(wn, x.x, an[4 * c + 0]); an[4 * c + 1] = fmaf(wn, x.y, an[4 * c + 1]);
        an[4 * c + 2] = fmaf(wn, x.z, an[4 * c + 2]); an[4 * c + 3] = fmaf(wn, x.w, an[4 * c + 3]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  // GRU cell, PyTorch gate order r,z,n (uisrnn.py:39-47):  h' = (h - n) * z + n
  const float gir = gi[j], giz = gi[H + j], gin = gi[2 * H + j];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const float4 hold = reinterpret_cast<const float4*>(XA + (size_t)j * MP)[c];
    const float ho[4] = {hold.x, hold.y, hold.z, hold.w};
    float hn[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = 4 * c + q;
      const float r = sigmoid_f32(__fadd_rn(gir, __fadd_rn(ar[m], bhr)));
      const float z = sigmoid_f32(__fadd_rn(giz, __fadd_rn(az[m], bhz)));
      const float n = tanhf(__fadd_rn(gin, __fmul_rn(r, __fadd_rn(an[m], bhn))));
      hn[q] = __fadd_rn(__fmul_rn(__fsub_rn(ho[q], n), z), n);
      if (m < M) pool_hidden[(size_t)colnew[m] * H + j] = hn[q];
    }
    reinterpret_cast<float4*>(XB + (size_t)j * MP)[c] = make_float4(hn[0], hn[1], hn[2], hn[3]);
  }
}

template <int H, int D, int MP, int NC>
__device__ __forceinline__ void mlp1_pass(const float* __restrict__ ring, uint64_t* full, uint64_t* empty,
                                          unsigned& it, const float* __restrict__ XB, float* __restrict__ XA,
                                          float b1j, int j, int lane) {
  using TL = Tiles<H, D>;
  float a[4 * NC];
#pragma unroll
  for (int i = 0; i < 4 * NC; ++i) a[i] = 0.f;
  for (int tile = 0; tile < TL::N_1; ++tile, ++it) {
    const unsigned s = it % kStages, ph = (it / kStages) & 1;
    mbar_wait(&full[s], ph);
    const float* wt = ring + (size_t)s * (kStageBytes / 4);
#pragma unroll 8
    for (int kk = 0; kk < TL::KT_1; ++kk) {
      const float w = wt[kk * H + j];
      const float4* xp = reinterpret_cast<const float4*>(XB + (size_t)(tile * TL::KT_1 + kk) * MP);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float4 x = xp[c];
        a[4 * c + 0] = fmaf(w, x.x, a[4 * c + 0]); a[4 * c + 1] = fmaf(w, x.y, a[4 * c + 1]);
        a[4 * c + 2] = fmaf(w, x.z, a[4 * c + 2]); a[4 * c + 3] = fmaf(w, x.w, a[4 * c + 3]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    float4 v;
    v.x = fmaxf(__fadd_rn(a[4 * c + 0], b1j), 0.f); v.y = fmaxf(__fadd_rn(a[4 * c + 1], b1j), 0.f);
    v.z = fmaxf(__fadd_rn(a[4 * c + 2], b1j), 0.f); v.w = fmaxf(__fadd_rn(a[4 * c + 3], b1j), 0.f);
    reinterpret_cast<float4*>(XA + (size_t)j * MP)[c] = v;
  }
}

// W2 pass: D output rows, NT = H threads => G2 = H / D thread groups split each tile's k-rows.
// Partial sums of groups 1.. go through `scratch` (aliases XB, dead after the W1 pass).
template <int H, int D, int MP, int NC>
__device__ __forceinline__ void mlp2_pass(const float* __restrict__ ring, uint64_t* full, uint64_t* empty,
                                          unsigned& it, const float* __restrict__ XA, float* __restrict__ scratch,
                                          float b2d, int M, const int* __restrict__ colsrc,
                                          const int* __restrict__ colnew, const int* __restrict__ colvis,
                                          float* __restrict__ pool_mean, int tid, int lane) {
  using TL = Tiles<H, D>;
  constexpr int G2 = H / D;
  constexpr int KPG = TL::KT_2 / G2;
  static_assert(H % D == 0 && TL::KT_2 % G2 == 0, "W2 split");
  const int d = tid % D, g = tid / D;
  float a[4 * NC];
#pragma unroll
  for (int i = 0; i < 4 * NC; ++i) a[i] = 0.f;
  // old means of the source slots (consumed in the epilogue; issued early to hide L2 latency)
  float mu_old[4 * NC];
  if (g == 0) {
#pragma unroll
    for (int m = 0; m < 4 * NC; ++m) mu_old[m] = (m < M) ? pool_mean[(size_t)colsrc[m] * D + d] : 0.f;
  }
  for (int tile = 0; tile < TL::N_2; ++tile, ++it) {
    const unsigned s = it % kStages, ph = (it / kStages) & 1;
    mbar_wait(&full[s], ph);
    const float* wt = ring + (size_t)s * (kStageBytes / 4);
#pragma unroll 8
    for (int kq = 0; kq < KPG; ++kq) {
      const int kk = g * KPG + kq;
      const float w = wt[kk * D + d];
      const float4* xp = reinterpret_cast<const float4*>(XA + (size_t)(tile * TL::KT_2 + kk) * MP);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float4 x = xp[c];
        a[4 * c + 0] = fmaf(w, x.x, a[4 * c + 0]); a[4 * c + 1] = fmaf(w, x.y, a[4 * c + 1]);
        a[4 * c + 2] = fmaf(w, x.z, a[4 * c + 2]); a[4 * c + 3] = fmaf(w, x.w, a[4 * c + 3]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  if (G2 > 1) {
    if (g > 0) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        reinterpret_cast<float4*>(scratch + ((size_t)(g - 1) * D + d) * MP)[c] =
            make_float4(a[4 * c + 0], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]);
    }
    named_bar_sync(1, H);
  }
  if (g == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      float v[4] = {a[4 * c + 0], a[4 * c + 1], a[4 * c + 2], a[4 * c + 3]};
      for (int gg = 1; gg < G2; ++gg) {
        const float4 o = reinterpret_cast<const float4*>(scratch + ((size_t)(gg - 1) * D + d) * MP)[c];
        v[0] = __fadd_rn(v[0], o.x); v[1] = __fadd_rn(v[1], o.y);
        v[2] = __fadd_rn(v[2], o.z); v[3] = __fadd_rn(v[3], o.w);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = 4 * c + q;
        if (m < M) {
          const float mval = __fadd_rn(v[q], b2d);
          const int n = colvis[m];  // visits BEFORE this one (uisrnn.py:425-429)
          // mean_set[c] = (mean_set[c] * (n - 1) + mean) / n   -- fp32, true division
          const float mu = (n == 0) ? mval
                                    : __fdiv_rn(__fadd_rn(__fmul_rn(mu_old[m], (float)(n - 1)), mval), (float)n);
          pool_mean[(size_t)colnew[m] * D + d] = mu;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ the kernel
template <int H, int D, int MP>
__global__ void __launch_bounds__(H + 32, 1) uis_beam_kernel(const BeamParams p) {
  constexpr int NT = H;        // consumer threads
  constexpr int NW = NT / 32;  // consumer warps
  static_assert(D % 4 == 0 && H % 32 == 0 && (3 * H / 4) + (D / 4) <= NT, "shape");
  static_assert(MP % 4 == 0 && MP <= 12, "MP");
  extern __shared__ __align__(128) unsigned char smem[];
  const SmemLayout L = make_layout<H, D, MP>(p.B, p.Kcap);
  float* ring = reinterpret_cast<float*>(smem + L.ring);
  float* XA = reinterpret_cast<float*>(smem + L.xa);
  float* XB = reinterpret_cast<float*>(smem + L.xb);
  float* xt = reinterpret_cast<float*>(smem + L.xt);
  float* gis = reinterpret_cast<float*>(smem + L.gi);
  float* wv = reinterpret_cast<float*>(smem + L.wv);
  TabEntry* tabs = reinterpret_cast<TabEntry*>(smem + L.tabs);
  int* meta = reinterpret_cast<int*>(smem + L.meta);
  int* candoff = reinterpret_cast<int*>(smem + L.candoff);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem + L.keys);
  float* svals = reinterpret_cast<float*>(smem + L.svals);
  int* wins = reinterpret_cast<int*>(smem + L.wins);
  int* cols = reinterpret_cast<int*>(smem + L.cols);
  unsigned* used = reinterpret_cast<unsigned*>(smem + L.used);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + kStages;
  volatile int* misc = reinterpret_cast<volatile int*>(smem + L.misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int B = p.B, Kcap = p.Kcap;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NW);
    }
    for (int i = 0; i < 16; ++i) misc[i] = 0;
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == NW) {  // ---------------- producer warp
    if (lane == 0) producer_loop<H, D>(p, ring, full, empty, misc);
    return;
  }

  // ---------------- consumer threads (tid < NT); they synchronise on named barrier 1
  float* pool_mean = p.pool_mean + (size_t)blockIdx.x * p.P * D;
  float* pool_hidden = p.pool_hidden + (size_t)blockIdx.x * p.P * H;
  unsigned* bp = p.bp + (size_t)blockIdx.x * p.maxN * B;
  const int j = tid;
  const float bhr = p.bhh[j], bhz = p.bhh[H + j], bhn = p.bhh[2 * H + j];
  const float b1j = p.b1[j];
  const float b2d = p.b2[tid % D];
  if (tid < D) {
    wv[tid] = p.wvec[tid];
    pool_mean[(size_t)kInitSlot * D + tid] = p.mean0[tid];
  }
  pool_hidden[(size_t)kInitSlot * H + j] = p.hidden0[j];
  const unsigned PW = (unsigned)(p.P + 31) / 32;

  unsigned it = 0;  // weight-ring tile counter (identical in every consumer thread)
  unsigned long long st_cols = 0, st_pass = 0, st_cand = 0, st_steps = 0;
  int st_maxk = 0;

  for (;;) {
    if (tid == 0) misc[MI_UIDX] = atomicAdd(p.queue, 1);
    named_bar_sync(1, NT);
    const int uidx = misc[MI_UIDX];
    if (uidx >= p.U) break;
    const int u = p.order[uidx];
    const long long row0 = p.row_off[u];
    const int N = (int)(p.row_off[u + 1] - row0);
    const int TN = p.T * N;
    const bool traced = (u == p.trace_utt);
    long long dbg_rows = 0;

    // generation 0 = one empty hypothesis (uisrnn.py:528)
    int gen = 0;
    if (tid == 0) {
      int* K = meta;  // meta layout: [gen][field][B], fields K,last,tot,nl
      K[0] = 0; K[B + 0] = -1; K[2 * B + 0] = 0; reinterpret_cast<float*>(K)[3 * B + 0] = 0.f;
      misc[MI_ERR] = 0;
      if (traced && p.dbg_off) p.dbg_off[0] = 0;
    }
    int nb = 1;
    // prefetch the first frame (x row and its input projection) into buffer 0
    if (N > 0) {
      if (tid < 3 * H / 4) cp_async16(gis + tid * 4, p.gi + (size_t)row0 * 3 * H + tid * 4);
      else if (tid < 3 * H / 4 + D / 4) cp_async16(xt + (tid - 3 * H / 4) * 4, p.x + (size_t)row0 * D + (tid - 3 * H / 4) * 4);
      cp_async_commit();
    }
    named_bar_sync(1, NT);
    bool failed = false;

    for (int t = 0; t < TN; ++t) {
      int* mK = meta + gen * 4 * B;       // current generation
      int* mLast = mK + B;
      int* mTot = mK + 2 * B;
      float* mNl = reinterpret_cast<float*>(mK + 3 * B);
      int* nK = meta + (gen ^ 1) * 4 * B;  // next generation
      int* nLast = nK + B;
      int* nTot = nK + 2 * B;
      float* nNl = reinterpret_cast<float*>(nK + 3 * B);
      const TabEntry* tab = tabs + (size_t)gen * B * Kcap;
      TabEntry* ntab = tabs + (size_t)(gen ^ 1) * B * Kcap;
      const int buf = t & 1;
      const float* xs = xt + buf * D;
      const float* gs = gis + buf * 3 * H;

      // ---- P0: publish this step's weight pass to the producer; land x_t / gi_t
      if (tid == 0) {
        __threadfence_block();
        misc[MI_PUBLISHED] = misc[MI_PUBLISHED] + 1;
        int off = 0, kmax = 0;
        for (int b = 0; b < nb; ++b) { candoff[b] = off; off += mK[b] + 1; kmax = max(kmax, mK[b]); }
        candoff[nb] = off;
        misc[MI_NFINITE] = 0;
        misc[MI_KMAX] = kmax;
      }
      for (unsigned w = tid; w < PW; w += NT) used[w] = (w == 0) ? 1u : 0u;  // slot 0 = INIT, always live
      cp_async_wait_all();
      named_bar_sync(1, NT);
      if (t + 1 < TN) {  // prefetch next frame into the other buffer
        const long long r = row0 + ((t + 1) % N);
        if (tid < 3 * H / 4) cp_async16(gis + (buf ^ 1) * 3 * H + tid * 4, p.gi + (size_t)r * 3 * H + tid * 4);
        else if (tid < 3 * H / 4 + D / 4)
          cp_async16(xt + (buf ^ 1) * D + (tid - 3 * H / 4) * 4, p.x + (size_t)r * D + (tid - 3 * H / 4) * 4);
        cp_async_commit();
      }

      // ---- P1: score every candidate (b, c <= K_b): one warp each (uisrnn.py:409-420, 434-446)
      const int NE = candoff[nb];
      const int kmax = misc[MI_KMAX];
      for (int e = warp; e < NE; e += NW) {
        int b = 0;
        while (candoff[b + 1] <= e) ++b;
        const int c = e - candoff[b];
        const int Kb = mK[b];
        const TabEntry en = (c < Kb) ? tab[(size_t)b * Kcap + c] : TabEntry{kInitSlot, 0, 0, 0};
        const float* mu = pool_mean + (size_t)en.slot * D;
        // weighted_mse_loss for one row (loss_func.py:33-41): sum_d fl(fl(diff^2) * w_d)
        float acc = 0.f, d0sq = 1.f;
        for (int d = lane * 4; d < D; d += 128) {
          const float4 m4 = *reinterpret_cast<const float4*>(mu + d);
          const float4 x4 = *reinterpret_cast<const float4*>(xs + d);
          const float4 w4 = *reinterpret_cast<const float4*>(wv + d);
          const float e0 = __fsub_rn(m4.x, x4.x), e1 = __fsub_rn(m4.y, x4.y);
          const float e2 = __fsub_rn(m4.z, x4.z), e3 = __fsub_rn(m4.w, x4.w);
          const float q0 = __fmul_rn(e0, e0);
          if (d == 0) d0sq = q0;
          acc = __fadd_rn(acc, __fmul_rn(q0, w4.x));
          acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(e1, e1), w4.y));
          acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(e2, e2), w4.z));
          acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(e3, e3), w4.w));
        }
        acc = warp_sum(acc);
        if (lane == 0) {
          if (c < Kb) atomicOr(&used[en.slot >> 5], 1u << (en.slot & 31));
          float mse = acc;
          if (d0sq == 0.f) mse = __fdiv_rn(acc, 0.f);  // zero "non-zero rows" (loss_func.py:36)
          double pen;
          if (c < Kb) {
            pen = (c == mLast[b]) ? p.log_1mp0 : (p.log_p0 + p.logn[en.blocks]) - p.logtot[mTot[b]];
          } else {
            pen = (p.log_p0 + p.log_alpha) - p.logtot[mTot[b]];
          }
          const float loss = __double2float_rn((double)mse - pen);
          const float S = __fadd_rn(mNl[b], loss);
          svals[e] = S;
          const unsigned flat = (unsigned)(b * (kmax + 1) + c);
          keys[e] = ((unsigned long long)float_order_key(S) << 32) | flat;
          if (S < __int_as_float(0x7f800000)) atomicAdd((int*)&misc[MI_NFINITE], 1);
        }
      }
      named_bar_sync(1, NT);

      // ---- P2: rank by counting; the best min(#finite, B) become the new hypotheses (:546-552)
      const int nwin = min((int)misc[MI_NFINITE], B);
      for (int e = tid; e < NE; e += NT) {
        const unsigned long long k = keys[e];
        int rank = 0;
        for (int q = 0; q < NE; ++q) rank += (keys[q] < k) ? 1 : 0;
        if (rank < nwin) {
          int b = 0;
          while (candoff[b + 1] <= e) ++b;
          wins[rank] = b;
          wins[B + rank] = e - candoff[b];
          reinterpret_cast<float*>(wins)[2 * B + rank] = svals[e];
        }
      }
      named_bar_sync(1, NT);

      // ---- P3: warp 0 assigns GRU columns (distinct source slots) and allocates new slots;
      //          the other warps copy the parents' tables
      int* col = cols; int* colsrc = cols + B; int* colnew = cols + 2 * B; int* colvis = cols + 3 * B;
      if (warp == 0) {
        // (beam_size <= 32 is enforced by the host)
        const int r = lane;
        int src = -1;
        if (r < nwin) {
          const int b = wins[r], c = wins[B + r];
          src = (c < mK[b]) ? tab[(size_t)b * Kcap + c].slot : kInitSlot;
        }
        int first = r;
        for (int q = 0; q < nwin; ++q) {
          const int sq = __shfl_sync(0xffffffffu, src, q);
          if (q < first && sq == src) first = q;
        }
        const bool isfirst = (r < nwin) && (first == r);
        const unsigned fm = __ballot_sync(0xffffffffu, isfirst);
        const int mycol = __popc(fm & ((1u << lane) - 1));
        const int M = __popc(fm);
        const int c_of_first = __shfl_sync(0xffffffffu, mycol, first);
        if (r < nwin) col[r] = c_of_first;
        if (isfirst) colsrc[mycol] = src;
        // allocate M free slots from the bitmap (any free slot will do)
        int cnt = 0;
        for (unsigned w = lane; w < PW; w += 32) {
          unsigned fr = ~used[w];
          if (w == PW - 1 && (p.P & 31)) fr &= (1u << (p.P & 31)) - 1;
          cnt += __popc(fr);
        }
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += v;
        }
        int idx = incl - cnt;
        for (unsigned w = lane; w < PW && idx < M; w += 32) {
          unsigned fr = ~used[w];
          if (w == PW - 1 && (p.P & 31)) fr &= (1u << (p.P & 31)) - 1;
          while (fr && idx < M) {
            const int bit = __ffs(fr) - 1;
            fr &= fr - 1;
            colnew[idx++] = (int)(w * 32 + bit);
          }
        }
        if (lane == 0) misc[MI_M] = M;
      } else {
        for (int r = warp - 1; r < nwin; r += NW - 1) {
          const int b = wins[r];
          const int Kb = mK[b];
          for (int c = lane; c < Kb; c += 32) ntab[(size_t)r * Kcap + c] = tab[(size_t)b * Kcap + c];
        }
      }
      named_bar_sync(1, NT);

      // ---- P4: patch the one changed table entry per child; back-pointers; hypothesis meta
      const int M = misc[MI_M];
      if (tid < nwin) {
        const int r = tid;
        const int b = wins[r], c = wins[B + r];
        const int Kb = mK[b];
        const bool isnew = (c == Kb);
        if (isnew && Kb >= Kcap) {
          misc[MI_ERR] = 1;
        } else {
          const TabEntry old = isnew ? TabEntry{kInitSlot, 0, 0, 0} : tab[(size_t)b * Kcap + c];
          const bool moved = isnew || (c != mLast[b]);
          TabEntry ne;
          ne.slot = colnew[col[r]];
          ne.blocks = old.blocks + (moved ? 1 : 0);  // uisrnn.py:431-432; new cluster starts at 1 (:76)
          ne.visits = old.visits + 1;
          ne.pad = 0;
          ntab[(size_t)r * Kcap + c] = ne;
          nK[r] = Kb + (isnew ? 1 : 0);
          nLast[r] = c;
          nTot[r] = mTot[b] + (moved ? 1 : 0);
          nNl[r] = reinterpret_cast<float*>(wins)[2 * B + r];
          if (col[r] >= 0 && colsrc[col[r]] == old.slot) colvis[col[r]] = old.visits;  // same slot => same count
        }
        if (t >= TN - N) bp[(size_t)(t - (TN - N)) * B + r] = ((unsigned)b << 16) | (unsigned)c;
        if (traced && p.dbg_win && dbg_rows + r < p.trace_capacity) {
          p.dbg_win[(dbg_rows + r) * 2 + 0] = b;
          p.dbg_win[(dbg_rows + r) * 2 + 1] = c;
          p.dbg_score[dbg_rows + r] = reinterpret_cast<float*>(wins)[2 * B + r];
        }
      }
      if (traced && tid == 0 && p.dbg_off) p.dbg_off[t + 1] = dbg_rows + nwin;
      dbg_rows += nwin;
      if (tid == 0) {
        st_cand += NE; st_cols += M; st_steps += 1;
        const int npass = max(1, (M + MP - 1) / MP);
        st_pass += npass;
        if (npass > 1) { __threadfence_block(); misc[MI_PUBLISHED] = misc[MI_PUBLISHED] + (npass - 1); }
      }
      named_bar_sync(1, NT);
      if (misc[MI_ERR]) {
        // cluster cap exceeded: the already-published weight pass must still be consumed
        failed = true;
      }

      // ---- P5: GRU + MLP for the M distinct source states, MP columns per weight pass
      if (M == 0) drain_pass<H, D>(full, empty, it, lane);
      for (int m0 = 0; m0 < M; m0 += MP) {
        const int Mp = min(MP, M - m0);
        {  // gather source hidden states, transposed: XA[k][m]
          float hv[MP];
#pragma unroll
          for (int m = 0; m < MP; ++m)
            hv[m] = (m < Mp) ? pool_hidden[(size_t)colsrc[m0 + m] * H + j] : 0.f;
#pragma unroll
          for (int c = 0; c < MP / 4; ++c)
            reinterpret_cast<float4*>(XA + (size_t)j * MP)[c] =
                make_float4(hv[4 * c], hv[4 * c + 1], hv[4 * c + 2], hv[4 * c + 3]);
        }
        named_bar_sync(1, NT);
        const int nc = (Mp + 3) / 4;
        if (nc == 1) gru_pass<H, D, MP, 1>(ring, full, empty, it, XA, XB, gs, bhr, bhz, bhn, Mp, colnew + m0, pool_hidden, j, lane);
        else if (nc == 2) gru_pass<H, D, MP, 2>(ring, full, empty, it, XA, XB, gs, bhr, bhz, bhn, Mp, colnew + m0, pool_hidden, j, lane);
        else gru_pass<H, D, MP, 3>(ring, full, empty, it, XA, XB, gs, bhr, bhz, bhn, Mp, colnew + m0, pool_hidden, j, lane);
        named_bar_sync(1, NT);
        if (nc == 1) mlp1_pass<H, D, MP, 1>(ring, full, empty, it, XB, XA, b1j, j, lane);
        else if (nc == 2) mlp1_pass<H, D, MP, 2>(ring, full, empty, it, XB, XA, b1j, j, lane);
        else mlp1_pass<H, D, MP, 3>(ring, full, empty, it, XB, XA, b1j, j, lane);
        named_bar_sync(1, NT);
        if (nc == 1) mlp2_pass<H, D, MP, 1>(ring, full, empty, it, XA, XB, b2d, Mp, colsrc + m0, colnew + m0, colvis + m0, pool_mean, tid, lane);
        else if (nc == 2) mlp2_pass<H, D, MP, 2>(ring, full, empty, it, XA, XB, b2d, Mp, colsrc + m0, colnew + m0, colvis + m0, pool_mean, tid, lane);
        else mlp2_pass<H, D, MP, 3>(ring, full, empty, it, XA, XB, b2d, Mp, colsrc + m0, colnew + m0, colvis + m0, pool_mean, tid, lane);
        named_bar_sync(1, NT);
      }
      if (tid == 0) { for (int r = 0; r < nwin; ++r) st_maxk = max(st_maxk, nK[r]); }
      nb = nwin;
      gen ^= 1;
      if (failed || nb == 0) break;
    }  // steps

    // drain a pending prefetch before the buffers are reused by the next utterance
    cp_async_wait_all();
    named_bar_sync(1, NT);

    // ---- utterance epilogue: back-track the best hypothesis (uisrnn.py:561)
    if (tid == 0) {
      if (failed || nb == 0) {
        p.status[u] = failed ? -4 : -1;
        for (int i = 0; i < N; ++i) p.labels[row0 + i] = -1;
      } else {
        p.status[u] = 0;
        int r = 0;
        for (int i = N - 1; i >= 0; --i) {
          const unsigned e = bp[(size_t)i * B + r];
          p.labels[row0 + i] = (int)(e & 0xffffu);
          r = (int)(e >> 16);
        }
      }
    }
    if (p.dbg_final_scores) {
      const float* fNl = reinterpret_cast<const float*>(meta + gen * 4 * B + 3 * B);
      if (tid < B) p.dbg_final_scores[(size_t)u * B + tid] = (tid < nb) ? fNl[tid] : __int_as_float(0x7f800000);
      if (tid == 0 && p.dbg_final_k) p.dbg_final_k[u] = (nb > 0) ? meta[gen * 4 * B] : 0;
    }
    if (traced && nb > 0 && !failed && p.dbg_best_mean) {
      const TabEntry* ftab = tabs + (size_t)gen * B * Kcap;  // best hypothesis = rank 0
      const int K0 = meta[gen * 4 * B];
      for (int c = 0; c < K0; ++c) {
        const TabEntry en = ftab[c];
        if (tid < D) p.dbg_best_mean[(size_t)c * D + tid] = pool_mean[(size_t)en.slot * D + tid];
        p.dbg_best_hidden[(size_t)c * H + j] = pool_hidden[(size_t)en.slot * H + j];
        if (tid == 0) p.dbg_best_blocks[c] = en.blocks;
      }
    }
    named_bar_sync(1, NT);
  }  // utterances

  if (tid == 0) {
    __threadfence_block();
    misc[MI_DONE] = 1;
    atomicAdd(&p.stats[0], st_cols);
    atomicAdd(&p.stats[1], st_pass);
    atomicAdd(&p.stats[2], st_cand);
    atomicAdd(&p.stats[3], st_steps);
    atomicMax(&p.stats[4], (unsigned long long)st_maxk);
  }
}

}  // namespace uis
