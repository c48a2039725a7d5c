// C ABI of libuisrnn_b200.so (see include/uisrnn_b200.h).  Host-side orchestration only:
// weight re-layout at model creation, workspace management, utterance scheduling (longest
// first), kernel launches.  No PyTorch types, no CPU compute fallback: if the kernels are not
// instantiated for a shape the call fails with UIS_ERR_UNSUPPORTED.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/uisrnn_b200.h"
#include "uis_beam.cuh"
#include "uis_beam_tree.cuh"
#include "uis_launch.cuh"
#include "uis_prepass.cuh"

namespace {
thread_local std::string g_err;
}

namespace uis {
// shared with uis_train.cu
int api_fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
}  // namespace uis

namespace {

template <class... Args>
int fail(int code, const char* fmt, Args... args) {
  return uis::api_fail(code, fmt, args...);
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(UIS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                  __LINE__);                                                                  \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      want = bytes;
      e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess) return fail(UIS_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    cap = want;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T> T* as() const { return static_cast<T*>(p); }
};


}  // namespace

namespace {
// Host threads that copy the pieces of one staging chunk from the caller's PAGEABLE arrays into pinned memory, side by
// side (a cudaMemcpy from pageable memory is staged by the driver in one thread at ~11 GB/s; several threads reach
// the PCIe rate).  run() hands the same piece list to every worker; worker i copies bytes [total * i / n, total * (i + 1) / n).
struct CopyPiece { const char* src; size_t dst_off, bytes; };
class CopyPool {
 public:
  explicit CopyPool(int n) : n_(n) {
    for (int i = 0; i < n; ++i) th_.emplace_back([this, i] { loop(i); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> lk(mu_); stop_ = true; ++gen_; }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  void run(const std::vector<CopyPiece>* pieces, char* dst, size_t total) {
    { std::lock_guard<std::mutex> lk(mu_); pieces_ = pieces; dst_ = dst; total_ = total; pending_ = n_; ++gen_; }
    cv_.notify_all();
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
  }
 private:
  void loop(int i) {
    unsigned long long seen = 0;
    for (;;) {
      const std::vector<CopyPiece>* pieces; char* dst; size_t total;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        pieces = pieces_; dst = dst_; total = total_;
      }
      const size_t b0 = total * (size_t)i / (size_t)n_, b1 = total * (size_t)(i + 1) / (size_t)n_;
      for (const CopyPiece& p : *pieces) {
        const size_t lo = std::max(b0, p.dst_off), hi = std::min(b1, p.dst_off + p.bytes);
        if (lo < hi) std::memcpy(dst + lo, p.src + (lo - p.dst_off), hi - lo);
      }
      std::lock_guard<std::mutex> lk(mu_);
      if (--pending_ == 0) done_.notify_one();
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  const std::vector<CopyPiece>* pieces_ = nullptr;
  char* dst_ = nullptr;
  size_t total_ = 0;
  int pending_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
};
}  // namespace

struct uis_model {
  int device = 0, D = 0, H = 0, depth = 1, num_sms = 0;  // D, H: the kernel shape the model runs in
  int D_user = 0, H_user = 0;  // the caller's shape (<= D, H): smaller models are zero-padded into the next kernel shape
  double p0 = 0, alpha = 0;
  // weights, k-major
  DevBuf wih_t, whh_t, w1_t, w2_t, bih, bhh, b1, b2, wvec, mean0, hidden0;
  DevBuf wih_up_t;  // [depth-1][H][3H]; whh_t is [depth][H][3H]; bih / bhh are [depth][3H]
  // tensor-core pass (uis_beam_tc.cuh): fp16 hi/lo planes of [W_hh; W1; W2] behind a tensor map, scales
  DevBuf tc_planes, tc_scratch, stat_bar, stat_scratch;
  alignas(64) CUtensorMap tc_map;
  bool tc_ready = false;
  float tc_sh = 0, tc_sa = 0, tc_inv_hh = 0, tc_inv_1 = 0, tc_inv_2 = 0;
  // log tables
  DevBuf logn, logtot;
  int log_cap = 0;
  // workspace
  DevBuf x64, x32, gi, row_off, order, pool_mean, pool_hidden, pool_mse, bp, queue_stats, labels, status;
  DevBuf dbg_win, dbg_score, dbg_off, dbg_final_scores, dbg_final_k, dbg_best_mean, dbg_best_hidden,
      dbg_best_blocks;
  // last call
  uis_stats stats{};
  int last_U = 0;
  cudaStream_t last_stream = nullptr;
  bool stats_pending = false;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};  // before prepass, after prepass, after beam kernel
  // host-buffer path (uis_predict): the float64 rows travel in chunks through a small ring of staging slots on a
  // copy stream of their own, so the H2D copy of chunk c + 1 runs under the cast + input projection of chunk c and
  // the fp64 staging is O(chunk), not O(input); labels come back in ONE copy into pinned memory
  cudaStream_t copy_stream = nullptr;
  static constexpr int kSlots = 3;
  cudaEvent_t ev_copied[kSlots] = {nullptr, nullptr, nullptr}, ev_free[kSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
  cudaEvent_t ev_pipe = nullptr;  // compute stream, before the first cast
  int32_t* labels_pin = nullptr;
  size_t labels_pin_cap = 0;
  // pageable inputs: pinned staging ring (kSlots chunks) filled by host threads, one DMA per chunk
  char* pin_stage = nullptr;
  size_t pin_stage_cap = 0;
  cudaEvent_t ev_dma[kSlots] = {nullptr, nullptr, nullptr};
  CopyPool* copy_pool = nullptr;
};

namespace {

unsigned smem_bytes(int H, int D, int B, int Kcap, int G) {
  if (H == 512 && D == 256) return uis::make_layout<512, 256>(B, Kcap, G).total;
  if (H == 256 && D == 128) return uis::make_layout<256, 128>(B, Kcap, G).total;
  if (H == 128 && D == 64) return uis::make_layout<128, 64>(B, Kcap, G).total;
  if (H == 1024 && D == 512) return uis::make_layout<1024, 512, uis::beam_cp<1024>()>(B, Kcap, G).total;
  return 0xffffffffu;
}

unsigned tree_smem_bytes(int H, int D, int B, int Kcap, int L, int NI, int NLF, int P) {
  if (H == 512 && D == 256) return uis::make_tree_layout<512, 256>(B, Kcap, L, NI, NLF, P).total;
  if (H == 256 && D == 128) return uis::make_tree_layout<256, 128>(B, Kcap, L, NI, NLF, P).total;
  if (H == 128 && D == 64) return uis::make_tree_layout<128, 64>(B, Kcap, L, NI, NLF, P).total;
  return 0xffffffffu;
}

int dispatch_tree(int H, int D, const uis::BeamParams& p, int ctas, cudaStream_t st) {
  const unsigned smem = tree_smem_bytes(H, D, p.B, p.Kcap, p.L, p.node_cap, p.leaf_cap, p.P);
  if (smem > 227u * 1024u)
    return fail(UIS_ERR_UNSUPPORTED, "look_ahead=%d beam_size=%d kcap=%d needs %u B of shared memory (> 227 KB)", p.L,
                p.B, p.Kcap, smem);
  cudaError_t e = cudaSuccess;
  if (!uis::launch_tree_large(H, D, p, ctas, smem, st, &e) && !uis::launch_tree_small(H, D, p, ctas, smem, st, &e))
    return fail(UIS_ERR_UNSUPPORTED, "no sm_100a kernel instantiated for hidden=%d dim=%d", H, D);
  if (e != cudaSuccess) return fail(UIS_ERR_CUDA, "look-ahead kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}

bool shape_supported(int H, int D) {
  return (H == 512 && D == 256) || (H == 256 && D == 128) || (H == 128 && D == 64) || (H == 1024 && D == 512);
}

// *cluster: in = planned cluster size, out = the one that was launched.  When the cluster size was chosen
// automatically and the cluster launch is refused (e.g. a partitioned GPU that cannot co-schedule the CTAs), the
// one-CTA-per-utterance kernel runs on the same grid instead: its extra CTAs find the utterance queue empty.
int dispatch_beam(int H, int D, const uis::BeamParams& p, int ctas, int* cluster, bool cluster_forced, cudaStream_t st) {
  if (*cluster > 1) {
    cudaError_t e = cudaSuccess;
    const bool have = uis::launch_beam_cluster(H, D, p, ctas, *cluster, uis::beam_cluster_smem(H, D, p.B, p.Kcap), st, &e);
    if (have && e == cudaSuccess) return 0;
    if (cluster_forced) {
      if (!have) return fail(UIS_ERR_UNSUPPORTED, "no cluster-mode kernel for hidden=%d dim=%d", H, D);
      return fail(UIS_ERR_CUDA, "cluster beam kernel launch failed: %s", cudaGetErrorString(e));
    }
    (void)cudaGetLastError();  // clear the launch error and fall back
    *cluster = 1;
  }
  const unsigned smem = smem_bytes(H, D, p.B, p.Kcap, p.G);
  if (smem > 227u * 1024u)
    return fail(UIS_ERR_UNSUPPORTED, "beam_size=%d kcap=%d lanes=%d needs %u B of shared memory (> 227 KB); lower kcap",
                p.B, p.Kcap, p.G, smem);
  cudaError_t e = cudaSuccess;
  if (!uis::launch_beam_large(H, D, p, ctas, smem, st, &e) && !uis::launch_beam_small(H, D, p, ctas, smem, st, &e))
    return fail(UIS_ERR_UNSUPPORTED, "no sm_100a kernel instantiated for hidden=%d dim=%d", H, D);
  if (e != cudaSuccess) return fail(UIS_ERR_CUDA, "beam kernel launch failed: %s", cudaGetErrorString(e));
  return 0;
}

int fetch(std::vector<float>& dst, const float* src, size_t n) {
  dst.resize(n);
  CU(cudaMemcpy(dst.data(), src, n * sizeof(float), cudaMemcpyDefault));
  return 0;
}

int upload(DevBuf& b, const void* src, size_t bytes) {
  if (int rc = b.ensure(bytes)) return rc;
  CU(cudaMemcpy(b.p, src, bytes, cudaMemcpyHostToDevice));
  return 0;
}

std::vector<float> transpose(const std::vector<float>& w, int rows, int cols) {
  std::vector<float> t((size_t)rows * cols);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) t[(size_t)c * rows + r] = w[(size_t)r * cols + c];
  return t;
}

// ---- tensor-core pass: one-off weight preparation --------------------------------------------------------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// largest power of two s with bound * s <= 2^14 (fp16 keeps 11 significant bits down to 2^-14; |hi| stays < 65504)
float tc_pow2_scale(double bound) {
  if (!(bound > 0.0) || !std::isfinite(bound)) return 0.f;
  int e = 0;
  std::frexp(16384.0 / bound, &e);   // 16384 / bound = f * 2^e, f in [0.5, 1)
  e -= 1;                            // 2^e <= 16384 / bound
  if (e > 40) e = 40;
  if (e < -40) return 0.f;
  return std::ldexp(1.0f, e);
}

// Splits w * scale into fp16 hi + lo (22 significant bits) -- the A operand planes of the tcgen05 pass.
void tc_split(const std::vector<float>& w, float scale, __half* hi, __half* lo) {
  for (size_t i = 0; i < w.size(); ++i) {
    const float v = w[i] * scale;
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn(v - __half2float(h));
  }
}

// w_hh [3H,H], w1 [H,H], w2 [D,H] are the row-major (= K-major) PyTorch tensors; hidden0 [H] = CoreRNN(0, h0).
// Leaves tc_ready false (the FFMA kernels serve the model) when the shape does not tile or a bound is not finite.
int tc_prepare(uis_model* m, const std::vector<float>& w_hh, const std::vector<float>& w1, const std::vector<float>& b1,
               const std::vector<float>& w2, const std::vector<float>& hidden0) {
  const int H = m->H, D = m->D;
  m->tc_ready = false;
  if (m->depth != 1 || !uis::beam_tc_supported(H, D, 48)) return 0;
  auto maxabs = [](const std::vector<float>& v) { double a = 0; for (float x : v) a = std::max(a, (double)std::fabs(x)); return a; };
  // |h'| <= max(1, |h|) by induction (h' is a convex combination of h and tanh(.)), starting from hidden0
  const double hmax = std::max(1.0, maxabs(hidden0));
  double amax = 0;  // a = relu(W1 h' + b1):  |a_i| <= |b1_i| + hmax * sum_j |W1_ij|
  for (int i = 0; i < H; ++i) {
    double srow = 0;
    for (int j = 0; j < H; ++j) srow += std::fabs(w1[(size_t)i * H + j]);
    amax = std::max(amax, std::fabs((double)b1[i]) + hmax * srow);
  }
  const float s_hh = tc_pow2_scale(maxabs(w_hh)), s_1 = tc_pow2_scale(maxabs(w1)), s_2 = tc_pow2_scale(maxabs(w2));
  const float s_h = tc_pow2_scale(hmax), s_a = tc_pow2_scale(std::max(amax, 1e-30));
  if (s_hh == 0.f || s_1 == 0.f || s_2 == 0.f || s_h == 0.f || s_a == 0.f) return 0;
  const size_t rows = (size_t)3 * H + H + D, n = rows * H;
  std::vector<__half> planes(2 * n);  // [plane 0 = lo | plane 1 = hi][rows][H]
  tc_split(w_hh, s_hh, planes.data() + n, planes.data());
  {
    std::vector<__half> hi((size_t)H * H), lo((size_t)H * H);
    tc_split(w1, s_1, hi.data(), lo.data());
    std::copy(lo.begin(), lo.end(), planes.begin() + (size_t)3 * H * H);
    std::copy(hi.begin(), hi.end(), planes.begin() + n + (size_t)3 * H * H);
  }
  {
    std::vector<__half> hi((size_t)D * H), lo((size_t)D * H);
    tc_split(w2, s_2, hi.data(), lo.data());
    std::copy(lo.begin(), lo.end(), planes.begin() + (size_t)4 * H * H);
    std::copy(hi.begin(), hi.end(), planes.begin() + n + (size_t)4 * H * H);
  }
  if (int r = upload(m->tc_planes, planes.data(), planes.size() * sizeof(__half))) return r;
  TensorMapEncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&encode), cudaEnableDefault, &qres) !=
          cudaSuccess || !encode || qres != cudaDriverEntryPointSuccess) {
    (void)cudaGetLastError();
    return 0;  // driver without tensor maps: the FFMA kernels serve the model
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)H, (cuuint64_t)(2 * rows)};
  const cuuint64_t gstr[1] = {(cuuint64_t)H * 2};
  const cuuint32_t box[2] = {64, 128};
  const cuuint32_t estr[2] = {1, 1};
  if (encode(&m->tc_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, m->tc_planes.p, gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return 0;
  m->tc_sh = s_h; m->tc_sa = s_a;
  m->tc_inv_hh = 1.0f / (s_hh * s_h); m->tc_inv_1 = 1.0f / (s_1 * s_h); m->tc_inv_2 = 1.0f / (s_2 * s_a);
  m->tc_ready = true;
  return 0;
}

int ensure_log_tables(uis_model* m, int max_tn) {
  const int need = max_tn + 2;
  if (need <= m->log_cap) return 0;
  const int cap = std::max(need, 4096);
  std::vector<double> ln(cap), lt(cap);
  ln[0] = -INFINITY;
  for (int i = 1; i < cap; ++i) ln[i] = std::log((double)i);             // np.log(block_counts[c])
  for (int i = 0; i < cap; ++i) lt[i] = std::log((double)i + m->alpha);  // np.log(sum(block_counts) + alpha)
  if (int rc = upload(m->logn, ln.data(), cap * sizeof(double))) return rc;
  if (int rc = upload(m->logtot, lt.data(), cap * sizeof(double))) return rc;
  m->log_cap = cap;
  return 0;
}

struct Plan {
  int B, L, T, Kcap, ctas, P, maxN, G;
  int cluster = 1;  // CTAs per utterance (thread-block cluster size); 1 = one CTA per lane group
  int stat = 0;     // > 0: stationary-weights mode with this many groups of 32 CTAs (uis_beam_stat.cuh)
  bool stat_forced = false;
  int tcn = 0;      // > 0: tensor-core beam kernel with this many columns per pass (uis_beam_tc.cuh)
  bool cluster_forced = false;
  int node_cap = 0, leaf_cap = 0, maxTN = 0, maxSteps = 0;  // look_ahead >= 2 only
  long long rows;
};

int make_plan(uis_model* m, const int64_t* off, int U, const uis_predict_opts* o, Plan* pl, bool has_taps = false) {
  if (!m || !o || (U > 0 && !off)) return fail(UIS_ERR_INVALID, "null argument");
  if (U < 0) return fail(UIS_ERR_INVALID, "U < 0");
  if (o->beam_size < 1 || o->look_ahead < 1 || o->test_iteration < 1)
    return fail(UIS_ERR_INVALID, "beam_size, look_ahead and test_iteration must be >= 1");
  if (o->look_ahead > 8) return fail(UIS_ERR_UNSUPPORTED, "look_ahead=%d > 8 not supported", o->look_ahead);
  if (o->beam_size > uis::kMaxBeam) return fail(UIS_ERR_UNSUPPORTED, "beam_size=%d > %d not supported", o->beam_size, uis::kMaxBeam);
  if (o->beam_size > 32 && o->look_ahead > 1)
    return fail(UIS_ERR_UNSUPPORTED, "beam_size=%d > 32 is supported with look_ahead 1 only (look_ahead=%d)", o->beam_size, o->look_ahead);
  if (o->engine < 0 || o->engine > 2) return fail(UIS_ERR_INVALID, "engine must be 0 (auto), 1 (FFMA) or 2 (tensor cores)");
  pl->B = o->beam_size;
  pl->L = o->look_ahead;
  pl->T = o->test_iteration;
  const bool tree = pl->L > 1;
  pl->Kcap = o->kcap > 0 ? o->kcap : (tree ? 16 : 32);
  if (o->kcap <= 0 && pl->B > 32)  // wide beams: the per-hypothesis tables (B * kcap entries) must fit shared memory
    while (pl->Kcap > 4 && smem_bytes(m->H, m->D, pl->B, pl->Kcap, 1) > 227u * 1024u) pl->Kcap /= 2;
  pl->tcn = 0;
  if (pl->Kcap > (pl->B > 32 ? 511 : 2047) || pl->B * pl->Kcap + pl->B + 1 > 65535) return fail(UIS_ERR_INVALID, "kcap too large");
  if (tree && pl->Kcap > 255) return fail(UIS_ERR_UNSUPPORTED, "look_ahead >= 2 supports kcap <= 255");
  pl->P = pl->B * pl->Kcap + pl->B + 1;
  pl->rows = U > 0 ? off[U] : 0;
  int maxN = 0;
  for (int u = 0; u < U; ++u) {
    const long long n = off[u + 1] - off[u];
    if (n < 0) return fail(UIS_ERR_INVALID, "frame_offsets not monotone");
    if (n * pl->T > (1ll << 30)) return fail(UIS_ERR_INVALID, "utterance too long");
    maxN = std::max<long long>(maxN, n);
  }
  pl->maxN = std::max(maxN, 1);
  int ctas = o->n_ctas > 0 ? o->n_ctas : m->num_sms;
  // lanes (utterances advanced together by one CTA, sharing each weight pass): 2 when there is
  // enough work to keep every CTA's lanes busy, else 1 (latency mode); opts->lanes overrides.
  int G = o->lanes > 0 ? std::min(o->lanes, 4) : ((long long)U >= 2ll * ctas ? 2 : 1);
  if (tree) {
    // look-ahead tree kernel: one utterance per CTA; size the on-chip node / leaf arrays to what
    // shared memory allows (internal nodes : leaves ~ 1 : 8, the typical fan-out K+2)
    G = 1;
    long long tn = 0;
    for (int u = 0; u < U; ++u) tn = std::max<long long>(tn, (off[u + 1] - off[u]) * pl->T);
    pl->maxTN = (int)std::max<long long>(tn, 1);
    pl->maxSteps = (pl->maxTN + pl->L - 1) / pl->L;
    int ni = 64;
    auto fits = [&](int n) {
      const int P = pl->B * pl->Kcap + n + pl->B + 1;
      return tree_smem_bytes(m->H, m->D, pl->B, pl->Kcap, pl->L, n, 8 * n, P) <= 227u * 1024u;
    };
    if (!fits(ni)) return fail(UIS_ERR_UNSUPPORTED, "look_ahead=%d beam_size=%d kcap=%d does not fit in shared memory", pl->L, pl->B, pl->Kcap);
    while (ni < 4096 && fits(ni + 32)) ni += 32;
    pl->node_cap = ni;
    pl->leaf_cap = 8 * ni;
    pl->P = pl->B * pl->Kcap + ni + pl->B + 1;
  } else {
    // Tensor-core engine (look_ahead 1, depth 1, 128-row-tileable shapes): the cost of a weight pass does not depend on
    // the number of columns, so a CTA advances up to N / 8 utterances together (a lane needs ~6 columns per step,
    // at most beam_size + 1).  Chosen automatically when some CTA gets more than one utterance; below that the
    // one-lane FFMA kernel or the cluster (latency) mode is faster.  Its device tables default to 16 clusters per
    // hypothesis (UIS_ERR_OVERFLOW asks the caller for more, as always).
    if (o->engine != 1 && m->tc_ready && o->cluster <= 0) {
      int N = 48;
      if (const char* env = std::getenv("UISRNN_B200_TC_N")) N = std::atoi(env);
      if (uis::beam_tc_supported(m->H, m->D, N)) {
        const int kc = o->kcap > 0 ? o->kcap : 16;
        int Gt = o->lanes > 0 ? std::min(o->lanes, (int)uis::kMaxLanes)
                              : (int)std::min<long long>(N / 8, ((long long)U + ctas - 1) / std::max(ctas, 1));
        Gt = std::max(Gt, 1);
        while (Gt > 1 && (uis::beam_tc_smem(m->H, m->D, N, pl->B, kc, Gt) > 227u * 1024u || Gt * pl->B > 256)) --Gt;
        const bool fits = uis::beam_tc_smem(m->H, m->D, N, pl->B, kc, Gt) <= 227u * 1024u &&
                          pl->B * kc + pl->B + 1 <= 65535;
        if (fits && (o->engine == 2 || (long long)U > ctas)) {
          pl->tcn = N;
          pl->Kcap = kc;
          pl->P = pl->B * kc + pl->B + 1;
          G = Gt;
        } else if (o->engine == 2) {
          return fail(UIS_ERR_UNSUPPORTED, "tensor-core engine: beam_size=%d kcap=%d does not fit in shared memory", pl->B, kc);
        }
      } else if (o->engine == 2) {
        return fail(UIS_ERR_UNSUPPORTED, "tensor-core engine: no kernel for hidden=%d dim=%d columns=%d", m->H, m->D, N);
      }
    } else if (o->engine == 2) {
      return fail(UIS_ERR_UNSUPPORTED, "tensor-core engine needs look_ahead 1, depth 1, hidden/dim multiples of 128 and no cluster mode");
    }
    if (!pl->tcn)
      while (G > 1 && smem_bytes(m->H, m->D, pl->B, pl->Kcap, G) > 227u * 1024u) --G;
    while (G > 1 && G * pl->B > 256) --G;  // phase P4 gives one consumer thread to every (lane, winner)
  }
  if (tree && o->engine == 2) return fail(UIS_ERR_UNSUPPORTED, "tensor-core engine: look_ahead must be 1");
  pl->G = G;
  pl->ctas = std::max(1, std::min(ctas, std::max((U + G - 1) / G, 1)));
  // Cluster (latency) mode: with fewer utterances than SMs, a thread-block cluster of 2/4/8 CTAs works on
  // each utterance (k-split of every weight matrix, uis_beam.cuh).  opts->cluster: 0 = auto (largest of 4, 2
  // that still gives every utterance its own cluster), -1 = off, 2/4/8 = forced; UISRNN_B200_CLUSTER=0 disables
  // the automatic choice.
  pl->cluster = 1;
  pl->stat = 0;
  // Stationary-weights mode (uis_beam_stat.cuh): 32 CTAs per utterance keep the weights in shared memory.  The fastest
  // way to decode up to #SMs / 32 utterances at a time; opts->cluster = 32 forces it, 0 picks it automatically,
  // UISRNN_B200_STAT=0 disables the automatic choice.
  if (!tree && U >= 1 && (o->cluster == 0 || o->cluster == uis::kStatGroup) && !has_taps && m->depth == 1 &&
      o->lanes <= 1 && o->engine != 2) {
    const char* env = std::getenv("UISRNN_B200_STAT");
    const bool want = o->cluster == uis::kStatGroup ||
                      (!(env && env[0] == '0') && (long long)U * uis::kStatGroup <= ctas && o->engine == 0 && o->n_ctas <= 0);
    const int kc = o->kcap > 0 ? o->kcap : 32;
    const bool can = ctas >= uis::kStatGroup && uis::beam_stat_smem(m->H, m->D, pl->B, kc) <= 227u * 1024u &&
                     pl->B * kc + pl->B + 1 <= 65535;
    if (want && can) {
      pl->stat = std::max(1, std::min(ctas / uis::kStatGroup, U));
      pl->stat_forced = o->cluster == uis::kStatGroup;
      pl->tcn = 0;
      pl->Kcap = kc;
      pl->P = pl->B * kc + pl->B + 1;
      pl->G = 1;
      pl->ctas = pl->stat * uis::kStatGroup;
      return 0;
    }
    if (o->cluster == uis::kStatGroup)
      return fail(UIS_ERR_UNSUPPORTED, "stationary-weights mode needs hidden=512 dim=256 depth=1, >= 32 CTAs and beam_size/kcap that fit in shared memory");
  }
  if (!tree && !pl->tcn && U >= 1 && o->cluster >= 0 && !has_taps && m->depth == 1 && o->lanes <= 1) {
    int cs = 0;
    if (o->cluster == 2 || o->cluster == 4 || o->cluster == 8) {
      cs = o->cluster;
    } else if (o->cluster == 0) {
      const char* env = std::getenv("UISRNN_B200_CLUSTER");
      if (!(env && env[0] == '0'))
        for (int c : {4, 2})
          if ((long long)U * c <= ctas) { cs = c; break; }
    } else {
      return fail(UIS_ERR_INVALID, "cluster must be -1, 0, 2, 4, 8 or 32");
    }
    if (cs > 1 && uis::beam_cluster_smem(m->H, m->D, pl->B, pl->Kcap) <= 227u * 1024u) {
      const int clusters = std::max(1, std::min(ctas / cs, U));
      pl->cluster = cs;
      pl->cluster_forced = o->cluster > 0;
      pl->G = 1;
      pl->ctas = clusters * cs;
    } else if (o->cluster > 0) {
      return fail(UIS_ERR_UNSUPPORTED, "cluster mode needs hidden=512 dim=256 depth=1 and beam_size/kcap that fit in shared memory");
    }
  }
  return 0;
}

size_t workspace_bytes(const uis_model* m, const Plan& pl, int U) {
  size_t b = 0;
  b += (size_t)pl.rows * 3 * m->H * 4;                                  // gi
  b += (size_t)pl.ctas * pl.G * pl.P * (m->D + m->depth * m->H + 1) * 4;  // slot pools (+ Gaussian term per slot)
  b += (size_t)pl.ctas * pl.G * (pl.L > 1 ? (size_t)pl.maxTN + pl.maxSteps : (size_t)pl.maxN) * pl.B * 4;  // back-pointers
  b += (size_t)(U + 1) * 8 + (size_t)U * 8 + 256;                       // offsets, order, status
  if (pl.tcn) b += (size_t)pl.ctas * pl.tcn * m->H * 4;                 // a = relu(W1 h' + b1) between two products
  return b;
}

int run_device(uis_model* m, const float* x_dev, const int64_t* off, int U, const Plan& pl, int32_t* labels_dev,
               const uis_debug_taps* taps, cudaStream_t st, bool gi_ready = false) {
  const int H = m->H, D = m->D;
  m->stats = uis_stats{};
  m->stats.utterances = U;
  m->stats.frames = pl.rows;
  m->stats.ctas = pl.ctas;
  m->stats.lanes = pl.G;
  m->stats.cluster = pl.cluster;
  m->last_U = U;
  m->last_stream = st;
  m->stats_pending = false;
  if (U == 0 || pl.rows == 0) {
    return 0;
  }
  long long max_tn = 0;
  for (int u = 0; u < U; ++u) max_tn = std::max<long long>(max_tn, (off[u + 1] - off[u]) * pl.T);
  if (int rc = ensure_log_tables(m, (int)max_tn)) return rc;

  // schedule: longest utterance first (LPT) -- steps are strictly sequential per utterance
  std::vector<int> order(U);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(),
                   [&](int a, int b) { return off[a + 1] - off[a] > off[b + 1] - off[b]; });
  std::vector<long long> off_ll(off, off + U + 1);

  if (int rc = m->row_off.ensure((U + 1) * sizeof(long long))) return rc;
  if (int rc = m->order.ensure(U * sizeof(int))) return rc;
  if (int rc = m->status.ensure(U * sizeof(int))) return rc;
  if (int rc = m->queue_stats.ensure(40 * sizeof(unsigned long long))) return rc;
  if (int rc = m->gi.ensure((size_t)pl.rows * 3 * H * sizeof(float))) return rc;
  if (int rc = m->pool_mean.ensure((size_t)pl.ctas * pl.G * pl.P * D * sizeof(float))) return rc;
  if (int rc = m->pool_hidden.ensure((size_t)pl.ctas * pl.G * pl.P * m->depth * H * sizeof(float))) return rc;
  if (int rc = m->pool_mse.ensure((size_t)pl.ctas * pl.G * pl.P * sizeof(float))) return rc;
  if (int rc = m->bp.ensure((size_t)pl.ctas * pl.G * (pl.L > 1 ? (size_t)pl.maxTN + pl.maxSteps : (size_t)pl.maxN) * pl.B *
                            sizeof(unsigned)))
    return rc;

  if (pl.tcn)
    if (int rc = m->tc_scratch.ensure((size_t)pl.ctas * pl.tcn * H * sizeof(float))) return rc;
  if (pl.stat) {
    if (int rc = m->stat_bar.ensure((size_t)pl.stat * uis::kStatGroup * sizeof(unsigned))) return rc;
    if (int rc = m->stat_scratch.ensure((size_t)pl.stat * uis::kCPCluster * H * sizeof(float))) return rc;
    CU(cudaMemsetAsync(m->stat_bar.p, 0, (size_t)pl.stat * uis::kStatGroup * sizeof(unsigned), st));
  }

  CU(cudaMemcpyAsync(m->row_off.p, off_ll.data(), (U + 1) * sizeof(long long), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(m->order.p, order.data(), U * sizeof(int), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(m->queue_stats.p, 0, 40 * sizeof(unsigned long long), st));
  CU(cudaMemsetAsync(m->status.p, 0xff, U * sizeof(int), st));

  uis::BeamParams p{};
  p.whh_t = m->whh_t.as<float>(); p.w1_t = m->w1_t.as<float>(); p.w2_t = m->w2_t.as<float>();
  p.depth = m->depth;
  for (int l = 1; l < m->depth; ++l) {
    p.wih_up_t[l - 1] = m->wih_up_t.as<float>() + (size_t)(l - 1) * H * 3 * H;
    p.whh_up_t[l - 1] = m->whh_t.as<float>() + (size_t)l * H * 3 * H;
  }
  p.bih_up = m->bih.as<float>() + 3 * H;   // layers >= 1
  p.bhh_up = m->bhh.as<float>() + 3 * H;
  p.bhh = m->bhh.as<float>(); p.b1 = m->b1.as<float>(); p.b2 = m->b2.as<float>();
  p.wvec = m->wvec.as<float>(); p.mean0 = m->mean0.as<float>(); p.hidden0 = m->hidden0.as<float>();
  p.log_p0 = std::log(m->p0);          // np.log(self.transition_bias)      uisrnn.py:418
  p.log_1mp0 = std::log(1.0 - m->p0);  // np.log(1 - self.transition_bias)  uisrnn.py:416
  p.log_alpha = std::log(m->alpha);    // np.log(self.crp_alpha)            uisrnn.py:445
  p.logn = m->logn.as<double>(); p.logtot = m->logtot.as<double>();
  p.x = x_dev; p.gi = m->gi.as<float>();
  p.row_off = m->row_off.as<long long>(); p.order = m->order.as<int>();
  p.U = U; p.B = pl.B; p.Kcap = pl.Kcap; p.T = pl.T; p.P = pl.P; p.maxN = pl.maxN; p.G = pl.G;
  p.L = pl.L; p.node_cap = pl.node_cap; p.leaf_cap = pl.leaf_cap; p.maxTN = pl.maxTN; p.maxSteps = pl.maxSteps;
  { const char* e = getenv("UIS_DBG_MODE"); p.dbg_mode = e ? atoi(e) : 0; }
  p.pool_mean = m->pool_mean.as<float>(); p.pool_hidden = m->pool_hidden.as<float>(); p.pool_mse = m->pool_mse.as<float>();
  p.bp = m->bp.as<unsigned>();
  p.queue = m->queue_stats.as<int>();
  p.stats = m->queue_stats.as<unsigned long long>() + 8;
  p.labels = labels_dev; p.status = m->status.as<int>();
  p.trace_utt = -1;
  if (pl.stat) {
    p.stat_bar = m->stat_bar.as<unsigned>();
    p.stat_scratch = m->stat_scratch.as<float>();
  }
  if (pl.tcn) {
    p.tc_wmap = m->tc_map;
    p.tc_sh = m->tc_sh; p.tc_sa = m->tc_sa;
    p.tc_inv_hh = m->tc_inv_hh; p.tc_inv_1 = m->tc_inv_1; p.tc_inv_2 = m->tc_inv_2;
    p.tc_scratch = m->tc_scratch.as<float>();
  }

  long long trace_steps = 0;
  if (taps) {
    if (taps->final_scores) {
      if (int rc = m->dbg_final_scores.ensure((size_t)U * pl.B * 4)) return rc;
      p.dbg_final_scores = m->dbg_final_scores.as<float>();
      if (int rc = m->dbg_final_k.ensure((size_t)U * 4)) return rc;
      p.dbg_final_k = m->dbg_final_k.as<int>();
    }
    if (taps->trace_utt >= 0 && taps->trace_utt < U) {
      p.trace_utt = taps->trace_utt;
      p.trace_capacity = std::max(taps->trace_capacity, 0);
      trace_steps = ((off[p.trace_utt + 1] - off[p.trace_utt]) * pl.T + pl.L - 1) / pl.L;
      if (taps->step_winners && p.trace_capacity > 0) {
        if (int rc = m->dbg_win.ensure((size_t)p.trace_capacity * 4 * (1 + pl.L))) return rc;
        if (int rc = m->dbg_score.ensure((size_t)p.trace_capacity * 4)) return rc;
        if (int rc = m->dbg_off.ensure((size_t)(trace_steps + 1) * 8)) return rc;
        p.dbg_win = m->dbg_win.as<int>(); p.dbg_score = m->dbg_score.as<float>();
        p.dbg_off = m->dbg_off.as<long long>();
      }
      if (taps->best_mean) {
        if (int rc = m->dbg_best_mean.ensure((size_t)pl.Kcap * D * 4)) return rc;
        if (int rc = m->dbg_best_hidden.ensure((size_t)pl.Kcap * m->depth * H * 4)) return rc;
        if (int rc = m->dbg_best_blocks.ensure((size_t)pl.Kcap * 4)) return rc;
        p.dbg_best_mean = m->dbg_best_mean.as<float>(); p.dbg_best_hidden = m->dbg_best_hidden.as<float>();
        p.dbg_best_blocks = m->dbg_best_blocks.as<int>();
      }
    }
  }

  for (auto& e : m->ev)
    if (!e) CU(cudaEventCreate(&e));
  CU(cudaEventRecord(m->ev[0], st));
  // kernel 1: input projection GEMM (the host-buffer path has already run it chunk by chunk, under the H2D copies)
  if (!gi_ready) {
    dim3 grid((3 * H + uis::PBN - 1) / uis::PBN, (unsigned)((pl.rows + uis::PBM - 1) / uis::PBM));
    uis::input_proj_kernel<<<grid, 256, 0, st>>>(x_dev, m->wih_t.as<float>(), m->bih.as<float>(), m->gi.as<float>(),
                                                (int)pl.rows, 3 * H, D);
    CU(cudaGetLastError());
  }
  CU(cudaEventRecord(m->ev[1], st));
  // kernel 2: persistent beam search
  int cluster_used = pl.cluster;
  bool stat_done = false;
  if (pl.stat) {
    cudaError_t e = cudaSuccess;
    const bool have = uis::launch_beam_stat(H, D, p, pl.ctas, uis::beam_stat_smem(H, D, pl.B, pl.Kcap), st, &e);
    if (have && e == cudaSuccess) {
      cluster_used = uis::kStatGroup;
      stat_done = true;
    } else if (pl.stat_forced) {
      if (!have) return fail(UIS_ERR_UNSUPPORTED, "no stationary-weights kernel for hidden=%d dim=%d", H, D);
      return fail(UIS_ERR_CUDA, "stationary-weights beam kernel launch failed: %s", cudaGetErrorString(e));
    } else {
      // chosen automatically and refused (e.g. the CTAs cannot all be co-resident on a shared / partitioned GPU): the
      // one-CTA-per-utterance kernel runs on the same grid instead; its extra CTAs find the utterance queue empty
      (void)cudaGetLastError();
      cluster_used = 1;
    }
  }
  if (stat_done) {
  } else if (pl.tcn) {
    cudaError_t e = cudaSuccess;
    if (!uis::launch_beam_tc(H, D, pl.tcn, p, pl.ctas, uis::beam_tc_smem(H, D, pl.tcn, pl.B, pl.Kcap, pl.G), st, &e))
      return fail(UIS_ERR_UNSUPPORTED, "no tensor-core kernel for hidden=%d dim=%d columns=%d", H, D, pl.tcn);
    if (e != cudaSuccess) return fail(UIS_ERR_CUDA, "tensor-core beam kernel launch failed: %s", cudaGetErrorString(e));
  } else if (int rc = (pl.L > 1 ? dispatch_tree(H, D, p, pl.ctas, st)
                                : dispatch_beam(H, D, p, pl.ctas, &cluster_used, pl.cluster_forced, st))) {
    return rc;
  }
  m->stats.cluster = cluster_used;
  m->stats.engine = pl.tcn ? 2 : 1;
  m->stats.tc_columns = pl.tcn;
  CU(cudaEventRecord(m->ev[2], st));
  m->stats.kernel_launches = 2;
  m->stats_pending = true;

  if (taps) {
    CU(cudaStreamSynchronize(st));
    if (p.dbg_final_scores) {
      CU(cudaMemcpy(taps->final_scores, p.dbg_final_scores, (size_t)U * pl.B * 4, cudaMemcpyDeviceToHost));
      if (taps->final_k) CU(cudaMemcpy(taps->final_k, p.dbg_final_k, (size_t)U * 4, cudaMemcpyDeviceToHost));
    }
    if (p.dbg_win) {
      CU(cudaMemcpy(taps->step_winners, p.dbg_win, (size_t)p.trace_capacity * 4 * (1 + pl.L), cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(taps->step_scores, p.dbg_score, (size_t)p.trace_capacity * 4, cudaMemcpyDeviceToHost));
      if (taps->step_offsets)
        CU(cudaMemcpy(taps->step_offsets, p.dbg_off, (size_t)(trace_steps + 1) * 8, cudaMemcpyDeviceToHost));
    }
    if (p.dbg_best_mean) {
      CU(cudaMemcpy2D(taps->best_mean, (size_t)m->D_user * 4, p.dbg_best_mean, (size_t)D * 4, (size_t)m->D_user * 4, pl.Kcap,
                      cudaMemcpyDeviceToHost));
      if (taps->best_hidden)
        CU(cudaMemcpy2D(taps->best_hidden, (size_t)m->H_user * 4, p.dbg_best_hidden, (size_t)H * 4, (size_t)m->H_user * 4,
                        (size_t)pl.Kcap * m->depth, cudaMemcpyDeviceToHost));
      if (taps->best_blocks)
        CU(cudaMemcpy(taps->best_blocks, p.dbg_best_blocks, (size_t)pl.Kcap * 4, cudaMemcpyDeviceToHost));
    }
  }
  return 0;
}

// Pull the device-side counters and per-utterance status of the last call (synchronises).
int collect(uis_model* m) {
  if (!m->stats_pending) return 0;
  CU(cudaStreamSynchronize(m->last_stream));
  unsigned long long s[24];
  CU(cudaMemcpy(s, m->queue_stats.as<unsigned long long>() + 8, sizeof s, cudaMemcpyDeviceToHost));
  for (int i = 0; i < 10; ++i) m->stats.phase_cycles[i] = (int64_t)s[8 + i];
  for (int i = 0; i < 4; ++i) m->stats.tc_cycles[i] = (int64_t)s[18 + i];
  m->stats.gru_columns = (int64_t)s[0];
  m->stats.weight_passes = (int64_t)s[1];
  m->stats.candidates = (int64_t)s[2];
  m->stats.beam_steps = (int64_t)s[3];
  m->stats.max_k = (int32_t)s[4];
  CU(cudaEventElapsedTime(&m->stats.prepass_ms, m->ev[0], m->ev[1]));
  CU(cudaEventElapsedTime(&m->stats.beam_ms, m->ev[1], m->ev[2]));
  m->stats_pending = false;
  std::vector<int> status(m->last_U);
  CU(cudaMemcpy(status.data(), m->status.p, (size_t)m->last_U * sizeof(int), cudaMemcpyDeviceToHost));
  int overflow = 0, bad = 0, capacity = 0;
  for (int v : status) {
    if (v == -4) ++overflow;
    else if (v == -5) ++capacity;
    else if (v != 0) ++bad;
  }
  if (capacity)
    return fail(UIS_ERR_CAPACITY, "%d utterance(s): the look-ahead tree of one beam step outgrew the on-chip node arrays "
                "(lower beam_size / look_ahead / kcap)", capacity);
  if (overflow)
    return fail(UIS_ERR_OVERFLOW, "%d utterance(s) opened more clusters than kcap; retry with a larger kcap", overflow);
  if (bad) return fail(UIS_ERR_INVALID, "%d utterance(s) ended with no finite hypothesis", bad);
  return 0;
}

}  // namespace

extern "C" {

int uis_version(void) { return UIS_ABI_VERSION; }
const char* uis_last_error(void) { return g_err.c_str(); }

static int model_create_impl(uis_model** out, int device, int D, int H, int depth, const float* w_ih, const float* w_hh,
                             const float* b_ih, const float* b_hh, const float* w1, const float* b1, const float* w2,
                             const float* b2, const float* h0, const float* sigma2, double transition_bias,
                             double crp_alpha, int D_user, int H_user);

// Any (hidden <= 512, dim <= 256) runs in the smallest instantiated kernel shape that holds it, zero-padded: a padded
// hidden unit has zero weights and biases (r = z = 1/2, n = 0, so it stays at its initial 0) and feeds nothing; a
// padded observation dimension has x = mean = 0 and adds (0 - 0)^2 * w = 0 to every Gaussian term.  Adding exact
// zeros does not change an fp32 sum, so the results are those of a kernel instantiated for the caller's shape.
int uis_model_create(uis_model** out, int device, int D, int H, int depth, const float* w_ih, const float* w_hh,
                     const float* b_ih, const float* b_hh, const float* w1, const float* b1, const float* w2,
                     const float* b2, const float* h0, const float* sigma2, double transition_bias,
                     double crp_alpha) {
  if (!out) return fail(UIS_ERR_INVALID, "out is NULL");
  *out = nullptr;
  if (!w_ih || !w_hh || !b_ih || !b_hh || !w1 || !b1 || !w2 || !b2 || !h0 || !sigma2)
    return fail(UIS_ERR_INVALID, "NULL weight pointer");
  if (depth < 1 || depth > uis::kMaxDepth)
    return fail(UIS_ERR_UNSUPPORTED, "rnn_depth=%d: the sm_100a kernels support 1..%d stacked GRU layers", depth, uis::kMaxDepth);
  if (D < 1 || H < 1) return fail(UIS_ERR_INVALID, "observation_dim and rnn_hidden_size must be >= 1");
  if ((H > 512 || D > 256) && depth > 1)
    return fail(UIS_ERR_UNSUPPORTED, "hidden=%d dim=%d with rnn_depth=%d: models above hidden=512 / dim=256 run with one GRU layer", H, D, depth);
  if (shape_supported(H, D))
    return model_create_impl(out, device, D, H, depth, w_ih, w_hh, b_ih, b_hh, w1, b1, w2, b2, h0, sigma2,
                             transition_bias, crp_alpha, D, H);
  static const int shapes[4][2] = {{128, 64}, {256, 128}, {512, 256}, {1024, 512}};
  int Hp = 0, Dp = 0;
  for (auto& sh : shapes)
    if (!Hp && H <= sh[0] && D <= sh[1]) { Hp = sh[0]; Dp = sh[1]; }
  if (!Hp)
    return fail(UIS_ERR_UNSUPPORTED, "hidden=%d dim=%d: the sm_100a kernels hold models up to hidden=1024 dim=512", H, D);
  uis::DeviceGuard device_guard_(device);
  CU(device_guard_.status);
  std::vector<float> v, p_wih((size_t)3 * Hp * Dp + (size_t)(depth - 1) * 3 * Hp * Hp, 0.f), p_whh((size_t)depth * 3 * Hp * Hp, 0.f),
      p_bih((size_t)depth * 3 * Hp, 0.f), p_bhh((size_t)depth * 3 * Hp, 0.f), p_w1((size_t)Hp * Hp, 0.f), p_b1(Hp, 0.f),
      p_w2((size_t)Dp * Hp, 0.f), p_b2(Dp, 0.f), p_h0((size_t)depth * Hp, 0.f), p_s2(Dp, 1.f);
  // gate blocks (r, z, n) keep their own row ranges: row g * H + j -> g * Hp + j
  if (int r = fetch(v, w_ih, (size_t)3 * H * D + (size_t)(depth - 1) * 3 * H * H)) return r;
  for (int g = 0; g < 3; ++g)
    for (int j = 0; j < H; ++j)
      std::copy(v.begin() + ((size_t)g * H + j) * D, v.begin() + ((size_t)g * H + j + 1) * D,
                p_wih.begin() + ((size_t)g * Hp + j) * Dp);
  for (int l = 1; l < depth; ++l)
    for (int g = 0; g < 3; ++g)
      for (int j = 0; j < H; ++j) {
        const float* src = v.data() + (size_t)3 * H * D + (size_t)(l - 1) * 3 * H * H + ((size_t)g * H + j) * H;
        std::copy(src, src + H, p_wih.begin() + (size_t)3 * Hp * Dp + (size_t)(l - 1) * 3 * Hp * Hp + ((size_t)g * Hp + j) * Hp);
      }
  if (int r = fetch(v, w_hh, (size_t)depth * 3 * H * H)) return r;
  for (int l = 0; l < depth; ++l)
    for (int g = 0; g < 3; ++g)
      for (int j = 0; j < H; ++j) {
        const float* src = v.data() + (size_t)l * 3 * H * H + ((size_t)g * H + j) * H;
        std::copy(src, src + H, p_whh.begin() + (size_t)l * 3 * Hp * Hp + ((size_t)g * Hp + j) * Hp);
      }
  for (int which = 0; which < 2; ++which) {
    if (int r = fetch(v, which ? b_hh : b_ih, (size_t)depth * 3 * H)) return r;
    std::vector<float>& dst = which ? p_bhh : p_bih;
    for (int l = 0; l < depth; ++l)
      for (int g = 0; g < 3; ++g)
        std::copy(v.begin() + ((size_t)l * 3 + g) * H, v.begin() + ((size_t)l * 3 + g + 1) * H,
                  dst.begin() + ((size_t)l * 3 + g) * Hp);
  }
  if (int r = fetch(v, w1, (size_t)H * H)) return r;
  for (int j = 0; j < H; ++j) std::copy(v.begin() + (size_t)j * H, v.begin() + (size_t)(j + 1) * H, p_w1.begin() + (size_t)j * Hp);
  if (int r = fetch(v, b1, H)) return r;
  std::copy(v.begin(), v.end(), p_b1.begin());
  if (int r = fetch(v, w2, (size_t)D * H)) return r;
  for (int d = 0; d < D; ++d) std::copy(v.begin() + (size_t)d * H, v.begin() + (size_t)(d + 1) * H, p_w2.begin() + (size_t)d * Hp);
  if (int r = fetch(v, b2, D)) return r;
  std::copy(v.begin(), v.end(), p_b2.begin());
  if (int r = fetch(v, h0, (size_t)depth * H)) return r;
  for (int l = 0; l < depth; ++l) std::copy(v.begin() + (size_t)l * H, v.begin() + (size_t)(l + 1) * H, p_h0.begin() + (size_t)l * Hp);
  if (int r = fetch(v, sigma2, D)) return r;
  std::copy(v.begin(), v.end(), p_s2.begin());
  return model_create_impl(out, device, Dp, Hp, depth, p_wih.data(), p_whh.data(), p_bih.data(), p_bhh.data(), p_w1.data(),
                           p_b1.data(), p_w2.data(), p_b2.data(), p_h0.data(), p_s2.data(), transition_bias, crp_alpha, D, H);
}

static int model_create_impl(uis_model** out, int device, int D, int H, int depth, const float* w_ih, const float* w_hh,
                             const float* b_ih, const float* b_hh, const float* w1, const float* b1, const float* w2,
                             const float* b2, const float* h0, const float* sigma2, double transition_bias,
                             double crp_alpha, int D_user, int H_user) {
  if (!(transition_bias > 0.0 && transition_bias < 1.0))
    return fail(UIS_ERR_INVALID, "transition_bias must be in (0,1), got %g", transition_bias);
  if (!(crp_alpha > 0.0)) return fail(UIS_ERR_INVALID, "crp_alpha must be > 0");
  uis::DeviceGuard device_guard_(device);
  CU(device_guard_.status);
  uis_model* m = new uis_model();
  m->device = device; m->D = D; m->H = H; m->depth = depth; m->p0 = transition_bias; m->alpha = crp_alpha;
  m->D_user = D_user; m->H_user = H_user;
  int rc = 0;
  auto body = [&]() -> int {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    m->num_sms = prop.multiProcessorCount;
    std::vector<float> v;
    // w_ih = [layer 0: 3H x D | layers >= 1: 3H x H each]; w_hh = depth x [3H x H]; b_ih, b_hh = depth x [3H]
    const size_t n_ih = (size_t)3 * H * D + (size_t)(depth - 1) * 3 * H * H;
    if (int r = fetch(v, w_ih, n_ih)) return r;
    {
      std::vector<float> l0(v.begin(), v.begin() + (size_t)3 * H * D);
      auto t = transpose(l0, 3 * H, D);
      if (int r = upload(m->wih_t, t.data(), t.size() * 4)) return r;
      std::vector<float> up;
      for (int l = 1; l < depth; ++l) {
        std::vector<float> w(v.begin() + (size_t)3 * H * D + (size_t)(l - 1) * 3 * H * H,
                             v.begin() + (size_t)3 * H * D + (size_t)l * 3 * H * H);
        auto tt = transpose(w, 3 * H, H);
        up.insert(up.end(), tt.begin(), tt.end());
      }
      if (up.empty()) up.resize(4, 0.f);
      if (int r = upload(m->wih_up_t, up.data(), up.size() * 4)) return r;
    }
    if (int r = fetch(v, w_hh, (size_t)depth * 3 * H * H)) return r;
    {
      std::vector<float> all;
      for (int l = 0; l < depth; ++l) {
        std::vector<float> w(v.begin() + (size_t)l * 3 * H * H, v.begin() + (size_t)(l + 1) * 3 * H * H);
        auto tt = transpose(w, 3 * H, H);
        all.insert(all.end(), tt.begin(), tt.end());
      }
      if (int r = upload(m->whh_t, all.data(), all.size() * 4)) return r;
    }
    if (int r = fetch(v, w1, (size_t)H * H)) return r;
    { auto t = transpose(v, H, H); if (int r = upload(m->w1_t, t.data(), t.size() * 4)) return r; }
    if (int r = fetch(v, w2, (size_t)D * H)) return r;
    { auto t = transpose(v, D, H); if (int r = upload(m->w2_t, t.data(), t.size() * 4)) return r; }
    if (int r = fetch(v, b_ih, (size_t)depth * 3 * H)) return r;
    if (int r = upload(m->bih, v.data(), v.size() * 4)) return r;
    if (int r = fetch(v, b_hh, (size_t)depth * 3 * H)) return r;
    if (int r = upload(m->bhh, v.data(), v.size() * 4)) return r;
    if (int r = fetch(v, b1, H)) return r;
    if (int r = upload(m->b1, v.data(), v.size() * 4)) return r;
    if (int r = fetch(v, b2, D)) return r;
    if (int r = upload(m->b2, v.data(), v.size() * 4)) return r;
    if (int r = fetch(v, sigma2, D)) return r;
    for (float& s : v) s = 1.0f / (2.0f * s);  // weight = 1 / (2 * sigma2), two fp32 ops (uisrnn.py:414)
    if (int r = upload(m->wvec, v.data(), v.size() * 4)) return r;
    std::vector<float> h0v;
    if (int r = fetch(h0v, h0, (size_t)depth * H)) return r;
    DevBuf h0d;
    if (int r = upload(h0d, h0v.data(), (size_t)depth * H * 4)) return r;
    if (int r = m->mean0.ensure(D * 4)) return r;
    if (int r = m->hidden0.ensure((size_t)depth * H * 4)) return r;
    uis::init_state_kernel<<<1, H, 3 * H * sizeof(float)>>>(m->whh_t.as<float>(), m->wih_up_t.as<float>(),
                                                           m->w1_t.as<float>(), m->w2_t.as<float>(),
                                                           m->bih.as<float>(), m->bhh.as<float>(), m->b1.as<float>(),
                                                           m->b2.as<float>(), h0d.as<float>(), H, D, depth,
                                                           m->mean0.as<float>(), m->hidden0.as<float>());
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    h0d.release();
    if (depth == 1) {  // tensor-core pass: fp16 hi/lo planes of the untransposed (K-major) matrices + tensor map
      std::vector<float> whh0, w1v, b1v, w2v, hid0((size_t)H);
      if (int r = fetch(whh0, w_hh, (size_t)3 * H * H)) return r;
      if (int r = fetch(w1v, w1, (size_t)H * H)) return r;
      if (int r = fetch(b1v, b1, H)) return r;
      if (int r = fetch(w2v, w2, (size_t)D * H)) return r;
      CU(cudaMemcpy(hid0.data(), m->hidden0.p, (size_t)H * 4, cudaMemcpyDeviceToHost));
      if (int r = tc_prepare(m, whh0, w1v, b1v, w2v, hid0)) return r;
    }
    return 0;
  };
  rc = body();
  if (rc) {
    uis_model_destroy(m);
    return rc;
  }
  *out = m;
  return 0;
}

int uis_model_destroy(uis_model* m) {
  if (!m) return 0;
  uis::DeviceGuard device_guard_(m->device);
  DevBuf* bufs[] = {&m->wih_t, &m->whh_t, &m->w1_t, &m->w2_t, &m->bih, &m->bhh, &m->b1, &m->b2, &m->wvec, &m->mean0,
                    &m->hidden0, &m->wih_up_t, &m->logn, &m->logtot, &m->x64, &m->x32, &m->gi, &m->row_off, &m->order,
                    &m->pool_mean, &m->pool_hidden, &m->bp, &m->queue_stats, &m->labels, &m->status, &m->dbg_win,
                    &m->dbg_score, &m->dbg_off, &m->dbg_final_scores, &m->dbg_final_k, &m->dbg_best_mean,
                    &m->dbg_best_hidden, &m->dbg_best_blocks, &m->tc_planes, &m->tc_scratch, &m->pool_mse, &m->stat_bar, &m->stat_scratch};
  for (DevBuf* b : bufs) b->release();
  for (auto& e : m->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : m->ev_copied)
    if (e) cudaEventDestroy(e);
  for (auto& e : m->ev_free)
    if (e) cudaEventDestroy(e);
  for (auto& e : m->ev_h2d)
    if (e) cudaEventDestroy(e);
  if (m->ev_pipe) cudaEventDestroy(m->ev_pipe);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->labels_pin) cudaFreeHost(m->labels_pin);
  for (auto& e : m->ev_dma)
    if (e) cudaEventDestroy(e);
  if (m->pin_stage) cudaFreeHost(m->pin_stage);
  delete m->copy_pool;
  delete m;
  return 0;
}

int uis_model_constants(uis_model* m, float* mean0, float* hidden0) {
  if (!m) return fail(UIS_ERR_INVALID, "model is NULL");
  uis::DeviceGuard device_guard_(m->device);
  CU(device_guard_.status);
  if (mean0) CU(cudaMemcpy(mean0, m->mean0.p, m->D_user * 4, cudaMemcpyDeviceToHost));
  if (hidden0)
    CU(cudaMemcpy2D(hidden0, (size_t)m->H_user * 4, m->hidden0.p, (size_t)m->H * 4, (size_t)m->H_user * 4, m->depth,
                    cudaMemcpyDeviceToHost));
  return 0;
}

size_t uis_predict_workspace_bytes(uis_model* m, const int64_t* frame_offsets, int U, const uis_predict_opts* opts) {
  Plan pl;
  if (make_plan(m, frame_offsets, U, opts, &pl)) return 0;
  return workspace_bytes(m, pl, U);
}

int uis_predict_device(uis_model* m, const float* x_dev, const int64_t* frame_offsets, int U,
                       const uis_predict_opts* opts, int32_t* labels_dev, const uis_debug_taps* taps, void* stream) {
  Plan pl;
  if (int rc = make_plan(m, frame_offsets, U, opts, &pl, taps != nullptr)) return rc;
  if (U > 0 && pl.rows > 0 && (!x_dev || !labels_dev)) return fail(UIS_ERR_INVALID, "null device buffer");
  uis::DeviceGuard device_guard_(m->device);
  CU(device_guard_.status);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (m->D != m->D_user && pl.rows > 0) {  // zero-pad the caller's rows to the kernel's row length
    if (int rc = m->x32.ensure((size_t)pl.rows * m->D * 4)) return rc;
    const size_t np = (size_t)pl.rows * m->D;
    const int blocks = (int)std::min<size_t>((np + 255) / 256, (size_t)m->num_sms * 16);
    uis::pad_rows_f32_kernel<<<blocks, 256, 0, st>>>(x_dev, m->x32.as<float>(), (size_t)pl.rows, m->D_user, m->D);
    CU(cudaGetLastError());
    x_dev = m->x32.as<float>();
  }
  return run_device(m, x_dev, frame_offsets, U, pl, labels_dev, taps, st);
}

}  // extern "C"

namespace {

int ensure_host_path(uis_model* m) {
  if (!m->copy_stream) CU(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < uis_model::kSlots; ++i) {
    if (!m->ev_copied[i]) CU(cudaEventCreateWithFlags(&m->ev_copied[i], cudaEventDisableTiming));
    if (!m->ev_free[i]) CU(cudaEventCreateWithFlags(&m->ev_free[i], cudaEventDisableTiming));
  }
  for (auto& e : m->ev_h2d)
    if (!e) CU(cudaEventCreate(&e));
  if (!m->ev_pipe) CU(cudaEventCreate(&m->ev_pipe));
  for (auto& e : m->ev)
    if (!e) CU(cudaEventCreate(&e));
  return 0;
}

// Rows of one staging chunk of the host-buffer path (UISRNN_B200_CHUNK_MB of float64, default 32 MB: long enough
// for full PCIe rate, short enough that the first cast starts ~0.6 ms after the call).
size_t staging_chunk_rows(int d_user) {
  long long mb = 32;
  if (const char* env = std::getenv("UISRNN_B200_CHUNK_MB")) mb = std::max(0ll, std::atoll(env));  // 0 = the 256-row floor
  return std::max<size_t>(256, (size_t)mb * (1u << 20) / ((size_t)d_user * 8));
}

// One group of utterances, host buffers in, host buffers out: chunked H2D on the copy stream || cast + input
// projection on `st`, then the beam kernel, then one D2H copy of all labels.
int predict_host_group_impl(uis_model* m, const double* const* seqs, const int64_t* n_frames, int U, const int64_t* off,
                            const Plan& pl, int32_t* const* labels_out, const uis_debug_taps* taps, cudaStream_t st);

int predict_host_group(uis_model* m, const double* const* seqs, const int64_t* n_frames, int U, const int64_t* off,
                       const Plan& pl, int32_t* const* labels_out, const uis_debug_taps* taps, cudaStream_t st) {
  const int rc = predict_host_group_impl(m, seqs, n_frames, U, off, pl, labels_out, taps, st);
  if (rc != 0 && rc != UIS_ERR_OVERFLOW && rc != UIS_ERR_CAPACITY) {
    // a failed call may leave copies / kernels in flight on either stream: drain them (the error already recorded in
    // uis_last_error() is the one reported) so that the staging ring and the workspace are quiescent for the next call
    const std::string keep = g_err;
    if (m->copy_stream) cudaStreamSynchronize(m->copy_stream);
    cudaStreamSynchronize(st);
    (void)cudaGetLastError();
    g_err = keep;
  }
  return rc;
}

int predict_host_group_impl(uis_model* m, const double* const* seqs, const int64_t* n_frames, int U, const int64_t* off,
                            const Plan& pl, int32_t* const* labels_out, const uis_debug_taps* taps, cudaStream_t st) {
  const int D = m->D_user, H = m->H;  // the caller's rows; the device rows are padded to m->D floats
  const size_t rows = (size_t)pl.rows;
  if (rows == 0) {
    m->stats = uis_stats{};
    m->stats.utterances = U;
    return 0;
  }
  if (int rc = ensure_host_path(m)) return rc;
  const size_t chunk = std::min(staging_chunk_rows(D), rows);
  const int n_chunks = (int)((rows + chunk - 1) / chunk);
  const int slots = std::min(n_chunks, (int)uis_model::kSlots);
  if (int rc = m->x64.ensure((size_t)slots * chunk * D * 8)) return rc;
  if (int rc = m->x32.ensure(rows * m->D * 4)) return rc;
  if (int rc = m->gi.ensure(rows * 3 * H * sizeof(float))) return rc;
  if (int rc = m->labels.ensure(rows * 4)) return rc;
  if (rows * 4 > m->labels_pin_cap) {
    if (m->labels_pin) cudaFreeHost(m->labels_pin);
    m->labels_pin = nullptr;
    m->labels_pin_cap = 0;
    const size_t want = rows * 4 + rows / 2 + 4096;
    if (cudaMallocHost(&m->labels_pin, want) != cudaSuccess) {
      (void)cudaGetLastError();
      return fail(UIS_ERR_NOMEM, "cudaMallocHost(%zu) for the label staging buffer failed", want);
    }
    m->labels_pin_cap = want;
  }
  cudaStream_t cs = m->copy_stream;
  // Pageable or pinned?  Ordinary numpy arrays are pageable: the driver would stage every cudaMemcpy itself, in one
  // thread (measured 10.8 GB/s against 44.6 GB/s from pinned memory).  Such inputs go through a pinned ring of our own,
  // filled by a few host threads while the previous chunk is on the bus.  UISRNN_B200_HOST_STAGING=0 / 1 overrides.
  bool staged = false;
  {
    const char* env = std::getenv("UISRNN_B200_HOST_STAGING");
    if (env && (env[0] == '0' || env[0] == '1')) {
      staged = env[0] == '1';
    } else if (rows * (size_t)D * 8 >= ((size_t)8 << 20)) {  // small inputs: not worth waking the threads
      for (int q = 0; q < U && !staged; q += std::max(1, U / 8)) {  // a sample of the list
        if (n_frames[q] <= 0) continue;
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, seqs[q]) != cudaSuccess) { (void)cudaGetLastError(); staged = true; }
        else if (attr.type == cudaMemoryTypeUnregistered) staged = true;
      }
    }
  }
  const size_t chunk_bytes = chunk * (size_t)D * 8;
  if (staged) {
    if ((size_t)slots * chunk_bytes > m->pin_stage_cap) {
      if (m->pin_stage) cudaFreeHost(m->pin_stage);
      m->pin_stage = nullptr;
      m->pin_stage_cap = 0;
      if (cudaMallocHost(&m->pin_stage, (size_t)slots * chunk_bytes) != cudaSuccess) {
        (void)cudaGetLastError();
        staged = false;  // no pinned memory to be had: let the driver stage the copies
      } else {
        m->pin_stage_cap = (size_t)slots * chunk_bytes;
      }
    }
    for (auto& e : m->ev_dma)
      if (!e) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (staged && !m->copy_pool) {
      const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
      int nthreads = (int)std::min(16u, std::max(2u, hw / 2));  // measured on the B200 box: 4 -> 40 ms, 8 -> 29 ms, 16 -> 23 ms for 909 MB
      if (const char* env = std::getenv("UISRNN_B200_COPY_THREADS")) nthreads = std::max(1, std::min(64, std::atoi(env)));
      m->copy_pool = new CopyPool(nthreads);
    }
  }
  CU(cudaEventRecord(m->ev_h2d[0], cs));
  CU(cudaEventRecord(m->ev_pipe, st));
  size_t r0 = 0;
  int u = 0;
  std::vector<CopyPiece> pieces;
  for (int c = 0; r0 < rows; ++c) {
    const size_t r1 = std::min(rows, r0 + chunk);
    const int slot = c % uis_model::kSlots;
    double* stage = m->x64.as<double>() + (size_t)slot * chunk * D;
    if (staged) {
      pieces.clear();
      for (size_t r = r0; r < r1;) {
        while (u < U && (size_t)off[u + 1] <= r) ++u;
        const size_t take = std::min((size_t)off[u + 1], r1) - r;
        pieces.push_back(CopyPiece{reinterpret_cast<const char*>(seqs[u] + (r - (size_t)off[u]) * D), (r - r0) * D * 8, take * D * 8});
        r += take;
      }
      char* pin = m->pin_stage + (size_t)slot * chunk_bytes;
      if (c >= uis_model::kSlots) CU(cudaEventSynchronize(m->ev_dma[slot]));  // the DMA out of this pinned slot is done
      m->copy_pool->run(&pieces, pin, (r1 - r0) * D * 8);
      if (c >= uis_model::kSlots) CU(cudaStreamWaitEvent(cs, m->ev_free[slot], 0));
      CU(cudaMemcpyAsync(stage, pin, (r1 - r0) * D * 8, cudaMemcpyHostToDevice, cs));
      CU(cudaEventRecord(m->ev_dma[slot], cs));
    } else {
      if (c >= uis_model::kSlots) CU(cudaStreamWaitEvent(cs, m->ev_free[slot], 0));  // cast of chunk c - kSlots has read the slot
      for (size_t r = r0; r < r1;) {
        while (u < U && (size_t)off[u + 1] <= r) ++u;  // the utterance that holds row r (empty ones are skipped)
        const size_t take = std::min((size_t)off[u + 1], r1) - r;
        CU(cudaMemcpyAsync(stage + (r - r0) * D, seqs[u] + (r - (size_t)off[u]) * D, take * D * 8, cudaMemcpyHostToDevice, cs));
        r += take;
      }
    }
    CU(cudaEventRecord(m->ev_copied[slot], cs));
    CU(cudaStreamWaitEvent(st, m->ev_copied[slot], 0));
    const size_t nr = r1 - r0, np = nr * m->D;
    const int blocks = (int)std::min<size_t>((np + 255) / 256, (size_t)m->num_sms * 16);
    float* x32 = m->x32.as<float>() + r0 * m->D;
    if (m->D == D) uis::cast_f64_f32_kernel<<<blocks, 256, 0, st>>>(stage, x32, np);
    else uis::cast_pad_f64_f32_kernel<<<blocks, 256, 0, st>>>(stage, x32, nr, D, m->D);
    dim3 grid((3 * H + uis::PBN - 1) / uis::PBN, (unsigned)((nr + uis::PBM - 1) / uis::PBM));
    uis::input_proj_kernel<<<grid, 256, 0, st>>>(x32, m->wih_t.as<float>(), m->bih.as<float>(),
                                                m->gi.as<float>() + r0 * 3 * H, (int)nr, 3 * H, m->D);
    CU(cudaGetLastError());
    CU(cudaEventRecord(m->ev_free[slot], st));
    r0 = r1;
  }
  CU(cudaEventRecord(m->ev_h2d[1], cs));
  if (int rc = run_device(m, m->x32.as<float>(), off, U, pl, m->labels.as<int32_t>(), taps, st, /*gi_ready=*/true)) return rc;
  m->stats.kernel_launches = 1 + 2 * (int64_t)n_chunks;
  m->stats.chunks = n_chunks;
  m->stats.staged = staged ? 1 : 0;
  CU(cudaMemcpyAsync(m->labels_pin, m->labels.p, rows * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int q = 0; q < U; ++q)
    if (n_frames[q] > 0) std::memcpy(labels_out[q], m->labels_pin + off[q], (size_t)n_frames[q] * 4);
  if (int rc = collect(m)) return rc;
  CU(cudaEventElapsedTime(&m->stats.h2d_ms, m->ev_h2d[0], m->ev_h2d[1]));
  CU(cudaEventElapsedTime(&m->stats.pipeline_ms, m->ev_pipe, m->ev[1]));  // first cast -> beam kernel start
  return 0;
}

void add_stats(uis_stats* a, const uis_stats& b) {
  a->utterances += b.utterances; a->frames += b.frames; a->beam_steps += b.beam_steps; a->gru_columns += b.gru_columns;
  a->weight_passes += b.weight_passes; a->candidates += b.candidates; a->kernel_launches += b.kernel_launches;
  a->ctas = std::max(a->ctas, b.ctas); a->max_k = std::max(a->max_k, b.max_k);
  a->prepass_ms += b.prepass_ms; a->beam_ms += b.beam_ms; a->h2d_ms += b.h2d_ms; a->pipeline_ms += b.pipeline_ms;
  a->lanes = std::max(a->lanes, b.lanes); a->cluster = std::max(a->cluster, b.cluster);
  a->engine = std::max(a->engine, b.engine); a->tc_columns = std::max(a->tc_columns, b.tc_columns);
  for (int i = 0; i < 10; ++i) a->phase_cycles[i] += b.phase_cycles[i];
  for (int i = 0; i < 4; ++i) a->tc_cycles[i] += b.tc_cycles[i];
  a->chunks += b.chunks; a->groups += b.groups; a->staged = std::max(a->staged, b.staged);
}

}  // namespace

extern "C" {

int uis_predict(uis_model* m, const double* const* seqs, const int64_t* n_frames, int U, const uis_predict_opts* opts,
                int32_t* const* labels_out, const uis_debug_taps* taps, void* stream) {
  if (!m) return fail(UIS_ERR_INVALID, "model is NULL");
  if (U < 0 || (U > 0 && (!seqs || !n_frames || !labels_out))) return fail(UIS_ERR_INVALID, "null argument");
  const auto t_begin = std::chrono::steady_clock::now();
  std::vector<int64_t> off(U + 1, 0);
  for (int u = 0; u < U; ++u) {
    if (n_frames[u] < 0) return fail(UIS_ERR_INVALID, "negative length");
    if (n_frames[u] > 0 && (!seqs[u] || !labels_out[u])) return fail(UIS_ERR_INVALID, "null utterance buffer");
    off[u + 1] = off[u] + n_frames[u];
  }
  Plan pl;
  if (int rc = make_plan(m, off.data(), U, opts, &pl, taps != nullptr)) return rc;
  uis::DeviceGuard device_guard_(m->device);
  CU(device_guard_.status);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pl.rows == 0) return 0;
  // Memory: the per-frame workspace (gi 12H B + fp32 rows 4D B + labels) is the part that grows with the input.
  // A list that does not fit the device at once is decoded in groups of whole utterances, one after the other
  // (utterances are independent, uisrnn.py:587-589); UISRNN_B200_MAX_ROWS forces a limit (tests).
  const size_t per_row = (size_t)3 * m->H * 4 + (size_t)m->D * 4 + 4;
  size_t max_rows = 0;
  if (const char* env = std::getenv("UISRNN_B200_MAX_ROWS")) max_rows = (size_t)std::max(1ll, std::atoll(env));
  if (!max_rows) {
    size_t free_b = 0, total_b = 0;
    CU(cudaMemGetInfo(&free_b, &total_b));
    const size_t held = m->gi.cap + m->x32.cap + m->labels.cap + m->x64.cap;  // re-used by this call
    const size_t fixed = workspace_bytes(m, pl, U) - (size_t)pl.rows * 3 * m->H * 4 +
                         (size_t)uis_model::kSlots * staging_chunk_rows(m->D_user) * m->D_user * 8;
    const double budget = 0.9 * (double)(free_b + held) - (double)fixed;
    max_rows = budget > (double)per_row ? (size_t)(budget / (double)per_row) : 1;
  }
  if ((size_t)pl.rows <= max_rows || taps || U <= 1) {
    if (int rc = predict_host_group(m, seqs, n_frames, U, off.data(), pl, labels_out, taps, st)) return rc;
    m->stats.groups = 1;
  } else {
    uis_stats total{};
    int u0 = 0;
    while (u0 < U) {
      int u1 = u0 + 1;
      while (u1 < U && (size_t)(off[u1 + 1] - off[u0]) <= max_rows) ++u1;
      std::vector<int64_t> goff(u1 - u0 + 1);
      for (int q = u0; q <= u1; ++q) goff[q - u0] = off[q] - off[u0];
      Plan gp;
      if (int rc = make_plan(m, goff.data(), u1 - u0, opts, &gp, false)) return rc;
      if (int rc = predict_host_group(m, seqs + u0, n_frames + u0, u1 - u0, goff.data(), gp, labels_out + u0, nullptr, st))
        return rc;
      m->stats.groups = 1;
      add_stats(&total, m->stats);
      u0 = u1;
    }
    m->stats = total;
  }
  m->stats.host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  return 0;
}

int uis_get_stats(uis_model* m, uis_stats* out) {
  if (!m || !out) return fail(UIS_ERR_INVALID, "null argument");
  uis::DeviceGuard device_guard_(m->device);
  CU(device_guard_.status);
  const int rc = collect(m);
  *out = m->stats;
  return rc;
}

}  // extern "C"
