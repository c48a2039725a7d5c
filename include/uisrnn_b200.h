/*
 * uisrnn_b200.h -- C ABI of libuisrnn_b200.so: the B200 (sm_100a) implementation of UIS-RNN's
 * predict() hot path (beam search over GRU hypotheses).
 *
 * The reference (google/uis-rnn) has NO native / FFI layer: its only boundary
 * is the Python API.  This header is the seam a maintainer would bind (ctypes stub in
 * INTEGRATION.md); each entry point names the reference code it replaces.  All paths are
 * relative to /root/reference.
 *
 * Conventions
 *   - return 0 on success, negative uis_status on failure; message via uis_last_error()
 *     (thread-local).  Nothing throws across the ABI.
 *   - a uis_model is bound to one CUDA device; calls on one handle must be serialised by the
 *     caller; different handles may be used from different threads/processes.
 *   - caller owns every input/output buffer; the library owns the handle and its workspace.
 *   - all device work is ordered on the `stream` argument (a cudaStream_t, NULL = default
 *     stream).  The *_device entry point does not synchronise; the host-buffer entry point
 *     returns after the labels have landed in the caller's host buffers.
 */
#ifndef UISRNN_B200_H_
#define UISRNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UIS_ABI_VERSION 4

typedef enum uis_status {
  UIS_OK = 0,
  UIS_ERR_INVALID = -1,     /* bad argument (shape, NULL, beam_size < 1, ...)                  */
  UIS_ERR_UNSUPPORTED = -2, /* shape / option the sm_100a kernels are not instantiated for     */
  UIS_ERR_CUDA = -3,        /* a CUDA runtime call failed; uis_last_error() has the string     */
  UIS_ERR_OVERFLOW = -4,    /* a hypothesis opened more than `kcap` clusters; retry with more  */
  UIS_ERR_NOMEM = -5,
  UIS_ERR_CAPACITY = -6     /* look_ahead >= 2: a beam step's candidate tree outgrew on-chip storage */
} uis_status;

typedef struct uis_model uis_model; /* opaque */

/* Inference options = the reference's inference_args (uisrnn/arguments.py:172-193). */
typedef struct uis_predict_opts {
  int32_t beam_size;      /* --beam_size      (arguments.py:175-180), 1..128 (1..32 when look_ahead >= 2) */
  int32_t look_ahead;     /* --look_ahead     (arguments.py:181-185), >= 1                     */
  int32_t test_iteration; /* --test_iteration (arguments.py:186-193), >= 1                     */
  int32_t kcap;           /* max clusters per hypothesis held on device; 0 = default (32; 16
                             when look_ahead >= 2 or on the tensor-core engine; beam_size > 32:
                             the largest of 32, 16, 8, 4 whose tables fit shared memory)        */
  int32_t n_ctas;         /* persistent CTAs to launch; 0 = one per SM                         */
  int32_t lanes;          /* utterances advanced together per CTA (share each weight pass);
                             0 = auto (FFMA engine: 2 when U >= 2 * CTAs, else 1, max 4;
                             tensor-core engine: up to columns / 8 = 6, max 8)                  */
  int32_t cluster;        /* CTAs per utterance (latency modes for few utterances; default shape,
                             depth 1, look_ahead 1): 0 = auto (32 when U * 32 <= CTAs, else 4 or 2 when
                             U * cluster <= CTAs), -1 = off, 2 / 4 / 8 = thread-block cluster (k-split of the
                             streamed weights), 32 = stationary-weights group (weights resident in the
                             shared memory of 32 CTAs, products split by rows, cooperative launch)         */
  int32_t engine;         /* matrix engine of the look_ahead-1 beam kernel: 0 = auto (tensor cores
                             when some CTA gets more than one utterance), 1 = fp32 FFMA kernels,
                             2 = tcgen05 tensor-core pass (fp16 hi/lo split operands, fp32-grade;
                             depth 1, hidden/dim multiples of 128; kcap defaults to 16)          */
} uis_predict_opts;

/* Optional per-call debug / parity taps.  Any pointer may be NULL.  All are HOST buffers
 * the library fills before uis_predict*() returns (it synchronises the stream if any tap is set). */
typedef struct uis_debug_taps {
  int32_t trace_utt;       /* utterance index to trace step by step, -1 = none                 */
  int32_t trace_capacity;  /* rows available in step_winners / step_scores                     */
  int32_t* step_winners;   /* [trace_capacity][1+look_ahead]: (parent beam, cluster...) rows,
                              ranked order, all steps concatenated -- what uisrnn.py:551-556
                              unravels                                                         */
  float* step_scores;      /* [trace_capacity] neg_likelihood of each new hypothesis           */
  int64_t* step_offsets;   /* [steps+1] row offsets per beam step                              */
  float* final_scores;     /* [U][beam_size] final neg_likelihood per hypothesis (+inf pad)    */
  int32_t* final_k;        /* [U] clusters in the best hypothesis                              */
  float* best_mean;        /* [kcap][D]  mean_set   of the best hypothesis of `trace_utt`      */
  float* best_hidden;      /* [kcap][depth][H] hidden_set of the best hypothesis of `trace_utt` */
  int32_t* best_blocks;    /* [kcap]     block_counts of the best hypothesis of `trace_utt`    */
} uis_debug_taps;

/* Work counters of the last uis_predict*() call on this handle (for bench.py's accounting). */
typedef struct uis_stats {
  int64_t utterances;
  int64_t frames;          /* un-tiled input rows                                              */
  int64_t beam_steps;      /* sum over utterances of test_iteration * N / look_ahead (ceil)    */
  int64_t gru_columns;     /* GRU+MLP evaluations actually performed                           */
  int64_t weight_passes;   /* full passes over (W_hh, W1, W2) streamed by all CTAs             */
  int64_t candidates;      /* scored (hypothesis, cluster) candidates                          */
  int64_t kernel_launches; /* CUDA kernels launched by the call                                */
  int32_t ctas;            /* persistent CTAs used                                             */
  int32_t max_k;           /* largest cluster count seen in any hypothesis                     */
  float prepass_ms;        /* device time of the input-projection GEMM (CUDA events on `stream`) */
  float beam_ms;           /* device time of the persistent beam-search kernel                 */
  int32_t lanes;           /* lanes per CTA used                                               */
  int32_t cluster;   /* CTAs per utterance the last call used: 1 = none, 2/4/8 = cluster, 32 = stationary-weights group */
  int32_t engine;          /* 1 = FFMA kernels, 2 = tensor-core pass                            */
  int32_t tc_columns;      /* tensor-core pass: columns per weight pass (0 otherwise)           */
  int64_t phase_cycles[10]; /* SM cycles summed over CTAs: [0] re-pack (P4), [1] gather, [2] GRU pass,
                               [3] W1 pass, [4] W2 pass, [5] advance/back-track, [6] frame landing
                               (P0), [7] scoring (P1), [8] ranking (P2), [9] column/slot assignment (P3) */
  int64_t tc_cycles[4];     /* tensor-core pass, SM cycles of the MMA-issuing thread summed over CTAs: stalled on [0] a
                               weight box not yet landed (TMA), [1] an accumulator slot not yet drained (epilogue),
                               [2] the B operand of the next product; [3] inside passes (first operand ready -> last issue) */
  /* host-buffer entry point (uis_predict) only, ABI 4: */
  float h2d_ms;             /* span of the chunked host->device copies on the copy stream                              */
  float pipeline_ms;        /* compute stream: first cast kernel -> start of the beam kernel (casts + input projections,
                               overlapped with the copies)                                                             */
  float host_ms;            /* wall time inside uis_predict()                                                          */
  int32_t chunks;           /* staging chunks the float64 rows travelled in                                            */
  int32_t groups;           /* > 1: the list did not fit the device at once and was decoded in this many groups        */
  int32_t staged;           /* 1: the inputs were pageable and went through the library's pinned staging ring (host
                               copy threads), 0: copied straight from the caller's (pinned) buffers                     */
} uis_stats;

int uis_version(void);
const char* uis_last_error(void);

/*
 * Replaces UISRNN.__init__ / load (uisrnn/uisrnn.py:83-107, 149-170) for the inference path:
 * takes the CoreRNN parameters (uisrnn.py:35-43, PyTorch state_dict layout, row-major fp32):
 *   w_ih [3H,D]  w_hh [3H,H]  b_ih [3H]  b_hh [3H]   (gru.*_l0, gate order r,z,n)
 *   w1 [H,H] b1 [H] (linear_mean1)   w2 [D,H] b2 [D] (linear_mean2)
 *   h0 [depth,H] (rnn_init_hidden)   sigma2 [D]
 * Stacked layers (depth 2..4): w_ih = [gru.weight_ih_l0 (3H x D) | gru.weight_ih_l1 (3H x H) | ...]
 * concatenated, w_hh / b_ih / b_hh = the per-layer tensors concatenated in layer order.
 * Pointers may be host or device memory (copied, never retained).  Precomputes the per-model
 * constants CoreRNN(zeros, rnn_init_hidden) that uisrnn.py:435-439 recomputes per candidate.
 */
int uis_model_create(uis_model** out, int device, int D, int H, int depth,
                     const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                     const float* w1, const float* b1, const float* w2, const float* b2,
                     const float* h0, const float* sigma2, double transition_bias,
                     double crp_alpha);
int uis_model_destroy(uis_model* m);

/* Copies the per-model constants back (host buffers, fp32): mean0 [D], hidden0 [depth][H]. */
int uis_model_constants(uis_model* m, float* mean0, float* hidden0);

/*
 * Replaces UISRNN.predict / predict_single / parallel_predict (uisrnn/uisrnn.py:479-623) for a
 * list of U utterances held in HOST memory:
 *   seqs[u]      -> row-major float64 [n_frames[u], D]   (the ndarray the reference validates at
 *                   uisrnn.py:511-521; cast to fp32 on the device, = uisrnn.py:525-526)
 *   labels_out[u]-> int32 [n_frames[u]]  = beam_set[0].trace[-N:] (uisrnn.py:561)
 * Host->device copies, the beam search and the device->host copy of the labels all happen
 * inside the call.  Utterances are independent; they are scheduled longest-first over
 * persistent CTAs.
 */
int uis_predict(uis_model* m, const double* const* seqs, const int64_t* n_frames, int U,
                const uis_predict_opts* opts, int32_t* const* labels_out,
                const uis_debug_taps* taps, void* stream);

/*
 * Same search with inputs already resident in HBM:
 *   x_dev      fp32 [frame_offsets[U], D] (all utterances concatenated), device memory
 *   frame_offsets  HOST int64 [U+1]
 *   labels_dev int32 [frame_offsets[U]], device memory
 * Asynchronous on `stream` unless taps != NULL.
 */
int uis_predict_device(uis_model* m, const float* x_dev, const int64_t* frame_offsets, int U,
                       const uis_predict_opts* opts, int32_t* labels_dev,
                       const uis_debug_taps* taps, void* stream);

/* Device bytes uis_predict_device() will hold for this problem (workspace is cached in the handle). */
size_t uis_predict_workspace_bytes(uis_model* m, const int64_t* frame_offsets, int U,
                                   const uis_predict_opts* opts);

int uis_get_stats(uis_model* m, uis_stats* out);

/* ------------------------------------------------------------------------------------------------
 * Training: one iteration of UISRNN.fit_concatenated (uisrnn/uisrnn.py:252-295) on the device.
 * Parameters are 4 * depth + 6 tensors, in this order, row-major fp32 (PyTorch layouts); depth 1:
 *   0 gru.weight_ih_l0 [3H,D]  1 gru.weight_hh_l0 [3H,H]  2 gru.bias_ih_l0 [3H]  3 gru.bias_hh_l0 [3H]
 *   4 linear_mean1.weight [H,H] 5 linear_mean1.bias [H]   6 linear_mean2.weight [D,H] 7 linear_mean2.bias [D]
 *   8 rnn_init_hidden [H]       9 sigma2 [D]
 * depth > 1: the four gru tensors of layer 0, then of layer 1 (weight_ih_l1 is [3H,H]), ..., then linear_mean1/2,
 * rnn_init_hidden [depth,H], sigma2.
 * (all but the last two = the "rnn parameters" group that is norm-clipped, uisrnn.py:120-133, 292.)
 */
typedef struct uis_trainer uis_trainer; /* opaque; owns parameters, gradients and Adam state */

typedef struct uis_train_hparams {          /* training_args, uisrnn/arguments.py:105-169 */
  float learning_rate;                      /* --learning_rate                                   */
  float sigma_alpha, sigma_beta;            /* --sigma_alpha / --sigma_beta                      */
  float regularization_weight;              /* --regularization_weight                           */
  float grad_max_norm;                      /* --grad_max_norm                                   */
  int32_t train_sigma2;                     /* 1 if sigma2 is estimated (model_args.sigma2 None) */
  /* ABI 4: stacked GRU layers (model_args, uisrnn/arguments.py:55-64; nn.GRU(num_layers, dropout), uisrnn.py:35-43) */
  int32_t rnn_depth;                        /* --rnn_depth, 1..4 (0 = 1)                         */
  float rnn_dropout;                        /* --rnn_dropout: applied to the output sequence of every layer but the
                                               last, in every training iteration (train mode), when rnn_depth > 1   */
  int64_t dropout_seed;                     /* seed of the dropout masks: keep(i) is a pure function of
                                               (seed, iteration, layer, element) -- see uis_train.cu dropout_hash   */
} uis_train_hparams;

/* params: 4 * hp->rnn_depth + 6 host (or device) pointers, copied.  Adam state starts at zero (a fresh optimiser per
 * fit_concatenated call, uisrnn.py:235-236). */
int uis_trainer_create(uis_trainer** out, int device, int D, int H, const float* const* params,
                       const uis_train_hparams* hp);
int uis_trainer_destroy(uis_trainer* t);

/* One iteration on one batch = what utils.pack_sequence builds (utils.py:237-246): x_host fp32
 * [L][B][D] zero-padded, time-major, row 0 all zeros; lengths[B] (incl. the zero row) sorted
 * descending with lengths[0] == L; any B >= 1 (the recurrence runs in groups of 32 columns).  mode 0: forward + backward + clip + Adam + clamp;
 * mode 1: forward + backward only (for gradient checks); mode 2: data-parallel shard (see below).  losses_out[3] (host, may be NULL) =
 * negative log likelihood, sigma2 prior, regularisation -- the three numbers uisrnn.py:297-310 logs.
 * With losses_out == NULL the call only enqueues work on `stream` (the host batch has been staged
 * when it returns); read the losses later with uis_trainer_losses(). */
int uis_trainer_step(uis_trainer* t, const float* x_host, const int32_t* lengths, int B, int L, int mode,
                     float* losses_out, void* stream);

/* what = 0: current parameters, 1: gradients of the last step.  out: 4 * depth + 6 host pointers (NULL = skip). */
int uis_trainer_get(uis_trainer* t, int what, float* const* out);

/* Losses of the last `count` (<= 4096) steps, oldest first: out[count][3] host floats.  Synchronises. */
int uis_trainer_losses(uis_trainer* t, int count, float* out);

/*
 * Training set resident on the device (replaces the per-iteration host work of
 * utils.pack_sequence, utils.py:204-250, and the num_permutations-fold float64 copy of
 * utils.resize_sequence, utils.py:172-201).  rows: host float64 [n_rows][D] = the concatenated training
 * sequence (cast to fp32 on the device); index: host int32 [n_index] = the row indices of every
 * sub-sequence back to back; offsets: host int64 [n_sub + 1].  uis_trainer_step_corpus() then runs one
 * iteration on the batch whose columns are sub-sequences chosen[0..B) (the caller keeps the reference's RNG
 * draw and passes the ids in pack_sequence's column order: lengths descending); the batch tensor is
 * gathered on the device, zero frame first, zero padded.
 */
int uis_trainer_set_corpus(uis_trainer* t, const double* rows, int64_t n_rows, const int32_t* index, int64_t n_index,
                           const int64_t* offsets, int32_t n_sub);
int uis_trainer_step_corpus(uis_trainer* t, const int32_t* chosen, int B, int mode, float* losses_out, void* stream);

/*
 * Data-parallel fit() (optional): every rank runs uis_trainer_step(mode = 2) on its
 * shard of the mini-batch (forward + backward with UN-normalised gradients), exports
 *   [gradients of all parameters but sigma2 | per-dimension squared-residual sums | per-dimension counts | row count]
 * (uis_trainer_comm_size() floats) into a caller-owned DEVICE buffer, all-reduces(sum) it (NCCL over
 * NVLink: one collective per iteration), and hands it back: uis_trainer_comm_apply() normalises by the
 * global row count, forms the sigma2 gradient and the three losses from the global statistics, adds the
 * regulariser, clips and takes the Adam step -- identically on every rank.
 */
int64_t uis_trainer_comm_size(uis_trainer* t);
int uis_trainer_comm_export(uis_trainer* t, float* dev_buf, void* stream);
int uis_trainer_comm_apply(uis_trainer* t, const float* dev_buf, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UISRNN_B200_H_ */
