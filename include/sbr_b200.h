/* sbr_b200.h -- C ABI of libsbr_b200.so: the B200-native replacement for the three callables
 * that theano.function compiles on the reference's RNN training hot path.
 *
 * Reference interface replaced (paths relative to rdevooght/sequence-based-recommendations):
 *   train_function(*theano_inputs) -> cost, in-place parameter/optimizer update
 *       built at neural_networks/rnn_base.py:175-186, called at rnn_base.py:290
 *       inputs: OneHot   [X, mask, Y, pop, exclude]           rnn_one_hot.py:61
 *               Sampling [X, mask, Y, samples, pop, exclude]  rnn_sampling.py:128
 *               Margin   [X, mask, Ymat, weight, exclude]     rnn_margin.py:100
 *   test_function(theano_inputs, k) -> ids[k]     rnn_base.py:196-213, rnn_sampling.py:140-157
 *   predict_function(X, mask) -> scores[1,N]      rnn_base.py:188-194
 *   lasagne.layers.get/set_all_param_values       rnn_base.py:476,515
 *
 * Conventions
 *   - plain C, no C++/torch types; every pointer argument is a HOST pointer, borrowed for the
 *     duration of the call (C-contiguous, caller-owned, e.g. a numpy buffer);
 *   - every function returns 0 on success, a negative sbr_status otherwise; the message is read
 *     with sbr_last_error(); errors coming from CUDA/NCCL are sticky on the handle;
 *   - a handle is NOT thread-safe: one handle per GPU rank, one caller thread per handle;
 *   - X is int32 [B, max_length, ids_per_step], left-aligned; mask is float32 [B, max_length]
 *     with mask[b, :len_b] = 1 exactly as _prepare_input builds it (rnn_one_hot.py:100-101).
 *     A mask that is not a left-aligned run of ones is rejected with SBR_E_MASK;
 *   - `exclude` of the reference's train inputs is not part of the ABI: the train cost never reads
 *     it (on_unused_input='ignore', rnn_base.py:185); at test time the excluded ids are passed
 *     as a ragged list (sbr_topk);
 *   - parameters are addressed by their index in the reference checkpoint order
 *     (lasagne get_all_param_values order, rnn_base.py:476), per-gate matrices, row-major.
 */
#ifndef SBR_B200_H
#define SBR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SBR_API __attribute__((visibility("default")))
#else
#define SBR_API
#endif

#define SBR_ABI_VERSION 2
#define SBR_MAX_LAYERS 8
#define SBR_NCCL_ID_BYTES 128

typedef struct sbr_model sbr_model;

typedef enum sbr_status {
  SBR_OK = 0,
  SBR_E_ARG = -1,       /* bad argument / unsupported configuration        */
  SBR_E_CUDA = -2,      /* CUDA runtime or driver error (sticky)           */
  SBR_E_NCCL = -3,      /* NCCL error or libnccl not loadable (sticky)     */
  SBR_E_MASK = -4,      /* mask is not a left-aligned run of ones          */
  SBR_E_RANGE = -5,     /* an id is outside [0, n_items + n_extra_ids)     */
  SBR_E_NOGPU = -6      /* no CUDA device: there is no CPU fallback        */
} sbr_status;

enum { SBR_CELL_LSTM = 0, SBR_CELL_GRU = 1, SBR_CELL_VANILLA = 2 };          /* --r_t, recurrent_layers.py:9 */
enum { SBR_LOSS_CCE = 0, SBR_LOSS_BPR = 1, SBR_LOSS_BPRI = 2, SBR_LOSS_TOP1 = 3,
       SBR_LOSS_BLACKOUT = 4, SBR_LOSS_HINGE = 5, SBR_LOSS_LOGIT = 6, SBR_LOSS_LOGSIG = 7 }; /* --loss */
enum { SBR_UPD_ADAM = 0, SBR_UPD_ADAGRAD = 1, SBR_UPD_ADADELTA = 2, SBR_UPD_RMSPROP = 3,
       SBR_UPD_NESTEROV = 4 };                                               /* --u_m, update_manager.py:4 */
/* arithmetic of the GEMM-shaped stages; both keep fp32 storage and fp32 accumulation */
enum { SBR_MATH_FP32 = 0,   /* fp32-accurate: CUDA-core FFMA or 3xTF32 split on tcgen05 */
       SBR_MATH_TF32 = 1 }; /* single-pass TF32 on tcgen05 (10-bit mantissa inputs)       */

typedef struct sbr_config {
  int32_t struct_size;              /* = sizeof(sbr_config), ABI guard                         */
  int32_t cell;                     /* SBR_CELL_*                                              */
  int32_t n_layers;                 /* --r_l "a-b-c"                                           */
  int32_t layers[SBR_MAX_LAYERS];
  int32_t n_items;                  /* dataset.n_items (rnn_base.py:109)                       */
  int32_t n_extra_ids;              /* optional-feature id rows appended after the items (10 with --rf) */
  int32_t ids_per_step;             /* K = RNNBase._input_size() (rnn_base.py:615-622)         */
  int32_t embedding;                /* --r_emb, 0 = gather-sum layer 0                         */
  int32_t max_length;               /* T, --max_length                                         */
  int32_t batch_size;               /* rows per call on THIS rank (local batch)                */
  int32_t loss;                     /* SBR_LOSS_*                                              */
  int32_t n_samples;                /* S of RNNSampling (rnn_sampling.py:105-108)              */
  int32_t last_layer_tanh;          /* rnn_sampling.py:19                                      */
  int32_t updater;                  /* SBR_UPD_*                                               */
  float lr, rho, beta1, beta2;      /* update_manager.py:5-8                                   */
  float grad_clip;                  /* always 100 in the reference (recurrent_layers.py:19)    */
  float regularization;             /* output-bias L2 (>0) / L1 (<0), rnn_one_hot.py:73-77      */
  int32_t math_mode;                /* SBR_MATH_*                                              */
  int32_t device;                   /* CUDA ordinal                                            */
  int32_t n_ranks;                  /* data-parallel world size (1 = no NCCL)                  */
  int32_t rank;
  int32_t global_batch;             /* rows over all ranks; 0 -> batch_size * n_ranks          */
  int32_t n_slots;                  /* device-resident batch slots (>=1), see sbr_stage_*      */
  int32_t bidirectional;            /* --r_bi: every depth = forward + backwards layer, concatenated (recurrent_layers.py:72-78) */
  uint8_t nccl_id[SBR_NCCL_ID_BYTES]; /* from sbr_nccl_unique_id on rank 0, shared by the host */
} sbr_config;

/* ---- life cycle ------------------------------------------------------------------------- */
SBR_API int sbr_abi_version(void);
SBR_API int sbr_device_count(void);                       /* 0 when no CUDA device is visible          */
SBR_API int sbr_nccl_unique_id(uint8_t out[SBR_NCCL_ID_BYTES]);
SBR_API int sbr_create(const sbr_config* cfg, sbr_model** out);   /* replaces _prepare_networks + _compile_* */
SBR_API void sbr_destroy(sbr_model* m);
SBR_API const char* sbr_last_error(const sbr_model* m);   /* m == NULL: error of the last failed sbr_create */

/* ---- parameters: lasagne get/set_all_param_values (rnn_base.py:476,515) ------------------ */
SBR_API int sbr_param_count(const sbr_model* m);
SBR_API int sbr_param_info(const sbr_model* m, int idx, char* name, int name_cap, int* ndim, int64_t shape[4]);
SBR_API int sbr_get_param(sbr_model* m, int idx, float* host);
SBR_API int sbr_set_param(sbr_model* m, int idx, const float* host);
SBR_API int sbr_get_grad(sbr_model* m, int idx, float* host);     /* gradient left by the last step when skip_update=1 */
SBR_API int64_t sbr_total_params(const sbr_model* m);
SBR_API int sbr_reset_optimizer(sbr_model* m);                    /* zero the updater state and its step counter */
SBR_API int sbr_set_skip_update(sbr_model* m, int flag);          /* 1: steps compute cost+gradients only (tests) */

/* ---- train_function --------------------------------------------------------------------- */
/* RNNOneHot: cost = mean_b(-log softmax(h W + b)[Y_b] / pop_b) (+ bias reg), rnn_one_hot.py:65-77.
 * On a single rank the call returns as soon as the cost is on the host; the backward pass and the update may still
 * be running on the handle's stream (every later call on the handle is ordered behind them). */
SBR_API int sbr_train_step_cce(sbr_model* m, const int32_t* X, const float* mask, const int32_t* Y,
                       const float* pop, int B, float* cost);
/* RNNSampling: cells = [Y_all; samples], rnn_sampling.py:68-91,137 and sparse_lstm.py:41-54.
 * Y_all holds the targets of the WHOLE global batch (n_all of them); this rank's rows are
 * Y_all[row_offset : row_offset+B].  Single rank: Y_all = Y, n_all = B, row_offset = 0. */
SBR_API int sbr_train_step_sampled(sbr_model* m, const int32_t* X, const float* mask, const int32_t* Y_all,
                           int n_all, int row_offset, const int32_t* samples, int S,
                           const float* pop, int B, float* cost);
/* RNNMargin with the reference's dense inputs Ymat/weight [B, n_items], rnn_margin.py:100,109 */
SBR_API int sbr_train_step_margin_dense(sbr_model* m, const int32_t* X, const float* mask, const float* Ymat,
                                const float* weight, int B, float* cost);
/* RNNMargin, ragged form of the same inputs: the device rebuilds what rnn_margin.py:121-149
 * fills -- weight = w_neg[b] everywhere, -1 on the row's targets, 0 on the items of X (when
 * exclude_seen); Y = default_target (NULL = 0) everywhere, 1 on targets, 0 on seen items. */
SBR_API int sbr_train_step_margin(sbr_model* m, const int32_t* X, const float* mask,
                          const int32_t* target_offsets /* [B+1] */, const int32_t* target_ids,
                          const float* w_neg /* [B] */, const float* default_target /* [n_items] or NULL */,
                          int exclude_seen, int B, float* cost);

/* Device-side batch assembly (replaces the per-item python loop of rnn_one_hot.py:90-101 /
 * rnn_base.py:396-415): upload the training sequences ONCE as a CSR of ids ([total, ids_per_step] int32, offsets
 * [n_seqs+1]); a mini-batch is then B (sequence, start, length) triples -- row b reads
 * ids[offsets[seq_b] + start_b : + len_b] -- and the padded X / lengths are built on the device.  Same arithmetic and
 * same results as sbr_train_step_cce on the equivalent X / mask. */
SBR_API int sbr_dataset_upload(sbr_model* m, int n_seqs, const int32_t* offsets, const int32_t* ids);
SBR_API int sbr_train_step_cce_rows(sbr_model* m, const int32_t* seq, const int32_t* start, const int32_t* len,
                                    const int32_t* Y, const float* pop, int B, float* cost);

/* Device-resident batches (bench `value`, prefetch): stage a batch into slot s once, then step on
 * it any number of times with no host->device traffic.  cost may be NULL (no sync, no D2H). */
SBR_API int sbr_stage_cce(sbr_model* m, int slot, const int32_t* X, const float* mask, const int32_t* Y,
                  const float* pop, int B);
SBR_API int sbr_train_step_staged(sbr_model* m, int slot, float* cost);
SBR_API int sbr_synchronize(sbr_model* m, float* last_cost /* may be NULL */);

/* ---- predict_function / test_function --------------------------------------------------- */
/* scores[B, n_items]: softmax probabilities for CCE, raw linear scores otherwise; softmax != 0
 * forces a softmax (RNNSampling test function, rnn_sampling.py:143). */
SBR_API int sbr_scores(sbr_model* m, const int32_t* X, const float* mask, int B, int softmax, float* scores);
/* Fused exclude + sorted top-k on the device.  excl_* is a ragged list of ids per row (may be
 * NULL).  mode bit0: softmax first; bit1: 0 = excluded scores are multiplied by 0 (test_function,
 * rnn_base.py:201-202), 1 = set to -inf (top_k_recommendations, rnn_base.py:154-156).
 * ids_out [B, k], best first (np.argpartition(-out, range(k))[:k], rnn_base.py:159,207). */
SBR_API int sbr_topk(sbr_model* m, const int32_t* X, const float* mask, int B, const int32_t* excl_offsets,
             const int32_t* excl_ids, int k, int mode, int32_t* ids_out);

/* ---- measurement ------------------------------------------------------------------------ */
#define SBR_N_STAGES 9
SBR_API const char* sbr_stage_name(int i);            /* "h2d","gather","rnn_fwd","output","rnn_bwd","wgrad","scatter","allreduce","optimizer" */
SBR_API int sbr_set_profiling(sbr_model* m, int on);  /* record a cudaEvent pair around every stage */
SBR_API int sbr_stage_times(sbr_model* m, float ms[SBR_N_STAGES]);   /* of the last profiled step   */
SBR_API int64_t sbr_kernel_launches(const sbr_model* m);             /* kernels launched since create */
/* Host-only: how the tcgen05 scan launchers would tile a batch with these lengths on `slots` co-resident 8-CTA
 * clusters (rnn_tc.cu::plan_tiles): tile height of the main launch (8 / 16 rows), its tiles in launch order (longest
 * first, order64 has room for 64), and the 16-row group that runs as a second launch in the mixed tiling (-1: none).
 * `lens` may be NULL (static rule).  Touches no device; used by the CPU tests. */
SBR_API int sbr_plan_scan_tiles(const int32_t* lens, int B, int t_max, int slots, float ratio8,
                                int* tile_rows, int* n_tiles, int* extra16, unsigned char* order64);
/* Diagnostics: C[M,N] = alpha * op(A) * op(B) (+ bias[n]) (+ beta * C) on the device the handle lives on, host buffers
 * in and out (row-major; ta: A stored [K,lda]; tb: B stored [N,ldb]; beta in {0,1}; bias may be NULL).  engine 0 = the
 * fp32 FFMA kernels (gemm.cu), 1 = the tcgen05 3xTF32 kernel (tc_gemm.cu), which is what every GEMM-shaped stage of
 * the path runs on.  Used by the GPU tests to check the tensor-core kernel against float64 in isolation; `ms` (may be
 * NULL) receives the device time of `reps` back-to-back launches. */
SBR_API int sbr_debug_gemm(sbr_model* m, int engine, int ta, int tb, int M, int N, int K, const float* A, int lda,
                           const float* B, int ldb, float* C, int ldc, float alpha, float beta, const float* bias,
                           int reps, float* ms);
/* device-side stopwatch on the handle's stream (cudaEvent pair): start synchronises the stream
 * first, stop blocks until the stop event has completed and returns the elapsed milliseconds */
SBR_API int sbr_timer_start(sbr_model* m);
SBR_API int sbr_timer_stop(sbr_model* m, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* SBR_B200_H */
