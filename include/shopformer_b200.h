/*
 * shopformer_b200.h -- C ABI of the B200-native Shopformer scoring path.
 *
 * The reference (cthadeufaria/computer-vision-shoplifting-detection) has no FFI: its
 * "plugin interface" for this path is a set of Python methods on nn.Module / Dataset
 * objects.  Every entry point below names the reference method it replaces
 * (file:line relative to the reference checkout).  INTEGRATION.md shows the ctypes
 * stub a maintainer would add on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - `*_dev` pointers are device (HBM) pointers, `*_host` pointers are host pointers;
 *   - outputs are caller-allocated; nothing is allocated per call except inside the
 *     explicit `sf_runner_*` object; no entry point synchronises the device unless
 *     its comment says so;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *   - every function returns SF_OK (0) or a negative SF_E* code and never throws;
 *     `sf_last_error()` returns a thread-local message for the last failure;
 *   - a `sf_model` is immutable after creation and may be used concurrently from
 *     several host threads on distinct streams (each call brings its own workspace).
 *
 * Tensor layouts (all contiguous, fp32 unless stated)
 *   poses   (B, C, T, V)      x-plane then y-plane, each T x V  (reference __getitem__ layout)
 *   tokens  (B, S, D)         D = latent_channels * V, feature index c*V + v
 *   recon   (B, S, D)
 *   scores  (B)               or (B, S) for SF_REDUCE_NONE (variant 2 only)
 */
#ifndef SHOPFORMER_B200_H
#define SHOPFORMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SF_ABI_VERSION 2

enum {
  SF_OK = 0,
  SF_E_INVALID = -1,      /* bad argument / shape / config                              */
  SF_E_MISSING = -2,      /* a state-dict key the config requires was not supplied      */
  SF_E_SHAPE = -3,        /* a supplied tensor has the wrong number of elements         */
  SF_E_UNSUPPORTED = -4,  /* valid config that no kernel in this build covers           */
  SF_E_CUDA = -5,         /* CUDA runtime error (message has the cudaError string)      */
  SF_E_NODEVICE = -6      /* no sm_100 device visible: there is NO CPU fallback         */
};

enum { SF_VARIANT_SHOPFORMER = 1, SF_VARIANT_SHOPFORMER_2 = 2 };
enum { SF_REDUCE_MEAN = 0, SF_REDUCE_NONE = 1 };
/* arithmetic of the tokenizer/transformer contractions: fp32 on the CUDA cores, or 16-bit operands with fp32
 * accumulation on the tensor cores (SF_PREC_BF16 names the tensor-core path; its operand FORMAT is sf_config.tc_format) */
enum { SF_PREC_FP32 = 0, SF_PREC_BF16 = 1, SF_PREC_TC16 = 1 };
/* 16-bit operand format of the tensor-core path.  AUTO = fp16 (11 significant bits: 8x smaller rounding error than
 * bf16 on weights and activations; activations saturate at +-65504) whenever every packed weight is inside fp16's
 * range, else bf16.  BF16 forces the wide-range format (checkpoints whose activations can exceed 6e4). */
enum { SF_TC_AUTO = 0, SF_TC_BF16 = 1, SF_TC_FP16 = 2 };

#define SF_MAX_BLOCKS 8

/* Everything that is not recoverable from tensor shapes.  Mirrors the constructor
 * arguments of shopformer/models/shopformer.py:35-50 and the `model:` section of
 * shopformer_2/configs/paper_config.yaml:5-26. */
typedef struct sf_config {
  int32_t variant;                 /* SF_VARIANT_*                                        */
  int32_t in_channels;             /* 2                                                   */
  int32_t num_keypoints;           /* V: 17 or 18 (<= 32)                                 */
  int32_t n_blocks;                /* ST-GCN blocks in the tokenizer (4)                  */
  int32_t channels[SF_MAX_BLOCKS + 1]; /* [in, hidden, hidden, hidden, latent]            */
  int32_t strides[SF_MAX_BLOCKS];  /* temporal stride of each block                       */
  int32_t pool_tokens;             /* variant 2 AdaptiveAvgPool target, 0 = no pooling    */
  int32_t d_model;                 /* transformer width (136 / 144)                       */
  int32_t n_heads;
  int32_t n_enc_layers;
  int32_t n_dec_layers;
  int32_t d_ff;
  int32_t tc_format;               /* SF_TC_*                                             */
  int32_t reserved[7];
} sf_config;

typedef struct sf_model sf_model;    /* packed, BatchNorm-folded weights resident in HBM */
typedef struct sf_runner sf_runner;  /* pinned staging + device buffer ring + 2 streams    */

/* ------------------------------------------------------------------ library ------- */
int sf_abi_version(void);
const char* sf_last_error(void);
/* Number of visible sm_100 devices; SF_E_NODEVICE if none. */
int sf_device_count(void);

/* ------------------------------------------------------------------ model --------- */
/* Replaces: Shopformer.load_state_dict + .to(device) + .eval()
 *   (shopformer/evaluate.py:75-78, shopformer/inference.py:59-62,
 *    shopformer_2/train.py:512-521).
 * `names[i]` are the reference's state-dict keys, `data_host[i]` fp32 host arrays of
 * `numel[i]` elements.  BatchNorm running statistics are folded into the adjacent
 * conv / affine here, once; the result is immutable.  Synchronises `device`. */
int sf_model_create(const sf_config* cfg, int32_t n_tensors, const char* const* names,
                    const float* const* data_host, const int64_t* numel, int32_t device,
                    sf_model** out);
void sf_model_destroy(sf_model* m);
/* Token count S and token width D for windows of T frames (conv length (T-1)/s+1 per block). */
int sf_model_token_shape(const sf_model* m, int32_t T, int32_t* S, int32_t* D);
/* Operand format the tensor-core kernels use for windows of T frames: *tokenizer_f16 / *transformer_f16 = 1 for
 * fp16, 0 for bf16 (the one-window-per-pass tokenizer that serves hidden-64 / pooled shapes is bf16 only). */
int sf_model_tc_formats(const sf_model* m, int32_t T, int32_t* tokenizer_f16, int32_t* transformer_f16);
/* Bytes of device workspace the calls below need for a batch of B windows of T frames.
 * Bounded: batches beyond 131,072 windows are processed in internal passes of that size. */
int64_t sf_workspace_bytes(const sf_model* m, int64_t B, int32_t T);

/* ------------------------------------------------------------------ hot path ------ */
/* Replaces: Shopformer.tokenize -> GCAEEncoder.forward
 *   (shopformer/models/shopformer.py:125-136, shopformer/models/gcae.py:331-366;
 *    shopformer_2/models/gcae.py:375-422). */
int sf_tokenize(const sf_model* m, const float* poses_dev, int64_t B, int32_t T, int32_t precision,
                float* tokens_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);

/* Replaces: Shopformer.reconstruct_tokens -> ShopformerTransformer.forward
 *   (shopformer/models/shopformer.py:138-148, shopformer/models/transformer.py:304-329;
 *    shopformer_2/models/transformer.py:147-194). */
int sf_reconstruct_tokens(const sf_model* m, const float* tokens_dev, int64_t B, int32_t S, int32_t precision,
                          float* recon_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);

/* Replaces: Shopformer.compute_normality_score (shopformer/models/shopformer.py:150-178)
 *   and the MSE of compute_anomaly_score (shopformer_2/models/shopformer.py:178-186).
 * Stand-alone HBM-bound kernel: reads tokens + recon, writes B (or B*S) floats. */
int sf_normality_score(const sf_model* m, const float* tokens_dev, const float* recon_dev, int64_t B,
                       int32_t S, int32_t reduction, float* scores_dev, void* stream);

/* Replaces: Shopformer.forward()['normality_score'] (shopformer/models/shopformer.py:180-220)
 *   and Shopformer.compute_anomaly_score (shopformer_2/models/shopformer.py:155-188).
 * poses -> scores in one call; `tokens_dev` / `recon_dev` may be NULL (then neither ever
 * reaches HBM as fp32 tensors).  `precision` is SF_PREC_*. */
int sf_score_windows(const sf_model* m, const float* poses_dev, int64_t B, int32_t T, int32_t reduction,
                     int32_t precision, float* scores_dev, float* tokens_dev, float* recon_dev,
                     void* workspace_dev, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ windowing ----- */
/* Packed tracks (one entry per detection of one person, grouped per person in the
 * reference's first-seen order, frames ascending inside a track):
 *   kp_dev         (F, K, 3) fp32   x, y, conf as PoseLift stores them, K >= 17
 *                  ((F, K, 2) x, y only when kp_channels == 2: ingest may drop the confidence, a third fewer bytes)
 *   frame_no_dev   (F) int32
 *   track_offsets  (n_tracks + 1) int64, host
 *   track_video    (n_tracks) int32, host        index into gt_offsets
 *   gt_dev         concatenated per-video uint8 frame labels (may be NULL -> label 0)
 *   gt_offsets     (n_videos + 1) int64, host
 */
typedef struct sf_tracks {
  const float* kp_dev;
  const int32_t* frame_no_dev;
  const int64_t* track_offsets_host;
  const int32_t* track_video_host;
  const uint8_t* gt_dev;
  const int64_t* gt_offsets_host;
  int64_t n_frames;
  int32_t n_tracks;
  int32_t n_videos;
  int32_t kp_per_frame;            /* K (17 for PoseLift)                                 */
  int32_t kp_channels;             /* floats per keypoint in kp_dev: 3 (or 0) = x, y, conf; 2 = x, y */
} sf_tracks;

typedef struct sf_window_params {
  int32_t seq_len;                 /* T                                                   */
  int32_t stride;
  int32_t max_gap;                 /* 5 in shopformer/, ctor arg in shopformer_2/         */
  int32_t num_keypoints;           /* V keypoints per frame in the output                 */
  int32_t normalize;               /* centre / scale normalisation on/off                 */
  int32_t add_neck;                /* 1: keypoint V-1 = shoulder midpoint (shopformer_2 add_neck_keypoint,
                                      data/poselift_dataset.py:57-91); 0: keypoints beyond the detection's are zero
                                      rows (shopformer/data/poselift_dataset.py:352-354)                            */
  int32_t include_confidence;      /* 1: third output plane = the detections' confidence (shopformer/ ctor flag,
                                      data/poselift_dataset.py:216-225,349); poses_dev is then (cap, 3, T, V)        */
  int32_t reserved[1];
} sf_window_params;

/* Upper bound on the number of windows (every candidate start position), host-only. */
int64_t sf_window_capacity(const sf_tracks* tr, const sf_window_params* p);
int64_t sf_window_workspace_bytes(const sf_tracks* tr, const sf_window_params* p);

/* Replaces: PoseLiftDataset._extract_sequences + _check_continuity +
 *   _extract_pose_sequence + _normalize_sequence + __getitem__ transpose
 *   (shopformer/data/poselift_dataset.py:256-400;
 *    shopformer_2/data/poselift_dataset.py:57-91,410-589).
 * Outputs sized by sf_window_capacity(); valid windows are compacted to the front in
 * the reference's order.  `n_windows_dev` (int64, device) receives the count; if
 * `n_windows_host` is non-NULL the call synchronises the stream and stores it there.
 *   poses_dev        (cap, 2, T, V) fp32   ((cap, 3, T, V) with include_confidence)
 *   labels_dev       (cap) int32           majority vote of GT[min(f, len-1)]
 *   window_track_dev (cap) int32           which track each window came from
 *   window_start_dev (cap) int32           start position inside the track
 *   frame_idx_dev    (cap, T) int32 or NULL (shopformer_2 `frame_indices`)
 */
int sf_window_normalize(const sf_tracks* tr, const sf_window_params* p, float* poses_dev,
                        int32_t* labels_dev, int32_t* window_track_dev, int32_t* window_start_dev,
                        int32_t* frame_idx_dev, int64_t* n_windows_dev, int64_t* n_windows_host,
                        void* workspace_dev, int64_t workspace_bytes, void* stream);

/* Replaces: _extract_pose_sequence + _normalize_sequence + the (T,V,C)->(C,T,V) transpose for windows that
 *   are ALREADY cut (shopformer/data/poselift_dataset.py:331-400): raw_dev is (B, T, K, 3) AoS keypoints,
 *   poses_dev the model's (B, 2, T, V) layout (V = 18 adds the synthetic neck).  No host work and no
 *   synchronisation, so a streaming tick can be captured in a CUDA graph. */
int sf_normalize_windows(const float* raw_dev, int64_t B, int32_t T, int32_t K, int32_t V, int32_t normalize,
                         float* poses_dev, void* stream);

/* Replaces the reference flow "dataset construction feeding the scoring loop":
 *   PoseLiftDataset.__init__ -> _extract_sequences / _normalize_sequence / __getitem__
 *   (shopformer/data/poselift_dataset.py:256-400) -> evaluate_model's loop (shopformer/evaluate.py:83-104;
 *   shopformer_2/evaluate.py:36-63).  Packed tracks in, per-window scores and the window index out, in the
 *   reference's window order.  The normalised windows never exist as a whole: they are cut, normalised and scored in
 *   passes of 131,072 windows through the workspace.  Outputs are sized by sf_window_capacity().
 *   Synchronises `stream` once, at the end (tensor-core kernels: every pass is launched for its capacity and clamps to the
 *   device-side count) or after the windowing kernels (fp32 kernels: the count sizes the launches); the count is returned in
 *   `n_windows_host` (required).  `labels_dev` / `window_track_dev` / `window_start_dev` are required (the index the
 *   gather reads). */
int64_t sf_score_from_tracks_workspace_bytes(const sf_model* m, const sf_tracks* tr, const sf_window_params* p);
int sf_score_from_tracks(const sf_model* m, const sf_tracks* tr, const sf_window_params* p, int32_t precision,
                         float* scores_dev, int32_t* labels_dev, int32_t* window_track_dev,
                         int32_t* window_start_dev, int64_t* n_windows_host, void* workspace_dev,
                         int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ pose decoder (optional output) */
/* Replaces: GCAEDecoder.forward in eval mode (shopformer/models/gcae.py:369-478; shopformer_2/models/gcae.py:425-534):
 *   Linear latent*V -> hidden*V per token, n_layers x (ConvTranspose2d k=(u,1) stride (u,1) | Conv2d 1x1) with BatchNorm2d +
 *   ReLU after all but the last, bilinear resize (align_corners=False) to seq_len when the stack's length differs.
 * Not on the scoring path: `Shopformer.forward` returns `gcae_reconstructed`, the facade produces it on demand with this.
 * `names` are the decoder's own state-dict keys ("initial_proj.weight", "layers.0.weight", "layers.1.running_mean", ...);
 * `upsample[i]` is the temporal factor of layer i (1 = 1x1 conv).  BatchNorm is folded here.  Synchronises `device`. */
typedef struct sf_decoder sf_decoder;
int sf_decoder_create(int32_t latent_channels, int32_t hidden_channels, int32_t out_channels, int32_t num_keypoints,
                      int32_t seq_len, int32_t n_layers, const int32_t* upsample, int32_t n_tensors,
                      const char* const* names, const float* const* data_host, const int64_t* numel, int32_t device,
                      sf_decoder** out);
void sf_decoder_destroy(sf_decoder* d);
int64_t sf_decoder_workspace_bytes(const sf_decoder* d, int64_t B, int32_t S);
/* tokens_dev (B, S, latent*V) -> poses_dev (B, out_channels, seq_len, V).  SF_E_UNSUPPORTED when a window's activations do
 * not fit shared memory.  No synchronisation. */
int sf_decode_poses(const sf_decoder* d, const float* tokens_dev, int64_t B, int32_t S, float* poses_dev,
                    void* workspace_dev, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ after the path: aggregation + ranking metrics */
/* Replaces: the per-video grouping of evaluate_video_level (shopformer_2/evaluate.py:65-118) and
 *   compute_video_level_metrics' aggregation (shopformer_2/utils/metrics.py:148-188): per-video max / mean / 95th
 *   percentile of the window scores, in float64 as numpy computes them from the python floats, and the video label =
 *   the label of the video's LAST window in dataset order (what the reference's `video_labels[video_id] = info['label']`
 *   leaves).  `video_id_dev` (n) int32 in [0, n_videos) (other ids are ignored); windows of a video need not be
 *   contiguous.  The percentile is numpy's default (linear interpolation).
 *   Outputs (n_videos): agg_max/agg_mean/agg_p95 float64 (NaN for a video without windows), video_label int32 (0 for a
 *   video without windows), count int32.  Any output may be NULL.  No synchronisation. */
int64_t sf_video_aggregate_workspace_bytes(int64_t n, int32_t n_videos);
int sf_video_aggregate(const float* scores_dev, const int32_t* video_id_dev, const int32_t* labels_dev, int64_t n,
                       int32_t n_videos, double* agg_max_dev, double* agg_mean_dev, double* agg_p95_dev,
                       int32_t* video_label_dev, int32_t* count_dev, void* workspace_dev, int64_t workspace_bytes,
                       void* stream);

/* Replaces: compute_auc_roc / compute_auc_pr (shopformer/utils/metrics.py:18-77; shopformer_2/utils/metrics.py:21-145),
 *   i.e. sklearn.metrics.roc_auc_score and average_precision_score, on the device: radix sort of the scores (descending),
 *   tie groups, prefix sums of positives, trapezoid / step integration in float64, and compute_metrics' thresholding:
 *   `threshold` NaN selects the Youden-J optimum of the ROC curve (first maximum of tpr - fpr in descending-score order,
 *   +inf if no point beats the origin), predictions = score >= threshold.
 *   out_host (8 doubles): auc_roc, auc_pr, threshold used, n_positive, tp, fp, tn, fn; auc_roc is NaN when only one class
 *   is present and auc_pr when there is no positive (the reference returns 0.5 / 0.0 there: the facade applies that
 *   convention).  Synchronises `stream`. */
int64_t sf_ranking_metrics_workspace_bytes(int64_t n);
int sf_ranking_metrics(const float* scores_dev, const int32_t* labels_dev, int64_t n, float threshold,
                       double* out_host, void* workspace_dev, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ host-buffer runner */
/* The call a reference-side loop makes per batch: poses on the HOST, scores back on the
 * HOST (shopformer/evaluate.py:90-99, shopformer/inference.py:67-94,
 * shopformer/train.py:311-319, shopformer_2/evaluate.py:52-58).  The runner owns pinned
 * staging, a ring of 4 device buffers, a copy stream and a compute stream, and pipelines
 * H2D / kernels / D2H in chunks of at most `max_chunk` windows (rounded down to whole
 * waves of the transformer grid when `max_chunk` allows). */
int sf_runner_create(const sf_model* m, int32_t T, int64_t max_chunk, sf_runner** out);
void sf_runner_destroy(sf_runner* r);
/* Blocking: returns after `scores_host[0..B)` is written.  `poses_host` need not be pinned: a pageable
 * buffer is staged through the runner's pinned ring; a page-locked one (cudaHostAlloc / cudaHostRegister /
 * torch pinned memory) is read in place by the kernels over PCIe, with no staging copy. */
int sf_runner_score(sf_runner* r, const float* poses_host, int64_t B, int32_t precision,
                    float* scores_host);
/* Host tracks in, host scores + window index out: the reference flow of sf_score_from_tracks for a caller whose
 * tracks live in host memory (what PoseLiftDataset reads from the pickle files).  Tracks are uploaded in groups of
 * whole tracks on the copy stream while the previous group is windowed and scored on the compute stream; uploading
 * the raw detections moves stride*K*kp_channels*4 bytes per window instead of the 2*T*V*4 of a pre-cut window.
 * `tr->kp_dev`, `frame_no_dev` and `gt_dev` are HOST pointers here (pinned or pageable); outputs are host arrays
 * sized by sf_window_capacity().  Blocking. */
int sf_runner_score_tracks(sf_runner* r, const sf_tracks* tr_host, const sf_window_params* p, int32_t precision,
                           float* scores_host, int32_t* labels_host, int32_t* window_track_host,
                           int32_t* window_start_host, int64_t* n_windows_host);

/* Pinned staging buffer of the runner (capacity one chunk of windows, slots 0..3) so that
 * producers can write windows straight into page-locked memory. */
float* sf_runner_pinned_poses(sf_runner* r, int32_t slot);

/* ------------------------------------------------------------------ diagnostics ---- */
/* One 128 x N x K bf16 tcgen05 product (fp32 accumulate in TMEM) on a single CTA, host in / host out.
 * Exists so that the tests can pin the shared-memory descriptor encodings the tensor-core kernels
 * rely on.  mode 0: B is (N,K) row-major (K-major operand); mode 1: B is (K,N) row-major (MN-major
 * operand); mode | 12: both operands are rounded to fp16 instead of bf16 (kind::f16 takes A and B in the same format).
 * A is (128+shift, K) row-major and the product uses rows [shift, shift+128).
 * Synchronises the device.  No reference counterpart. */
int sf_selftest_umma(int32_t mode, int32_t N, int32_t K, int32_t shift, const float* a_host,
                     const float* b_host, float* d_host);

#ifdef __cplusplus
}
#endif
#endif /* SHOPFORMER_B200_H */
