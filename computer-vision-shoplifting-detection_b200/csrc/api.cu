// C-ABI entry points of the scoring hot path + the host-buffer runner.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include <mutex>

#include "sf_internal.h"

using namespace sf;

#include <atomic>
namespace sf {
static std::atomic<long long> g_launches[LK_COUNT];
void count_launch(int kind) { g_launches[kind].fetch_add(1, std::memory_order_relaxed); }
}  // namespace sf
// Launches so far of [tokenizer v2, one-window tensor-core tokenizer, fp32 tokenizer, tensor-core transformer, fp32 transformer]
// (host-side counters; debugging / test aid, not part of the C ABI)
extern "C" int sfdbg_launch_counts(long long* out, int n) {
  for (int i = 0; i < n && i < LK_COUNT; ++i) out[i] = g_launches[i].load(std::memory_order_relaxed);
  return LK_COUNT;
}

namespace {
inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

int check_model(const sf_model* m) {
  SF_REQUIRE(m, SF_E_INVALID, "null model");
  SF_REQUIRE(m->device >= 0, SF_E_INVALID, "host-only model (emulator fixture): no entry point computes with it");
  return SF_OK;
}
int check_T(const sf_model* m, int T) {
  int rc = check_model(m);
  if (rc) return rc;
  SF_REQUIRE(T >= 1 && T <= 4096, SF_E_INVALID, "T=%d outside [1,4096]", T);
  return SF_OK;
}
// tensor-core tokenizer: the multi-window tile kernel (tokenizer2_bf16.cu) when it covers the shape, otherwise the
// one-window-per-pass kernel (tokenizer_bf16.cu)
int launch_tokenizer_tc(const sf_model* m, const float* poses, int64_t B, int T, float* tokens, cudaStream_t st, DevCount cnt = DevCount()) {
  if (tokenizer2_supported(m, T)) return launch_tokenizer2(m, poses, B, T, tokens, st, cnt);
  SF_REQUIRE(!cnt.n, SF_E_UNSUPPORTED, "device-side batch size needs tokenizer v2");
  return launch_tokenizer_bf16(m, poses, B, T, tokens, st);
}
// kernels that accept a device-side batch size (DevCount): tokenizer v2 + the tensor-core transformer
bool supports_dev_count(const sf_model* m, int T, int precision) {
  return precision == SF_PREC_BF16 && tokenizer2_supported(m, T) && transformer_bf16_supported(m, token_len(m, T));
}
}  // namespace

// Windows per internal pass of sf_score_windows when the caller does not ask for the tokens: bounds the workspace of a
// 10 M-window sweep to what 131,072 windows need, and a pass's tokens (<= 214 MB) largely stay in the 126 MB L2 between
// the two kernels.  A multiple of the transformer wave (sm_count x tile) would be marginally better; it only matters
// for B > 131,072.
constexpr int64_t kScoreChunk = 131072;

extern "C" int64_t sf_workspace_bytes(const sf_model* m, int64_t B, int32_t T) {
  if (!m || m->device < 0 || B < 0 || T < 1) return SF_E_INVALID;
  B = std::min(B, kScoreChunk);
  const int S = token_len(m, T);
  // tokens staged between the two kernels when the caller does not ask for them
  return align256(B * (int64_t)S * m->xf.d_tok * (int64_t)sizeof(float)) + align256(tokenizer_fp32_workspace(m, B, T));
}

extern "C" int sf_model_tc_formats(const sf_model* m, int32_t T, int32_t* tokenizer_f16, int32_t* transformer_f16) {
  SF_REQUIRE(m && T >= 1, SF_E_INVALID, "sf_model_tc_formats: bad argument");
  // tokenizer v2 follows the model's format; the one-window-per-pass kernel (hidden-64 / pooled shapes) is bf16 only
  if (tokenizer_f16) *tokenizer_f16 = (m->device >= 0 && tokenizer2_supported(m, T) && tokenizer2_f16(m)) ? 1 : 0;
  if (transformer_f16) *transformer_f16 = m->xfprog.f16 ? 1 : 0;
  return SF_OK;
}

extern "C" int sf_tokenize(const sf_model* m, const float* poses_dev, int64_t B, int32_t T, int32_t precision,
                           float* tokens_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
  int rc = check_T(m, T);
  if (rc) return rc;
  SF_REQUIRE(B >= 0 && (B == 0 || (poses_dev && tokens_dev)), SF_E_INVALID, "sf_tokenize: null buffer");
  SF_REQUIRE(precision == SF_PREC_FP32 || precision == SF_PREC_BF16, SF_E_INVALID, "unknown precision %d", precision);
  DeviceGuard guard;
  SF_CUDA_OK(guard.enter(m->device));
  if (precision == SF_PREC_BF16) return launch_tokenizer_tc(m, poses_dev, B, T, tokens_dev, (cudaStream_t)stream);
  return launch_tokenizer_fp32(m, poses_dev, B, T, tokens_dev, workspace_dev, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int sf_reconstruct_tokens(const sf_model* m, const float* tokens_dev, int64_t B, int32_t S, int32_t precision,
                                     float* recon_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
  (void)workspace_dev;
  (void)workspace_bytes;
  int rc = check_model(m);
  if (rc) return rc;
  SF_REQUIRE(B >= 0 && (B == 0 || (tokens_dev && recon_dev)), SF_E_INVALID, "sf_reconstruct_tokens: null buffer");
  SF_REQUIRE(S >= 1 && S <= 100, SF_E_INVALID, "S=%d outside [1,100]", S);
  SF_REQUIRE(precision == SF_PREC_FP32 || precision == SF_PREC_BF16, SF_E_INVALID, "unknown precision %d", precision);
  DeviceGuard guard;
  SF_CUDA_OK(guard.enter(m->device));
  if (precision == SF_PREC_BF16)
    return launch_transformer_bf16(m, tokens_dev, B, S, SF_REDUCE_MEAN, recon_dev, nullptr, (cudaStream_t)stream);
  return launch_transformer_fp32(m, tokens_dev, B, S, SF_REDUCE_MEAN, recon_dev, nullptr, (cudaStream_t)stream);
}

extern "C" int sf_normality_score(const sf_model* m, const float* tokens_dev, const float* recon_dev, int64_t B, int32_t S,
                                  int32_t reduction, float* scores_dev, void* stream) {
  int rc = check_model(m);
  if (rc) return rc;
  SF_REQUIRE(B >= 0 && (B == 0 || (tokens_dev && recon_dev && scores_dev)), SF_E_INVALID, "sf_normality_score: null buffer");
  SF_REQUIRE(S >= 1 && S <= 100, SF_E_INVALID, "S=%d outside [1,100]", S);
  DeviceGuard guard;
  SF_CUDA_OK(guard.enter(m->device));
  return launch_score(m, tokens_dev, recon_dev, B, S, reduction, scores_dev, (cudaStream_t)stream);
}

static int score_windows_impl(const sf_model* m, const float* poses_dev, int64_t B, int32_t T, int32_t reduction,
                              int32_t precision, float* scores_dev, float* tokens_dev, float* recon_dev,
                              void* workspace_dev, int64_t workspace_bytes, void* stream, DevCount cnt);

extern "C" int sf_score_windows(const sf_model* m, const float* poses_dev, int64_t B, int32_t T, int32_t reduction,
                                int32_t precision, float* scores_dev, float* tokens_dev, float* recon_dev,
                                void* workspace_dev, int64_t workspace_bytes, void* stream) {
  return score_windows_impl(m, poses_dev, B, T, reduction, precision, scores_dev, tokens_dev, recon_dev, workspace_dev, workspace_bytes,
                            stream, DevCount());
}

// `cnt`: optional device-side count of valid windows (B is then an upper bound; see DevCount)
static int score_windows_impl(const sf_model* m, const float* poses_dev, int64_t B, int32_t T, int32_t reduction,
                              int32_t precision, float* scores_dev, float* tokens_dev, float* recon_dev,
                              void* workspace_dev, int64_t workspace_bytes, void* stream, DevCount cnt) {
  int rc = check_T(m, T);
  if (rc) return rc;
  SF_REQUIRE(B >= 0 && (B == 0 || (poses_dev && scores_dev)), SF_E_INVALID, "sf_score_windows: null buffer");
  SF_REQUIRE(precision == SF_PREC_FP32 || precision == SF_PREC_BF16, SF_E_INVALID, "unknown precision %d", precision);
  if (B == 0) return SF_OK;
  DeviceGuard guard;
  SF_CUDA_OK(guard.enter(m->device));
  const int S = token_len(m, T);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t pose_elems = (int64_t)m->cfg.in_channels * T * m->cfg.num_keypoints, tok_elems = (int64_t)S * m->xf.d_tok;
  // with caller-provided token storage the whole batch goes in one pass; otherwise in passes of kScoreChunk windows
  const int64_t pass = tokens_dev ? B : std::min(B, kScoreChunk);
  const int64_t tok_bytes = align256(pass * tok_elems * (int64_t)sizeof(float));
  char* ws = (char*)workspace_dev;
  int64_t ws_left = workspace_bytes;
  float* tok_ws = nullptr;
  if (!tokens_dev) {
    SF_REQUIRE(ws && ws_left >= tok_bytes, SF_E_INVALID, "sf_score_windows: workspace too small (%lld < %lld)",
               (long long)workspace_bytes, (long long)sf_workspace_bytes(m, B, T));
    tok_ws = (float*)ws;
    ws += tok_bytes;
    ws_left -= tok_bytes;
  }
  const int64_t score_stride = reduction == SF_REDUCE_NONE ? S : 1;
  const bool xf_tc = precision == SF_PREC_BF16 && transformer_bf16_supported(m, S);
  SF_REQUIRE(!cnt.n || (precision == SF_PREC_BF16 && xf_tc), SF_E_UNSUPPORTED, "device-side batch size needs the tensor-core kernels");
  // one range of a pass on one stream: tokenizer, then transformer + score.  The two kernels only share the fp32 tokens in
  // HBM, so each picks its own path: a shape the tensor-core transformer does not cover (head width not a multiple of 4,
  // S > 4, d_model > 160) still gets the tensor-core tokenizer.
  auto tokenize = [&](int64_t off, int64_t n, float* tok, cudaStream_t s2) -> int {
    const DevCount c{cnt.n, cnt.off + off};
    return precision == SF_PREC_BF16 ? launch_tokenizer_tc(m, poses_dev + off * pose_elems, n, T, tok, s2, c)
                                     : launch_tokenizer_fp32(m, poses_dev + off * pose_elems, n, T, tok, ws, ws_left, s2);
  };
  auto reconstruct = [&](int64_t off, int64_t n, const float* tok, cudaStream_t s2) -> int {
    const DevCount c{cnt.n, cnt.off + off};
    float* rec = recon_dev ? recon_dev + off * tok_elems : nullptr;
    float* sc = scores_dev + off * score_stride;
    return xf_tc ? launch_transformer_bf16(m, tok, n, S, reduction, rec, sc, s2, c) : launch_transformer_fp32(m, tok, n, S, reduction, rec, sc, s2);
  };
  // Both tensor-core kernels are persistent (one CTA per SM walking equal tiles), so each ends in a partial wave: 65,536
  // config-A windows are 11.07 waves of transformer tiles = 12 tile times.  A large pass is therefore cut in two halves that
  // run on two streams: as the CTAs of one half's kernel retire, the SMs pick up the other half's CTAs, and only the
  // last kernel's tail stays exposed.  (Not under stream capture, not for small batches; SF_SPLIT_STREAMS=0 disables it.)
  static const bool split_on = !(getenv("SF_SPLIT_STREAMS") && atoi(getenv("SF_SPLIT_STREAMS")) == 0);
  const int64_t unit = 280 * 4;                         // whole tokenizer (7-window) and transformer (40-window) tiles
  for (int64_t off = 0; off < B; off += pass) {
    const int64_t n = std::min(pass, B - off);
    float* tok = tokens_dev ? tokens_dev + off * tok_elems : tok_ws;
    bool split = split_on && m->side && precision == SF_PREC_BF16 && xf_tc && n >= (int64_t)m->sm_count * 160;
    if (split) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
        cudaGetLastError();
        split = false;
      }
    }
    if (!split) {
      rc = tokenize(off, n, tok, st);
      if (rc) return rc;
      rc = reconstruct(off, n, tok, st);
      if (rc) return rc;
      continue;
    }
    const int64_t h1 = (n / 2 + unit - 1) / unit * unit, h2 = n - h1;
    sf::SideStream* sd = m->side;
    std::lock_guard<std::mutex> lock(sd->mu);
    SF_CUDA_OK(cudaEventRecord(sd->fork, st));
    SF_CUDA_OK(cudaStreamWaitEvent(sd->st, sd->fork, 0));
    rc = tokenize(off, h1, tok, st);
    if (!rc) rc = tokenize(off + h1, h2, tok + h1 * tok_elems, sd->st);
    if (!rc) rc = reconstruct(off, h1, tok, st);
    if (!rc) rc = reconstruct(off + h1, h2, tok + h1 * tok_elems, sd->st);
    // always join, so that the caller's stream never runs ahead of work enqueued on the side stream
    cudaError_t e = cudaEventRecord(sd->join, sd->st);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, sd->join, 0);
    if (rc) return rc;
    SF_CUDA_OK(e);
  }
  return SF_OK;
}

// ------------------------------------------------------------------------------------ tracks -> scores
namespace {
struct TrackScoreLayout {
  int64_t win_ws, count, poses, score_ws, total, pass, cap;
};
TrackScoreLayout track_score_layout(const sf_model* m, const sf_tracks* tr, const sf_window_params* p) {
  TrackScoreLayout L{};
  L.cap = window_candidates(tr, p);
  L.pass = std::max<int64_t>(1, std::min<int64_t>(L.cap, kScoreChunk));
  const int64_t per_w = (int64_t)m->cfg.in_channels * p->seq_len * m->cfg.num_keypoints;
  int64_t off = 0;
  L.win_ws = off; off += align256(sf_window_workspace_bytes(tr, p));
  L.count = off; off += 256;
  L.poses = off; off += align256(L.pass * per_w * (int64_t)sizeof(float));
  L.score_ws = off; off += align256(sf_workspace_bytes(m, L.pass, p->seq_len));
  L.total = off;
  return L;
}
int check_tracks_vs_model(const sf_model* m, const sf_tracks* tr, const sf_window_params* p) {
  int rc = check_model(m);
  if (rc) return rc;
  SF_REQUIRE(tr && p, SF_E_INVALID, "null tracks / window parameters");
  SF_REQUIRE(p->num_keypoints == m->cfg.num_keypoints, SF_E_INVALID, "window keypoints %d != model keypoints %d", p->num_keypoints,
             m->cfg.num_keypoints);
  SF_REQUIRE((p->include_confidence ? 3 : 2) == m->cfg.in_channels, SF_E_INVALID, "window channels do not match the model's in_channels=%d",
             m->cfg.in_channels);
  return check_T(m, p->seq_len);
}
}  // namespace

extern "C" int64_t sf_score_from_tracks_workspace_bytes(const sf_model* m, const sf_tracks* tr, const sf_window_params* p) {
  if (check_tracks_vs_model(m, tr, p) != SF_OK || sf_window_workspace_bytes(tr, p) < 0) return SF_E_INVALID;
  return track_score_layout(m, tr, p).total;
}

// Sync-free core (kernels that take the window count from device memory): every pass is launched for its capacity and
// clamps to the count on the device; the count stays in the workspace (`*n_windows_dev_out` points at it).
static int score_from_tracks_async(const sf_model* m, const sf_tracks* tr, const sf_window_params* p, int32_t precision,
                                   float* scores_dev, int32_t* labels_dev, int32_t* window_track_dev, int32_t* window_start_dev,
                                   const int64_t** n_windows_dev_out, void* workspace_dev, int64_t workspace_bytes, cudaStream_t st,
                                   bool tables_resident = false) {
  const TrackScoreLayout L = track_score_layout(m, tr, p);
  char* ws = (char*)workspace_dev;
  int64_t* n_dev = (int64_t*)(ws + L.count);
  *n_windows_dev_out = n_dev;
  int rc = window_index(tr, p, labels_dev, window_track_dev, window_start_dev, n_dev, ws + L.win_ws, L.poses - L.win_ws, st, tables_resident);
  if (rc) return rc;
  float* poses = (float*)(ws + L.poses);
  for (int64_t off = 0; off < L.cap; off += L.pass) {
    const int64_t cnt = std::min(L.pass, L.cap - off);
    rc = window_gather(tr, p, window_track_dev, window_start_dev, n_dev, off, cnt, poses, nullptr, ws + L.win_ws, st);
    if (rc) return rc;
    rc = score_windows_impl(m, poses, cnt, p->seq_len, SF_REDUCE_MEAN, precision, scores_dev + off, nullptr, nullptr, ws + L.score_ws,
                            L.total - L.score_ws, st, DevCount{n_dev, off});
    if (rc) return rc;
  }
  return SF_OK;
}

extern "C" int sf_score_from_tracks(const sf_model* m, const sf_tracks* tr, const sf_window_params* p, int32_t precision,
                                    float* scores_dev, int32_t* labels_dev, int32_t* window_track_dev, int32_t* window_start_dev,
                                    int64_t* n_windows_host, void* workspace_dev, int64_t workspace_bytes, void* stream) {
  int rc = check_tracks_vs_model(m, tr, p);
  if (rc) return rc;
  SF_REQUIRE(n_windows_host, SF_E_INVALID, "sf_score_from_tracks: n_windows_host is required");
  SF_REQUIRE(precision == SF_PREC_FP32 || precision == SF_PREC_BF16, SF_E_INVALID, "unknown precision %d", precision);
  SF_REQUIRE(sf_window_workspace_bytes(tr, p) >= 0, SF_E_INVALID, "bad tracks / window parameters");
  const TrackScoreLayout L = track_score_layout(m, tr, p);
  *n_windows_host = 0;
  if (L.cap == 0) return SF_OK;
  SF_REQUIRE(scores_dev && labels_dev && window_track_dev && window_start_dev, SF_E_INVALID, "sf_score_from_tracks: null output");
  SF_REQUIRE(workspace_dev && workspace_bytes >= L.total, SF_E_INVALID, "sf_score_from_tracks: workspace of %lld bytes needed, got %lld",
             (long long)L.total, (long long)workspace_bytes);
  DeviceGuard guard;
  SF_CUDA_OK(guard.enter(m->device));
  int dev = 0;
  rc = window_device_of(tr, &dev);
  if (rc) return rc;
  SF_REQUIRE(dev == m->device, SF_E_INVALID, "tracks live on device %d, the model on device %d", dev, m->device);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace_dev;
  int64_t* n_dev = (int64_t*)(ws + L.count);
  if (supports_dev_count(m, p->seq_len, precision)) {
    // everything is enqueued without knowing the count; it is read back once, behind the last kernel
    const int64_t* n_out = nullptr;
    rc = score_from_tracks_async(m, tr, p, precision, scores_dev, labels_dev, window_track_dev, window_start_dev, &n_out, workspace_dev,
                                 workspace_bytes, st);
    if (rc) return rc;
    int64_t n = 0;
    SF_CUDA_OK(cudaMemcpyAsync(&n, n_out, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SF_CUDA_OK(cudaStreamSynchronize(st));
    *n_windows_host = n;
    return SF_OK;
  }
  rc = window_index(tr, p, labels_dev, window_track_dev, window_start_dev, n_dev, ws + L.win_ws, L.poses - L.win_ws, st);
  if (rc) return rc;
  // the first pass does not need the count: its gather clamps on the device, and the count arrives while it runs
  float* poses = (float*)(ws + L.poses);
  rc = window_gather(tr, p, window_track_dev, window_start_dev, n_dev, 0, L.pass, poses, nullptr, ws + L.win_ws, st);
  if (rc) return rc;
  int64_t n = 0;
  SF_CUDA_OK(cudaMemcpyAsync(&n, n_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  SF_CUDA_OK(cudaStreamSynchronize(st));
  *n_windows_host = n;
  for (int64_t off = 0; off < n; off += L.pass) {
    const int64_t cnt = std::min(L.pass, n - off);
    if (off > 0) {
      rc = window_gather(tr, p, window_track_dev, window_start_dev, n_dev, off, L.pass, poses, nullptr, ws + L.win_ws, st);
      if (rc) return rc;
    }
    rc = sf_score_windows(m, poses, cnt, p->seq_len, SF_REDUCE_MEAN, precision, scores_dev + off, nullptr, nullptr, ws + L.score_ws,
                          L.total - L.score_ws, st);
    if (rc) return rc;
  }
  return SF_OK;
}

// ------------------------------------------------------------------------------------ runner
// Host-buffer scoring pipeline: a copy stream uploads chunk c+1.. while the compute stream scores chunk c.  kRing
// device/pinned slots let the DMA engine run ahead of the kernels; the kernels of consecutive chunks queue back
// to back on ONE stream (they fill the GPU on their own, so running two chunks side by side only adds tail waves).
constexpr int kRing = 4;
struct sf_runner {
  const sf_model* m;
  int T, S;
  int64_t chunk, first_chunk;     // windows per upload; a shorter first upload (one wave) when chunk spans several
  size_t pose_elems;              // floats per window
  cudaStream_t copy_st, comp_st;
  cudaEvent_t copied[kRing], computed[kRing];
  float* pin_in[kRing];
  float* pin_out[kRing];
  float* dev_in[kRing];
  float* dev_out[kRing];
  void* ws;
  int64_t ws_bytes;
  // page-locked sources are scored in place (see sf_runner_score): full-batch score buffer + workspace, grown on demand
  float* big_out;
  int64_t big_out_cap;
  void* big_ws;
  int64_t big_ws_bytes;
  // sf_runner_score_tracks: two groups of tracks in flight (upload of group g+1 under the kernels of group g); grow-only
  struct TrackSlot {
    float* kp;
    int32_t* frame_no;
    int64_t frames_cap;
    char* out_dev;         // scores | labels | window_track | window_start, `win_cap` entries each
    char* out_pin;
    int64_t win_cap;
    void* ws;
    int64_t ws_bytes;
    char* tab_pin;         // per-track tables of the group, packed for one H2D copy on the copy stream
    int64_t tab_cap;
    cudaEvent_t uploaded, done;
  } ts[2];
  uint8_t* gt_dev;
  int64_t gt_cap;
};

extern "C" int sf_runner_create(const sf_model* m, int32_t T, int64_t max_chunk, sf_runner** out) {
  int rc = check_T(m, T);
  if (rc) return rc;
  SF_REQUIRE(out && max_chunk >= 1, SF_E_INVALID, "sf_runner_create: bad argument");
  DeviceGuard guard;
  SF_CUDA_OK(guard.enter(m->device));
  sf_runner* r = new sf_runner();
  memset(r, 0, sizeof(*r));
  r->m = m;
  r->T = T;
  r->S = token_len(m, T);
  // whole waves: the tensor-core transformer walks tiles of 4 * floor(32 / S) windows with one CTA per SM, so a
  // chunk that is a multiple of sm_count * tile leaves no partial wave (and is a whole number of tokenizer waves)
  const int64_t wave = (int64_t)m->sm_count * 4 * (32 / std::max(1, std::min(r->S, 32)));
  r->chunk = max_chunk >= wave ? max_chunk / wave * wave : max_chunk;
  r->first_chunk = r->chunk > wave ? wave : 0;
  r->pose_elems = (size_t)m->cfg.in_channels * T * m->cfg.num_keypoints;
  r->ws_bytes = sf_workspace_bytes(m, r->chunk, T);
  cudaError_t e = cudaStreamCreateWithFlags(&r->copy_st, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&r->comp_st, cudaStreamNonBlocking);
  if (e == cudaSuccess && r->ws_bytes > 0) e = cudaMalloc(&r->ws, r->ws_bytes);
  for (int i = 0; i < kRing && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&r->copied[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->computed[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&r->pin_in[i], r->pose_elems * r->chunk * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost((void**)&r->pin_out[i], r->chunk * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->dev_in[i], r->pose_elems * r->chunk * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc((void**)&r->dev_out[i], r->chunk * sizeof(float));
  }
  if (e != cudaSuccess) {
    set_error("sf_runner_create: %s", cudaGetErrorString(e));
    sf_runner_destroy(r);
    return SF_E_CUDA;
  }
  *out = r;
  return SF_OK;
}

extern "C" void sf_runner_destroy(sf_runner* r) {
  if (!r) return;
  DeviceGuard guard;
  guard.enter(r->m->device);
  if (r->copy_st) cudaStreamSynchronize(r->copy_st);
  if (r->comp_st) cudaStreamSynchronize(r->comp_st);
  for (int i = 0; i < kRing; ++i) {
    if (r->pin_in[i]) cudaFreeHost(r->pin_in[i]);
    if (r->pin_out[i]) cudaFreeHost(r->pin_out[i]);
    if (r->dev_in[i]) cudaFree(r->dev_in[i]);
    if (r->dev_out[i]) cudaFree(r->dev_out[i]);
    if (r->copied[i]) cudaEventDestroy(r->copied[i]);
    if (r->computed[i]) cudaEventDestroy(r->computed[i]);
  }
  if (r->ws) cudaFree(r->ws);
  for (int i = 0; i < 2; ++i) {
    sf_runner::TrackSlot& s = r->ts[i];
    if (s.kp) cudaFree(s.kp);
    if (s.frame_no) cudaFree(s.frame_no);
    if (s.out_dev) cudaFree(s.out_dev);
    if (s.out_pin) cudaFreeHost(s.out_pin);
    if (s.ws) cudaFree(s.ws);
    if (s.tab_pin) cudaFreeHost(s.tab_pin);
    if (s.uploaded) cudaEventDestroy(s.uploaded);
    if (s.done) cudaEventDestroy(s.done);
  }
  if (r->gt_dev) cudaFree(r->gt_dev);
  if (r->big_out) cudaFree(r->big_out);
  if (r->big_ws) cudaFree(r->big_ws);
  if (r->copy_st) cudaStreamDestroy(r->copy_st);
  if (r->comp_st) cudaStreamDestroy(r->comp_st);
  delete r;
}

extern "C" float* sf_runner_pinned_poses(sf_runner* r, int32_t slot) {
  if (!r || slot < 0 || slot >= kRing) return nullptr;
  return r->pin_in[slot];
}

extern "C" int sf_runner_score(sf_runner* r, const float* poses_host, int64_t B, int32_t precision, float* scores_host) {
  SF_REQUIRE(r && (B == 0 || (poses_host && scores_host)), SF_E_INVALID, "sf_runner_score: null argument");
  DeviceGuard guard;
  SF_CUDA_OK(guard.enter(r->m->device));
  int64_t pending_off[kRing], pending_n[kRing];
  for (int i = 0; i < kRing; ++i) pending_off[i] = -1, pending_n[i] = 0;
  bool src_pinned = false;
  if (B > 0) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, poses_host) == cudaSuccess) src_pinned = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  // A page-locked source is DMA-ed chunk by chunk straight out of the caller's buffer (no staging memcpy) by the ring
  // below: the copy engine sustains ~55 GB/s over PCIe 5 against ~38 GB/s for kernel loads from mapped host memory, and
  // the upload of chunk c+1 runs under the kernels of chunk c.  SF_RUNNER_INPLACE=1 keeps the older zero-copy mode (one
  // pass of the kernels reading host memory in place), which wins only when the kernels, not PCIe, are the bottleneck.
  static const bool inplace = getenv("SF_RUNNER_INPLACE") != nullptr;
  if (src_pinned && inplace) {
    // Page-locked source: no staging copy at all.  The tokenizer's TMA (the precise path: its global loads) reads each
    // window straight out of host memory over PCIe one window ahead of use, so the whole batch is ONE pass of the two
    // kernels; only the scores come back.  (h2d bytes per call are the same B * window bytes, moved by the kernel.)
    const float* dptr = nullptr;
    if (cudaHostGetDevicePointer((void**)&dptr, (void*)poses_host, 0) == cudaSuccess && dptr) {
      const int64_t need_ws = sf_workspace_bytes(r->m, B, r->T);
      if (need_ws > r->big_ws_bytes) {
        if (r->big_ws) cudaFree(r->big_ws);
        r->big_ws = nullptr;
        r->big_ws_bytes = 0;
        SF_CUDA_OK(cudaMalloc(&r->big_ws, (size_t)need_ws));
        r->big_ws_bytes = need_ws;
      }
      if (B > r->big_out_cap) {
        if (r->big_out) cudaFree(r->big_out);
        r->big_out = nullptr;
        r->big_out_cap = 0;
        SF_CUDA_OK(cudaMalloc((void**)&r->big_out, (size_t)B * sizeof(float)));
        r->big_out_cap = B;
      }
      int rc = sf_score_windows(r->m, dptr, B, r->T, SF_REDUCE_MEAN, precision, r->big_out, nullptr, nullptr, r->big_ws,
                                r->big_ws_bytes, r->comp_st);
      if (rc) return rc;
      SF_CUDA_OK(cudaMemcpyAsync(scores_host, r->big_out, (size_t)B * sizeof(float), cudaMemcpyDeviceToHost, r->comp_st));
      SF_CUDA_OK(cudaStreamSynchronize(r->comp_st));
      return SF_OK;
    }
    cudaGetLastError();        // not mappable: fall through to the staged pipeline
  }
  // Chunk schedule.  The DMA engine is the bottleneck of this call (a config-A window is 3,264 bytes: 65,536 windows take
  // 3.9 ms of PCIe 5 against 2.9 ms of kernels), so what matters is the upload starting at once and as little work as possible
  // left after the last byte has landed: the remainder (B mod chunk) goes FIRST (a short upload, then the kernels run
  // under every later upload), full chunks follow, and the last full chunk is cut in two halves so that only half a chunk's
  // kernels run after the final upload.
  std::vector<std::pair<int64_t, int64_t>> sched;      // (first window, windows)
  {
    const int64_t ch = r->chunk, rem = B % ch, n_full = B / ch;
    int64_t o = 0;
    static const bool plain = getenv("SF_RUNNER_SCHED") && atoi(getenv("SF_RUNNER_SCHED")) == 0;      // A/B switch: chunks in order
    if (plain) {
      for (; o < B; o += ch) sched.emplace_back(o, std::min(ch, B - o));
    } else if (rem > 0) {
      sched.emplace_back(0, rem);
      o = rem;
    }
    for (int64_t k = 0; k < n_full && !plain; ++k, o += ch) {
      if (k == 0 && rem == 0 && n_full > 1 && r->first_chunk > 0) {      // no remainder to start with: one wave first
        sched.emplace_back(o, r->first_chunk);
        sched.emplace_back(o + r->first_chunk, ch - r->first_chunk);
      } else if (k + 1 == n_full && (n_full > 1 || rem > 0) && ch >= 2240) {
        const int64_t h = (ch / 2 + 279) / 280 * 280;  // whole tiles of both kernels (7- and 40-window tiles)
        sched.emplace_back(o, h);
        sched.emplace_back(o + h, ch - h);
      } else {
        sched.emplace_back(o, ch);
      }
    }
  }
  for (size_t c = 0; c < sched.size(); ++c) {
    const int s = (int)(c % kRing);
    const int64_t off = sched[c].first, n = sched[c].second;
    if (pending_off[s] >= 0) {                       // drain the slot before reusing its buffers
      SF_CUDA_OK(cudaEventSynchronize(r->computed[s]));
      memcpy(scores_host + pending_off[s], r->pin_out[s], pending_n[s] * sizeof(float));
      pending_off[s] = -1;
    }
    const float* src = poses_host + (size_t)off * r->pose_elems;
    // page-locked sources (the runner's own slots, cudaHostAlloc / cudaHostRegister / torch pinned memory) are
    // DMA-ed directly; pageable sources are staged through the runner's pinned slot first.
    const float* dma_src = src;
    if (!src_pinned) {
      memcpy(r->pin_in[s], src, (size_t)n * r->pose_elems * sizeof(float));
      dma_src = r->pin_in[s];
    }
    SF_CUDA_OK(cudaMemcpyAsync(r->dev_in[s], dma_src, (size_t)n * r->pose_elems * sizeof(float), cudaMemcpyHostToDevice, r->copy_st));
    SF_CUDA_OK(cudaEventRecord(r->copied[s], r->copy_st));
    SF_CUDA_OK(cudaStreamWaitEvent(r->comp_st, r->copied[s], 0));
    int rc = sf_score_windows(r->m, r->dev_in[s], n, r->T, SF_REDUCE_MEAN, precision, r->dev_out[s], nullptr, nullptr,
                              r->ws, r->ws_bytes, r->comp_st);
    if (rc) return rc;
    SF_CUDA_OK(cudaMemcpyAsync(r->pin_out[s], r->dev_out[s], (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, r->comp_st));
    SF_CUDA_OK(cudaEventRecord(r->computed[s], r->comp_st));
    pending_off[s] = off;
    pending_n[s] = n;
  }
  for (int s = 0; s < kRing; ++s)                   // drain (one compute stream: any order)
    if (pending_off[s] >= 0) {
      SF_CUDA_OK(cudaEventSynchronize(r->computed[s]));
      memcpy(scores_host + pending_off[s], r->pin_out[s], pending_n[s] * sizeof(float));
    }
  return SF_OK;
}

// Host tracks -> host scores (see the header).  Groups of whole tracks of about `chunk` windows' worth of frames: the raw
// detections of group g+1 are uploaded on the copy stream while group g is windowed and scored on the compute stream.
extern "C" int sf_runner_score_tracks(sf_runner* r, const sf_tracks* th, const sf_window_params* p, int32_t precision,
                                      float* scores_host, int32_t* labels_host, int32_t* window_track_host,
                                      int32_t* window_start_host, int64_t* n_windows_host) {
  SF_REQUIRE(r && th && p && n_windows_host, SF_E_INVALID, "sf_runner_score_tracks: null argument");
  SF_REQUIRE(p->seq_len == r->T, SF_E_INVALID, "the runner was created for T=%d, the windows have T=%d", r->T, p->seq_len);
  const sf_model* m = r->m;
  int rc = check_tracks_vs_model(m, th, p);
  if (rc) return rc;
  *n_windows_host = 0;
  const int64_t cap_total = sf_window_capacity(th, p);
  SF_REQUIRE(cap_total >= 0, SF_E_INVALID, "sf_runner_score_tracks: bad tracks / window parameters");
  if (cap_total == 0) return SF_OK;
  SF_REQUIRE(scores_host && th->kp_dev && th->frame_no_dev && th->track_offsets_host, SF_E_INVALID, "sf_runner_score_tracks: null buffer");
  DeviceGuard guard;
  SF_CUDA_OK(guard.enter(m->device));
  const int KC = th->kp_channels == 2 ? 2 : 3;
  const size_t frame_floats = (size_t)th->kp_per_frame * KC;
  const bool has_gt = th->gt_dev && th->gt_offsets_host && th->track_video_host && th->n_videos > 0;
  if (has_gt) {                                      // frame labels of every video: uploaded once
    const int64_t gt_bytes = th->gt_offsets_host[th->n_videos];
    if (gt_bytes > r->gt_cap) {
      if (r->gt_dev) cudaFree(r->gt_dev);
      r->gt_dev = nullptr;
      r->gt_cap = 0;
      SF_CUDA_OK(cudaMalloc((void**)&r->gt_dev, (size_t)std::max<int64_t>(gt_bytes, 1)));
      r->gt_cap = gt_bytes;
    }
    SF_CUDA_OK(cudaMemcpyAsync(r->gt_dev, th->gt_dev, (size_t)gt_bytes, cudaMemcpyHostToDevice, r->copy_st));
  }
  // ---- groups of whole tracks
  const int64_t target_frames = std::max<int64_t>(r->chunk * (int64_t)p->stride, 1);
  std::vector<int> g_begin;                          // first track of each group (+ sentinel)
  {
    int t = 0;
    while (t < th->n_tracks) {
      g_begin.push_back(t);
      const int64_t f0 = th->track_offsets_host[t];
      // the first upload is not hidden behind any kernel: start with a quarter-size group
      const int64_t target = g_begin.size() == 1 ? std::max<int64_t>(target_frames / 4, 1) : target_frames;
      int e = t + 1;
      while (e < th->n_tracks && th->track_offsets_host[e + 1] - f0 <= target) ++e;
      t = e;
    }
    g_begin.push_back(th->n_tracks);
  }
  const int n_groups = (int)g_begin.size() - 1;
  struct Group {
    sf_tracks tr;
    std::vector<int64_t> off;                        // track offsets relative to the group's first frame
    int64_t cap, n;
    int t0;
  };
  std::vector<Group> groups((size_t)n_groups);
  auto prepare = [&](int g) -> int {                 // size the slot and enqueue the upload of group g
    Group& G = groups[g];
    sf_runner::TrackSlot& s = r->ts[g & 1];
    const int t0 = g_begin[g], t1 = g_begin[g + 1];
    const int64_t f0 = th->track_offsets_host[t0], f1 = th->track_offsets_host[t1];
    G.t0 = t0;
    G.off.resize((size_t)(t1 - t0 + 1));
    for (int t = t0; t <= t1; ++t) G.off[(size_t)(t - t0)] = th->track_offsets_host[t] - f0;
    G.tr = *th;
    G.tr.n_frames = f1 - f0;
    G.tr.n_tracks = t1 - t0;
    G.tr.track_offsets_host = G.off.data();
    G.tr.track_video_host = has_gt ? th->track_video_host + t0 : nullptr;
    G.tr.gt_dev = has_gt ? r->gt_dev : nullptr;
    G.cap = sf_window_capacity(&G.tr, p);
    G.n = 0;
    if (f1 - f0 > s.frames_cap) {
      if (s.kp) cudaFree(s.kp);
      if (s.frame_no) cudaFree(s.frame_no);
      s.kp = nullptr;
      s.frame_no = nullptr;
      s.frames_cap = 0;
      const int64_t want = std::max(f1 - f0, target_frames + target_frames / 4);
      SF_CUDA_OK(cudaMalloc((void**)&s.kp, (size_t)want * frame_floats * sizeof(float)));
      SF_CUDA_OK(cudaMalloc((void**)&s.frame_no, (size_t)want * sizeof(int32_t)));
      s.frames_cap = want;
    }
    if (G.cap > s.win_cap) {
      if (s.out_dev) cudaFree(s.out_dev);
      if (s.out_pin) cudaFreeHost(s.out_pin);
      s.out_dev = s.out_pin = nullptr;
      s.win_cap = 0;
      const int64_t want = std::max(G.cap, r->chunk + r->chunk / 4);
      SF_CUDA_OK(cudaMalloc((void**)&s.out_dev, (size_t)want * 16));
      SF_CUDA_OK(cudaMallocHost((void**)&s.out_pin, (size_t)want * 16 + 256));
      s.win_cap = want;
    }
    G.tr.kp_dev = s.kp;
    G.tr.frame_no_dev = s.frame_no;
    if (G.cap > 0) {
      const int64_t need = sf_score_from_tracks_workspace_bytes(m, &G.tr, p);
      SF_REQUIRE(need >= 0, SF_E_INVALID, "sf_runner_score_tracks: bad group");
      if (need > s.ws_bytes) {
        if (s.ws) cudaFree(s.ws);
        s.ws = nullptr;
        s.ws_bytes = 0;
        SF_CUDA_OK(cudaMalloc(&s.ws, (size_t)(need + need / 4)));
        s.ws_bytes = need + need / 4;
      }
    }
    if (!s.uploaded) SF_CUDA_OK(cudaEventCreateWithFlags(&s.uploaded, cudaEventDisableTiming));
    if (!s.done) SF_CUDA_OK(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    // the slot's previous group (g - 2) has been drained by the caller before this point
    SF_CUDA_OK(cudaMemcpyAsync(s.kp, th->kp_dev + (size_t)f0 * frame_floats, (size_t)(f1 - f0) * frame_floats * sizeof(float),
                               cudaMemcpyHostToDevice, r->copy_st));
    SF_CUDA_OK(cudaMemcpyAsync(s.frame_no, th->frame_no_dev + f0, (size_t)(f1 - f0) * sizeof(int32_t), cudaMemcpyHostToDevice, r->copy_st));
    if (G.cap > 0) {
      // the group's per-track tables go through pinned staging on the copy stream too: a pageable copy on the compute
      // stream would make the host wait for the previous group's kernels before it can enqueue this one
      const int64_t tb = window_tables_bytes(&G.tr, p);
      if (tb > s.tab_cap) {
        if (s.tab_pin) cudaFreeHost(s.tab_pin);
        s.tab_pin = nullptr;
        s.tab_cap = 0;
        SF_CUDA_OK(cudaMallocHost((void**)&s.tab_pin, (size_t)(tb + tb / 4)));
        s.tab_cap = tb + tb / 4;
      }
      window_tables_pack(&G.tr, p, s.tab_pin);
      SF_CUDA_OK(cudaMemcpyAsync(s.ws, s.tab_pin, (size_t)tb, cudaMemcpyHostToDevice, r->copy_st));
    }
    SF_CUDA_OK(cudaEventRecord(s.uploaded, r->copy_st));
    return SF_OK;
  };
  int64_t out_off = 0;
  auto drain = [&](int g) -> int {                   // D2H results of group g -> the caller's arrays
    Group& G = groups[g];
    sf_runner::TrackSlot& s = r->ts[g & 1];
    SF_CUDA_OK(cudaEventSynchronize(s.done));
    if (G.n < 0) memcpy(&G.n, s.out_pin + s.win_cap * 16, sizeof(int64_t));
    const int64_t n = G.n, wc = s.win_cap;
    memcpy(scores_host + out_off, s.out_pin, (size_t)n * 4);
    if (labels_host) memcpy(labels_host + out_off, s.out_pin + wc * 4, (size_t)n * 4);
    if (window_track_host) {
      const int32_t* wt = (const int32_t*)(s.out_pin + wc * 8);
      for (int64_t i = 0; i < n; ++i) window_track_host[out_off + i] = wt[i] + G.t0;
    }
    if (window_start_host) memcpy(window_start_host + out_off, s.out_pin + wc * 12, (size_t)n * 4);
    out_off += n;
    return SF_OK;
  };
  // Two groups in flight.  With kernels that take the window count from device memory (tokenizer v2 + the tensor-core
  // transformer) a group is enqueued in full -- upload, windowing, scoring, D2H of capacity-sized results and the count --
  // without any host synchronisation, so the GPU never waits for the host between groups; otherwise sf_score_from_tracks
  // synchronises once per group.
  const bool async = supports_dev_count(m, p->seq_len, precision);
  auto launch = [&](int g) -> int {
    Group& G = groups[g];
    sf_runner::TrackSlot& s = r->ts[g & 1];
    SF_CUDA_OK(cudaStreamWaitEvent(r->comp_st, s.uploaded, 0));
    const int64_t wc = s.win_cap;
    if (G.cap > 0) {
      float* sc = (float*)s.out_dev;
      int32_t *lb = (int32_t*)(s.out_dev + wc * 4), *wt = (int32_t*)(s.out_dev + wc * 8), *wsr = (int32_t*)(s.out_dev + wc * 12);
      int64_t n_copy = 0;
      if (async) {
        const int64_t* n_dev = nullptr;
        int rc2 = score_from_tracks_async(m, &G.tr, p, precision, sc, lb, wt, wsr, &n_dev, s.ws, s.ws_bytes, r->comp_st, true);
        if (rc2) return rc2;
        SF_CUDA_OK(cudaMemcpyAsync(s.out_pin + wc * 16, n_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, r->comp_st));
        G.n = -1;                                      // read from the pinned tail when the group is drained
        n_copy = G.cap;
      } else {
        int rc2 = sf_score_from_tracks(m, &G.tr, p, precision, sc, lb, wt, wsr, &G.n, s.ws, s.ws_bytes, r->comp_st);
        if (rc2) return rc2;
        n_copy = G.n;
      }
      for (int k = 0; k < 4 && n_copy > 0; ++k)
        SF_CUDA_OK(cudaMemcpyAsync(s.out_pin + wc * 4 * k, s.out_dev + wc * 4 * k, (size_t)n_copy * 4, cudaMemcpyDeviceToHost, r->comp_st));
    }
    SF_CUDA_OK(cudaEventRecord(s.done, r->comp_st));
    return SF_OK;
  };
  for (int g = 0; g < n_groups; ++g) {
    if (g >= 2) {                                    // slot g & 1 still holds group g - 2
      rc = drain(g - 2);
      if (rc) return rc;
    }
    rc = prepare(g);                                 // upload of this group runs under the previous group's kernels
    if (rc) return rc;
    rc = launch(g);
    if (rc) return rc;
  }
  for (int g = std::max(0, n_groups - 2); g < n_groups; ++g) {
    rc = drain(g);
    if (rc) return rc;
  }
  *n_windows_host = out_off;
  return SF_OK;
}

// ------------------------------------------------------------------------------------ test infrastructure
// Runs the tokenizer-v2 tile program of `m` (which may be a HOST-ONLY model, device < 0) on the host emulator
// (tok2_emulate.cu).  Returns SF_E_UNSUPPORTED when the shape is outside tokenizer v2, SF_E_INVALID on an emulator
// failure (deadlock / hazard; message in sf_last_error).  Not part of the C ABI; no product path calls it.
#include "tok2_build.h"
namespace sf { const t2::Static* tok2_static(const Tok2State* s); }
extern "C" int sfdbg_tok2_emulate(const sf_model* m, const float* poses_host, int64_t B, int32_t T, float* tokens_host,
                                  uint32_t schedule_seed, int32_t* info_out) {
  SF_REQUIRE(m && m->tok2 && (B == 0 || (poses_host && tokens_host)), SF_E_INVALID, "sfdbg_tok2_emulate: bad argument");
  const t2::Static* st = tok2_static(m->tok2);
  t2::Program pr;
  t2::build_program(*st, T, m->max_smem_optin, &pr);
  SF_REQUIRE(pr.ok, SF_E_UNSUPPORTED, "tokenizer v2 does not cover this shape: %s", pr.why.c_str());
  if (info_out) {
    info_out[0] = pr.plan.n_groups;
    info_out[1] = pr.plan.n_stages[0] + pr.plan.n_stages[1];
    info_out[2] = pr.plan.n_loads;
    info_out[3] = pr.plan.n_mma;
    info_out[4] = (int32_t)pr.plan.smem_bytes;
    info_out[5] = pr.plan.WT;
  }
  std::string err;
  SF_REQUIRE(t2::emulate(*st, pr, poses_host, B, tokens_host, schedule_seed, &err), SF_E_INVALID, "%s", err.c_str());
  return SF_OK;
}

// Prints the tokenizer-v2 tile program of (m, T): groups (MMAs, waits), stages per team, loads.  Debugging aid that goes with
// profiles/tok2_timing.py (the stamps there are group / stage indices); not part of the C ABI.
extern "C" int sfdbg_tok2_describe(const sf_model* m, int32_t T) {
  SF_REQUIRE(m && m->tok2, SF_E_INVALID, "sfdbg_tok2_describe: bad argument");
  const t2::Static* st = tok2_static(m->tok2);
  t2::Program pr;
  t2::build_program(*st, T, m->max_smem_optin > 0 ? m->max_smem_optin - 256 : 232448 - 256, &pr);
  SF_REQUIRE(pr.ok, SF_E_UNSUPPORTED, "tokenizer v2 does not cover this shape: %s", pr.why.c_str());
  const t2::Plan& pl = pr.plan;
  printf("plan: WT=%d rows=%d smem=%u P@%u Q@%u W@%u xin@%u const=%u groups=%d stages=%d+%d loads=%d mma=%d\n", pl.WT, pl.rows, pl.smem_bytes,
         pl.off_P, pl.off_Q, pl.off_W, pl.off_xin, pl.const_bytes, pl.n_groups, pl.n_stages[0], pl.n_stages[1], pl.n_loads, pl.n_mma);
  for (size_t g = 0; g < pr.groups.size(); ++g) {
    const t2::Group& gr = pr.groups[g];
    int ncols = 0, dmin = 1 << 30, dmax = 0;
    for (int i = gr.first; i < gr.first + gr.count; ++i) {
      const int N = (int)((pr.mma[i].idesc >> 17) & 0x3F) * 8, d = (int)(pr.mma[i].d & 0x1FF);
      ncols += N;
      dmin = std::min(dmin, d);
      dmax = std::max(dmax, d + N);
    }
    printf("G%-3zu mma=%-3d sumN=%-5d tmem=[%d,%d) wait e0=%d e1=%d l=%d prev=%d.%d\n", g, gr.count, ncols, dmin, dmax, gr.wait_e[0], gr.wait_e[1],
           gr.wait_l, gr.prev_team, gr.prev_stage);
  }
  static const char* kType[] = {"G0", "CVT", "XEPI0", "TOKENS"};
  for (int t = 0; t < t2::kTeams; ++t)
    for (size_t e = 0; e < pr.stages[t].size(); ++e) {
      const t2::Stage& s = pr.stages[t][e];
      printf("E%d.%-3zu %-6s flags=%d wait g=%d l=%d eo=%d gprev=%d tmem_col=%d n_cg=%d dst=%u p=[%d,%d)\n", t, e, kType[s.type], s.flags, s.wait_g,
             s.wait_l, s.wait_eo, s.wait_g_prev, s.tmem_col, s.n_cg, s.dst_off, s.p0, s.p1);
    }
  for (size_t l = 0; l < pr.loads.size(); ++l) {
    const t2::Load& ld = pr.loads[l];
    printf("L%-2zu kind=%d bytes=%u dst=%u wait g=%d e0=%d e1=%d gprev=%d\n", l, ld.kind, ld.bytes, ld.dst_off, ld.wait_g, ld.wait_e[0], ld.wait_e[1],
           ld.wait_g_prev);
  }
  fflush(stdout);
  return SF_OK;
}
