// C-ABI entry points of the scoring hot path + the host-buffer runner.
#include <algorithm>
#include <cstring>

#include "sf_internal.h"

using namespace sf;

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
int launch_tokenizer_tc(const sf_model* m, const float* poses, int64_t B, int T, float* tokens, cudaStream_t st) {
  if (tokenizer2_supported(m, T)) return launch_tokenizer2(m, poses, B, T, tokens, st);
  return launch_tokenizer_bf16(m, poses, B, T, tokens, st);
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

extern "C" int sf_score_windows(const sf_model* m, const float* poses_dev, int64_t B, int32_t T, int32_t reduction,
                                int32_t precision, float* scores_dev, float* tokens_dev, float* recon_dev,
                                void* workspace_dev, int64_t workspace_bytes, void* stream) {
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
  for (int64_t off = 0; off < B; off += pass) {
    const int64_t n = std::min(pass, B - off);
    float* tok = tokens_dev ? tokens_dev + off * tok_elems : tok_ws;
    const float* x = poses_dev + off * pose_elems;
    rc = precision == SF_PREC_BF16 ? launch_tokenizer_tc(m, x, n, T, tok, st) : launch_tokenizer_fp32(m, x, n, T, tok, ws, ws_left, st);
    if (rc) return rc;
    float* rec = recon_dev ? recon_dev + off * tok_elems : nullptr;
    float* sc = scores_dev + off * score_stride;
    rc = precision == SF_PREC_BF16 ? launch_transformer_bf16(m, tok, n, S, reduction, rec, sc, st)
                                   : launch_transformer_fp32(m, tok, n, S, reduction, rec, sc, st);
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
  if (src_pinned) {
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
  int64_t off = 0;
  for (int64_t c = 0; off < B; ++c) {
    const int s = (int)(c % kRing);
    // the first upload is not hidden behind any kernel: start with a single wave, then full chunks
    const int64_t want = (c == 0 && r->first_chunk > 0) ? r->first_chunk : r->chunk;
    const int64_t n = std::min(want, B - off);
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
    off += n;
  }
  for (int s = 0; s < kRing; ++s)                   // drain (one compute stream: any order)
    if (pending_off[s] >= 0) {
      SF_CUDA_OK(cudaEventSynchronize(r->computed[s]));
      memcpy(scores_host + pending_off[s], r->pin_out[s], pending_n[s] * sizeof(float));
    }
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
