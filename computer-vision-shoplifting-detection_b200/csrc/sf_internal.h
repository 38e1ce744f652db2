// Internal structures shared by the translation units of libshopformer_b200.so.
// Nothing here is part of the C ABI (see include/shopformer_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <vector>

#include "shopformer_b200.h"

namespace sf {

constexpr int kMaxBlocks = SF_MAX_BLOCKS;
constexpr int kMaxLayers = 6;     // encoder / decoder depth supported by the by-value kernel params
constexpr int kMaxV = 32;         // keypoints live on the lanes of one warp
constexpr int kTaps = 9;          // temporal kernel size of the reference (gcae.py:223)
constexpr int kHalo = 4;          // its padding

// ---- error plumbing -----------------------------------------------------------------
void set_error(const char* fmt, ...);
// host-side launch counters (sfdbg_launch_counts): which kernel served a call -- the tests' proof that the tensor-core kernels ran
enum LaunchKind { LK_TOK2 = 0, LK_TOK_BF16 = 1, LK_TOK_FP32 = 2, LK_XF_TC = 3, LK_XF_FP32 = 4, LK_COUNT = 5 };
void count_launch(int kind);
#define SF_CUDA_OK(expr)                                                                    \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::sf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SF_E_CUDA;                                                                     \
    }                                                                                       \
  } while (0)
#define SF_REQUIRE(cond, code, ...)  \
  do {                               \
    if (!(cond)) {                   \
      ::sf::set_error(__VA_ARGS__);  \
      return (code);                 \
    }                                \
  } while (0)

// Every entry point that takes a model / tracks runs on THAT object's device and leaves the caller's current device
// untouched (a torch process has its own notion of the current device).
struct DeviceGuard {
  int prev = -1;
  bool active = false;
  cudaError_t enter(int device) {
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) return e;
    if (prev != device) {
      e = cudaSetDevice(device);
      if (e != cudaSuccess) return e;
      active = true;
    }
    return cudaSuccess;
  }
  ~DeviceGuard() {
    if (active) cudaSetDevice(prev);
  }
};

// ---- tokenizer (device view; all pointers into one HBM arena) ------------------------
struct TokBlock {
  int cin, cout, stride, identity_res;
  int ell_width;            // entries per row of the ELL adjacency
  const float* gcn_w;       // [cin][cout]
  const float* gcn_b;       // [cout]
  const float* tcn_w;       // [cout(in)][9][cout(out)]   BatchNorm scale folded in
  const float* res_w;       // [cin][cout] BN-folded 1x1 residual conv, nullptr if identity
  const float* out_b;       // [cout] folded tcn bias (+ folded residual bias)
  const float* ell_val;     // [V][ell_width]  normalised adjacency, row-compressed
  const int* ell_col;       // [V][ell_width]
};

struct Tokenizer {
  int n_blocks, V, c_in, pool_tokens;
  const float* in_scale;    // [c_in*V]  BatchNorm1d folded to scale/shift, index c*V+v
  const float* in_shift;
  TokBlock blk[kMaxBlocks];
};

// ---- transformer ----------------------------------------------------------------------
struct Linear {
  const float* wt;          // [K][N]  (transposed nn.Linear weight: n contiguous)
  const float* b;           // [N]
  int K, N;
};
struct Norm {
  const float* g;
  const float* b;
};
struct Attn {
  Linear qkv;               // packed in_proj: columns [0:d) q, [d:2d) k, [2d:3d) v
  Linear out;
};
struct EncLayer {
  Attn sa;
  Linear ff1, ff2;
  Norm n1, n2;
};
struct DecLayer {
  Attn sa, ca;
  Linear ff1, ff2;
  Norm n1, n2, n3;
};
struct Transformer {
  int variant, d_tok, d_model, heads, n_enc, n_dec, d_ff, has_io_proj;
  const float* pe;          // [100][d_model]  transformer.pos_encoder.pe
  const float* pe_score;    // [100][d_tok]    variant 1: the facade's pos_encoder.pe
  Linear in_proj;           // variant 2, iff has_io_proj
  Linear out_proj;          // variant 1 output_proj / variant 2 output_projection
  Norm enc_norm, dec_norm;  // variant 2 final norms
  EncLayer enc[kMaxLayers];
  DecLayer dec[kMaxLayers];
};
static_assert(sizeof(Transformer) <= 3800, "Transformer must fit in kernel parameter space");
static_assert(sizeof(Tokenizer) <= 3800, "Tokenizer must fit in kernel parameter space");

}  // namespace sf

namespace sf {
// bf16 operand images of the tokenizer for the tcgen05 path, pre-packed in the exact shared-memory
// layout the MMA descriptors expect ([K chunk of 8][N row][8], see tc_common.cuh) so that staging is
// a straight 16-byte copy.  Channel counts are padded up to multiples of 16 with zeros.
struct BfBlockW {
  const uint16_t* tcn;      // [9 taps x npad/8 chunks][npad][8]   BN-folded temporal conv
  const uint16_t* gcn;      // [kin_pad/8][npad][8]                graph-conv weight (nullptr for block 0)
  const uint16_t* res;      // [kin_pad/8][npad][8]                folded 1x1 residual (nullptr if identity / block 0)
  const float* gcn_b;       // [npad] zero padded
  const float* out_b;       // [npad] zero padded
  int npad, kin_pad;
};
struct BfTokenizerW {
  BfBlockW blk[kMaxBlocks];
};
}  // namespace sf

namespace sf {
// "Program" of the tensor-core transformer: the host flattens either variant (post-LN/ReLU/shifted target
// or pre-LN/GELU/final norms/projections) into a list of ops that one kernel interprets per row tile.
enum XfOpType { XF_INIT = 0, XF_GEMM = 1, XF_POST = 2 };
enum XfInit { XI_TOK_PE = 0, XI_SHIFT_TOK_PE = 1, XI_TOK_TO_AOP = 2 };
enum XfEpi { XE_NONE = 0, XE_ATTN = 1, XE_ACT_H = 2, XE_STREAM_ADD = 3, XE_STREAM_SET_PE = 4, XE_SCORE = 5 };
enum XfPost { XP_NONE = 0, XP_COPY_TO_AOP = 1, XP_LN_INPLACE_TO_AOP = 2, XP_LN_TO_AOP = 3, XP_LN_TO_MEM = 4, XP_LN_SCORE = 5 };
enum XfSrc { XS_AOP = 0, XS_HOP = 1, XS_MEM = 2 };
struct XfOp {
  int32_t type, init_mode, a_src, K, N, tmem_col, accumulate, epi, act, post, also_mem, chain;   // chain > 1: this GEMM and the next chain-1 GEMMs (no epilogue of their own) are issued in ONE phase
  const uint16_t* w;        // [K/8][N][8] 16-bit operand image (K-major B; fp16 or bf16, XfProgram::f16)
  const float* bias;        // [N] zero padded
  const float* ln_g;
  const float* ln_b;
};
struct XfProgram {
  const XfOp* ops;          // device
  int n_ops, dp, dtp, supported;
  int max_w_bytes;          // largest weight image (ring slot size)
  int f16;                  // the 16-bit operands (weight images, activations) are fp16 instead of bf16
};
}  // namespace sf

namespace sf { struct Tok2State; }
namespace sf {
// second stream of a model: sf_score_windows runs the two halves of a large batch side by side so that the partial last wave
// of one half's persistent kernels is filled by the other half's CTAs (api.cu)
struct SideStream {
  cudaStream_t st = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  std::mutex mu;              // one fork / join sequence is enqueued at a time (the events are re-recorded per call)
};
}

struct sf_model {
  sf_config cfg;
  int device;
  int sm_count;
  int max_smem_optin;
  float* arena;             // one cudaMalloc holding every packed fp32 tensor
  size_t arena_bytes;
  sf::Tokenizer tok;
  sf::Transformer xf;
  uint16_t* arena_bf16;     // bf16 operand images (tcgen05 path)
  size_t arena_bf16_bytes;
  sf::BfTokenizerW tokbf;
  sf::XfProgram xfprog;
  sf::XfOp* xfops_dev;
  sf::Tok2State* tok2;      // tokenizer v2: operand images + per-T tile programs (tokenizer2_bf16.cu)
  float* host_arena;        // device < 0 only: a HOST model for the program emulator (tests); no entry point computes with it
  sf::SideStream* side;     // nullptr when it could not be created (the batch then runs on the caller's stream alone)
};

namespace sf {
// launchers (fp32 CUDA-core path)
int launch_tokenizer_fp32(const sf_model* m, const float* poses, int64_t B, int T, float* tokens,
                          void* ws, int64_t ws_bytes, cudaStream_t st);
int64_t tokenizer_fp32_workspace(const sf_model* m, int64_t B, int T);
int launch_transformer_fp32(const sf_model* m, const float* tokens, int64_t B, int S, int reduction,
                            float* recon, float* scores, cudaStream_t st);
int launch_score(const sf_model* m, const float* tokens, const float* recon, int64_t B, int S,
                 int reduction, float* scores, cudaStream_t st);
int token_len(const sf_model* m, int T);
// windowing.cu: the index (flag / scan / compact) and gather halves of sf_window_normalize
int64_t window_candidates(const sf_tracks* tr, const sf_window_params* p);
int window_index(const sf_tracks* tr, const sf_window_params* p, int32_t* labels_dev, int32_t* window_track_dev,
                 int32_t* window_start_dev, int64_t* n_windows_dev, void* workspace_dev, int64_t workspace_bytes, cudaStream_t st,
                 bool tables_resident = false);
int64_t window_tables_bytes(const sf_tracks* tr, const sf_window_params* p);
void window_tables_pack(const sf_tracks* tr, const sf_window_params* p, char* host_dst);
int window_gather(const sf_tracks* tr, const sf_window_params* p, const int32_t* window_track_dev, const int32_t* window_start_dev,
                  const int64_t* n_windows_dev, int64_t w_begin, int64_t w_cap, float* poses_dev, int32_t* frame_idx_dev,
                  void* workspace_dev, cudaStream_t st);
int window_device_of(const sf_tracks* tr, int* device);
// bf16 tcgen05 tokenizer (returns SF_E_UNSUPPORTED for shapes it does not cover)
int launch_tokenizer_bf16(const sf_model* m, const float* poses, int64_t B, int T, float* tokens, cudaStream_t st);
bool tokenizer_bf16_supported(const sf_model* m, int T);
// tokenizer v2 (multi-window tiles, see tok2.h); `tokenizer_bf16` dispatches to it when the shape is covered
Tok2State* tok2_create(const Tokenizer& host_tok, int pool_tokens, bool upload, bool allow_f16);
bool tokenizer2_f16(const sf_model* m);
void tok2_destroy(Tok2State* s);
bool tokenizer2_supported(const sf_model* m, int T);
const char* tokenizer2_why(const sf_model* m, int T);
// Optional device-side batch size: the kernels process min(B, max(0, *n - off)) windows, so a caller that only knows an
// upper bound on the host (windows compacted on the device) can launch without reading the count back.
struct DevCount {
  const int64_t* n = nullptr;
  int64_t off = 0;
};
__device__ __forceinline__ int64_t dev_count_clamp(const DevCount& c, int64_t B) {
  if (!c.n) return B;
  const int64_t left = *c.n - c.off;
  return left < 0 ? 0 : (left < B ? left : B);
}
int launch_tokenizer2(const sf_model* m, const float* poses, int64_t B, int T, float* tokens, cudaStream_t st, DevCount cnt = DevCount());
// bf16 tcgen05 transformer + fused score
int launch_transformer_bf16(const sf_model* m, const float* tokens, int64_t B, int S, int reduction, float* recon,
                            float* scores, cudaStream_t st, DevCount cnt = DevCount());
bool transformer_bf16_supported(const sf_model* m, int S);
}  // namespace sf
