// fp32 encoder/decoder transformer + fused reconstruction-error score (CUDA cores).
//
// Maths: SURVEY Appendix A.2 / A.3.
//   variant 1  shopformer/models/transformer.py:60-118 (encoder layer, post-LN, ReLU),
//              :121-196 (decoder layer), :304-329 (zero start token + shifted target, no masks,
//              output_proj); score shopformer/models/shopformer.py:150-178 (vs tokens + PE).
//   variant 2  shopformer_2/models/transformer.py:105-136,147-194 (nn.Transformer{En,De}coder,
//              norm_first, exact GELU, final LayerNorms, optional 136<->144 projections);
//              score shopformer_2/models/shopformer.py:178-186 (vs raw tokens).
//
// One CTA owns a tile of windows (R = windows x S rows <= 48) and keeps four row-major fp32
// buffers on chip: the encoder stream, the decoder stream, a scratch row block and a wide
// block (QKV / FFN hidden).  Every nn.Linear is a register-tiled [R x K] x [K x N] product whose
// weights stream from L2 (they are shared by all CTAs; ~2 MB for config A); attention over the
// S = 2..3 tokens of a window runs on one warp per (row, head); LayerNorm is one warp per row.
// The reconstruction never leaves the SM unless the caller asks for it: the squared error
// against the score target is reduced in the same kernel and only B floats are written.
#include <algorithm>
#include <cmath>

#include "sf_internal.h"

namespace sf {
namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsMax = 48;
constexpr int kSMax = 16;
constexpr float kLnEps = 1e-5f;

enum Act { kActNone = 0, kActRelu = 1, kActGelu = 2 };

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ float act_apply(float x, int act) {
  if (act == kActRelu) return fmaxf(x, 0.f);
  if (act == kActGelu) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  return x;
}

// out[r][n] (+)= act(sum_k in[r][k] * wt[k*ldw + n] + b[n]) for r < R, n < N.  K, N multiples of 4.
template <int RT>
__device__ __forceinline__ void linear_rt(const float* __restrict__ in, int ldin, const float* __restrict__ wt, int ldw,
                                          const float* __restrict__ bias, int K, int N, float* __restrict__ out,
                                          int ldout, int R, int act, bool accumulate) {
  const int n_cg = N >> 2;
  const int n_rg = (R + RT - 1) / RT;
  const int items = n_cg * n_rg;
  for (int it = threadIdx.x; it < items; it += kThreads) {
    const int cg = it % n_cg, rg = it / n_cg;
    const int n0 = cg << 2, r0 = rg * RT;
    float acc[RT][4];
#pragma unroll
    for (int a = 0; a < RT; ++a) acc[a][0] = acc[a][1] = acc[a][2] = acc[a][3] = 0.f;
    const float* xp[RT];
#pragma unroll
    for (int a = 0; a < RT; ++a) xp[a] = in + (size_t)min(r0 + a, R - 1) * ldin;
    const float* wp = wt + n0;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
      const float4 w0 = ldg4(wp + (size_t)(k + 0) * ldw);
      const float4 w1 = ldg4(wp + (size_t)(k + 1) * ldw);
      const float4 w2 = ldg4(wp + (size_t)(k + 2) * ldw);
      const float4 w3 = ldg4(wp + (size_t)(k + 3) * ldw);
#pragma unroll
      for (int a = 0; a < RT; ++a) {
        const float4 x = *reinterpret_cast<const float4*>(xp[a] + k);
        acc[a][0] = fmaf(x.x, w0.x, acc[a][0]); acc[a][1] = fmaf(x.x, w0.y, acc[a][1]);
        acc[a][2] = fmaf(x.x, w0.z, acc[a][2]); acc[a][3] = fmaf(x.x, w0.w, acc[a][3]);
        acc[a][0] = fmaf(x.y, w1.x, acc[a][0]); acc[a][1] = fmaf(x.y, w1.y, acc[a][1]);
        acc[a][2] = fmaf(x.y, w1.z, acc[a][2]); acc[a][3] = fmaf(x.y, w1.w, acc[a][3]);
        acc[a][0] = fmaf(x.z, w2.x, acc[a][0]); acc[a][1] = fmaf(x.z, w2.y, acc[a][1]);
        acc[a][2] = fmaf(x.z, w2.z, acc[a][2]); acc[a][3] = fmaf(x.z, w2.w, acc[a][3]);
        acc[a][0] = fmaf(x.w, w3.x, acc[a][0]); acc[a][1] = fmaf(x.w, w3.y, acc[a][1]);
        acc[a][2] = fmaf(x.w, w3.z, acc[a][2]); acc[a][3] = fmaf(x.w, w3.w, acc[a][3]);
      }
    }
    const float4 b4 = ldg4(bias + n0);
#pragma unroll
    for (int a = 0; a < RT; ++a) {
      if (r0 + a >= R) break;
      float4* op = reinterpret_cast<float4*>(out + (size_t)(r0 + a) * ldout + n0);
      float4 v = make_float4(act_apply(acc[a][0] + b4.x, act), act_apply(acc[a][1] + b4.y, act),
                             act_apply(acc[a][2] + b4.z, act), act_apply(acc[a][3] + b4.w, act));
      if (accumulate) {
        const float4 o = *op;
        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
      }
      *op = v;
    }
  }
}

__device__ __forceinline__ void linear(const float* in, int ldin, const float* wt, int ldw, const float* bias, int K, int N,
                                       float* out, int ldout, int R, int act, bool accumulate) {
  const int n_cg = N >> 2;
  if (n_cg * ((R + 7) / 8) >= (kThreads * 3) / 4) linear_rt<8>(in, ldin, wt, ldw, bias, K, N, out, ldout, R, act, accumulate);
  else if (n_cg * ((R + 3) / 4) >= (kThreads * 3) / 4) linear_rt<4>(in, ldin, wt, ldw, bias, K, N, out, ldout, R, act, accumulate);
  else linear_rt<2>(in, ldin, wt, ldw, bias, K, N, out, ldout, R, act, accumulate);
  __syncthreads();
}

// out[r] = LayerNorm(in[r]) * g + b, one warp per row (two-pass variance).  in may alias out.
__device__ __forceinline__ void layernorm_rows(const float* in, float* out, int ld, int d, int R, const Norm& nm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < R; r += kWarps) {
    const float* x = in + (size_t)r * ld;
    float s = 0.f;
    for (int j = lane; j < d; j += 32) s += x[j];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)d;
    float q = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float c = x[j] - mean;
      q = fmaf(c, c, q);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)d + kLnEps);
    float* y = out + (size_t)r * ld;
    for (int j = lane; j < d; j += 32) y[j] = (x[j] - mean) * rstd * __ldg(nm.g + j) + __ldg(nm.b + j);
  }
  __syncthreads();
}

// softmax(q k^T / sqrt(hd)) v over the S rows of each window; q at wide[:, 0:d), k at [d:2d),
// v at [2d:3d); the context overwrites the q slot.  One warp per (row, head).
__device__ __forceinline__ void attention(float* wide, int ldw, int d, int heads, int S, int R) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hd = d / heads;
  const float scale = rsqrtf((float)hd);
  for (int it = warp; it < R * heads; it += kWarps) {
    const int r = it / heads, h = it % heads;
    const int w0 = (r / S) * S;                       // first row of this window
    float* q = wide + (size_t)r * ldw + h * hd;
    float sc[kSMax];
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < kSMax; ++s) {
      if (s < S) {
        const float* k = wide + (size_t)(w0 + s) * ldw + d + h * hd;
        float a = 0.f;
        for (int j = lane; j < hd; j += 32) a = fmaf(q[j], k[j], a);
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        sc[s] = a * scale;
        mx = fmaxf(mx, sc[s]);
      }
    }
    float den = 0.f;
#pragma unroll
    for (int s = 0; s < kSMax; ++s)
      if (s < S) {
        sc[s] = expf(sc[s] - mx);
        den += sc[s];
      }
    const float inv = 1.f / den;
    __syncwarp();
    for (int j = lane; j < hd; j += 32) {
      float a = 0.f;
#pragma unroll
      for (int s = 0; s < kSMax; ++s)
        if (s < S) a = fmaf(sc[s], wide[(size_t)(w0 + s) * ldw + 2 * d + h * hd + j], a);
      q[j] = a * inv;
    }
  }
  __syncthreads();
}

struct XfGeom {
  int S, R_tile, win_per_tile, ld, ldw;     // ld = d_model row stride, ldw = wide row stride
  int off_x, off_y, off_t, off_w, total;    // float offsets
};

// self-attention sub-block on stream `x` (normalised input `h`, may alias x): x += out_proj(attn(h))
__device__ __forceinline__ void self_attn_block(const Attn& at, const float* h, float* x, float* wide, const XfGeom& g,
                                                int d, int heads, int R) {
  linear(h, g.ld, at.qkv.wt, 3 * d, at.qkv.b, d, 3 * d, wide, g.ldw, R, kActNone, false);
  attention(wide, g.ldw, d, heads, g.S, R);
  linear(wide, g.ldw, at.out.wt, d, at.out.b, d, d, x, g.ld, R, kActNone, true);
}

// cross-attention: queries from `h`, keys/values from `mem`
__device__ __forceinline__ void cross_attn_block(const Attn& at, const float* h, const float* mem, float* x, float* wide,
                                                 const XfGeom& g, int d, int heads, int R) {
  linear(h, g.ld, at.qkv.wt, 3 * d, at.qkv.b, d, d, wide, g.ldw, R, kActNone, false);
  linear(mem, g.ld, at.qkv.wt + d, 3 * d, at.qkv.b + d, d, 2 * d, wide + d, g.ldw, R, kActNone, false);
  attention(wide, g.ldw, d, heads, g.S, R);
  linear(wide, g.ldw, at.out.wt, d, at.out.b, d, d, x, g.ld, R, kActNone, true);
}

__device__ __forceinline__ void ffn_block(const Linear& f1, const Linear& f2, const float* h, float* x, float* wide,
                                          const XfGeom& g, int d, int dff, int act, int R) {
  linear(h, g.ld, f1.wt, dff, f1.b, d, dff, wide, g.ldw, R, act, false);
  linear(wide, g.ldw, f2.wt, d, f2.b, dff, d, x, g.ld, R, kActNone, true);
}

__global__ void __launch_bounds__(kThreads, 1)
transformer_fp32_kernel(const __grid_constant__ Transformer xf, const __grid_constant__ XfGeom g,
                        const float* __restrict__ tokens, int64_t B, int reduction, float* __restrict__ recon_out,
                        float* __restrict__ scores) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem + g.off_x;
  float* ys = smem + g.off_y;
  float* ts = smem + g.off_t;
  float* wide = smem + g.off_w;
  const int d = xf.d_model, dt = xf.d_tok, S = g.S, H = xf.heads, dff = xf.d_ff;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_tiles = (B + g.win_per_tile - 1) / g.win_per_tile;

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t w_first = tile * g.win_per_tile;
    const int64_t left = B - w_first;
    const int nw = left < (int64_t)g.win_per_tile ? (int)left : g.win_per_tile;
    const int R = nw * S;
    const float* tok = tokens + (size_t)w_first * S * dt;

    if (xf.variant == SF_VARIANT_SHOPFORMER) {
      // ---------------- variant 1: post-LN
      for (int i = threadIdx.x; i < R * d; i += kThreads) {
        const int r = i / d, j = i % d, s = r % S;
        const float pe = __ldg(xf.pe + s * d + j);
        xs[r * g.ld + j] = __ldg(tok + (size_t)r * dt + j) + pe;
        ys[r * g.ld + j] = (s == 0 ? 0.f : __ldg(tok + (size_t)(r - 1) * dt + j)) + pe;   // zero start token + shift
      }
      __syncthreads();
      for (int l = 0; l < xf.n_enc; ++l) {
        const EncLayer& L = xf.enc[l];
        self_attn_block(L.sa, xs, xs, wide, g, d, H, R);
        layernorm_rows(xs, xs, g.ld, d, R, L.n1);
        ffn_block(L.ff1, L.ff2, xs, xs, wide, g, d, dff, kActRelu, R);
        layernorm_rows(xs, xs, g.ld, d, R, L.n2);
      }
      for (int l = 0; l < xf.n_dec; ++l) {
        const DecLayer& L = xf.dec[l];
        self_attn_block(L.sa, ys, ys, wide, g, d, H, R);
        layernorm_rows(ys, ys, g.ld, d, R, L.n1);
        cross_attn_block(L.ca, ys, xs, ys, wide, g, d, H, R);
        layernorm_rows(ys, ys, g.ld, d, R, L.n2);
        ffn_block(L.ff1, L.ff2, ys, ys, wide, g, d, dff, kActRelu, R);
        layernorm_rows(ys, ys, g.ld, d, R, L.n3);
      }
      linear(ys, g.ld, xf.out_proj.wt, d, xf.out_proj.b, d, d, ts, g.ld, R, kActNone, false);
    } else {
      // ---------------- variant 2: pre-LN, GELU, final norms, optional projections
      if (xf.has_io_proj) {
        for (int i = threadIdx.x; i < R * dt; i += kThreads) wide[(i / dt) * g.ldw + (i % dt)] = __ldg(tok + i);
        __syncthreads();
        linear(wide, g.ldw, xf.in_proj.wt, d, xf.in_proj.b, dt, d, xs, g.ld, R, kActNone, false);
      } else {
        for (int i = threadIdx.x; i < R * d; i += kThreads) xs[(i / d) * g.ld + (i % d)] = __ldg(tok + i);
        __syncthreads();
      }
      for (int i = threadIdx.x; i < R * d; i += kThreads) {
        const int r = i / d, j = i % d;
        const float v = xs[r * g.ld + j] + __ldg(xf.pe + (r % S) * d + j);
        xs[r * g.ld + j] = v;
        ys[r * g.ld + j] = v;                         // decoder target = encoder input
      }
      __syncthreads();
      for (int l = 0; l < xf.n_enc; ++l) {
        const EncLayer& L = xf.enc[l];
        layernorm_rows(xs, ts, g.ld, d, R, L.n1);
        self_attn_block(L.sa, ts, xs, wide, g, d, H, R);
        layernorm_rows(xs, ts, g.ld, d, R, L.n2);
        ffn_block(L.ff1, L.ff2, ts, xs, wide, g, d, dff, kActGelu, R);
      }
      layernorm_rows(xs, xs, g.ld, d, R, xf.enc_norm);   // memory
      for (int l = 0; l < xf.n_dec; ++l) {
        const DecLayer& L = xf.dec[l];
        layernorm_rows(ys, ts, g.ld, d, R, L.n1);
        self_attn_block(L.sa, ts, ys, wide, g, d, H, R);
        layernorm_rows(ys, ts, g.ld, d, R, L.n2);
        cross_attn_block(L.ca, ts, xs, ys, wide, g, d, H, R);
        layernorm_rows(ys, ts, g.ld, d, R, L.n3);
        ffn_block(L.ff1, L.ff2, ts, ys, wide, g, d, dff, kActGelu, R);
      }
      layernorm_rows(ys, ts, g.ld, d, R, xf.dec_norm);
      if (xf.has_io_proj) {
        linear(ts, g.ld, xf.out_proj.wt, dt, xf.out_proj.b, d, dt, wide, g.ldw, R, kActNone, false);
      }
    }
    // reconstruction now at `rec` with row stride `ldr`, width dt
    const float* rec = (xf.variant == SF_VARIANT_SHOPFORMER_2 && xf.has_io_proj) ? wide : ts;
    const int ldr = (xf.variant == SF_VARIANT_SHOPFORMER_2 && xf.has_io_proj) ? g.ldw : g.ld;
    if (recon_out) {
      float* dst = recon_out + (size_t)w_first * S * dt;
      for (int i = threadIdx.x; i < R * dt; i += kThreads) dst[i] = rec[(i / dt) * ldr + (i % dt)];
    }
    if (scores) {
      // fused score: one warp per window (mean) or per row (none)
      const bool per_row = reduction == SF_REDUCE_NONE;
      const int groups = per_row ? R : nw;
      const int rows_per = per_row ? 1 : S;
      for (int gi = warp; gi < groups; gi += kWarps) {
        float a = 0.f;
        for (int i = lane; i < rows_per * dt; i += 32) {
          const int r = gi * rows_per + i / dt, j = i % dt;
          float target = __ldg(tok + (size_t)r * dt + j);
          if (xf.variant == SF_VARIANT_SHOPFORMER) target += __ldg(xf.pe_score + (r % S) * dt + j);
          const float df = rec[r * ldr + j] - target;
          a = fmaf(df, df, a);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
          const int64_t idx = per_row ? (w_first * S + gi) : (w_first + gi);
          scores[idx] = a / (float)(rows_per * dt);
        }
      }
    }
    __syncthreads();
  }
}

// stand-alone MSE score: HBM-bound, one warp per output element, float4 loads
__global__ void __launch_bounds__(256)
score_kernel(const float* __restrict__ tokens, const float* __restrict__ recon, const float* __restrict__ pe_score,
             int64_t n_out, int n_per, int S, int dt, int per_row, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int n4 = n_per >> 2;
  for (int64_t o = warp0; o < n_out; o += n_warps) {
    const float4* t4 = reinterpret_cast<const float4*>(tokens + o * n_per);
    const float4* r4 = reinterpret_cast<const float4*>(recon + o * n_per);
    float a = 0.f;
    for (int i = lane; i < n4; i += 32) {
      float4 t = __ldcs(t4 + i);
      const float4 r = __ldcs(r4 + i);
      if (pe_score) {
        // position of element 4*i inside the (S, dt) block of its window
        const int e = per_row ? (int)((o % S) * dt) + 4 * i : 4 * i;
        const float4 p = __ldg(reinterpret_cast<const float4*>(pe_score + e));
        t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
      }
      const float dx = r.x - t.x, dy = r.y - t.y, dz = r.z - t.z, dw = r.w - t.w;
      a = fmaf(dx, dx, a); a = fmaf(dy, dy, a); a = fmaf(dz, dz, a); a = fmaf(dw, dw, a);
    }
#pragma unroll
    for (int s = 16; s; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) scores[o] = a / (float)n_per;
  }
}

}  // namespace

int launch_transformer_fp32(const sf_model* m, const float* tokens, int64_t B, int S, int reduction, float* recon,
                            float* scores, cudaStream_t st) {
  if (B == 0) return SF_OK;
  const Transformer& xf = m->xf;
  SF_REQUIRE(S >= 1 && S <= kSMax, SF_E_UNSUPPORTED, "token count S=%d outside [1,%d]", S, kSMax);
  count_launch(LK_XF_FP32);
  SF_REQUIRE(S <= 100, SF_E_INVALID, "S=%d exceeds the positional-encoding table (100)", S);
  SF_REQUIRE(reduction == SF_REDUCE_MEAN || (reduction == SF_REDUCE_NONE && xf.variant == SF_VARIANT_SHOPFORMER_2),
             SF_E_INVALID, "reduction %d not available for variant %d", reduction, xf.variant);
  XfGeom g;
  g.S = S;
  g.win_per_tile = std::max(1, kRowsMax / S);
  g.R_tile = g.win_per_tile * S;
  g.ld = xf.d_model;
  g.ldw = std::max(std::max(3 * xf.d_model, xf.d_ff), xf.d_tok);
  g.off_x = 0;
  g.off_y = g.off_x + g.R_tile * g.ld;
  g.off_t = g.off_y + g.R_tile * g.ld;
  g.off_w = g.off_t + g.R_tile * g.ld;
  g.total = g.off_w + g.R_tile * g.ldw;
  const size_t smem = (size_t)g.total * sizeof(float);
  SF_REQUIRE(smem <= (size_t)m->max_smem_optin, SF_E_UNSUPPORTED,
             "transformer tile needs %zu bytes of shared memory (d_model=%d, d_ff=%d), device allows %d", smem, xf.d_model,
             xf.d_ff, m->max_smem_optin);
  SF_CUDA_OK(cudaFuncSetAttribute(transformer_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (B + g.win_per_tile - 1) / g.win_per_tile;
  const int grid = (int)std::min<int64_t>(n_tiles, m->sm_count);
  transformer_fp32_kernel<<<grid, kThreads, smem, st>>>(xf, g, tokens, B, reduction, recon, scores);
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}

int launch_score(const sf_model* m, const float* tokens, const float* recon, int64_t B, int S, int reduction,
                 float* scores, cudaStream_t st) {
  if (B == 0) return SF_OK;
  const Transformer& xf = m->xf;
  SF_REQUIRE(reduction == SF_REDUCE_MEAN || (reduction == SF_REDUCE_NONE && xf.variant == SF_VARIANT_SHOPFORMER_2),
             SF_E_INVALID, "reduction %d not available for variant %d", reduction, xf.variant);
  const int per_row = reduction == SF_REDUCE_NONE;
  const int64_t n_out = per_row ? B * S : B;
  const int n_per = per_row ? xf.d_tok : S * xf.d_tok;
  const int64_t warps_needed = n_out;
  const int64_t blocks = std::min<int64_t>((warps_needed + 7) / 8, (int64_t)m->sm_count * 8);
  score_kernel<<<(int)std::max<int64_t>(blocks, 1), 256, 0, st>>>(
      tokens, recon, xf.variant == SF_VARIANT_SHOPFORMER ? xf.pe_score : nullptr, n_out, n_per, S, xf.d_tok, per_row, scores);
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}

}  // namespace sf
