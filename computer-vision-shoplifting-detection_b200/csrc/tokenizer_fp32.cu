// fp32 ST-GCN pose tokenizer (CUDA cores): the precise path (<=1e-3 vs the reference).
//
// Maths (eval mode, BatchNorms folded at pack time) -- SURVEY Appendix A.1,
// reference shopformer/models/gcae.py:124-154 (graph conv), :185-195 (temporal conv + BN),
// :242-259 (block), :331-366 (encoder); shopformer_2/models/gcae.py:375-422.
//
// One CTA owns one window at a time (persistent, grid = SMs x occupancy) and keeps every
// activation of that window on chip:
//   bufA / bufB   ping-pong block inputs/outputs, layout [c][t][v]
//   bufG          relu(gcn) output with a 4-row zero halo either side of t, [c][4+t+..][v],
//                 so the 9-tap temporal conv needs no bounds checks
// Block 0 (Cin = 2 or 3) never materialises its graph-conv output: each thread rebuilds the
// 9+ rows it needs from the adjacency-mixed input held in registers (2 FMAs + max per value).
// Later blocks: residual -> Y, adjacency mix of X in place with warp shuffles over the
// keypoint axis (one warp per (c,t) row, keypoints on lanes), X.W register-tiled GEMM into G,
// then the temporal conv as a register-tiled sliding window (TP outputs x OT channels per
// thread; consecutive outputs share taps so a thread loads S*(TP-1)+9 inputs for 9*TP*OT FMAs).
#include <algorithm>

#include "sf_internal.h"

namespace sf {
namespace {

constexpr int kThreads = 256;

struct TokGeom {
  int T[kMaxBlocks + 1];      // temporal length before block i (T[n_blocks] = tokenizer output length)
  int g_rows[kMaxBlocks];     // rows per channel plane of bufG for block i
  int tp[kMaxBlocks];         // tile choice per block: outputs per thread along t'
  int ot[kMaxBlocks];         //                         output channels per thread
  int offA, offB, offG;       // float offsets of the three buffers inside the scratch slab
  int slab;                   // floats per CTA
  int S_out;                  // tokens written per window (after optional pooling)
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---------------------------------------------------------------------------------------
// temporal conv tile.  kFusedCin > 0: inputs are the adjacency-mixed block input (planes
// [kFusedCin][rows][V], halo included) and the graph conv is rebuilt on the fly.
// kFusedCin == 0: inputs are read from the materialised G buffer.
template <int S, int TP, int OT, int kFusedCin>
__device__ __noinline__ void tcn_items(const TokBlock& bk, const float* __restrict__ gin_buf, int g_rows,
                                          const float* __restrict__ X, float* __restrict__ Y, int T_in, int T_out,
                                          int V, int stride_rt, bool y_has_residual) {
  // S == 0 means "runtime stride" and is only instantiated with TP == 1.
  const int s = (S == 0) ? stride_rt : S;
  constexpr int NIN = (S == 0 ? 0 : S * (TP - 1)) + kTaps;
  const int cg_n = bk.cout;                 // temporal conv is cout -> cout
  const int n_og = bk.cout / OT;
  const int n_tg = (T_out + TP - 1) / TP;
  const int items = n_og * n_tg * V;
  const int planeG = g_rows * V;
  const int planeX = T_in * V;
  const int planeY = T_out * V;
  for (int it = threadIdx.x; it < items; it += kThreads) {
    const int v = it % V;
    const int tg = (it / V) % n_tg;
    const int og = it / (V * n_tg);
    const int t0 = tg * TP;                 // first output row of this tile
    const int r0 = s * t0;                  // first input row in halo coordinates (t = r - 4)
    float acc[TP][OT];
#pragma unroll
    for (int a = 0; a < TP; ++a)
#pragma unroll
      for (int o = 0; o < OT; ++o) acc[a][o] = 0.f;

    if constexpr (kFusedCin > 0) {
      float m[kFusedCin][NIN];
      unsigned inr = 0;
#pragma unroll
      for (int j = 0; j < NIN; ++j) {
        const int t = r0 + j - kHalo;
        if (t >= 0 && t < T_in) inr |= 1u << j;
#pragma unroll
        for (int ci = 0; ci < kFusedCin; ++ci) m[ci][j] = gin_buf[ci * planeG + (r0 + j) * V + v];
      }
      for (int cg = 0; cg < cg_n; ++cg) {
        float gw[kFusedCin];
#pragma unroll
        for (int ci = 0; ci < kFusedCin; ++ci) gw[ci] = __ldg(bk.gcn_w + ci * cg_n + cg);
        const float gb = __ldg(bk.gcn_b + cg);
        float gin[NIN];
#pragma unroll
        for (int j = 0; j < NIN; ++j) {
          float a = gb;
#pragma unroll
          for (int ci = 0; ci < kFusedCin; ++ci) a = fmaf(gw[ci], m[ci][j], a);
          gin[j] = ((inr >> j) & 1u) ? fmaxf(a, 0.f) : 0.f;
        }
        const float* wp = bk.tcn_w + (size_t)cg * kTaps * bk.cout + og * OT;
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
          float w[OT];
#pragma unroll
          for (int q = 0; q < OT / 4; ++q) {
            const float4 w4 = ldg4(wp + k * bk.cout + q * 4);
            w[q * 4 + 0] = w4.x; w[q * 4 + 1] = w4.y; w[q * 4 + 2] = w4.z; w[q * 4 + 3] = w4.w;
          }
#pragma unroll
          for (int a = 0; a < TP; ++a)
#pragma unroll
            for (int o = 0; o < OT; ++o) acc[a][o] = fmaf(gin[(S == 0 ? 0 : S * a) + k], w[o], acc[a][o]);
        }
      }
    } else {
      for (int cg = 0; cg < cg_n; ++cg) {
        const float* gp = gin_buf + cg * planeG + r0 * V + v;
        float gin[NIN];
#pragma unroll
        for (int j = 0; j < NIN; ++j) gin[j] = gp[j * V];
        const float* wp = bk.tcn_w + (size_t)cg * kTaps * bk.cout + og * OT;
#pragma unroll
        for (int k = 0; k < kTaps; ++k) {
          float w[OT];
#pragma unroll
          for (int q = 0; q < OT / 4; ++q) {
            const float4 w4 = ldg4(wp + k * bk.cout + q * 4);
            w[q * 4 + 0] = w4.x; w[q * 4 + 1] = w4.y; w[q * 4 + 2] = w4.z; w[q * 4 + 3] = w4.w;
          }
#pragma unroll
          for (int a = 0; a < TP; ++a)
#pragma unroll
            for (int o = 0; o < OT; ++o) acc[a][o] = fmaf(gin[(S == 0 ? 0 : S * a) + k], w[o], acc[a][o]);
        }
      }
    }

    // epilogue: folded bias + residual, ReLU
#pragma unroll
    for (int a = 0; a < TP; ++a) {
      const int t = t0 + a;
      if (t >= T_out) continue;
#pragma unroll
      for (int o = 0; o < OT; ++o) {
        const int oc = og * OT + o;
        float r;
        if (y_has_residual) {
          r = Y[oc * planeY + t * V + v];       // bias + residual already there
        } else {
          r = __ldg(bk.out_b + oc);
          if (bk.identity_res) {
            r += X[oc * planeX + t * V + v];
          } else {
            for (int ci = 0; ci < bk.cin; ++ci) r = fmaf(__ldg(bk.res_w + ci * bk.cout + oc), X[ci * planeX + (s * t) * V + v], r);
          }
        }
        Y[oc * planeY + t * V + v] = fmaxf(acc[a][o] + r, 0.f);
      }
    }
  }
}

template <int S, int kFusedCin>
__device__ __forceinline__ void tcn_dispatch(const TokBlock& bk, int tp, int ot, const float* gin_buf, int g_rows,
                                             const float* X, float* Y, int T_in, int T_out, int V, bool y_res) {
  if constexpr (S == 0) {
    if (ot == 8) tcn_items<0, 1, 8, kFusedCin>(bk, gin_buf, g_rows, X, Y, T_in, T_out, V, bk.stride, y_res);
    else tcn_items<0, 1, 4, kFusedCin>(bk, gin_buf, g_rows, X, Y, T_in, T_out, V, bk.stride, y_res);
  } else {
    if (tp == 4 && ot == 8) tcn_items<S, 4, 8, kFusedCin>(bk, gin_buf, g_rows, X, Y, T_in, T_out, V, S, y_res);
    else if (tp == 2 && ot == 8) tcn_items<S, 2, 8, kFusedCin>(bk, gin_buf, g_rows, X, Y, T_in, T_out, V, S, y_res);
    else if (tp == 2 && ot == 4) tcn_items<S, 2, 4, kFusedCin>(bk, gin_buf, g_rows, X, Y, T_in, T_out, V, S, y_res);
    else tcn_items<S, 1, 4, kFusedCin>(bk, gin_buf, g_rows, X, Y, T_in, T_out, V, S, y_res);
  }
}

template <int kFusedCin>
__device__ __forceinline__ void tcn_any_stride(const TokBlock& bk, int tp, int ot, const float* gin_buf, int g_rows,
                                               const float* X, float* Y, int T_in, int T_out, int V, bool y_res) {
  if (tp == 1 || bk.stride > 3) {
    tcn_dispatch<0, kFusedCin>(bk, tp, ot, gin_buf, g_rows, X, Y, T_in, T_out, V, y_res);
  } else if (bk.stride == 1) {
    tcn_dispatch<1, kFusedCin>(bk, tp, ot, gin_buf, g_rows, X, Y, T_in, T_out, V, y_res);
  } else if (bk.stride == 2) {
    tcn_dispatch<2, kFusedCin>(bk, tp, ot, gin_buf, g_rows, X, Y, T_in, T_out, V, y_res);
  } else {
    tcn_dispatch<3, kFusedCin>(bk, tp, ot, gin_buf, g_rows, X, Y, T_in, T_out, V, y_res);
  }
}

// adjacency mix of one value held on lane v: sum_e val[v][e] * x[col[v][e]] via warp shuffles
__device__ __forceinline__ float ell_mix_shfl(const TokBlock& bk, int v, int V, float x) {
  float a = 0.f;
  const int w = bk.ell_width;
  for (int e = 0; e < w; ++e) {
    const int col = (v < V) ? __ldg(bk.ell_col + v * w + e) : 0;
    const float val = (v < V) ? __ldg(bk.ell_val + v * w + e) : 0.f;
    a = fmaf(val, __shfl_sync(0xffffffffu, x, col), a);
  }
  return a;
}

template <bool kGlobalScratch>
__global__ void __launch_bounds__(kThreads, 2)
tokenizer_fp32_kernel(const __grid_constant__ Tokenizer tk, const __grid_constant__ TokGeom geo,
                      const float* __restrict__ poses, float* __restrict__ tokens, int64_t B, float* gscratch) {
  extern __shared__ __align__(16) float smem[];
  float* base = kGlobalScratch ? gscratch + (size_t)blockIdx.x * geo.slab : smem;
  float* bufA = base + geo.offA;
  float* bufB = base + geo.offB;
  float* bufG = base + geo.offG;
  const int V = tk.V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kThreads / 32;

  for (int64_t w = blockIdx.x; w < B; w += gridDim.x) {
    // ---- E0: load + folded BatchNorm1d (index c*V+v)
    const int T0 = geo.T[0];
    {
      const float* src = poses + (size_t)w * tk.c_in * T0 * V;
      const int n = tk.c_in * T0 * V;
      for (int i = threadIdx.x; i < n; i += kThreads) {
        const int v = i % V, c = i / (T0 * V);
        const int j = c * V + v;
        bufB[i] = fmaf(__ldg(src + i), __ldg(tk.in_scale + j), __ldg(tk.in_shift + j));
      }
    }
    __syncthreads();
    float* X = bufB;
    float* Y = bufA;
    for (int b = 0; b < tk.n_blocks; ++b) {
      const TokBlock& bk = tk.blk[b];
      const int T_in = geo.T[b], T_out = geo.T[b + 1], g_rows = geo.g_rows[b];
      if (bk.cin == 2 || bk.cin == 3) {
        // ---- fused block: G <- adjacency-mixed X with zero halo, then conv rebuilds relu(gcn) on the fly
        const int n = bk.cin * g_rows * V;
        for (int i = threadIdx.x; i < n; i += kThreads) {
          const int v = i % V, r = (i / V) % g_rows, c = i / (V * g_rows);
          const int t = r - kHalo;
          float a = 0.f;
          if (t >= 0 && t < T_in) {
            const float* xr = X + c * T_in * V + t * V;
            for (int e = 0; e < bk.ell_width; ++e)
              a = fmaf(__ldg(bk.ell_val + v * bk.ell_width + e), xr[__ldg(bk.ell_col + v * bk.ell_width + e)], a);
          }
          bufG[i] = a;
        }
        __syncthreads();
        if (bk.cin == 2) tcn_any_stride<2>(bk, geo.tp[b], geo.ot[b], bufG, g_rows, X, Y, T_in, T_out, V, false);
        else tcn_any_stride<3>(bk, geo.tp[b], geo.ot[b], bufG, g_rows, X, Y, T_in, T_out, V, false);
      } else {
        // ---- (1) Y <- folded bias + residual(X)
        {
          const int n = bk.cout * T_out * V;
          const int planeX = T_in * V;
          for (int i = threadIdx.x; i < n; i += kThreads) {
            const int v = i % V, t = (i / V) % T_out, o = i / (V * T_out);
            float r = __ldg(bk.out_b + o);
            if (bk.identity_res) {
              r += X[o * planeX + t * V + v];
            } else {
              const float* xp = X + (bk.stride * t) * V + v;
#pragma unroll 4
              for (int c = 0; c < bk.cin; ++c) r = fmaf(__ldg(bk.res_w + c * bk.cout + o), xp[c * planeX], r);
            }
            Y[i] = r;
          }
        }
        __syncthreads();
        // ---- (2) X <- A_hat . X in place; keypoints on lanes, one warp per (c,t) row
        {
          const int rows = bk.cin * T_in;
          for (int r = warp; r < rows; r += kWarps) {
            float x = (lane < V) ? X[r * V + lane] : 0.f;
            x = ell_mix_shfl(bk, lane, V, x);
            if (lane < V) X[r * V + lane] = x;
          }
        }
        // ---- (3a) zero the halo / slack rows of G
        {
          const int lo = kHalo * V, hi = (kHalo + T_in) * V, plane = g_rows * V;
          const int n = bk.cout * plane;
          for (int i = threadIdx.x; i < n; i += kThreads) {
            const int p = i % plane;
            if (p < lo || p >= hi) bufG[i] = 0.f;
          }
        }
        __syncthreads();
        // ---- (3b) G[o][4+t][v] = relu(b[o] + sum_c X[c][t][v] W[c][o]); 4 positions x 8 channels per thread
        {
          const int P = T_in * V;                 // positions are contiguous inside a channel plane
          const int n_pg = (P + 3) / 4;
          const int plane = g_rows * V;
          const bool ot8 = (bk.cout % 8) == 0;
          const int OTg = ot8 ? 8 : 4;
          const int n_og = bk.cout / OTg;
          const int items = n_og * n_pg;
          for (int it = threadIdx.x; it < items; it += kThreads) {
            const int pg = it % n_pg, og = it / n_pg;
            const int p0 = pg * 4;
            float acc[4][8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
              for (int o = 0; o < 8; ++o) acc[a][o] = 0.f;
            for (int c = 0; c < bk.cin; ++c) {
              float xv[4];
#pragma unroll
              for (int a = 0; a < 4; ++a) xv[a] = (p0 + a < P) ? X[c * P + p0 + a] : 0.f;
              float wv[8];
              const float4 w0 = ldg4(bk.gcn_w + c * bk.cout + og * OTg);
              wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w;
              if (ot8) {
                const float4 w1 = ldg4(bk.gcn_w + c * bk.cout + og * OTg + 4);
                wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
              } else {
                wv[4] = wv[5] = wv[6] = wv[7] = 0.f;
              }
#pragma unroll
              for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int o = 0; o < 8; ++o) acc[a][o] = fmaf(xv[a], wv[o], acc[a][o]);
            }
#pragma unroll
            for (int o = 0; o < 8; ++o) {
              if (o >= OTg) break;
              const int oc = og * OTg + o;
              const float bb = __ldg(bk.gcn_b + oc);
#pragma unroll
              for (int a = 0; a < 4; ++a)
                if (p0 + a < P) bufG[oc * plane + kHalo * V + p0 + a] = fmaxf(acc[a][o] + bb, 0.f);
            }
          }
        }
        __syncthreads();
        // ---- (4) Y <- relu(Y + tcn(G))
        tcn_any_stride<0>(bk, geo.tp[b], geo.ot[b], bufG, g_rows, X, Y, T_in, T_out, V, true);
      }
      __syncthreads();
      float* t = X; X = Y; Y = t;
    }
    // ---- E5: tokens[w][t'][c*V+v] = X[c][t'][v]  (optionally adaptive-avg-pooled over t)
    {
      const int C = tk.blk[tk.n_blocks - 1].cout;
      const int Tl = geo.T[tk.n_blocks];
      const int S = geo.S_out;
      float* dst = tokens + (size_t)w * S * C * V;
      const int n = S * C * V;
      for (int i = threadIdx.x; i < n; i += kThreads) {
        const int v = i % V, c = (i / V) % C, s = i / (V * C);
        float val;
        if (tk.pool_tokens > 0) {
          const int lo = (s * Tl) / S, hi = ((s + 1) * Tl + S - 1) / S;   // AdaptiveAvgPool window
          float a = 0.f;
          for (int t = lo; t < hi; ++t) a += X[c * Tl * V + t * V + v];
          val = a / (float)(hi - lo);
        } else {
          val = X[c * Tl * V + s * V + v];
        }
        dst[i] = val;
      }
    }
    __syncthreads();
  }
}

int build_geom(const sf_model* m, int T, TokGeom* g) {
  const Tokenizer& tk = m->tok;
  const int V = tk.V;
  g->T[0] = T;
  size_t maxA = 0, maxB = 0, maxG = 0;
  maxB = (size_t)tk.c_in * T * V;
  for (int b = 0; b < tk.n_blocks; ++b) {
    const TokBlock& bk = tk.blk[b];
    const int T_in = g->T[b];
    const int T_out = (T_in - 1) / bk.stride + 1;
    g->T[b + 1] = T_out;
    // tile choice: biggest register tile that still gives ~one item per thread
    const int cand[4][2] = {{4, 8}, {2, 8}, {2, 4}, {1, 4}};
    int tp = 1, ot = 4;
    for (auto& c : cand) {
      if (bk.cout % c[1]) continue;
      const int items = (bk.cout / c[1]) * ((T_out + c[0] - 1) / c[0]) * V;
      if (items >= (kThreads * 3) / 4 || (c[0] == 1 && c[1] == 4)) {
        tp = c[0];
        ot = c[1];
        break;
      }
    }
    if (bk.stride > 3) tp = 1;
    g->tp[b] = tp;
    g->ot[b] = ot;
    const int t_tile = ((T_out + tp - 1) / tp) * tp;
    const int rows = std::max(T_in + 2 * kHalo, bk.stride * (t_tile - 1) + kTaps);
    g->g_rows[b] = rows;
    const size_t gsz = (size_t)((bk.cin == 2 || bk.cin == 3) ? bk.cin : bk.cout) * rows * V;
    maxG = std::max(maxG, gsz);
    const size_t osz = (size_t)bk.cout * T_out * V;
    if (b % 2 == 0) maxA = std::max(maxA, osz); else maxB = std::max(maxB, osz);
  }
  auto up4 = [](size_t x) { return (x + 3) & ~size_t(3); };
  g->offA = 0;
  g->offB = (int)up4(maxA);
  g->offG = g->offB + (int)up4(maxB);
  g->slab = g->offG + (int)up4(maxG);
  g->S_out = tk.pool_tokens > 0 ? tk.pool_tokens : g->T[tk.n_blocks];
  return SF_OK;
}

}  // namespace

int64_t tokenizer_fp32_workspace(const sf_model* m, int64_t B, int T) {
  TokGeom g;
  build_geom(m, T, &g);
  const size_t smem = (size_t)g.slab * sizeof(float);
  if (smem <= (size_t)m->max_smem_optin) return 0;
  const int64_t grid = std::min<int64_t>(B, (int64_t)m->sm_count * 2);
  return grid * (int64_t)smem;
}

int launch_tokenizer_fp32(const sf_model* m, const float* poses, int64_t B, int T, float* tokens, void* ws,
                          int64_t ws_bytes, cudaStream_t st) {
  if (B == 0) return SF_OK;
  count_launch(LK_TOK_FP32);
  TokGeom g;
  build_geom(m, T, &g);
  const size_t smem = (size_t)g.slab * sizeof(float);
  if (smem <= (size_t)m->max_smem_optin) {
    static thread_local size_t configured = 0;
    if (smem > configured) {
      SF_CUDA_OK(cudaFuncSetAttribute(tokenizer_fp32_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      m->max_smem_optin));
      configured = m->max_smem_optin;
    }
    int occ = 1;
    SF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tokenizer_fp32_kernel<false>, kThreads, smem));
    occ = std::max(occ, 1);
    const int grid = (int)std::min<int64_t>(B, (int64_t)m->sm_count * occ);
    tokenizer_fp32_kernel<false><<<grid, kThreads, smem, st>>>(m->tok, g, poses, tokens, B, nullptr);
  } else {
    const int64_t need = tokenizer_fp32_workspace(m, B, T);
    SF_REQUIRE(ws && ws_bytes >= need, SF_E_INVALID,
               "tokenizer needs %lld bytes of workspace for this config (per-window activations exceed shared memory), got %lld",
               (long long)need, (long long)ws_bytes);
    const int grid = (int)std::min<int64_t>(B, (int64_t)m->sm_count * 2);
    tokenizer_fp32_kernel<true><<<grid, kThreads, 0, st>>>(m->tok, g, poses, tokens, B, (float*)ws);
  }
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}

}  // namespace sf
