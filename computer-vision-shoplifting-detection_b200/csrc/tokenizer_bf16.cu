// bf16 tcgen05 ST-GCN pose tokenizer: the fast path (tolerance 1e-2 on the score, north star).
//
// Same maths as tokenizer_fp32.cu (reference shopformer/models/gcae.py:124-154,185-195,
// 242-259,331-366 and shopformer_2/models/gcae.py:375-422), mapped onto the 5th-gen tensor cores:
//
//   * activations live in shared memory as bf16 in the planar-chunk layout of tc_common.cuh,
//     rows = (window, time, keypoint), columns = channels; accumulators live in TMEM (fp32);
//   * the 9x1 temporal convolution is an implicit GEMM: the input of a stride-s block is stored split
//     into s time phases, so tap k is a SHIFTED VIEW (row offset o_k * V) of phase (k-4) mod s -- nine
//     smem descriptors into one buffer, no im2col, zero rows ("gaps") between windows give the padding;
//     taps that can only ever hit padding are skipped;
//   * the BN-folded 1x1 strided residual conv is two more MMAs into the same accumulator (phase 0 of x);
//   * graph conv = adjacency mix over keypoints on CUDA cores (bf16 rows in smem) followed by a
//     [rows x Cin] x [Cin x Cout] tensor-core GEMM; block 0 (Cin = 2) is rebuilt on CUDA cores in fp32;
//   * epilogues read TMEM with tcgen05.ld, add the folded bias, ReLU, and write the next operand
//     straight back to shared memory as bf16 -- nothing but the final tokens (fp32) goes to HBM.
//
// One CTA owns G windows at a time (persistent); phases are separated by mbarrier (MMA completion) and
// __syncthreads; two CTAs per SM overlap one CTA's MMAs with the other's CUDA-core phases.
#include <algorithm>
#include <cstdlib>

#include "sf_internal.h"
#include "tc_common.cuh"

namespace sf {
namespace {

using namespace tc;

constexpr int kThreads = 256;

struct BfBlk {
  int cin, kin, npad, cout;            // real / padded input channels, padded / real output channels
  int stride, Tin, Tout, gap, slot;    // slot = Tout + gap time rows per window in a phase buffer
  int rows, mrows, rtot;               // rows per phase, M-space rows (G*slot*V), total rows (stride*rows)
  int n_taps;
  int tap_k[kTaps], tap_phase[kTaps], tap_rowoff[kTaps];
  int identity_res, ell_width;
  const float* ell_val;
  const int* ell_col;
  const uint16_t *w_tcn, *w_gcn, *w_res;
  const float *gcn_b, *out_b;          // [npad], zero padded
  const float *gcn_w32, *res_w32;      // block 0 only: fp32 [cin][cout]
};

struct BfPlan {
  int n_blocks, V, c_in, G, T0, S_out, c_last;
  const float *in_scale, *in_shift;
  BfBlk blk[kMaxBlocks];
  uint32_t off_A, off_X0, off_X1, off_WT, off_WG, off_x0, off_m0, smem_bytes, tmem_cols;
};
static_assert(sizeof(BfPlan) <= 3900, "BfPlan must fit in kernel parameter space");

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void stage(unsigned char* dst, const uint16_t* src, int bytes) {
  const unsigned char* s = reinterpret_cast<const unsigned char*>(src);
  for (int i = threadIdx.x * 16; i < bytes; i += kThreads * 16) cp_async16(dst + i, s + i);
}

// decode a row of one phase buffer: returns false for gap rows
__device__ __forceinline__ bool decode_row(const BfBlk& b, int V, int rloc, int& w, int& t, int& v) {
  const int rr = rloc - b.gap * V;
  if (rr < 0) return false;
  const int slotV = b.slot * V;
  w = rr / slotV;
  const int q = rr - w * slotV;
  if (q >= b.Tout * V) return false;
  t = q / V;
  v = q - t * V;
  return true;
}

__device__ __forceinline__ void zero_fill(unsigned char* p, int bytes) {
  uint4* q = reinterpret_cast<uint4*>(p);
  for (int i = threadIdx.x; i < (bytes >> 4); i += kThreads) q[i] = make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(kThreads, 2)
tokenizer_bf16_kernel(const __grid_constant__ BfPlan pl, const float* __restrict__ poses, float* __restrict__ tokens,
                      int64_t B) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* sA = smem + pl.off_A;
  unsigned char* sX[2] = {smem + pl.off_X0, smem + pl.off_X1};
  unsigned char* sWT = smem + pl.off_WT;
  unsigned char* sWG = smem + pl.off_WG;
  float* x0 = reinterpret_cast<float*>(smem + pl.off_x0);      // [G][c_in][T0][V] fp32, BN folded
  float* m0 = reinterpret_cast<float*>(smem + pl.off_m0);      // adjacency-mixed copy
  const int V = pl.V, G = pl.G;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lane_grp = warp & 3, col_half = warp >> 2;

  if (warp == 0) tmem_alloc(&tmem_base_s, pl.tmem_cols);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint32_t parity = 0;

  const int64_t n_groups = (B + G - 1) / G;
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int64_t w_first = grp * G;
    const int nw = (int)((B - w_first) < (int64_t)G ? (B - w_first) : (int64_t)G);

    // =============================== block 0 ===============================
    {
      const BfBlk& b = pl.blk[0];
      stage(sWT, b.w_tcn, kTaps * b.npad * b.npad * 2);
      // x0 <- folded BN1d(poses); windows beyond the batch end are zero
      const int per_w = pl.c_in * pl.T0 * V;
      for (int i = threadIdx.x; i < G * per_w; i += kThreads) {
        const int w = i / per_w, r = i - w * per_w;
        const int v = r % V, c = r / (pl.T0 * V);
        const int j = c * V + v;
        x0[i] = w < nw ? fmaf(__ldg(poses + (size_t)(w_first + w) * per_w + r), __ldg(pl.in_scale + j), __ldg(pl.in_shift + j)) : 0.f;
      }
      __syncthreads();
      // m0 <- A_hat . x0 over the keypoint axis
      for (int i = threadIdx.x; i < G * per_w; i += kThreads) {
        const int v = i % V;
        const float* row = x0 + (i - v);
        float a = 0.f;
        for (int e = 0; e < b.ell_width; ++e) a = fmaf(__ldg(b.ell_val + v * b.ell_width + e), row[__ldg(b.ell_col + v * b.ell_width + e)], a);
        m0[i] = a;
      }
      // next block's input buffer: gaps must be zero
      if (pl.n_blocks > 1) zero_fill(sX[0], pl.blk[1].rtot * pl.blk[1].kin * 2);
      __syncthreads();
      // g0 = relu(W0 . m0 + b0) -> bf16, phase-split rows, 8 channels (one 16-byte granule) per item
      {
        const int chunks = b.npad >> 3;
        const int items = b.rtot * chunks;
        for (int it = threadIdx.x; it < items; it += kThreads) {
          const int r = it % b.rtot, j = it / b.rtot;
          const int ph = r / b.rows, rloc = r - ph * b.rows;
          int w, t2, v;
          uint4 out = make_uint4(0, 0, 0, 0);
          if (decode_row(b, V, rloc, w, t2, v) && w < nw) {
            const int t = b.stride * t2 + ph;
            float mv[4];
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) mv[ci] = ci < b.cin ? m0[((w * pl.c_in + ci) * pl.T0 + t) * V + v] : 0.f;
            float g[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int c = j * 8 + e;
              float a = 0.f;
              if (c < b.cout) {
                a = __ldg(b.gcn_b + c);
#pragma unroll
                for (int ci = 0; ci < 4; ++ci)
                  if (ci < b.cin) a = fmaf(__ldg(b.gcn_w32 + ci * b.cout + c), mv[ci], a);
                a = fmaxf(a, 0.f);
              }
              g[e] = a;
            }
            out = make_uint4(pack_bf16x2(g[0], g[1]), pack_bf16x2(g[2], g[3]), pack_bf16x2(g[4], g[5]), pack_bf16x2(g[6], g[7]));
          }
          *reinterpret_cast<uint4*>(sA + ((size_t)j * b.rtot + r) * 16) = out;
        }
      }
    }

    for (int bi = 0; bi < pl.n_blocks; ++bi) {
      const BfBlk& b = pl.blk[bi];
      unsigned char* sXin = sX[(bi + 1) & 1];      // x_b   (b >= 1)
      unsigned char* sXout = sX[bi & 1];           // x_{b+1}
      const uint32_t planeA = (uint32_t)b.rtot * 16u;
      const int dcol = b.npad < 32 ? 32 : b.npad;   // TMEM column stride between accumulator tiles
      if (bi > 0) {
        // ---- stage this block's weights, clear the output buffer, adjacency mix x_b -> A (bf16)
        stage(sWG, b.w_gcn, b.kin * b.npad * 2);
        stage(sWT, b.w_tcn, kTaps * b.npad * b.npad * 2);
        if (b.w_res) stage(sWT + kTaps * b.npad * b.npad * 2, b.w_res, b.kin * b.npad * 2);
        if (bi + 1 < pl.n_blocks) zero_fill(sXout, pl.blk[bi + 1].rtot * pl.blk[bi + 1].kin * 2);
        {
          const int chunks = b.kin >> 3;
          const int items = b.rtot * chunks;
          for (int it = threadIdx.x; it < items; it += kThreads) {
            const int r = it % b.rtot, j = it / b.rtot;
            const int ph = r / b.rows, rloc = r - ph * b.rows;
            int w, t2, v;
            uint4 out = make_uint4(0, 0, 0, 0);
            if (decode_row(b, V, rloc, w, t2, v)) {
              float a[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) a[e] = 0.f;
              const unsigned char* plane = sXin + (size_t)j * planeA;
              for (int e2 = 0; e2 < b.ell_width; ++e2) {
                const float val = __ldg(b.ell_val + v * b.ell_width + e2);
                const int u = __ldg(b.ell_col + v * b.ell_width + e2);
                const uint4 q = *reinterpret_cast<const uint4*>(plane + (size_t)(r + u - v) * 16);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = __bfloat1622float2(h[e]);
                  a[2 * e] = fmaf(val, f.x, a[2 * e]);
                  a[2 * e + 1] = fmaf(val, f.y, a[2 * e + 1]);
                }
              }
              out = make_uint4(pack_bf16x2(a[0], a[1]), pack_bf16x2(a[2], a[3]), pack_bf16x2(a[4], a[5]), pack_bf16x2(a[6], a[7]));
            }
            *reinterpret_cast<uint4*>(sA + ((size_t)j * b.rtot + r) * 16) = out;
          }
        }
        cp_async_wait_all();
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        // ---- P = M . W  (all rows of all phases), accumulators in TMEM
        const int p_tiles = (b.rtot + 127) >> 7;
        if (threadIdx.x == 0) {
          tc_fence_after();
          const uint32_t idesc = make_idesc(128, b.npad, false);
          const uint64_t bdesc0 = make_desc(smem_u32(sWG), (uint32_t)b.npad * 16u, 128u);
          for (int tile = 0; tile < p_tiles; ++tile) {
            const uint64_t adesc0 = make_desc(smem_u32(sA) + (uint32_t)tile * 2048u, planeA, 128u);
            for (int ks = 0; ks < (b.kin >> 4); ++ks)
              umma_bf16(tmem + (uint32_t)(tile * dcol), desc_advance(adesc0, (uint32_t)ks * 2u * planeA),
                        desc_advance(bdesc0, (uint32_t)ks * 2u * (uint32_t)b.npad * 16u), idesc, ks > 0);
          }
          umma_commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        tc_fence_after();
        // ---- g = relu(P + b) -> bf16 over M in place (gap rows -> 0)
        for (int tile = 0; tile < p_tiles; ++tile) {
          const int r = tile * 128 + lane_grp * 32 + lane;
          const int ph = r / b.rows, rloc = r - ph * b.rows;
          int w, t2, v;
          const bool data = r < b.rtot && decode_row(b, V, rloc, w, t2, v);
          const int groups = b.npad >> 4;
          for (int gq = (groups > 1 ? col_half : 0); gq < groups; gq += (groups > 1 ? 2 : 1)) {
            if (groups == 1 && col_half) break;
            float acc[16];
            tmem_ld16(tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(tile * dcol + gq * 16), acc);
            tmem_ld_wait();
            if (r < b.rtot) {
              uint32_t pk[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float lo = data ? fmaxf(acc[2 * e] + __ldg(b.gcn_b + gq * 16 + 2 * e), 0.f) : 0.f;
                const float hi = data ? fmaxf(acc[2 * e + 1] + __ldg(b.gcn_b + gq * 16 + 2 * e + 1), 0.f) : 0.f;
                pk[e] = pack_bf16x2(lo, hi);
              }
              *reinterpret_cast<uint4*>(sA + ((size_t)(gq * 2) * b.rtot + r) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(sA + ((size_t)(gq * 2 + 1) * b.rtot + r) * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
      } else {
        cp_async_wait_all();
      }
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      // ---- temporal conv (+ residual conv) as shifted-view MMAs
      const int m_tiles = (b.mrows + 127) >> 7;
      if (threadIdx.x == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc(128, b.npad, false);
        const uint32_t w_plane = (uint32_t)b.npad * 16u;
        const int ksteps = b.npad >> 4;
        for (int tile = 0; tile < m_tiles; ++tile) {
          const uint32_t d = tmem + (uint32_t)(tile * dcol);
          uint32_t acc_flag = 0;
          for (int tp = 0; tp < b.n_taps; ++tp) {
            const int row0 = b.tap_phase[tp] * b.rows + b.gap * V + tile * 128 + b.tap_rowoff[tp];
            const uint64_t adesc0 = make_desc(smem_u32(sA) + (uint32_t)row0 * 16u, planeA, 128u);
            const uint64_t bdesc0 = make_desc(smem_u32(sWT) + (uint32_t)(b.tap_k[tp] * (b.npad >> 3)) * w_plane, w_plane, 128u);
            for (int ks = 0; ks < ksteps; ++ks) {
              umma_bf16(d, desc_advance(adesc0, (uint32_t)ks * 2u * planeA), desc_advance(bdesc0, (uint32_t)ks * 2u * w_plane), idesc, acc_flag);
              acc_flag = 1;
            }
          }
          if (bi > 0 && !b.identity_res) {
            const uint32_t planeX = planeA;     // x_b shares the block's row geometry
            const int row0 = b.gap * V + tile * 128;                       // phase 0, offset 0: x[s*t']
            const uint64_t adesc0 = make_desc(smem_u32(sXin) + (uint32_t)row0 * 16u, planeX, 128u);
            const uint64_t bdesc0 = make_desc(smem_u32(sWT) + (uint32_t)(kTaps * b.npad * b.npad * 2), w_plane, 128u);
            for (int ks = 0; ks < (b.kin >> 4); ++ks)
              umma_bf16(d, desc_advance(adesc0, (uint32_t)ks * 2u * planeX), desc_advance(bdesc0, (uint32_t)ks * 2u * w_plane), idesc, 1u);
          }
        }
        umma_commit(&bar);
      }
      mbar_wait(&bar, parity);
      parity ^= 1;
      tc_fence_after();
      // ---- x_{b+1} = relu(acc + bias + residual): bf16 into the next block's phase layout, or fp32 tokens
      const bool last = bi + 1 == pl.n_blocks;
      for (int tile = 0; tile < m_tiles; ++tile) {
        const int mrow = tile * 128 + lane_grp * 32 + lane;
        const int slotV = b.slot * V;
        const int w = mrow / slotV;
        const int q = mrow - w * slotV;
        const bool data = mrow < b.mrows && q < b.Tout * V && w < nw;
        const int t = q / V, v = q - t * V;
        const int groups = b.npad >> 4;
        for (int gq = (groups > 1 ? col_half : 0); gq < groups; gq += (groups > 1 ? 2 : 1)) {
          if (groups == 1 && col_half) break;
          float acc[16];
          tmem_ld16(tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(tile * dcol + gq * 16), acc);
          tmem_ld_wait();
          if (!data) continue;
          float res[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) res[e] = __ldg(b.out_b + gq * 16 + e);
          if (bi == 0) {
            // K = Cin (2) residual conv on CUDA cores in fp32 from the un-mixed input
            for (int ci = 0; ci < b.cin; ++ci) {
              const float xv = x0[((w * pl.c_in + ci) * pl.T0 + b.stride * t) * V + v];
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (gq * 16 + e < b.cout) res[e] = fmaf(__ldg(b.res_w32 + ci * b.cout + gq * 16 + e), xv, res[e]);
            }
          } else if (b.identity_res) {
            const int r = b.gap * V + mrow;                                  // phase 0 (stride 1)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
              const uint4 qx = *reinterpret_cast<const uint4*>(sXin + ((size_t)(gq * 2 + h2) * b.rtot + r) * 16);
              const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&qx);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(hx[e]);
                res[h2 * 8 + 2 * e] += f.x;
                res[h2 * 8 + 2 * e + 1] += f.y;
              }
            }
          }
          float y[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) y[e] = fmaxf(acc[e] + res[e], 0.f);
          if (!last) {
            const BfBlk& nb = pl.blk[bi + 1];
            const int ph = t % nb.stride, t2 = t / nb.stride;
            const int r = ph * nb.rows + nb.gap * V + w * nb.slot * V + t2 * V + v;
            *reinterpret_cast<uint4*>(sXout + ((size_t)(gq * 2) * nb.rtot + r) * 16) =
                make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
            *reinterpret_cast<uint4*>(sXout + ((size_t)(gq * 2 + 1) * nb.rtot + r) * 16) =
                make_uint4(pack_bf16x2(y[8], y[9]), pack_bf16x2(y[10], y[11]), pack_bf16x2(y[12], y[13]), pack_bf16x2(y[14], y[15]));
          } else {
            float* dst = tokens + ((size_t)(w_first + w) * pl.S_out + t) * (size_t)(pl.c_last * V) + v;
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (gq * 16 + e < b.cout) dst[(gq * 16 + e) * V] = y[e];
          }
        }
      }
      tc_fence_before();
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, pl.tmem_cols);
}

inline int pad16(int x) { return (x + 15) & ~15; }

// Build the per-launch plan; returns false if this (model, T) is outside what the kernel covers.
bool build_plan(const sf_model* m, int T, int G, BfPlan* pl, const char** why) {
  static const char* reason = "";
  *why = reason;
  const Tokenizer& tk = m->tok;
  const int V = tk.V, nb = tk.n_blocks;
  if (tk.pool_tokens > 0) { *why = "adaptive pooling"; return false; }
  if (tk.c_in > 4) { *why = "more than 4 input channels"; return false; }
  pl->n_blocks = nb;
  pl->V = V;
  pl->c_in = tk.c_in;
  pl->G = G;
  pl->T0 = T;
  pl->in_scale = tk.in_scale;
  pl->in_shift = tk.in_shift;
  int Tin = T;
  size_t maxA = 0, maxX[2] = {0, 0}, maxWT = 0, maxWG = 0;
  int max_cols = 0;
  for (int i = 0; i < nb; ++i) {
    const TokBlock& tb = tk.blk[i];
    const BfBlockW& w = m->tokbf.blk[i];
    BfBlk& b = pl->blk[i];
    if (Tin % tb.stride) { *why = "temporal length not divisible by the block stride"; return false; }
    b.cin = tb.cin;
    b.kin = w.kin_pad;
    b.npad = w.npad;
    b.cout = tb.cout;
    b.stride = tb.stride;
    b.Tin = Tin;
    b.Tout = Tin / tb.stride;
    if (i > 0 && b.kin != pl->blk[i - 1].npad) { *why = "channel padding mismatch"; return false; }
    // taps: input row t = s*t' + k - 4 -> phase p = (k-4) mod s, offset o = (k-4-p)/s; live iff |o| < Tout
    b.n_taps = 0;
    b.gap = 0;
    for (int k = 0; k < kTaps; ++k) {
      const int d = k - kHalo;
      const int p = ((d % tb.stride) + tb.stride) % tb.stride;
      const int o = (d - p) / tb.stride;
      if (std::abs(o) >= b.Tout) continue;
      b.tap_k[b.n_taps] = k;
      b.tap_phase[b.n_taps] = p;
      b.tap_rowoff[b.n_taps] = o * V;
      b.gap = std::max(b.gap, std::abs(o));
      ++b.n_taps;
    }
    b.slot = b.Tout + b.gap;
    b.mrows = G * b.slot * V;
    b.rows = b.gap * V + b.mrows;
    b.rtot = b.stride * b.rows;
    b.identity_res = tb.identity_res;
    b.ell_width = tb.ell_width;
    b.ell_val = tb.ell_val;
    b.ell_col = tb.ell_col;
    b.w_tcn = w.tcn;
    b.w_gcn = w.gcn;
    b.w_res = w.res;
    b.gcn_b = w.gcn_b;
    b.out_b = w.out_b;
    b.gcn_w32 = tb.gcn_w;
    b.res_w32 = tb.res_w;
    if (i == 0 && tb.identity_res) { *why = "identity residual in block 0"; return false; }
    const size_t a_bytes = (size_t)b.rtot * std::max(b.npad, i > 0 ? b.kin : 0) * 2 + 4096;   // + MMA tile overrun slack
    maxA = std::max(maxA, a_bytes);
    if (i > 0) maxX[(i + 1) & 1] = std::max(maxX[(i + 1) & 1], (size_t)b.rtot * b.kin * 2 + 4096);
    maxWT = std::max(maxWT, (size_t)kTaps * b.npad * b.npad * 2 + (size_t)(w.res ? b.kin * b.npad * 2 : 0));
    if (i > 0) maxWG = std::max(maxWG, (size_t)b.kin * b.npad * 2);
    const int tiles = std::max((b.mrows + 127) / 128, i > 0 ? (b.rtot + 127) / 128 : 0);
    max_cols = std::max(max_cols, tiles * std::max(b.npad, 32));
    Tin = b.Tout;
  }
  pl->S_out = Tin;
  pl->c_last = tk.blk[nb - 1].cout;
  uint32_t cols = 32;
  while ((int)cols < max_cols) cols <<= 1;
  if (cols > 512) { *why = "accumulators exceed tensor memory"; return false; }
  pl->tmem_cols = cols;
  auto up = [](size_t x) { return (uint32_t)((x + 127) & ~size_t(127)); };
  uint32_t off = 0;
  pl->off_A = off; off += up(maxA);
  pl->off_X0 = off; off += up(std::max(maxX[0], (size_t)16));
  pl->off_X1 = off; off += up(std::max(maxX[1], (size_t)16));
  pl->off_WT = off; off += up(maxWT);
  pl->off_WG = off; off += up(std::max(maxWG, (size_t)16));
  const size_t xbytes = (size_t)G * tk.c_in * T * V * sizeof(float);
  pl->off_x0 = off; off += up(xbytes);
  pl->off_m0 = off; off += up(xbytes);
  pl->smem_bytes = off;
  if (off > (uint32_t)m->max_smem_optin) { *why = "activations + weights exceed shared memory"; return false; }
  // start-address field of the descriptor is 14 bits of 16-byte units = 256 KB: always fine on sm_100
  return true;
}

}  // namespace

bool tokenizer_bf16_supported(const sf_model* m, int T) {
  BfPlan pl;
  const char* why;
  return build_plan(m, T, 1, &pl, &why);
}

int launch_tokenizer_bf16(const sf_model* m, const float* poses, int64_t B, int T, float* tokens, cudaStream_t st) {
  if (B == 0) return SF_OK;
  BfPlan pl;
  const char* why = "";
  SF_REQUIRE(build_plan(m, T, 1, &pl, &why), SF_E_UNSUPPORTED, "bf16 tensor-core tokenizer does not cover this shape: %s", why);
  SF_CUDA_OK(cudaFuncSetAttribute(tokenizer_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
  int occ = 1;
  SF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tokenizer_bf16_kernel, kThreads, pl.smem_bytes));
  occ = std::max(1, std::min(occ, (int)(512 / pl.tmem_cols)));      // co-resident CTAs must all get their TMEM columns
  const int64_t n_groups = (B + pl.G - 1) / pl.G;
  const int grid = (int)std::min<int64_t>(n_groups, (int64_t)m->sm_count * occ);
  tokenizer_bf16_kernel<<<grid, kThreads, pl.smem_bytes, st>>>(pl, poses, tokens, B);
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}

}  // namespace sf
