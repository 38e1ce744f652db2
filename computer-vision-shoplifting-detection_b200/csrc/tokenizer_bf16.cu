// bf16 tcgen05 ST-GCN pose tokenizer: the fast path (tolerance 1e-2 on the score, north star).
//
// Same maths as tokenizer_fp32.cu (reference shopformer/models/gcae.py:124-154,185-195,
// 242-259,331-366 and shopformer_2/models/gcae.py:375-422), mapped onto the 5th-gen tensor cores:
//
//   * activations live in shared memory as bf16 in the planar-chunk layout of tc_common.cuh,
//     rows = (window, time, keypoint), columns = channels; accumulators live in TMEM (fp32);
//   * the 9x1 temporal convolution is an implicit GEMM: the input of a stride-s block is stored split
//     into s time phases, so tap k is a SHIFTED VIEW (row offset o_k * V) of phase (k-4) mod s -- nine
//     smem descriptors into one buffer, no im2col; zero rows ("gaps") between windows give the padding;
//     taps that can only ever hit padding are skipped;
//   * the BN-folded 1x1 strided residual conv is two more MMAs into the same accumulator (phase 0 of x);
//   * graph conv = adjacency mix over keypoints on CUDA cores (bf16 rows in smem, fp32 accumulate)
//     followed by a [rows x Cin] x [Cin x Cout] tensor-core GEMM; block 0 (Cin = 2 raw coordinates) feeds the
//     same GEMM with a K = 16 operand that carries the mixed poses split into bf16 hi + lo against split
//     weights, so its inputs are not rounded to bf16;
//   * epilogues read TMEM with tcgen05.ld, add the folded bias, ReLU, and write the next operand
//     straight back to shared memory as bf16 -- nothing but the final tokens (fp32) goes to HBM;
//   * every row -> (window, time, keypoint) decode is a shared-memory table built once per CTA, MMA issue is
//     spread over four warps;
//   * per-block weight images and the next window's poses arrive by TMA (cp.async.bulk, mbarrier complete_tx);
//     block b's temporal-conv weights land in x_{b+1}'s buffer, which is dead until block b's epilogue.
//
// One CTA owns one window at a time (persistent).  A window is a chain of 13 dependent phases separated by
// mbarrier waits (MMA completion) and __syncthreads, so SM throughput is the number of chains running side by
// side: the shared-memory plan (build_plan) fits THREE CTAs per SM for hidden 32 (76.5 KB, 128 TMEM columns,
// 80 registers x 256 threads each); shapes that only fit one CTA per SM run a 512-thread instantiation.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "sf_internal.h"
#include "tc_common.cuh"

namespace sf {
__device__ long long g_tok_timing[512];
__device__ int g_tok_timing_on = 0;
namespace {

using namespace tc;

constexpr int kMaxOcc = 3;            // CTAs per SM the register budget allows (3 x 256 threads x 80 registers)
constexpr int kThreadsMin = 256;      // 2 CTAs/SM shapes; shapes that only fit one CTA per SM run 512 threads
constexpr int kIssuers = 4;          // lane 0 of warps 0..3 issue the MMAs of tiles t = warp (mod 4)
constexpr int kEllMax = 8;
constexpr uint16_t kGap = 0xFFFF;

struct BfBlk {
  int cin, kin, npad, cout;            // real / padded input channels, padded / real output channels
  int stride, Tin, Tout, gap, slot;    // slot = Tout + gap time rows per window in a phase buffer
  int rows, mrows, rtot;               // rows between phase starts, M-space rows (G*slot*V), total rows of the buffer
  int n_taps;
  int tap_k[kTaps], tap_phase[kTaps], tap_rowoff[kTaps];
  int tap_arel[kTaps], tap_wrel[kTaps];   // per tap: A-view row offset and weight-slab offset, both in 16-byte units
  int identity_res, ell_width;
  const float* ell_val;
  const int* ell_col;
  const uint16_t *w_tcn, *w_gcn, *w_res;
  const float *gcn_b, *out_b;          // [npad], zero padded
  const float *gcn_w32, *res_w32;      // block 0 only: fp32 [cin][cout]
  uint32_t off_rowtab, off_mtab;       // smem tables (uint16)
  uint32_t off_drtab;                  // blocks >= 1: the data rows of the input buffer, row | keypoint << 11
  int n_data;
};

struct BfPlan {
  int n_blocks, V, c_in, G, T0, S_out, c_last;
  int bstride, ell_stride;             // floats per block in the bias tables (widest npad), ELL entries per keypoint row kept
  const float *in_scale, *in_shift;
  BfBlk blk[kMaxBlocks];
  uint32_t off_A, off_X0, off_X1, off_WG, off_x0, off_w0img;
  uint32_t off_bias_g, off_bias_o, off_r0, off_ellv, off_scale, off_shift;
  uint32_t smem_bytes, tmem_cols;
};
static_assert(sizeof(BfPlan) <= 3900, "BfPlan must fit in kernel parameter space");

__device__ __forceinline__ void zero_fill(unsigned char* p, int bytes) {
  uint4* q = reinterpret_cast<uint4*>(p);
  for (int i = threadIdx.x; i < (bytes >> 4); i += (int)blockDim.x) q[i] = make_uint4(0, 0, 0, 0);
}
// descriptor with separately tracked low word: advancing an operand is one 32-bit add
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint64_t desc_join(uint32_t lo) {      // SBO = 128 B, version 1
  return ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | lo;
}
__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  f[0] = __uint_as_float(q.x << 16); f[1] = __uint_as_float(q.x & 0xFFFF0000u);
  f[2] = __uint_as_float(q.y << 16); f[3] = __uint_as_float(q.y & 0xFFFF0000u);
  f[4] = __uint_as_float(q.z << 16); f[5] = __uint_as_float(q.z & 0xFFFF0000u);
  f[6] = __uint_as_float(q.w << 16); f[7] = __uint_as_float(q.w & 0xFFFF0000u);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
// the same through ReLU: one conversion instruction per pair, no separate max
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint4 pack8_relu(const float* f) {
  return make_uint4(pack_relu_bf16x2(f[0], f[1]), pack_relu_bf16x2(f[2], f[3]), pack_relu_bf16x2(f[4], f[5]), pack_relu_bf16x2(f[6], f[7]));
}


// ---- branch-free, loads-first inner loops (the CUDA-core phases are latency bound: keep independent loads in flight)
// packed ELL entry: .x = adjacency value, .y = row delta (u - v) as int bits; unused entries are (0, 0)
// Block 0 (Cin <= 4 raw coordinates): adjacency mix of the fp32 poses, written as the bf16 A operand of a K = 16
// tensor-core GEMM.  Each mixed value m is split m = hi + lo (two bf16) and each weight w = w_hi + w_lo, and the
// K axis carries the three significant products per input channel: columns [3c, 3c+1, 3c+2] = [hi, hi, lo] against
// weight rows [w_hi, w_lo, w_hi] -- the graph conv of block 0 stays at fp32 input accuracy on the tensor cores.
template <int W, int CIN>
__device__ __forceinline__ void mix_a0(const float* __restrict__ x0, unsigned char* __restrict__ a0, uint32_t plane_bytes,
                                       const uint16_t* __restrict__ rt, const float2* __restrict__ ell, int V, int tv,
                                       int per_w, int nw, int rtot) {
  for (int r = threadIdx.x; r < rtot; r += (int)blockDim.x) {
    const uint32_t en = rt[r];
    if (en == kGap) continue;             // gap rows of A0 only feed GEMM rows that the g0 epilogue replaces by zeros
    const bool ok = (int)(en >> 11) < nw;
    const int v = (int)(en & 31);
    const int base = ok ? (int)(en >> 11) * per_w + (int)((en >> 5) & 63) * V + v : v;
    float2 e[W];
#pragma unroll
    for (int k = 0; k < W; ++k) e[k] = ell[k * V + v];
    float m[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      float xv[W];
#pragma unroll
      for (int k = 0; k < W; ++k) xv[k] = x0[base + ci * tv + __float_as_int(e[k].y)];
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < W; ++k) acc = fmaf(e[k].x, xv[k], acc);
      m[ci] = ok ? acc : 0.f;
    }
    float f[16];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const float hi = __bfloat162float(__float2bfloat16_rn(m[ci]));
      f[3 * ci] = hi;
      f[3 * ci + 1] = hi;
      f[3 * ci + 2] = m[ci] - hi;
    }
    f[12] = f[13] = f[14] = f[15] = 0.f;
    *reinterpret_cast<uint4*>(a0 + (size_t)r * 16) = pack8(f);
    *reinterpret_cast<uint4*>(a0 + plane_bytes + (size_t)r * 16) = pack8(f + 8);
  }
}

// adjacency mix of one 8-channel granule column of the bf16 activation buffer: dst[r] = sum_k val_k * src[r + delta_k],
// over the DATA rows only (dr[i] = row | keypoint << 11).  Gap rows of the destination are left as they are: they only
// feed GEMM rows whose outputs the next epilogue replaces by zeros.
template <int W>
__device__ __forceinline__ void mix_rows(const unsigned char* __restrict__ plane, unsigned char* __restrict__ dstp,
                                         const uint16_t* __restrict__ dr, const float2* __restrict__ ell, int ellV,
                                         int i_first, int i_step, int n_data) {
  for (int i = i_first; i < n_data; i += i_step) {
    const uint32_t en = dr[i];
    const int r = (int)(en & 0x7FF), v = (int)(en >> 11);
    float2 e[W];
#pragma unroll
    for (int k = 0; k < W; ++k) e[k] = ell[k * ellV + v];
    uint4 q[W];
#pragma unroll
    for (int k = 0; k < W; ++k) q[k] = *reinterpret_cast<const uint4*>(plane + (size_t)(r + __float_as_int(e[k].y)) * 16);
    float a[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = 0.f;
#pragma unroll
    for (int k = 0; k < W; ++k) {
      float f[8];
      unpack8(q[k], f);
#pragma unroll
      for (int c = 0; c < 8; ++c) a[c] = fmaf(e[k].x, f[c], a[c]);
    }
    *reinterpret_cast<uint4*>(dstp + (size_t)r * 16) = pack8(a);
  }
}

// Straight-line MMA issue for one accumulator tile, executed by ONE thread: all descriptor words are computed
// first (independent adds), then the tcgen05.mma instructions go out back to back.  Measured on B200: ~50 cycles
// per 128x32x16 MMA this way versus ~180 when descriptor arithmetic / branches sit between the MMAs.
template <int NT, int KS, int RS>
__device__ __forceinline__ void conv_issue(uint32_t d, uint32_t a_lo0, uint32_t b_lo0, const int* __restrict__ arel,
                                           const int* __restrict__ wrel, uint32_t astep, uint32_t bstep, uint32_t idesc,
                                           uint32_t r_alo, uint32_t r_blo) {
  uint32_t al[NT], bl[NT];
#pragma unroll
  for (int tp = 0; tp < NT; ++tp) {
    al[tp] = a_lo0 + (uint32_t)arel[tp];
    bl[tp] = b_lo0 + (uint32_t)wrel[tp];
  }
#pragma unroll
  for (int tp = 0; tp < NT; ++tp)
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      umma_bf16(d, desc_join(al[tp] + ks * astep), desc_join(bl[tp] + ks * bstep), idesc, (tp | ks) ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < RS; ++ks) umma_bf16(d, desc_join(r_alo + ks * astep), desc_join(r_blo + ks * bstep), idesc, 1u);
}

// The clock64 stamps of the timeline tools (profiles/*_timing.py) are compiled in only with `make EXTRA=-DSF_STAMPS`: even
// disabled at run time they cost the production kernels 2-5 % (tokenizer v2 1.717 -> 1.631 ms, transformer 1.086 -> 1.065 ms,
// one-window tokenizer on config B 5.08 -> 4.89 ms per 65,536 windows).
#if !defined(SF_STAMPS) && !defined(SF_TOK2_FINE_STAMPS)
#define TOK_STAMP(id) do { } while (0)
#else
#define TOK_STAMP(id)                                                                   \
  do {                                                                                  \
    if (timing && threadIdx.x == 0 && stamp_i < 510) {                                  \
      g_tok_timing[stamp_i++] = (long long)(id);                                        \
      g_tok_timing[stamp_i++] = clock64();                                              \
    }                                                                                   \
  } while (0)
#endif

template <int kThreads>
__global__ void __launch_bounds__(kThreads, kThreads == 256 ? kMaxOcc : 1)
tokenizer_bf16_kernel(const __grid_constant__ BfPlan pl, const float* __restrict__ poses, float* __restrict__ tokens,
                      int64_t B) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar, wbar, pbar;        // MMA completion; TMA: weights of a block, poses of a window group
  __shared__ uint32_t tmem_base_s;
  unsigned char* sA = smem + pl.off_A;
  unsigned char* sX[2] = {smem + pl.off_X0, smem + pl.off_X1};
  unsigned char* sWG = smem + pl.off_WG;
  float* x0 = reinterpret_cast<float*>(smem + pl.off_x0);      // raw poses of the current window group (fp32)
  const float* bias_g = reinterpret_cast<const float*>(smem + pl.off_bias_g);       // [blk][64]
  const float* bias_o = reinterpret_cast<const float*>(smem + pl.off_bias_o);
  const unsigned char* w0img = smem + pl.off_w0img;                                 // block-0 graph-conv weights, [2][npad][8] bf16
  const float* r0s = reinterpret_cast<const float*>(smem + pl.off_r0);              // [4][64]
  const float2* ell2 = reinterpret_cast<const float2*>(smem + pl.off_ellv);         // [blk][kEllMax][V] (value, row delta)
  const float* scale_s = reinterpret_cast<const float*>(smem + pl.off_scale);
  const float* shift_s = reinterpret_cast<const float*>(smem + pl.off_shift);
  const int V = pl.V, G = pl.G;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform on purpose (uniform datapath)
  const int lane = threadIdx.x & 31;
  const int lane_grp = warp & 3, col_part = warp >> 2;
  constexpr int kParts = kThreads / 128;                 // warps sharing a 32-row lane group split (column group, tile) units
  const int per_w = pl.c_in * pl.T0 * V;

  // ------------------------------------------------------------------ one-time setup
  if (warp == 0) tmem_alloc(&tmem_base_s, pl.tmem_cols);
  if (threadIdx.x == 0) {
    mbar_init(&bar, kIssuers);
    mbar_init(&wbar, 1);
    mbar_init(&pbar, 1);
    fence_mbar_init();
  }
  for (int bi = 0; bi < pl.n_blocks; ++bi) {
    const BfBlk& b = pl.blk[bi];
    uint16_t* rt = reinterpret_cast<uint16_t*>(smem + b.off_rowtab);
    uint16_t* mt = reinterpret_cast<uint16_t*>(smem + b.off_mtab);
    const int slotV = b.slot * V;
    for (int r = threadIdx.x; r < b.rtot; r += kThreads) {          // row of the phase-split input buffer
      const int ph = r / b.rows, rr = r - ph * b.rows - b.gap * V;
      uint16_t e = kGap;
      if (rr >= 0) {
        const int w = rr / slotV, q = rr - w * slotV;
        if (q < b.Tout * V) {
          const int t2 = q / V, v = q - t2 * V;
          e = (uint16_t)(v | ((b.stride * t2 + ph) << 5) | (w << 11));       // v, input time index, window
        }
      }
      rt[r] = e;
    }
    if (bi > 0) {
      uint16_t* dr = reinterpret_cast<uint16_t*>(smem + b.off_drtab);
      const int per_phase = G * b.Tout * V;                             // data rows of one phase, in row order
      for (int i = threadIdx.x; i < b.n_data; i += kThreads) {
        const int ph = i / per_phase, q = i - ph * per_phase;
        const int w = q / (b.Tout * V), qq = q - w * b.Tout * V;
        const int r = ph * b.rows + b.gap * V + w * slotV + qq;
        dr[i] = (uint16_t)(r | ((qq % V) << 11));
      }
    }
    const bool last = bi + 1 == pl.n_blocks;
    for (int mrow = threadIdx.x; mrow < b.mrows; mrow += kThreads) {         // row of the conv output (M space)
      const int w = mrow / slotV, q = mrow - w * slotV;
      uint16_t e = kGap;
      if (q < b.Tout * V) {
        const int t = q / V, v = q - t * V;
        int target;
        if (!last) {
          const BfBlk& nb = pl.blk[bi + 1];
          target = (t % nb.stride) * nb.rows + nb.gap * V + w * nb.slot * V + (t / nb.stride) * V + v;
        } else {
          target = t * (pl.c_last * V) + v;                                   // offset inside the window's tokens
        }
        e = (uint16_t)(target | (w << 11));
      }
      mt[mrow] = e;
    }
    float* bg = reinterpret_cast<float*>(smem + pl.off_bias_g) + bi * pl.bstride;
    float* bo = reinterpret_cast<float*>(smem + pl.off_bias_o) + bi * pl.bstride;
    for (int i = threadIdx.x; i < pl.bstride; i += kThreads) {
      bg[i] = i < b.npad ? __ldg(b.gcn_b + i) : 0.f;
      bo[i] = i < b.npad ? __ldg(b.out_b + i) : 0.f;
    }
    float2* e2 = reinterpret_cast<float2*>(smem + pl.off_ellv) + bi * V * pl.ell_stride;
    for (int i = threadIdx.x; i < V * pl.ell_stride; i += kThreads) {
      const int k = i / V, v = i % V;                    // [k][v]: a warp's lanes read consecutive entries
      const bool ok = k < b.ell_width;
      const float val = ok ? __ldg(b.ell_val + v * b.ell_width + k) : 0.f;
      const int dl = (ok && val != 0.f) ? __ldg(b.ell_col + v * b.ell_width + k) - v : 0;
      e2[i] = make_float2(val, __int_as_float(dl));
    }
  }
  {
    const BfBlk& b0 = pl.blk[0];
    for (int i = threadIdx.x; i < pl.c_in * 64; i += kThreads) {
      const int ci = i >> 6, c = i & 63;
      const bool ok = ci < b0.cin && c < b0.cout;
      reinterpret_cast<float*>(smem + pl.off_r0)[i] = ok ? __ldg(b0.res_w32 + ci * b0.cout + c) : 0.f;
    }
    // K = 16 weight image of the block-0 graph conv (see mix_a0): rows [3c, 3c+1, 3c+2] = [w_hi, w_lo, w_hi]
    for (int i = threadIdx.x; i < 16 * b0.npad; i += kThreads) {
      const int k = i / b0.npad, n = i - k * b0.npad;
      const int ci = k / 3, j = k - 3 * ci;
      float val = 0.f;
      if (ci < b0.cin && n < b0.cout) {
        const float w = __ldg(b0.gcn_w32 + ci * b0.cout + n);
        const float hi = __bfloat162float(__float2bfloat16_rn(w));
        val = j == 1 ? w - hi : hi;
      }
      reinterpret_cast<__nv_bfloat16*>(smem + pl.off_w0img)[((k >> 3) * b0.npad + n) * 8 + (k & 7)] = __float2bfloat16_rn(val);
    }
    for (int i = threadIdx.x; i < pl.c_in * V; i += kThreads) {
      reinterpret_cast<float*>(smem + pl.off_scale)[i] = __ldg(pl.in_scale + i);
      reinterpret_cast<float*>(smem + pl.off_shift)[i] = __ldg(pl.in_shift + i);
    }
  }
  // operand buffers start out finite: an MMA may read stale bytes against zero weights, and NaN * 0 = NaN
  zero_fill(smem + pl.off_A, (int)(pl.off_x0 - pl.off_A));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint32_t parity = 0;
  const bool timing = g_tok_timing_on && blockIdx.x == 0;
  int stamp_i = 0;
  (void)timing; (void)stamp_i;                       // only used with -DSF_STAMPS

  const int64_t n_groups = (B + G - 1) / G;
  // raw poses of a window group are fetched by TMA one iteration ahead (one bulk copy, issued by thread 0)
  auto prefetch = [&](int64_t g_idx) {
    if (g_idx >= n_groups || threadIdx.x != 0) return;
    const int64_t wf = g_idx * G;
    const uint32_t bytes = (uint32_t)((B - wf) < (int64_t)G ? (B - wf) : (int64_t)G) * (uint32_t)per_w * 4u;
    mbar_expect_tx(&pbar, bytes);
    tma_load_1d(x0, poses + (size_t)wf * per_w, bytes, &pbar);
  };
  uint32_t wpar = 0, ppar = 0;                          // phase parities: weights barrier, pose barrier
  prefetch(blockIdx.x);
  for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int64_t w_first = grp * G;
    const int nw = (int)((B - w_first) < (int64_t)G ? (B - w_first) : (int64_t)G);
    TOK_STAMP(100);

    // =============================== block 0 prologue ===============================
    {
      const BfBlk& b = pl.blk[0];
      if (warp == 0) mbar_wait(&pbar, ppar);             // this group's poses (requested after the previous group's block 0)
      ppar ^= 1u;
      __syncthreads();
      if (threadIdx.x == 0) {                            // temporal-conv weights of block 0 into x1's (still unused) buffer
        const uint32_t bytes = (uint32_t)(kTaps * b.npad * b.npad * 2);
        mbar_expect_tx(&wbar, bytes);
        tma_load_1d(sX[0], b.w_tcn, bytes, &wbar);
      }
      const int n_valid = nw * per_w;
      for (int i = threadIdx.x; i < G * per_w; i += kThreads) {
        const int ii = i % per_w, j = (ii / (pl.T0 * V)) * V + ii % V;          // c*V + v of this pose element
        x0[i] = i < n_valid ? fmaf(x0[i], scale_s[j], shift_s[j]) : 0.f;      // folded BatchNorm1d, in place
      }
      __syncthreads();
      TOK_STAMP(101);
      // A0 <- split(A_hat . x0): K = 16 operand rows of the phase-split layout, in the A buffer (g0 later overwrites it)
      const uint16_t* rt0 = reinterpret_cast<const uint16_t*>(smem + b.off_rowtab);
      const uint32_t pa0 = (uint32_t)b.rtot * 16u;
      const int tv0 = pl.T0 * V;
#define SF_MIX_A0(W_) \
  switch (b.cin) { \
    case 1: mix_a0<W_, 1>(x0, sA, pa0, rt0, ell2, V, tv0, per_w, nw, b.rtot); break; \
    case 2: mix_a0<W_, 2>(x0, sA, pa0, rt0, ell2, V, tv0, per_w, nw, b.rtot); break; \
    case 3: mix_a0<W_, 3>(x0, sA, pa0, rt0, ell2, V, tv0, per_w, nw, b.rtot); break; \
    default: mix_a0<W_, 4>(x0, sA, pa0, rt0, ell2, V, tv0, per_w, nw, b.rtot); break; \
  }
      if (b.ell_width <= 5) { SF_MIX_A0(5) } else { SF_MIX_A0(kEllMax) }
#undef SF_MIX_A0
      TOK_STAMP(102);
    }

    for (int bi = 0; bi < pl.n_blocks; ++bi) {
      const BfBlk& b = pl.blk[bi];
      unsigned char* sXin = sX[(bi + 1) & 1];      // x_b   (b >= 1)
      unsigned char* sXout = sX[bi & 1];           // x_{b+1}
      unsigned char* sWT = sXout;                  // ... which holds this block's conv weights until the conv epilogue
      const uint32_t planeA = (uint32_t)b.rtot * 16u;
      const int dcol = b.npad < 32 ? 32 : b.npad;   // TMEM column stride between accumulator tiles
      const uint16_t* rt = reinterpret_cast<const uint16_t*>(smem + b.off_rowtab);
      const uint16_t* mt = reinterpret_cast<const uint16_t*>(smem + b.off_mtab);
      const int groups = b.npad >> 4;               // 16-column groups per accumulator tile
      // epilogue work split over the kParts warps of a lane group: by column group when there are enough of them,
      // otherwise also by accumulator tile
      const int g_first = groups >= kParts ? col_part : col_part % groups, g_step = groups >= kParts ? kParts : groups;
      const int t_first = groups >= kParts ? 0 : col_part / groups, t_step = groups >= kParts ? 1 : kParts / groups;
      TOK_STAMP(110 + bi * 10);
      if (bi > 0) {
        // ---- stage this block's weights, clear the output buffer, adjacency mix x_b -> A (bf16)
        if (threadIdx.x == 0) {                          // TMA: graph-conv, temporal-conv and residual weight images
          const uint32_t bg = (uint32_t)(b.kin * b.npad * 2), bt = (uint32_t)(kTaps * b.npad * b.npad * 2);
          const uint32_t br = b.w_res ? bg : 0u;
          mbar_expect_tx(&wbar, bg + bt + br);
          tma_load_1d(sWG, b.w_gcn, bg, &wbar);
          tma_load_1d(sWT, b.w_tcn, bt, &wbar);           // (sWT = x_{b+1}'s buffer)
          if (br) tma_load_1d(sWT + bt, b.w_res, br, &wbar);
        }
        {
          const int chunks = b.kin >> 3;
          const int tpc = kThreads / chunks;
          const int j = threadIdx.x / tpc;
          const unsigned char* plane = sXin + (size_t)j * planeA;
          unsigned char* dstp = sA + (size_t)j * planeA;
          const float2* el = ell2 + bi * V * pl.ell_stride;
          const uint16_t* dr = reinterpret_cast<const uint16_t*>(smem + b.off_drtab);
          if (b.ell_width <= 5) mix_rows<5>(plane, dstp, dr, el, V, threadIdx.x - j * tpc, tpc, b.n_data);
          else mix_rows<kEllMax>(plane, dstp, dr, el, V, threadIdx.x - j * tpc, tpc, b.n_data);
        }
        TOK_STAMP(111 + bi * 10);
        if (warp == 0) mbar_wait(&wbar, wpar);
        wpar ^= 1u;
      }
      {
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        TOK_STAMP(112 + bi * 10);
        // ---- P = M . W  (all rows of all phases), accumulators in TMEM.  Block 0: M = split A0 (K = 16) in x1's buffer
        const int p_tiles = (b.rtot + 127) >> 7;
        if (warp < kIssuers) {
          // warp-uniform control flow and descriptor arithmetic (uniform registers); only the MMA itself is
          // predicated on one lane
          tc_fence_after();
          const uint32_t idesc = make_idesc(128, b.npad, false);
          const uint32_t w_plane = (uint32_t)b.npad * 16u;
          const uint32_t blo0 = desc_lo(smem_u32(bi == 0 ? w0img : sWG), w_plane);
          const uint32_t a0 = smem_u32(sA);
          const int ksp = bi == 0 ? 1 : b.kin >> 4;
          if (elect_one()) {
            for (int tile = warp; tile < p_tiles; tile += kIssuers) {
              uint32_t alo = desc_lo(a0 + (uint32_t)tile * 2048u, planeA);
              uint32_t blo = blo0;
              for (int ks = 0; ks < ksp; ++ks) {
                umma_bf16(tmem + (uint32_t)(tile * dcol), desc_join(alo), desc_join(blo), idesc, ks > 0);
                alo += (2u * planeA) >> 4;
                blo += (2u * w_plane) >> 4;
              }
            }
            umma_commit(&bar);
          }
        }
        mbar_wait(&bar, parity);                    // every warp sleeps on the mbarrier itself (no polling warp + bar.sync hop)
        parity ^= 1;
        tc_fence_after();
        TOK_STAMP(113 + bi * 10);
        // ---- g = relu(P + b) -> bf16 over M in place (gap rows -> 0)
        const float* bgp = bias_g + bi * pl.bstride;
        for (int gq = g_first; gq < groups; gq += g_step) {
          float4 bb[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) bb[q4] = *reinterpret_cast<const float4*>(bgp + gq * 16 + q4 * 4);
          for (int t0 = t_first; t0 < p_tiles; t0 += t_step) {
            float acc[16];
            tmem_ld16(tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(t0 * dcol + gq * 16), acc);
            tmem_ld_wait();
            const int r = t0 * 128 + lane_grp * 32 + lane;
            if (r >= b.rtot) continue;
            const bool data = rt[r] != kGap;
            float y[16];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              y[q4 * 4 + 0] = data ? acc[q4 * 4 + 0] + bb[q4].x : 0.f;
              y[q4 * 4 + 1] = data ? acc[q4 * 4 + 1] + bb[q4].y : 0.f;
              y[q4 * 4 + 2] = data ? acc[q4 * 4 + 2] + bb[q4].z : 0.f;
              y[q4 * 4 + 3] = data ? acc[q4 * 4 + 3] + bb[q4].w : 0.f;
            }
            *reinterpret_cast<uint4*>(sA + ((size_t)(gq * 2) * b.rtot + r) * 16) = pack8_relu(y);        // ReLU inside the conversion
            *reinterpret_cast<uint4*>(sA + ((size_t)(gq * 2 + 1) * b.rtot + r) * 16) = pack8_relu(y + 8);
          }
        }
      }
      if (bi == 0) {                                  // temporal-conv weights (requested in the prologue)
        if (warp == 0) mbar_wait(&wbar, wpar);
        wpar ^= 1u;
      }
      TOK_STAMP(114 + bi * 10);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      TOK_STAMP(115 + bi * 10);
      // ---- temporal conv (+ residual conv) as shifted-view MMAs
      const int m_tiles = (b.mrows + 127) >> 7;
      if (warp < kIssuers) {
        tc_fence_after();
        const uint32_t idesc = make_idesc(128, b.npad, false);
        const uint32_t w_plane = (uint32_t)b.npad * 16u;
        const int ksteps = b.npad >> 4;
        const int n_taps = b.n_taps;
        const uint32_t a0 = smem_u32(sA), w0a = smem_u32(sWT);
        const uint32_t astep = (2u * planeA) >> 4, bstep = (2u * w_plane) >> 4;
        const int rsteps = (bi > 0 && !b.identity_res) ? (b.kin >> 4) : 0;
        for (int tile = warp; tile < m_tiles; tile += kIssuers) {
          const uint32_t d = tmem + (uint32_t)(tile * dcol);
          const uint32_t a_lo0 = desc_lo(a0, planeA) + (uint32_t)(tile * 128), b_lo0 = desc_lo(w0a, w_plane);
          const uint32_t r_alo = desc_lo(smem_u32(sXin) + (uint32_t)(b.gap * V + tile * 128) * 16u, planeA);   // phase 0 of x_b
          const uint32_t r_blo = desc_lo(w0a + (uint32_t)(kTaps * b.npad * b.npad * 2), w_plane);
          // (measured: issuing the conv MMAs from a plain lane-0 branch gives a 3% faster kernel than the
          // elect.sync / uniform-datapath form used for the other GEMMs -- the slower issue rate leaves more
          // shared-memory bandwidth to the co-resident CTA's CUDA-core phases)
          if (lane == 0) {
            const int* arel = b.tap_arel;
            const int* wrel = b.tap_wrel;
#define SF_ISSUE(NT, KS, RS) \
  if (n_taps == NT && ksteps == KS && rsteps == RS) conv_issue<NT, KS, RS>(d, a_lo0, b_lo0, arel, wrel, astep, bstep, idesc, r_alo, r_blo); else
            SF_ISSUE(9, 2, 0) SF_ISSUE(9, 2, 2) SF_ISSUE(7, 2, 2) SF_ISSUE(5, 2, 2) SF_ISSUE(5, 1, 2) SF_ISSUE(3, 1, 2)
            SF_ISSUE(9, 4, 0) SF_ISSUE(9, 4, 4) SF_ISSUE(7, 4, 4) SF_ISSUE(5, 4, 4) SF_ISSUE(3, 4, 4) SF_ISSUE(5, 1, 4) SF_ISSUE(3, 1, 4)
            SF_ISSUE(3, 2, 2) SF_ISSUE(5, 4, 0) SF_ISSUE(3, 4, 0)
            {   // generic fallback (any tap count / K split)
              uint32_t acc_flag = 0;
              for (int tp = 0; tp < n_taps; ++tp) {
                uint32_t alo = a_lo0 + (uint32_t)arel[tp], blo = b_lo0 + (uint32_t)wrel[tp];
                for (int ks = 0; ks < ksteps; ++ks) {
                  umma_bf16(d, desc_join(alo), desc_join(blo), idesc, acc_flag);
                  acc_flag = 1;
                  alo += astep;
                  blo += bstep;
                }
              }
              for (int ks = 0; ks < rsteps; ++ks) umma_bf16(d, desc_join(r_alo + ks * astep), desc_join(r_blo + ks * bstep), idesc, 1u);
            }
#undef SF_ISSUE
          }
        }
        if (elect_one()) umma_commit(&bar);
        TOK_STAMP(118 + bi * 10);
      }
      mbar_wait(&bar, parity);
      TOK_STAMP(119 + bi * 10);
      parity ^= 1;
      tc_fence_after();
      TOK_STAMP(116 + bi * 10);
      // ---- x_{b+1} = relu(acc + bias + residual): bf16 into the next block's phase layout, or fp32 tokens
      const bool last = bi + 1 == pl.n_blocks;
      if (!last) {
        // the conv weights that lived in x_{b+1}'s buffer are dead: zero its gap rows (the data rows are all written below)
        const BfBlk& nx = pl.blk[bi + 1];
        const uint16_t* rtn = reinterpret_cast<const uint16_t*>(smem + nx.off_rowtab);
        const int planes_n = nx.kin >> 3;
        for (int r = threadIdx.x; r < nx.rtot; r += kThreads)
          if (rtn[r] == kGap)
            for (int j = 0; j < planes_n; ++j)
              *reinterpret_cast<uint4*>(sXout + ((size_t)j * nx.rtot + r) * 16) = make_uint4(0, 0, 0, 0);
      }
      const float* bop = bias_o + bi * pl.bstride;
      const int nxt_rtot = last ? 0 : pl.blk[bi + 1].rtot;
      const int tv = pl.T0 * V;
      for (int gq = g_first; gq < groups; gq += g_step) {
        float4 bb[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) bb[q4] = *reinterpret_cast<const float4*>(bop + gq * 16 + q4 * 4);
        for (int t0 = t_first; t0 < m_tiles; t0 += t_step) {
          float acc[16];
          tmem_ld16(tmem + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(t0 * dcol + gq * 16), acc);
          tmem_ld_wait();
          const int mrow = t0 * 128 + lane_grp * 32 + lane;
          const uint32_t e = mrow < b.mrows ? mt[mrow] : kGap;
          const int w = e >> 11, target = e & 0x7FF;
          if (e == kGap || w >= nw) continue;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            acc[q4 * 4 + 0] += bb[q4].x; acc[q4 * 4 + 1] += bb[q4].y; acc[q4 * 4 + 2] += bb[q4].z; acc[q4 * 4 + 3] += bb[q4].w;
          }
          if (bi == 0) {
            // K = Cin (2) residual conv on CUDA cores in fp32 from the un-mixed input: raw-pose element of this output
            // row = window w, frame stride * t', keypoint v  (mrow = w*slot*V + t'*V + v)
            const int q0 = mrow - w * b.slot * V, t0r = q0 / V;
            const float* xp = x0 + w * per_w + (b.stride * t0r) * V + (q0 - t0r * V);
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
              if (ci < b.cin) {
                const float xv = xp[ci * tv];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  const float4 rr = *reinterpret_cast<const float4*>(r0s + ci * 64 + gq * 16 + q4 * 4);
                  acc[q4 * 4 + 0] = fmaf(rr.x, xv, acc[q4 * 4 + 0]); acc[q4 * 4 + 1] = fmaf(rr.y, xv, acc[q4 * 4 + 1]);
                  acc[q4 * 4 + 2] = fmaf(rr.z, xv, acc[q4 * 4 + 2]); acc[q4 * 4 + 3] = fmaf(rr.w, xv, acc[q4 * 4 + 3]);
                }
              }
          } else if (b.identity_res) {
            const int r = b.gap * V + mrow;                                  // phase 0 (stride 1)
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(sXin + ((size_t)(gq * 2) * b.rtot + r) * 16), f);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += f[q];
            unpack8(*reinterpret_cast<const uint4*>(sXin + ((size_t)(gq * 2 + 1) * b.rtot + r) * 16), f);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[8 + q] += f[q];
          }
          if (!last) {
            *reinterpret_cast<uint4*>(sXout + ((size_t)(gq * 2) * nxt_rtot + target) * 16) = pack8_relu(acc);
            *reinterpret_cast<uint4*>(sXout + ((size_t)(gq * 2 + 1) * nxt_rtot + target) * 16) = pack8_relu(acc + 8);
          } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) acc[q] = fmaxf(acc[q], 0.f);
            float* dst = tokens + (size_t)(w_first + w) * pl.S_out * (size_t)(pl.c_last * V) + target;
#pragma unroll
            for (int q = 0; q < 16; ++q)
              if (gq * 16 + q < b.cout) dst[(gq * 16 + q) * V] = acc[q];
          }
        }
      }
      fence_proxy_async();                            // x_{b+1} (generic stores) before later TMA / MMA accesses of these buffers
      tc_fence_before();
      __syncthreads();
      TOK_STAMP(117 + bi * 10);
      if (bi == 0) prefetch(grp + gridDim.x);         // the raw poses are dead: fetch the next group's
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, pl.tmem_cols);
}

// Build the per-launch plan; returns false if this (model, T) is outside what the kernel covers.
bool build_plan(const sf_model* m, int T, int G, BfPlan* pl, const char** why) {
  *why = "";
  const Tokenizer& tk = m->tok;
  const int V = tk.V, nb = tk.n_blocks;
  if (tk.pool_tokens > 0) { *why = "adaptive pooling"; return false; }
  if (tk.c_in > 4) { *why = "more than 4 input channels"; return false; }
  if (T > 63) { *why = "window longer than 63 frames"; return false; }
  if (G > 4 || tk.c_in * V > 255) { *why = "group / keypoint count outside table range"; return false; }
  pl->n_blocks = nb;
  pl->V = V;
  pl->c_in = tk.c_in;
  pl->G = G;
  pl->T0 = T;
  pl->in_scale = tk.in_scale;
  pl->in_shift = tk.in_shift;
  int Tin = T;
  size_t maxWG = 0;
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
    if (b.npad > 64 || (kThreadsMin % (b.npad >> 3)) || (i > 0 && (kThreadsMin % (b.kin >> 3)))) { *why = "channel count"; return false; }
    if (i > 0 && b.kin != pl->blk[i - 1].npad) { *why = "channel padding mismatch"; return false; }
    if (tb.ell_width > kEllMax) { *why = "adjacency rows with more than 8 non-zeros"; return false; }
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
      b.tap_wrel[b.n_taps] = k * (w.npad >> 3) * w.npad;
      b.gap = std::max(b.gap, std::abs(o));
      ++b.n_taps;
    }
    b.slot = b.Tout + b.gap;
    b.mrows = G * b.slot * V;
    // phase p = [gap frames][G window slots]; with one window per group the trailing gap of a phase is the leading gap
    // of the next one (shared), and one more gap closes the buffer
    b.rows = G == 1 ? (b.gap + b.Tout) * V : b.gap * V + b.mrows;
    b.rtot = G == 1 ? b.stride * b.rows + b.gap * V : b.stride * b.rows;
    for (int tp = 0; tp < b.n_taps; ++tp) b.tap_arel[tp] = b.tap_phase[tp] * b.rows + b.gap * V + b.tap_rowoff[tp];
    if (b.rtot > 2047) { *why = "operand buffer longer than 2047 rows"; return false; }
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
    if (i > 0) maxWG = std::max(maxWG, (size_t)b.kin * b.npad * 2);
    const int tiles = std::max((b.mrows + 127) / 128, (b.rtot + 127) / 128);
    max_cols = std::max(max_cols, tiles * std::max(b.npad, 32));
    Tin = b.Tout;
  }
  pl->S_out = Tin;
  pl->c_last = tk.blk[nb - 1].cout;
  if (Tin * pl->c_last * V > 2047) { *why = "token block larger than the table range"; return false; }
  uint32_t cols = 32;
  while ((int)cols < max_cols) cols <<= 1;
  if (cols > 512) { *why = "accumulators exceed tensor memory"; return false; }
  pl->tmem_cols = cols;
  auto up = [](size_t x) { return (uint32_t)((x + 127) & ~size_t(127)); };
  // Operand buffers first: an MMA tile that runs past the end of one buffer (rows that are discarded)
  // only ever reads the bytes of the next region, never past the allocation.
  // x2 (X1) is first written by block 1's epilogue, when the A buffer only holds M_b / g_b of blocks >= 1:
  // it lives in the tail of the A region that only g0 needs.
  // Shared-memory plan (three CTAs per SM for the default shapes):
  //   A   : block 0's split operand A0, then g0 in place; blocks >= 1: M_b / g_b in its first `A_late` bytes and the
  //         odd x buffer X1 in the tail that only g0 needs;
  //   X0/X1: x_{b+1} of even / odd blocks.  x_{b+1}'s buffer is dead until block b's conv epilogue writes it, so it
  //         first holds block b's temporal-conv (+ residual) weight images: W_b lands there by TMA, the conv MMAs read
  //         it, then the epilogue zeroes the gap rows and writes the data rows over it.
  size_t maxA_late = 0, need_x[2] = {16, 16};
  for (int i = 0; i < nb; ++i) {
    const BfBlk& bb = pl->blk[i];
    if (i > 0) {
      maxA_late = std::max(maxA_late, (size_t)bb.rtot * std::max(bb.npad, bb.kin) * 2);
      need_x[(i + 1) & 1] = std::max(need_x[(i + 1) & 1], (size_t)bb.rtot * bb.kin * 2);          // x_i = input of block i
    }
    const size_t wb = (size_t)kTaps * bb.npad * bb.npad * 2 + (size_t)(bb.w_res ? bb.kin * bb.npad * 2 : 0);
    need_x[i & 1] = std::max(need_x[i & 1], wb);                                                    // W_i in x_{i+1}'s buffer
  }
  const size_t a0_bytes = (size_t)pl->blk[0].rtot * 16 * 2, g0_bytes = (size_t)pl->blk[0].rtot * pl->blk[0].npad * 2;
  const size_t xbytes = (size_t)G * tk.c_in * T * V * sizeof(float);
  if (xbytes % 16) { *why = "window size not a multiple of 16 bytes"; return false; }
  uint32_t off = 0;
  pl->off_A = off;
  pl->off_X1 = off + up(maxA_late);
  off += std::max(up(std::max(a0_bytes, g0_bytes)), up(maxA_late) + up(need_x[1]));
  pl->off_X0 = off; off += up(need_x[0]);
  pl->off_WG = off; off += up(std::max(maxWG, (size_t)16));
  pl->off_x0 = off; off += up(xbytes);
  for (int i = 0; i < nb; ++i) {
    pl->blk[i].off_rowtab = off; off += up((size_t)pl->blk[i].rtot * 2);
    pl->blk[i].off_mtab = off; off += up((size_t)pl->blk[i].mrows * 2);
    pl->blk[i].n_data = i > 0 ? pl->blk[i].stride * G * pl->blk[i].Tout * V : 0;
    pl->blk[i].off_drtab = off;
    if (i > 0) off += up((size_t)pl->blk[i].n_data * 2);
  }
  pl->bstride = 16;
  pl->ell_stride = 1;
  for (int i = 0; i < nb; ++i) {
    pl->bstride = std::max(pl->bstride, pl->blk[i].npad);
    pl->ell_stride = std::max(pl->ell_stride, pl->blk[i].ell_width);
  }
  if (pl->ell_stride > 5) pl->ell_stride = kEllMax;      // the mix templates read 5 or 8 entries per row
  else pl->ell_stride = 5;
  pl->off_bias_g = off; off += up((size_t)nb * pl->bstride * 4);
  pl->off_bias_o = off; off += up((size_t)nb * pl->bstride * 4);
  pl->off_w0img = off; off += up((size_t)2 * pl->blk[0].npad * 16);
  pl->off_r0 = off; off += up((size_t)tk.c_in * 64 * 4);
  pl->off_ellv = off; off += up((size_t)nb * V * pl->ell_stride * 8);
  pl->off_scale = off; off += up((size_t)tk.c_in * V * 4);
  pl->off_shift = off; off += up((size_t)tk.c_in * V * 4);
  pl->smem_bytes = off;
  if (off > (uint32_t)m->max_smem_optin) { *why = "activations + weights exceed shared memory"; return false; }
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
  SF_REQUIRE(((uintptr_t)poses & 15) == 0, SF_E_INVALID, "pose buffer must be 16-byte aligned (TMA bulk copies)");
  count_launch(LK_TOK_BF16);
  // CTAs per SM: shared memory (1 KB driver reservation per CTA), registers (128 x 256 threads) and TMEM columns --
  // co-resident CTAs must all get their tensor-memory allocation or they would serialise on tcgen05.alloc.
  int smem_per_sm = 0;
  SF_CUDA_OK(cudaDeviceGetAttribute(&smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, m->device));
  auto occupancy = [&](const BfPlan& p) {
    const int o = smem_per_sm / (int)(p.smem_bytes + 1024 + 256);
    return std::max(1, std::min(std::min(o, kMaxOcc), (int)(512 / p.tmem_cols)));
  };
  int occ = occupancy(pl);
  // Windows per CTA pass.  A pass is a chain of fixed latencies (MMA completion, barriers, TMA) and SM throughput is the
  // number of windows in flight: shapes that fit two or three CTAs per SM run one window per pass in each; a shape that only
  // fits ONE CTA per SM (hidden 64) packs as many windows into its pass as shared and tensor memory allow instead
  // (config B: two windows per pass, 8.31 -> 5.08 ms per 65,536 windows; C and A' only fit one).
  if (occ == 1) {
    int g_max = 4;
    if (const char* e = getenv("SF_TOK_G")) g_max = std::max(1, atoi(e));
    for (int G = g_max; G > 1; --G) {
      BfPlan p2;
      const char* w2 = "";
      if (build_plan(m, T, G, &p2, &w2)) {
        pl = p2;
        break;
      }
    }
  }
  if (const char* dbg = getenv("SF_TOK_OCC")) occ = std::max(1, std::min(occ, atoi(dbg)));     // debugging aid
  const int64_t n_groups = (B + pl.G - 1) / pl.G;
  const int grid = (int)std::min<int64_t>(n_groups, (int64_t)m->sm_count * occ);
  if (getenv("SF_TOK_DEBUG")) fprintf(stderr, "tokenizer_bf16: smem %u B, tmem %u cols, %d CTAs/SM, grid %d\n", pl.smem_bytes, pl.tmem_cols, occ, grid);
  // two CTAs per SM run 256 threads each; a shape that only fits one CTA per SM gets 512 threads so that the
  // CUDA-core phases still have 16 warps per SM to hide latency with
  auto launch = [&](auto kernel, int threads) -> int {
    SF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    SF_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    kernel<<<grid, threads, pl.smem_bytes, st>>>(pl, poses, tokens, B);
    return SF_OK;
  };
  const bool wide = occ == 1 && !getenv("SF_TOK_NARROW");
  int rc = wide ? launch(tokenizer_bf16_kernel<512>, 512) : launch(tokenizer_bf16_kernel<256>, 256);
  if (rc) return rc;
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}

}  // namespace sf

// debugging aid (not part of the C ABI): phase timestamps of CTA 0, pairs of (phase id, clock64)
extern "C" int sfdbg_tokenizer_timing(int enable, long long* out_host, int n) {
  int on = enable;
  if (cudaMemcpyToSymbol(sf::g_tok_timing_on, &on, sizeof(int)) != cudaSuccess) return -1;
  if (out_host && n > 0) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out_host, sf::g_tok_timing, sizeof(long long) * (n < 512 ? n : 512)) != cudaSuccess) return -1;
  }
  return 0;
}
