// Tokenizer v2 (bf16 tcgen05 / TMEM): (window, keypoint) rows x (time, channel) columns, see tok2.h.
//
// One persistent CTA per SM interprets the tile program built by tok2_build.cu with three specialised roles:
//   warp 0      MMA issue: walks the G sequence, waits on the mbarriers of the E / L items a group depends on, issues the
//               group's tcgen05.mma from the descriptor table in shared memory and commits to the group's mbarrier;
//   warp 1      TMA: per-block temporal-conv weight images (L2 -> smem) and the NEXT tile's poses (HBM -> smem);
//   warps 4-11  epilogue / prep: block-0 operand preparation from the raw poses (BatchNorm1d fold, adjacency mix, bf16
//               hi/lo split), TMEM -> (bias, ReLU, bf16) -> shared-memory operand of the next MMA group, and the final
//               tokens (fp32) staged in shared memory and written by ONE bulk store per tile.
// Nothing but poses in and tokens out touches HBM; MMA groups of one chunk run under the epilogue of the previous one.
// Reference maths: shopformer/models/gcae.py:124-154,185-195,242-259,331-366; shopformer_2/models/gcae.py:375-422.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "sf_internal.h"
#include "tc_common.cuh"
#include "tok2_build.h"

namespace sf {
__device__ long long g_tok2_timing[4096];
__device__ int g_tok2_timing_on = 0;
namespace {

using namespace tc;
using namespace t2;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// two fp32 -> packed 16-bit pair (lo in the low half), optionally through ReLU; F16: fp16 (saturating) instead of bf16
template <bool F16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  uint32_t d;
  if (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack2_relu(float lo, float hi) {
  uint32_t d;
  if (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

// packed fp32 pairs (sm_100 FFMA2 / FADD2): two exact fp32 operations per issue slot
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n"
      ".reg .b64 ra, rb, rc, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "mov.b64 rc, {%6, %7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{\n"
      ".reg .b64 ra, rb, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "add.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// A MMA with a 64-bit B descriptor whose high word depends on the operand kind
__device__ __forceinline__ void issue_mma(uint32_t tmem, const Mma& m, uint32_t base16) {
  constexpr uint32_t kHiK = (128u >> 4) | (1u << 14);        // SBO = 128 B (next 8 rows), descriptor version 1
  constexpr uint32_t kHiMN = (kPlane >> 4) | (1u << 14);     // MN-major B: SBO = next 8-column chunk (one plane)
  const uint64_t ad = ((uint64_t)kHiK << 32) | (uint64_t)(m.a_lo + base16);
  const uint64_t bd = ((uint64_t)((m.d & (1u << 17)) ? kHiMN : kHiK) << 32) | (uint64_t)(m.b_lo + base16);
  umma_bf16(tmem + (m.d & 0x1FFu), ad, bd, m.idesc, (m.d >> 16) & 1u);
}

// per-stage phase stamps (table loaded / body done / arrived; one stamping warp per team) cost registers in the epilogue
// loop: only with `make EXTRA=-DSF_TOK2_FINE_STAMPS` (profiles/tok2_stage_timing.py)
#ifdef SF_TOK2_FINE_STAMPS
#define T2_FINE(id) T2_STAMP(id)
#else
#define T2_FINE(id) do { } while (0)
#endif
// The clock64 stamps of the timeline tools (profiles/*_timing.py) are compiled in only with `make EXTRA=-DSF_STAMPS`: even
// disabled at run time they cost the production kernels 2-5 % (tokenizer v2 1.717 -> 1.631 ms, transformer 1.086 -> 1.065 ms,
// one-window tokenizer on config B 5.08 -> 4.89 ms per 65,536 windows).
#if !defined(SF_STAMPS) && !defined(SF_TOK2_FINE_STAMPS)
#define T2_STAMP(id) do { } while (0)
#else
#define T2_STAMP(id)                                                      \
  do {                                                                    \
    if (timing && lane == 0 && stamp_i < stamp_end) {                     \
      g_tok2_timing[stamp_i++] = (long long)(id);                         \
      g_tok2_timing[stamp_i++] = clock64();                               \
    }                                                                     \
  } while (0)
#endif

// 16 fp32 -> two 16-byte granules of bf16 (optionally through ReLU) at columns [16 cg, 16 cg + 16) of a planar-chunk buffer
template <bool RELU, bool F16>
__device__ __forceinline__ void store16(unsigned char* dst_row, int cg, const float* a) {
  uint4 o0, o1;
  if (RELU) {
    o0 = make_uint4(pack2_relu<F16>(a[0], a[1]), pack2_relu<F16>(a[2], a[3]), pack2_relu<F16>(a[4], a[5]), pack2_relu<F16>(a[6], a[7]));
    o1 = make_uint4(pack2_relu<F16>(a[8], a[9]), pack2_relu<F16>(a[10], a[11]), pack2_relu<F16>(a[12], a[13]), pack2_relu<F16>(a[14], a[15]));
  } else {
    o0 = make_uint4(pack2<F16>(a[0], a[1]), pack2<F16>(a[2], a[3]), pack2<F16>(a[4], a[5]), pack2<F16>(a[6], a[7]));
    o1 = make_uint4(pack2<F16>(a[8], a[9]), pack2<F16>(a[10], a[11]), pack2<F16>(a[12], a[13]), pack2<F16>(a[14], a[15]));
  }
  *reinterpret_cast<uint4*>(dst_row + (size_t)(2 * cg) * kPlane) = o0;
  *reinterpret_cast<uint4*>(dst_row + (size_t)(2 * cg + 1) * kPlane) = o1;
}

// Block 0, one time slice [p0, p1) (at most two time steps) on the CUDA cores, fp32: thread = (row (window w, keypoint v),
// time step p0 + half):
//   m_c = sum_u A_hat[v][u] * bn(x[c][t][u])      (BatchNorm1d folded into the coefficient row, shopformer/models/gcae.py:351-355)
//   g_o = relu(m_x * W[x][o] + m_y * W[y][o] + b[o])   (graph conv, gcae.py:124-154)  -> bf16 operand slot of the temporal conv
// The weight table is read with broadcast shared-memory loads, the FMAs are packed fp32x2.  Rows beyond the tile's
// windows produce finite values nobody reads.
template <int KW, bool F16>
__device__ __forceinline__ void g0_stage(const Plan& pl, const StageK& s, unsigned char* smem, int row, int half, int my_w, int my_v, int nw,
                                         int* pz) {
  const int t = s.p0() + half;
  if (t >= s.p1()) return;
  const int V = pl.V, tv4 = pl.T0 * V * 4, cp0 = pl.cp0;
  const bool valid = row < pl.rows && my_w < nw;
  const bool two = pl.c_in > 1;
  const int chunks = cp0 / 8;                         // 8-channel granules per time step
  const int c8_lo = s.tmem_col(), c8_hi = s.tmem_col() + s.n_cg();     // this stage's share of the output channels
  unsigned char* dst_row = smem + s.dst_off + (size_t)(half * chunks) * kPlane + (size_t)row * 16;
  float2 mx = make_float2(0.f, 0.f), my = make_float2(0.f, 0.f);
  if (valid) {
    const float4* coef = reinterpret_cast<const float4*>(smem + pl.off_ell);
    const float2 hc = reinterpret_cast<const float2*>(smem + pl.off_hc)[my_v];
    const unsigned char* xp = smem + pl.off_xin + (size_t)(my_w * pl.per_w + my_v + t * V) * 4;
    float4 cf[KW];
#pragma unroll
    for (int k = 0; k < KW; ++k) cf[k] = coef[k * V + my_v];
    float x0[KW], x1[KW];
#pragma unroll
    for (int k = 0; k < KW; ++k) {
      const int db = __float_as_int(cf[k].z) * 4;           // neighbour's byte offset
      x0[k] = *reinterpret_cast<const float*>(xp + db);
      x1[k] = two ? *reinterpret_cast<const float*>(xp + tv4 + db) : 0.f;
    }
    float a0 = hc.x, a1 = hc.y;
#pragma unroll
    for (int k = 0; k < KW; ++k) {
      a0 = fmaf(cf[k].x, x0[k], a0);
      a1 = fmaf(cf[k].y, x1[k], a1);
    }
#ifdef SF_TOK2_G0_CHK_CHAIN
    float chk = 0.f;
#pragma unroll
    for (int k = 0; k < KW; ++k) chk = fmaf(x0[k], 0.f, fmaf(x1[k], 0.f, chk));
#else
    // stays 0 unless a gathered pose is inf or NaN: a non-finite input makes its mixed value non-finite whatever the
    // coefficient (0 * inf = NaN; unused ELL entries have coefficient 0 and gather the thread's own keypoint)
    const float chk = fmaf(a0, 0.f, a1 * 0.f);
#endif
    if (chk == 0.f) {
      mx = make_float2(a0, a0);
      my = make_float2(a1, a1);
    } else {
      // a window with a non-finite pose is reported as NaN tokens (the mix MMA would otherwise spread it over the tile)
      atomicOr(&pz[my_w], 1);
    }
  }
  if (F16) {
    // packed half arithmetic: y = relu(mx * w_x + (my * w_y + b)), two channels per HFMA2, ReLU fused into the second one
    uint32_t mx2, my2;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %1;" : "=r"(mx2) : "f"(mx.x));
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %1;" : "=r"(my2) : "f"(my.x));
    const uint4* tabh = reinterpret_cast<const uint4*>(smem + pl.off_g0tab_h);
#pragma unroll 2
    for (int c8 = c8_lo; c8 < c8_hi; ++c8) {
      const uint4 wx = tabh[c8 * 3], wy = tabh[c8 * 3 + 1], bb = tabh[c8 * 3 + 2];
      uint4 y;
      asm("{\n.reg .b32 t;\nfma.rn.f16x2 t, %1, %2, %3;\nfma.rn.relu.f16x2 %0, %4, %5, t;\n}" : "=r"(y.x) : "r"(my2), "r"(wy.x), "r"(bb.x), "r"(mx2), "r"(wx.x));
      asm("{\n.reg .b32 t;\nfma.rn.f16x2 t, %1, %2, %3;\nfma.rn.relu.f16x2 %0, %4, %5, t;\n}" : "=r"(y.y) : "r"(my2), "r"(wy.y), "r"(bb.y), "r"(mx2), "r"(wx.y));
      asm("{\n.reg .b32 t;\nfma.rn.f16x2 t, %1, %2, %3;\nfma.rn.relu.f16x2 %0, %4, %5, t;\n}" : "=r"(y.z) : "r"(my2), "r"(wy.z), "r"(bb.z), "r"(mx2), "r"(wx.z));
      asm("{\n.reg .b32 t;\nfma.rn.f16x2 t, %1, %2, %3;\nfma.rn.relu.f16x2 %0, %4, %5, t;\n}" : "=r"(y.w) : "r"(my2), "r"(wy.w), "r"(bb.w), "r"(mx2), "r"(wx.w));
      *reinterpret_cast<uint4*>(dst_row + (size_t)c8 * kPlane) = y;
    }
    return;
  }
  const float4* tab = reinterpret_cast<const float4*>(smem + pl.off_g0tab);
#pragma unroll 2
  for (int c8 = c8_lo; c8 < c8_hi; ++c8) {
    float4 w[6];                                      // (wx, wy, b) of outputs 8 c8 .. +4, then +4 .. +8
#pragma unroll
    for (int i = 0; i < 6; ++i) w[i] = tab[c8 * 6 + i];
    float2 y[4];
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      y[2 * g] = fma2(mx, make_float2(w[3 * g].x, w[3 * g].y), fma2(my, make_float2(w[3 * g + 1].x, w[3 * g + 1].y), make_float2(w[3 * g + 2].x, w[3 * g + 2].y)));
      y[2 * g + 1] = fma2(mx, make_float2(w[3 * g].z, w[3 * g].w), fma2(my, make_float2(w[3 * g + 1].z, w[3 * g + 1].w), make_float2(w[3 * g + 2].z, w[3 * g + 2].w)));
    }
    *reinterpret_cast<uint4*>(dst_row + (size_t)c8 * kPlane) =
        make_uint4(pack2_relu<F16>(y[0].x, y[0].y), pack2_relu<F16>(y[1].x, y[1].y), pack2_relu<F16>(y[2].x, y[2].y), pack2_relu<F16>(y[3].x, y[3].y));
  }
}

#ifndef SF_TOK2_XEPI_LOADS
#define SF_TOK2_XEPI_LOADS 1
#endif
constexpr int kXepiLoads = SF_TOK2_XEPI_LOADS;     // TMEM loads in flight per wait in block 0's output stage (1, 2 or 4)

// Block 0 output for output times [p0, p1): x1 = relu(acc + BN-folded strided 1x1 residual conv of the raw poses + bias)
// (gcae.py:237-259); the residual (2 input channels) is added in fp32 here instead of going through the tensor cores.
// The two column halves of a team take alternate 16-channel groups.
template <bool F16, int kMaxT>
__device__ __forceinline__ void xepi0_stage(const Plan& pl, const StageK& s, unsigned char* smem, uint32_t lane_base, int row, int half, int my_w,
                                            int my_v, int nw, int team) {
  const int V = pl.V, tv = pl.T0 * V, cp0 = pl.cp0;
  const bool valid = row < pl.rows && my_w < nw;
  const bool two = pl.c_in > 1;
  // kMaxT: output time steps per stage the instantiation supports (4 or 8; the host picks the smallest that covers the program)
  float xa[kMaxT], xb[kMaxT];
  const int ntp = (int)s.p1() - (int)s.p0();
  {
    const float* xw = reinterpret_cast<const float*>(smem + pl.off_xin) + my_w * pl.per_w + my_v;
    float sc0 = 0.f, sh0 = 0.f, sc1 = 0.f, sh1 = 0.f;
    if (valid) {
      const float* scale = reinterpret_cast<const float*>(smem + pl.off_scale);
      const float* shift = reinterpret_cast<const float*>(smem + pl.off_shift);
      sc0 = scale[my_v];
      sh0 = shift[my_v];
      if (two) {
        sc1 = scale[V + my_v];
        sh1 = shift[V + my_v];
      }
    }
#pragma unroll
    for (int i = 0; i < kMaxT; ++i) {
      xa[i] = xb[i] = 0.f;
      if (valid && i < ntp) {
        const float* xp = xw + pl.stride0 * ((int)s.p0() + i) * V;
        float u = xp[0], w = two ? xp[tv] : 0.f;
        if (!(fmaf(u, 0.f, w * 0.f) == 0.f)) u = w = 0.f;                 // inf / NaN (the window is already flagged by its G0 stages)
        xa[i] = fmaf(u, sc0, sh0);
        xb[i] = fmaf(w, sc1, sh1);
      }
    }
  }
  // this chunk of x1 overwrites the pose slot: every thread of the team has its poses in registers first
  if (s.flags() & SF_TEAM_SYNC) named_bar_sync(1 + team, kTeamWarps * 32);
  unsigned char* dst_row = smem + s.dst_off + (size_t)row * 16;
  const float4* tab = reinterpret_cast<const float4*>(smem + pl.off_r0tab);
  const int cgs = cp0 / 16;
  for (int cg = half; cg < cgs; cg += 2) {
#pragma unroll
    for (int i0 = 0; i0 < kMaxT; i0 += kXepiLoads) {
      if (i0 < ntp) {
        float a[kXepiLoads][16];
#pragma unroll
        for (int u = 0; u < kXepiLoads; ++u)
          if (i0 + u < ntp) tmem_ld16(lane_base + (uint32_t)(s.tmem_col() + (i0 + u) * cp0 + cg * 16), a[u]);
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < kXepiLoads; ++u) {
          const int i = i0 + u;
          if (i < ntp) {
            const float2 u2 = make_float2(xa[i], xa[i]), w2 = make_float2(xb[i], xb[i]);
#pragma unroll
            for (int g = 0; g < 4; ++g) {                    // 4 output channels per table entry (rx, ry, rb)
              const float4 rx = tab[cg * 12 + 3 * g], ry = tab[cg * 12 + 3 * g + 1], rb = tab[cg * 12 + 3 * g + 2];
              const float2 r0 = fma2(u2, make_float2(rx.x, rx.y), fma2(w2, make_float2(ry.x, ry.y), make_float2(rb.x, rb.y)));
              const float2 r1 = fma2(u2, make_float2(rx.z, rx.w), fma2(w2, make_float2(ry.z, ry.w), make_float2(rb.z, rb.w)));
              a[u][4 * g + 0] += r0.x; a[u][4 * g + 1] += r0.y; a[u][4 * g + 2] += r1.x; a[u][4 * g + 3] += r1.y;
            }
            store16<true, F16>(dst_row, i * cgs + cg, a[u]);
          }
        }
      }
    }
  }
}

// The stage tables are read from global memory (L1-resident: 5 KB per team) with one burst of five 16-byte loads per
// stage.  Reading them field by field from the kernel-parameter constant bank cost ~1.5 k cycles PER STAGE: the 20 KB plan
// does not fit the constant cache, so every stage paid several dependent constant-cache misses.
struct Tables {
  const StageK* stages[kTeams];
  const Group* groups;
  const Mma* mma;
};
__device__ __forceinline__ void load_stage(StageK* dst, const StageK* src) {
  const uint4* p = reinterpret_cast<const uint4*>(src);
  const uint4 a = __ldg(p), b = __ldg(p + 1);
  dst->w0 = a.x; dst->w1 = a.y; dst->w2 = a.z; dst->w3 = a.w;
  dst->dst_off = b.x; dst->bias_off = b.y;
}

#ifndef SF_TOK2_CVT_LOADS
#define SF_TOK2_CVT_LOADS 2
#endif
constexpr int kCvtLoads = SF_TOK2_CVT_LOADS;      // TMEM loads in flight per wait in a conversion stage

// POOL: adaptive average pooling in the token stage (shopformer_2 configs with num_tokens not dividing the last length); KW: ELL
// width of block 0's mix (5 covers the skeleton graphs, 8 the general case).  Template parameters so that the common
// instantiation carries neither path: the epilogue loop runs at the 96-register limit
template <bool F16, bool POOL, int KW, int MAXT>
__global__ void __launch_bounds__(kThreads, 1)
tokenizer2_kernel(const __grid_constant__ Plan pl, const float* __restrict__ poses, float* __restrict__ tokens, int64_t B_max, const DevCount cnt,
                  const Tables tabs) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.off_bars);
  int* poison = reinterpret_cast<int*>(smem + pl.off_flags);                 // [2][64]
  // 32-bit window / tile indices (the launcher bounds B): two registers less per index in the 96-register epilogue loop
  const int B = (int)dev_count_clamp(cnt, B_max);
  const int n_tiles = (B + pl.WT - 1) / pl.WT;

  // ------------------------------------------------------------------ one-time setup
  {
    const uint4* src = reinterpret_cast<const uint4*>(pl.const_src);
    uint4* dst = reinterpret_cast<uint4*>(smem + pl.off_const);
    for (uint32_t i = threadIdx.x; i < pl.const_bytes / 16; i += kThreads) dst[i] = __ldg(src + i);
    uint4* z = reinterpret_cast<uint4*>(smem + pl.off_P);                   // operand regions start out finite (0 * NaN = NaN)
    for (uint32_t i = threadIdx.x; i < (pl.off_bars - pl.off_P) / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 128) poison[threadIdx.x] = 0;
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (threadIdx.x == 0) {
    for (int i = 0; i < pl.n_groups; ++i) mbar_init(&bars[pl.bar_g0 + i], 1);
    for (int t = 0; t < kTeams; ++t)
      for (int i = 0; i < pl.n_stages[t]; ++i) mbar_init(&bars[pl.bar_e0[t] + i], kTeamWarps);
    for (int i = 0; i < pl.n_loads; ++i) mbar_init(&bars[pl.bar_l0 + i], 1);
    fence_mbar_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
#ifdef SF_TOK2_FINE_STAMPS
  const bool timing = g_tok2_timing_on && blockIdx.x == 0 && (warp == 0 || warp == kFirstEpiWarp || warp == kFirstEpiWarp + kTeamWarps);   // one stamping warp per role
#else
  const bool timing = g_tok2_timing_on && blockIdx.x == 0;
#endif
  // debug stamps: MMA warp in [0, 1024), first warp of team 0 in [1024, 2560), first warp of team 1 in [2560, 4096)
  int stamp_i = warp == 0 ? 0 : (warp == kFirstEpiWarp ? 1024 : 2560);
  const int stamp_end = warp == 0 ? 1022 : (warp == kFirstEpiWarp ? 2558 : 4094);
  const uint32_t stamp_it = (int)(blockIdx.x + gridDim.x) < n_tiles ? 1u : 0u;     // steady-state tile when there is one
  (void)timing; (void)stamp_i; (void)stamp_end; (void)stamp_it;                       // only used with -DSF_STAMPS

  if (warp == 0) {
    // =================================================================== MMA issue
    const uint32_t base16 = smem_u32(smem) >> 4;
    uint32_t it = 0;
    for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x, ++it) {
      const uint32_t par = it & 1u;
      for (int g = 0; g < pl.n_groups; ++g) {
        const Group gr = pl.groups[g];
        if (gr.wait_e[0] >= 0) mbar_wait(&bars[pl.bar_e0[0] + gr.wait_e[0]], par);
        if (gr.wait_e[1] >= 0) mbar_wait(&bars[pl.bar_e0[1] + gr.wait_e[1]], par);
        if (gr.wait_l >= 0) mbar_wait(&bars[pl.bar_l0 + gr.wait_l], par);
        if (gr.prev_stage >= 0 && it > 0) mbar_wait(&bars[pl.bar_e0[gr.prev_team] + gr.prev_stage], par ^ 1u);
        tc_fence_after();
        if (timing && it == stamp_it) T2_STAMP(1000 + g);
        if (elect_one()) {
          // Descriptors come straight from the parameter (constant) bank with a warp-uniform index (uniform datapath).
          // A group's issue takes 500-800 cycles however its descriptors are fetched -- measured: prefetching entry i + 1
          // before issuing MMA i changes nothing (1.82 vs 1.77 ms), and fetching the table with one coalesced load per lane
          // before the waits + shuffles to lane 0 is far slower (2.64 ms: the MMA then issues from per-thread registers in a
          // divergent branch) -- the issue rate is paced by the tensor pipe accepting the previous MMA.
          const int end = gr.first + gr.count;
          for (int i = gr.first; i < end; ++i) issue_mma(tmem, pl.mma[i], base16);
          umma_commit(&bars[pl.bar_g0 + g]);
        }
        __syncwarp();
        if (timing && it == stamp_it) T2_STAMP(3000 + g);
      }
    }
  } else if (warp == 1) {
    // =================================================================== TMA
    // (the whole warp walks the L sequence and waits -- a converged warp is parked by a blocking try_wait, a lone lane
    // spins through the scheduler's issue slots -- and lane 0 issues the copies)
    auto pose_load = [&](int tile) {
      if (tile >= n_tiles || lane != 0) return;
      const int w0 = tile * pl.WT;
      const uint32_t nw = (uint32_t)((B - w0) < pl.WT ? (B - w0) : pl.WT);
      const uint32_t bytes = nw * (uint32_t)pl.per_w * 4u;
      uint64_t* bar = &bars[pl.bar_l0 + pl.n_loads - 1];
      mbar_expect_tx(bar, bytes);
      tma_load_1d(smem + pl.off_xin, poses + (size_t)w0 * pl.per_w, bytes, bar);
    };
    pose_load(blockIdx.x);
    uint32_t it = 0;
    for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x, ++it) {
      const uint32_t par = it & 1u;
      for (int l = 0; l < pl.n_loads; ++l) {
        const Load ld = pl.loads[l];
        if (ld.wait_g >= 0) mbar_wait(&bars[pl.bar_g0 + ld.wait_g], par);
        if (ld.wait_e[0] >= 0) mbar_wait(&bars[pl.bar_e0[0] + ld.wait_e[0]], par);
        if (ld.wait_e[1] >= 0) mbar_wait(&bars[pl.bar_e0[1] + ld.wait_e[1]], par);
        if (ld.wait_g_prev >= 0 && it > 0) mbar_wait(&bars[pl.bar_g0 + ld.wait_g_prev], par ^ 1u);
        __syncwarp();
        if (ld.kind == LD_WEIGHTS) {
          if (lane == 0) {
            uint64_t* bar = &bars[pl.bar_l0 + l];
            mbar_expect_tx(bar, ld.bytes);
            tma_load_1d(smem + ld.dst_off, pl.const_src + ld.src_off, ld.bytes, bar);
          }
        } else {
          pose_load(tile + gridDim.x);          // the pose barrier is one completion ahead (the prologue load)
        }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // =================================================================== epilogue: two teams of eight warps
    // (a team = 4 TMEM lane quarters x 2 column halves; many resident warps are what hides the CUDA-core latencies)
    const int team = (warp - kFirstEpiWarp) >> 3, q = warp & 3, half = ((warp - kFirstEpiWarp) >> 2) & 1;
    const int tt = (int)threadIdx.x - kFirstEpiWarp * 32 - team * (kTeamWarps * 32);     // thread within the team
    const int row = q * 32 + lane;
    const int V = pl.V, rows = pl.rows;
    const int my_w = row / V, my_v = row - my_w * V;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const int n_st = pl.n_stages[team];
    uint32_t it = 0;
    for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x, ++it) {
      const uint32_t par = it & 1u;
      const int w_first = tile * pl.WT;
      const int nw = (B - w_first) < pl.WT ? (B - w_first) : pl.WT;
      int* pz = poison + par * 64;
      for (int e = 0; e < n_st; ++e) {
        StageK s;
        load_stage(&s, tabs.stages[team] + e);     // (hoisting this pointer out of the loop costs two registers -> spills: measured slower)
        if (timing && it == stamp_it && tt < 32) T2_FINE(5000 + e);
        // (every warp of the team waits on the mbarriers itself: try_wait carries a suspend-time hint, see tc_common.cuh; one
        // polling warp + a named barrier for the other seven measured slower, 1.93 vs 1.87 ms)
        if (s.bar_g()) mbar_wait(reinterpret_cast<uint64_t*>(smem + s.bar_g()), par);
        if (s.bar_l()) mbar_wait(reinterpret_cast<uint64_t*>(smem + s.bar_l()), par);
        if (s.bar_eo()) mbar_wait(reinterpret_cast<uint64_t*>(smem + s.bar_eo()), par);
        if (s.bar_g_prev() && it > 0) mbar_wait(reinterpret_cast<uint64_t*>(smem + s.bar_g_prev()), par ^ 1u);
        tc_fence_after();
        if (timing && it == stamp_it && tt < 32) T2_STAMP(2000 + e);
        if (s.type() == ST_CVT) {
          const bool relu = s.flags() & SF_RELU, bias = s.flags() & SF_BIAS;
          const float* bp = reinterpret_cast<const float*>(smem + s.bias_off);
          unsigned char* dst = smem + s.dst_off + (size_t)row * 16;
          // bias column of this warp's next column group, kept incrementally (no division in the loop): cg advances by 2
          const int period = s.bias_period() > 0 ? s.bias_period() : 16;
          int bcol = half * 16;
          while (bcol >= period) bcol -= period;
          // this warp's column groups: cg = half, half + 2, ...; kCvtLoads TMEM loads in flight before one wait (a TMEM load
          // takes 300+ cycles while the tensor pipe is busy with the next group: fewer, wider round trips per stage)
          for (int cg0 = half; cg0 < (int)s.n_cg(); cg0 += 2 * kCvtLoads) {
            float a[kCvtLoads][16];
#pragma unroll
            for (int b = 0; b < kCvtLoads; ++b)
              if (cg0 + 2 * b < (int)s.n_cg()) tmem_ld16(lane_base + (uint32_t)(s.tmem_col() + (cg0 + 2 * b) * 16), a[b]);
            tmem_ld_wait();
            if (timing && it == stamp_it && tt < 32 && cg0 == half) T2_FINE(8000 + e);
#pragma unroll
            for (int b = 0; b < kCvtLoads; ++b) {
              const int cg = cg0 + 2 * b;
              if (cg < (int)s.n_cg()) {
                if (bias) {
                  const float* b16 = bp + bcol;
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float4 bb = *reinterpret_cast<const float4*>(b16 + 4 * j);
                    const float2 s0 = add2(make_float2(a[b][4 * j], a[b][4 * j + 1]), make_float2(bb.x, bb.y));
                    const float2 s1 = add2(make_float2(a[b][4 * j + 2], a[b][4 * j + 3]), make_float2(bb.z, bb.w));
                    a[b][4 * j + 0] = s0.x; a[b][4 * j + 1] = s0.y; a[b][4 * j + 2] = s1.x; a[b][4 * j + 3] = s1.y;
                  }
                }
                if (relu) store16<true, F16>(dst, cg, a[b]);
                else store16<false, F16>(dst, cg, a[b]);
              }
              bcol += 32;
              while (bcol >= period) bcol -= period;
            }
            if (timing && it == stamp_it && tt < 32 && cg0 == half) T2_FINE(9000 + e);
          }
        } else if (s.type() == ST_G0) {
          g0_stage<KW, F16>(pl, s, smem, row, half, my_w, my_v, nw, pz);
        } else if (s.type() == ST_XEPI0) {
          xepi0_stage<F16, MAXT>(pl, s, smem, lane_base, row, half, my_w, my_v, nw, team);
        } else {   // ST_TOKENS
          const float* bp = reinterpret_cast<const float*>(smem + s.bias_off);
          float* stg = reinterpret_cast<float*>(smem + pl.off_stage_tok);
          const bool live = row < rows && my_w < nw;
          const bool poisoned = live && pz[my_w] != 0;
          // (the staging area is free: the previous tile's token stage completed only after its bulk store had read it)
          if (POOL) {
            // adaptive average pooling over time (T_last -> pool tokens): token j = mean of relu(acc + b) over the time
            // steps [floor(j T / P), ceil((j + 1) T / P)); every time step of a row sits in this thread's TMEM lane, so
            // the two column halves of the team take alternate output tokens and sum in time order
            const int cpt = pl.cp_last >> 4;                                // 16-column groups per time step
            for (int j = half; j < pl.pool; j += 2) {
              const int t0 = (j * pl.T_last) / pl.pool, t1 = ((j + 1) * pl.T_last + pl.pool - 1) / pl.pool;
              const float inv = 1.f / (float)(t1 - t0);
              for (int g = 0; g < cpt; ++g) {
                float acc[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) acc[e] = 0.f;
                for (int t = t0; t < t1; ++t) {
                  float a[16];
                  tmem_ld16(lane_base + (uint32_t)(s.tmem_col() + (t * cpt + g) * 16), a);
                  tmem_ld_wait();
#pragma unroll
                  for (int e = 0; e < 16; ++e) {
                    const int c = g * 16 + e;
                    acc[e] += fmaxf(a[e] + (c < pl.c_last ? bp[c] : 0.f), 0.f);
                  }
                }
                if (live) {
#pragma unroll
                  for (int e = 0; e < 16; ++e) {
                    const int c = g * 16 + e;
                    if (c < pl.c_last) {
                      float y = acc[e] * inv;
                      if (poisoned) y = __int_as_float(0x7fc00000);
                      stg[(my_w * pl.S_out + j) * pl.d_tok + c * V + my_v] = y;
                    }
                  }
                }
              }
            }
          } else
          for (int cg = half; cg < (int)s.n_cg(); cg += 2) {
            float a[16];
            tmem_ld16(lane_base + (uint32_t)(s.tmem_col() + cg * 16), a);
            tmem_ld_wait();
            if (live) {
              const int t = (cg * 16) / pl.cp_last, c0 = cg * 16 - t * pl.cp_last;      // cp_last is a multiple of 16
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int c = c0 + j;
                if (c < pl.c_last) {
                  float y = fmaxf(a[j] + bp[c], 0.f);
                  if (poisoned) y = __int_as_float(0x7fc00000);
                  stg[(my_w * pl.S_out + t) * pl.d_tok + c * V + my_v] = y;
                }
              }
            }
          }
          fence_proxy_async();
          named_bar_sync(1 + team, kTeamWarps * 32);
          if (tt < 64) pz[tt] = 0;          // every thread has read its flag; this buffer is next used two tiles from now
          if (tt == 0) {
            bulk_store(tokens + (size_t)w_first * pl.S_out * pl.d_tok, stg, (uint32_t)nw * (uint32_t)(pl.S_out * pl.d_tok) * 4u);
            bulk_wait_read();               // the staging area aliases activation storage: free it before the stage completes
          }
        }
        if (timing && it == stamp_it && tt < 32) T2_FINE(6000 + e);
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(reinterpret_cast<uint64_t*>(smem + s.bar_self()));
        if (timing && it == stamp_it && tt < 32) T2_FINE(7000 + e);
      }
    }
    if (tt == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- per-model cache of uploaded programs (one per window length)
struct Uploaded {
  Program prog;
  unsigned char* tables_dev = nullptr;       // [stages of team 0][stages of team 1]
  Tables tabs;
  ~Uploaded() {
    if (tables_dev) cudaFree(tables_dev);
  }
};
struct Cache {
  std::mutex mu;
  std::map<int, Uploaded*> by_T;
};

}  // namespace

struct Tok2State {
  t2::Static st;
  unsigned char* blob_dev = nullptr;
  Cache cache;
};

Tok2State* tok2_create(const Tokenizer& host_tok, int pool_tokens, bool upload, bool allow_f16) {
  Tok2State* s = new Tok2State();
  t2::build_static(host_tok, pool_tokens, allow_f16, &s->st);
  if (s->st.ok && upload) {
    if (cudaMalloc((void**)&s->blob_dev, s->st.blob.size()) != cudaSuccess ||
        cudaMemcpy(s->blob_dev, s->st.blob.data(), s->st.blob.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaGetLastError();
      s->st.ok = false;
      s->st.why = "uploading the operand images failed";
    }
  }
  return s;
}

void tok2_destroy(Tok2State* s) {
  if (!s) return;
  for (auto& kv : s->cache.by_T) delete kv.second;
  if (s->blob_dev) cudaFree(s->blob_dev);
  delete s;
}

const t2::Static* tok2_static(const Tok2State* s) { return s ? &s->st : nullptr; }

static Uploaded* tok2_program(const sf_model* m, int T) {
  Tok2State* s = m->tok2;
  if (!s || !s->st.ok || !s->blob_dev) return nullptr;
  std::lock_guard<std::mutex> lk(s->cache.mu);
  auto it = s->cache.by_T.find(T);
  if (it != s->cache.by_T.end()) return it->second;
  Uploaded* u = new Uploaded();
  t2::build_program(s->st, T, m->max_smem_optin - 256, &u->prog);    // minus the kernel's static shared memory (4 bytes) and the 128-byte alignment of the dynamic part
  if (u->prog.ok) {
    u->prog.plan.const_src = s->blob_dev;
    const size_t n0 = u->prog.stages[0].size(), n1 = u->prog.stages[1].size();
    const size_t ng = u->prog.groups.size(), nm = u->prog.mma.size();
    std::vector<StageK> sk(n0 + n1);
    bool packed = true;
    for (size_t i = 0; i < n0; ++i) packed &= pack_stage(u->prog.plan.stages[0][i], &sk[i]);
    for (size_t i = 0; i < n1; ++i) packed &= pack_stage(u->prog.plan.stages[1][i], &sk[n0 + i]);
    const size_t off_g = (n0 + n1) * sizeof(StageK), off_m = off_g + ng * sizeof(Group);
    if (!packed || cudaMalloc((void**)&u->tables_dev, off_m + (nm + 64) * sizeof(Mma)) != cudaSuccess ||
        cudaMemset(u->tables_dev, 0, off_m + (nm + 64) * sizeof(Mma)) != cudaSuccess ||
        cudaMemcpy(u->tables_dev, sk.data(), (n0 + n1) * sizeof(StageK), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(u->tables_dev + off_g, u->prog.groups.data(), ng * sizeof(Group), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(u->tables_dev + off_m, u->prog.mma.data(), nm * sizeof(Mma), cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaGetLastError();
      u->prog.ok = false;
      u->prog.why = packed ? "uploading the stage tables failed" : "a stage field does not fit the packed stage table";
    }
    u->tabs.stages[0] = reinterpret_cast<const StageK*>(u->tables_dev);
    u->tabs.stages[1] = reinterpret_cast<const StageK*>(u->tables_dev + n0 * sizeof(StageK));
    u->tabs.groups = reinterpret_cast<const Group*>(u->tables_dev + off_g);
    u->tabs.mma = reinterpret_cast<const Mma*>(u->tables_dev + off_m);
  }
  s->cache.by_T[T] = u;
  return u;
}

bool tokenizer2_f16(const sf_model* m) { return m->tok2 && m->tok2->st.ok && m->tok2->st.f16; }

bool tokenizer2_supported(const sf_model* m, int T) {
  if (getenv("SF_TOK2_OFF")) return false;
  Uploaded* u = tok2_program(m, T);
  return u && u->prog.ok;
}

const char* tokenizer2_why(const sf_model* m, int T) {
  if (!m->tok2) return "no tokenizer-v2 state";
  if (!m->tok2->st.ok) return m->tok2->st.why.c_str();
  Uploaded* u = tok2_program(m, T);
  return u ? u->prog.why.c_str() : "";
}

int launch_tokenizer2(const sf_model* m, const float* poses, int64_t B, int T, float* tokens, cudaStream_t st, DevCount cnt) {
  if (B == 0) return SF_OK;
  Uploaded* u = tok2_program(m, T);
  SF_REQUIRE(u && u->prog.ok, SF_E_UNSUPPORTED, "tokenizer v2 does not cover this shape: %s", tokenizer2_why(m, T));
  SF_REQUIRE(((uintptr_t)poses & 15) == 0 && ((uintptr_t)tokens & 15) == 0, SF_E_INVALID,
             "pose / token buffers must be 16-byte aligned (TMA bulk copies)");
  const Plan& pl = u->prog.plan;
  SF_REQUIRE(B <= (int64_t)0x7FFFFF00, SF_E_INVALID, "tokenizer v2 takes at most 2^31 - 256 windows per launch");
  const int64_t n_tiles = (B + pl.WT - 1) / pl.WT;
  const int grid = (int)std::min<int64_t>(n_tiles, (int64_t)m->sm_count);
  count_launch(LK_TOK2);
  auto go = [&](auto kern) -> int {
    SF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
    kern<<<grid, kThreads, pl.smem_bytes, st>>>(pl, poses, tokens, B, cnt, u->tabs);
    return SF_OK;
  };
  const bool f16 = m->tok2->st.f16, pool = pl.pool > 0, kw5 = pl.ell_width <= 5;
  int maxt = 0;                                   // longest block-0 output stage of this program
  for (int t = 0; t < kTeams; ++t)
    for (int e = 0; e < pl.n_stages[t]; ++e)
      if (pl.stages[t][e].type == ST_XEPI0) maxt = std::max(maxt, (int)(pl.stages[t][e].p1 - pl.stages[t][e].p0));
  const bool t4 = maxt <= 4;
  int rc;
#define SF_T2_GO(F, P) (kw5 ? (t4 ? go(tokenizer2_kernel<F, P, 5, 4>) : go(tokenizer2_kernel<F, P, 5, 8>)) \
                            : (t4 ? go(tokenizer2_kernel<F, P, 8, 4>) : go(tokenizer2_kernel<F, P, 8, 8>)))
  if (f16) rc = pool ? SF_T2_GO(true, true) : SF_T2_GO(true, false);
  else rc = pool ? SF_T2_GO(false, true) : SF_T2_GO(false, false);
#undef SF_T2_GO
  if (rc) return rc;
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}

}  // namespace sf

// debugging aid (not part of the C ABI): stage / group start stamps of CTA 0's second tile, pairs of (id, clock64)
extern "C" int sfdbg_tokenizer2_timing(int enable, long long* out_host, int n) {
  int on = enable;
  if (cudaMemcpyToSymbol(sf::g_tok2_timing_on, &on, sizeof(int)) != cudaSuccess) return -1;
  if (out_host && n > 0) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out_host, sf::g_tok2_timing, sizeof(long long) * (n < 4096 ? n : 4096)) != cudaSuccess) return -1;
  }
  return 0;
}
