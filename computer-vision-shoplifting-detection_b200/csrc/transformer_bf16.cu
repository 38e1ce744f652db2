// bf16 tcgen05 encoder/decoder transformer + fused reconstruction-error score.
//
// Same maths as transformer_fp32.cu (reference shopformer/models/transformer.py:60-196,
// 304-329, shopformer/models/shopformer.py:150-178; shopformer_2/models/transformer.py:105-194,
// shopformer_2/models/shopformer.py:178-186).  The host flattens either variant into an op list
// (model.cu: "tensor-core transformer program"); this kernel interprets it for one tile of rows at a time:
//
//   * one CTA = 128 TMEM lanes = 4 lane groups x 32 rows; a lane group holds floor(32/S) whole windows
//     (S = 2..4 tokens each), so the S x S attention of a window never leaves a warp: QK^T and PV are
//     computed straight out of TMEM with warp shuffles over the token axis;
//   * every nn.Linear is a chain of tcgen05.mma (M=128, N = padded width, K = 16 per instruction) with the
//     bf16 activation operand in shared memory (planar-chunk layout, tc_common.cuh) and the weight image
//     streamed L2 -> smem by TMA (cp.async.bulk + mbarrier complete_tx) through a 2-slot ring, one GEMM ahead;
//   * the fp32 residual stream of a row lives in REGISTERS (four warps share a row: every 4th 16-column
//     group each); bias, residual add, LayerNorm (two-pass, partial sums exchanged through smem), ReLU / exact
//     GELU and the bf16 down-conversion of the next operand all happen in the TMEM epilogue;
//   * q, k, v accumulate side by side in TMEM columns [0,dp) [dp,2dp) [2dp,3dp); the FFN hidden layer is
//     processed in chunks of <= 128 columns with the second GEMM accumulating in TMEM across chunks;
//   * the squared error against the score target is reduced in the last epilogue: only B floats are written.
#include <algorithm>
#include <cstdlib>

#include "sf_internal.h"
#include "tc_common.cuh"

namespace sf {
__device__ long long g_xf_timing[1024];
__device__ int g_xf_timing_on = 0;
namespace {

using namespace tc;

// The clock64 stamps of the timeline tools (profiles/*_timing.py) are compiled in only with `make EXTRA=-DSF_STAMPS`: even
// disabled at run time they cost the production kernels 2-5 % (tokenizer v2 1.717 -> 1.631 ms, transformer 1.086 -> 1.065 ms,
// one-window tokenizer on config B 5.08 -> 4.89 ms per 65,536 windows).
#if !defined(SF_STAMPS) && !defined(SF_TOK2_FINE_STAMPS)
#define XF_STAMP(id) do { } while (0)
#else
#define XF_STAMP(id)                                                             \
  do {                                                                           \
    if (timing && threadIdx.x == 0 && stamp_i < 1022) {                          \
      g_xf_timing[stamp_i++] = (long long)(id);                                  \
      g_xf_timing[stamp_i++] = clock64();                                        \
    }                                                                            \
  } while (0)
#endif

constexpr int kThreads = 512;
constexpr int kParts = kThreads / 128;  // warps sharing one 32-row lane group: each owns every kParts-th 16-column group
constexpr int kSlots = 3;              // 16-column groups per thread: widths up to 192
constexpr int kSMax = 4;
constexpr float kLnEps = 1e-5f;
// per-slot parameter block (floats): [bias | bias_q | bias_k | ln_gamma | ln_beta], 160 each
constexpr int kPW = 160, kParamFloats = 5 * kPW;

struct XfGeo {
  int S, wpw, rows_per_warp, win_per_tile;
  uint32_t off_aop, off_hop, off_mem, off_w, off_red, off_ops, smem_bytes;
  int ops_in_smem;                     // the op program is copied into shared memory when it fits (else read from global)
  int slot_bytes, plane;               // plane = 128 rows * 16 B
  int param_off;                       // byte offset of the fp32 parameter block inside a ring slot
};

__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
  return ((smem_addr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ uint64_t desc_join(uint32_t lo) { return ((uint64_t)((128u >> 4) | (1u << 14)) << 32) | lo; }
template <bool F16>
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_op2<F16>(f[0], f[1]), pack_op2<F16>(f[2], f[3]), pack_op2<F16>(f[4], f[5]), pack_op2<F16>(f[6], f[7]));
}
// packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2): two exact fp32 operations per issue slot -- the epilogues are bound by
// their instruction stream, and every result is bit-identical to the scalar form
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
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 d;
  asm("{\n"
      ".reg .b64 ra, rb, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "sub.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n"
      ".reg .b64 ra, rb, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "mul.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 pr(const float* p) { return make_float2(p[0], p[1]); }
__device__ __forceinline__ void put(float* p, float2 v) { p[0] = v.x; p[1] = v.y; }
// fp32 pairs -> packed 16-bit pair through ReLU (one conversion instruction, no separate max)
template <bool F16>
__device__ __forceinline__ uint32_t pack_relu2(float lo, float hi) {
  uint32_t d;
  if (F16) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <bool F16>
__device__ __forceinline__ uint4 pack8_relu(const float* f) {
  return make_uint4(pack_relu2<F16>(f[0], f[1]), pack_relu2<F16>(f[2], f[3]), pack_relu2<F16>(f[4], f[5]), pack_relu2<F16>(f[6], f[7]));
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <int KS>
__device__ __forceinline__ void gemm_issue(uint32_t d, uint32_t alo0, uint32_t blo0, uint32_t astep, uint32_t bstep,
                                           uint32_t idesc, uint32_t accumulate) {
  uint32_t al[KS], bl[KS];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    al[ks] = alo0 + ks * astep;
    bl[ks] = blo0 + ks * bstep;
  }
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) umma_bf16(d, desc_join(al[ks]), desc_join(bl[ks]), idesc, (accumulate || ks > 0) ? 1u : 0u);
}

// busy poll without a suspend hint: lowest wake-up latency when nothing else on the SM wants the issue slots
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}

// 16 consecutive floats of a 16-byte aligned global row; columns >= lim (lim % 4 == 0) or !ok read as zero
__device__ __forceinline__ void ldg16(const float* __restrict__ p, int c0, int lim, bool ok, float* out) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok && c0 + 4 * e < lim) v = __ldg(reinterpret_cast<const float4*>(p + c0 + 4 * e));
    out[4 * e + 0] = v.x;
    out[4 * e + 1] = v.y;
    out[4 * e + 2] = v.z;
    out[4 * e + 3] = v.w;
  }
}

// barrier over the kParts warps that share lane group `lane_grp` (ids 1..4; id 0 is __syncthreads)
__device__ __forceinline__ void group_barrier(int lane_grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + lane_grp), "r"(kParts * 32) : "memory");
}
__device__ __forceinline__ float red_sum(const float* red, int row) {
  float t = 0.f;
#pragma unroll
  for (int p = 0; p < kParts; ++p) t += red[p * 128 + row];
  return t;
}

// 16 consecutive floats of a 16-byte aligned shared-memory array (bias blocks): four 128-bit loads
__device__ __forceinline__ void lds16(const float* p, float* out) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float4 v = *reinterpret_cast<const float4*>(p + 4 * e);
    out[4 * e + 0] = v.x; out[4 * e + 1] = v.y; out[4 * e + 2] = v.z; out[4 * e + 3] = v.w;
  }
}

// write one 16-column group of a row into a planar-chunk operand buffer
template <bool F16>
__device__ __forceinline__ void store_group(unsigned char* buf, int plane, int row, int g, const float* y) {
  *reinterpret_cast<uint4*>(buf + (size_t)(2 * g) * plane + row * 16) = pack8<F16>(y);
  *reinterpret_cast<uint4*>(buf + (size_t)(2 * g + 1) * plane + row * 16) = pack8<F16>(y + 8);
}

template <bool F16>
__device__ __forceinline__ void store_group_relu(unsigned char* buf, int plane, int row, int g, const float* y) {
  *reinterpret_cast<uint4*>(buf + (size_t)(2 * g) * plane + row * 16) = pack8_relu<F16>(y);
  *reinterpret_cast<uint4*>(buf + (size_t)(2 * g + 1) * plane + row * 16) = pack8_relu<F16>(y + 8);
}

// F16: the 16-bit operands (activations written by the epilogues and the weight images) are fp16 instead of bf16
// (the reduction mode stays a run-time value: as a template parameter it made config C 0.9 % slower)
// V1: shopformer/ (post-LN, ReLU, positional encoding added to the score target); else shopformer_2/ (pre-LN, exact GELU).
// RECON: the reconstructed tokens are written out (sf_reconstruct_tokens / explicit recon buffer); the scoring path compiles the
// stores, the pointer and its predicates away (122 instead of 128 registers: 1.3 % on sf_score_windows)
template <bool F16, bool RECON, bool V1>
__global__ void __launch_bounds__(kThreads, 1)
transformer_bf16_kernel(const XfProgram prog, const __grid_constant__ XfGeo geo, const __grid_constant__ Transformer xf,
                        const float* __restrict__ tokens, int64_t B_max, int reduction, float* __restrict__ recon_out_param,
                        float* __restrict__ scores, const DevCount cnt, const int flags_rt) {
  // experiment switches of profiles/r2_summary.md (1: one polling warp + bar.sync, 2 / 8: busy test_wait spin, 4: q / k / v as three
  // phases): a run-time value only with -DSF_XF_FLAGS_RUNTIME (SF_XF_FLAGS in the environment), otherwise the constant 0 so
  // that the branches fold away
#ifdef SF_XF_FLAGS_RUNTIME
  const int flags = flags_rt;
#else
  constexpr int flags = 0;
  (void)flags_rt;
#endif
  float* const recon_out = RECON ? recon_out_param : nullptr;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar, cbar, wbar[2];      // MMA completion (phase / first GEMM of a chain); TMA completion per weight-ring slot
  __shared__ uint32_t tmem_base_s;
  unsigned char* sAop = smem + geo.off_aop;
  unsigned char* sHop = smem + geo.off_hop;
  unsigned char* sMem = smem + geo.off_mem;
  unsigned char* sW = smem + geo.off_w;                       // 2 ring slots
  float* red = reinterpret_cast<float*>(smem + geo.off_red);  // LayerNorm partial sums: [2 buffers][sum, sumsq][kParts][128 rows]
  float* xsc = red + 4 * kParts * 128;                            // [kParts][kSMax][128 rows] partial attention logits
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform on purpose (uniform datapath)
  const int lane = threadIdx.x & 31;
  const int lane_grp = warp & 3, part = warp >> 2;
  const int row = lane_grp * 32 + lane;
  const int S = geo.S, d = xf.d_model, dt = xf.d_tok, dp = prog.dp, plane = geo.plane;
  const int tok_s = lane % S, win_l = lane / S, wb = win_l * S;       // token index, window in warp, first lane of window
  const bool lane_ok = lane < geo.rows_per_warp;
  const uint32_t lane_addr = (uint32_t)(lane_grp * 32) << 16;

  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&cbar, 1);
    mbar_init(&wbar[0], 1);
    mbar_init(&wbar[1], 1);
    fence_mbar_init();
  }
  // zero the operand buffers once: padded columns / idle rows must stay finite (NaN * 0 = NaN in the MMA)
  // (also the parameter blocks of the ring slots: LayerNorm gamma / beta beyond d_model stay 0)
  for (int i = threadIdx.x; i < (int)((geo.off_red - geo.off_aop) >> 4); i += kThreads)
    reinterpret_cast<uint4*>(smem + geo.off_aop)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint32_t parity = 0, wpar = 0, cpar = 0;
  const bool timing = g_xf_timing_on && blockIdx.x == 0;
  int stamp_i = 0;
  (void)timing; (void)stamp_i;                       // only used with -DSF_STAMPS

  // index of the first / next GEMM op (for the weight ring)
  const int n_ops = prog.n_ops;
  const XfOp* __restrict__ ops = prog.ops;
  if (geo.ops_in_smem) {      // every op costs several dependent global loads otherwise (the interpreter walks the list per tile)
    static_assert(sizeof(XfOp) % 16 == 0, "XfOp is copied in 16-byte pieces");
    const uint4* src = reinterpret_cast<const uint4*>(prog.ops);
    uint4* dst = reinterpret_cast<uint4*>(smem + geo.off_ops);
    for (int i = threadIdx.x; i < n_ops * (int)(sizeof(XfOp) / 16); i += kThreads) dst[i] = __ldg(src + i);
    __syncthreads();
    ops = reinterpret_cast<const XfOp*>(smem + geo.off_ops);
  }

  // 32-bit window / tile indices (the launcher bounds B): fewer live registers in a kernel that runs at the 128-register limit
  const int B = (int)dev_count_clamp(cnt, B_max);
  const int n_tiles = (B + geo.win_per_tile - 1) / geo.win_per_tile;
  for (int tile = (int)blockIdx.x; tile < n_tiles; tile += (int)gridDim.x) {
    const int window = tile * geo.win_per_tile + lane_grp * geo.wpw + win_l;
    const bool valid = lane_ok && window < B;
    const float* tok_row = tokens + ((size_t)(valid ? window : 0) * S + tok_s) * dt;
    {   // the next tile's tokens (one contiguous range) start their way from HBM to L2 now, a whole tile ahead of use
      const int w0 = (tile + (int)gridDim.x) * geo.win_per_tile;
      if (w0 < B) {
        const int bytes = (B - w0 < geo.win_per_tile ? B - w0 : geo.win_per_tile) * S * dt * (int)sizeof(float);
        const char* base = reinterpret_cast<const char*>(tokens + (size_t)w0 * S * dt);
        for (int o = (int)threadIdx.x * 128; o < bytes; o += kThreads * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
      }
    }
    float st[kSlots][16];
#pragma unroll
    for (int i = 0; i < kSlots; ++i)
#pragma unroll
      for (int q = 0; q < 16; ++q) st[i][q] = 0.f;

    int slot = 0, ln_buf = 0;
    // weights + the op's fp32 parameters (biases, LayerNorm affine) go into ring slot `sl`
    // TMA (bulk async copies issued by thread 0, completion on the slot's mbarrier): the weight image and the op's
    // fp32 parameters (bias, q-bias for attention, LayerNorm gamma / beta padded to 160 columns)
    auto stage_op = [&](int idx, int sl) {
      if (threadIdx.x != 0) return;
      const XfOp o = ops[idx];
      unsigned char* base = sW + sl * geo.slot_bytes;
      float* pp = reinterpret_cast<float*>(base + geo.param_off);
      const uint32_t wbytes = (uint32_t)(o.K * o.N * 2), bbytes = (uint32_t)o.N * 4u, lbytes = kPW * 4u;
      const float* qb = o.epi == XE_ATTN ? ops[idx - 2].bias : nullptr;
      mbar_expect_tx(&wbar[sl], wbytes + (o.bias ? bbytes : 0u) + (qb ? bbytes : 0u) + (o.ln_g ? 2u * lbytes : 0u));
      tma_load_1d(base, o.w, wbytes, &wbar[sl]);
      if (o.bias) tma_load_1d(pp, o.bias, bbytes, &wbar[sl]);
      if (qb) tma_load_1d(pp + kPW, qb, bbytes, &wbar[sl]);
      if (o.ln_g) {
        tma_load_1d(pp + 3 * kPW, o.ln_g, lbytes, &wbar[sl]);
        tma_load_1d(pp + 4 * kPW, o.ln_b, lbytes, &wbar[sl]);
      }
    };
    {   // prefetch the first GEMM
      int nx = 0;
      while (nx < n_ops && ops[nx].type != XF_GEMM) ++nx;
      if (nx < n_ops) stage_op(nx, 0);
    }

    for (int oi = 0; oi < n_ops; ++oi) {
      XfOp op = ops[oi];
      int post = op.post;
      XF_STAMP(oi * 8 + 0);
      const float* pbase = nullptr;              // staged parameters of this op (GEMM ops only)
      // ------------------------------------------------------------------ stream initialisation
      if (op.type == XF_INIT) {
        if (op.init_mode == XI_TOK_TO_AOP) {
          const int ng = prog.dtp >> 4;
#pragma unroll
          for (int i = 0; i < kSlots; ++i) {
            const int g = kParts * i + part;
            if (g < ng) {
              float y[16];
              ldg16(tok_row, g * 16, dt, valid, y);
              store_group<F16>(sAop, plane, row, g, y);
            }
          }
        } else {
          const bool shift = op.init_mode == XI_SHIFT_TOK_PE;
          const int ng = dp >> 4;
          // every token load of the row goes out before anything waits on one (a tile's tokens are a cold, latency-bound
          // read; the positional encoding is cache resident)
#pragma unroll
          for (int i = 0; i < kSlots; ++i) {
            const int g = kParts * i + part;
            if (g < ng) ldg16(shift ? tok_row - dt : tok_row, g * 16, d, valid && (!shift || tok_s > 0), st[i]);   // zero start token + shift by one
          }
#pragma unroll
          for (int i = 0; i < kSlots; ++i) {
            const int g = kParts * i + part;
            if (g < ng) {
              float pe16[16];
              ldg16(xf.pe + tok_s * d, g * 16, d, valid, pe16);
#pragma unroll
              for (int q = 0; q < 16; q += 2) put(st[i] + q, add2(pr(st[i] + q), pr(pe16 + q)));
            }
          }
        }
      }
      // ------------------------------------------------------------------ GEMM + epilogue
      if (op.type == XF_GEMM) {
        // One MMA phase = `nch` GEMMs (1, or q / k / v of an attention block: XfOp::chain) issued back to back; only the
        // last one's completion is waited for.  Two weight slots: GEMM c of the phase uses slot ^ (c & 1); its image is
        // staged (TMA) as soon as the slot's previous reader has completed -- the GEMM before this phase for c = 1
        // (complete since its epilogue ran), GEMM c - 2 of this phase for c = 2 (its own commit on `cbar`).
        const int nch = (flags & 4) ? 1 : (op.chain > 1 ? op.chain : 1);
        if (warp == 0) mbar_wait(&wbar[slot], (wpar >> slot) & 1u);      // the first GEMM's weights + parameters have landed
        wpar ^= 1u << slot;
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        XF_STAMP(oi * 8 + 1);
        auto issue = [&](const XfOp& o, int sl, uint64_t* commit_to) {     // elected thread of warp 0
          tc_fence_after();
          const unsigned char* a = o.a_src == XS_AOP ? sAop : (o.a_src == XS_HOP ? sHop : sMem);
          const uint32_t idesc = make_idesc(128, o.N, false, F16, F16);
          const uint32_t w_plane = (uint32_t)o.N * 16u;
          uint32_t alo = desc_lo(smem_u32(a), (uint32_t)plane);
          uint32_t blo = desc_lo(smem_u32(sW + sl * geo.slot_bytes), w_plane);
          const uint32_t dd = tmem + (uint32_t)o.tmem_col, astep = (2u * (uint32_t)plane) >> 4, bstep = (2u * w_plane) >> 4;
          const uint32_t accf = o.accumulate ? 1u : 0u;
          switch (o.K >> 4) {
            case 9: gemm_issue<9>(dd, alo, blo, astep, bstep, idesc, accf); break;
            case 4: gemm_issue<4>(dd, alo, blo, astep, bstep, idesc, accf); break;
            case 8: gemm_issue<8>(dd, alo, blo, astep, bstep, idesc, accf); break;
            case 10: gemm_issue<10>(dd, alo, blo, astep, bstep, idesc, accf); break;
            case 2: gemm_issue<2>(dd, alo, blo, astep, bstep, idesc, accf); break;
            default:
              for (int ks = 0; ks < (o.K >> 4); ++ks) {
                umma_bf16(dd, desc_join(alo), desc_join(blo), idesc, (accf || ks > 0) ? 1u : 0u);
                alo += astep;
                blo += bstep;
              }
          }
          if (commit_to) umma_commit(commit_to);
        };
        if (nch == 1) {
          if (warp == 0 && elect_one()) {            // uniform-datapath issue: descriptors live in uniform registers
            issue(op, slot, &bar);
            XF_STAMP(oi * 8 + 2);
          }
          {   // prefetch the next GEMM's weights into the other ring slot (its previous user has completed)
            int nx = oi + 1;
            while (nx < n_ops && ops[nx].type != XF_GEMM) ++nx;
            if (nx < n_ops) stage_op(nx, slot ^ 1);
          }
          pbase = reinterpret_cast<const float*>(sW + slot * geo.slot_bytes + geo.param_off);
          slot ^= 1;
        } else {
          // chain of three (nch == 3 is the only other value the host emits): slots  s, s^1, s
          if (warp == 0) {
            stage_op(oi + 1, slot ^ 1);                                   // k: the slot of the GEMM before this phase
            __syncwarp();
            if (elect_one()) issue(op, slot, &cbar);                      // q
            __syncwarp();
            mbar_wait(&wbar[slot ^ 1], (wpar >> (slot ^ 1)) & 1u);
            if (elect_one()) issue(ops[oi + 1], slot ^ 1, nullptr);       // k
            __syncwarp();
            mbar_wait(&cbar, cpar);                                       // q has completed: its slot takes v
            stage_op(oi + 2, slot);
            __syncwarp();
            mbar_wait(&wbar[slot], (wpar >> slot) & 1u);                  // (wpar was flipped for q above: this is v's phase)
            if (elect_one()) {
              issue(ops[oi + 2], slot, &bar);                             // v
              XF_STAMP(oi * 8 + 2);
            }
            __syncwarp();
          }
          cpar ^= 1u;
          // every warp reads v's parameter block (q / v bias): observe its TMA completion directly
          if (warp != 0) mbar_wait(&wbar[slot], (wpar >> slot) & 1u);
          wpar ^= 1u << (slot ^ 1);                                       // k
          wpar ^= 1u << slot;                                             // v
          pbase = reinterpret_cast<const float*>(sW + slot * geo.slot_bytes + geo.param_off);
          oi += 2;
          op = ops[oi];
          post = op.post;
          slot ^= 1;                                                      // three GEMMs = one net toggle
        }
        if (op.epi == XE_SCORE) {
          // the score target (tokens, plus the positional encoding for shopformer/) is fetched into the stream registers
          // -- dead since the last LayerNorm wrote its operand -- while the projection's MMAs run
          const int ng = op.N >> 4;
#pragma unroll
          for (int i = 0; i < kSlots; ++i) {
            const int g = kParts * i + part;
            if (g < ng) ldg16(tok_row, g * 16, dt, valid, st[i]);
          }
          if (V1) {
#pragma unroll
            for (int i = 0; i < kSlots; ++i) {
              const int g = kParts * i + part;
              if (g < ng) {
                float pe16[16];
                ldg16(xf.pe_score + tok_s * dt, g * 16, dt, valid, pe16);
#pragma unroll
                for (int q = 0; q < 16; q += 2) put(st[i] + q, add2(pr(st[i] + q), pr(pe16 + q)));
              }
            }
          }
        }
        if (flags & 1) {
          if (warp == 0) {                           // one polling warp; the others block in the hardware barrier
            if (flags & 2) mbar_spin(&bar, parity);
            else mbar_wait(&bar, parity);
          }
          __syncthreads();
        } else {
          if (flags & 8) mbar_spin(&bar, parity);
          else mbar_wait(&bar, parity);              // every warp sleeps on the mbarrier itself (measured faster than one
        }                                            // polling warp + bar.sync: 1.197 -> 1.144 ms per 65,536 windows)
        parity ^= 1;
        if (nch > 1) {   // the next GEMM (out_proj) goes to the slot k used; k has completed with v
          int nx = oi + 1;
          while (nx < n_ops && ops[nx].type != XF_GEMM) ++nx;
          if (nx < n_ops) stage_op(nx, slot);
        }
        tc_fence_after();
        XF_STAMP(oi * 8 + 3);

        if (op.epi == XE_ATTN) {
          // ---- softmax(q k^T / sqrt(hd)) v per (row, head).  The kParts warps of a lane group split the heads
          // (H >= kParts) or the columns of a head (H < kParts; partial logits are summed through smem).
          const int H = xf.heads, hd = d / H;
          const int P = H >= kParts ? 1 : kParts / H;          // column parts per head
          const int hpw = H >= kParts ? H / kParts : 1;        // heads per warp
          const int n4 = hd >> 2;
          const float scale = rsqrtf((float)hd);
          int rot[kSMax];                                      // lane of the j-th next token of this window
#pragma unroll
          for (int j = 0; j < kSMax; ++j) rot[j] = (wb + (tok_s + j) % S) & 31;
          const float* bq = pbase + kPW;
          const float* bv = pbase;
          for (int hh = 0; hh < hpw; ++hh) {
            const int h = P == 1 ? part * hpw + hh : part / P;
            const int cp = P == 1 ? 0 : part % P;
            const int f_lo = (cp * n4) / P, f_hi = ((cp + 1) * n4) / P;
            const int c0 = h * hd + 4 * f_lo, ncol = 4 * (f_hi - f_lo);
            float sc[kSMax];
            float2 sc2[kSMax];
#pragma unroll
            for (int j = 0; j < kSMax; ++j) sc2[j] = make_float2(0.f, 0.f);
            // 16 accumulator columns per TMEM round trip (columns past the range are loaded but not used).
            // The key bias adds the same q.b_k to every logit of a row -- softmax cancels it, so it is skipped.
            // Own token first, then the other S-1 tokens of the window in rotating order (S-1 shuffles per value);
            // products accumulate as fp32 pairs (even / odd columns), summed once at the end.
            for (int c16 = 0; c16 < ncol; c16 += 16) {
              float q16[16], k16[16];
              tmem_ld16(tmem + lane_addr + (uint32_t)(c0 + c16), q16);
              tmem_ld16(tmem + lane_addr + (uint32_t)(dp + c0 + c16), k16);
              tmem_ld_wait();
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4)
                if (c16 + 4 * e4 < ncol) {
                  const float4 bq4 = *reinterpret_cast<const float4*>(bq + c0 + c16 + 4 * e4);
                  const float* k4 = k16 + 4 * e4;
                  const float2 qa = add2(pr(q16 + 4 * e4), make_float2(bq4.x, bq4.y));
                  const float2 qb = add2(pr(q16 + 4 * e4 + 2), make_float2(bq4.z, bq4.w));
                  sc2[0] = fma2(qa, pr(k4), sc2[0]);
                  sc2[0] = fma2(qb, pr(k4 + 2), sc2[0]);
#pragma unroll
                  for (int j = 1; j < kSMax; ++j)
                    if (j < S) {
                      float ks[4];
#pragma unroll
                      for (int e = 0; e < 4; ++e) ks[e] = __shfl_sync(0xffffffffu, k4[e], rot[j]);
                      sc2[j] = fma2(qa, pr(ks), sc2[j]);
                      sc2[j] = fma2(qb, pr(ks + 2), sc2[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < kSMax; ++j) sc[j] = sc2[j].x + sc2[j].y;
            XF_STAMP(oi * 8 + 5);
            if (P > 1) {   // uniform over the CTA (hpw == 1 here): one exchange per attention op
#pragma unroll
              for (int j = 0; j < kSMax; ++j)
                if (j < S) xsc[(part * kSMax + j) * 128 + row] = sc[j];
              group_barrier(lane_grp);
#pragma unroll
              for (int j = 0; j < kSMax; ++j)
                if (j < S) {
                  float t = 0.f;
                  for (int pp = 0; pp < P; ++pp) t += xsc[((h * P + pp) * kSMax + j) * 128 + row];
                  sc[j] = t;
                }
            }
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < kSMax; ++j)
              if (j < S) {
                sc[j] *= scale;
                mx = fmaxf(mx, sc[j]);
              }
            float den = 0.f;
#pragma unroll
            for (int j = 0; j < kSMax; ++j)
              if (j < S) {
                sc[j] = expf(sc[j] - mx);
                den += sc[j];
              }
            const float inv = 1.f / den;
            XF_STAMP(oi * 8 + 7);
            for (int c16 = 0; c16 < ncol; c16 += 16) {
              float v16[16];
              tmem_ld16(tmem + lane_addr + (uint32_t)(2 * dp + c0 + c16), v16);
              tmem_ld_wait();
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4)
                if (c16 + 4 * e4 < ncol) {
                  // sum_j p_j (v_j + b_v) = sum_j p_j v_j + b_v (the probabilities sum to 1): the bias is added once, after 1/den
                  const float4 bv4 = *reinterpret_cast<const float4*>(bv + c0 + c16 + 4 * e4);
                  const float* v4 = v16 + 4 * e4;
                  float2 oa = mul2(make_float2(sc[0], sc[0]), pr(v4)), ob = mul2(make_float2(sc[0], sc[0]), pr(v4 + 2));
#pragma unroll
                  for (int j = 1; j < kSMax; ++j)
                    if (j < S) {
                      float vs[4];
#pragma unroll
                      for (int e = 0; e < 4; ++e) vs[e] = __shfl_sync(0xffffffffu, v4[e], rot[j]);
                      oa = fma2(make_float2(sc[j], sc[j]), pr(vs), oa);
                      ob = fma2(make_float2(sc[j], sc[j]), pr(vs + 2), ob);
                    }
                  oa = fma2(oa, make_float2(inv, inv), make_float2(bv4.x, bv4.y));
                  ob = fma2(ob, make_float2(inv, inv), make_float2(bv4.z, bv4.w));
                  const int c = c0 + c16 + 4 * e4;
                  uint2 pk = make_uint2(pack_op2<F16>(oa.x, oa.y), pack_op2<F16>(ob.x, ob.y));
                  *reinterpret_cast<uint2*>(sHop + (size_t)(c >> 3) * plane + row * 16 + (c & 7) * 2) = pk;
                }
            }
          }
        } else if (op.epi == XE_ACT_H) {
          const int ng = op.N >> 4;
#pragma unroll
          for (int i = 0; i < kSlots; ++i) {
            const int g = kParts * i + part;
            if (g < ng) {
              float acc[16], bs[16];
              tmem_ld16(tmem + lane_addr + (uint32_t)(op.tmem_col + g * 16), acc);
              lds16(pbase + g * 16, bs);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 16; q += 2) put(acc + q, add2(pr(acc + q), pr(bs + q)));
              if (!V1 && op.act == 2) {
#pragma unroll
                for (int q = 0; q < 16; ++q) acc[q] = gelu_erf(acc[q]);
                store_group<F16>(sHop, plane, row, g, acc);
              } else {
                store_group_relu<F16>(sHop, plane, row, g, acc);      // ReLU inside the 16-bit conversion
              }
            }
          }
        } else if (op.epi == XE_STREAM_ADD || op.epi == XE_STREAM_SET_PE) {
          const int ng = op.N >> 4;
          const bool set_pe = op.epi == XE_STREAM_SET_PE;
#pragma unroll
          for (int i = 0; i < kSlots; ++i) {
            const int g = kParts * i + part;
            if (g < ng) {
              float acc[16], bs[16];
              tmem_ld16(tmem + lane_addr + (uint32_t)(op.tmem_col + g * 16), acc);
              lds16(pbase + g * 16, bs);
              if (set_pe) {
                float pe16[16];
                ldg16(xf.pe + tok_s * d, g * 16, d, valid, pe16);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 16; ++q) st[i][q] = (valid && g * 16 + q < d) ? acc[q] + bs[q] + pe16[q] : 0.f;
              } else {
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 16; q += 2) put(st[i] + q, add2(pr(st[i] + q), add2(pr(acc + q), pr(bs + q))));
              }
            }
          }
        } else if (op.epi == XE_SCORE) {
          // recon = D + bias (width dt); squared error against the score target
          const int ng = op.N >> 4;
          float sq = 0.f;
#pragma unroll
          for (int i = 0; i < kSlots; ++i) {
            const int g = kParts * i + part;
            if (g < ng) {
              float acc[16], bs[16];
              const float* target = st[i];
              tmem_ld16(tmem + lane_addr + (uint32_t)(op.tmem_col + g * 16), acc);
              lds16(pbase + g * 16, bs);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const int c = g * 16 + q;
                const float r = acc[q] + bs[q];
                const float df = (valid && c < dt) ? r - target[q] : 0.f;
                sq = fmaf(df, df, sq);
                acc[q] = r;
              }
              if (recon_out && valid) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (g * 16 + 4 * e < dt)
                    *reinterpret_cast<float4*>(recon_out + ((size_t)window * S + tok_s) * dt + g * 16 + 4 * e) =
                        make_float4(acc[4 * e], acc[4 * e + 1], acc[4 * e + 2], acc[4 * e + 3]);
              }
            }
          }
          xsc[part * 128 + row] = sq;
          group_barrier(lane_grp);         // the four column parts of a row live in the four warps of its lane group
          if (part == 0) {
            const float tot = red_sum(xsc, row);
            if (reduction == SF_REDUCE_NONE) {
              if (valid && scores) scores[(size_t)window * S + tok_s] = tot / (float)dt;
            } else {
              float wsum = 0.f;
#pragma unroll
              for (int j = 0; j < kSMax; ++j)
                if (j < S) wsum += __shfl_sync(0xffffffffu, tot, (wb + j) & 31);
              if (valid && tok_s == 0 && scores) scores[window] = wsum / (float)(S * dt);
            }
          }                                // (xsc is next written in the next tile, behind that tile's barriers)
          post = XP_NONE;
        }
      }
      XF_STAMP(oi * 8 + 4);
      // ------------------------------------------------------------------ post step on the register stream
      if (post == XP_COPY_TO_AOP) {
        const int ng = dp >> 4;
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
          const int g = kParts * i + part;
          if (g < ng) store_group<F16>(sAop, plane, row, g, st[i]);
        }
      } else if (post >= XP_LN_INPLACE_TO_AOP) {
        const int ng = dp >> 4;
        // one-pass statistics (padded columns of the stream are exactly 0, so they need no masking); the four
        // warps of a lane group exchange partial sums through smem under their own named barrier, and two
        // alternating buffers make a single barrier per LayerNorm enough
        float2 s1p = make_float2(0.f, 0.f), s2p = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
          const int g = kParts * i + part;
          if (g < ng) {
#pragma unroll
            for (int q = 0; q < 16; q += 2) {
              const float2 x2 = pr(st[i] + q);
              s1p = add2(s1p, x2);
              s2p = fma2(x2, x2, s2p);
            }
          }
        }
        const float s1 = s1p.x + s1p.y, s2 = s2p.x + s2p.y;
        float* rb = red + ln_buf * (2 * kParts * 128);
        ln_buf ^= 1;
        rb[part * 128 + row] = s1;
        rb[(kParts + part) * 128 + row] = s2;
        group_barrier(lane_grp);
        const float mean = red_sum(rb, row) / (float)d;
        const float var = fmaxf(red_sum(rb + kParts * 128, row) / (float)d - mean * mean, 0.f);
        const float rstd = rsqrtf(var + kLnEps);
        const float2 mean2 = make_float2(mean, mean), rstd2 = make_float2(rstd, rstd);
        XF_STAMP(oi * 8 + 6);
        const float* ln_g = pbase ? pbase + 3 * kPW : op.ln_g;
        const float* ln_b = pbase ? pbase + 4 * kPW : op.ln_b;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < kSlots; ++i) {
          const int g = kParts * i + part;
          if (g < ng) {
            float y[16];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // generic loads: staged copy in the ring slot (GEMM ops) or the arena itself (stand-alone norms)
              // gamma / beta are zero beyond d_model (padded arena / zeroed staging block): padded columns give 0
              const float4 gm = *reinterpret_cast<const float4*>(ln_g + g * 16 + 4 * e);
              const float4 bt = *reinterpret_cast<const float4*>(ln_b + g * 16 + 4 * e);
              put(y + 4 * e, fma2(mul2(sub2(pr(st[i] + 4 * e), mean2), rstd2), make_float2(gm.x, gm.y), make_float2(bt.x, bt.y)));
              put(y + 4 * e + 2, fma2(mul2(sub2(pr(st[i] + 4 * e + 2), mean2), rstd2), make_float2(gm.z, gm.w), make_float2(bt.z, bt.w)));
            }
            if (post == XP_LN_INPLACE_TO_AOP) {
#pragma unroll
              for (int q = 0; q < 16; ++q) st[i][q] = y[q];
            }
            if (post == XP_LN_INPLACE_TO_AOP || post == XP_LN_TO_AOP) store_group<F16>(sAop, plane, row, g, y);
            if (post == XP_LN_TO_MEM || op.also_mem) store_group<F16>(sMem, plane, row, g, y);
            if (post == XP_LN_SCORE) {
              float target[16];
              ldg16(tok_row, g * 16, dt, valid, target);
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                const float df = (valid && g * 16 + q < dt) ? target[q] - y[q] : 0.f;
                sq = fmaf(df, df, sq);
              }
              if (recon_out && valid) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (g * 16 + 4 * e < dt)
                    *reinterpret_cast<float4*>(recon_out + ((size_t)window * S + tok_s) * dt + g * 16 + 4 * e) =
                        make_float4(y[4 * e], y[4 * e + 1], y[4 * e + 2], y[4 * e + 3]);
              }
            }
          }
        }
        if (post == XP_LN_SCORE) {
          xsc[part * 128 + row] = sq;
          group_barrier(lane_grp);         // the four column parts of a row live in the four warps of its lane group
          if (part == 0) {
            const float tot = red_sum(xsc, row);
            if (reduction == SF_REDUCE_NONE) {
              if (valid && scores) scores[(size_t)window * S + tok_s] = tot / (float)dt;
            } else {
              float wsum = 0.f;
#pragma unroll
              for (int j = 0; j < kSMax; ++j)
                if (j < S) wsum += __shfl_sync(0xffffffffu, tot, (wb + j) & 31);
              if (valid && tok_s == 0 && scores) scores[window] = wsum / (float)(S * dt);
            }
          }                                // (xsc is next written in the next tile, behind that tile's barriers)
        }
      }
    }
    XF_STAMP(9999);
    tc_fence_before();
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

bool make_geo(const sf_model* m, int S, XfGeo* g) {
  const XfProgram& p = m->xfprog;
  if (!p.supported || S < 1 || S > kSMax) return false;
  if ((m->xf.d_model & 3) || (m->xf.d_tok & 3)) return false;    // float4 row accesses
  g->S = S;
  g->wpw = 32 / S;
  g->rows_per_warp = g->wpw * S;
  g->win_per_tile = 4 * g->wpw;
  g->plane = 128 * 16;
  const int width = std::max(std::max(p.dp, p.dtp), 128);       // operand buffers hold activations, ctx and FFN chunks
  const uint32_t op_bytes = (uint32_t)(width / 8) * g->plane;
  g->param_off = (p.max_w_bytes + 127) & ~127;
  g->slot_bytes = g->param_off + kParamFloats * (int)sizeof(float);
  uint32_t off = 0;
  g->off_aop = off; off += op_bytes;
  g->off_hop = off; off += op_bytes;
  g->off_mem = off; off += op_bytes;
  g->off_w = off; off += 2u * (uint32_t)g->slot_bytes;
  g->off_red = off; off += (4 * kParts + kParts * kSMax) * 128 * sizeof(float);
  g->off_ops = off;
  const uint32_t ops_bytes = ((uint32_t)p.n_ops * (uint32_t)sizeof(XfOp) + 127u) & ~127u;
  g->ops_in_smem = off + ops_bytes <= (uint32_t)m->max_smem_optin ? 1 : 0;
  if (g->ops_in_smem) off += ops_bytes;
  g->smem_bytes = off;
  return off <= (uint32_t)m->max_smem_optin;
}

}  // namespace

bool transformer_bf16_supported(const sf_model* m, int S) {
  XfGeo g;
  return make_geo(m, S, &g);
}

int launch_transformer_bf16(const sf_model* m, const float* tokens, int64_t B, int S, int reduction, float* recon,
                            float* scores, cudaStream_t st, DevCount cnt) {
  if (B == 0) return SF_OK;
  XfGeo g;
  SF_REQUIRE(make_geo(m, S, &g), SF_E_UNSUPPORTED,
             "bf16 tensor-core transformer does not cover this shape (d_model=%d, heads=%d, S=%d)", m->xf.d_model, m->xf.heads, S);
  SF_REQUIRE(reduction == SF_REDUCE_MEAN || (reduction == SF_REDUCE_NONE && m->xf.variant == SF_VARIANT_SHOPFORMER_2),
             SF_E_INVALID, "reduction %d not available for variant %d", reduction, m->xf.variant);
  SF_REQUIRE(((uintptr_t)tokens & 15) == 0 && ((uintptr_t)recon & 15) == 0, SF_E_INVALID,
             "token / reconstruction buffers must be 16-byte aligned");
  SF_REQUIRE(B <= (int64_t)0x7FFF0000, SF_E_INVALID, "the tensor-core transformer takes at most 2^31 - 65,536 windows per launch");
  const int64_t n_tiles = (B + g.win_per_tile - 1) / g.win_per_tile;
  const int grid = (int)std::min<int64_t>(n_tiles, m->sm_count);
  count_launch(LK_XF_TC);
  static const int flags = getenv("SF_XF_FLAGS") ? atoi(getenv("SF_XF_FLAGS")) : 0;     // experiment switches (profiles/r2_summary.md)
  auto go = [&](auto kern) -> int {
    SF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
    kern<<<grid, kThreads, g.smem_bytes, st>>>(m->xfprog, g, m->xf, tokens, B, reduction, recon, scores, cnt, flags);
    return SF_OK;
  };
  const bool v1 = m->xf.variant == SF_VARIANT_SHOPFORMER;
  int rc;
#define SF_XF_GO(F, R) (v1 ? go(transformer_bf16_kernel<F, R, true>) : go(transformer_bf16_kernel<F, R, false>))
  if (m->xfprog.f16) rc = recon ? SF_XF_GO(true, true) : SF_XF_GO(true, false);
  else rc = recon ? SF_XF_GO(false, true) : SF_XF_GO(false, false);
#undef SF_XF_GO
  if (rc) return rc;
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}

}  // namespace sf

// debugging aid (not part of the C ABI): per-op timestamps of CTA 0, pairs of (op*8 + phase, clock64)
extern "C" int sfdbg_transformer_timing(int enable, long long* out_host, int n) {
  int on = enable;
  if (cudaMemcpyToSymbol(sf::g_xf_timing_on, &on, sizeof(int)) != cudaSuccess) return -1;
  if (out_host && n > 0) {
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(out_host, sf::g_xf_timing, sizeof(long long) * (n < 1024 ? n : 1024)) != cudaSuccess) return -1;
  }
  return 0;
}
