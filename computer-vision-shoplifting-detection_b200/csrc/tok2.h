// Tokenizer v2 ("tok2"): the tile program shared by the host builder (tok2_build.cu), the sm_100a kernel
// (tokenizer2_bf16.cu) and the host emulator of the program (tok2_emulate.cu, test infrastructure).
//
// Layout idea (reference maths: shopformer/models/gcae.py:124-154,185-195,242-259,331-366):
//   * one tile = WT = floor(128 / V) whole windows; MMA row r = w * V + v  (window, keypoint),
//     MMA column = (time, channel).  Every activation buffer is the planar-chunk layout of tc_common.cuh with
//     128 rows: byte = (col / 8) * 2048 + row * 16 + (col % 8) * 2.
//   * block 0's graph conv (2 raw coordinates -> C channels) runs on the CUDA cores in fp32 straight from the poses
//     (BatchNorm1d folded into the adjacency coefficients), one time-slice at a time into a ring of bf16 operand slots;
//   * adjacency mix of the later blocks = one 128x128 block-diagonal A operand (I_WT (x) A_hat) times the activation
//     buffer used as an MN-major B operand (K = rows);
//   * graph-conv weights = per-time-step [128 x Cin] x [Cin x Cout] MMAs;
//   * temporal conv = Toeplitz product done WITHOUT a Toeplitz matrix: for input time t the taps that are valid for
//     consecutive output times t' are consecutive blocks of a tap image stored in descending-tap order per stride phase,
//     so ONE MMA with N = (#valid t') * Cout adds time t's contribution to all its outputs; no padding taps are computed;
//   * strided 1x1 residual conv = per-output-time MMAs into the same accumulator (they also initialise it); block 0's
//     (2 input channels) is added in fp32 by the epilogue that drains the accumulator.
// The host flattens a (model, T) pair into four in-order item sequences -- G (MMA groups, one issuing thread),
// E0 / E1 (epilogue stages of two 4-warp teams) and L (TMA loads, one thread) -- and derives every cross-sequence wait
// from the items' read / write sets (shared-memory byte ranges and TMEM column ranges).  Each item owns one mbarrier
// that completes exactly once per tile.
#pragma once
#include <stdint.h>

namespace sf {
namespace t2 {

constexpr int kRows = 128;               // MMA M
constexpr uint32_t kPlane = 2048;        // bytes of one 8-column chunk of a 128-row activation buffer
constexpr int kTeams = 2;                // epilogue teams: warps 4-11 and 12-19; a team = 4 TMEM lane quarters x 2 column halves
constexpr int kTeamWarps = 8;
// warp 0: MMA issue, warp 1: TMA, then the teams from warp kFirstEpiWarp on.  A warp reads the TMEM lane quarter warp % 4, and a
// team needs two warps per quarter: any even first warp works.  Starting at warp 2 (no idle warps, 576 threads) gives every
// thread 112 registers instead of 96 at 640 threads (SF_TOK2_FIRST_EPI_WARP=4 restores the older layout).
#ifndef SF_TOK2_FIRST_EPI_WARP
#define SF_TOK2_FIRST_EPI_WARP 2
#endif
constexpr int kFirstEpiWarp = SF_TOK2_FIRST_EPI_WARP;
static_assert(kFirstEpiWarp >= 2 && kFirstEpiWarp % 2 == 0, "teams start at an even warp behind the MMA and TMA warps");
constexpr int kThreads = kFirstEpiWarp * 32 + kTeams * kTeamWarps * 32;
constexpr int kMaxGroups = 80, kMaxStages = 72, kMaxLoads = 10, kMaxMma = 384;
constexpr int kMaxC0 = 64;               // block-0 output channels handled by the CUDA-core graph conv

struct Mma {                 // one tcgen05.mma (M = 128, K = 16)
  uint32_t a_lo;             // descriptor low word relative to the dynamic smem base: (offset >> 4) | (LBO >> 4) << 16
  uint32_t b_lo;
  uint32_t d;                // [0,9) TMEM column | bit 16 accumulate | bit 17 B is MN-major (SBO = plane instead of 128 B)
  uint32_t idesc;
};

struct Group {               // G item: a run of MMAs followed by one commit
  uint16_t first, count;
  int16_t wait_e[kTeams];    // stage of team 0 / 1 that must have completed this tile (-1: none)
  int16_t wait_l;            // L load
  int16_t prev_team, prev_stage;   // stage of the PREVIOUS tile that must have completed (-1: none): the token store
  int16_t pad;
};

enum StageType { ST_G0 = 0, ST_CVT = 1, ST_XEPI0 = 2, ST_TOKENS = 3 };
enum StageFlags { SF_RELU = 1, SF_BIAS = 2, SF_TEAM_SYNC = 4 };

struct Stage {               // E item (index = position in its TEAM's sequence); 32-bit fields: read with uniform constant loads
  int32_t type, flags;
  int32_t wait_g, wait_l, wait_eo, wait_g_prev;   // item indices (-1: none): G group / L load / other team's stage (this tile), G group (previous tile)
  // the same waits as shared-memory byte offsets of the mbarriers (0: none), and this stage's own barrier
  uint32_t bar_g, bar_l, bar_eo, bar_g_prev, bar_self;
  int32_t tmem_col, n_cg;    // CVT / XEPI0 / TOKENS: first accumulator column, number of 16-column groups
  uint32_t dst_off;          // smem byte offset of destination column 0 (G0: the ring slot)
  uint32_t bias_off;         // byte offset of the fp32 bias vector (period `bias_period` columns)
  int32_t bias_period;
  int32_t p0, p1;            // G0: input time steps [p0, p1); XEPI0: output time steps [p0, p1)
  int32_t pad[2];            // 80 bytes: the kernel loads a stage as five 16-byte words
};
static_assert(sizeof(Stage) == 80, "Stage is loaded as five 16-byte words");

// Kernel form of a Stage: 32 bytes = two 16-byte loads per stage instead of five, and 6 live registers instead of 14 (the
// epilogue loop runs at the 96-register limit of 18 resident warps).  Barrier offsets are stored / 8 (0: none).
struct StageK {
  uint32_t w0;               // type [0,2) | flags [2,5) | n_cg [5,13) | tmem_col [13,22) | bias_period [22,30)
  uint32_t w1;               // bar_g | bar_l << 16
  uint32_t w2;               // bar_eo | bar_g_prev << 16
  uint32_t w3;               // bar_self | p0 << 16 | (p1 - p0) << 24
  uint32_t dst_off, bias_off;
  uint32_t pad[2];
#ifdef __CUDACC__
#define SF_HD __host__ __device__ __forceinline__
#else
#define SF_HD inline
#endif
  SF_HD int type() const { return (int)(w0 & 3u); }
  SF_HD int flags() const { return (int)((w0 >> 2) & 7u); }
  SF_HD int n_cg() const { return (int)((w0 >> 5) & 0xFFu); }
  SF_HD int tmem_col() const { return (int)((w0 >> 13) & 0x1FFu); }
  SF_HD int bias_period() const { return (int)((w0 >> 22) & 0xFFu); }
  SF_HD uint32_t bar_g() const { return (w1 & 0xFFFFu) << 3; }
  SF_HD uint32_t bar_l() const { return (w1 >> 16) << 3; }
  SF_HD uint32_t bar_eo() const { return (w2 & 0xFFFFu) << 3; }
  SF_HD uint32_t bar_g_prev() const { return (w2 >> 16) << 3; }
  SF_HD uint32_t bar_self() const { return (w3 & 0xFFFFu) << 3; }
  SF_HD int p0() const { return (int)((w3 >> 16) & 0xFFu); }
  SF_HD int p1() const { return (int)((w3 >> 16) & 0xFFu) + (int)(w3 >> 24); }
#undef SF_HD
};
static_assert(sizeof(StageK) == 32, "StageK is loaded as two 16-byte words");
// false when a field does not fit its bit range
inline bool pack_stage(const Stage& s, StageK* k) {
  auto bar = [](uint32_t off) { return off >> 3; };
  const bool ok = s.type >= 0 && s.type < 4 && s.flags >= 0 && s.flags < 8 && s.n_cg >= 0 && s.n_cg < 256 && s.tmem_col >= 0 && s.tmem_col < 512 &&
                  s.bias_period >= 0 && s.bias_period < 256 && s.p0 >= 0 && s.p0 < 256 && s.p1 >= s.p0 && s.p1 - s.p0 < 256 &&
                  !((s.bar_g | s.bar_l | s.bar_eo | s.bar_g_prev | s.bar_self) & 7u) &&
                  bar(s.bar_g | s.bar_l | s.bar_eo | s.bar_g_prev | s.bar_self) < 65536u;
  k->w0 = (uint32_t)s.type | (uint32_t)s.flags << 2 | (uint32_t)s.n_cg << 5 | (uint32_t)s.tmem_col << 13 | (uint32_t)s.bias_period << 22;
  k->w1 = bar(s.bar_g) | bar(s.bar_l) << 16;
  k->w2 = bar(s.bar_eo) | bar(s.bar_g_prev) << 16;
  k->w3 = bar(s.bar_self) | (uint32_t)s.p0 << 16 | (uint32_t)(s.p1 - s.p0) << 24;
  k->dst_off = s.dst_off;
  k->bias_off = s.bias_off;
  k->pad[0] = k->pad[1] = 0;
  return ok;
}

enum LoadKind { LD_WEIGHTS = 0, LD_POSES = 1 };
struct Load {                // L item
  uint8_t kind, pad0;
  int16_t wait_g;            // this tile
  int16_t wait_e[kTeams];
  int16_t wait_g_prev;       // previous tile
  int16_t pad1;
  uint32_t dst_off, bytes;
  uint32_t src_off, pad2;    // LD_WEIGHTS: byte offset of the image inside the model's blob (Plan::const_src)
};

struct Plan {                // kernel parameter (by value)
  int V, WT, rows, c_in, T0, S_out, c_last, cp_last, d_tok;
  int T_last, pool;          // time steps of the last block's output; > 0: adaptive average pooling of them to `pool` = S_out tokens
                             // (shopformer_2/models/gcae.py:406-415)
  int per_w;                 // floats per pose window
  int n_groups, n_loads, n_mma;
  int n_stages[kTeams];
  const unsigned char* const_src;   // resident images + fp32 tables, copied to smem once per CTA
  uint32_t const_bytes;
  // shared-memory map (byte offsets from the dynamic smem base)
  uint32_t off_const, off_P, off_Q, off_W, off_stage_tok, off_bars, off_flags;
  uint32_t off_xin;
  uint32_t off_ell, off_hc, off_scale, off_shift;   // const blob: mix coefficients float4 (A_hat*scale_x, A_hat*scale_y, row delta, 0) [5|8][V],
                                                   // float2 [V] mixed BN shifts, BN1d scale / shift [c_in][V]
  int ell_width;
  int cp0, stride0;
  int bar_g0, bar_l0, n_bars;                   // barrier index bases
  int bar_e0[kTeams];
  uint32_t smem_bytes;
  // block 0 on the CUDA cores: fp32 tables in the const blob, per 4 output channels (w_x[4], w_y[4], bias[4]):
  // graph-conv weight / bias, and the BN-folded residual 1x1 conv weight / output bias
  uint32_t off_g0tab, off_r0tab;
  uint32_t off_g0tab_h;      // fp16 operand format: graph-conv table in halves, per 8 output channels (w_x[8], w_y[8], bias[8])
  // The tile program itself travels in the kernel's parameter space (constant bank): the MMA-issuing thread reads
  // descriptors with uniform-datapath constant loads, no shared-memory round trip and no register -> uniform moves.
  Group groups[kMaxGroups];
  Stage stages[kTeams][kMaxStages];
  Load loads[kMaxLoads];
  Mma mma[kMaxMma];
};
static_assert(sizeof(Plan) < 20000, "Plan travels as a kernel parameter");

}  // namespace t2
}  // namespace sf
