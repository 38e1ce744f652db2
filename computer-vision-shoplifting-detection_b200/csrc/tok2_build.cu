// Tokenizer v2, host side: operand images (once per model) and the tile program (once per window length T).
// See tok2.h for the layout.  Reference maths: shopformer/models/gcae.py:124-154 (graph conv), :185-195 (temporal
// conv + BN), :242-259 (block), :331-366 (encoder); shopformer_2/models/gcae.py:375-422.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "tok2_build.h"

namespace sf {
namespace t2 {
namespace {

inline uint16_t f2bf(float f) {       // round-to-nearest-even fp32 -> bf16
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline uint16_t f2h(float f) {        // round-to-nearest-even fp32 -> fp16 (finite inputs; overflow saturates to inf)
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
  const int e = (int)((u >> 23) & 0xFF) - 127 + 15;
  uint32_t m = u & 0x7FFFFFu;
  if (e >= 31) return (uint16_t)(sign | 0x7C00u);
  if (e <= 0) {                                    // subnormal half (or zero)
    if (e < -10) return sign;
    m |= 0x800000u;
    const int sh = 14 - e;                         // 24-bit significand -> 10 bits at exponent 1
    uint32_t h = m >> sh;
    const uint32_t rem = m & ((1u << sh) - 1), half = 1u << (sh - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)e << 10) | (m >> 13);
  const uint32_t rem = m & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;     // a carry into the exponent is the right result
  return (uint16_t)(sign | h);
}
inline int pad16(int x) { return (x + 15) & ~15; }
inline uint32_t up128(size_t x) { return (uint32_t)((x + 127) & ~size_t(127)); }

struct Blob {
  std::vector<unsigned char>& b;
  uint32_t alloc(size_t bytes) {
    const uint32_t off = up128(b.size());
    b.resize(off + bytes, 0);
    return off;
  }
  uint16_t* bf(uint32_t off) { return reinterpret_cast<uint16_t*>(b.data() + off); }
  float* f32(uint32_t off) { return reinterpret_cast<float*>(b.data() + off); }
};

// K-major image [K/8][nrows][8] of val(n, k), bf16 (or fp16 when `half`)
template <class F>
void fill_kmajor(uint16_t* dst, int nrows, int K, F val, bool half = false) {
  for (int kc = 0; kc < K / 8; ++kc)
    for (int n = 0; n < nrows; ++n)
      for (int e = 0; e < 8; ++e) {
        const float x = val(n, kc * 8 + e);
        dst[((size_t)kc * nrows + n) * 8 + e] = half ? f2h(x) : f2bf(x);
      }
}

}  // namespace

void build_static(const Tokenizer& tok, int pool_tokens, bool allow_f16, Static* out) {
  Static& s = *out;
  s = Static();
  const int V = tok.V, nb = tok.n_blocks;
  auto fail = [&](const char* w) { s.ok = false; s.why = w; };
  s.pool_tokens = pool_tokens > 0 ? pool_tokens : 0;
  if (tok.c_in > 2) return fail("more than 2 input channels");
  if (nb < 2) return fail("fewer than 2 blocks");
  if (V > 64 || V < 2) return fail("keypoint count outside [2,64]");
  s.V = V;
  s.WT = kRows / V;
  s.rows = s.WT * V;
  s.c_in = tok.c_in;
  s.n_blocks = nb;
  Blob bl{s.blob};

  // dense adjacency per block from the ELL rows; blocks >= 1 must share one adjacency (one block-diagonal mix operand)
  std::vector<std::vector<float>> adj(nb, std::vector<float>((size_t)V * V, 0.f));
  for (int b = 0; b < nb; ++b) {
    const TokBlock& tb = tok.blk[b];
    for (int v = 0; v < V; ++v)
      for (int k = 0; k < tb.ell_width; ++k) adj[b][(size_t)v * V + tb.ell_col[v * tb.ell_width + k]] += tb.ell_val[v * tb.ell_width + k];
  }
  for (int b = 2; b < nb; ++b)
    if (adj[b] != adj[1]) return fail("blocks use different adjacency matrices");
  if (tok.blk[0].ell_width > 8) return fail("adjacency rows with more than 8 non-zeros");
  if (tok.blk[0].identity_res) return fail("identity residual in block 0");

  // Operand format: fp16 (11 significant bits instead of bf16's 8 -- a weight's rounding error is a fixed perturbation
  // of the model that does not average out, and the normalised adjacency is a handful of irrational values every
  // activation passes through once per block) unless the caller wants bf16 or a folded weight leaves fp16's range.
  // tcgen05 kind::f16 takes A and B in the SAME format (a mixed descriptor is an illegal instruction), so the activations
  // the epilogues write are fp16 too, saturating at +-65504 (tc_common.cuh pack_f16x2).
  s.f16 = allow_f16;
  for (int b = 0; b < nb && s.f16; ++b) {
    const TokBlock& tb = tok.blk[b];
    auto fits = [&](const float* w, size_t n) {
      for (size_t i = 0; i < n; ++i)
        if (!(std::fabs(w[i]) < 32768.f)) return false;
      return true;
    };
    s.f16 = fits(tb.tcn_w, (size_t)tb.cout * kTaps * tb.cout);
    if (b > 0 && s.f16) s.f16 = fits(tb.gcn_w, (size_t)tb.cin * tb.cout) && (tb.identity_res || fits(tb.res_w, (size_t)tb.cin * tb.cout));
  }

  // ---- const part ------------------------------------------------------------------------------------------
  // block-diagonal mix operand (A, K-major): (m, k) = A_hat[v_m][u_k] iff same window
  s.off_ablk = bl.alloc((size_t)kRows * kRows * 2);
  fill_kmajor(bl.bf(s.off_ablk), kRows, kRows, [&](int m, int k) -> float {
    if (m >= s.rows || k >= s.rows || m / V != k / V) return 0.f;
    return adj[1][(size_t)(m % V) * V + (k % V)];
  }, s.f16);
  for (int b = 0; b < nb; ++b) {
    const TokBlock& tb = tok.blk[b];
    BlockStatic& o = s.blk[b];
    o.cin = tb.cin;
    o.cout = tb.cout;
    o.cin_p = b == 0 ? tb.cin : pad16(tb.cin);
    o.cp = pad16(tb.cout);
    o.stride = tb.stride;
    o.identity = tb.identity_res;
    if (o.cp > 128) return fail("more than 128 channels");
    if (b > 0 && o.cin_p != s.blk[b - 1].cp) return fail("channel padding mismatch");
  }
  {
    // block 0 runs on the CUDA cores from fp32 tables (they travel in the kernel parameters); the tensor cores only
    // need a zero A operand (two 8-column chunks) to clear the block-0 accumulator
    const TokBlock& t0 = tok.blk[0];
    if (s.blk[0].cp > kMaxC0) return fail("block 0 wider than 64 channels");
    // the all-zero A operand that clears block 0's accumulator: the mix image's last K chunk (columns 120..127) is zero in
    // every row whenever the tile leaves at least 8 padding rows; both K halves of the MMA then read that one plane (LBO = 0)
    if (kRows - s.rows >= 8) {
      s.off_zero = s.off_ablk + (uint32_t)(kRows / 8 - 1) * kPlane;
      s.zero_lbo = 0;
    } else {
      s.off_zero = bl.alloc((size_t)2 * kPlane);
      s.zero_lbo = kPlane;
    }
    const int cp = s.blk[0].cp;
    auto table = [&](const float* w, const float* bias) {
      const uint32_t off = bl.alloc((size_t)cp * 3 * 4);
      float* t = bl.f32(off);
      for (int o = 0; o < cp; ++o) {
        float* e = t + (o / 4) * 12 + (o % 4);
        e[0] = (o < t0.cout && t0.cin > 0) ? w[0 * t0.cout + o] : 0.f;
        e[4] = (o < t0.cout && t0.cin > 1) ? w[1 * t0.cout + o] : 0.f;
        e[8] = o < t0.cout ? bias[o] : 0.f;
      }
      return off;
    };
    s.off_g0tab = table(t0.gcn_w, t0.gcn_b);
    if (s.f16) {
      // fp16 format: the graph conv of block 0 runs in packed half arithmetic (HFMA2 with fused ReLU writes the fp16
      // operand directly: no fp32 -> fp16 conversions, half the table loads); its inputs (|m| of a few units, BN-folded
      // weights) are far inside fp16's range
      for (int i = 0; i < t0.cin * t0.cout; ++i)
        if (!(std::fabs(t0.gcn_w[i]) < 32768.f)) return fail("block-0 graph-conv weight outside fp16's range");
      s.off_g0tab_h = bl.alloc((size_t)cp * 3 * 2);
      uint16_t* t = bl.bf(s.off_g0tab_h);
      for (int o = 0; o < cp; ++o) {
        uint16_t* e = t + (o / 8) * 24 + (o % 8);
        e[0] = f2h((o < t0.cout && t0.cin > 0) ? t0.gcn_w[0 * t0.cout + o] : 0.f);
        e[8] = f2h((o < t0.cout && t0.cin > 1) ? t0.gcn_w[1 * t0.cout + o] : 0.f);
        e[16] = f2h(o < t0.cout ? t0.gcn_b[o] : 0.f);
      }
    }
    s.off_r0tab = table(t0.res_w, t0.out_b);
  }
  for (int b = 1; b < nb; ++b) {
    const TokBlock& tb = tok.blk[b];
    BlockStatic& o = s.blk[b];
    o.off_gcn = bl.alloc((size_t)o.cin_p * o.cp * 2);
    fill_kmajor(bl.bf(o.off_gcn), o.cp, o.cin_p, [&](int n, int k) -> float { return (n < tb.cout && k < tb.cin) ? tb.gcn_w[k * tb.cout + n] : 0.f; }, s.f16);
    o.off_res = bl.alloc((size_t)o.cin_p * o.cp * 2);
    fill_kmajor(bl.bf(o.off_res), o.cp, o.cin_p, [&](int n, int k) -> float {
      if (n >= tb.cout || k >= tb.cin) return 0.f;
      return tb.identity_res ? (n == k ? 1.f : 0.f) : tb.res_w[k * tb.cout + n];
    }, s.f16);
    o.off_bias_g = bl.alloc((size_t)o.cp * 4);
    o.off_bias_o = bl.alloc((size_t)o.cp * 4);
    for (int n = 0; n < tb.cout; ++n) {
      bl.f32(o.off_bias_g)[n] = tb.gcn_b[n];
      bl.f32(o.off_bias_o)[n] = tb.out_b[n];
    }
  }
  {
    // prep tables: block-0 adjacency row of keypoint v with the BatchNorm1d scale folded in -- float4 [k][V] =
    // (A_hat[v][u] * scale_x[u], A_hat[v][u] * scale_y[u], row delta u - v, 0) -- the mixed shifts float2 [V] =
    // sum_u A_hat[v][u] * shift_c[u], and the plain scale / shift for the un-mixed residual operand
    const TokBlock& t0 = tok.blk[0];
    s.ell_width = t0.ell_width <= 5 ? 5 : 8;
    s.off_ell = bl.alloc((size_t)s.ell_width * V * 16);
    s.off_hc = bl.alloc((size_t)V * 8);
    for (int v = 0; v < V; ++v) {
      float h[2] = {0.f, 0.f};
      for (int k = 0; k < s.ell_width; ++k) {
        float val = 0.f;
        int dl = 0;
        if (k < t0.ell_width) {
          val = t0.ell_val[v * t0.ell_width + k];
          dl = val != 0.f ? t0.ell_col[v * t0.ell_width + k] - v : 0;
        }
        float* e = bl.f32(s.off_ell) + (size_t)(k * V + v) * 4;
        for (int c = 0; c < 2; ++c) {
          e[c] = c < tok.c_in ? val * tok.in_scale[c * V + v + dl] : 0.f;
          if (c < tok.c_in) h[c] = std::fmaf(val, tok.in_shift[c * V + v + dl], h[c]);
        }
        memcpy(&e[2], &dl, 4);
        e[3] = 0.f;
      }
      bl.f32(s.off_hc)[v * 2] = h[0];
      bl.f32(s.off_hc)[v * 2 + 1] = h[1];
    }
    s.off_scale = bl.alloc((size_t)tok.c_in * V * 4);
    s.off_shift = bl.alloc((size_t)tok.c_in * V * 4);
    memcpy(bl.f32(s.off_scale), tok.in_scale, sizeof(float) * tok.c_in * V);
    memcpy(bl.f32(s.off_shift), tok.in_shift, sizeof(float) * tok.c_in * V);
  }
  s.const_bytes = up128(s.blob.size());
  s.blob.resize(s.const_bytes, 0);

  // ---- temporal-conv images: per stride phase p the taps k = p (mod s) in DESCENDING order, so that for one input
  // time the taps of consecutive output times are consecutive N blocks
  for (int b = 0; b < nb; ++b) {
    const TokBlock& tb = tok.blk[b];
    BlockStatic& o = s.blk[b];
    int pos = 0;
    for (int p = 0; p < tb.stride && p < kTaps; ++p) {
      int kmax = p;
      while (kmax + tb.stride < kTaps) kmax += tb.stride;
      for (int k = kmax; k >= 0; k -= tb.stride) o.tap_pos[k] = pos++;
    }
    std::vector<int> tap_at(kTaps);
    for (int k = 0; k < kTaps; ++k) tap_at[o.tap_pos[k]] = k;
    o.tcn_bytes = (uint32_t)(kTaps * o.cp * o.cp * 2);
    o.off_tcn = bl.alloc(o.tcn_bytes);
    const int co = tb.cout;
    fill_kmajor(bl.bf(o.off_tcn), kTaps * o.cp, o.cp, [&](int n, int c) -> float {
      const int k = tap_at[n / o.cp], oc = n % o.cp;
      return (c < co && oc < co) ? tb.tcn_w[((size_t)c * kTaps + k) * co + oc] : 0.f;
    }, s.f16);
  }
  s.blob.resize(up128(s.blob.size()), 0);
  s.ok = true;
}

// ---------------------------------------------------------------------------------------------------------------
namespace {

// SF_TOK2_SPLIT=1 splits every conversion / block-0 stage between the two epilogue teams instead of alternating whole
// stages (measured slower: 2.31 vs 1.91 ms -- the epilogue is bound by its instruction stream, not by per-stage latency,
// and both teams then repeat the pose gather of block 0)
const bool kSplitStages = getenv("SF_TOK2_SPLIT") != nullptr;

struct Rng {
  int space;        // 0 = shared-memory bytes, 1 = TMEM columns
  uint32_t lo, hi;
};
enum Side { SIDE_G = 0, SIDE_L = 1, SIDE_E0 = 2, SIDE_E1 = 3, N_SIDES = 4 };
struct Item {
  int side, idx;
  std::vector<Rng> rd, wr;
};
bool overlap(const std::vector<Rng>& a, const std::vector<Rng>& b) {
  for (const Rng& x : a)
    for (const Rng& y : b)
      if (x.space == y.space && x.lo < y.hi && y.lo < x.hi) return true;
  return false;
}

}  // namespace

void build_program(const Static& st, int T, int max_smem, Program* out) {
  Program& pr = *out;
  pr = Program();
  auto fail = [&](const std::string& w) { pr.ok = false; pr.why = w; };
  if (!st.ok) return fail(st.why);
  const int nb = st.n_blocks, V = st.V;
  if (T < 1 || T > 256) return fail("window length outside [1,256]");
  int Tin[kMaxBlocks], Tout[kMaxBlocks];
  {
    int t = T;
    for (int b = 0; b < nb; ++b) {
      Tin[b] = t;
      Tout[b] = (t - 1) / st.blk[b].stride + 1;
      t = Tout[b];
    }
  }
  const int per_w = st.c_in * T * V;
  if ((per_w * 4) % 16) return fail("window size not a multiple of 16 bytes");
  const uint32_t xin_bytes = (uint32_t)st.WT * per_w * 4, xin_alloc = up128(xin_bytes);
  const int cp0 = st.blk[0].cp;
  if (Tout[0] * cp0 > 512) return fail("block-0 accumulators exceed tensor memory");
  if (cp0 != 16 && cp0 != 32 && cp0 != 64) return fail("block-0 width not 16 / 32 / 64");

  // ---- block 0: time steps per CUDA-core slice and the ring of operand slots in Q
  const uint32_t ring_cap_q = 65536;
  int st0 = 2;                                          // time steps per slice
  while (st0 > 1 && (uint32_t)(st0 * cp0 / 8) * kPlane * 2 > ring_cap_q) --st0;
  const uint32_t ring0_slot = (uint32_t)(st0 * cp0 / 8) * kPlane;
  const int n_sl = (T + st0 - 1) / st0;
  const int ring0_slots = std::max(1, std::min(std::min(4, n_sl), (int)(ring_cap_q / ring0_slot)));

  // ---- chunking of blocks >= 1 (time steps per mix / graph-conv chunk) and region sizes
  int ct[kMaxBlocks] = {0}, nch[kMaxBlocks] = {0};
  uint32_t slot_bytes[kMaxBlocks] = {0};
  uint32_t x_bytes[kMaxBlocks + 1] = {0};          // x_b = input of block b (b >= 1), planar-chunk bf16
  for (int b = 1; b <= nb; ++b) x_bytes[b] = b < nb ? (uint32_t)(Tin[b] * st.blk[b].cin_p / 8) * kPlane : 0u;
  uint32_t P_need = xin_alloc, Q_need = ring0_slot * ring0_slots;
  for (int b = 1; b < nb; ++b) {
    const BlockStatic& k = st.blk[b];
    const bool x_in_P = (b & 1) != 0;                 // x1 in P, x2 in Q, ...
    (x_in_P ? P_need : Q_need) = std::max(x_in_P ? P_need : Q_need, x_bytes[b]);
    const int accw = Tout[b] * k.cp;
    // at least two chunks when the block is long enough: the second chunk's mix / graph conv run under the first
    // chunk's epilogues
    int c_max = Tin[b] >= 4 ? (Tin[b] + 1) / 2 : Tin[b];
    if (b == 1) c_max = std::min(c_max, 8);        // block 0's output epilogue keeps one chunk's poses in registers
    int best = 0;
    for (int c = c_max; c >= 1; --c) {
      const int chunks = (Tin[b] + c - 1) / c;
      const uint32_t sb = (uint32_t)(c * std::max(k.cin_p, k.cp) / 8) * kPlane;
      const uint32_t ring = sb * (uint32_t)std::min(2, chunks);
      if (c * k.cin_p > 256 || c * k.cp > 256) continue;
      if (accw + c * k.cin_p + c * k.cp > 512) continue;
      const uint32_t cap = x_in_P ? ring_cap_q : std::max(P_need, (uint32_t)98304);
      if (ring > cap) continue;
      best = c;
      break;
    }
    if (!best) return fail("no chunking fits tensor memory / shared memory");
    ct[b] = best;
    nch[b] = (Tin[b] + best - 1) / best;
    slot_bytes[b] = (uint32_t)(best * std::max(k.cin_p, k.cp) / 8) * kPlane;
    const uint32_t ring = slot_bytes[b] * (uint32_t)std::min(2, nch[b]);
    (x_in_P ? Q_need : P_need) = std::max(x_in_P ? Q_need : P_need, ring);
  }
  const int T_last = Tout[nb - 1], S_out = st.pool_tokens > 0 ? st.pool_tokens : T_last;
  const int c_last = st.blk[nb - 1].cout, d_tok = c_last * V;
  const uint32_t stage_bytes = up128((size_t)st.WT * S_out * d_tok * 4);
  const uint32_t P_size = up128(P_need), Q_size = up128(Q_need);
  uint32_t W_size = 0;
  for (int b = 0; b < nb; ++b) W_size = std::max(W_size, st.blk[b].tcn_bytes);

  Plan& pl = pr.plan;
  memset(&pl, 0, sizeof(pl));
  pl.V = V;
  pl.WT = st.WT;
  pl.rows = st.rows;
  pl.c_in = st.c_in;
  pl.T0 = T;
  pl.S_out = S_out;
  pl.T_last = T_last;
  pl.pool = st.pool_tokens;
  pl.c_last = c_last;
  pl.cp_last = st.blk[nb - 1].cp;
  pl.d_tok = d_tok;
  pl.per_w = per_w;
  pl.const_bytes = st.const_bytes;
  pl.ell_width = st.ell_width;
  pl.cp0 = cp0;
  pl.stride0 = st.blk[0].stride;
  uint32_t off = 0;
  pl.off_const = off; off += st.const_bytes;
  pl.off_P = off; off += P_size;
  pl.off_Q = off; off += Q_size;
  pl.off_W = off; off += up128(W_size);
  // The token staging area lives at the START of P (x1 / x3 / block-2 ring): by the token stage every MMA that reads P has
  // completed (the stage drains the last accumulator), the next tile's poses sit at the END of P, and the next writers of
  // P's head (block 0's output stages) come after the next tile's first MMA group, which waits for this stage.  The stage
  // only completes once the bulk store has READ the staging area (cp.async.bulk.wait_group.read in the storing thread).
  pl.off_stage_tok = pl.off_P;
  if (stage_bytes + xin_alloc > P_size) return fail("token staging and the pose slot do not both fit the first activation buffer");
  pl.off_ell = pl.off_const + st.off_ell;
  pl.off_hc = pl.off_const + st.off_hc;
  pl.off_scale = pl.off_const + st.off_scale;
  pl.off_shift = pl.off_const + st.off_shift;
  pl.off_g0tab = pl.off_const + st.off_g0tab;
  pl.off_g0tab_h = pl.off_const + st.off_g0tab_h;
  pl.off_r0tab = pl.off_const + st.off_r0tab;
  // the raw poses sit at the END of P: x1 (block 0's output, written last) only reaches them with its final columns
  pl.off_xin = pl.off_P + P_size - xin_alloc;

  // ---- emit the items ----------------------------------------------------------------------------------------
  std::vector<Item> order;          // one valid sequential schedule of a tile
  auto smem_r = [](uint32_t lo, uint32_t bytes) { return Rng{0, lo, lo + bytes}; };
  auto tmem_r = [](int col, int n) { return Rng{1, (uint32_t)col, (uint32_t)(col + n)}; };
  const uint32_t c_lo = pl.off_const, c_hi = pl.off_const + st.const_bytes;
  auto is_const = [&](uint32_t o) { return o >= c_lo && o < c_hi; };
  const std::vector<Rng> xin_rng{smem_r(pl.off_xin, xin_alloc)};

  auto new_group = [&]() {
    Group g;
    memset(&g, 0, sizeof(g));
    g.first = (uint16_t)pr.mma.size();
    g.wait_e[0] = g.wait_e[1] = g.wait_l = -1;
    g.prev_team = g.prev_stage = -1;
    pr.groups.push_back(g);
    order.push_back(Item{SIDE_G, (int)pr.groups.size() - 1, {}, {}});
  };
  // one MMA: A K-major at a_off (LBO = a_lbo), B at b_off: K-major (LBO = b_lbo, rows = N) or MN-major (activation
  // buffer used with K = rows: K step kk at +256 B, N chunks one plane apart)
  auto add_mma = [&](uint32_t a_off, uint32_t a_lbo, uint32_t b_off, uint32_t b_lbo, bool b_mn, int N, int dcol, bool acc) {
    const bool a_f16 = st.f16, b_f16 = st.f16;          // instruction descriptor formats: 0 = F16, 1 = BF16
    Mma m;
    m.a_lo = ((a_off >> 4) & 0x3FFFu) | ((a_lbo >> 4) << 16);
    m.b_lo = ((b_off >> 4) & 0x3FFFu) | (((b_mn ? 128u : b_lbo) >> 4) << 16);
    m.d = (uint32_t)dcol | (acc ? 1u << 16 : 0u) | (b_mn ? 1u << 17 : 0u);
    m.idesc = (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);
    pr.mma.push_back(m);
    pr.groups.back().count++;
    Item& it = order.back();
    if (!is_const(a_off)) {
      it.rd.push_back(smem_r(a_off, kPlane));
      it.rd.push_back(smem_r(a_off + a_lbo, kPlane));
    }
    if (!is_const(b_off)) {
      if (b_mn) it.rd.push_back(smem_r(b_off & ~(kPlane - 1), (uint32_t)(N / 8) * kPlane));
      else {
        it.rd.push_back(smem_r(b_off, (uint32_t)N * 16));
        it.rd.push_back(smem_r(b_off + b_lbo, (uint32_t)N * 16));
      }
    }
    it.wr.push_back(tmem_r(dcol, N));
  };
  int next_team = 0;                 // unsplit stages alternate between the two epilogue teams
  auto new_stage = [&](int type, int flags, int on_team = -1) -> Stage& {
    const int team = on_team >= 0 ? on_team : next_team;
    if (on_team < 0) next_team ^= 1;
    Stage s;
    memset(&s, 0, sizeof(s));
    s.type = type;
    s.flags = flags;
    s.wait_g = s.wait_l = s.wait_eo = s.wait_g_prev = -1;
    pr.stages[team].push_back(s);
    order.push_back(Item{SIDE_E0 + team, (int)pr.stages[team].size() - 1, {}, {}});
    return pr.stages[team].back();
  };
  // A conversion stage sits on the tile's critical path (MMA group -> stage -> MMA group), so its columns are split between
  // the two teams, which run their halves at the same time; the consuming group waits for both.  The split point is a
  // multiple of the bias period (a stage indexes its bias from its own first column).
  auto cvt_stage = [&](int tmem_col, int ncols, uint32_t dst_off, int flags, uint32_t bias_off, int period) {
    const int n_cg = ncols / 16;
    const int unit = (flags & SF_BIAS) && period > 16 ? period / 16 : 1;
    int n0 = (n_cg + 1) / 2;
    n0 = (n0 + unit - 1) / unit * unit;
    auto emit = [&](int cg0, int cgs, int team) {
      Stage& s = new_stage(ST_CVT, flags, team);
      s.tmem_col = tmem_col + cg0 * 16;
      s.n_cg = cgs;
      s.dst_off = dst_off + (uint32_t)(cg0 * 2) * kPlane;
      s.bias_off = bias_off;
      s.bias_period = period;
      order.back().rd.push_back(tmem_r(tmem_col + cg0 * 16, cgs * 16));
      order.back().wr.push_back(smem_r(dst_off + (uint32_t)(cg0 * 2) * kPlane, (uint32_t)(cgs * 2) * kPlane));
    };
    if (!kSplitStages || n0 >= n_cg) return emit(0, n_cg, -1);
    emit(0, n0, 0);
    emit(n0, n_cg - n0, 1);
  };
  auto new_load = [&](int kind, uint32_t dst_off, uint32_t bytes, uint32_t src) {
    Load l;
    memset(&l, 0, sizeof(l));
    l.kind = (uint8_t)kind;
    l.wait_g = l.wait_e[0] = l.wait_e[1] = l.wait_g_prev = -1;
    l.dst_off = dst_off;
    l.bytes = bytes;
    l.src_off = src;
    pr.loads.push_back(l);
    order.push_back(Item{SIDE_L, (int)pr.loads.size() - 1, {}, {smem_r(dst_off, bytes)}});
  };
  // temporal-conv MMAs of input time steps [t0, t0 + nt) whose bf16 activations sit in `slot` (column = tl * cp + c)
  auto tcn_mmas = [&](const BlockStatic& k, int b, uint32_t slot, int t0, int nt, int acc) {
    const int cp = k.cp, s = k.stride;
    const uint32_t w_plane = (uint32_t)kTaps * cp * 16;
    for (int tl = 0; tl < nt; ++tl) {
      const int t = t0 + tl;
      int lo = (t - kHalo + s - 1) / s;
      if (t - kHalo < 0) lo = 0;
      const int hi = std::min(Tout[b] - 1, (t + kHalo) / s);
      if (lo > hi) continue;
      const int k_lo = t - s * lo + kHalo;
      for (int n0 = lo; n0 <= hi;) {
        const int cnt = std::min(hi - n0 + 1, 256 / cp);
        const int ktap = k_lo - s * (n0 - lo);
        for (int ks = 0; ks < cp / 16; ++ks)
          add_mma(slot + (uint32_t)(tl * cp / 8 + 2 * ks) * kPlane, kPlane,
                  pl.off_W + (uint32_t)(k.tap_pos[ktap] * cp) * 16 + (uint32_t)(2 * ks) * w_plane, w_plane, false, cnt * cp, acc + n0 * cp, true);
        n0 += cnt;
      }
    }
  };
  auto drop_empty_group = [&]() {
    if (pr.groups.back().count == 0) {
      pr.groups.pop_back();
      order.pop_back();
    }
  };

  // ======================= block 0
  {
    const BlockStatic& k = st.blk[0];
    const int cp = k.cp, accw = Tout[0] * cp;
    new_load(LD_WEIGHTS, pl.off_W, k.tcn_bytes, k.off_tcn);
    {
      // clear the accumulator: zero A operand times the (always resident, finite) mix image, 128 columns per MMA
      new_group();
      for (int c0 = 0; c0 < accw; c0 += 128)
        add_mma(pl.off_const + st.off_zero, st.zero_lbo, pl.off_const + st.off_ablk, kPlane, false, std::min(128, accw - c0), c0, false);
    }
    for (int i = 0; i < n_sl; ++i) {
      const uint32_t slot = pl.off_Q + (uint32_t)(i % ring0_slots) * ring0_slot;
      const int t0 = i * st0, nt = std::min(st0, T - t0);
      // both teams work on every slice, half of the output channels each (8-channel planes [c8, c8 + n) of every time
      // step), so that two slices' MMAs are in flight while the third ring slot is being filled
      const int chunks = cp / 8, halves = kSplitStages && chunks >= 2 ? 2 : 1;
      for (int hh = 0; hh < halves; ++hh) {
        const int c8 = hh * (chunks / halves), n8 = hh + 1 == halves ? chunks - c8 : chunks / halves;
        Stage& s = new_stage(ST_G0, SF_RELU, halves == 2 ? hh : -1);
        s.p0 = t0;
        s.p1 = t0 + nt;
        s.dst_off = slot;
        s.tmem_col = c8;                       // G0: first 8-channel plane and number of planes this stage writes
        s.n_cg = n8;
        order.back().rd = xin_rng;
        for (int tl = 0; tl < nt; ++tl)
          order.back().wr.push_back(smem_r(slot + (uint32_t)(tl * chunks + c8) * kPlane, (uint32_t)n8 * kPlane));
      }
      new_group();
      tcn_mmas(k, 0, slot, t0, nt, 0);
      drop_empty_group();
    }
    new_load(LD_WEIGHTS, pl.off_W, st.blk[1].tcn_bytes, st.blk[1].off_tcn);
    // x1 = relu(acc + residual conv of the raw poses + bias) -> P, chunked like block 1's mix
    const int ctn = ct[1];
    for (int tp0 = 0; tp0 < Tout[0];) {
      int ntp = std::min(ctn, Tout[0] - tp0);
      const uint32_t dst = pl.off_P + (uint32_t)(tp0 * cp / 8) * kPlane;
      uint32_t bytes = (uint32_t)(ntp * cp / 8) * kPlane;
      // the first chunk of x1 that reaches the pose slot takes all remaining columns: ONE stage whose threads first pull
      // their poses into registers, then (team barrier) overwrite the slot -- no later stage needs the poses any more
      const bool hits_xin = dst + bytes > pl.off_xin;
      if (hits_xin) {
        ntp = Tout[0] - tp0;
        bytes = (uint32_t)(ntp * cp / 8) * kPlane;
        if (ntp > 8) return fail("block-0 output chunk over the pose slot longer than 8 time steps");
      }
      Stage& s = new_stage(ST_XEPI0, SF_RELU | (hits_xin ? SF_TEAM_SYNC : 0));
      s.tmem_col = tp0 * cp;
      s.n_cg = ntp * cp / 16;
      s.dst_off = dst;
      s.p0 = tp0;
      s.p1 = tp0 + ntp;
      order.back().rd = xin_rng;
      order.back().rd.push_back(tmem_r(tp0 * cp, ntp * cp));
      order.back().wr.push_back(smem_r(dst, bytes));
      tp0 += ntp;
    }
  }

  // ======================= blocks >= 1
  for (int b = 1; b < nb; ++b) {
    const BlockStatic& k = st.blk[b];
    const bool x_in_P = (b & 1) != 0;
    const uint32_t xb = x_in_P ? pl.off_P : pl.off_Q, ring = x_in_P ? pl.off_Q : pl.off_P;
    const int cin = k.cin_p, cp = k.cp, s = k.stride;
    const int accw = Tout[b] * cp, smw = ct[b] * cin;
    const int acc = x_in_P ? 512 - accw : 0;
    const int SM = x_in_P ? 0 : accw, SG = SM + smw;
    const uint32_t g_img = pl.off_const + k.off_gcn, r_img = pl.off_const + k.off_res, gr_plane = (uint32_t)cp * 16;
    const bool last = b + 1 == nb;
    const int n_slots = std::min(2, nch[b]);
    auto slot_of = [&](int c) { return ring + (uint32_t)(c % n_slots) * slot_bytes[b]; };
    auto nt_of = [&](int c) { return std::min(ct[b], Tin[b] - c * ct[b]); };
    auto mix = [&](int c) {
      new_group();
      const int nt = nt_of(c);
      for (int kk = 0; kk < kRows / 16; ++kk)
        add_mma(pl.off_const + st.off_ablk + (uint32_t)(2 * kk) * kPlane, kPlane, xb + (uint32_t)(c * ct[b] * cin / 8) * kPlane + (uint32_t)kk * 256u, 0,
                true, nt * cin, SM, kk > 0);
    };
    auto mepi = [&](int c) { cvt_stage(SM, nt_of(c) * cin, slot_of(c), 0, 0, 0); };
    auto gcn = [&](int c) {
      new_group();
      for (int tl = 0; tl < nt_of(c); ++tl)
        for (int ks = 0; ks < cin / 16; ++ks)
          add_mma(slot_of(c) + (uint32_t)(tl * cin / 8 + 2 * ks) * kPlane, kPlane, g_img + (uint32_t)(2 * ks) * gr_plane, gr_plane, false, cp,
                  SG + tl * cp, ks > 0);
    };
    auto gepi = [&](int c) { cvt_stage(SG, nt_of(c) * cp, slot_of(c), SF_RELU | SF_BIAS, pl.off_const + k.off_bias_g, cp); };
    auto res = [&]() {
      new_group();
      for (int tp = 0; tp < Tout[b]; ++tp)
        for (int ks = 0; ks < cin / 16; ++ks)
          add_mma(xb + (uint32_t)((s * tp) * cin / 8 + 2 * ks) * kPlane, kPlane, r_img + (uint32_t)(2 * ks) * gr_plane, gr_plane, false, cp,
                  acc + tp * cp, ks > 0);
    };
    mix(0);
    mepi(0);
    for (int c = 0; c < nch[b]; ++c) {
      gcn(c);
      if (c + 1 < nch[b]) mix(c + 1);
      if (c == 0) res();
      gepi(c);
      if (c + 1 < nch[b]) mepi(c + 1);
      new_group();
      tcn_mmas(k, b, slot_of(c), c * ct[b], nt_of(c), acc);
      drop_empty_group();
    }
    if (!last) {
      new_load(LD_WEIGHTS, pl.off_W, st.blk[b + 1].tcn_bytes, st.blk[b + 1].off_tcn);
      const uint32_t xn = x_in_P ? pl.off_Q : pl.off_P;
      const int cw = ct[b + 1] * cp;
      for (int c0 = 0; c0 < accw; c0 += cw)
        cvt_stage(acc + c0, std::min(cw, accw - c0), xn + (uint32_t)(c0 / 8) * kPlane, SF_RELU | SF_BIAS, pl.off_const + k.off_bias_o, cp);
    } else {
      Stage& s2 = new_stage(ST_TOKENS, SF_RELU | SF_BIAS);
      s2.tmem_col = acc;
      s2.n_cg = accw / 16;
      s2.bias_off = pl.off_const + k.off_bias_o;
      s2.bias_period = cp;
      order.back().rd.push_back(tmem_r(acc, accw));
      order.back().wr.push_back(smem_r(pl.off_stage_tok, stage_bytes));
    }
  }

  // ---- the next tile's poses: after the last item of this tile that touches the pose slot (and after every other load:
  // it is the last item of the L sequence)
  {
    size_t at = 0;
    for (size_t i = 0; i < order.size(); ++i)
      if (overlap(order[i].rd, xin_rng) || overlap(order[i].wr, xin_rng) || order[i].side == SIDE_L) at = i + 1;
    Load l;
    memset(&l, 0, sizeof(l));
    l.kind = LD_POSES;
    l.wait_g = l.wait_e[0] = l.wait_e[1] = l.wait_g_prev = -1;
    l.dst_off = pl.off_xin;
    l.bytes = xin_bytes;
    pr.loads.push_back(l);
    order.insert(order.begin() + at, Item{SIDE_L, (int)pr.loads.size() - 1, {}, xin_rng});
  }

  // ---- cross-sequence waits from the read / write sets
  for (size_t i = 0; i < order.size(); ++i) {
    Item& x = order[i];
    int w[N_SIDES] = {-1, -1, -1, -1};
    for (size_t j = 0; j < i; ++j) {
      const Item& y = order[j];
      if (y.side == x.side) continue;
      if (overlap(y.wr, x.rd) || overlap(y.rd, x.wr) || overlap(y.wr, x.wr)) w[y.side] = std::max(w[y.side], y.idx);
    }
    if (x.side == SIDE_G) {
      pr.groups[x.idx].wait_e[0] = (int16_t)w[SIDE_E0];
      pr.groups[x.idx].wait_e[1] = (int16_t)w[SIDE_E1];
      pr.groups[x.idx].wait_l = (int16_t)w[SIDE_L];
    } else if (x.side == SIDE_L) {
      pr.loads[x.idx].wait_g = (int16_t)w[SIDE_G];
      pr.loads[x.idx].wait_e[0] = (int16_t)w[SIDE_E0];
      pr.loads[x.idx].wait_e[1] = (int16_t)w[SIDE_E1];
    } else {
      const int team = x.side - SIDE_E0;
      Stage& s = pr.stages[team][x.idx];
      s.wait_g = w[SIDE_G];
      s.wait_l = w[SIDE_L];
      s.wait_eo = w[SIDE_E0 + (team ^ 1)];
    }
  }
  // a wait that an earlier item of the same sequence already performed (or implied: stages of one team and G commits
  // complete in order) is dropped -- most groups then issue without touching a barrier
  {
    int seen_e[kTeams] = {-1, -1};
    std::vector<char> seen_l(pr.loads.size(), 0);
    for (Group& g : pr.groups) {
      for (int t = 0; t < kTeams; ++t) {
        if (g.wait_e[t] >= 0 && g.wait_e[t] <= seen_e[t]) g.wait_e[t] = -1;
        seen_e[t] = std::max(seen_e[t], (int)g.wait_e[t]);
      }
      if (g.wait_l >= 0) {
        if (seen_l[g.wait_l]) g.wait_l = -1;
        else seen_l[g.wait_l] = 1;
      }
    }
    for (int t = 0; t < kTeams; ++t) {
      int seen_g = -1, seen_o = -1;
      for (Stage& s : pr.stages[t]) {
        if (s.wait_g >= 0 && s.wait_g <= seen_g) s.wait_g = -1;
        seen_g = std::max(seen_g, (int)s.wait_g);
        if (s.wait_eo >= 0 && s.wait_eo <= seen_o) s.wait_eo = -1;
        seen_o = std::max(seen_o, (int)s.wait_eo);
      }
    }
  }
  // tile boundary: each team's first stage overwrites operand regions the previous tile's last MMAs read, the first
  // weight load overwrites the weights they read; every stage that reads the poses waits for the load the previous
  // tile issued (the pose barrier is one completion ahead: the prologue load)
  const int last_g = (int)pr.groups.size() - 1;
  for (int t = 0; t < kTeams; ++t) {
    if (pr.stages[t].empty()) return fail("an epilogue team has no work");
    pr.stages[t][0].wait_g_prev = last_g;
    // (only the team's FIRST pose-reading stage of a tile waits: its later stages run behind it in program order)
    for (Stage& s : pr.stages[t])
      if (s.type == ST_G0 || s.type == ST_XEPI0) {
        s.wait_l = (int)pr.loads.size() - 1;
        break;
      }
  }
  pr.loads[0].wait_g_prev = (int16_t)last_g;
  // ... and the first MMA group overwrites accumulator columns the previous tile's token stage may still be reading
  for (int t = 0; t < kTeams; ++t)
    if (pr.stages[t].back().type == ST_TOKENS) {
      pr.groups[0].prev_team = (int16_t)t;
      pr.groups[0].prev_stage = (int16_t)((int)pr.stages[t].size() - 1);
    }

  // ---- tables + barriers
  pl.n_groups = (int)pr.groups.size();
  pl.n_stages[0] = (int)pr.stages[0].size();
  pl.n_stages[1] = (int)pr.stages[1].size();
  pl.n_loads = (int)pr.loads.size();
  pl.n_mma = (int)pr.mma.size();
  if (pl.n_groups > kMaxGroups || pl.n_stages[0] > kMaxStages || pl.n_stages[1] > kMaxStages || pl.n_loads > kMaxLoads || pl.n_mma > kMaxMma)
    return fail("tile program too long");
  pl.bar_g0 = 0;
  pl.bar_e0[0] = pl.n_groups;
  pl.bar_e0[1] = pl.bar_e0[0] + pl.n_stages[0];
  pl.bar_l0 = pl.bar_e0[1] + pl.n_stages[1];
  pl.n_bars = pl.bar_l0 + pl.n_loads;
  pl.off_bars = off; off += up128((size_t)pl.n_bars * 8);
  pl.off_flags = off; off += 512;
  for (int t = 0; t < kTeams; ++t)
    for (size_t i = 0; i < pr.stages[t].size(); ++i) {
      Stage& s = pr.stages[t][i];
      auto bar = [&](int base, int idx) { return idx >= 0 ? pl.off_bars + 8u * (uint32_t)(base + idx) : 0u; };
      s.bar_g = bar(pl.bar_g0, s.wait_g);
      s.bar_l = bar(pl.bar_l0, s.wait_l);
      s.bar_eo = bar(pl.bar_e0[t ^ 1], s.wait_eo);
      s.bar_g_prev = bar(pl.bar_g0, s.wait_g_prev);
      s.bar_self = bar(pl.bar_e0[t], (int)i);
    }
  pl.smem_bytes = off;
  if ((int)off > max_smem) return fail("tile does not fit shared memory (" + std::to_string(off) + " bytes)");
  if (off >= (1u << 18)) return fail("operand offsets exceed the descriptor range");
  memcpy(pl.groups, pr.groups.data(), pr.groups.size() * sizeof(Group));
  for (int t = 0; t < kTeams; ++t) memcpy(pl.stages[t], pr.stages[t].data(), pr.stages[t].size() * sizeof(Stage));
  memcpy(pl.loads, pr.loads.data(), pr.loads.size() * sizeof(Load));
  memcpy(pl.mma, pr.mma.data(), pr.mma.size() * sizeof(Mma));
  pr.ok = true;
}

}  // namespace t2
}  // namespace sf
