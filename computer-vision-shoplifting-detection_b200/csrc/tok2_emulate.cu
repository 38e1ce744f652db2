// Host emulator of the tokenizer-v2 tile program -- TEST INFRASTRUCTURE, never on a product path.
//
// It executes exactly what tokenizer2_kernel executes: the same descriptor table (decoded the way the tensor core
// decodes no-swizzle K-major / MN-major operands), the same epilogue / prep stages, the same mbarrier waits with the
// hardware's PARITY semantics (a barrier that runs a phase ahead of a waiter is reported, as it would hang the GPU).
// The three item sequences and the eight epilogue warps advance in a pseudo-random interleaving chosen by
// `schedule_seed`; asynchronous items (MMA groups, TMA loads) take effect either at issue or at a later, randomly
// chosen completion point.  A dependency the builder missed shows up as a result that changes with the seed, a
// dependency cycle as a reported deadlock.  Used by tests/test_tok2_program.py on GPU-less hosts.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <deque>

#include "tok2_build.h"

namespace sf {
namespace t2 {
namespace {

inline uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct Emu {
  const Static& st;
  const Program& pr;
  const Plan& pl;
  std::vector<unsigned char> smem;
  std::vector<float> tmem;             // [128][512]
  std::vector<long> done;              // completions per barrier
  std::vector<int> arrivals;           // pending arrivals of the E barriers
  uint64_t rng;
  Emu(const Static& s, const Program& p, uint32_t seed)
      : st(s), pr(p), pl(p.plan), smem(p.plan.smem_bytes, 0), tmem(128 * 512, 0.f), done(p.plan.n_bars, 0), arrivals(p.plan.n_bars, 0),
        rng(seed * 0x9E3779B97F4A7C15ull + 12345) {}
  uint32_t rnd() {
    rng ^= rng << 13;
    rng ^= rng >> 7;
    rng ^= rng << 17;
    return (uint32_t)(rng >> 11);
  }
  // hardware parity wait: passes iff the barrier's current phase parity differs from `par`
  // returns 1 pass, 0 not yet, -1 the barrier is a phase AHEAD of this waiter (would hang)
  int wait(int bar, long need_completions) {
    if (done[bar] == need_completions) return 1;
    if (done[bar] < need_completions) return 0;
    return -1;
  }
  float bf_at(uint32_t byte) const {
    uint16_t h;
    memcpy(&h, &smem[byte], 2);
    return bf2f(h);
  }
  float h_at(uint32_t byte) const {          // fp16 operand element
    uint16_t h;
    memcpy(&h, &smem[byte], 2);
    const int e = (h >> 10) & 31, m = h & 0x3FF;
    float v = e == 0 ? std::ldexp((float)m, -24) : (e == 31 ? (m ? NAN : INFINITY) : std::ldexp((float)(m | 0x400), e - 25));
    return (h & 0x8000) ? -v : v;
  }
  void mma(const Mma& m) {
    const uint32_t a_off = (m.a_lo & 0x3FFFu) << 4, a_lbo = ((m.a_lo >> 16) & 0x3FFFu) << 4;
    const uint32_t b_off = (m.b_lo & 0x3FFFu) << 4, b_lbo = ((m.b_lo >> 16) & 0x3FFFu) << 4;
    const bool mn = (m.idesc >> 16) & 1u;
    const int N = (int)((m.idesc >> 17) & 0x3Fu) << 3;
    const int dcol = (int)(m.d & 0x1FFu);
    const bool acc = (m.d >> 16) & 1u;
    const bool a_f16 = ((m.idesc >> 7) & 7u) == 0;     // instruction descriptor: A format 0 = F16, 1 = BF16
    std::vector<float> A(128 * 16), Bm((size_t)16 * N);
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < 16; ++k) {
        const uint32_t at = a_off + (uint32_t)(k / 8) * a_lbo + (uint32_t)r * 16 + (uint32_t)(k % 8) * 2;
        A[r * 16 + k] = a_f16 ? h_at(at) : bf_at(at);
      }
    const bool b_f16 = ((m.idesc >> 10) & 7u) == 0;
    for (int k = 0; k < 16; ++k)
      for (int n = 0; n < N; ++n) {
        const uint32_t at = mn ? b_off + (uint32_t)(n / 8) * kPlane + (uint32_t)(k / 8) * b_lbo + (uint32_t)(k % 8) * 16 + (uint32_t)(n % 8) * 2
                               : b_off + (uint32_t)(k / 8) * b_lbo + (uint32_t)n * 16 + (uint32_t)(k % 8) * 2;
        Bm[(size_t)k * N + n] = b_f16 ? h_at(at) : bf_at(at);
      }
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < N; ++n) {
        float s = 0.f;
        for (int k = 0; k < 16; ++k) s += A[r * 16 + k] * Bm[(size_t)k * N + n];
        float& d = tmem[(size_t)r * 512 + dcol + n];
        d = acc ? d + s : s;
      }
  }
  void group_effect(int g) {
    const Group& gr = pr.groups[g];
    for (int i = gr.first; i < gr.first + gr.count; ++i) mma(pr.mma[i]);
  }
  void load_effect(const Load& ld, const float* poses, int64_t B, int64_t tile) {
    if (ld.kind == LD_WEIGHTS) {
      memcpy(&smem[ld.dst_off], st.blob.data() + ld.src_off, ld.bytes);
    } else {
      const int64_t w0 = tile * pl.WT;
      const int64_t nw = std::min<int64_t>(pl.WT, B - w0);
      memcpy(&smem[pl.off_xin], poses + (size_t)w0 * pl.per_w, (size_t)nw * pl.per_w * 4);
    }
  }
};

inline uint16_t f2h_sat(float f) {     // round-to-nearest-even fp32 -> fp16, saturating at +-65504 (cvt.rn.satfinite.f16x2.f32)
  if (f != f) return 0x7FFF;
  if (f > 65504.f) f = 65504.f;
  if (f < -65504.f) f = -65504.f;
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
  const int e = (int)((u >> 23) & 0xFF) - 127 + 15;
  uint32_t m = u & 0x7FFFFFu;
  if (e <= 0) {
    if (e < -10) return sign;
    m |= 0x800000u;
    const int sh = 14 - e;
    uint32_t h = m >> sh;
    const uint32_t rem = m & ((1u << sh) - 1), half = 1u << (sh - 1);
    if (rem > half || (rem == half && (h & 1u))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)e << 10) | (m >> 13);
  const uint32_t rem = m & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) ++h;
  if ((h & 0x7FFFu) >= 0x7C00u) h = 0x7BFFu;
  return (uint16_t)(sign | h);
}
inline float h2f(uint16_t h) {
  const int e = (h >> 10) & 31, m = h & 0x3FF;
  float v = e == 0 ? std::ldexp((float)m, -24) : (e == 31 ? (m ? NAN : INFINITY) : std::ldexp((float)(m | 0x400), e - 25));
  return (h & 0x8000) ? -v : v;
}
uint16_t pack_one(float x, bool relu, bool f16) {
  if (relu && !(x > 0.f)) x = 0.f;
  return f16 ? f2h_sat(x) : f2bf(x);
}

}  // namespace

bool emulate(const Static& st, const Program& pr, const float* poses, int64_t B, float* tokens, uint32_t seed, std::string* err) {
  auto fail = [&](const std::string& m) {
    if (err) *err = m;
    return false;
  };
  if (!st.ok || !pr.ok) return fail("program not built: " + (st.ok ? pr.why : st.why));
  Emu E(st, pr, seed);
  const Plan& pl = pr.plan;
  memcpy(&E.smem[pl.off_const], st.blob.data(), st.const_bytes);
  // garbage (finite) in the operand regions: stale data must never reach a result
  for (uint32_t i = pl.off_P; i + 1 < pl.off_bars; i += 2) {
    const uint16_t h = f2bf((float)((int)(E.rnd() % 2001) - 1000) * 1e-3f);
    memcpy(&E.smem[i], &h, 2);
  }
  for (size_t i = 0; i < E.tmem.size(); ++i) E.tmem[i] = (float)((int)(E.rnd() % 2001) - 1000);
  const int64_t n_tiles = (B + pl.WT - 1) / pl.WT;
  if (n_tiles == 0) return true;
  const int V = pl.V, rows = pl.rows, T0 = pl.T0, tv = T0 * V, cp0 = pl.cp0;
  const float* scale = reinterpret_cast<const float*>(&E.smem[pl.off_scale]);
  const float* shift = reinterpret_cast<const float*>(&E.smem[pl.off_shift]);
  const float* coef = reinterpret_cast<const float*>(&E.smem[pl.off_ell]);
  const float* hcs = reinterpret_cast<const float*>(&E.smem[pl.off_hc]);
  std::vector<int> poison(128, 0);
  const float* g0tab = reinterpret_cast<const float*>(&E.smem[pl.off_g0tab]);
  const float* r0tab = reinterpret_cast<const float*>(&E.smem[pl.off_r0tab]);
  auto tabv = [](const float* t, int which, int o) { return t[(o / 4) * 12 + which * 4 + (o % 4)]; };

  struct Pending { int idx; int64_t tile; bool effect_done; };
  int g_next = 0;
  int64_t g_it = 0;
  std::deque<Pending> g_pending;
  int l_next = 0;
  int64_t l_it = 0;
  std::deque<Pending> l_pending;
  struct WarpState { int e = 0; int64_t it = 0; int sub = 0; };
  WarpState ws[kTeams][kTeamWarps];
  // named barrier of each team: generation counter + arrivals
  int named_arrived[kTeams] = {0, 0};
  long named_gen[kTeams] = {0, 0};
  long warp_gen[kTeams][kTeamWarps] = {{0}};
  bool store_pending = false;          // bulk store issued, staging not yet read
  int64_t store_tile = 0;
  int store_nw = 0;
  // XEPI0 keeps the poses of its chunk in "registers" across the team barrier
  std::vector<float> xreg((size_t)kTeams * kRows * 16 * 2, 0.f);

  {
    Load l0 = pr.loads.back();
    E.load_effect(l0, poses, B, 0);
    E.done[pl.bar_l0 + pl.n_loads - 1] = 1;
  }
  auto bar_ok = [&](int bar, long need, bool* ahead) {
    const int w = E.wait(bar, need);
    if (w < 0) *ahead = true;
    return w == 1;
  };
  auto do_store = [&]() {
    const float* stg = reinterpret_cast<const float*>(&E.smem[pl.off_stage_tok]);
    memcpy(tokens + (size_t)store_tile * pl.WT * pl.S_out * pl.d_tok, stg, (size_t)store_nw * pl.S_out * pl.d_tok * 4);
    store_pending = false;
  };
  auto put16 = [&](uint32_t dst_off, int row, int cg, const float* a, bool relu) {
    for (int j = 0; j < 16; ++j) {
      const uint16_t h = pack_one(a[j], relu, st.f16);
      const int col = cg * 16 + j;
      memcpy(&E.smem[dst_off + (uint32_t)(col / 8) * kPlane + (uint32_t)row * 16 + (uint32_t)(col % 8) * 2], &h, 2);
    }
  };
  const bool eager = seed == 0;
  long steps = 0;
  std::string ahead_msg;
  for (;;) {
    const bool g_done = g_it >= n_tiles && g_pending.empty();
    const bool l_done = l_it >= n_tiles && l_pending.empty();
    bool e_done = true;
    for (auto& tm : ws)
      for (auto& w : tm) e_done &= w.it >= n_tiles;
    if (g_done && l_done && e_done && !store_pending) break;
    if (++steps > 80000000) return fail("emulator: step limit");
    enum { A_G_ISSUE, A_G_DONE, A_L_ISSUE, A_L_DONE, A_E0, A_STORE = 100 };
    std::vector<int> acts;
    bool ahead = false;
    if (g_it < n_tiles && (int)g_pending.size() < 6) {
      const Group& gr = pr.groups[g_next];
      bool ok = true;
      for (int t = 0; t < kTeams; ++t)
        if (gr.wait_e[t] >= 0) ok &= bar_ok(pl.bar_e0[t] + gr.wait_e[t], g_it + 1, &ahead);
      if (gr.wait_l >= 0) ok &= bar_ok(pl.bar_l0 + gr.wait_l, g_it + 1, &ahead);
      if (gr.prev_stage >= 0 && g_it > 0) ok &= bar_ok(pl.bar_e0[gr.prev_team] + gr.prev_stage, g_it, &ahead);
      if (ahead) ahead_msg = "G group " + std::to_string(g_next);
      if (ok) acts.push_back(A_G_ISSUE);
    }
    if (!g_pending.empty()) acts.push_back(A_G_DONE);
    if (l_it < n_tiles && (int)l_pending.size() < 4) {
      const Load& ld = pr.loads[l_next];
      bool ok = true, ah = false;
      if (ld.wait_g >= 0) ok &= bar_ok(pl.bar_g0 + ld.wait_g, l_it + 1, &ah);
      for (int t = 0; t < kTeams; ++t)
        if (ld.wait_e[t] >= 0) ok &= bar_ok(pl.bar_e0[t] + ld.wait_e[t], l_it + 1, &ah);
      if (ld.wait_g_prev >= 0 && l_it > 0) ok &= bar_ok(pl.bar_g0 + ld.wait_g_prev, l_it, &ah);
      if (ah) { ahead = true; ahead_msg = "L load " + std::to_string(l_next); }
      if (ok) acts.push_back(A_L_ISSUE);
    }
    if (!l_pending.empty()) acts.push_back(A_L_DONE);
    for (int tm = 0; tm < kTeams; ++tm)
      for (int w = 0; w < kTeamWarps; ++w) {
        WarpState& W = ws[tm][w];
        if (W.it >= n_tiles) continue;
        const Stage& s = pr.stages[tm][W.e];
        bool ok = true, ah = false;
        if (W.sub == 0) {
          if (s.wait_g >= 0) ok &= bar_ok(pl.bar_g0 + s.wait_g, W.it + 1, &ah);
          if (s.wait_l >= 0) ok &= bar_ok(pl.bar_l0 + s.wait_l, W.it + 1, &ah);
          if (s.wait_eo >= 0) ok &= bar_ok(pl.bar_e0[tm ^ 1] + s.wait_eo, W.it + 1, &ah);
          if (s.wait_g_prev >= 0 && W.it > 0) ok &= bar_ok(pl.bar_g0 + s.wait_g_prev, W.it, &ah);
        } else if (W.sub == 3) {
          ok = !store_pending;                        // cp.async.bulk.wait_group.read
        } else if (s.type == ST_TOKENS && W.sub == 1) {
          ok = true;                                  // no barrier before the staging writes any more
        } else {
          ok = named_gen[tm] > warp_gen[tm][w];       // parked at the team's named barrier
        }
        if (ah) { ahead = true; ahead_msg = "E team " + std::to_string(tm) + " stage " + std::to_string(W.e) + " warp " + std::to_string(w); }
        if (ok) acts.push_back(A_E0 + tm * kTeamWarps + w);
      }
    if (store_pending) acts.push_back(A_STORE);
    if (ahead) return fail("emulator: a barrier ran a phase ahead of a waiter (" + ahead_msg + "): the GPU would hang");
    if (acts.empty()) {
      std::string m = "emulator: deadlock at G " + std::to_string(g_next) + "/it " + std::to_string(g_it) + ", L " + std::to_string(l_next) + ", E";
      for (auto& tm : ws)
        for (auto& w : tm) m += " " + std::to_string(w.e) + "." + std::to_string(w.sub);
      return fail(m);
    }
    const int act = eager ? acts[0] : acts[E.rnd() % acts.size()];
    if (act == A_G_ISSUE) {
      // MMAs execute in issue order: a group can only take effect at issue if nothing older is still in flight
      bool older_in_flight = false;
      for (const Pending& p : g_pending) older_in_flight |= !p.effect_done;
      const bool now = eager || (!older_in_flight && (E.rnd() & 1));
      if (now) E.group_effect(g_next);
      g_pending.push_back(Pending{g_next, g_it, now});
      if (++g_next == pl.n_groups) { g_next = 0; ++g_it; }
    } else if (act == A_G_DONE) {
      Pending p = g_pending.front();
      g_pending.pop_front();
      if (!p.effect_done) E.group_effect(p.idx);
      E.done[pl.bar_g0 + p.idx]++;
    } else if (act == A_L_ISSUE) {
      const Load& ld = pr.loads[l_next];
      const bool skip = ld.kind == LD_POSES && l_it + 1 >= n_tiles;
      if (!skip) {
        const bool now = eager || (E.rnd() & 1);
        if (now) E.load_effect(ld, poses, B, l_it + 1);
        l_pending.push_back(Pending{l_next, l_it + 1, now});
      }
      if (++l_next == pl.n_loads) { l_next = 0; ++l_it; }
    } else if (act == A_L_DONE) {
      Pending p = l_pending.front();
      l_pending.pop_front();
      if (!p.effect_done) E.load_effect(pr.loads[p.idx], poses, B, p.tile);
      E.done[pl.bar_l0 + p.idx]++;
    } else if (act == A_STORE) {
      do_store();
    } else {
      const int tm = (act - A_E0) / kTeamWarps, w = (act - A_E0) % kTeamWarps;
      WarpState& W = ws[tm][w];
      const Stage& s = pr.stages[tm][W.e];
      const int q = w & 3, half = w >> 2;
      const int64_t tile = W.it;
      const int nw = (int)std::min<int64_t>(pl.WT, B - tile * pl.WT);
      const int par = (int)(W.it & 1);
      auto arrive_named = [&]() {
        warp_gen[tm][w] = named_gen[tm];
        if (++named_arrived[tm] == kTeamWarps) {
          named_arrived[tm] = 0;
          named_gen[tm]++;
        }
      };
      const float* xin = reinterpret_cast<const float*>(&E.smem[pl.off_xin]);
      bool finish = false;
      if (s.type == ST_CVT) {
        for (int cg = half; cg < s.n_cg; cg += 2)
          for (int ln = 0; ln < 32; ++ln) {
            const int row = q * 32 + ln;
            float a[16];
            for (int j = 0; j < 16; ++j) {
              a[j] = E.tmem[(size_t)row * 512 + s.tmem_col + cg * 16 + j];
              if (s.flags & SF_BIAS) a[j] += reinterpret_cast<const float*>(&E.smem[s.bias_off])[(cg * 16) % s.bias_period + j];
            }
            put16(s.dst_off, row, cg, a, s.flags & SF_RELU);
          }
        finish = true;
      } else if (s.type == ST_G0) {
        for (int ln = 0; ln < 32; ++ln) {
          const int row = q * 32 + ln, ww = row / V, v = row - ww * V;
          const bool valid = row < rows && ww < nw;
          bool bad = false;
          const float* xw = xin + ww * pl.per_w + v;
          const int t = s.p0 + half;                           // thread = (row, time step p0 + half)
          if (t < s.p1) {
            float m[2] = {0.f, 0.f};
            if (valid) {
              m[0] = hcs[v * 2];
              m[1] = hcs[v * 2 + 1];
              for (int k = 0; k < pl.ell_width; ++k) {
                const float* e = coef + (size_t)(k * V + v) * 4;
                int dl;
                memcpy(&dl, &e[2], 4);
                for (int c = 0; c < pl.c_in; ++c) {
                  const float xv = xw[t * V + c * tv + dl];
                  bad |= !(std::fabs(xv) <= 3.0e38f);
                  m[c] = std::fmaf(e[c], xv, m[c]);
                }
              }
              if (bad) m[0] = m[1] = 0.f;
            }
            for (int o = s.tmem_col * 8; o < (s.tmem_col + s.n_cg) * 8; ++o) {
              uint16_t hb;
              if (st.f16) {
                // packed half arithmetic of the kernel: two fused multiply-adds, each rounded once to fp16
                const uint16_t* th = reinterpret_cast<const uint16_t*>(&E.smem[pl.off_g0tab_h]) + (o / 8) * 24 + (o % 8);
                const double mxh = h2f(f2h_sat(m[0])), myh = h2f(f2h_sat(m[1]));
                const double inner = h2f(f2h_sat((float)(myh * h2f(th[8]) + h2f(th[16]))));
                double y = mxh * h2f(th[0]) + inner;
                if (!(y > 0.0)) y = 0.0;
                hb = f2h_sat((float)y);
              } else {
                const float y = std::fmaf(m[0], tabv(g0tab, 0, o), std::fmaf(m[1], tabv(g0tab, 1, o), tabv(g0tab, 2, o)));
                hb = pack_one(y, true, false);
              }
              const int col = (t - s.p0) * cp0 + o;
              memcpy(&E.smem[s.dst_off + (uint32_t)(col / 8) * kPlane + (uint32_t)row * 16 + (uint32_t)(col % 8) * 2], &hb, 2);
            }
          }
          if (bad) poison[par * 64 + ww] = 1;
        }
        finish = true;
      } else if (s.type == ST_XEPI0) {
        float* xr = &xreg[(size_t)tm * kRows * 32];
        auto body = [&]() {
          for (int ln = 0; ln < 32; ++ln) {
            const int row = q * 32 + ln;
            for (int i = 0; i < s.p1 - s.p0; ++i)
              for (int cg = half; cg < cp0 / 16; cg += 2) {
                float a[16];
                for (int j = 0; j < 16; ++j) {
                  const int o = cg * 16 + j;
                  a[j] = E.tmem[(size_t)row * 512 + s.tmem_col + i * cp0 + o] +
                         std::fmaf(xr[row * 32 + i], tabv(r0tab, 0, o), std::fmaf(xr[row * 32 + 16 + i], tabv(r0tab, 1, o), tabv(r0tab, 2, o)));
                }
                put16(s.dst_off, row, i * (cp0 / 16) + cg, a, true);
              }
          }
        };
        if (W.sub == 0) {
          for (int ln = 0; ln < 32; ++ln) {
            const int row = q * 32 + ln, ww = row / V, v = row - ww * V;
            const bool valid = row < rows && ww < nw;
            for (int i = 0; i < s.p1 - s.p0; ++i) {
              float xa = 0.f, xb = 0.f;
              if (valid) {
                const float* xp = xin + ww * pl.per_w + v + pl.stride0 * (s.p0 + i) * V;
                float u = xp[0], wv = pl.c_in > 1 ? xp[tv] : 0.f;
                if (!(std::fabs(u) <= 3.0e38f) || !(std::fabs(wv) <= 3.0e38f)) u = wv = 0.f;
                xa = std::fmaf(u, scale[v], shift[v]);
                xb = pl.c_in > 1 ? std::fmaf(wv, scale[V + v], shift[V + v]) : 0.f;
              }
              xr[row * 32 + i] = xa;
              xr[row * 32 + 16 + i] = xb;
            }
          }
          if (s.flags & SF_TEAM_SYNC) {
            arrive_named();
            W.sub = 1;
            continue;
          }
          body();
          finish = true;
        } else {
          body();
          finish = true;
        }
      } else {   // ST_TOKENS: drain + barrier, staging writes + barrier, store
        if (W.sub == 0) {
          if (store_pending) return fail("emulator: token stage entered while the previous bulk store has not read the staging area");
          W.sub = 1;
          continue;
        } else if (W.sub == 1) {
          float* stg = reinterpret_cast<float*>(&E.smem[pl.off_stage_tok]);
          const float* bp = reinterpret_cast<const float*>(&E.smem[s.bias_off]);
          if (pl.pool > 0) {
            for (int j = half; j < pl.pool; j += 2) {
              const int t0 = (j * pl.T_last) / pl.pool, t1 = ((j + 1) * pl.T_last + pl.pool - 1) / pl.pool;
              const float inv = 1.f / (float)(t1 - t0);
              for (int ln = 0; ln < 32; ++ln) {
                const int row = q * 32 + ln, mw = row / V, mv = row - mw * V;
                if (!(row < rows && mw < nw)) continue;
                for (int c = 0; c < pl.c_last; ++c) {
                  float acc = 0.f;
                  for (int t = t0; t < t1; ++t) {
                    const float y = E.tmem[(size_t)row * 512 + s.tmem_col + t * pl.cp_last + c] + bp[c];
                    acc += y > 0.f ? y : 0.f;
                  }
                  float y = acc * inv;
                  if (poison[par * 64 + mw]) y = NAN;
                  stg[(mw * pl.S_out + j) * pl.d_tok + c * V + mv] = y;
                }
              }
            }
          } else
          for (int cg = half; cg < s.n_cg; cg += 2)
            for (int ln = 0; ln < 32; ++ln) {
              const int row = q * 32 + ln, mw = row / V, mv = row - mw * V;
              if (!(row < rows && mw < nw)) continue;
              for (int j = 0; j < 16; ++j) {
                const int col = cg * 16 + j, t = col / pl.cp_last, c = col - t * pl.cp_last;
                if (c >= pl.c_last) continue;
                float y = E.tmem[(size_t)row * 512 + s.tmem_col + col] + bp[c];
                y = y > 0.f ? y : 0.f;
                if (poison[par * 64 + mw]) y = NAN;
                stg[(mw * pl.S_out + t) * pl.d_tok + c * V + mv] = y;
              }
            }
          arrive_named();
          W.sub = 2;
          continue;
        } else {
          if (w < 2)
            for (int i = w * 32; i < w * 32 + 32; ++i) poison[par * 64 + i] = 0;
          if (w == 0 && W.sub == 2) {
            // thread 0 issues the bulk store and waits for it to have read the staging area (which aliases activation storage)
            if (store_pending) return fail("emulator: token staging overwritten while a bulk store is pending");
            store_pending = true;
            store_tile = tile;
            store_nw = nw;
            W.sub = 3;
            continue;
          }
          finish = true;
        }
      }
      if (finish) {
        const int bar = pl.bar_e0[tm] + W.e;
        if (++E.arrivals[bar] == kTeamWarps) {
          E.arrivals[bar] = 0;
          E.done[bar]++;
        }
        W.sub = 0;
        if (++W.e == pl.n_stages[tm]) { W.e = 0; ++W.it; }
      }
    }
  }
  return true;
}

}  // namespace t2
}  // namespace sf
