// sf_model_create: validate the state dict against the config, fold every eval-mode
// BatchNorm into its neighbour, transpose the linears for coalesced reads and upload one
// immutable arena.  Reference for the folded maths: SURVEY Appendix A.1-A.3
// (shopformer/models/gcae.py:242-259,331-366; shopformer/models/transformer.py:60-196,304-329;
//  shopformer_2/models/transformer.py:105-194).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "sf_internal.h"

namespace sf {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

namespace {

constexpr double kBnEps = 1e-5;

struct HostTensor {
  const float* p;
  int64_t n;
};

struct Packer {
  std::map<std::string, HostTensor> sd;
  std::vector<float> arena;      // staged host copy, 64-float (256 B) aligned sections
  std::vector<int> iarena_dummy;
  int err = SF_OK;

  const float* get(const std::string& key, int64_t want) {
    auto it = sd.find(key);
    if (it == sd.end()) {
      if (err == SF_OK) {
        set_error("state dict is missing '%s'", key.c_str());
        err = SF_E_MISSING;
      }
      return nullptr;
    }
    if (it->second.n != want) {
      if (err == SF_OK) {
        set_error("'%s' has %lld elements, config implies %lld", key.c_str(), (long long)it->second.n,
                  (long long)want);
        err = SF_E_SHAPE;
      }
      return nullptr;
    }
    return it->second.p;
  }
  bool has(const std::string& key) const { return sd.count(key) != 0; }

  // reserve n floats, returns offset (in floats)
  size_t alloc(size_t n) {
    size_t off = (arena.size() + 63) & ~size_t(63);
    arena.resize(off + n, 0.f);
    return off;
  }
};

// BatchNorm (eval) -> per-channel scale/shift in double
static void bn_fold(Packer& pk, const std::string& p, int n, std::vector<double>& scale,
                    std::vector<double>& shift) {
  const float* w = pk.get(p + "weight", n);
  const float* b = pk.get(p + "bias", n);
  const float* rm = pk.get(p + "running_mean", n);
  const float* rv = pk.get(p + "running_var", n);
  scale.assign(n, 1.0);
  shift.assign(n, 0.0);
  if (!w || !b || !rm || !rv) return;
  for (int i = 0; i < n; ++i) {
    double s = (double)w[i] / std::sqrt((double)rv[i] + kBnEps);
    scale[i] = s;
    shift[i] = (double)b[i] - (double)rm[i] * s;
  }
}

static inline uint16_t f2bf(float f) {       // round-to-nearest-even fp32 -> bf16
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline uint16_t f2h(float f) {        // round-to-nearest-even fp32 -> fp16 (|f| < 65520)
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint16_t sign = (uint16_t)((u >> 16) & 0x8000u);
  const int e = (int)((u >> 23) & 0xFF) - 127 + 15;
  uint32_t m = u & 0x7FFFFFu;
  if (e >= 31) return (uint16_t)(sign | 0x7C00u);
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
  return (uint16_t)(sign | h);
}
static inline int pad16(int x) { return (x + 15) & ~15; }

struct LinOff {
  size_t wt, b;
  int K, N;
};

// nn.Linear weight (N,K) row-major -> [K][N]
static LinOff pack_linear(Packer& pk, const std::string& wkey, const std::string& bkey, int K, int N) {
  LinOff o{0, 0, K, N};
  const float* w = pk.get(wkey, (int64_t)K * N);
  const float* b = pk.get(bkey, N);
  o.wt = pk.alloc((size_t)K * N);
  o.b = pk.alloc(N);
  if (!w || !b) return o;
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) pk.arena[o.wt + (size_t)k * N + n] = w[(size_t)n * K + k];
  memcpy(&pk.arena[o.b], b, sizeof(float) * N);
  return o;
}

struct NormOff {
  size_t g, b;
};
static NormOff pack_norm(Packer& pk, const std::string& p, int d) {
  // padded to the widest stream (160 columns) with zeros: the tensor-core path applies gamma / beta to padded
  // columns without masking (0 * x + 0)
  NormOff o{pk.alloc(std::max(d, 160)), 0};
  o.b = pk.alloc(std::max(d, 160));
  const float* g = pk.get(p + "weight", d);
  const float* b = pk.get(p + "bias", d);
  if (g && b) {
    memcpy(&pk.arena[o.g], g, sizeof(float) * d);
    memcpy(&pk.arena[o.b], b, sizeof(float) * d);
  }
  return o;
}

struct AttnOff {
  LinOff qkv, out;
};
static AttnOff pack_attn(Packer& pk, const std::string& p, int d) {
  AttnOff a;
  a.qkv = pack_linear(pk, p + "in_proj_weight", p + "in_proj_bias", d, 3 * d);
  a.out = pack_linear(pk, p + "out_proj.weight", p + "out_proj.bias", d, d);
  return a;
}

}  // namespace
}  // namespace sf

using namespace sf;

extern "C" int sf_abi_version(void) { return SF_ABI_VERSION; }
extern "C" const char* sf_last_error(void) { return sf::last_error(); }

extern "C" int sf_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    set_error("no CUDA device visible (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return SF_E_NODEVICE;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ++ok;
  }
  if (!ok) {
    set_error("no sm_100 device among %d CUDA devices; kernels are built for sm_100a only", n);
    return SF_E_NODEVICE;
  }
  return ok;
}

extern "C" int sf_model_create(const sf_config* cfg, int32_t n_tensors, const char* const* names,
                               const float* const* data_host, const int64_t* numel, int32_t device,
                               sf_model** out) {
  SF_REQUIRE(cfg && names && data_host && numel && out, SF_E_INVALID, "sf_model_create: null argument");
  *out = nullptr;
  SF_REQUIRE(cfg->variant == SF_VARIANT_SHOPFORMER || cfg->variant == SF_VARIANT_SHOPFORMER_2, SF_E_INVALID,
             "unknown variant %d", cfg->variant);
  const int V = cfg->num_keypoints, nb = cfg->n_blocks, d = cfg->d_model, H = cfg->n_heads;
  SF_REQUIRE(V >= 1 && V <= kMaxV, SF_E_UNSUPPORTED, "num_keypoints=%d outside [1,%d]", V, kMaxV);
  SF_REQUIRE(nb >= 1 && nb <= kMaxBlocks, SF_E_UNSUPPORTED, "n_blocks=%d outside [1,%d]", nb, kMaxBlocks);
  SF_REQUIRE(cfg->n_enc_layers >= 1 && cfg->n_enc_layers <= kMaxLayers && cfg->n_dec_layers >= 1 &&
                 cfg->n_dec_layers <= kMaxLayers,
             SF_E_UNSUPPORTED, "transformer depth %d+%d outside [1,%d]", cfg->n_enc_layers, cfg->n_dec_layers,
             kMaxLayers);
  SF_REQUIRE(cfg->channels[0] == cfg->in_channels && cfg->in_channels >= 1, SF_E_INVALID,
             "channels[0] must equal in_channels");
  for (int i = 0; i < nb; ++i) {
    SF_REQUIRE(cfg->channels[i + 1] >= 1 && cfg->channels[i + 1] % 4 == 0, SF_E_UNSUPPORTED,
               "block %d: out channels %d must be a positive multiple of 4", i, cfg->channels[i + 1]);
    SF_REQUIRE(cfg->strides[i] >= 1 && cfg->strides[i] <= 8, SF_E_UNSUPPORTED, "block %d: stride %d outside [1,8]",
               i, cfg->strides[i]);
  }
  const int d_tok = cfg->channels[nb] * V;
  SF_REQUIRE(d >= 4 && d % 4 == 0 && H >= 1 && d % H == 0, SF_E_UNSUPPORTED,
             "d_model=%d must be a multiple of 4 and of n_heads=%d", d, H);
  SF_REQUIRE(cfg->d_ff >= 4 && cfg->d_ff % 4 == 0, SF_E_UNSUPPORTED, "d_ff=%d must be a multiple of 4", cfg->d_ff);
  SF_REQUIRE(d_tok % 4 == 0, SF_E_UNSUPPORTED, "token width %d must be a multiple of 4", d_tok);
  if (cfg->variant == SF_VARIANT_SHOPFORMER)
    SF_REQUIRE(d == d_tok, SF_E_INVALID, "variant 1 needs d_model == latent*V (%d vs %d)", d, d_tok);

  // device < 0: HOST-ONLY model for the tile-program emulator (sfdbg_tok2_emulate, test infrastructure).  Every compute
  // entry point rejects such a model; nothing below touches CUDA for it.
  const bool host_only = device < 0;
  cudaDeviceProp prop;
  memset(&prop, 0, sizeof(prop));
  DeviceGuard guard;
  if (!host_only) {
    int rc = sf_device_count();
    if (rc < 0) return rc;
    SF_CUDA_OK(guard.enter(device));
    SF_CUDA_OK(cudaGetDeviceProperties(&prop, device));
    SF_REQUIRE(prop.major == 10, SF_E_NODEVICE, "device %d is sm_%d%d; this build targets sm_100a only", device,
               prop.major, prop.minor);
  } else {
    prop.multiProcessorCount = 148;
    prop.sharedMemPerBlockOptin = 232448;
  }

  Packer pk;
  for (int i = 0; i < n_tensors; ++i) pk.sd[names[i]] = HostTensor{data_host[i], numel[i]};

  // ------------------------------------------------------------- tokenizer
  const std::string enc = "gcae.encoder.";
  const int c0 = cfg->in_channels;
  size_t off_in_scale = pk.alloc((size_t)c0 * V), off_in_shift = pk.alloc((size_t)c0 * V);
  {
    std::vector<double> s, t;
    bn_fold(pk, enc + "bn_input.", c0 * V, s, t);
    for (int i = 0; i < c0 * V; ++i) {
      pk.arena[off_in_scale + i] = (float)s[i];
      pk.arena[off_in_shift + i] = (float)t[i];
    }
  }
  struct BlkOff {
    size_t gw, gb, tw, rw, ob, ev, ec;
    int ellw, identity;
  } bo[kMaxBlocks];
  std::vector<std::vector<int>> ell_cols(nb);
  for (int i = 0; i < nb; ++i) {
    const int ci = cfg->channels[i], co = cfg->channels[i + 1];
    const std::string p = enc + "layers." + std::to_string(i) + ".";
    BlkOff& b = bo[i];
    // gcn
    b.gw = pk.alloc((size_t)ci * co);
    b.gb = pk.alloc(co);
    if (const float* w = pk.get(p + "gcn.weight", (int64_t)ci * co)) memcpy(&pk.arena[b.gw], w, sizeof(float) * ci * co);
    if (const float* w = pk.get(p + "gcn.bias", co)) memcpy(&pk.arena[b.gb], w, sizeof(float) * co);
    // adjacency -> ELL (exactly the non-zeros of the checkpoint's buffer)
    const float* adj = pk.get(p + "gcn.adj", (int64_t)V * V);
    int width = 1;
    if (adj)
      for (int v = 0; v < V; ++v) {
        int nnz = 0;
        for (int u = 0; u < V; ++u) nnz += adj[v * V + u] != 0.f;
        width = nnz > width ? nnz : width;
      }
    b.ellw = width;
    b.ev = pk.alloc((size_t)V * width);
    b.ec = pk.alloc((size_t)V * width);   // ints stored in the float arena (bit pattern)
    if (adj)
      for (int v = 0; v < V; ++v) {
        int j = 0;
        for (int u = 0; u < V; ++u)
          if (adj[v * V + u] != 0.f) {
            pk.arena[b.ev + (size_t)v * width + j] = adj[v * V + u];
            int col = u;
            memcpy(&pk.arena[b.ec + (size_t)v * width + j], &col, 4);
            ++j;
          }
        for (; j < width; ++j) {
          int col = v;
          pk.arena[b.ev + (size_t)v * width + j] = 0.f;
          memcpy(&pk.arena[b.ec + (size_t)v * width + j], &col, 4);
        }
      }
    // temporal conv (co,co,9,1) + BN2d -> [c][k][o] * scale[o]
    std::vector<double> ts, tt;
    bn_fold(pk, p + "tcn.bn.", co, ts, tt);
    b.tw = pk.alloc((size_t)co * kTaps * co);
    b.ob = pk.alloc(co);
    const float* tw = pk.get(p + "tcn.conv.weight", (int64_t)co * co * kTaps);
    const float* tb = pk.get(p + "tcn.conv.bias", co);
    if (tw && tb)
      for (int o = 0; o < co; ++o) {
        for (int c = 0; c < co; ++c)
          for (int k = 0; k < kTaps; ++k)
            pk.arena[b.tw + ((size_t)c * kTaps + k) * co + o] = (float)((double)tw[((size_t)o * co + c) * kTaps + k] * ts[o]);
        pk.arena[b.ob + o] = (float)((double)tb[o] * ts[o] + tt[o]);
      }
    // residual: identity iff no conv in the state dict (gcae.py:234-240)
    b.identity = !pk.has(p + "residual.0.weight");
    b.rw = 0;
    if (b.identity) {
      SF_REQUIRE(ci == co && cfg->strides[i] == 1, SF_E_MISSING,
                 "block %d has no residual conv in the state dict but cin!=cout or stride!=1", i);
    } else {
      std::vector<double> rs, rt;
      bn_fold(pk, p + "residual.1.", co, rs, rt);
      b.rw = pk.alloc((size_t)ci * co);
      const float* rw = pk.get(p + "residual.0.weight", (int64_t)co * ci);
      const float* rb = pk.get(p + "residual.0.bias", co);
      if (rw && rb)
        for (int o = 0; o < co; ++o) {
          for (int c = 0; c < ci; ++c) pk.arena[b.rw + (size_t)c * co + o] = (float)((double)rw[(size_t)o * ci + c] * rs[o]);
          pk.arena[b.ob + o] = (float)((double)pk.arena[b.ob + o] + (double)rb[o] * rs[o] + rt[o]);
        }
    }
  }

  // ------------------------------------------------------------- bf16 operand images (tcgen05 path)
  std::vector<uint16_t> bf;                       // 16-byte aligned sections
  struct BfOff { size_t tcn, gcn, res, gb, ob; int npad, kin; bool has_gcn, has_res; } bfo[kMaxBlocks];
  auto bf_alloc = [&](size_t n) { size_t o = (bf.size() + 63) & ~size_t(63); bf.resize(o + n, 0); return o; };
  if (pk.err == SF_OK)
    for (int i = 0; i < nb; ++i) {
      const int ci = cfg->channels[i], co = cfg->channels[i + 1];
      const int npad = pad16(co), kin = i == 0 ? ci : pad16(ci);
      BfOff& o = bfo[i];
      o.npad = npad;
      o.kin = kin;
      o.has_gcn = i > 0;
      o.has_res = i > 0 && !bo[i].identity;
      // temporal conv: chunk (k, jc) -> [n][8], element = Kf[c = jc*8+e][k][n]
      o.tcn = bf_alloc((size_t)kTaps * npad * npad);
      for (int k = 0; k < kTaps; ++k)
        for (int jc = 0; jc < npad / 8; ++jc)
          for (int n = 0; n < npad; ++n)
            for (int e = 0; e < 8; ++e) {
              const int c = jc * 8 + e;
              const float v = (c < co && n < co) ? pk.arena[bo[i].tw + ((size_t)c * kTaps + k) * co + n] : 0.f;
              bf[o.tcn + (((size_t)k * (npad / 8) + jc) * npad + n) * 8 + e] = f2bf(v);
            }
      auto pack_kn = [&](size_t src, size_t dst) {   // fp32 [c][o] -> chunks [jc][n][8]
        for (int jc = 0; jc < kin / 8; ++jc)
          for (int n = 0; n < npad; ++n)
            for (int e = 0; e < 8; ++e) {
              const int c = jc * 8 + e;
              bf[dst + ((size_t)jc * npad + n) * 8 + e] = f2bf((c < ci && n < co) ? pk.arena[src + (size_t)c * co + n] : 0.f);
            }
      };
      o.gcn = o.res = 0;
      if (o.has_gcn) {
        o.gcn = bf_alloc((size_t)kin * npad);
        pack_kn(bo[i].gw, o.gcn);
      }
      if (o.has_res) {
        o.res = bf_alloc((size_t)kin * npad);
        pack_kn(bo[i].rw, o.res);
      }
      o.gb = pk.alloc(npad);
      o.ob = pk.alloc(npad);
      for (int n = 0; n < co; ++n) {
        pk.arena[o.gb + n] = pk.arena[bo[i].gb + n];
        pk.arena[o.ob + n] = pk.arena[bo[i].ob + n];
      }
    }

  // ------------------------------------------------------------- transformer
  const std::string tp = "transformer.";
  const bool v1 = cfg->variant == SF_VARIANT_SHOPFORMER;
  const int pe_rows = 100;
  size_t off_pe = pk.alloc((size_t)pe_rows * d), off_pe_score = 0;
  if (const float* pe = pk.get(tp + "pos_encoder.pe", (int64_t)pe_rows * d)) memcpy(&pk.arena[off_pe], pe, sizeof(float) * pe_rows * d);
  if (v1) {
    off_pe_score = pk.alloc((size_t)pe_rows * d_tok);
    if (const float* pe = pk.get("pos_encoder.pe", (int64_t)pe_rows * d_tok))
      memcpy(&pk.arena[off_pe_score], pe, sizeof(float) * pe_rows * d_tok);
  }
  const bool io_proj = !v1 && pk.has(tp + "input_projection.weight");
  if (!v1 && !io_proj)
    SF_REQUIRE(d == d_tok, SF_E_MISSING, "variant 2 without input_projection needs d_model == latent*V (%d vs %d)", d, d_tok);
  LinOff in_proj{}, out_proj{};
  if (io_proj) {
    in_proj = pack_linear(pk, tp + "input_projection.weight", tp + "input_projection.bias", d_tok, d);
    out_proj = pack_linear(pk, tp + "output_projection.weight", tp + "output_projection.bias", d, d_tok);
  } else if (v1) {
    out_proj = pack_linear(pk, tp + "output_proj.weight", tp + "output_proj.bias", d, d);
  }
  struct EncOff { AttnOff sa; LinOff f1, f2; NormOff n1, n2; } eo[kMaxLayers];
  struct DecOff { AttnOff sa, ca; LinOff f1, f2; NormOff n1, n2, n3; } dof[kMaxLayers];
  const std::string encp = v1 ? tp + "encoder_layers." : tp + "encoder.layers.";
  const std::string decp = v1 ? tp + "decoder_layers." : tp + "decoder.layers.";
  for (int i = 0; i < cfg->n_enc_layers; ++i) {
    const std::string p = encp + std::to_string(i) + ".";
    eo[i].sa = pack_attn(pk, p + "self_attn.", d);
    eo[i].f1 = pack_linear(pk, p + "linear1.weight", p + "linear1.bias", d, cfg->d_ff);
    eo[i].f2 = pack_linear(pk, p + "linear2.weight", p + "linear2.bias", cfg->d_ff, d);
    eo[i].n1 = pack_norm(pk, p + "norm1.", d);
    eo[i].n2 = pack_norm(pk, p + "norm2.", d);
  }
  for (int i = 0; i < cfg->n_dec_layers; ++i) {
    const std::string p = decp + std::to_string(i) + ".";
    dof[i].sa = pack_attn(pk, p + "self_attn.", d);
    dof[i].ca = pack_attn(pk, p + "multihead_attn.", d);
    dof[i].f1 = pack_linear(pk, p + "linear1.weight", p + "linear1.bias", d, cfg->d_ff);
    dof[i].f2 = pack_linear(pk, p + "linear2.weight", p + "linear2.bias", cfg->d_ff, d);
    dof[i].n1 = pack_norm(pk, p + "norm1.", d);
    dof[i].n2 = pack_norm(pk, p + "norm2.", d);
    dof[i].n3 = pack_norm(pk, p + "norm3.", d);
  }
  NormOff enc_norm{}, dec_norm{};
  if (!v1) {
    enc_norm = pack_norm(pk, tp + "encoder.norm.", d);
    dec_norm = pack_norm(pk, tp + "decoder.norm.", d);
  }
  if (pk.err != SF_OK) return pk.err;

  // ------------------------------------------------------------- tensor-core transformer program
  // Host-side op list with OFFSETS (bf16 image offset into `bf`, fp32 offsets into pk.arena); pointers are
  // fixed up after the upload.  dp / dtp: widths padded to multiples of 16 (MMA K and N granularity).
  struct HostOp { XfOp op; size_t w_off, b_off, g_off, lb_off; bool has_w, has_b, has_ln; };
  std::vector<HostOp> prog;
  const int dp = pad16(d), dtp = pad16(d_tok), dff = cfg->d_ff;
  const int hd = d / H;
  const int ffc = std::min(pad16(dff), 128);                  // FFN hidden processed in chunks of <= 128 columns
  bool xf_ok = dp <= 160 && dtp <= 160 && hd % 4 == 0 && (H == 1 || H == 2 || H % 4 == 0) && d % 8 == 0 && d_tok % 8 == 0;
  int max_w_bytes = 0;
  // 16-bit operand format of the transformer: fp16 (11 significant bits; activations saturate at +-65504) unless the
  // config asks for bf16 or a packed value is outside fp16's range.  (tcgen05 kind::f16 needs A and B in the SAME
  // format -- a mixed fp16 x bf16 descriptor is an illegal instruction -- so activations and weights switch together.)
  bool xf_f16 = cfg->tc_format != SF_TC_BF16;
  for (size_t i = 0; i < pk.arena.size() && xf_f16; ++i)
    if (!(std::fabs(pk.arena[i]) < 32768.f)) xf_f16 = false;
  auto image = [&](const LinOff& lin, int n0, int n_cnt, int k0, int k_cnt, int Npad, int Kpad) {
    const size_t off = bf_alloc((size_t)Npad * Kpad);
    const bool half = xf_f16;
    for (int kc = 0; kc < Kpad / 8; ++kc)
      for (int n = 0; n < Npad; ++n)
        for (int e = 0; e < 8; ++e) {
          const int k = kc * 8 + e;
          const float v = (n < n_cnt && k < k_cnt) ? pk.arena[lin.wt + (size_t)(k0 + k) * lin.N + (n0 + n)] : 0.f;
          bf[off + ((size_t)kc * Npad + n) * 8 + e] = half ? f2h(v) : f2bf(v);
        }
    max_w_bytes = std::max(max_w_bytes, Npad * Kpad * 2);
    return off;
  };
  auto bias_pad = [&](const LinOff& lin, int n0, int n_cnt, int Npad) {
    const size_t off = pk.alloc(Npad);
    for (int n = 0; n < n_cnt; ++n) pk.arena[off + n] = pk.arena[lin.b + n0 + n];
    return off;
  };
  auto op_gemm = [&](int a_src, const LinOff& lin, int n0, int n_cnt, int k0, int k_cnt, int Npad, int Kpad, int col,
                     int accumulate, int epi, int act, bool with_bias) {
    HostOp h{};
    h.op.type = XF_GEMM; h.op.a_src = a_src; h.op.K = Kpad; h.op.N = Npad; h.op.tmem_col = col;
    h.op.accumulate = accumulate; h.op.epi = epi; h.op.act = act; h.op.post = XP_NONE;
    h.w_off = image(lin, n0, n_cnt, k0, k_cnt, Npad, Kpad); h.has_w = true;
    h.has_b = with_bias;
    if (with_bias) h.b_off = bias_pad(lin, n0, n_cnt, Npad);
    prog.push_back(h);
    return prog.size() - 1;
  };
  auto set_post = [&](size_t idx, int post, const NormOff* nm, int also_mem) {
    prog[idx].op.post = post; prog[idx].op.also_mem = also_mem;
    if (nm) { prog[idx].has_ln = true; prog[idx].g_off = nm->g; prog[idx].lb_off = nm->b; }
  };
  auto op_init = [&](int mode) {
    HostOp h{};
    h.op.type = XF_INIT; h.op.init_mode = mode;
    prog.push_back(h);
    return prog.size() - 1;
  };
  // attention block: q, k, v into TMEM columns [0,dp) [dp,2dp) [2dp,3dp), attention epilogue -> Hop, out_proj -> stream
  auto attention = [&](const AttnOff& at, int kv_src) {
    // q, k, v are issued in one phase (XfOp::chain = 3): only v's completion is waited for
    prog[op_gemm(XS_AOP, at.qkv, 0, d, 0, d, dp, dp, 0, 0, XE_NONE, 0, true)].op.chain = 3;
    op_gemm(kv_src, at.qkv, d, d, 0, d, dp, dp, dp, 0, XE_NONE, 0, true);
    op_gemm(kv_src, at.qkv, 2 * d, d, 0, d, dp, dp, 2 * dp, 0, XE_ATTN, 0, true);
    return op_gemm(XS_HOP, at.out, 0, d, 0, d, dp, dp, 0, 0, XE_STREAM_ADD, 0, true);
  };
  // FFN: hidden in chunks of ffc columns; ff2 accumulates in TMEM columns [0,dp), ff1 chunk lands at column 256
  auto ffn = [&](const LinOff& f1, const LinOff& f2, int act) {
    size_t last = 0;
    const int n_chunks = (dff + ffc - 1) / ffc;
    for (int c = 0; c < n_chunks; ++c) {
      const int cnt = std::min(ffc, dff - c * ffc), cpad = pad16(cnt);
      op_gemm(XS_AOP, f1, c * ffc, cnt, 0, d, cpad, dp, 256, 0, XE_ACT_H, act, true);
      last = op_gemm(XS_HOP, f2, 0, d, c * ffc, cnt, dp, cpad, 0, c > 0, c + 1 == n_chunks ? XE_STREAM_ADD : XE_NONE, 0,
                     c + 1 == n_chunks);
    }
    return last;
  };
  if (xf_ok) {
    if (v1) {
      set_post(op_init(XI_TOK_PE), XP_COPY_TO_AOP, nullptr, 0);
      for (int i = 0; i < cfg->n_enc_layers; ++i) {
        set_post(attention(eo[i].sa, XS_AOP), XP_LN_INPLACE_TO_AOP, &eo[i].n1, 0);
        set_post(ffn(eo[i].f1, eo[i].f2, 1), XP_LN_INPLACE_TO_AOP, &eo[i].n2, i + 1 == cfg->n_enc_layers);
      }
      set_post(op_init(XI_SHIFT_TOK_PE), XP_COPY_TO_AOP, nullptr, 0);
      for (int i = 0; i < cfg->n_dec_layers; ++i) {
        set_post(attention(dof[i].sa, XS_AOP), XP_LN_INPLACE_TO_AOP, &dof[i].n1, 0);
        set_post(attention(dof[i].ca, XS_MEM), XP_LN_INPLACE_TO_AOP, &dof[i].n2, 0);
        set_post(ffn(dof[i].f1, dof[i].f2, 1), XP_LN_INPLACE_TO_AOP, &dof[i].n3, 0);
      }
      op_gemm(XS_AOP, out_proj, 0, d, 0, d, dp, dp, 0, 0, XE_SCORE, 0, true);
    } else {
      auto embed = [&](const NormOff* first_norm) {
        if (io_proj) {
          op_init(XI_TOK_TO_AOP);
          set_post(op_gemm(XS_AOP, in_proj, 0, d, 0, d_tok, dp, dtp, 0, 0, XE_STREAM_SET_PE, 0, true), XP_LN_TO_AOP, first_norm, 0);
        } else {
          set_post(op_init(XI_TOK_PE), XP_LN_TO_AOP, first_norm, 0);
        }
      };
      embed(&eo[0].n1);
      for (int i = 0; i < cfg->n_enc_layers; ++i) {
        const bool lastl = i + 1 == cfg->n_enc_layers;
        set_post(attention(eo[i].sa, XS_AOP), XP_LN_TO_AOP, &eo[i].n2, 0);
        set_post(ffn(eo[i].f1, eo[i].f2, 2), lastl ? XP_LN_TO_MEM : XP_LN_TO_AOP, lastl ? &enc_norm : &eo[i + 1].n1, 0);
      }
      embed(&dof[0].n1);
      for (int i = 0; i < cfg->n_dec_layers; ++i) {
        const bool lastl = i + 1 == cfg->n_dec_layers;
        set_post(attention(dof[i].sa, XS_AOP), XP_LN_TO_AOP, &dof[i].n2, 0);
        set_post(attention(dof[i].ca, XS_MEM), XP_LN_TO_AOP, &dof[i].n3, 0);
        const size_t f = ffn(dof[i].f1, dof[i].f2, 2);
        if (!lastl) set_post(f, XP_LN_TO_AOP, &dof[i + 1].n1, 0);
        else set_post(f, io_proj ? XP_LN_TO_AOP : XP_LN_SCORE, &dec_norm, 0);
      }
      if (io_proj) op_gemm(XS_AOP, out_proj, 0, d_tok, 0, d, dtp, dp, 0, 0, XE_SCORE, 0, true);
    }
  }

  // ------------------------------------------------------------- upload + pointer fix-up
  sf_model* m = new sf_model();
  m->cfg = *cfg;
  m->device = device;
  m->sm_count = prop.multiProcessorCount;
  m->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  m->arena_bytes = pk.arena.size() * sizeof(float);
  m->tok2 = nullptr;
  m->host_arena = nullptr;
  m->arena_bf16 = nullptr;
  m->xfops_dev = nullptr;
  m->side = nullptr;
  auto fill_tok = [&](const float* A, Tokenizer& T) {
    T.n_blocks = nb;
    T.V = V;
    T.c_in = c0;
    T.pool_tokens = cfg->pool_tokens;
    T.in_scale = A + off_in_scale;
    T.in_shift = A + off_in_shift;
    for (int i = 0; i < nb; ++i) {
      TokBlock& b = T.blk[i];
      b.cin = cfg->channels[i];
      b.cout = cfg->channels[i + 1];
      b.stride = cfg->strides[i];
      b.identity_res = bo[i].identity;
      b.ell_width = bo[i].ellw;
      b.gcn_w = A + bo[i].gw;
      b.gcn_b = A + bo[i].gb;
      b.tcn_w = A + bo[i].tw;
      b.res_w = bo[i].identity ? nullptr : A + bo[i].rw;
      b.out_b = A + bo[i].ob;
      b.ell_val = A + bo[i].ev;
      b.ell_col = reinterpret_cast<const int*>(A + bo[i].ec);
    }
  };
  {
    Tokenizer host_tok;
    fill_tok(pk.arena.data(), host_tok);
    m->tok2 = tok2_create(host_tok, cfg->pool_tokens, !host_only, cfg->tc_format != SF_TC_BF16);
  }
  if (host_only) {
    m->host_arena = (float*)malloc(m->arena_bytes);
    memcpy(m->host_arena, pk.arena.data(), m->arena_bytes);
    m->arena = nullptr;
    fill_tok(m->host_arena, m->tok);
    memset(&m->xf, 0, sizeof(m->xf));
    m->xf.d_tok = d_tok;
    m->xfprog = XfProgram{nullptr, 0, 0, 0, 0, 0, 0};
    *out = m;
    return SF_OK;
  }
  cudaError_t e = cudaMalloc((void**)&m->arena, m->arena_bytes);
  if (e == cudaSuccess) e = cudaMemcpy(m->arena, pk.arena.data(), m->arena_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("uploading %zu bytes of packed weights failed: %s", m->arena_bytes, cudaGetErrorString(e));
    if (m->arena) cudaFree(m->arena);
    tok2_destroy(m->tok2);
    delete m;
    return SF_E_CUDA;
  }
  {
    sf::SideStream* sd = new sf::SideStream();
    if (cudaStreamCreateWithFlags(&sd->st, cudaStreamNonBlocking) == cudaSuccess &&
        cudaEventCreateWithFlags(&sd->fork, cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags(&sd->join, cudaEventDisableTiming) == cudaSuccess) {
      m->side = sd;
    } else {
      cudaGetLastError();
      if (sd->fork) cudaEventDestroy(sd->fork);
      if (sd->st) cudaStreamDestroy(sd->st);
      delete sd;
    }
  }
  m->arena_bf16 = nullptr;
  m->arena_bf16_bytes = bf.size() * sizeof(uint16_t);
  e = cudaMalloc((void**)&m->arena_bf16, m->arena_bf16_bytes + 256);
  if (e == cudaSuccess) e = cudaMemcpy(m->arena_bf16, bf.data(), m->arena_bf16_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("uploading %zu bytes of bf16 operand images failed: %s", m->arena_bf16_bytes, cudaGetErrorString(e));
    cudaFree(m->arena);
    if (m->arena_bf16) cudaFree(m->arena_bf16);
    tok2_destroy(m->tok2);
    delete m;
    return SF_E_CUDA;
  }
  const float* A = m->arena;
  for (int i = 0; i < nb; ++i) {
    BfBlockW& w = m->tokbf.blk[i];
    w.tcn = m->arena_bf16 + bfo[i].tcn;
    w.gcn = bfo[i].has_gcn ? m->arena_bf16 + bfo[i].gcn : nullptr;
    w.res = bfo[i].has_res ? m->arena_bf16 + bfo[i].res : nullptr;
    w.gcn_b = A + bfo[i].gb;
    w.out_b = A + bfo[i].ob;
    w.npad = bfo[i].npad;
    w.kin_pad = bfo[i].kin;
  }
  {
    std::vector<XfOp> ops(prog.size());
    for (size_t i = 0; i < prog.size(); ++i) {
      ops[i] = prog[i].op;
      ops[i].w = prog[i].has_w ? m->arena_bf16 + prog[i].w_off : nullptr;
      ops[i].bias = prog[i].has_b ? A + prog[i].b_off : nullptr;
      ops[i].ln_g = prog[i].has_ln ? A + prog[i].g_off : nullptr;
      ops[i].ln_b = prog[i].has_ln ? A + prog[i].lb_off : nullptr;
    }
    m->xfops_dev = nullptr;
    m->xfprog = XfProgram{nullptr, (int)ops.size(), dp, dtp, xf_ok && !ops.empty() ? 1 : 0, max_w_bytes, xf_f16 ? 1 : 0};
    if (!ops.empty()) {
      e = cudaMalloc((void**)&m->xfops_dev, ops.size() * sizeof(XfOp));
      if (e == cudaSuccess) e = cudaMemcpy(m->xfops_dev, ops.data(), ops.size() * sizeof(XfOp), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) {
        set_error("uploading the transformer program failed: %s", cudaGetErrorString(e));
        cudaFree(m->arena);
        cudaFree(m->arena_bf16);
        tok2_destroy(m->tok2);
        delete m;
        return SF_E_CUDA;
      }
      m->xfprog.ops = m->xfops_dev;
    }
  }
  auto lin = [&](const LinOff& o) { return Linear{A + o.wt, A + o.b, o.K, o.N}; };
  auto nrm = [&](const NormOff& o) { return Norm{A + o.g, A + o.b}; };
  auto att = [&](const AttnOff& o) { return Attn{lin(o.qkv), lin(o.out)}; };

  fill_tok(A, m->tok);
  Transformer& X = m->xf;
  X.variant = cfg->variant;
  X.d_tok = d_tok;
  X.d_model = d;
  X.heads = H;
  X.n_enc = cfg->n_enc_layers;
  X.n_dec = cfg->n_dec_layers;
  X.d_ff = cfg->d_ff;
  X.has_io_proj = io_proj;
  X.pe = A + off_pe;
  X.pe_score = v1 ? A + off_pe_score : nullptr;
  X.in_proj = io_proj ? lin(in_proj) : Linear{nullptr, nullptr, 0, 0};
  X.out_proj = (io_proj || v1) ? lin(out_proj) : Linear{nullptr, nullptr, 0, 0};
  X.enc_norm = v1 ? Norm{nullptr, nullptr} : nrm(enc_norm);
  X.dec_norm = v1 ? Norm{nullptr, nullptr} : nrm(dec_norm);
  for (int i = 0; i < X.n_enc; ++i) X.enc[i] = EncLayer{att(eo[i].sa), lin(eo[i].f1), lin(eo[i].f2), nrm(eo[i].n1), nrm(eo[i].n2)};
  for (int i = 0; i < X.n_dec; ++i)
    X.dec[i] = DecLayer{att(dof[i].sa), att(dof[i].ca), lin(dof[i].f1), lin(dof[i].f2), nrm(dof[i].n1), nrm(dof[i].n2), nrm(dof[i].n3)};
  *out = m;
  return SF_OK;
}

extern "C" void sf_model_destroy(sf_model* m) {
  if (!m) return;
  if (m->device < 0) {
    tok2_destroy(m->tok2);
    free(m->host_arena);
    delete m;
    return;
  }
  DeviceGuard guard;
  guard.enter(m->device);
  tok2_destroy(m->tok2);
  if (m->side) {
    cudaStreamSynchronize(m->side->st);
    cudaEventDestroy(m->side->fork);
    cudaEventDestroy(m->side->join);
    cudaStreamDestroy(m->side->st);
    delete m->side;
  }
  if (m->arena) cudaFree(m->arena);
  if (m->arena_bf16) cudaFree(m->arena_bf16);
  if (m->xfops_dev) cudaFree(m->xfops_dev);
  delete m;
}

int sf::token_len(const sf_model* m, int T) {
  int t = T;
  for (int i = 0; i < m->cfg.n_blocks; ++i) t = (t - 1) / m->cfg.strides[i] + 1;
  if (m->cfg.pool_tokens > 0) t = m->cfg.pool_tokens;
  return t;
}

extern "C" int sf_model_token_shape(const sf_model* m, int32_t T, int32_t* S, int32_t* D) {
  SF_REQUIRE(m && T >= 1, SF_E_INVALID, "sf_model_token_shape: bad argument");
  if (S) *S = sf::token_len(m, T);
  if (D) *D = m->xf.d_tok;
  return SF_OK;
}
