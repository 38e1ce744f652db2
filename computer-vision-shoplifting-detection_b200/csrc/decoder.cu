// GCAE pose decoder: tokens (B, S, latent * V) -> reconstructed poses (B, C_out, seq_len, V), eval mode.
//
// Replaces GCAEDecoder.forward -- shopformer/models/gcae.py:369-478, shopformer_2/models/gcae.py:425-534:
//   h = initial_proj(tokens)                       Linear latent*V -> hidden*V per token
//   h -> (B, hidden, S, V)
//   n_layers x [ConvTranspose2d(k=(u,1), stride (u,1)) | Conv2d 1x1] (+ BatchNorm2d + ReLU except after the last)
//   bilinear resize (align_corners=False) to (seq_len, V) when the stack's length differs from seq_len
// The decoder is NOT on the scoring path (the score only needs tokens and their reconstruction); `Shopformer.forward`
// returns it, so the facade produces it on demand through this kernel pair instead of an ATen composition:
//   dec_proj_kernel   fp32 tiled GEMM (64 x 64 x 16 tiles, 4 x 4 per thread): M = B*S tokens, K = latent*V, N = hidden*V
//   dec_stack_kernel  one window per CTA iteration, activations ping-pong between two shared-memory buffers
//                     ([channel][time][keypoint]), each layer's BN-folded weights staged in shared memory; a transposed
//                     conv with kernel = stride is a 1x1 conv whose weight slice is picked by t' mod u.
// FFMA-bound (2.0 MFLOP per config-A window); both kernels are fp32 end to end.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "sf_internal.h"

namespace sf {
namespace {

constexpr int kDecMaxLayers = 8;
constexpr double kBnEpsDec = 1e-5;

struct PoseDecLayer {
  int cin, cout, up, relu;
  const float* w;       // [up][cin][cout4] BN-folded, cout padded to a multiple of 4
  const float* b;       // [cout4]
  int cout4;
};
struct PoseDecPlan {
  int V, S, hidden, c_out, seq_len, n_layers;
  int len[kDecMaxLayers + 1];        // temporal length of the input of layer i; len[n_layers] = the stack's output length
  PoseDecLayer layer[kDecMaxLayers];
  uint32_t off_buf[2], off_w, smem_bytes;
  float t_scale;                     // stack length / seq_len (bilinear source scale), as PyTorch computes it
};

// ---- h[m][n] = sum_k a[m][k] * wt[k][n] + bias[n]
constexpr int kTM = 64, kTN = 64, kTK = 16;
__global__ void __launch_bounds__(256)
dec_proj_kernel(const float* __restrict__ a, const float* __restrict__ wt, const float* __restrict__ bias, float* __restrict__ h,
                int64_t M, int K, int N) {
  __shared__ float sa[kTK][kTM + 4];
  __shared__ float sw[kTK][kTN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * kTM;
  const int n0 = blockIdx.x * kTN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += kTK) {
    for (int i = threadIdx.x; i < kTM * kTK; i += 256) {            // a tile: consecutive threads walk k (contiguous in a row)
      const int mm = i / kTK, kk = i - mm * kTK;
      const int64_t m = m0 + mm;
      sa[kk][mm] = (m < M && k0 + kk < K) ? __ldg(a + m * K + k0 + kk) : 0.f;
    }
    for (int i = threadIdx.x; i < kTK * kTN; i += 256) {
      const int kk = i / kTN, nn = i - kk * kTN;
      sw[kk][nn] = (k0 + kk < K && n0 + nn < N) ? __ldg(wt + (size_t)(k0 + kk) * N + n0 + nn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&sa[kk][ty * 4]);
      const float4 wv = *reinterpret_cast<const float4*>(&sw[kk][tx * 4]);
      const float aa[4] = {av.x, av.y, av.z, av.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], ww[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) h[m * N + n] = acc[i][j] + __ldg(bias + n);
    }
  }
}

// ---- conv stack + resize, one window per CTA iteration
__global__ void __launch_bounds__(256)
dec_stack_kernel(const __grid_constant__ PoseDecPlan pl, const float* __restrict__ h, float* __restrict__ out, int64_t B) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* buf[2] = {reinterpret_cast<float*>(smem + pl.off_buf[0]), reinterpret_cast<float*>(smem + pl.off_buf[1])};
  float* sw = reinterpret_cast<float*>(smem + pl.off_w);
  const int V = pl.V, S = pl.S;
  for (int64_t w = blockIdx.x; w < B; w += gridDim.x) {
    // h (S, hidden * V) of this window -> buf[0] as [c][s][v]  (feature index of the Linear output = c * V + v)
    const float* hw = h + (size_t)w * S * pl.hidden * V;
    const int hv = pl.hidden * V;
    for (int i = threadIdx.x; i < S * hv; i += 256) {
      const int s = i / hv, f = i - s * hv;
      const int c = f / V, v = f - c * V;
      buf[0][(c * S + s) * V + v] = __ldg(hw + i);
    }
    for (int li = 0; li < pl.n_layers; ++li) {
      const PoseDecLayer L = pl.layer[li];
      const int Lin = pl.len[li], Lout = pl.len[li + 1];
      const float* in = buf[li & 1];
      float* o = buf[(li + 1) & 1];
      __syncthreads();                                              // the previous layer's output is complete; sw is free
      const int nw = L.up * L.cin * L.cout4;
      for (int i = threadIdx.x * 4; i < nw; i += 256 * 4) *reinterpret_cast<float4*>(sw + i) = __ldg(reinterpret_cast<const float4*>(L.w + i));
      for (int i = threadIdx.x; i < L.cout4; i += 256) sw[nw + i] = __ldg(L.b + i);
      __syncthreads();
      const int P = Lout * V, groups = L.cout4 >> 2;
      for (int item = threadIdx.x; item < P * groups; item += 256) {
        const int og = item / P, pos = item - og * P;
        const int tp = pos / V, v = pos - tp * V;
        const int t = tp / L.up, j = tp - t * L.up;
        const float4 bb = *reinterpret_cast<const float4*>(sw + nw + og * 4);
        float a0 = bb.x, a1 = bb.y, a2 = bb.z, a3 = bb.w;
        const float* xin = in + t * V + v;
        const float* wj = sw + (size_t)j * L.cin * L.cout4 + og * 4;
#pragma unroll 4
        for (int c = 0; c < L.cin; ++c) {
          const float x = xin[c * Lin * V];
          const float4 ww = *reinterpret_cast<const float4*>(wj + c * L.cout4);
          a0 = fmaf(x, ww.x, a0); a1 = fmaf(x, ww.y, a1); a2 = fmaf(x, ww.z, a2); a3 = fmaf(x, ww.w, a3);
        }
        if (L.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
        const int oc = og * 4;
        if (oc + 0 < L.cout) o[((oc + 0) * Lout + tp) * V + v] = a0;
        if (oc + 1 < L.cout) o[((oc + 1) * Lout + tp) * V + v] = a1;
        if (oc + 2 < L.cout) o[((oc + 2) * Lout + tp) * V + v] = a2;
        if (oc + 3 < L.cout) o[((oc + 3) * Lout + tp) * V + v] = a3;
      }
    }
    __syncthreads();
    // ---- (C_out, Lf, V) -> out (C_out, seq_len, V); bilinear along time when the lengths differ (the keypoint axis keeps its
    // size, so its interpolation is the identity), align_corners=False: src = scale * (dst + 0.5) - 0.5 clamped at 0
    const float* fin = buf[pl.n_layers & 1];
    const int Lf = pl.len[pl.n_layers], T = pl.seq_len;
    float* ow = out + (size_t)w * pl.c_out * T * V;
    for (int i = threadIdx.x; i < pl.c_out * T * V; i += 256) {
      const int c = i / (T * V), r = i - c * T * V;
      const int t = r / V, v = r - t * V;
      float val;
      if (Lf == T) {
        val = fin[(c * Lf + t) * V + v];
      } else {
        float src = pl.t_scale * ((float)t + 0.5f) - 0.5f;
        if (src < 0.f) src = 0.f;
        int i0 = (int)src;
        if (i0 > Lf - 1) i0 = Lf - 1;
        const int i1 = i0 + (i0 < Lf - 1 ? 1 : 0);
        const float l1 = src - (float)i0, l0 = 1.f - l1;
        val = l0 * fin[(c * Lf + i0) * V + v] + l1 * fin[(c * Lf + i1) * V + v];
      }
      __stcs(ow + i, val);
    }
    __syncthreads();                                                // buf[0] is rewritten by the next window
  }
}

}  // namespace
}  // namespace sf

using namespace sf;

struct sf_decoder {
  int device, sm_count, max_smem;
  int latent, hidden, c_out, V, seq_len, n_layers;
  int up[kDecMaxLayers];
  float* arena;                 // one allocation: proj Wt [K][N], proj bias [N], per layer w / b
  size_t off_wt, off_pb;
  size_t off_w[kDecMaxLayers], off_b[kDecMaxLayers];
  int cout[kDecMaxLayers], cout4[kDecMaxLayers], relu[kDecMaxLayers];
};

extern "C" int sf_decoder_create(int32_t latent_channels, int32_t hidden_channels, int32_t out_channels, int32_t num_keypoints,
                                 int32_t seq_len, int32_t n_layers, const int32_t* upsample, int32_t n_tensors,
                                 const char* const* names, const float* const* data_host, const int64_t* numel, int32_t device,
                                 sf_decoder** out) {
  SF_REQUIRE(out && upsample && names && data_host && numel, SF_E_INVALID, "sf_decoder_create: null argument");
  SF_REQUIRE(latent_channels >= 1 && hidden_channels >= 1 && out_channels >= 1 && num_keypoints >= 1 && seq_len >= 1 && n_layers >= 1 &&
                 n_layers <= kDecMaxLayers,
             SF_E_INVALID, "sf_decoder_create: bad shape");
  int rc = sf_device_count();
  if (rc < 0) return rc;
  SF_REQUIRE(device >= 0 && device < rc, SF_E_INVALID, "sf_decoder_create: device %d out of range", device);
  std::map<std::string, std::pair<const float*, int64_t>> sd;
  for (int i = 0; i < n_tensors; ++i) sd[names[i]] = {data_host[i], numel[i]};
  auto get = [&](const std::string& k, int64_t want, const float** p) -> int {
    auto it = sd.find(k);
    SF_REQUIRE(it != sd.end(), SF_E_MISSING, "decoder: missing state-dict key %s", k.c_str());
    SF_REQUIRE(it->second.second == want, SF_E_SHAPE, "decoder: %s has %lld elements, expected %lld", k.c_str(), (long long)it->second.second,
               (long long)want);
    *p = it->second.first;
    return SF_OK;
  };
  const int V = num_keypoints, H = hidden_channels, K = latent_channels * V, N = H * V;
  std::vector<float> host;
  auto alloc = [&](size_t n) { size_t o = (host.size() + 63) & ~size_t(63); host.resize(o + n, 0.f); return o; };
  sf_decoder* d = new sf_decoder();
  memset(d, 0, sizeof(*d));
  d->device = device;
  d->latent = latent_channels; d->hidden = H; d->c_out = out_channels; d->V = V; d->seq_len = seq_len; d->n_layers = n_layers;
  auto fail = [&](int code) { delete d; return code; };
  const float *pw = nullptr, *pb = nullptr;
  if ((rc = get("initial_proj.weight", (int64_t)N * K, &pw)) || (rc = get("initial_proj.bias", N, &pb))) return fail(rc);
  d->off_wt = alloc((size_t)K * N);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) host[d->off_wt + (size_t)k * N + n] = pw[(size_t)n * K + k];
  d->off_pb = alloc(N);
  memcpy(&host[d->off_pb], pb, sizeof(float) * N);
  int idx = 0;                                       // position inside the reference's nn.Sequential
  for (int i = 0; i < n_layers; ++i) {
    const int u = upsample[i];
    SF_REQUIRE(u >= 1 && u <= 8, SF_E_INVALID, "decoder: upsample factor %d", u);
    const bool last = i + 1 == n_layers;
    const int co = last ? out_channels : H, co4 = (co + 3) & ~3;
    d->up[i] = u; d->cout[i] = co; d->cout4[i] = co4; d->relu[i] = last ? 0 : 1;
    const std::string p = "layers." + std::to_string(idx) + ".";
    const float *w = nullptr, *b = nullptr;
    if ((rc = get(p + "weight", (int64_t)H * co * u, &w)) || (rc = get(p + "bias", co, &b))) return fail(rc);
    std::vector<double> scale(co, 1.0), shift(co, 0.0);
    if (!last) {
      const std::string q = "layers." + std::to_string(idx + 1) + ".";
      const float *g = nullptr, *be = nullptr, *rm = nullptr, *rv = nullptr;
      if ((rc = get(q + "weight", co, &g)) || (rc = get(q + "bias", co, &be)) || (rc = get(q + "running_mean", co, &rm)) ||
          (rc = get(q + "running_var", co, &rv)))
        return fail(rc);
      for (int o = 0; o < co; ++o) {
        scale[o] = (double)g[o] / std::sqrt((double)rv[o] + kBnEpsDec);
        shift[o] = (double)be[o] - (double)rm[o] * scale[o];
      }
    }
    d->off_w[i] = alloc((size_t)u * H * co4);
    d->off_b[i] = alloc(co4);
    for (int j = 0; j < u; ++j)
      for (int c = 0; c < H; ++c)
        for (int o = 0; o < co; ++o) {
          // ConvTranspose2d weight (in, out, u, 1); Conv2d 1x1 weight (out, in, 1, 1)
          const double wv = u > 1 ? (double)w[((size_t)c * co + o) * u + j] : (double)w[(size_t)o * H + c];
          host[d->off_w[i] + ((size_t)j * H + c) * co4 + o] = (float)(wv * scale[o]);
        }
    for (int o = 0; o < co; ++o) host[d->off_b[i] + o] = (float)((double)b[o] * scale[o] + shift[o]);
    idx += last ? 1 : 4;                             // conv, BatchNorm2d, ReLU, Dropout
  }
  DeviceGuard guard;
  cudaError_t e = guard.enter(device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&d->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d->arena, host.size() * sizeof(float));
  if (e == cudaSuccess) e = cudaMemcpy(d->arena, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    set_error("sf_decoder_create: %s", cudaGetErrorString(e));
    if (d->arena) cudaFree(d->arena);
    return fail(SF_E_CUDA);
  }
  *out = d;
  return SF_OK;
}

extern "C" void sf_decoder_destroy(sf_decoder* d) {
  if (!d) return;
  DeviceGuard guard;
  guard.enter(d->device);
  if (d->arena) cudaFree(d->arena);
  delete d;
}

static bool dec_plan(const sf_decoder* d, int S, PoseDecPlan* pl) {
  PoseDecPlan& p = *pl;
  memset(&p, 0, sizeof(p));
  p.V = d->V; p.S = S; p.hidden = d->hidden; p.c_out = d->c_out; p.seq_len = d->seq_len; p.n_layers = d->n_layers;
  p.len[0] = S;
  size_t sz[2] = {0, 0}, wmax = 0;
  for (int i = 0; i < d->n_layers; ++i) {
    p.len[i + 1] = p.len[i] * d->up[i];
    PoseDecLayer& L = p.layer[i];
    L.cin = d->hidden; L.cout = d->cout[i]; L.cout4 = d->cout4[i]; L.up = d->up[i]; L.relu = d->relu[i];
    L.w = d->arena + d->off_w[i];
    L.b = d->arena + d->off_b[i];
    sz[i & 1] = std::max(sz[i & 1], (size_t)d->hidden * p.len[i] * d->V);
    wmax = std::max(wmax, (size_t)L.up * L.cin * L.cout4 + L.cout4);
  }
  sz[d->n_layers & 1] = std::max(sz[d->n_layers & 1], (size_t)d->c_out * p.len[d->n_layers] * d->V);
  auto up16 = [](size_t x) { return (x + 15) & ~size_t(15); };
  uint32_t off = 0;
  p.off_buf[0] = off; off += (uint32_t)up16(sz[0] * 4);
  p.off_buf[1] = off; off += (uint32_t)up16(sz[1] * 4);
  p.off_w = off; off += (uint32_t)up16(wmax * 4);
  p.smem_bytes = off;
  p.t_scale = (float)p.len[d->n_layers] / (float)d->seq_len;
  return (int)off <= d->max_smem;
}

extern "C" int64_t sf_decoder_workspace_bytes(const sf_decoder* d, int64_t B, int32_t S) {
  if (!d || B < 0 || S < 1) return SF_E_INVALID;
  return (B * (int64_t)S * d->hidden * d->V * (int64_t)sizeof(float) + 255) & ~int64_t(255);
}

extern "C" int sf_decode_poses(const sf_decoder* d, const float* tokens_dev, int64_t B, int32_t S, float* poses_dev,
                               void* workspace_dev, int64_t workspace_bytes, void* stream) {
  SF_REQUIRE(d && B >= 0 && S >= 1 && S <= 100, SF_E_INVALID, "sf_decode_poses: bad argument");
  if (B == 0) return SF_OK;
  SF_REQUIRE(tokens_dev && poses_dev, SF_E_INVALID, "sf_decode_poses: null buffer");
  SF_REQUIRE(workspace_dev && workspace_bytes >= sf_decoder_workspace_bytes(d, B, S), SF_E_INVALID,
             "sf_decode_poses: workspace of %lld bytes needed, got %lld", (long long)sf_decoder_workspace_bytes(d, B, S),
             (long long)workspace_bytes);
  PoseDecPlan pl;
  SF_REQUIRE(dec_plan(d, S, &pl), SF_E_UNSUPPORTED, "decoder activations (%u bytes) exceed shared memory for S=%d", pl.smem_bytes, S);
  DeviceGuard guard;
  SF_CUDA_OK(guard.enter(d->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int K = d->latent * d->V, N = d->hidden * d->V;
  const int64_t M = B * S;
  float* h = (float*)workspace_dev;
  // the GEMM's grid.y is limited to 65,535 row tiles: walk larger batches in slabs
  const int64_t slab = (int64_t)65535 * kTM;
  for (int64_t m0 = 0; m0 < M; m0 += slab) {
    const int64_t mm = std::min(slab, M - m0);
    dim3 grid((unsigned)((N + kTN - 1) / kTN), (unsigned)((mm + kTM - 1) / kTM));
    dec_proj_kernel<<<grid, 256, 0, st>>>(tokens_dev + m0 * K, d->arena + d->off_wt, d->arena + d->off_pb, h + m0 * N, mm, K, N);
  }
  SF_CUDA_OK(cudaGetLastError());
  SF_CUDA_OK(cudaFuncSetAttribute(dec_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes));
  int occ = 1;
  SF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dec_stack_kernel, 256, pl.smem_bytes));
  const int grid = (int)std::min<int64_t>(B, (int64_t)d->sm_count * std::max(occ, 1));
  dec_stack_kernel<<<grid, 256, pl.smem_bytes, st>>>(pl, h, poses_dev, B);
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}
