// Pose-track windowing + centre/scale normalisation on the GPU (HBM-bound).
//
// Replaces the pure-Python hot loop of the reference's dataset construction:
//   shopformer/data/poselift_dataset.py:297-329 (sliding windows, continuity, majority label),
//   :331-365 (keypoint gather), :367-388 (normalise), :393-400 ((T,V,C)->(C,T,V));
//   shopformer_2/data/poselift_dataset.py:57-91 (neck), :451-589 (same pipeline + frame_indices).
//
// Integer results (which windows exist, order, frame numbers, labels) are bit-exact.
//
// Pipeline (all on `stream`, no host sync unless the caller asks for the count):
//   K_flag    one thread per candidate start position: continuity test over T-1 gaps,
//             majority label, per-1024-candidate block sums
//   K_scan    single CTA exclusive scan of the block sums
//   K_compact per-block scan + scatter of (track, start, label) to the compacted order
//   K_gather  one warp per window: the T*K*3 floats of a window are CONTIGUOUS in the packed
//             track array, so they are staged into shared memory with 16-byte loads
//             (scalar head/tail for the 204-byte frame pitch); ONE pass over the staged keypoints
//             collects count / sum / min / max of the valid ones (centre = mean, scale = largest
//             |coord - centre| = max(max - centre, centre - min)), a second pass writes the
//             normalised planes into a shared-memory slab in output order, and the slab goes to
//             HBM with 16-byte streaming stores.  No division or modulo per element: every lane
//             walks (t, v) incrementally.
#include <algorithm>
#include <cstring>
#include <vector>

#include "sf_internal.h"

namespace sf {
namespace {

constexpr int kScanBlock = 1024;

struct WinCtx {
  const float* kp;
  const int32_t* frame_no;
  const int64_t* track_off;     // device copies
  const int64_t* cand_off;
  const int32_t* track_video;
  const int64_t* gt_off;
  const uint8_t* gt;
  int n_tracks, K, T, stride, max_gap, V, normalize;
  int KC;                       // floats per source keypoint: 3 (x, y, conf as PoseLift stores them) or 2 (x, y)
  int add_neck, conf;           // synthetic 18th keypoint (variant 2); third output plane = raw confidence (variant 1 include_confidence)
  int64_t n_cand;
};

__device__ __forceinline__ int find_track(const int64_t* __restrict__ cand_off, int n_tracks, int64_t c) {
  int lo = 0, hi = n_tracks;             // cand_off[lo] <= c < cand_off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cand_off + mid) <= c) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kScanBlock)
k_flag(const WinCtx cx, uint8_t* __restrict__ flag, uint8_t* __restrict__ label, int32_t* __restrict__ block_sum) {
  __shared__ int warp_sums[kScanBlock / 32];
  const int64_t c = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  int ok = 0;
  if (c < cx.n_cand) {
    const int tr = find_track(cx.cand_off, cx.n_tracks, c);
    const int64_t f0 = __ldg(cx.track_off + tr) + (c - __ldg(cx.cand_off + tr)) * cx.stride;
    ok = 1;
    int prev = __ldg(cx.frame_no + f0);
    int votes = 0;
    const bool has_gt = cx.gt != nullptr;
    int64_t g0 = 0;
    int glen = 0;
    if (has_gt) {
      const int vid = __ldg(cx.track_video + tr);
      g0 = __ldg(cx.gt_off + vid);
      glen = (int)(__ldg(cx.gt_off + vid + 1) - g0);
    }
    if (has_gt && glen > 0) votes += cx.gt[g0 + min(prev, glen - 1)];
    for (int t = 1; t < cx.T; ++t) {
      const int cur = __ldg(cx.frame_no + f0 + t);
      if (cur - prev > cx.max_gap) ok = 0;
      if (has_gt && glen > 0) votes += cx.gt[g0 + min(cur, glen - 1)];
      prev = cur;
    }
    flag[c] = (uint8_t)ok;
    label[c] = (uint8_t)(votes > cx.T / 2 ? 1 : 0);
  }
  // block sum of flags
  int s = ok;
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = warp_sums[threadIdx.x];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) block_sum[blockIdx.x] = v;
  }
}

// exclusive scan of n block sums into int64 offsets; writes the grand total to *total
__global__ void __launch_bounds__(1024) k_scan(const int32_t* __restrict__ block_sum, int64_t* __restrict__ block_off,
                                               int n, int64_t* __restrict__ total) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + threadIdx.x;
    const int64_t v = i < n ? block_sum[i] : 0;
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_tot[warp] = x;
    __syncthreads();
    if (warp == 0) {
      int64_t w = warp_tot[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o) w += y;
      }
      warp_tot[lane] = w;            // inclusive over warps
    }
    __syncthreads();
    const int64_t before = carry + (warp ? warp_tot[warp - 1] : 0) + (x - v);
    if (i < n) block_off[i] = before;
    __syncthreads();
    if (threadIdx.x == 0) carry += warp_tot[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kScanBlock)
k_compact(const WinCtx cx, const uint8_t* __restrict__ flag, const uint8_t* __restrict__ label,
          const int64_t* __restrict__ block_off, int32_t* __restrict__ labels_out, int32_t* __restrict__ win_track,
          int32_t* __restrict__ win_start) {
  __shared__ int warp_tot[kScanBlock / 32];
  const int64_t c = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = (c < cx.n_cand) ? flag[c] : 0;
  int x = f;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  if (f) {
    const int64_t pos = block_off[blockIdx.x] + (warp ? warp_tot[warp - 1] : 0) + (x - 1);
    const int tr = find_track(cx.cand_off, cx.n_tracks, c);
    labels_out[pos] = label[c];
    win_track[pos] = tr;
    win_start[pos] = (int)((c - __ldg(cx.cand_off + tr)) * cx.stride);
  }
}

constexpr int kGatherWarps = 8;

// shared-memory floats per warp: staged raw keypoints (+ phase slack), neck, output slab
__host__ __device__ inline int gather_raw_floats(int T, int K, int KC) { return (T * K * KC + 7) & ~3; }
__host__ __device__ inline int gather_neck_floats(int T) { return (2 * T + 3) & ~3; }
__host__ __device__ inline int gather_out_floats(int T, int V, int C) { return (C * T * V + 3) & ~3; }
inline int gather_per_warp(int T, int K, int KC, int V, int C) { return gather_raw_floats(T, K, KC) + gather_neck_floats(T) + gather_out_floats(T, V, C); }

// One warp per window.  Windows [w_begin, min(count, w_begin + w_cap)) are written to poses[0 .. ) (a pass of a larger
// sweep gathers its own range into a pass-sized buffer).
// FAST: the detection has exactly the model's keypoints (K == V), no synthetic neck and no confidence plane -- the usual
// COCO-17 case.  Output element i = (t, v) then sits at raw[i * KC], so both passes are flat loops without (t, v)
// bookkeeping or per-element branches, and the normalised planes go straight to HBM with coalesced streaming stores
// (no output slab: 4.9 KB of shared memory per warp instead of 8.4 KB, 5 resident blocks per SM instead of 3).
template <bool FAST>
__global__ void __launch_bounds__(kGatherWarps * 32)
k_gather(const WinCtx cx, const int64_t* __restrict__ n_windows, int64_t n_fixed, const int32_t* __restrict__ win_track,
         const int32_t* __restrict__ win_start, float* __restrict__ poses, int32_t* __restrict__ frame_idx, int per_warp,
         int64_t w_begin, int64_t w_cap) {
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* slab = smem + (size_t)warp * per_warp;
  const int T = cx.T, K = cx.K, V = cx.V, KC = cx.KC;
  const int n_raw = T * K * KC;
  const int C = cx.conf ? 3 : 2;
  float* neck = slab + gather_raw_floats(T, K, KC); // [T][2], only used with add_neck
  float* outs = neck + gather_neck_floats(T);       // [C][T][V]: the window in output order
  const bool add_neck = cx.add_neck != 0;
  const int Vsrc = add_neck ? min(V - 1, K) : min(V, K);      // keypoints taken from the detection itself
  const int n_el = T * V;
  // this lane's walk over (t, v): element i = lane, lane + 32, ...  (one division per kernel, none per element)
  const int t_first = lane / V, v_first = lane - t_first * V, dt = 32 / V, dv = 32 - dt * V;
  int64_t nw = n_windows ? *n_windows : n_fixed;            // pre-cut windows: the count is known on the host
  if (nw > w_begin + w_cap) nw = w_begin + w_cap;
  const int64_t warps_total = (int64_t)gridDim.x * kGatherWarps;
  for (int64_t w = w_begin + (int64_t)blockIdx.x * kGatherWarps + warp; w < nw; w += warps_total) {
    // pre-cut mode (win_track == nullptr): window w is frames [w*T, (w+1)*T)
    const int64_t f0 = win_track ? __ldg(cx.track_off + __ldg(win_track + w)) + __ldg(win_start + w) : w * (int64_t)cx.T;
    // ---- stage the contiguous chunk: scalar head, 16-byte body, scalar tail
    const float* src = cx.kp + (size_t)f0 * K * KC;
    const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
    float* raw = slab + mis;                        // same 16-byte phase in smem as in HBM
    const int head = mis ? min(4 - mis, n_raw) : 0;
    if (lane < head) raw[lane] = __ldg(src + lane);
    const int body4 = (n_raw - head) >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src + head);
    float4* d4 = reinterpret_cast<float4*>(raw + head);
    for (int i = lane; i < body4; i += 32) d4[i] = __ldg(s4 + i);
    for (int i = head + 4 * body4 + lane; i < n_raw; i += 32) raw[i] = __ldg(src + i);
    if (frame_idx && cx.frame_no)
      for (int t = lane; t < T; t += 32) frame_idx[(w - w_begin) * T + t] = __ldg(cx.frame_no + f0 + t);
    __syncwarp();
    if (FAST) {
      float cxm = 0.f, cym = 0.f, inv = 1.f;
      if (cx.normalize) {
        float sx = 0.f, sy = 0.f, lox = 3.0e38f, hix = -3.0e38f, loy = 3.0e38f, hiy = -3.0e38f;
        int cnt = 0;
        for (int i = lane; i < n_el; i += 32) {
          const float x = raw[i * KC], y = raw[i * KC + 1];
          if (x != 0.f || y != 0.f) {
            sx += x; sy += y; ++cnt;
            lox = fminf(lox, x); hix = fmaxf(hix, x); loy = fminf(loy, y); hiy = fmaxf(hiy, y);
          }
        }
        double dsx = (double)sx, dsy = (double)sy;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          dsx += __shfl_xor_sync(0xffffffffu, dsx, o);
          dsy += __shfl_xor_sync(0xffffffffu, dsy, o);
          cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
          lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o));
          hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
          loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o));
          hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
        }
        if (cnt > 0) {
          cxm = (float)(dsx / (double)cnt);
          cym = (float)(dsy / (double)cnt);
          const float mx = fmaxf(fmaxf(hix - cxm, cxm - lox), fmaxf(hiy - cym, cym - loy));
          inv = 1.f / (mx + 1e-6f);
        }
      }
      float* dst = poses + (size_t)(w - w_begin) * 2 * n_el;
      for (int i = lane; i < n_el; i += 32) {
        float x = raw[i * KC], y = raw[i * KC + 1];
        if (cx.normalize) {
          x = (x - cxm) * inv;
          y = (y - cym) * inv;
          if (!(fabsf(x) <= 3.0e38f)) x = 0.f;        // nan_to_num(nan=0, posinf=0, neginf=0)
          if (!(fabsf(y) <= 3.0e38f)) y = 0.f;
        }
        __stcs(dst + i, x);
        __stcs(dst + n_el + i, y);
      }
      __syncwarp();
      continue;
    }
    if (add_neck) {
      // add_neck_keypoint: midpoint of shoulders 5/6; np.allclose(.,0) == |x|,|y| <= 1e-8
      for (int t = lane; t < T; t += 32) {
        const float lx = raw[(t * K + 5) * KC], ly = raw[(t * K + 5) * KC + 1];
        const float rx = raw[(t * K + 6) * KC], ry = raw[(t * K + 6) * KC + 1];
        const bool l0 = fabsf(lx) <= 1e-8f && fabsf(ly) <= 1e-8f;
        const bool r0 = fabsf(rx) <= 1e-8f && fabsf(ry) <= 1e-8f;
        float nx = (lx + rx) * 0.5f, ny = (ly + ry) * 0.5f;
        if (l0 && r0) { nx = 0.f; ny = 0.f; }
        else if (l0) { nx = rx; ny = ry; }
        else if (r0) { nx = lx; ny = ly; }
        neck[2 * t] = nx;
        neck[2 * t + 1] = ny;
      }
      __syncwarp();
    }
    auto load_xy = [&](int t, int v, float& x, float& y) {
      if (v < Vsrc) {
        const float* p = raw + (t * K + v) * KC;
        x = p[0];
        y = p[1];
      } else if (add_neck && v == V - 1) {
        x = neck[2 * t];
        y = neck[2 * t + 1];
      } else {
        x = y = 0.f;                                // zero padding when the detection has fewer keypoints
      }
    };
    float cxm = 0.f, cym = 0.f, inv = 1.f;
    if (cx.normalize) {
      // one pass: count, sums (fp32 per lane, fp64 across lanes), min / max of the valid keypoints
      float sx = 0.f, sy = 0.f, lox = 3.0e38f, hix = -3.0e38f, loy = 3.0e38f, hiy = -3.0e38f;
      int cnt = 0;
      for (int t = t_first, v = v_first; t < T;) {
        float x, y;
        load_xy(t, v, x, y);
        if (x != 0.f || y != 0.f) {
          sx += x; sy += y; ++cnt;
          lox = fminf(lox, x); hix = fmaxf(hix, x); loy = fminf(loy, y); hiy = fmaxf(hiy, y);
        }
        v += dv; t += dt;
        if (v >= V) { v -= V; ++t; }
      }
      double dsx = (double)sx, dsy = (double)sy;
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        dsx += __shfl_xor_sync(0xffffffffu, dsx, o);
        dsy += __shfl_xor_sync(0xffffffffu, dsy, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o));
        hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
        loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o));
        hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
      }
      if (cnt > 0) {
        cxm = (float)(dsx / (double)cnt);
        cym = (float)(dsy / (double)cnt);
        // max |coord - centre| over the valid keypoints = the larger distance of the centre to the per-axis extremes
        // (fp32 subtraction is monotone, so this is exactly the maximum of the per-element fp32 differences)
        const float mx = fmaxf(fmaxf(hix - cxm, cxm - lox), fmaxf(hiy - cym, cym - loy));
        inv = 1.f / (mx + 1e-6f);
      }
    }
    // ---- second pass: the window in output order ((C, T, V) planes) into the shared-memory slab
    for (int t = t_first, v = v_first, i = lane; t < T; i += 32) {
      float x, y;
      load_xy(t, v, x, y);
      if (cx.normalize) {
        x = (x - cxm) * inv;
        y = (y - cym) * inv;
        if (!(fabsf(x) <= 3.0e38f)) x = 0.f;        // nan_to_num(nan=0, posinf=0, neginf=0)
        if (!(fabsf(y) <= 3.0e38f)) y = 0.f;
      }
      outs[i] = x;
      outs[n_el + i] = y;
      if (C == 3) outs[2 * n_el + i] = v < Vsrc ? raw[(t * K + v) * KC + 2] : 0.f;
      v += dv; t += dt;
      if (v >= V) { v -= V; ++t; }
    }
    __syncwarp();
    // ---- slab -> HBM
    float* dst = poses + (size_t)(w - w_begin) * C * n_el;
    const int n_out = C * n_el;
    if (((n_out & 3) == 0)) {
      const float4* o4 = reinterpret_cast<const float4*>(outs);
      for (int q = lane; q < (n_out >> 2); q += 32) __stcs(reinterpret_cast<float4*>(dst) + q, o4[q]);
    } else {
      for (int i = lane; i < n_out; i += 32) dst[i] = outs[i];
    }
    __syncwarp();
  }
}

struct WsLayout {
  size_t track_off, cand_off, track_video, gt_off, flag, label, block_sum, block_off, total;
  int64_t n_cand;
  int n_blocks;
};

int64_t candidates(const sf_tracks* tr, const sf_window_params* p, std::vector<int64_t>* cand_off) {
  int64_t n = 0;
  if (cand_off) cand_off->assign(tr->n_tracks + 1, 0);
  for (int i = 0; i < tr->n_tracks; ++i) {
    const int64_t len = tr->track_offsets_host[i + 1] - tr->track_offsets_host[i];
    if (len >= p->seq_len) n += (len - p->seq_len) / p->stride + 1;
    if (cand_off) (*cand_off)[i + 1] = n;
  }
  return n;
}

WsLayout layout(const sf_tracks* tr, const sf_window_params* p) {
  WsLayout L{};
  L.n_cand = candidates(tr, p, nullptr);
  L.n_blocks = (int)((L.n_cand + kScanBlock - 1) / kScanBlock);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += (bytes + 255) & ~size_t(255);
    return o;
  };
  L.track_off = take(sizeof(int64_t) * (tr->n_tracks + 1));
  L.cand_off = take(sizeof(int64_t) * (tr->n_tracks + 1));
  L.track_video = take(sizeof(int32_t) * std::max(tr->n_tracks, 1));
  L.gt_off = take(sizeof(int64_t) * (tr->n_videos + 1));
  L.flag = take((size_t)std::max<int64_t>(L.n_cand, 1));
  L.label = take((size_t)std::max<int64_t>(L.n_cand, 1));
  L.block_sum = take(sizeof(int32_t) * std::max(L.n_blocks, 1));
  L.block_off = take(sizeof(int64_t) * std::max(L.n_blocks, 1));
  L.total = off;
  return L;
}

// Launch geometry of k_gather on the CURRENT device.  The opt-in shared-memory size is a per-device function attribute:
// it is (re)applied whenever the device or the size changes, never cached across devices.
bool gather_fast(const WinCtx& cx) { return cx.K == cx.V && !cx.add_neck && !cx.conf; }

int gather_launch_config(bool fast, size_t smem, int64_t blocks_wanted, int* grid) {
  int dev = 0, max_smem = 0, sms = 0, occ = 1;
  SF_CUDA_OK(cudaGetDevice(&dev));
  SF_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  SF_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SF_REQUIRE(smem <= (size_t)max_smem, SF_E_UNSUPPORTED, "windowing: the window needs %zu bytes of shared memory per block", smem);
  if (fast) {
    SF_CUDA_OK(cudaFuncSetAttribute(k_gather<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gather<true>, kGatherWarps * 32, smem));
  } else {
    SF_CUDA_OK(cudaFuncSetAttribute(k_gather<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SF_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gather<false>, kGatherWarps * 32, smem));
  }
  *grid = (int)std::max<int64_t>(1, std::min<int64_t>(blocks_wanted, (int64_t)sms * std::max(occ, 1)));
  return SF_OK;
}

int check(const sf_tracks* tr, const sf_window_params* p) {
  SF_REQUIRE(tr && p, SF_E_INVALID, "windowing: null argument");
  SF_REQUIRE(p->seq_len >= 1 && p->stride >= 1 && p->max_gap >= 0, SF_E_INVALID, "windowing: bad seq_len/stride/max_gap");
  SF_REQUIRE(tr->n_tracks >= 0 && tr->kp_per_frame >= 1, SF_E_INVALID, "windowing: bad track table");
  SF_REQUIRE(p->num_keypoints >= 1 && p->num_keypoints <= kMaxV, SF_E_UNSUPPORTED, "windowing: V=%d outside [1,%d]",
             p->num_keypoints, kMaxV);
  SF_REQUIRE(!p->add_neck || (p->num_keypoints >= 8 && tr->kp_per_frame >= 7), SF_E_INVALID,
             "neck synthesis needs the shoulder keypoints 5 and 6 in the source and a slot for the neck");
  SF_REQUIRE(tr->kp_channels == 0 || tr->kp_channels == 2 || tr->kp_channels == 3, SF_E_INVALID, "windowing: kp_channels must be 2 or 3");
  SF_REQUIRE(!p->include_confidence || tr->kp_channels != 2, SF_E_INVALID, "include_confidence needs (x, y, conf) source keypoints");
  return SF_OK;
}

}  // namespace
}  // namespace sf

using namespace sf;

extern "C" int64_t sf_window_capacity(const sf_tracks* tr, const sf_window_params* p) {
  if (check(tr, p) != SF_OK) return SF_E_INVALID;
  return candidates(tr, p, nullptr);
}

extern "C" int64_t sf_window_workspace_bytes(const sf_tracks* tr, const sf_window_params* p) {
  if (check(tr, p) != SF_OK) return SF_E_INVALID;
  return (int64_t)layout(tr, p).total;
}

// ---- the two halves of the pipeline, shared by sf_window_normalize and sf_score_from_tracks (api.cu)
namespace sf {

int64_t window_candidates(const sf_tracks* tr, const sf_window_params* p) { return candidates(tr, p, nullptr); }

static void fill_ctx(const sf_tracks* tr, const sf_window_params* p, const WsLayout& L, char* ws, WinCtx* out) {
  WinCtx& cx = *out;
  const bool has_gt = tr->gt_dev && tr->gt_offsets_host && tr->track_video_host && tr->n_videos > 0;
  cx.kp = tr->kp_dev;
  cx.frame_no = tr->frame_no_dev;
  cx.track_off = (const int64_t*)(ws + L.track_off);
  cx.cand_off = (const int64_t*)(ws + L.cand_off);
  cx.track_video = (const int32_t*)(ws + L.track_video);
  cx.gt_off = (const int64_t*)(ws + L.gt_off);
  cx.gt = has_gt ? tr->gt_dev : nullptr;
  cx.n_tracks = tr->n_tracks;
  cx.K = tr->kp_per_frame;
  cx.KC = tr->kp_channels == 2 ? 2 : 3;
  cx.T = p->seq_len;
  cx.stride = p->stride;
  cx.max_gap = p->max_gap;
  cx.V = p->num_keypoints;
  cx.normalize = p->normalize;
  cx.add_neck = p->add_neck ? 1 : 0;
  cx.conf = p->include_confidence ? 1 : 0;
  cx.n_cand = L.n_cand;
}

// K_flag / K_scan / K_compact: which windows exist, in the reference's order, with their labels.  The per-track tables
// are uploaded into the workspace (they stay there for window_gather).  No synchronisation.
// The per-track tables (track offsets, candidate offsets, track -> video, ground-truth offsets) occupy the head of the
// workspace.  A caller that uploads them itself (the host-track runner: through pinned staging on its copy stream, so that
// no pageable copy ever sits on the compute stream) packs them with window_tables_pack and passes tables_resident.
int64_t window_tables_bytes(const sf_tracks* tr, const sf_window_params* p) { return (int64_t)layout(tr, p).flag; }

void window_tables_pack(const sf_tracks* tr, const sf_window_params* p, char* host_dst) {
  const WsLayout L = layout(tr, p);
  memset(host_dst, 0, L.flag);
  memcpy(host_dst + L.track_off, tr->track_offsets_host, sizeof(int64_t) * (tr->n_tracks + 1));
  int64_t n = 0;
  int64_t* cand = reinterpret_cast<int64_t*>(host_dst + L.cand_off);
  cand[0] = 0;
  for (int i = 0; i < tr->n_tracks; ++i) {
    const int64_t len = tr->track_offsets_host[i + 1] - tr->track_offsets_host[i];
    if (len >= p->seq_len) n += (len - p->seq_len) / p->stride + 1;
    cand[i + 1] = n;
  }
  if (tr->gt_dev && tr->gt_offsets_host && tr->track_video_host && tr->n_videos > 0) {
    memcpy(host_dst + L.track_video, tr->track_video_host, sizeof(int32_t) * tr->n_tracks);
    memcpy(host_dst + L.gt_off, tr->gt_offsets_host, sizeof(int64_t) * (tr->n_videos + 1));
  }
}

int window_index(const sf_tracks* tr, const sf_window_params* p, int32_t* labels_dev, int32_t* window_track_dev,
                 int32_t* window_start_dev, int64_t* n_windows_dev, void* workspace_dev, int64_t workspace_bytes, cudaStream_t st,
                 bool tables_resident) {
  int rc = check(tr, p);
  if (rc != SF_OK) return rc;
  SF_REQUIRE(n_windows_dev, SF_E_INVALID, "windowing: n_windows_dev is required");
  const WsLayout L = layout(tr, p);
  SF_REQUIRE(workspace_dev && workspace_bytes >= (int64_t)L.total, SF_E_INVALID,
             "windowing: workspace of %lld bytes needed, got %lld", (long long)L.total, (long long)workspace_bytes);
  if (L.n_cand == 0) {
    SF_CUDA_OK(cudaMemsetAsync(n_windows_dev, 0, sizeof(int64_t), st));
    return SF_OK;
  }
  SF_REQUIRE(labels_dev && window_track_dev && window_start_dev, SF_E_INVALID, "windowing: null output");
  char* ws = (char*)workspace_dev;
  if (!tables_resident) {
    std::vector<int64_t> cand_off;
    candidates(tr, p, &cand_off);
    SF_CUDA_OK(cudaMemcpyAsync(ws + L.track_off, tr->track_offsets_host, sizeof(int64_t) * (tr->n_tracks + 1), cudaMemcpyHostToDevice, st));
    // cand_off lives on this stack frame: a pageable source is staged before the call returns
    SF_CUDA_OK(cudaMemcpyAsync(ws + L.cand_off, cand_off.data(), sizeof(int64_t) * (tr->n_tracks + 1), cudaMemcpyHostToDevice, st));
    const bool has_gt = tr->gt_dev && tr->gt_offsets_host && tr->track_video_host && tr->n_videos > 0;
    if (has_gt) {
      SF_CUDA_OK(cudaMemcpyAsync(ws + L.track_video, tr->track_video_host, sizeof(int32_t) * tr->n_tracks, cudaMemcpyHostToDevice, st));
      SF_CUDA_OK(cudaMemcpyAsync(ws + L.gt_off, tr->gt_offsets_host, sizeof(int64_t) * (tr->n_videos + 1), cudaMemcpyHostToDevice, st));
    }
  }
  WinCtx cx;
  fill_ctx(tr, p, L, ws, &cx);
  uint8_t* flag = (uint8_t*)(ws + L.flag);
  uint8_t* label = (uint8_t*)(ws + L.label);
  int32_t* block_sum = (int32_t*)(ws + L.block_sum);
  int64_t* block_off = (int64_t*)(ws + L.block_off);
  k_flag<<<L.n_blocks, kScanBlock, 0, st>>>(cx, flag, label, block_sum);
  k_scan<<<1, 1024, 0, st>>>(block_sum, block_off, L.n_blocks, n_windows_dev);
  k_compact<<<L.n_blocks, kScanBlock, 0, st>>>(cx, flag, label, block_off, labels_dev, window_track_dev, window_start_dev);
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}

// K_gather for windows [w_begin, min(*n_windows_dev, w_begin + w_cap)) -> poses_dev[0 ..) (and frame_idx_dev[0 ..)).
// `workspace_dev` must be the one window_index filled for the same tracks / parameters.
int window_gather(const sf_tracks* tr, const sf_window_params* p, const int32_t* window_track_dev, const int32_t* window_start_dev,
                  const int64_t* n_windows_dev, int64_t w_begin, int64_t w_cap, float* poses_dev, int32_t* frame_idx_dev,
                  void* workspace_dev, cudaStream_t st) {
  if (w_cap <= 0) return SF_OK;
  const WsLayout L = layout(tr, p);
  WinCtx cx;
  fill_ctx(tr, p, L, (char*)workspace_dev, &cx);
  const bool fast = gather_fast(cx);
  const int per_warp = fast ? gather_raw_floats(cx.T, cx.K, cx.KC) : gather_per_warp(cx.T, cx.K, cx.KC, cx.V, cx.conf ? 3 : 2);
  const size_t smem = (size_t)per_warp * kGatherWarps * sizeof(float);
  int grid = 0;
  int rc = gather_launch_config(fast, smem, (w_cap + kGatherWarps - 1) / kGatherWarps, &grid);
  if (rc != SF_OK) return rc;
  if (fast)
    k_gather<true><<<grid, kGatherWarps * 32, smem, st>>>(cx, n_windows_dev, 0, window_track_dev, window_start_dev, poses_dev, frame_idx_dev,
                                                          per_warp, w_begin, w_cap);
  else
    k_gather<false><<<grid, kGatherWarps * 32, smem, st>>>(cx, n_windows_dev, 0, window_track_dev, window_start_dev, poses_dev, frame_idx_dev,
                                                           per_warp, w_begin, w_cap);
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}

int window_device_of(const sf_tracks* tr, int* device) {
  cudaPointerAttributes attr;
  SF_CUDA_OK(cudaPointerGetAttributes(&attr, tr->kp_dev));
  SF_REQUIRE(attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged, SF_E_INVALID, "windowing: kp_dev is not a device pointer");
  *device = attr.device;
  return SF_OK;
}

}  // namespace sf

extern "C" int sf_window_normalize(const sf_tracks* tr, const sf_window_params* p, float* poses_dev, int32_t* labels_dev,
                                   int32_t* window_track_dev, int32_t* window_start_dev, int32_t* frame_idx_dev,
                                   int64_t* n_windows_dev, int64_t* n_windows_host, void* workspace_dev,
                                   int64_t workspace_bytes, void* stream) {
  int rc = check(tr, p);
  if (rc != SF_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  SF_REQUIRE(n_windows_dev, SF_E_INVALID, "windowing: n_windows_dev is required");
  // run on the device that owns the track buffers, whatever the caller's current device is (no track long enough for a
  // window: there may be no keypoint buffer at all -- the count's device then)
  DeviceGuard guard;
  int dev = 0;
  if (candidates(tr, p, nullptr) == 0) {
    cudaPointerAttributes attr;
    SF_CUDA_OK(cudaPointerGetAttributes(&attr, n_windows_dev));
    SF_REQUIRE(attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged, SF_E_INVALID, "windowing: n_windows_dev is not a device pointer");
    dev = attr.device;
  } else {
    rc = window_device_of(tr, &dev);
    if (rc != SF_OK) return rc;
  }
  SF_CUDA_OK(guard.enter(dev));
  rc = window_index(tr, p, labels_dev, window_track_dev, window_start_dev, n_windows_dev, workspace_dev, workspace_bytes, st, false);
  if (rc != SF_OK) return rc;
  const int64_t n_cand = candidates(tr, p, nullptr);
  if (n_cand > 0) {
    SF_REQUIRE(poses_dev, SF_E_INVALID, "windowing: null output");
    rc = window_gather(tr, p, window_track_dev, window_start_dev, n_windows_dev, 0, n_cand, poses_dev, frame_idx_dev, workspace_dev, st);
    if (rc != SF_OK) return rc;
  }
  if (n_windows_host) {
    SF_CUDA_OK(cudaMemcpyAsync(n_windows_host, n_windows_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    SF_CUDA_OK(cudaStreamSynchronize(st));
  }
  return SF_OK;
}

// Normalisation of PRE-CUT windows (streaming mode / callers that window on the host): raw_dev is
// (B, T, K, 3) AoS keypoints, the output is the model's (B, 2, T, V) layout.  No host work, no sync:
// capturable in a CUDA graph.
extern "C" int sf_normalize_windows(const float* raw_dev, int64_t B, int32_t T, int32_t K, int32_t V, int32_t normalize,
                                    float* poses_dev, void* stream) {
  SF_REQUIRE(B >= 0 && T >= 1 && K >= 1 && V >= 1 && V <= kMaxV && (B == 0 || (raw_dev && poses_dev)), SF_E_INVALID,
             "sf_normalize_windows: bad argument");
  SF_REQUIRE(V != 18 || K >= 17, SF_E_INVALID, "neck synthesis needs >= 17 source keypoints");
  if (B == 0) return SF_OK;
  DeviceGuard guard;
  {
    cudaPointerAttributes attr;
    SF_CUDA_OK(cudaPointerGetAttributes(&attr, raw_dev));
    SF_REQUIRE(attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged, SF_E_INVALID, "sf_normalize_windows: raw_dev is not a device pointer");
    SF_CUDA_OK(guard.enter(attr.device));
  }
  WinCtx cx{};
  cx.kp = raw_dev;
  cx.K = K;
  cx.KC = 3;
  cx.T = T;
  cx.V = V;
  cx.normalize = normalize;
  cx.add_neck = V == 18 ? 1 : 0;          // this entry point keeps the variant-2 convention: 18 keypoints = 17 + synthetic neck
  cx.conf = 0;
  const bool fast = gather_fast(cx);
  const int per_warp = fast ? gather_raw_floats(T, K, 3) : gather_per_warp(T, K, 3, V, 2);
  const size_t smem = (size_t)per_warp * kGatherWarps * sizeof(float);
  int grid = 0;
  int rc = gather_launch_config(fast, smem, (B + kGatherWarps - 1) / kGatherWarps, &grid);
  if (rc != SF_OK) return rc;
  if (fast)
    k_gather<true><<<grid, kGatherWarps * 32, smem, (cudaStream_t)stream>>>(cx, nullptr, B, nullptr, nullptr, poses_dev, nullptr, per_warp, 0, B);
  else
    k_gather<false><<<grid, kGatherWarps * 32, smem, (cudaStream_t)stream>>>(cx, nullptr, B, nullptr, nullptr, poses_dev, nullptr, per_warp, 0, B);
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}
