// What follows the scoring path on the device: per-video aggregation of window scores and the
// ranking metrics (AUC-ROC, average precision, Youden threshold, confusion counts).
//
// Replaces the host post-processing of a sweep:
//   shopformer_2/evaluate.py:65-118 + shopformer_2/utils/metrics.py:148-188 (group window scores per video; max / mean /
//   95th percentile; the video's label is the label of its LAST window, as the reference's dict assignment leaves it),
//   shopformer/utils/metrics.py:18-77 and shopformer_2/utils/metrics.py:21-145 (sklearn roc_auc_score,
//   average_precision_score, roc_curve + Youden J threshold, `scores >= threshold` confusion counts).
//
// Both are built on one LSD radix sort (8-bit digits, stable: one warp walks its 2048-key slice in order and ranks equal
// digits with match.any) and on tiled device-wide scans.  Everything after the sort is HBM-streaming integer / fp64 work.
#include <algorithm>
#include <cmath>

#include "sf_internal.h"

namespace sf {
namespace {

constexpr int kSortWarps = 8;                  // warps per block
constexpr int kWarpKeys = 2048;                // consecutive keys owned by one warp

__device__ __forceinline__ uint32_t orderable(float f) {          // monotone float -> uint32 (-inf < ... < +inf)
  const uint32_t u = __float_as_uint(f + 0.0f);                    // -0.0 -> +0.0: they compare equal, so they must tie
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// ---- radix sort pass: count[digit][warp] -> exclusive scan -> stable scatter
__global__ void __launch_bounds__(kSortWarps * 32)
k_radix_count(const uint64_t* __restrict__ keys, int64_t n, int shift, uint32_t* __restrict__ counts, int64_t n_warps) {
  __shared__ uint32_t hist[kSortWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 256; i += 32) hist[warp][i] = 0;
  __syncwarp();
  const int64_t gw = (int64_t)blockIdx.x * kSortWarps + warp;
  const int64_t base = gw * kWarpKeys;
  for (int c = 0; c < kWarpKeys; c += 32) {
    const int64_t i = base + c + lane;
    const bool in = i < n;
    const uint32_t d = in ? (uint32_t)((keys[i] >> shift) & 0xFF) : 0x100u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    if (in && lane == __ffs(peers) - 1) hist[warp][d] += __popc(peers);
    __syncwarp();
  }
  if (gw < n_warps)
    for (int i = lane; i < 256; i += 32) counts[(int64_t)i * n_warps + gw] = hist[warp][i];
}

__global__ void __launch_bounds__(kSortWarps * 32)
k_radix_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, int shift,
                const int64_t* __restrict__ offsets, int64_t n_warps, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  __shared__ int64_t off[kSortWarps][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t gw = (int64_t)blockIdx.x * kSortWarps + warp;
  if (gw >= n_warps) return;
  for (int i = lane; i < 256; i += 32) off[warp][i] = offsets[(int64_t)i * n_warps + gw];
  __syncwarp();
  const int64_t base = gw * kWarpKeys;
  for (int c = 0; c < kWarpKeys; c += 32) {
    const int64_t i = base + c + lane;
    const bool in = i < n;
    const uint64_t k = in ? keys[i] : 0;
    const uint32_t d = in ? (uint32_t)((k >> shift) & 0xFF) : 0x100u;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int64_t pos = 0;
    if (in) pos = off[warp][d] + rank;
    __syncwarp();
    if (in && lane == __ffs(peers) - 1) off[warp][d] += __popc(peers);
    __syncwarp();
    if (in) {
      keys_out[pos] = k;
      if (vals) vals_out[pos] = vals[i];
    }
  }
}

// ---- device-wide scans over int64: tiles of 1024 x 4, tile totals scanned by one CTA, then applied
constexpr int kScanThreads = 1024, kScanItems = 4, kScanTile = kScanThreads * kScanItems;
struct OpAdd {
  __device__ static int64_t id() { return 0; }
  __device__ static int64_t f(int64_t a, int64_t b) { return a + b; }
};
struct OpMax {
  __device__ static int64_t id() { return INT64_MIN; }
  __device__ static int64_t f(int64_t a, int64_t b) { return a > b ? a : b; }
};

template <class Op>
__device__ __forceinline__ int64_t block_scan_inclusive(int64_t x, int64_t* warp_tot /*[32]*/, int64_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x = Op::f(x, y);
  }
  if (lane == 31) warp_tot[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int64_t w = warp_tot[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w = Op::f(w, y);
    }
    warp_tot[lane] = w;
  }
  __syncthreads();
  if (warp > 0) x = Op::f(warp_tot[warp - 1], x);
  if (total) *total = warp_tot[31];
  __syncthreads();
  return x;
}

// in[i] is produced by `load(i)`; pass 0 writes tile totals, pass 1 writes the inclusive scan given scanned tile prefixes
template <class Op, class Load>
__device__ __forceinline__ void tile_scan(int64_t n, Load load, int64_t* __restrict__ tile_tot, const int64_t* __restrict__ tile_pre,
                                          int64_t* __restrict__ out) {
  __shared__ int64_t warp_tot[32];
  const int64_t t0 = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int64_t v[kScanItems];
  int64_t acc = Op::id();
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    v[j] = (t0 + j < n) ? load(t0 + j) : Op::id();
    acc = Op::f(acc, v[j]);
    v[j] = acc;
  }
  int64_t total;
  const int64_t incl = block_scan_inclusive<Op>(acc, warp_tot, &total);
  if (tile_tot) {
    if (threadIdx.x == 0) tile_tot[blockIdx.x] = total;
    return;
  }
  // exclusive prefix of this thread = inclusive of the previous thread; recover it without a subtraction (max has none)
  const int64_t prev_incl = __shfl_up_sync(0xffffffffu, incl, 1);
  __shared__ int64_t last_of_warp[32];
  if ((threadIdx.x & 31) == 31) last_of_warp[threadIdx.x >> 5] = incl;
  __syncthreads();
  int64_t before = Op::id();
  if ((threadIdx.x & 31) > 0) before = prev_incl;
  else if (threadIdx.x > 0) before = last_of_warp[(threadIdx.x >> 5) - 1];
  if (blockIdx.x > 0) before = Op::f(tile_pre[blockIdx.x - 1], before);
#pragma unroll
  for (int j = 0; j < kScanItems; ++j)
    if (t0 + j < n) out[t0 + j] = Op::f(before, v[j]);
}

// inclusive scan of tile totals in place (single CTA, sequential over chunks of 1024)
template <class Op>
__global__ void __launch_bounds__(1024) k_scan_totals(int64_t* __restrict__ tot, int64_t n) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) carry = Op::id();
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    int64_t x = i < n ? tot[i] : Op::id();
    int64_t total;
    x = block_scan_inclusive<Op>(x, warp_tot, &total);
    if (i < n) tot[i] = Op::f(carry, x);
    __syncthreads();
    if (threadIdx.x == 0) carry = Op::f(carry, total);
    __syncthreads();
  }
}

// scan #1: exclusive offsets of the radix counts (uint32 counts -> int64 exclusive offsets)
__global__ void __launch_bounds__(kScanThreads) k_counts_tot(const uint32_t* __restrict__ c, int64_t n, int64_t* tile_tot) {
  tile_scan<OpAdd>(n, [&](int64_t i) { return (int64_t)c[i]; }, tile_tot, nullptr, nullptr);
}
__global__ void __launch_bounds__(kScanThreads) k_counts_apply(const uint32_t* __restrict__ c, int64_t n, const int64_t* tile_pre, int64_t* out) {
  tile_scan<OpAdd>(n, [&](int64_t i) { return (int64_t)c[i]; }, nullptr, tile_pre, out);
}
__global__ void k_incl_to_excl(const uint32_t* __restrict__ c, int64_t n, int64_t* __restrict__ io) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) io[i] -= (int64_t)c[i];
}

struct SortWs {
  uint64_t *keys_a, *keys_b;
  uint32_t *vals_a, *vals_b;
  uint32_t* counts;
  int64_t* offsets;
  int64_t* tile_tot;
  int64_t n_warps, n_blocks, n_cnt, n_tiles;
};
int64_t sort_ws_layout(int64_t n, bool with_vals, char* base, SortWs* w) {
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    const int64_t o = off;
    off += (bytes + 255) & ~int64_t(255);
    return base ? base + o : (char*)nullptr;
  };
  SortWs t{};
  t.n_warps = std::max<int64_t>(1, (n + kWarpKeys - 1) / kWarpKeys);
  t.n_blocks = (t.n_warps + kSortWarps - 1) / kSortWarps;
  t.n_cnt = 256 * t.n_warps;
  t.n_tiles = (t.n_cnt + kScanTile - 1) / kScanTile;
  t.keys_a = (uint64_t*)take(n * 8);
  t.keys_b = (uint64_t*)take(n * 8);
  t.vals_a = (uint32_t*)take(with_vals ? n * 4 : 0);
  t.vals_b = (uint32_t*)take(with_vals ? n * 4 : 0);
  t.counts = (uint32_t*)take(t.n_cnt * 4);
  t.offsets = (int64_t*)take(t.n_cnt * 8);
  t.tile_tot = (int64_t*)take(t.n_tiles * 8);
  if (w) *w = t;
  return off;
}
// sorts keys_a (ascending, stable) over key bits [0, bits); the result is in *keys_out / *vals_out (a or b)
int radix_sort(SortWs& w, int64_t n, int bits, bool with_vals, cudaStream_t st, uint64_t** keys_out, uint32_t** vals_out) {
  uint64_t *ka = w.keys_a, *kb = w.keys_b;
  uint32_t *va = w.vals_a, *vb = w.vals_b;
  for (int shift = 0; shift < bits; shift += 8) {
    k_radix_count<<<(unsigned)w.n_blocks, kSortWarps * 32, 0, st>>>(ka, n, shift, w.counts, w.n_warps);
    k_counts_tot<<<(unsigned)w.n_tiles, kScanThreads, 0, st>>>(w.counts, w.n_cnt, w.tile_tot);
    k_scan_totals<OpAdd><<<1, 1024, 0, st>>>(w.tile_tot, w.n_tiles);
    k_counts_apply<<<(unsigned)w.n_tiles, kScanThreads, 0, st>>>(w.counts, w.n_cnt, w.tile_tot, w.offsets);
    k_incl_to_excl<<<(unsigned)((w.n_cnt + 255) / 256), 256, 0, st>>>(w.counts, w.n_cnt, w.offsets);
    k_radix_scatter<<<(unsigned)w.n_blocks, kSortWarps * 32, 0, st>>>(ka, with_vals ? va : nullptr, n, shift, w.offsets, w.n_warps, kb, vb);
    std::swap(ka, kb);
    std::swap(va, vb);
  }
  SF_CUDA_OK(cudaGetLastError());
  *keys_out = ka;
  *vals_out = va;
  return SF_OK;
}

// ====================================================================== video-level aggregation
__global__ void k_video_keys(const float* __restrict__ scores, const int32_t* __restrict__ vid, int64_t n, int n_videos,
                             uint64_t* __restrict__ keys, int32_t* __restrict__ last_idx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v = vid[i];
  const bool ok = v >= 0 && v < n_videos;
  if (!ok) v = n_videos;                                    // out-of-range ids sort behind every video and are ignored
  keys[i] = ((uint64_t)(uint32_t)v << 32) | orderable(scores[i]);
  if (ok) atomicMax(&last_idx[v], (int32_t)i);              // dataset order: the reference keeps the LAST window's label
}
__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__device__ __forceinline__ int64_t lower_bound_u64(const uint64_t* __restrict__ a, int64_t n, uint64_t key) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// one warp per video over its sorted segment
__global__ void __launch_bounds__(256)
k_video_reduce(const uint64_t* __restrict__ keys, int64_t n, int n_videos, const int32_t* __restrict__ labels,
               const int32_t* __restrict__ last_idx, double* __restrict__ agg_max, double* __restrict__ agg_mean,
               double* __restrict__ agg_p95, int32_t* __restrict__ video_label, int32_t* __restrict__ count) {
  const int v = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= n_videos) return;
  const int64_t s = lower_bound_u64(keys, n, (uint64_t)(uint32_t)v << 32);
  const int64_t e = lower_bound_u64(keys, n, (uint64_t)((uint32_t)v + 1u) << 32);
  const int64_t cnt = e - s;
  double sum = 0.0;
  for (int64_t i = s + lane; i < e; i += 32) sum += (double)from_orderable((uint32_t)keys[i]);
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane != 0) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  if (count) count[v] = (int32_t)cnt;
  if (video_label) video_label[v] = (cnt > 0 && labels) ? labels[last_idx[v]] : 0;
  if (cnt == 0) {
    if (agg_max) agg_max[v] = nan;
    if (agg_mean) agg_mean[v] = nan;
    if (agg_p95) agg_p95[v] = nan;
    return;
  }
  if (agg_max) agg_max[v] = (double)from_orderable((uint32_t)keys[e - 1]);
  if (agg_mean) agg_mean[v] = sum / (double)cnt;
  if (agg_p95) {
    // numpy.percentile(x, 95), method 'linear': virtual index (n - 1) * 0.95, lerp between the neighbours
    const double pos = (double)(cnt - 1) * 0.95;
    const int64_t lo = (int64_t)floor(pos);
    const int64_t hi = lo + 1 < cnt ? lo + 1 : cnt - 1;
    const double t = pos - (double)lo;
    const double a = (double)from_orderable((uint32_t)keys[s + lo]), b = (double)from_orderable((uint32_t)keys[s + hi]);
    double r = a + (b - a) * t;                                     // numpy's _lerp
    if (t >= 0.5) r = b - (b - a) * (1.0 - t);
    if (t == 0.0 || a == b) r = a;
    agg_p95[v] = r;
  }
}

int bits_for(uint32_t max_value) {
  int b = 0;
  while (b < 32 && (max_value >> b)) ++b;
  return b;
}

// ====================================================================== ranking metrics
__global__ void k_rank_keys(const float* __restrict__ scores, const int32_t* __restrict__ labels, int64_t n, uint64_t* __restrict__ keys,
                            uint32_t* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = (uint64_t)(~orderable(scores[i]));                 // ascending sort of the complement = descending scores
  vals[i] = labels[i] != 0 ? 1u : 0u;
}
// scans over the sorted order: cumulative positives; start index of the tie group an element belongs to
__global__ void __launch_bounds__(kScanThreads) k_tp_tot(const uint32_t* __restrict__ lab, int64_t n, int64_t* tile_tot) {
  tile_scan<OpAdd>(n, [&](int64_t i) { return (int64_t)lab[i]; }, tile_tot, nullptr, nullptr);
}
__global__ void __launch_bounds__(kScanThreads) k_tp_apply(const uint32_t* __restrict__ lab, int64_t n, const int64_t* tile_pre, int64_t* out) {
  tile_scan<OpAdd>(n, [&](int64_t i) { return (int64_t)lab[i]; }, nullptr, tile_pre, out);
}
__global__ void __launch_bounds__(kScanThreads) k_gs_tot(const uint64_t* __restrict__ keys, int64_t n, int64_t* tile_tot) {
  tile_scan<OpMax>(n, [&](int64_t i) { return (i == 0 || keys[i] != keys[i - 1]) ? i : INT64_MIN; }, tile_tot, nullptr, nullptr);
}
__global__ void __launch_bounds__(kScanThreads) k_gs_apply(const uint64_t* __restrict__ keys, int64_t n, const int64_t* tile_pre, int64_t* out) {
  tile_scan<OpMax>(n, [&](int64_t i) { return (i == 0 || keys[i] != keys[i - 1]) ? i : INT64_MIN; }, nullptr, tile_pre, out);
}

// One contribution per tie group (taken at the group's LAST element): trapezoid of the ROC curve, step of the PR curve,
// Youden J.  Block partials in fp64 (fixed order), best-J as (J, -index) so the FIRST maximum in descending-score order wins.
struct RankPartial {
  double roc, ap, best_j;
  int64_t best_i;
};
__global__ void __launch_bounds__(256)
k_rank_contrib(const uint64_t* __restrict__ keys, const int64_t* __restrict__ tp_cum, const int64_t* __restrict__ grp_start, int64_t n,
               RankPartial* __restrict__ partial) {
  __shared__ double s_roc[256], s_ap[256], s_j[256];
  __shared__ int64_t s_i[256];
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  double roc = 0.0, ap = 0.0, bj = -2.0;
  int64_t bi = INT64_MAX;
  if (i < n && (i == n - 1 || keys[i] != keys[i + 1])) {
    const int64_t P = tp_cum[n - 1], N = n - P;
    const int64_t s = grp_start[i];
    const int64_t tp = tp_cum[i], tp0 = s > 0 ? tp_cum[s - 1] : 0;
    const int64_t fp = (i + 1) - tp, fp0 = s - tp0;
    if (P > 0 && N > 0) {
      roc = ((double)(fp - fp0) / (double)N) * (((double)tp + (double)tp0) / (double)P) * 0.5;
      bj = (double)tp / (double)P - (double)fp / (double)N;
      bi = i;
    }
    if (P > 0) ap = ((double)(tp - tp0) / (double)P) * ((double)tp / (double)(i + 1));
  }
  s_roc[threadIdx.x] = roc; s_ap[threadIdx.x] = ap; s_j[threadIdx.x] = bj; s_i[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) {
      s_roc[threadIdx.x] += s_roc[threadIdx.x + o];
      s_ap[threadIdx.x] += s_ap[threadIdx.x + o];
      const double j2 = s_j[threadIdx.x + o];
      const int64_t i2 = s_i[threadIdx.x + o];
      if (j2 > s_j[threadIdx.x] || (j2 == s_j[threadIdx.x] && i2 < s_i[threadIdx.x])) { s_j[threadIdx.x] = j2; s_i[threadIdx.x] = i2; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = RankPartial{s_roc[0], s_ap[0], s_j[0], s_i[0]};
}
// final: sums the partials in order, picks the threshold, leaves [auc_roc, auc_pr, threshold, P] in out4
__global__ void __launch_bounds__(256)
k_rank_final(const RankPartial* __restrict__ partial, int64_t n_part, const uint64_t* __restrict__ keys, const int64_t* __restrict__ tp_cum,
             int64_t n, float threshold_in, double* __restrict__ out4) {
  __shared__ double s_roc[256], s_ap[256], s_j[256];
  __shared__ int64_t s_i[256];
  double roc = 0.0, ap = 0.0, bj = -2.0;
  int64_t bi = INT64_MAX;
  for (int64_t k = threadIdx.x; k < n_part; k += 256) {
    const RankPartial p = partial[k];
    roc += p.roc; ap += p.ap;
    if (p.best_j > bj || (p.best_j == bj && p.best_i < bi)) { bj = p.best_j; bi = p.best_i; }
  }
  s_roc[threadIdx.x] = roc; s_ap[threadIdx.x] = ap; s_j[threadIdx.x] = bj; s_i[threadIdx.x] = bi;
  __syncthreads();
  for (int o = 128; o; o >>= 1) {
    if (threadIdx.x < o) {
      s_roc[threadIdx.x] += s_roc[threadIdx.x + o];
      s_ap[threadIdx.x] += s_ap[threadIdx.x + o];
      const double j2 = s_j[threadIdx.x + o];
      const int64_t i2 = s_i[threadIdx.x + o];
      if (j2 > s_j[threadIdx.x] || (j2 == s_j[threadIdx.x] && i2 < s_i[threadIdx.x])) { s_j[threadIdx.x] = j2; s_i[threadIdx.x] = i2; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int64_t P = tp_cum[n - 1], N = n - P;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    out4[0] = (P > 0 && N > 0) ? s_roc[0] : nan;
    out4[1] = P > 0 ? s_ap[0] : nan;
    double thr = (double)threshold_in;
    if (threshold_in != threshold_in) {
      // roc_curve's first point is (0, 0) at threshold +inf with J = 0: a later point must beat it strictly
      thr = (s_i[0] != INT64_MAX && s_j[0] > 0.0) ? (double)from_orderable(~(uint32_t)keys[s_i[0]]) : (double)INFINITY;
    }
    out4[2] = thr;
    out4[3] = (double)P;
  }
}
// confusion counts at the threshold: predictions = scores >= threshold.  In descending order that is a prefix.
__global__ void k_rank_confusion(const uint64_t* __restrict__ keys, const int64_t* __restrict__ tp_cum, int64_t n, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double thr = out[2];
  // first index whose score < thr (scores descending)
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const double s = (double)from_orderable(~(uint32_t)keys[mid]);
    if (s >= thr) lo = mid + 1; else hi = mid;
  }
  const int64_t P = tp_cum[n - 1];
  const int64_t tp = lo > 0 ? tp_cum[lo - 1] : 0, fp = lo - tp;
  out[4] = (double)tp;
  out[5] = (double)fp;
  out[6] = (double)((n - P) - fp);
  out[7] = (double)(P - tp);
}

struct RankWs {
  SortWs sort;
  int64_t *tp_cum, *grp_start, *tile_tot;
  RankPartial* partial;
  double* out_dev;
  int64_t n_tiles, n_part;
};
int64_t rank_ws_layout(int64_t n, char* base, RankWs* w) {
  RankWs t{};
  int64_t off = sort_ws_layout(n, true, base, &t.sort);
  auto take = [&](int64_t bytes) {
    const int64_t o = off;
    off += (bytes + 255) & ~int64_t(255);
    return base ? base + o : (char*)nullptr;
  };
  t.n_tiles = std::max<int64_t>(1, (n + kScanTile - 1) / kScanTile);
  t.n_part = std::max<int64_t>(1, (n + 255) / 256);
  t.tp_cum = (int64_t*)take(n * 8);
  t.grp_start = (int64_t*)take(n * 8);
  t.tile_tot = (int64_t*)take(t.n_tiles * 8);
  t.partial = (RankPartial*)take(t.n_part * (int64_t)sizeof(RankPartial));
  t.out_dev = (double*)take(8 * 8);
  if (w) *w = t;
  return off;
}

}  // namespace
}  // namespace sf

using namespace sf;

extern "C" int64_t sf_video_aggregate_workspace_bytes(int64_t n, int32_t n_videos) {
  if (n < 0 || n_videos < 0) return SF_E_INVALID;
  return sort_ws_layout(std::max<int64_t>(n, 1), false, nullptr, nullptr) + (((int64_t)std::max(n_videos, 1) * 4 + 255) & ~int64_t(255));
}

extern "C" int sf_video_aggregate(const float* scores_dev, const int32_t* video_id_dev, const int32_t* labels_dev, int64_t n,
                                  int32_t n_videos, double* agg_max_dev, double* agg_mean_dev, double* agg_p95_dev,
                                  int32_t* video_label_dev, int32_t* count_dev, void* workspace_dev, int64_t workspace_bytes,
                                  void* stream) {
  SF_REQUIRE(n >= 0 && n_videos >= 0 && n < (int64_t)1 << 31, SF_E_INVALID, "sf_video_aggregate: bad sizes");
  if (n_videos == 0) return SF_OK;
  SF_REQUIRE(n == 0 || (scores_dev && video_id_dev), SF_E_INVALID, "sf_video_aggregate: null input");
  SF_REQUIRE(workspace_dev && workspace_bytes >= sf_video_aggregate_workspace_bytes(n, n_videos), SF_E_INVALID,
             "sf_video_aggregate: workspace of %lld bytes needed, got %lld", (long long)sf_video_aggregate_workspace_bytes(n, n_videos),
             (long long)workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard guard;
  {
    cudaPointerAttributes attr;
    SF_CUDA_OK(cudaPointerGetAttributes(&attr, workspace_dev));
    SF_REQUIRE(attr.type == cudaMemoryTypeDevice, SF_E_INVALID, "sf_video_aggregate: workspace is not device memory");
    SF_CUDA_OK(guard.enter(attr.device));
  }
  const int64_t nn = std::max<int64_t>(n, 1);
  SortWs w;
  const int64_t sort_bytes = sort_ws_layout(nn, false, (char*)workspace_dev, &w);
  int32_t* last_idx = (int32_t*)((char*)workspace_dev + sort_bytes);
  k_fill_i32<<<(n_videos + 255) / 256, 256, 0, st>>>(last_idx, n_videos, -1);
  uint64_t* sorted = w.keys_a;
  uint32_t* unused = nullptr;
  if (n > 0) {
    k_video_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scores_dev, video_id_dev, n, n_videos, w.keys_a, last_idx);
    int rc = radix_sort(w, n, 32 + bits_for((uint32_t)n_videos), false, st, &sorted, &unused);
    if (rc) return rc;
  }
  k_video_reduce<<<(unsigned)(((int64_t)n_videos * 32 + 255) / 256), 256, 0, st>>>(sorted, n, n_videos, labels_dev, last_idx, agg_max_dev,
                                                                                  agg_mean_dev, agg_p95_dev, video_label_dev, count_dev);
  SF_CUDA_OK(cudaGetLastError());
  return SF_OK;
}

extern "C" int64_t sf_ranking_metrics_workspace_bytes(int64_t n) {
  if (n < 0) return SF_E_INVALID;
  return rank_ws_layout(std::max<int64_t>(n, 1), nullptr, nullptr);
}

extern "C" int sf_ranking_metrics(const float* scores_dev, const int32_t* labels_dev, int64_t n, float threshold, double* out_host,
                                  void* workspace_dev, int64_t workspace_bytes, void* stream) {
  SF_REQUIRE(n >= 1 && n < (int64_t)1 << 31 && scores_dev && labels_dev && out_host, SF_E_INVALID, "sf_ranking_metrics: bad argument");
  SF_REQUIRE(workspace_dev && workspace_bytes >= sf_ranking_metrics_workspace_bytes(n), SF_E_INVALID,
             "sf_ranking_metrics: workspace of %lld bytes needed, got %lld", (long long)sf_ranking_metrics_workspace_bytes(n),
             (long long)workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard guard;
  {
    cudaPointerAttributes attr;
    SF_CUDA_OK(cudaPointerGetAttributes(&attr, workspace_dev));
    SF_REQUIRE(attr.type == cudaMemoryTypeDevice, SF_E_INVALID, "sf_ranking_metrics: workspace is not device memory");
    SF_CUDA_OK(guard.enter(attr.device));
  }
  RankWs w;
  rank_ws_layout(n, (char*)workspace_dev, &w);
  k_rank_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scores_dev, labels_dev, n, w.sort.keys_a, w.sort.vals_a);
  uint64_t* keys = nullptr;
  uint32_t* lab = nullptr;
  int rc = radix_sort(w.sort, n, 32, true, st, &keys, &lab);
  if (rc) return rc;
  k_tp_tot<<<(unsigned)w.n_tiles, kScanThreads, 0, st>>>(lab, n, w.tile_tot);
  k_scan_totals<OpAdd><<<1, 1024, 0, st>>>(w.tile_tot, w.n_tiles);
  k_tp_apply<<<(unsigned)w.n_tiles, kScanThreads, 0, st>>>(lab, n, w.tile_tot, w.tp_cum);
  k_gs_tot<<<(unsigned)w.n_tiles, kScanThreads, 0, st>>>(keys, n, w.tile_tot);
  k_scan_totals<OpMax><<<1, 1024, 0, st>>>(w.tile_tot, w.n_tiles);
  k_gs_apply<<<(unsigned)w.n_tiles, kScanThreads, 0, st>>>(keys, n, w.tile_tot, w.grp_start);
  k_rank_contrib<<<(unsigned)w.n_part, 256, 0, st>>>(keys, w.tp_cum, w.grp_start, n, w.partial);
  k_rank_final<<<1, 256, 0, st>>>(w.partial, w.n_part, keys, w.tp_cum, n, threshold, w.out_dev);
  k_rank_confusion<<<1, 32, 0, st>>>(keys, w.tp_cum, n, w.out_dev);
  SF_CUDA_OK(cudaGetLastError());
  SF_CUDA_OK(cudaMemcpyAsync(out_host, w.out_dev, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
  SF_CUDA_OK(cudaStreamSynchronize(st));
  return SF_OK;
}
