// sf_selftest_umma: one 128 x N x K tcgen05 product on a single CTA, used by the GPU tests to pin the
// descriptor encodings (no-swizzle K-major A, K-major / MN-major B, arbitrary 16-byte row shifts of A),
// the TMEM lane/column mapping of tcgen05.ld and the commit/mbarrier handshake against numpy.
#include <cuda_fp16.h>

#include <vector>

#include "sf_internal.h"
#include "tc_common.cuh"

namespace sf {
namespace {

using namespace tc;

// A: (a_rows, K) fp32 row-major in global; the MMA uses rows [shift, shift+128).
// B: mode 0 -> (N, K) row-major (K-major operand); mode 1 -> (K, N) row-major (MN-major operand).
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                     int a_rows, int shift, int mode, int a_f16, int b_f16) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem_raw);               // [K/8][a_rows][8]
  const int b_rows = mode == 0 ? N : K;                                          // rows of the B buffer
  const int b_cols = mode == 0 ? K : N;
  __nv_bfloat16* sB = sA + (size_t)(K / 8) * a_rows * 8;                         // [b_cols/8][b_rows][8]
  const int warp = threadIdx.x >> 5;

  if (warp == 0) tmem_alloc(&tmem_base, 64);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < a_rows * K; i += blockDim.x) {
    const int r = i / K, c = i % K;
    if (a_f16) reinterpret_cast<__half*>(sA)[pc_index(r, c, a_rows)] = __float2half_rn(A[i]);
    else sA[pc_index(r, c, a_rows)] = __float2bfloat16(A[i]);
  }
  for (int i = threadIdx.x; i < b_rows * b_cols; i += blockDim.x) {
    const int r = i / b_cols, c = i % b_cols;
    if (b_f16) reinterpret_cast<__half*>(sB)[pc_index(r, c, b_rows)] = __float2half_rn(B[i]);
    else sB[pc_index(r, c, b_rows)] = __float2bfloat16(B[i]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base;

  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N, mode == 1, a_f16 != 0, b_f16 != 0);
    const uint32_t a_plane = (uint32_t)a_rows * 16u, b_plane = (uint32_t)b_rows * 16u;
    const uint64_t adesc0 = make_desc(smem_u32(sA) + (uint32_t)shift * 16u, /*lbo=*/a_plane, /*sbo=*/128u);
    // K-major B: LBO = next K chunk (plane), SBO = next 8 N rows (128 B).
    // MN-major B: LBO = next 8 K rows (128 B), SBO = next N chunk (plane).
    const uint64_t bdesc0 = mode == 0 ? make_desc(smem_u32(sB), b_plane, 128u) : make_desc(smem_u32(sB), 128u, b_plane);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t ad = desc_advance(adesc0, (uint32_t)k * 2u * a_plane);                 // two K chunks per MMA
      const uint64_t bd = mode == 0 ? desc_advance(bdesc0, (uint32_t)k * 2u * b_plane)      // two K chunks
                                    : desc_advance(bdesc0, (uint32_t)k * 256u);             // 16 K rows
      umma_bf16(tbase, ad, bd, idesc, k > 0);
    }
    umma_commit(&bar);
  }
  {
    // bounded wait: a wrong descriptor must fail the test, not hang the GPU
    bool done = false;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin) done = mbar_try_wait(&bar, 0);
    if (!done) {
      for (int c = 0; c < N; ++c) D[(warp * 32 + (threadIdx.x & 31)) * N + c] = __int_as_float(0x7fc00000);
      return;
    }
  }
  tc_fence_after();
  const int row = warp * 32 + (threadIdx.x & 31);
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) D[row * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 64);
}

}  // namespace
}  // namespace sf

extern "C" int sf_selftest_umma(int32_t mode, int32_t N, int32_t K, int32_t shift, const float* a_host,
                                const float* b_host, float* d_host) {
  using namespace sf;
  // mode bits: 0 = B is MN-major, 2 and 3 (value 12) = both operands hold fp16 instead of bf16
  const int a_f16 = (mode >> 2) & 1, b_f16 = (mode >> 3) & 1;
  mode &= 1;
  SF_REQUIRE(a_f16 == b_f16, SF_E_INVALID, "tcgen05 kind::f16 takes A and B in the same format (a mixed descriptor is an illegal instruction)");
  SF_REQUIRE(N >= 16 && N <= 64 && N % 16 == 0 && K >= 16 && K % 16 == 0 && K <= 512 &&
                 shift >= 0 && shift <= 64 && a_host && b_host && d_host,
             SF_E_INVALID, "sf_selftest_umma: bad argument");
  int rc = sf_device_count();
  if (rc < 0) return rc;
  const int a_rows = 128 + shift;
  float *dA = nullptr, *dB = nullptr, *dD = nullptr;
  SF_CUDA_OK(cudaMalloc(&dA, sizeof(float) * a_rows * K));
  SF_CUDA_OK(cudaMalloc(&dB, sizeof(float) * N * K));
  SF_CUDA_OK(cudaMalloc(&dD, sizeof(float) * 128 * N));
  SF_CUDA_OK(cudaMemcpy(dA, a_host, sizeof(float) * a_rows * K, cudaMemcpyHostToDevice));
  SF_CUDA_OK(cudaMemcpy(dB, b_host, sizeof(float) * N * K, cudaMemcpyHostToDevice));
  const size_t smem = ((size_t)a_rows * K + (size_t)N * K) * 2 + 256;
  SF_CUDA_OK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_selftest_kernel<<<1, 128, smem>>>(dA, dB, dD, N, K, a_rows, shift, mode, a_f16, b_f16);
  SF_CUDA_OK(cudaGetLastError());
  SF_CUDA_OK(cudaDeviceSynchronize());
  SF_CUDA_OK(cudaMemcpy(d_host, dD, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  return SF_OK;
}

// ---------------------------------------------------------------------------------------------------------
// debugging aid (not part of the C ABI): cycles for `n_mma` back-to-back 128 x N x 16 MMAs issued by one thread
// on no-swizzle operands (garbage data), measured from first issue to mbarrier completion.
namespace sf {
namespace {
__global__ void __launch_bounds__(128, 1) umma_timing_kernel(int N, int n_mma, int a_rows_shift, int mode, long long* out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  using namespace tc;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 256);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(128, N, false);
    const uint32_t a_plane = 1024u * 16u, b_plane = (uint32_t)N * 16u;
    uint64_t ad = make_desc(smem_u32(smem_raw) + (uint32_t)a_rows_shift * 16u, a_plane, 128u);
    uint64_t bd = make_desc(smem_u32(smem_raw) + 40960u, b_plane, 128u);
    if (mode >= 2) {
      // SWIZZLE_128B K-major: rows of 128 B (64 bf16 of K), 8-row atoms of 1024 B: SBO = 1024, LBO unused (1), layout type 2
      ad = make_desc(smem_u32(smem_raw), 16u, 1024u) | ((uint64_t)2 << 61);
      bd = make_desc(smem_u32(smem_raw) + 32768u, 16u, 1024u) | ((uint64_t)2 << 61);
    }
    if (mode == 4 || mode == 5) {
      // conv-like: LBO = 544 rows * 16 B, A start = (34 + 17 * tap_offset) rows (+ second K chunk pair), weight slab per tap
      const uint32_t pa = 544u * 16u;
      ad = make_desc(smem_u32(smem_raw), pa, 128u);
      bd = make_desc(smem_u32(smem_raw) + 40960u, b_plane, 128u);
    }
    if (mode >= 10) {
      // straight-line: 16 descriptor pairs prepared BEFORE the timed region, then 16 back-to-back MMAs (x repeats)
      uint64_t av[16], bv[16];
      const uint32_t pa = 544u * 16u;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int tap = (k >> 1) % 9, ks = k & 1;
        if (mode == 10) { av[k] = ad; bv[k] = bd; }
        else if (mode == 11) {   // conv-like, unaligned row offsets
          av[k] = desc_advance(make_desc(smem_u32(smem_raw), pa, 128u), (uint32_t)(34 + 17 * (tap / 2)) * 16u + (uint32_t)ks * 2u * pa);
          bv[k] = desc_advance(bd, (uint32_t)(tap * 4 + ks * 2) * 512u);
        } else {                 // conv-like, 128-byte aligned row offsets
          av[k] = desc_advance(make_desc(smem_u32(smem_raw), pa, 128u), (uint32_t)(32 + 16 * (tap / 2)) * 16u + (uint32_t)ks * 2u * pa);
          bv[k] = desc_advance(bd, (uint32_t)(tap * 4 + ks * 2) * 512u);
        }
      }
      const long long t0 = clock64();
      for (int rep = 0; rep < n_mma / 16; ++rep) {
#pragma unroll
        for (int k = 0; k < 16; ++k) umma_bf16(tmem_base, av[k], bv[k], idesc, 1u);
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      while (!mbar_try_wait(&bar, 0)) {
      }
      const long long t2 = clock64();
      out[0] = t1 - t0;
      out[1] = t2 - t0;
    } else {
    const long long t0 = clock64();
    for (int k = 0; k < n_mma; ++k) {
      uint64_t a_k = ad, b_k = bd;
      if (mode == 1 || mode == 3) a_k = desc_advance(ad, (uint32_t)(k & 7) * 2048u);   // a different A tile every MMA
      if (mode == 4) {
        const int tap = (k >> 1) % 9, ks = k & 1;
        a_k = desc_advance(ad, (uint32_t)(34 + 17 * (tap / 2)) * 16u + (uint32_t)ks * 2u * 544u * 16u);
        b_k = desc_advance(bd, (uint32_t)(tap * 4 + ks * 2) * 512u);
      }
      if (mode == 5) {     // same as 4 but 128-byte aligned row offsets (multiples of 8 rows)
        const int tap = (k >> 1) % 9, ks = k & 1;
        a_k = desc_advance(ad, (uint32_t)(32 + 16 * (tap / 2)) * 16u + (uint32_t)ks * 2u * 544u * 16u);
        b_k = desc_advance(bd, (uint32_t)(tap * 4 + ks * 2) * 512u);
      }
      umma_bf16(tmem_base, a_k, b_k, idesc, k > 0);
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    while (!mbar_try_wait(&bar, 0)) {
    }
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}
}  // namespace
}  // namespace sf

extern "C" int sfdbg_umma_timing(int N, int n_mma, int shift, int mode, long long* out_host) {
  long long* d = nullptr;
  if (cudaMalloc(&d, 16) != cudaSuccess) return -1;
  cudaFuncSetAttribute(sf::umma_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  sf::umma_timing_kernel<<<1, 128, 64 * 1024>>>(N, n_mma, shift, mode, d);
  if (cudaDeviceSynchronize() != cudaSuccess) return -2;
  cudaMemcpy(out_host, d, 16, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// TMEM -> register read bandwidth (tcgen05.ld.32x32b): `warps` warps (warp w reads lane quarter w & 3) each issue `n_ld` loads
// of COLS columns, PW loads in flight per tcgen05.wait::ld; mode 1 also converts every value to fp16 and stores it to shared
// memory (what a conversion stage of the tokenizer / transformer epilogues does).  out[0] = cycles of the slowest warp,
// out[1] = bytes read.  Every accumulator element of both tensor-core kernels moves through this path once per GEMM, so
// bytes / cycles here bounds their epilogues.
namespace sf {
namespace {
template <int COLS, int PW, int MODE>
__global__ void __launch_bounds__(544, 1) tmem_ld_timing_kernel(int n_ld, int n_mma, int mma_n, long long* out) {
  extern __shared__ __align__(128) unsigned char sbuf[];     // mode 1: [col / 8][128 rows][16 B] planar-chunk buffer, 512 columns
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t bar;
  __shared__ long long t_end[17];
  using namespace tc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_conv = (int)(blockDim.x >> 5) - (n_mma > 0 ? 1 : 0);       // with n_mma > 0 the LAST warp issues MMAs meanwhile
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == n_conv) {
    // n_mma back-to-back 128 x mma_n x 16 MMAs on no-swizzle operands inside the buffer the other warps write (garbage data),
    // accumulating into TMEM columns [256, 256 + mma_n): tensor-pipe operand reads from shared memory + accumulator writes
    const long long t0 = clock64();
    if (elect_one()) {
      const uint64_t ad = make_desc(smem_u32(sbuf), 2048u, 128u), bd = make_desc(smem_u32(sbuf) + 8192u, (uint32_t)mma_n * 16u, 128u);
      const uint32_t idesc = make_idesc(128, mma_n, false, true, true);
      for (int k = 0; k < n_mma; ++k) umma_bf16(tmem_base_s + 256u, ad, bd, idesc, 1u);
      umma_commit(&bar);
      while (!mbar_try_wait(&bar, 0)) {
      }
    }
    __syncwarp();
    if (lane == 0) t_end[16] = clock64() - t0;
  }
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  const int row = (warp & 3) * 32 + lane;
  float acc = 0.f;
  const long long t0 = clock64();
  const uint32_t cmask = n_mma > 0 ? 255u : 511u;                       // stay clear of the MMA's accumulator columns
  uint32_t col = (uint32_t)((warp >> 2) * COLS * PW) & cmask;
  if (warp < n_conv)
  for (int i = 0; i < n_ld; i += PW) {
    float v[PW][COLS];
#pragma unroll
    for (int j = 0; j < PW; ++j) {
      const uint32_t c = (col + (uint32_t)(j * COLS)) & cmask;
      if (COLS == 16) tmem_ld16(base + c, v[j]);
      else tmem_ld32(base + c, v[j]);
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < PW; ++j) {
      if (MODE == 1) {
        const uint32_t c = (col + (uint32_t)(j * COLS)) & cmask;
#pragma unroll
        for (int q = 0; q < COLS; q += 8)
          *reinterpret_cast<uint4*>(sbuf + (size_t)((c + q) >> 3) * 2048 + row * 16) =
              make_uint4(pack_f16x2(v[j][q], v[j][q + 1]), pack_f16x2(v[j][q + 2], v[j][q + 3]), pack_f16x2(v[j][q + 4], v[j][q + 5]),
                         pack_f16x2(v[j][q + 6], v[j][q + 7]));
      } else {
        acc += v[j][0] + v[j][COLS - 1];
      }
    }
    col = (col + (uint32_t)(COLS * PW * 4)) & cmask;
  }
  const long long t1 = clock64();
  if (lane == 0 && warp < n_conv) t_end[warp] = t1 - t0;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long mx = 0;
    for (int w = 0; w < n_conv; ++w) mx = t_end[w] > mx ? t_end[w] : mx;
    out[0] = mx;
    out[1] = (long long)n_conv * n_ld * COLS * 32 * 4;
    out[2] = (long long)(acc == 12345.f);      // keeps the loads alive
    out[3] = n_mma > 0 ? t_end[16] : 0;        // cycles of the MMA stream (issue -> completion)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}
template <int COLS, int PW, int MODE>
int run_tmem_ld(int warps, int n_ld, int n_mma, int mma_n, long long* d) {
  const int smem = 64 * 2048;
  cudaFuncSetAttribute(tmem_ld_timing_kernel<COLS, PW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  tmem_ld_timing_kernel<COLS, PW, MODE><<<1, (warps + (n_mma > 0 ? 1 : 0)) * 32, smem>>>(n_ld / PW * PW, n_mma, mma_n, d);
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : -2;
}
}  // namespace
}  // namespace sf

extern "C" int sfdbg_tmem_ld_timing(int warps, int n_ld, int cols, int per_wait, int mode, int n_mma, int mma_n, long long* out_host) {
  if (warps < 1 || warps > 16 || n_ld < 4 || n_mma < 0 || (n_mma > 0 && (mma_n < 16 || mma_n > 256 || mma_n % 16))) return -3;
  long long* d = nullptr;
  if (cudaMalloc(&d, 32) != cudaSuccess) return -1;
  int rc = -3;
#define SF_CASE(C_, P_, M_) if (cols == C_ && per_wait == P_ && mode == M_) rc = sf::run_tmem_ld<C_, P_, M_>(warps, n_ld, n_mma, mma_n, d);
  SF_CASE(16, 1, 0) SF_CASE(16, 2, 0) SF_CASE(16, 4, 0) SF_CASE(32, 1, 0) SF_CASE(32, 2, 0)
  SF_CASE(16, 1, 1) SF_CASE(16, 2, 1) SF_CASE(32, 1, 1)
#undef SF_CASE
  if (rc == 0) cudaMemcpy(out_host, d, 32, cudaMemcpyDeviceToHost);
  cudaFree(d);
  return rc;
}
