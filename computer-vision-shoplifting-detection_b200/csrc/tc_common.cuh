// Hand-written sm_100a tensor-core plumbing: tcgen05.mma / TMEM / mbarrier PTX wrappers and the
// shared-memory matrix descriptors for the NO-SWIZZLE ("interleaved") canonical layouts.
//
// Operand layout used everywhere in this library (bf16, 16-byte granules):
//
//     buf[chunk j = col/8][row r][8 cols]          byte address = j*plane + r*16 + (col%8)*2
//
// i.e. one 16-byte granule per (row, 8-column chunk), granules of one chunk contiguous over rows.
// The same buffer is
//   * a K-major  A/B operand when the columns are the contraction (K) axis and rows are M/N:
//       8 rows x 16 B = one 128-byte core matrix, SBO (next 8 rows) = 128 B, LBO (next K chunk) = plane
//   * an MN-major B operand when the ROWS are the contraction axis and columns are N:
//       core matrix = 8 k-rows x 16 B, LBO (next 8 k-rows) = 128 B, SBO (next N chunk) = plane
// and any row offset (multiple of 16 B) is a valid operand start, which is what lets the 9 taps of the
// temporal convolution be nine shifted views of one activation buffer.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace sf {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---------------------------------------------------------------- descriptors
// 64-bit shared-memory matrix descriptor (PTX ISA "tcgen05 matrix descriptor"):
//   [0,14) start address >> 4   [16,30) leading-dim byte offset >> 4   [32,46) stride-dim byte offset >> 4
//   [46,48) version = 1 (Blackwell)   [49,52) base offset = 0   [61,64) swizzle mode = 0 (none)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// advance the start-address field by `bytes` (must keep it inside the 14-bit field)
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }

// 32-bit instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate:
//   [4,6) D format = 1 (f32)  [7,10) A format = 1 (bf16)  [10,13) B format = 1 (bf16)
//   [15] A major (0 = K)  [16] B major (0 = K, 1 = MN)  [17,23) N >> 3   [24,29) M >> 4
// A / B formats are independent: 0 = fp16, 1 = bf16 (activations are bf16 for range; constant operands may be fp16).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool b_mn_major, bool a_f16 = false, bool b_f16 = false) {
  return (1u << 4) | ((a_f16 ? 0u : 1u) << 7) | ((b_f16 ? 0u : 1u) << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// one lane of a converged warp (ptxas keeps code predicated on elect.sync in the uniform datapath)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// suspend-time hint of try_wait: the thread sleeps until the phase completes or this many nanoseconds pass, instead of the
// (short) system default after which the wait loop spins through issue slots the working warps need
#ifndef SF_MBAR_SUSPEND_NS
#define SF_MBAR_SUSPEND_NS 100000
#endif
constexpr uint32_t kMbarSuspendNs = SF_MBAR_SUSPEND_NS;
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(kMbarSuspendNs)
      : "memory");
  return ok != 0;
}
// SF_MBAR_SLEEP_NS > 0: a failed try_wait backs off with nanosleep before polling again (experiment switch: a warp that
// spins in the try_wait loop takes issue slots from the working warps of its scheduler)
#ifndef SF_MBAR_SLEEP_NS
#define SF_MBAR_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
    if (SF_MBAR_SLEEP_NS > 0) __nanosleep(SF_MBAR_SLEEP_NS);
  }
}

// ---------------------------------------------------------------- TMA (bulk async copy, 1-D)
// One thread arms the barrier with the byte count and issues the copies; completion flips the barrier phase.
// dst / src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- TMEM
// One full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *dst (smem).
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy smem writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- MMA issue (one thread)
// D[tmem] (+)= A[smem] * B[smem];  M = 128, K = 16 per instruction, N from the instruction descriptor.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` once every previously issued MMA of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------- TMEM -> registers
// Thread i of warp w reads TMEM lane 32*(w%4)+i, 16 / 32 consecutive fp32 columns starting at taddr's column.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- bf16 packing
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// fp16 operands: saturating (an activation beyond +-65504 becomes the largest finite half instead of inf, so no NaN
// can come out of the next MMA; weights are range-checked when the model is packed)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_op2(float lo, float hi) { return F16 ? pack_f16x2(lo, hi) : pack_bf16x2(lo, hi); }
// element (row, col) of the planar-chunk layout, in bf16 elements from the buffer base
__device__ __forceinline__ uint32_t pc_index(int row, int col, int plane_rows) {
  return (uint32_t)((col >> 3) * plane_rows + row) * 8u + (uint32_t)(col & 7);
}

}  // namespace tc
}  // namespace sf
