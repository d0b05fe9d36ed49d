// tcgen05 / tensor-memory helpers shared by the INT8 tensor-core kernels (int8_fused.cu, series_i8.cu): shared-memory
// and instruction descriptors of kind::i8 with K-major 128-byte-swizzled operands, elected-lane MMA issue, commit,
// tensor-memory loads and bounded mbarrier spins.
#pragma once
#include "mgb_common.cuh"

namespace mgb {

constexpr int F8_BM = 128;                // rows per tile = UMMA M = TMEM lanes

// Byte offset of (row r, 16-byte chunk c) inside a K-major tile with the 128-byte swizzle: rows are 128 bytes, groups of
// eight rows are 1024 bytes, and the chunk index is XORed with the row index inside the group.
__device__ __forceinline__ int f8_swizzled(int r, int c16) { return (r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4); }

// ---- tcgen05 / tensor memory ---------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major, 128-byte swizzle: start address >> 4 in [0, 14), leading byte offset
// (unused for swizzled K-major, 1) in [16, 30), stride byte offset (eight rows = 1024 bytes) >> 4 in [32, 46),
// descriptor version 1 in [46, 48), layout type 2 (128-byte swizzle) in [61, 64).
__device__ __forceinline__ uint64_t f8_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor, kind::i8: D int32 (2 in [4, 6)), A and B signed 8-bit (1 in [7, 10) and [10, 13)), both
// K-major (0 in bits 15, 16), N >> 3 in [17, 23), M >> 4 in [24, 29).
__host__ __device__ constexpr uint32_t f8_idesc(int n) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(F8_BM >> 4) << 24);
}

// One lane of a converged warp (the issuing code stays warp-uniform, so the descriptors live in uniform registers).
__device__ __forceinline__ bool f8_elect() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// Upper word of the shared-memory descriptor (constant), lower word = (address >> 4) | leading byte offset 1.
constexpr uint32_t F8_DESC_HI = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t f8_desc_from_lo(uint32_t lo) { return ((uint64_t)F8_DESC_HI << 32) | lo; }
__device__ __forceinline__ uint32_t f8_desc_lo(uint32_t saddr) { return (saddr >> 4) | (1u << 16); }

__device__ __forceinline__ void f8_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  if (f8_elect()) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// mbarrier arrive once every tcgen05 operation issued so far by this thread has completed
__device__ __forceinline__ void f8_commit(uint64_t* bar) {
  if (f8_elect()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void f8_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void f8_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 16 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void f8_tmem_ld16(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// 8 consecutive columns of this thread's TMEM lane
__device__ __forceinline__ void f8_tmem_ld8(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void f8_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool f8_mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol error traps (launch failure reported to the host) instead of hanging the device.
__device__ __forceinline__ void f8_mbar_spin(uint64_t* bar, uint32_t parity) {
  for (unsigned spins = 0; !f8_mbar_test(bar, parity); ++spins) {
    if (spins > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void f8_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

}  // namespace mgb
