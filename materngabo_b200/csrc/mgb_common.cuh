// Shared helpers for libmgb.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <atomic>
#include "../../include/mgb.h"

namespace mgb {

extern thread_local char g_err[512];
extern std::atomic<long long> g_launches;

inline int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return MGB_OK;
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
  return MGB_ERR_CUDA;
}

inline int after_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}

// SM count of the CURRENT device, cached per device (a process may drive several GPUs through the _host entry points)
inline int device_sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  int v = cache[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cache[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

#define MGB_TRY(expr)                 \
  do {                                \
    int _rc = (expr);                 \
    if (_rc != MGB_OK) return _rc;    \
  } while (0)

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming (evict-first) stores for write-once outputs
__device__ __forceinline__ void st_stream(double* p, double a) { __stcs(p, a); }
__device__ __forceinline__ void st_stream2(double* p, double a, double b) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Branch-free exp for the quadrature inner phases (arguments <= 0 up to rounding; any |x| < 1e9 is handled).
// Same scheme as the CUDA library's fast path - round x log2(e) with the 1.5 * 2^52 trick, two-constant reduction,
// degree-11 polynomial (Chebyshev-node interpolant on |r| <= ln2 / 2, approximation error 1.5e-17) - but the scaling
// by 2^n is done without the library's out-of-range branch: the exponent is added in two pieces (n1 >= -1000 into the
// exponent field, the rest as a power-of-two factor that reaches 0 for results below the denormal range).  Without the
// branch the compiler interleaves the independent exponentials of an unrolled loop, which is what hides the FP64
// latency at 8-16 warps per SM.
__device__ __forceinline__ double exp_bf(double x) {
  double t = fma(x, 1.4426950408889634, 6755399441055744.0);
  const int n = __double2loint(t);
  t -= 6755399441055744.0;
  double r = fma(t, -6.9314718055994529e-1, x);
  r = fma(t, -2.3190468138462996e-17, r);
  double q = 2.5110095611886237e-08;
  q = fma(q, r, 2.7632715071632255e-07);
  q = fma(q, r, 2.7557240761739257e-06);
  q = fma(q, r, 2.4801485278370062e-05);
  q = fma(q, r, 0.00019841269890193705);
  q = fma(q, r, 0.0013888888952505406);
  q = fma(q, r, 0.008333333333319546);
  q = fma(q, r, 0.04166666666648738);
  q = fma(q, r, 0.1666666666666668);
  q = fma(q, r, 0.5000000000000019);
  q = fma(q, r, 1.0);
  q = fma(q, r, 1.0);
  const int n1 = max(n, -1000);
  const int n2 = max(n - n1, -1023);
  const double v = __hiloint2double(__double2hiint(q) + (n1 << 20), __double2loint(q));
  return v * __hiloint2double((n2 + 1023) << 20, 0);
}

}  // namespace mgb
