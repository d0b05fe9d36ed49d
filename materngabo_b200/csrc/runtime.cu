// Library runtime: error slot, launch counter, FP64 peak micro-benchmarks (bench.py's roofline
// denominator: MEASURED_PEAKS.json has no FP64 figure, so it is measured in the same run).
#include "mgb_common.cuh"

namespace mgb {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

// 8 independent DFMA chains per thread, register resident.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed) {
  double a[8], b = seed, c = 1.0 - seed * 1e-9;
#pragma unroll
  for (int q = 0; q < 8; ++q) a[q] = seed * (q + 1) + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = fma(a[q], c, b);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += a[q];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // defeat dead-code elimination
}

// 8 independent DMMA m8n8k4 accumulator chains per warp.
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters, double seed) {
  double c0[8], c1[8];
  double a = seed + (threadIdx.x & 31) * 1e-3, b = 1.0 - seed * 1e-3;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    c0[q] = q * 1e-3;
    c1[q] = -q * 1e-3;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[q]), "+d"(c1[q])
                     : "d"(a), "d"(b));
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += c0[q] + c1[q];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA interleaved in one warp with equal FMA counts (32 DMMA = 8192 FMA per warp, 256 DFMA per thread):
// if the two instruction kinds executed on separate data paths the combined rate would approach the sum of the
// two peaks; on a shared FP64 data path it stays at one peak.
__global__ void __launch_bounds__(256) mixed_peak_kernel(double* out, int iters, double seed) {
  double c0[8], c1[8], f[8];
  double a = seed + (threadIdx.x & 31) * 1e-3, b = 1.0 - seed * 1e-3, c = 1.0 - seed * 1e-9;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    c0[q] = q * 1e-3;
    c1[q] = -q * 1e-3;
    f[q] = seed * (q + 1) + threadIdx.x * 1e-6;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[q]), "+d"(c1[q])
                     : "d"(a), "d"(b));
#pragma unroll
        for (int r2 = 0; r2 < 8; ++r2) f[r2] = fma(f[r2], c, b);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int q = 0; q < 8; ++q) s += c0[q] + c1[q] + f[q];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace mgb

using namespace mgb;

extern "C" const char* mgb_last_error(void) { return g_err; }
extern "C" int mgb_version(void) { return 202; }
extern "C" long long mgb_launch_count(void) { return g_launches.load(); }

extern "C" int mgb_fp64_peak(int kind, int repeats, double* tflops, double* elapsed_ms) {
  if (!tflops || (kind < 0 || kind > 2) || repeats < 1) return fail(MGB_ERR_ARG, "mgb_fp64_peak: bad argument");
  int dev = 0, sms = 0;
  MGB_TRY(check_cuda(cudaGetDevice(&dev), "cudaGetDevice"));
  MGB_TRY(check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute"));
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  MGB_TRY(check_cuda(cudaMalloc(&out, sizeof(double) * blocks * threads), "cudaMalloc"));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  int rc = MGB_OK;
  for (int r = 0; r < repeats + 1 && rc == MGB_OK; ++r) {  // first pass is the warm-up
    cudaEventRecord(e0, 0);
    if (kind == 0) dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.5);
    else if (kind == 1) dmma_peak_kernel<<<blocks, threads>>>(out, iters, 0.5);
    else mixed_peak_kernel<<<blocks, threads>>>(out, iters, 0.5);
    rc = after_launch("fp64 peak kernel");
    cudaEventRecord(e1, 0);
    if (rc == MGB_OK) rc = check_cuda(cudaEventSynchronize(e1), "cudaEventSynchronize");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (rc != MGB_OK) return rc;
  double flops;
  if (kind == 0) flops = 2.0 * 64.0 * iters * (double)blocks * threads;                 // 64 DFMA / thread / iter
  else flops = 2.0 * 256.0 * 32.0 * iters * (double)blocks * (threads / 32);            // 32 DMMA(8x8x4) / warp / iter
  if (kind == 2) flops *= 2.0;                                                          // + the same FMA count as DFMA
  *tflops = flops / (best * 1e-3) / 1e12;
  if (elapsed_ms) *elapsed_ms = best;
  return MGB_OK;
}
