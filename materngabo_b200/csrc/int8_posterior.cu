// Posterior variance through the INT8 tensor cores: error-free slicing of the N^2 contraction (opt-in path).
//
//   V = K* L^-T  (m x n),   var_i = kss - sum_j V_ij^2,   mean_i = mean_const + sum_j K*_ij alpha_j
//
// The DMMA kernel (posterior.cu) runs this contraction at 92 % of the FP64 tensor peak; the only way past that roof
// is another pipe.  Here every row of K* and every row of L^-1 is scaled by a power of two and cut into S = 7 signed
// 7-bit slices (round to nearest: |q| <= 64, remainders zero-mean),
//        a_ik = 2^ea_i sum_p A_p[i][k] 2^(-7 (p + 1)),     b_jk = 2^eb_j sum_q B_q[j][k] 2^(-7 (q + 1)),
// so that  V_ij = 2^(ea_i + eb_j) sum_d 2^(-7 (d + 2)) C_d[i][j],   C_d = sum_{p + q = d} A_p B_q^T   (d <= S - 1),
// with C_d an INTEGER matrix product (exact in int32: 7 * 5000 * 64^2 < 2^31).  The slices are stored side by side,
// A as [A_0 | ... | A_6] and B REVERSED as [B_6 | ... | B_0]: anti-diagonal d is then ONE int8 GEMM with
// K' = (d + 1) Kpad over a prefix of A and a suffix of B - 7 GEMMs (28 slice products) per candidate block.
// Dropped pairs (p + q >= 7) and slice remainders leave 8.5e-12 of max var at n = 5000 (tools/ozaki_prototype.py),
// inside the 1e-10 parity bar.
//
// The GEMM itself is a LIBRARY kernel (CUTLASS collective builder from the header tree vendored in the image,
// tcgen05 kind::i8, 2-SM 256 x 256 tiles: tools/int8_probe measures 3.1 POP/s on this shape); slicing and the
// recombination + sum of squares are this file's kernels.  Built only when the CUTLASS headers are found
// (materngabo_b200/build.py defines MGB_HAVE_CUTLASS); otherwise the entry points report "not available".
#include "mgb_common.cuh"

#ifdef MGB_HAVE_CUTLASS
#include "cutlass/cutlass.h"
#include "cute/tensor.hpp"
#include "cutlass/gemm/device/gemm_universal_adapter.h"
#include "cutlass/gemm/kernel/gemm_universal.hpp"
#include "cutlass/gemm/collective/collective_builder.hpp"
#include "cutlass/epilogue/collective/collective_builder.hpp"
#endif

namespace mgb {

constexpr int I8_SLICES = 7;
constexpr int I8_BITS = 7;
constexpr int I8_PAD = 256;

__host__ __device__ inline long long i8_round_up(long long v, long long m) { return (v + m - 1) / m * m; }

// One warp per row: e = floor(log2 max|row|) + 2 (so |row| 2^-e <= 1/2), then S round-to-nearest slices of 7 bits.
// dst row = [slice 0 | slice 1 | ... ] (reversed = 0) or [slice S-1 | ... | slice 0] (reversed = 1), each Kpad wide,
// zero beyond `cols`.
__global__ void __launch_bounds__(256) i8_slice_rows_kernel(const double* src, long long ld, long long rows, int cols,
                                                            int Kpad, int reversed, signed char* dst, int* exps) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const double* s = src + row * ld;
  double mx = 0.0;
  for (int c = lane; c < cols; c += 32) mx = fmax(mx, fabs(s[c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e = 0;
  if (mx > 0.0) {
    (void)frexp(mx, &e);      // mx = f 2^e, f in [1/2, 1)
    e += 1;                   // |row| 2^-e < 1/2
  }
  if (lane == 0) exps[row] = e;
  signed char* d = dst + row * (long long)(I8_SLICES * Kpad);
  for (int c0 = 4 * lane; c0 < Kpad; c0 += 128) {      // four columns per lane: one 32-bit store per slice
    double r[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) r[t] = (c0 + t < cols) ? ldexp(s[c0 + t], -e) : 0.0;
#pragma unroll
    for (int p = 0; p < I8_SLICES; ++p) {
      char4 q;
      signed char* qq = reinterpret_cast<signed char*>(&q);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double x = r[t] * (double)(1 << I8_BITS);
        const double qi = rint(x);
        qq[t] = (signed char)(int)qi;
        r[t] = x - qi;
      }
      const int slot = reversed ? (I8_SLICES - 1 - p) : p;
      *reinterpret_cast<char4*>(d + (long long)slot * Kpad + c0) = q;
    }
  }
}

// One warp per candidate: v_ij = 2^(ea_i + eb_j) sum_d 2^(-7 (d + 2)) C_d[i][j], var_i = kss - sum_j v_ij^2,
// mean_i = mean_const + sum_j K*_ij alpha_j.
__global__ void __launch_bounds__(256) i8_combine_kernel(const int* C, long long m, int n, int Npad, const int* ea,
                                                         const int* eb, const double* kstar, long long ldk,
                                                         const double* alpha, double kss, double mean_const, double* mean,
                                                         double* var) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= m) return;
  const long long plane = m * (long long)Npad;
  const int* c = C + row * (long long)Npad;
  const double* kr = kstar + row * ldk;
  const int ei = ea[row];
  double ss = 0.0, mu = 0.0;
  for (int j = lane; j < n; j += 32) {
    double acc = 0.0;
#pragma unroll
    for (int d = I8_SLICES - 1; d >= 0; --d)      // smallest weights first
      acc = fma((double)c[d * plane + j], exp2((double)(-I8_BITS * (d + 2))), acc);
    const double v = ldexp(acc, ei + eb[j]);
    ss = fma(v, v, ss);
    mu = fma(kr[j], alpha[j], mu);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    mu += __shfl_xor_sync(0xffffffffu, mu, o);
  }
  if (lane == 0) {
    var[row] = kss - ss;
    mean[row] = mean_const + mu;
  }
}

#ifdef MGB_HAVE_CUTLASS
using namespace cute;
using I8TileShape = Shape<_256, _256, _128>;
using I8ClusterShape = Shape<_2, _1, _1>;
using I8Epilogue = typename cutlass::epilogue::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, I8TileShape, I8ClusterShape,
    cutlass::epilogue::collective::EpilogueTileAuto, int32_t, int32_t, int32_t, cutlass::layout::RowMajor, 4, int32_t,
    cutlass::layout::RowMajor, 4, cutlass::epilogue::collective::EpilogueScheduleAuto>::CollectiveOp;
using I8Mainloop = typename cutlass::gemm::collective::CollectiveBuilder<
    cutlass::arch::Sm100, cutlass::arch::OpClassTensorOp, int8_t, cutlass::layout::RowMajor, 16, int8_t,
    cutlass::layout::ColumnMajor, 16, int32_t, I8TileShape, I8ClusterShape,
    cutlass::gemm::collective::StageCountAutoCarveout<static_cast<int>(sizeof(typename I8Epilogue::SharedStorage))>,
    cutlass::gemm::collective::KernelScheduleAuto>::CollectiveOp;
using I8Kernel = cutlass::gemm::kernel::GemmUniversal<Shape<int, int, int, int>, I8Mainloop, I8Epilogue, void>;
using I8Gemm = cutlass::gemm::device::GemmUniversalAdapter<I8Kernel>;

// C (M x N int32, row-major, ldc) = A (M x K int8, row-major, lda) * B^T (B: N x K int8, row-major, ldb)
static int i8_gemm(const signed char* A, long long lda, const signed char* B, long long ldb, int* C, long long ldc, int M, int N,
                   int K, void* ws, size_t ws_bytes, cudaStream_t s) {
  using StrideA = typename I8Gemm::GemmKernel::StrideA;
  using StrideB = typename I8Gemm::GemmKernel::StrideB;
  using StrideC = typename I8Gemm::GemmKernel::StrideC;
  using StrideD = typename I8Gemm::GemmKernel::StrideD;
  StrideA sa{};
  StrideB sb{};
  StrideC sc{};
  StrideD sd{};
  get<0>(sa) = lda;
  get<0>(sb) = ldb;
  get<0>(sc) = ldc;
  get<0>(sd) = ldc;
  typename I8Gemm::Arguments args{cutlass::gemm::GemmUniversalMode::kGemm,
                                  {M, N, K, 1},
                                  {reinterpret_cast<const int8_t*>(A), sa, reinterpret_cast<const int8_t*>(B), sb},
                                  {{1, 0}, C, sc, C, sd}};
  I8Gemm gemm;
  if (gemm.can_implement(args) != cutlass::Status::kSuccess) return fail(MGB_ERR_ARG, "int8 gemm: shape not supported");
  if (I8Gemm::get_workspace_size(args) > ws_bytes) return fail(MGB_ERR_ARG, "int8 gemm: workspace too small");
  if (gemm.initialize(args, ws, s) != cutlass::Status::kSuccess) return fail(MGB_ERR_CUDA, "int8 gemm: initialize failed");
  if (gemm.run(s) != cutlass::Status::kSuccess) return fail(MGB_ERR_CUDA, "int8 gemm: launch failed");
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MGB_OK;
}
#endif

constexpr size_t I8_GEMM_WS = 1 << 20;

}  // namespace mgb

using namespace mgb;

extern "C" int mgb_int8_available(void) {
#ifdef MGB_HAVE_CUTLASS
  return 1;
#else
  return 0;
#endif
}

extern "C" long long mgb_posterior_int8_pack_bytes(long long n) {
  if (n < 1) return 0;
  const long long np = i8_round_up(n, I8_PAD);
  return 4 * np + np * (long long)I8_SLICES * np;
}

extern "C" int mgb_posterior_int8_pack(const double* rinv, long long n, long long ld, void* packed, void* stream) {
  if (!rinv || !packed || n < 1 || ld < n) return fail(MGB_ERR_ARG, "posterior_int8_pack: bad argument");
  if (n > 32768) return fail(MGB_ERR_ARG, "posterior_int8_pack: n too large for exact int32 accumulation");
  const long long np = i8_round_up(n, I8_PAD);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int* eb = static_cast<int*>(packed);
  signed char* B = static_cast<signed char*>(packed) + 4 * np;
  MGB_TRY(check_cuda(cudaMemsetAsync(packed, 0, (size_t)mgb_posterior_int8_pack_bytes(n), s), "cudaMemsetAsync"));
  i8_slice_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(rinv, ld, n, (int)n, (int)np, 1, B, eb);
  return after_launch("i8_slice_rows_kernel (pack)");
}

extern "C" long long mgb_posterior_int8_workspace_bytes(long long m, long long n) {
  if (m < 1 || n < 1) return 0;
  const long long np = i8_round_up(n, I8_PAD);
  return i8_round_up(4 * m, 256) + i8_round_up(m * (long long)I8_SLICES * np, 256) + (long long)I8_SLICES * m * np * 4 +
         (long long)I8_GEMM_WS;
}

extern "C" int mgb_posterior_int8_reduce(const double* kstar, long long m, long long n, long long ldk, const void* packed,
                                         const double* alpha, double kss, double mean_const, double* mean, double* var,
                                         void* workspace, long long workspace_bytes, void* stream) {
#ifndef MGB_HAVE_CUTLASS
  (void)kstar; (void)m; (void)n; (void)ldk; (void)packed; (void)alpha; (void)kss; (void)mean_const; (void)mean; (void)var;
  (void)workspace; (void)workspace_bytes; (void)stream;
  return fail(MGB_ERR_ARG, "posterior_int8_reduce: libmgb was built without the CUTLASS headers (mgb_int8_available() == 0)");
#else
  if (m == 0) return MGB_OK;
  if (!kstar || !packed || !alpha || !mean || !var || !workspace || m < 0 || n < 1 || ldk < n)
    return fail(MGB_ERR_ARG, "posterior_int8_reduce: bad argument");
  if (m > 0x7fffffffLL / 8 || n > 32768) return fail(MGB_ERR_ARG, "posterior_int8_reduce: block too large");
  if (workspace_bytes < mgb_posterior_int8_workspace_bytes(m, n)) return fail(MGB_ERR_ARG, "posterior_int8_reduce: workspace too small");
  const long long np = i8_round_up(n, I8_PAD);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  int* ea = reinterpret_cast<int*>(w);
  w += i8_round_up(4 * m, 256);
  signed char* A = reinterpret_cast<signed char*>(w);
  w += i8_round_up(m * (long long)I8_SLICES * np, 256);
  int* C = reinterpret_cast<int*>(w);
  w += (long long)I8_SLICES * m * np * 4;
  void* gws = w;
  const int* eb = static_cast<const int*>(packed);
  const signed char* B = static_cast<const signed char*>(packed) + 4 * np;
  i8_slice_rows_kernel<<<(unsigned)((m + 7) / 8), 256, 0, s>>>(kstar, ldk, m, (int)n, (int)np, 0, A, ea);
  MGB_TRY(after_launch("i8_slice_rows_kernel"));
  const long long lda = (long long)I8_SLICES * np;
  for (int d = 0; d < I8_SLICES; ++d) {
    // pairs (p, d - p), p = 0..d: A slices 0..d against B slices d..0 = columns [(S-1-d) np, S np) of the reversed B
    MGB_TRY(i8_gemm(A, lda, B + (long long)(I8_SLICES - 1 - d) * np, lda, C + (long long)d * m * np, np, (int)m, (int)np,
                    (int)((d + 1) * np), gws, I8_GEMM_WS, s));
  }
  i8_combine_kernel<<<(unsigned)((m + 7) / 8), 256, 0, s>>>(C, m, (int)n, (int)np, ea, eb, kstar, ldk, alpha, kss, mean_const,
                                                           mean, var);
  return after_launch("i8_combine_kernel");
#endif
}
