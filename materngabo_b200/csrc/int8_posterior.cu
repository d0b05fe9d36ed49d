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
// L^-1 is lower triangular: its rows are split into NB column blocks of the product (rows j in [J nb, (J + 1) nb) only
// meet k < (J + 1) nb), and each block gets its own compacted copy of the operands ([A_0[:, :kJ] | ... | A_6[:, :kJ]],
// kJ = (J + 1) nb), so that the prefix / suffix trick works per block: 7 NB GEMMs with 60 % of the dense work at NB = 5.
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

// Column blocks of the triangular operand: NB = 5 from 2048 training points on, Npad a multiple of 256 NB.
struct I8Geom {
  int NB;
  long long np, nb;                       // padded n, block width
  __host__ __device__ long long kJ(int J) const { return (J + 1) * nb; }
  // byte offset of block J's operand copy in a buffer of `rows`-row copies: sum_{I < J} rows * S * kI
  __host__ __device__ long long a_off(int J, long long rows) const { return rows * I8_SLICES * nb * ((long long)J * (J + 1) / 2); }
  __host__ __device__ long long a_bytes(long long rows) const { return a_off(NB, rows); }
  // B copies hold only the nb rows of their block
  __host__ __device__ long long b_off(int J) const { return nb * I8_SLICES * nb * ((long long)J * (J + 1) / 2); }
  __host__ __device__ long long b_bytes() const { return b_off(NB); }
};

__host__ __device__ inline I8Geom i8_geom(long long n) {
  I8Geom g;
  g.NB = (n >= 2048) ? 5 : 1;      // 5 blocks: 60 % of the dense work, block width 1024 = 4 GEMM tiles at n = 5000
  g.np = i8_round_up(n, (long long)I8_PAD * g.NB);
  g.nb = g.np / g.NB;
  return g;
}

// One warp per row: e = floor(log2 max|row|) + 2 (so |row| 2^-e <= 1/2), then S round-to-nearest slices of 7 bits.
// A operand (is_b = 0): the row goes into EVERY block copy J >= column block, at [row][p kJ + c].
// B operand (is_b = 1, L^-1): row j belongs to block J = j / nb only, slices reversed: [j - J nb][(S - 1 - p) kJ + c];
// entries right of the diagonal are zero by construction and are written as zeros.
// For the A operand the same pass over the row also forms mean_i = mean_const + sum_j K*_ij alpha_j (alpha != nullptr).
__global__ void __launch_bounds__(256) i8_slice_rows_kernel(const double* src, long long ld, long long rows, int cols, I8Geom g,
                                                            int is_b, signed char* dst, int* exps, const double* alpha,
                                                            double mean_const, double* mean) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const double* s = src + row * ld;
  const int used = is_b ? (int)min((long long)cols, row + 1) : cols;      // lower triangle of L^-1
  double mx = 0.0, mu = 0.0;
  if (alpha != nullptr) {
    for (int c = lane; c < used; c += 32) {
      const double v = s[c];
      mx = fmax(mx, fabs(v));
      mu = fma(v, alpha[c], mu);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mu += __shfl_xor_sync(0xffffffffu, mu, o);
    if (lane == 0) mean[row] = mean_const + mu;
  } else {
    for (int c = lane; c < used; c += 32) mx = fmax(mx, fabs(s[c]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e = 0;
  if (mx > 0.0) {
    (void)frexp(mx, &e);      // mx = f 2^e, f in [1/2, 1)
    e += 1;                   // |row| 2^-e < 1/2
  }
  if (lane == 0) exps[row] = e;
  const int Jrow = is_b ? (int)(row / g.nb) : 0;
  const long long kmax = is_b ? g.kJ(Jrow) : g.np;
  // sixteen columns per lane and iteration: one 16-byte store per slice and operand copy (nb is a multiple of 256, so the
  // sixteen columns lie in one column block)
  for (long long c0 = 16 * lane; c0 < kmax; c0 += 512) {
    double r[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) r[t] = (c0 + t < used) ? ldexp(s[c0 + t], -e) : 0.0;
    const int kb = (int)(c0 / g.nb);
#pragma unroll 1
    for (int p = 0; p < I8_SLICES; ++p) {
      uint4 q;
      signed char* qq = reinterpret_cast<signed char*>(&q);
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const double x = r[t] * (double)(1 << I8_BITS);
        const double qi = rint(x);
        qq[t] = (signed char)(int)qi;
        r[t] = x - qi;
      }
      if (is_b) {
        const long long k = g.kJ(Jrow);
        *reinterpret_cast<uint4*>(dst + g.b_off(Jrow) + (row - Jrow * g.nb) * (I8_SLICES * k) + (long long)(I8_SLICES - 1 - p) * k + c0) = q;
      } else {
        for (int J = kb; J < g.NB; ++J) {
          const long long k = g.kJ(J);
          *reinterpret_cast<uint4*>(dst + g.a_off(J, rows) + row * (I8_SLICES * k) + (long long)p * k + c0) = q;
        }
      }
    }
  }
}

// One warp per candidate: v_ij = 2^(ea_i + eb_j) sum_d 2^(-7 (d + 2)) C_d[i][j], var_i = kss - sum_j v_ij^2.
__global__ void __launch_bounds__(256) i8_combine_kernel(const int* C, long long m, int n, int Npad, const int* ea,
                                                         const int* eb, double kss, double* var) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= m) return;
  const long long plane = m * (long long)Npad;
  const int* c = C + row * (long long)Npad;
  const int ei = ea[row];
  double wgt[I8_SLICES];
#pragma unroll
  for (int d = 0; d < I8_SLICES; ++d) wgt[d] = exp2((double)(-I8_BITS * (d + 2)));
  double ss = 0.0;
  // four columns per lane and iteration: 16-byte loads from each of the seven accumulator planes (Npad % 4 == 0)
  for (int j0 = 4 * lane; j0 < n; j0 += 128) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int d = I8_SLICES - 1; d >= 0; --d) {      // smallest weights first
      const int4 q = *reinterpret_cast<const int4*>(c + d * plane + j0);
      acc[0] = fma((double)q.x, wgt[d], acc[0]);
      acc[1] = fma((double)q.y, wgt[d], acc[1]);
      acc[2] = fma((double)q.z, wgt[d], acc[2]);
      acc[3] = fma((double)q.w, wgt[d], acc[3]);
    }
    const int4 e4 = *reinterpret_cast<const int4*>(eb + j0);
    const int ee[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (j0 + t < n) {
        const double v = ldexp(acc[t], ei + ee[t]);
        ss = fma(v, v, ss);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) var[row] = kss - ss;
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
  const I8Geom g = i8_geom(n);
  return 4 * g.np + g.b_bytes();
}

extern "C" int mgb_posterior_int8_pack(const double* rinv, long long n, long long ld, void* packed, void* stream) {
  if (!rinv || !packed || n < 1 || ld < n) return fail(MGB_ERR_ARG, "posterior_int8_pack: bad argument");
  if (n > 32768) return fail(MGB_ERR_ARG, "posterior_int8_pack: n too large for exact int32 accumulation");
  const I8Geom g = i8_geom(n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int* eb = static_cast<int*>(packed);
  signed char* B = static_cast<signed char*>(packed) + 4 * g.np;
  MGB_TRY(check_cuda(cudaMemsetAsync(packed, 0, (size_t)mgb_posterior_int8_pack_bytes(n), s), "cudaMemsetAsync"));
  i8_slice_rows_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(rinv, ld, n, (int)n, g, 1, B, eb, nullptr, 0.0, nullptr);
  return after_launch("i8_slice_rows_kernel (pack)");
}

extern "C" long long mgb_posterior_int8_workspace_bytes(long long m, long long n) {
  if (m < 1 || n < 1) return 0;
  const I8Geom g = i8_geom(n);
  return i8_round_up(4 * m, 256) + i8_round_up(g.a_bytes(m), 256) + (long long)I8_SLICES * m * g.np * 4 + (long long)I8_GEMM_WS;
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
  const I8Geom g = i8_geom(n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  char* w = static_cast<char*>(workspace);
  int* ea = reinterpret_cast<int*>(w);
  w += i8_round_up(4 * m, 256);
  signed char* A = reinterpret_cast<signed char*>(w);
  w += i8_round_up(g.a_bytes(m), 256);
  int* C = reinterpret_cast<int*>(w);
  w += (long long)I8_SLICES * m * g.np * 4;
  void* gws = w;
  const int* eb = static_cast<const int*>(packed);
  const signed char* B = static_cast<const signed char*>(packed) + 4 * g.np;
  i8_slice_rows_kernel<<<(unsigned)((m + 7) / 8), 256, 0, s>>>(kstar, ldk, m, (int)n, g, 0, A, ea, alpha, mean_const, mean);
  MGB_TRY(after_launch("i8_slice_rows_kernel"));
  for (int J = 0; J < g.NB; ++J) {
    const long long k = g.kJ(J), ldab = (long long)I8_SLICES * k;
    const signed char* AJ = A + g.a_off(J, m);
    const signed char* BJ = B + g.b_off(J);
    for (int d = 0; d < I8_SLICES; ++d) {
      // pairs (p, d - p), p = 0..d: A slices 0..d against B slices d..0 = columns [(S-1-d) kJ, S kJ) of the reversed B
      MGB_TRY(i8_gemm(AJ, ldab, BJ + (long long)(I8_SLICES - 1 - d) * k, ldab, C + (long long)d * m * g.np + J * g.nb, g.np,
                      (int)m, (int)g.nb, (int)((d + 1) * k), gws, I8_GEMM_WS, s));
    }
  }
  i8_combine_kernel<<<(unsigned)((m + 7) / 8), 256, 0, s>>>(C, m, (int)n, (int)g.np, ea, eb, kss, var);
  return after_launch("i8_combine_kernel");
#endif
}
