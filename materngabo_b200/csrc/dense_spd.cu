// Dense SPD toolkit for training sets beyond the single-CTA kernel of mll.cu (160 < n <= ~16k): in-place blocked
// Cholesky, triangular inverse, K^-1 = L^-T L^-1 and the exact-GP marginal likelihood / posterior fit built from them.
//
// What it replaces: the Cholesky / cholesky_solve / solve_triangular calls (cuSOLVER, cuBLAS trsm through torch.linalg)
// of ExactGPPosterior.fit at n = 5000 and of the marginal-likelihood step for 160 < n <= 1000 - i.e. gpytorch's
// ExactMarginalLogLikelihood + prediction-strategy caches (examples/bo_manifolds_benchmark_examples/bo_sphere.py:228,
// 257-265) at the sizes SURVEY.md section 8f rank 4 names.
//
//   chol_coop_kernel   right-looking, 32-column panels, ONE cooperative launch for the whole factorisation: per panel
//                      (1) every CTA factorises the 32 x 32 diagonal block redundantly in shared memory (elimination
//                      without pivot-column scaling: one CTA barrier per column; a first version kept the block in the
//                      registers of ONE warp and moved pivots by shuffles - latency bound at ~40 us per panel, ncu:
//                      profiles/r02_ncu_chol_v1.txt), then solves its share of the rows below (one thread per row, L11
//                      broadcast from shared memory); grid barrier; (2) rank-32 update of the trailing lower triangle
//                      in 64 x 64 tiles on the FP64 tensor cores; grid barrier.  2 grid barriers per 32 columns
//                      instead of a launch per column.
//   tri_inv_diag_kernel + dense_gemm_kernel
//                      X = L^-1 by recursive doubling: invert the 32 x 32 diagonal blocks (warp per block), then per
//                      level s = 32, 64, ...: X21 = -X22 (L21 X11) for all aligned block pairs at once (two batched
//                      GEMMs per level, k-ranges clipped to the triangles) - log2(n / 32) levels, all GEMM-shaped.
//   dense_gemm_kernel  64 x 64 x 16 tiles on the FP64 tensor cores (DMMA m8n8k4), generic strides (N / T operands), batched over
//                      gridDim.z, k-range clipped by the triangular structure of the operands; also K^-1 = X^T X.
//   glue               r = y - mean, z = X r, alpha = X^T z, G = s/2 (K^-1 - alpha alpha^T), reductions for dnll/ds,
//                      dnll/dnoise, dnll/dmean; dnll/dcoef through mgb_series_gram_bwd on G.
// FP64 throughout; accuracy is that of a standard Cholesky (backward error n eps |K|).
#include "mgb_common.cuh"
#include <cooperative_groups.h>
#include <cmath>

namespace cg = cooperative_groups;

namespace mgb {

constexpr int CB = 32;          // Cholesky panel width = warp width
constexpr int CHOL_THREADS = 256;
constexpr int UT = 64;          // trailing-update tile edge

// ---------------------------------------------------------------------------------------------------------------
// Cholesky
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_chol(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(CHOL_THREADS) chol_coop_kernel(double* __restrict__ A, int n, long long ld, int* status,
                                                                 double* sum_log_diag) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double Ls[CB][CB + 1];
  __shared__ double inv_s[CB];
  __shared__ double dg_s[CB];
  __shared__ __align__(16) double Pi[CB / 4][UT][4];      // [k / 4][row][k % 4]: DMMA fragment loads are 32 consecutive doubles
  __shared__ __align__(16) double Pj[CB / 4][UT][4];
  __shared__ int bad_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double logsum = 0.0;      // thread j < 32 accumulates log L_jj of its panel column (CTA 0 reports)
  if (tid == 0) bad_s = 0x7fffffff;

  for (int k0 = 0; k0 < n; k0 += CB) {
    const int nb = (n - k0 < CB) ? (n - k0) : CB;
    const int k1 = k0 + nb;
    // ---- (1a) diagonal block, whole CTA, in shared memory (identity padding beyond nb).  Right-looking elimination
    // WITHOUT scaling the pivot column (a_ic -= a_ij a_cj / a_jj), so that one barrier per column suffices; the
    // columns are scaled by 1 / sqrt(a_jj) at the end.  Every CTA does this redundantly: no broadcast, no extra barrier.
    for (int e = tid; e < CB * CB; e += CHOL_THREADS) {
      const int i = e / CB, c = e - i * CB;
      double v = (i == c) ? 1.0 : 0.0;
      if (i < nb && c < nb && c <= i) v = A[(long long)(k0 + i) * ld + k0 + c];
      Ls[i][c] = v;
    }
    __syncthreads();
    for (int j = 0; j < CB - 1; ++j) {
      const double rj = 1.0 / Ls[j][j];
      const int cnt = CB - 1 - j;                       // rows / columns j+1 .. 31
      for (int e = tid; e < cnt * cnt; e += CHOL_THREADS) {
        const int i = j + 1 + e / cnt, c = j + 1 + e % cnt;
        if (c <= i) Ls[i][c] = fma(-Ls[i][j] * rj, Ls[c][j], Ls[i][c]);
      }
      __syncthreads();
    }
    if (tid < CB) {
      const double piv = Ls[tid][tid];
      if (!(piv > 0.0) && tid < nb) atomicMin(&bad_s, k0 + tid + 1);      // first non-positive pivot (NaNs propagate)
      const double d = sqrt(piv);
      dg_s[tid] = d;
      inv_s[tid] = 1.0 / d;
      if (tid < nb) logsum += log(d);
    }
    __syncthreads();
    for (int e = tid; e < CB * CB; e += CHOL_THREADS) {
      const int i = e / CB, c = e - i * CB;
      if (c < i) Ls[i][c] *= inv_s[c];
      else if (c == i) Ls[i][c] = dg_s[c];
      else Ls[i][c] = 0.0;
    }
    __syncthreads();
    // ---- (1b) rows below the panel: x L11^T = a_r, one thread per row ----------------------------------------
    for (long long r = (long long)k1 + (long long)blockIdx.x * CHOL_THREADS + tid; r < n; r += (long long)gridDim.x * CHOL_THREADS) {
      double* row = A + r * ld + k0;
      double x[CB];
#pragma unroll
      for (int c = 0; c < CB; ++c) x[c] = (c < nb) ? row[c] : 0.0;
#pragma unroll
      for (int j = 0; j < CB; ++j) {
        double sacc = x[j];
#pragma unroll
        for (int q = 0; q < j; ++q) sacc = fma(-x[q], Ls[j][q], sacc);
        x[j] = sacc * inv_s[j];
      }
#pragma unroll
      for (int c = 0; c < CB; ++c)
        if (c < nb) row[c] = x[c];
    }
    grid.sync();
    // CTA 0 publishes L11 now: no CTA reads the diagonal block any more
    if (blockIdx.x == 0)
      for (int e = tid; e < nb * nb; e += CHOL_THREADS) {
        const int i = e / nb, c = e - i * nb;
        if (c <= i) A[(long long)(k0 + i) * ld + k0 + c] = Ls[i][c];
      }
    // ---- (2) trailing update A22 -= L21 L21^T (lower triangle): 64 x 64 tiles on the FP64 tensor cores, 8 warps as
    // 2 x 4, warp tile 32 x 16 = 4 x 2 m8n8k4 fragments, k = 32 ------------------------------------------------
    const int rem = n - k1;
    const int T = (rem + UT - 1) / UT;
    const int ntiles = T * (T + 1) / 2;
    const int wr = warp >> 2, wc = warp & 3;
    const int frow = lane >> 2, fk = lane & 3;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      int ti = (int)((sqrt(8.0 * t + 1.0) - 1.0) * 0.5);
      while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
      while (ti * (ti + 1) / 2 > t) --ti;
      const int tj = t - ti * (ti + 1) / 2;
      const int i0 = k1 + ti * UT, j0 = k1 + tj * UT;
      __syncthreads();     // the previous tile's reads of Pi / Pj are done
      // the C tile is requested first: its global-memory latency overlaps the panel loads and the tensor-core work
      double2 cur[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int gi = i0 + wr * 32 + u * 8 + frow;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int gj = j0 + wc * 16 + v * 8 + 2 * fk;
          cur[u][v] = make_double2(0.0, 0.0);
          if (gi < n && gj + 1 <= gi) cur[u][v] = *reinterpret_cast<const double2*>(A + (long long)gi * ld + gj);
          else if (gi < n && gj <= gi) cur[u][v].x = A[(long long)gi * ld + gj];
        }
      }
      for (int e = tid; e < UT * CB; e += CHOL_THREADS) {
        const int rr = e / CB, pc = e - rr * CB;
        const int gi = i0 + rr, gj = j0 + rr;
        Pi[pc >> 2][rr][pc & 3] = (gi < n && pc < nb) ? A[(long long)gi * ld + k0 + pc] : 0.0;
        Pj[pc >> 2][rr][pc & 3] = (gj < n && pc < nb) ? A[(long long)gj * ld + k0 + pc] : 0.0;
      }
      __syncthreads();
      double acc0[4][2], acc1[4][2];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) acc0[u][v] = acc1[u][v] = 0.0;
#pragma unroll
      for (int q = 0; q < CB / 4; ++q) {
        double af[4], bf[2];
#pragma unroll
        for (int u = 0; u < 4; ++u) af[u] = Pi[q][wr * 32 + u * 8 + frow][fk];
#pragma unroll
        for (int v = 0; v < 2; ++v) bf[v] = Pj[q][wc * 16 + v * 8 + frow][fk];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 2; ++v) dmma_chol(acc0[u][v], acc1[u][v], af[u], bf[v]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int gi = i0 + wr * 32 + u * 8 + frow;
        if (gi >= n) continue;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int gj = j0 + wc * 16 + v * 8 + 2 * fk;
          double* dst = A + (long long)gi * ld + gj;
          if (gj + 1 <= gi) {                                    // ld even, gj even, base 16-byte aligned
            *reinterpret_cast<double2*>(dst) = make_double2(cur[u][v].x - acc0[u][v], cur[u][v].y - acc1[u][v]);
          } else if (gj <= gi) {
            dst[0] = cur[u][v].x - acc0[u][v];
          }
        }
      }
    }
    grid.sync();
  }
  if (blockIdx.x == 0) {
    __syncthreads();
    if (warp == 0) {
      logsum = warp_sum(logsum);
      if (lane == 0) {
        if (sum_log_diag) *sum_log_diag = logsum;
        if (status) *status = (bad_s == 0x7fffffff) ? 0 : bad_s;
      }
    }
  }
}

// inverse of every 32 x 32 diagonal block of the lower-triangular L: one warp per block, lane c = column c
__global__ void __launch_bounds__(128) tri_inv_diag_kernel(const double* __restrict__ L, int n, long long ld,
                                                           double* __restrict__ X, long long ldx) {
  __shared__ double Ls[4][CB][CB + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int blk = blockIdx.x * 4 + w;
  const int b0 = blk * CB;
  if (b0 >= n) return;
  const int nb = (n - b0 < CB) ? (n - b0) : CB;
#pragma unroll
  for (int c = 0; c < CB; ++c) {
    double v = (c == lane) ? 1.0 : 0.0;
    if (lane < nb && c < nb && c <= lane) v = L[(long long)(b0 + lane) * ld + b0 + c];
    Ls[w][lane][c] = v;
  }
  __syncwarp();
  double x[CB];
#pragma unroll
  for (int r = 0; r < CB; ++r) {
    double sacc = (r == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int q = 0; q < r; ++q) sacc = fma(-Ls[w][r][q], (q >= lane) ? x[q] : 0.0, sacc);
    x[r] = (r >= lane) ? sacc / Ls[w][r][r] : 0.0;
  }
#pragma unroll
  for (int r = 0; r < CB; ++r)
    if (r < nb && lane < nb && r >= lane) X[(long long)(b0 + r) * ldx + b0 + lane] = x[r];
}

// ---------------------------------------------------------------------------------------------------------------
// batched GEMM with triangular k-clipping:  C_b = alpha * A'_b B_b,  A'(i, k) = A[i a_rs + k a_cs], B(k, j) = B[k b_rs + j b_cs]
// ---------------------------------------------------------------------------------------------------------------
struct GemmParams {
  const double* A; long long a_rs, a_cs, a_bs;
  const double* B; long long b_rs, b_cs, b_bs;
  double* C; long long ldc, c_bs;
  int M, N, K;
  double alpha;
  int klo_mode;       // 0: 0;  1: first column of the C tile (B lower triangular);  2: max(first row, first column) (C = X^T X)
  int khi_mode;       // 0: K;  1: last row of the C tile + 1 (A' lower triangular)
  int m_clip_base, m_clip_step;   // batch b: M_b = min(M, m_clip_base - b * m_clip_step) (<= 0: nothing); 0 step = no clipping
  int k_eq_m;         // K_b = M_b
};

constexpr int GT = 64, GK = 16;
constexpr int GEMM_THREADS = 128;

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 64 x 64 CTA tile, 4 warps as 2 x 2, warp tile 32 x 32 = 4 x 4 FP64 tensor-core fragments (mma.m8n8k4: lane holds
// A[lane / 4][lane % 4], B[lane % 4][lane / 4], C[lane / 4][2 (lane % 4) + {0, 1}]).  Operand tiles are stored
// [k / 4][row or column][k % 4]: a fragment load touches 32 consecutive doubles - bank-conflict free.
__global__ void __launch_bounds__(GEMM_THREADS, 4) dense_gemm_kernel(GemmParams p) {
  __shared__ __align__(16) double As[GK / 4][GT][4];
  __shared__ __align__(16) double Bs[GK / 4][GT][4];
  const int b = blockIdx.z;
  int M = p.M;
  if (p.m_clip_step > 0) {
    const int avail = p.m_clip_base - b * p.m_clip_step;
    M = (avail < M) ? avail : M;
  }
  if (M <= 0) return;
  const int K = p.k_eq_m ? M : p.K;
  const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
  if (i0 >= M || j0 >= p.N) return;
  const double* A = p.A + (long long)b * p.a_bs;
  const double* B = p.B + (long long)b * p.b_bs;
  double* C = p.C + (long long)b * p.c_bs;
  int klo = 0, khi = K;
  if (p.klo_mode == 1) klo = j0;
  else if (p.klo_mode == 2) klo = (i0 > j0) ? i0 : j0;
  if (p.khi_mode == 1) khi = (i0 + GT < K) ? i0 + GT : K;
  klo = klo / GK * GK;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wr = warp >> 1, wc = warp & 1;
  const int frow = lane >> 2, fk = lane & 3;
  double acc0[4][4], acc1[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc0[u][v] = acc1[u][v] = 0.0;
  const bool a_kfast = (p.a_cs == 1), b_kfast = (p.b_rs == 1);
  for (int k0 = klo; k0 < khi; k0 += GK) {
    __syncthreads();
#pragma unroll
    for (int rep = 0; rep < (GT * GK) / GEMM_THREADS; ++rep) {
      const int e = tid + rep * GEMM_THREADS;
      int ai, ak, bj, bk;
      if (a_kfast) { ai = e / GK; ak = e % GK; } else { ak = e / GT; ai = e % GT; }
      if (b_kfast) { bj = e / GK; bk = e % GK; } else { bk = e / GT; bj = e % GT; }
      const int gi = i0 + ai, gka = k0 + ak, gj = j0 + bj, gkb = k0 + bk;
      As[ak >> 2][ai][ak & 3] = (gi < M && gka < khi) ? A[(long long)gi * p.a_rs + (long long)gka * p.a_cs] : 0.0;
      Bs[bk >> 2][bj][bk & 3] = (gj < p.N && gkb < khi) ? B[(long long)gkb * p.b_rs + (long long)gj * p.b_cs] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < GK / 4; ++q) {
      double af[4], bf[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) af[u] = As[q][wr * 32 + u * 8 + frow][fk];
#pragma unroll
      for (int v = 0; v < 4; ++v) bf[v] = Bs[q][wc * 32 + v * 8 + frow][fk];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) dmma_884(acc0[u][v], acc1[u][v], af[u], bf[v]);
    }
  }
  const bool vec = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int gi = i0 + wr * 32 + u * 8 + frow;
    if (gi >= M) continue;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int gj = j0 + wc * 32 + v * 8 + 2 * fk;
      double* dst = C + (long long)gi * p.ldc + gj;
      if (vec && gj + 1 < p.N) {
        *reinterpret_cast<double2*>(dst) = make_double2(p.alpha * acc0[u][v], p.alpha * acc1[u][v]);
      } else {
        if (gj < p.N) dst[0] = p.alpha * acc0[u][v];
        if (gj + 1 < p.N) dst[1] = p.alpha * acc1[u][v];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// glue
// ---------------------------------------------------------------------------------------------------------------
// A = s * K + noise * I (lower triangle and diagonal are what the factorisation reads; the whole matrix is written)
__global__ void scale_add_diag_kernel(const double* __restrict__ K, int n, long long ldk, const double* __restrict__ hyper,
                                      double* __restrict__ A, long long lda) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)n * n) return;
  const int i = (int)(e / n), j = (int)(e - (long long)i * n);
  A[(long long)i * lda + j] = hyper[0] * K[(long long)i * ldk + j] + ((i == j) ? hyper[1] : 0.0);
}

__global__ void residual_kernel(const double* __restrict__ y, const double* __restrict__ hyper, int n, double* __restrict__ r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) r[i] = y[i] - hyper[2];
}

// z = X r (X lower triangular): warp per row
__global__ void __launch_bounds__(256) tri_matvec_kernel(const double* __restrict__ X, int n, long long ldx,
                                                         const double* __restrict__ r, double* __restrict__ z) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  double s = 0.0;
  for (int j = lane; j <= row; j += 32) s = fma(X[(long long)row * ldx + j], r[j], s);
  s = warp_sum(s);
  if (lane == 0) z[row] = s;
}

// alpha = X^T z: CTA per 32 columns (lane = column, coalesced 256-byte row segments), the 8 warps stride over the rows
__global__ void __launch_bounds__(256) tri_matvec_t_kernel(const double* __restrict__ X, int n, long long ldx,
                                                           const double* __restrict__ z, double* __restrict__ alpha) {
  __shared__ double red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = blockIdx.x * 32, j = j0 + lane;
  double s0 = 0.0, s1 = 0.0;
  if (j < n) {
    int i = j0 + warp;
    for (; i + 8 < n; i += 16) {
      if (i >= j) s0 = fma(X[(long long)i * ldx + j], z[i], s0);
      s1 = fma((i + 8 >= j) ? X[(long long)(i + 8) * ldx + j] : 0.0, z[i + 8], s1);
    }
    if (i < n && i >= j) s0 = fma(X[(long long)i * ldx + j], z[i], s0);
  }
  red[warp][lane] = s0 + s1;
  __syncthreads();
  if (warp == 0 && j < n) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][lane];
    alpha[j] = t;
  }
}

// G = s/2 (C - alpha alpha^T) (full symmetric); out[1] += sum_ij 1/2 (C - aa^T)_ij K_ij, out[2] += tr 1/2 (C - aa^T)
__global__ void __launch_bounds__(256) mll_weights_kernel(const double* __restrict__ Cm, long long ldc, const double* __restrict__ alpha,
                                                          const double* __restrict__ K, long long ldk, const double* __restrict__ hyper,
                                                          int n, double* __restrict__ G, long long ldg, double* out) {
  __shared__ double red[2][8];
  const double s = hyper[0];
  double gs = 0.0, gn = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < (long long)n * n; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (long long)i * n);
    const double w = 0.5 * (Cm[(long long)i * ldc + j] - alpha[i] * alpha[j]);
    G[(long long)i * ldg + j] = s * w;
    gs = fma(w, K[(long long)i * ldk + j], gs);
    if (i == j) gn += w;
  }
  gs = warp_sum(gs);
  gn = warp_sum(gn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = gs; red[1][warp] = gn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0;
    for (int q = 0; q < 8; ++q) { a += red[0][q]; b2 += red[1][q]; }
    atomicAdd(out + 1, a);
    atomicAdd(out + 2, b2);
  }
}

// out[0] = 1/2 r^T alpha + sum log L_ii + n/2 log(2 pi); out[3] = -sum alpha   (one CTA)
__global__ void __launch_bounds__(256) mll_scalars_kernel(const double* __restrict__ r, const double* __restrict__ alpha, int n,
                                                          const double* __restrict__ sum_log_diag, double* out, int write_grad) {
  __shared__ double red[2][8];
  double q = 0.0, sa = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    q = fma(r[i], alpha[i], q);
    sa += alpha[i];
  }
  q = warp_sum(q);
  sa = warp_sum(sa);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = q; red[1][warp] = sa; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b2 = 0.0;
    for (int w = 0; w < 8; ++w) { a += red[0][w]; b2 += red[1][w]; }
    out[0] = 0.5 * a + *sum_log_diag + 0.5 * n * 1.8378770664093453;
    if (write_grad) out[3] = -b2;
  }
}

__global__ void transpose_lower_kernel(const double* __restrict__ X, int n, long long ldx, double* __restrict__ Xt, long long ldt) {
  __shared__ double tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = by + r, j = bx + threadIdx.x;
    tile[r][threadIdx.x] = (i < n && j < n) ? X[(long long)i * ldx + j] : 0.0;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bx + r, j = by + threadIdx.x;     // Xt[i][j] = X[j][i]
    if (i < n && j < n) Xt[(long long)i * ldt + j] = tile[threadIdx.x][r];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host-side drivers
// ---------------------------------------------------------------------------------------------------------------
static int launch_gemm(const GemmParams& p, int batch, cudaStream_t s) {
  if (p.M <= 0 || p.N <= 0 || batch <= 0) return MGB_OK;
  dim3 grid((unsigned)((p.N + GT - 1) / GT), (unsigned)((p.M + GT - 1) / GT), (unsigned)batch);
  if (grid.y > 65535 || grid.z > 65535) return fail(MGB_ERR_ARG, "dense_gemm: matrix too large");
  dense_gemm_kernel<<<grid, GEMM_THREADS, 0, s>>>(p);
  return after_launch("dense_gemm_kernel");
}

static int chol_inplace(double* A, int n, long long ld, int* status, double* sum_log_diag, cudaStream_t s) {
  int per_sm = 0;
  MGB_TRY(check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, chol_coop_kernel, CHOL_THREADS, 0),
                     "cudaOccupancyMaxActiveBlocksPerMultiprocessor"));
  if (per_sm < 1) return fail(MGB_ERR_CUDA, "cholesky: kernel does not fit an SM");
  const int resident = device_sm_count() * (per_sm > 2 ? 2 : per_sm);
  const int T = (n + UT - 1) / UT;
  long long want = (long long)T * (T + 1) / 2;
  const long long rows = (n + CHOL_THREADS - 1) / CHOL_THREADS;
  if (rows > want) want = rows;
  if (want < 1) want = 1;
  const int grid = (int)(want < resident ? want : resident);
  void* args[] = {&A, &n, &ld, &status, &sum_log_diag};
  MGB_TRY(check_cuda(cudaLaunchCooperativeKernel((const void*)chol_coop_kernel, dim3(grid), dim3(CHOL_THREADS), args, 0, s),
                     "cudaLaunchCooperativeKernel"));
  return after_launch("chol_coop_kernel");
}

// X = L^-1 (X zero-initialised by the caller above the diagonal blocks; T: n x n scratch, same leading dimension ld)
static int tri_inverse(const double* L, int n, long long ld, double* X, double* T, cudaStream_t s) {
  const int nblk = (n + CB - 1) / CB;
  tri_inv_diag_kernel<<<(unsigned)((nblk + 3) / 4), 128, 0, s>>>(L, n, ld, X, ld);
  MGB_TRY(after_launch("tri_inv_diag_kernel"));
  for (long long sz = CB; sz < n; sz *= 2) {
    const int pairs = (int)((n + 2 * sz - 1) / (2 * sz));
    const long long stride = 2 * sz * (ld + 1);
    // T[r0:r0+r, l0:l0+s] = L[r0:r0+r, l0:l0+s] X[l0:l0+s, l0:l0+s]       (X11 lower: k >= column)
    GemmParams g1{};
    g1.A = L + sz * ld; g1.a_rs = ld; g1.a_cs = 1; g1.a_bs = stride;
    g1.B = X; g1.b_rs = ld; g1.b_cs = 1; g1.b_bs = stride;
    g1.C = T + sz * ld; g1.ldc = ld; g1.c_bs = stride;
    g1.M = (int)sz; g1.N = (int)sz; g1.K = (int)sz; g1.alpha = 1.0;
    g1.klo_mode = 1; g1.khi_mode = 0; g1.m_clip_base = (int)(n - sz); g1.m_clip_step = (int)(2 * sz); g1.k_eq_m = 0;
    MGB_TRY(launch_gemm(g1, pairs, s));
    // X[r0:r0+r, l0:l0+s] = -X[r0:r0+r, r0:r0+r] T[r0:r0+r, l0:l0+s]      (X22 lower: k <= row)
    GemmParams g2{};
    g2.A = X + sz * ld + sz; g2.a_rs = ld; g2.a_cs = 1; g2.a_bs = stride;
    g2.B = T + sz * ld; g2.b_rs = ld; g2.b_cs = 1; g2.b_bs = stride;
    g2.C = X + sz * ld; g2.ldc = ld; g2.c_bs = stride;
    g2.M = (int)sz; g2.N = (int)sz; g2.K = (int)sz; g2.alpha = -1.0;
    g2.klo_mode = 0; g2.khi_mode = 1; g2.m_clip_base = (int)(n - sz); g2.m_clip_step = (int)(2 * sz); g2.k_eq_m = 1;
    MGB_TRY(launch_gemm(g2, pairs, s));
  }
  return MGB_OK;
}

}  // namespace mgb

using namespace mgb;

static long long ws_doubles(long long n, int matrices) {
  const long long ld = (n + 1) / 2 * 2;
  return (long long)matrices * n * ld + 4 * ld + 8;
}

extern "C" long long mgb_dense_max_n(void) { return 16384; }

extern "C" long long mgb_dense_fit_workspace_bytes(long long n) {
  if (n < 1) return 0;
  return ws_doubles(n, 2) * (long long)sizeof(double);      // A (Kt -> L), T
}

extern "C" long long mgb_dense_mll_workspace_bytes(long long n) {
  if (n < 1) return 0;
  return ws_doubles(n, 5) * (long long)sizeof(double);      // A, X, T, C, K
}

// Kt = hyper[0] K + hyper[1] I -> L -> rinv = L^-1 (n x n row-major, ldo, zeros above the diagonal), optional transpose,
// alpha = Kt^-1 (y - hyper[2]), *nll.  hyper: DEVICE scalars (outputscale, noise, mean).
extern "C" int mgb_dense_fit_large(const double* K, long long n, long long ldk, const double* y, const double* hyper,
                                   double* rinv, double* rinv_t, long long ldo, double* alpha, double* nll, int* status,
                                   void* workspace, long long workspace_bytes, void* stream) {
  if (!K || !y || !hyper || !rinv || !alpha || !workspace || n < 1 || ldk < n || ldo < n)
    return fail(MGB_ERR_ARG, "dense_fit_large: bad argument");
  if (n > mgb_dense_max_n()) return fail(MGB_ERR_ARG, "dense_fit_large: n too large");
  if (workspace_bytes < mgb_dense_fit_workspace_bytes(n)) return fail(MGB_ERR_ARG, "dense_fit_large: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 15u) return fail(MGB_ERR_ARG, "dense_fit_large: workspace must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int ni = (int)n;
  const long long ld = (n + 1) / 2 * 2;
  double* A = static_cast<double*>(workspace);
  double* T = A + n * ld;
  double* r = T + n * ld;
  double* z = r + ld;
  double* logd = z + ld;
  const long long cells = n * n;
  scale_add_diag_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(K, ni, ldk, hyper, A, ld);
  MGB_TRY(after_launch("scale_add_diag_kernel"));
  MGB_TRY(chol_inplace(A, ni, ld, status, logd, s));
  // rinv doubles as X when its leading dimension matches the scratch; otherwise invert into T's sibling and copy
  if (ldo != ld) return fail(MGB_ERR_ARG, "dense_fit_large: ldo must equal n rounded up to an even number");
  MGB_TRY(check_cuda(cudaMemsetAsync(rinv, 0, sizeof(double) * (size_t)(n * ld), s), "cudaMemsetAsync"));
  MGB_TRY(tri_inverse(A, ni, ld, rinv, T, s));
  residual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y, hyper, ni, r);
  MGB_TRY(after_launch("residual_kernel"));
  tri_matvec_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(rinv, ni, ld, r, z);
  MGB_TRY(after_launch("tri_matvec_kernel"));
  tri_matvec_t_kernel<<<(unsigned)((n + 31) / 32), 256, 0, s>>>(rinv, ni, ld, z, alpha);
  MGB_TRY(after_launch("tri_matvec_t_kernel"));
  if (nll) {
    mll_scalars_kernel<<<1, 256, 0, s>>>(r, alpha, ni, logd, nll, 0);
    MGB_TRY(after_launch("mll_scalars_kernel"));
  }
  if (rinv_t) {
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((n + 31) / 32));
    transpose_lower_kernel<<<grid, dim3(32, 8), 0, s>>>(rinv, ni, ld, rinv_t, ld);
    MGB_TRY(after_launch("transpose_lower_kernel"));
  }
  return MGB_OK;
}

// Marginal likelihood + gradients for a Gram matrix any kernel produced (n beyond the single-CTA kernel):
//   out[0] = nll, out[1] = dnll/ds, out[2] = dnll/dnoise, out[3] = dnll/dmean, G (n x n, ldg) = dnll/dK = s/2 (Kt^-1 - aa^T)
extern "C" int mgb_dense_mll_large(const double* K, long long n, long long ldk, const double* y, const double* hyper,
                                   double* out, double* G, long long ldg, int* status, void* workspace,
                                   long long workspace_bytes, void* stream) {
  if (!K || !y || !hyper || !out || !G || !workspace || n < 1 || ldk < n || ldg < n)
    return fail(MGB_ERR_ARG, "dense_mll_large: bad argument");
  if (n > mgb_dense_max_n()) return fail(MGB_ERR_ARG, "dense_mll_large: n too large");
  if (workspace_bytes < mgb_dense_mll_workspace_bytes(n)) return fail(MGB_ERR_ARG, "dense_mll_large: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 15u) return fail(MGB_ERR_ARG, "dense_mll_large: workspace must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int ni = (int)n;
  const long long ld = (n + 1) / 2 * 2;
  double* A = static_cast<double*>(workspace);
  double* X = A + n * ld;
  double* T = X + n * ld;
  double* Cm = T + n * ld;
  double* r = Cm + 2 * n * ld;          // the fifth matrix slot (K copy) is unused in dense mode: K is the caller's
  double* z = r + ld;
  double* al = z + ld;
  double* logd = al + ld;
  const long long cells = n * n;
  MGB_TRY(check_cuda(cudaMemsetAsync(out, 0, sizeof(double) * 4, s), "cudaMemsetAsync"));
  scale_add_diag_kernel<<<(unsigned)((cells + 255) / 256), 256, 0, s>>>(K, ni, ldk, hyper, A, ld);
  MGB_TRY(after_launch("scale_add_diag_kernel"));
  MGB_TRY(chol_inplace(A, ni, ld, status, logd, s));
  MGB_TRY(check_cuda(cudaMemsetAsync(X, 0, sizeof(double) * (size_t)(n * ld), s), "cudaMemsetAsync"));
  MGB_TRY(tri_inverse(A, ni, ld, X, T, s));
  residual_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(y, hyper, ni, r);
  MGB_TRY(after_launch("residual_kernel"));
  tri_matvec_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(X, ni, ld, r, z);
  MGB_TRY(after_launch("tri_matvec_kernel"));
  tri_matvec_t_kernel<<<(unsigned)((n + 31) / 32), 256, 0, s>>>(X, ni, ld, z, al);
  MGB_TRY(after_launch("tri_matvec_t_kernel"));
  // Kt^-1 = X^T X
  GemmParams g{};
  g.A = X; g.a_rs = 1; g.a_cs = ld; g.a_bs = 0;
  g.B = X; g.b_rs = ld; g.b_cs = 1; g.b_bs = 0;
  g.C = Cm; g.ldc = ld; g.c_bs = 0;
  g.M = ni; g.N = ni; g.K = ni; g.alpha = 1.0; g.klo_mode = 2; g.khi_mode = 0; g.m_clip_step = 0; g.k_eq_m = 0;
  MGB_TRY(launch_gemm(g, 1, s));
  long long wblocks = (cells + 255) / 256;
  if (wblocks > 148 * 8) wblocks = 148 * 8;
  mll_weights_kernel<<<(unsigned)wblocks, 256, 0, s>>>(Cm, ld, al, K, ldk, hyper, ni, G, ldg, out);
  MGB_TRY(after_launch("mll_weights_kernel"));
  mll_scalars_kernel<<<1, 256, 0, s>>>(r, al, ni, logd, out, 1);
  return after_launch("mll_scalars_kernel");
}

// Series kernel: Gram + the above + dnll/dcoef (out[4 + k]) in one call (the n > single-CTA branch of mgb_series_mll).
extern "C" int mgb_series_mll_large(const mgb_series_desc* desc, const double* x, long long n, long long ld, const double* y,
                                    const double* coef, const double* hyper, double* out, int* status, void* workspace,
                                    long long workspace_bytes, void* stream) {
  if (!desc || !x || !y || !coef || !hyper || !out || !workspace || n < 1)
    return fail(MGB_ERR_ARG, "series_mll_large: bad argument");
  if (workspace_bytes < mgb_dense_mll_workspace_bytes(n)) return fail(MGB_ERR_ARG, "series_mll_large: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long ldw = (n + 1) / 2 * 2;
  double* base = static_cast<double*>(workspace);
  double* Kmat = base + 4 * n * ldw;      // fifth matrix slot
  const long long T = (long long)desc->n_factors * desc->n_terms;
  MGB_TRY(mgb_series_gram_fwd(desc, x, n, ld, x, n, ld, coef, Kmat, ldw, stream));
  // G is written over T's slot (free after the triangular inverse)
  double* Gm = base + 2 * n * ldw;
  MGB_TRY(mgb_dense_mll_large(Kmat, n, ldw, y, hyper, out, Gm, ldw, status, workspace, workspace_bytes, stream));
  MGB_TRY(check_cuda(cudaMemsetAsync(out + 4, 0, sizeof(double) * (size_t)T, s), "cudaMemsetAsync"));
  return mgb_series_gram_bwd(desc, x, n, ld, x, n, ld, coef, Gm, ldw, 1, out + 4, nullptr, nullptr, stream);
}
