// Fused exact-GP marginal likelihood for launch-bound training sizes.
//
// What it replaces: one evaluation of gpytorch's ExactMarginalLogLikelihood + its backward, which
// botorch.fit_gpytorch_model calls ~100 times per BO iteration (examples/bo_manifolds_benchmark_examples/
// bo_sphere.py:228,257-262).  At the reference's sizes (5 ... 100 training points) that is ~70 tiny launches (Gram
// kernel, Cholesky, two triangular solves, log-determinant, the Cholesky backward and the Gram backward) whose
// durations, not their work, add up to ~250 us.  Here ONE CTA does all of it in shared memory:
//
//   Kt   = s * K(X, X) + noise * I            series kernel of include/mgb.h, single factor
//   L    = chol(Kt)                           left-looking, 4 columns per round; a group of G lanes per row shares the dots
//   U    = L^-T                               rows of U are independent: one lane group per row, 4 entries per round
//   z    = L^-1 (y - mean), alpha = U z
//   nll  = 1/2 (y - mean)^T alpha + sum log L_ii + n/2 log(2 pi)
//   G    = dnll/dKt = 1/2 (Kt^-1 - alpha alpha^T),  Kt^-1 = U U^T
//   dnll/dcoef_k = s sum_ij G_ij phi_k(u_ij),  dnll/ds = sum_ij G_ij K_ij,  dnll/dnoise = tr G,  dnll/dmean = -sum alpha
//
// Storage: two packed triangles in shared memory (n (n + 1) doubles in total: n <= 160 fits one SM).  The hyper-
// parameter chain (coef as a function of the lengthscale / smoothness) stays in torch autograd, as everywhere else.
#include "mgb_common.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace mgb {

constexpr int MLL_THREADS = 512;
constexpr int MLL_WARPS = MLL_THREADS / 32;
constexpr int MLL_TCHUNK = 16;

struct MllParams {
  const double* x; int n; long long ld;
  const double* y; const double* coef; const double* hyper;
  double* out; int* status;
  int W, T, basis;
  double scale, shift, lo, hi;
  // fit mode (rinv != nullptr): write L^-1 (row-major, lower), its transpose and alpha instead of the gradients
  double* rinv; double* rinv_t; double* alpha_out; long long ldo;
  // dense mode (Kin != nullptr): Kt = s * Kin + noise I from a Gram matrix any kernel produced; writes G = dnll/dKin
  const double* Kin; long long ldk; double* Gout; long long ldg;
};

__device__ __forceinline__ int low_idx(int i, int j) { return i * (i + 1) / 2 + j; }              // j <= i
__device__ __forceinline__ int up_off(int j, int n) { return j * n - j * (j - 1) / 2 - j; }       // + k, k >= j

// p(u) and the pair scalar for rows i, j of the staged inputs
__device__ __forceinline__ double pair_u(const MllParams& p, const double* xs, int i, int j, bool& inside) {
  double d = 0.0;
  for (int k = 0; k < p.W; ++k) d = fma(xs[i * p.W + k], xs[j * p.W + k], d);
  const double raw = fma(d, p.scale, p.shift);
  inside = (raw >= p.lo) && (raw <= p.hi);
  return (raw < p.lo) ? p.lo : ((raw > p.hi) ? p.hi : raw);
}

__device__ __forceinline__ double series_value(const MllParams& p, const double* cf, double u) {
  if (p.basis == MGB_BASIS_MONOMIAL) {
    double s = cf[p.T - 1];
    for (int k = p.T - 2; k >= 0; --k) s = fma(s, u, cf[k]);
    return s;
  }
  double t0 = 1.0, t1 = u, s = cf[0] + ((p.T > 1) ? cf[1] * u : 0.0);
  for (int k = 2; k < p.T; ++k) {
    const double t2 = fma(2.0 * u, t1, -t0);
    s = fma(cf[k], t2, s);
    t0 = t1;
    t1 = t2;
  }
  return s;
}

__global__ void __launch_bounds__(MLL_THREADS, 1) series_mll_kernel(MllParams p) {
  extern __shared__ __align__(16) double sm[];
  const int n = p.n, tri = n * (n + 1) / 2;
  double* A = sm;                 // packed lower: Kt -> L (strict lower part; diagonal in dg) -> weights
  double* U = A + tri;            // packed upper: L^-T
  double* xs = U + tri;           // n * W
  double* cfs = xs + n * p.W;     // T
  double* ry = cfs + p.T;         // y - mean
  double* z = ry + n;             // L^-1 ry
  double* al = z + n;             // alpha
  double* dg = al + n;            // diagonal of L
  double* dgi = dg + n;           // its reciprocal
  double* acc = dgi + n;          // 4 + T block accumulators
  __shared__ int bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double s = p.hyper[0], noise = p.hyper[1], mean = p.hyper[2];

  for (int e = tid; e < n * p.W; e += MLL_THREADS) xs[e] = p.x[(long long)(e / p.W) * p.ld + (e % p.W)];
  for (int e = tid; e < p.T; e += MLL_THREADS) cfs[e] = p.coef[e];
  for (int e = tid; e < n; e += MLL_THREADS) ry[e] = p.y[e] - mean;
  for (int e = tid; e < 4 + p.T; e += MLL_THREADS) acc[e] = 0.0;
  if (tid == 0) bad = 0;
  __syncthreads();

  // ---- 1. Kt (lower triangle): warp per row, lanes over columns -------------------------------------
  for (int i = warp; i < n; i += MLL_WARPS) {
    for (int j = lane; j <= i; j += 32) {
      double kij;
      if (p.Kin != nullptr) {
        kij = p.Kin[(long long)i * p.ldk + j];
      } else {
        bool inside;
        const double u = pair_u(p, xs, i, j, inside);
        kij = series_value(p, cfs, u);
      }
      A[low_idx(i, j)] = s * kij + ((i == j) ? noise : 0.0);
    }
  }
  __syncthreads();

  // Thread groups of G adjacent lanes (G = 2^g <= 32, n * G <= MLL_THREADS) share one dot product: lane q of a group
  // takes every G-th term and the group combines with g shuffle steps.
  int G = 1;
  while (G < 32 && 2 * G * n <= MLL_THREADS) G *= 2;
  const int grp = tid / G, q = tid % G;
  auto group_sum = [&](double v) {
    for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };

  // ---- 2. Cholesky, left-looking, PB columns per round.  Group r owns row r.  Per round: (a) every active row takes
  // its PB dot products against the PB panel rows over the columns left of the panel (independent: 4-way ILP);
  // (b) the PB x PB diagonal block is finished by the groups of the panel rows themselves in PB short steps through
  // shared memory; (c) after one barrier every row below the panel finishes its PB entries in registers.
  // One barrier pair per PB columns instead of per column, reciprocal pivots instead of divisions.
  constexpr int PB = 4;
  for (int j0 = 0; j0 < n; j0 += PB) {
    const int pb = (n - j0 < PB) ? (n - j0) : PB;
    const int r = grp;
    const bool active = (r < n) && (r >= j0);
    double v[PB];
#pragma unroll
    for (int c = 0; c < PB; ++c) v[c] = 0.0;
    if (active) {
      const double* rowr = A + low_idx(r, 0);
      for (int k = q; k < j0; k += G) {
        const double a = rowr[k];
#pragma unroll
        for (int c = 0; c < PB; ++c)
          if (c < pb) v[c] = fma(a, A[low_idx(j0 + c, 0) + k], v[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < PB; ++c) v[c] = group_sum(v[c]);
    if (active) {
#pragma unroll
      for (int c = 0; c < PB; ++c)
        if (c < pb && j0 + c <= r) v[c] = A[low_idx(r, j0 + c)] - v[c];
    }
    // (b) diagonal block: column c of the panel is final once rows j0..j0+c-1 of the block are; pb tiny steps
    for (int c = 0; c < pb; ++c) {
      if (active && q == 0 && r < j0 + pb && r >= j0 + c) {
        // subtract the in-panel part: sum_{k < c} L[r][j0+k] L[j0+c][j0+k]
        double t = v[c];
        for (int k = 0; k < c; ++k) t = fma(-A[low_idx(r, j0 + k)], A[low_idx(j0 + c, j0 + k)], t);
        if (r == j0 + c) {
          const double rs = rsqrt(t);
          dg[r] = t * rs;                         // sqrt(t); NaN for a non-positive pivot: flagged, propagates
          dgi[r] = rs;
          if (!(t > 0.0) && bad == 0) bad = r + 1;
        } else {
          v[c] = t;                               // finished after the barrier below (needs dgi of the pivot row)
        }
      }
      __syncthreads();
      if (active && q == 0 && r < j0 + pb && r > j0 + c) A[low_idx(r, j0 + c)] = v[c] * dgi[j0 + c];
      __syncthreads();
    }
    // (c) rows below the panel: forward substitution inside the panel, in registers
    if (active && q == 0 && r >= j0 + pb) {
#pragma unroll
      for (int c = 0; c < PB; ++c) {
        if (c < pb) {
          double t = v[c];
#pragma unroll
          for (int k = 0; k < PB; ++k)
            if (k < c) t = fma(-v[k], A[low_idx(j0 + c, j0 + k)], t);
          v[c] = t * dgi[j0 + c];
        }
      }
#pragma unroll
      for (int c = 0; c < PB; ++c)
        if (c < pb) A[low_idx(r, j0 + c)] = v[c];
    }
    __syncthreads();
  }

  // ---- 3. U = L^-T: group c owns row c of U (= column c of L^-1): forward substitution, communication only inside
  // the group.  PB entries per round: the dot products over k < i0 are independent (ILP), the in-block part is a
  // short recurrence in registers of lane 0. ----------------------------------------------------------------
  {
    const int c = grp;
    const bool active = c < n;
    double* urow = U + (active ? up_off(c, n) : 0);  // urow[k], k >= c
    if (active && q == 0) urow[c] = dgi[c];
    __syncwarp();
    for (int i0 = 1; i0 < n; i0 += PB) {            // uniform trip count: the shuffles need every lane
      double d[PB];
#pragma unroll
      for (int e = 0; e < PB; ++e) d[e] = 0.0;
      if (active && i0 + PB - 1 > c) {
        const int kend = (i0 > c) ? i0 : c;          // terms k in [c, i0) exist only when i0 > c
        for (int k = c + q; k < kend; k += G) {
          const double u = urow[k];
#pragma unroll
          for (int e = 0; e < PB; ++e)
            if (i0 + e < n) d[e] = fma(A[low_idx(i0 + e, 0) + k], u, d[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < PB; ++e) d[e] = group_sum(d[e]);
      if (active && q == 0) {
        double w[PB];
#pragma unroll
        for (int e = 0; e < PB; ++e) {
          const int i = i0 + e;
          w[e] = 0.0;
          if (i < n && i > c) {
            double t = d[e];
            // in-block terms k in [max(i0, c), i): entries finished earlier in this round (or the diagonal entry)
#pragma unroll
            for (int f = 0; f < PB; ++f) {
              const int k = i0 + f;
              if (f < e && k >= c) t = fma(A[low_idx(i, k)], (k == c) ? dgi[c] : w[f], t);
            }
            w[e] = -t * dgi[i];
            urow[i] = w[e];
          } else if (i == c) {
            w[e] = dgi[c];
          }
        }
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // ---- 4. z = L^-1 ry (column access of U), alpha = U z, log-determinant ------------------------------
  for (int i = tid; i < n; i += MLL_THREADS) {
    double a = 0.0;
    for (int c = 0; c <= i; ++c) a = fma(U[up_off(c, n) + i], ry[c], a);
    z[i] = a;
  }
  __syncthreads();
  {
    double quad = 0.0, gmean = 0.0, logdet = 0.0;
    for (int i = tid; i < n; i += MLL_THREADS) {
      const double* urow = U + up_off(i, n);
      double a = 0.0;
      for (int k = i; k < n; ++k) a = fma(urow[k], z[k], a);
      al[i] = a;
      quad = fma(ry[i], a, quad);
      gmean -= a;
      logdet += log(dg[i]);
    }
    quad = warp_sum(quad);
    gmean = warp_sum(gmean);
    logdet = warp_sum(logdet);
    if (lane == 0) {
      atomicAdd(acc + 0, 0.5 * quad + logdet);
      atomicAdd(acc + 3, gmean);
    }
  }
  __syncthreads();

  if (p.rinv != nullptr) {
    // ---- fit mode: L^-1 = U^T, alpha; out[0] = nll ------------------------------------------------------
    for (int e = tid; e < n * n; e += MLL_THREADS) {
      const int i = e / n, j = e - i * n;                        // rinv[i][j] = L^-1_ij = U[j][i] (j <= i)
      p.rinv[(long long)i * p.ldo + j] = (j <= i) ? U[up_off(j, n) + i] : 0.0;
      p.rinv_t[(long long)i * p.ldo + j] = (j >= i) ? U[up_off(i, n) + j] : 0.0;
    }
    for (int e = tid; e < n; e += MLL_THREADS) p.alpha_out[e] = al[e];
    if (tid == 0) {
      p.out[0] = acc[0] + 0.5 * n * 1.8378770664093453;
      if (p.status) *p.status = bad;
    }
    return;
  }

  // ---- 5. weights: A[i][j] <- (i == j ? 1/2 : 1) * (Kt^-1_ij - alpha_i alpha_j), Kt^-1 = U U^T ------
  for (int i = warp; i < n; i += MLL_WARPS) {
    const double* ui = U + up_off(i, n);
    for (int j = lane; j <= i; j += 32) {
      const double* uj = U + up_off(j, n);
      double k0 = 0.0, k1 = 0.0;
      int k = i;
      for (; k + 1 < n; k += 2) {
        k0 = fma(ui[k], uj[k], k0);
        k1 = fma(ui[k + 1], uj[k + 1], k1);
      }
      if (k < n) k0 = fma(ui[k], uj[k], k0);
      A[low_idx(i, j)] = ((i == j) ? 0.5 : 1.0) * ((k0 + k1) - al[i] * al[j]);
    }
  }
  __syncthreads();

  if (p.Kin != nullptr) {
    // ---- 6 (dense mode). G = dnll/dKin = s/2 (Kt^-1 - alpha alpha^T) as a full symmetric matrix; dnll/ds, dnll/dnoise
    double gs = 0.0, gn = 0.0;
    for (int i = warp; i < n; i += MLL_WARPS) {
      for (int j = lane; j <= i; j += 32) {
        const double w = A[low_idx(i, j)];
        gs = fma(w, p.Kin[(long long)i * p.ldk + j], gs);
        if (i == j) gn += w;
        const double g = s * ((i == j) ? w : 0.5 * w);
        p.Gout[(long long)i * p.ldg + j] = g;
        p.Gout[(long long)j * p.ldg + i] = g;
      }
    }
    gs = warp_sum(gs);
    gn = warp_sum(gn);
    if (lane == 0) {
      atomicAdd(acc + 1, gs);
      atomicAdd(acc + 2, gn);
    }
    __syncthreads();
    for (int e = tid; e < 4; e += MLL_THREADS) p.out[e] = acc[e] + ((e == 0) ? 0.5 * n * 1.8378770664093453 : 0.0);
    if (tid == 0 && p.status) *p.status = bad;
    return;
  }

  // ---- 6. gradient reductions over the pairs --------------------------------------------------------
  for (int c0 = 0; c0 < p.T; c0 += MLL_TCHUNK) {
    double gk[MLL_TCHUNK];
#pragma unroll
    for (int q = 0; q < MLL_TCHUNK; ++q) gk[q] = 0.0;
    double gs = 0.0, gn = 0.0;
    for (int i = warp; i < n; i += MLL_WARPS) {
      for (int j = lane; j <= i; j += 32) {
        const double w = A[low_idx(i, j)];
        bool inside;
        const double u = pair_u(p, xs, i, j, inside);
        if (c0 == 0) {
          gs = fma(w, series_value(p, cfs, u), gs);
          if (i == j) gn += w;
        }
        // phi_k(u) for k in [c0, c0 + MLL_TCHUNK)
        if (p.basis == MGB_BASIS_MONOMIAL) {
          double pw = 1.0;
          for (int k = 0; k < c0; ++k) pw *= u;
#pragma unroll
          for (int q = 0; q < MLL_TCHUNK; ++q) {
            if (c0 + q < p.T) gk[q] = fma(w, pw, gk[q]);
            pw *= u;
          }
        } else {
          double t0 = 1.0, t1 = u;
          for (int k = 0; k < c0; ++k) {
            const double t2 = fma(2.0 * u, t1, -t0);
            t0 = t1;
            t1 = t2;
          }
#pragma unroll
          for (int q = 0; q < MLL_TCHUNK; ++q) {
            if (c0 + q < p.T) gk[q] = fma(w, t0, gk[q]);
            const double t2 = fma(2.0 * u, t1, -t0);
            t0 = t1;
            t1 = t2;
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < MLL_TCHUNK; ++q) {
      const double v = warp_sum(gk[q]);
      if (lane == 0 && c0 + q < p.T) atomicAdd(acc + 4 + c0 + q, s * v);
    }
    if (c0 == 0) {
      gs = warp_sum(gs);
      gn = warp_sum(gn);
      if (lane == 0) {
        atomicAdd(acc + 1, gs);
        atomicAdd(acc + 2, gn);
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < 4 + p.T; e += MLL_THREADS)
    p.out[e] = acc[e] + ((e == 0) ? 0.5 * n * 1.8378770664093453 : 0.0);   // + n/2 log(2 pi)
  if (tid == 0 && p.status) *p.status = bad;
}

// ------------------------------------------------------------------------------------------------
// The same marginal likelihood and gradients through the symmetric sweep (Gauss-Jordan on the full n x n matrix in
// shared memory) instead of Cholesky -> triangular inverse -> U U^T: sweeping pivot k
//     d = m_kk;   m_kk <- -1/d;   m_ik, m_ki <- m_ik / d;   m_ij <- m_ij - m_ik m_kj / d   (i, j != k)
// for k = 0 .. n-1 turns Kt into -Kt^-1, the pivots are the squared Cholesky diagonal (log-determinant, positive
// definiteness flag), and every step is one rank-one update of all n^2 entries, which live in the REGISTERS of the 512
// threads (thread (warp, lane) owns rows warp + 16 a and columns lane + 32 c) from the Gram evaluation to the gradient
// contractions - there is no n x n array in shared memory: one barrier per pivot, independent FMAs, no triangular
// dependency chains.  Used for the gradient modes (series and dense) up to n = 128 (RA x CA <= 8 x 4
// doubles per thread); beyond that and for the fit mode, which must hand out L^-1, series_mll_kernel.
// ------------------------------------------------------------------------------------------------
template <int RA, int CA, bool SERIES>
__global__ void __launch_bounds__(MLL_THREADS, 1) series_mll_sweep_kernel(MllParams p) {
  extern __shared__ __align__(16) double sm[];
  const int n = p.n;
  double* xs = sm;                // n * W
  double* cfs = xs + n * p.W;     // T
  double* ry = cfs + p.T;         // y - mean, zero up to 32 CA
  double* al = ry + 32 * CA;      // alpha, zero up to max(16 RA, 32 CA) = 32 CA
  double* ck = al + 32 * CA;      // row k before its sweep: two buffers of 32 CA doubles (zero beyond n)
  double* dg = ck + 2 * 32 * CA;  // pivots
  double* acc = dg + n;           // 4 + T block accumulators
  __shared__ int bad;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double s = p.hyper[0], noise = p.hyper[1], mean = p.hyper[2];

  for (int e = tid; e < n * p.W; e += MLL_THREADS) xs[e] = p.x[(long long)(e / p.W) * p.ld + (e % p.W)];
  for (int e = tid; e < p.T; e += MLL_THREADS) cfs[e] = p.coef[e];
  for (int e = tid; e < 32 * CA; e += MLL_THREADS) {
    ry[e] = (e < n) ? p.y[e] - mean : 0.0;
    al[e] = 0.0;
  }
  for (int e = tid; e < 4 + p.T; e += MLL_THREADS) acc[e] = 0.0;
  for (int e = tid; e < 2 * 32 * CA; e += MLL_THREADS) ck[e] = 0.0;
  if (tid == 0) bad = 0;
  __syncthreads();

  // Thread (warp, lane) owns the entries (i_a, j_c) = (warp + 16 a, lane + 32 c), a < RA, c < CA, of the FULL symmetric
  // matrix, in registers for the whole kernel.  Pair scalars u_ij of a register row: W broadcast loads of x_i against
  // CA loads of x_j (rows clamped to n - 1: those entries are masked out afterwards).
  constexpr bool series = SERIES;                 // false: dense mode (p.Kin != nullptr)
  const bool mono = (p.basis == MGB_BASIS_MONOMIAL);
  auto row_u = [&](int a, double (&u)[CA]) {
    const int i = min(warp + MLL_WARPS * a, n - 1);
#pragma unroll
    for (int c = 0; c < CA; ++c) u[c] = 0.0;
    for (int k = 0; k < p.W; ++k) {
      const double xi = xs[i * p.W + k];
#pragma unroll
      for (int c = 0; c < CA; ++c) u[c] = fma(xi, xs[min(lane + 32 * c, n - 1) * p.W + k], u[c]);
    }
#pragma unroll
    for (int c = 0; c < CA; ++c) {
      const double raw = fma(u[c], p.scale, p.shift);
      u[c] = (raw < p.lo) ? p.lo : ((raw > p.hi) ? p.hi : raw);
    }
  };

  // ---- 1. Kt ----------------------------------------------------------------------------------------------
  double m[RA][CA];
#pragma unroll
  for (int a = 0; a < RA; ++a) {
    const int i = warp + MLL_WARPS * a;
    double kv[CA];
    if (series) {
      double u[CA];
      row_u(a, u);
      if (mono) {
#pragma unroll
        for (int c = 0; c < CA; ++c) kv[c] = cfs[p.T - 1];
        for (int k = p.T - 2; k >= 0; --k) {
          const double ckf = cfs[k];
#pragma unroll
          for (int c = 0; c < CA; ++c) kv[c] = fma(kv[c], u[c], ckf);
        }
      } else {
#pragma unroll
        for (int c = 0; c < CA; ++c) kv[c] = series_value(p, cfs, u[c]);
      }
    } else {
#pragma unroll
      for (int c = 0; c < CA; ++c) {
        const int j = lane + 32 * c;
        kv[c] = (i < n && j < n) ? p.Kin[(long long)max(i, j) * p.ldk + min(i, j)] : 0.0;      // the lower triangle is the reference
      }
    }
#pragma unroll
    for (int c = 0; c < CA; ++c) {
      const int j = lane + 32 * c;
      m[a][c] = (i < n && j < n) ? s * kv[c] + ((i == j) ? noise : 0.0) : 0.0;
    }
  }

  // ---- 2. sweep every pivot.  Row k (= column k, by symmetry) is published through shared memory by the warp that owns
  // it - all lanes, CA stores each - and gives every thread its column factors m_kj; the row factors m_ik of a thread's
  // own rows sit in its own warp (lane k % 32) and come by shuffle.  One barrier per pivot (two alternating buffers),
  // then RA x CA independent FMAs per thread; the pivot row / column are patched afterwards.  The pivot loop is split as
  // k = 32 kc + 16 kh + kq with kc, kh unrolled, so that the register indices of the pivot column (kc) and pivot row
  // (2 kc + kh) are compile-time constants - no register selects in the loop body.
#pragma unroll
  for (int kc = 0; kc < CA; ++kc) {
#pragma unroll
    for (int kh = 0; kh < 2; ++kh) {
      constexpr int KQ = MLL_WARPS;                     // 16 pivots per (kc, kh): pivot row in warp kq, register row ka
      const int ka = 2 * kc + kh;
      if (ka < RA) {
        for (int kq = 0; kq < KQ; ++kq) {
          const int k = 32 * kc + 16 * kh + kq;
          if (k >= n) break;
          const int kl = k & 31;
          double* cb = ck + (k & 1) * (32 * CA);
          if (warp == kq) {
#pragma unroll
            for (int c = 0; c < CA; ++c) {
              const int j = lane + 32 * c;
              if (j < n) cb[j] = m[ka][c];
            }
          }
          __syncthreads();
          const double d = cb[k];
          if (tid == 0) {
            dg[k] = d;
            if (!(d > 0.0) && bad == 0) bad = k + 1;    // not positive definite: flagged, NaN / garbage propagates
          }
          const double dinv = 1.0 / d;
          double cj[CA];
#pragma unroll
          for (int c = 0; c < CA; ++c) cj[c] = cb[lane + 32 * c];                 // zero beyond n
#pragma unroll
          for (int a = 0; a < RA; ++a) {
            const double f = __shfl_sync(0xffffffffu, m[a][kc], kl) * dinv;      // m_ik / d for the thread's row i
#pragma unroll
            for (int c = 0; c < CA; ++c) m[a][c] = fma(-f, cj[c], m[a][c]);
            if (lane == kl) m[a][kc] = f;                                          // pivot column
          }
          if (warp == kq) {                                                        // pivot row
#pragma unroll
            for (int c = 0; c < CA; ++c) m[ka][c] = cj[c] * dinv;
            if (lane == kl) m[ka][kc] = -dinv;
          }
        }
      }
    }
  }

  // ---- 3. alpha = Kt^-1 ry = -(M ry) from the registers, quadratic form, log-determinant, dnll/dmean -------------
  {
    double ryj[CA];
#pragma unroll
    for (int c = 0; c < CA; ++c) ryj[c] = ry[lane + 32 * c];
    double quad = 0.0, gmean = 0.0;
#pragma unroll
    for (int a = 0; a < RA; ++a) {
      const int i = warp + MLL_WARPS * a;
      double part = 0.0;
#pragma unroll
      for (int c = 0; c < CA; ++c) part = fma(m[a][c], ryj[c], part);
      part = -warp_sum(part);
      if (lane == 0 && i < n) {
        al[i] = part;
        quad = fma(ry[i], part, quad);
        gmean -= part;
      }
    }
    double logdet = (tid < n) ? 0.5 * log(dg[tid]) : 0.0;          // pivot = L_ii^2
    logdet = warp_sum(logdet);
    if (lane == 0) {
      atomicAdd(acc + 0, 0.5 * quad + logdet);
      atomicAdd(acc + 3, gmean);
    }
  }
  __syncthreads();
  // weights over the FULL matrix: wf_ij = 1/2 (Kt^-1_ij - alpha_i alpha_j); sum over all (i, j) = the lower-triangle sum
  // with the diagonal halved
  double alj[CA];
#pragma unroll
  for (int c = 0; c < CA; ++c) alj[c] = al[lane + 32 * c];

  if constexpr (!series) {
    // ---- dense mode: G = dnll/dKin = s wf as a full symmetric matrix; dnll/ds = sum wf Kin, dnll/dnoise = tr wf
    double gs = 0.0, gn = 0.0;
#pragma unroll
    for (int a = 0; a < RA; ++a) {
      const int i = warp + MLL_WARPS * a;
      const double ali = al[min(i, 32 * CA - 1)];
#pragma unroll
      for (int c = 0; c < CA; ++c) {
        const int j = lane + 32 * c;
        if (i < n && j < n) {
          const double wf = 0.5 * (-m[a][c] - ali * alj[c]);
          gs = fma(wf, p.Kin[(long long)max(i, j) * p.ldk + min(i, j)], gs);
          if (i == j) gn += wf;
          p.Gout[(long long)i * p.ldg + j] = s * wf;
        }
      }
    }
    gs = warp_sum(gs);
    gn = warp_sum(gn);
    if (lane == 0) {
      atomicAdd(acc + 1, gs);
      atomicAdd(acc + 2, gn);
    }
    __syncthreads();
    for (int e = tid; e < 4; e += MLL_THREADS) p.out[e] = acc[e] + ((e == 0) ? 0.5 * n * 1.8378770664093453 : 0.0);
    if (tid == 0 && p.status) *p.status = bad;
    return;
  } else {

  // ---- series mode: dnll/dcoef_k = s sum wf phi_k(u), dnll/ds = sum wf K, dnll/dnoise = tr wf, one register row at a time
  constexpr int SWCH = 8;                         // coefficient gradients per pass (register budget)
  for (int c0 = 0; c0 < p.T; c0 += SWCH) {
    double gk[SWCH];
#pragma unroll
    for (int q = 0; q < SWCH; ++q) gk[q] = 0.0;
    double gs = 0.0, gn = 0.0;
#pragma unroll
    for (int a = 0; a < RA; ++a) {
      const int i = warp + MLL_WARPS * a;
      const double ali = al[min(i, 32 * CA - 1)];
      double u[CA], wf[CA];
      row_u(a, u);
#pragma unroll
      for (int c = 0; c < CA; ++c) {
        const int j = lane + 32 * c;
        wf[c] = (i < n && j < n) ? 0.5 * (-m[a][c] - ali * alj[c]) : 0.0;
        if (c0 == 0 && i == j) gn += wf[c];
      }
      if (mono) {
        double pw[CA], val[CA];
#pragma unroll
        for (int c = 0; c < CA; ++c) {
          pw[c] = 1.0;
          val[c] = 0.0;
        }
        if (c0 == 0) {                      // K_ij = sum_k coef_k u^k rides along with the powers of the first chunk
          for (int k = 0; k < p.T; ++k) {
            const double ckf = cfs[k];
#pragma unroll
            for (int c = 0; c < CA; ++c) {
              val[c] = fma(ckf, pw[c], val[c]);
              pw[c] *= u[c];
            }
          }
#pragma unroll
          for (int c = 0; c < CA; ++c) {
            gs = fma(wf[c], val[c], gs);
            pw[c] = 1.0;
          }
        } else {
          for (int k = 0; k < c0; ++k) {
#pragma unroll
            for (int c = 0; c < CA; ++c) pw[c] *= u[c];
          }
        }
#pragma unroll
        for (int q = 0; q < SWCH; ++q) {
          if (c0 + q < p.T) {
#pragma unroll
            for (int c = 0; c < CA; ++c) {
              gk[q] = fma(wf[c], pw[c], gk[q]);
              pw[c] *= u[c];
            }
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < CA; ++c) {
          if (c0 == 0) gs = fma(wf[c], series_value(p, cfs, u[c]), gs);
          double t0 = 1.0, t1 = u[c];
          for (int k = 0; k < c0; ++k) {
            const double t2 = fma(2.0 * u[c], t1, -t0);
            t0 = t1;
            t1 = t2;
          }
#pragma unroll
          for (int q = 0; q < SWCH; ++q) {
            if (c0 + q < p.T) gk[q] = fma(wf[c], t0, gk[q]);
            const double t2 = fma(2.0 * u[c], t1, -t0);
            t0 = t1;
            t1 = t2;
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < SWCH; ++q) {
      const double v = warp_sum(gk[q]);
      if (lane == 0 && c0 + q < p.T) atomicAdd(acc + 4 + c0 + q, s * v);
    }
    if (c0 == 0) {
      gs = warp_sum(gs);
      gn = warp_sum(gn);
      if (lane == 0) {
        atomicAdd(acc + 1, gs);
        atomicAdd(acc + 2, gn);
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < 4 + p.T; e += MLL_THREADS)
    p.out[e] = acc[e] + ((e == 0) ? 0.5 * n * 1.8378770664093453 : 0.0);   // + n/2 log(2 pi)
  if (tid == 0 && p.status) *p.status = bad;
  }
}

// ------------------------------------------------------------------------------------------------
// Spectral coefficients of the sphere / SO(3) series kernels and their Jacobian in ONE launch.
//   a_n   = c_n f(lambda_n; kappa, nu),   f = ((b + lambda) / b)^-(nu + h), b = 2 nu / kappa^2   (Matern, mode 0)
//                                         f = exp(-kappa^2 lambda / 2)                            (heat, mode 1)
//   w     = a / sum_n a_n g_n,   coef = w M            (M: T x Tout basis-change matrix)
//   jac[0] = d coef / d kappa,  jac[1] = d coef / d nu
// (kernels_sphere.py:126-139,212-224 with c = cst_nd, g = zero_gpolynomials, lambda_n = n (n + d - 1), h = d / 2;
//  kernels_so.py:264-273,101 with c = g = 2l + 1, lambda_l = l (l + 1), h = 3/2.)  In torch this chain and its
// backward are ~60 tiny kernels per marginal-likelihood evaluation - half of a graph-replayed step at N = 100.
// ------------------------------------------------------------------------------------------------
struct CoefParams {
  int mode, T, Tout;
  const double* lam; const double* c; const double* g; const double* M;
  double half_dim;
  const double* kappa; const double* nu;
  double* coef; double* jac;
};

__global__ void __launch_bounds__(128) spectral_coef_kernel(CoefParams p) {
  __shared__ double a[MGB_MAX_TERMS], da_k[MGB_MAX_TERMS], da_n[MGB_MAX_TERMS];
  __shared__ double red[3];
  const int tid = threadIdx.x;
  const double kappa = *p.kappa, nu = (p.nu != nullptr) ? *p.nu : 0.0;
  if (tid < 3) red[tid] = 0.0;
  __syncthreads();
  for (int n = tid; n < p.T; n += blockDim.x) {
    const double lam = p.lam[n], c = p.c[n];
    double e, dek, den;
    if (p.mode == 0) {
      const double b = 2.0 * nu / (kappa * kappa), pw = nu + p.half_dim;
      const double lr = log1p(lam / b);                     // log((b + lambda) / b)
      e = exp(-pw * lr);
      const double deb = -pw * e * (1.0 / (b + lam) - 1.0 / b);
      dek = deb * (-2.0 * b / kappa);
      den = deb * (2.0 / (kappa * kappa)) - e * lr;
    } else {
      e = exp(-0.5 * kappa * kappa * lam);
      dek = -kappa * lam * e;
      den = 0.0;
    }
    a[n] = c * e;
    da_k[n] = c * dek;
    da_n[n] = c * den;
  }
  __syncthreads();
  if (tid < 3) {   // fixed summation order: deterministic coefficients
    const double* src = (tid == 0) ? a : ((tid == 1) ? da_k : da_n);
    double acc = 0.0;
    for (int n = 0; n < p.T; ++n) acc = fma(src[n], p.g[n], acc);
    red[tid] = acc;
  }
  __syncthreads();
  const double S = red[0], dSk = red[1], dSn = red[2];
  for (int o = tid; o < p.Tout; o += blockDim.x) {
    double v = 0.0, vk = 0.0, vn = 0.0;
    for (int n = 0; n < p.T; ++n) {
      const double m = p.M[n * p.Tout + o];
      const double w = a[n] / S;
      v = fma(w, m, v);
      vk = fma((da_k[n] - w * dSk) / S, m, vk);
      vn = fma((da_n[n] - w * dSn) / S, m, vn);
    }
    p.coef[o] = v;
    p.jac[o] = vk;
    p.jac[p.Tout + o] = vn;
  }
}

// ------------------------------------------------------------------------------------------------
// Chebyshev coefficients of the circle kernels (torus factors) and their Jacobian in ONE launch.
//   mode 0, integrated Matern (kernels_sphere.py:551-614): a_n = (2 - delta_n0) sum_a tw_a t_a^(nu - 1/2)
//           exp(-2 nu t_a / kappa^2) exp(-4 pi^2 t_a n^2)   over the caller's t grid (the reference's logspace grid,
//           built on the host from the detached lengthscale) with trapezoid weights tw;
//   mode 1, heat (kernels_sphere.py:443-472): a_0 = 1, a_n = 2 exp(-2 pi^2 kappa^2 n^2);
//   coef = a / sum_n a_n (the value at phi = 0),  jac[0] = d coef / d kappa,  jac[1] = d coef / d nu.
// In torch this is ~20 launches forward and ~25 backward per torus factor and marginal-likelihood evaluation.
// ------------------------------------------------------------------------------------------------
constexpr int CC_THREADS = 256;
constexpr int CC_MAXP = 1024;

struct CircleCoefParams {
  int mode, T, P;
  const double* t; const double* tw;
  const double* kappa; const double* nu;
  double* coef; double* jac;
};

__global__ void __launch_bounds__(CC_THREADS) circle_coef_kernel(CircleCoefParams p) {
  __shared__ double ts[CC_MAXP], lw[CC_MAXP], dk[CC_MAXP], dn[CC_MAXP];
  __shared__ double a[CC_THREADS], ak[CC_THREADS], an[CC_THREADS];
  __shared__ double red[3];
  const int tid = threadIdx.x;
  const double kappa = *p.kappa, nu = (p.nu != nullptr) ? *p.nu : 0.0;
  const double four_pi2 = 39.478417604357434;   // 4 pi^2
  if (p.mode == 0) {
    for (int e = tid; e < p.P; e += CC_THREADS) {
      const double t = p.t[e], lt = log(t);
      ts[e] = t;
      // log of the pair-independent part of node e; its derivatives with respect to kappa and nu
      lw[e] = (nu - 0.5) * lt - 2.0 * nu * t / (kappa * kappa);
      dk[e] = 4.0 * nu * t / (kappa * kappa * kappa);
      dn[e] = lt - 2.0 * t / (kappa * kappa);
    }
  }
  __syncthreads();
  double v = 0.0, vk = 0.0, vn = 0.0;
  if (tid < p.T) {
    const double n2 = (double)tid * (double)tid;
    if (p.mode == 0) {
      for (int e = 0; e < p.P; ++e) {
        const double f = p.tw[e] * exp(lw[e] - four_pi2 * ts[e] * n2);
        v += f;
        vk = fma(f, dk[e], vk);
        vn = fma(f, dn[e], vn);
      }
      if (tid > 0) { v *= 2.0; vk *= 2.0; vn *= 2.0; }
    } else {
      if (tid == 0) {
        v = 1.0;
      } else {
        v = 2.0 * exp(-0.5 * four_pi2 * kappa * kappa * n2);
        vk = v * (-four_pi2 * kappa * n2);
      }
    }
  }
  a[tid] = v; ak[tid] = vk; an[tid] = vn;
  __syncthreads();
  if (tid < 3) {   // fixed summation order: deterministic coefficients
    const double* src = (tid == 0) ? a : ((tid == 1) ? ak : an);
    double acc = 0.0;
    for (int n = 0; n < p.T; ++n) acc += src[n];
    red[tid] = acc;
  }
  __syncthreads();
  // Device-side truncation (the host rule of _weights._truncate without its read-back): coefficients whose tail sum
  // sum_{n >= k} |c_n| is not above 1e-16 of the total are written as exact zeros - |T_n| <= 1, so the kernel value moves
  // by less than one ulp - and the Gram kernels skip trailing zeros (series_gram.cu: effective_terms).  Every thread sums
  // its own tail from the far end (the order of the host's flipped cumsum).
  const double S = red[0], c = v / S;
  __syncthreads();                       // a[] is read above by the three summing threads
  a[tid] = (tid < p.T) ? fabs(c) : 0.0;
  __syncthreads();
  double tail = 0.0;
  for (int n = p.T - 1; n >= tid; --n) tail += a[n];
  if (tid == 0) red[0] = tail;
  __syncthreads();
  if (tid < p.T) {
    const bool live = tid < 2 || tail > 1e-16 * red[0];
    p.coef[tid] = live ? c : 0.0;
    p.jac[tid] = live ? (vk - c * red[1]) / S : 0.0;
    p.jac[p.T + tid] = live ? (vn - c * red[2]) / S : 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// Normalised t-node weights of the integrated-Matern quadrature kernels and their Jacobian in ONE launch:
//   w_a = tw_a t_a^(nu - power_offset) exp(-2 nu t_a / kappa^2),  tcoef_a = w_a / sum_b w_b h0_b
// (kernels_hyperbolic.py:367-383, kernels_spd.py:262-289, kernels_euclidean.py:186-207; h0 = the inner integral at
// zero distance on the same grid, NULL = 1).  The grid t is the caller's (detached, as the reference's .item()).
// ------------------------------------------------------------------------------------------------
struct TWeightParams {
  int P;
  const double* t; const double* tw; const double* h0;
  double power_offset;
  const double* kappa; const double* nu;
  double* tcoef; double* jac;
};

__global__ void __launch_bounds__(CC_THREADS) matern_t_weights_kernel(TWeightParams p) {
  __shared__ double w[CC_MAXP], wk[CC_MAXP], wn[CC_MAXP];
  __shared__ double red[3];
  const int tid = threadIdx.x;
  const double kappa = *p.kappa, nu = *p.nu;
  for (int e = tid; e < p.P; e += CC_THREADS) {
    const double t = p.t[e], lt = log(t);
    const double v = p.tw[e] * exp((nu - p.power_offset) * lt - 2.0 * nu * t / (kappa * kappa));
    w[e] = v;
    wk[e] = v * (4.0 * nu * t / (kappa * kappa * kappa));
    wn[e] = v * (lt - 2.0 * t / (kappa * kappa));
  }
  __syncthreads();
  if (tid < 3) {   // fixed summation order
    const double* src = (tid == 0) ? w : ((tid == 1) ? wk : wn);
    double acc = 0.0;
    for (int e = 0; e < p.P; ++e) acc = fma(src[e], (p.h0 != nullptr) ? p.h0[e] : 1.0, acc);
    red[tid] = acc;
  }
  __syncthreads();
  const double S = red[0];
  for (int e = tid; e < p.P; e += CC_THREADS) {
    const double c = w[e] / S;
    p.tcoef[e] = c;
    p.jac[e] = (wk[e] - c * red[1]) / S;
    p.jac[p.P + e] = (wn[e] - c * red[2]) / S;
  }
}

// t[i] = 10^(lo + log10(kappa) + i step) with torch.logspace's float64 formula (first half from the start, second half
// from the end), trapezoid weights tw; one CTA.
__global__ void __launch_bounds__(CC_THREADS) matern_t_grid_kernel(const double* kappa, double lo, double hi, int P, double* t,
                                                                  double* tw) {
  __shared__ double ts[CC_MAXP];
  const double shift = log10(*kappa);
  const double start = lo + shift, end = hi + shift;
  const double step = (P > 1) ? (end - start) / (double)(P - 1) : 0.0;
  const int halfway = P / 2;
  for (int i = threadIdx.x; i < P; i += CC_THREADS) {
    const double v = (i < halfway) ? pow(10.0, start + step * i) : pow(10.0, end - step * (P - i - 1));
    ts[i] = v;
    t[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < P; i += CC_THREADS) {
    double w = 0.0;
    if (i + 1 < P) w += 0.5 * (ts[i + 1] - ts[i]);
    if (i > 0) w += 0.5 * (ts[i] - ts[i - 1]);
    tw[i] = w;
  }
}

}  // namespace mgb

using namespace mgb;

extern "C" int mgb_matern_t_grid(const double* kappa, double lo_exp, double hi_exp, int P, double* t, double* tw,
                                 void* stream) {
  if (!kappa || !t || !tw) return fail(MGB_ERR_ARG, "matern_t_grid: null pointer");
  if (P < 2 || P > CC_MAXP) return fail(MGB_ERR_ARG, "matern_t_grid: P out of range [2,1024]");
  matern_t_grid_kernel<<<1, CC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(kappa, lo_exp, hi_exp, P, t, tw);
  return after_launch("matern_t_grid_kernel");
}

extern "C" int mgb_series_fit(const mgb_series_desc* d, const double* x, long long n, long long ld, const double* y,
                              const double* coef, const double* hyper, double* rinv, double* rinv_t, long long ldo,
                              double* alpha, double* nll, int* status, void* stream) {
  if (!d || !x || !y || !coef || !hyper || !rinv || !rinv_t || !alpha || !nll) return fail(MGB_ERR_ARG, "series_fit: null pointer");
  if (d->n_factors != 1 || d->pair_square) return fail(MGB_ERR_ARG, "series_fit: single-factor kernels in the affine form only");
  if (d->factor_width < 1 || d->n_terms < 1 || d->n_terms > MGB_MAX_TERMS) return fail(MGB_ERR_ARG, "series_fit: bad descriptor");
  if (d->basis != MGB_BASIS_MONOMIAL && d->basis != MGB_BASIS_CHEBYSHEV) return fail(MGB_ERR_ARG, "series_fit: bad basis");
  if (n < 1 || ld < d->factor_width || ldo < n) return fail(MGB_ERR_ARG, "series_fit: bad sizes");
  if (n > mgb_series_mll_single_cta_max_n(d->factor_width, d->n_terms))
    return fail(MGB_ERR_ARG, "series_fit: n exceeds the single-CTA limit: use mgb_dense_fit_large");
  MllParams p{x, (int)n, ld, y, coef, hyper, nll, status, d->factor_width, d->n_terms, d->basis,
              d->scale, d->shift, d->lo, d->hi, rinv, rinv_t, alpha, ldo};
  const size_t smem = sizeof(double) * (size_t)(n * (n + 1) + n * d->factor_width + d->n_terms + 5 * n + 4 + d->n_terms);
  if (smem > 48 * 1024)
    MGB_TRY(check_cuda(cudaFuncSetAttribute(series_mll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute"));
  series_mll_kernel<<<1, MLL_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("series_mll_kernel (fit)");
}

extern "C" int mgb_spectral_coef(int mode, int n_terms, int n_out, const double* lam, const double* c, const double* g,
                                 const double* M, double half_dim, const double* kappa, const double* nu, double* coef,
                                 double* jac, void* stream) {
  if (!lam || !c || !g || !M || !kappa || !coef || !jac) return fail(MGB_ERR_ARG, "spectral_coef: null pointer");
  if (mode != 0 && mode != 1) return fail(MGB_ERR_ARG, "spectral_coef: mode must be 0 (Matern) or 1 (heat)");
  if (mode == 0 && !nu) return fail(MGB_ERR_ARG, "spectral_coef: the Matern form needs nu");
  if (n_terms < 1 || n_terms > MGB_MAX_TERMS || n_out < 1 || n_out > MGB_MAX_TERMS) return fail(MGB_ERR_ARG, "spectral_coef: bad sizes");
  CoefParams p{mode, n_terms, n_out, lam, c, g, M, half_dim, kappa, nu, coef, jac};
  spectral_coef_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("spectral_coef_kernel");
}


extern "C" long long mgb_series_mll_max_n(int factor_width, int n_terms) {
  (void)factor_width; (void)n_terms;
  return mgb_dense_max_n();     // beyond the single-CTA limit: the multi-CTA path (csrc/dense_spd.cu, *_large entry points)
}

extern "C" long long mgb_series_mll_single_cta_max_n(int factor_width, int n_terms) {
  // two packed triangles + inputs + vectors within 220 KB of shared memory
  for (long long n = 200; n >= 1; --n) {
    const long long doubles = n * (n + 1) + n * factor_width + n_terms + 5 * n + 4 + n_terms;
    if (doubles * 8 <= 220 * 1024) return n;
  }
  return 0;
}

// gradient modes: the sweep kernel (MGB_MLL_PATH=chol: the Cholesky-based kernel, kept for comparison)
template <int RA, int CA, bool SERIES>
static int launch_sweep_mode(const MllParams& p, size_t smem, cudaStream_t s) {
  if (smem > 48 * 1024)
    MGB_TRY(check_cuda(cudaFuncSetAttribute(series_mll_sweep_kernel<RA, CA, SERIES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute"));
  series_mll_sweep_kernel<RA, CA, SERIES><<<1, MLL_THREADS, smem, s>>>(p);
  return MGB_OK;
}

template <int RA, int CA>
static int launch_sweep(const MllParams& p, cudaStream_t s) {
  const size_t n = (size_t)p.n;
  const size_t smem = sizeof(double) * (n * p.W + p.T + 4 * 32 * CA + n + 4 + p.T);      // no n x n array: the matrix lives in registers
  return (p.Kin == nullptr) ? launch_sweep_mode<RA, CA, true>(p, smem, s) : launch_sweep_mode<RA, CA, false>(p, smem, s);
}

static int launch_mll_grad(const MllParams& p, size_t smem, cudaStream_t s, const char* what) {
  static const bool chol = [] {
    const char* e = getenv("MGB_MLL_PATH");
    return e && strcmp(e, "chol") == 0;
  }();
  if (chol || p.n > 128) {
    if (smem > 48 * 1024)
      MGB_TRY(check_cuda(cudaFuncSetAttribute(series_mll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute"));
    series_mll_kernel<<<1, MLL_THREADS, smem, s>>>(p);
  } else if (p.n <= 32) {
    MGB_TRY((launch_sweep<2, 1>(p, s)));
  } else if (p.n <= 64) {
    MGB_TRY((launch_sweep<4, 2>(p, s)));
  } else if (p.n <= 80) {
    MGB_TRY((launch_sweep<5, 3>(p, s)));
  } else if (p.n <= 96) {
    MGB_TRY((launch_sweep<6, 3>(p, s)));
  } else if (p.n <= 112) {
    MGB_TRY((launch_sweep<7, 4>(p, s)));
  } else {
    MGB_TRY((launch_sweep<8, 4>(p, s)));
  }
  return after_launch(what);
}

extern "C" int mgb_series_mll(const mgb_series_desc* d, const double* x, long long n, long long ld, const double* y,
                              const double* coef, const double* hyper, double* out, int* status, void* stream) {
  if (!d || !x || !y || !coef || !hyper || !out) return fail(MGB_ERR_ARG, "series_mll: null pointer");
  if (d->n_factors != 1) return fail(MGB_ERR_ARG, "series_mll: single-factor kernels only");
  if (d->pair_square) return fail(MGB_ERR_ARG, "series_mll: pair_square form not supported");
  if (d->factor_width < 1 || d->n_terms < 1 || d->n_terms > MGB_MAX_TERMS) return fail(MGB_ERR_ARG, "series_mll: bad descriptor");
  if (d->basis != MGB_BASIS_MONOMIAL && d->basis != MGB_BASIS_CHEBYSHEV) return fail(MGB_ERR_ARG, "series_mll: bad basis");
  if (n < 1 || ld < d->factor_width) return fail(MGB_ERR_ARG, "series_mll: bad sizes");
  if (n > mgb_series_mll_single_cta_max_n(d->factor_width, d->n_terms))
    return fail(MGB_ERR_ARG, "series_mll: n exceeds the single-CTA limit: use mgb_series_mll_large");
  MllParams p{x, (int)n, ld, y, coef, hyper, out, status, d->factor_width, d->n_terms, d->basis,
              d->scale, d->shift, d->lo, d->hi, nullptr, nullptr, nullptr, 0};
  const size_t smem = sizeof(double) * (size_t)(n * (n + 1) + n * d->factor_width + d->n_terms + 5 * n + 4 + d->n_terms);
  return launch_mll_grad(p, smem, static_cast<cudaStream_t>(stream), "series_mll kernel");
}

extern "C" int mgb_dense_mll(const double* K, long long n, long long ldk, const double* y, const double* hyper, double* out,
                             double* G, long long ldg, int* status, void* stream) {
  if (!K || !y || !hyper || !out || !G) return fail(MGB_ERR_ARG, "dense_mll: null pointer");
  if (n < 1 || ldk < n || ldg < n) return fail(MGB_ERR_ARG, "dense_mll: bad sizes");
  if (n > mgb_series_mll_single_cta_max_n(0, 0)) return fail(MGB_ERR_ARG, "dense_mll: n exceeds the single-CTA limit: use mgb_dense_mll_large");
  MllParams p{};
  p.n = (int)n; p.y = y; p.hyper = hyper; p.out = out; p.status = status;
  p.Kin = K; p.ldk = ldk; p.Gout = G; p.ldg = ldg;
  const size_t smem = sizeof(double) * (size_t)(n * (n + 1) + 5 * n + 4);
  return launch_mll_grad(p, smem, static_cast<cudaStream_t>(stream), "series_mll kernel (dense)");
}

extern "C" int mgb_circle_coef(int mode, int n_terms, const double* t, const double* tw, int P, const double* kappa,
                               const double* nu, double* coef, double* jac, void* stream) {
  if (!kappa || !coef || !jac) return fail(MGB_ERR_ARG, "circle_coef: null pointer");
  if (mode != 0 && mode != 1) return fail(MGB_ERR_ARG, "circle_coef: mode must be 0 (integrated Matern) or 1 (heat)");
  if (n_terms < 1 || n_terms > CC_THREADS) return fail(MGB_ERR_ARG, "circle_coef: n_terms out of range [1,256]");
  if (mode == 0 && (!t || !tw || !nu || P < 2 || P > CC_MAXP)) return fail(MGB_ERR_ARG, "circle_coef: the integrated form needs nu and a t grid of 2..1024 nodes");
  CircleCoefParams p{mode, n_terms, P, t, tw, kappa, nu, coef, jac};
  circle_coef_kernel<<<1, CC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("circle_coef_kernel");
}

extern "C" int mgb_matern_t_weights(const double* t, const double* tw, const double* h0, int P, double power_offset,
                                    const double* kappa, const double* nu, double* tcoef, double* jac, void* stream) {
  if (!t || !tw || !kappa || !nu || !tcoef || !jac) return fail(MGB_ERR_ARG, "matern_t_weights: null pointer");
  if (P < 1 || P > CC_MAXP) return fail(MGB_ERR_ARG, "matern_t_weights: P out of range [1,1024]");
  TWeightParams p{P, t, tw, h0, power_offset, kappa, nu, tcoef, jac};
  matern_t_weights_kernel<<<1, CC_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("matern_t_weights_kernel");
}

extern "C" int mgb_dense_fit(const double* K, long long n, long long ldk, const double* y, const double* hyper, double* rinv,
                             double* rinv_t, long long ldo, double* alpha, double* nll, int* status, void* stream) {
  if (!K || !y || !hyper || !rinv || !rinv_t || !alpha || !nll) return fail(MGB_ERR_ARG, "dense_fit: null pointer");
  if (n < 1 || ldk < n || ldo < n) return fail(MGB_ERR_ARG, "dense_fit: bad sizes");
  if (n > mgb_series_mll_single_cta_max_n(0, 0)) return fail(MGB_ERR_ARG, "dense_fit: n exceeds the single-CTA limit: use mgb_dense_fit_large");
  MllParams p{};
  p.n = (int)n; p.y = y; p.hyper = hyper; p.out = nll; p.status = status;
  p.rinv = rinv; p.rinv_t = rinv_t; p.alpha_out = alpha; p.ldo = ldo;
  p.Kin = K; p.ldk = ldk;
  const size_t smem = sizeof(double) * (size_t)(n * (n + 1) + 5 * n + 4);
  if (smem > 48 * 1024)
    MGB_TRY(check_cuda(cudaFuncSetAttribute(series_mll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute"));
  series_mll_kernel<<<1, MLL_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("series_mll_kernel (dense fit)");
}
