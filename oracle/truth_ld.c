/* TEST INFRASTRUCTURE ONLY - extended-precision "truth" leg for the exact-GP posterior parity tests.
 *
 * The exact-GP posterior arithmetic of the reference lives in unvendored gpytorch / botorch (call sites
 * examples/bo_manifolds_benchmark_examples/bo_sphere.py:215-233,265 and
 * manifold_optimization/manifold_optimize.py:195,254,337); the float64 oracle (oracle/restated.py: gp_fit /
 * gp_predict) restates the textbook equations.  When the CUDA path and that oracle differ by 1e-11 at a
 * training-set condition number of 1e6, neither is "right": both carry cond * eps.  This file computes the same
 * quantities in x87 long double (64-bit mantissa, eps = 5.4e-20) so that a test can attribute the error:
 *
 *     Kt   = s K(X, X) + noise I            K(x, y) = P(clamp(<x, y>, lo, hi)),  P = sum_n w[n] sum_p G[n][p] u^p
 *     L    = chol(Kt)                       (blocked right-looking, OpenMP over rows)
 *     mean = m + s K(X*, X) Kt^-1 (y - m)
 *     var  = s P(hi) - || L^-1 s K(X*, X)^T ||^2
 *
 * P is the reference's own polynomial: w[n] are its float64 series weights (kernels_sphere.py:126-139) and G[n][p]
 * the float64 power-sum coefficients of math_utils/gegenbauer_polynomials.py:31-33, both taken as exact numbers; all
 * arithmetic from the inner product on is long double.  Only tests/ may load this library (oracle/build_truth.py).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long double ld;

static ld poly(const ld* mono, int T, ld u) {
  ld s = mono[T - 1];
  for (int p = T - 2; p >= 0; --p) s = s * u + mono[p];
  return s;
}

static ld pair_scalar(const double* a, const double* b, int D, double lo, double hi) {
  ld s = 0.0L;
  for (int k = 0; k < D; ++k) s += (ld)a[k] * (ld)b[k];
  if (s < (ld)lo) s = (ld)lo;
  if (s > (ld)hi) s = (ld)hi;
  return s;
}

/* in-place lower Cholesky of the n x n row-major matrix a (upper part ignored); returns 0 or the failing pivot + 1 */
static long chol_ld(ld* a, long n) {
  const long B = 48;
  for (long k0 = 0; k0 < n; k0 += B) {
    const long kb = (n - k0 < B) ? (n - k0) : B;
    /* diagonal block, unblocked */
    for (long j = k0; j < k0 + kb; ++j) {
      ld d = a[j * n + j];
      for (long p = k0; p < j; ++p) d -= a[j * n + p] * a[j * n + p];
      if (!(d > 0.0L)) return j + 1;
      d = sqrtl(d);
      a[j * n + j] = d;
      for (long i = j + 1; i < k0 + kb; ++i) {
        ld s = a[i * n + j];
        for (long p = k0; p < j; ++p) s -= a[i * n + p] * a[j * n + p];
        a[i * n + j] = s / d;
      }
    }
    /* panel: rows below the block, L21 = A21 L11^-T */
#pragma omp parallel for schedule(static)
    for (long i = k0 + kb; i < n; ++i) {
      for (long j = k0; j < k0 + kb; ++j) {
        ld s = a[i * n + j];
        for (long p = k0; p < j; ++p) s -= a[i * n + p] * a[j * n + p];
        a[i * n + j] = s / a[j * n + j];
      }
    }
    /* trailing update: A22 -= L21 L21^T (lower part) */
#pragma omp parallel for schedule(dynamic, 8)
    for (long i = k0 + kb; i < n; ++i) {
      const ld* li = a + i * n + k0;
      for (long j = k0 + kb; j <= i; ++j) {
        const ld* lj = a + j * n + k0;
        ld s = 0.0L;
        for (long p = 0; p < kb; ++p) s += li[p] * lj[p];
        a[i * n + j] -= s;
      }
    }
  }
  return 0;
}

/* Returns 0, -1 (allocation) or pivot + 1 (not positive definite).  mean / var: m doubles each (rounded from long
 * double).  G: T x T row-major, G[n][p] = coefficient of u^p in the n-th basis polynomial. */
long mgb_truth_series_posterior(const double* xt, long n, const double* y, const double* xc, long m, int D,
                                const double* w, const double* G, int T, double lo, double hi, double outputscale,
                                double noise, double mean_const, double* mean, double* var, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  ld* mono = (ld*)calloc((size_t)T, sizeof(ld));
  ld* a = (ld*)malloc(sizeof(ld) * (size_t)n * (size_t)n);
  ld* alpha = (ld*)malloc(sizeof(ld) * (size_t)n);
  if (!mono || !a || !alpha) { free(mono); free(a); free(alpha); return -1; }
  for (int k = 0; k < T; ++k)
    for (int p = 0; p < T; ++p) mono[p] += (ld)w[k] * (ld)G[k * T + p];
  const ld s = (ld)outputscale;
#pragma omp parallel for schedule(dynamic, 16)
  for (long i = 0; i < n; ++i)
    for (long j = 0; j <= i; ++j) {
      ld v = s * poly(mono, T, pair_scalar(xt + i * D, xt + j * D, D, lo, hi));
      if (i == j) v += (ld)noise;
      a[i * n + j] = v;
    }
  long rc = chol_ld(a, n);
  if (rc != 0) { free(mono); free(a); free(alpha); return rc; }
  /* alpha = Kt^-1 (y - m): forward then backward substitution */
  for (long i = 0; i < n; ++i) {
    ld v = (ld)y[i] - (ld)mean_const;
    for (long p = 0; p < i; ++p) v -= a[i * n + p] * alpha[p];
    alpha[i] = v / a[i * n + i];
  }
  for (long i = n - 1; i >= 0; --i) {
    ld v = alpha[i];
    for (long p = i + 1; p < n; ++p) v -= a[p * n + i] * alpha[p];
    alpha[i] = v / a[i * n + i];
  }
  const ld kss = s * poly(mono, T, (ld)hi);
#pragma omp parallel
  {
    ld* ks = (ld*)malloc(sizeof(ld) * (size_t)n);
#pragma omp for schedule(dynamic, 4)
    for (long c = 0; c < m; ++c) {
      ld mu = (ld)mean_const;
      for (long j = 0; j < n; ++j) {
        ks[j] = s * poly(mono, T, pair_scalar(xc + c * D, xt + j * D, D, lo, hi));
        mu += ks[j] * alpha[j];
      }
      ld ss = 0.0L;
      for (long i = 0; i < n; ++i) {       /* v = L^-1 ks in place */
        ld v = ks[i];
        const ld* li = a + i * n;
        for (long p = 0; p < i; ++p) v -= li[p] * ks[p];
        v /= li[i];
        ks[i] = v;
        ss += v * v;
      }
      mean[c] = (double)mu;
      var[c] = (double)(kss - ss);
    }
    free(ks);
  }
  free(mono); free(a); free(alpha);
  return 0;
}

/* 64 on x86-64 (x87 extended); 113 where long double is IEEE quad; 53 would mean this leg adds nothing */
int mgb_truth_long_double_digits(void) { return LDBL_MANT_DIG; }
