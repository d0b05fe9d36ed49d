// Fused analytic Expected Improvement with gradient and Hessian-vector product at a batch of candidates
// (the trust-region inner loop).
//
// What it replaces: the pymanopt cost / egrad / ehess closures of the reference
// (manifold_optimization/manifold_optimize.py:184-217, pymanopt_addons/tools/autodiff/_pytorch.py:89-116), each of
// which runs botorch's ExpectedImprovement on gpytorch's exact-GP posterior at ONE point and differentiates it with
// torch autograd (once for egrad, twice for ehess): ~150 tiny launches per call, thousands of calls per BO iteration
// (num_restarts x trust-region iterations x truncated-CG steps).  Here one CTA per candidate evaluates, for a
// single-factor series kernel k_j = p(u_j), u_j = clamp(scale <x, x_j> + shift):
//     mu = m + k.alpha,   w = R k,   var = kss - |w|^2,   beta = R^T w                       (R = L^-1, lower)
//     grad mu  = sum_j alpha_j k'_j x_j,        grad var = -2 sum_j beta_j k'_j x_j           (k' includes scale and the
//     Hess mu v  = sum_j alpha_j k''_j (x_j.v) x_j                                             clamp mask, as autograd does)
//     Hess var v = -2 [ sum_j (R^T R dk)_j k'_j x_j + sum_j beta_j k''_j (x_j.v) x_j ],  dk_j = k'_j (x_j.v)
// and chains them through EI(mu, var) = sigma (phi(u) + u Phi(u)), sigma = sqrt(max(var, 1e-9)), analytically.
// The candidates of all restarts go into one launch (grid = number of candidates).
#include "mgb_common.cuh"
#include <cmath>

namespace mgb {

constexpr int AQ_THREADS = 256;
constexpr int AQ_WARPS = AQ_THREADS / 32;
constexpr int AQ_MAXD = 32;

struct AcqParams {
  const double* xc; long long b, ldc;
  const double* v; long long ldv;              // directions for the Hessian-vector product (may be null)
  const double* xt; int n; long long ldt;
  const double* coef;
  const double* R; long long ldr;              // L^-1, row-major lower triangle
  const double* Rt; long long ldrt;            // its transpose, row-major (upper triangle)
  const double* alpha;
  double kss, mean_const, best_f; int maximize;
  double* ei; double* grad; double* hvp; double* mean; double* var;
  int W, T, basis;
  double scale, shift, lo, hi;
};

// p(u), p'(u), p''(u)
__device__ __forceinline__ void series_d2(const AcqParams& p, const double* cf, double u, double& f, double& f1, double& f2) {
  if (p.basis == MGB_BASIS_MONOMIAL) {
    double b = cf[p.T - 1], d1 = 0.0, d2 = 0.0;
    for (int k = p.T - 2; k >= 0; --k) {
      d2 = fma(d2, u, d1);
      d1 = fma(d1, u, b);
      b = fma(b, u, cf[k]);
    }
    f = b; f1 = d1; f2 = 2.0 * d2;
    return;
  }
  // Chebyshev: T, T', T'' by the differentiated three-term recurrence
  double t0 = 1.0, t1 = u, a0 = 0.0, a1 = 1.0, c0 = 0.0, c1 = 0.0;
  f = cf[0] + ((p.T > 1) ? cf[1] * u : 0.0);
  f1 = (p.T > 1) ? cf[1] : 0.0;
  f2 = 0.0;
  for (int k = 2; k < p.T; ++k) {
    const double t2 = fma(2.0 * u, t1, -t0);
    const double a2 = fma(2.0 * u, a1, 2.0 * t1 - a0);
    const double c2 = fma(2.0 * u, c1, 4.0 * a1 - c0);
    f = fma(cf[k], t2, f);
    f1 = fma(cf[k], a2, f1);
    f2 = fma(cf[k], c2, f2);
    t0 = t1; t1 = t2; a0 = a1; a1 = a2; c0 = c1; c1 = c2;
  }
}

// out[i] = sum_j M[i][j] in[j] over j in [jlo(i), jhi(i)): warp per row, lanes over columns (coalesced rows)
__device__ __forceinline__ void tri_matvec(const double* M, long long ld, const double* in, double* out, int n, bool lower,
                                           int warp, int lane) {
  for (int i = warp; i < n; i += AQ_WARPS) {
    const double* row = M + (long long)i * ld;
    const int j0 = lower ? 0 : i, j1 = lower ? i + 1 : n;
    double s = 0.0;
    for (int j = j0 + lane; j < j1; j += 32) s = fma(__ldg(row + j), in[j], s);
    s = warp_sum(s);
    if (lane == 0) out[i] = s;
  }
}

__global__ void __launch_bounds__(AQ_THREADS) series_acquisition_kernel(AcqParams p) {
  extern __shared__ __align__(16) double sm[];
  const int n = p.n;
  double* k0 = sm;            // k_j
  double* k1 = k0 + n;        // k'_j (scale and clamp mask included)
  double* k2 = k1 + n;        // k''_j
  double* w = k2 + n;         // R k
  double* be = w + n;         // R^T w
  double* dk = be + n;        // k'_j (x_j . v)
  double* w2 = dk + n;        // R dk
  double* b2 = w2 + n;        // R^T R dk
  double* xv = b2 + n;        // x_j . v
  double* cfs = xv + n;       // T
  double* red = cfs + p.T;    // 2 + 4 * W reduction slots
  __shared__ double xs[AQ_MAXD], vs[AQ_MAXD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long c = blockIdx.x;
  const bool want_h = (p.v != nullptr) && (p.hvp != nullptr);
  const int W = p.W;

  for (int e = tid; e < p.T; e += AQ_THREADS) cfs[e] = p.coef[e];
  if (tid < W) {
    xs[tid] = p.xc[c * p.ldc + tid];
    vs[tid] = want_h ? p.v[c * p.ldv + tid] : 0.0;
  }
  for (int e = tid; e < 2 + 4 * W; e += AQ_THREADS) red[e] = 0.0;
  __syncthreads();

  // ---- kernel row and its derivatives in the pair scalar -------------------------------------------
  for (int j = tid; j < n; j += AQ_THREADS) {
    const double* xj = p.xt + (long long)j * p.ldt;
    double d = 0.0, pv = 0.0;
    for (int k = 0; k < W; ++k) {
      const double t = __ldg(xj + k);
      d = fma(xs[k], t, d);
      pv = fma(vs[k], t, pv);
    }
    const double raw = fma(d, p.scale, p.shift);
    const bool inside = (raw >= p.lo) && (raw <= p.hi);
    const double u = (raw < p.lo) ? p.lo : ((raw > p.hi) ? p.hi : raw);
    double f, f1, f2;
    series_d2(p, cfs, u, f, f1, f2);
    k0[j] = f;
    k1[j] = inside ? f1 * p.scale : 0.0;
    k2[j] = inside ? f2 * p.scale * p.scale : 0.0;
    xv[j] = pv;
    dk[j] = k1[j] * pv;
  }
  __syncthreads();
  tri_matvec(p.R, p.ldr, k0, w, n, true, warp, lane);
  if (want_h) tri_matvec(p.R, p.ldr, dk, w2, n, true, warp, lane);
  __syncthreads();
  tri_matvec(p.Rt, p.ldrt, w, be, n, false, warp, lane);
  if (want_h) tri_matvec(p.Rt, p.ldrt, w2, b2, n, false, warp, lane);
  __syncthreads();

  // ---- reductions: mu - m, |w|^2, and the four W-vectors -------------------------------------------
  // red[0] = k.alpha, red[1] = |w|^2, red[2 + k] = grad mu, [2 + W + k] = -1/2 grad var,
  // red[2 + 2W + k] = Hess mu v, red[2 + 3W + k] = -1/2 Hess var v
  {
    double s_mu = 0.0, s_w = 0.0;
    for (int j = tid; j < n; j += AQ_THREADS) {
      s_mu = fma(k0[j], __ldg(p.alpha + j), s_mu);
      s_w = fma(w[j], w[j], s_w);
    }
    s_mu = warp_sum(s_mu);
    s_w = warp_sum(s_w);
    if (lane == 0) {
      atomicAdd(red + 0, s_mu);
      atomicAdd(red + 1, s_w);
    }
  }
  for (int k = 0; k < W; ++k) {
    double g_mu = 0.0, g_var = 0.0, h_mu = 0.0, h_var = 0.0;
    for (int j = tid; j < n; j += AQ_THREADS) {
      const double xjk = __ldg(p.xt + (long long)j * p.ldt + k);
      const double a = __ldg(p.alpha + j);
      g_mu = fma(a * k1[j], xjk, g_mu);
      g_var = fma(be[j] * k1[j], xjk, g_var);
      if (want_h) {
        h_mu = fma(a * k2[j] * xv[j], xjk, h_mu);
        h_var = fma(fma(b2[j], k1[j], be[j] * k2[j] * xv[j]), xjk, h_var);
      }
    }
    g_mu = warp_sum(g_mu);
    g_var = warp_sum(g_var);
    h_mu = warp_sum(h_mu);
    h_var = warp_sum(h_var);
    if (lane == 0) {
      atomicAdd(red + 2 + k, g_mu);
      atomicAdd(red + 2 + W + k, g_var);
      atomicAdd(red + 2 + 2 * W + k, h_mu);
      atomicAdd(red + 2 + 3 * W + k, h_var);
    }
  }
  __syncthreads();

  // ---- EI and its derivatives in (mu, var) ----------------------------------------------------------
  if (tid < W || tid == 0) {
    const double mu = p.mean_const + red[0];
    const double var = p.kss - red[1];
    const bool clamped = !(var > 1e-9);
    const double sigma = sqrt(clamped ? 1e-9 : var);
    const double sg = p.maximize ? 1.0 : -1.0;
    const double u = sg * (mu - p.best_f) / sigma;
    const double pdf = exp(-0.5 * u * u) * 0.3989422804014327;
    const double cdf = 0.5 * erfc(-u * 0.7071067811865476);
    const double ei = sigma * (pdf + u * cdf);
    // first derivatives
    const double e_m = sg * cdf;
    const double e_v = clamped ? 0.0 : pdf / (2.0 * sigma);
    // second derivatives
    const double e_mm = pdf / sigma;
    const double e_mv = clamped ? 0.0 : -sg * u * pdf / (2.0 * sigma * sigma);
    const double e_vv = clamped ? 0.0 : pdf * (u * u - 1.0) / (4.0 * sigma * sigma * sigma);
    if (tid == 0) {
      p.ei[c] = ei;
      if (p.mean) p.mean[c] = mu;
      if (p.var) p.var[c] = var;
    }
    if (tid < W) {
      const double gm = red[2 + tid], gv = -2.0 * red[2 + W + tid];
      if (p.grad) p.grad[c * W + tid] = e_m * gm + e_v * gv;
      if (want_h) {
        // directional derivatives of mu and var along v
        double dmu = 0.0, dvar = 0.0;
        for (int k = 0; k < W; ++k) {
          dmu = fma(red[2 + k], vs[k], dmu);
          dvar = fma(-2.0 * red[2 + W + k], vs[k], dvar);
        }
        const double hm = red[2 + 2 * W + tid], hv = -2.0 * red[2 + 3 * W + tid];
        p.hvp[c * W + tid] = (e_mm * dmu + e_mv * dvar) * gm + (e_mv * dmu + e_vv * dvar) * gv + e_m * hm + e_v * hv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Multi-factor variant (torus T^d = product of circles, kernel_utils/kernels_torus.py): k_j = prod_f p_f(u_fj) with
// u_fj from coordinate block f (W wide) of x and x_j.  Per training point and factor:
//     G_f = (prod_{g != f} p_g) p_f'          coefficient of x_j^f in grad k_j
//     H_f = d G_f along v = (prod_{g != f} p_g) p_f'' t_f + p_f' sum_{g != f} (prod_{h != f,g} p_h) p_g' t_g,   t_f = x_j^f . v^f
// (p', p'' include scale and the clamp mask).  Everything else is the single-factor algebra with k'_j x_j replaced by
// the blocks G_f x_j^f.  Products are formed by loops (no division: a factor may vanish).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AQ_THREADS) series_acquisition_multi_kernel(AcqParams p, int F) {
  extern __shared__ __align__(16) double sm[];
  const int n = p.n, W = p.W, D = F * W;
  double* k0 = sm;                 // k_j
  double* w = k0 + n;              // R k
  double* be = w + n;              // R^T w
  double* dk = be + n;             // grad k_j . v
  double* w2 = dk + n;             // R dk
  double* b2 = w2 + n;             // R^T R dk
  double* P = b2 + n;              // [F][n] p_f
  double* P1 = P + F * n;          // [F][n] p_f'
  double* P2 = P1 + F * n;         // [F][n] p_f''
  double* Tf = P2 + F * n;         // [F][n] t_f
  double* G = Tf + F * n;          // [F][n]
  double* H = G + F * n;           // [F][n]
  double* cfs = H + F * n;         // F * T
  double* red = cfs + F * p.T;     // 2 + 4 * D
  __shared__ double xs[AQ_MAXD], vs[AQ_MAXD];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long c = blockIdx.x;
  const bool want_h = (p.v != nullptr) && (p.hvp != nullptr);

  for (int e = tid; e < F * p.T; e += AQ_THREADS) cfs[e] = p.coef[e];
  if (tid < D) {
    xs[tid] = p.xc[c * p.ldc + tid];
    vs[tid] = want_h ? p.v[c * p.ldv + tid] : 0.0;
  }
  for (int e = tid; e < 2 + 4 * D; e += AQ_THREADS) red[e] = 0.0;
  __syncthreads();

  for (int j = tid; j < n; j += AQ_THREADS) {
    const double* xj = p.xt + (long long)j * p.ldt;
    double prod = 1.0;
    for (int f = 0; f < F; ++f) {
      double d = 0.0, pv = 0.0;
      for (int k = 0; k < W; ++k) {
        const double t = __ldg(xj + f * W + k);
        d = fma(xs[f * W + k], t, d);
        pv = fma(vs[f * W + k], t, pv);
      }
      const double raw = fma(d, p.scale, p.shift);
      const bool inside = (raw >= p.lo) && (raw <= p.hi);
      const double u = (raw < p.lo) ? p.lo : ((raw > p.hi) ? p.hi : raw);
      double f0, f1, f2;
      series_d2(p, cfs + f * p.T, u, f0, f1, f2);
      P[f * n + j] = f0;
      P1[f * n + j] = inside ? f1 * p.scale : 0.0;
      P2[f * n + j] = inside ? f2 * p.scale * p.scale : 0.0;
      Tf[f * n + j] = pv;
      prod *= f0;
    }
    k0[j] = prod;
    double dkj = 0.0;
    for (int f = 0; f < F; ++f) {
      double rest = 1.0;
      for (int g = 0; g < F; ++g)
        if (g != f) rest *= P[g * n + j];
      const double gf = rest * P1[f * n + j];
      G[f * n + j] = gf;
      dkj = fma(gf, Tf[f * n + j], dkj);
      double hf = 0.0;
      if (want_h) {
        hf = rest * P2[f * n + j] * Tf[f * n + j];
        double cross = 0.0;
        for (int g = 0; g < F; ++g) {
          if (g == f) continue;
          double r2 = 1.0;
          for (int h = 0; h < F; ++h)
            if (h != f && h != g) r2 *= P[h * n + j];
          cross = fma(r2 * P1[g * n + j], Tf[g * n + j], cross);
        }
        hf = fma(P1[f * n + j], cross, hf);
      }
      H[f * n + j] = hf;
    }
    dk[j] = dkj;
  }
  __syncthreads();
  tri_matvec(p.R, p.ldr, k0, w, n, true, warp, lane);
  if (want_h) tri_matvec(p.R, p.ldr, dk, w2, n, true, warp, lane);
  __syncthreads();
  tri_matvec(p.Rt, p.ldrt, w, be, n, false, warp, lane);
  if (want_h) tri_matvec(p.Rt, p.ldrt, w2, b2, n, false, warp, lane);
  __syncthreads();

  {
    double s_mu = 0.0, s_w = 0.0;
    for (int j = tid; j < n; j += AQ_THREADS) {
      s_mu = fma(k0[j], __ldg(p.alpha + j), s_mu);
      s_w = fma(w[j], w[j], s_w);
    }
    s_mu = warp_sum(s_mu);
    s_w = warp_sum(s_w);
    if (lane == 0) {
      atomicAdd(red + 0, s_mu);
      atomicAdd(red + 1, s_w);
    }
  }
  for (int q = 0; q < D; ++q) {
    const int f = q / W;
    double g_mu = 0.0, g_var = 0.0, h_mu = 0.0, h_var = 0.0;
    for (int j = tid; j < n; j += AQ_THREADS) {
      const double xjq = __ldg(p.xt + (long long)j * p.ldt + q);
      const double a = __ldg(p.alpha + j);
      const double gf = G[f * n + j];
      g_mu = fma(a * gf, xjq, g_mu);
      g_var = fma(be[j] * gf, xjq, g_var);
      if (want_h) {
        const double hf = H[f * n + j];
        h_mu = fma(a * hf, xjq, h_mu);
        h_var = fma(fma(b2[j], gf, be[j] * hf), xjq, h_var);
      }
    }
    g_mu = warp_sum(g_mu);
    g_var = warp_sum(g_var);
    h_mu = warp_sum(h_mu);
    h_var = warp_sum(h_var);
    if (lane == 0) {
      atomicAdd(red + 2 + q, g_mu);
      atomicAdd(red + 2 + D + q, g_var);
      atomicAdd(red + 2 + 2 * D + q, h_mu);
      atomicAdd(red + 2 + 3 * D + q, h_var);
    }
  }
  __syncthreads();

  if (tid < D || tid == 0) {
    const double mu = p.mean_const + red[0];
    const double var = p.kss - red[1];
    const bool clamped = !(var > 1e-9);
    const double sigma = sqrt(clamped ? 1e-9 : var);
    const double sg = p.maximize ? 1.0 : -1.0;
    const double u = sg * (mu - p.best_f) / sigma;
    const double pdf = exp(-0.5 * u * u) * 0.3989422804014327;
    const double cdf = 0.5 * erfc(-u * 0.7071067811865476);
    const double e_m = sg * cdf;
    const double e_v = clamped ? 0.0 : pdf / (2.0 * sigma);
    const double e_mm = pdf / sigma;
    const double e_mv = clamped ? 0.0 : -sg * u * pdf / (2.0 * sigma * sigma);
    const double e_vv = clamped ? 0.0 : pdf * (u * u - 1.0) / (4.0 * sigma * sigma * sigma);
    if (tid == 0) {
      p.ei[c] = sigma * (pdf + u * cdf);
      if (p.mean) p.mean[c] = mu;
      if (p.var) p.var[c] = var;
    }
    if (tid < D) {
      const double gm = red[2 + tid], gv = -2.0 * red[2 + D + tid];
      if (p.grad) p.grad[c * D + tid] = e_m * gm + e_v * gv;
      if (want_h) {
        double dmu = 0.0, dvar = 0.0;
        for (int k = 0; k < D; ++k) {
          dmu = fma(red[2 + k], vs[k], dmu);
          dvar = fma(-2.0 * red[2 + D + k], vs[k], dvar);
        }
        const double hm = red[2 + 2 * D + tid], hv = -2.0 * red[2 + 3 * D + tid];
        p.hvp[c * D + tid] = (e_mm * dmu + e_mv * dvar) * gm + (e_mv * dmu + e_vv * dvar) * gv + e_m * hm + e_v * hv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Radial kernels (hyperbolic H^1..H^3, Euclidean): k_j = s P(z_j) with z_j = d(x, x_j) (geodesic distance, Lorentz
// model) or |x - x_j|^2.  P, P', P'' at the b x n pair scalars come from mgb_radial_profile (the per-pair quadrature and
// its derivatives, csrc/quad_profile.cu); this kernel adds the pair geometry
//     u = max(0, <D,D>_L) + 1e-8,  d = 2 asinh(sqrt(u) / 2),  grad_x d = h'(u) 2 J D,  Hess d v = h''(u) (2JD.v) 2JD + h'(u) 2 J v
// (zero where the max(0, .) of the reference is active), the four triangular mat-vecs and the chain through EI.
// ------------------------------------------------------------------------------------------------
struct RadialAcqParams {
  int geom, width;               // geom 0: Lorentz rows (width = dim + 1); 1: Euclidean rows (z = squared distance)
  const double* xc; long long b, ldc;
  const double* v; long long ldv;
  const double* xt; int n; long long ldt;
  const double* P0; const double* P1; const double* P2;   // b x n each (P2 may be null without directions)
  double outputscale;
  const double* R; long long ldr; const double* Rt; long long ldrt; const double* alpha;
  double kss, mean_const, best_f; int maximize;
  double* ei; double* grad; double* hvp;
};

constexpr int RA_MAXW = 4;

// e = 2 J D (J = diag(-1, 1, ..) for the Lorentz model, identity otherwise), hp = dz/du-type factors; returns whether the
// pair contributes a gradient
__device__ __forceinline__ bool radial_geometry(const RadialAcqParams& p, const double* xs, const double* xj, double* e,
                                                double& hp, double& hpp) {
  double mink = 0.0;
  for (int k = 0; k < p.width; ++k) {
    const double dlt = xs[k] - __ldg(xj + k);
    const double sgn = (p.geom == 0 && k == 0) ? -1.0 : 1.0;
    e[k] = 2.0 * sgn * dlt;
    mink = fma(sgn * dlt, dlt, mink);
  }
  if (p.geom == 1) {
    hp = 1.0;
    hpp = 0.0;
    return true;
  }
  if (!(mink >= 0.0)) {                        // max(0, .) of the reference active (or NaN): no gradient
    hp = hpp = 0.0;
    return false;
  }
  // torch.maximum splits the gradient evenly at a tie: a candidate that coincides with a training point (mink == 0
  // exactly) passes HALF of it, which matters for the Hessian-vector product (h' 2 J v does not vanish there)
  const double tie = (mink == 0.0) ? 0.5 : 1.0;
  const double u = mink + 1e-8;
  const double q = u * (4.0 + u);
  const double h1 = rsqrt(q);                  // d/du 2 asinh(sqrt(u)/2) = (u (4 + u))^-1/2
  hp = tie * h1;
  hpp = tie * (-0.5 * h1 / q * (4.0 + 2.0 * u));
  return true;
}

__global__ void __launch_bounds__(AQ_THREADS) radial_acquisition_kernel(RadialAcqParams p) {
  extern __shared__ __align__(16) double sm[];
  const int n = p.n, W = p.width;
  double* k0 = sm;            // s P
  double* k1 = k0 + n;        // s P'
  double* k2 = k1 + n;        // s P''
  double* dd = k2 + n;        // grad d . v
  double* w = dd + n;
  double* be = w + n;
  double* dk = be + n;
  double* w2 = dk + n;
  double* b2 = w2 + n;
  double* red = b2 + n;       // 2 + 4 * W
  __shared__ double xs[RA_MAXW], vs[RA_MAXW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long c = blockIdx.x;
  const bool want_h = (p.v != nullptr) && (p.hvp != nullptr) && (p.P2 != nullptr);
  if (tid < W) {
    xs[tid] = p.xc[c * p.ldc + tid];
    vs[tid] = want_h ? p.v[c * p.ldv + tid] : 0.0;
  }
  for (int e = tid; e < 2 + 4 * W; e += AQ_THREADS) red[e] = 0.0;
  __syncthreads();
  for (int j = tid; j < n; j += AQ_THREADS) {
    k0[j] = p.outputscale * p.P0[c * n + j];
    k1[j] = p.outputscale * p.P1[c * n + j];
    k2[j] = want_h ? p.outputscale * p.P2[c * n + j] : 0.0;
    double e[RA_MAXW], hp, hpp;
    const bool act = radial_geometry(p, xs, p.xt + (long long)j * p.ldt, e, hp, hpp);
    double ev = 0.0;
    for (int k = 0; k < W; ++k) ev = fma(e[k], vs[k], ev);
    dd[j] = act ? hp * ev : 0.0;
    dk[j] = k1[j] * dd[j];
  }
  __syncthreads();
  tri_matvec(p.R, p.ldr, k0, w, n, true, warp, lane);
  if (want_h) tri_matvec(p.R, p.ldr, dk, w2, n, true, warp, lane);
  __syncthreads();
  tri_matvec(p.Rt, p.ldrt, w, be, n, false, warp, lane);
  if (want_h) tri_matvec(p.Rt, p.ldrt, w2, b2, n, false, warp, lane);
  __syncthreads();
  {
    double s_mu = 0.0, s_w = 0.0;
    for (int j = tid; j < n; j += AQ_THREADS) {
      s_mu = fma(k0[j], __ldg(p.alpha + j), s_mu);
      s_w = fma(w[j], w[j], s_w);
    }
    s_mu = warp_sum(s_mu);
    s_w = warp_sum(s_w);
    if (lane == 0) {
      atomicAdd(red + 0, s_mu);
      atomicAdd(red + 1, s_w);
    }
  }
  {
    double g_mu[RA_MAXW], g_var[RA_MAXW], h_mu[RA_MAXW], h_var[RA_MAXW];
#pragma unroll
    for (int k = 0; k < RA_MAXW; ++k) g_mu[k] = g_var[k] = h_mu[k] = h_var[k] = 0.0;
    for (int j = tid; j < n; j += AQ_THREADS) {
      double e[RA_MAXW], hp, hpp;
      const bool act = radial_geometry(p, xs, p.xt + (long long)j * p.ldt, e, hp, hpp);
      if (!act) continue;
      const double a = __ldg(p.alpha + j);
      double ev = 0.0;
      for (int k = 0; k < W; ++k) ev = fma(e[k], vs[k], ev);
#pragma unroll
      for (int k = 0; k < RA_MAXW; ++k) {
        if (k < W) {
          const double gd = hp * e[k];                                           // (grad d)_k
          const double sgn = (p.geom == 0 && k == 0) ? -1.0 : 1.0;
          const double hd = hpp * ev * e[k] + hp * 2.0 * sgn * vs[k];            // (Hess d v)_k
          g_mu[k] = fma(a * k1[j], gd, g_mu[k]);
          g_var[k] = fma(be[j] * k1[j], gd, g_var[k]);
          if (want_h) {
            const double hk = fma(k2[j] * dd[j], gd, k1[j] * hd);               // d/dv of (k' grad d)_k
            h_mu[k] = fma(a, hk, h_mu[k]);
            h_var[k] = fma(b2[j] * k1[j], gd, fma(be[j], hk, h_var[k]));
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < RA_MAXW; ++k) {
      if (k < W) {
        const double a0 = warp_sum(g_mu[k]), a1 = warp_sum(g_var[k]), a2 = warp_sum(h_mu[k]), a3 = warp_sum(h_var[k]);
        if (lane == 0) {
          atomicAdd(red + 2 + k, a0);
          atomicAdd(red + 2 + W + k, a1);
          atomicAdd(red + 2 + 2 * W + k, a2);
          atomicAdd(red + 2 + 3 * W + k, a3);
        }
      }
    }
  }
  __syncthreads();
  if (tid < W || tid == 0) {
    const double mu = p.mean_const + red[0];
    const double var = p.kss - red[1];
    const bool clamped = !(var > 1e-9);
    const double sigma = sqrt(clamped ? 1e-9 : var);
    const double sg = p.maximize ? 1.0 : -1.0;
    const double u = sg * (mu - p.best_f) / sigma;
    const double pdf = exp(-0.5 * u * u) * 0.3989422804014327;
    const double cdf = 0.5 * erfc(-u * 0.7071067811865476);
    const double e_m = sg * cdf;
    const double e_v = clamped ? 0.0 : pdf / (2.0 * sigma);
    const double e_mm = pdf / sigma;
    const double e_mv = clamped ? 0.0 : -sg * u * pdf / (2.0 * sigma * sigma);
    const double e_vv = clamped ? 0.0 : pdf * (u * u - 1.0) / (4.0 * sigma * sigma * sigma);
    if (tid == 0) p.ei[c] = sigma * (pdf + u * cdf);
    if (tid < W) {
      const double gm = red[2 + tid], gv = -2.0 * red[2 + W + tid];
      if (p.grad) p.grad[c * W + tid] = e_m * gm + e_v * gv;
      if (want_h) {
        double dmu = 0.0, dvar = 0.0;
        for (int k = 0; k < W; ++k) {
          dmu = fma(red[2 + k], vs[k], dmu);
          dvar = fma(-2.0 * red[2 + W + k], vs[k], dvar);
        }
        const double hm = red[2 + 2 * W + tid], hv = -2.0 * red[2 + 3 * W + tid];
        p.hvp[c * W + tid] = (e_mm * dmu + e_mv * dvar) * gm + (e_mv * dmu + e_vv * dvar) * gv + e_m * hm + e_v * hv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// SPD(2) (Mandel 3-vectors): k_j = s P(Delta_j, r2_j) with (Delta, r2) from the log singular values H1 >= H2 of
// A(x) B_j^-1 (kernels_spd.py:230-256; Cholesky factors for the heat kernel, :90-102).  Value + gradient only: the
// reference's SPD trust region uses a finite-difference Hessian (bo_spd.py:383).  P, dP/dDelta, dP/dr2 at the b x n
// pairs come from mgb_spd2_profile; this kernel differentiates the closed-form 2 x 2 singular values.
// ------------------------------------------------------------------------------------------------
struct SpdAcqParams {
  int use_chol;
  const double* xc; long long b, ldc;
  const double* xt; int n; long long ldt;
  const double* P0; const double* Pd; const double* Pr;   // b x n each
  double outputscale;
  const double* R; long long ldr; const double* Rt; long long ldrt; const double* alpha;
  double kss, mean_const, best_f; int maximize;
  double* ei; double* grad;
};

// Mandel (s11, s22, sqrt2 s12) -> 2 x 2 matrix (optionally its lower Cholesky factor), row-major a b / c d
__device__ __forceinline__ void spd2_mat(const double* v, int chol, double& a, double& b, double& c, double& d) {
  const double s11 = v[0], s22 = v[1], s12 = v[2] * 0.7071067811865476;
  if (!chol) {
    a = s11; b = s12; c = s12; d = s22;
  } else {
    a = sqrt(s11);
    b = 0.0;
    c = s12 / a;
    d = sqrt(s22 - c * c);
  }
}

// d Delta / d x and d r2 / d x (3-vectors) for candidate x against training point xj
__device__ __forceinline__ void spd2_pair_grad(const double* x, const double* xj, int chol, double* gD, double* gR) {
  double a, b, c, d, e, f, g, h;
  spd2_mat(x, chol, a, b, c, d);
  spd2_mat(xj, chol, e, f, g, h);
  const double detB = e * h - f * g;
  const double i11 = h / detB, i12 = -f / detB, i21 = -g / detB, i22 = e / detB;
  const double p11 = a * i11 + b * i21, p12 = a * i12 + b * i22;
  const double p21 = c * i11 + d * i21, p22 = c * i12 + d * i22;
  const double t1 = p11 + p22, t2 = p12 - p21, t3 = p11 - p22, t4 = p12 + p21;
  const double q1 = sqrt(t1 * t1 + t2 * t2), q2sq = t3 * t3 + t4 * t4;
  const double q2 = sqrt(q2sq);
  const double det = p11 * p22 - p12 * p21;
  const double s1 = 0.5 * (q1 + q2);
  const double H1 = log(s1), H2 = log(fabs(det) / s1);
  // dH1/dP = (dq1 + dq2) / (q1 + q2); zero sub-gradient of q2 at q2 = 0 (equal singular values)
  const double w1 = 1.0 / (q1 * (q1 + q2)), w2 = (q2sq > 0.0) ? 1.0 / (q2 * (q1 + q2)) : 0.0;
  const double h11 = t1 * w1 + t3 * w2, h22 = t1 * w1 - t3 * w2;
  const double h12 = t2 * w1 + t4 * w2, h21 = -t2 * w1 + t4 * w2;
  // d log|det| / dP = cof(P) / det
  const double l11 = p22 / det, l12 = -p21 / det, l21 = -p12 / det, l22 = p11 / det;
  // Delta = 2 H1 - log|det|,  r2 = H1^2 + H2^2,  H2 = log|det| - H1
  const double cD1 = 2.0, cDl = -1.0;
  const double cR1 = 2.0 * H1 - 2.0 * H2, cRl = 2.0 * H2;
  double GP[2][4];   // [Delta | r2][p11 p12 p21 p22]
  GP[0][0] = cD1 * h11 + cDl * l11; GP[0][1] = cD1 * h12 + cDl * l12; GP[0][2] = cD1 * h21 + cDl * l21; GP[0][3] = cD1 * h22 + cDl * l22;
  GP[1][0] = cR1 * h11 + cRl * l11; GP[1][1] = cR1 * h12 + cRl * l12; GP[1][2] = cR1 * h21 + cRl * l21; GP[1][3] = cR1 * h22 + cRl * l22;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    // dF/dA = dF/dP Binv^T
    const double A11 = GP[q][0] * i11 + GP[q][1] * i12, A12 = GP[q][0] * i21 + GP[q][1] * i22;
    const double A21 = GP[q][2] * i11 + GP[q][3] * i12, A22 = GP[q][2] * i21 + GP[q][3] * i22;
    double* out = q ? gR : gD;
    if (!chol) {
      out[0] = A11;
      out[1] = A22;
      out[2] = (A12 + A21) * 0.7071067811865476;
    } else {
      // A = chol(X): a = sqrt(s11), c = s12 / a, d = sqrt(s22 - c^2)  (b = 0: A12 plays no role)
      const double fs22 = A22 / (2.0 * d);
      const double fc = A21 - A22 * c / d;
      const double s12 = x[2] * 0.7071067811865476;
      const double fs12 = fc / a;
      const double fa = A11 - fc * s12 / (a * a);
      out[0] = fa / (2.0 * a);
      out[1] = fs22;
      out[2] = fs12 * 0.7071067811865476;
    }
  }
}

__global__ void __launch_bounds__(AQ_THREADS) spd2_acquisition_kernel(SpdAcqParams p) {
  extern __shared__ __align__(16) double sm[];
  const int n = p.n;
  double* k0 = sm;
  double* w = k0 + n;
  double* be = w + n;
  double* red = be + n;       // 2 + 6
  __shared__ double xs[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long c = blockIdx.x;
  if (tid < 3) xs[tid] = p.xc[c * p.ldc + tid];
  if (tid < 8) red[tid] = 0.0;
  __syncthreads();
  for (int j = tid; j < n; j += AQ_THREADS) k0[j] = p.outputscale * p.P0[c * n + j];
  __syncthreads();
  tri_matvec(p.R, p.ldr, k0, w, n, true, warp, lane);
  __syncthreads();
  tri_matvec(p.Rt, p.ldrt, w, be, n, false, warp, lane);
  __syncthreads();
  double s_mu = 0.0, s_w = 0.0, gm[3] = {0.0, 0.0, 0.0}, gv[3] = {0.0, 0.0, 0.0};
  for (int j = tid; j < n; j += AQ_THREADS) {
    const double a = __ldg(p.alpha + j);
    s_mu = fma(k0[j], a, s_mu);
    s_w = fma(w[j], w[j], s_w);
    if (p.grad) {
      double gD[3], gR[3];
      spd2_pair_grad(xs, p.xt + (long long)j * p.ldt, p.use_chol, gD, gR);
      const double pd = p.outputscale * p.Pd[c * n + j], pr = p.outputscale * p.Pr[c * n + j];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double gk = fma(pd, gD[k], pr * gR[k]);      // d k_j / d x_k
        gm[k] = fma(a, gk, gm[k]);
        gv[k] = fma(be[j], gk, gv[k]);
      }
    }
  }
  s_mu = warp_sum(s_mu);
  s_w = warp_sum(s_w);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    gm[k] = warp_sum(gm[k]);
    gv[k] = warp_sum(gv[k]);
  }
  if (lane == 0) {
    atomicAdd(red + 0, s_mu);
    atomicAdd(red + 1, s_w);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      atomicAdd(red + 2 + k, gm[k]);
      atomicAdd(red + 5 + k, gv[k]);
    }
  }
  __syncthreads();
  if (tid < 3) {
    const double mu = p.mean_const + red[0];
    const double var = p.kss - red[1];
    const bool clamped = !(var > 1e-9);
    const double sigma = sqrt(clamped ? 1e-9 : var);
    const double sg = p.maximize ? 1.0 : -1.0;
    const double u = sg * (mu - p.best_f) / sigma;
    const double pdf = exp(-0.5 * u * u) * 0.3989422804014327;
    const double cdf = 0.5 * erfc(-u * 0.7071067811865476);
    if (tid == 0) p.ei[c] = sigma * (pdf + u * cdf);
    if (p.grad) p.grad[c * 3 + tid] = sg * cdf * red[2 + tid] + (clamped ? 0.0 : pdf / (2.0 * sigma)) * (-2.0 * red[5 + tid]);
  }
}

}  // namespace mgb

using namespace mgb;

extern "C" int mgb_spd2_acquisition_ei(int use_cholesky, const double* xc, long long b, long long ldc, const double* xtrain,
                                       long long n, long long ldt, const double* P0, const double* Pd, const double* Pr,
                                       double outputscale, const double* rinv, long long ldr, const double* rinv_t,
                                       long long ldrt, const double* alpha, double kss, double mean_const, double best_f,
                                       int maximize, double* ei, double* grad, void* stream) {
  if (b == 0) return MGB_OK;
  if (!xc || !xtrain || !P0 || !rinv || !rinv_t || !alpha || !ei) return fail(MGB_ERR_ARG, "spd2_acquisition_ei: null pointer");
  if (grad && (!Pd || !Pr)) return fail(MGB_ERR_ARG, "spd2_acquisition_ei: the gradient needs both profile derivatives");
  if (b < 0 || n < 1 || ldc < 3 || ldt < 3 || ldr < n || ldrt < n) return fail(MGB_ERR_ARG, "spd2_acquisition_ei: bad sizes");
  if (b > 0x7fffffffLL) return fail(MGB_ERR_ARG, "spd2_acquisition_ei: too many candidates for one launch");
  const size_t smem = sizeof(double) * (size_t)(3 * n + 8);
  if (smem > 200 * 1024) return fail(MGB_ERR_ARG, "spd2_acquisition_ei: n too large for the shared-memory vectors");
  if (smem > 48 * 1024)
    MGB_TRY(check_cuda(cudaFuncSetAttribute(spd2_acquisition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute"));
  SpdAcqParams p{use_cholesky, xc, b, ldc, xtrain, (int)n, ldt, P0, Pd, Pr, outputscale, rinv, ldr, rinv_t, ldrt, alpha,
                 kss, mean_const, best_f, maximize, ei, grad};
  spd2_acquisition_kernel<<<(unsigned)b, AQ_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("spd2_acquisition_kernel");
}

namespace mgb {
// EI = sigma (phi(u) + u Phi(u)), sigma = sqrt(max(var, 1e-9)), u = +-(mean - best_f) / sigma: botorch's analytic
// ExpectedImprovement on posterior means / variances, one launch instead of a dozen elementwise torch kernels.
__global__ void expected_improvement_kernel(const double* mean, const double* var, long long m, double best_f, int maximize,
                                            double* out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  const double sigma = sqrt(fmax(var[c], 1e-9));
  double u = (mean[c] - best_f) / sigma;
  if (!maximize) u = -u;
  const double pdf = exp(-0.5 * u * u) * 0.3989422804014327;
  const double cdf = 0.5 * erfc(-u * 0.7071067811865476);
  out[c] = sigma * (pdf + u * cdf);
}
}  // namespace mgb

extern "C" int mgb_expected_improvement(const double* mean, const double* var, long long m, double best_f, int maximize,
                                        double* ei, void* stream) {
  if (m == 0) return MGB_OK;
  if (!mean || !var || !ei || m < 0) return fail(MGB_ERR_ARG, "expected_improvement: bad argument");
  expected_improvement_kernel<<<(unsigned)((m + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(mean, var, m, best_f,
                                                                                                           maximize, ei);
  return after_launch("expected_improvement_kernel");
}

extern "C" int mgb_radial_acquisition_ei(int geom, int width, const double* xc, long long b, long long ldc,
                                         const double* v, long long ldv, const double* xtrain, long long n, long long ldt,
                                         const double* P0, const double* P1, const double* P2, double outputscale,
                                         const double* rinv, long long ldr, const double* rinv_t, long long ldrt,
                                         const double* alpha, double kss, double mean_const, double best_f, int maximize,
                                         double* ei, double* grad, double* hvp, void* stream) {
  if (b == 0) return MGB_OK;
  if (!xc || !xtrain || !P0 || !P1 || !rinv || !rinv_t || !alpha || !ei) return fail(MGB_ERR_ARG, "radial_acquisition_ei: null pointer");
  if ((geom != 0 && geom != 1) || width < 1 || width > RA_MAXW) return fail(MGB_ERR_ARG, "radial_acquisition_ei: geometry must be H^1..H^3 or R^1..R^4");
  if (b < 0 || n < 1 || ldc < width || ldt < width || ldr < n || ldrt < n || (v && ldv < width))
    return fail(MGB_ERR_ARG, "radial_acquisition_ei: bad sizes");
  if (hvp && (!v || !P2)) return fail(MGB_ERR_ARG, "radial_acquisition_ei: hvp needs directions and the second profile derivative");
  if (b > 0x7fffffffLL) return fail(MGB_ERR_ARG, "radial_acquisition_ei: too many candidates for one launch");
  const size_t smem = sizeof(double) * (size_t)(9 * n + 2 + 4 * width);
  if (smem > 200 * 1024) return fail(MGB_ERR_ARG, "radial_acquisition_ei: n too large for the shared-memory vectors");
  if (smem > 48 * 1024)
    MGB_TRY(check_cuda(cudaFuncSetAttribute(radial_acquisition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute"));
  RadialAcqParams p{geom, width, xc, b, ldc, v, ldv, xtrain, (int)n, ldt, P0, P1, P2, outputscale, rinv, ldr, rinv_t, ldrt,
                    alpha, kss, mean_const, best_f, maximize, ei, grad, hvp};
  radial_acquisition_kernel<<<(unsigned)b, AQ_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("radial_acquisition_kernel");
}


extern "C" int mgb_series_acquisition_ei(const mgb_series_desc* d, const double* xc, long long b, long long ldc,
                                         const double* v, long long ldv, const double* xtrain, long long n,
                                         long long ldt, const double* coef, const double* rinv, long long ldr,
                                         const double* rinv_t, long long ldrt, const double* alpha, double kss,
                                         double mean_const, double best_f, int maximize, double* ei, double* grad,
                                         double* hvp, double* mean, double* var, void* stream) {
  if (b == 0) return MGB_OK;
  if (!d || !xc || !xtrain || !coef || !rinv || !rinv_t || !alpha || !ei) return fail(MGB_ERR_ARG, "series_acquisition_ei: null pointer");
  if (d->pair_square) return fail(MGB_ERR_ARG, "series_acquisition_ei: affine pair scalar only");
  if (d->n_factors < 1 || d->n_factors > MGB_MAX_FACTORS) return fail(MGB_ERR_ARG, "series_acquisition_ei: n_factors out of range");
  if (d->factor_width < 1 || d->n_factors * d->factor_width > AQ_MAXD) return fail(MGB_ERR_ARG, "series_acquisition_ei: row width out of range [1,32]");
  if (d->n_terms < 1 || d->n_terms > MGB_MAX_TERMS) return fail(MGB_ERR_ARG, "series_acquisition_ei: bad n_terms");
  if (d->basis != MGB_BASIS_MONOMIAL && d->basis != MGB_BASIS_CHEBYSHEV) return fail(MGB_ERR_ARG, "series_acquisition_ei: bad basis");
  const int W = d->factor_width, F = d->n_factors, D = F * W;
  if (b < 0 || n < 1 || ldc < D || ldt < D || ldr < n || ldrt < n || (v && ldv < D)) return fail(MGB_ERR_ARG, "series_acquisition_ei: bad sizes");
  if (hvp && !v) return fail(MGB_ERR_ARG, "series_acquisition_ei: hvp requested without directions");
  if (n > 16000) return fail(MGB_ERR_ARG, "series_acquisition_ei: n too large for the shared-memory vectors");
  if (b > 0x7fffffffLL) return fail(MGB_ERR_ARG, "series_acquisition_ei: too many candidates for one launch");
  AcqParams p{xc, b, ldc, v, ldv, xtrain, (int)n, ldt, coef, rinv, ldr, rinv_t, ldrt, alpha, kss, mean_const, best_f,
              maximize, ei, grad, hvp, mean, var, W, d->n_terms, d->basis, d->scale, d->shift, d->lo, d->hi};
  if (F > 1) {
    const size_t smem_m = sizeof(double) * (size_t)(6 * n + 6 * (size_t)F * n + (size_t)F * d->n_terms + 2 + 4 * D);
    if (smem_m > 200 * 1024) return fail(MGB_ERR_ARG, "series_acquisition_ei: n * n_factors too large for the shared-memory vectors");
    if (smem_m > 48 * 1024)
      MGB_TRY(check_cuda(cudaFuncSetAttribute(series_acquisition_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem_m), "cudaFuncSetAttribute"));
    series_acquisition_multi_kernel<<<(unsigned)b, AQ_THREADS, smem_m, static_cast<cudaStream_t>(stream)>>>(p, F);
    return after_launch("series_acquisition_multi_kernel");
  }
  const size_t smem = sizeof(double) * (size_t)(9 * n + d->n_terms + 2 + 4 * W);
  if (smem > 200 * 1024) return fail(MGB_ERR_ARG, "series_acquisition_ei: n too large for the shared-memory vectors");
  if (smem > 48 * 1024)
    MGB_TRY(check_cuda(cudaFuncSetAttribute(series_acquisition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute"));
  series_acquisition_kernel<<<(unsigned)b, AQ_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("series_acquisition_kernel");
}
