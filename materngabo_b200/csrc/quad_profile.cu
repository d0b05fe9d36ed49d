// Radial profiles of the quadrature kernels and their derivatives with respect to the pair scalar(s).
//
// The hyperbolic / Euclidean / SPD(2) kernels of the reference are functions of one or two cheap per-pair
// scalars (geodesic distance d, squared Euclidean distance, (Delta, r^2) of the log singular values) composed
// with an expensive quadrature.  The trust-region optimisers of the reference differentiate the acquisition
// function with respect to the INPUT point (first order for SPD, second order for hyperbolic:
// manifold_optimization/manifold_optimize.py:184-217, pymanopt_addons/tools/autodiff/_pytorch.py:89-116), which
// by the chain rule needs d^k/dd^k of the quadrature.  These kernels evaluate exactly that:
//
//   hyperbolic / Euclidean :  out[e] = sum_a tcoef[a] * d^order/dz^order heat(z_e, tval[a]),  order in {0,1,2}
//        heat as in mgb.h (kernels_hyperbolic.py:258-330; kernels_euclidean.py:82-89), z = distance (or |x-y|^2);
//        for H^2 the s grid moves with d (s = s0 + k ds + d, kernels_hyperbolic.py:299), which the derivative
//        includes, as torch autograd does in the reference.
//   SPD(2)                 :  out[e] = sum_a tcoef[a] * D h_a(Delta_e, r2_e),  D in {1, d/dDelta, d/dr2}
//        h_a as in mgb.h (kernels_spd.py:200-209).
//
// One warp per element, outer (t) nodes over the lanes, inner nodes in a loop over per-warp tables.  Element
// counts on this path are small (one candidate row against the training set), so clarity wins over the
// factorised-exponential tricks of the Gram kernels in quad_gram.cu.
#include "mgb_common.cuh"

namespace mgb {

constexpr int PTHREADS = 128;
constexpr int PWARPS = PTHREADS / 32;
constexpr int PMAXP = 1024;
constexpr int PMAXT = 128;
constexpr int PNQ = PMAXT / 32;

enum ProfKind { PK_H1 = 0, PK_H2 = 1, PK_H3 = 2, PK_EUCLID = 3 };
enum ProfMode { PM_SUM = 0, PM_BWD_COEF = 1 };

struct ProfParams {
  int kind, order, mode;
  const double* in; long long count;
  const double* tval; const double* tcoef; int Pt;
  double s0, ds; int Ps;
  const double* G;
  double* out;
};

// q(z) = z / sinh z and its first two derivatives, stable near 0
__device__ __forceinline__ void z_over_sinh(double z, double& q, double& q1, double& q2) {
  if (fabs(z) < 0.1) {
    const double z2 = z * z;
    q = 1.0 + z2 * (-1.0 / 6 + z2 * (7.0 / 360 + z2 * (-31.0 / 15120 + z2 * (127.0 / 604800 + z2 * (-73.0 / 3421440)))));
    q1 = z * (-1.0 / 3 + z2 * (7.0 / 90 + z2 * (-31.0 / 2520 + z2 * (127.0 / 75600 + z2 * (-73.0 / 342144)))));
    q2 = -1.0 / 3 + z2 * (7.0 / 30 + z2 * (-31.0 / 504 + z2 * (127.0 / 10800 + z2 * (-73.0 * 9 / 342144))));
  } else {
    const double sh = sinh(z), ch = cosh(z);
    q = z / sh;
    q1 = (sh - z * ch) / (sh * sh);
    q2 = -2.0 * ch / (sh * sh) - z / sh + 2.0 * z * ch * ch / (sh * sh * sh);
  }
}

__global__ void __launch_bounds__(PTHREADS) hyp_profile_kernel(ProfParams p) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tab = sm + (size_t)warp * 4 * p.Ps;   // [s^2 | T0 | T1 | T2] for this warp's element (H^2 only)
  double i2t[PNQ], tc[PNQ], gacc[PNQ];
#pragma unroll
  for (int q = 0; q < PNQ; ++q) {
    const int a = lane + 32 * q;
    i2t[q] = (a < p.Pt) ? 0.5 / p.tval[a] : 0.0;
    tc[q] = (a < p.Pt && p.tcoef != nullptr) ? p.tcoef[a] : 0.0;
    gacc[q] = 0.0;
  }
  const long long wstride = (long long)gridDim.x * PWARPS;
  for (long long e = (long long)blockIdx.x * PWARPS + warp; e < p.count; e += wstride) {
    const double z = p.in[e];
    double val[PNQ];
    if (p.kind == PK_H2) {
      const double chd = cosh(z), shd = sinh(z);
      __syncwarp();
      for (int k = lane; k < p.Ps; k += 32) {
        const double b = p.s0 + k * p.ds;
        const double s = b + z;
        const double hs = sinh(0.5 * b);
        const double cm1 = 2.0 * hs * hs, shb = sinh(b);
        const double D = fma(chd, cm1, shd * shb);    // cosh(s) - cosh(d), no cancellation
        const double N1 = fma(shd, cm1, chd * shb);   // sinh(s) - sinh(d) = dD/dd
        const double w = (k == 0 || k == p.Ps - 1) ? 0.5 * p.ds : p.ds;
        const double F = rsqrt(D);
        const double F1 = -0.5 * F * F * F * N1;                                   // dF/dd
        const double F2 = 0.75 * F * F * F * F * F * N1 * N1 - 0.5 * F;            // d2F/dd2 (d2D/dd2 = D)
        double t0, t1 = 0.0, t2 = 0.0;
        if (p.order == 0) {
          t0 = w * s * F;
        } else if (p.order == 1) {          // I' = E [ (F + s F') - (1/2t) s^2 F ]
          t0 = w * (F + s * F1);
          t1 = -w * s * s * F;
        } else {                            // I'' = E [ (2F' + sF'') + (1/2t)(-3sF - 2s^2F') + (1/2t)^2 s^3 F ]
          t0 = w * (2.0 * F1 + s * F2);
          t1 = w * (-3.0 * s * F - 2.0 * s * s * F1);
          t2 = w * s * s * s * F;
        }
        tab[k] = s * s;
        tab[p.Ps + k] = t0;
        tab[2 * p.Ps + k] = t1;
        tab[3 * p.Ps + k] = t2;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < PNQ; ++q) {
        val[q] = 0.0;
        if (lane + 32 * q < p.Pt) {
          const double h = i2t[q];
          double acc = 0.0;
          for (int k = 0; k < p.Ps; ++k) {
            const double poly = fma(h, fma(h, tab[3 * p.Ps + k], tab[2 * p.Ps + k]), tab[p.Ps + k]);
            acc = fma(exp(-0.5 * h * tab[k]), poly, acc);   // exp(-s^2 / 4t)
          }
          val[q] = acc;
        }
      }
    } else {
#pragma unroll
      for (int q = 0; q < PNQ; ++q) {
        val[q] = 0.0;
        if (lane + 32 * q < p.Pt) {
          const double h = i2t[q];
          if (p.kind == PK_EUCLID) {       // exp(-z / 4t), z = squared distance
            const double E = exp(-0.5 * h * z);
            val[q] = (p.order == 0) ? E : (p.order == 1) ? -0.5 * h * E : 0.25 * h * h * E;
          } else {
            const double E = exp(-0.5 * h * z * z);
            const double E1 = -z * h;                       // E'/E
            const double E2 = -h + z * z * h * h;           // E''/E
            if (p.kind == PK_H1) {
              val[q] = (p.order == 0) ? E : (p.order == 1) ? E1 * E : E2 * E;
            } else {
              double qq, q1, q2;
              z_over_sinh(z + 1e-8, qq, q1, q2);
              val[q] = (p.order == 0) ? E * qq : (p.order == 1) ? E * (E1 * qq + q1)
                                                               : E * (E2 * qq + 2.0 * E1 * q1 + q2);
            }
          }
        }
      }
    }
    if (p.mode == PM_SUM) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < PNQ; ++q) s = fma(tc[q], val[q], s);
      s = warp_sum(s);
      if (lane == 0) p.out[e] = s;
    } else {
      const double g = p.G[e];
#pragma unroll
      for (int q = 0; q < PNQ; ++q) gacc[q] = fma(g, val[q], gacc[q]);
    }
  }
  if (p.mode == PM_BWD_COEF) {
#pragma unroll
    for (int q = 0; q < PNQ; ++q) {
      const int a = lane + 32 * q;
      if (a < p.Pt) atomicAdd(p.out + a, gacc[q]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
struct SpdProfParams {
  int which, mode;   // which: 0 value, 1 d/dDelta, 2 d/dr2
  const double* delta; const double* r2; long long count;
  const double* tcoef; const double* rscale; const double* bscale; int Pt;
  const double* bval; const double* bw; int Pb;
  const double* G;
  double* out;
};

__global__ void __launch_bounds__(PTHREADS) spd2_profile_kernel(SpdProfParams p) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tab = sm + (size_t)warp * 3 * p.Pb;   // [c_b | qA | qB]
  double tc[PNQ], rs[PNQ], bs[PNQ], gacc[PNQ];
#pragma unroll
  for (int q = 0; q < PNQ; ++q) {
    const int a = lane + 32 * q;
    tc[q] = (a < p.Pt && p.tcoef != nullptr) ? p.tcoef[a] : 0.0;
    rs[q] = (a < p.Pt) ? p.rscale[a] : 0.0;
    bs[q] = (a < p.Pt) ? p.bscale[a] : 0.0;
    gacc[q] = 0.0;
  }
  const long long wstride = (long long)gridDim.x * PWARPS;
  for (long long e = (long long)blockIdx.x * PWARPS + warp; e < p.count; e += wstride) {
    const double delta = p.delta[e], r2 = p.r2[e];
    __syncwarp();
    for (int k = lane; k < p.Pb; k += 32) {
      const double b = p.bval[k];
      const double sb = sinh(b), sbd = sinh(b + delta);
      const double base = p.bw[k] * rsqrt(sb * sbd);
      const double lin = 2.0 * b + delta;
      tab[k] = b * (b + delta);
      if (p.which == 1) {   // d/dDelta [ lin e^{-b(b+Delta) bs} base ] = e^{..} base [ 1 - lin coth(b+Delta)/2 - lin b bs ]
        tab[p.Pb + k] = base * (1.0 - 0.5 * lin * cosh(b + delta) / sbd);
        tab[2 * p.Pb + k] = -base * lin * b;
      } else {
        tab[p.Pb + k] = base * lin;
        tab[2 * p.Pb + k] = 0.0;
      }
    }
    __syncwarp();
    double val[PNQ];
#pragma unroll
    for (int q = 0; q < PNQ; ++q) {
      val[q] = 0.0;
      if (lane + 32 * q < p.Pt) {
        double acc = 0.0;
        for (int k = 0; k < p.Pb; ++k)
          acc = fma(exp(-tab[k] * bs[q]), fma(bs[q], tab[2 * p.Pb + k], tab[p.Pb + k]), acc);
        acc *= exp(-r2 * rs[q]);
        val[q] = (p.which == 2) ? -rs[q] * acc : acc;
      }
    }
    if (p.mode == PM_SUM) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < PNQ; ++q) s = fma(tc[q], val[q], s);
      s = warp_sum(s);
      if (lane == 0) p.out[e] = s;
    } else {
      const double g = p.G[e];
#pragma unroll
      for (int q = 0; q < PNQ; ++q) gacc[q] = fma(g, val[q], gacc[q]);
    }
  }
  if (p.mode == PM_BWD_COEF) {
#pragma unroll
    for (int q = 0; q < PNQ; ++q) {
      const int a = lane + 32 * q;
      if (a < p.Pt) atomicAdd(p.out + a, gacc[q]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// All derivative orders in one launch (order = 3 / which = 3 of the C ABI): what the trust-region step asks for
// (mgb_*_acquisition_ei_step).  Same tables and sums as the single-order kernels above; the exponentials - the
// cost of these kernels - are shared by the three results.  out is order-major: out[o * count + e].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PTHREADS) hyp_profile_all_kernel(ProfParams p) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tab = sm + (size_t)warp * 7 * p.Ps;   // [s^2 | a0 | b0 b1 | c0 c1 c2] (H^2 only)
  double i2t[PNQ], tc[PNQ];
#pragma unroll
  for (int q = 0; q < PNQ; ++q) {
    const int a = lane + 32 * q;
    i2t[q] = (a < p.Pt) ? 0.5 / p.tval[a] : 0.0;
    tc[q] = (a < p.Pt) ? p.tcoef[a] : 0.0;
  }
  const long long wstride = (long long)gridDim.x * PWARPS;
  for (long long e = (long long)blockIdx.x * PWARPS + warp; e < p.count; e += wstride) {
    const double z = p.in[e];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    if (p.kind == PK_H2) {
      const double chd = cosh(z), shd = sinh(z);
      __syncwarp();
      for (int k = lane; k < p.Ps; k += 32) {
        const double b = p.s0 + k * p.ds;
        const double s = b + z;
        const double hs = sinh(0.5 * b);
        const double cm1 = 2.0 * hs * hs, shb = sinh(b);
        const double D = fma(chd, cm1, shd * shb);
        const double N1 = fma(shd, cm1, chd * shb);
        const double w = (k == 0 || k == p.Ps - 1) ? 0.5 * p.ds : p.ds;
        const double F = rsqrt(D);
        const double F1 = -0.5 * F * F * F * N1;
        const double F2 = 0.75 * F * F * F * F * F * N1 * N1 - 0.5 * F;
        tab[k] = s * s;
        tab[p.Ps + k] = w * s * F;
        tab[2 * p.Ps + k] = w * (F + s * F1);
        tab[3 * p.Ps + k] = -w * s * s * F;
        tab[4 * p.Ps + k] = w * (2.0 * F1 + s * F2);
        tab[5 * p.Ps + k] = w * (-3.0 * s * F - 2.0 * s * s * F1);
        tab[6 * p.Ps + k] = w * s * s * s * F;
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < PNQ; ++q) {
        if (lane + 32 * q < p.Pt) {
          const double h = i2t[q], mh = -0.5 * h;
          double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll 2
          for (int k = 0; k < p.Ps; ++k) {
            const double ex = exp_bf(mh * tab[k]);   // exp(-s^2 / 4t)
            a0 = fma(ex, tab[p.Ps + k], a0);
            a1 = fma(ex, fma(h, tab[3 * p.Ps + k], tab[2 * p.Ps + k]), a1);
            a2 = fma(ex, fma(h, fma(h, tab[6 * p.Ps + k], tab[5 * p.Ps + k]), tab[4 * p.Ps + k]), a2);
          }
          s0 = fma(tc[q], a0, s0);
          s1 = fma(tc[q], a1, s1);
          s2 = fma(tc[q], a2, s2);
        }
      }
    } else {
      double qq = 1.0, q1 = 0.0, q2 = 0.0;
      if (p.kind == PK_H3) z_over_sinh(z + 1e-8, qq, q1, q2);
#pragma unroll
      for (int q = 0; q < PNQ; ++q) {
        if (lane + 32 * q < p.Pt) {
          const double h = i2t[q];
          double v0, v1, v2;
          if (p.kind == PK_EUCLID) {
            const double E = exp_bf(-0.5 * h * z);
            v0 = E; v1 = -0.5 * h * E; v2 = 0.25 * h * h * E;
          } else {
            const double E = exp_bf(-0.5 * h * z * z);
            const double E1 = -z * h, E2 = -h + z * z * h * h;
            if (p.kind == PK_H1) {
              v0 = E; v1 = E1 * E; v2 = E2 * E;
            } else {
              v0 = E * qq; v1 = E * (E1 * qq + q1); v2 = E * (E2 * qq + 2.0 * E1 * q1 + q2);
            }
          }
          s0 = fma(tc[q], v0, s0);
          s1 = fma(tc[q], v1, s1);
          s2 = fma(tc[q], v2, s2);
        }
      }
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      p.out[e] = s0;
      p.out[p.count + e] = s1;
      p.out[2 * p.count + e] = s2;
    }
  }
}

__global__ void __launch_bounds__(PTHREADS) spd2_profile_all_kernel(SpdProfParams p) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* tab = sm + (size_t)warp * 4 * p.Pb;   // [c_b | base lin | d/dDelta constant part | d/dDelta bscale part]
  double tc[PNQ], rs[PNQ], bs[PNQ];
#pragma unroll
  for (int q = 0; q < PNQ; ++q) {
    const int a = lane + 32 * q;
    tc[q] = (a < p.Pt) ? p.tcoef[a] : 0.0;
    rs[q] = (a < p.Pt) ? p.rscale[a] : 0.0;
    bs[q] = (a < p.Pt) ? p.bscale[a] : 0.0;
  }
  const long long wstride = (long long)gridDim.x * PWARPS;
  for (long long e = (long long)blockIdx.x * PWARPS + warp; e < p.count; e += wstride) {
    const double delta = p.delta[e], r2 = p.r2[e];
    __syncwarp();
    for (int k = lane; k < p.Pb; k += 32) {
      const double b = p.bval[k];
      const double sb = sinh(b), sbd = sinh(b + delta);
      const double base = p.bw[k] * rsqrt(sb * sbd);
      const double lin = 2.0 * b + delta;
      tab[k] = b * (b + delta);
      tab[p.Pb + k] = base * lin;
      tab[2 * p.Pb + k] = base * (1.0 - 0.5 * lin * cosh(b + delta) / sbd);
      tab[3 * p.Pb + k] = -base * lin * b;
    }
    __syncwarp();
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int q = 0; q < PNQ; ++q) {
      if (lane + 32 * q < p.Pt) {
        double a0 = 0.0, a1 = 0.0;
        const double mb = -bs[q];
#pragma unroll 4
        for (int k = 0; k < p.Pb; ++k) {
          const double ex = exp_bf(mb * tab[k]);
          a0 = fma(ex, tab[p.Pb + k], a0);
          a1 = fma(ex, fma(bs[q], tab[3 * p.Pb + k], tab[2 * p.Pb + k]), a1);
        }
        const double er = tc[q] * exp_bf(-r2 * rs[q]);
        s0 = fma(er, a0, s0);
        s1 = fma(er, a1, s1);
        s2 = fma(-rs[q] * er, a0, s2);
      }
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) {
      p.out[e] = s0;
      p.out[p.count + e] = s1;
      p.out[2 * p.count + e] = s2;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tabulated radial profile (opt-in, NOT the per-entry quadrature of the reference).
//
// For the hyperbolic and Euclidean kernels the Gram entry is a fixed one-dimensional function of the pair scalar,
// K_ij = P(z_ij): the reference (and hyp_gram_kernel) re-evaluates a 64 x 64 quadrature for every pair although P
// does not depend on the pair.  Here P is sampled once with the quadrature kernel (mgb_radial_profile) at Chebyshev
// nodes of a piecewise partition and every entry evaluates a degree-`deg` polynomial instead.  Partition: v = z + 1
// (z = 2 sinh(d/2) = sqrt(max(0, <D,D>_L) + 1e-8) for H^n, |x - y| for R^d) is split by its binary exponent and
// the top `sub_bits` mantissa bits, so interval index and local coordinate come from integer operations on the bit
// pattern - no logarithm, and no asinh per entry.  The host validates the table against the quadrature (error
// bound reported to the caller); parity tests compare with the oracle at the usual 1e-10.
// ------------------------------------------------------------------------------------------------
struct TableParams {
  int geom, dim;          // geom 0: Lorentz rows (dim + 1 wide); 1: Euclidean rows (dim wide)
  const double* x1; long long m, ld1;
  const double* x2; long long n, ld2;
  int sub_bits, n_int, deg;
  const double* coef;     // [n_int][deg + 1], monomials in s = 2 t - 1, t = position inside the interval
  double* K; long long ldk;
};

constexpr int TB_THREADS = 256;
constexpr int TB_ROWS = 64;      // rows of K per CTA
constexpr int TB_MAXW = 4;

__global__ void __launch_bounds__(TB_THREADS) radial_table_gram_kernel(TableParams p) {
  extern __shared__ __align__(16) double tab[];   // [n_int][deg + 2] (row pitch deg + 2: odd number of 8-byte banks)
  const int pitch = p.deg + 2;
  for (int e = threadIdx.x; e < p.n_int * (p.deg + 1); e += blockDim.x) {
    const int r = e / (p.deg + 1), k = e - r * (p.deg + 1);
    tab[r * pitch + k] = p.coef[e];
  }
  __syncthreads();
  const int W = (p.geom == 0) ? p.dim + 1 : p.dim;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long col = (long long)blockIdx.x * 64 + 2 * lane;
  const long long row0 = (long long)blockIdx.y * TB_ROWS;
  double xa[TB_MAXW], xb[TB_MAXW];
  const bool va = col < p.n, vb = col + 1 < p.n;
#pragma unroll
  for (int k = 0; k < TB_MAXW; ++k) {
    xa[k] = (va && k < W) ? __ldg(p.x2 + col * p.ld2 + k) : 0.0;
    xb[k] = (vb && k < W) ? __ldg(p.x2 + (col + 1) * p.ld2 + k) : 0.0;
  }
  const bool vec_ok = vb && ((p.ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 15u) == 0);
  const int shift = 52 - p.sub_bits;
  const long long sub_mask = (1LL << p.sub_bits) - 1, rem_mask = (1LL << shift) - 1;
  for (int r = warp; r < TB_ROWS; r += TB_THREADS / 32) {
    const long long i = row0 + r;
    if (i >= p.m) break;
    const double* xr = p.x1 + i * p.ld1;
    double xrow[TB_MAXW];
#pragma unroll
    for (int k = 0; k < TB_MAXW; ++k) xrow[k] = (k < W) ? __ldg(xr + k) : 0.0;
    double out[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double z;
      if (p.geom == 0) {   // reference operation order, no FMA contraction (hyperbolic_utils_torch.py:38-44,59-60)
        const double d0 = __dsub_rn(xrow[0], h ? xb[0] : xa[0]);
        double sacc = 0.0;
#pragma unroll
        for (int k = 1; k < TB_MAXW; ++k) {
          if (k < W) {
            const double dk = __dsub_rn(xrow[k], h ? xb[k] : xa[k]);
            sacc = (k == 1) ? __dmul_rn(dk, dk) : __dadd_rn(sacc, __dmul_rn(dk, dk));
          }
        }
        const double mink = __dadd_rn(-__dmul_rn(d0, d0), sacc);
        z = sqrt(__dadd_rn(fmax(0.0, mink), 1e-8));
      } else {
        z = 0.0;
#pragma unroll
        for (int k = 0; k < TB_MAXW; ++k) {
          if (k < W) {
            const double df = xrow[k] - (h ? xb[k] : xa[k]);
            z = fma(df, df, z);
          }
        }
        z = sqrt(z);   // tabulated in the distance: the small-t heat nodes make P sharp in |x - y|^2 near 0
      }
      const long long bits = __double_as_longlong(z + 1.0);
      int idx = (int)((((bits >> 52) - 1023) << p.sub_bits) | ((bits >> shift) & sub_mask));
      double sloc = 2.0 * __longlong_as_double(((bits & rem_mask) << p.sub_bits) | 0x3ff0000000000000LL) - 3.0;
      if (idx >= p.n_int) {   // beyond the tabulated range (caller's bound violated): clamp to the last interval's end
        idx = p.n_int - 1;
        sloc = 1.0;
      }
      const double* c = tab + idx * pitch;
      double acc = c[p.deg];
      for (int k = p.deg - 1; k >= 0; --k) acc = fma(acc, sloc, c[k]);
      out[h] = acc;
    }
    double* dst = p.K + i * p.ldk + col;
    if (vec_ok) {
      st_stream2(dst, out[0], out[1]);
    } else {
      if (va) st_stream(dst, out[0]);
      if (vb) st_stream(dst + 1, out[1]);
    }
  }
}

static long long profile_blocks(long long count) {
  long long blocks = (count + PWARPS - 1) / PWARPS;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * 8;
  return blocks > cap ? cap : blocks;
}

static int hyp_profile_launch(ProfParams p, cudaStream_t s) {
  if (p.kind < PK_H1 || p.kind > PK_EUCLID) return fail(MGB_ERR_ARG, "profile: kind must be 0 (H^1), 1 (H^2), 2 (H^3) or 3 (Euclidean)");
  const bool all = (p.order == 3 && p.mode == PM_SUM);
  if (!all && (p.order < 0 || p.order > 2)) return fail(MGB_ERR_ARG, "profile: derivative order must be 0, 1, 2 (or 3 = all three, forward only)");
  if (p.Pt < 1 || p.Pt > PMAXT) return fail(MGB_ERR_ARG, "profile: Pt out of range [1,128]");
  if (p.kind == PK_H2 && (p.Ps < 2 || p.Ps > PMAXP)) return fail(MGB_ERR_ARG, "profile: Ps out of range [2,1024]");
  if (p.count == 0) return MGB_OK;
  if (all) {
    const size_t smem_all = (p.kind == PK_H2) ? sizeof(double) * (size_t)PWARPS * 7 * p.Ps : 0;
    if (smem_all > 200 * 1024) {                       // very long s grids: three single-order launches
      for (int o = 0; o < 3; ++o) {
        ProfParams q = p;
        q.order = o;
        q.out = p.out + o * p.count;
        MGB_TRY(hyp_profile_launch(q, s));
      }
      return MGB_OK;
    }
    if (smem_all > 48 * 1024)
      MGB_TRY(check_cuda(cudaFuncSetAttribute(hyp_profile_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem_all), "cudaFuncSetAttribute"));
    hyp_profile_all_kernel<<<(unsigned)profile_blocks(p.count), PTHREADS, smem_all, s>>>(p);
    return after_launch("hyp_profile_all_kernel");
  }
  const size_t smem = (p.kind == PK_H2) ? sizeof(double) * (size_t)PWARPS * 4 * p.Ps : 0;
  if (smem > 48 * 1024)
    MGB_TRY(check_cuda(cudaFuncSetAttribute(hyp_profile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute"));
  hyp_profile_kernel<<<(unsigned)profile_blocks(p.count), PTHREADS, smem, s>>>(p);
  return after_launch("hyp_profile_kernel");
}

static int spd_profile_launch(SpdProfParams p, cudaStream_t s) {
  const bool all = (p.which == 3 && p.mode == PM_SUM);
  if (!all && (p.which < 0 || p.which > 2))
    return fail(MGB_ERR_ARG, "spd2 profile: which must be 0 (value), 1 (d/dDelta), 2 (d/dr2) (or 3 = all three, forward only)");
  if (p.Pt < 1 || p.Pt > PMAXT) return fail(MGB_ERR_ARG, "spd2 profile: Pt out of range [1,128]");
  if (p.Pb < 2 || p.Pb > PMAXP) return fail(MGB_ERR_ARG, "spd2 profile: Pb out of range [2,1024]");
  if (p.count == 0) return MGB_OK;
  if (all) {
    const size_t smem_all = sizeof(double) * (size_t)PWARPS * 4 * p.Pb;
    if (smem_all > 48 * 1024)
      MGB_TRY(check_cuda(cudaFuncSetAttribute(spd2_profile_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)smem_all), "cudaFuncSetAttribute"));
    spd2_profile_all_kernel<<<(unsigned)profile_blocks(p.count), PTHREADS, smem_all, s>>>(p);
    return after_launch("spd2_profile_all_kernel");
  }
  const size_t smem = sizeof(double) * (size_t)PWARPS * 3 * p.Pb;
  if (smem > 48 * 1024)
    MGB_TRY(check_cuda(cudaFuncSetAttribute(spd2_profile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute"));
  spd2_profile_kernel<<<(unsigned)profile_blocks(p.count), PTHREADS, smem, s>>>(p);
  return after_launch("spd2_profile_kernel");
}

}  // namespace mgb

using namespace mgb;

extern "C" int mgb_radial_profile(int kind, int order, const double* z, long long count, const double* tval,
                                  const double* tcoef, int Pt, double s0, double ds, int Ps, double* out,
                                  void* stream) {
  if (!z || !tval || !tcoef || !out || count < 0) return fail(MGB_ERR_ARG, "radial_profile: bad argument");
  ProfParams p{};
  p.kind = kind; p.order = order; p.mode = PM_SUM; p.in = z; p.count = count; p.tval = tval; p.tcoef = tcoef;
  p.Pt = Pt; p.s0 = s0; p.ds = ds; p.Ps = Ps; p.out = out;
  return hyp_profile_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_radial_profile_bwd_coef(int kind, int order, const double* z, long long count,
                                           const double* tval, int Pt, double s0, double ds, int Ps,
                                           const double* G, double* gtcoef, void* stream) {
  if (!z || !tval || !G || !gtcoef || count < 0) return fail(MGB_ERR_ARG, "radial_profile_bwd_coef: bad argument");
  ProfParams p{};
  p.kind = kind; p.order = order; p.mode = PM_BWD_COEF; p.in = z; p.count = count; p.tval = tval; p.tcoef = nullptr;
  p.Pt = Pt; p.s0 = s0; p.ds = ds; p.Ps = Ps; p.G = G; p.out = gtcoef;
  return hyp_profile_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_spd2_profile(int which, const double* delta, const double* r2, long long count,
                                const double* tcoef, const double* rscale, const double* bscale, int Pt,
                                const double* bval, const double* bw, int Pb, double* out, void* stream) {
  if (!delta || !r2 || !tcoef || !rscale || !bscale || !bval || !bw || !out || count < 0)
    return fail(MGB_ERR_ARG, "spd2_profile: bad argument");
  SpdProfParams p{};
  p.which = which; p.mode = PM_SUM; p.delta = delta; p.r2 = r2; p.count = count; p.tcoef = tcoef; p.rscale = rscale;
  p.bscale = bscale; p.Pt = Pt; p.bval = bval; p.bw = bw; p.Pb = Pb; p.out = out;
  return spd_profile_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_spd2_profile_bwd_coef(int which, const double* delta, const double* r2, long long count,
                                         const double* rscale, const double* bscale, int Pt, const double* bval,
                                         const double* bw, int Pb, const double* G, double* gtcoef, void* stream) {
  if (!delta || !r2 || !rscale || !bscale || !bval || !bw || !G || !gtcoef || count < 0)
    return fail(MGB_ERR_ARG, "spd2_profile_bwd_coef: bad argument");
  SpdProfParams p{};
  p.which = which; p.mode = PM_BWD_COEF; p.delta = delta; p.r2 = r2; p.count = count; p.tcoef = nullptr;
  p.rscale = rscale; p.bscale = bscale; p.Pt = Pt; p.bval = bval; p.bw = bw; p.Pb = Pb; p.G = G; p.out = gtcoef;
  return spd_profile_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_radial_table_gram(int geom, int dim, const double* x1, long long m, long long ld1, const double* x2,
                                     long long n, long long ld2, int sub_bits, int n_intervals, int degree,
                                     const double* coef, double* K, long long ldk, void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "radial_table_gram: negative size") : MGB_OK;
  if (!x1 || !x2 || !coef || !K) return fail(MGB_ERR_ARG, "radial_table_gram: null pointer");
  const int W = (geom == 0) ? dim + 1 : dim;
  if ((geom != 0 && geom != 1) || dim < 1 || W > TB_MAXW) return fail(MGB_ERR_ARG, "radial_table_gram: geometry must be H^1..H^3 or R^1..R^4");
  if (m < 0 || n < 0 || ld1 < W || ld2 < W || ldk < n) return fail(MGB_ERR_ARG, "radial_table_gram: bad sizes");
  if (sub_bits < 0 || sub_bits > 8 || n_intervals < 1 || degree < 1 || degree > 31)
    return fail(MGB_ERR_ARG, "radial_table_gram: bad table shape");
  const size_t smem = sizeof(double) * (size_t)n_intervals * (degree + 2);
  if (smem > 200 * 1024) return fail(MGB_ERR_ARG, "radial_table_gram: table exceeds shared memory");
  if (smem > 48 * 1024)
    MGB_TRY(check_cuda(cudaFuncSetAttribute(radial_table_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute"));
  TableParams p{geom, dim, x1, m, ld1, x2, n, ld2, sub_bits, n_intervals, degree, coef, K, ldk};
  const long long gy = (m + TB_ROWS - 1) / TB_ROWS;
  if (gy > 65535) {
    // split tall problems into row bands of 65535 * 64 rows
    for (long long r0 = 0; r0 < m; r0 += 65535LL * TB_ROWS) {
      const long long rows = (m - r0 < 65535LL * TB_ROWS) ? (m - r0) : 65535LL * TB_ROWS;
      TableParams q = p;
      q.x1 = x1 + r0 * ld1; q.m = rows; q.K = K + r0 * ldk;
      dim3 grid((unsigned)((n + 63) / 64), (unsigned)((rows + TB_ROWS - 1) / TB_ROWS));
      radial_table_gram_kernel<<<grid, TB_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(q);
      MGB_TRY(after_launch("radial_table_gram_kernel"));
    }
    return MGB_OK;
  }
  dim3 grid((unsigned)((n + 63) / 64), (unsigned)gy);
  radial_table_gram_kernel<<<grid, TB_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("radial_table_gram_kernel");
}
