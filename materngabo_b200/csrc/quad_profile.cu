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

static long long profile_blocks(long long count) {
  long long blocks = (count + PWARPS - 1) / PWARPS;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long cap = (long long)sms * 8;
  return blocks > cap ? cap : blocks;
}

static int hyp_profile_launch(ProfParams p, cudaStream_t s) {
  if (p.kind < PK_H1 || p.kind > PK_EUCLID) return fail(MGB_ERR_ARG, "profile: kind must be 0 (H^1), 1 (H^2), 2 (H^3) or 3 (Euclidean)");
  if (p.order < 0 || p.order > 2) return fail(MGB_ERR_ARG, "profile: derivative order must be 0, 1 or 2");
  if (p.Pt < 1 || p.Pt > PMAXT) return fail(MGB_ERR_ARG, "profile: Pt out of range [1,128]");
  if (p.kind == PK_H2 && (p.Ps < 2 || p.Ps > PMAXP)) return fail(MGB_ERR_ARG, "profile: Ps out of range [2,1024]");
  if (p.count == 0) return MGB_OK;
  const size_t smem = (p.kind == PK_H2) ? sizeof(double) * (size_t)PWARPS * 4 * p.Ps : 0;
  if (smem > 48 * 1024)
    MGB_TRY(check_cuda(cudaFuncSetAttribute(hyp_profile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                       "cudaFuncSetAttribute"));
  hyp_profile_kernel<<<(unsigned)profile_blocks(p.count), PTHREADS, smem, s>>>(p);
  return after_launch("hyp_profile_kernel");
}

static int spd_profile_launch(SpdProfParams p, cudaStream_t s) {
  if (p.which < 0 || p.which > 2) return fail(MGB_ERR_ARG, "spd2 profile: which must be 0 (value), 1 (d/dDelta) or 2 (d/dr2)");
  if (p.Pt < 1 || p.Pt > PMAXT) return fail(MGB_ERR_ARG, "spd2 profile: Pt out of range [1,128]");
  if (p.Pb < 2 || p.Pb > PMAXP) return fail(MGB_ERR_ARG, "spd2 profile: Pb out of range [2,1024]");
  if (p.count == 0) return MGB_OK;
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
