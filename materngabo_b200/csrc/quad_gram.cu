// Quadrature Gram kernels: hyperbolic space H^n (Lorentz model, n = 1,2,3) and SPD(2).
// Reference: kernel_utils/kernels_hyperbolic.py:58-88,258-389, Riemannian_utils/hyperbolic_utils_torch.py:11-60,
//            kernel_utils/kernels_spd.py:65-136,200-289, Riemannian_utils/spd_utils_torch.py:336-371.
//
// Mapping: ONE WARP PER MATRIX ENTRY.  The outer quadrature (t nodes) is spread over the lanes, the
// inner quadrature (s or b nodes) is a register loop fed by per-warp tables in shared memory, and the
// t-sum is a warp-shuffle reduction.  Pair geometry (Lorentz distance, 2x2 singular values) is
// recomputed in registers from the two input rows.
//
// H^2: the reference evaluates exp(-(b_k + d)^2 / 4t) at Ps x Pt nodes per entry.  Because the s grid is
// uniform (b_k = s0 + k ds), exp(-(b_k+d)^2/4t) = E[a][k] * g_a(d) * r_a(d)^k with a pair-independent
// table E[a][k] = exp(-b_k^2/4t_a) (shared memory) and two exponentials per (entry, t node); the sum over k is a
// Horner recurrence in r_a(d) (one multiplication + one FMA per node; relative rounding <= Ps * 2^-53).
// cosh(b_k + d) - cosh(d) is formed as cosh(d)(cosh b_k - 1) + sinh(d) sinh(b_k): all terms positive, no
// cancellation.
#include "mgb_common.cuh"
#include <cstdlib>
#include <cstring>

namespace mgb {

constexpr int QTHREADS = 256;
constexpr int QWARPS = QTHREADS / 32;
constexpr int QMAXP = 1024;   // max inner nodes
constexpr int QMAXT = 128;    // max outer nodes (4 per lane)

enum QuadMode { Q_FWD = 0, Q_BWD_COEF = 1, Q_NODES = 2, Q_BWD_TVAL = 3, Q_BWD_SCALES = 4 };

// d = 2 asinh( sqrt(max(0, -D0^2 + sum_k Dk^2) + 1e-8) / 2 ), reference operation order, no FMA contraction
__device__ __forceinline__ double lorentz_dist(const double* a, const double* b, int dim) {
  const double d0 = __dsub_rn(a[0], b[0]);
  double neg = -__dmul_rn(d0, d0);
  double s = 0.0;
  for (int k = 1; k <= dim; ++k) {
    const double dk = __dsub_rn(a[k], b[k]);
    s = (k == 1) ? __dmul_rn(dk, dk) : __dadd_rn(s, __dmul_rn(dk, dk));
  }
  const double mink = __dadd_rn(neg, s);
  const double nrm = sqrt(__dadd_rn(fmax(0.0, mink), 1e-8));
  return 2.0 * asinh(0.5 * nrm);
}

struct HypParams {
  int dim, mode;
  int geom, width;   // geom 1: Euclidean rows of `width` doubles, heat(d, t) = exp(-|x1 - x2|^2 / (4 t))
  int deriv;         // 1: evaluate d heat / d t instead of heat (Q_NODES with derivative, Q_BWD_TVAL)
  const double* x1; long long m, ld1;
  const double* x2; long long n, ld2;
  const double* dist; long long count;        // Q_NODES
  const double* tval; const double* tcoef; int Pt;
  double s0, ds; int Ps;
  const double* G; long long ldg;
  double* K; long long ldk;                    // Q_FWD
  double* out;                                 // Q_BWD_COEF: gtcoef[Pt]; Q_NODES: out[count*Pt]
};

// PAD: the table E is stored lane-major and zero-padded to 32 * NQ columns (for even NQ as double2 = the two t nodes
// a lane owns in a node pair), and the two entries' inner weights are interleaved as double2: the Horner loop is then
// two 16-byte shared loads + 4 NQ FP64 instructions per inner node with compile-time strides and no predicates.
// (NQ odd > 1 is launched as NQ + 1.)  Without PAD (large Ps, e.g. the heat kernel's Ps = 1000, Pt = 1) E is Ps x Pt.
template <int NQ, bool PAD>
__global__ void __launch_bounds__(QTHREADS) hyp_gram_kernel(HypParams p) {
  static_assert(!PAD || NQ == 1 || NQ % 2 == 0, "padded table: one or an even number of t nodes per lane");
  extern __shared__ __align__(16) double sm[];
  // layout: E | per-warp qw [2 * QWARPS][Ps] (two entries in flight) | cm1 [Ps] | sh [Ps] | wq [Ps]
  double* Es = sm;
  const int nE = (p.dim == 2) ? (PAD ? p.Ps * 32 * NQ : p.Ps * p.Pt + ((p.Ps * p.Pt) & 1)) : 0;
  const int nP = (p.dim == 2) ? p.Ps : 0;
  double* qw_all = Es + nE;
  double* cm1 = qw_all + 2 * QWARPS * nP;
  double* sh = cm1 + nP;
  double* wq = sh + nP;
  // pair-independent tables, rebuilt per CTA (Ps*Pt exponentials: negligible next to the entry loop)
  if (PAD) {
    for (int e = threadIdx.x; e < nE; e += blockDim.x) {
      int k, a;
      if (NQ == 1) {
        k = e >> 5;
        a = e & 31;
      } else {
        const int d2 = e >> 1, comp = e & 1, ln = d2 & 31, t = d2 >> 5;
        const int qp = t % (NQ / 2);
        k = t / (NQ / 2);
        a = ln + 32 * (2 * qp + comp);
      }
      const double b = p.s0 + k * p.ds;
      Es[e] = (a < p.Pt) ? exp(-(b * b) / (4.0 * p.tval[a])) : 0.0;
    }
  } else {
    for (int e = threadIdx.x; e < p.Ps * p.Pt && p.dim == 2; e += blockDim.x) {
      const int k = e / p.Pt, a = e - k * p.Pt;
      const double b = p.s0 + k * p.ds;
      Es[e] = exp(-(b * b) / (4.0 * p.tval[a]));  // E[k][a] = exp(-b_k^2 / (4 t_a))
    }
  }
  for (int e = threadIdx.x; e < nP; e += blockDim.x) {
    const double b = p.s0 + e * p.ds;
    const double hs = sinh(0.5 * b);
    cm1[e] = 2.0 * hs * hs;  // cosh(b) - 1 without cancellation
    sh[e] = sinh(b);
    wq[e] = (e == 0 || e == p.Ps - 1) ? 0.5 * p.ds : p.ds;  // trapezoid weights of the uniform s grid
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* qw = qw_all + warp * nP;
  const long long total = (p.mode == Q_NODES) ? p.count : p.m * p.n;
  const long long wstride = (long long)gridDim.x * QWARPS;

  // per-lane outer nodes a = lane + 32 q
  double tv[NQ], tc[NQ], gacc[NQ], i4t[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int a = lane + 32 * q;
    tv[q] = (a < p.Pt) ? p.tval[a] : 1.0;
    tc[q] = (a < p.Pt && p.tcoef != nullptr) ? p.tcoef[a] : 0.0;
    gacc[q] = 0.0;
    i4t[q] = 0.25 / tv[q];
  }

  // Two entries per warp iteration: the pair-independent table values E[k][a] are fetched once and used for both.
  // The pair geometry (distance, cosh d, sinh d) of the warp's next 64 entries is computed one entry pair per lane
  // and broadcast by shuffles, instead of 32 redundant copies per entry.
  double* qw0 = qw;
  double* qw1 = qw_all + (QWARPS + warp) * nP;
  double2* qw01 = reinterpret_cast<double2*>(qw_all + 2 * warp * nP);   // PAD: {entry 0, entry 1} per inner node
  for (long long ebase = (long long)blockIdx.x * QWARPS + warp; ebase < total; ebase += 64 * wstride) {
    double l_dd[2] = {0.0, 0.0}, l_dsq[2] = {0.0, 0.0}, l_ch[2] = {1.0, 1.0}, l_sh[2] = {0.0, 0.0};
    long long l_off[2] = {0, 0};   // output (Q_FWD) / upstream-gradient (backward modes) offset of the lane's entries
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      const long long el = ebase + (2 * lane + s2) * wstride;
      if (el < total) {
        if (p.mode == Q_NODES) {
          l_dd[s2] = p.dist[el];
        } else {
          const long long il = el / p.n, jl = el - il * p.n;
          l_off[s2] = il * (p.mode == Q_FWD ? p.ldk : p.ldg) + jl;
          if (p.geom == 1) {
            const double* xa = p.x1 + il * p.ld1;
            const double* xb = p.x2 + jl * p.ld2;
            for (int k = 0; k < p.width; ++k) {
              const double df = xa[k] - xb[k];
              l_dsq[s2] = fma(df, df, l_dsq[s2]);
            }
          } else {
            l_dd[s2] = lorentz_dist(p.x1 + il * p.ld1, p.x2 + jl * p.ld2, p.dim);
          }
        }
        if (p.geom != 1) l_dsq[s2] = l_dd[s2] * l_dd[s2];
        if (p.dim == 2 && p.geom != 1) {
          l_ch[s2] = cosh(l_dd[s2]);
          l_sh[s2] = sinh(l_dd[s2]);
        } else if (p.dim == 3 && p.geom != 1) {
          l_ch[s2] = (l_dd[s2] + 1e-8) / sinh(l_dd[s2] + 1e-8);   // heat kernel of H^3: exp(-d^2 / 4t) d / sinh d
        }
      }
    }
   for (int sub = 0; sub < 32; ++sub) {
    const long long e0 = ebase + 2 * sub * wstride;
    if (e0 >= total) break;
    const long long e1 = e0 + wstride;
    const bool has1 = e1 < total;
    long long ei[2] = {e0, has1 ? e1 : e0};
    long long off[2];
    double dd[2], dsq[2], chd2[2], shd2[2];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      dd[s2] = __shfl_sync(0xffffffffu, l_dd[s2], sub);
      dsq[s2] = __shfl_sync(0xffffffffu, l_dsq[s2], sub);
      chd2[s2] = __shfl_sync(0xffffffffu, l_ch[s2], sub);
      shd2[s2] = __shfl_sync(0xffffffffu, l_sh[s2], sub);
      off[s2] = __shfl_sync(0xffffffffu, l_off[s2], sub);
    }
    if (!has1) { dd[1] = dd[0]; dsq[1] = dsq[0]; chd2[1] = chd2[0]; shd2[1] = shd2[0]; }
    double h[2][NQ];
    if (p.dim == 2 && p.geom != 1) {
      __syncwarp();
      if (PAD) {
        for (int k = lane; k < p.Ps; k += 32) {
          const double bk = p.s0 + k * p.ds;
          const double c = cm1[k], sn = sh[k], wk = wq[k];
          double2 w2;
          {
            const double sk = bk + dd[0];
            w2.x = wk * sk * rsqrt(fma(chd2[0], c, shd2[0] * sn)) * (p.deriv ? sk * sk : 1.0);
          }
          {
            const double sk = bk + dd[1];
            w2.y = wk * sk * rsqrt(fma(chd2[1], c, shd2[1] * sn)) * (p.deriv ? sk * sk : 1.0);
          }
          qw01[k] = w2;
        }
      } else {
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const double chd = chd2[s2], shd = shd2[s2];
          double* dst = s2 ? qw1 : qw0;
          for (int k = lane; k < p.Ps; k += 32) {
            const double bk = p.s0 + k * p.ds;
            const double den = fma(chd, cm1[k], shd * sh[k]);
            const double sk = bk + dd[s2];
            dst[k] = wq[k] * sk * rsqrt(den) * (p.deriv ? sk * sk : 1.0);   // d/dt brings down s^2 / (4 t^2)
          }
        }
      }
      __syncwarp();
      // sum_k w_k E[k][a] r_a^k by Horner in r_a = exp(-ds d / 2t) from the last node down: one multiplication and
      // one FMA per (entry, node).  All terms are positive; rounding grows like Ps * 2^-53 (1e-13 at Ps = 1000).
      double g[2][NQ], r[2][NQ], acc[2][NQ], i2t[NQ];
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        i2t[q] = 0.5 / tv[q];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          g[s2][q] = exp(-(dsq[s2] + 2.0 * p.s0 * dd[s2]) * 0.5 * i2t[q]);
          r[s2][q] = exp(-p.ds * dd[s2] * i2t[q]);
          acc[s2][q] = 0.0;
        }
      }
      if (PAD) {
        if (NQ == 1) {
          const double* ep = Es + lane + (p.Ps - 1) * 32;
          const double2* wp = qw01 + (p.Ps - 1);
#pragma unroll 4
          for (int k = p.Ps - 1; k >= 0; --k, ep -= 32, --wp) {
            const double2 w = *wp;
            const double ev = *ep;
            acc[0][0] = fma(acc[0][0], r[0][0], w.x * ev);
            acc[1][0] = fma(acc[1][0], r[1][0], w.y * ev);
          }
        } else {
          constexpr int H = (NQ > 1) ? NQ / 2 : 1;
          const double2* ep = reinterpret_cast<const double2*>(Es) + lane + (p.Ps - 1) * (32 * H);
          const double2* wp = qw01 + (p.Ps - 1);
#pragma unroll 4
          for (int k = p.Ps - 1; k >= 0; --k, ep -= 32 * H, --wp) {
            const double2 w = *wp;
#pragma unroll
            for (int qp = 0; qp < H; ++qp) {
              const double2 ev = ep[32 * qp];
              acc[0][2 * qp] = fma(acc[0][2 * qp], r[0][2 * qp], w.x * ev.x);
              acc[1][2 * qp] = fma(acc[1][2 * qp], r[1][2 * qp], w.y * ev.x);
              if (2 * qp + 1 < NQ) {
                acc[0][2 * qp + 1] = fma(acc[0][2 * qp + 1], r[0][2 * qp + 1], w.x * ev.y);
                acc[1][2 * qp + 1] = fma(acc[1][2 * qp + 1], r[1][2 * qp + 1], w.y * ev.y);
              }
            }
          }
        }
      } else {
        for (int k = p.Ps - 1; k >= 0; --k) {
          const double w0 = qw0[k], w1 = qw1[k];
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const int a = lane + 32 * q;
            const double ev = (a < p.Pt) ? Es[k * p.Pt + a] : 0.0;
            acc[0][q] = fma(acc[0][q], r[0][q], w0 * ev);
            acc[1][q] = fma(acc[1][q], r[1][q], w1 * ev);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
          h[s2][q] = (lane + 32 * q < p.Pt) ? g[s2][q] * acc[s2][q] * (p.deriv ? i2t[q] * i2t[q] : 1.0) : 0.0;
    } else {
#pragma unroll
      for (int s2 = 0; s2 < 2; ++s2) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int a = lane + 32 * q;
          h[s2][q] = 0.0;
          if (a < p.Pt) {
            double v = exp_bf(-dsq[s2] * i4t[q]);
            if (p.dim == 3 && p.geom != 1) v = v * chd2[s2];
            if (p.deriv) v = v * dsq[s2] * (4.0 * i4t[q] * i4t[q]);
            h[s2][q] = v;
          }
        }
      }
    }
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      if (s2 == 1 && !has1) break;
      if (p.mode == Q_FWD) {
        double sum = 0.0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) sum = fma(tc[q], h[s2][q], sum);
        sum = warp_sum(sum);
        if (lane == 0) st_stream(p.K + off[s2], sum);
      } else if (p.mode == Q_BWD_COEF) {
        const double gg = p.G[off[s2]];
#pragma unroll
        for (int q = 0; q < NQ; ++q) gacc[q] = fma(gg, h[s2][q], gacc[q]);
      } else if (p.mode == Q_BWD_TVAL) {
        const double gg = p.G[off[s2]];
#pragma unroll
        for (int q = 0; q < NQ; ++q) gacc[q] = fma(gg * tc[q], h[s2][q], gacc[q]);
      } else {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int a = lane + 32 * q;
          if (a < p.Pt) p.out[ei[s2] * p.Pt + a] = h[s2][q];
        }
      }
    }
   }
  }
  if (p.mode == Q_BWD_COEF || p.mode == Q_BWD_TVAL) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int a = lane + 32 * q;
      if (a < p.Pt) atomicAdd(p.out + a, gacc[q]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// H^2 forward Gram, four entries per warp iteration.  Same nodes, tables and Horner recurrence as hyp_gram_kernel
// (PAD form); what changes is the reuse of the shared-memory traffic: with two entries in flight the node loop
// needs 5 shared-memory wavefronts (4 for the lane-major table row, 1 for the weights) per 8 FP64 instructions
// = 4 SM cycles of the FP64 pipe, i.e. the kernel is shared-memory bound (ncu: 65 % LSU wavefronts, 56 % FP64 pipe);
// with four entries it is 6 wavefronts per 16 FP64 instructions.  The pair geometry (distance, cosh, sinh, output
// offset) of the warp's next 128 entries is computed one entry per lane and parked in shared memory, and the four
// t-sums are reduced by a split butterfly (12 shuffles instead of 40).
// ------------------------------------------------------------------------------------------------
constexpr int H4E = 4;              // entries per warp iteration
constexpr int H4ROUND = 128;        // entries per geometry round (4 per lane)

template <int NQ>
__global__ void __launch_bounds__(QTHREADS, 2) h2_gram_fwd4_kernel(HypParams p) {
  static_assert(NQ == 1 || NQ == 2, "one or two t nodes per lane");
  extern __shared__ __align__(16) double sm[];
  // layout: E [Ps][32 NQ] | qw [QWARPS][Ps][4] | geo [QWARPS][4][128] | cm1 [Ps] | sh [Ps] | wq [Ps]
  double* Es = sm;
  const int nE = p.Ps * 32 * NQ;
  double* qw_all = Es + nE;
  double* geo_all = qw_all + QWARPS * H4E * p.Ps;
  double* cm1 = geo_all + QWARPS * 4 * H4ROUND;
  double* sh = cm1 + p.Ps;
  double* wq = sh + p.Ps;
  for (int e = threadIdx.x; e < nE; e += blockDim.x) {
    int k, a;
    if (NQ == 1) {
      k = e >> 5;
      a = e & 31;
    } else {
      const int d2 = e >> 1;
      k = d2 >> 5;
      a = (d2 & 31) + 32 * (e & 1);
    }
    const double b = p.s0 + k * p.ds;
    Es[e] = (a < p.Pt) ? exp(-(b * b) / (4.0 * p.tval[a])) : 0.0;
  }
  for (int e = threadIdx.x; e < p.Ps; e += blockDim.x) {
    const double b = p.s0 + e * p.ds;
    const double hs = sinh(0.5 * b);
    cm1[e] = 2.0 * hs * hs;
    sh[e] = sinh(b);
    wq[e] = (e == 0 || e == p.Ps - 1) ? 0.5 * p.ds : p.ds;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2* qw = reinterpret_cast<double2*>(qw_all + warp * H4E * p.Ps);     // [k][2] double2 = 4 entries
  double* geo = geo_all + warp * 4 * H4ROUND;
  double* g_dd = geo;
  double* g_ch = geo + H4ROUND;
  double* g_sh = geo + 2 * H4ROUND;
  long long* g_off = reinterpret_cast<long long*>(geo + 3 * H4ROUND);
  const long long total = p.m * p.n;
  const long long wstride = (long long)gridDim.x * QWARPS;

  double i2t[NQ], tc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int a = lane + 32 * q;
    i2t[q] = 0.5 / ((a < p.Pt) ? p.tval[a] : 1.0);
    tc[q] = (a < p.Pt) ? p.tcoef[a] : 0.0;
  }

  for (long long ebase = (long long)blockIdx.x * QWARPS + warp; ebase < total; ebase += H4ROUND * wstride) {
    __syncwarp();
#pragma unroll 1
    for (int mth = 0; mth < H4ROUND / 32; ++mth) {
      const int t = lane + 32 * mth;
      const long long el = ebase + t * wstride;
      double d = 0.0, c = 1.0, sn = 0.0;
      long long off = -1;
      if (el < total) {
        const long long il = el / p.n, jl = el - il * p.n;
        d = lorentz_dist(p.x1 + il * p.ld1, p.x2 + jl * p.ld2, 2);
        c = cosh(d);
        sn = sinh(d);
        off = il * p.ldk + jl;
      }
      g_dd[t] = d;
      g_ch[t] = c;
      g_sh[t] = sn;
      g_off[t] = off;
    }
    __syncwarp();
#pragma unroll 1
    for (int sub = 0; sub < H4ROUND / H4E; ++sub) {
      if (ebase + (long long)(H4E * sub) * wstride >= total) break;
      double dd[H4E];
      {
        const double2 a01 = reinterpret_cast<const double2*>(g_dd)[2 * sub], a23 = reinterpret_cast<const double2*>(g_dd)[2 * sub + 1];
        dd[0] = a01.x; dd[1] = a01.y; dd[2] = a23.x; dd[3] = a23.y;
      }
      __syncwarp();          // the previous iteration's node loop has finished reading qw
      {
        const double2 c01 = reinterpret_cast<const double2*>(g_ch)[2 * sub], c23 = reinterpret_cast<const double2*>(g_ch)[2 * sub + 1];
        const double2 s01 = reinterpret_cast<const double2*>(g_sh)[2 * sub], s23 = reinterpret_cast<const double2*>(g_sh)[2 * sub + 1];
        for (int k = lane; k < p.Ps; k += 32) {
          const double bk = p.s0 + k * p.ds;
          const double c = cm1[k], sn = sh[k], wk = wq[k];
          double2 w01, w23;
          w01.x = wk * (bk + dd[0]) * rsqrt(fma(c01.x, c, s01.x * sn));
          w01.y = wk * (bk + dd[1]) * rsqrt(fma(c01.y, c, s01.y * sn));
          w23.x = wk * (bk + dd[2]) * rsqrt(fma(c23.x, c, s23.x * sn));
          w23.y = wk * (bk + dd[3]) * rsqrt(fma(c23.y, c, s23.y * sn));
          qw[2 * k] = w01;
          qw[2 * k + 1] = w23;
        }
      }
      __syncwarp();
      double r[H4E][NQ], acc[H4E][NQ];
#pragma unroll
      for (int j = 0; j < H4E; ++j)
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          r[j][q] = exp_bf(-p.ds * dd[j] * i2t[q]);
          acc[j][q] = 0.0;
        }
      {
        const double2* wp = qw + 2 * (p.Ps - 1);
        if (NQ == 1) {
          const double* ep = Es + lane + (p.Ps - 1) * 32;
#pragma unroll 4
          for (int k = p.Ps - 1; k >= 0; --k, ep -= 32, wp -= 2) {
            const double2 w01 = wp[0], w23 = wp[1];
            const double ev = *ep;
            acc[0][0] = fma(acc[0][0], r[0][0], w01.x * ev);
            acc[1][0] = fma(acc[1][0], r[1][0], w01.y * ev);
            acc[2][0] = fma(acc[2][0], r[2][0], w23.x * ev);
            acc[3][0] = fma(acc[3][0], r[3][0], w23.y * ev);
          }
        } else {
          const double2* ep = reinterpret_cast<const double2*>(Es) + lane + (p.Ps - 1) * 32;
#pragma unroll 4
          for (int k = p.Ps - 1; k >= 0; --k, ep -= 32, wp -= 2) {
            const double2 w01 = wp[0], w23 = wp[1];
            const double2 ev = *ep;
            acc[0][0] = fma(acc[0][0], r[0][0], w01.x * ev.x);
            acc[1][0] = fma(acc[1][0], r[1][0], w01.y * ev.x);
            acc[2][0] = fma(acc[2][0], r[2][0], w23.x * ev.x);
            acc[3][0] = fma(acc[3][0], r[3][0], w23.y * ev.x);
            acc[0][NQ - 1] = fma(acc[0][NQ - 1], r[0][NQ - 1], w01.x * ev.y);
            acc[1][NQ - 1] = fma(acc[1][NQ - 1], r[1][NQ - 1], w01.y * ev.y);
            acc[2][NQ - 1] = fma(acc[2][NQ - 1], r[2][NQ - 1], w23.x * ev.y);
            acc[3][NQ - 1] = fma(acc[3][NQ - 1], r[3][NQ - 1], w23.y * ev.y);
          }
        }
      }
      double sum[H4E];
#pragma unroll
      for (int j = 0; j < H4E; ++j) {
        sum[j] = 0.0;
        const double cexp = -(dd[j] * dd[j] + 2.0 * p.s0 * dd[j]) * 0.5;
#pragma unroll
        for (int q = 0; q < NQ; ++q) sum[j] = fma(tc[q], exp_bf(cexp * i2t[q]) * acc[j][q], sum[j]);
      }
      // split butterfly: after the xor-16 and xor-8 stages each group of 8 lanes owns one entry
      {
        const bool hi = lane & 16;
        const double keep0 = hi ? sum[2] : sum[0], keep1 = hi ? sum[3] : sum[1];
        const double send0 = hi ? sum[0] : sum[2], send1 = hi ? sum[1] : sum[3];
        const double a0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 16);
        const double a1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 16);
        const bool mid = lane & 8;
        const double keep = mid ? a1 : a0, send = mid ? a0 : a1;
        double v = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        if ((lane & 7) == 0) {
          const long long off = g_off[H4E * sub + (lane >> 3)];
          if (off >= 0) st_stream(p.K + off, v);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// SPD(2)
// ------------------------------------------------------------------------------------------------
struct SpdParams {
  int use_chol, mode;
  const double* x1; long long m, ld1;
  const double* x2; long long n, ld2;
  const double* tcoef; const double* rscale; const double* bscale; int Pt;
  const double* bval; const double* bw; int Pb;
  const double* G; long long ldg;
  double* K; long long ldk;
  double* out;
  double* out2;
};

// Mandel (s11, s22, sqrt2 s12) -> 2x2 (optionally its lower Cholesky factor); returns a,b,c,d row-major
__device__ __forceinline__ void spd2_load(const double* v, int chol, double& a, double& b, double& c, double& d) {
  const double s11 = v[0], s22 = v[1], s12 = v[2] / 1.4142135623730951;
  if (!chol) {
    a = s11; b = s12; c = s12; d = s22;
  } else {
    a = sqrt(s11);
    b = 0.0;
    c = s12 / a;
    d = sqrt(s22 - c * c);
  }
}

// log singular values (H1 >= H2) of P = A * B^{-1}
__device__ __forceinline__ void spd2_logsv(const double* va, const double* vb, int chol, double& H1, double& H2) {
  double a, b, c, d, e, f, g, h;
  spd2_load(va, chol, a, b, c, d);
  spd2_load(vb, chol, e, f, g, h);
  const double det = e * h - f * g;
  const double i11 = h / det, i12 = -f / det, i21 = -g / det, i22 = e / det;
  const double p11 = a * i11 + b * i21, p12 = a * i12 + b * i22;
  const double p21 = c * i11 + d * i21, p22 = c * i12 + d * i22;
  const double q1 = hypot(p11 + p22, p12 - p21), q2 = hypot(p11 - p22, p12 + p21);
  const double s1 = 0.5 * (q1 + q2);
  const double s2 = fabs(p11 * p22 - p12 * p21) / s1;
  H1 = log(s1);
  H2 = log(s2);
}

__global__ void __launch_bounds__(QTHREADS) spd2_gram_kernel(SpdParams p) {
  extern __shared__ __align__(16) double sm[];
  // layout: bval [Pb] | bw [Pb] | sinh b [Pb] | cosh b [Pb] | per-warp (qb [Pb], cb [Pb])
  double* bv = sm;
  double* bw = bv + p.Pb;
  double* shb = bw + p.Pb;
  double* chb = shb + p.Pb;
  double* scratch = chb + p.Pb;
  for (int e = threadIdx.x; e < p.Pb; e += blockDim.x) {
    const double b = p.bval[e];
    bv[e] = b;
    bw[e] = p.bw[e];
    shb[e] = sinh(b);
    chb[e] = cosh(b);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* qb = scratch + (size_t)warp * 2 * p.Pb;
  double* cb = qb + p.Pb;
  const long long total = p.m * p.n;
  const long long wstride = (long long)gridDim.x * QWARPS;

  double tc[QMAXT / 32], rs[QMAXT / 32], bs[QMAXT / 32], gacc[QMAXT / 32], gacc2[QMAXT / 32];
#pragma unroll
  for (int q = 0; q < QMAXT / 32; ++q) {
    const int a = lane + 32 * q;
    tc[q] = (a < p.Pt && p.tcoef != nullptr) ? p.tcoef[a] : 0.0;
    rs[q] = (a < p.Pt) ? p.rscale[a] : 0.0;
    bs[q] = (a < p.Pt) ? p.bscale[a] : 0.0;
    gacc[q] = 0.0;
    gacc2[q] = 0.0;
  }

  for (long long e = (long long)blockIdx.x * QWARPS + warp; e < total; e += wstride) {
    const long long i = e / p.n, j = e - i * p.n;
    double H1, H2;
    spd2_logsv(p.x1 + i * p.ld1, p.x2 + j * p.ld2, p.use_chol, H1, H2);
    const double delta = H1 - H2, r2 = H1 * H1 + H2 * H2;
    const double shd = sinh(delta), chd = cosh(delta);
    __syncwarp();
    for (int k = lane; k < p.Pb; k += 32) {
      const double b = bv[k];
      const double sbd = fma(shb[k], chd, chb[k] * shd);  // sinh(b + delta)
      qb[k] = bw[k] * (2.0 * b + delta) * rsqrt(shb[k] * sbd);
      cb[k] = b * (b + delta);
    }
    __syncwarp();
    double h[QMAXT / 32], hc[QMAXT / 32];
#pragma unroll
    for (int q = 0; q < QMAXT / 32; ++q) {
      const int a = lane + 32 * q;
      h[q] = 0.0;
      hc[q] = 0.0;
      if (a < p.Pt) {
        const double sc = -bs[q];
        double acc = 0.0, accc = 0.0;
        if (p.mode == Q_BWD_SCALES) {
          for (int k = 0; k < p.Pb; ++k) {
            const double ev = qb[k] * exp(cb[k] * sc);
            acc += ev;
            accc = fma(cb[k], ev, accc);   // sum_b q_b c_b exp(-c_b bscale)
          }
        } else {
          for (int k = 0; k < p.Pb; ++k) acc = fma(qb[k], exp(cb[k] * sc), acc);
        }
        const double er = exp(-r2 * rs[q]);
        h[q] = acc * er;
        hc[q] = accc * er;
      }
    }
    if (p.mode == Q_FWD) {
      double s = 0.0;
#pragma unroll
      for (int q = 0; q < QMAXT / 32; ++q) s = fma(tc[q], h[q], s);
      s = warp_sum(s);
      if (lane == 0) st_stream(p.K + i * p.ldk + j, s);
    } else if (p.mode == Q_BWD_COEF) {
      const double g = p.G[i * p.ldg + j];
#pragma unroll
      for (int q = 0; q < QMAXT / 32; ++q) gacc[q] = fma(g, h[q], gacc[q]);
    } else {
      // dK/drscale_a = -r2 tcoef_a h_a ; dK/dbscale_a = -tcoef_a exp(-r2 rscale_a) sum_b q_b c_b exp(-c_b bscale_a)
      const double g = p.G[i * p.ldg + j];
#pragma unroll
      for (int q = 0; q < QMAXT / 32; ++q) {
        gacc[q] = fma(-g * tc[q] * r2, h[q], gacc[q]);
        gacc2[q] = fma(-g * tc[q], hc[q], gacc2[q]);
      }
    }
  }
  if (p.mode == Q_BWD_COEF || p.mode == Q_BWD_SCALES) {
#pragma unroll
    for (int q = 0; q < QMAXT / 32; ++q) {
      const int a = lane + 32 * q;
      if (a < p.Pt) {
        atomicAdd(p.out + a, gacc[q]);
        if (p.mode == Q_BWD_SCALES) atomicAdd(p.out2 + a, gacc2[q]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// SPD(2), lattice form of the forward kernel.  The reference's grids are both geometric (b = logspace(-5, 1, Pb),
// t = logspace(.., .., Pt): kernels_spd.py:262-265), so the pair-dependent exponent of node (a, b),
//     -b_b (b_b + Delta) bscale_a = -b_b^2 bscale_a  -  Delta * (b_b bscale_a),
// has a pair-INDEPENDENT first part (table E[b][a], shared by the CTA) and a second part whose coefficient
// b_b bscale_a = g_min tau^k lives on a one-dimensional lattice k = sb*b + sa*(Pt-1-a) whenever the two log-steps are
// commensurate (12 : 7 for Pb = Pt).  A warp therefore evaluates exp(-Delta g_min tau^k) once per lattice point
// (1198 exponentials for 64 x 64 nodes instead of 4096) and the node loop becomes two shared-memory reads, one
// multiplication and one FMA per node.  Same nodes, same weights, same trapezoid sums as spd2_gram_kernel.
// ------------------------------------------------------------------------------------------------
struct SpdLatticeParams {
  int use_chol;
  const double* x1; long long m, ld1;
  const double* x2; long long n, ld2;
  const double* tcoef; const double* rscale; const double* bscale; int Pt;
  const double* bval; const double* bw; int Pb;
  int sb, sa, nk;
  double g_min, log_tau;
  double* K; long long ldk;
};

constexpr int LTHREADS = 384;   // 12 warps: per-warp lattice tables (nk + Pb doubles) keep it to one CTA per SM
constexpr int LWARPS = LTHREADS / 32;
constexpr int LNQ = 2;   // outer nodes per lane (Pt <= 64)

__global__ void __launch_bounds__(LTHREADS, 1) spd2_lattice_kernel(SpdLatticeParams p) {
  extern __shared__ __align__(16) double sm[];
  // layout: E [Pb*Pt] | G [nk] | bv, bw, shb, chb [Pb each] | per warp: X [nk] | qb [Pb]
  double* Es = sm;
  double* Gs = Es + (size_t)p.Pb * p.Pt;
  double* bv = Gs + p.nk;
  double* bw = bv + p.Pb;
  double* shb = bw + p.Pb;
  double* chb = shb + p.Pb;
  double* wbase = chb + p.Pb;
  for (int e = threadIdx.x; e < p.Pb * p.Pt; e += blockDim.x) {
    const int b = e / p.Pt, a = e - b * p.Pt;
    const double bb = p.bval[b];
    Es[e] = exp(-bb * bb * p.bscale[a]);
  }
  for (int k = threadIdx.x; k < p.nk; k += blockDim.x) Gs[k] = -p.g_min * exp(k * p.log_tau);
  for (int e = threadIdx.x; e < p.Pb; e += blockDim.x) {
    const double b = p.bval[e];
    bv[e] = b;
    bw[e] = p.bw[e];
    shb[e] = sinh(b);
    chb[e] = cosh(b);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* X = wbase + (size_t)warp * (p.nk + p.Pb);
  double* qb = X + p.nk;
  double tc[LNQ], rs[LNQ];
  int koff[LNQ];
#pragma unroll
  for (int q = 0; q < LNQ; ++q) {
    const int a = lane + 32 * q;
    tc[q] = (a < p.Pt) ? p.tcoef[a] : 0.0;
    rs[q] = (a < p.Pt) ? p.rscale[a] : 0.0;
    koff[q] = (a < p.Pt) ? p.sa * (p.Pt - 1 - a) : 0;
  }
  const long long total = p.m * p.n;
  const long long wstride = (long long)gridDim.x * LWARPS;
  for (long long ebase = (long long)blockIdx.x * LWARPS + warp; ebase < total; ebase += 32 * wstride) {
   // pair geometry of this warp's next 32 entries, one entry per lane (instead of 32 redundant copies per entry)
   double g_delta = 0.0, g_r2 = 0.0, g_sh = 0.0, g_ch = 1.0;
   {
     const long long el = ebase + lane * wstride;
     if (el < total) {
       const long long il = el / p.n, jl = el - il * p.n;
       double H1, H2;
       spd2_logsv(p.x1 + il * p.ld1, p.x2 + jl * p.ld2, p.use_chol, H1, H2);
       g_delta = H1 - H2;
       g_r2 = H1 * H1 + H2 * H2;
       g_sh = sinh(g_delta);
       g_ch = cosh(g_delta);
     }
   }
   for (int sub = 0; sub < 32; ++sub) {
    const long long e = ebase + sub * wstride;
    if (e >= total) break;
    const long long i = e / p.n, j = e - i * p.n;
    const double delta = __shfl_sync(0xffffffffu, g_delta, sub), r2 = __shfl_sync(0xffffffffu, g_r2, sub);
    const double shd = __shfl_sync(0xffffffffu, g_sh, sub), chd = __shfl_sync(0xffffffffu, g_ch, sub);
    __syncwarp();
    // lattice exponentials, four independent chains per lane
    int k = lane;
    for (; k + 96 < p.nk; k += 128) {
      const double e0 = exp(delta * Gs[k]), e1 = exp(delta * Gs[k + 32]);
      const double e2 = exp(delta * Gs[k + 64]), e3 = exp(delta * Gs[k + 96]);
      X[k] = e0; X[k + 32] = e1; X[k + 64] = e2; X[k + 96] = e3;
    }
    for (; k < p.nk; k += 32) X[k] = exp(delta * Gs[k]);
    for (int b = lane; b < p.Pb; b += 32) {
      const double bb = bv[b];
      const double sbd = fma(shb[b], chd, chb[b] * shd);  // sinh(b + delta)
      qb[b] = bw[b] * (2.0 * bb + delta) * rsqrt(shb[b] * sbd);
    }
    __syncwarp();
    double acc[LNQ];
#pragma unroll
    for (int q = 0; q < LNQ; ++q) acc[q] = 0.0;
    for (int b = 0; b < p.Pb; ++b) {
      const double qv = qb[b];
      const double* Eb = Es + b * p.Pt;
      const double* Xb = X + p.sb * b;
#pragma unroll
      for (int q = 0; q < LNQ; ++q) {
        const int a = lane + 32 * q;
        if (a < p.Pt) acc[q] = fma(qv * Eb[a], Xb[koff[q]], acc[q]);
      }
    }
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < LNQ; ++q) {
      if (lane + 32 * q < p.Pt) s = fma(tc[q] * exp(-r2 * rs[q]), acc[q], s);
    }
    s = warp_sum(s);
    if (lane == 0) st_stream(p.K + i * p.ldk + j, s);
   }
  }
}

// Two entries per warp iteration.  spd2_lattice_kernel is bound by shared-memory wavefronts (ncu: 69 % LSU, 43 % FP64
// pipe): per inner node b it reads 1 (weight) + 4 (table row E[b][.]) + 4 (lattice gather) wavefronts for 4 FP64
// instructions.  Here the table row and the weight pair serve two entries (13 wavefronts per 8 FP64 instructions), the
// table is stored lane-major as double2 = the two t nodes of a lane (zero padded, no predicates), the two entries'
// lattice exponentials are interleaved ({X0[k], X1[k]}: one 16-byte gather per t node) and so are the weights.
constexpr int L2THREADS = 256;
constexpr int L2WARPS = L2THREADS / 32;

__global__ void __launch_bounds__(L2THREADS, 1) spd2_lattice2_kernel(SpdLatticeParams p) {
  extern __shared__ __align__(16) double sm[];
  // layout: E2 [Pb][32] double2 | G [nk (+1)] | bv, bw, shb, chb [Pb each (+pad)] | per warp: X01 [nk] double2 | qb01 [Pb] double2
  double2* E2 = reinterpret_cast<double2*>(sm);
  double* Gs = sm + (size_t)p.Pb * 64;
  const int nkp = p.nk + (p.nk & 1), Pbp = p.Pb + (p.Pb & 1);
  double* bv = Gs + nkp;
  double* bw = bv + Pbp;
  double* shb = bw + Pbp;
  double* chb = shb + Pbp;
  double2* wbase = reinterpret_cast<double2*>(chb + Pbp);
  for (int e = threadIdx.x; e < p.Pb * 64; e += blockDim.x) {
    const int d2 = e >> 1, b = d2 >> 5, a = (d2 & 31) + 32 * (e & 1);
    const double bb = p.bval[b];
    sm[e] = (a < p.Pt) ? exp(-bb * bb * p.bscale[a]) : 0.0;
  }
  for (int k = threadIdx.x; k < p.nk; k += blockDim.x) Gs[k] = -p.g_min * exp(k * p.log_tau);
  for (int e = threadIdx.x; e < p.Pb; e += blockDim.x) {
    const double b = p.bval[e];
    bv[e] = b;
    bw[e] = p.bw[e];
    shb[e] = sinh(b);
    chb[e] = cosh(b);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double2* X = wbase + (size_t)warp * (p.nk + p.Pb);
  double2* qb = X + p.nk;
  double tc[LNQ], rs[LNQ];
  int koff[LNQ];
#pragma unroll
  for (int q = 0; q < LNQ; ++q) {
    const int a = lane + 32 * q;
    tc[q] = (a < p.Pt) ? p.tcoef[a] : 0.0;
    rs[q] = (a < p.Pt) ? p.rscale[a] : 0.0;
    koff[q] = (a < p.Pt) ? p.sa * (p.Pt - 1 - a) : 0;
  }
  const long long total = p.m * p.n;
  const long long wstride = (long long)gridDim.x * L2WARPS;
  for (long long ebase = (long long)blockIdx.x * L2WARPS + warp; ebase < total; ebase += 32 * wstride) {
    // pair geometry of this warp's next 32 entries, one entry per lane
    double g_delta = 0.0, g_r2 = 0.0, g_sh = 0.0, g_ch = 1.0;
    long long g_off = -1;
    {
      const long long el = ebase + lane * wstride;
      if (el < total) {
        const long long il = el / p.n, jl = el - il * p.n;
        double H1, H2;
        spd2_logsv(p.x1 + il * p.ld1, p.x2 + jl * p.ld2, p.use_chol, H1, H2);
        g_delta = H1 - H2;
        g_r2 = H1 * H1 + H2 * H2;
        g_sh = sinh(g_delta);
        g_ch = cosh(g_delta);
        g_off = il * p.ldk + jl;
      }
    }
#pragma unroll 1
    for (int sub = 0; sub < 16; ++sub) {
      if (ebase + (long long)(2 * sub) * wstride >= total) break;
      const double d0 = __shfl_sync(0xffffffffu, g_delta, 2 * sub), d1 = __shfl_sync(0xffffffffu, g_delta, 2 * sub + 1);
      __syncwarp();          // the previous iteration's node loop has finished reading X and qb
      {
        int k = lane;
        for (; k + 96 < p.nk; k += 128) {
          const double ga = Gs[k], gb = Gs[k + 32], gc = Gs[k + 64], gd = Gs[k + 96];
          double2 xa, xb, xc, xd;
          xa.x = exp_bf(d0 * ga); xa.y = exp_bf(d1 * ga);
          xb.x = exp_bf(d0 * gb); xb.y = exp_bf(d1 * gb);
          xc.x = exp_bf(d0 * gc); xc.y = exp_bf(d1 * gc);
          xd.x = exp_bf(d0 * gd); xd.y = exp_bf(d1 * gd);
          X[k] = xa;
          X[k + 32] = xb;
          X[k + 64] = xc;
          X[k + 96] = xd;
        }
        for (; k + 32 < p.nk; k += 64) {
          const double ga = Gs[k], gb = Gs[k + 32];
          double2 xa, xb;
          xa.x = exp_bf(d0 * ga); xa.y = exp_bf(d1 * ga);
          xb.x = exp_bf(d0 * gb); xb.y = exp_bf(d1 * gb);
          X[k] = xa;
          X[k + 32] = xb;
        }
        for (; k < p.nk; k += 32) {
          const double ga = Gs[k];
          double2 xa;
          xa.x = exp_bf(d0 * ga); xa.y = exp_bf(d1 * ga);
          X[k] = xa;
        }
      }
      {
        const double sh0 = __shfl_sync(0xffffffffu, g_sh, 2 * sub), sh1 = __shfl_sync(0xffffffffu, g_sh, 2 * sub + 1);
        const double ch0 = __shfl_sync(0xffffffffu, g_ch, 2 * sub), ch1 = __shfl_sync(0xffffffffu, g_ch, 2 * sub + 1);
        for (int b = lane; b < p.Pb; b += 32) {
          const double bb = bv[b], sb_ = shb[b], cb_ = chb[b], wb = bw[b];
          double2 qv;
          qv.x = wb * (2.0 * bb + d0) * rsqrt(sb_ * fma(sb_, ch0, cb_ * sh0));   // sinh(b + delta) = sinh b cosh delta + cosh b sinh delta
          qv.y = wb * (2.0 * bb + d1) * rsqrt(sb_ * fma(sb_, ch1, cb_ * sh1));
          qb[b] = qv;
        }
      }
      __syncwarp();
      double acc[2][LNQ];
#pragma unroll
      for (int q = 0; q < LNQ; ++q) acc[0][q] = acc[1][q] = 0.0;
      {
        const double2* Ep = E2 + lane;
        const double2* X0 = X + koff[0];
        const double2* X1 = X + koff[1];
        const int sb = p.sb;
#pragma unroll 4
        for (int b = 0; b < p.Pb; ++b, Ep += 32, X0 += sb, X1 += sb) {
          const double2 qv = qb[b];
          const double2 ev = *Ep;
          const double2 xa = *X0, xb = *X1;
          acc[0][0] = fma(qv.x * ev.x, xa.x, acc[0][0]);
          acc[1][0] = fma(qv.y * ev.x, xa.y, acc[1][0]);
          acc[0][1] = fma(qv.x * ev.y, xb.x, acc[0][1]);
          acc[1][1] = fma(qv.y * ev.y, xb.y, acc[1][1]);
        }
      }
      const double r20 = __shfl_sync(0xffffffffu, g_r2, 2 * sub), r21 = __shfl_sync(0xffffffffu, g_r2, 2 * sub + 1);
      const long long off0 = __shfl_sync(0xffffffffu, g_off, 2 * sub), off1 = __shfl_sync(0xffffffffu, g_off, 2 * sub + 1);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int q = 0; q < LNQ; ++q) {
        s0 = fma(tc[q] * exp_bf(-r20 * rs[q]), acc[0][q], s0);
        s1 = fma(tc[q] * exp_bf(-r21 * rs[q]), acc[1][q], s1);
      }
      // split butterfly: lanes 0-15 end up with entry 0, lanes 16-31 with entry 1
      const bool hi = lane & 16;
      double v = (hi ? s1 : s0) + __shfl_xor_sync(0xffffffffu, hi ? s0 : s1, 16);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      if ((lane & 15) == 0) {
        const long long off = hi ? off1 : off0;
        if (off >= 0) st_stream(p.K + off, v);
      }
    }
  }
}

// The same two-entry scheme with PAIRS of warps: spd2_lattice2_kernel is latency-bound at 8 warps per SM (the per-warp
// lattice tables fill the shared memory; ncu: issue slots 50 %, FP64 pipe 48 %, top stall "wait").  Here two warps
// share one pair of lattice tables: each computes half of the lattice exponentials and half of the inner weights,
// then owns one t node per lane (warp 0 of the pair a < 32, warp 1 a >= 32) in the node loop.  Same tables per SM,
// twice the warps; the two warps meet at a 64-thread named barrier twice per entry pair.
constexpr int LPTHREADS = 512;
constexpr int LPPAIRS = LPTHREADS / 64;

__device__ __forceinline__ void pair_barrier(int pair) {
  asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory");
}

__global__ void __launch_bounds__(LPTHREADS, 1) spd2_lattice2p_kernel(SpdLatticeParams p) {
  extern __shared__ __align__(16) double sm[];
  // layout: E [Pb][64] | G [nk (+1)] | bv, bw, shb, chb [Pb each (+pad)] | per pair: X01 [nk] double2 | qb01 [Pb] double2 | part [2]
  double* Es = sm;
  double* Gs = sm + (size_t)p.Pb * 64;
  const int nkp = p.nk + (p.nk & 1), Pbp = p.Pb + (p.Pb & 1);
  double* bv = Gs + nkp;
  double* bw = bv + Pbp;
  double* shb = bw + Pbp;
  double* chb = shb + Pbp;
  double2* wbase = reinterpret_cast<double2*>(chb + Pbp);
  for (int e = threadIdx.x; e < p.Pb * 64; e += blockDim.x) {
    const int b = e >> 6, a = e & 63;
    const double bb = p.bval[b];
    Es[e] = (a < p.Pt) ? exp(-bb * bb * p.bscale[a]) : 0.0;
  }
  for (int k = threadIdx.x; k < p.nk; k += blockDim.x) Gs[k] = -p.g_min * exp(k * p.log_tau);
  for (int e = threadIdx.x; e < p.Pb; e += blockDim.x) {
    const double b = p.bval[e];
    bv[e] = b;
    bw[e] = p.bw[e];
    shb[e] = sinh(b);
    chb[e] = cosh(b);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, pair = warp >> 1, half = warp & 1;
  double2* X = wbase + (size_t)pair * (p.nk + p.Pb + 1);
  double2* qb = X + p.nk;
  double* part = reinterpret_cast<double*>(qb + p.Pb);
  const int a = lane + 32 * half;
  const double tc = (a < p.Pt) ? p.tcoef[a] : 0.0;
  const double rs = (a < p.Pt) ? p.rscale[a] : 0.0;
  const int koff = (a < p.Pt) ? p.sa * (p.Pt - 1 - a) : 0;
  const long long total = p.m * p.n;
  const long long wstride = (long long)gridDim.x * LPPAIRS;
  for (long long ebase = (long long)blockIdx.x * LPPAIRS + pair; ebase < total; ebase += 32 * wstride) {
    // pair geometry of the next 32 entries, one per lane (both warps of the pair hold a copy)
    double g_delta = 0.0, g_r2 = 0.0, g_sh = 0.0, g_ch = 1.0;
    long long g_off = -1;
    {
      const long long el = ebase + lane * wstride;
      if (el < total) {
        const long long il = el / p.n, jl = el - il * p.n;
        double H1, H2;
        spd2_logsv(p.x1 + il * p.ld1, p.x2 + jl * p.ld2, p.use_chol, H1, H2);
        g_delta = H1 - H2;
        g_r2 = H1 * H1 + H2 * H2;
        g_sh = sinh(g_delta);
        g_ch = cosh(g_delta);
        g_off = il * p.ldk + jl;
      }
    }
#pragma unroll 1
    for (int sub = 0; sub < 16; ++sub) {
      if (ebase + (long long)(2 * sub) * wstride >= total) break;      // uniform over the pair: same ebase, same total
      const double d0 = __shfl_sync(0xffffffffu, g_delta, 2 * sub), d1 = __shfl_sync(0xffffffffu, g_delta, 2 * sub + 1);
      {
        // this warp's half of the lattice: k = lane + 32 (2 i + half)
        int k = lane + 32 * half;
        for (; k + 192 < p.nk; k += 256) {
          const double ga = Gs[k], gb = Gs[k + 64], gc = Gs[k + 128], gd = Gs[k + 192];
          double2 xa, xb, xc, xd;
          xa.x = exp_bf(d0 * ga); xa.y = exp_bf(d1 * ga);
          xb.x = exp_bf(d0 * gb); xb.y = exp_bf(d1 * gb);
          xc.x = exp_bf(d0 * gc); xc.y = exp_bf(d1 * gc);
          xd.x = exp_bf(d0 * gd); xd.y = exp_bf(d1 * gd);
          X[k] = xa;
          X[k + 64] = xb;
          X[k + 128] = xc;
          X[k + 192] = xd;
        }
        for (; k < p.nk; k += 64) {
          const double ga = Gs[k];
          double2 xa;
          xa.x = exp_bf(d0 * ga); xa.y = exp_bf(d1 * ga);
          X[k] = xa;
        }
      }
      {
        const double sh0 = __shfl_sync(0xffffffffu, g_sh, 2 * sub), sh1 = __shfl_sync(0xffffffffu, g_sh, 2 * sub + 1);
        const double ch0 = __shfl_sync(0xffffffffu, g_ch, 2 * sub), ch1 = __shfl_sync(0xffffffffu, g_ch, 2 * sub + 1);
        for (int b = lane + 32 * half; b < p.Pb; b += 64) {
          const double bb = bv[b], sb_ = shb[b], cb_ = chb[b], wb = bw[b];
          double2 qv;
          qv.x = wb * (2.0 * bb + d0) * rsqrt(sb_ * fma(sb_, ch0, cb_ * sh0));
          qv.y = wb * (2.0 * bb + d1) * rsqrt(sb_ * fma(sb_, ch1, cb_ * sh1));
          qb[b] = qv;
        }
      }
      pair_barrier(pair);      // (B) tables complete
      double acc0 = 0.0, acc1 = 0.0;
      {
        const double* Ep = Es + a;
        const double2* Xp = X + koff;
        const int sb = p.sb;
#pragma unroll 4
        for (int b = 0; b < p.Pb; ++b, Ep += 64, Xp += sb) {
          const double2 qv = qb[b];
          const double ev = *Ep;
          const double2 xv = *Xp;
          acc0 = fma(qv.x * ev, xv.x, acc0);
          acc1 = fma(qv.y * ev, xv.y, acc1);
        }
      }
      const double r20 = __shfl_sync(0xffffffffu, g_r2, 2 * sub), r21 = __shfl_sync(0xffffffffu, g_r2, 2 * sub + 1);
      const double s0 = tc * exp_bf(-r20 * rs) * acc0, s1 = tc * exp_bf(-r21 * rs) * acc1;
      // split butterfly: lanes 0-15 end up with entry 0, lanes 16-31 with entry 1
      const bool hi = lane & 16;
      double v = (hi ? s1 : s0) + __shfl_xor_sync(0xffffffffu, hi ? s0 : s1, 16);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      const long long off0 = __shfl_sync(0xffffffffu, g_off, 2 * sub), off1 = __shfl_sync(0xffffffffu, g_off, 2 * sub + 1);
      if (half == 1 && (lane & 15) == 0) part[lane >> 4] = v;
      pair_barrier(pair);      // (A) node loop finished on both warps (tables may be rewritten), partial sums visible
      if (half == 0 && (lane & 15) == 0) {
        const long long off = hi ? off1 : off0;
        if (off >= 0) st_stream(p.K + off, v + part[lane >> 4]);
      }
    }
  }
}

static size_t spd2_lattice2p_smem(int Pb, int nk) {
  const int nkp = nk + (nk & 1), Pbp = Pb + (Pb & 1);
  return sizeof(double) * ((size_t)Pb * 64 + nkp + 4 * (size_t)Pbp + (size_t)LPPAIRS * 2 * (nk + Pb + 1));
}

static size_t spd2_lattice2_smem(int Pb, int nk) {
  const int nkp = nk + (nk & 1), Pbp = Pb + (Pb & 1);
  return sizeof(double) * ((size_t)Pb * 64 + nkp + 4 * (size_t)Pbp + (size_t)L2WARPS * 2 * (nk + Pb));
}

// ------------------------------------------------------------------------------------------------
static int sm_count() { return device_sm_count(); }

// MGB_H2_PATH=generic keeps the two-entries-per-iteration kernel for the forward Gram as well (A/B measurements, tests)
static bool hyp_force_generic() {
  const char* e = getenv("MGB_H2_PATH");
  return e != nullptr && strcmp(e, "generic") == 0;
}

static int hyp_launch(HypParams p, cudaStream_t s) {
  if (p.dim < 1 || p.dim > 3) return fail(MGB_ERR_ARG, "hyperbolic: dim must be 1, 2 or 3 (NotImplementedError in the reference)");
  if (p.Pt < 1 || p.Pt > QMAXT) return fail(MGB_ERR_ARG, "hyperbolic: Pt out of range [1,128]");
  if (p.dim == 2 && (p.Ps < 2 || p.Ps > QMAXP)) return fail(MGB_ERR_ARG, "hyperbolic: Ps out of range [2,1024]");
  const long long total = (p.mode == Q_NODES) ? p.count : p.m * p.n;
  if (total == 0) return MGB_OK;
  size_t smem = 0;
  const int nq = (p.Pt + 31) / 32;
  bool pad = false;
  if (p.dim == 2) {
    const size_t rest = 3 * (size_t)p.Ps + (size_t)2 * QWARPS * p.Ps;
    const int nqp = (nq == 3) ? 4 : nq;
    const size_t padded = sizeof(double) * ((size_t)p.Ps * 32 * nqp + rest);
    // padded lane-major table while two CTAs per SM still fit (the register file allows no more)
    pad = p.geom != 1 && padded <= 100 * 1024;
    const size_t nE = (size_t)p.Ps * p.Pt;
    smem = pad ? padded : sizeof(double) * (nE + (nE & 1) + rest);
    if (smem > 200 * 1024) return fail(MGB_ERR_ARG, "hyperbolic: Ps*Pt table exceeds shared memory");
  }
  if (p.dim == 2 && p.geom != 1 && p.mode == Q_FWD && !p.deriv && nq <= 2 && total >= 4096 && !hyp_force_generic()) {
    const size_t smem4 = sizeof(double) * ((size_t)p.Ps * 32 * nq + (size_t)QWARPS * H4E * p.Ps + (size_t)QWARPS * 4 * H4ROUND +
                                           3 * (size_t)p.Ps);
    if (smem4 <= 110 * 1024) {
      long long blocks4 = (total + QWARPS * H4E - 1) / (QWARPS * H4E);
      if (blocks4 > 2LL * sm_count()) blocks4 = 2LL * sm_count();
      if (nq == 1) {
        MGB_TRY(check_cuda(cudaFuncSetAttribute(h2_gram_fwd4_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4),
                           "cudaFuncSetAttribute"));
        h2_gram_fwd4_kernel<1><<<(unsigned)blocks4, QTHREADS, smem4, s>>>(p);
      } else {
        MGB_TRY(check_cuda(cudaFuncSetAttribute(h2_gram_fwd4_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4),
                           "cudaFuncSetAttribute"));
        h2_gram_fwd4_kernel<2><<<(unsigned)blocks4, QTHREADS, smem4, s>>>(p);
      }
      return after_launch("h2_gram_fwd4_kernel");
    }
  }
  long long blocks = (total + QWARPS - 1) / QWARPS;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
#define MGB_HYP_LAUNCH(NQV, PADV)                                                                                       \
  do {                                                                                                                  \
    if (smem > 48 * 1024)                                                                                               \
      MGB_TRY(check_cuda(cudaFuncSetAttribute(hyp_gram_kernel<NQV, PADV>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                              (int)smem), "cudaFuncSetAttribute"));                                     \
    hyp_gram_kernel<NQV, PADV><<<(unsigned)blocks, QTHREADS, smem, s>>>(p);                                             \
  } while (0)
  if (pad) {
    if (nq == 1) MGB_HYP_LAUNCH(1, true);
    else if (nq == 2) MGB_HYP_LAUNCH(2, true);
    else MGB_HYP_LAUNCH(4, true);
  } else {
    if (nq == 1) MGB_HYP_LAUNCH(1, false);
    else if (nq == 2) MGB_HYP_LAUNCH(2, false);
    else if (nq == 3) MGB_HYP_LAUNCH(3, false);
    else MGB_HYP_LAUNCH(4, false);
  }
#undef MGB_HYP_LAUNCH
  return after_launch("hyp_gram_kernel");
}

static int spd_launch(SpdParams p, cudaStream_t s) {
  if (p.Pt < 1 || p.Pt > QMAXT) return fail(MGB_ERR_ARG, "spd2: Pt out of range [1,128]");
  if (p.Pb < 2 || p.Pb > QMAXP) return fail(MGB_ERR_ARG, "spd2: Pb out of range [2,1024]");
  const long long total = p.m * p.n;
  if (total == 0) return MGB_OK;
  const size_t smem = sizeof(double) * (4 * (size_t)p.Pb + (size_t)QWARPS * 2 * p.Pb);
  MGB_TRY(check_cuda(cudaFuncSetAttribute(spd2_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                     "cudaFuncSetAttribute"));
  long long blocks = (total + QWARPS - 1) / QWARPS;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  spd2_gram_kernel<<<(unsigned)blocks, QTHREADS, smem, s>>>(p);
  return after_launch("spd2_gram_kernel");
}

// Pair scalars of the quadrature kernels for a candidate block against the training set (acquisition path: the
// profile kernels take them as flat lists).  kind 0: Lorentz distance (rows width wide); 1: squared Euclidean
// distance; 2 / 3: SPD(2) (Delta, r2) from the matrices / from their Cholesky factors.
__global__ void __launch_bounds__(256) pair_scalars_kernel(int kind, int width, const double* x1, long long m, long long ld1,
                                                           const double* x2, long long n, long long ld2, double* out0,
                                                           double* out1) {
  const long long total = m * n;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / n, j = e - i * n;
    const double* a = x1 + i * ld1;
    const double* b = x2 + j * ld2;
    if (kind == 0) {
      out0[e] = lorentz_dist(a, b, width - 1);
    } else if (kind == 1) {
      double d = 0.0;
      for (int k = 0; k < width; ++k) {
        const double df = a[k] - b[k];
        d = fma(df, df, d);
      }
      out0[e] = d;
    } else {
      double H1, H2;
      spd2_logsv(a, b, kind == 3, H1, H2);
      out0[e] = H1 - H2;
      out1[e] = H1 * H1 + H2 * H2;
    }
  }
}

}  // namespace mgb

using namespace mgb;

extern "C" int mgb_pair_scalars(int kind, int width, const double* x1, long long m, long long ld1, const double* x2,
                                long long n, long long ld2, double* out0, double* out1, void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "pair_scalars: negative size") : MGB_OK;
  if (!x1 || !x2 || !out0 || (kind >= 2 && !out1)) return fail(MGB_ERR_ARG, "pair_scalars: null pointer");
  if (kind < 0 || kind > 3) return fail(MGB_ERR_ARG, "pair_scalars: kind must be 0..3");
  if (kind >= 2) width = 3;
  if (width < 1 || (kind == 0 && (width < 2 || width > 4)) || ld1 < width || ld2 < width) return fail(MGB_ERR_ARG, "pair_scalars: bad widths");
  long long blocks = (m * n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pair_scalars_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(kind, width, x1, m, ld1, x2, n, ld2, out0, out1);
  return after_launch("pair_scalars_kernel");
}


extern "C" int mgb_hyperbolic_gram_fwd(int dim, const double* x1, long long m, long long ld1, const double* x2,
                                       long long n, long long ld2, const double* tval, const double* tcoef, int Pt,
                                       double s0, double ds, int Ps, double* K, long long ldk, void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "mgb_hyperbolic_gram_fwd: negative size") : MGB_OK;  // empty block
  if (!x1 || !x2 || !tval || !tcoef || !K) return fail(MGB_ERR_ARG, "hyperbolic_gram_fwd: null pointer");
  if (m < 0 || n < 0 || ld1 < dim + 1 || ld2 < dim + 1 || ldk < n) return fail(MGB_ERR_ARG, "hyperbolic_gram_fwd: bad sizes");
  HypParams p{};
  p.dim = dim; p.mode = Q_FWD; p.x1 = x1; p.m = m; p.ld1 = ld1; p.x2 = x2; p.n = n; p.ld2 = ld2;
  p.tval = tval; p.tcoef = tcoef; p.Pt = Pt; p.s0 = s0; p.ds = ds; p.Ps = Ps; p.K = K; p.ldk = ldk;
  return hyp_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_hyperbolic_gram_bwd_coef(int dim, const double* x1, long long m, long long ld1, const double* x2,
                                            long long n, long long ld2, const double* tval, int Pt, double s0, double ds,
                                            int Ps, const double* G, long long ldg, double* gtcoef, void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "mgb_hyperbolic_gram_bwd_coef: negative size") : MGB_OK;  // empty block
  if (!x1 || !x2 || !tval || !G || !gtcoef) return fail(MGB_ERR_ARG, "hyperbolic_gram_bwd_coef: null pointer");
  if (m < 0 || n < 0 || ld1 < dim + 1 || ld2 < dim + 1 || ldg < n) return fail(MGB_ERR_ARG, "hyperbolic_gram_bwd_coef: bad sizes");
  HypParams p{};
  p.dim = dim; p.mode = Q_BWD_COEF; p.x1 = x1; p.m = m; p.ld1 = ld1; p.x2 = x2; p.n = n; p.ld2 = ld2;
  p.tval = tval; p.tcoef = nullptr; p.Pt = Pt; p.s0 = s0; p.ds = ds; p.Ps = Ps; p.G = G; p.ldg = ldg; p.out = gtcoef;
  return hyp_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_hyperbolic_gram_bwd_tval(int dim, const double* x1, long long m, long long ld1, const double* x2,
                                            long long n, long long ld2, const double* tval, const double* tcoef, int Pt,
                                            double s0, double ds, int Ps, const double* G, long long ldg, double* gtval,
                                            void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "mgb_hyperbolic_gram_bwd_tval: negative size") : MGB_OK;  // empty block
  if (!x1 || !x2 || !tval || !tcoef || !G || !gtval) return fail(MGB_ERR_ARG, "hyperbolic_gram_bwd_tval: null pointer");
  if (m < 0 || n < 0 || ld1 < dim + 1 || ld2 < dim + 1 || ldg < n) return fail(MGB_ERR_ARG, "hyperbolic_gram_bwd_tval: bad sizes");
  HypParams p{};
  p.dim = dim; p.mode = Q_BWD_TVAL; p.deriv = 1; p.x1 = x1; p.m = m; p.ld1 = ld1; p.x2 = x2; p.n = n; p.ld2 = ld2;
  p.tval = tval; p.tcoef = tcoef; p.Pt = Pt; p.s0 = s0; p.ds = ds; p.Ps = Ps; p.G = G; p.ldg = ldg; p.out = gtval;
  return hyp_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_hyperbolic_heat_nodes(int dim, int derivative, const double* dist, long long count, const double* tval,
                                         int Pt, double s0, double ds, int Ps, double* out, void* stream) {
  if (!dist || !tval || !out || count < 0) return fail(MGB_ERR_ARG, "hyperbolic_heat_nodes: bad argument");
  HypParams p{};
  p.deriv = derivative ? 1 : 0;
  p.dim = dim; p.mode = Q_NODES; p.dist = dist; p.count = count; p.tval = tval; p.Pt = Pt;
  p.s0 = s0; p.ds = ds; p.Ps = Ps; p.out = out;
  return hyp_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_euclidean_gram_fwd(const double* x1, long long m, long long ld1, const double* x2, long long n,
                                      long long ld2, int width, const double* tval, const double* tcoef, int Pt,
                                      double* K, long long ldk, void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "mgb_euclidean_gram_fwd: negative size") : MGB_OK;  // empty block
  if (!x1 || !x2 || !tval || !tcoef || !K) return fail(MGB_ERR_ARG, "euclidean_gram_fwd: null pointer");
  if (m < 0 || n < 0 || width < 1 || ld1 < width || ld2 < width || ldk < n) return fail(MGB_ERR_ARG, "euclidean_gram_fwd: bad sizes");
  HypParams p{};
  p.dim = 1; p.geom = 1; p.width = width; p.mode = Q_FWD; p.x1 = x1; p.m = m; p.ld1 = ld1; p.x2 = x2; p.n = n; p.ld2 = ld2;
  p.tval = tval; p.tcoef = tcoef; p.Pt = Pt; p.K = K; p.ldk = ldk;
  return hyp_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_euclidean_gram_bwd_coef(const double* x1, long long m, long long ld1, const double* x2, long long n,
                                           long long ld2, int width, const double* tval, int Pt, const double* G,
                                           long long ldg, double* gtcoef, void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "mgb_euclidean_gram_bwd_coef: negative size") : MGB_OK;  // empty block
  if (!x1 || !x2 || !tval || !G || !gtcoef) return fail(MGB_ERR_ARG, "euclidean_gram_bwd_coef: null pointer");
  if (m < 0 || n < 0 || width < 1 || ld1 < width || ld2 < width || ldg < n) return fail(MGB_ERR_ARG, "euclidean_gram_bwd_coef: bad sizes");
  HypParams p{};
  p.dim = 1; p.geom = 1; p.width = width; p.mode = Q_BWD_COEF; p.x1 = x1; p.m = m; p.ld1 = ld1; p.x2 = x2; p.n = n; p.ld2 = ld2;
  p.tval = tval; p.Pt = Pt; p.G = G; p.ldg = ldg; p.out = gtcoef;
  return hyp_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_spd2_gram_fwd(int use_cholesky, const double* x1, long long m, long long ld1, const double* x2,
                                 long long n, long long ld2, const double* tcoef, const double* rscale,
                                 const double* bscale, int Pt, const double* bval, const double* bw, int Pb, double* K,
                                 long long ldk, void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "mgb_spd2_gram_fwd: negative size") : MGB_OK;  // empty block
  if (!x1 || !x2 || !tcoef || !rscale || !bscale || !bval || !bw || !K) return fail(MGB_ERR_ARG, "spd2_gram_fwd: null pointer");
  if (m < 0 || n < 0 || ld1 < 3 || ld2 < 3 || ldk < n) return fail(MGB_ERR_ARG, "spd2_gram_fwd: bad sizes");
  SpdParams p{};
  p.use_chol = use_cholesky; p.mode = Q_FWD; p.x1 = x1; p.m = m; p.ld1 = ld1; p.x2 = x2; p.n = n; p.ld2 = ld2;
  p.tcoef = tcoef; p.rscale = rscale; p.bscale = bscale; p.Pt = Pt; p.bval = bval; p.bw = bw; p.Pb = Pb; p.K = K; p.ldk = ldk;
  return spd_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_spd2_gram_fwd_lattice(int use_cholesky, const double* x1, long long m, long long ld1,
                                         const double* x2, long long n, long long ld2, const double* tcoef,
                                         const double* rscale, const double* bscale, int Pt, const double* bval,
                                         const double* bw, int Pb, int sb, int sa, double g_min, double log_tau,
                                         double* K, long long ldk, void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "mgb_spd2_gram_fwd_lattice: negative size") : MGB_OK;  // empty block
  if (!x1 || !x2 || !tcoef || !rscale || !bscale || !bval || !bw || !K) return fail(MGB_ERR_ARG, "spd2_gram_fwd_lattice: null pointer");
  if (m < 0 || n < 0 || ld1 < 3 || ld2 < 3 || ldk < n) return fail(MGB_ERR_ARG, "spd2_gram_fwd_lattice: bad sizes");
  if (Pt < 1 || Pt > 32 * LNQ) return fail(MGB_ERR_ARG, "spd2_gram_fwd_lattice: Pt out of range [1,64]");
  if (Pb < 2 || Pb > QMAXP) return fail(MGB_ERR_ARG, "spd2_gram_fwd_lattice: Pb out of range [2,1024]");
  if (sb < 1 || sa < 0 || !(g_min > 0.0) || !(log_tau > 0.0)) return fail(MGB_ERR_ARG, "spd2_gram_fwd_lattice: bad lattice");
  if (m == 0 || n == 0) return MGB_OK;
  SpdLatticeParams p{};
  p.use_chol = use_cholesky; p.x1 = x1; p.m = m; p.ld1 = ld1; p.x2 = x2; p.n = n; p.ld2 = ld2;
  p.tcoef = tcoef; p.rscale = rscale; p.bscale = bscale; p.Pt = Pt; p.bval = bval; p.bw = bw; p.Pb = Pb;
  p.sb = sb; p.sa = sa; p.nk = sb * (Pb - 1) + sa * (Pt - 1) + 1; p.g_min = g_min; p.log_tau = log_tau;
  p.K = K; p.ldk = ldk;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    const char* force = getenv("MGB_SPD2_PATH");       // "lattice1" / "lattice2" / "pairs": pick a kernel (A/B measurements, tests)
    const size_t smem2p = spd2_lattice2p_smem(Pb, p.nk);
    if (smem2p <= 227 * 1024 && (force == nullptr || strcmp(force, "pairs") == 0)) {
      MGB_TRY(check_cuda(cudaFuncSetAttribute(spd2_lattice2p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2p),
                         "cudaFuncSetAttribute"));
      const long long totalp = m * n;
      long long blocksp = (totalp + 2 * LPPAIRS - 1) / (2 * LPPAIRS);
      if (blocksp > (long long)sm_count()) blocksp = sm_count();
      spd2_lattice2p_kernel<<<(unsigned)blocksp, LPTHREADS, smem2p, s>>>(p);
      return after_launch("spd2_lattice2p_kernel");
    }
    const size_t smem2 = spd2_lattice2_smem(Pb, p.nk);
    if (smem2 <= 227 * 1024 && !(force != nullptr && strcmp(force, "lattice1") == 0)) {
      MGB_TRY(check_cuda(cudaFuncSetAttribute(spd2_lattice2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2),
                         "cudaFuncSetAttribute"));
      const long long total2 = m * n;
      long long blocks2 = (total2 + 2 * L2WARPS - 1) / (2 * L2WARPS);
      if (blocks2 > (long long)sm_count()) blocks2 = sm_count();
      spd2_lattice2_kernel<<<(unsigned)blocks2, L2THREADS, smem2, s>>>(p);
      return after_launch("spd2_lattice2_kernel");
    }
  }
  const size_t smem = sizeof(double) * ((size_t)Pb * Pt + p.nk + 4 * (size_t)Pb + (size_t)LWARPS * (p.nk + Pb));
  if (smem > 200 * 1024) return fail(MGB_ERR_ARG, "spd2_gram_fwd_lattice: lattice too large for shared memory");
  MGB_TRY(check_cuda(cudaFuncSetAttribute(spd2_lattice_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                     "cudaFuncSetAttribute"));
  const long long total = m * n;
  long long blocks = (total + LWARPS - 1) / LWARPS;
  const long long cap = (long long)sm_count();
  if (blocks > cap) blocks = cap;
  spd2_lattice_kernel<<<(unsigned)blocks, LTHREADS, smem, s>>>(p);
  return after_launch("spd2_lattice_kernel");
}

extern "C" int mgb_spd2_gram_bwd_scales(int use_cholesky, const double* x1, long long m, long long ld1, const double* x2,
                                        long long n, long long ld2, const double* tcoef, const double* rscale,
                                        const double* bscale, int Pt, const double* bval, const double* bw, int Pb,
                                        const double* G, long long ldg, double* g_rscale, double* g_bscale, void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "mgb_spd2_gram_bwd_scales: negative size") : MGB_OK;  // empty block
  if (!x1 || !x2 || !tcoef || !rscale || !bscale || !bval || !bw || !G || !g_rscale || !g_bscale)
    return fail(MGB_ERR_ARG, "spd2_gram_bwd_scales: null pointer");
  if (m < 0 || n < 0 || ld1 < 3 || ld2 < 3 || ldg < n) return fail(MGB_ERR_ARG, "spd2_gram_bwd_scales: bad sizes");
  SpdParams p{};
  p.use_chol = use_cholesky; p.mode = Q_BWD_SCALES; p.x1 = x1; p.m = m; p.ld1 = ld1; p.x2 = x2; p.n = n; p.ld2 = ld2;
  p.tcoef = tcoef; p.rscale = rscale; p.bscale = bscale; p.Pt = Pt; p.bval = bval; p.bw = bw; p.Pb = Pb;
  p.G = G; p.ldg = ldg; p.out = g_rscale; p.out2 = g_bscale;
  return spd_launch(p, static_cast<cudaStream_t>(stream));
}

extern "C" int mgb_spd2_gram_bwd_coef(int use_cholesky, const double* x1, long long m, long long ld1, const double* x2,
                                      long long n, long long ld2, const double* rscale, const double* bscale, int Pt,
                                      const double* bval, const double* bw, int Pb, const double* G, long long ldg,
                                      double* gtcoef, void* stream) {
  if (m == 0 || n == 0) return (m < 0 || n < 0) ? fail(MGB_ERR_ARG, "mgb_spd2_gram_bwd_coef: negative size") : MGB_OK;  // empty block
  if (!x1 || !x2 || !rscale || !bscale || !bval || !bw || !G || !gtcoef) return fail(MGB_ERR_ARG, "spd2_gram_bwd_coef: null pointer");
  if (m < 0 || n < 0 || ld1 < 3 || ld2 < 3 || ldg < n) return fail(MGB_ERR_ARG, "spd2_gram_bwd_coef: bad sizes");
  SpdParams p{};
  p.use_chol = use_cholesky; p.mode = Q_BWD_COEF; p.x1 = x1; p.m = m; p.ld1 = ld1; p.x2 = x2; p.n = n; p.ld2 = ld2;
  p.tcoef = nullptr; p.rscale = rscale; p.bscale = bscale; p.Pt = Pt; p.bval = bval; p.bw = bw; p.Pb = Pb;
  p.G = G; p.ldg = ldg; p.out = gtcoef;
  return spd_launch(p, static_cast<cudaStream_t>(stream));
}
