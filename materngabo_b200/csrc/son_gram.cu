// SO(n) character-sum Gram kernel for n != 3 (n = 2, 4, 5, 6, 7; rank r = n / 2 <= 3).
//
// Reference: kernel_utils/kernels_so.py:305-362 (Matern) / :121-177 (heat), dim != 3 branch: for every pair it forms
// R = R1 R2^T, calls torch.linalg.eigvals, keeps the first member of each complex-conjugate pair IN LAPACK'S ORDER
// (the Python loop at :335-346) and substitutes those r numbers into a sympy-lambdified character polynomial.  That
// value depends on the order in which the eigen-solver lists the pairs, so there is no function
// of (R1, R2) to port.  This kernel evaluates the SAME polynomial at a canonical argument - the eigen-angles
// theta_1 >= theta_2 >= ... in (0, pi) - which makes the result a function of the pair and agrees with the reference
// wherever LAPACK lists the pairs in that order (always the positive-imaginary member first; descending in ~50 % of
// random pairs for r = 2, tests/golden/son_*.npz record which).
//
// Per pair, in registers (no eigen-solver):
//   1. R = R1 R2^T (n x n) and its symmetric part S = (R + R^T) / 2, whose eigenvalues are cos theta_i, each twice
//      (plus 1 for odd n);
//   2. eigenvalues of S by cyclic Jacobi rotations in registers (absolute accuracy ~1e-16 also for coinciding angles -
//      a closed form through tr R^k / symmetric functions loses half the digits there, e.g. on the diagonal of a Gram
//      matrix where R = I), sorted ascending, paired (the largest dropped for odd n): c_1 <= c_2 <= ... = angles
//      descending; s_i = +sqrt(1 - c_i^2);
//   3. cos(a theta_i) = T_a(c_i), sin(a theta_i) = s_i U_{a-1}(c_i) by three-term recurrences, a <= J;
//   4. K = sum over sign patterns and exponent grids G_s[a] prod_i (cos | sin)(a_i theta_i): the host
//      (materngabo_b200/_so_characters.py) folds hyper-parameters, multiplicities and the normalisation into G.
// mode 1 writes the products themselves (the basis, for hyper-parameter gradients through torch).
// Accuracy: 1e-14 for generic pairs; sin theta_i = sqrt(1 - c_i^2) carries 1e-16 / theta_i for tiny angles.
#include "mgb_common.cuh"
#include <cmath>

namespace mgb {

constexpr int kSonJMax = 12;     // largest exponent supported (n = 2 with 10 summands: 9)
constexpr int kSonTile = 16;     // 16 x 16 pairs per CTA

struct SonParams {
  const double* x1; long long m, ld1;
  const double* x2; long long n, ld2;
  const double* grids;     // [patterns][(J+1)^rank], normalised
  int J;
  double* out; long long ldo;   // mode 0: K (m x n, ld = ldo); mode 1: basis [m][n][patterns * (J+1)^rank]
  int mode;
};

template <int R>
struct SonPatterns;
template <>
struct SonPatterns<1> { static constexpr int count = 1; };
template <>
struct SonPatterns<2> { static constexpr int count = 2; };   // (c,c), (s,s)
template <>
struct SonPatterns<3> { static constexpr int count = 4; };   // (c,c,c), (c,s,s), (s,c,s), (s,s,c)

// eigenvalues of the symmetric D x D matrix a (destroyed) by cyclic Jacobi rotations, ascending in ev
template <int D>
__device__ __forceinline__ void sym_eigenvalues(double (&a)[D][D], double (&ev)[D]) {
  if constexpr (D > 1) {
    for (int sweep = 0; sweep < 15; ++sweep) {
      double off = 0.0, dg = 0.0;
#pragma unroll
      for (int p = 0; p < D; ++p) {
        dg = fma(a[p][p], a[p][p], dg);
#pragma unroll
        for (int q = p + 1; q < D; ++q) off = fma(a[p][q], a[p][q], off);
      }
      if (off <= 1e-33 * dg) break;
#pragma unroll
      for (int p = 0; p < D - 1; ++p)
#pragma unroll
        for (int q = p + 1; q < D; ++q) {
          const double apq = a[p][q];
          if (apq != 0.0) {
            const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
            const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(fma(theta, theta, 1.0)));
            const double c = rsqrt(fma(t, t, 1.0)), sn = t * c;
            a[p][p] = fma(-t, apq, a[p][p]);
            a[q][q] = fma(t, apq, a[q][q]);
            a[p][q] = a[q][p] = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
              if (k != p && k != q) {
                const double akp = a[k][p], akq = a[k][q];
                a[k][p] = a[p][k] = c * akp - sn * akq;
                a[k][q] = a[q][k] = sn * akp + c * akq;
              }
            }
          }
        }
    }
  }
#pragma unroll
  for (int k = 0; k < D; ++k) ev[k] = a[k][k];
#pragma unroll
  for (int pass = 0; pass < D - 1; ++pass)
#pragma unroll
    for (int k = 0; k < D - 1 - pass; ++k) {
      const double lo = fmin(ev[k], ev[k + 1]), hi = fmax(ev[k], ev[k + 1]);
      ev[k] = lo;
      ev[k + 1] = hi;
    }
}

template <int DIM>
__global__ void __launch_bounds__(kSonTile * kSonTile) son_gram_kernel(SonParams p) {
  constexpr int R = DIM / 2;
  constexpr int NP = SonPatterns<R>::count;
  constexpr int D2 = DIM * DIM;
  extern __shared__ __align__(16) double smem[];
  double* a_s = smem;                          // kSonTile x D2
  double* b_s = a_s + kSonTile * D2;           // kSonTile x D2
  double* g_s = b_s + kSonTile * D2;           // NP x (J+1)^R
  const int j1 = p.J + 1;
  int gsz = 1;
#pragma unroll
  for (int i = 0; i < R; ++i) gsz *= j1;
  const int tid = threadIdx.y * kSonTile + threadIdx.x;
  const long long row0 = (long long)blockIdx.y * kSonTile, col0 = (long long)blockIdx.x * kSonTile;
  for (int e = tid; e < kSonTile * D2; e += kSonTile * kSonTile) {
    const int r = e / D2, k = e - r * D2;
    a_s[e] = (row0 + r < p.m) ? p.x1[(row0 + r) * p.ld1 + k] : 0.0;
    b_s[e] = (col0 + r < p.n) ? p.x2[(col0 + r) * p.ld2 + k] : 0.0;
  }
  if (p.mode == 0)
    for (int e = tid; e < NP * gsz; e += kSonTile * kSonTile) g_s[e] = p.grids[e];
  __syncthreads();
  const long long i = row0 + threadIdx.y, j = col0 + threadIdx.x;
  if (i >= p.m || j >= p.n) return;
  const double* A = a_s + threadIdx.y * D2;
  const double* B = b_s + threadIdx.x * D2;

  // ---- R = A B^T ------------------------------------------------------------------------------------
  double Rm[DIM][DIM];
#pragma unroll
  for (int a = 0; a < DIM; ++a)
#pragma unroll
    for (int b = 0; b < DIM; ++b) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) s = fma(A[a * DIM + k], B[b * DIM + k], s);
      Rm[a][b] = s;
    }
  double c[R], s[R];
  {
    double S[DIM][DIM], ev[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a)
#pragma unroll
      for (int b = 0; b < DIM; ++b) S[a][b] = 0.5 * (Rm[a][b] + Rm[b][a]);
    sym_eigenvalues<DIM>(S, ev);
    // cos theta_i twice each (the single eigenvalue 1 of odd n is the largest and is dropped)
#pragma unroll
    for (int q = 0; q < R; ++q) c[q] = 0.5 * (ev[2 * q] + ev[2 * q + 1]);
  }
  // ascending cosines = descending angles: (c[0], theta_1 largest) is the first argument of the polynomial
#pragma unroll
  for (int q = 0; q < R; ++q) {
    c[q] = fmin(1.0, fmax(-1.0, c[q]));
    s[q] = sqrt(fmax(0.0, (1.0 - c[q]) * (1.0 + c[q])));
  }
  // ---- cos(a theta_q), sin(a theta_q), a = 0..J --------------------------------------------------
  double ct[R][kSonJMax + 1], st[R][kSonJMax + 1];
#pragma unroll
  for (int q = 0; q < R; ++q) {
    ct[q][0] = 1.0;
    st[q][0] = 0.0;
    ct[q][1] = c[q];
    st[q][1] = s[q];
    for (int a = 2; a <= p.J; ++a) {
      ct[q][a] = 2.0 * c[q] * ct[q][a - 1] - ct[q][a - 2];
      st[q][a] = 2.0 * c[q] * st[q][a - 1] - st[q][a - 2];
    }
  }
  if (p.mode == 0) {
    double acc = 0.0;
    if constexpr (R == 1) {
      for (int a = 0; a <= p.J; ++a) acc = fma(g_s[a], ct[0][a], acc);
    } else if constexpr (R == 2) {
      for (int a = 0; a <= p.J; ++a) {
        double pc = 0.0, ps = 0.0;
        for (int b = 0; b <= p.J; ++b) {
          pc = fma(g_s[a * j1 + b], ct[1][b], pc);
          ps = fma(g_s[gsz + a * j1 + b], st[1][b], ps);
        }
        acc = fma(ct[0][a], pc, acc);
        acc = fma(st[0][a], ps, acc);
      }
    } else {
      for (int a = 0; a <= p.J; ++a) {
        double ra = 0.0, rs = 0.0;      // multiplied by cos(a t1) / sin(a t1)
        for (int b = 0; b <= p.J; ++b) {
          double ccc = 0.0, css = 0.0, scs = 0.0, ssc = 0.0;
          const int base = (a * j1 + b) * j1;
          for (int d = 0; d <= p.J; ++d) {
            ccc = fma(g_s[base + d], ct[2][d], ccc);
            css = fma(g_s[gsz + base + d], st[2][d], css);
            scs = fma(g_s[2 * gsz + base + d], st[2][d], scs);
            ssc = fma(g_s[3 * gsz + base + d], ct[2][d], ssc);
          }
          ra = fma(ct[1][b], ccc, ra);
          ra = fma(st[1][b], css, ra);
          rs = fma(ct[1][b], scs, rs);
          rs = fma(st[1][b], ssc, rs);
        }
        acc = fma(ct[0][a], ra, acc);
        acc = fma(st[0][a], rs, acc);
      }
    }
    p.out[i * p.ldo + j] = acc;
  } else {
    double* o = p.out + (i * p.n + j) * (long long)(NP * gsz);
    if constexpr (R == 1) {
      for (int a = 0; a <= p.J; ++a) o[a] = ct[0][a];
    } else if constexpr (R == 2) {
      for (int a = 0; a <= p.J; ++a)
        for (int b = 0; b <= p.J; ++b) {
          o[a * j1 + b] = ct[0][a] * ct[1][b];
          o[gsz + a * j1 + b] = st[0][a] * st[1][b];
        }
    } else {
      for (int a = 0; a <= p.J; ++a)
        for (int b = 0; b <= p.J; ++b)
          for (int d = 0; d <= p.J; ++d) {
            const int f = (a * j1 + b) * j1 + d;
            o[f] = ct[0][a] * ct[1][b] * ct[2][d];
            o[gsz + f] = ct[0][a] * st[1][b] * st[2][d];
            o[2 * gsz + f] = st[0][a] * ct[1][b] * st[2][d];
            o[3 * gsz + f] = st[0][a] * st[1][b] * ct[2][d];
          }
    }
  }
}

template <int DIM>
static int son_launch(const SonParams& p, cudaStream_t s) {
  constexpr int R = DIM / 2;
  int gsz = 1;
  for (int i = 0; i < R; ++i) gsz *= p.J + 1;
  const size_t smem = sizeof(double) * (2 * kSonTile * DIM * DIM + (size_t)SonPatterns<R>::count * gsz);
  if (smem > 200 * 1024) return fail(MGB_ERR_ARG, "son_gram: coefficient grids too large for shared memory");
  MGB_TRY(check_cuda(cudaFuncSetAttribute(son_gram_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                     "cudaFuncSetAttribute"));
  dim3 grid((unsigned)((p.n + kSonTile - 1) / kSonTile), (unsigned)((p.m + kSonTile - 1) / kSonTile));
  if (grid.y > 65535) return fail(MGB_ERR_ARG, "son_gram: too many rows for one launch (split the call)");
  son_gram_kernel<DIM><<<grid, dim3(kSonTile, kSonTile), smem, s>>>(p);
  return after_launch("son_gram_kernel");
}

}  // namespace mgb

using namespace mgb;

extern "C" long long mgb_son_basis_size(int dim, int J) {
  if (dim < 2 || dim > 7 || dim == 3 || J < 0 || J > kSonJMax) return 0;
  const int r = dim / 2;
  long long g = 1;
  for (int i = 0; i < r; ++i) g *= J + 1;
  return g * (r == 1 ? 1 : (r == 2 ? 2 : 4));
}

extern "C" int mgb_son_gram(int dim, int mode, const double* x1, long long m, long long ld1, const double* x2,
                            long long n, long long ld2, const double* grids, int J, double* out, long long ldo,
                            void* stream) {
  if (dim < 2 || dim > 7 || dim == 3) return fail(MGB_ERR_ARG, "son_gram: dim must be 2, 4, 5, 6 or 7 (SO(3) has its own series kernel)");
  if (J < 0 || J > kSonJMax) return fail(MGB_ERR_ARG, "son_gram: exponent bound out of range");
  if (m < 0 || n < 0 || ld1 < dim * dim || ld2 < dim * dim) return fail(MGB_ERR_ARG, "son_gram: bad sizes / leading dimensions");
  if (mode != 0 && mode != 1) return fail(MGB_ERR_ARG, "son_gram: mode must be 0 (kernel) or 1 (basis)");
  if (m == 0 || n == 0) return MGB_OK;
  if (!x1 || !x2 || !out || (mode == 0 && (!grids || ldo < n))) return fail(MGB_ERR_ARG, "son_gram: null pointer / bad ldo");
  SonParams p{x1, m, ld1, x2, n, ld2, grids, J, out, ldo, mode};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dim) {
    case 2: return son_launch<2>(p, s);
    case 4: return son_launch<4>(p, s);
    case 5: return son_launch<5>(p, s);
    case 6: return son_launch<6>(p, s);
    default: return son_launch<7>(p, s);
  }
}
