// Posterior variance on the INT8 tensor cores, hand-written: tcgen05.mma kind::i8 with the accumulators in tensor
// memory, operands staged by the bulk-copy engine, float64 recombination in the epilogue (opt-in path).
//
//   V = K* L^-T  (m x n),   var_i = kss - sum_j V_ij^2          (reference: gpytorch's exact prediction strategy behind
//                                                                 model.posterior(X), manifold_optimize.py:195,254,337)
//
// Error-free slicing as in int8_posterior.cu: every row of K* and of L^-1 is scaled by a power of two and cut into
// S = 7 signed 8-bit slices (round to nearest, a carry pass keeps every digit in [-128, 127]),
// a_ik = 2^ea_i sum_p A_p[i][k] 2^(-8 (p + 1)) (likewise b_jk), so that
//      V_ij = 2^(ea_i + eb_j) sum_{d < 7} 2^(-8 (d + 2)) C_d[i][j],     C_d = sum_{p + q = d} A_p B_q^T   (int32, exact:
//      7 n 128^2 < 2^31 for n <= 16384).  56 bits per operand and 28 of the 49 slice products: the variance lands within
//      ~1e-13 of max var of the exact result at n = 5000 (7-bit slices as in int8_posterior.cu: ~2e-11).
// What differs from int8_posterior.cu (library GEMMs + a recombination pass over 7 int32 planes in HBM):
//   * ONE kernel per candidate block.  A CTA owns a 128-candidate x 64-column tile of V; the seven planes C_0..C_6 of
//     that tile live in tensor memory (7 x 64 = 448 of the 512 columns) for the whole k loop, the 28 slice products of a
//     k-block are 112 tcgen05.mma (M 128, N 64, K 32) issued by one thread, and the epilogue warps read the planes back
//     with tcgen05.ld, recombine in float64 and reduce to sum_j V_ij^2 - no int32 plane ever reaches HBM;
//   * the k loop of a column tile stops at the diagonal of L^-1 (rows j only meet k <= j): the exact triangle instead
//     of five column blocks (50 % of the dense work instead of 60 %), with no compacted operand copies;
//   * the slicing kernels write the operands directly in the tensor core's shared-memory image (K-major, 128-byte
//     swizzle, one contiguous 16 KB / 8 KB tile per slice and k-block), so a tile arrives with ONE cp.async.bulk and no
//     tensor map.
// Pipeline per CTA (192 threads): warp 0 = producer (bulk copies: the seven B tiles of a k-block as one 56 KB copy into
// a 2-deep ring, the A tiles one by one into a 6-deep ring), warp 1 = MMA issuer (lane 0), warps 2-5 = epilogue (one
// TMEM lane quarter each).  mbarriers: a_full / a_empty, b_full / b_empty (tcgen05.commit releases a slot when the
// MMAs reading it have retired), tmem_full / tmem_empty between the issuer and the epilogue.
// Work items (row block, column tile) are laid out so that 8 row blocks (their A slices: 37 MB at n = 5000) stay in
// L2 while the column tiles sweep over them, longest k loops first, and dealt round-robin to the persistent CTAs.
#include <cstdlib>

#include "mgb_common.cuh"
#include "mgb_tcgen05.cuh"

namespace mgb {

constexpr int F8_S = 7;                   // slices
constexpr int F8_BITS = 8;                // bits per slice: digits in [-128, 127] after the carry pass of the slicing kernel
// F8_BM = 128 candidates per tile = UMMA M = TMEM lanes (mgb_tcgen05.cuh)
constexpr int F8_BN = 64;                 // columns per tile = UMMA N
constexpr int F8_BK = 128;                // int8 per k-block = one swizzle row
constexpr int F8_A_TILE = F8_BM * F8_BK;  // 16 KB
constexpr int F8_B_TILE = F8_BN * F8_BK;  // 8 KB
constexpr int F8_A_STAGES = 6;
constexpr int F8_B_STAGES = 2;
constexpr int F8_THREADS = 192;
constexpr int F8_TMEM_COLS = 512;
constexpr long long F8_MAX_N = 16384;     // 7 n 128^2 < 2^31
constexpr size_t F8_SMEM = 1024 + (size_t)F8_B_STAGES * F8_S * F8_B_TILE + (size_t)F8_A_STAGES * F8_A_TILE + 256;

__host__ __device__ inline long long f8_round_up(long long v, long long m) { return (v + m - 1) / m * m; }

struct F8Geom {
  long long np;     // n rounded up to 128
  int KB, CT;       // k-blocks of 128, column tiles of 64 that hold columns < n
  __host__ __device__ long long b_tile(int ct, int kb, int q) const { return (((long long)ct * KB + kb) * F8_S + q) * F8_B_TILE; }
  __host__ __device__ long long a_tile(long long rb, int kb, int p) const { return ((rb * KB + kb) * F8_S + p) * (long long)F8_A_TILE; }
  __host__ __device__ long long b_bytes() const { return b_tile((int)(np / F8_BN), 0, 0); }
  __host__ __device__ long long a_bytes(long long m) const { return a_tile(f8_round_up(m, F8_BM) / F8_BM, 0, 0); }
  __host__ __device__ static int nkb(int ct) { return ct / 2 + 1; }      // k-blocks up to the diagonal of column tile ct
};

__host__ __device__ inline F8Geom f8_geom(long long n) {
  F8Geom g;
  g.np = f8_round_up(n, F8_BK);
  g.KB = (int)(g.np / F8_BK);
  g.CT = (int)((n + F8_BN - 1) / F8_BN);
  return g;
}


// One warp per row: e = exponent with |row| 2^-e <= 126/256, then S round-to-nearest slices of 8 bits, written into the
// tiled operand image.  is_b = 0: K* rows (128-row tiles; also mean_i = mean_const + sum_j K*_ij alpha_j);
// is_b = 1: rows of L^-1 (64-row tiles, lower triangle only, k-blocks beyond the tile's diagonal are never read).
// Rows in [rows, rows_pad) are written as zeros; row0 = position of src's first row in the operand.  Per iteration the warp takes 128 consecutive columns - lane l the four
// columns 4 l .. 4 l + 3 (two 16-byte loads, 1 KB per warp) - and stores one 32-bit word per slice: the 32 words of a
// warp are the 128 bytes of one (swizzled) tile row.  Rounding by the 1.5 * 2^52 trick: t = x + M is x rounded to an
// integer in the low mantissa bits (|x| <= 64), t - M that integer as a double, x - (t - M) the exact remainder.
__global__ void __launch_bounds__(256) f8_slice_kernel(const double* src, long long ld, long long row0, long long rows, long long rows_pad,
                                                       int cols, F8Geom g, int is_b, signed char* dst, int* exps, double* scales,
                                                       const double* alpha, double mean_const, double* mean) {
  // src holds rows [row0, row0 + rows) of the operand (a sub-block of the candidate chunk); rows up to rows_pad are padding
  const int lane = threadIdx.x & 31;
  const long long lrow = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (lrow >= rows_pad) return;
  const long long row = row0 + lrow;
  const bool valid = lrow < rows;
  const double* s = src + (valid ? lrow : 0) * ld;
  const int used = !valid ? 0 : (is_b ? (int)min((long long)cols, row + 1) : cols);
  const bool vec = ((reinterpret_cast<uintptr_t>(s) & 15) == 0);      // 16-byte loads need an aligned row
  double mx = 0.0, mu = 0.0;
  if (alpha != nullptr) {
    for (int c = lane; c < used; c += 32) {
      const double v = s[c];
      mx = fmax(mx, fabs(v));
      mu = fma(v, alpha[c], mu);
    }
    mu = warp_sum(mu);
    if (lane == 0 && valid) mean[row] = mean_const + mu;
  } else {
    for (int c = lane; c < used; c += 32) mx = fmax(mx, fabs(s[c]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  int e = 0;
  if (mx > 0.0) {
    const double f = frexp(mx, &e);      // mx = f 2^e, f in [1/2, 1)
    e += (f > 0.984375) ? 2 : 1;         // |row| 2^-e <= 126 / 256: the leading digit stays below 127 and can take a carry
  }
  if (lane == 0) {
    if (exps != nullptr) exps[row] = e;
    if (scales != nullptr) scales[row] = ldexp(1.0, e);
  }
  const double down = ldexp(1.0, -e);                // exact power of two
  const int tile_rows = is_b ? F8_BN : F8_BM;
  const long long tile_bytes = (long long)tile_rows * F8_BK;
  const long long rt = row / tile_rows;
  const int rr = (int)(row % tile_rows);
  const int kb_end = is_b ? min(g.KB, F8Geom::nkb((int)rt)) : g.KB;
  // this lane's word inside a tile row: 16-byte chunk lane / 4 (swizzled with the row), word lane % 4
  signed char* base = dst + rt * g.KB * F8_S * tile_bytes + f8_swizzled(rr, lane >> 2) + ((lane & 3) << 2);
  constexpr double kMagic = 6755399441055744.0;      // 1.5 * 2^52
  for (int kb = 0; kb < kb_end; ++kb) {
    const int c0 = kb * F8_BK + 4 * lane;
    double r[4];
    if (vec && c0 + 4 <= used) {
      const double2 v0 = *reinterpret_cast<const double2*>(s + c0);
      const double2 v1 = *reinterpret_cast<const double2*>(s + c0 + 2);
      r[0] = v0.x * down; r[1] = v0.y * down; r[2] = v1.x * down; r[3] = v1.y * down;
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t) r[t] = (c0 + t < used) ? s[c0 + t] * down : 0.0;
    }
    signed char* tile = base + (long long)kb * F8_S * tile_bytes;
    int q[F8_S][4];
#pragma unroll
    for (int p = 0; p < F8_S; ++p) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double tt = fma(r[t], (double)(1 << F8_BITS), kMagic);      // r 2^8 is exact: one rounding, to an integer
        q[p][t] = __double2loint(tt);            // round to nearest: [-128, 128]
        r[t] = fma(r[t], (double)(1 << F8_BITS), -(tt - kMagic));        // exact remainder
      }
    }
    // a digit of +128 does not fit int8: write it as -128 and carry one into the next more significant digit
#pragma unroll
    for (int p = F8_S - 1; p >= 1; --p) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int c = (q[p][t] == 128) ? 1 : 0;
        q[p][t] -= c << 8;
        q[p - 1][t] += c;
      }
    }
#pragma unroll
    for (int p = 0; p < F8_S; ++p) {
      const unsigned word = ((unsigned)q[p][0] & 0xffu) | (((unsigned)q[p][1] & 0xffu) << 8) | (((unsigned)q[p][2] & 0xffu) << 16) |
                            (((unsigned)q[p][3] & 0xffu) << 24);
      *reinterpret_cast<unsigned*>(tile + p * tile_bytes) = word;
    }
  }
}

// Kernel rows straight into slices for single-factor monomial series kernels (sphere S^d, SO(3)): one CTA per
// (128-candidate row block, 128-column k-block) evaluates K*[i][j] = p(clamp(scale <x_i, t_j>(^2) + shift)) from shared
// memory panels and writes the block's seven 16 KB slice tiles - the float64 kernel rows never exist in HBM.  All rows
// share one exponent (from sup |p| on [lo, hi]; a row far from every training point loses the bits its own maximum would
// have gained, which the parity bar - relative to the largest variance - does not see).  Per (row, four
// columns): W dot-product FMAs per column from a broadcast candidate value and two 16-byte training loads, Horner, then
// the same digit extraction + carry pass as f8_slice_kernel.  The mean's k-block partial sums go to meanpart[kb][row]
// (f8_finish_kernel adds them in a fixed order).
constexpr int F8G_MAXW = 16, F8G_MAXT = 32, F8G_THREADS = 256;

struct F8GramParams {
  const double* xc; long long m, ldc;        // candidates
  const double* xt; long long n, ldt;        // training points
  const double* coef; const double* alpha;
  int W, T, square;
  double scale, shift, lo, hi;
  signed char* A; int* ea; double* meanpart;
  long long m_pad;
  F8Geom g;
};

// WT / TT > 0: width and term count at compile time (unrolled dot products and Horner); 0: run-time loops.
template <int WT, int TT>
__global__ void __launch_bounds__(F8G_THREADS) f8_gram_slice_kernel(const F8GramParams p) {
  __shared__ __align__(16) double xts[F8G_MAXW][F8_BK];      // training panel, transposed: [k][column]
  __shared__ __align__(16) double xcs[F8_BM][F8G_MAXW];      // candidate rows
  __shared__ double cfs[F8G_MAXT];
  __shared__ double als[F8_BK];
  const int W = WT > 0 ? WT : p.W, T = TT > 0 ? TT : p.T;
  const int kb = blockIdx.x;
  const long long rb = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int e = threadIdx.x; e < W * F8_BK; e += F8G_THREADS) {
    const int c = e / W, k = e - c * W;
    const long long j = (long long)kb * F8_BK + c;
    xts[k][c] = (j < p.n) ? p.xt[j * p.ldt + k] : 0.0;
  }
  for (int e = threadIdx.x; e < W * F8_BM; e += F8G_THREADS) {
    const int r = e / W, k = e - r * W;
    const long long i = rb * F8_BM + r;
    xcs[r][k] = (i < p.m) ? p.xc[i * p.ldc + k] : 0.0;
  }
  for (int e = threadIdx.x; e < T; e += F8G_THREADS) cfs[e] = p.coef[e];
  for (int e = threadIdx.x; e < F8_BK; e += F8G_THREADS) {
    const long long j = (long long)kb * F8_BK + e;
    als[e] = (j < p.n) ? p.alpha[j] : 0.0;
  }
  __syncthreads();
  double cf[TT > 0 ? TT : 1];
  // One exponent for every row from sup |p| on [lo, hi]: thread t evaluates p at the t-th of 256 Chebyshev nodes of the
  // interval, and for a polynomial of degree < 32 the supremum is below max_t |p(u_t)| / cos(pi 31 / 512) (Ehlich -
  // Zeller), so 1.25 x the sampled maximum is a bound.  (sum |coef_k| would be up to 100 x too large at small
  // lengthscales - the monomial coefficients cancel - and cost 7 of the 56 bits.)
  __shared__ double wmax[F8G_THREADS / 32];
  {
    const double mid = 0.5 * (p.hi + p.lo), half = 0.5 * (p.hi - p.lo);
    const double u = fma(half, cospi((threadIdx.x + 0.5) / (double)F8G_THREADS), mid);
    double sv = cfs[T - 1];
    for (int k = T - 2; k >= 0; --k) sv = fma(sv, u, cfs[k]);
    sv = fabs(sv);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sv = fmax(sv, __shfl_xor_sync(0xffffffffu, sv, o));
    if (lane == 0) wmax[warp] = sv;
  }
  __syncthreads();
  double bound = 0.0;
#pragma unroll
  for (int w8 = 0; w8 < F8G_THREADS / 32; ++w8) bound = fmax(bound, wmax[w8]);
  bound *= 1.25;
  if (TT > 0) {
#pragma unroll
    for (int k = 0; k < TT; ++k) cf[k] = cfs[k];
  }
  int e2 = 0;
  if (bound > 0.0) {
    const double f = frexp(bound, &e2);
    e2 += (f > 0.984375) ? 2 : 1;
  }
  const double down = ldexp(1.0, -e2);
  constexpr double kMagic = 6755399441055744.0;      // 1.5 * 2^52
  const long long tile_bytes = (long long)F8_BM * F8_BK;
  signed char* tile0 = p.A + p.g.a_tile(rb, kb, 0) + ((lane & 3) << 2);
  const int c0 = 4 * lane;
  double al4[4];
  bool colok[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    al4[t] = als[c0 + t];
    colok[t] = (long long)kb * F8_BK + c0 + t < p.n;
  }
  constexpr int RPT = 2;                               // rows per iteration: eight independent chains per lane
  for (int r0 = RPT * warp; r0 < F8_BM; r0 += RPT * (F8G_THREADS / 32)) {
    double d[RPT][4];
#pragma unroll
    for (int h = 0; h < RPT; ++h)
#pragma unroll
      for (int t = 0; t < 4; ++t) d[h][t] = 0.0;
#pragma unroll
    for (int k = 0; k < (WT > 0 ? WT : F8G_MAXW); ++k) {
      if (k < W) {
        const double2 t01 = *reinterpret_cast<const double2*>(&xts[k][c0]);
        const double2 t23 = *reinterpret_cast<const double2*>(&xts[k][c0 + 2]);
#pragma unroll
        for (int h = 0; h < RPT; ++h) {
          const double v = xcs[r0 + h][k];
          d[h][0] = fma(v, t01.x, d[h][0]);
          d[h][1] = fma(v, t01.y, d[h][1]);
          d[h][2] = fma(v, t23.x, d[h][2]);
          d[h][3] = fma(v, t23.y, d[h][3]);
        }
      }
    }
    double kv[RPT][4];
#pragma unroll
    for (int h = 0; h < RPT; ++h) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        double u = fma(p.square ? d[h][t] * d[h][t] : d[h][t], p.scale, p.shift);
        u = (u < p.lo) ? p.lo : ((u > p.hi) ? p.hi : u);
        kv[h][t] = u;
      }
    }
    if (TT > 0) {
      double sacc[RPT][4];
#pragma unroll
      for (int h = 0; h < RPT; ++h)
#pragma unroll
        for (int t = 0; t < 4; ++t) sacc[h][t] = cf[TT - 1];
#pragma unroll
      for (int k = TT - 2; k >= 0; --k)
#pragma unroll
        for (int h = 0; h < RPT; ++h)
#pragma unroll
          for (int t = 0; t < 4; ++t) sacc[h][t] = fma(sacc[h][t], kv[h][t], cf[k]);
#pragma unroll
      for (int h = 0; h < RPT; ++h)
#pragma unroll
        for (int t = 0; t < 4; ++t) kv[h][t] = sacc[h][t];
    } else {
#pragma unroll
      for (int h = 0; h < RPT; ++h)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          double sv = cfs[T - 1];
          for (int k = T - 2; k >= 0; --k) sv = fma(sv, kv[h][t], cfs[k]);
          kv[h][t] = sv;
        }
    }
#pragma unroll
    for (int h = 0; h < RPT; ++h) {
      const int r = r0 + h;
      const long long i = rb * F8_BM + r;
      const bool rowok = i < p.m;
      double mu = 0.0, rr[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double kval = (rowok && colok[t]) ? kv[h][t] : 0.0;
        mu = fma(kval, al4[t], mu);
        rr[t] = kval * down;
      }
      mu = warp_sum(mu);
      if (lane == 0 && rowok) p.meanpart[(long long)kb * p.m_pad + i] = mu;
      if (kb == 0 && lane == 0) p.ea[i] = e2;
      // Digits on the integer pipe (the FP64 pipe is this kernel's bound): r 2^56 as two exact integers - the leading 32
      // bits (|r| 2^32 <= 126 2^24 fits int32) and the following 24 of the exact remainder, 4 FP64 instructions instead of
      // three per digit - then balanced base-256 digits from the least significant end: the digit is the low byte read as
      // int8, what is left is (v + 128) >> 8, and the leftover of the lower word carries into the upper one.  Only the low
      // byte of q[s][t] is used below.
      int q[F8_S][4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double t1 = fma(rr[t], 4294967296.0, kMagic);                         // round(r 2^32)
        int hi4 = __double2loint(t1);
        const double rem = fma(rr[t], 4294967296.0, -(t1 - kMagic));                // exact, |rem| <= 1/2
        int lo3 = __double2loint(fma(rem, 16777216.0, kMagic));                     // round(rem 2^24), |lo3| <= 2^23
        q[6][t] = lo3;
        lo3 = (lo3 + 128) >> 8;
        q[5][t] = lo3;
        lo3 = (lo3 + 128) >> 8;
        q[4][t] = lo3;
        hi4 += (lo3 + 128) >> 8;
        q[3][t] = hi4;
        hi4 = (hi4 + 128) >> 8;
        q[2][t] = hi4;
        hi4 = (hi4 + 128) >> 8;
        q[1][t] = hi4;
        q[0][t] = (hi4 + 128) >> 8;
      }
      signed char* dst = tile0 + f8_swizzled(r, lane >> 2);
#pragma unroll
      for (int s7 = 0; s7 < F8_S; ++s7) {
        const unsigned word = ((unsigned)q[s7][0] & 0xffu) | (((unsigned)q[s7][1] & 0xffu) << 8) | (((unsigned)q[s7][2] & 0xffu) << 16) |
                              (((unsigned)q[s7][3] & 0xffu) << 24);
        *reinterpret_cast<unsigned*>(dst + s7 * tile_bytes) = word;
      }
    }
  }
}

struct F8Params {
  const signed char* A;      // tiled K* slices  [row block][k-block][slice][128 x 128 swizzled]
  const signed char* B;      // tiled L^-1 slices [column tile][k-block][slice][64 x 128 swizzled]
  const int* ea;             // exponent per candidate
  const double* sb;          // 2^eb per column
  double* partial;           // [column tile][m_pad] sums of squares
  int* planes;               // debug: [row block][column tile][7][128][64] raw accumulators, or nullptr
  long long m, m_pad;
  long long n_rb;            // row blocks
  long long items;           // work items incl. the padding of the last group
  int group;                 // row blocks per L2 group
  int debug;                 // timing experiments (MGB_I8F_DEBUG): 1 = no copies, 2 = no MMAs (results are then garbage)
  F8Geom g;
};

struct F8Item {
  long long rb;
  int ct;
  bool live;
};
__device__ __forceinline__ F8Item f8_item(const F8Params& P, long long it) {
  const long long per_group = (long long)P.group * P.g.CT;
  const long long grp = it / per_group;
  const int r = (int)(it % per_group);
  F8Item w;
  w.ct = P.g.CT - 1 - r / P.group;               // longest k loops first
  w.rb = grp * P.group + r % P.group;
  w.live = w.rb < P.n_rb;
  return w;
}

__global__ void __launch_bounds__(F8_THREADS, 1) f8_posterior_kernel(const F8Params P) {
  extern __shared__ unsigned char f8_smem_raw[];
  const uint32_t raw = smem_u32(f8_smem_raw);
  unsigned char* smem = f8_smem_raw + (((raw + 1023u) & ~1023u) - raw);      // the swizzle pattern is on absolute address bits
  unsigned char* Bs = smem;
  unsigned char* As = smem + (size_t)F8_B_STAGES * F8_S * F8_B_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(As + (size_t)F8_A_STAGES * F8_A_TILE);
  uint64_t* a_full = bars;                          // [F8_A_STAGES]
  uint64_t* a_empty = a_full + F8_A_STAGES;         // [F8_A_STAGES]
  uint64_t* b_full = a_empty + F8_A_STAGES;         // [F8_B_STAGES]
  uint64_t* b_empty = b_full + F8_B_STAGES;         // [F8_B_STAGES]
  uint64_t* tmem_full = b_empty + F8_B_STAGES;
  uint64_t* tmem_empty = tmem_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;      // warp-uniform for the compiler
  if (threadIdx.x == 0) {
    for (int s = 0; s < F8_A_STAGES; ++s) {
      mbar_init(a_full + s, 1);
      mbar_init(a_empty + s, 1);
    }
    for (int s = 0; s < F8_B_STAGES; ++s) {
      mbar_init(b_full + s, 1);
      mbar_init(b_empty + s, 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(F8_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  f8_fence_before();
  __syncthreads();
  f8_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== producer =====
    if (lane == 0) {
      uint32_t a_it = 0, b_it = 0;
      for (long long it = blockIdx.x; it < P.items; it += gridDim.x) {
        const F8Item w = f8_item(P, it);
        if (!w.live) continue;
        const int nkb = F8Geom::nkb(w.ct);
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t bs = b_it % F8_B_STAGES, bph = (b_it / F8_B_STAGES) & 1u;
          f8_mbar_spin(b_empty + bs, bph ^ 1u);
          if (P.debug & 1) {
            f8_mbar_arrive(b_full + bs);
          } else {
            mbar_expect_tx(b_full + bs, F8_S * F8_B_TILE);
            bulk_g2s(Bs + (size_t)bs * F8_S * F8_B_TILE, P.B + P.g.b_tile(w.ct, kb, 0), F8_S * F8_B_TILE, b_full + bs);
          }
          ++b_it;
          for (int p = 0; p < F8_S; ++p) {
            const uint32_t as = a_it % F8_A_STAGES, aph = (a_it / F8_A_STAGES) & 1u;
            f8_mbar_spin(a_empty + as, aph ^ 1u);
            if (P.debug & 1) {
              f8_mbar_arrive(a_full + as);
            } else {
              mbar_expect_tx(a_full + as, F8_A_TILE);
              bulk_g2s(As + (size_t)as * F8_A_TILE, P.A + P.g.a_tile(w.rb, kb, p), F8_A_TILE, a_full + as);
            }
            ++a_it;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    uint32_t a_it = 0, b_it = 0, n_item = 0;
    for (long long it = blockIdx.x; it < P.items; it += gridDim.x) {
      const F8Item w = f8_item(P, it);
      if (!w.live) continue;
      const int nkb = F8Geom::nkb(w.ct);
      f8_mbar_spin(tmem_empty, (n_item & 1u) ^ 1u);       // the epilogue has drained the previous tile
      f8_fence_after();
      for (int kb = 0; kb < nkb; ++kb) {
        const uint32_t bs = b_it % F8_B_STAGES, bph = (b_it / F8_B_STAGES) & 1u;
        f8_mbar_spin(b_full + bs, bph);
        const uint32_t b_addr = smem_u32(Bs + (size_t)bs * F8_S * F8_B_TILE);
        // last k-block of an even column tile: columns [64 ct, 64 ct + 64) only meet k < 64 ct + 64, its first half
        const int ks_end = (kb == nkb - 1 && (w.ct & 1) == 0) ? F8_BK / 64 : F8_BK / 32;
        const uint32_t b_lo = f8_desc_lo(b_addr);
        // Planes p .. 6 are adjacent in tensor memory and the B tiles of a k-block are adjacent in shared memory (same
        // swizzle pattern, 64 rows each): the 7 - p slice products of A_p are ONE MMA with N = 64 (7 - p) against the
        // stacked tiles B_0 .. B_{6-p}, split in two where N would exceed 256 - 10 MMAs per 32-byte k step instead of
        // 28.  The whole warp runs this code (one elected lane issues): descriptors stay in uniform registers.
#pragma unroll
        for (int p = 0; p < F8_S; ++p) {
          const uint32_t as = a_it % F8_A_STAGES, aph = (a_it / F8_A_STAGES) & 1u;
          f8_mbar_spin(a_full + as, aph);
          f8_fence_after();
          const uint32_t a_lo = f8_desc_lo(smem_u32(As + (size_t)as * F8_A_TILE));
          const int n_all = F8_BN * (F8_S - p);
          const int n0 = n_all < 256 ? n_all : 256, n1 = n_all - n0;
          const uint32_t d0 = tmem_base + (uint32_t)(p * F8_BN);
          if (!(P.debug & 2)) {
#pragma unroll
            for (int ks = 0; ks < F8_BK / 32; ++ks) {
              if (ks < ks_end) f8_mma(d0, f8_desc_from_lo(a_lo + 2 * ks), f8_desc_from_lo(b_lo + 2 * ks), f8_idesc(n0), (kb | p | ks) != 0 ? 1u : 0u);
            }
            if (n1 > 0) {
#pragma unroll
              for (int ks = 0; ks < F8_BK / 32; ++ks) {
                if (ks < ks_end)
                  f8_mma(d0 + 256u, f8_desc_from_lo(a_lo + 2 * ks), f8_desc_from_lo(b_lo + (4 * F8_B_TILE >> 4) + 2 * ks), f8_idesc(n1),
                         (kb | p | ks) != 0 ? 1u : 0u);
              }
            }
          }
          f8_commit(a_empty + as);                        // slot free once these MMAs have read it
          ++a_it;
        }
        f8_commit(b_empty + bs);
        ++b_it;
      }
      f8_commit(tmem_full);                               // all seven planes of the tile are final
      ++n_item;
    }
  } else {
    // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    double wgt[F8_S];
#pragma unroll
    for (int d = 0; d < F8_S; ++d) wgt[d] = exp2((double)(-F8_BITS * (d + 2)));
    uint32_t n_item = 0;
    for (long long it = blockIdx.x; it < P.items; it += gridDim.x) {
      const F8Item w = f8_item(P, it);
      if (!w.live) continue;
      const long long row = w.rb * F8_BM + quarter * 32 + lane;
      f8_mbar_spin(tmem_full, n_item & 1u);
      f8_fence_after();
      double ss = 0.0;
#pragma unroll 1
      for (int c = 0; c < F8_BN / 16; ++c) {
        double acc[16];
#pragma unroll
        for (int t = 0; t < 16; ++t) acc[t] = 0.0;
#pragma unroll
        for (int d = F8_S - 1; d >= 0; --d) {             // smallest weights first
          int v[16];
          f8_tmem_ld16(lane_addr + (uint32_t)(d * F8_BN + c * 16), v);
          f8_tmem_ld_wait();
          if (P.planes != nullptr) {
            int* dbg = P.planes + (((w.rb * P.g.CT + w.ct) * F8_S + d) * F8_BM + (quarter * 32 + lane)) * F8_BN + c * 16;
#pragma unroll
            for (int t = 0; t < 16; ++t) dbg[t] = v[t];
          }
#pragma unroll
          for (int t = 0; t < 16; ++t) acc[t] = fma((double)v[t], wgt[d], acc[t]);
        }
        const double* sb = P.sb + (long long)w.ct * F8_BN + c * 16;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const double x = acc[t] * __ldg(sb + t);
          ss = fma(x, x, ss);
        }
      }
      f8_fence_before();
      __syncwarp();
      if (lane == 0) f8_mbar_arrive(tmem_empty);
      if (row < P.m) P.partial[(long long)w.ct * P.m_pad + row] = ldexp(ss, 2 * P.ea[row]);
      ++n_item;
    }
  }
  f8_fence_before();
  __syncthreads();
  if (warp == 1) {
    f8_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(F8_TMEM_COLS) : "memory");
  }
}

// var_i = kss - sum over column tiles (fixed order: reproducible)
__global__ void __launch_bounds__(256) f8_finish_kernel(const double* partial, long long m, long long m_pad, int CT, double kss, double* var,
                                                        const double* meanpart, int KB, double mean_const, double* mean) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double s = 0.0;
  for (int ct = 0; ct < CT; ++ct) s += partial[(long long)ct * m_pad + i];
  var[i] = kss - s;
  if (meanpart != nullptr) {       // f8_gram_slice_kernel's k-block partial sums of K* alpha
    double mu = 0.0;
    for (int kb = 0; kb < KB; ++kb) mu += meanpart[(long long)kb * m_pad + i];
    mean[i] = mean_const + mu;
  }
}

}  // namespace mgb

using namespace mgb;

extern "C" long long mgb_posterior_i8f_pack_bytes(long long n) {
  if (n < 1) return 0;
  const F8Geom g = f8_geom(n);
  return 8 * g.np + g.b_bytes();
}

extern "C" int mgb_posterior_i8f_pack(const double* rinv, long long n, long long ld, void* packed, void* stream) {
  if (!rinv || !packed || n < 1 || ld < n) return fail(MGB_ERR_ARG, "posterior_i8f_pack: bad argument");
  if (n > F8_MAX_N) return fail(MGB_ERR_ARG, "posterior_i8f_pack: n too large for exact int32 accumulation (16384)");
  const F8Geom g = f8_geom(n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* sb = static_cast<double*>(packed);
  signed char* B = static_cast<signed char*>(packed) + 8 * g.np;
  MGB_TRY(check_cuda(cudaMemsetAsync(packed, 0, (size_t)mgb_posterior_i8f_pack_bytes(n), s), "cudaMemsetAsync"));
  f8_slice_kernel<<<(unsigned)((g.np + 7) / 8), 256, 0, s>>>(rinv, ld, 0, n, g.np, (int)n, g, 1, B, nullptr, sb, nullptr, 0.0, nullptr);
  return after_launch("f8_slice_kernel (pack)");
}

// Row blocks whose A slices are kept L2-resident while the column tiles sweep over them (MGB_I8F_GROUP overrides).
static int f8_group() {
  static const int g = [] {
    const char* e = std::getenv("MGB_I8F_GROUP");
    const int v = e ? std::atoi(e) : 0;
    return v >= 1 && v <= 4096 ? v : 8;       // 8: 37 MB of A slices at n = 5000; 16 costs 3 % under the power cap (more HBM reads), 32 8 %
  }();
  return g;
}

static long long f8_ws_ea(long long m) { return f8_round_up(4 * f8_round_up(m, F8_BM), 256); }

extern "C" long long mgb_posterior_i8f_workspace_bytes(long long m, long long n) {
  if (m < 1 || n < 1) return 0;
  const F8Geom g = f8_geom(n);
  const long long m_pad = f8_round_up(m, F8_BM);
  return f8_ws_ea(m) + f8_round_up(g.a_bytes(m), 256) + f8_round_up((long long)g.CT * m_pad * 8, 256);
}

struct F8Workspace {
  int* ea;
  signed char* A;
  double* partial;
  double* meanpart;   // series entry point, fused Gram + slicing: [k-block][m_pad] partial sums of K* alpha
  double* krows;      // series entry point, two-kernel route: one sub-block of float64 kernel rows
};

static F8Workspace f8_carve(void* workspace, long long m, const F8Geom& g) {
  F8Workspace w;
  char* p = static_cast<char*>(workspace);
  w.ea = reinterpret_cast<int*>(p);
  p += f8_ws_ea(m);
  w.A = reinterpret_cast<signed char*>(p);
  p += f8_round_up(g.a_bytes(m), 256);
  w.partial = reinterpret_cast<double*>(p);
  p += f8_round_up((long long)g.CT * f8_round_up(m, F8_BM) * 8, 256);
  w.meanpart = reinterpret_cast<double*>(p);
  p += f8_round_up((long long)g.KB * f8_round_up(m, F8_BM) * 8, 256);
  w.krows = reinterpret_cast<double*>(p);
  return w;
}

// rows [row0, row0 + rows) of K* -> slices; the last sub-block of a chunk also writes the zero rows up to the 128 boundary
static int f8_slice_rows(const double* kstar, long long ldk, long long row0, long long rows, bool last, long long n, const F8Geom& g,
                         const F8Workspace& w, const double* alpha, double mean_const, double* mean, cudaStream_t s) {
  const long long rows_pad = last ? f8_round_up(row0 + rows, F8_BM) - row0 : rows;
  f8_slice_kernel<<<(unsigned)((rows_pad + 7) / 8), 256, 0, s>>>(kstar, ldk, row0, rows, rows_pad, (int)n, g, 0, w.A, w.ea, nullptr, alpha,
                                                                   mean_const, mean);
  return after_launch("f8_slice_kernel");
}

// slices of m candidates (already in the workspace) x packed factor -> var
static int f8_contract(long long m, long long n, const F8Geom& g, const F8Workspace& w, const void* packed, double kss, double* var,
                       int* planes, cudaStream_t s, const double* meanpart = nullptr, double mean_const = 0.0, double* mean = nullptr) {
  const long long m_pad = f8_round_up(m, F8_BM);
  F8Params P;
  P.A = w.A;
  P.B = static_cast<const signed char*>(packed) + 8 * g.np;
  P.ea = w.ea;
  P.sb = static_cast<const double*>(packed);
  P.partial = w.partial;
  P.planes = planes;
  P.m = m;
  P.m_pad = m_pad;
  P.n_rb = m_pad / F8_BM;
  P.group = f8_group();
  {
    const char* e = std::getenv("MGB_I8F_DEBUG");
    P.debug = e ? std::atoi(e) : 0;
  }
  P.items = f8_round_up(P.n_rb, P.group) * g.CT;
  P.g = g;
  MGB_TRY(check_cuda(cudaFuncSetAttribute(f8_posterior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F8_SMEM), "cudaFuncSetAttribute"));
  const long long live = P.n_rb * g.CT;
  const unsigned grid = (unsigned)std::min<long long>(device_sm_count(), live);
  f8_posterior_kernel<<<grid, F8_THREADS, F8_SMEM, s>>>(P);
  MGB_TRY(after_launch("f8_posterior_kernel"));
  f8_finish_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(w.partial, m, m_pad, g.CT, kss, var, meanpart, g.KB, mean_const, mean);
  return after_launch("f8_finish_kernel");
}

static int f8_reduce(const double* kstar, long long m, long long n, long long ldk, const void* packed, const double* alpha,
                     double kss, double mean_const, double* mean, double* var, void* workspace, long long workspace_bytes,
                     int* planes, void* stream) {
  if (m == 0) return MGB_OK;
  if (!kstar || !packed || !alpha || !mean || !var || !workspace || m < 0 || n < 1 || ldk < n)
    return fail(MGB_ERR_ARG, "posterior_i8f_reduce: bad argument");
  if (m > 0x7fffffffLL / 8 || n > F8_MAX_N) return fail(MGB_ERR_ARG, "posterior_i8f_reduce: block too large (n <= 16384)");
  if (workspace_bytes < mgb_posterior_i8f_workspace_bytes(m, n)) return fail(MGB_ERR_ARG, "posterior_i8f_reduce: workspace too small");
  const F8Geom g = f8_geom(n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const F8Workspace w = f8_carve(workspace, m, g);
  MGB_TRY(f8_slice_rows(kstar, ldk, 0, m, true, n, g, w, alpha, mean_const, mean, s));
  return f8_contract(m, n, g, w, packed, kss, var, planes, s);
}

extern "C" int mgb_posterior_i8f_reduce(const double* kstar, long long m, long long n, long long ldk, const void* packed,
                                        const double* alpha, double kss, double mean_const, double* mean, double* var,
                                        void* workspace, long long workspace_bytes, void* stream) {
  return f8_reduce(kstar, m, n, ldk, packed, alpha, kss, mean_const, mean, var, workspace, workspace_bytes, nullptr, stream);
}

extern "C" long long mgb_posterior_i8f_planes_bytes(long long m, long long n) {
  if (m < 1 || n < 1) return 0;
  const F8Geom g = f8_geom(n);
  return f8_round_up(m, F8_BM) / F8_BM * g.CT * (long long)F8_S * F8_BM * F8_BN * 4;
}

extern "C" int mgb_posterior_i8f_reduce_debug(const double* kstar, long long m, long long n, long long ldk, const void* packed,
                                              const double* alpha, double kss, double mean_const, double* mean, double* var,
                                              void* workspace, long long workspace_bytes, int* planes, void* stream) {
  if (!planes) return fail(MGB_ERR_ARG, "posterior_i8f_reduce_debug: planes is null");
  return f8_reduce(kstar, m, n, ldk, packed, alpha, kss, mean_const, mean, var, workspace, workspace_bytes, planes, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// Series kernels end to end (mgb.h): candidates -> kernel rows -> slices -> contraction, for one chunk of m candidates.
// MGB_I8F_SUB = rows per Gram + slicing sub-block (default: the whole chunk).  Sub-blocks small enough to keep the float64
// kernel rows in L2 between the Gram kernel's writes and the slicing kernel's reads were measured and lost: 1280 rows
// 3.29e6, 2560 rows 3.50e6, whole chunk 3.68e6 candidates/s (both kernels are sized for large launches).
// ------------------------------------------------------------------------------------------------------------------
static long long f8_sub_rows() {
  static const long long v = [] {
    const char* e = std::getenv("MGB_I8F_SUB");
    const long long r = e ? std::atoll(e) : 0;
    return r >= 128 ? r / 128 * 128 : (1LL << 40);
  }();
  return v;
}

extern "C" long long mgb_series_posterior_i8f_workspace_bytes(long long m, long long n) {
  if (m < 1 || n < 1) return 0;
  const long long sub = std::min<long long>(f8_sub_rows(), f8_round_up(m, F8_BM));
  const F8Geom g = f8_geom(n);
  return mgb_posterior_i8f_workspace_bytes(m, n) + f8_round_up((long long)g.KB * f8_round_up(m, F8_BM) * 8, 256) + sub * n * 8;
}

// f8_gram_slice_kernel serves single-factor monomial series of moderate width (MGB_I8F_GRAM=0: two-kernel route)
static bool f8_gram_fusable(const mgb_series_desc* d) {
  static const bool enabled = [] {
    const char* e = std::getenv("MGB_I8F_GRAM");
    return !(e && e[0] == '0');
  }();
  return enabled && d->n_factors == 1 && d->basis == MGB_BASIS_MONOMIAL && d->factor_width <= F8G_MAXW && d->n_terms <= F8G_MAXT;
}

extern "C" int mgb_series_posterior_i8f(const mgb_series_desc* desc, const double* xc, long long m, long long ldc, const double* xtrain,
                                        long long n, long long ldt, const double* coef, const void* packed, const double* alpha,
                                        double kss, double mean_const, double* mean, double* var, void* workspace,
                                        long long workspace_bytes, void* stream) {
  if (m == 0) return MGB_OK;
  if (!desc || !xc || !xtrain || !coef || !packed || !alpha || !mean || !var || !workspace || m < 0 || n < 1)
    return fail(MGB_ERR_ARG, "series_posterior_i8f: bad argument");
  if (m > 0x7fffffffLL / 8 || n > F8_MAX_N) return fail(MGB_ERR_ARG, "series_posterior_i8f: block too large (n <= 16384)");
  if (workspace_bytes < mgb_series_posterior_i8f_workspace_bytes(m, n)) return fail(MGB_ERR_ARG, "series_posterior_i8f: workspace too small");
  const F8Geom g = f8_geom(n);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const F8Workspace w = f8_carve(workspace, m, g);
  if (f8_gram_fusable(desc)) {
    const long long m_pad = f8_round_up(m, F8_BM);
    F8GramParams q;
    q.xc = xc; q.m = m; q.ldc = ldc;
    q.xt = xtrain; q.n = n; q.ldt = ldt;
    q.coef = coef; q.alpha = alpha;
    q.W = desc->factor_width; q.T = desc->n_terms; q.square = desc->pair_square;
    q.scale = desc->scale; q.shift = desc->shift; q.lo = desc->lo; q.hi = desc->hi;
    q.A = w.A; q.ea = w.ea; q.meanpart = w.meanpart;
    q.m_pad = m_pad;
    q.g = g;
    const dim3 grid((unsigned)g.KB, (unsigned)(m_pad / F8_BM));
    if (q.W == 11 && q.T == 10) f8_gram_slice_kernel<11, 10><<<grid, F8G_THREADS, 0, s>>>(q);        // S^10
    else if (q.W == 4 && q.T == 10) f8_gram_slice_kernel<4, 10><<<grid, F8G_THREADS, 0, s>>>(q);     // S^3, SO(3) quaternions
    else if (q.W == 3 && q.T == 10) f8_gram_slice_kernel<3, 10><<<grid, F8G_THREADS, 0, s>>>(q);     // S^2
    else f8_gram_slice_kernel<0, 0><<<grid, F8G_THREADS, 0, s>>>(q);
    MGB_TRY(after_launch("f8_gram_slice_kernel"));
    return f8_contract(m, n, g, w, packed, kss, var, nullptr, s, w.meanpart, mean_const, mean);
  }
  const long long sub = std::min<long long>(f8_sub_rows(), f8_round_up(m, F8_BM));
  for (long long r0 = 0; r0 < m; r0 += sub) {
    const long long rows = std::min<long long>(sub, m - r0);
    MGB_TRY(mgb_series_gram_fwd(desc, xc + r0 * ldc, rows, ldc, xtrain, n, ldt, coef, w.krows, n, stream));
    MGB_TRY(f8_slice_rows(w.krows, n, r0, rows, r0 + rows >= m, n, g, w, alpha, mean_const, mean, s));
  }
  return f8_contract(m, n, g, w, packed, kss, var, nullptr, s);
}
