// Series Gram matrix with the inner products on the INT8 tensor cores (tcgen05.mma kind::i8, accumulators in tensor
// memory) and the polynomial on the FP64 pipe: K_ij = p(clamp(scale <x1_i, x2_j>(^2) + shift)) for single-factor
// monomial series kernels - sphere S^d, SO(3) in trace or quaternion form (reference: kernel_utils/kernels_sphere.py:
// 102-139,188-224,343-384 and kernels_so.py:305-362 with dim = 3; same contract as mgb_series_gram_fwd).
//
// Why: DFMA and DMMA share one FP64 data path on B200, and the float64 kernel spends W (padded to a multiple of 4) of
// its W + T FMA slots per entry on the inner product: 12 of 21 for S^10 and for SO(3) in trace form.  Under sustained
// load the board sits at its power cap and that kernel reaches 64 % (S^10) of the HBM-write roof.  Here the inner
// product leaves the FP64 pipe: 3 + (T - 1) FP64 instructions per entry remain (two to recombine, one scale / shift,
// Horner), the rest is integer work and tensor-core time that is an order of magnitude below its roof.
// MEASURED (1 M x 2 k, one B200): exact - planes equal to an integer matmul, K within 2 - 5e-16 of a long-double evaluation -
// but 3.6e11 entries/s on S^2, S^10 and SO(3) alike against 5.3 - 6.5e11 for the float64 kernels: recombining seven
// overlapping 21-bit planes takes ~10 integer instructions per entry, which is more issue time than the FP64 slots it
// saves.  Opt-in (MGB_GRAM_PATH=int8); the float64 kernels stay the default.
//
// Error-free slicing (Ozaki scheme, as in int8_fused.cu): every row is scaled by a power of two 2^-e (|x| 2^-e <=
// 126/256) and cut into S = 7 signed 8-bit digits (round to nearest + carry pass), x = 2^e sum_p d_p 2^(-8 (p + 1)) up
// to 2^(e - 57), so
//      <x, y> = 2^(ea + eb) sum_{s < 7} 2^(-8 (s + 2)) P_s,     P_s = sum_{p + q = s} <d_p(x), d_q(y)>   (int32, exact)
// with the products of p + q >= 7 dropped (<= 6 W 2^14 2^-72 2^(ea + eb) ~ 3e-16 2^(ea + eb) for W = 16; measured
// against the float64 kernel: a few 1e-16).  The seven planes come out of ONE small GEMM per tile: an operand row of x is its digit vectors
// side by side, [d_0 | d_1 | ... | d_6 | 0] (eight 16-byte chunks, W <= 16 int8 each: one 128-byte swizzle row), and the
// B operand row of (plane s, column j) is [d_s(y_j) | d_{s-1}(y_j) | ... | d_0(y_j) | 0 ...]: the K = 128 contraction
// of the two is exactly P_s.  Chunk pair a = 0 .. 3 (one K = 32 MMA step) only meets planes s >= 2 a, which are adjacent
// in tensor memory and in the stacked B tile: 4 MMAs per 128 x 32 tile, N = 224, 160, 96, 32.
//
// CTA (640 threads, one per SM, persistent over 128-row panels):
//   warp 0      B producer: one cp.async.bulk per stacked B tile (28 KB, pre-built by g8_slice_b_kernel) into a 4-deep ring
//   warp 1      MMA issuer (one elected lane), accumulators double-buffered in tensor memory (2 x 224 columns)
//   warp 2      A slicer: digits of the next row panel straight into the swizzled shared-memory operand image
//   warps 4-19  epilogue (32 rows x 8 columns of a tile each): tcgen05.ld (thread = row; the next tile's loads are issued
//               before this tile's Horner phase), integer recombination of the planes into two exact 64-bit sums,
//               magic-number conversion, two FMAs -> the correctly rounded inner product, scale / shift, clamp (integer
//               test, exact slow path), Horner, transpose through shared memory, 64-byte row segments as streaming stores.
#include "mgb_tcgen05.cuh"

#include <cstring>
#include <cmath>

namespace mgb {

constexpr int G8_S = 7;                       // digits per value
constexpr int G8_NT = 32;                     // columns per tile
constexpr int G8_BROWS = G8_S * G8_NT;        // rows of a stacked B tile
constexpr int G8_B_TILE = G8_BROWS * 128;     // 28 KB
constexpr int G8_A_TILE = F8_BM * 128;        // 16 KB
constexpr int G8_B_STAGES = 4;
constexpr int G8_A_BUFS = 2;
constexpr int G8_ACC_STRIDE = 256;            // tensor-memory columns between the two accumulator buffers
constexpr int G8_EPI_WARPS = 16;
constexpr int G8_CW = G8_NT * 4 / G8_EPI_WARPS;   // columns of a tile per epilogue warp (8: one tcgen05.ld.x8 per plane)
constexpr int G8_THREADS = 32 * (4 + G8_EPI_WARPS);
constexpr int G8_STG_ROW = 80;                // bytes per staging row: 8 doubles + 16 bytes (conflict-free 16-byte accesses)
constexpr int G8_STG_WARP = 32 * G8_STG_ROW;
constexpr int G8_MAXW = 16;
constexpr int G8_BAD_COLUMN = 0x7fffffff;     // exponent word of a training row with a non-finite entry
constexpr int G8_MAX_COLS = 4096;             // columns per launch (exponent table in shared memory)
constexpr size_t G8_SMEM = 1024 + (size_t)G8_B_STAGES * G8_B_TILE + (size_t)G8_A_BUFS * G8_A_TILE + (size_t)G8_EPI_WARPS * G8_STG_WARP +
                           (size_t)G8_MAX_COLS * 4 + (size_t)G8_A_BUFS * F8_BM * 8 + 256;

// a * b + c with a 32 x 32 -> 64-bit product (IMAD.WIDE)
__device__ __forceinline__ long long g8_mad_wide(int a, int b, long long c) {
  long long d;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
}

// mbarrier operations on precomputed shared-memory addresses (no generic -> shared conversion in the tile loop)
__device__ __forceinline__ bool g8_mbar_test_u(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void g8_mbar_arrive_u(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

__device__ __forceinline__ long long g8_pack(int hi, int lo) {
  long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// Waits of the three single-warp roles: back off between polls so that their spinning does not take issue slots from
// the epilogue warps of the same scheduler (bounded like f8_mbar_spin: a protocol error traps instead of hanging).
__device__ __forceinline__ void g8_mbar_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
  for (unsigned spins = 0; !f8_mbar_test(bar, parity); ++spins) {
    __nanosleep(ns);
    if (spins > (1u << 26)) __trap();
  }
}

// exponent e with |v| 2^-e <= 126/256 for |v| <= mx (0 for a zero row); non-finite rows are flagged by the caller
__device__ __forceinline__ int g8_exponent(double mx) {
  int e = 0;
  if (mx > 0.0) {
    const double f = frexp(mx, &e);
    e += (f > 0.984375) ? 2 : 1;
  }
  return e;
}

// seven digits of r (|r| <= 126/256): round to nearest by the 1.5 * 2^52 trick, exact remainders, then the carry pass
// that rewrites a digit of +128 as -128 and increments its more significant neighbour
__device__ __forceinline__ void g8_digits(double r, int (&q)[G8_S]) {
  constexpr double kMagic = 6755399441055744.0;
#pragma unroll
  for (int p = 0; p < G8_S; ++p) {
    const double tt = fma(r, 256.0, kMagic);
    q[p] = __double2loint(tt);
    r = fma(r, 256.0, -(tt - kMagic));
  }
#pragma unroll
  for (int p = G8_S - 1; p >= 1; --p) {
    const int c = (q[p] == 128) ? 1 : 0;
    q[p] -= c << 8;
    q[p - 1] += c;
  }
}

// Training side, once per call: thread j slices row j of x2 and writes its rows of the stacked B tile j / 32 (plane s,
// chunk c holds digit vector s - c, zeros where s < c) plus the column's exponent word for the epilogue's constants.
__global__ void __launch_bounds__(128) g8_slice_b_kernel(const double* __restrict__ x2, long long n, long long ld2, int W, long long n_pad,
                                                         unsigned char* __restrict__ img, int* __restrict__ expw) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_pad) return;
  double v[G8_MAXW];
  double mx = 0.0;
  bool finite = true;
#pragma unroll
  for (int k = 0; k < G8_MAXW; ++k) {
    v[k] = (j < n && k < W) ? x2[j * ld2 + k] : 0.0;
    mx = fmax(mx, fabs(v[k]));
    finite = finite && (fabs(v[k]) < 1e300);      // false for NaN too
  }
  int e = finite ? g8_exponent(mx) : 0;
  const double down = ldexp(1.0, -e);
  unsigned w[G8_S][4];
#pragma unroll
  for (int p = 0; p < G8_S; ++p)
#pragma unroll
    for (int g = 0; g < 4; ++g) w[p][g] = 0u;
#pragma unroll
  for (int k = 0; k < G8_MAXW; ++k) {
    int q[G8_S];
    g8_digits(finite ? v[k] * down : 0.0, q);
#pragma unroll
    for (int p = 0; p < G8_S; ++p) w[p][k >> 2] |= ((unsigned)q[p] & 0xffu) << (8 * (k & 3));
  }
  unsigned char* tile = img + (j / G8_NT) * (long long)G8_B_TILE;
  const int jj = (int)(j % G8_NT);
#pragma unroll
  for (int s = 0; s < G8_S; ++s) {
    const int row = s * G8_NT + jj;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (c <= s) val = make_uint4(w[s - c][0], w[s - c][1], w[s - c][2], w[s - c][3]);
      *reinterpret_cast<uint4*>(tile + f8_swizzled(row, c)) = val;
    }
  }
  // exponent word: added to the high word of the row factor in the epilogue (times 2^eb); a non-finite row is flagged
  // and its column overwritten with NaN by g8_poison_kernel
  expw[j] = finite ? e * 1048576 : G8_BAD_COLUMN;
}

// Columns whose training row holds a NaN / Inf: NaN down the column, as the float64 kernel's arithmetic would give.
__global__ void __launch_bounds__(256) g8_poison_kernel(const int* __restrict__ expw, long long n, long long m, double* __restrict__ K, long long ldk) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || expw[j] != G8_BAD_COLUMN) return;
  const double nan = __longlong_as_double(0x7ff8000000000000LL);
  for (long long i = blockIdx.y; i < m; i += gridDim.y) K[i * ldk + j] = nan;
}

struct G8Params {
  const double* x1; long long m, ld1;
  const unsigned char* bimg;       // stacked B tiles of this launch's columns
  const int* expw;                 // exponent words of this launch's columns (padded to a multiple of 32)
  const double* coef;
  double* K; long long ldk;        // K already points at this launch's first column
  long long n;                     // columns of this launch (<= G8_MAX_COLS)
  int W, T, square;
  double scale, shift, lo, hi;
  unsigned clamp_word;
  int* planes;                     // debug: [panel][tile][7][128][32] raw accumulators, or nullptr
};

template <int TREG, bool SQUARE, bool DEBUG>
__global__ void __launch_bounds__(G8_THREADS, 1) series_gram_i8_kernel(const G8Params P) {
  extern __shared__ unsigned char g8_smem_raw[];
  const uint32_t raw = smem_u32(g8_smem_raw);
  unsigned char* smem = g8_smem_raw + (((raw + 1023u) & ~1023u) - raw);      // the swizzle pattern is on absolute address bits
  unsigned char* Bs = smem;
  unsigned char* As = Bs + (size_t)G8_B_STAGES * G8_B_TILE;
  unsigned char* Stg = As + (size_t)G8_A_BUFS * G8_A_TILE;
  int* expw_s = reinterpret_cast<int*>(Stg + (size_t)G8_EPI_WARPS * G8_STG_WARP);
  double* sa_s = reinterpret_cast<double*>(expw_s + G8_MAX_COLS);           // [G8_A_BUFS][128] row factors
  uint64_t* bars = reinterpret_cast<uint64_t*>(sa_s + G8_A_BUFS * F8_BM);
  uint64_t* b_full = bars;                          // [G8_B_STAGES]
  uint64_t* b_empty = b_full + G8_B_STAGES;         // [G8_B_STAGES]
  uint64_t* a_full = b_empty + G8_B_STAGES;         // [G8_A_BUFS]
  uint64_t* a_empty = a_full + G8_A_BUFS;           // [G8_A_BUFS]
  uint64_t* t_full = a_empty + G8_A_BUFS;           // [2]
  uint64_t* t_empty = t_full + 2;                   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const long long npanels = (P.m + F8_BM - 1) / F8_BM;
  const int ntiles = (int)((P.n + G8_NT - 1) / G8_NT);

  if (threadIdx.x == 0) {
    for (int s = 0; s < G8_B_STAGES; ++s) {
      mbar_init(b_full + s, 1);
      mbar_init(b_empty + s, 1);
    }
    for (int s = 0; s < G8_A_BUFS; ++s) {
      mbar_init(a_full + s, 1);
      mbar_init(a_empty + s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full + s, 1);
      mbar_init(t_empty + s, G8_EPI_WARPS);
    }
    fence_mbar_init();
  }
  for (int c = threadIdx.x; c < ntiles * G8_NT; c += G8_THREADS) {
    const int w = P.expw[c];
    expw_s[c] = (w == G8_BAD_COLUMN) ? 0 : w;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  f8_fence_before();
  __syncthreads();
  f8_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== B producer =====
    if (lane == 0) {
      uint32_t b_it = 0;
      for (long long pn = blockIdx.x; pn < npanels; pn += gridDim.x) {
        for (int t = 0; t < ntiles; ++t) {
          const uint32_t bs = b_it % G8_B_STAGES, bph = (b_it / G8_B_STAGES) & 1u;
          g8_mbar_sleep(b_empty + bs, bph ^ 1u, 200);
          mbar_expect_tx(b_full + bs, G8_B_TILE);
          bulk_g2s(Bs + (size_t)bs * G8_B_TILE, P.bimg + (long long)t * G8_B_TILE, G8_B_TILE, b_full + bs);
          ++b_it;
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    uint32_t b_it = 0, acc_it = 0, a_it = 0;
    for (long long pn = blockIdx.x; pn < npanels; pn += gridDim.x) {
      const uint32_t ab = a_it % G8_A_BUFS, aph = (a_it / G8_A_BUFS) & 1u;
      g8_mbar_sleep(a_full + ab, aph, 100);
      f8_fence_after();
      const uint32_t a_lo = f8_desc_lo(smem_u32(As + (size_t)ab * G8_A_TILE));
      for (int t = 0; t < ntiles; ++t) {
        const uint32_t bs = b_it % G8_B_STAGES, bph = (b_it / G8_B_STAGES) & 1u;
        const uint32_t acc = acc_it & 1u, cph = (acc_it >> 1) & 1u;
        g8_mbar_sleep(t_empty + acc, cph ^ 1u, 32);        // the epilogue has read this accumulator buffer
        g8_mbar_sleep(b_full + bs, bph, 32);
        f8_fence_after();
        const uint32_t b_lo = f8_desc_lo(smem_u32(Bs + (size_t)bs * G8_B_TILE));
        const uint32_t d0 = tmem_base + acc * G8_ACC_STRIDE;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          // chunk pair a against planes 2 a .. 6: B rows from 64 a on (8 KB per 64 rows), D columns from 64 a on
          f8_mma(d0 + 64u * a, f8_desc_from_lo(a_lo + 2 * a), f8_desc_from_lo(b_lo + ((64 * a * 128) >> 4) + 2 * a),
                 f8_idesc(G8_NT * (G8_S - 2 * a)), a != 0 ? 1u : 0u);
        }
        f8_commit(b_empty + bs);
        f8_commit(t_full + acc);
        ++b_it;
        ++acc_it;
      }
      f8_commit(a_empty + ab);
      ++a_it;
    }
  } else if (warp == 2) {
    // ===== A slicer: rows lane, lane + 32, lane + 64, lane + 96 of each panel =====
    uint32_t a_it = 0;
    for (long long pn = blockIdx.x; pn < npanels; pn += gridDim.x) {
      const uint32_t ab = a_it % G8_A_BUFS, aph = (a_it / G8_A_BUFS) & 1u;
      g8_mbar_sleep(a_empty + ab, aph ^ 1u, 500);
      unsigned char* tile = As + (size_t)ab * G8_A_TILE;
      for (int rr = 0; rr < 4; ++rr) {
        const int r = lane + 32 * rr;
        const long long row = pn * F8_BM + r;
        const double* src = P.x1 + (row < P.m ? row : 0) * P.ld1;
        double mx = 0.0;
        bool finite = true;
        for (int k = 0; k < P.W; ++k) {
          const double v = (row < P.m) ? src[k] : 0.0;
          mx = fmax(mx, fabs(v));
          finite = finite && (fabs(v) < 1e300);
        }
        const int e = finite ? g8_exponent(mx) : 0;
        const double down = ldexp(1.0, -e);
        // row factor of the epilogue: u = v * scale 2^ea + shift, or (v * sqrt(scale) 2^ea)^2 + shift
        const double fac = ldexp(SQUARE ? sqrt(P.scale) : P.scale, e);
        sa_s[ab * F8_BM + r] = finite ? fac : __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          unsigned w[G8_S];
#pragma unroll
          for (int p = 0; p < G8_S; ++p) w[p] = 0u;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int k = 4 * g + kk;
            const double v = (finite && row < P.m && k < P.W) ? src[k] * down : 0.0;
            int q[G8_S];
            g8_digits(v, q);
#pragma unroll
            for (int p = 0; p < G8_S; ++p) w[p] |= ((unsigned)q[p] & 0xffu) << (8 * kk);
          }
#pragma unroll
          for (int p = 0; p < G8_S; ++p) *reinterpret_cast<unsigned*>(tile + f8_swizzled(r, p) + 4 * g) = w[p];
          *reinterpret_cast<unsigned*>(tile + f8_swizzled(r, 7) + 4 * g) = 0u;
        }
      }
      fence_proxy_async();          // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) f8_mbar_arrive(a_full + ab);
      ++a_it;
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM lane quarter = warp % 4, column group (8 of the tile's 32 columns) = (warp - 4) / 4 =====
    const int ew = warp - 4;
    const int quarter = warp & 3, cg = ew >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    unsigned char* stg = Stg + (size_t)ew * G8_STG_WARP;
    double c[TREG];
#pragma unroll
    for (int k = 0; k < TREG; ++k) c[k] = (k < P.T) ? __ldg(P.coef + k) : 0.0;
    const unsigned thr = P.clamp_word;
    // M = 1.5 * 2^52 (bits 0x4338000000000000): the double with bits M + v is M + v for an integer |v| < 2^51
    // <x, y> 2^-(ea + eb) = hi 2^-32 + lo 2^-64 with hi = P0 2^16 + P1 2^8 + P2, lo = P3 2^24 + P4 2^16 + P5 2^8 + P6
    // (64-bit integers, exact).  As doubles M + 2^31 + hi and M + lo: t1 = (M + 2^31 + hi) 2^-32 + kC0 = hi 2^-32 - M 2^-64
    // is exact (36 bits), v = (M + lo) 2^-64 + t1 is the sum rounded ONCE.
    constexpr double kC1 = 2.3283064365386963e-10;            // 2^-32
    constexpr double kC2 = 5.421010862427522e-20;             // 2^-64
    constexpr double kC0 = -(1572864.5 + 0.0003662109375);    // -(M + 2^31) 2^-32 - M 2^-64: 2^20 .. 2^-13, exact
    const bool aligned = ((P.ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(P.K) & 15u) == 0);
    // Flattened (panel, tile) sequence of this CTA with the tensor-memory loads of the NEXT tile issued between the
    // integer recombination of the current one (after which the plane registers are free) and its Horner / store phase.
    const long long my_panels = (npanels - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const long long items = my_panels * ntiles;
    const uint32_t t_full_u = smem_u32(t_full), t_empty_u = smem_u32(t_empty);
    const int r = quarter * 32 + lane;
    const uint32_t tm_addr = lane_addr + (uint32_t)(G8_CW * cg);
    int pl[G8_S][G8_CW];
    auto wait_full = [&](uint32_t acc_i) {
      const uint32_t bar = t_full_u + 8u * (acc_i & 1u), par = (acc_i >> 1) & 1u;
      for (unsigned spins = 0; !g8_mbar_test_u(bar, par); ++spins) {
        __nanosleep(40);
        if (spins > (1u << 26)) __trap();
      }
      f8_fence_after();
    };
    auto issue_loads = [&](uint32_t acc_i) {
#pragma unroll
      for (int s = 0; s < G8_S; ++s) f8_tmem_ld8(tm_addr + (acc_i & 1u) * G8_ACC_STRIDE + (uint32_t)(s * G8_NT), pl[s]);
    };
    if (items > 0) {
      wait_full(0);
      issue_loads(0);
    }
    int fac_hi = 0, fac_lo = 0, t = 0;
    bool row_bad = false;
    long long pn = blockIdx.x;
    uint32_t ab = 0;
    for (long long it = 0; it < items; ++it) {
      const uint32_t acc_i = (uint32_t)it;
      const int colbase = t * G8_NT + G8_CW * cg;
      int ew8[G8_CW];
      *reinterpret_cast<int4*>(ew8) = *reinterpret_cast<const int4*>(expw_s + colbase);          // broadcast
      *reinterpret_cast<int4*>(ew8 + 4) = *reinterpret_cast<const int4*>(expw_s + colbase + 4);
      if (t == 0) {                                             // written by the slicer before a_full, which the MMAs waited for
        double fac = sa_s[ab * F8_BM + r];
        row_bad = fac != fac;                                   // NaN: the row holds a non-finite entry
        if (row_bad) fac = 0.0;
        fac_hi = __double2hiint(fac);
        fac_lo = __double2loint(fac);
      }
      f8_tmem_ld_wait();
      // every tensor-memory read of this warp is complete: hand the buffer back before the arithmetic
      f8_fence_before();
      __syncwarp();
      if (lane == 0) g8_mbar_arrive_u(t_empty_u + 8u * (acc_i & 1u));
      if constexpr (DEBUG) {
#pragma unroll
        for (int s = 0; s < G8_S; ++s) {
          int* dbg = P.planes + ((((pn * ntiles + t) * G8_S + s) * F8_BM + r) * G8_NT + G8_CW * cg);
#pragma unroll
          for (int e8 = 0; e8 < G8_CW; ++e8) dbg[e8] = pl[s][e8];
        }
      }
      double u[G8_CW];
      unsigned need = 0;
#pragma unroll
      for (int e8 = 0; e8 < G8_CW; ++e8) {
        // hi: the low word of M carries t + 2^31 (mod 2^32) = t ^ 0x80000000 for a signed 32-bit t, i.e. the double
        // M + 2^31 + t, no sign extension; lo: M + sext(t) (the 2^31 offset of a second biased word would make the
        // constant below a 54-bit number).  The wide multiply-adds put the upper planes on top.
        const int t12 = pl[1][e8] * 256 + pl[2][e8], t56 = pl[5][e8] * 256 + pl[6][e8];
        const long long hi = g8_mad_wide(pl[0][e8], 65536, g8_pack(0x43380000, t12 ^ 0x80000000));
        const long long lo = g8_mad_wide(pl[3][e8] * 256 + pl[4][e8], 65536, g8_pack(0x43380000 + (t56 >> 31), t56));
        const double t1 = fma(__longlong_as_double(hi), kC1, kC0);             // exact
        const double v = fma(__longlong_as_double(lo), kC2, t1);               // <x, y> 2^-(ea + eb), one rounding
        const double fij = __hiloint2double(fac_hi + ew8[e8], fac_lo);         // row factor times 2^eb of the column
        if constexpr (SQUARE) {
          const double w = v * fij;
          u[e8] = fma(w, w, P.shift);
        } else {
          u[e8] = fma(v, fij, P.shift);
        }
        need |= (unsigned)(((unsigned)__double2hiint(u[e8]) & 0x7fffffffu) >= thr);
      }
      const long long row0 = pn * F8_BM + quarter * 32;
      // the plane registers are free: start the next tile's loads now, they land during Horner and the stores
      if (it + 1 < items) {
        wait_full(acc_i + 1u);
        issue_loads(acc_i + 1u);
      }
      auto tail = [&](const double (&uu)[G8_CW]) {
        double pv[G8_CW];
#pragma unroll
        for (int e8 = 0; e8 < G8_CW; ++e8) pv[e8] = c[TREG - 1];
#pragma unroll
        for (int k = TREG - 2; k >= 0; --k) {
#pragma unroll
          for (int e8 = 0; e8 < G8_CW; ++e8) pv[e8] = fma(pv[e8], uu[e8], c[k]);
        }
        if (row_bad) {
#pragma unroll
          for (int e8 = 0; e8 < G8_CW; ++e8) pv[e8] = __longlong_as_double(0x7ff8000000000000LL);
        }
        double2* srow = reinterpret_cast<double2*>(stg + lane * G8_STG_ROW);
#pragma unroll
        for (int e8 = 0; e8 < G8_CW / 2; ++e8) srow[e8] = make_double2(pv[2 * e8], pv[2 * e8 + 1]);
      };
      if (need) {                                               // rare: the reference's clamp is active somewhere in these 8 columns
        double uc[G8_CW];
#pragma unroll
        for (int e8 = 0; e8 < G8_CW; ++e8) uc[e8] = (u[e8] < P.lo) ? P.lo : ((u[e8] > P.hi) ? P.hi : u[e8]);
        tail(uc);
      } else {
        tail(u);
      }
      __syncwarp();
      // transposed read: lane -> row lane % 8 (+ 8 i), 16-byte chunk lane / 8: a quarter warp reads one chunk of eight
      // rows (conflict-free with the 80-byte pitch); the four lanes of a row write its 64 contiguous bytes
      const bool full = aligned && (row0 + 32 <= P.m) && (colbase + G8_CW <= P.n);
      if (full) {
        double* outp = P.K + (row0 + (lane & 7)) * P.ldk + colbase + 2 * (lane >> 3);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double2 val = *reinterpret_cast<const double2*>(stg + (8 * i + (lane & 7)) * G8_STG_ROW + 16 * (lane >> 3));
          st_stream2(outp + (long long)(8 * i) * P.ldk, val.x, val.y);
        }
      } else {
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3), cc = lane & 7;
          const long long grow = row0 + rr, gcol = colbase + cc;
          if (grow < P.m && gcol < P.n) st_stream(P.K + grow * P.ldk + gcol, *reinterpret_cast<const double*>(stg + rr * G8_STG_ROW + 8 * cc));
        }
      }
      __syncwarp();
      if (++t == ntiles) {
        t = 0;
        pn += gridDim.x;
        ab ^= 1u;
      }
    }
  }
  f8_fence_before();
  __syncthreads();
  if (warp == 1) {
    f8_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

static long long g8_round_up(long long v, long long m) { return (v + m - 1) / m * m; }

static bool g8_supported(const mgb_series_desc* d) {
  return d != nullptr && d->n_factors == 1 && d->basis == MGB_BASIS_MONOMIAL && d->n_terms >= 1 && d->n_terms <= 12 &&
         d->factor_width >= 1 && d->factor_width <= G8_MAXW && (!d->pair_square || d->scale > 0.0);
}

}  // namespace mgb

using namespace mgb;

extern "C" int mgb_series_gram_i8_supported(const mgb_series_desc* d) { return g8_supported(d) ? 1 : 0; }

extern "C" long long mgb_series_gram_i8_workspace_bytes(long long n) {
  if (n < 1) return 0;
  const long long n_pad = g8_round_up(n, G8_NT);
  return n_pad / G8_NT * (long long)G8_B_TILE + g8_round_up(n_pad * 4, 256);
}

static int g8_launch(const mgb_series_desc* d, const double* x1, long long m, long long ld1, const double* x2, long long n, long long ld2,
                     const double* coef, double* K, long long ldk, void* workspace, long long workspace_bytes, int* planes, void* stream) {
  if (!g8_supported(d)) return fail(MGB_ERR_ARG, "series_gram_fwd_i8: single-factor monomial series with <= 12 terms and rows <= 16 wide only");
  if (m < 0 || n < 0 || (m > 0 && !x1) || (n > 0 && !x2) || !coef || ld1 < d->factor_width || ld2 < d->factor_width)
    return fail(MGB_ERR_ARG, "series_gram_fwd_i8: bad inputs");
  if (m == 0 || n == 0) return MGB_OK;
  if (!K || ldk < n) return fail(MGB_ERR_ARG, "series_gram_fwd_i8: bad K / ldk");
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u) || workspace_bytes < mgb_series_gram_i8_workspace_bytes(n))
    return fail(MGB_ERR_ARG, "series_gram_fwd_i8: workspace missing, not 256-byte aligned or smaller than mgb_series_gram_i8_workspace_bytes(n)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long n_pad = g8_round_up(n, G8_NT);
  unsigned char* img = static_cast<unsigned char*>(workspace);
  int* expw = reinterpret_cast<int*>(img + n_pad / G8_NT * (long long)G8_B_TILE);
  g8_slice_b_kernel<<<(unsigned)((n_pad + 127) / 128), 128, 0, s>>>(x2, n, ld2, d->factor_width, n_pad, img, expw);
  MGB_TRY(after_launch("g8_slice_b_kernel"));
  G8Params P{};
  P.x1 = x1; P.m = m; P.ld1 = ld1;
  P.coef = coef; P.ldk = ldk;
  P.W = d->factor_width; P.T = d->n_terms; P.square = d->pair_square ? 1 : 0;
  P.scale = d->scale; P.shift = d->shift; P.lo = d->lo; P.hi = d->hi;
  P.clamp_word = 0u;
  if (d->lo < 0.0 && d->hi > 0.0) {
    const double bound = fmin(-d->lo, d->hi);
    unsigned long long bits;
    memcpy(&bits, &bound, sizeof(bits));
    P.clamp_word = (unsigned)(bits >> 32);
  }
  const long long npanels = (m + F8_BM - 1) / F8_BM;
  long long grid = device_sm_count();
  if (grid > npanels) grid = npanels;
  void (*kernel)(const G8Params);
  if (planes != nullptr) kernel = d->pair_square ? series_gram_i8_kernel<12, true, true> : series_gram_i8_kernel<12, false, true>;
  else if (d->n_terms <= 10) kernel = d->pair_square ? series_gram_i8_kernel<10, true, false> : series_gram_i8_kernel<10, false, false>;
  else kernel = d->pair_square ? series_gram_i8_kernel<12, true, false> : series_gram_i8_kernel<12, false, false>;
  MGB_TRY(check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G8_SMEM), "cudaFuncSetAttribute"));
  for (long long c0 = 0; c0 < n; c0 += G8_MAX_COLS) {
    P.n = (n - c0 < G8_MAX_COLS) ? (n - c0) : G8_MAX_COLS;
    P.bimg = img + c0 / G8_NT * (long long)G8_B_TILE;
    P.expw = expw + c0;
    P.K = K + c0;
    P.planes = planes;
    kernel<<<(unsigned)grid, G8_THREADS, G8_SMEM, s>>>(P);
    MGB_TRY(after_launch("series_gram_i8_kernel"));
  }
  g8_poison_kernel<<<dim3((unsigned)((n + 255) / 256), 64), 256, 0, s>>>(expw, n, m, K, ldk);
  return after_launch("g8_poison_kernel");
}

extern "C" int mgb_series_gram_fwd_i8(const mgb_series_desc* d, const double* x1, long long m, long long ld1, const double* x2, long long n,
                                      long long ld2, const double* coef, double* K, long long ldk, void* workspace,
                                      long long workspace_bytes, void* stream) {
  return g8_launch(d, x1, m, ld1, x2, n, ld2, coef, K, ldk, workspace, workspace_bytes, nullptr, stream);
}

// Test hook: additionally dumps the raw int32 accumulator planes [panel][tile][7][128][32] (n <= 4096: one launch).
extern "C" int mgb_series_gram_fwd_i8_debug(const mgb_series_desc* d, const double* x1, long long m, long long ld1, const double* x2,
                                            long long n, long long ld2, const double* coef, double* K, long long ldk, void* workspace,
                                            long long workspace_bytes, int* planes, void* stream) {
  if (n > G8_MAX_COLS) return fail(MGB_ERR_ARG, "series_gram_fwd_i8_debug: n <= 4096");
  return g8_launch(d, x1, m, ld1, x2, n, ld2, coef, K, ldk, workspace, workspace_bytes, planes, stream);
}
