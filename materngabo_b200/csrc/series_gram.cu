// Spectral-series Gram kernels: K_ij = prod_f p_f(clamp(scale <x1_i[f], x2_j[f]> + shift)).
//
// One kernel family serves the sphere S^d (Gegenbauer series re-expanded on the host), SO(3)
// (character sum = Chebyshev series in cos(theta)), the circle and the torus T^d (product of
// cosine series).  Reference: kernel_utils/kernels_sphere.py:102-139,188-224,443-472,551-614,
// kernels_so.py:305-362, kernels_torus.py:142-177 (see include/mgb.h).
//
// Data layout / tiling (B200: 148 SMs, HBM3e write-bound at 8 B per entry):
//   - a CTA owns a 128-row x 64-column tile of K; 8 warps, warp w owns rows [16w, 16w+16);
//   - lane l owns columns 2l, 2l+1 of the tile: one 16-byte streaming store per row, a warp writes
//     512 contiguous bytes = 4 full 128-byte lines;
//   - the x1 row panel (128 x D doubles, contiguous in HBM) is staged in shared memory by ONE
//     cp.async.bulk (TMA engine, mbarrier completion) and read back as warp-wide broadcasts;
//   - the two x2 rows of a lane live in registers for the whole tile (fast path) or in a
//     transposed shared-memory panel (general path);
//   - inner products, clamp and the polynomial are evaluated in registers: D + T + 3 DFMA per entry
//     (monomial/Horner, T <= 12) or D + 2T + 3 (Chebyshev three-term recurrence).
#include "mgb_common.cuh"
#include <cstring>
#include <cstdlib>
#include <cmath>

namespace mgb {

constexpr int kTM = 128;   // rows per tile
constexpr int kTN = 64;    // cols per tile
constexpr int kThreads = 256;
constexpr int kRowsPerWarp = kTM / (kThreads / 32);

struct SeriesParams {
  const double* x1; long long m, ld1;
  const double* x2; long long n, ld2;
  const double* coef;
  double* K; long long ldk;
  int F, W, T;
  double scale, shift, lo, hi;
  long long ncc;        // column chunks
  unsigned clamp_word;  // high 32 bits of min(|lo|, |hi|): |u| below this word can never need clamping
  int row_splits;       // CTAs per column chunk; CTA s walks row panels s, s + row_splits, ...
  int tiled;            // 1: write the tile-major swizzled layout consumed by posterior_reduce_kernel
  long long nch;        // tiled: 16-column chunks per row (n rounded up to 16, / 16)
  int square;           // u = scale * dot^2 + shift instead of scale * dot + shift (mgb_series_desc::pair_square)
  long long bs1, bs2, bsk;  // batched launch (gridDim.z = batch): element strides between the batch members' x1 / x2 / K
  int batch;
};

// gridDim.z = batch members (gpytorch batch dimensions b x m x D against b x n x D): same tile code, shifted bases
__device__ __forceinline__ void batch_offset(SeriesParams& p) {
  const long long b = blockIdx.z;
  p.x1 += b * p.bs1;
  p.x2 += b * p.bs2;
  p.K += b * p.bsk;
}

// Address of K[grow][col].  Row-major, or the posterior's operand layout: 128 x 16 blocks stored
// [candidate tile][column chunk][row][16], each row's eight 16-byte chunks XOR-swizzled by 2*(row & 3) so that the
// DMMA fragment loads of the consumer are bank-conflict free after ONE 16 KB bulk copy per block.
__device__ __forceinline__ double* out_ptr(const SeriesParams& p, long long grow, long long col) {
  if (!p.tiled) return p.K + grow * p.ldk + col;
  const long long ctile = grow >> 7;
  const int r = (int)(grow & 127);
  const long long ch = col >> 4;
  const int chunk = (int)(col & 15) >> 1;
  return p.K + ((((ctile * p.nch + ch) << 7) + r) << 4) + (((chunk ^ ((r & 3) << 1)) << 1) | (int)(col & 1));
}

// row-major results are write-once (evict-first); the tiled chunk is re-read right away by the posterior kernel
__device__ __forceinline__ void st_keep2(double* p, double a, double b, int keep) {
  if (keep) *reinterpret_cast<double2*>(p) = make_double2(a, b);
  else st_stream2(p, a, b);
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Start staging rows [row0, row0+rows) of x into dst[r*D + k].  Contiguous, 16-byte aligned panels go through ONE
// cp.async.bulk (TMA engine) that completes on `bar` (returns true: the consumer must wait on the barrier);
// anything else is copied with plain loads (returns false: visible after the next __syncthreads()).
__device__ __forceinline__ bool stage_panel_async(double* dst, const double* x, long long row0, int rows, int D,
                                                  long long ld, uint64_t* bar) {
  const double* src = x + row0 * ld;
  const uint32_t bytes = static_cast<uint32_t>(rows) * D * 8u;
  const bool bulk = (ld == D) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) && ((bytes & 15u) == 0);
  if (bulk) {
    if (threadIdx.x == 0) {
      mbar_expect_tx(bar, bytes);
      bulk_g2s(dst, src, bytes, bar);
    }
  } else {
    for (int e = threadIdx.x; e < rows * D; e += blockDim.x) {
      int r = e / D, k = e - r * D;
      dst[e] = src[(long long)r * ld + k];
    }
  }
  return bulk;
}

// ---------------------------------------------------------------------------------------------
// Fast path: one factor, width W compile-time, monomial basis with T <= TREG coefficients in registers.
//   - scale is folded into the x2 registers and shift is the initial value of the dot accumulator;
//   - RB rows x 2 columns = 2*RB independent Horner chains per thread (DFMA latency hiding);
//   - the clamp costs two integer instructions per entry: a pair can only need clamping when the high
//     word of |u| reaches the high word of the bound, which is tested on the integer pipe; the exact
//     (torch.clamp-equivalent) clamp runs in a rarely taken slow path.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double clamp_exact(double u, double lo, double hi) {
  return (u < lo) ? lo : ((u > hi) ? hi : u);  // NaN stays NaN, like torch.clamp
}

template <int W, int TREG>
__global__ void __launch_bounds__(kThreads, 2) series_gram_fwd_fast(SeriesParams p) {
  batch_offset(p);
  constexpr int RB = 4;                       // rows per iteration: 8 independent chains per thread
  constexpr bool kRowsInRegs = (W <= 6);      // small rows are fetched with 16-byte broadcasts up front
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t bars[2];
  // two x1 panels (kTM * W doubles each): panel i+1 is in flight while panel i is evaluated

  const long long cc = blockIdx.x % p.ncc;
  const int split = (int)(blockIdx.x / p.ncc);
  const long long npanels = (p.m + kTM - 1) / kTM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  double c[TREG];
#pragma unroll
  for (int k = 0; k < TREG; ++k) c[k] = (k < p.T) ? __ldg(p.coef + k) : 0.0;

  const long long col = cc * kTN + 2 * lane;
  // square mode: sqrt(scale) is folded into x2 and u = dot * dot + shift
  const double fold = p.square ? sqrt(p.scale) : p.scale;
  const double init = p.square ? 0.0 : p.shift;
  double xa[W], xb[W];
  {
    const bool va = col < p.n, vb = col + 1 < p.n;
#pragma unroll
    for (int k = 0; k < W; ++k) {
      xa[k] = va ? fold * __ldg(p.x2 + col * p.ld2 + k) : 0.0;
      xb[k] = vb ? fold * __ldg(p.x2 + (col + 1) * p.ld2 + k) : 0.0;
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const bool vec_ok = (col + 1 < p.n) && (p.tiled || (((p.ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 15u) == 0)));
  const bool va = col < p.n, vb = col + 1 < p.n;
  const int r_begin = warp * kRowsPerWarp;
  const unsigned thr = p.clamp_word;

  long long rp = split;
  if (rp >= npanels) return;
  unsigned parity = 0;   // bit b = phase parity of bars[b]
  bool bulk_cur = stage_panel_async(smem, p.x1, rp * kTM, (int)min((long long)kTM, p.m - rp * kTM), W, p.ld1, &bars[0]);
  for (int it = 0; rp < npanels; rp += p.row_splits, ++it) {
    const int buf = it & 1;
    double* x1s = smem + buf * (kTM * W);
    const long long row0 = rp * kTM;
    const int rows = (int)min((long long)kTM, p.m - row0);
    // prefetch the next panel into the other buffer (every thread left it at the barrier ending iteration it-1)
    const long long rp_next = rp + p.row_splits;
    bool bulk_next = false;
    if (rp_next < npanels)
      bulk_next = stage_panel_async(smem + (buf ^ 1) * (kTM * W), p.x1, rp_next * kTM,
                                    (int)min((long long)kTM, p.m - rp_next * kTM), W, p.ld1, &bars[buf ^ 1]);
    if (bulk_cur) {
      while (!mbar_try_wait(&bars[buf], (parity >> buf) & 1u)) {
      }
      parity ^= 1u << buf;
    } else if (it == 0) {
      __syncthreads();   // plain-load staging of the very first panel
    }
#pragma unroll 1
  for (int r = r_begin; r < r_begin + kRowsPerWarp; r += RB) {
    if (r >= rows) break;
    double ua[RB], ub[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) ua[q] = ub[q] = init;
    if constexpr (kRowsInRegs) {
      double v[RB * W];
      const double2* src = reinterpret_cast<const double2*>(x1s + r * W);  // r % RB == 0 -> 16-byte aligned
#pragma unroll
      for (int i = 0; i < RB * W / 2; ++i) {
        const double2 t = src[i];  // warp-wide broadcast
        v[2 * i] = t.x;
        v[2 * i + 1] = t.y;
      }
#pragma unroll
      for (int k = 0; k < W; ++k) {
#pragma unroll
        for (int q = 0; q < RB; ++q) {
          ua[q] = fma(v[q * W + k], xa[k], ua[q]);
          ub[q] = fma(v[q * W + k], xb[k], ub[q]);
        }
      }
    } else {
      // wide rows: k-outer, operands straight from shared memory (broadcast), 2*RB chains in flight
      const double* xr = x1s + r * W;
#pragma unroll
      for (int k = 0; k < W; ++k) {
#pragma unroll
        for (int q = 0; q < RB; ++q) {
          const double v = xr[q * W + k];
          ua[q] = fma(v, xa[k], ua[q]);
          ub[q] = fma(v, xb[k], ub[q]);
        }
      }
    }
    if (p.square) {
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        ua[q] = fma(ua[q], ua[q], p.shift);
        ub[q] = fma(ub[q], ub[q], p.shift);
      }
    }
    unsigned need = 0;
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      need |= (unsigned)(((unsigned)__double2hiint(ua[q]) & 0x7fffffffu) >= thr);
      need |= (unsigned)(((unsigned)__double2hiint(ub[q]) & 0x7fffffffu) >= thr);
    }
    if (need) {
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        ua[q] = clamp_exact(ua[q], p.lo, p.hi);
        ub[q] = clamp_exact(ub[q], p.lo, p.hi);
      }
    }
    double pa[RB], pb[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) pa[q] = pb[q] = c[TREG - 1];
#pragma unroll
    for (int k = TREG - 2; k >= 0; --k) {
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        pa[q] = fma(pa[q], ua[q], c[k]);
        pb[q] = fma(pb[q], ub[q], c[k]);
      }
    }
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      if (r + q < rows) {
        double* out = out_ptr(p, row0 + r + q, col);
        if (vec_ok) {
          st_keep2(out, pa[q], pb[q], p.tiled);
        } else {
          if (va) st_stream(out, pa[q]);
          if (vb) st_stream(out + 1, pb[q]);
        }
      }
    }
  }
    __syncthreads();     // everyone is done with x1s[buf]; plain-load staging of the next panel is visible
    bulk_cur = bulk_next;
  }
}

// ---------------------------------------------------------------------------------------------
// Small problems (m * n below ~4 M entries: BASELINE configs[0], the 1000 x 1000 S^2 Gram; GP-fit sizes): the
// persistent-panel kernels above launch ~128 CTAs whose prologue (mbarrier set-up, bulk copy, two dependent DRAM
// round trips) is most of the 10 us they take.  Here nothing is staged and nothing is synchronised: a warp owns
// 8 (W <= 4) or 4 rows x 64 columns, a lane keeps its two x2 rows and the coefficients in registers, x1 rows arrive as warp-wide
// broadcast loads, and every load of the prologue is issued before the first use.  32-row x 64-column CTAs of 4
// warps: 512 CTAs for 1000 x 1000 - one wave over the 148 SMs with ~14 warps each to hide the FP64 latency.
// Same arithmetic as series_gram_fwd_fast (DFMA inner product with scale folded into x2, integer-pipe clamp test,
// Horner), so the results agree bit for bit with that kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kSmallThreads = 128;
__host__ __device__ constexpr int small_rows_per_warp(int W) { return W <= 4 ? 8 : 4; }   // wide rows: registers

template <int W, int TREG>
__global__ void __launch_bounds__(kSmallThreads) series_gram_fwd_small(SeriesParams p) {
  batch_offset(p);
  constexpr int RB = 4;
  constexpr int kSmallRowsPerWarp = small_rows_per_warp(W);
  constexpr int kSmallRows = (kSmallThreads / 32) * kSmallRowsPerWarp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long col = (long long)blockIdx.x * kTN + 2 * lane;
  const long long row_w = (long long)blockIdx.y * kSmallRows + warp * kSmallRowsPerWarp;
  if (row_w >= p.m) return;
  const bool va = col < p.n, vb = col + 1 < p.n;
  // every global load of the prologue is in flight before the first one is needed
  double xa[W], xb[W], c[TREG], v[kSmallRowsPerWarp][W];
  {
    const double* pa = p.x2 + (va ? col : 0) * p.ld2;
    const double* pb = p.x2 + (vb ? col + 1 : 0) * p.ld2;
#pragma unroll
    for (int k = 0; k < W; ++k) {
      xa[k] = __ldg(pa + k);
      xb[k] = __ldg(pb + k);
    }
#pragma unroll
    for (int q = 0; q < kSmallRowsPerWarp; ++q) {
      const long long r = (row_w + q < p.m) ? row_w + q : p.m - 1;
#pragma unroll
      for (int k = 0; k < W; ++k) v[q][k] = __ldg(p.x1 + r * p.ld1 + k);      // warp-wide broadcast
    }
#pragma unroll
    for (int k = 0; k < TREG; ++k) c[k] = (k < p.T) ? __ldg(p.coef + k) : 0.0;
  }
  const double fold = p.square ? sqrt(p.scale) : p.scale;
  const double init = p.square ? 0.0 : p.shift;
#pragma unroll
  for (int k = 0; k < W; ++k) {
    xa[k] = va ? fold * xa[k] : 0.0;
    xb[k] = vb ? fold * xb[k] : 0.0;
  }
  const bool vec_ok = vb && ((p.ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 15u) == 0);
  const unsigned thr = p.clamp_word;
#pragma unroll
  for (int r0 = 0; r0 < kSmallRowsPerWarp; r0 += RB) {
    double ua[RB], ub[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) ua[q] = ub[q] = init;
#pragma unroll
    for (int k = 0; k < W; ++k) {
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        ua[q] = fma(v[r0 + q][k], xa[k], ua[q]);
        ub[q] = fma(v[r0 + q][k], xb[k], ub[q]);
      }
    }
    if (p.square) {
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        ua[q] = fma(ua[q], ua[q], p.shift);
        ub[q] = fma(ub[q], ub[q], p.shift);
      }
    }
    unsigned need = 0;
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      need |= (unsigned)(((unsigned)__double2hiint(ua[q]) & 0x7fffffffu) >= thr);
      need |= (unsigned)(((unsigned)__double2hiint(ub[q]) & 0x7fffffffu) >= thr);
    }
    if (need) {
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        ua[q] = clamp_exact(ua[q], p.lo, p.hi);
        ub[q] = clamp_exact(ub[q], p.lo, p.hi);
      }
    }
    double pa[RB], pb[RB];
#pragma unroll
    for (int q = 0; q < RB; ++q) pa[q] = pb[q] = c[TREG - 1];
#pragma unroll
    for (int k = TREG - 2; k >= 0; --k) {
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        pa[q] = fma(pa[q], ua[q], c[k]);
        pb[q] = fma(pb[q], ub[q], c[k]);
      }
    }
#pragma unroll
    for (int q = 0; q < RB; ++q) {
      const long long r = row_w + r0 + q;
      if (r < p.m) {
        double* out = p.K + r * p.ldk + col;
        if (vec_ok) st_stream2(out, pa[q], pb[q]);
        else {
          if (va) st_stream(out, pa[q]);
          if (vb) st_stream(out + 1, pb[q]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tensor-pipe variant of the fast path: the W-wide inner products run as FP64 DMMA m8n8k4 fragments (x2 columns
// resident in registers as B fragments, x1 rows fetched from the shared-memory panel as A fragments, k padded
// with zeros to a multiple of 4), which moves W of the W + T + 1 FMAs per entry off the FP64 DFMA pipe: DMMA
// issues on the tensor pipe and overlaps the Horner chains.  The 8 warps tile the 128 x 64 panel as 4 x 2; a warp
// owns four 8-row fragments x four 8-column fragments and a lane ends up with columns 8b + 2(lane%4), +1 of row
// lane/4: 16-byte stores again.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int W, int TREG>
__global__ void __launch_bounds__(kThreads, 2) series_gram_fwd_mma(SeriesParams p) {
  batch_offset(p);
  constexpr int KS = (W + 3) / 4;
  constexpr int NB = 4;   // 8-column fragments per warp (32 columns)
  constexpr int NA = 4;   // 8-row fragments per warp (32 rows)
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t bars[2];

  const long long cc = blockIdx.x % p.ncc;
  const int split = (int)(blockIdx.x / p.ncc);
  const long long npanels = (p.m + kTM - 1) / kTM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int frow = lane >> 2, fk = lane & 3;
  const int wr = warp >> 1;   // rows [32 wr, 32 wr + 32) of the panel
  const int wq = warp & 1;    // columns [32 wq, 32 wq + 32) of the chunk

  double c[TREG];
#pragma unroll
  for (int k = 0; k < TREG; ++k) c[k] = (k < p.T) ? __ldg(p.coef + k) : 0.0;

  // B fragments: element (k = 4 s + fk, n = frow) of column fragment b is scale * x2[col0 + 8 b + frow][k]
  // (square mode: sqrt(scale) folded in, u = dot * dot + shift)
  const long long col0 = cc * kTN + 32 * wq;
  const double fold = p.square ? sqrt(p.scale) : p.scale;
  const double init = p.square ? 0.0 : p.shift;
  double bf[NB][KS];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const long long cj = col0 + 8 * b + frow;
#pragma unroll
    for (int sidx = 0; sidx < KS; ++sidx) {
      const int k = 4 * sidx + fk;
      bf[b][sidx] = (cj < p.n && k < W) ? fold * __ldg(p.x2 + cj * p.ld2 + k) : 0.0;
    }
  }
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const bool aligned = p.tiled || (((p.ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 15u) == 0));
  const unsigned thr = p.clamp_word;

  long long rp = split;
  if (rp >= npanels) return;
  unsigned parity = 0;
  bool bulk_cur = stage_panel_async(smem, p.x1, rp * kTM, (int)min((long long)kTM, p.m - rp * kTM), W, p.ld1, &bars[0]);
  for (int it = 0; rp < npanels; rp += p.row_splits, ++it) {
    const int buf = it & 1;
    const double* x1s = smem + buf * (kTM * W);
    const long long row0 = rp * kTM;
    const int rows = (int)min((long long)kTM, p.m - row0);
    const long long rp_next = rp + p.row_splits;
    bool bulk_next = false;
    if (rp_next < npanels)
      bulk_next = stage_panel_async(smem + (buf ^ 1) * (kTM * W), p.x1, rp_next * kTM,
                                    (int)min((long long)kTM, p.m - rp_next * kTM), W, p.ld1, &bars[buf ^ 1]);
    if (bulk_cur) {
      while (!mbar_try_wait(&bars[buf], (parity >> buf) & 1u)) {
      }
      parity ^= 1u << buf;
    } else if (it == 0) {
      __syncthreads();
    }
    // software pipeline over the four 8-row fragments: the DMMAs (tensor pipe) of fragment a+1 are issued before
    // the Horner chains (FP64 pipe) of fragment a, so both pipes stay busy inside one warp
    double u0[2][NB], u1[2][NB];
    auto inner_products = [&](int a, double (&o0)[NB], double (&o1)[NB]) {
      const int r = 32 * wr + 8 * a + frow;
      double af[KS];
#pragma unroll
      for (int sidx = 0; sidx < KS; ++sidx) {
        const int k = 4 * sidx + fk;
        af[sidx] = (k < W) ? x1s[r * W + k] : 0.0;
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        o0[b] = init;
        o1[b] = init;
#pragma unroll
        for (int sidx = 0; sidx < KS; ++sidx) dmma884(o0[b], o1[b], af[sidx], bf[b][sidx]);
      }
      if (p.square) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          o0[b] = fma(o0[b], o0[b], p.shift);
          o1[b] = fma(o1[b], o1[b], p.shift);
        }
      }
    };
    inner_products(0, u0[0], u1[0]);
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      const int cur = a & 1;
      if (a + 1 < NA) inner_products(a + 1, u0[cur ^ 1], u1[cur ^ 1]);
      unsigned need = 0;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        need |= (unsigned)(((unsigned)__double2hiint(u0[cur][b]) & 0x7fffffffu) >= thr);
        need |= (unsigned)(((unsigned)__double2hiint(u1[cur][b]) & 0x7fffffffu) >= thr);
      }
      if (need) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          u0[cur][b] = clamp_exact(u0[cur][b], p.lo, p.hi);
          u1[cur][b] = clamp_exact(u1[cur][b], p.lo, p.hi);
        }
      }
      double p0[NB], p1[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) p0[b] = p1[b] = c[TREG - 1];
#pragma unroll
      for (int k = TREG - 2; k >= 0; --k) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          p0[b] = fma(p0[b], u0[cur][b], c[k]);
          p1[b] = fma(p1[b], u1[cur][b], c[k]);
        }
      }
      const int r = 32 * wr + 8 * a + frow;
      if (r < rows) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const long long col = col0 + 8 * b + 2 * fk;
          double* out = out_ptr(p, row0 + r, col);
          if (aligned && col + 1 < p.n) {
            st_keep2(out, p0[b], p1[b], p.tiled);
          } else {
            if (col < p.n) st_stream(out, p0[b]);
            if (col + 1 < p.n) st_stream(out + 1, p1[b]);
          }
        }
      }
    }
    __syncthreads();
    bulk_cur = bulk_next;
  }
}

// ---------------------------------------------------------------------------------------------
// Warp-autonomous variant of the tensor-pipe kernel.  Same tiling, fragment mapping and arithmetic order as
// series_gram_fwd_mma (results are bit-identical), but every warp stages ITS OWN 32-row slice of the x1 panel
// (32 x W doubles, contiguous in HBM) with its own cp.async.bulk + mbarrier pair, double buffered, and never meets
// the other warps at a CTA barrier.  Why: DFMA and DMMA share one FP64 data path (tools/probe_peaks.py), and with
// only 2 CTAs x 8 warps per SM the per-panel __syncthreads of the CTA-staged version keeps the warps of a CTA in
// phase - they do address arithmetic, loads and stores together and leave the FP64 pipe idle together (ncu: pipe
// busy 2/3 of the time).  Autonomous warps drift out of phase and fill each other's gaps.
// ---------------------------------------------------------------------------------------------
template <int W, int TREG, bool TILED>
__global__ void __launch_bounds__(kThreads, 2) series_gram_fwd_warp(SeriesParams p) {
  batch_offset(p);
  constexpr int KS = (W + 3) / 4;
  constexpr int NB = 4;
  constexpr int NA = 4;
  constexpr int RBLK = 32;                     // rows per warp per panel
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(16) double smem[];   // [warp][2][RBLK * W]
  __shared__ __align__(8) uint64_t bars[kWarps][2];

  const long long cc = blockIdx.x % p.ncc;
  const int split = (int)(blockIdx.x / p.ncc);
  const long long npanels = (p.m + kTM - 1) / kTM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int frow = lane >> 2, fk = lane & 3;
  const int wr = warp >> 1;
  const int wq = warp & 1;

  double c[TREG];
#pragma unroll
  for (int k = 0; k < TREG; ++k) c[k] = (k < p.T) ? __ldg(p.coef + k) : 0.0;

  const long long col0 = cc * kTN + 32 * wq;
  const double fold = p.square ? sqrt(p.scale) : p.scale;
  const double init = p.square ? 0.0 : p.shift;
  double bf[NB][KS];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const long long cj = col0 + 8 * b + frow;
#pragma unroll
    for (int sidx = 0; sidx < KS; ++sidx) {
      const int k = 4 * sidx + fk;
      bf[b][sidx] = (cj < p.n && k < W) ? fold * __ldg(p.x2 + cj * p.ld2 + k) : 0.0;
    }
  }
  if (col0 >= p.n) return;                     // nothing to write for this warp (ragged last column chunk)
  uint64_t* mybar = bars[warp];
  if (lane == 0) {
    mbar_init(&mybar[0], 1);
    mbar_init(&mybar[1], 1);
    fence_mbar_init();
  }
  __syncwarp();
  double* wbuf = smem + (size_t)warp * 2 * RBLK * W;
  const bool can_bulk = (p.ld1 == W) && ((reinterpret_cast<uintptr_t>(p.x1) & 15u) == 0);

  // stage rows [rp*128 + 32 wr, +32) into buffer b; returns true when the copy completes on the mbarrier
  auto stage = [&](long long rp, int b) -> bool {
    const long long r0 = rp * kTM + RBLK * wr;
    const long long left = p.m - r0;
    double* dst = wbuf + b * (RBLK * W);
    if (left >= RBLK && can_bulk) {
      if (lane == 0) {
        mbar_expect_tx(&mybar[b], RBLK * W * 8u);
        bulk_g2s(dst, p.x1 + r0 * p.ld1, RBLK * W * 8u, &mybar[b]);
      }
      return true;
    }
    const int rows = left < 0 ? 0 : (left < RBLK ? (int)left : RBLK);
    for (int e = lane; e < RBLK * W; e += 32) {
      const int r = e / W, k = e - r * W;
      dst[e] = (r < rows) ? p.x1[(r0 + r) * p.ld1 + k] : 0.0;
    }
    __syncwarp();
    return false;
  };

  const bool aligned = TILED || (((p.ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 15u) == 0));
  const bool full_cols = aligned && (col0 + 32 <= p.n);   // warp-uniform: all 32 columns of this warp exist
  const unsigned thr = p.clamp_word;
  // Full tiles (the common case) store through one base pointer per fragment row with compile-time offsets: layout
  // (row-major or the posterior's tile-major swizzle) is a template parameter, alignment / bounds tests are hoisted.
  long long lane_off, a_step;
  int sw0 = 0, sw1 = 0;
  if constexpr (TILED) {
    sw0 = ((fk ^ ((frow & 3) << 1)) << 1);
    sw1 = (((4 + fk) ^ ((frow & 3) << 1)) << 1);
    lane_off = (((col0 >> 4) << 7) + 32 * wr + frow) << 4;   // + (row panel) * nch * 2048
    a_step = 8 * 16;
  } else {
    lane_off = (long long)frow * p.ldk + col0 + 2 * fk;
    a_step = 8 * p.ldk;
  }

  long long rp = split;
  if (rp >= npanels) return;
  unsigned parity = 0;
  bool bulk_cur = stage(rp, 0);
  for (int it = 0; rp < npanels; rp += p.row_splits, ++it) {
    const int buf = it & 1;
    const double* x1s = wbuf + buf * (RBLK * W);
    const long long row0 = rp * kTM + RBLK * wr;
    const long long left = p.m - row0;
    const int rows = left < 0 ? 0 : (left < RBLK ? (int)left : RBLK);
    const long long rp_next = rp + p.row_splits;
    bool bulk_next = false;
    if (rp_next < npanels) bulk_next = stage(rp_next, buf ^ 1);   // every lane finished with buf^1 at the __syncwarp below
    if (bulk_cur) {
      while (!mbar_try_wait(&mybar[buf], (parity >> buf) & 1u)) {
      }
      parity ^= 1u << buf;
    }
    if (rows > 0) {
      // Measured on B200 (1M x 2k): hoisted addressing helps the wide rows and the tiled layout (S^10 5.59e11 ->
      // 5.89e11 entries/s, posterior Gram chunk 0.40 -> 0.33 ms) but costs the narrow row-major ones (S^2 7.09e11 ->
      // 6.2e11): those keep the generic store path.
      constexpr bool kHoist = TILED || (W > 4);
      const bool fast = kHoist && full_cols && rows == RBLK;
      // row0 = rp * 128 + 32 * wr: tile-major layout addresses the 128-row panel rp, row-major the row itself
      double* outp = p.K + lane_off + (TILED ? rp * p.nch * 2048 : row0 * p.ldk);
      double u0[2][NB], u1[2][NB];
      auto inner_products = [&](int a, double (&o0)[NB], double (&o1)[NB]) {
        const int r = 8 * a + frow;
        double af[KS];
#pragma unroll
        for (int sidx = 0; sidx < KS; ++sidx) {
          const int k = 4 * sidx + fk;
          af[sidx] = (k < W) ? x1s[r * W + k] : 0.0;
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          o0[b] = init;
          o1[b] = init;
#pragma unroll
          for (int sidx = 0; sidx < KS; ++sidx) dmma884(o0[b], o1[b], af[sidx], bf[b][sidx]);
        }
        if (p.square) {
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            o0[b] = fma(o0[b], o0[b], p.shift);
            o1[b] = fma(o1[b], o1[b], p.shift);
          }
        }
      };
      inner_products(0, u0[0], u1[0]);
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        const int cur = a & 1;
        if (a + 1 < NA) inner_products(a + 1, u0[cur ^ 1], u1[cur ^ 1]);
        unsigned need = 0;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          need |= (unsigned)(((unsigned)__double2hiint(u0[cur][b]) & 0x7fffffffu) >= thr);
          need |= (unsigned)(((unsigned)__double2hiint(u1[cur][b]) & 0x7fffffffu) >= thr);
        }
        if (need) {
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            u0[cur][b] = clamp_exact(u0[cur][b], p.lo, p.hi);
            u1[cur][b] = clamp_exact(u1[cur][b], p.lo, p.hi);
          }
        }
        double p0[NB], p1[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) p0[b] = p1[b] = c[TREG - 1];
#pragma unroll
        for (int k = TREG - 2; k >= 0; --k) {
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            p0[b] = fma(p0[b], u0[cur][b], c[k]);
            p1[b] = fma(p1[b], u1[cur][b], c[k]);
          }
        }
        if (fast) {
          double* o = outp + a * a_step;
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            if constexpr (TILED) {
              *reinterpret_cast<double2*>(o + (b >> 1) * 2048 + ((b & 1) ? sw1 : sw0)) = make_double2(p0[b], p1[b]);
            } else {
              st_stream2(o + 8 * b, p0[b], p1[b]);
            }
          }
        } else {
          const int r = 8 * a + frow;
          if (r < rows) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const long long col = col0 + 8 * b + 2 * fk;
              double* out = out_ptr(p, row0 + r, col);
              if (aligned && col + 1 < p.n) {
                st_keep2(out, p0[b], p1[b], TILED);
              } else {
                if (col < p.n) st_stream(out, p0[b]);
                if (col + 1 < p.n) st_stream(out + 1, p1[b]);
              }
            }
          }
        }
      }
    }
    __syncwarp();
    bulk_cur = bulk_next;
  }
}

// Terms worth evaluating: trailing exact zeros of every factor's coefficient row are skipped (circle_coef_kernel writes
// the negligible tail of a torus factor as zeros on the device, so a CUDA-graph-captured fit needs no host read to
// truncate).  Warp-uniform; call with all lanes converged after the coefficients are in shared memory.
__device__ __forceinline__ int effective_terms(const double* cfs, int F, int T) {
  const int lane = threadIdx.x & 31;
  int te = 1;
  for (int f = 0; f < F; ++f) {
    int last = 0;
    for (int k0 = 0; k0 < T; k0 += 32) {
      const int k = k0 + lane;
      const unsigned b = __ballot_sync(0xffffffffu, k < T && cfs[f * T + k] != 0.0);
      if (b) last = k0 + 31 - __clz(b);
    }
    te = max(te, last + 1);
  }
  return te;
}

// ---------------------------------------------------------------------------------------------
// General path: F factors of runtime width W, T terms per factor, either basis; coefficients and the
// transposed x2 panel in shared memory.  Each thread evaluates 2 columns x 2 rows per coefficient read.
// ---------------------------------------------------------------------------------------------
template <int BASIS, int NE>
__device__ __forceinline__ void eval_n(const double* __restrict__ cf, int T, const double (&u)[NE], double (&s)[NE]) {
  if (BASIS == MGB_BASIS_MONOMIAL) {
    double c = cf[T - 1];
#pragma unroll
    for (int q = 0; q < NE; ++q) s[q] = c;
    for (int k = T - 2; k >= 0; --k) {
      c = cf[k];
#pragma unroll
      for (int q = 0; q < NE; ++q) s[q] = fma(s[q], u[q], c);
    }
  } else {
    // forward Chebyshev recurrence T_k = 2u T_{k-1} - T_{k-2}, accumulated on the fly
    double t0[NE], t1[NE], u2[NE];
    const double c0 = cf[0], c1 = (T > 1) ? cf[1] : 0.0;
#pragma unroll
    for (int q = 0; q < NE; ++q) {
      t0[q] = 1.0;
      t1[q] = u[q];
      u2[q] = 2.0 * u[q];
      s[q] = fma(c1, u[q], c0);
    }
    for (int k = 2; k < T; ++k) {
      const double c = cf[k];
#pragma unroll
      for (int q = 0; q < NE; ++q) {
        const double t2 = fma(u2[q], t1[q], -t0[q]);
        s[q] = fma(c, t2, s[q]);
        t0[q] = t1[q];
        t1[q] = t2;
      }
    }
  }
}

// RPT rows x 2 columns per thread and coefficient read: 8 independent chains for the monomial basis (one FMA per term:
// the chain latency, not the pipe, limited the 4-chain version), 4 for the Chebyshev basis (three live values per chain).
template <int BASIS>
__global__ void __launch_bounds__(kThreads, 2) series_gram_fwd_general(SeriesParams p) {
  batch_offset(p);
  constexpr int RPT = (BASIS == MGB_BASIS_MONOMIAL) ? 4 : 2;
  constexpr int NE = 2 * RPT;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) uint64_t bar;
  const int D = p.F * p.W;
  double* x1s = smem;                 // kTM * D
  double* x2t = x1s + kTM * D;        // D * kTN (transposed: [k][col])
  double* cfs = x2t + D * kTN;        // F * T

  const long long cc = blockIdx.x % p.ncc;
  const int split = (int)(blockIdx.x / p.ncc);
  const long long npanels = (p.m + kTM - 1) / kTM;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  // staged once per CTA: coefficients and the transposed x2 panel of this column chunk
  for (int e = threadIdx.x; e < p.F * p.T; e += blockDim.x) cfs[e] = p.coef[e];
  for (int e = threadIdx.x; e < D * kTN; e += blockDim.x) {
    const int cidx = e / D, k = e - cidx * D;
    const long long gc = cc * kTN + cidx;
    x2t[k * kTN + cidx] = (gc < p.n) ? p.x2[gc * p.ld2 + k] : 0.0;
  }
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int te = effective_terms(cfs, p.F, p.T);

  const long long col = cc * kTN + 2 * lane;
  const bool vec_ok = (col + 1 < p.n) && (p.tiled || (((p.ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 15u) == 0)));
  const int r_begin = warp * kRowsPerWarp;
  unsigned parity = 0;
  for (long long rp = split; rp < npanels; rp += p.row_splits) {
    const long long row0 = rp * kTM;
    const int rows = (int)min((long long)kTM, p.m - row0);
    if (stage_panel_async(x1s, p.x1, row0, rows, D, p.ld1, &bar)) {
      while (!mbar_try_wait(&bar, parity)) {
      }
      parity ^= 1u;
    } else {
      __syncthreads();
    }
    for (int rr = 0; rr < kRowsPerWarp; rr += RPT) {
      const int r0 = r_begin + rr;
      if (r0 >= rows) break;
      int rix[RPT];
#pragma unroll
      for (int h = 0; h < RPT; ++h) rix[h] = min(r0 + h, rows - 1);
      double prod[NE];
#pragma unroll
      for (int q = 0; q < NE; ++q) prod[q] = 1.0;
      for (int f = 0; f < p.F; ++f) {
        double d[NE];
#pragma unroll
        for (int q = 0; q < NE; ++q) d[q] = 0.0;
        for (int k = 0; k < p.W; ++k) {
          const int kk = f * p.W + k;
          const double2 xc = *reinterpret_cast<const double2*>(x2t + kk * kTN + 2 * lane);
#pragma unroll
          for (int h = 0; h < RPT; ++h) {
            const double v = x1s[rix[h] * D + kk];
            d[2 * h] = fma(v, xc.x, d[2 * h]);
            d[2 * h + 1] = fma(v, xc.y, d[2 * h + 1]);
          }
        }
        double u[NE], s[NE];
        unsigned need = 0;
#pragma unroll
        for (int q = 0; q < NE; ++q) {
          u[q] = fma(p.square ? d[q] * d[q] : d[q], p.scale, p.shift);
          need |= (unsigned)(((unsigned)__double2hiint(u[q]) & 0x7fffffffu) >= p.clamp_word);
        }
        if (need) {
#pragma unroll
          for (int q = 0; q < NE; ++q) u[q] = clamp_exact(u[q], p.lo, p.hi);
        }
        eval_n<BASIS, NE>(cfs + f * p.T, te, u, s);
#pragma unroll
        for (int q = 0; q < NE; ++q) prod[q] *= s[q];
      }
#pragma unroll
      for (int h = 0; h < RPT; ++h) {
        const int r = r0 + h;
        if (r >= rows) break;
        double* out = out_ptr(p, row0 + r, col);
        if (vec_ok) {
          st_keep2(out, prod[2 * h], prod[2 * h + 1], p.tiled);
        } else {
          if (col < p.n) st_stream(out, prod[2 * h]);
          if (col + 1 < p.n) st_stream(out + 1, prod[2 * h + 1]);
        }
      }
    }
    __syncthreads();   // x1s is restaged for the next panel
  }
}

// ---------------------------------------------------------------------------------------------
// Backward.  One thread per (row i, 32-column strip) walks its strip; per-coefficient sums are reduced
// warp -> block -> one atomicAdd per coefficient per CTA; input gradients are reduced along the row.
// Sizes here are GP-fit sizes (N <= a few thousand): latency-bound, kept simple and exact.
//   order = 1 : standard backward (gcoef and gx1 with p')
//   order = 2 : gx1-like contraction with p'' (Hessian-vector support), gcoef untouched
// G is addressed with two strides so that the x2-side gradient is the same kernel with roles swapped.
// ---------------------------------------------------------------------------------------------
struct SeriesBwdParams {
  const double* xa; long long ma, lda;   // "row" side
  const double* xb; long long nb, ldb;   // "column" side
  const double* coef;
  const double* G; long long gs_row, gs_col;
  double* gcoef;   // may be null
  double* gxa;     // may be null, ma x D contiguous
  int F, W, T, basis, order;
  double scale, shift, lo, hi;
};

constexpr int kBwdThreads = 128;
constexpr int kBwdMaxW = 32;

// One row i of factor f against all columns, TACC coefficient accumulators in registers (te <= TACC terms are live).
template <int TACC>
__device__ __forceinline__ void series_bwd_row(const SeriesBwdParams& p, int te, const double* cfs, double* gc_blk, const double* xi,
                                               long long i, int f) {
  const int D = p.F * p.W;
  const int lane = threadIdx.x & 31;
  double acc[TACC];
#pragma unroll
  for (int k = 0; k < TACC; ++k) acc[k] = 0.0;
  double gx[kBwdMaxW];
#pragma unroll
  for (int k = 0; k < kBwdMaxW; ++k) gx[k] = 0.0;
  const bool want_coef = (p.gcoef != nullptr) && (p.order == 1);
  const double* cf = cfs + f * p.T;

  for (long long j0 = 0; j0 < p.nb; j0 += 32) {
    const long long j = j0 + lane;
    const bool valid = j < p.nb;
    const long long jj = valid ? j : 0;
    double go = valid ? p.G[i * p.gs_row + jj * p.gs_col] : 0.0;
    const double* xj = p.xb + jj * p.ldb;
    // product of the other factors' values
    for (int h = 0; h < p.F; ++h) {
      if (h == f) continue;
      double dot = 0.0;
      for (int k = 0; k < p.W; ++k) dot = fma(xi[h * p.W + k], xj[h * p.W + k], dot);
      const double u = fmin(fmax(fma(dot, p.scale, p.shift), p.lo), p.hi);
      const double* ch = cfs + h * p.T;
      double s;
      if (p.basis == MGB_BASIS_MONOMIAL) {
        s = ch[te - 1];
        for (int k = te - 2; k >= 0; --k) s = fma(s, u, ch[k]);
      } else {
        double t0 = 1.0, t1 = u;
        s = ch[0] + ((te > 1) ? ch[1] * u : 0.0);
        for (int k = 2; k < te; ++k) {
          const double t2 = fma(2.0 * u, t1, -t0);
          s = fma(ch[k], t2, s);
          t0 = t1;
          t1 = t2;
        }
      }
      go *= s;
    }
    // this factor: basis values (for gcoef) and the order-th derivative (for gx)
    double dot = 0.0;
    for (int k = 0; k < p.W; ++k) dot = fma(xi[f * p.W + k], xj[f * p.W + k], dot);
    const double raw = fma(dot, p.scale, p.shift);
    const double u = fmin(fmax(raw, p.lo), p.hi);
    const bool inside = (raw >= p.lo) && (raw <= p.hi);  // torch.clamp passes gradient on the boundary
    double d1 = 0.0, d2 = 0.0;
    if (p.basis == MGB_BASIS_MONOMIAL) {
      double pw = 1.0, pw1 = 0.0, pw2 = 0.0;  // u^k, u^(k-1), u^(k-2)
#pragma unroll
      for (int k = 0; k < TACC; ++k) {
        if (k < te) {
          if (want_coef) acc[k] = fma(go, pw, acc[k]);
          d1 = fma((double)k * cf[k], pw1, d1);
          d2 = fma((double)(k * (k - 1)) * cf[k], pw2, d2);
          pw2 = pw1;
          pw1 = pw;
          pw *= u;
          if (k == 0) pw1 = 1.0;
          if (k == 1) pw2 = 1.0;
        }
      }
    } else {
      double t0 = 1.0, t1 = u, a0 = 0.0, a1 = 1.0, b0 = 0.0, b1 = 0.0;
      if (want_coef) {
        acc[0] += go;
        if (te > 1) acc[1] = fma(go, u, acc[1]);
      }
      d1 = (te > 1) ? cf[1] : 0.0;
#pragma unroll
      for (int k = 2; k < TACC; ++k) {
        if (k < te) {
          const double t2 = fma(2.0 * u, t1, -t0);
          const double a2 = fma(2.0 * u, a1, 2.0 * t1 - a0);
          const double b2 = fma(2.0 * u, b1, 4.0 * a1 - b0);
          if (want_coef) acc[k] = fma(go, t2, acc[k]);
          d1 = fma(cf[k], a2, d1);
          d2 = fma(cf[k], b2, d2);
          t0 = t1; t1 = t2; a0 = a1; a1 = a2; b0 = b1; b1 = b2;
        }
      }
    }
    if (p.gxa != nullptr && inside) {
      const double w = go * ((p.order == 1) ? d1 * p.scale : d2 * p.scale * p.scale);
#pragma unroll
      for (int k = 0; k < kBwdMaxW; ++k)
        if (k < p.W) gx[k] = fma(w, xj[f * p.W + k], gx[k]);
    }
  }
  if (want_coef) {
#pragma unroll
    for (int k = 0; k < TACC; ++k) {
      if (k < te) {
        const double v = warp_sum(acc[k]);
        if (lane == 0) atomicAdd(gc_blk + k, v);
      }
    }
  }
  if (p.gxa != nullptr) {
#pragma unroll
    for (int k = 0; k < kBwdMaxW; ++k) {
      if (k < p.W) {
        const double v = warp_sum(gx[k]);
        if (lane == 0) p.gxa[i * D + f * p.W + k] = v;
      }
    }
  }
}

// The untruncated case (more than 64 live terms: accumulators in local memory) as a call, so that its stack frame does not
// shape the register allocation of the common paths.
template <int TACC>
__device__ __noinline__ void series_bwd_row_big(const SeriesBwdParams& p, int te, const double* cfs, double* gc_blk, const double* xi,
                                                long long i, int f) {
  series_bwd_row<TACC>(p, te, cfs, gc_blk, xi, i, f);
}

// grid = (row groups, factors).  Block (g, f): warp w handles row i = 4g + w for factor f.  TMAX bounds the host's
// n_terms; the live term count (trailing zero coefficients skipped) picks the accumulator count at run time, so a
// truncated 201-term torus factor runs with 64 register accumulators instead of 256 spilled ones.
template <int TMAX>
__global__ void __launch_bounds__(kBwdThreads) series_gram_bwd_kernel(SeriesBwdParams p) {
  extern __shared__ double sm[];
  const int D = p.F * p.W;
  const int f = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* cfs = sm;                          // F*T
  double* gc_blk = cfs + p.F * p.T;          // T
  double* xrows = gc_blk + p.T;              // warps * D
  for (int e = threadIdx.x; e < p.F * p.T; e += blockDim.x) cfs[e] = p.coef[e];
  for (int e = threadIdx.x; e < p.T; e += blockDim.x) gc_blk[e] = 0.0;
  const long long i = (long long)blockIdx.x * (kBwdThreads / 32) + warp;
  const bool row_ok = i < p.ma;
  double* xi = xrows + warp * D;
  if (row_ok)
    for (int k = lane; k < D; k += 32) xi[k] = p.xa[i * p.lda + k];
  __syncthreads();
  const int te = effective_terms(cfs, p.F, p.T);
  if (row_ok) {
    if (TMAX > 64 && te > 64) series_bwd_row_big<TMAX>(p, te, cfs, gc_blk, xi, i, f);
    else if (TMAX > 16 && te > 16) series_bwd_row<(TMAX < 64 ? TMAX : 64)>(p, te, cfs, gc_blk, xi, i, f);
    else series_bwd_row<(TMAX < 16 ? TMAX : 16)>(p, te, cfs, gc_blk, xi, i, f);
  }
  __syncthreads();
  if (p.gcoef != nullptr && p.order == 1)
    for (int e = threadIdx.x; e < p.T; e += blockDim.x) atomicAdd(p.gcoef + f * p.T + e, gc_blk[e]);
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
static int validate(const mgb_series_desc* d, const void* x1, long long m, long long ld1, const void* x2, long long n,
                    long long ld2, const void* coef) {
  if (!d || !coef || (m > 0 && !x1) || (n > 0 && !x2)) return fail(MGB_ERR_ARG, "series_gram: null pointer");
  if (d->n_factors < 1 || d->n_factors > MGB_MAX_FACTORS) return fail(MGB_ERR_ARG, "series_gram: n_factors out of range");
  if (d->factor_width < 1) return fail(MGB_ERR_ARG, "series_gram: factor_width < 1");
  if (d->n_terms < 1 || d->n_terms > MGB_MAX_TERMS) return fail(MGB_ERR_ARG, "series_gram: n_terms out of range");
  if (d->basis != MGB_BASIS_MONOMIAL && d->basis != MGB_BASIS_CHEBYSHEV) return fail(MGB_ERR_ARG, "series_gram: bad basis");
  const long long D = (long long)d->n_factors * d->factor_width;
  if (m < 0 || n < 0 || ld1 < D || ld2 < D) return fail(MGB_ERR_ARG, "series_gram: bad sizes / leading dimensions");
  if (!(d->lo <= d->hi)) return fail(MGB_ERR_ARG, "series_gram: lo > hi");
  return MGB_OK;
}

template <int W, int TREG>
static int launch_fast(const SeriesParams& p, long long tiles, cudaStream_t s) {
  const size_t smem = sizeof(double) * 2 * kTM * W;
  series_gram_fwd_fast<W, TREG><<<dim3((unsigned)tiles, 1, (unsigned)p.batch), kThreads, smem, s>>>(p);
  return after_launch("series_gram_fwd_fast");
}

template <int W, int TREG>
static int launch_mma(const SeriesParams& p, long long tiles, cudaStream_t s) {
  const size_t smem = sizeof(double) * 2 * kTM * W;
  series_gram_fwd_mma<W, TREG><<<dim3((unsigned)tiles, 1, (unsigned)p.batch), kThreads, smem, s>>>(p);
  return after_launch("series_gram_fwd_mma");
}

template <int W, int TREG>
static int launch_warp(const SeriesParams& p, long long tiles, cudaStream_t s) {
  const size_t smem = sizeof(double) * (kThreads / 32) * 2 * 32 * W;
  if (p.tiled) series_gram_fwd_warp<W, TREG, true><<<dim3((unsigned)tiles, 1, (unsigned)p.batch), kThreads, smem, s>>>(p);
  else series_gram_fwd_warp<W, TREG, false><<<dim3((unsigned)tiles, 1, (unsigned)p.batch), kThreads, smem, s>>>(p);
  return after_launch("series_gram_fwd_warp");
}

// 0 = DFMA inner products, 1 = DMMA inner products (CTA-staged panels), 2 = DMMA, warp-autonomous staging.  Default: DMMA wherever it is instantiated - measured on B200
// (1M x 2k): S^2 6.95e11 vs 6.38e11 entries/s, S^10 5.46e11 vs 4.46e11, SO(3) 5.28e11 vs 4.65e11 (profiles/).
// MGB_SERIES_PATH=dfma|mma|warp overrides (tuning knob).
static int series_path(int W) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("MGB_SERIES_PATH");
    forced = (e == nullptr) ? -1 : (strcmp(e, "warp") == 0 ? 2 : (strcmp(e, "mma") == 0 ? 1 : (strcmp(e, "dfma") == 0 ? 0 : -1)));
  }
  if (forced >= 0) return forced;
  (void)W;
  return 2;   // measured on B200, 1M x 2k: S^2 7.01e11 vs 6.79e11 entries/s, SO(3) (quaternion form) 6.26e11 vs 6.05e11,
              // S^10 5.58e11 vs 5.56e11 (gpurun_out/series_paths2.log)
}

template <int W, int TREG>
static int launch_small(const SeriesParams& p, cudaStream_t s) {
  constexpr int kSmallRows = (kSmallThreads / 32) * small_rows_per_warp(W);
  dim3 grid((unsigned)p.ncc, (unsigned)((p.m + kSmallRows - 1) / kSmallRows), (unsigned)p.batch);
  series_gram_fwd_small<W, TREG><<<grid, kSmallThreads, 0, s>>>(p);
  return after_launch("series_gram_fwd_small");
}

// MGB_SERIES_SMALL=0 disables the small-problem kernel (tuning / comparison knob)
static bool series_small_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("MGB_SERIES_SMALL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

template <int TREG>
static int dispatch_small(const SeriesParams& p, cudaStream_t s, bool* handled) {
  *handled = true;
  switch (p.W) {
    case 2: return launch_small<2, TREG>(p, s);
    case 3: return launch_small<3, TREG>(p, s);
    case 4: return launch_small<4, TREG>(p, s);
    case 9: return launch_small<9, TREG>(p, s);
    case 11: return launch_small<11, TREG>(p, s);
    default: *handled = false; return MGB_OK;
  }
}

template <int TREG>
static int dispatch_fast(const SeriesParams& p, long long tiles, cudaStream_t s, bool* handled) {
  *handled = true;
  if (series_path(p.W) == 2) {
    switch (p.W) {
      case 3: return launch_warp<3, TREG>(p, tiles, s);
      case 4: return launch_warp<4, TREG>(p, tiles, s);
      case 9: return launch_warp<9, TREG>(p, tiles, s);
      case 11: return launch_warp<11, TREG>(p, tiles, s);
      default: break;
    }
  }
  if (series_path(p.W) == 1) {
    switch (p.W) {
      case 3: return launch_mma<3, TREG>(p, tiles, s);
      case 4: return launch_mma<4, TREG>(p, tiles, s);
      case 9: return launch_mma<9, TREG>(p, tiles, s);
      case 11: return launch_mma<11, TREG>(p, tiles, s);
      default: break;
    }
  }
  switch (p.W) {
    case 2: return launch_fast<2, TREG>(p, tiles, s);
    case 3: return launch_fast<3, TREG>(p, tiles, s);
    case 4: return launch_fast<4, TREG>(p, tiles, s);
    case 5: return launch_fast<5, TREG>(p, tiles, s);
    case 6: return launch_fast<6, TREG>(p, tiles, s);
    case 9: return launch_fast<9, TREG>(p, tiles, s);
    case 11: return launch_fast<11, TREG>(p, tiles, s);
    default: *handled = false; return MGB_OK;
  }
}

}  // namespace mgb

using namespace mgb;

// Shepperd's method; one thread per rotation.  defect: max(|R R^T - I|_max, |det R - 1|) over all rows (atomicMax on
// the bit pattern: non-negative doubles order like unsigned integers).
__global__ void __launch_bounds__(256) so3_quat_kernel(const double* __restrict__ R, long long m, long long ld,
                                                       double* __restrict__ Q, double* defect) {
  double worst = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
    const double* r = R + i * ld;
    const double a = r[0], b = r[1], c = r[2], d = r[3], e = r[4], f = r[5], g = r[6], h = r[7], k = r[8];
    const double tr = a + e + k;
    double w, x, y, z;
    if (tr >= a && tr >= e && tr >= k) {
      w = 1.0 + tr; x = h - f; y = c - g; z = d - b;
    } else if (a >= e && a >= k) {
      w = h - f; x = 1.0 + a - e - k; y = b + d; z = c + g;
    } else if (e >= k) {
      w = c - g; x = b + d; y = 1.0 - a + e - k; z = f + h;
    } else {
      w = d - b; x = c + g; y = f + h; z = 1.0 - a - e + k;
    }
    const double inv = rsqrt(w * w + x * x + y * y + z * z);
    double* q = Q + i * 4;
    q[0] = w * inv; q[1] = x * inv; q[2] = y * inv; q[3] = z * inv;
    double dv = fabs(a * a + b * b + c * c - 1.0);
    dv = fmax(dv, fabs(d * d + e * e + f * f - 1.0));
    dv = fmax(dv, fabs(g * g + h * h + k * k - 1.0));
    dv = fmax(dv, fabs(a * d + b * e + c * f));
    dv = fmax(dv, fabs(a * g + b * h + c * k));
    dv = fmax(dv, fabs(d * g + e * h + f * k));
    const double det = a * (e * k - f * h) - b * (d * k - f * g) + c * (d * h - e * g);
    dv = fmax(dv, fabs(det - 1.0));
    if (!(dv == dv)) dv = 1.0;   // NaN input: not a rotation
    worst = fmax(worst, dv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
  if ((threadIdx.x & 31) == 0 && worst > 0.0)
    atomicMax(reinterpret_cast<unsigned long long*>(defect), (unsigned long long)__double_as_longlong(worst));
}

extern "C" int mgb_so3_to_quaternion(const double* R, long long m, long long ld, double* Q, double* defect,
                                     void* stream) {
  if (m < 0 || ld < 9 || !defect || (m > 0 && (!R || !Q))) return fail(MGB_ERR_ARG, "so3_to_quaternion: bad argument");
  if (m == 0) return MGB_OK;
  long long blocks = (m + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  so3_quat_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(R, m, ld, Q, defect);
  return after_launch("so3_quat_kernel");
}

static int series_sm_count() { return device_sm_count(); }

static int series_fwd_impl(const mgb_series_desc* d, const double* x1, long long m, long long ld1, const double* x2,
                           long long n, long long ld2, const double* coef, double* K, long long ldk, int tiled,
                           void* stream, long long batch = 1, long long bs1 = 0, long long bs2 = 0, long long bsk = 0) {
  MGB_TRY(validate(d, x1, m, ld1, x2, n, ld2, coef));
  if (m == 0 || n == 0) return MGB_OK;
  if (!K || (!tiled && ldk < n)) return fail(MGB_ERR_ARG, "series_gram_fwd: bad K / ldk");
  if (tiled && (reinterpret_cast<uintptr_t>(K) & 15u)) return fail(MGB_ERR_ARG, "series_gram_tiled: K must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SeriesParams p{x1, m, ld1, x2, n, ld2, coef, K, ldk, d->n_factors, d->factor_width, d->n_terms,
                 d->scale, d->shift, d->lo, d->hi, (n + kTN - 1) / kTN, 0u, 1, tiled, (n + 15) / 16,
                 d->pair_square ? 1 : 0, bs1, bs2, bsk, (int)batch};
  if (batch < 1 || batch > 65535) return fail(MGB_ERR_ARG, "series_gram_fwd: batch must be in [1, 65535] per launch");
  if (d->pair_square && !(d->scale > 0.0)) return fail(MGB_ERR_ARG, "series_gram_fwd: pair_square needs scale > 0");
  if (d->lo < 0.0 && d->hi > 0.0) {
    const double bound = fmin(-d->lo, d->hi);
    unsigned long long bits;
    memcpy(&bits, &bound, sizeof(bits));
    p.clamp_word = (unsigned)(bits >> 32);
  }
  // persistent over row panels: about 8 CTAs per SM in total, each walking a strided set of 128-row panels of
  // its 64-column chunk with the x2 rows / coefficients resident in registers
  const long long npanels = (m + kTM - 1) / kTM;
  long long splits = (8LL * series_sm_count() + p.ncc * batch - 1) / (p.ncc * batch);
  if (splits > npanels) splits = npanels;
  if (splits < 1) splits = 1;
  p.row_splits = (int)splits;
  const long long tiles = p.ncc * splits;
  if (tiles > 0x7fffffffLL) return fail(MGB_ERR_ARG, "series_gram_fwd: too many tiles for one launch");
  if (d->n_factors == 1 && d->basis == MGB_BASIS_MONOMIAL && d->n_terms <= 12 && !tiled && series_small_enabled() &&
      m * n * batch < (4LL << 20) && m <= 16 * 65535) {
    bool handled = false;
    int rc = (d->n_terms <= 10) ? dispatch_small<10>(p, s, &handled) : dispatch_small<12>(p, s, &handled);
    if (handled) return rc;
  }
  if (d->n_factors == 1 && d->basis == MGB_BASIS_MONOMIAL && d->n_terms <= 12) {
    bool handled = false;
    int rc = (d->n_terms <= 10) ? dispatch_fast<10>(p, tiles, s, &handled) : dispatch_fast<12>(p, tiles, s, &handled);
    if (handled) return rc;
  }
  const int D = d->n_factors * d->factor_width;
  const size_t smem = sizeof(double) * ((size_t)kTM * D + (size_t)D * kTN + (size_t)d->n_factors * d->n_terms);
  if (smem > 200 * 1024) return fail(MGB_ERR_ARG, "series_gram_fwd: row width too large for the shared-memory panels");
  if (d->basis == MGB_BASIS_MONOMIAL) {
    MGB_TRY(check_cuda(cudaFuncSetAttribute(series_gram_fwd_general<MGB_BASIS_MONOMIAL>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute"));
    series_gram_fwd_general<MGB_BASIS_MONOMIAL><<<dim3((unsigned)tiles, 1, (unsigned)p.batch), kThreads, smem, s>>>(p);
  } else {
    MGB_TRY(check_cuda(cudaFuncSetAttribute(series_gram_fwd_general<MGB_BASIS_CHEBYSHEV>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute"));
    series_gram_fwd_general<MGB_BASIS_CHEBYSHEV><<<dim3((unsigned)tiles, 1, (unsigned)p.batch), kThreads, smem, s>>>(p);
  }
  return after_launch("series_gram_fwd_general");
}

extern "C" int mgb_series_gram_fwd(const mgb_series_desc* d, const double* x1, long long m, long long ld1,
                                   const double* x2, long long n, long long ld2, const double* coef, double* K,
                                   long long ldk, void* stream) {
  return series_fwd_impl(d, x1, m, ld1, x2, n, ld2, coef, K, ldk, 0, stream);
}

extern "C" int mgb_series_gram_fwd_batched(const mgb_series_desc* d, const double* x1, long long m, long long ld1,
                                           long long stride1, const double* x2, long long n, long long ld2,
                                           long long stride2, const double* coef, double* K, long long ldk,
                                           long long stridek, long long batch, void* stream) {
  if (batch < 0 || stride1 < 0 || stride2 < 0 || stridek < 0) return fail(MGB_ERR_ARG, "series_gram_fwd_batched: bad batch / strides");
  for (long long b0 = 0; b0 < batch; b0 += 65535) {     // gridDim.z limit
    const long long nb = (batch - b0 < 65535) ? (batch - b0) : 65535;
    MGB_TRY(series_fwd_impl(d, x1 + b0 * stride1, m, ld1, x2 + b0 * stride2, n, ld2, coef, K + b0 * stridek, ldk, 0, stream, nb,
                            stride1, stride2, stridek));
  }
  return MGB_OK;
}

namespace mgb {
// out[i] = prod_f p_f(clamp(scale <x1_i[f], x2_i[f]> + shift)): the diagonal of the Gram matrix of two equally long
// point lists, one thread per row (gpytorch's diag=True: the test x test block of the posterior variance).
template <int BASIS>
__global__ void __launch_bounds__(256) series_gram_diag_kernel(SeriesParams p, double* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.m) return;
  const double* a = p.x1 + i * p.ld1;
  const double* b = p.x2 + i * p.ld2;
  double prod = 1.0;
  for (int f = 0; f < p.F; ++f) {
    double dsum = 0.0;
    for (int k = 0; k < p.W; ++k) dsum = fma(a[f * p.W + k], b[f * p.W + k], dsum);
    double u[1], sv[1];
    u[0] = clamp_exact(fma(p.square ? dsum * dsum : dsum, p.scale, p.shift), p.lo, p.hi);
    eval_n<BASIS, 1>(p.coef + f * p.T, p.T, u, sv);
    prod *= sv[0];
  }
  out[i] = prod;
}
}  // namespace mgb

extern "C" int mgb_series_gram_diag(const mgb_series_desc* d, const double* x1, long long ld1, const double* x2,
                                    long long ld2, long long rows, const double* coef, double* out, void* stream) {
  MGB_TRY(validate(d, x1, rows, ld1, x2, rows, ld2, coef));
  if (rows == 0) return MGB_OK;
  if (!out) return fail(MGB_ERR_ARG, "series_gram_diag: null output");
  SeriesParams p{x1, rows, ld1, x2, rows, ld2, coef, nullptr, 0, d->n_factors, d->factor_width, d->n_terms,
                 d->scale, d->shift, d->lo, d->hi, 0, 0u, 1, 0, 0, d->pair_square ? 1 : 0, 0, 0, 0, 1};
  const unsigned blocks = (unsigned)((rows + 255) / 256);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d->basis == MGB_BASIS_MONOMIAL) series_gram_diag_kernel<MGB_BASIS_MONOMIAL><<<blocks, 256, 0, s>>>(p, out);
  else series_gram_diag_kernel<MGB_BASIS_CHEBYSHEV><<<blocks, 256, 0, s>>>(p, out);
  return after_launch("series_gram_diag_kernel");
}

extern "C" long long mgb_posterior_tiled_bytes(long long m, long long n) {
  if (m < 0 || n < 0) return 0;
  return ((m + 127) / 128 * 128) * ((n + 15) / 16 * 16) * (long long)sizeof(double);
}

extern "C" int mgb_series_gram_tiled(const mgb_series_desc* d, const double* x1, long long m, long long ld1,
                                     const double* x2, long long n, long long ld2, const double* coef, double* Kb,
                                     void* stream) {
  return series_fwd_impl(d, x1, m, ld1, x2, n, ld2, coef, Kb, 0, 1, stream);
}

static int launch_bwd(const mgb_series_desc* d, int order, const double* xa, long long ma, long long lda,
                      const double* xb, long long nb, long long ldb, const double* coef, const double* G,
                      long long gs_row, long long gs_col, double* gcoef, double* gxa, cudaStream_t s) {
  if (ma == 0) return MGB_OK;
  SeriesBwdParams p{xa, ma, lda, xb, nb, ldb, coef, G, gs_row, gs_col, gcoef, gxa,
                    d->n_factors, d->factor_width, d->n_terms, d->basis, order, d->scale, d->shift, d->lo, d->hi};
  const int D = d->n_factors * d->factor_width;
  const size_t smem = sizeof(double) * ((size_t)d->n_factors * d->n_terms + d->n_terms + (size_t)(kBwdThreads / 32) * D);
  const long long blocks = (ma + (kBwdThreads / 32) - 1) / (kBwdThreads / 32);
  dim3 grid((unsigned)blocks, (unsigned)d->n_factors);
  if (smem > 48 * 1024) return fail(MGB_ERR_ARG, "series_gram_bwd: F*T too large");
  if (d->n_terms <= 16) series_gram_bwd_kernel<16><<<grid, kBwdThreads, smem, s>>>(p);
  else if (d->n_terms <= 64) series_gram_bwd_kernel<64><<<grid, kBwdThreads, smem, s>>>(p);
  else series_gram_bwd_kernel<MGB_MAX_TERMS><<<grid, kBwdThreads, smem, s>>>(p);
  return after_launch("series_gram_bwd_kernel");
}

extern "C" int mgb_series_gram_bwd(const mgb_series_desc* d, const double* x1, long long m, long long ld1,
                                   const double* x2, long long n, long long ld2, const double* coef, const double* G,
                                   long long ldg_row, long long ldg_col, double* gcoef, double* gx1, double* gx2,
                                   void* stream) {
  MGB_TRY(validate(d, x1, m, ld1, x2, n, ld2, coef));
  if (d->pair_square) return fail(MGB_ERR_ARG, "series_gram_bwd: pair_square is a forward-only form");
  if (!G) return fail(MGB_ERR_ARG, "series_gram_bwd: null G");
  if (d->factor_width > kBwdMaxW) return fail(MGB_ERR_ARG, "series_gram_bwd: factor width > 32");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (m == 0 || n == 0) return MGB_OK;
  if (gcoef || gx1) MGB_TRY(launch_bwd(d, 1, x1, m, ld1, x2, n, ld2, coef, G, ldg_row, ldg_col, gcoef, gx1, s));
  if (gx2) MGB_TRY(launch_bwd(d, 1, x2, n, ld2, x1, m, ld1, coef, G, ldg_col, ldg_row, nullptr, gx2, s));
  return MGB_OK;
}

extern "C" int mgb_series_gram_dx(const mgb_series_desc* d, int order, const double* x1, long long m, long long ld1,
                                  const double* x2, long long n, long long ld2, const double* coef, const double* G,
                                  long long ldg_row, long long ldg_col, double* out1, double* out2, void* stream) {
  MGB_TRY(validate(d, x1, m, ld1, x2, n, ld2, coef));
  if (d->pair_square) return fail(MGB_ERR_ARG, "series_gram_dx: pair_square is a forward-only form");
  if (!G) return fail(MGB_ERR_ARG, "series_gram_dx: null G");
  if (order != 1 && order != 2) return fail(MGB_ERR_ARG, "series_gram_dx: order must be 1 or 2");
  if (d->n_factors != 1) return fail(MGB_ERR_ARG, "series_gram_dx: single-factor kernels only");
  if (d->factor_width > kBwdMaxW) return fail(MGB_ERR_ARG, "series_gram_dx: factor width > 32");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (m == 0 || n == 0) return MGB_OK;
  if (out1) MGB_TRY(launch_bwd(d, order, x1, m, ld1, x2, n, ld2, coef, G, ldg_row, ldg_col, nullptr, out1, s));
  if (out2) MGB_TRY(launch_bwd(d, order, x2, n, ld2, x1, m, ld1, coef, G, ldg_col, ldg_row, nullptr, out2, s));
  return MGB_OK;
}
