// Fused exact-GP posterior reduction for a block of candidates (see include/mgb.h):
//   V[c][i]  = sum_{j<=i} Kc[c][j] * Rinv[i][j]        (FP64 tensor-core DMMA m8n8k4 tiles)
//   var[c]   = kss[c] - sum_i V[c][i]^2                (squared and reduced in registers; V never stored)
//   mean[c]  = mean_const + sum_j Kc[c][j] * alpha[j]  (alpha rides along as row n of the packed factor)
//
// This is the FLOP-dominant stage of the acquisition posterior (n^2 per candidate):
// FP64-bound, so the design goal is to keep the DMMA (FP64 tensor) pipe busy:
//   - CTA tile 128 candidates x 128 factor rows, 8 DMMA warps as 2 x 4, warp tile 64 x 32
//     (8 x 4 m8n8k4 fragments, 64 accumulator doubles per lane) + 1 producer warp;
//   - the producer streams 128 x 16 double operand blocks with cp.async.bulk (TMA engine, ONE 16 KB copy per operand
//     per stage: the blocks are stored tile-major and XOR-swizzled in HBM, completion counted in bytes on a per-stage
//     `full` mbarrier) into a 6-stage ring, so that every fragment load is bank-conflict free; consumers release
//     stages through `empty` mbarriers - there is no CTA-wide barrier in the main loop; fragment loads are software
//     pipelined across stage boundaries;
//   - a CTA walks a balanced subset of the row tiles of the packed factor for its 128 candidates
//     (triangular: row tile t only needs columns j < 128 (t+1)); square-sums stay in registers;
//   - the CTAs that share a candidate tile are adjacent in launch order and share its Gram panel through L2
//     (otherwise each panel would be re-streamed from HBM once per row tile); their partial sums are combined
//     in a fixed order by posterior_finalize_kernel (deterministic, no atomics).
// The packed factor (mgb_posterior_pack) is the lower triangle of Rinv padded with zeros to
// (n+1 rounded up to 128) x (n rounded up to 16), row n holding alpha.
#include "mgb_common.cuh"
#include <cmath>
#include <cstdlib>

namespace mgb {

constexpr int PT = 128;       // tile edge (candidates and factor rows)
constexpr int PK = 16;        // columns per pipeline stage
constexpr int PBLK = PT * PK;  // doubles per 128 x 16 operand block (16 KB), tile-major + swizzled in HBM
constexpr int PSTAGES = 6;
constexpr int PCONSUMERS = 8;                       // DMMA warps
constexpr int PTHREADS = (PCONSUMERS + 4) * 32;     // + one producer warpgroup (one warp of it issues copies)
constexpr int PREGS_PRODUCER = 40;                  // setmaxnreg budgets: 4*32*40 + 8*32*232 = 64512 <= 65536
constexpr int PREGS_CONSUMER = 232;
constexpr size_t kPostSmem = sizeof(double) * (size_t)PSTAGES * 2 * PBLK + 2 * PSTAGES * sizeof(uint64_t);
constexpr uint32_t kBlockBytes = PBLK * sizeof(double);

__host__ __device__ inline long long round_up(long long v, long long q) { return (v + q - 1) / q * q; }

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_test(bar, parity)) {
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct PostParams {
  const double* Kb; long long m, n;           // Gram chunk, tiled layout [ctile][chunk][128][16] (swizzled)
  const double* Rp; long long rows_p, ldp;   // packed factor, tiled layout [row tile][chunk][128][16] (swizzled)
  const double* kss;
  double mean_const;
  double* mean; double* var;
  double* partial;   // [splits][m] square-sum partials when splits > 1
  int n_tiles;       // row tiles of the packed factor
  int splits;        // CTAs sharing one candidate tile (they share its Gram panel through L2)
};

// Row tile handled in round j by split r of S: tiles sorted by work (descending = index descending) and
// dealt in snake order so that every split gets the same amount of work to within one tile.
__device__ __forceinline__ int tile_of(int j, int r, int S, int n_tiles) {
  const int idx = j * S + ((j & 1) ? (S - 1 - r) : r);
  return (idx < n_tiles) ? (n_tiles - 1 - idx) : -1;
}

// Warp-specialised: warp 8 streams 128 x 16 operand panels with cp.async.bulk (TMA engine) into a 6-stage
// ring, signalling per-stage `full` mbarriers by transaction bytes; warps 0-7 wait, issue DMMAs and release the
// stage through `empty` mbarriers.  No CTA-wide barrier in the main loop: warps may drift by up to 4 stages.
__global__ void __launch_bounds__(PTHREADS, 1) posterior_reduce_kernel(PostParams p) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                            // [stage][PT][PK]
  double* Bs = smem + (size_t)PSTAGES * PBLK;   // [stage][PT][PK]
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)PSTAGES * 2 * PBLK);
  uint64_t* empty = full + PSTAGES;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int S = p.splits;
  const long long ctile = blockIdx.x / S;
  const int split = (int)(blockIdx.x % S);
  const long long c0 = ctile * PT;
  const long long n_pad = p.ldp;
  const int rounds = (p.n_tiles + S - 1) / S;
  auto chunks_of = [&](int t) -> int {
    long long cols = (long long)(t + 1) * PT;
    if (cols > n_pad) cols = n_pad;
    return (int)(cols / PK);
  };

  if (tid == 0) {
    for (int s = 0; s < PSTAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, PCONSUMERS);
    }
    fence_mbar_init();
  }
  __syncthreads();

  if (warp >= PCONSUMERS) {
    // ===================== producer warpgroup =====================
    // hands its registers to the DMMA warps (168 per thread at launch: 384 threads), one warp issues copies
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PREGS_PRODUCER));
    if (warp != PCONSUMERS) return;
    int stage = 0;
    uint32_t phase = 0;
    const long long nch_all = n_pad / PK;
    const double* a_blocks = p.Kb + ctile * nch_all * PBLK;
    if (lane == 0) {
      for (int j = 0; j < rounds; ++j) {
        const int t = tile_of(j, split, S, p.n_tiles);
        if (t < 0) continue;
        const int nch = chunks_of(t);
        const double* b_blocks = p.Rp + (long long)t * nch_all * PBLK;
        for (int ch = 0; ch < nch; ++ch) {
          mbar_spin(empty + stage, phase ^ 1u);
          mbar_expect_tx(full + stage, 2u * kBlockBytes);
          bulk_g2s(As + (size_t)stage * PBLK, a_blocks + (long long)ch * PBLK, kBlockBytes, full + stage);
          bulk_g2s(Bs + (size_t)stage * PBLK, b_blocks + (long long)ch * PBLK, kBlockBytes, full + stage);
          if (++stage == PSTAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    return;
  }

  // ===================== consumer warps =====================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PREGS_CONSUMER));
  const int wc = warp >> 2;   // 0..1 : candidate half
  const int wi = warp & 3;    // 0..3 : interleaved factor-row group
  double acc0[8][4], acc1[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc0[a][b] = acc1[a][b] = 0.0;
  double sumsq[8], meanv[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) sumsq[a] = meanv[a] = 0.0;

  const int frow = lane >> 2, fk = lane & 3;
  // element (row, col = 4 k4 + fk) of a block lives at row*16 + ((chunk ^ 2 (row & 3)) * 2 + (col & 1)), chunk = col / 2;
  // all fragment rows of this lane are congruent to frow modulo 4
  int koff[PK / 4];
#pragma unroll
  for (int k4 = 0; k4 < PK / 4; ++k4) koff[k4] = (((2 * k4 + (fk >> 1)) ^ ((frow & 3) << 1)) << 1) | (fk & 1);
  int stage = 0;
  uint32_t phase = 0;
  bool owns_mean = false;
  // Factor rows are dealt to the four row-warps in interleaved groups of 8: fragment b of warp wi holds tile
  // rows 32 b + 8 wi + [0, 8).  Every warp therefore sees the same share of (a) the zero upper triangle of a
  // diagonal tile and (b) the zero padding rows of the last tile, and skips those DMMAs in lock step.
  for (int j = 0; j < rounds; ++j) {
    const int t = tile_of(j, split, S, p.n_tiles);
    if (t < 0) continue;
    const int nch = chunks_of(t);
    const long long row_base = (long long)t * PT + 8 * wi;   // first row of fragment 0
    if ((long long)t * PT <= p.n && p.n < (long long)(t + 1) * PT) owns_mean = true;
    int b_hi = 4;                                            // fragments starting beyond row n are all zero
    while (b_hi > 0 && row_base + 32 * (b_hi - 1) > p.n) --b_hi;
    // Column chunks [0, n_full) need all four fragments of this warp (they lie left of the diagonal block); they run
    // software-pipelined ACROSS stage boundaries: the fragments of the next stage's first k-step are loaded before the
    // last 32 DMMAs of the current stage are issued, so the barrier check, address arithmetic and shared-memory
    // latency of a stage change hide behind tensor work instead of idling the pipe (both warps of an SM sub-partition
    // change stage at the same time).
    int n_full = 0;
    if (b_hi == 4) {
      // chunk ch is full iff row_base + 7 >= ch * PK  (b_lo == 0)
      const long long lim = (row_base + 7) / PK + 1;
      n_full = (int)(lim < nch ? lim : nch);
    }
    int ch = 0;
    if (n_full > 0) {
      double af[8], bf[4];
      mbar_spin(full + stage, phase);
      {
        const double* at = As + ((size_t)stage * PT + wc * 64 + frow) * PK;
        const double* bt = Bs + ((size_t)stage * PT + 8 * wi + frow) * PK;
#pragma unroll
        for (int a = 0; a < 8; ++a) af[a] = at[(a * 8) * PK + koff[0]];
#pragma unroll
        for (int b = 0; b < 4; ++b) bf[b] = bt[(b * 32) * PK + koff[0]];
      }
      for (; ch < n_full; ++ch) {
        const double* at = As + ((size_t)stage * PT + wc * 64 + frow) * PK;
        const double* bt = Bs + ((size_t)stage * PT + 8 * wi + frow) * PK;
        int nstage = stage + 1;
        uint32_t nphase = phase;
        if (nstage == PSTAGES) {
          nstage = 0;
          nphase ^= 1u;
        }
#pragma unroll
        for (int k4 = 0; k4 < PK / 4; ++k4) {
          double an[8], bn[4];
          if (k4 + 1 < PK / 4) {
#pragma unroll
            for (int a = 0; a < 8; ++a) an[a] = at[(a * 8) * PK + koff[(k4 + 1) % (PK / 4)]];
#pragma unroll
            for (int b = 0; b < 4; ++b) bn[b] = bt[(b * 32) * PK + koff[(k4 + 1) % (PK / 4)]];
          } else if (ch + 1 < n_full) {
            mbar_spin(full + nstage, nphase);      // the producer runs several stages ahead: normally already complete
            const double* atn = As + ((size_t)nstage * PT + wc * 64 + frow) * PK;
            const double* btn = Bs + ((size_t)nstage * PT + 8 * wi + frow) * PK;
#pragma unroll
            for (int a = 0; a < 8; ++a) an[a] = atn[(a * 8) * PK + koff[0]];
#pragma unroll
            for (int b = 0; b < 4; ++b) bn[b] = btn[(b * 32) * PK + koff[0]];
          } else {
#pragma unroll
            for (int a = 0; a < 8; ++a) an[a] = 0.0;
#pragma unroll
            for (int b = 0; b < 4; ++b) bn[b] = 0.0;
          }
#pragma unroll
          for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma(acc0[a][b], acc1[a][b], af[a], bf[b]);
#pragma unroll
          for (int a = 0; a < 8; ++a) af[a] = an[a];
#pragma unroll
          for (int b = 0; b < 4; ++b) bf[b] = bn[b];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + stage);   // every read of this stage has been consumed by a DMMA
        stage = nstage;
        phase = nphase;
      }
    }
    for (; ch < nch; ++ch) {
      // fragment b is needed for this column chunk iff its last row >= first column of the chunk
      const long long j0 = (long long)ch * PK;
      int b_lo = 0;
      while (b_lo < b_hi && row_base + 32 * b_lo + 7 < j0) ++b_lo;
      mbar_spin(full + stage, phase);
      const double* at = As + ((size_t)stage * PT + wc * 64 + frow) * PK;
      const double* bt = Bs + ((size_t)stage * PT + 8 * wi + frow) * PK;
      if (b_lo == 0 && b_hi == 4) {
#pragma unroll
        for (int k4 = 0; k4 < PK / 4; ++k4) {
          double af[8], bf[4];
#pragma unroll
          for (int a = 0; a < 8; ++a) af[a] = at[(a * 8) * PK + koff[k4]];
#pragma unroll
          for (int b = 0; b < 4; ++b) bf[b] = bt[(b * 32) * PK + koff[k4]];
#pragma unroll
          for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma(acc0[a][b], acc1[a][b], af[a], bf[b]);
        }
      } else if (b_lo < b_hi) {
#pragma unroll
        for (int k4 = 0; k4 < PK / 4; ++k4) {
          double af[8];
#pragma unroll
          for (int a = 0; a < 8; ++a) af[a] = at[(a * 8) * PK + koff[k4]];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (b >= b_lo && b < b_hi) {
              const double bfv = bt[(b * 32) * PK + koff[k4]];
#pragma unroll
              for (int a = 0; a < 8; ++a) dmma(acc0[a][b], acc1[a][b], af[a], bfv);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + stage);   // this warp is done reading the stage
      if (++stage == PSTAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
    // ---- tile epilogue: square-sum (rows i < n), mean (row i == n) -------------------------
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const long long i0 = row_base + 32 * b + 2 * fk;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const double v0 = acc0[a][b], v1 = acc1[a][b];
        if (i0 == p.n) meanv[a] = v0; else sumsq[a] = fma(v0, v0, sumsq[a]);
        if (i0 + 1 == p.n) meanv[a] = v1; else sumsq[a] = fma(v1, v1, sumsq[a]);
        acc0[a][b] = 0.0;
        acc1[a][b] = 0.0;
      }
    }
  }
  // all stages this CTA ever requested have been consumed by every warp that reaches this barrier
  asm volatile("bar.sync 1, %0;" ::"n"(PCONSUMERS * 32) : "memory");

  // ---- CTA reduction: 4 lanes per fragment row, then the 4 row-group warps -------------------------
  double* red = smem;            // [4][PT] sums, then [PT] means
  double* mred = smem + 4 * PT;
  const int mean_wi = (int)((p.n % 32) / 8), mean_fk = (int)((p.n % 8) / 2);
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    double s = sumsq[a];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    const int cl = wc * 64 + a * 8 + frow;
    if (fk == 0) red[wi * PT + cl] = s;
    if (owns_mean && wi == mean_wi && fk == mean_fk) mred[cl] = meanv[a];
  }
  asm volatile("bar.sync 1, %0;" ::"n"(PCONSUMERS * 32) : "memory");
  if (tid < PT) {
    const long long c = c0 + tid;
    if (c < p.m) {
      const double tot = (red[tid] + red[PT + tid]) + (red[2 * PT + tid] + red[3 * PT + tid]);
      if (S == 1) p.var[c] = p.kss[c] - tot;
      else p.partial[(long long)split * p.m + c] = tot;
      if (owns_mean) p.mean[c] = p.mean_const + mred[tid];
    }
  }
}

// var[c] = kss[c] - sum_s partial[s][c], fixed summation order (deterministic)
__global__ void posterior_finalize_kernel(const double* kss, const double* partial, int splits, long long m, double* var) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= m) return;
  double tot = 0.0;
  for (int s = 0; s < splits; ++s) tot += partial[(long long)s * m + c];
  var[c] = kss[c] - tot;
}

// Pack Rinv (lower triangle) + alpha into the padded, tile-major, swizzled operand.
__global__ void posterior_pack_kernel(const double* Rinv, long long n, long long ldr, const double* alpha,
                                      double* Rp, long long rows_p, long long ldp) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j >= ldp) return;
  double v = 0.0;
  if (i < n && j <= i) v = Rinv[i * ldr + j];
  else if (i == n && j < n) v = alpha[j];
  const long long t = i >> 7, ch = j >> 4;
  const int r = (int)(i & 127), chunk = (int)(j & 15) >> 1;
  Rp[((((t * (ldp >> 4) + ch) << 7) + r) << 4) + (((chunk ^ ((r & 3) << 1)) << 1) | (int)(j & 1))] = v;
}

// Row-major K (m x n, ld = ldk) -> the tile-major swizzled operand layout (for Gram matrices produced by the
// quadrature kernels, which write row-major).  One thread per 16-byte pair.
__global__ void posterior_retile_kernel(const double* K, long long m, long long n, long long ldk, double* Kb, long long nch) {
  const long long pair = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pairs_per_row = (n + 1) / 2;
  if (pair >= m * pairs_per_row) return;
  const long long row = pair / pairs_per_row, col = 2 * (pair - row * pairs_per_row);
  const long long ctile = row >> 7, ch = col >> 4;
  const int r = (int)(row & 127), chunk = (int)(col & 15) >> 1;
  double* dst = Kb + ((((ctile * nch + ch) << 7) + r) << 4) + ((chunk ^ ((r & 3) << 1)) << 1);
  dst[0] = K[row * ldk + col];
  dst[1] = (col + 1 < n) ? K[row * ldk + col + 1] : 0.0;
}

__global__ void fill_kernel(double* dst, double v, long long count) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = v;
}

}  // namespace mgb

using namespace mgb;

extern "C" long long mgb_posterior_pack_bytes(long long n) {
  if (n < 0) return 0;
  return round_up(n + 1, PT) * round_up(n, PK) * (long long)sizeof(double);
}

extern "C" int mgb_posterior_pack(const double* Rinv, long long n, long long ldr, const double* alpha,
                                  double* packed, void* stream) {
  if (!Rinv || !alpha || !packed || n < 1 || ldr < n) return fail(MGB_ERR_ARG, "posterior_pack: bad argument");
  const long long rows_p = round_up(n + 1, PT), ldp = round_up(n, PK);
  if (rows_p > 65535) return fail(MGB_ERR_ARG, "posterior_pack: n too large");
  dim3 grid((unsigned)((ldp + 255) / 256), (unsigned)rows_p);
  posterior_pack_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(Rinv, n, ldr, alpha, packed, rows_p, ldp);
  return after_launch("posterior_pack_kernel");
}

static int post_sm_count() { return device_sm_count(); }

// How many CTAs share one 128-candidate tile.  (a) Few candidates: split so that the grid fills the SMs.
// (b) Many: the 128 x n Gram panel of a candidate tile is re-read once per row tile; CTAs working on the SAME
// panel at the same time share it through L2, so enough splits are used that the panels of all concurrently
// resident CTAs fit in a ~40 MB slice of the 126 MB L2 (the factor streams through the rest).
static double post_l2_mb() {   // MGB_POST_L2MB: tuning knob for the L2 slice the in-flight Gram panels may occupy
  static double l2_mb = 0.0;
  if (l2_mb == 0.0) {
    const char* e = getenv("MGB_POST_L2MB");
    l2_mb = (e != nullptr && atof(e) > 0.0) ? atof(e) : 40.0;
  }
  return l2_mb;
}
static int choose_splits(long long m, long long n) {
  const int sms = post_sm_count();
  const long long ctiles = (m + PT - 1) / PT;
  const int n_tiles = (int)(round_up(n + 1, PT) / PT);
  const double panel_bytes = (double)PT * (double)round_up(n, PK) * 8.0;  // one candidate tile's Gram panel
  long long s_l2 = (long long)ceil(sms * panel_bytes / (post_l2_mb() * 1024 * 1024));
  long long s_fill = (2LL * sms + ctiles - 1) / ctiles;
  long long s = s_l2 > s_fill ? s_l2 : s_fill;
  if (s < 1) s = 1;
  if (s > n_tiles) s = n_tiles;
  return (int)s;
}

extern "C" int mgb_posterior_retile(const double* K, long long m, long long n, long long ldk, double* Kb, void* stream) {
  if (!K || !Kb || m < 0 || n < 1 || ldk < n) return fail(MGB_ERR_ARG, "posterior_retile: bad argument");
  if (reinterpret_cast<uintptr_t>(Kb) & 15u) return fail(MGB_ERR_ARG, "posterior_retile: Kb must be 16-byte aligned");
  if (m == 0) return MGB_OK;
  const long long pairs = m * ((n + 1) / 2);
  posterior_retile_kernel<<<(unsigned)((pairs + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      K, m, n, ldk, Kb, round_up(n, PK) / PK);
  return after_launch("posterior_retile_kernel");
}

// Upper bound of splits(mc, n) * mc over every block size mc <= m: a workspace sized for `m` candidates also serves the
// shorter tail block of a chunked evaluation (fewer candidate tiles -> MORE splits per tile).  splits = max(s_l2,
// s_fill(mc)) clipped to n_tiles; s_l2 does not depend on mc and s_fill(mc) * mc <= (2 sms + ctiles(mc)) * 128.
extern "C" long long mgb_posterior_reduce_workspace_bytes(long long m, long long n) {
  if (m <= 0 || n <= 0) return 0;
  const long long m_pad = round_up(m, PT), ctiles = m_pad / PT;
  const long long n_tiles = round_up(n + 1, PT) / PT;
  const double panel_bytes = (double)PT * (double)round_up(n, PK) * 8.0;
  const long long s_l2 = (long long)ceil(post_sm_count() * panel_bytes / (post_l2_mb() * 1024 * 1024));
  long long doubles = (s_l2 > 1 ? s_l2 : 1) * m_pad + (2LL * post_sm_count() + ctiles) * PT;
  if (doubles > n_tiles * m_pad) doubles = n_tiles * m_pad;
  return doubles * (long long)sizeof(double);
}

extern "C" int mgb_posterior_reduce(const double* Kb, long long m, long long n, const double* packed,
                                    const double* kss, double mean_const, double* mean, double* var, void* workspace,
                                    long long workspace_bytes, void* stream) {
  if (!Kb || !packed || !kss || !mean || !var) return fail(MGB_ERR_ARG, "posterior_reduce: null pointer");
  if (m < 0 || n < 1) return fail(MGB_ERR_ARG, "posterior_reduce: bad sizes");
  if (reinterpret_cast<uintptr_t>(Kb) & 15u) return fail(MGB_ERR_ARG, "posterior_reduce: Kb must be 16-byte aligned");
  if (m == 0) return MGB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // per device, not per process: set on every call (cheap, legal during stream capture)
  MGB_TRY(check_cuda(cudaFuncSetAttribute(posterior_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)kPostSmem), "cudaFuncSetAttribute"));
  PostParams p;
  p.Kb = Kb; p.m = m; p.n = n;
  p.Rp = packed; p.rows_p = round_up(n + 1, PT); p.ldp = round_up(n, PK);
  p.kss = kss; p.mean_const = mean_const; p.mean = mean; p.var = var;
  p.n_tiles = (int)(p.rows_p / PT);
  p.splits = choose_splits(m, n);
  p.partial = nullptr;
  if (p.splits > 1) {
    if (!workspace || workspace_bytes < (long long)p.splits * m * (long long)sizeof(double))
      return fail(MGB_ERR_ARG, "posterior_reduce: workspace smaller than mgb_posterior_reduce_workspace_bytes(m, n)");
    p.partial = static_cast<double*>(workspace);
  }
  const long long ctiles = (m + PT - 1) / PT;
  if (ctiles * p.splits > 0x7fffffffLL) return fail(MGB_ERR_ARG, "posterior_reduce: m too large for one launch");
  posterior_reduce_kernel<<<(unsigned)(ctiles * p.splits), PTHREADS, kPostSmem, s>>>(p);
  MGB_TRY(after_launch("posterior_reduce_kernel"));
  if (p.splits > 1) {
    posterior_finalize_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(kss, p.partial, p.splits, m, var);
    MGB_TRY(after_launch("posterior_finalize_kernel"));
  }
  return MGB_OK;
}

extern "C" long long mgb_posterior_workspace_bytes(long long chunk, long long n) {
  if (chunk < 0 || n < 0) return 0;
  // tiled Gram chunk + self-kernel values + split partials
  return round_up(chunk, PT) * round_up(n, PK) * (long long)sizeof(double) + round_up(chunk, 2) * (long long)sizeof(double) +
         mgb_posterior_reduce_workspace_bytes(chunk, n);
}

extern "C" int mgb_series_posterior(const mgb_series_desc* desc, const double* xc, long long m, long long ldc,
                                    const double* xtrain, long long n, long long ldt, const double* coef,
                                    const double* packed, double kss_value, double mean_const, double* mean,
                                    double* var, void* workspace, long long workspace_bytes, long long chunk,
                                    void* stream) {
  if (!workspace || chunk < 1) return fail(MGB_ERR_ARG, "series_posterior: bad workspace / chunk");
  if (workspace_bytes < mgb_posterior_workspace_bytes(chunk, n)) return fail(MGB_ERR_ARG, "series_posterior: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 15u) return fail(MGB_ERR_ARG, "series_posterior: workspace must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* Kb = static_cast<double*>(workspace);
  const long long kb_doubles = round_up(chunk, PT) * round_up(n, PK);
  double* kss = Kb + kb_doubles;
  double* part = kss + round_up(chunk, 2);
  const long long part_bytes = mgb_posterior_reduce_workspace_bytes(chunk, n);
  // Rows past the last candidate of a tile and the padding columns [n, n rounded up to 16) of the tiled chunk are
  // read by the bulk copies (and multiplied by zeros of the factor / never stored): they must be finite, so the
  // part of the buffer this call uses is cleared once.
  const long long used_rows = round_up(chunk < m ? chunk : m, PT);
  MGB_TRY(check_cuda(cudaMemsetAsync(Kb, 0, sizeof(double) * (size_t)used_rows * round_up(n, PK), s), "cudaMemsetAsync"));
  fill_kernel<<<(unsigned)((chunk + 255) / 256), 256, 0, s>>>(kss, kss_value, chunk);
  MGB_TRY(after_launch("fill_kernel"));
  for (long long c0 = 0; c0 < m; c0 += chunk) {
    const long long mc = (m - c0 < chunk) ? (m - c0) : chunk;
    MGB_TRY(mgb_series_gram_tiled(desc, xc + c0 * ldc, mc, ldc, xtrain, n, ldt, coef, Kb, stream));
    MGB_TRY(mgb_posterior_reduce(Kb, mc, n, packed, kss, mean_const, mean + c0, var + c0, part, part_bytes, stream));
  }
  return MGB_OK;
}
