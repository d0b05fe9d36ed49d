// Fused exact-GP posterior reduction for a block of candidates (see include/mgb.h):
//   V[c][i]  = sum_{j<=i} Kc[c][j] * Rinv[i][j]        (FP64 tensor-core DMMA m8n8k4 tiles)
//   var[c]   = kss[c] - sum_i V[c][i]^2                (squared and reduced in registers; V never stored)
//   mean[c]  = mean_const + sum_j Kc[c][j] * alpha[j]  (alpha rides along as row n of the packed factor)
//
// This is the FLOP-dominant stage of the acquisition posterior (n^2 per candidate):
// FP64-bound, so the design goal is to keep the DMMA pipe busy:
//   - CTA tile 128 candidates x 128 factor rows, 8 warps as 2 x 4, warp tile 64 x 32
//     (8 x 4 DMMA fragments, 64 accumulator doubles per lane);
//   - one CTA walks ALL row tiles of the packed factor for its 128 candidates (triangular: row tile t
//     only needs columns j < 128 (t+1)), so the square-sum stays in registers and is stored once;
//   - operands stream through a 4-stage cp.async pipeline of 128 x 16 double panels, stored with a
//     row pitch of 20 doubles so that every fragment load is bank-conflict free;
//   - the (tile, column-chunk) loop is flattened so the pipeline never drains between row tiles.
// The packed factor (mgb_posterior_pack) is the lower triangle of Rinv padded with zeros to
// (n+1 rounded up to 128) x (n rounded up to 16), row n holding alpha.
#include "mgb_common.cuh"

namespace mgb {

constexpr int PT = 128;       // tile edge (candidates and factor rows)
constexpr int PK = 16;        // columns per pipeline stage
constexpr int PLD = PK + 4;   // smem row pitch in doubles (bank-conflict-free fragment loads)
constexpr int PSTAGES = 4;
constexpr int PTHREADS = 256;
constexpr size_t kPostSmem = sizeof(double) * (size_t)PSTAGES * 2 * PT * PLD;

__host__ __device__ inline long long round_up(long long v, long long q) { return (v + q - 1) / q * q; }

__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

struct PostParams {
  const double* Kc; long long m, n, ldk;
  const double* Rp; long long rows_p, ldp;   // packed factor
  const double* kss;
  double mean_const;
  double* mean; double* var;
  int n_tiles;   // row tiles of the packed factor
  int splits;    // CTAs sharing one candidate tile (row tiles dealt round-robin)
};

__global__ void posterior_init_kernel(const double* kss, double* var, long long m) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) var[i] = kss[i];
}

__global__ void __launch_bounds__(PTHREADS, 1) posterior_reduce_kernel(PostParams p) {
  extern __shared__ __align__(16) double smem[];
  double* As = smem;                               // [stage][PT][PLD]
  double* Bs = smem + (size_t)PSTAGES * PT * PLD;  // [stage][PT][PLD]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wc = warp >> 2;   // 0..1 : candidate half
  const int wi = warp & 3;    // 0..3 : factor-row quarter
  const long long c0 = (long long)blockIdx.x * PT;
  const int split = blockIdx.y;

  // ---- flattened (row tile, column chunk) schedule for this CTA ------------------------------
  // tile t has chunks_t = min(128 (t+1), n_pad) / PK column chunks
  const long long n_pad = p.ldp;
  auto chunks_of = [&](int t) -> int {
    long long cols = (long long)(t + 1) * PT;
    if (cols > n_pad) cols = n_pad;
    return (int)(cols / PK);
  };

  // loader mapping: thread -> (row = tid / 2, half = tid & 1) : 8 doubles = four 16-byte cp.async
  const int lrow = tid >> 1, lhalf = tid & 1;
  const long long a_row = c0 + lrow;
  const bool a_valid = a_row < p.m;
  const double* a_src_row = p.Kc + (a_valid ? a_row : 0) * p.ldk;

  auto issue = [&](int stage, int t, int ch) {
    const long long j0 = (long long)ch * PK + lhalf * 8;
    double* ad = As + ((size_t)stage * PT + lrow) * PLD + lhalf * 8;
    double* bd = Bs + ((size_t)stage * PT + lrow) * PLD + lhalf * 8;
    const double* bs = p.Rp + ((long long)t * PT + lrow) * p.ldp + j0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long long j = j0 + 2 * q;
      int bytes = 0;
      if (a_valid) {
        long long rem = p.n - j;
        bytes = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
      }
      cp_async16(ad + 2 * q, a_src_row + (bytes ? j : 0), bytes);
      cp_async16(bd + 2 * q, bs + 2 * q, 16);
    }
  };

  // iteration state for the producer side
  int pt = split, pch = 0;
  bool p_done = pt >= p.n_tiles;
  auto advance_producer = [&]() {
    if (++pch >= chunks_of(pt)) {
      pch = 0;
      pt += p.splits;
      if (pt >= p.n_tiles) p_done = true;
    }
  };

#pragma unroll
  for (int s = 0; s < PSTAGES - 1; ++s) {
    if (!p_done) {
      issue(s, pt, pch);
      advance_producer();
    }
    cp_async_commit();
  }

  double acc0[8][4], acc1[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc0[a][b] = acc1[a][b] = 0.0;
  double sumsq[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) sumsq[a] = 0.0;
  double meanv[8];
#pragma unroll
  for (int a = 0; a < 8; ++a) meanv[a] = 0.0;

  const int frow = lane >> 2, fk = lane & 3;
  int stage = 0;
  // Factor rows are dealt to the four row-warps in interleaved groups of 8: fragment b of warp wi holds tile
  // rows 32 b + 8 wi + [0, 8).  Every warp therefore sees the same share of (a) the zero upper triangle of a
  // diagonal tile and (b) the zero padding rows of the last tile, and skips those DMMAs in lock step.
  for (int t = split; t < p.n_tiles; t += p.splits) {
    const int nch = chunks_of(t);
    const long long row_base = (long long)t * PT + 8 * wi;   // first row of fragment 0
    // fragments whose first row is beyond row n (alpha) are all zero
    int b_hi = 4;
    while (b_hi > 0 && row_base + 32 * (b_hi - 1) > p.n) --b_hi;
    for (int ch = 0; ch < nch; ++ch) {
      cp_async_wait<PSTAGES - 2>();
      __syncthreads();
      {
        const int fill = (stage + PSTAGES - 1) % PSTAGES;
        if (!p_done) {
          issue(fill, pt, pch);
          advance_producer();
        }
        cp_async_commit();
      }
      // fragment b is needed for this column chunk iff its last row >= first column of the chunk
      const long long j0 = (long long)ch * PK;
      int b_lo = 0;
      while (b_lo < b_hi && row_base + 32 * b_lo + 7 < j0) ++b_lo;
      const double* at = As + ((size_t)stage * PT + wc * 64 + frow) * PLD + fk;
      const double* bt = Bs + ((size_t)stage * PT + 8 * wi + frow) * PLD + fk;
      if (b_lo == 0 && b_hi == 4) {
#pragma unroll
        for (int k4 = 0; k4 < PK / 4; ++k4) {
          double af[8], bf[4];
#pragma unroll
          for (int a = 0; a < 8; ++a) af[a] = at[(a * 8) * PLD + k4 * 4];
#pragma unroll
          for (int b = 0; b < 4; ++b) bf[b] = bt[(b * 32) * PLD + k4 * 4];
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
          for (int a = 0; a < 8; ++a) af[a] = at[(a * 8) * PLD + k4 * 4];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (b >= b_lo && b < b_hi) {
              const double bfv = bt[(b * 32) * PLD + k4 * 4];
#pragma unroll
              for (int a = 0; a < 8; ++a) dmma(acc0[a][b], acc1[a][b], af[a], bfv);
            }
          }
        }
      }
      stage = (stage + 1) % PSTAGES;
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
  cp_async_wait<0>();
  __syncthreads();

  // ---- CTA reduction: 4 lanes per fragment row, then the 4 row-quarter warps ------------------
  double* red = smem;            // [4][PT] sums, then [PT] means
  double* mred = smem + 4 * PT;
  for (int e = tid; e < PT; e += PTHREADS) mred[e] = 0.0;
  __syncthreads();
  const long long alpha_tile = p.n / PT;  // row tile holding row n
  const bool owns_mean = (alpha_tile % p.splits) == split;
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
  __syncthreads();
  for (int e = tid; e < PT; e += PTHREADS) {
    const long long c = c0 + e;
    if (c < p.m) {
      const double tot = (red[e] + red[PT + e]) + (red[2 * PT + e] + red[3 * PT + e]);
      if (p.splits == 1) {
        p.var[c] = p.kss[c] - tot;
      } else {
        atomicAdd(p.var + c, -tot);
      }
      if (owns_mean) p.mean[c] = p.mean_const + mred[e];
    }
  }
}

// Pack Rinv (lower triangle) + alpha into the padded operand.
__global__ void posterior_pack_kernel(const double* Rinv, long long n, long long ldr, const double* alpha,
                                      double* Rp, long long rows_p, long long ldp) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long i = blockIdx.y;
  if (j >= ldp) return;
  double v = 0.0;
  if (i < n && j <= i) v = Rinv[i * ldr + j];
  else if (i == n && j < n) v = alpha[j];
  Rp[i * ldp + j] = v;
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

extern "C" int mgb_posterior_reduce(const double* Kc, long long m, long long n, long long ldk, const double* packed,
                                    const double* kss, double mean_const, double* mean, double* var, void* stream) {
  if (!Kc || !packed || !kss || !mean || !var) return fail(MGB_ERR_ARG, "posterior_reduce: null pointer");
  if (m < 0 || n < 1 || ldk < n) return fail(MGB_ERR_ARG, "posterior_reduce: bad sizes");
  if ((ldk & 1) || (reinterpret_cast<uintptr_t>(Kc) & 15u)) return fail(MGB_ERR_ARG, "posterior_reduce: Kc must be 16-byte aligned with even ldk");
  if (m == 0) return MGB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    MGB_TRY(check_cuda(cudaGetDevice(&dev), "cudaGetDevice"));
    MGB_TRY(check_cuda(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute"));
    MGB_TRY(check_cuda(cudaFuncSetAttribute(posterior_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)kPostSmem), "cudaFuncSetAttribute"));
  }
  PostParams p;
  p.Kc = Kc; p.m = m; p.n = n; p.ldk = ldk;
  p.Rp = packed; p.rows_p = round_up(n + 1, PT); p.ldp = round_up(n, PK);
  p.kss = kss; p.mean_const = mean_const; p.mean = mean; p.var = var;
  p.n_tiles = (int)(p.rows_p / PT);
  const long long ctiles = (m + PT - 1) / PT;
  int splits = 1;
  if (ctiles < 2LL * sm_count) {
    splits = (int)((2LL * sm_count + ctiles - 1) / ctiles);
    if (splits > p.n_tiles) splits = p.n_tiles;
  }
  p.splits = splits;
  if (ctiles > 0x7fffffffLL) return fail(MGB_ERR_ARG, "posterior_reduce: m too large for one launch");
  if (splits > 1) {
    posterior_init_kernel<<<(unsigned)((m + 255) / 256), 256, 0, s>>>(kss, var, m);
    MGB_TRY(after_launch("posterior_init_kernel"));
  }
  dim3 grid((unsigned)ctiles, (unsigned)splits);
  posterior_reduce_kernel<<<grid, PTHREADS, kPostSmem, s>>>(p);
  return after_launch("posterior_reduce_kernel");
}

extern "C" long long mgb_posterior_workspace_bytes(long long chunk, long long n) {
  if (chunk < 0 || n < 0) return 0;
  // Kc chunk (ld = n rounded up to 16) + kss chunk
  return (chunk * round_up(n, PK) + chunk) * (long long)sizeof(double);
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
  const long long ldk = round_up(n, PK);
  double* Kc = static_cast<double*>(workspace);
  double* kss = Kc + chunk * ldk;
  fill_kernel<<<(unsigned)((chunk + 255) / 256), 256, 0, s>>>(kss, kss_value, chunk);
  MGB_TRY(after_launch("fill_kernel"));
  for (long long c0 = 0; c0 < m; c0 += chunk) {
    const long long mc = (m - c0 < chunk) ? (m - c0) : chunk;
    MGB_TRY(mgb_series_gram_fwd(desc, xc + c0 * ldc, mc, ldc, xtrain, n, ldt, coef, Kc, ldk, stream));
    MGB_TRY(mgb_posterior_reduce(Kc, mc, n, ldk, packed, kss, mean_const, mean + c0, var + c0, stream));
  }
  return MGB_OK;
}
