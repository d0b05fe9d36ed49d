// Host-buffer entry points: the reference-facing calls (numpy / CPU tensors in, numpy out).
// Device memory, streams and host<->device copies are managed here; the compute goes through the same
// device entry points as the torch-side bindings.  Candidate rows are double-buffered: the copy stream
// uploads block b+1 and downloads results of block b-1 while the compute stream works on block b.
#include "mgb_common.cuh"
#include <vector>

namespace mgb {

struct DevBuf {
  void* p = nullptr;
  int alloc(size_t bytes) { return check_cuda(cudaMalloc(&p, bytes ? bytes : 16), "cudaMalloc"); }
  ~DevBuf() { if (p) cudaFree(p); }
  template <class T> T* as() { return static_cast<T*>(p); }
};

struct StreamPair {
  cudaStream_t compute = nullptr, copy = nullptr;
  int init() {
    MGB_TRY(check_cuda(cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking), "cudaStreamCreate"));
    return check_cuda(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking), "cudaStreamCreate");
  }
  ~StreamPair() {
    if (compute) cudaStreamDestroy(compute);
    if (copy) cudaStreamDestroy(copy);
  }
};

struct Events {
  std::vector<cudaEvent_t> ev;
  int init(int count) {
    ev.resize(count, nullptr);
    for (auto& e : ev) MGB_TRY(check_cuda(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate"));
    return MGB_OK;
  }
  ~Events() { for (auto e : ev) if (e) cudaEventDestroy(e); }
};

static int select_device(int device) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail(MGB_ERR_NO_DEVICE, "no CUDA device (libmgb has no CPU fallback)");
  if (device < 0 || device >= count) return fail(MGB_ERR_ARG, "device index out of range");
  return check_cuda(cudaSetDevice(device), "cudaSetDevice");
}

}  // namespace mgb

using namespace mgb;

extern "C" int mgb_series_gram_host(const mgb_series_desc* desc, const double* x1, long long m, const double* x2,
                                    long long n, const double* coef, double* K, int device) {
  if (!desc || !x1 || !x2 || !coef || !K || m < 0 || n < 0) return fail(MGB_ERR_ARG, "series_gram_host: bad argument");
  MGB_TRY(select_device(device));
  if (m == 0 || n == 0) return MGB_OK;
  const long long D = (long long)desc->n_factors * desc->factor_width;
  const long long T = (long long)desc->n_factors * desc->n_terms;
  StreamPair st;
  MGB_TRY(st.init());
  // row blocks sized to ~256 MB of output so that downloads overlap the next block's kernel
  long long rows_blk = (256LL << 20) / (8 * n);
  rows_blk = rows_blk < 128 ? 128 : rows_blk / 128 * 128;
  if (rows_blk > m) rows_blk = m;
  DevBuf d_x1, d_x2, d_coef, d_k[2];
  MGB_TRY(d_x1.alloc(sizeof(double) * m * D));
  MGB_TRY(d_x2.alloc(sizeof(double) * n * D));
  MGB_TRY(d_coef.alloc(sizeof(double) * T));
  const long long ldk = (n + 1) / 2 * 2;
  MGB_TRY(d_k[0].alloc(sizeof(double) * rows_blk * ldk));
  MGB_TRY(d_k[1].alloc(sizeof(double) * rows_blk * ldk));
  Events done, freed;
  MGB_TRY(done.init(2));
  MGB_TRY(freed.init(2));
  MGB_TRY(check_cuda(cudaMemcpyAsync(d_x1.p, x1, sizeof(double) * m * D, cudaMemcpyHostToDevice, st.compute), "H2D x1"));
  MGB_TRY(check_cuda(cudaMemcpyAsync(d_x2.p, x2, sizeof(double) * n * D, cudaMemcpyHostToDevice, st.compute), "H2D x2"));
  MGB_TRY(check_cuda(cudaMemcpyAsync(d_coef.p, coef, sizeof(double) * T, cudaMemcpyHostToDevice, st.compute), "H2D coef"));
  int b = 0;
  for (long long r0 = 0; r0 < m; r0 += rows_blk, ++b) {
    const long long rows = (m - r0 < rows_blk) ? (m - r0) : rows_blk;
    const int buf = b & 1;
    if (b >= 2) MGB_TRY(check_cuda(cudaStreamWaitEvent(st.compute, freed.ev[buf], 0), "cudaStreamWaitEvent"));
    MGB_TRY(mgb_series_gram_fwd(desc, d_x1.as<double>() + r0 * D, rows, D, d_x2.as<double>(), n, D, d_coef.as<double>(),
                                d_k[buf].as<double>(), ldk, st.compute));
    MGB_TRY(check_cuda(cudaEventRecord(done.ev[buf], st.compute), "cudaEventRecord"));
    MGB_TRY(check_cuda(cudaStreamWaitEvent(st.copy, done.ev[buf], 0), "cudaStreamWaitEvent"));
    MGB_TRY(check_cuda(cudaMemcpy2DAsync(K + r0 * n, sizeof(double) * n, d_k[buf].p, sizeof(double) * ldk, sizeof(double) * n,
                                         rows, cudaMemcpyDeviceToHost, st.copy), "D2H K"));
    MGB_TRY(check_cuda(cudaEventRecord(freed.ev[buf], st.copy), "cudaEventRecord"));
  }
  MGB_TRY(check_cuda(cudaStreamSynchronize(st.compute), "cudaStreamSynchronize"));
  return check_cuda(cudaStreamSynchronize(st.copy), "cudaStreamSynchronize");
}

extern "C" int mgb_series_posterior_host(const mgb_series_desc* desc, const double* xc, long long m,
                                         const double* xtrain, long long n, const double* coef, const double* Rinv,
                                         const double* alpha, double kss_value, double mean_const, double* mean,
                                         double* var, long long chunk, int device) {
  if (!desc || !xc || !xtrain || !coef || !Rinv || !alpha || !mean || !var || m < 0 || n < 1)
    return fail(MGB_ERR_ARG, "series_posterior_host: bad argument");
  MGB_TRY(select_device(device));
  if (m == 0) return MGB_OK;
  if (chunk <= 0) chunk = 37888;  // 2 waves of 128-candidate tiles on 148 SMs
  if (chunk > m) chunk = (m + 127) / 128 * 128;
  const long long D = (long long)desc->n_factors * desc->factor_width;
  const long long T = (long long)desc->n_factors * desc->n_terms;
  StreamPair st;
  MGB_TRY(st.init());
  DevBuf d_xt, d_coef, d_R, d_alpha, d_pack, d_xc[2], d_mean, d_var, d_ws;
  MGB_TRY(d_xt.alloc(sizeof(double) * n * D));
  MGB_TRY(d_coef.alloc(sizeof(double) * T));
  MGB_TRY(d_R.alloc(sizeof(double) * n * n));
  MGB_TRY(d_alpha.alloc(sizeof(double) * n));
  MGB_TRY(d_pack.alloc((size_t)mgb_posterior_pack_bytes(n)));
  MGB_TRY(d_xc[0].alloc(sizeof(double) * chunk * D));
  MGB_TRY(d_xc[1].alloc(sizeof(double) * chunk * D));
  MGB_TRY(d_mean.alloc(sizeof(double) * m));
  MGB_TRY(d_var.alloc(sizeof(double) * m));
  const long long ws_bytes = mgb_posterior_workspace_bytes(chunk, n);
  MGB_TRY(d_ws.alloc((size_t)ws_bytes));
  Events up, done;
  MGB_TRY(up.init(2));
  MGB_TRY(done.init(2));
  MGB_TRY(check_cuda(cudaMemcpyAsync(d_xt.p, xtrain, sizeof(double) * n * D, cudaMemcpyHostToDevice, st.compute), "H2D xtrain"));
  MGB_TRY(check_cuda(cudaMemcpyAsync(d_coef.p, coef, sizeof(double) * T, cudaMemcpyHostToDevice, st.compute), "H2D coef"));
  MGB_TRY(check_cuda(cudaMemcpyAsync(d_R.p, Rinv, sizeof(double) * n * n, cudaMemcpyHostToDevice, st.compute), "H2D Rinv"));
  MGB_TRY(check_cuda(cudaMemcpyAsync(d_alpha.p, alpha, sizeof(double) * n, cudaMemcpyHostToDevice, st.compute), "H2D alpha"));
  MGB_TRY(mgb_posterior_pack(d_R.as<double>(), n, n, d_alpha.as<double>(), d_pack.as<double>(), st.compute));
  const long long nblk = (m + chunk - 1) / chunk;
  auto upload = [&](long long blk) -> int {
    const long long c0 = blk * chunk;
    const long long rows = (m - c0 < chunk) ? (m - c0) : chunk;
    const int buf = (int)(blk & 1);
    // the buffer was last read by the kernels of block blk-2
    if (blk >= 2) MGB_TRY(check_cuda(cudaStreamWaitEvent(st.copy, done.ev[buf], 0), "cudaStreamWaitEvent"));
    MGB_TRY(check_cuda(cudaMemcpyAsync(d_xc[buf].p, xc + c0 * D, sizeof(double) * rows * D, cudaMemcpyHostToDevice, st.copy), "H2D xc"));
    return check_cuda(cudaEventRecord(up.ev[buf], st.copy), "cudaEventRecord");
  };
  MGB_TRY(upload(0));
  for (long long blk = 0; blk < nblk; ++blk) {
    const long long c0 = blk * chunk;
    const long long rows = (m - c0 < chunk) ? (m - c0) : chunk;
    const int buf = (int)(blk & 1);
    MGB_TRY(check_cuda(cudaStreamWaitEvent(st.compute, up.ev[buf], 0), "cudaStreamWaitEvent"));
    MGB_TRY(mgb_series_posterior(desc, d_xc[buf].as<double>(), rows, D, d_xt.as<double>(), n, D, d_coef.as<double>(),
                                 d_pack.as<double>(), kss_value, mean_const, d_mean.as<double>() + c0,
                                 d_var.as<double>() + c0, d_ws.p, ws_bytes, chunk, st.compute));
    MGB_TRY(check_cuda(cudaEventRecord(done.ev[buf], st.compute), "cudaEventRecord"));
    if (blk + 1 < nblk) MGB_TRY(upload(blk + 1));  // next block's upload overlaps this block's kernels
    MGB_TRY(check_cuda(cudaStreamWaitEvent(st.copy, done.ev[buf], 0), "cudaStreamWaitEvent"));
    MGB_TRY(check_cuda(cudaMemcpyAsync(mean + c0, d_mean.as<double>() + c0, sizeof(double) * rows, cudaMemcpyDeviceToHost, st.copy), "D2H mean"));
    MGB_TRY(check_cuda(cudaMemcpyAsync(var + c0, d_var.as<double>() + c0, sizeof(double) * rows, cudaMemcpyDeviceToHost, st.copy), "D2H var"));
  }
  MGB_TRY(check_cuda(cudaStreamSynchronize(st.compute), "cudaStreamSynchronize"));
  return check_cuda(cudaStreamSynchronize(st.copy), "cudaStreamSynchronize");
}

// ------------------------------------------------------------------------------------------------
// Trust-region step of the quadrature kernels as one call (mgb.h): the existing entry points enqueued back to back.
// ------------------------------------------------------------------------------------------------
extern "C" int mgb_radial_acquisition_ei_step(int kind, int width, const double* xc, long long b, long long ldc,
                                              const double* v, long long ldv, const double* xtrain, long long n,
                                              long long ldt, const double* tval, const double* tcoef, int Pt, double s0,
                                              double ds, int Ps, double outputscale, const double* rinv, long long ldr,
                                              const double* rinv_t, long long ldrt, const double* alpha, double kss,
                                              double mean_const, double best_f, int maximize, double* work,
                                              long long work_doubles, double* ei, double* grad, double* hvp,
                                              void* stream) {
  if (kind < 0 || kind > 3) return mgb::fail(MGB_ERR_ARG, "radial_acquisition_ei_step: kind must be 0..3");
  if (b < 0 || n < 1) return mgb::fail(MGB_ERR_ARG, "radial_acquisition_ei_step: bad sizes");
  if (b == 0) return MGB_OK;
  const long long bn = b * n;
  if (!work || work_doubles < 4 * bn) return mgb::fail(MGB_ERR_ARG, "radial_acquisition_ei_step: scratch too small");
  if (hvp && !v) return mgb::fail(MGB_ERR_ARG, "radial_acquisition_ei_step: hvp requested without directions");
  const int geom = (kind == 3) ? 1 : 0;
  const int w = (kind == 3) ? width : kind + 2;
  double* z = work;
  double* P[3] = {work + bn, work + 2 * bn, work + 3 * bn};
  MGB_TRY(mgb_pair_scalars(geom, w, xc, b, ldc, xtrain, n, ldt, z, nullptr, stream));
  MGB_TRY(mgb_radial_profile(kind, 3, z, bn, tval, tcoef, Pt, s0, ds, Ps, P[0], stream));    // all three orders, one launch
  return mgb_radial_acquisition_ei(geom, w, xc, b, ldc, v, ldv, xtrain, n, ldt, P[0], P[1], hvp ? P[2] : nullptr, outputscale, rinv, ldr,
                                   rinv_t, ldrt, alpha, kss, mean_const, best_f, maximize, ei, grad, hvp, stream);
}

extern "C" int mgb_spd2_acquisition_ei_step(int use_cholesky, const double* xc, long long b, long long ldc,
                                            const double* xtrain, long long n, long long ldt, const double* tcoef,
                                            const double* rscale, const double* bscale, int Pt, const double* bval,
                                            const double* bw, int Pb, double outputscale, const double* rinv,
                                            long long ldr, const double* rinv_t, long long ldrt, const double* alpha,
                                            double kss, double mean_const, double best_f, int maximize, double* work,
                                            long long work_doubles, double* ei, double* grad, void* stream) {
  if (b < 0 || n < 1) return mgb::fail(MGB_ERR_ARG, "spd2_acquisition_ei_step: bad sizes");
  if (b == 0) return MGB_OK;
  const long long bn = b * n;
  if (!work || work_doubles < 5 * bn) return mgb::fail(MGB_ERR_ARG, "spd2_acquisition_ei_step: scratch too small");
  double* delta = work;
  double* r2 = work + bn;
  double* P[3] = {work + 2 * bn, work + 3 * bn, work + 4 * bn};
  MGB_TRY(mgb_pair_scalars(2 + (use_cholesky ? 1 : 0), 3, xc, b, ldc, xtrain, n, ldt, delta, r2, stream));
  MGB_TRY(mgb_spd2_profile(3, delta, r2, bn, tcoef, rscale, bscale, Pt, bval, bw, Pb, P[0], stream));   // value, d/dDelta, d/dr2
  return mgb_spd2_acquisition_ei(use_cholesky, xc, b, ldc, xtrain, n, ldt, P[0], P[1], P[2], outputscale, rinv, ldr, rinv_t,
                                 ldrt, alpha, kss, mean_const, best_f, maximize, ei, grad, stream);
}
