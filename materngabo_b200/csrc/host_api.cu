// Host-buffer entry points: the reference-facing calls (numpy / CPU tensors in, numpy out).
// Device memory, streams and host<->device copies are managed here; the compute goes through the same
// device entry points as the torch-side bindings.  Candidate rows are double-buffered: the copy stream
// uploads block b+1 and downloads results of block b-1 while the compute stream works on block b.
#include "mgb_common.cuh"
#include <cstdlib>
#include <cstring>
#include <new>
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
// Persistent posterior handle (mgb.h): the fitted state (training points, coefficients, packed factor) is uploaded and
// packed ONCE; every evaluation reuses the device buffers, workspace, streams and events of the handle.  This is the
// call pattern of the reference: one fit per BO iteration, then hundreds of acq(X) calls (manifold_optimize.py:229-236,
// 329-339).
// ------------------------------------------------------------------------------------------------
struct mgb_posterior {
  int device = 0;
  mgb_series_desc desc{};
  long long n = 0, D = 0, T = 0, chunk = 0;
  double kss = 0.0, mean_const = 0.0;
  bool owns_state = false;
  double *d_xt = nullptr, *d_coef = nullptr, *d_pack = nullptr;   // owned (create) or borrowed (create_from_device)
  long long ldt = 0;
  double* d_xc[2] = {nullptr, nullptr};
  double* d_out[2] = {nullptr, nullptr};    // [mean chunk | var chunk]
  void* d_ws = nullptr;
  long long ws_bytes = 0;
  double* h_stage = nullptr;                // pinned: small blocks from pageable user memory go through it
  long long stage_rows = 0;
  cudaStream_t compute = nullptr, h2d = nullptr, d2h = nullptr;
  cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, down[2] = {nullptr, nullptr};
  // opt-in INT8 tensor-core contraction (csrc/int8_fused.cu): sliced L^-1, alpha, workspace
  bool int8 = false, owns_i8 = false;
  void* d_pack8 = nullptr;
  double* d_alpha = nullptr;
  void* d_ws8 = nullptr;
  long long ws8_bytes = 0;
};

namespace mgb {
static void posterior_release(mgb_posterior* h) {
  if (!h) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  if (h->compute) cudaStreamSynchronize(h->compute);
  if (h->h2d) cudaStreamSynchronize(h->h2d);
  if (h->d2h) cudaStreamSynchronize(h->d2h);
  if (h->owns_state) {
    cudaFree(h->d_xt);
    cudaFree(h->d_coef);
    cudaFree(h->d_pack);
  }
  for (int b = 0; b < 2; ++b) {
    cudaFree(h->d_xc[b]);
    cudaFree(h->d_out[b]);
    if (h->up[b]) cudaEventDestroy(h->up[b]);
    if (h->done[b]) cudaEventDestroy(h->done[b]);
    if (h->down[b]) cudaEventDestroy(h->down[b]);
  }
  cudaFree(h->d_ws);
  if (h->owns_i8) {
    cudaFree(h->d_pack8);
    cudaFree(h->d_alpha);
  }
  cudaFree(h->d_ws8);
  if (h->h_stage) cudaFreeHost(h->h_stage);
  if (h->compute) cudaStreamDestroy(h->compute);
  if (h->h2d) cudaStreamDestroy(h->h2d);
  if (h->d2h) cudaStreamDestroy(h->d2h);
  cudaSetDevice(prev);
  delete h;
}

static int dev_alloc(double** p, size_t doubles) {
  return check_cuda(cudaMalloc(reinterpret_cast<void**>(p), sizeof(double) * (doubles ? doubles : 2)), "cudaMalloc");
}

// buffers every handle needs, whatever the origin of its state
static int posterior_common_init(mgb_posterior* h, long long max_chunk) {
  long long chunk = max_chunk > 0 ? max_chunk : 37888;     // 2 waves of 128-candidate tiles on 148 SMs
  chunk = (chunk + 127) / 128 * 128;
  h->chunk = chunk;
  for (int b = 0; b < 2; ++b) {
    MGB_TRY(dev_alloc(&h->d_xc[b], (size_t)(chunk * h->D)));
    MGB_TRY(dev_alloc(&h->d_out[b], (size_t)(2 * chunk)));
    MGB_TRY(check_cuda(cudaEventCreateWithFlags(&h->up[b], cudaEventDisableTiming), "cudaEventCreate"));
    MGB_TRY(check_cuda(cudaEventCreateWithFlags(&h->done[b], cudaEventDisableTiming), "cudaEventCreate"));
    MGB_TRY(check_cuda(cudaEventCreateWithFlags(&h->down[b], cudaEventDisableTiming), "cudaEventCreate"));
  }
  h->ws_bytes = mgb_posterior_workspace_bytes(chunk, h->n);
  MGB_TRY(check_cuda(cudaMalloc(&h->d_ws, (size_t)h->ws_bytes), "cudaMalloc"));
  h->stage_rows = chunk < 8192 ? chunk : 8192;
  MGB_TRY(check_cuda(cudaHostAlloc(reinterpret_cast<void**>(&h->h_stage), sizeof(double) * (size_t)(h->stage_rows * (h->D + 2)),
                                   cudaHostAllocDefault), "cudaHostAlloc"));
  MGB_TRY(check_cuda(cudaStreamCreateWithFlags(&h->compute, cudaStreamNonBlocking), "cudaStreamCreate"));
  MGB_TRY(check_cuda(cudaStreamCreateWithFlags(&h->h2d, cudaStreamNonBlocking), "cudaStreamCreate"));
  return check_cuda(cudaStreamCreateWithFlags(&h->d2h, cudaStreamNonBlocking), "cudaStreamCreate");
}
// INT8 mode of a handle: the workspace of mgb_series_posterior_i8f for one chunk.
static int posterior_int8_buffers(mgb_posterior* h) {
  if (h->n > 16384) return fail(MGB_ERR_ARG, "posterior handle: the INT8 path needs n <= 16384");
  if (!h->d_ws8) {
    h->ws8_bytes = mgb_series_posterior_i8f_workspace_bytes(h->chunk, h->n);
    MGB_TRY(check_cuda(cudaMalloc(&h->d_ws8, (size_t)h->ws8_bytes), "cudaMalloc"));
  }
  return MGB_OK;
}

// mean / var of up to `chunk`-sized blocks of candidate rows already on the device, on `stream`
static int posterior_eval_rows(mgb_posterior* h, const double* xc_dev, long long m, long long ldc, double* mean_dev, double* var_dev,
                               long long chunk, void* stream) {
  if (!h->int8)
    return mgb_series_posterior(&h->desc, xc_dev, m, ldc, h->d_xt, h->n, h->ldt, h->d_coef, h->d_pack, h->kss, h->mean_const,
                                mean_dev, var_dev, h->d_ws, h->ws_bytes, chunk, stream);
  for (long long c0 = 0; c0 < m; c0 += h->chunk) {
    const long long rows = m - c0 < h->chunk ? m - c0 : h->chunk;
    MGB_TRY(mgb_series_posterior_i8f(&h->desc, xc_dev + c0 * ldc, rows, ldc, h->d_xt, h->n, h->ldt, h->d_coef, h->d_pack8, h->d_alpha,
                                     h->kss, h->mean_const, mean_dev + c0, var_dev + c0, h->d_ws8, h->ws8_bytes, stream));
  }
  return MGB_OK;
}

static bool env_int8() {
  const char* e = std::getenv("MGB_POSTERIOR_PATH");
  return e && std::strcmp(e, "int8") == 0;
}
}  // namespace mgb

extern "C" int mgb_posterior_create(const mgb_series_desc* desc, const double* xtrain, long long n, const double* coef,
                                    const double* Rinv, const double* alpha, double kss_value, double mean_const,
                                    long long max_chunk, int device, mgb_posterior** out) {
  if (!desc || !xtrain || !coef || !Rinv || !alpha || !out || n < 1) return fail(MGB_ERR_ARG, "posterior_create: bad argument");
  *out = nullptr;
  MGB_TRY(select_device(device));
  mgb_posterior* h = new (std::nothrow) mgb_posterior();
  if (!h) return fail(MGB_ERR_ARG, "posterior_create: out of host memory");
  h->device = device; h->desc = *desc; h->n = n;
  h->D = (long long)desc->n_factors * desc->factor_width;
  h->T = (long long)desc->n_factors * desc->n_terms;
  h->ldt = h->D; h->kss = kss_value; h->mean_const = mean_const; h->owns_state = true;
  int rc = MGB_OK;
  double* d_R = nullptr;
  double* d_alpha = nullptr;
  do {
    if ((rc = posterior_common_init(h, max_chunk)) != MGB_OK) break;
    if ((rc = dev_alloc(&h->d_xt, (size_t)(n * h->D))) != MGB_OK) break;
    if ((rc = dev_alloc(&h->d_coef, (size_t)h->T)) != MGB_OK) break;
    if ((rc = check_cuda(cudaMalloc(reinterpret_cast<void**>(&h->d_pack), (size_t)mgb_posterior_pack_bytes(n)), "cudaMalloc")) != MGB_OK) break;
    if ((rc = dev_alloc(&d_R, (size_t)(n * n))) != MGB_OK) break;
    if ((rc = dev_alloc(&d_alpha, (size_t)n)) != MGB_OK) break;
    cudaStream_t s = h->compute;
    if ((rc = check_cuda(cudaMemcpyAsync(h->d_xt, xtrain, sizeof(double) * n * h->D, cudaMemcpyHostToDevice, s), "H2D xtrain")) != MGB_OK) break;
    if ((rc = check_cuda(cudaMemcpyAsync(h->d_coef, coef, sizeof(double) * h->T, cudaMemcpyHostToDevice, s), "H2D coef")) != MGB_OK) break;
    if ((rc = check_cuda(cudaMemcpyAsync(d_R, Rinv, sizeof(double) * n * n, cudaMemcpyHostToDevice, s), "H2D Rinv")) != MGB_OK) break;
    if ((rc = check_cuda(cudaMemcpyAsync(d_alpha, alpha, sizeof(double) * n, cudaMemcpyHostToDevice, s), "H2D alpha")) != MGB_OK) break;
    if ((rc = mgb_posterior_pack(d_R, n, n, d_alpha, h->d_pack, s)) != MGB_OK) break;
    if (env_int8()) {
      // MGB_POSTERIOR_PATH=int8: the handle also keeps the sliced factor and alpha and evaluates through int8_fused.cu
      if ((rc = check_cuda(cudaMalloc(&h->d_pack8, (size_t)mgb_posterior_i8f_pack_bytes(n)), "cudaMalloc")) != MGB_OK) break;
      h->owns_i8 = true;
      if ((rc = mgb_posterior_i8f_pack(d_R, n, n, h->d_pack8, s)) != MGB_OK) break;
      h->d_alpha = d_alpha;
      d_alpha = nullptr;
      if ((rc = posterior_int8_buffers(h)) != MGB_OK) break;
      h->int8 = true;
    }
    rc = check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize");
  } while (0);
  cudaFree(d_R);
  cudaFree(d_alpha);
  if (rc != MGB_OK) {
    posterior_release(h);
    return rc;
  }
  *out = h;
  return MGB_OK;
}

extern "C" int mgb_posterior_create_from_device(const mgb_series_desc* desc, const double* xtrain_dev, long long n,
                                                long long ldt, const double* coef_dev, const double* packed_dev,
                                                double kss_value, double mean_const, long long max_chunk, int device,
                                                mgb_posterior** out) {
  if (!desc || !xtrain_dev || !coef_dev || !packed_dev || !out || n < 1) return fail(MGB_ERR_ARG, "posterior_create_from_device: bad argument");
  *out = nullptr;
  MGB_TRY(select_device(device));
  mgb_posterior* h = new (std::nothrow) mgb_posterior();
  if (!h) return fail(MGB_ERR_ARG, "posterior_create_from_device: out of host memory");
  h->device = device; h->desc = *desc; h->n = n;
  h->D = (long long)desc->n_factors * desc->factor_width;
  h->T = (long long)desc->n_factors * desc->n_terms;
  if (ldt < h->D) { delete h; return fail(MGB_ERR_ARG, "posterior_create_from_device: ldt smaller than the row width"); }
  h->ldt = ldt; h->kss = kss_value; h->mean_const = mean_const; h->owns_state = false;
  h->d_xt = const_cast<double*>(xtrain_dev);
  h->d_coef = const_cast<double*>(coef_dev);
  h->d_pack = const_cast<double*>(packed_dev);
  const int rc = posterior_common_init(h, max_chunk);
  if (rc != MGB_OK) {
    posterior_release(h);
    return rc;
  }
  *out = h;
  return MGB_OK;
}

extern "C" int mgb_posterior_destroy(mgb_posterior* h) {
  posterior_release(h);
  return MGB_OK;
}

extern "C" long long mgb_posterior_chunk(const mgb_posterior* h) { return h ? h->chunk : 0; }

extern "C" int mgb_posterior_use_int8(mgb_posterior* h, const void* packed_i8_dev, const double* alpha_dev) {
  if (!h) return fail(MGB_ERR_ARG, "posterior_use_int8: null handle");
  if (!packed_i8_dev && !alpha_dev) {      // back to the FP64 kernel
    h->int8 = false;
    return MGB_OK;
  }
  if (!packed_i8_dev || !alpha_dev) return fail(MGB_ERR_ARG, "posterior_use_int8: both the sliced factor and alpha are needed");
  int prev = 0;
  cudaGetDevice(&prev);
  MGB_TRY(check_cuda(cudaSetDevice(h->device), "cudaSetDevice"));
  struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
  if (h->owns_i8) {
    cudaFree(h->d_pack8);
    cudaFree(h->d_alpha);
    h->owns_i8 = false;
  }
  h->d_pack8 = const_cast<void*>(packed_i8_dev);
  h->d_alpha = const_cast<double*>(alpha_dev);
  MGB_TRY(posterior_int8_buffers(h));
  h->int8 = true;
  return MGB_OK;
}

extern "C" int mgb_posterior_is_int8(const mgb_posterior* h) { return h && h->int8 ? 1 : 0; }

extern "C" int mgb_posterior_eval_device(mgb_posterior* h, const double* xc_dev, long long m, long long ldc,
                                         double* mean_dev, double* var_dev, void* stream) {
  if (!h || m < 0) return fail(MGB_ERR_ARG, "posterior_eval_device: bad argument");
  if (m == 0) return MGB_OK;
  if (!xc_dev || !mean_dev || !var_dev || ldc < h->D) return fail(MGB_ERR_ARG, "posterior_eval_device: null pointer / bad ldc");
  const long long chunk = m < h->chunk ? (m + 127) / 128 * 128 : h->chunk;
  return posterior_eval_rows(h, xc_dev, m, ldc, mean_dev, var_dev, chunk, stream);
}

extern "C" int mgb_posterior_eval_host(mgb_posterior* h, const double* xc, long long m, double* mean, double* var) {
  if (!h || m < 0) return fail(MGB_ERR_ARG, "posterior_eval_host: bad argument");
  if (m == 0) return MGB_OK;
  if (!xc || !mean || !var) return fail(MGB_ERR_ARG, "posterior_eval_host: null pointer");
  int prev = 0;
  cudaGetDevice(&prev);
  MGB_TRY(check_cuda(cudaSetDevice(h->device), "cudaSetDevice"));
  struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev};
  const long long D = h->D;
  if (m <= h->stage_rows) {
    // latency path (the reference's acq(X) calls: b = 1 ... a few hundred): one pinned staging block, one stream,
    // one synchronisation; pageable user memory never meets the DMA engine
    double* st_x = h->h_stage;
    double* st_out = h->h_stage + h->stage_rows * D;
    memcpy(st_x, xc, sizeof(double) * (size_t)(m * D));
    cudaStream_t s = h->compute;
    MGB_TRY(check_cuda(cudaMemcpyAsync(h->d_xc[0], st_x, sizeof(double) * m * D, cudaMemcpyHostToDevice, s), "H2D xc"));
    MGB_TRY(mgb_posterior_eval_device(h, h->d_xc[0], m, D, h->d_out[0], h->d_out[0] + m, s));
    MGB_TRY(check_cuda(cudaMemcpyAsync(st_out, h->d_out[0], sizeof(double) * 2 * m, cudaMemcpyDeviceToHost, s), "D2H mean/var"));
    MGB_TRY(check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize"));
    memcpy(mean, st_out, sizeof(double) * (size_t)m);
    memcpy(var, st_out + m, sizeof(double) * (size_t)m);
    return MGB_OK;
  }
  // throughput path: blocks of `chunk` rows, uploads / kernels / downloads on three streams, two buffers in flight
  const long long chunk = h->chunk;
  const long long nblk = (m + chunk - 1) / chunk;
  auto rows_of = [&](long long blk) { const long long c0 = blk * chunk; return (m - c0 < chunk) ? (m - c0) : chunk; };
  auto upload = [&](long long blk) -> int {
    const int buf = (int)(blk & 1);
    if (blk >= 2) MGB_TRY(check_cuda(cudaStreamWaitEvent(h->h2d, h->done[buf], 0), "cudaStreamWaitEvent"));  // kernels of blk-2 read it
    MGB_TRY(check_cuda(cudaMemcpyAsync(h->d_xc[buf], xc + blk * chunk * D, sizeof(double) * rows_of(blk) * D,
                                       cudaMemcpyHostToDevice, h->h2d), "H2D xc"));
    return check_cuda(cudaEventRecord(h->up[buf], h->h2d), "cudaEventRecord");
  };
  MGB_TRY(upload(0));
  for (long long blk = 0; blk < nblk; ++blk) {
    const long long c0 = blk * chunk, rows = rows_of(blk);
    const int buf = (int)(blk & 1);
    MGB_TRY(check_cuda(cudaStreamWaitEvent(h->compute, h->up[buf], 0), "cudaStreamWaitEvent"));
    if (blk >= 2) MGB_TRY(check_cuda(cudaStreamWaitEvent(h->compute, h->down[buf], 0), "cudaStreamWaitEvent"));  // results of blk-2 left
    MGB_TRY(posterior_eval_rows(h, h->d_xc[buf], rows, D, h->d_out[buf], h->d_out[buf] + chunk, chunk, h->compute));
    MGB_TRY(check_cuda(cudaEventRecord(h->done[buf], h->compute), "cudaEventRecord"));
    if (blk + 1 < nblk) MGB_TRY(upload(blk + 1));
    MGB_TRY(check_cuda(cudaStreamWaitEvent(h->d2h, h->done[buf], 0), "cudaStreamWaitEvent"));
    MGB_TRY(check_cuda(cudaMemcpyAsync(mean + c0, h->d_out[buf], sizeof(double) * rows, cudaMemcpyDeviceToHost, h->d2h), "D2H mean"));
    MGB_TRY(check_cuda(cudaMemcpyAsync(var + c0, h->d_out[buf] + chunk, sizeof(double) * rows, cudaMemcpyDeviceToHost, h->d2h), "D2H var"));
    MGB_TRY(check_cuda(cudaEventRecord(h->down[buf], h->d2h), "cudaEventRecord"));
  }
  MGB_TRY(check_cuda(cudaStreamSynchronize(h->compute), "cudaStreamSynchronize"));
  MGB_TRY(check_cuda(cudaStreamSynchronize(h->h2d), "cudaStreamSynchronize"));
  return check_cuda(cudaStreamSynchronize(h->d2h), "cudaStreamSynchronize");
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
