/*
 * mgb.h - C ABI of libmgb.so: B200 (sm_100a) implementation of the MaternGaBO hot path.
 *
 * The reference (NoemieJaquier/MaternGaBO) is pure Python: its "FFI" for this path is the body of
 * each gpytorch Kernel.forward, a composition of torch ops.  Every entry point below replaces one
 * such composition; the comment above it cites the reference lines (relative to
 * /root/reference/BoManifolds).  INTEGRATION.md shows the ctypes binding a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; all matrices row-major float64 with an explicit leading
 *     dimension (elements, not bytes);
 *   - "device" entry points take device pointers and a cudaStream_t (passed as void*); they never
 *     allocate, never synchronise and never throw; they return MGB_OK or an error code, and
 *     mgb_last_error() returns a static message for the calling thread;
 *   - "_host" entry points take HOST pointers, stage through device memory internally (host<->device
 *     copies included) and synchronise before returning;
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef MGB_H_
#define MGB_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MGB_OK 0
#define MGB_ERR_ARG 1      /* invalid argument (null pointer, bad size, unsupported width ...) */
#define MGB_ERR_CUDA 2     /* a CUDA runtime call or kernel launch failed */
#define MGB_ERR_NO_DEVICE 3

#define MGB_BASIS_MONOMIAL 0   /* p(u) = sum_k c_k u^k, Horner */
#define MGB_BASIS_CHEBYSHEV 1  /* p(u) = sum_k c_k T_k(u), Clenshaw */

#define MGB_MAX_FACTORS 16
#define MGB_MAX_TERMS 256

const char* mgb_last_error(void);
int mgb_version(void);
/* number of kernel launches issued by this library in this process (bench.py's gpu_launches) */
long long mgb_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Spectral-series Gram matrix (sphere S^d, SO(3), circle, torus T^d)
 *
 *   K[i][j] = prod_{f < n_factors} p_f( clamp(scale * <x1[i][f*W : (f+1)*W], x2[j][f*W : (f+1)*W]> + shift, lo, hi) )
 *   p_f(u)  = sum_{k < n_terms} coef[f*n_terms + k] * phi_k(u),  phi_k = u^k or T_k(u)
 *
 * replaces sphere_distance_torch + gegenbauer_polynomial + the term loop of
 *   kernel_utils/kernels_sphere.py:102-139 (Matern), :188-224 (heat), :343-384 (integrated Matern),
 *   :443-472 / :551-614 (circle; Jacobi theta_3 series), kernel_utils/kernels_torus.py:61-90,142-177
 *   (product over circles), kernel_utils/kernels_so.py:305-362 and :149-177 with dim = 3
 *   (W = 9, scale = 0.5, shift = -0.5, clamp +-(1-1e-8)).
 * The hyper-parameter dependence lives entirely in coef (host side, torch autograd).
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int n_factors;    /* F >= 1 */
  int factor_width; /* W >= 1; rows of x1/x2 are F*W wide */
  int n_terms;      /* T in [1, MGB_MAX_TERMS] */
  int basis;        /* MGB_BASIS_* */
  double scale, shift, lo, hi;
  int pair_square;  /* 0: u = scale * <x1, x2> + shift;  1: u = scale * <x1, x2>^2 + shift (forward only).
                       SO(3) through unit quaternions: (tr(R1 R2^T) - 1) / 2 = 2 <q1, q2>^2 - 1, so rows of 4
                       (mgb_so3_to_quaternion) with scale 2, shift -1 replace rows of 9 with scale 1/2, shift -1/2 */
} mgb_series_desc;

int mgb_series_gram_fwd(const mgb_series_desc* desc, const double* x1, long long m, long long ld1,
                        const double* x2, long long n, long long ld2, const double* coef,
                        double* K, long long ldk, void* stream);

/* gpytorch batch dimensions (b x m x D against b x n x D, e.g. a materialised expansion of the training inputs in the
 * posterior call): `batch` Gram matrices from ONE launch (gridDim.z = batch); stride1 / stride2 / stridek
 * = elements between consecutive batch members (0 = shared operand). */
int mgb_series_gram_fwd_batched(const mgb_series_desc* desc, const double* x1, long long m, long long ld1,
                                long long stride1, const double* x2, long long n, long long ld2, long long stride2,
                                const double* coef, double* K, long long ldk, long long stridek, long long batch,
                                void* stream);
/* diag=True of Kernel.forward: out[i] = k(x1[i], x2[i]) for two equally long point lists, one thread per row
 * (the reference computes a full bmm on expanded pairs: Riemannian_utils/sphere_utils_torch.py:50-53). */
int mgb_series_gram_diag(const mgb_series_desc* desc, const double* x1, long long ld1, const double* x2, long long ld2,
                         long long rows, const double* coef, double* out, void* stream);

/* Backward of the above for L = sum(K * G):
 *   gcoef[f*T + k] += dL/dcoef (accumulated with atomics: caller zeroes it),
 *   gx1 (m x F*W, ld = F*W) = dL/dx1 (overwritten; may be NULL), gx2 likewise (may be NULL).
 * Pairs where the clamp is active get zero input gradient, as torch.clamp does in the reference
 * (Riemannian_utils/sphere_utils_torch.py:57).  G is addressed as G[i*ldg_row + j*ldg_col]. */
int mgb_series_gram_bwd(const mgb_series_desc* desc, const double* x1, long long m, long long ld1,
                        const double* x2, long long n, long long ld2, const double* coef,
                        const double* G, long long ldg_row, long long ldg_col,
                        double* gcoef, double* gx1, double* gx2, void* stream);

/* Rotation matrices (rows of 9, row-major 3 x 3) -> unit quaternions (rows of 4: w, x, y, z) by Shepperd's method, for
 * the pair_square form of the SO(3) kernels.  *defect (device scalar, caller zeroes) receives the maximum over rows
 * of max(|R R^T - I|_max, |det R - 1|): the identity above holds only for rotations, so the caller falls back to
 * the 9-wide form when the inputs are not orthogonal to rounding (kernel_utils/kernels_so.py:323-327 evaluates the
 * trace formula on whatever matrices it is given). */
int mgb_so3_to_quaternion(const double* R, long long m, long long ld, double* Q, double* defect, void* stream);

/* Second-order support for trust-region Hessian-vector products (pymanopt_addons/tools/autodiff/
 * _pytorch.py:92-116): same contraction as gx1 of the backward but with the k-th derivative of p_f,
 *   out1[i][:] = sum_j G[i][j] * p^(order)(u_ij) * mask_ij * scale^order * x2[j][:]   (n_factors == 1)
 * order in {1, 2}.  out2 is the x2-side sum (may be NULL). */
int mgb_series_gram_dx(const mgb_series_desc* desc, int order, const double* x1, long long m, long long ld1,
                       const double* x2, long long n, long long ld2, const double* coef,
                       const double* G, long long ldg_row, long long ldg_col,
                       double* out1, double* out2, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SO(n) character sums for n != 3 (n = 2, 4, 5, 6, 7; rank r = n / 2), rows of n*n (row-major rotation matrices).
 *
 * Replaces the dim != 3 branch of kernel_utils/kernels_so.py:329-362 (Matern) / :149-177 (heat): bmm -> torch.linalg.
 * eigvals -> Python loop pairing conjugate eigenvalues -> sympy-lambdified character polynomial.  The polynomial is the
 * reference's (host side: materngabo_b200/_so_characters.py rebuilds it exactly); its argument is the vector of
 * eigen-angles of R1 R2^T in (0, pi) in DESCENDING order, obtained in registers from tr R, tr R^2, tr R^3 (closed form
 * for r <= 2, trigonometric cubic for r = 3).  The reference feeds the angles in the order LAPACK lists the conjugate
 * pairs, which makes its value order dependent; the two agree where that order is descending.
 *   grids: [patterns][(J+1)^r] normalised coefficients G_s[a] of prod_i (cos | sin)(a_i theta_i); patterns = sign
 *          patterns with an even number of sines in lexicographic order: r = 1: (c); r = 2: (c,c), (s,s);
 *          r = 3: (c,c,c), (c,s,s), (s,c,s), (s,s,c).  J <= 12.
 *   mode 0: out = K (m x n, ld = ldo).   mode 1: out[(i*n + j)*B + b] = the b-th basis product, B = mgb_son_basis_size
 *          (hyper-parameter gradients: K = basis . grids in torch; grids unused).
 * ------------------------------------------------------------------------------------------------ */
long long mgb_son_basis_size(int dim, int J);
int mgb_son_gram(int dim, int mode, const double* x1, long long m, long long ld1, const double* x2, long long n,
                 long long ld2, const double* grids, int J, double* out, long long ldo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Quadrature Gram matrices.
 *
 * Hyperbolic space H^dim in Lorentz coordinates (rows dim+1 wide), dim in {1,2,3}:
 *   d_ij = 2 asinh( sqrt(max(0, <D,D>_L) + 1e-8) / 2 ),  D = x1_i - x2_j
 *            (Riemannian_utils/hyperbolic_utils_torch.py:11-60)
 *   K_ij = sum_{a < Pt} tcoef[a] * heat(d_ij, tval[a])
 *   heat(d, t) = exp(-d^2/(4t))                                       dim 1 (kernels_hyperbolic.py:258-275)
 *              = exp(-d^2/(4t)) (d+1e-8)/sinh(d+1e-8)                 dim 3 (:306-330)
 *              = trapz_s  s exp(-s^2/(4t)) / sqrt(cosh s - cosh d),  s = s0 + k*ds + d, k < Ps   dim 2 (:281-304)
 * The integrated Matern kernel (:336-389) is this with tval = its t grid and
 * tcoef[a] = trapezoid weight * t^(nu-1) exp(-2 nu t / kappa^2) / normaliser (host side);
 * the heat kernel (:58-88) is the same with Pt = 1, tval[0] = kappa^2 / 2.
 * mgb_hyperbolic_heat_nodes writes heat(dist[e], tval[a]) to out[e*Pt + a] for given distances
 * (used for the normaliser at d = 0 and for tests).
 * ------------------------------------------------------------------------------------------------ */
int mgb_hyperbolic_gram_fwd(int dim, const double* x1, long long m, long long ld1, const double* x2,
                            long long n, long long ld2, const double* tval, const double* tcoef, int Pt,
                            double s0, double ds, int Ps, double* K, long long ldk, void* stream);
/* gtcoef[a] += sum_ij G_ij heat(d_ij, tval[a]) (caller zeroes) */
int mgb_hyperbolic_gram_bwd_coef(int dim, const double* x1, long long m, long long ld1, const double* x2,
                                 long long n, long long ld2, const double* tval, int Pt, double s0, double ds,
                                 int Ps, const double* G, long long ldg, double* gtcoef, void* stream);
/* gtval[a] += sum_ij G_ij tcoef[a] d heat(d_ij, t)/dt at t = tval[a] (caller zeroes): lengthscale gradient of the
 * heat kernels, whose t = kappa^2 / 2 is differentiable (the integrated Matern grids are detached in the reference) */
int mgb_hyperbolic_gram_bwd_tval(int dim, const double* x1, long long m, long long ld1, const double* x2,
                                 long long n, long long ld2, const double* tval, const double* tcoef, int Pt,
                                 double s0, double ds, int Ps, const double* G, long long ldg, double* gtval,
                                 void* stream);
/* derivative != 0: d heat / d t instead of heat */
int mgb_hyperbolic_heat_nodes(int dim, int derivative, const double* dist, long long count, const double* tval,
                              int Pt, double s0, double ds, int Ps, double* out, void* stream);

/* Euclidean factor of the SPD product kernels (kernel_utils/kernels_euclidean.py:82-89,162-207):
 *   K_ij = sum_a tcoef[a] exp(-|x1_i - x2_j|^2 / (4 tval[a])), rows `width` wide. */
int mgb_euclidean_gram_fwd(const double* x1, long long m, long long ld1, const double* x2, long long n,
                           long long ld2, int width, const double* tval, const double* tcoef, int Pt,
                           double* K, long long ldk, void* stream);
int mgb_euclidean_gram_bwd_coef(const double* x1, long long m, long long ld1, const double* x2, long long n,
                                long long ld2, int width, const double* tval, int Pt, const double* G,
                                long long ldg, double* gtcoef, void* stream);

/* SPD(2) in Mandel coordinates (rows 3 wide: s11, s22, sqrt2*s12)
 *   (Riemannian_utils/spd_utils_torch.py:336-371; kernel_utils/kernels_spd.py:200-289 and :65-136)
 *   H1 >= H2 = log singular values of A_i B_j^{-1}; A, B = the matrices (use_cholesky = 0, Matern :230-243)
 *   or their lower Cholesky factors (use_cholesky = 1, heat kernel :90-102); Delta = H1-H2, r2 = H1^2+H2^2
 *   K_ij = sum_a tcoef[a] exp(-r2 * rscale[a]) * sum_b bw[b] (2 b_b + Delta) exp(-b_b (b_b + Delta) * bscale[a])
 *                                                     / sqrt(sinh b_b * sinh(b_b + Delta))
 *   Matern: rscale[a] = 1/(4 t_a), bscale[a] = 1/(2 t_a); heat: Pt = 1, rscale = 1/(2 k^2), bscale = 1/k^2. */
int mgb_spd2_gram_fwd(int use_cholesky, const double* x1, long long m, long long ld1, const double* x2,
                      long long n, long long ld2, const double* tcoef, const double* rscale,
                      const double* bscale, int Pt, const double* bval, const double* bw, int Pb,
                      double* K, long long ldk, void* stream);
/* Same result as mgb_spd2_gram_fwd for geometric grids whose log-steps are commensurate:
 *   bval[b] * bscale[a] = g_min * exp(log_tau * (sb * b + sa * (Pt - 1 - a)))   for all a < Pt, b < Pb
 * (the reference's grids, kernels_spd.py:262-265: 12 : 7 when Pb == Pt).  The pair-dependent exponential
 * exp(-Delta bval[b] bscale[a]) is then evaluated once per lattice point instead of once per node.  The caller
 * verifies the lattice (host side, materngabo_b200/kernels_spd.py); Pt <= 64. */
int mgb_spd2_gram_fwd_lattice(int use_cholesky, const double* x1, long long m, long long ld1, const double* x2,
                              long long n, long long ld2, const double* tcoef, const double* rscale,
                              const double* bscale, int Pt, const double* bval, const double* bw, int Pb,
                              int sb, int sa, double g_min, double log_tau, double* K, long long ldk, void* stream);
/* g_rscale[a] += dL/drscale[a], g_bscale[a] += dL/dbscale[a] (caller zeroes): lengthscale gradient of the SPD(2) heat
 * kernel, whose scales 1/(2 kappa^2), 1/kappa^2 are differentiable */
int mgb_spd2_gram_bwd_scales(int use_cholesky, const double* x1, long long m, long long ld1, const double* x2,
                             long long n, long long ld2, const double* tcoef, const double* rscale,
                             const double* bscale, int Pt, const double* bval, const double* bw, int Pb,
                             const double* G, long long ldg, double* g_rscale, double* g_bscale, void* stream);
int mgb_spd2_gram_bwd_coef(int use_cholesky, const double* x1, long long m, long long ld1, const double* x2,
                           long long n, long long ld2, const double* rscale, const double* bscale, int Pt,
                           const double* bval, const double* bw, int Pb, const double* G, long long ldg,
                           double* gtcoef, void* stream);

/* Eigen-decomposition front-end of the SPD product kernels (spd_input=True; Riemannian_utils/spd_utils_torch.py:293-333
 * called from kernel_utils/kernels_spd.py:520-533,746-757; the reference uses torch.symeig, removed from torch >= 1.13):
 * rows of x are Mandel vectors of SPD(dim) matrices, dim in {2, 3} (3 / 6 wide).  Per row, one thread: cyclic Jacobi in
 * registers -> eigenvalues ascending, eigenvectors as columns, sign of each eigenvector fixed by "largest-magnitude
 * component positive", then Q = det(V) V as in the reference.
 *   target 0: out row = [eigenvalues (dim) | Q row-major (dim^2)]                  (R^d x SO(d) kernels)
 *   target 1: out row = [eigenvalues (dim) | first row of Q (dim 2) or quaternion (dim 3, Corke's algorithm as in
 *             Riemannian_utils/sphere_utils_torch.py:127-190)]                        (R^d x S kernels) */
int mgb_spd_decompose(int dim, int target, const double* x, long long m, long long ld, double* out, long long ldo,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * Radial profiles of the quadrature kernels and their derivatives in the pair scalar (csrc/quad_profile.cu).
 *
 * Input gradients (and Hessian-vector products) of the acquisition function, which the reference obtains by
 * torch autograd through Kernel.forward (manifold_optimization/manifold_optimize.py:184-217,
 * pymanopt_addons/tools/autodiff/_pytorch.py:89-116), factor by the chain rule into cheap pair geometry
 * (distance / log singular values as functions of the inputs) and derivatives of the expensive quadrature with
 * respect to the pair scalar.  These entry points evaluate the latter for a flat list of pair scalars:
 *
 *   mgb_radial_profile:  out[e] = sum_a tcoef[a] * d^order/dz^order heat(z[e], tval[a]),  order in {0, 1, 2}
 *     kind 0/1/2 = H^1/H^2/H^3 with z = geodesic distance and heat as above (for H^2 the s grid moves with d,
 *     kernels_hyperbolic.py:299, and the derivative includes it, as autograd does in the reference);
 *     kind 3 = Euclidean factor, z = |x - y|^2, heat = exp(-z / (4 t)) (kernels_euclidean.py:82-89).
 *     order 3 = all three orders from one launch, out[o * count + e] (the exponentials are shared);
 *   mgb_radial_profile_bwd_coef:  gtcoef[a] += sum_e G[e] * d^order/dz^order heat(z[e], tval[a])  (caller zeroes)
 *   mgb_spd2_profile:  out[e] = sum_a tcoef[a] * D h_a(delta[e], r2[e]),  which = 0: D = 1, 1: d/dDelta, 2: d/dr2,
 *     h_a(Delta, r2) = exp(-r2 rscale[a]) sum_b bw[b] (2 b_b + Delta) exp(-b_b (b_b + Delta) bscale[a])
 *                                                   / sqrt(sinh b_b sinh(b_b + Delta))   (kernels_spd.py:200-209)
 *     which 3 = all three from one launch, out[which * count + e];
 *   mgb_spd2_profile_bwd_coef:  gtcoef[a] += sum_e G[e] * D h_a(delta[e], r2[e])  (caller zeroes)
 * ------------------------------------------------------------------------------------------------ */
int mgb_radial_profile(int kind, int order, const double* z, long long count, const double* tval,
                       const double* tcoef, int Pt, double s0, double ds, int Ps, double* out, void* stream);
int mgb_radial_profile_bwd_coef(int kind, int order, const double* z, long long count, const double* tval, int Pt,
                                double s0, double ds, int Ps, const double* G, double* gtcoef, void* stream);
int mgb_spd2_profile(int which, const double* delta, const double* r2, long long count, const double* tcoef,
                     const double* rscale, const double* bscale, int Pt, const double* bval, const double* bw,
                     int Pb, double* out, void* stream);
int mgb_spd2_profile_bwd_coef(int which, const double* delta, const double* r2, long long count,
                              const double* rscale, const double* bscale, int Pt, const double* bval,
                              const double* bw, int Pb, const double* G, double* gtcoef, void* stream);

/* Opt-in tabulated form of a radial kernel (hyperbolic H^1..H^3: geom 0, rows dim+1 wide; Euclidean: geom 1):
 *   K_ij = poly_{I(v_ij)}(s_ij),  v = z + 1,  z = sqrt(max(0, <D,D>_L) + 1e-8) (= 2 sinh(d/2)) or |x1_i - x2_j|,
 * interval I = (binary exponent of v) * 2^sub_bits + (top sub_bits mantissa bits of v), s in [-1, 1) the position
 * inside the interval, coef[I][0..degree] monomial coefficients in s.  The table is built on the host side from
 * mgb_radial_profile (the same quadrature as the per-entry kernels) and validated against it
 * (materngabo_b200/tabulated.py).  This is NOT how the reference evaluates these kernels (it integrates per entry,
 * as mgb_hyperbolic_gram_fwd does); it is offered for large candidate sets and reported separately in bench.py. */
int mgb_radial_table_gram(int geom, int dim, const double* x1, long long m, long long ld1, const double* x2,
                          long long n, long long ld2, int sub_bits, int n_intervals, int degree, const double* coef,
                          double* K, long long ldk, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Exact-GP posterior for a block of candidates (gpytorch ExactGP prediction + botorch call sites:
 * examples/bo_manifolds_benchmark_examples/bo_sphere.py:215-233,265;
 * manifold_optimization/manifold_optimize.py:195,254,337)
 *
 *   Kc    = outputscale * K(Xc, Xtrain)                      (m x n, written by the caller with
 *                                                             mgb_series_gram_fwd etc. into Kc, ld = ldk)
 *   mean_c = mean_const + sum_j Kc[c][j] alpha[j]
 *   var_c  = kss[c] - sum_i ( sum_{j<=i} Rinv[i][j] Kc[c][j] )^2,   Rinv = L^{-1}, L = chol(s K + noise I)
 *
 * mgb_posterior_pack builds, once per GP fit, the packed factor: lower triangle of Rinv padded with zeros
 * to (n+1 rounded up to 128) x (n rounded up to 16) doubles, row n holding alpha (so the mean falls out of
 * the same tiles), stored tile-major: 128 x 16 blocks [row tile][column chunk][row][16], the eight 16-byte
 * pieces of a block row XOR-swizzled by 2*(row & 3).  One block = one 16 KB TMA bulk copy into shared memory,
 * after which the DMMA fragment loads are bank-conflict free.
 * mgb_series_gram_tiled writes K(Xc, Xtrain) for a block of candidates in the same layout
 * ([candidate tile][column chunk][row][16], mgb_posterior_tiled_bytes(m, n) bytes, 16-byte aligned; rows beyond
 * m and columns beyond n are not written and must hold finite values - clear the buffer once).
 * mgb_posterior_reduce is the fused triangular product + square-sum + mean on FP64 tensor-core DMMA tiles;
 * V = Rinv Kc^T is never written to memory.  kss: m values (k(x_c, x_c) * outputscale).
 * Results are deterministic (fixed summation order).
 * ------------------------------------------------------------------------------------------------ */
long long mgb_posterior_pack_bytes(long long n);
int mgb_posterior_pack(const double* Rinv, long long n, long long ldr, const double* alpha, double* packed,
                       void* stream);
long long mgb_posterior_tiled_bytes(long long m, long long n);
int mgb_series_gram_tiled(const mgb_series_desc* desc, const double* x1, long long m, long long ld1,
                          const double* x2, long long n, long long ld2, const double* coef, double* Kb,
                          void* stream);
/* row-major K (m x n) -> the tiled layout (Gram matrices of the quadrature kernels, which write row-major) */
int mgb_posterior_retile(const double* K, long long m, long long n, long long ldk, double* Kb, void* stream);
/* bytes of scratch that suffice for mgb_posterior_reduce on ANY block of at most m candidates (a chunked caller sizes it
 * once for its chunk and reuses it for the shorter tail block, which needs more splits per candidate tile) */
long long mgb_posterior_reduce_workspace_bytes(long long m, long long n);
int mgb_posterior_reduce(const double* Kb, long long m, long long n, const double* packed, const double* kss,
                         double mean_const, double* mean, double* var, void* workspace,
                         long long workspace_bytes, void* stream);
/* scratch mgb_series_posterior needs for `chunk` candidates x n training points (bytes) */
long long mgb_posterior_workspace_bytes(long long chunk, long long n);

/* Whole posterior for a series kernel, candidates in device memory, processed `chunk` rows at a time
 * through `workspace` (16-byte aligned, >= mgb_posterior_workspace_bytes(chunk, n)): Gram chunk -> reduce.
 * coef already contains outputscale / normaliser.  kss_value: k(x, x) * outputscale (the kernels are
 * stationary and normalised: one number). */
int mgb_series_posterior(const mgb_series_desc* desc, const double* xc, long long m, long long ldc,
                         const double* xtrain, long long n, long long ldt, const double* coef,
                         const double* packed, double kss_value, double mean_const, double* mean,
                         double* var, void* workspace, long long workspace_bytes, long long chunk,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused exact-GP marginal likelihood for launch-bound training sizes (csrc/mll.cu): one CTA, everything in
 * shared memory.  Replaces one evaluation + backward of gpytorch's ExactMarginalLogLikelihood as called by
 * botorch.fit_gpytorch_model (examples/bo_manifolds_benchmark_examples/bo_sphere.py:228,257-262) for a
 * single-factor series kernel:
 *   Kt = hyper[0] * K(X, X) + hyper[1] * I,  r = y - hyper[2]      (outputscale, noise, constant mean: DEVICE scalars)
 *   out[0] = 1/2 r^T Kt^-1 r + 1/2 log det Kt + n/2 log(2 pi)
 *   out[1..3] = d out[0] / d (outputscale, noise, mean),  out[4 + k] = d out[0] / d coef[k]
 *   *status = 0, or j + 1 when the j-th pivot of the Cholesky factorisation was not positive (outputs are NaN).
 * n <= mgb_series_mll_single_cta_max_n(factor_width, n_terms) (160 for S^3 with 10 terms): ONE launch, everything in
 * shared memory (mgb_series_mll / mgb_dense_mll / mgb_series_fit / mgb_dense_fit).  Up to mgb_series_mll_max_n (16384):
 * the multi-CTA path below (mgb_series_mll_large / mgb_dense_mll_large / mgb_dense_fit_large, caller-owned workspace).
 * ------------------------------------------------------------------------------------------------ */
long long mgb_series_mll_max_n(int factor_width, int n_terms);
long long mgb_series_mll_single_cta_max_n(int factor_width, int n_terms);

/* Multi-CTA dense SPD path (csrc/dense_spd.cu) for 160 < n <= mgb_dense_max_n(): replaces torch.linalg.cholesky /
 * cholesky_solve / solve_triangular (cuSOLVER, cuBLAS) in the marginal-likelihood step and in the posterior fit.
 *   Cholesky: one COOPERATIVE launch for the whole factorisation (32-column panels; diagonal block in one warp by
 *   shuffles; grid barriers instead of launches);  L^-1: recursive doubling, two batched GEMMs per level;
 *   Kt^-1 = L^-T L^-1: one GEMM with the k-range clipped to the triangles.
 * Same outputs as the single-CTA entry points.  hyper = DEVICE scalars (outputscale, noise, mean).  Workspaces are
 * caller-owned, 16-byte aligned, mgb_dense_{fit,mll}_workspace_bytes(n) bytes.  mgb_dense_fit_large: rinv (and rinv_t,
 * may be NULL) are n x ldo with ldo = n rounded up to an even number; zeros above the diagonal.  *status = 0 or the
 * failing pivot + 1. */
long long mgb_dense_max_n(void);
long long mgb_dense_fit_workspace_bytes(long long n);
long long mgb_dense_mll_workspace_bytes(long long n);
int mgb_dense_fit_large(const double* K, long long n, long long ldk, const double* y, const double* hyper, double* rinv,
                        double* rinv_t, long long ldo, double* alpha, double* nll, int* status, void* workspace,
                        long long workspace_bytes, void* stream);
int mgb_dense_mll_large(const double* K, long long n, long long ldk, const double* y, const double* hyper, double* out,
                        double* G, long long ldg, int* status, void* workspace, long long workspace_bytes, void* stream);
int mgb_series_mll_large(const mgb_series_desc* desc, const double* x, long long n, long long ld, const double* y,
                         const double* coef, const double* hyper, double* out, int* status, void* workspace,
                         long long workspace_bytes, void* stream);
int mgb_series_mll(const mgb_series_desc* desc, const double* x, long long n, long long ld, const double* y,
                   const double* coef, const double* hyper, double* out, int* status, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused analytic Expected Improvement with gradient and Hessian-vector product (csrc/acquisition.cu): what the
 * pymanopt cost / egrad / ehess closures of the reference compute with botorch + torch autograd at one point per call
 * (manifold_optimization/manifold_optimize.py:184-217; pymanopt_addons/tools/autodiff/_pytorch.py:89-116), for a batch
 * of b candidates (e.g. all restarts) in one launch.  Series kernels with rows up to 32 wide (sphere, SO(3), circle;
 * multi-factor = torus in (cos, sin) coordinates); coef includes the outputscale (on factor 0 for the torus).
 *   rinv = L^-1 (n x n row-major, lower triangle used), rinv_t its transpose (upper triangle used), alpha = Kt^-1 (y - m)
 *   EI = sigma (phi(u) + u Phi(u)), sigma = sqrt(max(var, 1e-9)), u = +-(mean - best_f) / sigma (+: maximize != 0)
 *   ei[c];  grad[c*D + k] = d EI / d xc[c][k] (may be NULL);  hvp[c*D + k] = sum_l d2 EI / dx_k dx_l v[c][l], D = F*W
 *   (v, hvp may be NULL);  mean, var: optional posterior outputs.  Pairs where the clamp of the pair scalar is
 *   active contribute zero gradient, as torch.clamp does in the reference.  n <= ~2800 (shared-memory vectors).
 * ------------------------------------------------------------------------------------------------ */
int mgb_series_acquisition_ei(const mgb_series_desc* desc, const double* xc, long long b, long long ldc,
                              const double* v, long long ldv, const double* xtrain, long long n, long long ldt,
                              const double* coef, const double* rinv, long long ldr, const double* rinv_t,
                              long long ldrt, const double* alpha, double kss, double mean_const, double best_f,
                              int maximize, double* ei, double* grad, double* hvp, double* mean, double* var,
                              void* stream);

/* Chebyshev coefficients (in cos(phi)) of the circle kernels - the torus factors - and their Jacobian in one launch:
 *   mode 0, integrated Matern (kernels_sphere.py:551-614): a_n = (2 - delta_n0) sum_a tw[a] t[a]^(nu - 1/2)
 *           exp(-2 nu t[a] / kappa^2) exp(-4 pi^2 t[a] n^2), t = the reference's logspace grid (P nodes, built by the
 *           caller from the detached lengthscale), tw = its trapezoid weights;
 *   mode 1, heat (kernels_sphere.py:443-472, math_utils/jacobi_theta_functions.py:16-20): a_0 = 1, a_n = 2 q^(n^2),
 *           q = exp(-2 pi^2 kappa^2) (t, tw, nu unused);
 *   coef[n] = a_n / sum_m a_m, n < n_terms (<= 256; the reference sums 201),  jac[0][n] = d coef_n / d kappa,
 *   jac[1][n] = d coef_n / d nu.  kappa, nu: device scalars. */
int mgb_circle_coef(int mode, int n_terms, const double* t, const double* tw, int P, const double* kappa,
                    const double* nu, double* coef, double* jac, void* stream);

/* Normalised t-node weights of the integrated-Matern quadrature kernels (hyperbolic, SPD(2), Euclidean) and their
 * Jacobian in one launch: w_a = tw[a] t[a]^(nu - power_offset) exp(-2 nu t[a] / kappa^2),
 * tcoef[a] = w_a / sum_b w_b h0[b] (h0 = the inner integral at zero distance on the same grid, NULL = 1;
 * kernels_hyperbolic.py:367-383, kernels_spd.py:262-289, kernels_euclidean.py:186-207),
 * jac[0][a] = d tcoef_a / d kappa, jac[1][a] = d tcoef_a / d nu.  P <= 1024; kappa, nu: device scalars. */
/* The t grid itself from a DEVICE lengthscale: t = logspace(lo_exp + log10 kappa, hi_exp + log10 kappa, P) with
 * torch.logspace's float64 formula, tw = its trapezoid weights (kernels_hyperbolic.py:367-369, kernels_spd.py:262-265,
 * kernels_euclidean.py:186-188 build it on the host from lengthscale.item()).  With this the whole hyper-parameter ->
 * node-weight chain of the quadrature kernels runs without a host synchronisation (CUDA-graph capturable). */
int mgb_matern_t_grid(const double* kappa, double lo_exp, double hi_exp, int P, double* t, double* tw, void* stream);
int mgb_matern_t_weights(const double* t, const double* tw, const double* h0, int P, double power_offset,
                         const double* kappa, const double* nu, double* tcoef, double* jac, void* stream);

/* The same single-CTA evaluation for a Gram matrix ANY kernel produced (quadrature kernels, torus, products):
 *   Kt = hyper[0] * K + hyper[1] * I,  out[0] = nll as above, out[1] = dnll/ds, out[2] = dnll/dnoise, out[3] = dnll/dmean,
 *   G (n x n, symmetric) = dnll/dK = s/2 (Kt^-1 - alpha alpha^T): the upstream gradient for the kernel's own backward.
 * Replaces the Cholesky / solve / log-determinant ops of gpytorch's ExactMarginalLogLikelihood and their autograd
 * backward (~60 launches) by one launch; n <= mgb_series_mll_max_n(0, 0). */
int mgb_dense_mll(const double* K, long long n, long long ldk, const double* y, const double* hyper, double* out,
                  double* G, long long ldg, int* status, void* stream);

/* Posterior fit for launch-bound training sizes with the same single-CTA kernel: rinv = L^-1 (n x n row-major, lower
 * triangle, zeros above), rinv_t = its transpose, alpha = Kt^-1 (y - mean), *nll as out[0] of mgb_series_mll.  Replaces
 * the Gram + Cholesky + solves of ExactGPPosterior.fit (gpytorch's prediction-strategy caches) at n <= mgb_series_mll_max_n. */
int mgb_series_fit(const mgb_series_desc* desc, const double* x, long long n, long long ld, const double* y,
                   const double* coef, const double* hyper, double* rinv, double* rinv_t, long long ldo, double* alpha,
                   double* nll, int* status, void* stream);

/* mgb_series_fit for a Gram matrix any kernel produced: Kt = hyper[0] * K + hyper[1] * I (n <= mgb_series_mll_max_n(0, 0)). */
int mgb_dense_fit(const double* K, long long n, long long ldk, const double* y, const double* hyper, double* rinv,
                  double* rinv_t, long long ldo, double* alpha, double* nll, int* status, void* stream);

/* Pair scalars of the quadrature kernels for a candidate block (m rows) against the training set (n rows), row-major
 * m x n: kind 0 Lorentz distance (rows `width` = dim + 1 wide), 1 squared Euclidean distance, 2 / 3 SPD(2) (Delta -> out0,
 * r2 -> out1) from the matrices / from their Cholesky factors (the geometry of mgb_hyperbolic_gram_fwd,
 * mgb_euclidean_gram_fwd, mgb_spd2_gram_fwd as a stand-alone step for the acquisition path). */
int mgb_pair_scalars(int kind, int width, const double* x1, long long m, long long ld1, const double* x2, long long n,
                     long long ld2, double* out0, double* out1, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Opt-in INT8 tensor-core path for the posterior's N^2 contraction (csrc/int8_posterior.cu): every row of K* and of
 * L^-1 is scaled by a power of two and cut into seven signed 7-bit slices; the 28 slice products with p + q <= 6 are
 * seven int8 GEMMs (exact int32 accumulation), recombined in float64.  Same quantities as mgb_posterior_reduce, to
 * ~1e-11 of max var at n = 5000 (the FP64 kernel: ~1e-14).  The GEMM is a library kernel (CUTLASS headers at build
 * time); mgb_int8_available() == 0 when the library was built without them.
 *   pack:    rinv (n x n, lower triangular L^-1, row-major, ld) -> packed (mgb_posterior_int8_pack_bytes(n) bytes)
 *   reduce:  kstar (m x n float64 kernel rows, row-major, ldk), packed, alpha -> mean[m], var[m];
 *            workspace of mgb_posterior_int8_workspace_bytes(m, n) bytes (caller-owned).
 * ------------------------------------------------------------------------------------------------ */
int mgb_int8_available(void);
long long mgb_posterior_int8_pack_bytes(long long n);
int mgb_posterior_int8_pack(const double* rinv, long long n, long long ld, void* packed, void* stream);
long long mgb_posterior_int8_workspace_bytes(long long m, long long n);
int mgb_posterior_int8_reduce(const double* kstar, long long m, long long n, long long ldk, const void* packed,
                              const double* alpha, double kss, double mean_const, double* mean, double* var,
                              void* workspace, long long workspace_bytes, void* stream);

/* The same contraction in ONE hand-written tcgen05 kernel (csrc/int8_fused.cu): tcgen05.mma kind::i8 on 128-candidate x
 * 64-column tiles with the seven int32 planes in tensor memory, operands pre-tiled in the tensor core's shared-memory
 * image and staged with cp.async.bulk, exact triangular k range, float64 recombination and the sum of squares in the
 * epilogue (no int32 plane reaches HBM).  Same semantics as the mgb_posterior_int8_* calls with 8-bit instead of 7-bit
 * slices (a carry pass keeps the digits in int8): ~2e-13 of max var at n = 5000 instead of ~2e-11, n <= 16384 for exact
 * int32 accumulation; needs no CUTLASS.
 * The _debug form additionally dumps the raw accumulators, planes[row block][column tile][7][128][64] int32
 * (mgb_posterior_i8f_planes_bytes(m, n) bytes), for the tests. */
long long mgb_posterior_i8f_pack_bytes(long long n);
int mgb_posterior_i8f_pack(const double* rinv, long long n, long long ld, void* packed, void* stream);
long long mgb_posterior_i8f_workspace_bytes(long long m, long long n);
int mgb_posterior_i8f_reduce(const double* kstar, long long m, long long n, long long ldk, const void* packed,
                             const double* alpha, double kss, double mean_const, double* mean, double* var,
                             void* workspace, long long workspace_bytes, void* stream);
/* Series kernels end to end on that kernel: candidates -> kernel rows -> slices -> contraction for one chunk of m
 * candidates (the INT8 counterpart of mgb_series_posterior): Gram kernel into the workspace, slicing, contraction;
 * workspace of mgb_series_posterior_i8f_workspace_bytes(m, n) bytes. */
long long mgb_series_posterior_i8f_workspace_bytes(long long m, long long n);
int mgb_series_posterior_i8f(const mgb_series_desc* desc, const double* xc, long long m, long long ldc, const double* xtrain,
                             long long n, long long ldt, const double* coef, const void* packed, const double* alpha,
                             double kss, double mean_const, double* mean, double* var, void* workspace,
                             long long workspace_bytes, void* stream);
long long mgb_posterior_i8f_planes_bytes(long long m, long long n);
int mgb_posterior_i8f_reduce_debug(const double* kstar, long long m, long long n, long long ldk, const void* packed,
                                   const double* alpha, double kss, double mean_const, double* mean, double* var,
                                   void* workspace, long long workspace_bytes, int* planes, void* stream);

/* Series Gram matrix with the inner products on the INT8 tensor cores (csrc/series_i8.cu; same contract and the same
 * reference lines as mgb_series_gram_fwd: kernel_utils/kernels_sphere.py:102-139,188-224,343-384, kernels_so.py:305-362
 * with dim = 3).  Single-factor monomial series with <= 12 terms and rows <= 16 wide (mgb_series_gram_i8_supported):
 * every row of x1 / x2 is cut into seven signed 8-bit digits (error-free up to 2^-57 of the row's largest entry), the
 * seven int32 planes of digit products come out of tcgen05.mma kind::i8 into tensor memory, the epilogue recombines them
 * exactly in 64-bit integers and rounds ONCE to float64 (the inner product is at least as accurate as the float64
 * kernel's FMA chain), then scale / shift, the reference's clamp and Horner run on the FP64 pipe as before.
 * Rows with a NaN / Inf / |x| >= 1e300 entry give NaN rows / columns.  workspace: 256-byte aligned,
 * >= mgb_series_gram_i8_workspace_bytes(n) bytes (the stacked operand image of x2); K: ldk >= n. */
int mgb_series_gram_i8_supported(const mgb_series_desc* desc);
long long mgb_series_gram_i8_workspace_bytes(long long n);
int mgb_series_gram_fwd_i8(const mgb_series_desc* desc, const double* x1, long long m, long long ld1, const double* x2,
                           long long n, long long ld2, const double* coef, double* K, long long ldk, void* workspace,
                           long long workspace_bytes, void* stream);
/* test hook: also writes the raw accumulator planes [row panel][column tile][7][128][32] (int32; n <= 4096) */
int mgb_series_gram_fwd_i8_debug(const mgb_series_desc* desc, const double* x1, long long m, long long ld1, const double* x2,
                                 long long n, long long ld2, const double* coef, double* K, long long ldk, void* workspace,
                                 long long workspace_bytes, int* planes, void* stream);

/* Analytic EI from posterior means / variances (botorch ExpectedImprovement: sigma = sqrt(max(var, 1e-9)),
 * u = +-(mean - best_f) / sigma, EI = sigma (phi(u) + u Phi(u))), elementwise over m candidates. */
int mgb_expected_improvement(const double* mean, const double* var, long long m, double best_f, int maximize,
                             double* ei, void* stream);

/* The same for the radial quadrature kernels (geom 0: hyperbolic H^1..H^3, rows dim + 1 wide; geom 1: Euclidean):
 * P0, P1, P2 (b x n each) are the profile and its first two derivatives at the pair scalars (geodesic distance or
 * |x - y|^2), as returned by mgb_radial_profile; this entry point adds the pair geometry, the triangular mat-vecs and
 * the chain through EI.  P2, v, hvp may be NULL (value + gradient only). */
int mgb_radial_acquisition_ei(int geom, int width, const double* xc, long long b, long long ldc, const double* v,
                              long long ldv, const double* xtrain, long long n, long long ldt, const double* P0,
                              const double* P1, const double* P2, double outputscale, const double* rinv,
                              long long ldr, const double* rinv_t, long long ldrt, const double* alpha, double kss,
                              double mean_const, double best_f, int maximize, double* ei, double* grad, double* hvp,
                              void* stream);

/* SPD(2) (Mandel 3-vectors): value + gradient (the reference's SPD trust region uses a finite-difference Hessian,
 * bo_spd.py:383).  P0, Pd, Pr (b x n each): the profile and its derivatives in Delta and r2 at the pairs, from
 * mgb_spd2_profile (which = 0, 1, 2); use_cholesky as in mgb_spd2_gram_fwd.  grad (b x 3) may be NULL. */
int mgb_spd2_acquisition_ei(int use_cholesky, const double* xc, long long b, long long ldc, const double* xtrain,
                            long long n, long long ldt, const double* P0, const double* Pd, const double* Pr,
                            double outputscale, const double* rinv, long long ldr, const double* rinv_t, long long ldrt,
                            const double* alpha, double kss, double mean_const, double best_f, int maximize, double* ei,
                            double* grad, void* stream);

/* The whole trust-region step of the quadrature kernels in ONE call: pair scalars (mgb_pair_scalars) -> profile and
 * its derivatives (mgb_radial_profile / mgb_spd2_profile) -> acquisition kernel, enqueued back to back on `stream`
 * with the b x n intermediates in caller-owned scratch `work` (work_doubles >= 4 * b * n for the radial kernels,
 * >= 5 * b * n for SPD(2)).  Same results as the three-step sequence; what it removes is the host time between the
 * launches (five binding calls and allocations per step at the reference's b = 1 ... 5 restarts, n <= a few hundred).
 * kind: 0/1/2 = H^1/H^2/H^3 (rows kind + 2 wide), 3 = Euclidean (rows `width` wide). */
int mgb_radial_acquisition_ei_step(int kind, int width, const double* xc, long long b, long long ldc, const double* v,
                                   long long ldv, const double* xtrain, long long n, long long ldt, const double* tval,
                                   const double* tcoef, int Pt, double s0, double ds, int Ps, double outputscale,
                                   const double* rinv, long long ldr, const double* rinv_t, long long ldrt,
                                   const double* alpha, double kss, double mean_const, double best_f, int maximize,
                                   double* work, long long work_doubles, double* ei, double* grad, double* hvp,
                                   void* stream);
int mgb_spd2_acquisition_ei_step(int use_cholesky, const double* xc, long long b, long long ldc, const double* xtrain,
                                 long long n, long long ldt, const double* tcoef, const double* rscale,
                                 const double* bscale, int Pt, const double* bval, const double* bw, int Pb,
                                 double outputscale, const double* rinv, long long ldr, const double* rinv_t,
                                 long long ldrt, const double* alpha, double kss, double mean_const, double best_f,
                                 int maximize, double* work, long long work_doubles, double* ei, double* grad,
                                 void* stream);

/* Series coefficients of the sphere / SO(3) kernels and their Jacobian with respect to (kappa, nu) in one launch
 * (the hyper-parameter -> coefficient chain of kernels_sphere.py:126-139,212-224 and kernels_so.py:264-273,101, which
 * torch autograd otherwise spreads over ~60 tiny kernels per marginal-likelihood evaluation):
 *   a_n = c[n] f(lam[n]),  f = ((b + lam)/b)^-(nu + half_dim), b = 2 nu / kappa^2 (mode 0) or exp(-kappa^2 lam / 2) (mode 1)
 *   coef = (a / sum_n a_n g[n]) M,  M row-major n_terms x n_out;  jac[0..n_out) = d coef / d kappa, jac[n_out..) = d coef / d nu.
 * kappa, nu: device scalars (nu may be NULL in mode 1). */
int mgb_spectral_coef(int mode, int n_terms, int n_out, const double* lam, const double* c, const double* g,
                      const double* M, double half_dim, const double* kappa, const double* nu, double* coef,
                      double* jac, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer entry points (the reference-facing calls: numpy / CPU tensors in, numpy out).
 * ------------------------------------------------------------------------------------------------ */
int mgb_series_gram_host(const mgb_series_desc* desc, const double* x1, long long m, const double* x2,
                         long long n, const double* coef, double* K, int device);
/* Rinv: n x n row-major (lower triangle used), alpha: n.  Candidates stream through pinned staging
 * buffers in blocks of `chunk` rows; copies overlap the kernels on a second stream. */
int mgb_series_posterior_host(const mgb_series_desc* desc, const double* xc, long long m,
                              const double* xtrain, long long n, const double* coef, const double* Rinv,
                              const double* alpha, double kss_value, double mean_const, double* mean,
                              double* var, long long chunk, int device);

/* ------------------------------------------------------------------------------------------------
 * Persistent posterior handle.  The reference fits once per BO iteration and then calls acq(X) hundreds of times
 * (manifold_optimization/manifold_optimize.py:229-236 trust-region restarts, :329-339 initial-condition scoring;
 * gpytorch keeps the Cholesky / root-inverse caches of its prediction strategy between those calls).  The handle is
 * that cache: the fitted state is uploaded and packed ONCE, and every evaluation reuses the handle's device buffers,
 * workspace, pinned staging block, streams and events - no allocation and no factor upload per call.
 *   mgb_posterior_create             host pointers (Rinv: n x n row-major, lower triangle used; alpha: n); synchronises
 *   mgb_posterior_create_from_device device pointers that the CALLER keeps alive (e.g. the packed factor received
 *                                    over NCCL / NVLink from the rank that did the fit: no PCIe upload per rank);
 *                                    packed = output of mgb_posterior_pack
 *   mgb_posterior_eval_host          host candidates (m x D, row-major, dense) -> host mean / var; synchronous.
 *                                    m <= 8192: one pinned staging block, one stream, one synchronisation (latency);
 *                                    larger: blocks of mgb_posterior_chunk(h) rows, uploads / kernels / downloads on
 *                                    three streams (throughput; pinned user buffers make the copies asynchronous)
 *   mgb_posterior_eval_device        device candidates -> device mean / var on `stream`; never allocates / synchronises
 *   mgb_posterior_destroy            releases everything the handle owns (NULL is allowed)
 * max_chunk <= 0: default 37888 candidates per block.  A handle is bound to `device` and is not thread-safe.
 * ------------------------------------------------------------------------------------------------ */
typedef struct mgb_posterior mgb_posterior;
int mgb_posterior_create(const mgb_series_desc* desc, const double* xtrain, long long n, const double* coef,
                         const double* Rinv, const double* alpha, double kss_value, double mean_const,
                         long long max_chunk, int device, mgb_posterior** out);
int mgb_posterior_create_from_device(const mgb_series_desc* desc, const double* xtrain_dev, long long n, long long ldt,
                                     const double* coef_dev, const double* packed_dev, double kss_value,
                                     double mean_const, long long max_chunk, int device, mgb_posterior** out);
long long mgb_posterior_chunk(const mgb_posterior* h);
/* Opt-in INT8 tensor-core contraction for a handle (csrc/int8_fused.cu; ~2e-13 of max var instead of ~1e-14, ~2.7 x the
 * candidates/s; n <= 16384).  packed_i8_dev = the sliced factor written by mgb_posterior_i8f_pack, alpha_dev = n doubles; both on the
 * handle's device, borrowed, and must outlive the handle (or the next call).  Both null = back to the FP64 kernel.
 * A handle made by mgb_posterior_create with MGB_POSTERIOR_PATH=int8 in the environment slices its own copy of L^-1 and
 * starts in this mode (the same switch as the Python host's ExactGPPosterior). */
int mgb_posterior_use_int8(mgb_posterior* h, const void* packed_i8_dev, const double* alpha_dev);
int mgb_posterior_is_int8(const mgb_posterior* h);
int mgb_posterior_eval_host(mgb_posterior* h, const double* xc, long long m, double* mean, double* var);
int mgb_posterior_eval_device(mgb_posterior* h, const double* xc_dev, long long m, long long ldc, double* mean_dev,
                              double* var_dev, void* stream);
int mgb_posterior_destroy(mgb_posterior* h);

/* ------------------------------------------------------------------------------------------------
 * Measurement helpers (bench.py): register-resident FP64 peak micro-benchmarks, result in TFLOP/s.
 * kind 0 = DFMA chains, 1 = DMMA m8n8k4 chains, 2 = both interleaved with equal FMA counts (do the two
 * instruction kinds share one FP64 data path?).  Synchronises.
 * ------------------------------------------------------------------------------------------------ */
int mgb_fp64_peak(int kind, int repeats, double* tflops, double* elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* MGB_H_ */
