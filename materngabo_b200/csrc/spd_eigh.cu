// Batched eigen-decomposition front-end of the SPD product kernels (spd_input=True).
//
// Reference: Riemannian_utils/spd_utils_torch.py:293-333 (spd_to_so_and_euclidean_torch /
// spd_to_sphere_and_euclidean_torch) as called by kernel_utils/kernels_spd.py:520-533,746-757:
//   Mandel vector -> symmetric matrix (spd_utils_torch.py:336-371) -> torch.symeig -> Q = det(V) V, eigenvalues ->
//   [eigenvalues | vec(Q)]  (R^d x SO(d) kernels)   or   [eigenvalues | sphere point of Q]  (R^d x S kernels; SO(2) -> S^1
//   = first row of Q, SO(3) -> S^3 = quaternion by Riemannian_utils/sphere_utils_torch.py:127-190).
// torch.symeig no longer exists (torch >= 1.13), and LAPACK fixes neither the sign of an eigenvector nor anything else
// the factor kernels would be invariant under, so the reference output cannot be reproduced bit for bit by ANY solver.
// This kernel states its convention instead:
//   - cyclic Jacobi rotations in registers, one thread per matrix (d = 2: one rotation, exact; d = 3: sweeps until the
//     off-diagonal mass is below 1e-32 of the diagonal, at most 12) - orthogonality and reconstruction at 1e-15;
//   - eigenvalues ascending (symeig's order), eigenvectors as columns;
//   - each eigenvector's sign is chosen so that its largest-magnitude component is positive (first such component);
//   - then Q = det(V) V as in the reference (:311).
#include "mgb_common.cuh"
#include <cmath>

namespace mgb {

template <int D>
__device__ __forceinline__ void jacobi_rotate(double (&a)[D][D], double (&v)[D][D], int p, int q) {
  const double apq = a[p][q];
  if (apq == 0.0) return;
  const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
  const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));   // smaller root: |angle| <= pi/4
  const double c = rsqrt(t * t + 1.0), s = t * c;
  const double app = a[p][p], aqq = a[q][q];
  a[p][p] = app - t * apq;
  a[q][q] = aqq + t * apq;
  a[p][q] = a[q][p] = 0.0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    if (k != p && k != q) {
      const double akp = a[k][p], akq = a[k][q];
      a[k][p] = a[p][k] = c * akp - s * akq;
      a[k][q] = a[q][k] = s * akp + c * akq;
    }
    const double vkp = v[k][p], vkq = v[k][q];
    v[k][p] = c * vkp - s * vkq;
    v[k][q] = s * vkp + c * vkq;
  }
}

// quaternion of a 3 x 3 rotation, operation for operation as Riemannian_utils/sphere_utils_torch.py:141-185
__device__ __forceinline__ void rotation_to_quaternion(const double (&R)[3][3], double* q) {
  const double qs = fmin(sqrt(R[0][0] + R[1][1] + R[2][2] + 1.0) / 2.0, 1.0);
  const double kx = R[2][1] - R[1][2], ky = R[0][2] - R[2][0], kz = R[1][0] - R[0][1];
  double kx1, ky1, kz1;
  bool add;
  if (R[0][0] >= R[1][1] && R[0][0] >= R[2][2]) {
    kx1 = R[0][0] - R[1][1] - R[2][2] + 1.0; ky1 = R[1][0] + R[0][1]; kz1 = R[2][0] + R[0][2];
    add = kx >= 0.0;
  } else if (R[1][1] >= R[2][2]) {
    kx1 = R[1][0] + R[0][1]; ky1 = R[1][1] - R[0][0] - R[2][2] + 1.0; kz1 = R[2][1] + R[1][2];
    add = ky >= 0.0;
  } else {
    kx1 = R[2][0] + R[0][2]; ky1 = R[2][1] + R[1][2]; kz1 = R[2][2] - R[0][0] - R[1][1] + 1.0;
    add = kz >= 0.0;
  }
  const double kxi = add ? kx + kx1 : kx - kx1, kyi = add ? ky + ky1 : ky - ky1, kzi = add ? kz + kz1 : kz - kz1;
  const double nm = sqrt(kxi * kxi + kyi * kyi + kzi * kzi);
  q[0] = qs;
  q[1] = q[2] = q[3] = 0.0;
  if (nm != 0.0) {
    const double s = sqrt(1.0 - qs * qs) / nm;
    q[1] = s * kxi; q[2] = s * kyi; q[3] = s * kzi;
  }
}

template <int D>
__global__ void __launch_bounds__(128) spd_decompose_kernel(const double* __restrict__ x, long long m, long long ld, int target,
                                                            double* __restrict__ out, long long ldo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double* vin = x + i * ld;
  const double r2 = 0.70710678118654752440;
  double a[D][D], v[D][D];
  // Mandel: diagonal first, then the k-th super-diagonals scaled by sqrt 2 (spd_utils_torch.py:355-365)
  if constexpr (D == 2) {
    a[0][0] = vin[0]; a[1][1] = vin[1]; a[0][1] = a[1][0] = vin[2] * r2;
  } else {
    a[0][0] = vin[0]; a[1][1] = vin[1]; a[2][2] = vin[2];
    a[0][1] = a[1][0] = vin[3] * r2; a[1][2] = a[2][1] = vin[4] * r2; a[0][2] = a[2][0] = vin[5] * r2;
  }
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) v[r][c] = (r == c) ? 1.0 : 0.0;
  if constexpr (D == 2) {
    jacobi_rotate<2>(a, v, 0, 1);
  } else {
    for (int sweep = 0; sweep < 12; ++sweep) {
      const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
      const double dg = a[0][0] * a[0][0] + a[1][1] * a[1][1] + a[2][2] * a[2][2];
      if (off <= 1e-32 * dg) break;
      jacobi_rotate<3>(a, v, 0, 1);
      jacobi_rotate<3>(a, v, 0, 2);
      jacobi_rotate<3>(a, v, 1, 2);
    }
  }
  // ascending eigenvalues (selection sort on D <= 3 columns)
  double ev[D];
  int ord[D];
#pragma unroll
  for (int k = 0; k < D; ++k) { ev[k] = a[k][k]; ord[k] = k; }
#pragma unroll
  for (int s = 0; s < D - 1; ++s)
#pragma unroll
    for (int t = s + 1; t < D; ++t)
      if (ev[t] < ev[s]) {
        const double te = ev[s]; ev[s] = ev[t]; ev[t] = te;
        const int to = ord[s]; ord[s] = ord[t]; ord[t] = to;
      }
  double Q[D][D];
#pragma unroll
  for (int c = 0; c < D; ++c) {
    double col[D];
#pragma unroll
    for (int r = 0; r < D; ++r) {
      double val = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) val = (ord[c] == k) ? v[r][k] : val;     // static indexing only
      col[r] = val;
    }
    int big = 0;
#pragma unroll
    for (int r = 1; r < D; ++r) if (fabs(col[r]) > fabs(col[big])) big = r;
    double pivot = col[0];
#pragma unroll
    for (int r = 1; r < D; ++r) pivot = (big == r) ? col[r] : pivot;
    const double sg = (pivot < 0.0) ? -1.0 : 1.0;
#pragma unroll
    for (int r = 0; r < D; ++r) Q[r][c] = sg * col[r];
  }
  double det;
  if constexpr (D == 2) det = Q[0][0] * Q[1][1] - Q[0][1] * Q[1][0];
  else det = Q[0][0] * (Q[1][1] * Q[2][2] - Q[1][2] * Q[2][1]) - Q[0][1] * (Q[1][0] * Q[2][2] - Q[1][2] * Q[2][0]) +
             Q[0][2] * (Q[1][0] * Q[2][1] - Q[1][1] * Q[2][0]);
#pragma unroll
  for (int r = 0; r < D; ++r)
#pragma unroll
    for (int c = 0; c < D; ++c) Q[r][c] *= det;        // x_so = det(V) V  (spd_utils_torch.py:309-311)
  double* o = out + i * ldo;
#pragma unroll
  for (int k = 0; k < D; ++k) o[k] = ev[k];
  if (target == 0) {
#pragma unroll
    for (int r = 0; r < D; ++r)
#pragma unroll
      for (int c = 0; c < D; ++c) o[D + r * D + c] = Q[r][c];
  } else if constexpr (D == 2) {
    o[2] = Q[0][0];          // first row of the rotation (sphere_utils_torch.py:118-123)
    o[3] = Q[0][1];
  } else {
    rotation_to_quaternion(Q, o + 3);
  }
}

}  // namespace mgb

using namespace mgb;

extern "C" int mgb_spd_decompose(int dim, int target, const double* x, long long m, long long ld, double* out,
                                 long long ldo, void* stream) {
  if (dim != 2 && dim != 3) return fail(MGB_ERR_ARG, "spd_decompose: dim must be 2 or 3");
  if (target != 0 && target != 1) return fail(MGB_ERR_ARG, "spd_decompose: target must be 0 (rotation) or 1 (sphere point)");
  const int width = dim + (target == 0 ? dim * dim : (dim == 2 ? 2 : 4));
  if (m < 0 || ld < dim * (dim + 1) / 2 || ldo < width) return fail(MGB_ERR_ARG, "spd_decompose: bad sizes / leading dimensions");
  if (m == 0) return MGB_OK;
  if (!x || !out) return fail(MGB_ERR_ARG, "spd_decompose: null pointer");
  const unsigned blocks = (unsigned)((m + 127) / 128);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dim == 2) spd_decompose_kernel<2><<<blocks, 128, 0, s>>>(x, m, ld, target, out, ldo);
  else spd_decompose_kernel<3><<<blocks, 128, 0, s>>>(x, m, ld, target, out, ldo);
  return after_launch("spd_decompose_kernel");
}
