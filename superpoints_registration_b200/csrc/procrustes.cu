// procrustes.cu -- batched weighted Kabsch / Procrustes on sm_100a.
//
// Replaces compute_rigid_transform (reference: utils/se3_torch.py:109-163), which calls torch.svd on a
// (..., 3, 3) covariance (LAPACK on CPU, cuSOLVER on GPU) after several reductions launched from Python.
// Here: one CTA of 4 warps per pair accumulates the 16 weighted moments in fp64 (pivoted on the pair's
// first correspondence to avoid cancellation at KITTI-scale coordinates), warp 0 folds them in a fixed
// order and finishes the pair: covariance, 3x3 one-sided Jacobi SVD in fp64, R = V diag(1,1,d) U^T with
// d = sign(det(V U^T)) (the reference negates V[:,2] when det <= 0, :154-158), t = -R ca + cb.
#include "spr_common.cuh"

namespace spr {
namespace {

constexpr int kProcThreads = 128;
constexpr int kMoments = 16;  // W, a(3), b(3), a b^T (9)

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// One-sided (Hestenes) Jacobi SVD of a 3x3 matrix: A V = U S.  Columns of G = A V are made mutually
// orthogonal by plane rotations applied from the right; singular values are the column norms.
__device__ void svd3(const double A[3][3], double U[3][3], double S[3], double V[3][3]) {
  double G[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      G[i][j] = A[i][j];
      V[i][j] = i == j ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      double alpha = 0, beta = 0, gamma = 0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        alpha += G[i][p] * G[i][p];
        beta += G[i][q] * G[i][q];
        gamma += G[i][p] * G[i][q];
      }
      const double lim = 1e-30 + 1e-17 * sqrt(alpha * beta);
      if (fabs(gamma) <= lim) continue;
      off = fmax(off, fabs(gamma) / (sqrt(alpha * beta) + 1e-300));
      const double zeta = (beta - alpha) / (2.0 * gamma);
      const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double gp = G[i][p], gq = G[i][q];
        G[i][p] = c * gp - s * gq;
        G[i][q] = s * gp + c * gq;
        const double vp = V[i][p], vq = V[i][q];
        V[i][p] = c * vp - s * vq;
        V[i][q] = s * vp + c * vq;
      }
    }
    if (off < 1e-15) break;
  }
  // column norms, sort descending (selection on 3 elements, swapping columns of G and V)
  double nrm[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) nrm[j] = sqrt(G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j]);
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = a + 1; b < 3; ++b)
      if (nrm[b] > nrm[a]) {
        const double t = nrm[a];
        nrm[a] = nrm[b];
        nrm[b] = t;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          double x = G[i][a];
          G[i][a] = G[i][b];
          G[i][b] = x;
          x = V[i][a];
          V[i][a] = V[i][b];
          V[i][b] = x;
        }
      }
  S[0] = nrm[0];
  S[1] = nrm[1];
  S[2] = nrm[2];
  // U columns; rank-deficient directions are completed to an orthonormal frame
  const double tiny = 1e-14 * (nrm[0] > 0 ? nrm[0] : 1.0);
  if (nrm[0] > tiny) {
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][0] = G[i][0] / nrm[0];
  } else {
    U[0][0] = 1;
    U[1][0] = 0;
    U[2][0] = 0;
  }
  if (nrm[1] > tiny) {
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][1] = G[i][1] / nrm[1];
  } else {
    // any unit vector orthogonal to U[:,0]
    const int m = fabs(U[0][0]) < fabs(U[1][0]) ? (fabs(U[0][0]) < fabs(U[2][0]) ? 0 : 2)
                                                : (fabs(U[1][0]) < fabs(U[2][0]) ? 1 : 2);
    double e[3] = {0, 0, 0};
    e[m] = 1.0;
    const double d = U[m][0];
    double w[3] = {e[0] - d * U[0][0], e[1] - d * U[1][0], e[2] - d * U[2][0]};
    const double wn = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][1] = w[i] / wn;
  }
  if (nrm[2] > tiny) {
#pragma unroll
    for (int i = 0; i < 3; ++i) U[i][2] = G[i][2] / nrm[2];
  } else {
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
  }
}

__device__ __forceinline__ double det3(const double M[3][3]) {
  return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
         M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

__global__ void __launch_bounds__(kProcThreads)
    k_procrustes(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ w,
                 const int* __restrict__ offs, float* __restrict__ out) {
  const int p = blockIdx.x;
  const int beg = offs[p], end = offs[p + 1];
  const int n = end - beg;
  __shared__ double s_part[kProcThreads / 32][kMoments];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double m[kMoments];
#pragma unroll
  for (int i = 0; i < kMoments; ++i) m[i] = 0.0;
  double pa[3] = {0, 0, 0}, pb[3] = {0, 0, 0};
  if (n > 0) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      pa[i] = (double)a[3 * (size_t)beg + i];
      pb[i] = (double)b[3 * (size_t)beg + i];
    }
  }
  for (int i = beg + (int)threadIdx.x; i < end; i += kProcThreads) {
    const double wi = w ? (double)w[i] : 1.0;
    const double ax = (double)a[3 * (size_t)i] - pa[0], ay = (double)a[3 * (size_t)i + 1] - pa[1],
                 az = (double)a[3 * (size_t)i + 2] - pa[2];
    const double bx = (double)b[3 * (size_t)i] - pb[0], by = (double)b[3 * (size_t)i + 1] - pb[1],
                 bz = (double)b[3 * (size_t)i + 2] - pb[2];
    m[0] += wi;
    m[1] += wi * ax;
    m[2] += wi * ay;
    m[3] += wi * az;
    m[4] += wi * bx;
    m[5] += wi * by;
    m[6] += wi * bz;
    const double wax = wi * ax, way = wi * ay, waz = wi * az;
    m[7] += wax * bx;
    m[8] += wax * by;
    m[9] += wax * bz;
    m[10] += way * bx;
    m[11] += way * by;
    m[12] += way * bz;
    m[13] += waz * bx;
    m[14] += waz * by;
    m[15] += waz * bz;
  }
#pragma unroll
  for (int i = 0; i < kMoments; ++i) {
    m[i] = warp_sum_d(m[i]);
    if (lane == 0) s_part[warp][i] = m[i];
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
#pragma unroll
  for (int i = 0; i < kMoments; ++i) {
    double t = s_part[0][i];
    for (int ww = 1; ww < kProcThreads / 32; ++ww) t += s_part[ww][i];
    m[i] = t;
  }
  // weights_normalized = w / clamp_min(sum w, 1e-6)  (se3_torch.py:137-138); weights None -> 1/N (torch.mean :146)
  const double Wn = w ? fmax(m[0], 1e-6) : (n > 0 ? (double)n : 1.0);
  const double sfrac = m[0] / Wn;  // = sum of normalised weights (1 unless sum w < 1e-6)
  double ca[3] = {m[1] / Wn, m[2] / Wn, m[3] / Wn};  // centroids of the PIVOTED points
  double cb[3] = {m[4] / Wn, m[5] / Wn, m[6] / Wn};
  // cov = sum w~ (a - ca_full)(b - cb_full)^T, with ca_full = ca + sfrac'*pa ... expanded on pivoted points:
  // a - ca_full = (a' + pa) - (ca' + sfrac*pa) ; when sfrac == 1 the pivot cancels exactly.  For the degenerate
  // sfrac < 1 case (all weights ~ 0) we reproduce the reference expression literally.
  double cov[3][3];
  if (sfrac > 1.0 - 1e-12 || !w) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) cov[i][j] = m[7 + 3 * i + j] / Wn - ca[i] * cb[j];
    if (!w) {
      // reference's unweighted branch (:145-150) does not divide the covariance by N
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) cov[i][j] *= (double)n;
    }
  } else {
    // centroid_full = ca' + sfrac*pa ; x - centroid_full = x' + (1 - sfrac) * pa
    const double r = 1.0 - sfrac;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const double ai = r * pa[i] - ca[i], bj = r * pb[j] - cb[j];  // constant offsets added to a', b'
        // sum w~ (a'_i + ai)(b'_j + bj) = M_ij/Wn + ai*cb'_j + ca'_i*bj + ai*bj*sfrac
        cov[i][j] = m[7 + 3 * i + j] / Wn + ai * cb[j] + ca[i] * bj + ai * bj * sfrac;
      }
  }
  double U[3][3], S[3], V[3][3];
  svd3(cov, U, S, V);
  // R = V U^T, flip the third column of V when det <= 0
  double R[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) R[i][j] = V[i][0] * U[j][0] + V[i][1] * U[j][1] + V[i][2] * U[j][2];
  if (!(det3(R) > 0.0)) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) R[i][j] = V[i][0] * U[j][0] + V[i][1] * U[j][1] - V[i][2] * U[j][2];
  }
  // full centroids (se3_torch.py:139-140): sum w~ x = ca' + sfrac * pivot
  const double caf[3] = {ca[0] + sfrac * pa[0], ca[1] + sfrac * pa[1], ca[2] + sfrac * pa[2]};
  const double cbf[3] = {cb[0] + sfrac * pb[0], cb[1] + sfrac * pb[1], cb[2] + sfrac * pb[2]};
  float* o = out + 12 * (size_t)p;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double t = -(R[i][0] * caf[0] + R[i][1] * caf[1] + R[i][2] * caf[2]) + cbf[i];
    o[4 * i + 0] = (float)R[i][0];
    o[4 * i + 1] = (float)R[i][1];
    o[4 * i + 2] = (float)R[i][2];
    o[4 * i + 3] = (float)t;
  }
}

}  // namespace
}  // namespace spr

using namespace spr;

extern "C" int spr_weighted_procrustes(const float* d_a, const float* d_b, const float* d_w, const int32_t* d_offsets,
                                       int n_pairs, float* d_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(n_pairs > 0, "weighted_procrustes: no pairs");
  SPR_CHECK_ARG(d_a && d_b && d_offsets && d_out, "weighted_procrustes: null pointer");
  k_procrustes<<<n_pairs, kProcThreads, 0, stream>>>(d_a, d_b, d_w, d_offsets, d_out);
  SPR_LAUNCH_CHECK("k_procrustes");
  return SPR_OK;
}
