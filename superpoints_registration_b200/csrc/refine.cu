// refine.cu -- the optional correspondence refinements of RegTR.softmax_correlation (all off in the shipped configs):
//   ratio test            models/qk_regtr_full.py:370-384
//   inlier re-weighting   models/qk_regtr_full.py:386-391  (recompute_weights; the loop of :393-398 alternates it with
//                                                          spr_weighted_procrustes)
//   RANSAC scoring        models/qk_regtr_full.py:400-421  (mean residual of every hypothesis, first strict minimum)
// All work on pairs packed back to back, one launch for the whole batch.
#include <cstdint>

#include "spr_common.cuh"

namespace spr {
namespace {

constexpr unsigned kFull = 0xffffffffu;

struct Top2 {
  float v1, v2;  // largest, second largest
  int i1;        // position of the largest (lowest position among equals)
};

__device__ __forceinline__ void top2_push(Top2& t, float v, int i) {
  if (v > t.v1 || (v == t.v1 && i < t.i1)) {
    t.v2 = t.v1;
    t.v1 = v;
    t.i1 = i;
  } else if (v > t.v2) {
    t.v2 = v;
  }
}

__device__ __forceinline__ void top2_merge(Top2& t, float v1, float v2, int i1) {
  top2_push(t, v1, i1);
  if (v2 > t.v2) t.v2 = v2;  // v2 <= v1 and v1 is already in: it can only be the runner-up
}

// One warp per output element.  Pair p reduces over the source axis (one output per target) when N_p > M_p, over the
// target axis otherwise, like the argmax of the matching kernel.  val = v1 if v2 / v1 < thres else 0 (a NaN ratio
// fails the comparison, as in torch.where).
__global__ void __launch_bounds__(256)
    k_top2_ratio(const float* __restrict__ attn, const long long* __restrict__ co, const int* __restrict__ so,
                 const int* __restrict__ to, const int* __restrict__ oo, int P, int total_out, float thres,
                 float* __restrict__ val, long long* __restrict__ ind) {
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= total_out) return;
  int lo = 0, hi = P - 1;  // last pair with oo[p] <= o and a non-empty output range
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(oo + mid) <= o) lo = mid; else hi = mid - 1;
  }
  const int p = lo;
  const int n = so[p + 1] - so[p], m = to[p + 1] - to[p];
  const int k = o - oo[p];
  const float* a = attn + co[p];
  Top2 t{-INFINITY, -INFINITY, 0x7fffffff};
  if (n > m) {
    for (int r = lane; r < n; r += 32) top2_push(t, __ldg(a + (size_t)r * m + k), r);
  } else {
    for (int c = lane; c < m; c += 32) top2_push(t, __ldg(a + (size_t)k * m + c), c);
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    const float v1 = __shfl_xor_sync(kFull, t.v1, s), v2 = __shfl_xor_sync(kFull, t.v2, s);
    const int i1 = __shfl_xor_sync(kFull, t.i1, s);
    top2_merge(t, v1, v2, i1);
  }
  if (lane == 0) {
    val[o] = (t.v2 / t.v1 < thres) ? t.v1 : 0.f;
    ind[o] = t.i1;
  }
}

__device__ __forceinline__ int find_segment(const int* __restrict__ offs, int P, int row) {
  int lo = 0, hi = P - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(offs + mid) <= row) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ float residual(const float* __restrict__ T, const float* __restrict__ a,
                                          const float* __restrict__ b) {
  // se3_transform (R a + t) and the Euclidean norm, fp32 like the reference (utils/se3_torch.py:50-66)
  const float ax = a[0], ay = a[1], az = a[2];
  const float x = T[0] * ax + T[1] * ay + T[2] * az + T[3];
  const float y = T[4] * ax + T[5] * ay + T[6] * az + T[7];
  const float z = T[8] * ax + T[9] * ay + T[10] * az + T[11];
  const float dx = b[0] - x, dy = b[1] - y, dz = b[2] - z;
  return sqrtf(dx * dx + dy * dy + dz * dz);
}

__global__ void __launch_bounds__(256)
    k_inlier_reweight(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ w,
                      const float* __restrict__ poses, const int* __restrict__ offs, int P, int total, float radius,
                      float* __restrict__ w_out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= total) return;
  const int p = find_segment(offs, P, r);
  const float res = residual(poses + 12 * (size_t)p, a + 3 * (size_t)r, b + 3 * (size_t)r);
  w_out[r] = w[r] * (res < radius ? 1.f : 0.f);
}

// loss[p][h] = mean over the rows of pair p of |b - T_{p,h} a|; one block per (hypothesis, pair), fixed-order folds
__global__ void __launch_bounds__(256)
    k_hypothesis_loss(const float* __restrict__ a, const float* __restrict__ b, const int* __restrict__ offs,
                      const float* __restrict__ poses, int n_hyp, float* __restrict__ loss) {
  const int h = blockIdx.x, p = blockIdx.y;
  const int r0 = offs[p], r1 = offs[p + 1];
  __shared__ float sT[12];
  __shared__ float sred[8];
  if (threadIdx.x < 12) sT[threadIdx.x] = poses[12 * ((size_t)p * n_hyp + h) + threadIdx.x];
  __syncthreads();
  float acc = 0.f;
  for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) acc += residual(sT, a + 3 * (size_t)r, b + 3 * (size_t)r);
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(kFull, acc, s);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sred[i];
    loss[(size_t)p * n_hyp + h] = t / (float)(r1 - r0);  // 0/0 = NaN for an empty pair, as torch.mean
  }
}

// the first hypothesis whose loss is strictly below every earlier one (:415-419); NaN never wins, hypothesis 0 is the
// fallback
__global__ void k_select_hypothesis(const float* __restrict__ loss, const float* __restrict__ poses, int n_hyp, int P,
                                    float* __restrict__ out, int* __restrict__ best) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  int bi = 0;
  float bl = loss[(size_t)p * n_hyp];
  for (int h = 1; h < n_hyp; ++h) {
    const float l = loss[(size_t)p * n_hyp + h];
    if (l < bl) {
      bl = l;
      bi = h;
    }
  }
  for (int i = 0; i < 12; ++i) out[12 * (size_t)p + i] = poses[12 * ((size_t)p * n_hyp + bi) + i];
  if (best) best[p] = bi;
}

}  // namespace
}  // namespace spr

using namespace spr;

extern "C" int spr_top2_ratio(const float* d_attn, const int64_t* d_corr_offsets, const int32_t* d_src_offsets,
                              const int32_t* d_tgt_offsets, const int32_t* d_out_offsets, int n_pairs, int total_out,
                              float lowe_thres, float* d_val, int64_t* d_ind, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(n_pairs > 0 && total_out > 0, "top2_ratio: empty input");
  SPR_CHECK_ARG(d_attn && d_corr_offsets && d_src_offsets && d_tgt_offsets && d_out_offsets && d_val && d_ind,
                "top2_ratio: null pointer");
  k_top2_ratio<<<(total_out + 7) / 8, 256, 0, stream>>>(d_attn, reinterpret_cast<const long long*>(d_corr_offsets),
                                                        d_src_offsets, d_tgt_offsets, d_out_offsets, n_pairs, total_out,
                                                        lowe_thres, d_val, reinterpret_cast<long long*>(d_ind));
  SPR_LAUNCH_CHECK("k_top2_ratio");
  return SPR_OK;
}

extern "C" int spr_inlier_reweight(const float* d_a, const float* d_b, const float* d_w, const float* d_poses,
                                   const int32_t* d_offsets, int n_pairs, int total, float acceptance_radius,
                                   float* d_w_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(n_pairs > 0 && total > 0, "inlier_reweight: empty input");
  SPR_CHECK_ARG(d_a && d_b && d_w && d_poses && d_offsets && d_w_out, "inlier_reweight: null pointer");
  k_inlier_reweight<<<(total + 255) / 256, 256, 0, stream>>>(d_a, d_b, d_w, d_poses, d_offsets, n_pairs, total,
                                                             acceptance_radius, d_w_out);
  SPR_LAUNCH_CHECK("k_inlier_reweight");
  return SPR_OK;
}

extern "C" int spr_select_hypothesis(const float* d_a, const float* d_b, const int32_t* d_offsets, int n_pairs,
                                     const float* d_poses, int n_hypotheses, float* d_loss, float* d_out,
                                     int32_t* d_best, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(n_pairs > 0 && n_hypotheses > 0, "select_hypothesis: empty input");
  SPR_CHECK_ARG(n_pairs <= 65535, "select_hypothesis: at most 65535 pairs per call");
  SPR_CHECK_ARG(d_a && d_b && d_offsets && d_poses && d_loss && d_out, "select_hypothesis: null pointer");
  dim3 grid(n_hypotheses, n_pairs);
  k_hypothesis_loss<<<grid, 256, 0, stream>>>(d_a, d_b, d_offsets, d_poses, n_hypotheses, d_loss);
  SPR_LAUNCH_CHECK("k_hypothesis_loss");
  k_select_hypothesis<<<(n_pairs + 63) / 64, 64, 0, stream>>>(d_loss, d_poses, n_hypotheses, n_pairs, d_out, d_best);
  SPR_LAUNCH_CHECK("k_select_hypothesis");
  return SPR_OK;
}
