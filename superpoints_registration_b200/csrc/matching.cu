// matching.cu -- superpoint matching on sm_100a: correlation, dual-softmax + argmax, Sinkhorn weighted targets.
//
// Reference: RegTR.softmax_correlation (models/qk_regtr_full.py:423-672), a Python loop over pairs (:445) of
// torch.matmul / softmax / max calls; utils/se3_torch.py:166-239 for the Sinkhorn branch.
// All pairs of the batch are processed by the same launches (blockIdx.y = pair) with packed inputs.
//
//   k_corr        corr = S T^T / sqrt(D)            fp32 tiled GEMM (64x64 tiles, 4x4 register blocks)
//   k_row_stats   per row   max_j, sum_j exp(.)     one warp per row, coalesced
//   k_col_stats   per column max_i, sum_i exp(.)    32 columns per CTA (lane = column), warps stride over rows
//   k_match_rows / k_match_cols   val, ind = max over the dual-softmax product along the reference's axis
//   k_attn        optional materialisation of the dual-softmax matrix (outputs['attn'], :295)
//   Sinkhorn: potentials form of se3_torch.py:166-202 -- log_alpha_ij = A_ij - u_i - v_j with the slack
//   row/column kept implicit (A = 0 there); each half-iteration is one pass over corr.
#include "spr_common.cuh"

#include <algorithm>

namespace spr {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__global__ void __launch_bounds__(256)
    k_corr(const float* __restrict__ src, const float* __restrict__ tgt, const int* __restrict__ so,
           const int* __restrict__ to, const long long* __restrict__ co, int D, float inv_sqrt_d,
           float* __restrict__ corr) {
  const int p = blockIdx.y;
  const int N = so[p + 1] - so[p], M = to[p + 1] - to[p];
  const int tiles_n = (M + TN - 1) / TN, tiles_m = (N + TM - 1) / TM;
  if ((int)blockIdx.x >= tiles_m * tiles_n) return;
  const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
  const float* S = src + (size_t)so[p] * D;
  const float* T = tgt + (size_t)to[p] * D;
  float* C = corr + co[p];
  __shared__ float sS[TK][TM + 4];
  __shared__ float sT[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  // loader: 64 rows x 16 k = 1024 floats per operand = 256 threads x float4 along k
  const int lr = tid / 4, lk = (tid % 4) * 4;
  for (int k0 = 0; k0 < D; k0 += TK) {
    float4 vs = make_float4(0.f, 0.f, 0.f, 0.f), vt = vs;
    const int gi = tm * TM + lr, gj = tn * TN + lr;
    if (gi < N) vs = __ldg(reinterpret_cast<const float4*>(S + (size_t)gi * D + k0 + lk));
    if (gj < M) vt = __ldg(reinterpret_cast<const float4*>(T + (size_t)gj * D + k0 + lk));
    sS[lk + 0][lr] = vs.x;
    sS[lk + 1][lr] = vs.y;
    sS[lk + 2][lr] = vs.z;
    sS[lk + 3][lr] = vs.w;
    sT[lk + 0][lr] = vt.x;
    sT[lk + 1][lr] = vt.y;
    sT[lk + 2][lr] = vt.z;
    sT[lk + 3][lr] = vt.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&sS[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sT[k][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gi = tm * TM + ty * 4 + i;
    if (gi >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gj = tn * TN + tx * 4 + j;
      if (gj < M) C[(size_t)gi * M + gj] = acc[i][j] * inv_sqrt_d;
    }
  }
}

// ---- generic row / column reductions over a per-pair matrix --------------------------------------------
// Element functor returns the value whose (max, sum exp) statistics are wanted.
struct CorrElem {
  __device__ __forceinline__ float operator()(float c, int, int) const { return c; }
};

// Sinkhorn element: A_ij - u_i - v_j with A = -(max(c, lo) - sp_alpha) / denom.  The model path (qk_regtr_full.py:
// 532-536) has lo = 0; a caller-supplied affinity matrix is the case lo = -inf, sp_alpha = 0, denom = -1 (A = c).
struct SinkElem {
  const float* u;
  const float* v;
  float sp_alpha, inv_denom, lo;
  __device__ __forceinline__ float operator()(float c, int gi, int gj) const {
    const float a = -(fmaxf(c, lo) - sp_alpha) * inv_denom;
    return a - u[gi] - v[gj];
  }
};


// Two-pass (max then sum) statistics per row: exact expf, matches torch.softmax / logsumexp numerics closely.
template <typename F>
__global__ void __launch_bounds__(256)
    k_row_stats(const float* __restrict__ corr, const int* __restrict__ so, const int* __restrict__ to,
                const long long* __restrict__ co, F f, float* __restrict__ rmax, float* __restrict__ rsum,
                float extra /* value of an implicit extra column per row (slack); -inf = none */, const float* extra_u) {
  const int p = blockIdx.y;
  const int N = so[p + 1] - so[p], M = to[p + 1] - to[p];
  const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const float* C = corr + co[p];
  for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < N; i += gridDim.x * warps) {
    const int gi = so[p] + i;
    const float* row = C + (size_t)i * M;
    float m = -INFINITY;
    for (int j = lane; j < M; j += 32) m = fmaxf(m, f(row[j], gi, to[p] + j));
    m = warp_maxf(m);
    float ex = -INFINITY;
    if (extra != -INFINITY) {
      ex = extra - (extra_u ? extra_u[gi] : 0.f);
      m = fmaxf(m, ex);
    }
    float s = 0.f;
    for (int j = lane; j < M; j += 32) s += expf(f(row[j], gi, to[p] + j) - m);
    s = warp_sum(s);
    if (ex != -INFINITY) s += expf(ex - m);
    if (lane == 0) {
      rmax[gi] = m;
      rsum[gi] = s;
    }
  }
}

// Per column: lane = column (coalesced across the warp), the CTA's warps stride over rows.
template <typename F>
__global__ void __launch_bounds__(256)
    k_col_stats(const float* __restrict__ corr, const int* __restrict__ so, const int* __restrict__ to,
                const long long* __restrict__ co, F f, float* __restrict__ cmax, float* __restrict__ csum, float extra,
                const float* extra_v) {
  const int p = blockIdx.y;
  const int N = so[p + 1] - so[p], M = to[p + 1] - to[p];
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  if ((int)blockIdx.x * 32 >= M) return;
  const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const float* C = corr + co[p];
  __shared__ float s_m[8][32], s_s[8][32];
  const int gj = to[p] + j;
  // pass 1: max
  float m = -INFINITY;
  if (j < M)
    for (int i = warp; i < N; i += warps) m = fmaxf(m, f(C[(size_t)i * M + j], so[p] + i, gj));
  s_m[warp][threadIdx.x & 31] = m;
  __syncthreads();
  m = s_m[0][threadIdx.x & 31];
  for (int w = 1; w < warps; ++w) m = fmaxf(m, s_m[w][threadIdx.x & 31]);
  float ex = -INFINITY;
  if (extra != -INFINITY && j < M) {
    ex = extra - (extra_v ? extra_v[gj] : 0.f);
    m = fmaxf(m, ex);
  }
  // pass 2: sum
  float s = 0.f;
  if (j < M)
    for (int i = warp; i < N; i += warps) s += expf(f(C[(size_t)i * M + j], so[p] + i, gj) - m);
  s_s[warp][threadIdx.x & 31] = s;
  __syncthreads();
  if (warp == 0 && j < M) {
    float t = 0.f;
    for (int w = 0; w < warps; ++w) t += s_s[w][threadIdx.x & 31];
    if (ex != -INFINITY) t += expf(ex - m);
    cmax[gj] = m;
    csum[gj] = t;
  }
}

// attn_ij = softmax over rows-of-a-column (dim=-2) * softmax over a row (dim=-1)   qk_regtr_full.py:457-459
__device__ __forceinline__ float dual_softmax(float c, float rm, float rs, float cm, float cs) {
  return (expf(c - cm) / cs) * (expf(c - rm) / rs);
}

// torch.max semantics for the arg-max (qk_regtr_full.py:468,576): NaN is the maximum, ties and NaNs keep the FIRST
// index, so the index is always in range (a row of NaNs must not leave a sentinel that a later gather dereferences).
__device__ __forceinline__ bool better(float a, int ai, float b, int bi) {
  const bool an = a != a, bn = b != b;
  if (an || bn) return an && (!bn || ai < bi);
  return a > b || (a == b && ai < bi);
}

// N <= M : val, ind = max(attn, dim=2): one warp per source row.
__global__ void __launch_bounds__(256)
    k_match_rows(const float* __restrict__ corr, const int* __restrict__ so, const int* __restrict__ to,
                 const long long* __restrict__ co, const int* __restrict__ oo, const float* __restrict__ rmax,
                 const float* __restrict__ rsum, const float* __restrict__ cmax, const float* __restrict__ csum,
                 float* __restrict__ val, long long* __restrict__ ind) {
  const int p = blockIdx.y;
  const int N = so[p + 1] - so[p], M = to[p + 1] - to[p];
  if (N > M) return;
  const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const float* C = corr + co[p];
  for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < N; i += gridDim.x * warps) {
    const int gi = so[p] + i;
    const float rm = rmax[gi], rs = rsum[gi];
    float best = -INFINITY;
    int bj = 0x7fffffff;
    for (int j = lane; j < M; j += 32) {
      const float a = dual_softmax(C[(size_t)i * M + j], rm, rs, cmax[to[p] + j], csum[to[p] + j]);
      if (better(a, j, best, bj)) {
        best = a;
        bj = j;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(kFull, best, o);
      const int oj = __shfl_xor_sync(kFull, bj, o);
      if (better(ob, oj, best, bj)) {
        best = ob;
        bj = oj;
      }
    }
    if (lane == 0) {
      val[oo[p] + i] = best;
      ind[oo[p] + i] = bj;
    }
  }
}

// N > M : val, ind = max(attn, dim=1): lane = target column, warps stride over source rows.
__global__ void __launch_bounds__(256)
    k_match_cols(const float* __restrict__ corr, const int* __restrict__ so, const int* __restrict__ to,
                 const long long* __restrict__ co, const int* __restrict__ oo, const float* __restrict__ rmax,
                 const float* __restrict__ rsum, const float* __restrict__ cmax, const float* __restrict__ csum,
                 float* __restrict__ val, long long* __restrict__ ind) {
  const int p = blockIdx.y;
  const int N = so[p + 1] - so[p], M = to[p + 1] - to[p];
  if (N <= M) return;
  if ((int)blockIdx.x * 32 >= M) return;
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const float* C = corr + co[p];
  __shared__ float s_b[8][32];
  __shared__ int s_i[8][32];
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (j < M) {
    const float cm = cmax[to[p] + j], cs = csum[to[p] + j];
    for (int i = warp; i < N; i += warps) {
      const float a = dual_softmax(C[(size_t)i * M + j], rmax[so[p] + i], rsum[so[p] + i], cm, cs);
      if (better(a, i, best, bi)) {
        best = a;
        bi = i;
      }
    }
  }
  s_b[warp][threadIdx.x & 31] = best;
  s_i[warp][threadIdx.x & 31] = bi;
  __syncthreads();
  if (warp == 0 && j < M) {
    for (int w = 1; w < warps; ++w) {
      const float ob = s_b[w][threadIdx.x & 31];
      const int oi = s_i[w][threadIdx.x & 31];
      if (better(ob, oi, best, bi)) {
        best = ob;
        bi = oi;
      }
    }
    val[oo[p] + j] = best;
    ind[oo[p] + j] = bi;
  }
}

__global__ void __launch_bounds__(256)
    k_attn(const float* __restrict__ corr, const int* __restrict__ so, const int* __restrict__ to,
           const long long* __restrict__ co, const float* __restrict__ rmax, const float* __restrict__ rsum,
           const float* __restrict__ cmax, const float* __restrict__ csum, float* __restrict__ attn) {
  const int p = blockIdx.y;
  const int N = so[p + 1] - so[p], M = to[p + 1] - to[p];
  const size_t total = (size_t)N * M;
  const float* C = corr + co[p];
  float* A = attn + co[p];
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / M), j = (int)(e % M);
    A[e] = dual_softmax(C[e], rmax[so[p] + i], rsum[so[p] + i], cmax[to[p] + j], csum[to[p] + j]);
  }
}

// u_i += logsumexp_j(...)   /   v_j += logsumexp_i(...)
__global__ void k_add_lse(float* __restrict__ pot, const float* __restrict__ mx, const float* __restrict__ sm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) pot[i] += mx[i] + logf(sm[i]);
}

// Final Sinkhorn pass: P_ij = exp(A_ij - u_i - v_j); w_i = sum_j P_ij; wt_i = (sum_j P_ij tgt_j) / (w_i + 1e-6)
__global__ void __launch_bounds__(256)
    k_sink_finish(const float* __restrict__ corr, const int* __restrict__ so, const int* __restrict__ to,
                  const long long* __restrict__ co, SinkElem f, const float* __restrict__ tgt_xyz,
                  float* __restrict__ wt, float* __restrict__ wsum) {
  const int p = blockIdx.y;
  const int N = so[p + 1] - so[p], M = to[p + 1] - to[p];
  const int lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const float* C = corr + co[p];
  for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < N; i += gridDim.x * warps) {
    const int gi = so[p] + i;
    float s = 0.f, x = 0.f, y = 0.f, z = 0.f;
    for (int j = lane; j < M; j += 32) {
      const int gj = to[p] + j;
      const float pij = expf(f(C[(size_t)i * M + j], gi, gj));
      s += pij;
      x = fmaf(pij, tgt_xyz[3 * (size_t)gj], x);
      y = fmaf(pij, tgt_xyz[3 * (size_t)gj + 1], y);
      z = fmaf(pij, tgt_xyz[3 * (size_t)gj + 2], z);
    }
    s = warp_sum(s);
    x = warp_sum(x);
    y = warp_sum(y);
    z = warp_sum(z);
    if (lane == 0) {
      const float d = s + 1e-6f;
      wsum[gi] = s;
      wt[3 * (size_t)gi] = x / d;
      wt[3 * (size_t)gi + 1] = y / d;
      wt[3 * (size_t)gi + 2] = z / d;
    }
  }
}

// log of the (near) doubly stochastic matrix, se3_torch.py:200: log_alpha_ij = A_ij - u_i - v_j
__global__ void __launch_bounds__(256)
    k_sink_log(const float* __restrict__ mat, const int* __restrict__ so, const int* __restrict__ to,
               const long long* __restrict__ co, SinkElem f, float* __restrict__ out) {
  const int p = blockIdx.y;
  const int N = so[p + 1] - so[p], M = to[p + 1] - to[p];
  const size_t total = (size_t)N * M;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / M), j = (int)(e % M);
    out[co[p] + e] = f(mat[co[p] + e], so[p] + i, to[p] + j);
  }
}

__global__ void k_gather_rows3(const float* __restrict__ src, int n_src, const long long* __restrict__ ind,
                               const int* __restrict__ base, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  long long r = (long long)base[i] + ind[i];
  r = r < 0 ? 0 : (r >= n_src ? n_src - 1 : r);  // never read outside src, whatever the index says
  out[3 * (size_t)i] = src[3 * r];
  out[3 * (size_t)i + 1] = src[3 * r + 1];
  out[3 * (size_t)i + 2] = src[3 * r + 2];
}

}  // namespace
}  // namespace spr

using namespace spr;

extern "C" size_t spr_match_workspace_bytes(int total_src, int total_tgt, int n_pairs) {
  (void)n_pairs;
  if (total_src < 0 || total_tgt < 0) return 0;
  return align_up((size_t)total_src * 8, 256) + align_up((size_t)total_tgt * 8, 256) + 512;
}

extern "C" int spr_dual_softmax_match(const float* d_src, const float* d_tgt, const int32_t* d_src_offsets,
                                      const int32_t* d_tgt_offsets, const int64_t* d_corr_offsets,
                                      const int32_t* d_out_offsets, int n_pairs, int total_src, int total_tgt, int D,
                                      int max_n, int max_m, float* d_corr, float* d_attn, float* d_val, int64_t* d_ind,
                                      void* d_workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(n_pairs > 0 && total_src > 0 && total_tgt > 0 && max_n > 0 && max_m > 0, "dual_softmax_match: empty input");
  SPR_CHECK_ARG(D > 0 && D % TK == 0, "dual_softmax_match: feature dim %d must be a multiple of %d", D, TK);
  SPR_CHECK_ARG(d_src && d_tgt && d_src_offsets && d_tgt_offsets && d_corr_offsets && d_out_offsets && d_corr && d_val &&
                    d_ind && d_workspace,
                "dual_softmax_match: null pointer");
  if (workspace_bytes < spr_match_workspace_bytes(total_src, total_tgt, n_pairs)) {
    set_error("dual_softmax_match: workspace too small");
    return SPR_ENOSPACE;
  }
  Carver ws(d_workspace, workspace_bytes);
  float* rmax = ws.take<float>((size_t)total_src * 2);
  float* rsum = rmax + total_src;
  float* cmax = ws.take<float>((size_t)total_tgt * 2);
  float* csum = cmax + total_tgt;
  const long long* co = reinterpret_cast<const long long*>(d_corr_offsets);
  long long* ind = reinterpret_cast<long long*>(d_ind);
  const int tiles = ((max_n + TM - 1) / TM) * ((max_m + TN - 1) / TN);
  k_corr<<<dim3(tiles, n_pairs), 256, 0, stream>>>(d_src, d_tgt, d_src_offsets, d_tgt_offsets, co, D,
                                                   1.0f / sqrtf((float)D), d_corr);
  SPR_LAUNCH_CHECK("k_corr");
  const int rb = min((max_n + 7) / 8, kNumSMs * 8);
  const int cb = (max_m + 31) / 32;
  k_row_stats<CorrElem><<<dim3(rb, n_pairs), 256, 0, stream>>>(d_corr, d_src_offsets, d_tgt_offsets, co, CorrElem(),
                                                               rmax, rsum, -INFINITY, nullptr);
  SPR_LAUNCH_CHECK("k_row_stats");
  k_col_stats<CorrElem><<<dim3(cb, n_pairs), 256, 0, stream>>>(d_corr, d_src_offsets, d_tgt_offsets, co, CorrElem(),
                                                               cmax, csum, -INFINITY, nullptr);
  SPR_LAUNCH_CHECK("k_col_stats");
  k_match_rows<<<dim3(rb, n_pairs), 256, 0, stream>>>(d_corr, d_src_offsets, d_tgt_offsets, co, d_out_offsets, rmax,
                                                      rsum, cmax, csum, d_val, ind);
  SPR_LAUNCH_CHECK("k_match_rows");
  k_match_cols<<<dim3(cb, n_pairs), 256, 0, stream>>>(d_corr, d_src_offsets, d_tgt_offsets, co, d_out_offsets, rmax,
                                                      rsum, cmax, csum, d_val, ind);
  SPR_LAUNCH_CHECK("k_match_cols");
  if (d_attn) {
    k_attn<<<dim3(kNumSMs * 2, n_pairs), 256, 0, stream>>>(d_corr, d_src_offsets, d_tgt_offsets, co, rmax, rsum, cmax,
                                                           csum, d_attn);
    SPR_LAUNCH_CHECK("k_attn");
  }
  return SPR_OK;
}

extern "C" size_t spr_sinkhorn_workspace_bytes(int total_src, int total_tgt, int n_pairs) {
  (void)n_pairs;
  if (total_src < 0 || total_tgt < 0) return 0;
  return align_up((size_t)total_src * 12, 256) + align_up((size_t)total_tgt * 12, 256) + 512;
}

static int sinkhorn_run(const float* d_mat, const int64_t* d_corr_offsets, const int32_t* d_src_offsets,
                        const int32_t* d_tgt_offsets, int n_pairs, int total_src, int total_tgt, int max_n, int max_m,
                        float sp_alpha, float inv_denom, float lo, int n_iters, const float* d_tgt_xyz,
                        float* d_weighted_tgt, float* d_weights, float* d_log_perm, void* d_workspace,
                        size_t workspace_bytes, cudaStream_t stream) {
  SPR_CHECK_ARG(n_pairs > 0 && total_src > 0 && total_tgt > 0 && max_n > 0 && max_m > 0, "sinkhorn: empty input");
  SPR_CHECK_ARG(n_iters >= 0, "sinkhorn: n_iters < 0");
  SPR_CHECK_ARG(d_mat && d_corr_offsets && d_src_offsets && d_tgt_offsets && d_workspace, "sinkhorn: null pointer");
  SPR_CHECK_ARG((d_weighted_tgt == nullptr) == (d_weights == nullptr) && (!d_weighted_tgt || d_tgt_xyz),
                "sinkhorn: weighted targets need tgt_xyz, weighted_tgt and weights together");
  SPR_CHECK_ARG(d_weighted_tgt || d_log_perm, "sinkhorn: no output requested");
  if (workspace_bytes < spr_sinkhorn_workspace_bytes(total_src, total_tgt, n_pairs)) {
    set_error("sinkhorn: workspace too small");
    return SPR_ENOSPACE;
  }
  Carver ws(d_workspace, workspace_bytes);
  float* u = ws.take<float>((size_t)total_src * 3);
  float* rmx = u + total_src;
  float* rsm = rmx + total_src;
  float* v = ws.take<float>((size_t)total_tgt * 3);
  float* cmx = v + total_tgt;
  float* csm = cmx + total_tgt;
  SPR_CUDA(cudaMemsetAsync(u, 0, (size_t)total_src * 4, stream));
  SPR_CUDA(cudaMemsetAsync(v, 0, (size_t)total_tgt * 4, stream));
  const long long* co = reinterpret_cast<const long long*>(d_corr_offsets);
  SinkElem f;
  f.u = u;
  f.v = v;
  f.sp_alpha = sp_alpha;
  f.inv_denom = inv_denom;
  f.lo = lo;
  const int rb = min((max_n + 7) / 8, kNumSMs * 8);
  const int cb = (max_m + 31) / 32;
  // With slack (se3_torch.py:183-199): rows 0..N-1 are normalised over M+1 columns (the slack column holds
  // 0 - u_i - v_M with v_M never updated, i.e. -u_i); columns 0..M-1 over N+1 rows (slack row holds -v_j).
  // Without slack the reference still zero-pads (:182-184), so the implicit entries are identical.
  for (int it = 0; it < n_iters; ++it) {
    k_row_stats<SinkElem><<<dim3(rb, n_pairs), 256, 0, stream>>>(d_mat, d_src_offsets, d_tgt_offsets, co, f, rmx, rsm,
                                                                 0.f, u);
    SPR_LAUNCH_CHECK("k_row_stats<sink>");
    k_add_lse<<<(total_src + 255) / 256, 256, 0, stream>>>(u, rmx, rsm, total_src);
    SPR_LAUNCH_CHECK("k_add_lse");
    k_col_stats<SinkElem><<<dim3(cb, n_pairs), 256, 0, stream>>>(d_mat, d_src_offsets, d_tgt_offsets, co, f, cmx, csm,
                                                                 0.f, v);
    SPR_LAUNCH_CHECK("k_col_stats<sink>");
    k_add_lse<<<(total_tgt + 255) / 256, 256, 0, stream>>>(v, cmx, csm, total_tgt);
    SPR_LAUNCH_CHECK("k_add_lse");
  }
  if (d_log_perm) {
    const long long per = (long long)max_n * max_m;
    const int lb = (int)std::min<long long>((per + 255) / 256, (long long)kNumSMs * 8);
    k_sink_log<<<dim3(lb, n_pairs), 256, 0, stream>>>(d_mat, d_src_offsets, d_tgt_offsets, co, f, d_log_perm);
    SPR_LAUNCH_CHECK("k_sink_log");
  }
  if (d_weighted_tgt) {
    k_sink_finish<<<dim3(rb, n_pairs), 256, 0, stream>>>(d_mat, d_src_offsets, d_tgt_offsets, co, f, d_tgt_xyz,
                                                         d_weighted_tgt, d_weights);
    SPR_LAUNCH_CHECK("k_sink_finish");
  }
  return SPR_OK;
}

extern "C" int spr_sinkhorn_weighted_targets(const float* d_corr, const int64_t* d_corr_offsets,
                                             const int32_t* d_src_offsets, const int32_t* d_tgt_offsets, int n_pairs,
                                             int total_src, int total_tgt, int max_n, int max_m,
                                             const float* d_tgt_xyz, float softplus_alpha, float exp_beta, int n_iters,
                                             int slack, float* d_weighted_tgt, float* d_weights, void* d_workspace,
                                             size_t workspace_bytes, void* stream_) {
  (void)slack;  // the reference zero-pads whether or not `slack` is set (se3_torch.py:182-184)
  SPR_CHECK_ARG(d_tgt_xyz && d_weighted_tgt && d_weights, "sinkhorn: null pointer");
  return sinkhorn_run(d_corr, d_corr_offsets, d_src_offsets, d_tgt_offsets, n_pairs, total_src, total_tgt, max_n, max_m,
                      softplus_alpha, 1.0f / (exp_beta + 0.02f), 0.f, n_iters, d_tgt_xyz, d_weighted_tgt, d_weights,
                      nullptr, d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream_));
}

extern "C" int spr_sinkhorn_affinity(const float* d_affinity, const int64_t* d_mat_offsets, const int32_t* d_src_offsets,
                                     const int32_t* d_tgt_offsets, int n_pairs, int total_src, int total_tgt, int max_n,
                                     int max_m, int n_iters, int slack, float* d_log_perm, const float* d_tgt_xyz,
                                     float* d_weighted_tgt, float* d_weights, void* d_workspace, size_t workspace_bytes,
                                     void* stream_) {
  (void)slack;
  return sinkhorn_run(d_affinity, d_mat_offsets, d_src_offsets, d_tgt_offsets, n_pairs, total_src, total_tgt, max_n,
                      max_m, 0.f, -1.0f, -INFINITY, n_iters, d_tgt_xyz, d_weighted_tgt, d_weights, d_log_perm,
                      d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream_));
}

extern "C" int spr_gather_rows3(const float* d_src, int n_src, const int64_t* d_ind, const int32_t* d_row_base,
                                int n_rows, float* d_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(n_rows > 0 && n_src > 0 && d_src && d_ind && d_row_base && d_out, "gather_rows3: bad argument");
  k_gather_rows3<<<(n_rows + 255) / 256, 256, 0, stream>>>(d_src, n_src, reinterpret_cast<const long long*>(d_ind),
                                                           d_row_base, n_rows, d_out);
  SPR_LAUNCH_CHECK("k_gather_rows3");
  return SPR_OK;
}
