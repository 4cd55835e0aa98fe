// blocks.cu -- the cheap epilogues between KPConv layers: per-cloud instance norm + LeakyReLU (+ residual),
// and the max-pool shortcut of strided blocks.
//
// Reference: BatchNormBlock.forward with nn.InstanceNorm1d (kpconv_blocks.py:474-530; a Python loop over
// clouds at :516-517, one InstanceNorm1d call per cloud), nn.LeakyReLU(0.1) (:556-561, :645, :727, :741),
// max_pool (:127-143).
//
// Instance norm is deterministic (no floating-point atomics): fixed row chunks per cloud produce fp64
// partial sums, a second kernel folds them in chunk order, a third normalises.
#include "spr_common.cuh"

namespace spr {
namespace {

constexpr int kRowsPerChunk = 128;

// chunk id -> (cloud, first row, row count). Chunks never straddle clouds.
__device__ __forceinline__ bool locate_chunk(const int* __restrict__ lens, int B, int chunk, int& cloud, int& row0,
                                             int& rows) {
  int acc_chunks = 0, acc_rows = 0;
  for (int b = 0; b < B; ++b) {
    const int len = __ldg(lens + b);
    const int nch = (len + kRowsPerChunk - 1) / kRowsPerChunk;
    if (chunk < acc_chunks + nch) {
      const int local = chunk - acc_chunks;
      cloud = b;
      row0 = acc_rows + local * kRowsPerChunk;
      rows = min(kRowsPerChunk, len - local * kRowsPerChunk);
      return true;
    }
    acc_chunks += nch;
    acc_rows += len;
  }
  return false;
}

// part[chunk][c] = (sum, sumsq) over the chunk's rows, fp64.
__global__ void __launch_bounds__(256) k_in_partial(const float* __restrict__ x, const int* __restrict__ lens, int B,
                                                    int c, double2* __restrict__ part) {
  __shared__ int s_loc[3];
  __shared__ double2 s_red[256];
  if (threadIdx.x == 0) {
    int cloud, row0, rows;
    if (!locate_chunk(lens, B, blockIdx.x, cloud, row0, rows)) rows = 0;
    s_loc[0] = cloud;
    s_loc[1] = row0;
    s_loc[2] = rows;
  }
  __syncthreads();
  const int row0 = s_loc[1], rows = s_loc[2];
  if (rows == 0) return;
  const int cw = c < 256 ? ((c + 31) / 32) * 32 : 256;  // channel lanes
  const int rl = 256 / cw;                              // row lanes
  const int tc = threadIdx.x % cw, tr = threadIdx.x / cw;
  for (int c0 = 0; c0 < c; c0 += cw) {
    const int ch = c0 + tc;
    double s1 = 0.0, s2 = 0.0;
    if (ch < c && tr < rl) {
      for (int r = tr; r < rows; r += rl) {
        const double v = (double)x[(size_t)(row0 + r) * c + ch];
        s1 += v;
        s2 += v * v;
      }
    }
    s_red[threadIdx.x] = make_double2(s1, s2);
    __syncthreads();
    if (tr == 0 && ch < c) {
      double a1 = s1, a2 = s2;
      for (int t = 1; t < rl; ++t) {  // fixed order
        a1 += s_red[t * cw + tc].x;
        a2 += s_red[t * cw + tc].y;
      }
      part[(size_t)blockIdx.x * c + ch] = make_double2(a1, a2);
    }
    __syncthreads();
  }
}

// stats[b][c] = (mean, rstd): fold the cloud's chunks in order.
__global__ void __launch_bounds__(256) k_in_stats(const double2* __restrict__ part, const int* __restrict__ lens, int B,
                                                  int c, float eps, float2* __restrict__ stats) {
  const int b = blockIdx.y;
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  int chunk0 = 0;
  for (int i = 0; i < b; ++i) chunk0 += (__ldg(lens + i) + kRowsPerChunk - 1) / kRowsPerChunk;
  const int len = __ldg(lens + b);
  const int nch = (len + kRowsPerChunk - 1) / kRowsPerChunk;
  double s1 = 0.0, s2 = 0.0;
  for (int k = 0; k < nch; ++k) {
    const double2 p = part[(size_t)(chunk0 + k) * c + ch];
    s1 += p.x;
    s2 += p.y;
  }
  const double n = len > 0 ? (double)len : 1.0;
  const double mean = s1 / n;
  double var = s2 / n - mean * mean;  // biased variance (InstanceNorm1d uses the batch statistics, unbiased=False)
  if (var < 0.0) var = 0.0;
  stats[(size_t)b * c + ch] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

__global__ void __launch_bounds__(256)
    k_in_apply(const float* __restrict__ x, const int* __restrict__ offs, int B, int n, int c,
               const float2* __restrict__ stats, float slope, const float* __restrict__ residual,
               float* __restrict__ out) {
  const int c4 = c >> 2;
  const size_t total = (size_t)n * c4;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int row = (int)(e / c4), cc = (int)(e % c4) * 4;
    const int b = find_cloud(offs, B, row);
    const float4 v = *reinterpret_cast<const float4*>(x + (size_t)row * c + cc);
    const float2* st = stats + (size_t)b * c + cc;
    const float2 s0 = __ldg(st), s1 = __ldg(st + 1), s2 = __ldg(st + 2), s3 = __ldg(st + 3);
    float4 y = make_float4((v.x - s0.x) * s0.y, (v.y - s1.x) * s1.y, (v.z - s2.x) * s2.y, (v.w - s3.x) * s3.y);
    if (residual) {
      const float4 r = *reinterpret_cast<const float4*>(residual + (size_t)row * c + cc);
      y.x += r.x;
      y.y += r.y;
      y.z += r.z;
      y.w += r.w;
    }
    y.x = y.x >= 0.f ? y.x : y.x * slope;
    y.y = y.y >= 0.f ? y.y : y.y * slope;
    y.z = y.z >= 0.f ? y.z : y.z * slope;
    y.w = y.w >= 0.f ? y.w : y.w * slope;
    *reinterpret_cast<float4*>(out + (size_t)row * c + cc) = y;
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
    k_max_pool(const float* __restrict__ x, const IdxT* __restrict__ idx, int row_stride, int H, int nq, int ns, int c,
               float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int c4 = c >> 2;
  for (int n = blockIdx.x * warps + (threadIdx.x >> 5); n < nq; n += gridDim.x * warps) {
    const IdxT* row = idx + (size_t)n * row_stride;
    for (int cc = lane; cc < c4; cc += 32) {
      float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      for (int h = 0; h < H; ++h) {
        const int j = (int)__ldg(row + h);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);  // the appended zero row (kpconv_blocks.py:135)
        if (j >= 0 && j < ns) v = __ldg(reinterpret_cast<const float4*>(x + (size_t)j * c) + cc);
        m.x = fmaxf(m.x, v.x);
        m.y = fmaxf(m.y, v.y);
        m.z = fmaxf(m.z, v.z);
        m.w = fmaxf(m.w, v.w);
      }
      reinterpret_cast<float4*>(out + (size_t)n * c)[cc] = m;
    }
  }
}

}  // namespace
}  // namespace spr

using namespace spr;

static size_t in_chunks_upper(int n, int B) { return (size_t)(n + kRowsPerChunk - 1) / kRowsPerChunk + (size_t)B; }

extern "C" size_t spr_instance_norm_workspace_bytes(int n_rows, int n_clouds, int c) {
  if (n_rows < 0 || n_clouds < 0 || c < 0) return 0;
  size_t bytes = align_up(in_chunks_upper(n_rows, n_clouds) * (size_t)c * sizeof(double2), 256);
  bytes += align_up((size_t)n_clouds * c * sizeof(float2), 256);
  bytes += align_up(((size_t)n_clouds + 1) * 4, 256);
  return bytes + 512;
}

extern "C" int spr_instance_norm_lrelu(const float* d_x, const int32_t* d_lengths, int n, int n_clouds, int c, float eps,
                                       float slope, const float* d_residual, float* d_out, void* d_workspace,
                                       size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(n > 0 && n_clouds > 0 && c > 0, "instance_norm: empty input (n=%d, clouds=%d, c=%d)", n, n_clouds, c);
  SPR_CHECK_ARG(c % 4 == 0, "instance_norm: channel count %d must be a multiple of 4", c);
  SPR_CHECK_ARG(d_x && d_lengths && d_out && d_workspace, "instance_norm: null pointer");
  if (workspace_bytes < spr_instance_norm_workspace_bytes(n, n_clouds, c)) {
    set_error("instance_norm: workspace too small");
    return SPR_ENOSPACE;
  }
  Carver ws(d_workspace, workspace_bytes);
  const size_t chunks = in_chunks_upper(n, n_clouds);
  double2* part = ws.take<double2>(chunks * (size_t)c);
  float2* stats = ws.take<float2>((size_t)n_clouds * c);
  int* offs = ws.take<int>((size_t)n_clouds + 1);
  int rc = cloud_offsets(d_lengths, n_clouds, offs, stream);
  if (rc) return rc;
  k_in_partial<<<(unsigned)chunks, 256, 0, stream>>>(d_x, d_lengths, n_clouds, c, part);
  SPR_LAUNCH_CHECK("k_in_partial");
  dim3 gs((c + 255) / 256, n_clouds);
  k_in_stats<<<gs, 256, 0, stream>>>(part, d_lengths, n_clouds, c, eps, stats);
  SPR_LAUNCH_CHECK("k_in_stats");
  size_t work = (size_t)n * (c / 4);
  int blocks = (int)((work + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  k_in_apply<<<blocks, 256, 0, stream>>>(d_x, offs, n_clouds, n, c, stats, slope, d_residual, d_out);
  SPR_LAUNCH_CHECK("k_in_apply");
  return SPR_OK;
}

extern "C" int spr_max_pool(const float* d_x, const void* d_idx, int idx_is_64, int row_stride, int H, int nq, int ns,
                            int c, float* d_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(nq > 0 && ns > 0 && c > 0 && H > 0, "max_pool: empty input");
  SPR_CHECK_ARG(c % 4 == 0, "max_pool: channel count %d must be a multiple of 4", c);
  SPR_CHECK_ARG(row_stride >= H, "max_pool: row_stride < H");
  SPR_CHECK_ARG(d_x && d_idx && d_out, "max_pool: null pointer");
  int blocks = (nq + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (idx_is_64)
    k_max_pool<long long><<<blocks, 256, 0, stream>>>(d_x, static_cast<const long long*>(d_idx), row_stride, H, nq, ns,
                                                      c, d_out);
  else
    k_max_pool<int><<<blocks, 256, 0, stream>>>(d_x, static_cast<const int*>(d_idx), row_stride, H, nq, ns, c, d_out);
  SPR_LAUNCH_CHECK("k_max_pool");
  return SPR_OK;
}
