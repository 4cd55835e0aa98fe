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

#include <cuda_fp16.h>

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

// part[chunk][c] = (sum, sumsq) over the chunk's rows, fp64.  A thread owns 4 consecutive channels (16-byte loads),
// c/4 threads cover a row, 256/(c/4) rows are read per step with the whole chunk's loads issued up front; the rows
// of a thread are folded in a fixed order, then the row lanes are folded in a fixed order through shared memory.
__global__ void __launch_bounds__(256) k_in_partial(const float* __restrict__ x, const int* __restrict__ lens, int B,
                                                    int c, double2* __restrict__ part) {
  __shared__ int s_loc[3];
  __shared__ double s_red[256 * 8];
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
  const int c4 = c >> 2;
  const int cw = c4 < 256 ? c4 : 256;   // channel-quad lanes
  const int rl = 256 / cw;              // row lanes
  const int tc = threadIdx.x % cw, tr = threadIdx.x / cw;
  for (int q0 = 0; q0 < c4; q0 += cw) {
    const int qd = q0 + tc;
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (qd < c4 && tr < rl) {
      const float4* col = reinterpret_cast<const float4*>(x + (size_t)row0 * c) + qd;
      int r = tr;
      for (; r + 3 * rl < rows; r += 4 * rl) {  // four independent 16-byte loads in flight
        const float4 v0 = col[(size_t)r * c4], v1 = col[(size_t)(r + rl) * c4], v2 = col[(size_t)(r + 2 * rl) * c4],
                     v3 = col[(size_t)(r + 3 * rl) * c4];
        const float4 vv[4] = {v0, v1, v2, v3};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          a[0] += (double)vv[k].x; a[1] += (double)vv[k].x * vv[k].x;
          a[2] += (double)vv[k].y; a[3] += (double)vv[k].y * vv[k].y;
          a[4] += (double)vv[k].z; a[5] += (double)vv[k].z * vv[k].z;
          a[6] += (double)vv[k].w; a[7] += (double)vv[k].w * vv[k].w;
        }
      }
      for (; r < rows; r += rl) {
        const float4 v = col[(size_t)r * c4];
        a[0] += (double)v.x; a[1] += (double)v.x * v.x;
        a[2] += (double)v.y; a[3] += (double)v.y * v.y;
        a[4] += (double)v.z; a[5] += (double)v.z * v.z;
        a[6] += (double)v.w; a[7] += (double)v.w * v.w;
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s_red[k * 256 + threadIdx.x] = a[k];
    __syncthreads();
    if (tr == 0 && qd < c4) {
      for (int t = 1; t < rl; ++t)  // fixed order
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] += s_red[k * 256 + t * cw + tc];
      double2* dst = part + (size_t)blockIdx.x * c + qd * 4;
      dst[0] = make_double2(a[0], a[1]);
      dst[1] = make_double2(a[2], a[3]);
      dst[2] = make_double2(a[4], a[5]);
      dst[3] = make_double2(a[6], a[7]);
    }
    __syncthreads();
  }
}

// stats[b][c] = (mean, rstd): fold the cloud's chunks in order.
__global__ void __launch_bounds__(256) k_in_stats(const double2* __restrict__ part, const int* __restrict__ lens, int B,
                                                  int c, float eps, float2* __restrict__ stats,
                                                  int* __restrict__ offs_out) {
  const int b = blockIdx.y;
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  int chunk0 = 0, row0 = 0;
  for (int i = 0; i < b; ++i) {
    const int l = __ldg(lens + i);
    chunk0 += (l + kRowsPerChunk - 1) / kRowsPerChunk;
    row0 += l;
  }
  const int len = __ldg(lens + b);
  if (offs_out && ch == 0) {  // row offsets of the clouds for the apply kernel (no separate launch)
    offs_out[b] = row0;
    offs_out[b + 1] = row0 + len;
  }
  if (ch >= c) return;
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

// stats[b][c] from the producer's 16-row block sums (spr_gemm_tc / spr_kpconv_forward_prepared `stats16`): blocks that
// lie inside the cloud are taken from part16, the ragged rows at the cloud's two ends are read from x.  One block per
// (32 channels, cloud): 16 channel-pair lanes x 16 row lanes, row lanes folded in a fixed order.
__global__ void __launch_bounds__(256)
    k_in_stats16(const float2* __restrict__ part16, const float* __restrict__ x, const int* __restrict__ lens, int c,
                 float eps, float2* __restrict__ stats, int* __restrict__ offs_out) {
  const int b = blockIdx.y;
  const int cp = threadIdx.x & 15, rl = threadIdx.x >> 4;
  const int ch = blockIdx.x * 32 + cp * 2;
  int s = 0;
  for (int i = 0; i < b; ++i) s += __ldg(lens + i);
  const int e = s + __ldg(lens + b);
  if (blockIdx.x == 0 && threadIdx.x == 0) {  // row offsets of the clouds for the apply kernel (no separate launch)
    offs_out[b] = s;
    offs_out[b + 1] = e;
  }
  __shared__ double s_red[16][16][4];
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  if (ch < c && e > s) {
    const int fb0 = (s + 15) >> 4, fb1 = e >> 4;  // full blocks [fb0, fb1)
    int j = fb0 + rl;
    for (; j + 48 < fb1; j += 64) {               // four independent 16-byte loads in flight
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(part16 + (size_t)(j + 16 * u) * c + ch);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a[0] += (double)v[u].x; a[1] += (double)v[u].y; a[2] += (double)v[u].z; a[3] += (double)v[u].w;
      }
    }
    for (; j < fb1; j += 16) {
      const float4 v = *reinterpret_cast<const float4*>(part16 + (size_t)j * c + ch);
      a[0] += (double)v.x; a[1] += (double)v.y; a[2] += (double)v.z; a[3] += (double)v.w;
    }
    // ragged ends: [s, head_end) and [tail_begin, e); all of [s, e) when the cloud holds no full block
    const int head_end = fb0 < fb1 ? fb0 << 4 : e;
    const int tail_begin = fb0 < fb1 ? fb1 << 4 : e;
    for (int r = s + rl; r < head_end; r += 16) {
      const float2 v = *reinterpret_cast<const float2*>(x + (size_t)r * c + ch);
      a[0] += (double)v.x; a[1] += (double)v.x * v.x; a[2] += (double)v.y; a[3] += (double)v.y * v.y;
    }
    for (int r = tail_begin + rl; r < e; r += 16) {
      const float2 v = *reinterpret_cast<const float2*>(x + (size_t)r * c + ch);
      a[0] += (double)v.x; a[1] += (double)v.x * v.x; a[2] += (double)v.y; a[3] += (double)v.y * v.y;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) s_red[rl][cp][k] = a[k];
  __syncthreads();
  if (rl == 0 && ch < c) {
    for (int t = 1; t < 16; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] += s_red[t][cp][k];
    const double n = e > s ? (double)(e - s) : 1.0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const double mean = a[2 * h] / n;
      double var = a[2 * h + 1] / n - mean * mean;
      if (var < 0.0) var = 0.0;
      stats[(size_t)b * c + ch + h] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
  }
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

// sticky numeric flags of this translation unit (bit 0: an fp16 operand image overflowed), see spr_numeric_flags
__device__ unsigned int g_blocks_flags;

// byte offset of 16-byte chunk j of row r inside a (rows x 128 B) SWIZZLE_128B operand tile (tc05.cuh)
__device__ __forceinline__ uint32_t sw128_off(uint32_t r, uint32_t j) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((j ^ (r & 7u)) << 4);
}

// Normalise + residual + LeakyReLU with format-aware outputs, so that the consumer needs no conversion pass:
//   out_f32   plain rows (block outputs, residual inputs)
//   out_img   operand image of the next tensor-core GEMM (gemm_tc.cu A image, K = c, scaled by a_scale)
//   out_x16 / out_pts4 / amax_bits   pre-split feature rows, packed support points and max|y| of the next KPConv
//             (kpconv_tc.cu pre-pass outputs)
// A group of G = min(c/8, 32) lanes owns one row, a lane 8 consecutive channels per step of 8*G; STEPS = c / (8 G)
// (1 for c <= 256, where the smaller register footprint buys more rows in flight per SM).
template <int STEPS>
__global__ void __launch_bounds__(256, STEPS == 1 ? 6 : 4)
    k_in_apply_ex(const float* __restrict__ x, const int* __restrict__ offs, int B, int n, int c,
                  const float2* __restrict__ stats, float slope, const float* __restrict__ residual,
                  const float2* __restrict__ res_stats, float* __restrict__ out_f32,
                  unsigned char* __restrict__ out_img, float a_scale,
                  uint32_t* __restrict__ out_x16, float4* __restrict__ out_pts4, const float* __restrict__ s_pts,
                  unsigned int* __restrict__ amax_bits, int x16_planar) {
  const int lane = threadIdx.x & 31;
  const int G = c / 8 < 32 ? c / 8 : 32;
  const int rpw = 32 / G;
  const int gl = lane % G;
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rpw + lane / G;
  const bool live = row < n;
  const int KA = (c + 63) / 64;
  float rmax = 0.f, rsum = 0.f;
  float y[STEPS][8];
  // the first step's rows are requested before the (dependent-load) cloud search
  float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
  if (live) {
    p0 = *reinterpret_cast<const float4*>(x + (size_t)row * c + gl * 8);
    p1 = *reinterpret_cast<const float4*>(x + (size_t)row * c + gl * 8 + 4);
  }
  const int b = live ? find_cloud(offs, B, row) : 0;
#pragma unroll
  for (int st = 0; st < STEPS; ++st) {
    const int ch = (st * G + gl) * 8;
    if (live) {
      const float4 v0 = st == 0 ? p0 : *reinterpret_cast<const float4*>(x + (size_t)row * c + ch);
      const float4 v1 = st == 0 ? p1 : *reinterpret_cast<const float4*>(x + (size_t)row * c + ch + 4);
      const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      const float4* st4 = reinterpret_cast<const float4*>(stats + (size_t)b * c + ch);
      float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (residual) {
        const float4 r0 = *reinterpret_cast<const float4*>(residual + (size_t)row * c + ch);
        const float4 r1 = *reinterpret_cast<const float4*>(residual + (size_t)row * c + ch + 4);
        r[0] = r0.x; r[1] = r0.y; r[2] = r0.z; r[3] = r0.w; r[4] = r1.x; r[5] = r1.y; r[6] = r1.z; r[7] = r1.w;
        if (res_stats) {  // the residual is a raw producer output too: normalise it here (its own statistics, no
                          // activation) instead of in a pass of its own -- same roundings as the two-pass route
          const float4* rs4 = reinterpret_cast<const float4*>(res_stats + (size_t)b * c + ch);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 s2 = __ldg(rs4 + i);
            r[2 * i] = __fmul_rn(r[2 * i] - s2.x, s2.y);
            r[2 * i + 1] = __fmul_rn(r[2 * i + 1] - s2.z, s2.w);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 s2 = __ldg(st4 + i);  // (mean, rstd) of channels 2i, 2i+1
        float a = (v[2 * i] - s2.x) * s2.y + r[2 * i], bb = (v[2 * i + 1] - s2.z) * s2.w + r[2 * i + 1];
        a = a >= 0.f ? a : a * slope;
        bb = bb >= 0.f ? bb : bb * slope;
        y[st][2 * i] = a;
        y[st][2 * i + 1] = bb;
        rsum += a + bb;
        rmax = fmaxf(rmax, fmaxf(fabsf(a), fabsf(bb)));
      }
      if (out_f32) {
        *reinterpret_cast<float4*>(out_f32 + (size_t)row * c + ch) = make_float4(y[st][0], y[st][1], y[st][2], y[st][3]);
        *reinterpret_cast<float4*>(out_f32 + (size_t)row * c + ch + 4) = make_float4(y[st][4], y[st][5], y[st][6], y[st][7]);
      }
      if (out_img) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float y0 = y[st][2 * i] * a_scale, y1 = y[st][2 * i + 1] * a_scale;
          const __half2 hh = __floats2half2_rn(y0, y1);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
          hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
          lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        const int chunk = ch >> 3;
        unsigned char* blk = out_img + ((size_t)(row >> 6) * KA + (chunk >> 3)) * 16384;
        const uint32_t r2 = 2 * (row & 63);
        *reinterpret_cast<uint4*>(blk + sw128_off(r2, chunk & 7)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(blk + sw128_off(r2 + 1, chunk & 7)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        if (st == 0 && gl < KA * 8 - c / 8) {  // K is padded to whole 64-wide atoms: the padding chunks must be finite
          unsigned char* last = out_img + ((size_t)(row >> 6) * KA + (KA - 1)) * 16384;
          const int pc = (c / 8 + gl) & 7;
          *reinterpret_cast<uint4*>(last + sw128_off(r2, pc)) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(last + sw128_off(r2 + 1, pc)) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
  }
  // (rmax is still this lane's own maximum here; the comparison is false for NaN, so NaN raises the flag too)
  if (out_img && live && !(rmax * a_scale <= 65504.f)) atomicOr(&g_blocks_flags, SPR_FLAG_FP16_OVERFLOW);
  if (out_x16) {
    for (int o = G >> 1; o > 0; o >>= 1) {
      rsum += __shfl_xor_sync(kFull, rsum, o);
      rmax = fmaxf(rmax, __shfl_xor_sync(kFull, rmax, o));
    }
    const int e = scale_exp(rmax, 14);
    const float rs = pow2i(e);
    if (live) {
#pragma unroll
      for (int st = 0; st < STEPS; ++st) {
        const int ch = (st * G + gl) * 8;
        uint32_t o8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v = y[st][i] * rs;
          const __half hi = __float2half_rn(v);
          const __half lo = __float2half_rn(v - __half2float(hi));
          o8[i] = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
        }
        if (x16_planar) {
          // planar per group of 32 channels: [32 hi | 32 lo] fp16 (the gather kernel, kpconv_g.cu, copies the two
          // 64-byte halves into separate operand blocks); this lane's 8 channels are 16 bytes of each half
          unsigned char* grp = reinterpret_cast<unsigned char*>(out_x16 + (size_t)row * c) + (ch >> 5) * 128 + (ch & 31) * 2;
          uint32_t hh[4], ll[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            hh[i] = __byte_perm(o8[2 * i], o8[2 * i + 1], 0x5410);
            ll[i] = __byte_perm(o8[2 * i], o8[2 * i + 1], 0x7632);
          }
          *reinterpret_cast<uint4*>(grp) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          *reinterpret_cast<uint4*>(grp + 64) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
        } else {
          *reinterpret_cast<uint4*>(out_x16 + (size_t)row * c + ch) = make_uint4(o8[0], o8[1], o8[2], o8[3]);
          *reinterpret_cast<uint4*>(out_x16 + (size_t)row * c + ch + 4) = make_uint4(o8[4], o8[5], o8[6], o8[7]);
        }
      }
      if (gl == 0) {
        const float inv = pow2i(-e);
        out_pts4[row] = make_float4(s_pts[3 * (size_t)row], s_pts[3 * (size_t)row + 1], s_pts[3 * (size_t)row + 2],
                                    rsum > 0.f ? inv : -inv);
      }
    }
    rmax = warp_maxf(live ? rmax : 0.f);
    if (lane == 0 && __float_as_uint(rmax) > *reinterpret_cast<volatile unsigned int*>(amax_bits))
      atomicMax(amax_bits, __float_as_uint(rmax));
  }
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
    k_max_pool(const float* __restrict__ x, const IdxT* __restrict__ idx, int row_stride, int H, int nq, int ns, int c,
               float* __restrict__ out, const int* __restrict__ order) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int c4 = c >> 2;
  for (int i = blockIdx.x * warps + (threadIdx.x >> 5); i < nq; i += gridDim.x * warps) {
    // `order` (optional) walks the pooled points in cell order: the eight rows of a CTA are spatial neighbours, their
    // pooling neighbourhoods overlap and the 512-byte feature rows they gather hit L1 / L2
    const int n = order ? __ldg(order + i) : i;
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

unsigned int blocks_numeric_flags(bool reset) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_blocks_flags, sizeof(v)) != cudaSuccess) return 0;
  if (reset && v) {
    const unsigned int zero = 0;
    cudaMemcpyToSymbol(g_blocks_flags, &zero, sizeof(zero));
  }
  return v;
}

}  // namespace spr

using namespace spr;

static size_t in_chunks_upper(int n, int B) { return (size_t)(n + kRowsPerChunk - 1) / kRowsPerChunk + (size_t)B; }

extern "C" size_t spr_instance_norm_workspace_bytes(int n_rows, int n_clouds, int c) {
  if (n_rows < 0 || n_clouds < 0 || c < 0) return 0;
  size_t bytes = align_up(in_chunks_upper(n_rows, n_clouds) * (size_t)c * sizeof(double2), 256);
  bytes += 2 * align_up((size_t)n_clouds * c * sizeof(float2), 256);  // statistics of x and of a raw residual
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
  k_in_stats<<<gs, 256, 0, stream>>>(part, d_lengths, n_clouds, c, eps, stats, nullptr);
  SPR_LAUNCH_CHECK("k_in_stats");
  size_t work = (size_t)n * (c / 4);
  int blocks = (int)((work + 255) / 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  k_in_apply<<<blocks, 256, 0, stream>>>(d_x, offs, n_clouds, n, c, stats, slope, d_residual, d_out);
  SPR_LAUNCH_CHECK("k_in_apply");
  return SPR_OK;
}

extern "C" int spr_instance_norm_lrelu_ex(const float* d_x, const int32_t* d_lengths, int n, int n_clouds, int c,
                                          float eps, float slope, const float* d_residual, float* d_out_f32,
                                          void* d_out_img, float a_scale, void* d_out_x16, void* d_out_pts4,
                                          const float* d_points, void* d_amax, const float* d_stats16,
                                          const float* d_residual_stats16, int x16_planar, void* d_workspace,
                                          size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(n > 0 && n_clouds > 0 && c > 0, "instance_norm_ex: empty input (n=%d, clouds=%d, c=%d)", n, n_clouds, c);
  SPR_CHECK_ARG(c % 32 == 0 && c <= 1024 && (c <= 256 || c % 256 == 0),
                "instance_norm_ex: channel count %d must be a multiple of 32 up to 256, or 512, 768, 1024", c);
  SPR_CHECK_ARG(d_x && d_lengths && d_workspace, "instance_norm_ex: null pointer");
  SPR_CHECK_ARG(d_out_f32 || d_out_img || d_out_x16, "instance_norm_ex: no output requested");
  SPR_CHECK_ARG(!d_out_x16 || (d_out_pts4 && d_points && d_amax), "instance_norm_ex: KPConv outputs need pts4, points, amax");
  SPR_CHECK_ARG(!d_residual_stats16 || d_residual, "instance_norm_ex: residual statistics without a residual");
  if (workspace_bytes < spr_instance_norm_workspace_bytes(n, n_clouds, c)) {
    set_error("instance_norm_ex: workspace too small");
    return SPR_ENOSPACE;
  }
  Carver ws(d_workspace, workspace_bytes);
  const size_t chunks = in_chunks_upper(n, n_clouds);
  double2* part = ws.take<double2>(chunks * (size_t)c);
  float2* stats = ws.take<float2>((size_t)n_clouds * c);
  float2* res_stats = ws.take<float2>((size_t)n_clouds * c);
  int* offs = ws.take<int>((size_t)n_clouds + 1);
  // the statistics kernels also write the clouds' row offsets (offs) that the apply kernel looks rows up in
  if (d_stats16) {  // the producer already summed 16-row blocks: no pass over x for the statistics
    dim3 g16((c + 31) / 32, n_clouds);
    k_in_stats16<<<g16, 256, 0, stream>>>(reinterpret_cast<const float2*>(d_stats16), d_x, d_lengths, c, eps, stats,
                                          offs);
    SPR_LAUNCH_CHECK("k_in_stats16");
  } else {
    k_in_partial<<<(unsigned)chunks, 256, 0, stream>>>(d_x, d_lengths, n_clouds, c, part);
    SPR_LAUNCH_CHECK("k_in_partial");
    dim3 gs((c + 255) / 256, n_clouds);
    k_in_stats<<<gs, 256, 0, stream>>>(part, d_lengths, n_clouds, c, eps, stats, offs);
    SPR_LAUNCH_CHECK("k_in_stats");
  }
  if (d_residual_stats16) {  // raw residual (a producer's output with its 16-row block sums): normalised in the apply kernel
    dim3 g16((c + 31) / 32, n_clouds);
    k_in_stats16<<<g16, 256, 0, stream>>>(reinterpret_cast<const float2*>(d_residual_stats16), d_residual, d_lengths, c,
                                          eps, res_stats, offs);
    SPR_LAUNCH_CHECK("k_in_stats16 (residual)");
  }
  if (d_amax) SPR_CUDA(cudaMemsetAsync(d_amax, 0, sizeof(unsigned int), stream));
  const int G = c / 8 < 32 ? c / 8 : 32;
  const int rows_per_block = 8 * (32 / G);
  const int steps = c / (8 * G);
  const unsigned blocks = (unsigned)((n + rows_per_block - 1) / rows_per_block);
#define SPR_IN_APPLY(S)                                                                                              \
  k_in_apply_ex<S><<<blocks, 256, 0, stream>>>(d_x, offs, n_clouds, n, c, stats, slope, d_residual,                   \
                                               d_residual_stats16 ? res_stats : nullptr, d_out_f32,                  \
                                               static_cast<unsigned char*>(d_out_img), a_scale,                      \
                                               static_cast<uint32_t*>(d_out_x16), static_cast<float4*>(d_out_pts4),  \
                                               d_points, static_cast<unsigned int*>(d_amax), x16_planar)
  switch (steps) {
    case 1: SPR_IN_APPLY(1); break;
    case 2: SPR_IN_APPLY(2); break;
    case 3: SPR_IN_APPLY(3); break;
    default: SPR_IN_APPLY(4); break;
  }
#undef SPR_IN_APPLY
  SPR_LAUNCH_CHECK("k_in_apply_ex");
  return SPR_OK;
}

extern "C" int spr_max_pool(const float* d_x, const void* d_idx, int idx_is_64, int row_stride, int H, int nq, int ns,
                            int c, float* d_out, const int32_t* d_order, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(nq > 0 && ns > 0 && c > 0 && H > 0, "max_pool: empty input");
  SPR_CHECK_ARG(c % 4 == 0, "max_pool: channel count %d must be a multiple of 4", c);
  SPR_CHECK_ARG(row_stride >= H, "max_pool: row_stride < H");
  SPR_CHECK_ARG(d_x && d_idx && d_out, "max_pool: null pointer");
  int blocks = (nq + 7) / 8;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  if (idx_is_64)
    k_max_pool<long long><<<blocks, 256, 0, stream>>>(d_x, static_cast<const long long*>(d_idx), row_stride, H, nq, ns,
                                                      c, d_out, d_order);
  else
    k_max_pool<int><<<blocks, 256, 0, stream>>>(d_x, static_cast<const int*>(d_idx), row_stride, H, nq, ns, c, d_out,
                                                d_order);
  SPR_LAUNCH_CHECK("k_max_pool");
  return SPR_OK;
}
