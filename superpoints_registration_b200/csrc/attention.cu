// attention.cu -- variable-length multi-head attention for the cross-encoder (reference:
// models/transformer/transformers.py:184-245 via nn.MultiheadAttention, head_dim = d_model / nhead = 32).
//
//   O[q, h*32:(h+1)*32] = softmax_k( Q[q,h] . K[k,h] / sqrt(32) ) V[k,h]      for k in the key segment of q's segment
//
// One launch serves every (segment, head): self-attention of all clouds, or both directions of the cross-attention,
// with no padding -- the reference pads every cloud to the longest one and masks (transformers.py:18-59).
// Flash-attention dataflow on the warp-level tensor path: a CTA owns 64 query rows (4 warps x 16), streams the key /
// value rows of its segment through shared memory in tiles of 64 (cp.async double buffer), keeps S, P and O in
// registers with an online soft-max, and never materialises the N x M score matrix.
// fp32-level accuracy from fp16 tensor-core products: every operand is an fp16 (hi, lo) pair (x = hi + lo to ~22
// bits, produced by spr_split_f16 or by the GEMM epilogue) and each product is three MMAs hi*hi + lo*hi + hi*lo
// accumulated in fp32.  Q arrives pre-multiplied by log2(e)/sqrt(32), so the soft-max is a bare exp2.
#include "spr_common.cuh"

#include <cuda_fp16.h>

namespace spr {
namespace {

// sticky numeric flags of this translation unit (bit 0: an fp16 operand image overflowed), see spr_numeric_flags
__device__ unsigned int g_attention_flags;


constexpr int HD = 32;       // head dimension
constexpr int BK = 64;       // key rows per shared-memory tile
constexpr int PLANE_BYTES = BK * HD * 2;  // one fp16 plane of a tile: 4 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of 16-byte chunk c (0..3) of row r in a [64 rows x 64 B] tile; two rows share a 128-byte line and the
// chunk index is XOR-swizzled with the line index so that the 8 rows of one ldmatrix hit 8 different bank groups
__device__ __forceinline__ uint32_t tile_off(int r, int c) {
  return (uint32_t)((r >> 1) * 128 + (((r & 1) << 2) + (c ^ ((r >> 1) & 3))) * 16);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// d[16x8] += a[16x16] * b[16x8], fp16 operands, fp32 accumulate
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t h2u(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnTile {
  int q_row0;   // first query row (global row index into the planes)
  int q_rows;   // valid query rows in this tile (1..64)
  int kv_row0;  // first key/value row of the segment
  int kv_len;   // key/value rows of the segment
};

// planes: hi and lo fp16 matrices with `ld` halves per row; q_col / k_col / v_col = first column of head 0
__global__ void __launch_bounds__(128, 4)
    k_attention(const __half* __restrict__ hi, const __half* __restrict__ lo, int ld, int q_col, int k_col, int v_col,
                const AttnTile* __restrict__ tiles, float* __restrict__ out, int out_ld,
                unsigned char* __restrict__ out_img, int img_katoms, float img_scale) {
  __shared__ __align__(128) unsigned char smem[2 * 4 * PLANE_BYTES];  // 2 stages x {K hi, K lo, V hi, V lo}
  const AttnTile tl = tiles[blockIdx.x];
  if (tl.q_rows <= 0) return;  // padding entry of a bucketed tile list (CUDA-graph replays keep the grid fixed)
  const int head = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;

  // ---- Q fragments (A operand, 2 k-steps of 16 along d), hi and lo ----
  uint32_t qh[2][4], ql[2][4];
  {
    const int r0 = min(warp * 16 + g, tl.q_rows - 1), r1 = min(warp * 16 + g + 8, tl.q_rows - 1);
    const size_t o0 = (size_t)(tl.q_row0 + r0) * ld + q_col + head * HD;
    const size_t o1 = (size_t)(tl.q_row0 + r1) * ld + q_col + head * HD;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const int c = ks * 16 + 2 * t;
      qh[ks][0] = *reinterpret_cast<const uint32_t*>(hi + o0 + c);
      qh[ks][1] = *reinterpret_cast<const uint32_t*>(hi + o1 + c);
      qh[ks][2] = *reinterpret_cast<const uint32_t*>(hi + o0 + c + 8);
      qh[ks][3] = *reinterpret_cast<const uint32_t*>(hi + o1 + c + 8);
      ql[ks][0] = *reinterpret_cast<const uint32_t*>(lo + o0 + c);
      ql[ks][1] = *reinterpret_cast<const uint32_t*>(lo + o1 + c);
      ql[ks][2] = *reinterpret_cast<const uint32_t*>(lo + o0 + c + 8);
      ql[ks][3] = *reinterpret_cast<const uint32_t*>(lo + o1 + c + 8);
    }
  }

  const int n_kv_tiles = (tl.kv_len + BK - 1) / BK;
  const uint32_t sbase = smem_u32(smem);
  // stage a key/value tile: 4 planes x 64 rows x 4 chunks of 16 B = 1024 chunks, 8 per thread
  auto stage_tile = [&](int kt, int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int id = i * 128 + tid;
      const int plane = id >> 8;          // 0 K hi, 1 K lo, 2 V hi, 3 V lo
      const int r = (id >> 2) & 63, c = id & 3;
      const int kv = kt * BK + r;
      const bool ok = kv < tl.kv_len;
      const __half* src = ((plane & 1) ? lo : hi) + (size_t)(tl.kv_row0 + (ok ? kv : 0)) * ld +
                          ((plane & 2) ? v_col : k_col) + head * HD + c * 8;
      cp_async16(sbase + (buf * 4 + plane) * PLANE_BYTES + tile_off(r, c), src, ok);
    }
    cp_async_commit();
  };

  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  stage_tile(0, 0);
  for (int kt = 0; kt < n_kv_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_kv_tiles) {
      stage_tile(kt + 1, buf ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t kh_s = sbase + (buf * 4 + 0) * PLANE_BYTES, kl_s = kh_s + PLANE_BYTES;
    const uint32_t vh_s = kh_s + 2 * PLANE_BYTES, vl_s = kh_s + 3 * PLANE_BYTES;

    // ---- S = Q K^T : 8 key tiles of 8, 2 k-steps of 16 ----
    // Issue order is product-major over groups of 4 key tiles: consecutive MMAs write different accumulators, so
    // the tensor pipe is not stalled on its own result latency (6 chained MMAs per tile otherwise).
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) s[j][e] = 0.f;
#pragma unroll
    for (int j0 = 0; j0 < 8; j0 += 4) {
      uint32_t bh[4][4], bl[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        // matrices: (rows 8j.., d chunk 0), (chunk 1), (chunk 2), (chunk 3) -> b0,b1 of k-step 0 ; b0,b1 of k-step 1
        const uint32_t off = tile_off(8 * (j0 + j) + (lane & 7), lane >> 3);
        ldsm_x4(kh_s + off, bh[j]);
        ldsm_x4(kl_s + off, bl[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) mma16816(s[j0 + j], ql[0], bh[j][0], bh[j][1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma16816(s[j0 + j], qh[0], bl[j][0], bl[j][1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma16816(s[j0 + j], ql[1], bh[j][2], bh[j][3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma16816(s[j0 + j], qh[1], bl[j][2], bl[j][3]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma16816(s[j0 + j], qh[0], bh[j][0], bh[j][1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma16816(s[j0 + j], qh[1], bh[j][2], bh[j][3]);
    }
    // mask the tail of the segment
    if ((kt + 1) * BK > tl.kv_len) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = kt * BK + 8 * j + 2 * t;
        if (c >= tl.kv_len) s[j][0] = s[j][2] = -INFINITY;
        if (c + 1 >= tl.kv_len) s[j][1] = s[j][3] = -INFINITY;
      }
    }
    // ---- online soft-max (rows g and g+8; a row lives in the 4 lanes of a quad) ----
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(kFull, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(kFull, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(kFull, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(kFull, mx1, 2));
    const float a0 = ex2(m0 - mx0), a1 = ex2(m1 - mx1);  // first tile: exp2(-inf) = 0
    m0 = mx0;
    m1 = mx1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = ex2(s[j][0] - mx0);
      s[j][1] = ex2(s[j][1] - mx0);
      s[j][2] = ex2(s[j][2] - mx1);
      s[j][3] = ex2(s[j][3] - mx1);
      rs0 += s[j][0] + s[j][1];
      rs1 += s[j][2] + s[j][3];
    }
    l0 = l0 * a0 + rs0;
    l1 = l1 * a1 + rs1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[i][0] *= a0;
      o[i][1] *= a0;
      o[i][2] *= a1;
      o[i][3] *= a1;
    }
    // ---- O += P V : 4 k-steps of 16 keys, 4 d tiles of 8 ----
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ph[4], pl[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // a0 = S[2kk] c0,c1 ; a1 = S[2kk] c2,c3 ; a2 = S[2kk+1] c0,c1 ; a3 = S[2kk+1] c2,c3
        const float v0 = s[2 * kk + (e >> 1)][(e & 1) * 2], v1 = s[2 * kk + (e >> 1)][(e & 1) * 2 + 1];
        const __half2 hh = __floats2half2_rn(v0, v1);
        const float2 hf = __half22float2(hh);
        ph[e] = h2u(hh);
        pl[e] = h2u(__floats2half2_rn(v0 - hf.x, v1 - hf.y));
      }
      uint32_t bh[2][4], bl[2][4];
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {
        // matrices: (keys 16kk.., d chunk 2dp), (keys 16kk+8.., chunk 2dp), (keys 16kk.., chunk 2dp+1), (+8, 2dp+1)
        const int r = 16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8;
        const uint32_t off = tile_off(r, 2 * dp + (lane >> 4));
        ldsm_x4_trans(vh_s + off, bh[dp]);
        ldsm_x4_trans(vl_s + off, bl[dp]);
      }
      // product-major: the four d tiles are independent accumulators
#pragma unroll
      for (int i = 0; i < 4; ++i) mma16816(o[i], pl, bh[i >> 1][2 * (i & 1)], bh[i >> 1][2 * (i & 1) + 1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) mma16816(o[i], ph, bl[i >> 1][2 * (i & 1)], bl[i >> 1][2 * (i & 1) + 1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) mma16816(o[i], ph, bh[i >> 1][2 * (i & 1)], bh[i >> 1][2 * (i & 1) + 1]);
    }
    __syncthreads();  // the buffer is re-filled two iterations later
  }
  // ---- normalise and store ----
  l0 += __shfl_xor_sync(kFull, l0, 1);
  l0 += __shfl_xor_sync(kFull, l0, 2);
  l1 += __shfl_xor_sync(kFull, l1, 1);
  l1 += __shfl_xor_sync(kFull, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  if (out) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = head * HD + 8 * i + 2 * t;
      if (r0 < tl.q_rows)
        *reinterpret_cast<float2*>(out + (size_t)(tl.q_row0 + r0) * out_ld + c) = make_float2(o[i][0] * i0, o[i][1] * i0);
      if (r1 < tl.q_rows)
        *reinterpret_cast<float2*>(out + (size_t)(tl.q_row0 + r1) * out_ld + c) = make_float2(o[i][2] * i1, o[i][3] * i1);
    }
  }
  if (out_img) {
    // A image of the output projection (gemm_tc.cu): token -> tile token/64, stacked rows 2r (hi), 2r+1 (lo);
    // column c -> K atom c/64, 16-byte chunk (c%64)/8, SWIZZLE_128B
    float amax16 = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = half ? r1 : r0;
      if (r >= tl.q_rows) continue;
      const int token = tl.q_row0 + r;
      const float inv = (half ? i1 : i0) * img_scale;
      const uint32_t r2 = 2 * (token & 63);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = head * HD + 8 * i + 2 * t;
        const float v0 = o[i][2 * half] * inv, v1 = o[i][2 * half + 1] * inv;
        amax16 = fmaxf(amax16, fmaxf(fabsf(v0), fabsf(v1)));
        const __half2 hh = __floats2half2_rn(v0, v1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
        unsigned char* blk = out_img + ((size_t)(token >> 6) * img_katoms + (c >> 6)) * 16384;
        const uint32_t chunk = (c & 63) >> 3, inb = (c & 7) * 2;
        const uint32_t o_hi = (r2 >> 3) * 1024u + (r2 & 7u) * 128u + ((chunk ^ (r2 & 7u)) << 4) + inb;
        const uint32_t o_lo = ((r2 + 1) >> 3) * 1024u + ((r2 + 1) & 7u) * 128u + ((chunk ^ ((r2 + 1) & 7u)) << 4) + inb;
        *reinterpret_cast<uint32_t*>(blk + o_hi) = h2u(hh);
        *reinterpret_cast<uint32_t*>(blk + o_lo) = h2u(ll);
      }
    }
    if (!(amax16 <= 65504.f)) atomicOr(&g_attention_flags, SPR_FLAG_FP16_OVERFLOW);  // false for NaN as well
  }
}

// x (fp32, row stride ld_in) -> fp16 hi / lo planes (row stride ld_out); columns < n_scaled are multiplied by
// `scale` first (the query block of a packed QKV projection)
__global__ void __launch_bounds__(256) k_split_f16(const float* __restrict__ x, int rows, int cols, int ld_in,
                                                    __half* __restrict__ hi, __half* __restrict__ lo, int ld_out,
                                                    int n_scaled, float scale) {
  const int c4 = cols >> 2;
  const size_t total = (size_t)rows * c4;
  float amax16 = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / c4), c = (int)(i % c4) * 4;
    float4 v = *reinterpret_cast<const float4*>(x + (size_t)r * ld_in + c);
    if (c < n_scaled) {
      v.x *= scale;
      v.y *= scale;
      v.z *= scale;
      v.w *= scale;
    }
    amax16 = fmaxf(amax16, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    *reinterpret_cast<uint2*>(hi + (size_t)r * ld_out + c) = make_uint2(h2u(h0), h2u(h1));
    *reinterpret_cast<uint2*>(lo + (size_t)r * ld_out + c) = make_uint2(h2u(l0), h2u(l1));
  }
  if (!(amax16 <= 65504.f)) atomicOr(&g_attention_flags, SPR_FLAG_FP16_OVERFLOW);
}

}  // namespace

unsigned int attention_numeric_flags(bool reset) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_attention_flags, sizeof(v)) != cudaSuccess) return 0;
  if (reset && v) {
    const unsigned int zero = 0;
    cudaMemcpyToSymbol(g_attention_flags, &zero, sizeof(zero));
  }
  return v;
}

}  // namespace spr

using namespace spr;

extern "C" int spr_split_f16(const float* d_x, int rows, int cols, int ld_in, void* d_hi, void* d_lo, int ld_out,
                             int n_scaled, float scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(rows > 0 && cols > 0 && (cols & 3) == 0 && (ld_in & 3) == 0 && (ld_out & 3) == 0,
                "split_f16: rows/cols must be positive and cols, ld_in, ld_out multiples of 4");
  SPR_CHECK_ARG(d_x && d_hi && d_lo, "split_f16: null pointer");
  const size_t total = (size_t)rows * (cols >> 2);
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  k_split_f16<<<grid, 256, 0, stream>>>(d_x, rows, cols, ld_in, static_cast<__half*>(d_hi), static_cast<__half*>(d_lo),
                                        ld_out, n_scaled, scale);
  SPR_LAUNCH_CHECK("k_split_f16");
  return SPR_OK;
}

extern "C" int spr_attention_varlen(const void* d_hi, const void* d_lo, int ld, int q_col, int k_col, int v_col,
                                    int n_heads, int head_dim, const int32_t* d_tiles, int n_tiles, float* d_out,
                                    int out_ld, void* d_out_img, float img_scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(head_dim == HD, "attention_varlen: head_dim must be %d (got %d)", HD, head_dim);
  SPR_CHECK_ARG(n_tiles > 0 && n_heads > 0, "attention_varlen: empty launch");
  SPR_CHECK_ARG((ld & 7) == 0 && (q_col & 7) == 0 && (k_col & 7) == 0 && (v_col & 7) == 0 && (out_ld & 1) == 0,
                "attention_varlen: row strides / column offsets must keep 16-byte alignment");
  SPR_CHECK_ARG(d_hi && d_lo && d_tiles && (d_out || d_out_img), "attention_varlen: null pointer");
  dim3 grid(n_tiles, n_heads);
  k_attention<<<grid, 128, 0, stream>>>(static_cast<const __half*>(d_hi), static_cast<const __half*>(d_lo), ld, q_col,
                                        k_col, v_col, reinterpret_cast<const AttnTile*>(d_tiles), d_out, out_ld,
                                        static_cast<unsigned char*>(d_out_img), (n_heads * HD + 63) / 64, img_scale);
  SPR_LAUNCH_CHECK("k_attention");
  return SPR_OK;
}
