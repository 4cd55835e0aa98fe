// kpconv.cu -- fused KPConv layer forward on sm_100a (rigid kernel, linear influence, sum aggregation).
//
// Replaces KPConv.forward (reference: models/backbone_kpconv/kpconv_blocks.py:269-414), which materialises
// [N,H,3], [N,H,K,3], [N,H,K], [N,K,H], [N,H,Cin], [N,K,Cin] and [K,N,Cout] tensors in global memory.
// Here one kernel does, per tile of TQ query points:
//   phase 1 (one warp per query)
//     a. lanes = neighbours: load idx, support xyz, compute the K=15 linear influences
//        w[h][k] = max(0, 1 - |(s[idx]-q) - kp[k]| / extent) into a per-warp shared buffer, count the
//        neighbours whose feature row-sum is > 0 (the reference's normaliser, :409-412), and ballot the
//        neighbours with at least one non-zero influence (shadow neighbours never qualify);
//     b. lanes = channels: for every active neighbour, one coalesced (vectorised) load of its feature row
//        and K fused multiply-adds per channel into register accumulators wf[k][c];
//     c. wf -> shared A tile [TQ][K*Cin].
//   phase 2 (whole CTA): A[TQ x K*Cin] @ W[K*Cin x Cout] on the fp32 pipe, 8x4 register tiles, split-K across
//     thread groups, reduced through shared memory; epilogue divides by the neighbour count and stores.
// Nothing but q, s, idx, x, W are read and out written: no intermediate ever reaches global memory.
#include "spr_common.cuh"

namespace spr {
namespace {

constexpr int KP = 15;          // kernel points handled by the fused kernels
constexpr int kWStride = 20;    // floats per neighbour in the influence buffer (16 used; 20 keeps LDS.128 aligned)

// ---------------------------------------------------------------------------------------------
// feature row-sum flags: flag[j] = (sum_c x[j,c] > 0), flag[ns] = 0 (shadow row)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_rowsum_flags(const float* __restrict__ x, int ns, int cin,
                                                       unsigned char* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row > ns) return;
  if (row == ns) {
    if (lane == 0) flag[ns] = 0;
    return;
  }
  float acc = 0.f;
  for (int c = lane; c < cin; c += 32) acc += x[(size_t)row * cin + c];
  acc = warp_sum(acc);
  if (lane == 0) flag[row] = acc > 0.f ? 1 : 0;
}

template <typename IdxT>
__device__ __forceinline__ int load_idx(const IdxT* __restrict__ p) {
  return (int)__ldg(p);
}

// influence of kernel point (kx,ky,kz) on the centred neighbour (cx,cy,cz); fp32, reference op order
// (kpconv_blocks.py:325-329 differences**2 summed over xyz, :368 clamp(1 - sqrt(d2)/extent, min=0))
__device__ __forceinline__ float influence(float cx, float cy, float cz, float kx, float ky, float kz, float extent) {
  const float dx = cx - kx, dy = cy - ky, dz = cz - kz;
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  const float v = 1.f - __fdiv_rn(__fsqrt_rn(d2), extent);
  return fmaxf(v, 0.f);
}

// ---------------------------------------------------------------------------------------------
// Cin == 1 (first encoder block, features = ones): one warp per query, lanes = neighbours.
// ---------------------------------------------------------------------------------------------
template <typename IdxT>
__global__ void __launch_bounds__(256)
    k_kpconv_cin1(const float* __restrict__ q, const float* __restrict__ s, const IdxT* __restrict__ idx,
                  int row_stride, int H, const float* __restrict__ x, const float* __restrict__ w, int cout,
                  const float* __restrict__ kp, float extent, float* __restrict__ out, int nq, int ns) {
  extern __shared__ float sm[];
  float* s_w = sm;                  // [KP][cout]
  float* s_kp = sm + KP * cout;     // [KP*3]
  for (int i = threadIdx.x; i < KP * cout; i += blockDim.x) s_w[i] = w[i];
  for (int i = threadIdx.x; i < KP * 3; i += blockDim.x) s_kp[i] = kp[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int n = blockIdx.x * warps + (threadIdx.x >> 5); n < nq; n += gridDim.x * warps) {
    const float qx = __ldg(q + 3 * (size_t)n), qy = __ldg(q + 3 * (size_t)n + 1), qz = __ldg(q + 3 * (size_t)n + 2);
    float wf[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) wf[k] = 0.f;
    int nn = 0;
    for (int h0 = 0; h0 < H; h0 += 32) {
      const int h = h0 + lane;
      int j = ns;
      if (h < H) j = load_idx(idx + (size_t)n * row_stride + h);
      const bool valid = j >= 0 && j < ns;
      float xv = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
      if (valid) {
        xv = __ldg(x + j);
        cx = __ldg(s + 3 * (size_t)j) - qx;
        cy = __ldg(s + 3 * (size_t)j + 1) - qy;
        cz = __ldg(s + 3 * (size_t)j + 2) - qz;
      }
      nn += __popc(__ballot_sync(kFull, valid && xv > 0.f));
      if (valid) {
#pragma unroll
        for (int k = 0; k < KP; ++k)
          wf[k] = fmaf(influence(cx, cy, cz, s_kp[3 * k], s_kp[3 * k + 1], s_kp[3 * k + 2], extent), xv, wf[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < KP; ++k) wf[k] = warp_sum(wf[k]);
    const float inv = 1.f / (float)max(nn, 1);
    for (int o = lane; o < cout; o += 32) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k) acc = fmaf(wf[k], s_w[k * cout + o], acc);
      out[(size_t)n * cout + o] = acc * inv;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// main fused kernel, Cin = Cout = C in {32, 64, 128, 256}
// ---------------------------------------------------------------------------------------------
template <int C>
struct Cfg {
  static constexpr int CT = C > 128 ? 128 : C;       // channels per pass
  static constexpr int PASSES = C / CT;
  static constexpr int CPL = CT / 32;                 // channels per lane in phase 1
  static constexpr int TQ = CT == 128 ? 16 : 32;      // queries per tile
  static constexpr int KC = KP * CT;                  // contraction length per pass
  static constexpr int AS = KC + 4;                   // A row stride (floats)
  static constexpr int THREADS = 256;
  static constexpr int WARPS = THREADS / 32;
  static constexpr int RQ = 8, RC = 4;                // phase-2 register tile
  static constexpr int TILE_THREADS = (TQ / RQ) * (C / RC);
  static constexpr int G = THREADS / TILE_THREADS;    // split-K groups
  static constexpr int KC_G = KC / G;
  static_assert(TILE_THREADS <= THREADS && THREADS % TILE_THREADS == 0, "tile/threads mismatch");
  static_assert(KC % G == 0 && KC_G % 4 == 0, "split-K must be a multiple of 4");
  static constexpr size_t SMEM_A = (size_t)TQ * AS * 4;
  static constexpr size_t SMEM_W = (size_t)WARPS * 32 * kWStride * 4;
  static constexpr size_t SMEM_RED = (size_t)G * TQ * C * 4;
  static constexpr size_t SMEM = (SMEM_A > SMEM_RED ? SMEM_A : SMEM_RED) + SMEM_W + 64 * 4 + KP * 3 * 4 + 64;
};

template <int CPL>
struct VecLoad;
template <>
struct VecLoad<1> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
};
template <>
struct VecLoad<2> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[2]) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x;
    v[1] = t.y;
  }
};
template <>
struct VecLoad<4> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x;
    v[1] = t.y;
    v[2] = t.z;
    v[3] = t.w;
  }
};

template <int C, typename IdxT>
__global__ void __launch_bounds__(Cfg<C>::THREADS, 1)
    k_kpconv_fused(const float* __restrict__ q, const float* __restrict__ s, const IdxT* __restrict__ idx,
                   int row_stride, int H, const float* __restrict__ x, const float* __restrict__ w,
                   const float* __restrict__ kp, const unsigned char* __restrict__ rowflag, float extent,
                   float* __restrict__ out, int nq, int ns, int n_tiles) {
  using K = Cfg<C>;
  constexpr int CT = K::CT, CPL = K::CPL, TQ = K::TQ, AS = K::AS, G = K::G, RQ = K::RQ, RC = K::RC;
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                                                          // [TQ][AS]   (aliased by the split-K reduction)
  float* sW = smem + (K::SMEM_A > K::SMEM_RED ? K::SMEM_A : K::SMEM_RED) / 4;  // [WARPS][32][kWStride]
  float* sInv = sW + K::WARPS * 32 * kWStride;                               // [TQ] 1/neighbour_num  (64 reserved)
  float* sKp = sInv + 64;                                                    // [KP*3]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < KP * 3; i += K::THREADS) sKp[i] = kp[i];
  __syncthreads();
  float* wbuf = sW + warp * 32 * kWStride;

  // phase-2 coordinates
  const int grp = tid / K::TILE_THREADS;
  const int tt = tid % K::TILE_THREADS;
  const int tx = tt % (C / RC), ty = tt / (C / RC);

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int q0 = tile * TQ;
    float acc2[RQ][RC];
#pragma unroll
    for (int a = 0; a < RQ; ++a)
#pragma unroll
      for (int b = 0; b < RC; ++b) acc2[a][b] = 0.f;

#pragma unroll 1
    for (int pass = 0; pass < K::PASSES; ++pass) {
      const int cbase = pass * CT;  // first input channel of this pass
      // ------------------------------ phase 1 ------------------------------
      for (int ql = warp; ql < TQ; ql += K::WARPS) {
        const int n = q0 + ql;
        float acc[KP][CPL];
#pragma unroll
        for (int k = 0; k < KP; ++k)
#pragma unroll
          for (int c = 0; c < CPL; ++c) acc[k][c] = 0.f;
        int nn = 0;
        if (n < nq) {
          const float qx = __ldg(q + 3 * (size_t)n), qy = __ldg(q + 3 * (size_t)n + 1),
                      qz = __ldg(q + 3 * (size_t)n + 2);
          for (int h0 = 0; h0 < H; h0 += 32) {
            const int h = h0 + lane;
            int j = ns;
            if (h < H) j = load_idx(idx + (size_t)n * row_stride + h);
            const bool valid = j >= 0 && j < ns;
            bool active = false;
            if (valid) {
              const float cx = __ldg(s + 3 * (size_t)j) - qx, cy = __ldg(s + 3 * (size_t)j + 1) - qy,
                          cz = __ldg(s + 3 * (size_t)j + 2) - qz;
              float wv[16];
#pragma unroll
              for (int k = 0; k < KP; ++k) {
                wv[k] = influence(cx, cy, cz, sKp[3 * k], sKp[3 * k + 1], sKp[3 * k + 2], extent);
                active |= wv[k] > 0.f;
              }
              wv[15] = 0.f;
              float4* dst = reinterpret_cast<float4*>(wbuf + lane * kWStride);
              dst[0] = make_float4(wv[0], wv[1], wv[2], wv[3]);
              dst[1] = make_float4(wv[4], wv[5], wv[6], wv[7]);
              dst[2] = make_float4(wv[8], wv[9], wv[10], wv[11]);
              dst[3] = make_float4(wv[12], wv[13], wv[14], wv[15]);
            }
            nn += __popc(__ballot_sync(kFull, valid && rowflag[j] != 0));
            unsigned am = __ballot_sync(kFull, active);
            __syncwarp();
            while (am) {
              const int a = __ffs(am) - 1;
              am &= am - 1;
              const int ja = __shfl_sync(kFull, j, a);
              float xv[CPL];
              VecLoad<CPL>::ld(x + (size_t)ja * C + cbase + lane * CPL, xv);
              const float4* wp = reinterpret_cast<const float4*>(wbuf + a * kWStride);
              const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
              const float wk[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                                    w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
              for (int k = 0; k < KP; ++k)
#pragma unroll
                for (int c = 0; c < CPL; ++c) acc[k][c] = fmaf(wk[k], xv[c], acc[k][c]);
            }
            __syncwarp();
          }
        }
        // wf -> A tile (zero rows for queries past the end)
        float* arow = sA + ql * AS;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          if (CPL == 1) {
            arow[k * CT + lane] = acc[k][0];
          } else if (CPL == 2) {
            *reinterpret_cast<float2*>(arow + k * CT + lane * 2) = make_float2(acc[k][0], acc[k][1]);
          } else {
            *reinterpret_cast<float4*>(arow + k * CT + lane * 4) =
                make_float4(acc[k][0], acc[k][1], acc[k][CPL > 2 ? 2 : 0], acc[k][CPL > 3 ? 3 : 0]);
          }
        }
        if (pass == 0 && lane == 0) sInv[ql] = 1.f / (float)max(nn, 1);
      }
      __syncthreads();
      // ------------------------------ phase 2 ------------------------------
      {
        const int kc0 = grp * K::KC_G;
        const float* wrow = w + ((size_t)0) + (size_t)tx * RC;  // W[k][cin][cout] row-major: row (k*C + cin)
#pragma unroll 2
        for (int kk = 0; kk < K::KC_G; kk += 4) {
          const int kc = kc0 + kk;                  // index inside this pass: k*CT + c
          const int kidx = kc / CT, cc = kc % CT;   // 4 consecutive kc never straddle a kernel point (CT % 4 == 0)
          const float* wp = wrow + ((size_t)kidx * C + cbase + cc) * C;
          float4 wv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) wv[i] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)i * C));
#pragma unroll
          for (int a = 0; a < RQ; ++a) {
            const float4 av = *reinterpret_cast<const float4*>(sA + (ty * RQ + a) * AS + kc);
            const float ar[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc2[a][0] = fmaf(ar[i], wv[i].x, acc2[a][0]);
              acc2[a][1] = fmaf(ar[i], wv[i].y, acc2[a][1]);
              acc2[a][2] = fmaf(ar[i], wv[i].z, acc2[a][2]);
              acc2[a][3] = fmaf(ar[i], wv[i].w, acc2[a][3]);
            }
          }
        }
      }
      __syncthreads();  // A tile free for the next pass / the reduction
    }
    // ------------------------------ split-K reduction + epilogue ------------------------------
    float* sRed = smem;  // [G][TQ][C]
#pragma unroll
    for (int a = 0; a < RQ; ++a)
      *reinterpret_cast<float4*>(sRed + ((size_t)grp * TQ + ty * RQ + a) * C + tx * RC) =
          make_float4(acc2[a][0], acc2[a][1], acc2[a][2], acc2[a][3]);
    __syncthreads();
    for (int e = tid; e < TQ * C / 4; e += K::THREADS) {
      const int ql = e / (C / 4), c4 = e % (C / 4);
      float4 r = *reinterpret_cast<const float4*>(sRed + (size_t)ql * C + c4 * 4);
#pragma unroll
      for (int g = 1; g < G; ++g) {
        const float4 t = *reinterpret_cast<const float4*>(sRed + ((size_t)g * TQ + ql) * C + c4 * 4);
        r.x += t.x;
        r.y += t.y;
        r.z += t.z;
        r.w += t.w;
      }
      const int n = q0 + ql;
      if (n < nq) {
        const float inv = sInv[ql];
        r.x *= inv;
        r.y *= inv;
        r.z *= inv;
        r.w *= inv;
        *reinterpret_cast<float4*>(out + (size_t)n * C + c4 * 4) = r;
      }
    }
    __syncthreads();
  }
}

template <int C, typename IdxT>
int launch_fused(const float* q, const float* s, const void* idx, int row_stride, int H, const float* x, const float* w,
                 const float* kp, const unsigned char* rowflag, float extent, float* out, int nq, int ns,
                 cudaStream_t stream) {
  using K = Cfg<C>;
  const int n_tiles = (nq + K::TQ - 1) / K::TQ;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    SPR_CUDA(cudaFuncSetAttribute(k_kpconv_fused<C, IdxT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K::SMEM));
    attr_set = true;
  }
  const int grid = n_tiles;  // one tile per CTA; the hardware scheduler balances ragged tiles
  k_kpconv_fused<C, IdxT><<<grid, K::THREADS, K::SMEM, stream>>>(q, s, static_cast<const IdxT*>(idx), row_stride, H, x,
                                                                 w, kp, rowflag, extent, out, nq, ns, n_tiles);
  SPR_LAUNCH_CHECK("k_kpconv_fused");
  return SPR_OK;
}

}  // namespace
}  // namespace spr

using namespace spr;

extern "C" size_t spr_kpconv_workspace_bytes(int nq, int ns, int cin, int cout, int n_kernel_points) {
  (void)nq;
  (void)cin;
  (void)cout;
  (void)n_kernel_points;
  return align_up((size_t)(ns > 0 ? ns : 0) + 1, 256) + 256;
}

extern "C" int spr_kpconv_forward(const float* d_q, const float* d_s, const void* d_idx, int idx_is_64, int row_stride,
                                  int H, const float* d_x, int cin, const float* d_w, int cout, const float* d_kp,
                                  int n_kernel_points, float extent, float* d_out, int nq, int ns, int mode,
                                  void* d_workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(nq > 0 && ns > 0, "kpconv_forward: empty input (nq=%d, ns=%d)", nq, ns);
  SPR_CHECK_ARG(H > 0 && row_stride >= H, "kpconv_forward: bad neighbour matrix shape (H=%d, row_stride=%d)", H,
                row_stride);
  SPR_CHECK_ARG(extent > 0.f, "kpconv_forward: extent must be > 0");
  SPR_CHECK_ARG(d_q && d_s && d_idx && d_x && d_w && d_kp && d_out, "kpconv_forward: null pointer");
  if (n_kernel_points != KP) {
    set_error("kpconv_forward: only %d kernel points are supported (got %d)", KP, n_kernel_points);
    return SPR_EUNSUPPORTED;
  }
  if (mode != 0) {
    set_error("kpconv_forward: mode %d not built in", mode);
    return SPR_EUNSUPPORTED;
  }
  if (cin == 1) {
    const size_t smem = (size_t)(KP * cout + KP * 3) * 4;
    SPR_CHECK_ARG(smem <= 48 * 1024, "kpconv_forward: cout %d too large for the Cin=1 kernel", cout);
    const int warps = 8;
    int grid = (nq + warps - 1) / warps;
    if (grid > kNumSMs * 16) grid = kNumSMs * 16;
    if (idx_is_64)
      k_kpconv_cin1<long long><<<grid, warps * 32, smem, stream>>>(d_q, d_s, static_cast<const long long*>(d_idx),
                                                                   row_stride, H, d_x, d_w, cout, d_kp, extent, d_out,
                                                                   nq, ns);
    else
      k_kpconv_cin1<int><<<grid, warps * 32, smem, stream>>>(d_q, d_s, static_cast<const int*>(d_idx), row_stride, H,
                                                             d_x, d_w, cout, d_kp, extent, d_out, nq, ns);
    SPR_LAUNCH_CHECK("k_kpconv_cin1");
    return SPR_OK;
  }
  if (cin != cout || !(cin == 32 || cin == 64 || cin == 128 || cin == 256)) {
    set_error("kpconv_forward: unsupported channel shape Cin=%d Cout=%d (supported: Cin=1, or Cin=Cout in {32,64,128,256})",
              cin, cout);
    return SPR_EUNSUPPORTED;
  }
  if (!d_workspace || workspace_bytes < spr_kpconv_workspace_bytes(nq, ns, cin, cout, n_kernel_points)) {
    set_error("kpconv_forward: workspace too small");
    return SPR_ENOSPACE;
  }
  unsigned char* rowflag = static_cast<unsigned char*>(d_workspace);
  k_rowsum_flags<<<(ns + 1 + 7) / 8, 256, 0, stream>>>(d_x, ns, cin, rowflag);
  SPR_LAUNCH_CHECK("k_rowsum_flags");
#define SPR_DISPATCH(CC)                                                                                          \
  case CC:                                                                                                        \
    return idx_is_64 ? launch_fused<CC, long long>(d_q, d_s, d_idx, row_stride, H, d_x, d_w, d_kp, rowflag, extent, \
                                                   d_out, nq, ns, stream)                                         \
                     : launch_fused<CC, int>(d_q, d_s, d_idx, row_stride, H, d_x, d_w, d_kp, rowflag, extent,     \
                                             d_out, nq, ns, stream);
  switch (cin) {
    SPR_DISPATCH(32)
    SPR_DISPATCH(64)
    SPR_DISPATCH(128)
    SPR_DISPATCH(256)
  }
#undef SPR_DISPATCH
  return SPR_EUNSUPPORTED;
}
