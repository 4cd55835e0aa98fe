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

#include <cstdlib>

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

// Influence of kernel point (kx,ky,kz) on the centred neighbour (cx,cy,cz): kpconv_blocks.py:325-329 (differences**2
// summed over xyz) and :368 (clamp(1 - sqrt(d2)/extent, min=0)), evaluated with an FMA-contracted distance, the
// MUFU square root and a multiplication by 1/extent.  Differs from the exactly-rounded expression by a few ulp of the influence
// (<= ~3e-7 absolute on values in [0,1]), far inside the feature tolerance, and costs ~10 instructions
// instead of ~28 (IEEE sqrt and division are multi-instruction sequences).
__device__ __forceinline__ float influence_fast(float cx, float cy, float cz, float kx, float ky, float kz,
                                                float inv_extent) {
  const float dx = cx - kx, dy = cy - ky, dz = cz - kz;
  const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  float d;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(d2));  // MUFU.SQRT, rel. error ~2^-23, sqrt(0) = 0
  return fmaxf(fmaf(-d, inv_extent, 1.f), 0.f);
}

// ---------------------------------------------------------------------------------------------
// Cin == 1 (first encoder block, features = ones): one warp per query.
// Lane (g = lane / 4, t = lane % 4) owns kernel points g and g + 8 and, in every block of 8 neighbours, slots 2t and
// 2t + 1: the 16 x 8 influence evaluations of a block are spread over the warp without redundancy, a lane
// accumulates its two kernel points directly, and the sum over neighbours is two shuffles at the end (instead of
// one butterfly per kernel point).  wf[15] is then broadcast and contracted with W[15, Cout] held in registers
// (Cout <= 64) or shared memory.
// ---------------------------------------------------------------------------------------------
// support point and its scalar feature in one 16-byte record: one load per neighbour instead of four
__global__ void __launch_bounds__(256) k_pack_points(const float* __restrict__ s, const float* __restrict__ x, int ns,
                                                      float4* __restrict__ packed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ns) packed[i] = make_float4(s[3 * (size_t)i], s[3 * (size_t)i + 1], s[3 * (size_t)i + 2], x[i]);
}

template <typename IdxT, int NO>  // NO = Cout / 32 outputs per lane
__global__ void __launch_bounds__(256)
    k_kpconv_cin1(const float* __restrict__ q, const float4* __restrict__ packed, const IdxT* __restrict__ idx,
                  int row_stride, int H, const float* __restrict__ w, const float* __restrict__ kp, float extent,
                  float* __restrict__ out, int nq, int ns) {
  constexpr int COUT = NO * 32;
  extern __shared__ float sm[];
  float* s_w = sm;  // [KP][COUT], only used when NO > 2
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  if (NO > 2) {
    for (int i = threadIdx.x; i < KP * COUT; i += blockDim.x) s_w[i] = w[i];
    __syncthreads();
  }
  float wreg[NO <= 2 ? KP : 1][NO <= 2 ? NO : 1];
  if (NO <= 2) {
#pragma unroll
    for (int k = 0; k < KP; ++k)
#pragma unroll
      for (int o = 0; o < NO; ++o) wreg[k][o] = __ldg(w + k * COUT + lane + 32 * o);
  }
  const float inv_extent = 1.f / extent;
  const float k0x = __ldg(kp + 3 * g), k0y = __ldg(kp + 3 * g + 1), k0z = __ldg(kp + 3 * g + 2);
  const int g1 = g < 7 ? g + 8 : 0;
  const float k1x = __ldg(kp + 3 * g1), k1y = __ldg(kp + 3 * g1 + 1), k1z = __ldg(kp + 3 * g1 + 2);
  const float k1_on = g < 7 ? 1.f : 0.f;
  const int stride_q = gridDim.x * warps;
  int n = blockIdx.x * warps + (threadIdx.x >> 5);
  if (n >= nq) return;

  // neighbour row of the NEXT query is requested while the current one is processed; the two packed points of
  // block b+1 are in flight while block b is evaluated
  int jr[3], jn[3];
  auto load_row = [&](int nn_, int (&dst)[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int h = 32 * i + lane;
      dst[i] = -1;
      if (32 * i < H && h < H && nn_ < nq) dst[i] = load_idx(idx + (size_t)nn_ * row_stride + h);
    }
  };
  load_row(n, jn);
  float qnx = __ldg(q + 3 * (size_t)n), qny = __ldg(q + 3 * (size_t)n + 1), qnz = __ldg(q + 3 * (size_t)n + 2);
  for (; n < nq; n += stride_q) {
    const float qx = qnx, qy = qny, qz = qnz;
    unsigned bm = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const bool valid = jn[i] >= 0 && jn[i] < ns;
      jr[i] = valid ? jn[i] : -1;
      if (32 * i < H) {
        const unsigned m = __ballot_sync(kFull, valid);
        bm |= (((m & 0xffu) ? 1u : 0u) | ((m & 0xff00u) ? 2u : 0u) | ((m & 0xff0000u) ? 4u : 0u) |
               ((m & 0xff000000u) ? 8u : 0u)) << (4 * i);
      }
    }
    const int n_next = n + stride_q;
    load_row(n_next, jn);
    if (n_next < nq) {
      qnx = __ldg(q + 3 * (size_t)n_next);
      qny = __ldg(q + 3 * (size_t)n_next + 1);
      qnz = __ldg(q + 3 * (size_t)n_next + 2);
    }
    float acc0 = 0.f, acc1 = 0.f, cnt = 0.f;
    auto fetch = [&](float4& pa, float4& pb) {
      const int b = __ffs(bm) - 1;
      bm &= bm - 1;
      const int src = (b & 3) * 8 + 2 * t;
      const int jsel = (b >> 2) == 0 ? jr[0] : ((b >> 2) == 1 ? jr[1] : jr[2]);
      const int ja = __shfl_sync(kFull, jsel, src);
      const int jb = __shfl_sync(kFull, jsel, src + 1);
      pa = make_float4(0.f, 0.f, 0.f, 0.f);
      pb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ja >= 0) pa = __ldg(packed + ja);
      if (jb >= 0) pb = __ldg(packed + jb);
    };
    float4 ca, cb, na, nb;
    bool have = bm != 0;  // warp-uniform: the row has at least one neighbour
    if (have) fetch(na, nb);
    while (have) {
      ca = na;
      cb = nb;
      have = bm != 0;
      if (have) fetch(na, nb);
      // an absent neighbour has w = x = 0 and contributes nothing
      const float ax = ca.x - qx, ay = ca.y - qy, az = ca.z - qz;
      const float bx = cb.x - qx, by = cb.y - qy, bz = cb.z - qz;
      acc0 = fmaf(influence_fast(ax, ay, az, k0x, k0y, k0z, inv_extent), ca.w, acc0);
      acc0 = fmaf(influence_fast(bx, by, bz, k0x, k0y, k0z, inv_extent), cb.w, acc0);
      acc1 = fmaf(influence_fast(ax, ay, az, k1x, k1y, k1z, inv_extent), ca.w, acc1);
      acc1 = fmaf(influence_fast(bx, by, bz, k1x, k1y, k1z, inv_extent), cb.w, acc1);
      cnt += (ca.w > 0.f ? 1.f : 0.f) + (cb.w > 0.f ? 1.f : 0.f);  // neighbour_num: rowsum(x) = x for Cin = 1
    }
    acc0 += __shfl_xor_sync(kFull, acc0, 1);
    acc0 += __shfl_xor_sync(kFull, acc0, 2);
    acc1 += __shfl_xor_sync(kFull, acc1, 1);
    acc1 += __shfl_xor_sync(kFull, acc1, 2);
    acc1 *= k1_on;
    cnt += __shfl_xor_sync(kFull, cnt, 1);
    cnt += __shfl_xor_sync(kFull, cnt, 2);   // every g group now holds the full count
    const float inv = 1.f / fmaxf(cnt, 1.f);
    float o[NO];
#pragma unroll
    for (int i = 0; i < NO; ++i) o[i] = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const float wf = __shfl_sync(kFull, k < 8 ? acc0 : acc1, 4 * (k & 7));
#pragma unroll
      for (int i = 0; i < NO; ++i) o[i] = fmaf(wf, NO <= 2 ? wreg[k][i] : s_w[k * COUT + lane + 32 * i], o[i]);
    }
#pragma unroll
    for (int i = 0; i < NO; ++i) out[(size_t)n * COUT + lane + 32 * i] = o[i] * inv;
  }
}

// ---------------------------------------------------------------------------------------------
// Cin == 1, second version (round 2): ONE THREAD per query.  The warp-per-query kernel above spends ~535 warp
// instructions per query, of which only ~200 are influence arithmetic: the rest is the per-block choreography
// (ballots, shuffles of indices, the butterfly at the end, a padded 16th kernel point).  With a thread per query there
// is no cross-lane traffic at all: the 15 kernel points live in registers, a neighbour costs one 16-byte gather and
// 15 x 9 arithmetic instructions (3 FADD, FMUL, 2 FFMA, MUFU.SQRT, FFMA.SAT, FFMA), the neighbours of the NEXT group
// of four are in flight while the current group is evaluated, and the 15 x Cout mat-vec reads W from shared memory
// with broadcast 16-byte loads.  A group of four columns that is padding in all 32 rows of the warp is skipped.
// ---------------------------------------------------------------------------------------------
template <typename IdxT, int COUT>
__global__ void __launch_bounds__(128, 4)
    k_kpconv_cin1_t(const float* __restrict__ q, const float4* __restrict__ packed, const IdxT* __restrict__ idx,
                    int row_stride, int H, const float* __restrict__ w, const float* __restrict__ kp, float extent,
                    float* __restrict__ out, int nq, int ns, const int* __restrict__ order) {
  constexpr int NP = KP / 2;  // kernel-point pairs (2p, 2p + 1); the last kernel point is handled alone
  // The first version of this kernel was bound by the L1 data pipe (79 % of its wavefront peak, ncu): a scalar index
  // load, a gather or a 16-byte row store of 32 threads that work on 32 different rows is 32 wavefronts.  Index rows are
  // therefore read 16 bytes at a time, and (Cout <= 64) the output rows leave through a shared-memory transposition as
  // contiguous 512-byte stores.
  constexpr bool STAGED_OUT = COUT <= 64;
  constexpr int OSTRIDE = COUT + 4;  // floats; 16-byte accesses of 8 consecutive rows fall into distinct banks
  __shared__ __align__(16) float s_w[KP * COUT];
  __shared__ __align__(16) float s_out[STAGED_OUT ? 4 * 32 * OSTRIDE : 4];
  for (int i = threadIdx.x; i < KP * COUT; i += blockDim.x) s_w[i] = __ldg(w + i);
  f2_t nkx[NP], nky[NP], nkz[NP];  // NEGATED kernel points (there is no packed subtraction)
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    nkx[p] = f2_pack(-__ldg(kp + 6 * p), -__ldg(kp + 6 * p + 3));
    nky[p] = f2_pack(-__ldg(kp + 6 * p + 1), -__ldg(kp + 6 * p + 4));
    nkz[p] = f2_pack(-__ldg(kp + 6 * p + 2), -__ldg(kp + 6 * p + 5));
  }
  const float lx = __ldg(kp + 3 * (KP - 1)), ly = __ldg(kp + 3 * (KP - 1) + 1), lz = __ldg(kp + 3 * (KP - 1) + 2);
  __syncthreads();
  const float inv_extent = 1.f / extent;
  const int n_groups = (H + 3) >> 2;
  // 16-byte index loads: 32-bit indices, rows that start on 16-byte boundaries
  const bool vec_idx = sizeof(IdxT) == 4 && (row_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(idx) & 15) == 0;
  for (int base = blockIdx.x * blockDim.x; base < nq; base += gridDim.x * blockDim.x) {
    // `order` (optional) walks the queries in cell order: the 32 queries of a warp are spatial neighbours, their
    // neighbourhoods overlap, and a gather instruction touches a third of the lines
    const bool live = base + (int)threadIdx.x < nq;
    const int n = live ? (order ? __ldg(order + base + threadIdx.x) : base + (int)threadIdx.x) : nq - 1;
    const IdxT* row = idx + (size_t)n * row_stride;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) {
      qx = __ldg(q + 3 * (size_t)n);
      qy = __ldg(q + 3 * (size_t)n + 1);
      qz = __ldg(q + 3 * (size_t)n + 2);
    }
    f2_t acc2[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) acc2[p] = 0ull;
    float acc_l = 0.f, cnt = 0.f;
    // group of four neighbours: indices (requested TWO groups ahead) -> packed points (one group ahead); an absent
    // neighbour is the zero record (x = 0).  Index load and gather are a dependent pair: with both in the same
    // pipeline stage the warp waited for the index before it could request the point.
    auto load_j = [&](int grp, int (&j4)[4]) {
#pragma unroll
      for (int e = 0; e < 4; ++e) j4[e] = -1;
      if (grp >= n_groups || !live) return;
      if (vec_idx && 4 * grp + 4 <= H) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(row) + grp);
        j4[0] = v.x; j4[1] = v.y; j4[2] = v.z; j4[3] = v.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (4 * grp + e < H) j4[e] = load_idx(row + 4 * grp + e);
      }
    };
    auto gather = [&](const int (&j4)[4], float4 (&p)[4]) -> bool {
      bool any = false;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        p[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j4[e] >= 0 && j4[e] < ns) {
          p[e] = __ldg(packed + j4[e]);
          any = true;
        }
      }
      return any;
    };
    float4 cur[4], nxt[4];
    int jn[4];
    load_j(0, jn);
    bool cur_any = __any_sync(kFull, gather(jn, cur));
    load_j(1, jn);
    for (int grp = 0; grp < n_groups; ++grp) {
      const bool nxt_any = __any_sync(kFull, gather(jn, nxt));  // group grp + 1 (all absent past the end)
      load_j(grp + 2, jn);
      // (warp-uniform) a group without a neighbour in any of the warp's 32 rows -- the padded tails -- costs 4 loads
      if (cur_any) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float cx = cur[e].x - qx, cy = cur[e].y - qy, cz = cur[e].z - qz, x = cur[e].w;
          cnt += x > 0.f ? 1.f : 0.f;  // neighbour_num: rowsum(x) = x for Cin = 1
          const f2_t cx2 = f2_pack(cx, cx), cy2 = f2_pack(cy, cy), cz2 = f2_pack(cz, cz), x2 = f2_pack(x, x);
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const f2_t dx = f2_add(cx2, nkx[p]), dy = f2_add(cy2, nky[p]), dz = f2_add(cz2, nkz[p]);
            const f2_t d2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
            float d2a, d2b, da, db;
            f2_unpack(d2, d2a, d2b);
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(da) : "f"(d2a));
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(db) : "f"(d2b));
            // 1 - d/extent <= 1: saturation to [0, 1] is max(0, .)
            const f2_t w2 = f2_pack(__saturatef(fmaf(-da, inv_extent, 1.f)), __saturatef(fmaf(-db, inv_extent, 1.f)));
            acc2[p] = f2_fma(w2, x2, acc2[p]);
          }
          {
            const float dx = cx - lx, dy = cy - ly, dz = cz - lz;
            const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            float d;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(d2));
            acc_l = fmaf(__saturatef(fmaf(-d, inv_extent, 1.f)), x, acc_l);
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) cur[e] = nxt[e];
      cur_any = nxt_any;
    }
    if (!STAGED_OUT && !live) continue;
    float acc[KP];
#pragma unroll
    for (int p = 0; p < NP; ++p) f2_unpack(acc2[p], acc[2 * p], acc[2 * p + 1]);
    acc[KP - 1] = acc_l;
    const float inv = 1.f / fmaxf(cnt, 1.f);
    float* stage = s_out + ((threadIdx.x >> 5) * 32 + (threadIdx.x & 31)) * OSTRIDE;
    float* orow = STAGED_OUT ? stage : out + (size_t)n * COUT;
#pragma unroll 1
    for (int oc = 0; oc < COUT; oc += 16) {
      float o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = 0.f;
#pragma unroll
      for (int k = 0; k < KP; ++k) {
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 wv = *reinterpret_cast<const float4*>(s_w + k * COUT + oc + 4 * v);
          o[4 * v] = fmaf(acc[k], wv.x, o[4 * v]);
          o[4 * v + 1] = fmaf(acc[k], wv.y, o[4 * v + 1]);
          o[4 * v + 2] = fmaf(acc[k], wv.z, o[4 * v + 2]);
          o[4 * v + 3] = fmaf(acc[k], wv.w, o[4 * v + 3]);
        }
      }
#pragma unroll
      for (int v = 0; v < 4; ++v)
        *reinterpret_cast<float4*>(orow + oc + 4 * v) =
            make_float4(o[4 * v] * inv, o[4 * v + 1] * inv, o[4 * v + 2] * inv, o[4 * v + 3] * inv);
    }
    if (STAGED_OUT) {  // the warp's 32 rows are contiguous in `out`: write them 512 bytes per instruction
      __syncwarp();
      constexpr int LPR = COUT / 4;  // lanes per row
      const int lane = threadIdx.x & 31;
      const float* wstage = s_out + (threadIdx.x >> 5) * 32 * OSTRIDE;
      const int row0 = base + (threadIdx.x & ~31);
#pragma unroll
      for (int i = 0; i < LPR; ++i) {
        const int f = i * 32 + lane, r = f / LPR, c4 = (f % LPR) * 4;
        const int nr = __shfl_sync(kFull, n, r);  // the row that staged row r belongs to
        if (row0 + r < nq)
          *reinterpret_cast<float4*>(out + (size_t)nr * COUT + c4) =
              *reinterpret_cast<const float4*>(wstage + r * OSTRIDE + c4);
      }
      __syncwarp();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// main fused kernel, Cin = Cout = C in {32, 64, 128, 256}
// ---------------------------------------------------------------------------------------------
// V selects the tile shape: V = 0 -> one big CTA per SM (TQ rows, 512 threads for C >= 64),
//                           V = 1 -> half-size tiles, 256 threads, two CTAs per SM (phases of different CTAs overlap).
template <int C, int V>
struct Cfg {
  static constexpr int CT = C > 128 ? 128 : C;       // channels per pass
  static constexpr int PASSES = C / CT;
  static constexpr int CPL = CT / 32;                 // channels per lane in phase 1
  static constexpr int TQ0 = CT == 128 ? 16 : 32;
  static constexpr int TQ = V == 0 ? TQ0 : TQ0 / 2;   // queries per tile
  static constexpr int KC = KP * CT;                  // contraction length per pass
  static constexpr int AS = KC + 4;                   // A row stride (floats)
  static constexpr int THREADS = (V == 0 && C >= 64) ? 512 : 256;
  static constexpr int MIN_BLOCKS = (V == 1 || C == 32) ? 2 : 1;
  static constexpr int UNROLL = CPL >= 4 ? 4 : 8;     // neighbour rows in flight per warp in phase 1b
  static constexpr int WARPS = THREADS / 32;
  static constexpr int RQ = (C == 32 && V == 1) ? 4 : 8, RC = 4;  // phase-2 register tile
  static constexpr int TILE_THREADS = (TQ / RQ) * (C / RC);
  static constexpr int G = THREADS / TILE_THREADS;    // split-K groups
  static constexpr int KC_G = KC / G;
  static_assert(TQ % RQ == 0, "tile rows");
  static_assert(TILE_THREADS <= THREADS && THREADS % TILE_THREADS == 0, "tile/threads mismatch");
  static_assert(KC % G == 0 && KC_G % 4 == 0, "split-K must be a multiple of 4");
  static constexpr size_t SMEM_A = (size_t)TQ * AS * 4;
  static constexpr size_t SMEM_W = (size_t)WARPS * 32 * kWStride * 4;
  static constexpr size_t SMEM_RED = (size_t)G * TQ * C * 4;
  static constexpr size_t SMEM_J = (size_t)WARPS * 32 * 4;
  static constexpr size_t SMEM = (SMEM_A > SMEM_RED ? SMEM_A : SMEM_RED) + SMEM_W + SMEM_J + 64 * 4 + KP * 3 * 4 + 64;
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

template <int C, int V, typename IdxT>
__global__ void __launch_bounds__(Cfg<C, V>::THREADS, Cfg<C, V>::MIN_BLOCKS)
    k_kpconv_fused(const float* __restrict__ q, const float* __restrict__ s, const IdxT* __restrict__ idx,
                   int row_stride, int H, const float* __restrict__ x, const float* __restrict__ w,
                   const float* __restrict__ kp, const unsigned char* __restrict__ rowflag, float extent,
                   float* __restrict__ out, int nq, int ns, int n_tiles) {
  using K = Cfg<C, V>;
  constexpr int CT = K::CT, CPL = K::CPL, TQ = K::TQ, AS = K::AS, G = K::G, RQ = K::RQ, RC = K::RC;
  constexpr int QROWS = TQ / RQ;  // thread rows of the phase-2 tile; a thread owns rows ty, ty+QROWS, ...
  const float inv_extent = 1.0f / extent;
  extern __shared__ __align__(16) float smem[];
  float* sA = smem;                                                          // [TQ][AS]   (aliased by the split-K reduction)
  float* sW = smem + (K::SMEM_A > K::SMEM_RED ? K::SMEM_A : K::SMEM_RED) / 4;  // [WARPS][32][kWStride] influences
  int* sJ = reinterpret_cast<int*>(sW + K::WARPS * 32 * kWStride);           // [WARPS][32] compacted neighbour ids
  float* sInv = reinterpret_cast<float*>(sJ + K::WARPS * 32);                // [TQ] 1/neighbour_num  (64 reserved)
  float* sKp = sInv + 64;                                                    // [KP*3]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < KP * 3; i += K::THREADS) sKp[i] = kp[i];
  __syncthreads();
  float* wbuf = sW + warp * 32 * kWStride;
  int* jbuf = sJ + warp * 32;

  // phase-2 coordinates
  const int grp = tid / K::TILE_THREADS;
  const int tt = tid % K::TILE_THREADS;
  const int tx = tt % (C / RC), ty = tt / (C / RC);

  // Software prefetch of the first neighbour round of this warp's FIRST query of the next tile: the index is
  // requested before phase 2 and the support coordinates after it, so both global latencies are hidden behind
  // the contraction of the current tile.
  int pf_j = ns;
  float pf_x = 0.f, pf_y = 0.f, pf_z = 0.f;
  auto prefetch_idx = [&](int tile) {
    pf_j = ns;
    const int n = tile * TQ + warp;
    if (tile < n_tiles && warp < TQ && n < nq && lane < H) pf_j = load_idx(idx + (size_t)n * row_stride + lane);
  };
  auto prefetch_xyz = [&]() {
    if (pf_j >= 0 && pf_j < ns) {
      pf_x = __ldg(s + 3 * (size_t)pf_j);
      pf_y = __ldg(s + 3 * (size_t)pf_j + 1);
      pf_z = __ldg(s + 3 * (size_t)pf_j + 2);
    }
  };
  prefetch_idx(blockIdx.x);
  prefetch_xyz();

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
          const bool use_pf = (ql == warp) && (pass == 0);
          for (int h0 = 0; h0 < H; h0 += 32) {
            const int h = h0 + lane;
            int j = ns;
            float sx = 0.f, sy = 0.f, sz = 0.f;
            if (use_pf && h0 == 0) {
              j = pf_j;
              sx = pf_x;
              sy = pf_y;
              sz = pf_z;
            } else {
              if (h < H) j = load_idx(idx + (size_t)n * row_stride + h);
              if (j >= 0 && j < ns) {
                sx = __ldg(s + 3 * (size_t)j);
                sy = __ldg(s + 3 * (size_t)j + 1);
                sz = __ldg(s + 3 * (size_t)j + 2);
              }
            }
            const bool valid = j >= 0 && j < ns;
            if (!__any_sync(kFull, valid)) break;  // rows are padded at the end: nothing valid from here on
            bool active = false;
            float wv[16];
            if (valid) {
              const float cx = sx - qx, cy = sy - qy, cz = sz - qz;
#pragma unroll
              for (int k = 0; k < KP; ++k) {
                wv[k] = influence_fast(cx, cy, cz, sKp[3 * k], sKp[3 * k + 1], sKp[3 * k + 2], inv_extent);
                active |= wv[k] > 0.f;
              }
              wv[15] = 0.f;
            }
            nn += __popc(__ballot_sync(kFull, valid && rowflag[j] != 0));
            const unsigned am = __ballot_sync(kFull, active);
            const int na = __popc(am);
            if (active) {  // compacted: slot = number of active lanes below this one
              const int pos = __popc(am & ((1u << lane) - 1u));
              float4* dst = reinterpret_cast<float4*>(wbuf + pos * kWStride);
              dst[0] = make_float4(wv[0], wv[1], wv[2], wv[3]);
              dst[1] = make_float4(wv[4], wv[5], wv[6], wv[7]);
              dst[2] = make_float4(wv[8], wv[9], wv[10], wv[11]);
              dst[3] = make_float4(wv[12], wv[13], wv[14], wv[15]);
              jbuf[pos] = j;
            }
            __syncwarp();
            // Batches of UNROLL active neighbours: all feature-row loads of a batch are issued before the first
            // FMA consumes one, so a warp keeps UNROLL independent L2 requests in flight.
            for (int a0 = 0; a0 < na; a0 += K::UNROLL) {
              float xv[K::UNROLL][CPL];
#pragma unroll
              for (int u = 0; u < K::UNROLL; ++u) {
                if (a0 + u < na) {
                  const int ja = jbuf[a0 + u];
                  VecLoad<CPL>::ld(x + (size_t)ja * C + cbase + lane * CPL, xv[u]);
                }
              }
#pragma unroll
              for (int u = 0; u < K::UNROLL; ++u) {
                if (a0 + u < na) {
                  const float4* wp = reinterpret_cast<const float4*>(wbuf + (a0 + u) * kWStride);
                  const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
                  const float wk[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                                        w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
                  for (int k = 0; k < KP; ++k)
#pragma unroll
                    for (int c = 0; c < CPL; ++c) acc[k][c] = fmaf(wk[k], xv[u][c], acc[k][c]);
                }
              }
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
      // ------------------------------ phase 2 ------------------------------
      // W[k][cin][cout] row-major: contraction row kc of this pass is global row (kc / CT) * C + cbase + kc % CT.
      // Four consecutive kc never straddle a kernel point (CT % 4 == 0).  The W rows of step kk+4 are loaded
      // while step kk is multiplied (register double buffer); the first step is requested before the barrier.
      const int kc0 = grp * K::KC_G;
      const float* wcol = w + (size_t)tx * RC;
      auto wrow_ptr = [&](int kc) { return wcol + ((size_t)(kc / CT) * C + cbase + (kc % CT)) * C; };
      float4 wn[4];
      {
        const float* wp = wrow_ptr(kc0);
#pragma unroll
        for (int i = 0; i < 4; ++i) wn[i] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)i * C));
      }
      if (pass == K::PASSES - 1) prefetch_idx(tile + gridDim.x);
      __syncthreads();
#pragma unroll 2
      for (int kk = 0; kk < K::KC_G; kk += 4) {
        const int kc = kc0 + kk;
        float4 wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wv[i] = wn[i];
        if (kk + 4 < K::KC_G) {
          const float* wp = wrow_ptr(kc + 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) wn[i] = __ldg(reinterpret_cast<const float4*>(wp + (size_t)i * C));
        }
#pragma unroll
        for (int a = 0; a < RQ; ++a) {
          const float4 av = *reinterpret_cast<const float4*>(sA + (a * QROWS + ty) * AS + kc);
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
      if (pass == K::PASSES - 1) prefetch_xyz();
      __syncthreads();  // A tile free for the next pass / the reduction
    }
    // ------------------------------ split-K reduction + epilogue ------------------------------
    float* sRed = smem;  // [G][TQ][C]
#pragma unroll
    for (int a = 0; a < RQ; ++a)
      *reinterpret_cast<float4*>(sRed + ((size_t)grp * TQ + a * QROWS + ty) * C + tx * RC) =
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

template <int C, int V, typename IdxT>
int launch_fused(const float* q, const float* s, const void* idx, int row_stride, int H, const float* x, const float* w,
                 const float* kp, const unsigned char* rowflag, float extent, float* out, int nq, int ns,
                 cudaStream_t stream) {
  using K = Cfg<C, V>;
  const int n_tiles = (nq + K::TQ - 1) / K::TQ;
  SPR_CUDA(ensure_max_dynamic_smem(reinterpret_cast<const void*>(k_kpconv_fused<C, V, IdxT>), K::SMEM));
  // persistent CTAs: one (or MIN_BLOCKS) per SM, tiles dealt round-robin, so that the next tile's first loads
  // can be prefetched behind the current tile's contraction
  const int resident = kNumSMs * K::MIN_BLOCKS;
  const int grid = n_tiles < resident ? n_tiles : resident;
  k_kpconv_fused<C, V, IdxT><<<grid, K::THREADS, K::SMEM, stream>>>(q, s, static_cast<const IdxT*>(idx), row_stride, H,
                                                                    x, w, kp, rowflag, extent, out, nq, ns, n_tiles);
  SPR_LAUNCH_CHECK("k_kpconv_fused");
  return SPR_OK;
}

// Tile-shape choice per channel count (measured on B200, profiles/): overridable for experiments with
// SPR_KPCONV_VARIANT=<0|1>.
int pick_variant(int c) {
  static int forced = -2;
  if (forced == -2) {
    const char* e = getenv("SPR_KPCONV_VARIANT");
    forced = e ? atoi(e) : -1;
  }
  if (forced >= 0) return forced;
  (void)c;
  return 0;
}

}  // namespace
}  // namespace spr

using namespace spr;

extern "C" size_t spr_kpconv_workspace_bytes(int nq, int ns, int cin, int cout, int n_kernel_points) {
  (void)nq;
  (void)n_kernel_points;
  const size_t simt = cin == 1 ? align_up((size_t)(ns > 0 ? ns : 0) * 16, 256) + 256
                               : align_up((size_t)(ns > 0 ? ns : 0) + 1, 256) + 256;
  const size_t tc = cin == cout ? kpconv_tc_workspace_bytes(ns > 0 ? ns : 0, cin) : 0;
  return simt > tc ? simt : tc;
}

extern "C" int spr_kpconv_forward(const float* d_q, const float* d_s, const void* d_idx, int idx_is_64, int row_stride,
                                  int H, const float* d_x, int cin, const float* d_w, int cout, const float* d_kp,
                                  int n_kernel_points, float extent, float* d_out, int nq, int ns, int mode,
                                  void* d_workspace, size_t workspace_bytes, const int32_t* d_order, void* stream_) {
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
  if (mode != 0 && mode != 1) {
    set_error("kpconv_forward: mode %d not built in", mode);
    return SPR_EUNSUPPORTED;
  }
  if (cin == 1) {
    SPR_CHECK_ARG(cout == 32 || cout == 64 || cout == 128 || cout == 256,
                  "kpconv_forward: the Cin=1 kernel supports Cout in {32,64,128,256} (got %d)", cout);
    SPR_CHECK_ARG(H <= 96, "kpconv_forward: at most 96 neighbour columns are supported (got %d)", H);
    if (!d_workspace || workspace_bytes < spr_kpconv_workspace_bytes(nq, ns, cin, cout, n_kernel_points)) {
      set_error("kpconv_forward: workspace too small");
      return SPR_ENOSPACE;
    }
    float4* packed = static_cast<float4*>(d_workspace);
    k_pack_points<<<(ns + 255) / 256, 256, 0, stream>>>(d_s, d_x, ns, packed);
    SPR_LAUNCH_CHECK("k_pack_points");
    // thread-per-query kernel (round 2); SPR_STEM_GEN=1 selects the warp-per-query kernel of round 1 for A/B runs
    static const bool stem_gen1 = [] {
      const char* e = getenv("SPR_STEM_GEN");
      return e && e[0] == '1';
    }();
    if (!stem_gen1 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0) {
      int grid_t = (nq + 127) / 128;
      if (grid_t > kNumSMs * 4) grid_t = kNumSMs * 4;
#define SPR_CIN1T(CO)                                                                                                \
  do {                                                                                                               \
    if (idx_is_64)                                                                                                   \
      k_kpconv_cin1_t<long long, CO><<<grid_t, 128, 0, stream>>>(d_q, packed, static_cast<const long long*>(d_idx),  \
                                                                 row_stride, H, d_w, d_kp, extent, d_out, nq, ns,    \
                                                                 d_order);                                           \
    else                                                                                                             \
      k_kpconv_cin1_t<int, CO><<<grid_t, 128, 0, stream>>>(d_q, packed, static_cast<const int*>(d_idx), row_stride, H, \
                                                           d_w, d_kp, extent, d_out, nq, ns, d_order);               \
  } while (0)
      switch (cout) {
        case 32: SPR_CIN1T(32); break;
        case 64: SPR_CIN1T(64); break;
        case 128: SPR_CIN1T(128); break;
        default: SPR_CIN1T(256); break;
      }
#undef SPR_CIN1T
      SPR_LAUNCH_CHECK("k_kpconv_cin1_t");
      return SPR_OK;
    }
    const int warps = 8;
    int grid = (nq + warps - 1) / warps;
    if (grid > kNumSMs * 6) grid = kNumSMs * 6;
    const size_t smem = cout > 64 ? (size_t)KP * cout * 4 : 0;
#define SPR_CIN1(NO)                                                                                              \
  do {                                                                                                            \
    if (idx_is_64)                                                                                                \
      k_kpconv_cin1<long long, NO><<<grid, warps * 32, smem, stream>>>(d_q, packed, static_cast<const long long*>(d_idx), \
                                                                        row_stride, H, d_w, d_kp, extent, d_out, nq, \
                                                                        ns);                                       \
    else                                                                                                          \
      k_kpconv_cin1<int, NO><<<grid, warps * 32, smem, stream>>>(d_q, packed, static_cast<const int*>(d_idx), row_stride, \
                                                                  H, d_w, d_kp, extent, d_out, nq, ns);            \
  } while (0)
    switch (cout) {
      case 32: SPR_CIN1(1); break;
      case 64: SPR_CIN1(2); break;
      case 128: SPR_CIN1(4); break;
      default: SPR_CIN1(8); break;
    }
#undef SPR_CIN1
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
  if (mode == 1)
    return kpconv_tc_forward(d_q, d_s, d_idx, idx_is_64, row_stride, H, d_x, cin, d_w, d_kp, extent, d_out, nq, ns,
                             d_workspace, stream);
  unsigned char* rowflag = static_cast<unsigned char*>(d_workspace);
  k_rowsum_flags<<<(ns + 1 + 7) / 8, 256, 0, stream>>>(d_x, ns, cin, rowflag);
  SPR_LAUNCH_CHECK("k_rowsum_flags");
#define SPR_DISPATCH2(CC, VV)                                                                                       \
  (idx_is_64 ? launch_fused<CC, VV, long long>(d_q, d_s, d_idx, row_stride, H, d_x, d_w, d_kp, rowflag, extent, d_out, \
                                               nq, ns, stream)                                                         \
             : launch_fused<CC, VV, int>(d_q, d_s, d_idx, row_stride, H, d_x, d_w, d_kp, rowflag, extent, d_out, nq,   \
                                         ns, stream))
#define SPR_DISPATCH(CC) \
  case CC:               \
    return pick_variant(CC) == 1 ? SPR_DISPATCH2(CC, 1) : SPR_DISPATCH2(CC, 0);
  switch (cin) {
    SPR_DISPATCH(32)
    SPR_DISPATCH(64)
    SPR_DISPATCH(128)
    SPR_DISPATCH(256)
  }
#undef SPR_DISPATCH2
#undef SPR_DISPATCH
  return SPR_EUNSUPPORTED;
}
