// gemm_tc.cu -- dense layers of the cross-encoder (and any Linear on the path) on the Blackwell tensor cores with
// fp32-level accuracy:   Y[T, N] = act( X[T, K] W[N, K]^T + b ) (+ residual)
// (reference: nn.Linear inside nn.MultiheadAttention in/out projections and the FFN, transformers.py:184-245).
//
// Operands are fp16 (hi, lo) pairs, x = hi + lo to ~22 bits.  The A operand stacks the two halves of a token as
// adjacent rows (2r = hi, 2r+1 = lo), the B operand concatenates [W_hi | W_lo] along N, so ONE tcgen05.mma per
// K step yields all four partial products; the epilogue adds the two rows (adjacent TMEM lanes -> one shuffle) and
// the two column halves.  Both operands live in global memory as ready-made shared-memory images (canonical
// K-major SWIZZLE_128B tiles, see tc05.cuh), so the producer is a bare bulk async copy per tile:
//   A image : [m_tile = token / 64][k_atom = k / 64] x (128 rows x 128 B)         written by the previous kernel
//   W image : [n_tile = n / 128][k_atom][sub = hi | lo] x (128 rows x 128 B)      built once per weight
// K <= 256 (every projection of the cross-encoder except FFN2): the CTA is pinned to ONE n tile, loads that tile's
// whole W image (<= 128 KB) into shared memory once and streams only A tiles afterwards -- the kernel is bound by
// L2 -> SM operand traffic, and this cuts it 3x.  Larger K streams W through the ring with A.
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2-17 = epilogue in two
// sets of 8, one per accumulator buffer (2 x 256 TMEM columns): the tiles of a CTA alternate between the sets, so
// two tiles are drained concurrently while the MMAs of the next run (the epilogue, not the MMA, bounds the encoder's
// tall-skinny GEMMs).
// The epilogue can emit fp32 rows, fp16 hi/lo planes (for the attention kernel), or the A image of the next GEMM.
#include "spr_common.cuh"
#include "tc05.cuh"

namespace spr {
namespace {

using namespace tc;

constexpr int BM_TOK = 64;                 // tokens per tile (128 stacked rows)
constexpr int BN = 128;                    // output columns per tile (256 B rows: hi | lo)
constexpr int A_STAGE = 128 * 128;         // 16 KB
constexpr int B_STAGE = 2 * 128 * 128;     // 32 KB
constexpr int NSTAGES = 3;
constexpr int EPI_SET = 8;                 // warps per epilogue set: 2 per TMEM lane quadrant, each owning 64 tile columns
constexpr int EPI_WARPS = 2 * EPI_SET;     // two sets, one per accumulator buffer: tiles i and i+1 are drained concurrently
constexpr int GEMM_THREADS = (2 + EPI_WARPS) * 32;
constexpr int RES_STAGES = 3;               // A ring depth when W is resident
constexpr int RES_KA = 4;                   // W atoms kept resident (K <= 256)
constexpr int EPI_ROW = 36;                 // floats per staged row (32 + 4 padding: conflict-free 16-byte accesses)
constexpr int EPI_STAGE_BYTES = 16 * EPI_ROW * 4;   // per epilogue warp: 16 tokens x 32 columns
constexpr size_t GEMM_SMEM = 1024 + (size_t)RES_STAGES * A_STAGE + (size_t)RES_KA * B_STAGE + 256 + EPI_WARPS * EPI_STAGE_BYTES;
static_assert((size_t)NSTAGES * (A_STAGE + B_STAGE) <= (size_t)RES_STAGES * A_STAGE + (size_t)RES_KA * B_STAGE, "smem");

enum { OUT_F32 = 0, OUT_PLANES = 1, OUT_AIMG = 2 };

struct GemmArgs {
  const unsigned char* a_img;
  const unsigned char* w_img;
  const float* bias;       // [N] or null
  const float* residual;   // [T, ld_res] or null (OUT_F32 only)
  float* out_f32;          // OUT_F32: [T, ld_out]
  __half* out_hi;          // OUT_PLANES: [T, ld_out] ; OUT_AIMG: image base (as bytes)
  __half* out_lo;          // OUT_PLANES only
  int T, N, K;             // K padded to a multiple of 64 in both images
  int ld_out, ld_res;
  float out_scale;         // 1 / (a_scale * w_scale)
  int relu;
  int n_scaled;            // OUT_PLANES: columns < n_scaled are multiplied by col_scale (query pre-scaling)
  float col_scale;
  float next_scale;        // OUT_AIMG: activation scale of the next GEMM's A operand
  float2* stats16;         // OUT_F32, optional: [ceil(T/16), N] (sum, sum of squares) of each 16-row block of the output
};

// sticky numeric flags of this translation unit (bit 0: an fp16 operand image overflowed), see spr_numeric_flags
__device__ unsigned int g_gemm_flags;

__device__ __forceinline__ void flag_overflow(float amax_scaled) {
  if (!(amax_scaled <= 65504.f)) atomicOr(&g_gemm_flags, SPR_FLAG_FP16_OVERFLOW);  // also catches NaN
}

__global__ void __launch_bounds__(GEMM_THREADS, 1) k_gemm_tc(const GemmArgs g, int mode) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int KA = g.K / 64;
  const bool resident = KA <= RES_KA;
  const int nstages = resident ? RES_STAGES : NSTAGES;
  unsigned char* sA = smem;                       // [nstages] A tiles
  unsigned char* sB = smem + nstages * A_STAGE;   // streaming: [NSTAGES] W stages ; resident: [KA <= 4] W atoms
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + RES_STAGES * A_STAGE + RES_KA * B_STAGE);
  uint64_t* bar_full = bars;            // [<= 8]
  uint64_t* bar_empty = bars + 8;       // [<= 8]
  uint64_t* bar_accf = bars + 16;       // [2] accumulator full
  uint64_t* bar_acce = bars + 18;       // [2] accumulator empty
  uint64_t* bar_w = bars + 20;          // resident W image has landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 22);
  float* s_epi = reinterpret_cast<float*>(smem + RES_STAGES * A_STAGE + RES_KA * B_STAGE + 256);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m_tiles = (g.T + BM_TOK - 1) / BM_TOK, n_tiles = (g.N + BN - 1) / BN;
  // tile walk: resident mode pins the CTA to n tile (blockIdx % n_tiles) and strides over m tiles; streaming mode
  // walks the (m, n) grid with n fastest.  Both are expressed as  tile_i = first + i * step,  i < count.
  const int nt_fixed = resident ? (int)blockIdx.x % n_tiles : 0;
  const int first = resident ? (int)blockIdx.x / n_tiles : (int)blockIdx.x;
  const int step = resident ? (int)gridDim.x / n_tiles : (int)gridDim.x;
  const int limit = resident ? m_tiles : m_tiles * n_tiles;

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_accf[i], 1);
      mbar_init(&bar_acce[i], EPI_SET);
    }
    mbar_init(bar_w, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(s_tmem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      if (resident && first < limit) {
        mbar_arrive_expect_tx(bar_w, (uint32_t)KA * B_STAGE);
        for (int a = 0; a < KA; ++a)
          bulk_g2s(sB + a * B_STAGE, g.w_img + ((size_t)nt_fixed * KA + a) * B_STAGE, B_STAGE, bar_w);
      }
      for (int tile = first; tile < limit; tile += step) {
        const int mt = resident ? tile : tile / n_tiles, nt = resident ? nt_fixed : tile % n_tiles;
        for (int a = 0; a < KA; ++a) {
          mbar_wait_park(&bar_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bar_full[stage], resident ? A_STAGE : A_STAGE + B_STAGE);
          bulk_g2s(sA + stage * A_STAGE, g.a_img + ((size_t)mt * KA + a) * A_STAGE, A_STAGE, &bar_full[stage]);
          if (!resident)
            bulk_g2s(sB + stage * B_STAGE, g.w_img + ((size_t)nt * KA + a) * B_STAGE, B_STAGE, &bar_full[stage]);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      // ONE N = 256 MMA per K step covers [W_hi | W_lo] (the two 128-row halves of a stage are contiguous row
      // groups of the same K-major tile): the A tile is read from shared memory once instead of twice, and SS-mode
      // MMAs are shared-memory-bandwidth bound (measured 104 cycles per N=128 MMA vs 128 per N=256)
      constexpr uint32_t idesc = idesc_f16_f32(128, 256);
      const uint64_t adesc0 = desc_sw128_kmajor(smem_u32(sA));
      const uint64_t bdesc0 = desc_sw128_kmajor(smem_u32(sB));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (resident && first < limit) mbar_wait_park(bar_w, 0);
      for (int tile = first; tile < limit; tile += step, ++it) {
        const int buf = it & 1;
        mbar_wait_park(&bar_acce[buf], ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        for (int a = 0; a < KA; ++a) {
          mbar_wait_park(&bar_full[stage], phase);
          tc_fence_after();
          const int bslot = resident ? a : stage;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = adesc0 + (uint64_t)((stage * A_STAGE + kk * 32) >> 4);
            const uint64_t bd = bdesc0 + (uint64_t)((bslot * B_STAGE + kk * 32) >> 4);
            umma_f16(tmem + buf * 256, ad, bd, idesc, (a | kk) != 0);
          }
          umma_commit(&bar_empty[stage]);
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&bar_accf[buf]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------ epilogue ------------------------------------
    // Software-pipelined over 8-column groups: the TMEM loads of group i+1 and the bias / residual loads of group
    // i are in flight while group i is reduced, so the per-tile latency chain is one TMEM round trip, not sixteen.
    const int qd = warp & 3;                       // TMEM lane quadrant this warp may read
    const int ew = warp - 2;
    const int set = ew / EPI_SET;                  // accumulator buffer (= tile parity) this warp drains
    const int chalf = (ew >> 2) & 1;               // which 64 columns of the tile
    // TMEM lane qd * 32 + lane is stacked row 2r (hi) / 2r + 1 (lo) of token r
    const int half_sel = lane & 1;                 // even lane stores columns c0..c0+3, odd lane c0+4..c0+7
    float amax16 = 0.f;                            // largest |value| this lane wrote into an fp16 output
    for (int it = set, tile = first + set * step; tile < limit; tile += 2 * step, it += 2) {
      const int buf = set;
      const int mt = resident ? tile : tile / n_tiles, nt = resident ? nt_fixed : tile % n_tiles;
      mbar_wait_park(&bar_accf[buf], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16) + buf * 256 + chalf * 64;
      // The accumulator rows are per-lane (lane = stacked token row): writing them directly costs one memory
      // transaction per token per instruction.  Each group of 32 columns is transposed through a small per-warp
      // shared-memory stage so that global accesses (bias, residual, output) are row-contiguous: a warp
      // instruction covers 4 tokens x 128 bytes.
      float* stg = s_epi + (warp - 2) * (EPI_STAGE_BYTES / 4);
#pragma unroll
      for (int grp = 0; grp < 2; ++grp) {
        // the 8 TMEM loads of this group's 32 columns (hi and lo halves) are issued back to back, then ONE wait
        float v1[4][8], v2[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          tmem_ld8(trow + (grp * 4 + i) * 8, v1[i]);
          tmem_ld8(trow + 128 + (grp * 4 + i) * 8, v2[i]);
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i) tmem_ld_fence(v1[i], v2[i]);
        if (grp == 1) {
          // the accumulator is in registers: hand the TMEM buffer back to the MMA warp before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_acce[buf]);
        }
        // bias and residual of this group's coalesced phase are requested first: their latency overlaps the transpose
        const int pcol = nt * BN + chalf * 64 + grp * 32 + (lane & 7) * 4;
        float4 pbias = make_float4(0.f, 0.f, 0.f, 0.f), pres[4];
        if (g.bias && pcol < g.N) pbias = __ldg(reinterpret_cast<const float4*>(g.bias + pcol));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          pres[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int ptok = mt * BM_TOK + qd * 16 + k * 4 + (lane >> 3);
          if (mode == OUT_F32 && g.residual && ptok < g.T && pcol < g.N)
            pres[k] = *reinterpret_cast<const float4*>(g.residual + (size_t)ptok * g.ld_res + pcol);
        }
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          float sum[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            sum[e] = v1[ii][e] + v2[ii][e];
            sum[e] += __shfl_xor_sync(kFull, sum[e], 1);
          }
          const float4 y = half_sel ? make_float4(sum[4], sum[5], sum[6], sum[7]) : make_float4(sum[0], sum[1], sum[2], sum[3]);
          *reinterpret_cast<float4*>(stg + (lane >> 1) * EPI_ROW + ii * 8 + half_sel * 4) = y;
        }
        __syncwarp();
        const int ncol0 = nt * BN + chalf * 64 + grp * 32 + (lane & 7) * 4;  // first of this lane's 4 columns
        float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};      // column sums of the warp's 16 tokens
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int tl = k * 4 + (lane >> 3);                    // token within the warp's 16
          const int token = mt * BM_TOK + qd * 16 + tl;
          const bool ok = token < g.T && ncol0 < g.N;
          if (!ok) continue;
          const float4 a = *reinterpret_cast<const float4*>(stg + tl * EPI_ROW + (lane & 7) * 4);
          float y[4] = {a.x * g.out_scale, a.y * g.out_scale, a.z * g.out_scale, a.w * g.out_scale};
          y[0] += pbias.x; y[1] += pbias.y; y[2] += pbias.z; y[3] += pbias.w;
          if (mode == OUT_F32) {
            y[0] += pres[k].x; y[1] += pres[k].y; y[2] += pres[k].z; y[3] += pres[k].w;
            if (g.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) y[e] = fmaxf(y[e], 0.f);
            }
            *reinterpret_cast<float4*>(g.out_f32 + (size_t)token * g.ld_out + ncol0) = make_float4(y[0], y[1], y[2], y[3]);
            if (g.stats16) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                cs[e] += y[e];
                cq[e] = fmaf(y[e], y[e], cq[e]);
              }
            }
          } else {
            if (g.relu) {
#pragma unroll
              for (int e = 0; e < 4; ++e) y[e] = fmaxf(y[e], 0.f);
            }
            const float sc = mode == OUT_AIMG ? g.next_scale : (ncol0 < g.n_scaled ? g.col_scale : 1.f);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              y[e] *= sc;
              amax16 = fmaxf(amax16, fabsf(y[e]));
            }
            const __half2 h0 = __floats2half2_rn(y[0], y[1]), h1 = __floats2half2_rn(y[2], y[3]);
            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
            const __half2 l0 = __floats2half2_rn(y[0] - f0.x, y[1] - f0.y), l1 = __floats2half2_rn(y[2] - f1.x, y[3] - f1.y);
            const uint2 hv = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            const uint2 lv = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
            if (mode == OUT_PLANES) {
              *reinterpret_cast<uint2*>(g.out_hi + (size_t)token * g.ld_out + ncol0) = hv;
              *reinterpret_cast<uint2*>(g.out_lo + (size_t)token * g.ld_out + ncol0) = lv;
            } else {  // A image of the next GEMM, whose K is this N
              const int KA2 = (g.N + 63) / 64;
              unsigned char* blk = reinterpret_cast<unsigned char*>(g.out_hi) +
                                   ((size_t)(token >> 6) * KA2 + (ncol0 >> 6)) * A_STAGE;
              const uint32_t r2 = 2 * (token & 63), chunk = (ncol0 & 63) >> 3, inb = (ncol0 & 7) * 2;
              *reinterpret_cast<uint2*>(blk + sw128_offset(r2, chunk) + inb) = hv;
              *reinterpret_cast<uint2*>(blk + sw128_offset(r2 + 1, chunk) + inb) = lv;
            }
          }
        }
        if (mode == OUT_F32 && g.stats16) {
          // InstanceNorm statistics of the consumer: fold the 4 token groups (lanes 8 apart) in a fixed order and
          // write (sum, sumsq) of this warp's 16-token block; the normalisation kernel never re-reads the rows
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            cs[e] += __shfl_xor_sync(kFull, cs[e], 8);
            cq[e] += __shfl_xor_sync(kFull, cq[e], 8);
            cs[e] += __shfl_xor_sync(kFull, cs[e], 16);
            cq[e] += __shfl_xor_sync(kFull, cq[e], 16);
          }
          const int tok0 = mt * BM_TOK + qd * 16;
          if (lane < 8 && tok0 < g.T && ncol0 < g.N) {
            float4* dst = reinterpret_cast<float4*>(g.stats16 + (size_t)(tok0 >> 4) * g.N + ncol0);
            dst[0] = make_float4(cs[0], cq[0], cs[1], cq[1]);
            dst[1] = make_float4(cs[2], cq[2], cs[3], cq[3]);
          }
        }
        __syncwarp();
      }
    }
    if (mode != OUT_F32) flag_overflow(amax16);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// LayerNorm over 256 channels (+ positional embedding) -> A image (scaled fp16 hi/lo stacked rows); one warp per token
__global__ void __launch_bounds__(256) k_ln_to_aimg(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, const float* __restrict__ pos,
                                                     int T, float eps, float a_scale, unsigned char* __restrict__ img,
                                                     float* __restrict__ out_f32) {
  const int lane = threadIdx.x & 31;
  const int token = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (token >= T) return;
  const float* xr = x + (size_t)token * 256 + lane * 8;
  const float4 a = *reinterpret_cast<const float4*>(xr), b = *reinterpret_cast<const float4*>(xr + 4);
  float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.f / 256.f);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[i] -= mean;
    q += v[i] * v[i];
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / 256.f) + eps);
  if (gamma) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = v[i] * rstd * __ldg(gamma + lane * 8 + i) + __ldg(beta + lane * 8 + i);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= rstd;
  }
  if (pos) {
    const float* pr = pos + (size_t)token * 256 + lane * 8;
    const float4 pa = *reinterpret_cast<const float4*>(pr), pb = *reinterpret_cast<const float4*>(pr + 4);
    v[0] += pa.x; v[1] += pa.y; v[2] += pa.z; v[3] += pa.w;
    v[4] += pb.x; v[5] += pb.y; v[6] += pb.z; v[7] += pb.w;
  }
  if (out_f32) {
    float* o = out_f32 + (size_t)token * 256 + lane * 8;
    *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (img) {
    uint32_t hi[4], lo[4];
    float amax16 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float y0 = v[2 * i] * a_scale, y1 = v[2 * i + 1] * a_scale;
      amax16 = fmaxf(amax16, fmaxf(fabsf(y0), fabsf(y1)));
      const __half2 hh = __floats2half2_rn(y0, y1);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(y0 - hf.x, y1 - hf.y);
      hi[i] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    // 256 channels = 4 K atoms; lane l owns chunk l % 8 of atom l / 8
    unsigned char* blk = img + ((size_t)(token >> 6) * 4 + (lane >> 3)) * A_STAGE;
    const uint32_t r2 = 2 * (token & 63);
    *reinterpret_cast<uint4*>(blk + sw128_offset(r2, lane & 7)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(blk + sw128_offset(r2 + 1, lane & 7)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    flag_overflow(amax16);
  }
}

// fp32 rows -> A image (generic K, multiple of 8); one thread per 8 consecutive channels
__global__ void __launch_bounds__(256) k_f32_to_aimg(const float* __restrict__ x, int T, int K, int ld, float a_scale,
                                                      unsigned char* __restrict__ img) {
  const int KA = (K + 63) / 64;
  const int chunks_per_row = KA * 8;
  const size_t total = (size_t)T * chunks_per_row;
  float amax16 = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int token = (int)(i / chunks_per_row), ch = (int)(i % chunks_per_row);
    const int k0 = ch * 8;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = (k0 + e < K) ? x[(size_t)token * ld + k0 + e] * a_scale : 0.f;
      amax16 = fmaxf(amax16, fabsf(v[e]));
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
      hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    unsigned char* blk = img + ((size_t)(token >> 6) * KA + (ch >> 3)) * A_STAGE;
    const uint32_t r2 = 2 * (token & 63);
    *reinterpret_cast<uint4*>(blk + sw128_offset(r2, ch & 7)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(blk + sw128_offset(r2 + 1, ch & 7)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  flag_overflow(amax16);
}

// W [N, K] fp32 -> W image (scaled fp16 hi | lo), one thread per 16-byte chunk
__global__ void __launch_bounds__(256) k_weight_to_img(const float* __restrict__ w, int N, int K, float w_scale,
                                                        unsigned char* __restrict__ img) {
  const int KA = (K + 63) / 64, n_tiles = (N + BN - 1) / BN;
  const size_t total = (size_t)n_tiles * KA * 2 * 128 * 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);
    const int r = (int)((i >> 3) & 127);
    const int sub = (int)((i >> 10) & 1);
    const size_t blk = i >> 11;  // nt * KA + a
    const int a = (int)(blk % KA), nt = (int)(blk / KA);
    const int n = nt * BN + r;
    __align__(16) __half h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = a * 64 + j * 8 + e;
      const float v = (n < N && k < K) ? w[(size_t)n * K + k] * w_scale : 0.f;
      const __half hi = __float2half_rn(v);
      h[e] = sub ? __float2half_rn(v - __half2float(hi)) : hi;
    }
    *reinterpret_cast<uint4*>(img + blk * B_STAGE + (size_t)sub * A_STAGE + sw128_offset(r, j)) =
        *reinterpret_cast<const uint4*>(h);
  }
}

}  // namespace

unsigned int gemm_numeric_flags(bool reset) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_gemm_flags, sizeof(v)) != cudaSuccess) return 0;
  if (reset && v) {
    const unsigned int zero = 0;
    cudaMemcpyToSymbol(g_gemm_flags, &zero, sizeof(zero));
  }
  return v;
}

}  // namespace spr

using namespace spr;

extern "C" size_t spr_gemm_a_image_bytes(int T, int K) {
  return (size_t)((T + BM_TOK - 1) / BM_TOK) * ((K + 63) / 64) * A_STAGE;
}
extern "C" size_t spr_gemm_w_image_bytes(int N, int K) {
  return (size_t)((N + BN - 1) / BN) * ((K + 63) / 64) * B_STAGE;
}

extern "C" int spr_gemm_prepare_weight(const float* d_w, int N, int K, float w_scale, void* d_img, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(d_w && d_img && N > 0 && K > 0, "gemm_prepare_weight: bad arguments");
  const size_t total = spr_gemm_w_image_bytes(N, K) / 16;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  k_weight_to_img<<<grid, 256, 0, stream>>>(d_w, N, K, w_scale, static_cast<unsigned char*>(d_img));
  SPR_LAUNCH_CHECK("k_weight_to_img");
  return SPR_OK;
}

extern "C" int spr_gemm_prepare_input(const float* d_x, int T, int K, int ld, float a_scale, void* d_img, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(d_x && d_img && T > 0 && K > 0 && ld >= K, "gemm_prepare_input: bad arguments");
  const size_t total = (size_t)T * ((K + 63) / 64) * 8;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  k_f32_to_aimg<<<grid, 256, 0, stream>>>(d_x, T, K, ld, a_scale, static_cast<unsigned char*>(d_img));
  SPR_LAUNCH_CHECK("k_f32_to_aimg");
  return SPR_OK;
}

extern "C" int spr_layernorm256_prepare(const float* d_x, const float* d_gamma, const float* d_beta, const float* d_pos,
                                        int T, float eps, float a_scale, void* d_img, float* d_out_f32, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(d_x && T > 0 && (d_img || d_out_f32), "layernorm256_prepare: bad arguments");
  SPR_CHECK_ARG((d_gamma == nullptr) == (d_beta == nullptr), "layernorm256_prepare: gamma and beta go together");
  k_ln_to_aimg<<<(T + 7) / 8, 256, 0, stream>>>(d_x, d_gamma, d_beta, d_pos, T, eps, a_scale,
                                                static_cast<unsigned char*>(d_img), d_out_f32);
  SPR_LAUNCH_CHECK("k_ln_to_aimg");
  return SPR_OK;
}

extern "C" int spr_gemm_tc(const void* d_a_img, const void* d_w_img, const float* d_bias, const float* d_residual,
                           int ld_res, int T, int N, int K, float out_scale, int relu, int out_mode, void* d_out,
                           void* d_out_lo, int ld_out, int n_scaled, float col_scale, float next_scale,
                           float* d_stats16, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(d_a_img && d_w_img && d_out && T > 0 && N > 0 && K > 0, "gemm_tc: bad arguments");
  SPR_CHECK_ARG(!d_stats16 || out_mode == OUT_F32, "gemm_tc: block statistics only with fp32 output");
  SPR_CHECK_ARG((N & 3) == 0, "gemm_tc: N must be a multiple of 4 (got %d)", N);
  SPR_CHECK_ARG(out_mode >= 0 && out_mode <= 2, "gemm_tc: unknown output mode %d", out_mode);
  SPR_CHECK_ARG(out_mode != OUT_PLANES || d_out_lo, "gemm_tc: plane output needs both planes");
  SPR_CHECK_ARG(out_mode == OUT_AIMG || (ld_out >= N && (ld_out & 3) == 0), "gemm_tc: bad output row stride");
  SPR_CHECK_ARG(!d_residual || (out_mode == OUT_F32 && (ld_res & 3) == 0), "gemm_tc: residual only with fp32 output");
  // the next GEMM reads whole 64-wide K atoms of the image this one writes: no pad columns may stay unwritten
  SPR_CHECK_ARG(out_mode != OUT_AIMG || (N & 63) == 0, "gemm_tc: operand-image output needs N to be a multiple of 64 (got %d)", N);
  GemmArgs g;
  g.a_img = static_cast<const unsigned char*>(d_a_img);
  g.w_img = static_cast<const unsigned char*>(d_w_img);
  g.bias = d_bias;
  g.residual = d_residual;
  g.out_f32 = static_cast<float*>(d_out);
  g.out_hi = static_cast<__half*>(d_out);
  g.out_lo = static_cast<__half*>(d_out_lo);
  g.T = T;
  g.N = N;
  g.K = (K + 63) / 64 * 64;
  g.ld_out = ld_out;
  g.ld_res = ld_res;
  g.out_scale = out_scale;
  g.relu = relu;
  g.n_scaled = n_scaled;
  g.col_scale = col_scale;
  g.next_scale = next_scale;
  g.stats16 = reinterpret_cast<float2*>(d_stats16);
  SPR_CUDA(ensure_max_dynamic_smem(reinterpret_cast<const void*>(k_gemm_tc), GEMM_SMEM));
  const int m_tiles = (T + BM_TOK - 1) / BM_TOK, n_tiles = (N + BN - 1) / BN;
  int grid;
  if (g.K / 64 <= RES_KA) {  // W-resident walk: a multiple of n_tiles CTAs, each pinned to one n tile
    int per = kNumSMs / n_tiles;
    if (per < 1) per = 1;
    if (per > m_tiles) per = m_tiles;
    grid = per * n_tiles;
  } else {
    const int total = m_tiles * n_tiles;
    grid = total < kNumSMs ? total : kNumSMs;
  }
  k_gemm_tc<<<grid, GEMM_THREADS, GEMM_SMEM, stream>>>(g, out_mode);
  SPR_LAUNCH_CHECK("k_gemm_tc");
  return SPR_OK;
}
