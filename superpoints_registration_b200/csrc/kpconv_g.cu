// kpconv_g.cu -- fused KPConv forward, second generation: BOTH matrix products of the layer run on tcgen05 and the
// neighbour-feature gather is asynchronous, global -> shared memory (reference: kpconv_blocks.py:269-414).
//
//   phase 1 (per query, per pass of 32 input channels)
//        wf[c][k] = sum_h X[idx[h]][c] * I[h][k]          I = kernel-point influences (geometry only)
//     A operand = the gathered, pre-split feature rows AS THEY LIE in global memory: a row is 32 channels x (hi, lo) fp16
//       = 64 contiguous "M elements" (even = hi, odd = lo), one row per neighbour = one K index  ->  MN-major SWIZZLE_128B
//       tile, filled by 16-byte asynchronous copies (cp.async through L2, zero-fill for shadow rows) whose completion
//       arrives on the slot's mbarrier: no feature ever passes through a register.  (The TMA row gather, cp.async.bulk.
//       tensor tile::gather4, produces the same tile -- tools/umma_mn_test.cu -- but sustained only one 128-byte row
//       per ~18 clocks per SM on B200, 5x below what the copies through the LSU path deliver.)
//     B operand = [I_hi | I_lo] (32 rows: 16 kernel points x {hi, lo}; K = neighbours), K-major SWIZZLE_128B, written
//       by the producer warp that evaluated the influences (pass 0) or copied back from an L2-resident scratch by the
//       TMA engine (later passes: the influences depend on geometry only);
//     D1 = 64 rows x 32 columns per query: row 2c / 2c+1 = hi / lo half of channel c, column k / 16+k = I_hi / I_lo.  An
//       M = 64 accumulator occupies 16 lanes of each TMEM lane quadrant (row m -> quadrant m / 16, lane m % 16), so two
//       queries share 32 columns: the slots of a pair use lane offsets 0 and 16;
//   readback: four warps (one per quadrant) take the D1 of a slot pair at once -- lanes 0..15 the first query, 16..31 the
//     second -- sum the four partial products (columns k and 16+k, lanes 2c and 2c+1), split the result into fp16
//     (hi, lo) and store it as two rows of the A tile of phase 2 (canonical K-major SWIZZLE_128B);
//   phase 2 (per tile of 64 queries, per pass): D2[128 x 2C] += A2[128 x 512] * [W_hi | W_lo]^T, weights streamed through
//     a shared-memory ring by the TMA engine -- as in kpconv_tc.cu, with the K index ordered channel-major so that a
//     readback lane writes 16 contiguous bytes;
//   epilogue: four warps (one per TMEM lane quadrant) drain D2 while the next tile is produced.
// Warp roles (2 NSLOT + 16 warps): two readback groups of four warps (slot pair p is served by group p % 2), two
// producer warps per operand slot (A1 + B1, 8-12 KB; each warp evaluates every other 8-neighbour block), the four
// epilogue warps, the phase-2 MMA thread, the weight-stream thread, and one phase-1 MMA thread per readback group (a
// hand-over costs a thread ~100 clocks per mbarrier operation, so the per-query chain is split over two issuers and
// handled a slot PAIR at a time).  Queries of a tile are dealt statically (query i -> slot i % NSLOT), so every hand-over is
// a plain in-order mbarrier wait.
#include "spr_common.cuh"
#include "tc05.cuh"

namespace spr {
namespace {

using namespace tc;

constexpr int KP = 15;
constexpr int kSmemMax = 232448;  // 227 KB of dynamic shared memory per CTA

template <int C, int KS>  // KS = 16-neighbour K steps of phase 1 (H <= 16 KS)
struct GCfg {
  static constexpr int PASSES = C / 32;
  static constexpr int NCOL = 2 * C;
  static constexpr int NS = NCOL < 128 ? NCOL : 128;  // N of one phase-2 MMA = rows of one ring stage
  static constexpr int NSUB = NCOL / NS;
  static constexpr int STAGE_BYTES = NS * 128;
  static constexpr int NSTAGES = 2;
  static constexpr int BLOCKS_PER_PASS = 8 * NSUB;
  static constexpr int TQ = 64;
  static constexpr int NB = 2 * KS;                   // 8-neighbour blocks of a row
  static constexpr int A1_BYTES = KS * 2048;          // 16 KS gathered rows of 128 B
  static constexpr int B1_BYTES = 4096;               // 32 rows x 128 B (K <= 64)
  static constexpr int SLOT_BYTES = A1_BYTES + B1_BYTES;
  static constexpr int A_ATOM_BYTES = 128 * 128;
  static constexpr int A_BYTES = 8 * A_ATOM_BYTES;
  static constexpr int RING_BYTES = NSTAGES * STAGE_BYTES;
  static constexpr int MISC_BYTES = 2048;
  static constexpr int NSLOT_FIT = (kSmemMax - 1024 - A_BYTES - RING_BYTES - MISC_BYTES) / SLOT_BYTES;
  static constexpr int NSLOT = NSLOT_FIT >= 8 ? 8 : (NSLOT_FIT & ~1);
  static constexpr int NRG = 2;                       // readback groups (4 warps each)
  static constexpr int W_PROD = 4 * NRG;              // first producer warp
  static constexpr int W_EPI = W_PROD + 2 * NSLOT;    // first epilogue warp (two producer warps per slot)
  static constexpr int W_MMA = W_EPI + 4;             // phase-2 MMA issuer (+ TMEM owner); W_MMA + 1 = weight stream
  static constexpr int W_ISS = W_MMA + 2;             // two phase-1 MMA issuers, one per readback group
  static constexpr int WARPS = W_ISS + NRG;
  static constexpr int THREADS = WARPS * 32;
  static constexpr int OFF_RING = A_BYTES;
  static constexpr int OFF_SLOTS = OFF_RING + RING_BYTES;
  static constexpr int OFF_MISC = OFF_SLOTS + NSLOT * SLOT_BYTES;
  static constexpr size_t SMEM = 1024 + OFF_MISC + MISC_BYTES;
  static constexpr int D1_COL0 = NCOL;                // D1 of slot s: TMEM columns NCOL + 32 (s / 2), lane offset 16 (s % 2)
  static constexpr size_t IMG_BYTES = (size_t)PASSES * BLOCKS_PER_PASS * STAGE_BYTES;
  static_assert(NSLOT >= 4 && NCOL + 16 * NSLOT <= 512, "slot / TMEM budget");
  static_assert(SMEM <= kSmemMax, "shared memory budget");
};

// ---- PTX pieces that tc05.cuh does not have ---------------------------------------------------------------------
// mbarrier wait that parks the thread in hardware (suspend-time hint) instead of spinning: a polling warp costs issue
// slots -- the first profile of this kernel spent 64 % of its issued instructions in try_wait loops
__device__ __forceinline__ void mbar_wait_park(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra.uni WAIT_DONE;\n\t"
      "bra.uni WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(20000u)
      : "memory");
}
// 16-byte asynchronous copy global -> shared through L2 (LDGSTS); src_bytes = 0 writes zeros
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// one arrival on the mbarrier once every cp.async this thread has issued so far has landed (the barrier's expected
// count must already include it: .noinc)
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// SWIZZLE_128B descriptor of an MN-major operand: 64 contiguous M elements per K row, 8 K rows per 1024-byte group
__device__ __forceinline__ uint64_t desc_sw128_mnmajor(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16_amn(int m, int n) {  // A MN-major, B K-major
  return (1u << 4) | (1u << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  // the wait is tied to the loaded registers so that no use of them can be scheduled above it
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float influence_g(float cx, float cy, float cz, float kx, float ky, float kz, float inv_extent) {
  const float dx = cx - kx, dy = cy - ky, dz = cz - kz;
  const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  float d;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(d2));  // MUFU.SQRT, rel. error ~2^-23, sqrt(0) = 0
  return fmaxf(fmaf(-d, inv_extent, 1.f), 0.f);
}
__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }

// ---------------------------------------------------------------------------------------------
// weight image, channel-major K order.  Block (pass, atom, sub) = NS rows x 128 B (SWIZZLE_128B K-major): row r <->
// column n' = sub*NS + r of [W_hi | W_lo]; K element kk of the atom <-> input channel pass*32 + atom*4 + kk/16, kernel
// point kk%16 (15 = zero padding).  One thread per 16-byte chunk.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) k_weight_image_g(const float* __restrict__ w,
                                                         const unsigned int* __restrict__ amax_w_bits,
                                                         unsigned char* __restrict__ img) {
  constexpr int NCOL = 2 * C, NS = NCOL < 128 ? NCOL : 128, NSUB = NCOL / NS;
  constexpr int CHUNKS = (C / 32) * 8 * NSUB * NS * 8;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= CHUNKS) return;
  const int j = t & 7;
  const int r = (t >> 3) % NS;
  const int blk = t / (8 * NS);
  const int sub = blk % NSUB;
  const int atom = (blk / NSUB) & 7;
  const int pass = blk / (NSUB * 8);
  const int ncol = sub * NS + r;
  const bool lo_part = ncol >= C;
  const int o = lo_part ? ncol - C : ncol;
  const int cin = pass * 32 + atom * 4 + (j >> 1);  // two 8-element chunks per channel
  const float tscale = pow2i(scale_exp(__uint_as_float(*amax_w_bits), 14));
  __align__(16) __half h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = (j & 1) * 8 + e;
    float v = 0.f;
    if (k < KP) v = __ldg(w + ((size_t)k * C + cin) * C + o) * tscale;
    const __half hi = __float2half_rn(v);
    h[e] = lo_part ? __float2half_rn(v - __half2float(hi)) : hi;
  }
  *reinterpret_cast<uint4*>(img + (size_t)blk * (NS * 128) + sw128_offset(r, j)) = *reinterpret_cast<const uint4*>(h);
}

__global__ void __launch_bounds__(256) k_absmax_g(const float* __restrict__ w, int n, unsigned int* __restrict__ amax_bits) {
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
  m = warp_maxf(m);
  if ((threadIdx.x & 31) == 0 && __float_as_uint(m) > *reinterpret_cast<volatile unsigned int*>(amax_bits))
    atomicMax(amax_bits, __float_as_uint(m));
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
#ifdef SPR_G_TRACE
__device__ long long g_trace[8 * 4096];
#define SPR_TR(role, i, v)                                                                         \
  do {                                                                                             \
    if (blockIdx.x == 0 && (i) < 4096 && lane == 0) g_trace[(role) * 4096 + (i)] = (long long)(v); \
  } while (0)
#else
#define SPR_TR(role, i, v) \
  do {                     \
  } while (0)
#endif

template <int C, typename IdxT, int KS>
__global__ void __launch_bounds__(GCfg<C, KS>::THREADS, 1)
    k_kpconv_g(const uint32_t* __restrict__ x16, const float* __restrict__ q, const IdxT* __restrict__ idx,
               int row_stride, int H, const unsigned char* __restrict__ wimg, const float* __restrict__ kp,
               const float4* __restrict__ pts4, const unsigned int* __restrict__ amax_x_bits,
               const unsigned int* __restrict__ amax_w_bits, float extent, float* __restrict__ out, int nq, int ns, int tq,
               int n_tiles, unsigned char* __restrict__ scratch, const int* __restrict__ order) {
  using K = GCfg<C, KS>;
  constexpr int NSLOT = K::NSLOT;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;
  unsigned char* sRing = smem + K::OFF_RING;
  unsigned char* sSlots = smem + K::OFF_SLOTS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_MISC);
  // per slot PAIR (the two slots of a pair are always filled, multiplied and read back together):
  uint64_t* bar_full = bars;               // [4]  operands of both slots in place (4 producer warps + TMA bytes)
  uint64_t* bar_d1full = bars + 8;         // [4]  phase-1 MMAs of the pair complete (D1 valid, slot memory reusable)
  uint64_t* bar_d1free = bars + 16;        // [4]  the four readback warps have D1 in registers
  uint64_t* bar_wfull = bars + 24;         // [NSTAGES] weight ring
  uint64_t* bar_wempty = bars + 26;        // [NSTAGES]
  uint64_t* bar_afull = bars + 28;         // A2 rows of the pass are written (all readback warps)
  uint64_t* bar_done = bars + 29;          // phase-2 MMAs of the pass complete (A2 reusable)
  uint64_t* bar_d2full = bars + 30;        // phase-2 MMAs of the tile's last pass complete (D2 valid)
  uint64_t* bar_d2free = bars + 31;        // the four epilogue warps have drained D2
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 32);
  float* sInv = reinterpret_cast<float*>(s_tmem + 4);  // [3][TQ] 1 / neighbour count, by tile number % 3
  float* sKp = sInv + 3 * K::TQ;                        // [45] (48 reserved)
  float* sCnt = sKp + 48;                               // [8][2] partial neighbour counts of a slot's two producer warps
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bar_full[i], 132);  // per producer lane (4 warps) when its asynchronous copies have landed + 4 x lane 0
      mbar_init(&bar_d1full[i], 1);
      mbar_init(&bar_d1free[i], 4);
    }
    for (int i = 0; i < K::NSTAGES; ++i) {
      mbar_init(&bar_wfull[i], 1);
      mbar_init(&bar_wempty[i], 1);
    }
    mbar_init(bar_afull, 4 * K::NRG);
    mbar_init(bar_done, 1);
    mbar_init(bar_d2full, 1);
    mbar_init(bar_d2free, 4);
    fence_mbar_init();
  }
  constexpr int W_MMA = K::W_MMA;  // the weight-stream warp is W_MMA + 1
  if (warp == W_MMA) tmem_alloc(s_tmem, 512);
  for (int i = tid; i < KP * 3; i += K::THREADS) sKp[i] = kp[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const int es = scale_exp((float)H * __uint_as_float(*amax_x_bits), 15);
  const int et = scale_exp(__uint_as_float(*amax_w_bits), 14);

  if (warp >= K::W_PROD && warp < K::W_EPI) {
    // =========================================== producers ===========================================
    // two warps per slot: warp `half` evaluates the 8-neighbour blocks b = half, half + 2, ... of every query of the
    // slot and copies every other group of four feature rows
    const int slot = (warp - K::W_PROD) >> 1, half = (warp - K::W_PROD) & 1;
    unsigned char* sA1 = sSlots + slot * K::SLOT_BYTES;
    unsigned char* sB1 = sA1 + K::A1_BYTES;
    const float inv_extent = 1.0f / extent;
    const float a_scale = pow2i(es);
    const int g = lane >> 2, t = lane & 3;
    const float k0x = sKp[3 * g], k0y = sKp[3 * g + 1], k0z = sKp[3 * g + 2];
    const float k1x = g < 7 ? sKp[3 * (g + 8)] : 0.f, k1y = g < 7 ? sKp[3 * (g + 8) + 1] : 0.f,
                k1z = g < 7 ? sKp[3 * (g + 8) + 2] : 0.f;
    const float k1_on = g < 7 ? 1.f : 0.f;  // kernel point 15 is padding
    uint32_t use = 0;                        // fills of this slot so far
    // The producer's work list is the sequence of its (tile, pass, query) items.  Its global loads are software
    // pipelined ACROSS items -- the index row and the query point of item i+1 are requested before item i is
    // evaluated, the packed support points of item i+1 right after -- so that no L2 round trip sits on the warp's
    // critical path (three dependent round trips per item otherwise: index row, support points, and the first block).
    struct Work {
      int tile, titer, pass, ql, q0, cnt;
      bool valid;
    };
    struct Loads {
      int j0, j1;          // the neighbour row, lane = column (two rounds of 32)
      float qx, qy, qz;    // the query point
      float4 p0, p1;       // packed support points of this lane's neighbours (x, y, z, +-2^-e), zero when absent
    };
    // A slot pair is filled as a unit: when the first query of a pair exists and the second does not (odd tail of a
    // tile), the second slot gets a dummy fill -- nothing is copied, the stale operands are multiplied, nobody reads
    // the result -- so that every hand-over of a pair has the same number of participants.
    const int pair = slot >> 1;
    auto settle = [&](Work& w) {  // move forward to the next (tile, pass) in which this warp's pair is active
      while (w.valid && w.ql - (slot & 1) >= w.cnt) {
        w.ql = slot;
        if (++w.pass == K::PASSES) {
          w.pass = 0;
          w.tile += gridDim.x;
          ++w.titer;
          w.valid = w.tile < n_tiles;
          w.q0 = w.tile * tq;
          w.cnt = w.valid ? min(nq, w.q0 + tq) - w.q0 : 0;
        }
      }
    };
    auto stage_a = [&](const Work& w, Loads& l) {
      const int qq = w.ql < w.cnt ? w.ql : w.cnt - 1;  // (a dummy fill looks at the tile's last query and ignores it)
      const int n = order ? __ldg(order + w.q0 + qq) : w.q0 + qq;
      l.j0 = l.j1 = -1;
      if (lane < H) l.j0 = (int)__ldg(idx + (size_t)n * row_stride + lane);
      if (KS > 2 && 32 + lane < H) l.j1 = (int)__ldg(idx + (size_t)n * row_stride + 32 + lane);
      l.qx = __ldg(q + 3 * (size_t)n);
      l.qy = __ldg(q + 3 * (size_t)n + 1);
      l.qz = __ldg(q + 3 * (size_t)n + 2);
    };
    auto stage_b = [&](const Work& w, Loads& l) {
      if (l.j0 >= ns) l.j0 = -1;
      if (l.j1 >= ns) l.j1 = -1;
      l.p0 = make_float4(0.f, 0.f, 0.f, 0.f);
      l.p1 = l.p0;
      if (w.pass == 0) {
        if (l.j0 >= 0) l.p0 = __ldg(pts4 + l.j0);
        if (KS > 2 && l.j1 >= 0) l.p1 = __ldg(pts4 + l.j1);
      }
    };
    Work cur;
    cur.tile = blockIdx.x;
    cur.titer = 0;
    cur.pass = 0;
    cur.ql = slot;
    cur.valid = cur.tile < n_tiles;
    cur.q0 = cur.tile * tq;
    cur.cnt = cur.valid ? min(nq, cur.q0 + tq) - cur.q0 : 0;
    settle(cur);
    Loads lc;
    if (cur.valid) {
      stage_a(cur, lc);
      stage_b(cur, lc);
    }
    while (cur.valid) {
      Work nxt = cur;
      nxt.ql += NSLOT;
      settle(nxt);
      Loads ln;
      if (nxt.valid) stage_a(nxt, ln);
      const int pass = cur.pass, ql = cur.ql, titer = cur.titer;
      const bool dummy = ql >= cur.cnt;
      unsigned char* tile_scratch = scratch + ((size_t)blockIdx.x * 2 + (titer & 1)) * K::TQ * K::B1_BYTES;
      const int j0 = lc.j0, j1 = lc.j1;
      uint32_t ih0[KS], ih1[KS], il0[KS], il1[KS];  // this warp's KS blocks
      float fcount = 0.f;
      if (pass == 0) {
        // ---- influences of the 16 kernel points on every neighbour, as fp16 (hi, lo) pairs in registers ----
        // lane (g, t) evaluates kernel points g, g + 8 on neighbours 8 b + 2 t, 8 b + 2 t + 1 of block b; the neighbours'
        // packed points sit in the lanes that loaded them (lane = column) and are fetched by shuffle
        const float qx = lc.qx, qy = lc.qy, qz = lc.qz;
        const unsigned m0 = __ballot_sync(kFull, j0 >= 0), m1 = __ballot_sync(kFull, j1 >= 0);
        const unsigned bm = ((m0 & 0xffu) ? 1u : 0u) | ((m0 & 0xff00u) ? 2u : 0u) | ((m0 & 0xff0000u) ? 4u : 0u) |
                            ((m0 & 0xff000000u) ? 8u : 0u) | ((m1 & 0xffu) ? 16u : 0u) | ((m1 & 0xff00u) ? 32u : 0u) |
                            ((m1 & 0xff0000u) ? 64u : 0u) | ((m1 & 0xff000000u) ? 128u : 0u);
#pragma unroll
        for (int bb = 0; bb < KS; ++bb) {
          const int b = 2 * bb + half;
          ih0[bb] = ih1[bb] = il0[bb] = il1[bb] = 0u;
          if ((bm >> b) & 1u) {  // warp-uniform
            const int src = (b & 3) * 8 + 2 * t;
            const float4 ps = b < 4 ? lc.p0 : lc.p1;
            const float pax = __shfl_sync(kFull, ps.x, src), pay = __shfl_sync(kFull, ps.y, src),
                        paz = __shfl_sync(kFull, ps.z, src), paw = __shfl_sync(kFull, ps.w, src);
            const float pbx = __shfl_sync(kFull, ps.x, src + 1), pby = __shfl_sync(kFull, ps.y, src + 1),
                        pbz = __shfl_sync(kFull, ps.z, src + 1), pbw = __shfl_sync(kFull, ps.w, src + 1);
            const float ax = pax - qx, ay = pay - qy, az = paz - qz;
            const float bx = pbx - qx, by = pby - qy, bz = pbz - qz;
            const float sa = fabsf(paw) * a_scale, sb = fabsf(pbw) * a_scale;  // an absent neighbour has w = 0
            fcount += (paw > 0.f ? 1.f : 0.f) + (pbw > 0.f ? 1.f : 0.f);
            const float f00 = influence_g(ax, ay, az, k0x, k0y, k0z, inv_extent) * sa;
            const float f01 = influence_g(bx, by, bz, k0x, k0y, k0z, inv_extent) * sb;
            const float f10 = influence_g(ax, ay, az, k1x, k1y, k1z, inv_extent) * (sa * k1_on);
            const float f11 = influence_g(bx, by, bz, k1x, k1y, k1z, inv_extent) * (sb * k1_on);
            const __half2 h0 = __floats2half2_rn(f00, f01), h1 = __floats2half2_rn(f10, f11);
            const float2 h0f = __half22float2(h0), h1f = __half22float2(h1);
            ih0[bb] = h2_bits(h0);
            ih1[bb] = h2_bits(h1);
            il0[bb] = h2_bits(__floats2half2_rn(f00 - h0f.x, f01 - h0f.y));
            il1[bb] = h2_bits(__floats2half2_rn(f10 - h1f.x, f11 - h1f.y));
          }
        }
      }
      if (nxt.valid) stage_b(nxt, ln);  // the next item's index row has arrived meanwhile
      // ---- the slot: free once the MMAs of its previous fill have completed ----
      if (slot == 0 && half == 0) SPR_TR(0, use, clock64());
      if (use > 0) mbar_wait_park(&bar_d1full[pair], (use - 1) & 1);
      if (slot == 0 && half == 0) SPR_TR(1, use, clock64());
      if (dummy) {
        cp_async_arrive_noinc(&bar_full[pair]);
        if (lane == 0) mbar_arrive(&bar_full[pair]);
        ++use;
        cur = nxt;
        lc = ln;
        continue;
      }
      if (K::PASSES > 1 && half == 0 && lane == 0) {
        if (pass == 0) bulk_wait_read0();  // the scratch copy of the previous pass-0 fill has left shared memory
        else bulk_wait0();                 // this thread's scratch copies have landed in global memory
      }
      // A1: the neighbours' 128-byte row segments (32 channels x hi/lo) of this pass, one row per neighbour, copied
      // asynchronously through L2; lane = (row % 4, 16-byte chunk); an absent neighbour's row is zero-filled
      {
        const int rsub = lane >> 3, ch = lane & 7;
        const uint32_t a1 = smem_u32(sA1);
        const uint32_t* xcol = x16 + pass * 32 + ch * 4;
#pragma unroll
        for (int ii = 0; ii < 2 * KS; ++ii) {
          const int h = 4 * (2 * ii + half) + rsub;
          const int j = __shfl_sync(kFull, h < 32 ? j0 : j1, h & 31);
          cp_async16(a1 + sw128_offset(h, ch), xcol + (size_t)(j >= 0 ? j : 0) * C, j >= 0 ? 16u : 0u);
        }
      }
      if (pass == 0) {
        // B1: row = kernel point (+16 for the lo half), K element = neighbour 8 b + 2 t (+1)
#pragma unroll
        for (int bb = 0; bb < KS; ++bb) {
          const int b = 2 * bb + half;
          *reinterpret_cast<uint32_t*>(sB1 + sw128_offset(g, b) + 4 * t) = ih0[bb];
          *reinterpret_cast<uint32_t*>(sB1 + sw128_offset(8 + g, b) + 4 * t) = ih1[bb];
          *reinterpret_cast<uint32_t*>(sB1 + sw128_offset(16 + g, b) + 4 * t) = il0[bb];
          *reinterpret_cast<uint32_t*>(sB1 + sw128_offset(24 + g, b) + 4 * t) = il1[bb];
        }
        // a neighbour is replicated over g: count the g == 0 copies (lanes 0..3) of this warp's blocks
        float c = g == 0 ? fcount : 0.f;
        c += __shfl_xor_sync(kFull, c, 1);
        c += __shfl_xor_sync(kFull, c, 2);
        if (lane == 0) sCnt[2 * slot + half] = c;
        if (K::PASSES > 1) fence_proxy_async_smem();  // the scratch copy below reads B1 through the async proxy
        asm volatile("bar.sync %0, 64;" ::"r"(1 + slot) : "memory");  // both warps of the slot have written B1 / sCnt
        if (half == 0 && lane == 0) {
          sInv[(titer % 3) * K::TQ + ql] = 1.f / fmaxf(sCnt[2 * slot] + sCnt[2 * slot + 1], 1.f);
          if (K::PASSES > 1) bulk_s2g(tile_scratch + (size_t)ql * K::B1_BYTES, sB1, K::B1_BYTES);
        }
      } else if (half == 0 && lane == 0) {
        mbar_expect_tx(&bar_full[pair], K::B1_BYTES);
        bulk_g2s(sB1, tile_scratch + (size_t)ql * K::B1_BYTES, K::B1_BYTES, &bar_full[pair]);
      }
      cp_async_arrive_noinc(&bar_full[pair]);
      if (lane == 0) mbar_arrive(&bar_full[pair]);  // release: this warp's shared-memory stores, in program order
      if (slot == 0 && half == 0) SPR_TR(2, use, clock64());
      ++use;
      cur = nxt;
      lc = ln;
    }
    if (K::PASSES > 1 && half == 0 && lane == 0) bulk_wait0();
  } else if (warp < K::W_PROD) {
    // =========================================== readback ===========================================
    const int rg = warp >> 2, qd = warp & 3;   // group, TMEM lane quadrant
    const int second = lane >> 4;               // lanes 0..15: first slot of the pair, 16..31: second slot
    const int cl = 8 * qd + ((lane & 15) >> 1);  // channel (within the pass) of this lane pair: D1 row 2 cl (+1)
    const int odd = lane & 1;                   // even lane: X_hi partials and kernel points 0..7, odd: X_lo and 8..15
    uint32_t par = 0;                           // bit p: parity of the next completion of slot pair p
    uint32_t seq = 0;
    int rtr = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int q0 = tile * tq;
      const int cnt = min(nq, q0 + tq) - q0;
#pragma unroll 1
      for (int pass = 0; pass < K::PASSES; ++pass, ++seq) {
        bool first = true;
#pragma unroll 1
        for (int base = 0; base < cnt; base += NSLOT) {
#pragma unroll 1
          for (int pr = rg; pr < NSLOT / 2; pr += K::NRG) {
            const int qa = base + 2 * pr;
            if (qa >= cnt) continue;
            const bool has_b = qa + 1 < cnt;
            mbar_wait_park(&bar_d1full[pr], (par >> pr) & 1u);
            par ^= 1u << pr;
            tc_fence_after();
            if (warp == 0) SPR_TR(6, rtr, clock64());
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(32 * qd) << 16) + K::D1_COL0 + 32 * pr, v);
            if (warp == 0) SPR_TR(7, rtr, clock64());
            ++rtr;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_d1free[pr]);
            if (first) {  // the A tile still feeds the phase-2 MMAs of the previous pass until bar_done completes
              if (seq > 0) mbar_wait_park(bar_done, (seq - 1) & 1);
              first = false;
            }
            // columns k and 16 + k (I_hi, I_lo), lanes 2c and 2c + 1 (X_hi, X_lo): the even lane finishes kernel points
            // 0..7, the odd lane 8..15
            float mine[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float lo_half = v[k] + v[16 + k], hi_half = v[8 + k] + v[24 + k];
              const float give = odd ? lo_half : hi_half;
              const float keep = odd ? hi_half : lo_half;
              mine[k] = keep + __shfl_xor_sync(kFull, give, 1);
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int p2 = 0; p2 < 4; ++p2) {
              const __half2 hh = __floats2half2_rn(mine[2 * p2], mine[2 * p2 + 1]);
              const float2 hf = __half22float2(hh);
              hi[p2] = h2_bits(hh);
              lo[p2] = h2_bits(__floats2half2_rn(mine[2 * p2] - hf.x, mine[2 * p2 + 1] - hf.y));
            }
            if (!second || has_b) {
              // K index (channel-major) = cl * 16 + k: atom cl / 4, 16-byte chunk (cl % 4) * 2 + odd
              const int ql = qa + second;
              unsigned char* atom = sA + (cl >> 2) * K::A_ATOM_BYTES;
              const uint32_t j = (cl & 3) * 2 + odd;
              *reinterpret_cast<uint4*>(atom + sw128_offset(2 * ql, j)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(atom + sw128_offset(2 * ql + 1, j)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
        }
        if (first && seq > 0) mbar_wait_park(bar_done, (seq - 1) & 1);  // keep the phases of bar_done in step
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_afull);
      }
    }
  } else if (warp < K::W_MMA) {
    // =========================================== epilogue ===========================================
    const int qd = warp & 3;  // TMEM lane quadrant: stacked rows 32 qd .. 32 qd + 31 = queries 16 qd .. 16 qd + 15
    const float o_scale = pow2i(-(es + et));
    int titer = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
      const int q0 = tile * tq;
      const int cnt = min(nq, q0 + tq) - q0;
      mbar_wait_park(bar_d2full, titer & 1);
      tc_fence_after();
      const int ql = qd * 16 + (lane >> 1);
      const bool ok = ql < cnt;
      const int n = ok ? (order ? __ldg(order + q0 + ql) : q0 + ql) : 0;
      const float scale = ok ? sInv[(titer % 3) * K::TQ + ql] * o_scale : 0.f;
      const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < C; c0 += 8) {
        float v1[8], v2[8];
        tmem_ld8(trow + c0, v1);
        tmem_ld8(trow + C + c0, v2);
        tmem_ld_wait(v1, v2);
        float sum[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          sum[i] = v1[i] + v2[i];
          sum[i] += __shfl_xor_sync(kFull, sum[i], 1);
        }
        if (ok) {
          const int off = (lane & 1) * 4;
          const float4 r = (lane & 1) ? make_float4(sum[4] * scale, sum[5] * scale, sum[6] * scale, sum[7] * scale)
                                      : make_float4(sum[0] * scale, sum[1] * scale, sum[2] * scale, sum[3] * scale);
          *reinterpret_cast<float4*>(out + (size_t)n * C + c0 + off) = r;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_d2free);
    }
  } else if (warp == W_MMA) {
    // ===================================== phase-2 MMA issuer =====================================
    if (lane == 0) {
      constexpr uint32_t idesc2 = idesc_f16_f32(128, K::NS);
      const uint64_t adesc0 = desc_sw128_kmajor(smem_u32(sA));
      const uint64_t bdesc0 = desc_sw128_kmajor(smem_u32(sRing));
      uint32_t seq = 0;
      int stage = 0;
      uint32_t phase = 0;
      int titer = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
        const int q0 = tile * tq;
        const int cnt = min(nq, q0 + tq) - q0;
        for (int pass = 0; pass < K::PASSES; ++pass, ++seq) {
          // ---- phase 2: D2 += A2 * [W_hi | W_lo]^T for this pass ----
          mbar_wait_park(bar_afull, seq & 1);
          if (pass == 0 && titer > 0) mbar_wait_park(bar_d2free, (titer - 1) & 1);
          tc_fence_after();
          for (int a = 0; a < 8; ++a) {
            for (int sub = 0; sub < K::NSUB; ++sub) {
              mbar_wait_park(&bar_wfull[stage], phase);
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ad = adesc0 + (uint64_t)((a * K::A_ATOM_BYTES + kk * 32) >> 4);
                const uint64_t bd = bdesc0 + (uint64_t)((stage * K::STAGE_BYTES + kk * 32) >> 4);
                umma_f16(tmem + sub * K::NS, ad, bd, idesc2, (pass | a | kk) != 0);
              }
              umma_commit(&bar_wempty[stage]);
              if (++stage == K::NSTAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          umma_commit(bar_done);
          if (pass == K::PASSES - 1) umma_commit(bar_d2full);
        }
      }
    }
    __syncwarp();
  } else if (warp >= K::W_ISS) {
    // ===================================== phase-1 MMA issuers =====================================
    // issuer g serves the slot pairs p = g, g + 2, ... (the pairs of readback group g), in the order the queries were
    // dealt: per pair one wait for the operands, one for the previous readback, 2 KS MMAs, one commit
    if (lane == 0) {
      const int rg = warp - K::W_ISS;
      constexpr uint32_t idesc1 = idesc_f16_amn(64, 32);
      uint32_t par = 0, used = 0;  // per pair: parity of the next fill, pair filled before
      int trq = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int q0 = tile * tq;
        const int cnt = min(nq, q0 + tq) - q0;
        for (int pass = 0; pass < K::PASSES; ++pass) {
          for (int base = 0; base < cnt; base += NSLOT) {
#pragma unroll
            for (int pr = 0; pr < NSLOT / 2; ++pr) {
              if ((pr % K::NRG) != rg || base + 2 * pr >= cnt) continue;
              const uint32_t p = (par >> pr) & 1u;
              mbar_wait_park(&bar_full[pr], p);
              if (rg == 0) SPR_TR(3, trq, clock64());
              if ((used >> pr) & 1u) mbar_wait_park(&bar_d1free[pr], p ^ 1u);  // readback of the previous fill has D1
              if (rg == 0) SPR_TR(4, trq, clock64());
              fence_proxy_async_smem();  // the operands were written through the generic proxy (cp.async, st.shared)
              tc_fence_after();
#pragma unroll
              for (int o = 0; o < 2; ++o) {
                const uint32_t a1 = smem_u32(sSlots + (2 * pr + o) * K::SLOT_BYTES);
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                  umma_f16(tmem + ((uint32_t)(o * 16) << 16) + K::D1_COL0 + 32 * pr, desc_sw128_mnmajor(a1 + ks * 2048),
                           desc_sw128_kmajor(a1 + K::A1_BYTES + ks * 32), idesc1, ks != 0);
              }
              umma_commit(&bar_d1full[pr]);
              if (rg == 0) SPR_TR(5, trq, clock64());
              ++trq;
              par ^= 1u << pr;
              used |= 1u << pr;
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // =========================================== weight stream ===========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int blk = 0; blk < K::PASSES * K::BLOCKS_PER_PASS; ++blk) {
          mbar_wait_park(&bar_wempty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bar_wfull[stage], K::STAGE_BYTES);
          bulk_g2s(sRing + stage * K::STAGE_BYTES, wimg + (size_t)blk * K::STAGE_BYTES, K::STAGE_BYTES,
                   &bar_wfull[stage]);
          if (++stage == K::NSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA) tmem_dealloc(tmem, 512);
}

// ---- host ----------------------------------------------------------------------------------------------------------
template <int C, typename IdxT>
int launch_g(const float* q, const void* idx, int row_stride, int H, const uint32_t* x16, const unsigned char* img,
             const float* kp, const float4* pts4, const unsigned int* amax_x_bits, const unsigned int* amax_w_bits,
             float extent, float* out, int nq, int ns, void* scratch, const int* order, cudaStream_t stream) {
  SPR_CHECK_ARG(H <= 64, "kpconv_forward_gather: at most 64 neighbour columns are supported (got %d)", H);
  SPR_CHECK_ARG(C <= 32 || scratch, "kpconv_forward_gather: scratch buffer missing");
  int tq = 64;
  {
    const int waves = (nq + kNumSMs * 64 - 1) / (kNumSMs * 64);
    const int per = (nq + kNumSMs * waves - 1) / (kNumSMs * waves);
    tq = per < 16 ? 16 : (per > 64 ? 64 : per);
  }
  const int n_tiles = (nq + tq - 1) / tq;
  const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  const IdxT* idx_t = static_cast<const IdxT*>(idx);
  unsigned char* scr = static_cast<unsigned char*>(scratch);
#define SPR_G(KS_)                                                                                                        \
  do {                                                                                                                    \
    using K = GCfg<C, KS_>;                                                                                               \
    SPR_CUDA(ensure_max_dynamic_smem(reinterpret_cast<const void*>(k_kpconv_g<C, IdxT, KS_>), K::SMEM));                  \
    k_kpconv_g<C, IdxT, KS_><<<grid, K::THREADS, K::SMEM, stream>>>(x16, q, idx_t, row_stride, H, img, kp, pts4,         \
                                                                    amax_x_bits, amax_w_bits, extent, out, nq, ns, tq,    \
                                                                    n_tiles, scr, order);                                 \
  } while (0)
  if (H <= 32)
    SPR_G(2);
  else if (H <= 48)
    SPR_G(3);
  else
    SPR_G(4);
#undef SPR_G
  SPR_LAUNCH_CHECK("k_kpconv_g");
  return SPR_OK;
}

template <int C>
int prepare_weights_g(const float* w, unsigned char* img, unsigned int* amax_w_bits, cudaStream_t stream) {
  using K = GCfg<C, 3>;
  SPR_CUDA(cudaMemsetAsync(amax_w_bits, 0, sizeof(unsigned int), stream));
  k_absmax_g<<<(KP * C * C + 1023) / 1024, 256, 0, stream>>>(w, KP * C * C, amax_w_bits);
  SPR_LAUNCH_CHECK("k_absmax_g");
  constexpr int chunks = (int)(K::IMG_BYTES / 16);
  k_weight_image_g<C><<<(chunks + 255) / 256, 256, 0, stream>>>(w, amax_w_bits, img);
  SPR_LAUNCH_CHECK("k_weight_image_g");
  return SPR_OK;
}

}  // namespace
}  // namespace spr

using namespace spr;

#ifdef SPR_G_TRACE
extern "C" __attribute__((visibility("default"))) int spr_kpconv_g_trace(long long* host_buf) {
  return (int)cudaMemcpyFromSymbol(host_buf, g_trace, sizeof(long long) * 8 * 4096);
}
#endif

extern "C" int spr_kpconv_gather_supported(int c, int H) { return (c == 32 || c == 64 || c == 128) && H > 0 && H <= 64; }

extern "C" size_t spr_kpconv_gather_weight_image_bytes(int c) {
  switch (c) {
    case 32: return GCfg<32, 3>::IMG_BYTES;
    case 64: return GCfg<64, 3>::IMG_BYTES;
    case 128: return GCfg<128, 3>::IMG_BYTES;
  }
  return 0;
}

// per CTA: two tile buffers x 64 queries x one B1 image (the influence operand of phase 1, reused by the later passes)
extern "C" size_t spr_kpconv_gather_scratch_bytes(int c) {
  if (c <= 32) return 0;
  return (size_t)kNumSMs * 2 * 64 * 4096;
}

extern "C" int spr_kpconv_gather_prepare_weights(const float* d_w, int c, void* d_img, void* d_amax_w, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(d_w && d_img && d_amax_w, "kpconv_gather_prepare_weights: null pointer");
  unsigned char* img = static_cast<unsigned char*>(d_img);
  unsigned int* am = static_cast<unsigned int*>(d_amax_w);
  switch (c) {
    case 32: return prepare_weights_g<32>(d_w, img, am, stream);
    case 64: return prepare_weights_g<64>(d_w, img, am, stream);
    case 128: return prepare_weights_g<128>(d_w, img, am, stream);
  }
  set_error("kpconv_gather_prepare_weights: unsupported channel count %d", c);
  return SPR_EUNSUPPORTED;
}

extern "C" int spr_kpconv_forward_gather(const float* d_q, const void* d_idx, int idx_is_64, int row_stride, int H,
                                         const void* d_pts4, const void* d_x16, const void* d_amax_x, int c,
                                         const void* d_wimg, const void* d_amax_w, const float* d_kp, float extent,
                                         float* d_out, int nq, int ns, void* d_scratch, const int32_t* d_order,
                                         void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(nq > 0 && ns > 0 && H > 0 && row_stride >= H, "kpconv_forward_gather: bad shape");
  SPR_CHECK_ARG(extent > 0.f, "kpconv_forward_gather: extent must be > 0");
  SPR_CHECK_ARG(d_q && d_idx && d_pts4 && d_x16 && d_amax_x && d_wimg && d_amax_w && d_kp && d_out,
                "kpconv_forward_gather: null pointer");
  const float4* pts4 = static_cast<const float4*>(d_pts4);
  const uint32_t* x16 = static_cast<const uint32_t*>(d_x16);
  const unsigned char* img = static_cast<const unsigned char*>(d_wimg);
  const unsigned int* ax = static_cast<const unsigned int*>(d_amax_x);
  const unsigned int* aw = static_cast<const unsigned int*>(d_amax_w);
#define SPR_GP(CC)                                                                                                     \
  case CC:                                                                                                             \
    return idx_is_64 ? launch_g<CC, long long>(d_q, d_idx, row_stride, H, x16, img, d_kp, pts4, ax, aw, extent, d_out, \
                                               nq, ns, d_scratch, d_order, stream)                                    \
                     : launch_g<CC, int>(d_q, d_idx, row_stride, H, x16, img, d_kp, pts4, ax, aw, extent, d_out, nq,   \
                                         ns, d_scratch, d_order, stream);
  switch (c) {
    SPR_GP(32)
    SPR_GP(64)
    SPR_GP(128)
  }
#undef SPR_GP
  set_error("kpconv_forward_gather: unsupported channel count %d", c);
  return SPR_EUNSUPPORTED;
}
