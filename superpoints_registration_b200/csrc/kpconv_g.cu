// kpconv_g.cu -- fused KPConv forward, second generation (opt-in, SPR_KPCONV_GEN=2): BOTH matrix products of the layer
// run on tcgen05 and the neighbour-feature gather is asynchronous, global -> shared memory (reference:
// kpconv_blocks.py:269-414).  Same results as kpconv_tc.cu within fp32 rounding; measured 0.54-0.68x its speed on B200
// (profiles/r2b_kpconv_gen_bench.log) -- see "why it is slower" below -- so generation 1 stays the default.
//
//   phase 1 (per query, per pass of 32 input channels)
//        wf[c][k] = sum_h X[idx[h]][c] * I[h][k]          I = kernel-point influences (geometry only)
//     A operand = the gathered feature rows in the PLANAR pre-split format (per 32-channel group 32 fp16 hi halves, then
//       32 lo halves): per neighbour a 64-byte hi row and a 64-byte lo row = 32 contiguous "M elements" per K index  ->
//       two MN-major SWIZZLE_64B tiles, filled by 16-byte asynchronous copies (cp.async through L2; rows beyond the
//       last neighbour keep stale finite data and meet zero influences) whose completion arrives on the slot pair's
//       mbarrier: no feature ever passes through a register.  (The TMA row gather, cp.async.bulk.tensor tile::gather4,
//       produces such a tile too -- tools/umma_mn_test.cu -- but sustained only one 128-byte row per ~18 clocks per SM
//       on B200, 5x below what the copies through the LSU path deliver.)
//     B operand = [I_hi | I_lo] (32 rows: 16 kernel points x {hi, lo}; K = neighbours), K-major SWIZZLE_128B, written
//       by the producer warp that evaluated the influences (pass 0) or took them back from an L2-resident scratch
//       (later passes: the influences depend on geometry only);
//     D1 = 32 useful rows x 32 columns per query: row c = channel, column k / 16+k = I_hi / I_lo product; the hi and the
//       lo feature halves are two MMA chains accumulating into the same D1.  An M = 64 accumulator occupies 16 lanes of
//       each TMEM lane quadrant (row m -> quadrant m / 16, lane m % 16), so two queries share 32 columns: the slots of a
//       pair use lane offsets 0 and 16;
//   readback: two warps (TMEM lane quadrants 0 and 1) take the D1 of a slot pair at once -- lanes 0..15 the first query,
//     16..31 the second; thread = channel -- add columns k and 16+k, split the result into fp16 (hi, lo) and store it as
//     two rows of the A tile of phase 2 (canonical K-major SWIZZLE_128B);
//   phase 2 (per tile of 64 queries, per pass): D2[128 x 2C] += A2[128 x 512] * [W_hi | W_lo]^T, weights streamed through
//     a shared-memory ring by the TMA engine -- as in kpconv_tc.cu, with the K index ordered channel-major so that a
//     readback thread writes 16 contiguous bytes;
//   epilogue: four warps (one per TMEM lane quadrant) drain D2 while the next tile is produced.
// Warp roles (2 NSLOT + 16 warps): up to four readback groups of two warps (group p serves slot pair p), two producer
// warps per operand slot taking its fills strictly in turn, the four epilogue warps, the phase-2 MMA thread, the weight-
// stream thread, and two phase-1 MMA threads (a hand-over costs a thread ~100 clocks per mbarrier operation, so the
// per-query chain is split over two issuers and handled a slot PAIR at a time).  Queries of a tile are dealt statically
// (query i -> slot i % NSLOT), so every hand-over is a plain in-order mbarrier wait.
//
// Why it is slower than generation 1 (measured, B200): a tcgen05.mma instruction occupies the tensor pipe for at least
// ~44 clocks whatever its shape (tools/umma_rate_test.cu, profiles/r2b_umma_rate.log: M 64/128, N 32/64, K- or MN-major
// A all cost 44-48 clk; N = 128 costs 64, N = 256 128).  Phase 1 is block-diagonal -- every query has its own A and its
// own B -- so it needs 2 KS (4-8) instructions per query and pass: >= 260 clocks of tensor pipe per query-pass at H = 40,
// where generation 1 spends ~450 clocks per query-pass in total with its warp-level mma.sync phase 1.  The hand-overs
// (producer -> issuer -> readback -> producer, ~100 clk per mbarrier operation) add a latency chain of ~5000 clocks per
// slot fill that eight slots do not hide.
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
  static constexpr int A1_PART = KS * 1024;           // 16 KS rows of 64 B: the hi (or lo) halves of 32 channels
  static constexpr int A1_BYTES = 2 * A1_PART;
  static constexpr int B1_BYTES = 4096;               // 32 rows x 128 B (K <= 64)
  static constexpr int SLOT_BYTES = A1_BYTES + B1_BYTES;
  static constexpr int A_ATOM_BYTES = 128 * 128;
  static constexpr int A_BYTES = 8 * A_ATOM_BYTES;
  static constexpr int RING_BYTES = NSTAGES * STAGE_BYTES;
  static constexpr int MISC_BYTES = 2048;
  static constexpr int NSLOT_FIT = (kSmemMax - 1024 - A_BYTES - RING_BYTES - MISC_BYTES) / SLOT_BYTES;
  static constexpr int NSLOT = NSLOT_FIT >= 8 ? 8 : (NSLOT_FIT & ~1);
  static constexpr int NPAIR = NSLOT / 2;
  // warps 0..15: (w & 3) < 2 -> readback group w >> 2 (TMEM lane quadrant w & 3), else producer; then the remaining
  // producers, four epilogue warps, the phase-2 MMA thread, the weight stream, two phase-1 MMA threads
  static constexpr int NPROD = 2 * NSLOT;
  static constexpr int W_EPI = 16 + (NPROD - 8);
  static constexpr int W_MMA = W_EPI + 4;
  static constexpr int W_ISS = W_MMA + 2;
  static constexpr int WARPS = W_ISS + 2;
  static constexpr int THREADS = WARPS * 32;
  static constexpr int OFF_RING = A_BYTES;
  static constexpr int OFF_SLOTS = OFF_RING + RING_BYTES;
  static constexpr int OFF_MISC = OFF_SLOTS + NSLOT * SLOT_BYTES;
  static constexpr size_t SMEM = 1024 + OFF_MISC + MISC_BYTES;
  static constexpr int D1_COL0 = NCOL;                // D1 of slot pair p: TMEM columns NCOL + 32 p, lane offsets 0 / 16
  static constexpr size_t IMG_BYTES = (size_t)PASSES * BLOCKS_PER_PASS * STAGE_BYTES;
  static_assert(NSLOT >= 4 && NSLOT <= 8 && NCOL + 16 * NSLOT <= 512, "slot / TMEM budget");
  static_assert(SMEM <= kSmemMax, "shared memory budget");
};

// ---- PTX pieces that tc05.cuh does not have ---------------------------------------------------------------------
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
// SWIZZLE_64B descriptor of an MN-major operand: 32 contiguous M elements (64 B) per K row, 8 K rows per 512-byte
// group; the leading byte offset (M elements 32..63 of an M = 64 MMA) is 0: those accumulator rows repeat rows 0..31 and
// are ignored (tools/umma_sw64_test.cu)
__device__ __forceinline__ uint64_t desc_sw64_mnmajor(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// byte offset of 16-byte chunk j (0..3) of row r inside a (rows x 64 B) SWIZZLE_64B tile
__host__ __device__ constexpr uint32_t sw64_offset(uint32_t r, uint32_t j) {
  return (r >> 3) * 512u + (r & 7u) * 64u + ((j ^ ((r >> 1) & 3u)) << 4);
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
// s * max(0, 1 - d / extent) with s and s / extent given: one instruction less per influence than scaling afterwards
__device__ __forceinline__ float influence_scaled(float cx, float cy, float cz, float kx, float ky, float kz, float s_ie,
                                                  float s) {
  const float dx = cx - kx, dy = cy - ky, dz = cz - kz;
  const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  float d;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(d2));  // MUFU.SQRT, rel. error ~2^-23, sqrt(0) = 0
  return fmaxf(fmaf(-d, s_ie, s), 0.f);
}
__device__ __forceinline__ void st_shared_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
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
    k_kpconv_g(const unsigned char* __restrict__ x16p, const float* __restrict__ q, const IdxT* __restrict__ idx,
               int row_stride, int H, const unsigned char* __restrict__ wimg, const float* __restrict__ kp,
               const float4* __restrict__ pts4, const unsigned int* __restrict__ amax_x_bits,
               const unsigned int* __restrict__ amax_w_bits, float extent, float* __restrict__ out, int nq, int ns, int tq,
               int n_tiles, unsigned char* __restrict__ scratch, const int* __restrict__ order) {
  using K = GCfg<C, KS>;
  constexpr int NSLOT = K::NSLOT, NPAIR = K::NPAIR;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;
  unsigned char* sRing = smem + K::OFF_RING;
  unsigned char* sSlots = smem + K::OFF_SLOTS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_MISC);
  // per slot PAIR (the two slots of a pair are always filled, multiplied and read back together):
  uint64_t* bar_full = bars;               // [4]  operands of both slots in place (2 producer warps + TMA bytes)
  uint64_t* bar_d1full = bars + 8;         // [4][2] phase-1 MMAs of the pair complete (D1 valid, slot memory reusable);
                                           //        fills alternate between the two barriers of a pair, so that a
                                           //        producer warp that takes every other fill waits on consecutive phases
  uint64_t* bar_d1free = bars + 16;        // [4]  the two readback warps have D1 in registers
  uint64_t* bar_wfull = bars + 24;         // [NSTAGES] weight ring
  uint64_t* bar_wempty = bars + 26;        // [NSTAGES]
  uint64_t* bar_afull = bars + 28;         // A2 rows of the pass are written (all readback warps)
  uint64_t* bar_done = bars + 29;          // phase-2 MMAs of the pass complete (A2 reusable)
  uint64_t* bar_d2full = bars + 30;        // phase-2 MMAs of the tile's last pass complete (D2 valid)
  uint64_t* bar_d2free = bars + 31;        // the four epilogue warps have drained D2
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 32);
  float* sInv = reinterpret_cast<float*>(s_tmem + 4);  // [3][TQ] 1 / neighbour count, by tile number % 3
  float* sKp = sInv + 3 * K::TQ;                        // [45] (48 reserved)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) {
      mbar_init(&bar_full[i], 66);   // per fill: two producer warps (one per slot), 32 copy arrivals + lane 0 each
      mbar_init(&bar_d1full[2 * i], 1);
      mbar_init(&bar_d1full[2 * i + 1], 1);
      mbar_init(&bar_d1free[i], 2);
    }
    for (int i = 0; i < K::NSTAGES; ++i) {
      mbar_init(&bar_wfull[i], 1);
      mbar_init(&bar_wempty[i], 1);
    }
    mbar_init(bar_afull, 8);
    mbar_init(bar_done, 1);
    mbar_init(bar_d2full, 1);
    mbar_init(bar_d2free, 4);
    fence_mbar_init();
  }
  constexpr int W_MMA = K::W_MMA;  // the weight-stream warp is W_MMA + 1
  if (warp == W_MMA) tmem_alloc(s_tmem, 512);
  for (int i = tid; i < KP * 3; i += K::THREADS) sKp[i] = kp[i];
  // the operand slots start as zeros: rows the producers skip (beyond a query's last neighbour) are multiplied by
  // zero influences and must therefore hold finite values from the first use on
  for (int i = tid; i < NSLOT * K::SLOT_BYTES / 16; i += K::THREADS)
    reinterpret_cast<uint4*>(sSlots)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const int es = scale_exp((float)H * __uint_as_float(*amax_x_bits), 15);
  const int et = scale_exp(__uint_as_float(*amax_w_bits), 14);

  const bool is_reader = warp < 16 && (warp & 3) < 2;
  const int prod_index = warp < 16 ? ((warp & 3) >= 2 ? (warp >> 2) * 2 + (warp & 3) - 2 : -1)
                                   : (warp < K::W_EPI ? 8 + (warp - 16) : -1);

  if (prod_index >= 0 && prod_index < K::NPROD) {
    // =========================================== producers ===========================================
    // Two warps share a slot and take its fills strictly in turn (fill number % 2, counted over the whole kernel: a warp
    // waiting for the other warp's fill must never be two barrier phases behind): while one warp's operands are
    // multiplied the other evaluates the next query.  Per query: kernel-point influences -> B1 (pass 0) or scratch -> B1
    // (later passes); feature rows -> A1.
    const int slot = prod_index >> 1, half = prod_index & 1, pair = slot >> 1;
    unsigned char* sA1 = sSlots + slot * K::SLOT_BYTES;
    unsigned char* sB1 = sA1 + K::A1_BYTES;
    const float inv_extent = 1.0f / extent;
    const float a_scale = pow2i(es);
    const int g = lane >> 2, t = lane & 3;
    const float k0x = sKp[3 * g], k0y = sKp[3 * g + 1], k0z = sKp[3 * g + 2];
    // kernel point 15 is padding: it sits infinitely far away, so that its influence is exactly 0
    const float k1x = g < 7 ? sKp[3 * (g + 8)] : 1.0e18f, k1y = g < 7 ? sKp[3 * (g + 8) + 1] : 0.f,
                k1z = g < 7 ? sKp[3 * (g + 8) + 2] : 0.f;
    // copy role of this lane: neighbour row 4 i + rsub, hi (0) or lo (1) half, 16-byte chunk ch of the 64-byte half row
    const int rsub = lane >> 3, part = (lane >> 2) & 1, ch = lane & 3;
    uint32_t dst_par[2];
#pragma unroll
    for (int e = 0; e < 2; ++e)
      dst_par[e] = smem_u32(sA1) + part * K::A1_PART + ((4 * e + rsub) & 7) * 64 + ((ch ^ ((2 * e + (rsub >> 1)) & 3)) << 4);
    const unsigned char* src_lane = x16p + part * 64 + ch * 16;
    const uint32_t b1_lane = smem_u32(sB1) + g * 128 + 4 * t;  // + ((b ^ g) << 4) + 1024 * {0, 1, 2, 3}

    struct Work {
      int tile, titer, pass, r, q0, cnt;
      uint32_t ubase;  // fills of the slot pair before this (tile, pass)
      bool valid;
    };
    struct Loads {
      int j0, j1;          // the neighbour row, lane = column (two rounds of 32)
      float qx, qy, qz;    // the query point
      float4 p0, p1;       // packed support points of this lane's neighbours (x, y, z, +-2^-e), zero when absent
    };
    // rounds of a (tile, pass) in which the pair is filled: r * NSLOT + 2 * pair < cnt
    auto rounds = [&](int cnt) { return cnt > 2 * pair ? (cnt - 2 * pair + NSLOT - 1) / NSLOT : 0; };
    auto settle = [&](Work& w) {
      while (w.valid && w.r >= rounds(w.cnt)) {
        w.ubase += rounds(w.cnt);
        w.r = (half ^ w.ubase) & 1;
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
      int ql = w.r * NSLOT + slot;
      ql = ql < w.cnt ? ql : w.cnt - 1;  // (a dummy fill looks at the tile's last query and ignores it)
      const int n = order ? __ldg(order + w.q0 + ql) : w.q0 + ql;
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
    cur.r = half;
    cur.ubase = 0;
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
      nxt.r += 2;
      settle(nxt);
      Loads ln;
      if (nxt.valid) stage_a(nxt, ln);
      const int pass = cur.pass, ql = cur.r * NSLOT + slot, titer = cur.titer;
      const uint32_t use = cur.ubase + cur.r;
      const bool dummy = ql >= cur.cnt;
      unsigned char* tile_scratch = scratch + ((size_t)blockIdx.x * 2 + (titer & 1)) * K::TQ * K::B1_BYTES;
      const int j0 = lc.j0, j1 = lc.j1;
      uint32_t ih0[K::NB], ih1[K::NB], il0[K::NB], il1[K::NB];
      const unsigned m0 = __ballot_sync(kFull, j0 >= 0), m1 = __ballot_sync(kFull, j1 >= 0);
      int n_rows = 0;  // neighbour columns up to the last present one (rows beyond keep their stale, finite contents)
      if (m0) n_rows = 32 - __clz(m0);
      if (m1) n_rows = 64 - __clz(m1);
      if (pass == 0 && !dummy) {
        // ---- influences of the 16 kernel points on every neighbour, as fp16 (hi, lo) pairs in registers ----
        // a lane first prepares ITS neighbours (lane = column): offset from the query, influence scale s = 2^-e * a_scale
        // and s / extent; lane (g, t) then evaluates kernel points g, g + 8 on neighbours 8 b + 2 t, 8 b + 2 t + 1 of
        // every block b, fetching the prepared values by shuffle
        const float r0x = lc.p0.x - lc.qx, r0y = lc.p0.y - lc.qy, r0z = lc.p0.z - lc.qz;
        const float r1x = lc.p1.x - lc.qx, r1y = lc.p1.y - lc.qy, r1z = lc.p1.z - lc.qz;
        const float s0 = fabsf(lc.p0.w) * a_scale, s1 = fabsf(lc.p1.w) * a_scale;  // an absent neighbour has w = 0
        const int fcount = __popc(__ballot_sync(kFull, lc.p0.w > 0.f)) + __popc(__ballot_sync(kFull, lc.p1.w > 0.f));
        if (lane == 0) sInv[(titer % 3) * K::TQ + ql] = 1.f / (float)max(fcount, 1);
        const unsigned bm = ((m0 & 0xffu) ? 1u : 0u) | ((m0 & 0xff00u) ? 2u : 0u) | ((m0 & 0xff0000u) ? 4u : 0u) |
                            ((m0 & 0xff000000u) ? 8u : 0u) | ((m1 & 0xffu) ? 16u : 0u) | ((m1 & 0xff00u) ? 32u : 0u) |
                            ((m1 & 0xff0000u) ? 64u : 0u) | ((m1 & 0xff000000u) ? 128u : 0u);
#pragma unroll
        for (int b = 0; b < K::NB; ++b) {
          ih0[b] = ih1[b] = il0[b] = il1[b] = 0u;
          if ((bm >> b) & 1u) {  // warp-uniform
            const int src = (b & 3) * 8 + 2 * t;
            const float rx = b < 4 ? r0x : r1x, ry = b < 4 ? r0y : r1y, rz = b < 4 ? r0z : r1z, sv = b < 4 ? s0 : s1;
            const float ax = __shfl_sync(kFull, rx, src), ay = __shfl_sync(kFull, ry, src),
                        az = __shfl_sync(kFull, rz, src), sa = __shfl_sync(kFull, sv, src);
            const float bx = __shfl_sync(kFull, rx, src + 1), by = __shfl_sync(kFull, ry, src + 1),
                        bz = __shfl_sync(kFull, rz, src + 1), sb = __shfl_sync(kFull, sv, src + 1);
            const float sae = sa * inv_extent, sbe = sb * inv_extent;
            const float f00 = influence_scaled(ax, ay, az, k0x, k0y, k0z, sae, sa);
            const float f01 = influence_scaled(bx, by, bz, k0x, k0y, k0z, sbe, sb);
            const float f10 = influence_scaled(ax, ay, az, k1x, k1y, k1z, sae, sa);
            const float f11 = influence_scaled(bx, by, bz, k1x, k1y, k1z, sbe, sb);
            const __half2 h0 = __floats2half2_rn(f00, f01), h1 = __floats2half2_rn(f10, f11);
            const float2 h0f = __half22float2(h0), h1f = __half22float2(h1);
            ih0[b] = h2_bits(h0);
            ih1[b] = h2_bits(h1);
            il0[b] = h2_bits(__floats2half2_rn(f00 - h0f.x, f01 - h0f.y));
            il1[b] = h2_bits(__floats2half2_rn(f10 - h1f.x, f11 - h1f.y));
          }
        }
      }
      if (nxt.valid) stage_b(nxt, ln);  // the next item's index row has arrived meanwhile
      // ---- the slot: free once the MMAs of its previous fill have completed ----
      if (slot == 0 && half == 0) SPR_TR(0, use >> 1, clock64());
      if (use > 0) mbar_wait_park(&bar_d1full[2 * pair + ((use - 1) & 1)], ((use - 1) >> 1) & 1);
      if (slot == 0 && half == 0) SPR_TR(1, use >> 1, clock64());
      if (!dummy) {
        // A1: per neighbour the 64-byte hi half and the 64-byte lo half of this pass's 32 channels, copied asynchronously
        // through L2 into the two blocks of the tile (rows beyond the last present neighbour are left alone)
        const unsigned char* src_pass = src_lane + pass * 128;
#pragma unroll
        for (int i = 0; i < 4 * KS; ++i) {
          if (4 * i < n_rows) {  // warp-uniform
            const int h = 4 * i + rsub;
            const int j = __shfl_sync(kFull, h < 32 ? j0 : j1, h & 31);
            cp_async16(dst_par[i & 1] + (i >> 1) * 512, src_pass + (size_t)(j >= 0 ? j : 0) * (4 * C), j >= 0 ? 16u : 0u);
          }
        }
        // the influence fragments depend on the geometry only: pass 0 parks them in an L2-resident scratch (one 16-byte
        // vector per lane and block), the later passes of the tile take them back (possibly the slot's other warp: the
        // stores are fenced and the barrier chain producer -> MMA -> producer orders them before the loads)
        uint4* frag = reinterpret_cast<uint4*>(tile_scratch) + (size_t)ql * (K::NB * 32) + lane;
        if (K::PASSES > 1) {
          if (pass == 0) {
#pragma unroll
            for (int b = 0; b < K::NB; ++b) frag[b * 32] = make_uint4(ih0[b], ih1[b], il0[b], il1[b]);
            __threadfence_block();
          } else {
#pragma unroll
            for (int b = 0; b < K::NB; ++b) {
              const uint4 f = frag[b * 32];  // ordinary (L1-coherent within the SM) load
              ih0[b] = f.x;
              ih1[b] = f.y;
              il0[b] = f.z;
              il1[b] = f.w;
            }
          }
        }
        // B1: row = kernel point (+16 for the lo half), K element = neighbour 8 b + 2 t (+1)
#pragma unroll
        for (int b = 0; b < K::NB; ++b) {
          const uint32_t cb = b1_lane + ((uint32_t)(b ^ g) << 4);
          st_shared_u32(cb, ih0[b]);
          st_shared_u32(cb + 1024, ih1[b]);
          st_shared_u32(cb + 2048, il0[b]);
          st_shared_u32(cb + 3072, il1[b]);
        }
      }
      cp_async_arrive_noinc(&bar_full[pair]);
      if (lane == 0) mbar_arrive(&bar_full[pair]);  // release: this warp's shared-memory stores, in program order
      if (slot == 0 && half == 0) SPR_TR(2, use >> 1, clock64());
      cur = nxt;
      lc = ln;
    }
  } else if (is_reader) {
    // =========================================== readback ===========================================
    // group rg (two warps: TMEM lane quadrants 0 and 1) serves slot pair rg.  An M = 64 accumulator keeps row m in
    // quadrant m / 16, lane m % 16; rows 0..31 are the 32 channels; the pair's second slot sits at lane offset 16.
    const int rg = warp >> 2, qd = warp & 3;
    const int second = lane >> 4;               // lanes 0..15: first slot of the pair, 16..31: second slot
    const int cl = 16 * qd + (lane & 15);       // channel (within the pass) of this thread
    uint32_t fill = 0;                          // fills of the pair read so far
    uint32_t seq = 0;
    int rtr = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int q0 = tile * tq;
      const int cnt = min(nq, q0 + tq) - q0;
#pragma unroll 1
      for (int pass = 0; pass < K::PASSES; ++pass, ++seq) {
        bool first = true;
        if (rg < NPAIR) {
#pragma unroll 1
          for (int qa = 2 * rg; qa < cnt; qa += NSLOT) {
            const bool has_b = qa + 1 < cnt;
            mbar_wait_park(&bar_d1full[2 * rg + (fill & 1)], (fill >> 1) & 1);
            ++fill;
            tc_fence_after();
            if (warp == 0) SPR_TR(6, rtr, clock64());
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(32 * qd) << 16) + K::D1_COL0 + 32 * rg, v);
            if (warp == 0) SPR_TR(7, rtr, clock64());
            ++rtr;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_d1free[rg]);
            if (first) {  // the A tile still feeds the phase-2 MMAs of the previous pass until bar_done completes
              if (seq > 0) mbar_wait_park(bar_done, (seq - 1) & 1);
              first = false;
            }
            // this thread's channel, all 16 kernel points: columns k and 16 + k are the I_hi and I_lo products (the
            // X_hi and X_lo halves were already summed by the MMAs); split into fp16 (hi, lo)
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int p2 = 0; p2 < 8; ++p2) {
              const float w0 = v[2 * p2] + v[16 + 2 * p2], w1 = v[2 * p2 + 1] + v[17 + 2 * p2];
              const __half2 hh = __floats2half2_rn(w0, w1);
              const float2 hf = __half22float2(hh);
              hi[p2] = h2_bits(hh);
              lo[p2] = h2_bits(__floats2half2_rn(w0 - hf.x, w1 - hf.y));
            }
            if (!second || has_b) {
              // K index (channel-major) = cl * 16 + k: atom cl / 4, 16-byte chunks (cl % 4) * 2 and + 1
              const int ql = qa + second;
              unsigned char* atom = sA + (cl >> 2) * K::A_ATOM_BYTES;
              const uint32_t j = (cl & 3) * 2;
              *reinterpret_cast<uint4*>(atom + sw128_offset(2 * ql, j)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(atom + sw128_offset(2 * ql, j + 1)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
              *reinterpret_cast<uint4*>(atom + sw128_offset(2 * ql + 1, j)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              *reinterpret_cast<uint4*>(atom + sw128_offset(2 * ql + 1, j + 1)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
          }
        }
        if (first && seq > 0) mbar_wait_park(bar_done, (seq - 1) & 1);  // keep the phases of bar_done in step
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_afull);
      }
    }
  } else if (warp >= K::W_EPI && warp < K::W_MMA) {
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
        for (int pass = 0; pass < K::PASSES; ++pass, ++seq) {
          // D2 += A2 * [W_hi | W_lo]^T for this pass
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
    // issuer g serves the slot pairs p = g, g + 2, in the order the queries were dealt: per pair one wait for the
    // operands, one for the previous readback, 4 KS MMAs (two slots x {hi, lo} feature halves), one commit
    if (lane == 0) {
      const int ig = warp - K::W_ISS;
      constexpr uint32_t idesc1 = idesc_f16_amn(64, 32);
      uint32_t fills[NPAIR];  // per pair: fills issued so far
#pragma unroll
      for (int pr = 0; pr < NPAIR; ++pr) fills[pr] = 0;
      int trq = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int q0 = tile * tq;
        const int cnt = min(nq, q0 + tq) - q0;
        for (int pass = 0; pass < K::PASSES; ++pass) {
          for (int base = 0; base < cnt; base += NSLOT) {
#pragma unroll
            for (int pr = 0; pr < NPAIR; ++pr) {
              if ((pr & 1) != ig || base + 2 * pr >= cnt) continue;
              const uint32_t f = fills[pr];
              mbar_wait_park(&bar_full[pr], f & 1u);
              if (ig == 0) SPR_TR(3, trq, clock64());
              if (f > 0) mbar_wait_park(&bar_d1free[pr], (f - 1) & 1u);  // readback of the previous fill has D1
              if (ig == 0) SPR_TR(4, trq, clock64());
              fence_proxy_async_smem();  // the operands were written through the generic proxy (cp.async, st.shared)
              tc_fence_after();
#pragma unroll
              for (int o = 0; o < 2; ++o) {
                const uint32_t a1 = smem_u32(sSlots + (2 * pr + o) * K::SLOT_BYTES);
#pragma unroll
                for (int hl = 0; hl < 2; ++hl)
#pragma unroll
                  for (int ks = 0; ks < KS; ++ks)
                    umma_f16(tmem + ((uint32_t)(o * 16) << 16) + K::D1_COL0 + 32 * pr,
                             desc_sw64_mnmajor(a1 + hl * K::A1_PART + ks * 1024),
                             desc_sw128_kmajor(a1 + K::A1_BYTES + ks * 32), idesc1, (hl | ks) != 0);
              }
              umma_commit(&bar_d1full[2 * pr + (f & 1u)]);
              if (ig == 0) SPR_TR(5, trq, clock64());
              ++trq;
              fills[pr] = f + 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == W_MMA + 1) {
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
    k_kpconv_g<C, IdxT, KS_><<<grid, K::THREADS, K::SMEM, stream>>>(reinterpret_cast<const unsigned char*>(x16), q, idx_t, row_stride, H, img, kp, pts4,         \
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
