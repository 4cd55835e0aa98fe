// kpconv_g.cu -- fused KPConv forward, second generation: BOTH matrix products of the layer run on tcgen05 and the
// neighbour-feature gather is done by the TMA engine (reference: kpconv_blocks.py:269-414).
//
//   phase 1 (per query, per pass of 32 input channels)
//        wf[c][k] = sum_h X[idx[h]][c] * I[h][k]          I = kernel-point influences (geometry only)
//     A operand = the gathered, pre-split feature rows AS THEY LIE in global memory: a row is 32 channels x (hi, lo) fp16
//       = 64 contiguous "M elements" (even = hi, odd = lo), one row per neighbour = one K index  ->  MN-major SWIZZLE_128B
//       tile, filled by cp.async.bulk.tensor ... tile::gather4 (four rows per instruction, out-of-range = shadow rows
//       arrive as zeros), no thread ever touches a feature;
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
// Warp roles (NSLOT + 14 warps): two readback groups of four warps (slot pair p is served by group p % 2), NSLOT
// producer warps, each owning one operand slot (A1 + B1, 8-12 KB), the four epilogue warps, the MMA-issuing thread and
// the weight-stream thread.  Queries of a tile are dealt statically (query i -> slot i % NSLOT), so every hand-over is
// a plain in-order mbarrier wait.
#include <cuda.h>

#include <mutex>

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
  static constexpr int W_EPI = W_PROD + NSLOT;        // first epilogue warp
  static constexpr int W_MMA = W_EPI + 4;
  static constexpr int WARPS = W_MMA + 2;
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
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* tmap, uint64_t* bar, int col, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
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
template <int C, typename IdxT, int KS>
__global__ void __launch_bounds__(GCfg<C, KS>::THREADS, 1)
    k_kpconv_g(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ q, const IdxT* __restrict__ idx,
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
  uint64_t* bar_full = bars;               // [8]  slot operands in place (producer arrive + TMA bytes)
  uint64_t* bar_d1full = bars + 8;         // [8]  phase-1 MMAs of the slot complete (D1 valid, slot memory reusable)
  uint64_t* bar_d1free = bars + 16;        // [8]  the four readback warps have D1 in registers
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
    for (int i = 0; i < 8; ++i) {
      mbar_init(&bar_full[i], 1);
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
    const int slot = warp - K::W_PROD;
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
    int titer = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
      const int q0 = tile * tq;
      const int cnt = min(nq, q0 + tq) - q0;
      unsigned char* tile_scratch = scratch + ((size_t)blockIdx.x * 2 + (titer & 1)) * K::TQ * K::B1_BYTES;
#pragma unroll 1
      for (int pass = 0; pass < K::PASSES; ++pass) {
#pragma unroll 1
        for (int ql = slot; ql < cnt; ql += NSLOT) {
          const int n = order ? __ldg(order + q0 + ql) : q0 + ql;
          // the neighbour row: lanes = slots (two rounds cover 64 columns)
          int j0 = -1, j1 = -1;
          if (lane < H) j0 = (int)__ldg(idx + (size_t)n * row_stride + lane);
          if (KS > 2 && 32 + lane < H) j1 = (int)__ldg(idx + (size_t)n * row_stride + 32 + lane);
          if (j0 >= ns) j0 = -1;
          if (j1 >= ns) j1 = -1;
          // gather coordinates: lane i < 4 KS fetches rows 4i .. 4i+3 (absent neighbour -> row ns: out of range -> zeros)
          int rr[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int src = (4 * lane + e) & 31;
            const int v0 = __shfl_sync(kFull, j0, src), v1 = __shfl_sync(kFull, j1, src);
            const int v = (4 * lane + e) < 32 ? v0 : v1;
            rr[e] = v >= 0 ? v : ns;
          }
          uint32_t ih0[K::NB], ih1[K::NB], il0[K::NB], il1[K::NB];
          float fcount = 0.f;
          if (pass == 0) {
            // ---- influences of the 16 kernel points on every neighbour, as fp16 (hi, lo) pairs in registers ----
            const float qx = __ldg(q + 3 * (size_t)n), qy = __ldg(q + 3 * (size_t)n + 1), qz = __ldg(q + 3 * (size_t)n + 2);
            const unsigned m0 = __ballot_sync(kFull, j0 >= 0), m1 = __ballot_sync(kFull, j1 >= 0);
            unsigned bm = ((m0 & 0xffu) ? 1u : 0u) | ((m0 & 0xff00u) ? 2u : 0u) | ((m0 & 0xff0000u) ? 4u : 0u) |
                          ((m0 & 0xff000000u) ? 8u : 0u) | ((m1 & 0xffu) ? 16u : 0u) | ((m1 & 0xff00u) ? 32u : 0u) |
                          ((m1 & 0xff0000u) ? 64u : 0u) | ((m1 & 0xff000000u) ? 128u : 0u);
            float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = pa, na = pa, nb = pa;
            auto fetch = [&](int b, float4& fa, float4& fb) {
              const int src = (b & 3) * 8 + 2 * t;
              const int jsel = b < 4 ? j0 : j1;
              const int ja = __shfl_sync(kFull, jsel, src), jb = __shfl_sync(kFull, jsel, src + 1);
              fa = make_float4(0.f, 0.f, 0.f, 0.f);
              fb = fa;
              if (ja >= 0) fa = __ldg(pts4 + ja);
              if (jb >= 0) fb = __ldg(pts4 + jb);
            };
            if (bm & 1u) fetch(0, na, nb);
#pragma unroll
            for (int b = 0; b < K::NB; ++b) {
              ih0[b] = ih1[b] = il0[b] = il1[b] = 0u;
              pa = na;
              pb = nb;
              if (b + 1 < K::NB && ((bm >> (b + 1)) & 1u)) fetch(b + 1, na, nb);  // warp-uniform
              if ((bm >> b) & 1u) {
                const float ax = pa.x - qx, ay = pa.y - qy, az = pa.z - qz;
                const float bx = pb.x - qx, by = pb.y - qy, bz = pb.z - qz;
                const float sa = fabsf(pa.w) * a_scale, sb = fabsf(pb.w) * a_scale;  // an absent neighbour has w = 0
                fcount += (pa.w > 0.f ? 1.f : 0.f) + (pb.w > 0.f ? 1.f : 0.f);
                const float f00 = influence_g(ax, ay, az, k0x, k0y, k0z, inv_extent) * sa;
                const float f01 = influence_g(bx, by, bz, k0x, k0y, k0z, inv_extent) * sb;
                const float f10 = influence_g(ax, ay, az, k1x, k1y, k1z, inv_extent) * (sa * k1_on);
                const float f11 = influence_g(bx, by, bz, k1x, k1y, k1z, inv_extent) * (sb * k1_on);
                const __half2 h0 = __floats2half2_rn(f00, f01), h1 = __floats2half2_rn(f10, f11);
                const float2 h0f = __half22float2(h0), h1f = __half22float2(h1);
                ih0[b] = h2_bits(h0);
                ih1[b] = h2_bits(h1);
                il0[b] = h2_bits(__floats2half2_rn(f00 - h0f.x, f01 - h0f.y));
                il1[b] = h2_bits(__floats2half2_rn(f10 - h1f.x, f11 - h1f.y));
              }
            }
          }
          // ---- the slot: free once the MMAs of its previous fill have completed ----
          if (use > 0) mbar_wait(&bar_d1full[slot], (use - 1) & 1);
          if (lane == 0) {
            if (K::PASSES > 1) {
              if (pass == 0) bulk_wait_read0();  // the scratch copy of the previous fill has left shared memory
              else bulk_wait0();                 // this thread's scratch copies have landed in global memory
            }
            mbar_expect_tx(&bar_full[slot], K::A1_BYTES + (pass > 0 ? K::B1_BYTES : 0));
          }
          __syncwarp();
          if (lane < 4 * KS) tma_gather4(sA1 + lane * 512, &tmap, &bar_full[slot], pass * 64, rr[0], rr[1], rr[2], rr[3]);
          if (pass == 0) {
            // B1: row = kernel point (+16 for the lo half), K element = neighbour 8 b + 2 t (+1)
#pragma unroll
            for (int b = 0; b < K::NB; ++b) {
              *reinterpret_cast<uint32_t*>(sB1 + sw128_offset(g, b) + 4 * t) = ih0[b];
              *reinterpret_cast<uint32_t*>(sB1 + sw128_offset(8 + g, b) + 4 * t) = ih1[b];
              *reinterpret_cast<uint32_t*>(sB1 + sw128_offset(16 + g, b) + 4 * t) = il0[b];
              *reinterpret_cast<uint32_t*>(sB1 + sw128_offset(24 + g, b) + 4 * t) = il1[b];
            }
            // a neighbour is replicated over g: count the g == 0 copies (lanes 0..3)
            float c = g == 0 ? fcount : 0.f;
            c += __shfl_xor_sync(kFull, c, 1);
            c += __shfl_xor_sync(kFull, c, 2);
            if (lane == 0) sInv[(titer % 3) * K::TQ + ql] = 1.f / fmaxf(c, 1.f);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              if (K::PASSES > 1) bulk_s2g(tile_scratch + (size_t)ql * K::B1_BYTES, sB1, K::B1_BYTES);
              mbar_arrive(&bar_full[slot]);
            }
          } else {
            if (lane == 0) {
              bulk_g2s(sB1, tile_scratch + (size_t)ql * K::B1_BYTES, K::B1_BYTES, &bar_full[slot]);
              mbar_arrive(&bar_full[slot]);
            }
          }
          ++use;
        }
      }
    }
    if (K::PASSES > 1 && lane == 0) bulk_wait0();
  } else if (warp < K::W_PROD) {
    // =========================================== readback ===========================================
    const int rg = warp >> 2, qd = warp & 3;   // group, TMEM lane quadrant
    const int second = lane >> 4;               // lanes 0..15: first slot of the pair, 16..31: second slot
    const int cl = 8 * qd + ((lane & 15) >> 1);  // channel (within the pass) of this lane pair: D1 row 2 cl (+1)
    const int odd = lane & 1;                   // even lane: X_hi partials and kernel points 0..7, odd: X_lo and 8..15
    uint32_t par = 0;                           // bit s: parity of the next completion of slot s
    uint32_t seq = 0;
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
            mbar_wait(&bar_d1full[2 * pr], (par >> (2 * pr)) & 1u);
            par ^= 1u << (2 * pr);
            if (has_b) {
              mbar_wait(&bar_d1full[2 * pr + 1], (par >> (2 * pr + 1)) & 1u);
              par ^= 1u << (2 * pr + 1);
            }
            tc_fence_after();
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(32 * qd) << 16) + K::D1_COL0 + 32 * pr, v);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&bar_d1free[2 * pr]);
              if (has_b) mbar_arrive(&bar_d1free[2 * pr + 1]);
            }
            if (first) {  // the A tile still feeds the phase-2 MMAs of the previous pass until bar_done completes
              if (seq > 0) mbar_wait(bar_done, (seq - 1) & 1);
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
        if (first && seq > 0) mbar_wait(bar_done, (seq - 1) & 1);  // keep the phases of bar_done in step
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
      mbar_wait_sleep(bar_d2full, titer & 1);
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
    // =========================================== MMA issuer ===========================================
    if (lane == 0) {
      constexpr uint32_t idesc1 = idesc_f16_amn(64, 32);
      constexpr uint32_t idesc2 = idesc_f16_f32(128, K::NS);
      const uint64_t adesc0 = desc_sw128_kmajor(smem_u32(sA));
      const uint64_t bdesc0 = desc_sw128_kmajor(smem_u32(sRing));
      uint32_t seq = 0;
      uint32_t par = 0, used = 0;  // per slot: parity of the next fill, slot filled before
      int stage = 0;
      uint32_t phase = 0;
      int titer = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
        const int q0 = tile * tq;
        const int cnt = min(nq, q0 + tq) - q0;
        for (int pass = 0; pass < K::PASSES; ++pass, ++seq) {
          // ---- phase 1: one group of KS MMAs per query, in query order (query i lives in slot i % NSLOT) ----
          for (int i = 0; i < cnt; ++i) {
            const int slot = i % NSLOT;
            const uint32_t p = (par >> slot) & 1u;
            mbar_wait(&bar_full[slot], p);
            if ((used >> slot) & 1u) mbar_wait(&bar_d1free[slot], p ^ 1u);  // readback of the previous fill has D1
            tc_fence_after();
            const uint32_t a1 = smem_u32(sSlots + slot * K::SLOT_BYTES);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              umma_f16(tmem + ((uint32_t)((slot & 1) * 16) << 16) + K::D1_COL0 + 32 * (slot >> 1),
                       desc_sw128_mnmajor(a1 + ks * 2048),
                       desc_sw128_kmajor(a1 + K::A1_BYTES + ks * 32), idesc1, ks != 0);
            umma_commit(&bar_d1full[slot]);
            par ^= 1u << slot;
            used |= 1u << slot;
          }
          // ---- phase 2: D2 += A2 * [W_hi | W_lo]^T for this pass ----
          mbar_wait(bar_afull, seq & 1);
          if (pass == 0 && titer > 0) mbar_wait(bar_d2free, (titer - 1) & 1);
          tc_fence_after();
          for (int a = 0; a < 8; ++a) {
            for (int sub = 0; sub < K::NSUB; ++sub) {
              mbar_wait(&bar_wfull[stage], phase);
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
  } else {
    // =========================================== weight stream ===========================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int blk = 0; blk < K::PASSES * K::BLOCKS_PER_PASS; ++blk) {
          mbar_wait_sleep(&bar_wempty[stage], phase ^ 1);
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
  static std::once_flag once;
  static EncodeTiledFn fn = nullptr;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

template <int C, typename IdxT>
int launch_g(const float* q, const void* idx, int row_stride, int H, const uint32_t* x16, const unsigned char* img,
             const float* kp, const float4* pts4, const unsigned int* amax_x_bits, const unsigned int* amax_w_bits,
             float extent, float* out, int nq, int ns, void* scratch, const int* order, cudaStream_t stream) {
  SPR_CHECK_ARG(H <= 64, "kpconv_forward_gather: at most 64 neighbour columns are supported (got %d)", H);
  SPR_CHECK_ARG(C <= 32 || scratch, "kpconv_forward_gather: scratch buffer missing");
  EncodeTiledFn enc = encode_tiled();
  SPR_CHECK_ARG(enc, "kpconv_forward_gather: cuTensorMapEncodeTiled is not available from this driver");
  // the pre-split feature rows as a [ns, 2C] fp16 matrix; a box is one row of 64 elements (32 channels x hi/lo)
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {(cuuint64_t)(2 * C), (cuuint64_t)ns};
  const cuuint64_t strides[1] = {(cuuint64_t)(4 * C)};
  const cuuint32_t box[2] = {64, 1};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<uint32_t*>(x16), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SPR_CHECK_ARG(r == CUDA_SUCCESS, "kpconv_forward_gather: cuTensorMapEncodeTiled failed (%d)", (int)r);
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
    k_kpconv_g<C, IdxT, KS_><<<grid, K::THREADS, K::SMEM, stream>>>(tmap, q, idx_t, row_stride, H, img, kp, pts4,         \
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
