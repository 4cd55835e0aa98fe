// kpconv_s.cu -- fused KPConv forward, third generation ("staged"): the structure of kpconv_tc.cu (warp-level mma.sync
// phase 1 per query, tcgen05 phase 2 per tile of 64 queries, TMA weight ring; reference: kpconv_blocks.py:269-414) with
// a different way of getting the neighbours' rows to the producer warps.
//
// What the profile of kpconv_tc.cu showed (profiles/README.md, round 2): 27 % of its warp time is spent waiting for
// global loads that were requested only one 8-neighbour block ahead -- deeper prefetching costs registers there, and
// the 640 threads x 96 registers leave none -- and each block pays 8 PRMT + 2 LDG.128 + 64-bit address arithmetic to
// assemble the B fragments of the warp-level MMAs.  Here
//   * every producer warp owns a ring of three block slots in shared memory (8 feature rows of 128 B + 8 packed
//     points + a 16-byte header; two slots from C = 64 on, where the third weight-ring stage needs the room); the rows
//     of the next block(s) are requested with 16-byte asynchronous copies (cp.async through L2, zero-fill for absent
//     neighbours) while block i is multiplied: no register holds data in flight;
//   * the feature rows are in the PLANAR pre-split format (per 32-channel group 32 fp16 hi halves, then 32 lo halves),
//     so the B fragments of the four channel tiles are ONE ldmatrix.x4.trans for the hi parts and one for the lo parts
//     (rows XOR-swizzled in the slot: conflict-free);
//   * the header carries the query point and (query slot, block, first / last of its query), so the multiply side keeps
//     no per-item state either and the blocks of consecutive queries of a pass stream through the ring back to back.
//     (Letting the stream run on across passes and tiles -- dispensers of later passes drawn early, pass structure
//     carried by header flags -- was built and measured 12-17 % SLOWER, as was the influence-fragment cache of
//     kpconv_tc.cu: both are gone.)
// A lane's accumulators of channel tile i are channels 8 i + 2 t, 8 i + 2 t + 1; the K index of phase 2 orders the 32
// channels of a pass as position 8 t + 2 i + e <-> channel 8 i + 2 t + e, so that a lane still writes 16 contiguous bytes
// of the A tile; the weight image is permuted to match (k_weight_image_s).
// Everything else -- operand scaling, A tile, MMA issuer, weight stream, deferred epilogue, barriers -- is kpconv_tc.cu's.
#include "spr_common.cuh"
#include "tc05.cuh"

namespace spr {
namespace {

using namespace tc;

constexpr int KP = 15;

template <int C>
struct SCfg {
  static constexpr int PASSES = C / 32;
  static constexpr int NCOL = 2 * C;
  static constexpr int NS = NCOL < 128 ? NCOL : 128;  // N of one MMA = rows of one ring stage
  static constexpr int NSUB = NCOL / NS;
  static constexpr int STAGE_BYTES = NS * 128;
  // The shared memory left by the A tile is split between the weight ring and the producers' block slots.  Measured
  // (profiles/r2c_kpconv_gen_bench_32pairs.log and the runs before it): a third ring stage is worth 6 % at C = 128 and
  // 16 % at C = 256 -- the MMAs of a pass wait for 256 KB and more of weights through the ring while the producers wait
  // for them (bar_done) -- whereas a third block slot per warp is worth 1 % (C = 32) or nothing (C = 64).
  static constexpr int NSTAGES = 3;
  static constexpr int BLOCKS_PER_PASS = 8 * NSUB;  // 8 K atoms of 64 fp16 per pass
  static constexpr int TQ = 64;
  static constexpr int WORKERS = 18;  // (a 21st warp makes six on one scheduler: 16384 / 6 caps the registers at 80 and spills; measured 5 % slower)
  static constexpr int THREADS = (WORKERS + 2) * 32;
  static constexpr int A_ATOM_BYTES = 128 * 128;
  static constexpr int A_BYTES = 8 * A_ATOM_BYTES;
  static constexpr int DEPTH = C <= 32 ? 3 : 2;       // block slots per producer warp
  static constexpr int SLOT_X = 1024, SLOT_P = 128;   // 8 rows x 128 B, 8 packed points
  static constexpr int SLOT_BYTES = SLOT_X + SLOT_P + 16;
  static constexpr int STAGING_BYTES = WORKERS * DEPTH * SLOT_BYTES;
  static constexpr int TMEM_COLS = NCOL < 32 ? 32 : NCOL;
  static constexpr int OFF_RING = A_BYTES;
  static constexpr int OFF_STAGING = OFF_RING + NSTAGES * STAGE_BYTES;
  static constexpr int OFF_MISC = OFF_STAGING + STAGING_BYTES;
  static constexpr int MISC_BYTES = 16 * 8 + 32 + 2 * TQ * 4 + 48 * 4;
  static constexpr size_t SMEM = 1024 + OFF_MISC + MISC_BYTES;
  static constexpr size_t IMG_BYTES = (size_t)PASSES * BLOCKS_PER_PASS * STAGE_BYTES;
  static_assert(SMEM <= 232448, "shared memory budget");
};

__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }
// d[16x8] += a[16x8] * b[8x8], fp16 operands, fp32 accumulate (warp-level tensor path)
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
// 16-byte asynchronous copy global -> shared through L2 (LDGSTS); src_bytes = 0 writes zeros without touching memory.
// (Absent neighbours must NOT be redirected to a shared all-zero row instead: every SM then reads the same line for a
// quarter of its copies -- measured 2x slower for the whole kernel.)
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
// keeps a per-lane constant in its register: without it the compiler re-derives these offsets from the lane number in
// every iteration (~100 of the ~1150 warp instructions per query in the first profile of this kernel)
__device__ __forceinline__ uint32_t pinned(uint32_t v) {
  asm volatile("" : "+r"(v));
  return v;
}
__device__ __forceinline__ void sts128u(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// four 8 x 8 fp16 matrices, transposed: lane (g, t) receives {M[2t][g], M[2t+1][g]} of matrix i in r[i] -- the B fragment
// of mma.m16n8k8 when the rows of M are K (neighbours) and its columns N (channels)
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---------------------------------------------------------------------------------------------
// weight image, as kpconv_tc.cu's k_weight_image with the channels of a pass permuted: K element kk of the atom <->
// kernel point 2*atom + kk/32 (15 = zero padding), position p = kk % 32 <-> input channel pass*32 + 8*((p%8)/2) + 2*(p/8)
// + p%2.  One thread per 16-byte chunk.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) k_weight_image_s(const float* __restrict__ w,
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
  const int k = atom * 2 + (j >> 2);  // kernel point of this chunk (15 = zero padding)
  const float tscale = pow2i(scale_exp(__uint_as_float(*amax_w_bits), 14));
  __align__(16) __half h[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int cin = pass * 32 + 8 * (e >> 1) + 2 * (j & 3) + (e & 1);
    float v = 0.f;
    if (k < KP) v = __ldg(w + ((size_t)k * C + cin) * C + o) * tscale;
    const __half hi = __float2half_rn(v);
    h[e] = lo_part ? __float2half_rn(v - __half2float(hi)) : hi;
  }
  *reinterpret_cast<uint4*>(img + (size_t)blk * (NS * 128) + sw128_offset(r, j)) = *reinterpret_cast<const uint4*>(h);
}

__global__ void __launch_bounds__(256) k_absmax_s(const float* __restrict__ w, int n, unsigned int* __restrict__ amax_bits) {
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
  m = warp_maxf(m);
  if ((threadIdx.x & 31) == 0 && __float_as_uint(m) > *reinterpret_cast<volatile unsigned int*>(amax_bits))
    atomicMax(amax_bits, __float_as_uint(m));
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
template <int C, typename IdxT, int HR>  // HR = ceil(H / 32): 32-slot rounds of a neighbour row
__global__ void __launch_bounds__(SCfg<C>::THREADS, 1)
    k_kpconv_s(const float* __restrict__ q, const IdxT* __restrict__ idx, int row_stride, int H,
               const unsigned char* __restrict__ x16p, const unsigned char* __restrict__ wimg,
               const float* __restrict__ kp, const float4* __restrict__ pts4, const unsigned int* __restrict__ amax_x_bits,
               const unsigned int* __restrict__ amax_w_bits, float extent, float* __restrict__ out, int nq, int ns, int tq,
               int n_tiles, const int* __restrict__ order) {
  using K = SCfg<C>;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;
  unsigned char* sRing = smem + K::OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_MISC);
  uint64_t* bar_full = bars;                 // [NSTAGES]
  uint64_t* bar_empty = bars + 4;            // [NSTAGES]
  uint64_t* bar_afull = bars + 8;
  uint64_t* bar_done = bars + 9;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);
  int* s_ctr = reinterpret_cast<int*>(s_tmem + 4);      // [4] query dispensers, indexed by pass sequence & 3
  float* sInv = reinterpret_cast<float*>(s_tmem + 8);  // [2][TQ]
  float* sKp = sInv + 2 * K::TQ;                       // [45] (48 reserved)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int i = 0; i < K::NSTAGES; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    mbar_init(bar_afull, K::WORKERS);
    mbar_init(bar_done, 1);
    fence_mbar_init();
    for (int i = 0; i < 4; ++i) s_ctr[i] = 0;
  }
  if (warp == K::WORKERS) tmem_alloc(s_tmem, K::TMEM_COLS);
  for (int i = tid; i < KP * 3; i += K::THREADS) sKp[i] = kp[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  const int es = scale_exp((float)H * __uint_as_float(*amax_x_bits), 15);
  const int et = scale_exp(__uint_as_float(*amax_w_bits), 14);

  if (warp < K::WORKERS) {
    // =========================================== producers ===========================================
    const float inv_extent = 1.0f / extent;
    const float a_scale = pow2i(es);
    const float o_scale = pow2i(-(es + et));
    const int g = lane >> 2, t = lane & 3;
    const float k0x = sKp[3 * g], k0y = sKp[3 * g + 1], k0z = sKp[3 * g + 2];
    // row 15 of the A fragment is padding: a kernel point infinitely far away has influence exactly 0
    const float k1x = g < 7 ? sKp[3 * (g + 8)] : 1.0e18f, k1y = g < 7 ? sKp[3 * (g + 8) + 1] : 0.f,
                k1z = g < 7 ? sKp[3 * (g + 8) + 2] : 0.f;
    // negated and duplicated for the packed (two neighbours at a time) influence arithmetic
    const f2_t nk0x = f2_pack(-k0x, -k0x), nk0y = f2_pack(-k0y, -k0y), nk0z = f2_pack(-k0z, -k0z);
    const f2_t nk1x = f2_pack(-k1x, -k1x), nk1y = f2_pack(-k1y, -k1y), nk1z = f2_pack(-k1z, -k1z);
    // ---- the warp's ring of block slots ----
    const uint32_t ring = pinned(smem_u32(smem + K::OFF_STAGING + warp * (K::DEPTH * K::SLOT_BYTES)));
    const uint32_t sA32 = smem_u32(sA);
    // copy role: rows crow and crow + 4 of a block, 16-byte chunk cch of the 128-byte row (chunks 0..3 = hi halves of
    // channels 0-7 .. 24-31, chunks 4..7 = lo halves); chunk c of row r is stored at chunk position c ^ r
    const int crow = lane >> 3, cch = lane & 7;
    // (row crow + 4: 512 bytes further, chunk position with bit 2 flipped)
    const uint32_t xdst0 = pinned(crow * 128 + ((cch ^ crow) << 4));
    const uint32_t pdst0 = pinned(K::SLOT_X + crow * 16);
    // read role: ldmatrix row address of matrix lane / 8 (= channel tile), row lane % 8 (= neighbour of the block)
    const int lrow = lane & 7, lm = lane >> 3;
    const uint32_t hi_off = pinned(lrow * 128 + ((lm ^ lrow) << 4));  // lo parts: chunk 4 + lm, i.e. bit 6 flipped
    const uint32_t pa_off = pinned(K::SLOT_X + 32 * t);
    constexpr uint32_t hdr_off = K::SLOT_X + K::SLOT_P;

    uint32_t seq = 0;
    int titer = 0;
    // Epilogue of a finished tile: D (TMEM) -> registers, add the hi/lo rows (adjacent lanes) and the two column
    // halves, scale, store.  Run DEFERRED: a warp executes it for tile t-1 just before its first A-tile store of
    // tile t, when the MMAs of tile t-1 have long completed, so no producer ever idles on the tensor pipe.
    auto epilogue = [&](int q0, int cnt, const float* inv_buf) {
      tc_fence_after();
      if (warp < 16) {  // 4 TMEM lane quadrants x 4 column groups
        const int qd = warp & 3, cg = warp >> 2;
        const int ql = qd * 16 + (lane >> 1);
        const bool ok = ql < cnt;
        const int n = ok ? (order ? __ldg(order + q0 + ql) : q0 + ql) : 0;
        const float scale = ok ? inv_buf[ql] * o_scale : 0.f;
        const uint32_t trow = tmem + ((uint32_t)(qd * 32) << 16);
#pragma unroll 1
        for (int c0 = cg * (C / 4); c0 < (cg + 1) * (C / 4); c0 += 8) {
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
      }
      tc_fence_before();
    };
    int prev_q0 = 0, prev_cnt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++titer) {
      const int q0 = tile * tq;
      const int cnt = min(nq, q0 + tq) - q0;  // queries in this tile
      float* inv_buf = sInv + (titer & 1) * K::TQ;
#pragma unroll 1
      for (int pass = 0; pass < K::PASSES; ++pass, ++seq) {
        const unsigned char* xpass = x16p + pass * 128 + cch * 16;
        bool first = true;
        // queries are dealt to the warps dynamically (neighbourhood sizes vary); the dispenser of pass seq+2 is
        // reset by whoever draws query 0 of pass seq (no warp can still be in pass seq-2, see bar_done)
        int* ctr = s_ctr + (seq & 3);
        auto grab = [&]() {
          int v = 0;
          if (lane == 0) v = atomicAdd(ctr, 1);
          return __shfl_sync(kFull, v, 0);
        };
        // ---- issue side: the query whose blocks are being requested, and the row of the one after it ----
        int jr[HR], jrn[HR];
        float qx = 0.f, qy = 0.f, qz = 0.f, qnx = 0.f, qny = 0.f, qnz = 0.f;
        unsigned bm = 0;
        bool firstblk = true, row_pending = false;
        int ql_iss = grab(), ql_next = 0;
        if (ql_iss == 0 && lane == 0) s_ctr[(seq + 2) & 3] = 0;
        bool more = ql_iss < cnt;
        auto issue_row = [&](int qq) {
          // `order` (optional) walks the queries in cell order: the queries in flight on an SM are spatial
          // neighbours, so their neighbourhoods overlap and the gathers hit L2
          const int n = order ? __ldg(order + q0 + qq) : q0 + qq;
#pragma unroll
          for (int i = 0; i < HR; ++i) {
            const int h = 32 * i + lane;
            jrn[i] = -1;
            if (h < H) jrn[i] = (int)__ldg(idx + (size_t)n * row_stride + h);
          }
          qnx = __ldg(q + 3 * (size_t)n);
          qny = __ldg(q + 3 * (size_t)n + 1);
          qnz = __ldg(q + 3 * (size_t)n + 2);
        };
        auto start_query = [&]() {  // the pending row becomes the issue query; request the row after it
          bm = 0;
#pragma unroll
          for (int i = 0; i < HR; ++i) {
            jr[i] = -1;
            if (32 * i < H) {
              const bool valid = jrn[i] >= 0 && jrn[i] < ns;
              jr[i] = valid ? jrn[i] : -1;
              const unsigned m = __ballot_sync(kFull, valid);
              const unsigned b4 = ((m & 0xffu) ? 1u : 0u) | ((m & 0xff00u) ? 2u : 0u) | ((m & 0xff0000u) ? 4u : 0u) |
                                  ((m & 0xff000000u) ? 8u : 0u);
              bm |= b4 << (4 * i);
            }
          }
          if (bm == 0) bm = 1;  // a query without neighbours still runs one (all-padding) block
          qx = qnx;
          qy = qny;
          qz = qnz;
          firstblk = true;
          ql_next = grab();
          row_pending = ql_next < cnt;
          if (row_pending) issue_row(ql_next);
        };
        int n_iss = 0, n_done = 0;
        // request one block (or nothing, when the warp has no more queries in this pass) into ring slot s
        auto issue_one = [&](int s) {
          if (more) {
            const int b = __ffs(bm) - 1;
            bm &= bm - 1;
            int jsel = jr[0];
#pragma unroll
            for (int i = 1; i < HR; ++i)
              if ((b >> 2) == i) jsel = jr[i];
            const int src = (b & 3) * 8 + crow;
            const int j0 = __shfl_sync(kFull, jsel, src);
            const int j1 = __shfl_sync(kFull, jsel, src + 4);
            const uint32_t slot = ring + s * K::SLOT_BYTES;
            const uint32_t n0 = j0 >= 0 ? 16u : 0u, n1 = j1 >= 0 ? 16u : 0u;  // absent neighbours: zero fill
            const unsigned r0 = max(j0, 0), r1 = max(j1, 0);
            cp_async16(slot + xdst0, xpass + (size_t)r0 * (4 * C), n0);
            cp_async16(slot + ((xdst0 ^ 64u) + 512u), xpass + (size_t)r1 * (4 * C), n1);
            if (cch == 0) {
              cp_async16(slot + pdst0, pts4 + r0, n0);
              cp_async16(slot + pdst0 + 64, pts4 + r1, n1);
            }
            if (lane == 0) {
              const int meta = ql_iss | (b << 8) | (bm == 0 ? 1 << 16 : 0) | (firstblk ? 1 << 17 : 0);
              sts128(slot + hdr_off, make_float4(qx, qy, qz, __int_as_float(meta)));
            }
            firstblk = false;
            ++n_iss;
            if (bm == 0) {  // that was the query's last block
              if (row_pending) {
                ql_iss = ql_next;
                start_query();
              } else {
                more = false;
              }
            }
          }
          cp_async_commit();  // (possibly empty: the group count per iteration stays uniform)
        };
        if (more) {
          issue_row(ql_iss);
          start_query();
        }
        issue_one(0);
        if (K::DEPTH > 2) issue_one(1);
        float d[4][4];
        float fcount = 0.f;
        int cs = 0, is = K::DEPTH - 1;  // ring slots of the block being multiplied / requested
#pragma unroll 1
        while (n_done < n_iss) {
          cp_async_wait<K::DEPTH - 2>();  // this lane's copies of the oldest block have landed ...
          __syncwarp();        // ... and so have everybody's; the slot freed by the previous iteration is reusable
          issue_one(is);
          is = is == K::DEPTH - 1 ? 0 : is + 1;
          const uint32_t slot = ring + cs * K::SLOT_BYTES;
          cs = cs == K::DEPTH - 1 ? 0 : cs + 1;
          ++n_done;
          const float4 hdr = lds128(slot + hdr_off);
          const int meta = __float_as_int(hdr.w);
          if (meta & (1 << 17)) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int e = 0; e < 4; ++e) d[i][e] = 0.f;
            fcount = 0.f;
          }
          const float4 pa = lds128(slot + pa_off), pb = lds128(slot + pa_off + 16);
          uint32_t bh[4], bl[4];
          ldmatrix_x4_trans(slot + hi_off, bh);
          ldmatrix_x4_trans(slot + (hi_off ^ 64u), bl);
          {
            const float ax = pa.x - hdr.x, ay = pa.y - hdr.y, az = pa.z - hdr.z;
            const float bx = pb.x - hdr.x, by = pb.y - hdr.y, bz = pb.z - hdr.z;
            const float sa = fabsf(pa.w) * a_scale, sb = fabsf(pb.w) * a_scale;  // an absent neighbour has w = 0
            const float sae = sa * inv_extent, sbe = sb * inv_extent;
            if (pass == 0) fcount += (pa.w > 0.f ? 1.f : 0.f) + (pb.w > 0.f ? 1.f : 0.f);
            // A fragment: a0 = (k = g; h = 2t, 2t+1), a1 = (k = g+8; h = 2t, 2t+1), fp16 hi + lo.  The influences are
            // re-evaluated in every channel pass: caching the fragments of pass 0 (as kpconv_tc.cu does) was measured
            // slower here -- a fragment load one block ahead stalls longer than the 4 influences take
            // two neighbours per instruction (packed fp32): s * max(0, 1 - d / extent) = max(0, s - d * (s / extent))
            const f2_t cx = f2_pack(ax, bx), cy = f2_pack(ay, by), cz = f2_pack(az, bz);
            const f2_t s2 = f2_pack(sa, sb), nse2 = f2_pack(-sae, -sbe);
            float f00, f01, f10, f11;
            {
              const f2_t dx = f2_add(cx, nk0x), dy = f2_add(cy, nk0y), dz = f2_add(cz, nk0z);
              float d2a, d2b, da, db;
              f2_unpack(f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx))), d2a, d2b);
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(da) : "f"(d2a));  // MUFU.SQRT, rel. error ~2^-23, sqrt(0) = 0
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(db) : "f"(d2b));
              f2_unpack(f2_fma(f2_pack(da, db), nse2, s2), f00, f01);
              f00 = fmaxf(f00, 0.f);
              f01 = fmaxf(f01, 0.f);
            }
            {
              const f2_t dx = f2_add(cx, nk1x), dy = f2_add(cy, nk1y), dz = f2_add(cz, nk1z);
              float d2a, d2b, da, db;
              f2_unpack(f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx))), d2a, d2b);
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(da) : "f"(d2a));
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(db) : "f"(d2b));
              f2_unpack(f2_fma(f2_pack(da, db), nse2, s2), f10, f11);
              f10 = fmaxf(f10, 0.f);
              f11 = fmaxf(f11, 0.f);
            }
            const __half2 h0 = __floats2half2_rn(f00, f01), h1 = __floats2half2_rn(f10, f11);
            const float2 h0f = __half22float2(h0), h1f = __half22float2(h1);
            const uint32_t ah0 = h2_bits(h0), ah1 = h2_bits(h1);
            const uint32_t al0 = h2_bits(__floats2half2_rn(f00 - h0f.x, f01 - h0f.y));
            const uint32_t al1 = h2_bits(__floats2half2_rn(f10 - h1f.x, f11 - h1f.y));
            // product-major order: consecutive MMAs write different accumulators
#pragma unroll
            for (int i = 0; i < 4; ++i) mma_f16(d[i], al0, al1, bh[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) mma_f16(d[i], ah0, ah1, bl[i]);
#pragma unroll
            for (int i = 0; i < 4; ++i) mma_f16(d[i], ah0, ah1, bh[i]);
          }
          if (!(meta & (1 << 16))) continue;
          // ---- the query is complete: split to fp16 pairs and store its two A rows ----
          // the A tile still feeds the MMAs of the previous pass until bar_done completes
          if (first) {
            if (seq > 0) mbar_wait_park(bar_done, (seq - 1) & 1);
            if (pass == 0 && titer > 0) epilogue(prev_q0, prev_cnt, sInv + ((titer - 1) & 1) * K::TQ);
            first = false;
          }
          // this lane holds, for kernel points k = g (d[i][0], d[i][1]) and k = g + 8 (d[i][2], d[i][3]), channels
          // 8 i + 2 t, 8 i + 2 t + 1 = K positions 8 t + 2 i, 8 t + 2 i + 1: one 16-byte chunk of fp16 each,
          // K element = k * 32 + position  ->  atom k / 2, chunk (k % 2) * 4 + t
          const int ql = meta & 0xff;
          const uint32_t r0 = 2 * ql, r1 = 2 * ql + 1;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float v0 = d[i][2 * half], v1 = d[i][2 * half + 1];
              const __half2 hh = __floats2half2_rn(v0, v1);
              const float2 hf = __half22float2(hh);
              hi[i] = h2_bits(hh);
              lo[i] = h2_bits(__floats2half2_rn(v0 - hf.x, v1 - hf.y));
            }
            const int k = g + 8 * half;
            const uint32_t atom = sA32 + (k >> 1) * K::A_ATOM_BYTES;
            const uint32_t j = (k & 1) * 4 + t;
            sts128u(atom + sw128_offset(r0, j), hi[0], hi[1], hi[2], hi[3]);
            sts128u(atom + sw128_offset(r1, j), lo[0], lo[1], lo[2], lo[3]);
          }
          if (pass == 0) {
            // a neighbour is replicated over g: count the g == 0 copies (lanes 0..3)
            float c = g == 0 ? fcount : 0.f;
            c += __shfl_xor_sync(kFull, c, 1);
            c += __shfl_xor_sync(kFull, c, 2);
            if (lane == 0) inv_buf[ql] = 1.f / fmaxf(c, 1.f);
          }
        }
        cp_async_wait<0>();  // (only empty groups are left)
        if (first) {         // warp without a query in this pass
          if (seq > 0) mbar_wait_park(bar_done, (seq - 1) & 1);
          if (pass == 0 && titer > 0) epilogue(prev_q0, prev_cnt, sInv + ((titer - 1) & 1) * K::TQ);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_afull);
      }
      prev_q0 = q0;
      prev_cnt = cnt;
    }
    if (titer > 0) {  // the last tile's epilogue
      mbar_wait_park(bar_done, (seq - 1) & 1);
      epilogue(prev_q0, prev_cnt, sInv + ((titer - 1) & 1) * K::TQ);
    }
  } else if (warp == K::WORKERS) {
    // =========================================== MMA issuer ===========================================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_f16_f32(128, K::NS);
      const uint64_t adesc0 = desc_sw128_kmajor(smem_u32(sA));
      const uint64_t bdesc0 = desc_sw128_kmajor(smem_u32(sRing));
      uint32_t seq = 0;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int pass = 0; pass < K::PASSES; ++pass, ++seq) {
          mbar_wait_park(bar_afull, seq & 1);
          tc_fence_after();
          for (int a = 0; a < 8; ++a) {
            for (int sub = 0; sub < K::NSUB; ++sub) {
              mbar_wait_park(&bar_full[stage], phase);
              tc_fence_after();
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ad = adesc0 + (uint64_t)((a * K::A_ATOM_BYTES + kk * 32) >> 4);
                const uint64_t bd = bdesc0 + (uint64_t)((stage * K::STAGE_BYTES + kk * 32) >> 4);
                umma_f16(tmem + sub * K::NS, ad, bd, idesc, (pass | a | kk) != 0);
              }
              umma_commit(&bar_empty[stage]);
              if (++stage == K::NSTAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
          umma_commit(bar_done);
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
          mbar_wait_park(&bar_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&bar_full[stage], K::STAGE_BYTES);
          bulk_g2s(sRing + stage * K::STAGE_BYTES, wimg + (size_t)blk * K::STAGE_BYTES, K::STAGE_BYTES,
                   &bar_full[stage]);
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
  if (warp == K::WORKERS) tmem_dealloc(tmem, K::TMEM_COLS);
}

// ---- host ----------------------------------------------------------------------------------------------------------
template <int C, typename IdxT>
int launch_s(const float* q, const void* idx, int row_stride, int H, const unsigned char* x16p, const unsigned char* img,
             const float* kp, const float4* pts4, const unsigned int* amax_x_bits, const unsigned int* amax_w_bits,
             float extent, float* out, int nq, int ns, const int* order, cudaStream_t stream) {
  using K = SCfg<C>;
  SPR_CHECK_ARG(H <= 96, "kpconv_forward_staged: at most 96 neighbour columns are supported (got %d)", H);
  // Tile size: the largest tq <= 64 that deals every SM the same number of tiles (as kpconv_tc.cu)
  int tq = K::TQ;
  {
    const int waves = (nq + kNumSMs * K::TQ - 1) / (kNumSMs * K::TQ);
    const int per = (nq + kNumSMs * waves - 1) / (kNumSMs * waves);
    tq = per < 16 ? 16 : (per > K::TQ ? K::TQ : per);
  }
  const int n_tiles = (nq + tq - 1) / tq;
  const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  const IdxT* idx_t = static_cast<const IdxT*>(idx);
#define SPR_S(HR_)                                                                                                       \
  do {                                                                                                                   \
    SPR_CUDA(ensure_max_dynamic_smem(reinterpret_cast<const void*>(k_kpconv_s<C, IdxT, HR_>), K::SMEM));                 \
    k_kpconv_s<C, IdxT, HR_><<<grid, K::THREADS, K::SMEM, stream>>>(q, idx_t, row_stride, H, x16p, img, kp, pts4,        \
                                                                    amax_x_bits, amax_w_bits, extent, out, nq, ns, tq,   \
                                                                    n_tiles, order);                                     \
  } while (0)
  if (H <= 32)
    SPR_S(1);
  else if (H <= 64)
    SPR_S(2);
  else
    SPR_S(3);
#undef SPR_S
  SPR_LAUNCH_CHECK("k_kpconv_s");
  return SPR_OK;
}

template <int C>
int prepare_weights_s(const float* w, unsigned char* img, unsigned int* amax_w_bits, cudaStream_t stream) {
  using K = SCfg<C>;
  SPR_CUDA(cudaMemsetAsync(amax_w_bits, 0, sizeof(unsigned int), stream));
  k_absmax_s<<<(KP * C * C + 1023) / 1024, 256, 0, stream>>>(w, KP * C * C, amax_w_bits);
  SPR_LAUNCH_CHECK("k_absmax_s");
  constexpr int chunks = (int)(K::IMG_BYTES / 16);
  k_weight_image_s<C><<<(chunks + 255) / 256, 256, 0, stream>>>(w, amax_w_bits, img);
  SPR_LAUNCH_CHECK("k_weight_image_s");
  return SPR_OK;
}

}  // namespace
}  // namespace spr

using namespace spr;

extern "C" int spr_kpconv_staged_supported(int c, int H) {
  return (c == 32 || c == 64 || c == 128 || c == 256) && H > 0 && H <= 96;
}

extern "C" size_t spr_kpconv_staged_weight_image_bytes(int c) {
  switch (c) {
    case 32: return SCfg<32>::IMG_BYTES;
    case 64: return SCfg<64>::IMG_BYTES;
    case 128: return SCfg<128>::IMG_BYTES;
    case 256: return SCfg<256>::IMG_BYTES;
  }
  return 0;
}

extern "C" int spr_kpconv_staged_prepare_weights(const float* d_w, int c, void* d_img, void* d_amax_w, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(d_w && d_img && d_amax_w, "kpconv_staged_prepare_weights: null pointer");
  unsigned char* img = static_cast<unsigned char*>(d_img);
  unsigned int* am = static_cast<unsigned int*>(d_amax_w);
  switch (c) {
    case 32: return prepare_weights_s<32>(d_w, img, am, stream);
    case 64: return prepare_weights_s<64>(d_w, img, am, stream);
    case 128: return prepare_weights_s<128>(d_w, img, am, stream);
    case 256: return prepare_weights_s<256>(d_w, img, am, stream);
  }
  set_error("kpconv_staged_prepare_weights: unsupported channel count %d", c);
  return SPR_EUNSUPPORTED;
}

extern "C" int spr_kpconv_forward_staged(const float* d_q, const void* d_idx, int idx_is_64, int row_stride, int H,
                                         const void* d_pts4, const void* d_x16, const void* d_amax_x, int c,
                                         const void* d_wimg, const void* d_amax_w, const float* d_kp, float extent,
                                         float* d_out, int nq, int ns, const int32_t* d_order, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(nq > 0 && ns > 0 && H > 0 && row_stride >= H, "kpconv_forward_staged: bad shape");
  SPR_CHECK_ARG(extent > 0.f, "kpconv_forward_staged: extent must be > 0");
  SPR_CHECK_ARG(d_q && d_idx && d_pts4 && d_x16 && d_amax_x && d_wimg && d_amax_w && d_kp && d_out,
                "kpconv_forward_staged: null pointer");
  const float4* pts4 = static_cast<const float4*>(d_pts4);
  const unsigned char* x16p = static_cast<const unsigned char*>(d_x16);
  const unsigned char* img = static_cast<const unsigned char*>(d_wimg);
  const unsigned int* ax = static_cast<const unsigned int*>(d_amax_x);
  const unsigned int* aw = static_cast<const unsigned int*>(d_amax_w);
#define SPR_SP(CC)                                                                                                         \
  case CC:                                                                                                                 \
    return idx_is_64 ? launch_s<CC, long long>(d_q, d_idx, row_stride, H, x16p, img, d_kp, pts4, ax, aw, extent, d_out, nq, \
                                               ns, d_order, stream)                                                        \
                     : launch_s<CC, int>(d_q, d_idx, row_stride, H, x16p, img, d_kp, pts4, ax, aw, extent, d_out, nq, ns,  \
                                         d_order, stream);
  switch (c) {
    SPR_SP(32)
    SPR_SP(64)
    SPR_SP(128)
    SPR_SP(256)
  }
#undef SPR_SP
  set_error("kpconv_forward_staged: unsupported channel count %d", c);
  return SPR_EUNSUPPORTED;
}
