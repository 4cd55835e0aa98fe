// kpconv_tc.cu -- fused KPConv forward with the (K*Cin) x Cout contraction on the Blackwell tensor cores
// (tcgen05.mma, accumulators in tensor memory, weights streamed by the TMA engine).  mode 1 of
// spr_kpconv_forward; same mathematics as kpconv.cu (reference: kpconv_blocks.py:269-414).
//
// Work decomposition (one persistent CTA per SM, 20 warps):
//   warps 0-17  producers.  One warp per query and pass of 32 input channels:
//                    wf[k][c] = sum_h infl[h][k] * x[idx[h]][c]
//               is a 16 x H x 32 matrix product done with warp-level mma.sync on fp16 (hi, lo) pairs (fp32
//               accumulate); each lane evaluates exactly the influences of its A fragment and loads its B
//               fragments as 16-byte pieces of the gathered, pre-split feature rows: nothing staged or broadcast.
//               The 15 x 32 block is split into fp16 (hi, lo) pairs and written to the A tile in shared memory
//               (canonical K-major SWIZZLE_128B layout): row 2*ql holds hi, row 2*ql+1 holds lo, so a tile of
//               up to 64 queries fills the M = 128 rows of one tcgen05.mma.
//   warp 18     one thread issues the MMAs:  D[128 x 2C] += A[128 x 512] * B'[2C x 512]^T per pass of 32 input
//               channels, B' = [W_hi | W_lo] (fp16 pairs of the fp32 weights), so that
//                    out = hi*W_hi + hi*W_lo + lo*W_hi (+ lo*W_lo)
//               carries ~22 significant bits per operand: fp32-level accuracy from fp16 tensor-core products
//               accumulated in fp32.  D stays in TMEM across the C/32 passes of a tile.
//   warp 19     one thread streams the pre-swizzled weight image through a 3-stage shared-memory ring with bulk
//               async copies (TMA engine), completion on mbarriers.
//   epilogue    producers read D with tcgen05.ld, add the hi/lo rows (adjacent lanes) and the two column
//               halves, scale by 1/neighbour_count and store -- deferred into the next tile so that it never waits.
// Operand scaling: fp16 has a 5-bit exponent, so wf is computed pre-multiplied by a power of two derived from
// H*max|x| (kept <= 2^15) and W by one derived from max|W|; both are exact and undone in the epilogue.
// Synchronisation is mbarrier-only in steady state: a_full (18 producer warps -> MMA), mma_done (MMA -> producers,
// frees the A tile and publishes D), full/empty per ring stage.  The MMAs of pass p run while the producers
// already gather and accumulate the first queries of pass p+1; they only wait before overwriting the A tile.
#include "spr_common.cuh"
#include "tc05.cuh"

// scratch of the influence-fragment cache: one 16-byte vector per lane, neighbour block and query of a tile, per CTA
extern "C" size_t spr_kpconv_scratch_bytes(int H, int c) {
  if (c <= 32) return 0;
  const size_t nb = (size_t)((H > 0 ? H : 1) + 7) / 8;
  // per CTA (the grid never exceeds kNumSMs): two tile buffers x 64 queries x nb blocks x 32 lanes x uint4
  return (size_t)spr::kNumSMs * 2 * 64 * nb * 32 * 16;
}

namespace spr {

namespace {

using namespace tc;

constexpr int KP = 15;

struct TcScales {
  unsigned int amax_x_bits;  // max |x| as float bits (atomicMax on non-negative floats)
  unsigned int amax_w_bits;
};

// ---------------------------------------------------------------------------------------------
// pre-pass 1 (one warp per support row): the feature row pre-split into fp16 (hi, lo) pairs of x * 2^e with a
// per-row power-of-two scale; the packed support point (x, y, z, +-2^-e) whose sign carries rowsum(x) > 0 (the
// reference's neighbour_num mask, :409-412); max|x| and max|W| for the global scales of the tcgen05 operands
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) k_flags_absmax(const float* __restrict__ x, const float* __restrict__ s, int ns,
                                                       float4* __restrict__ pts4, uint32_t* __restrict__ x16,
                                                       const float* __restrict__ w, int n_w, int n_xblocks,
                                                       TcScales* __restrict__ sc) {
  const int lane = threadIdx.x & 31;
  if ((int)blockIdx.x >= n_xblocks) {  // weights: 4 floats per thread
    const int i = ((blockIdx.x - n_xblocks) * blockDim.x + threadIdx.x) * 4;
    float m = 0.f;
    if (i + 3 < n_w) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(w + i));
      m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    } else {
      for (int k = i; k < n_w; ++k) m = fmaxf(m, fabsf(w[k]));
    }
    m = warp_maxf(m);
    if (lane == 0 && __float_as_uint(m) > *reinterpret_cast<volatile unsigned int*>(&sc->amax_w_bits))
      atomicMax(&sc->amax_w_bits, __float_as_uint(m));
    return;
  }
  // a group of G = min(C/4, 32) lanes owns one row, each lane 4 consecutive channels per step of 4*G
  constexpr int G = C / 4 < 32 ? C / 4 : 32;
  constexpr int RPW = 32 / G;  // rows per warp
  constexpr int STEPS = C / (4 * G);
  const int row = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / G;
  const int gl = lane % G;
  const bool live = row < ns;
  float4 v[STEPS];
  float acc = 0.f, m = 0.f;
#pragma unroll
  for (int i = 0; i < STEPS; ++i) {
    v[i] = live ? __ldg(reinterpret_cast<const float4*>(x + (size_t)row * C + (i * G + gl) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    acc += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
  }
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(kFull, acc, o);
    m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
  }
  // the row as (hi, lo) fp16 pairs of x * 2^e, e chosen per row so that max|x| * 2^e <= 2^14
  const int e = scale_exp(m, 14);
  const float rs = pow2i(e);
  if (live) {
#pragma unroll
    for (int i = 0; i < STEPS; ++i) {
      const float a[4] = {v[i].x * rs, v[i].y * rs, v[i].z * rs, v[i].w * rs};
      uint32_t o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __half hi = __float2half_rn(a[k]);
        const __half lo = __float2half_rn(a[k] - __half2float(hi));
        o[k] = (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
      }
      *reinterpret_cast<uint4*>(x16 + (size_t)row * C + (i * G + gl) * 4) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (gl == 0) {
      const float inv = pow2i(-e);
      pts4[row] = make_float4(s[3 * (size_t)row], s[3 * (size_t)row + 1], s[3 * (size_t)row + 2], acc > 0.f ? inv : -inv);
    }
  }
  // one address for the whole grid: reduce over the warp and only touch it when the running maximum rises
  m = warp_maxf(m);
  if (lane == 0 && __float_as_uint(m) > *reinterpret_cast<volatile unsigned int*>(&sc->amax_x_bits))
    atomicMax(&sc->amax_x_bits, __float_as_uint(m));
}

// ---------------------------------------------------------------------------------------------
// pre-pass 2: weight image.  Block (pass, atom, sub) = NS rows x 128 B in the SWIZZLE_128B K-major layout the
// MMA reads: row r <-> column n' = sub*NS + r of B' = [W_hi | W_lo], K element kk of the atom <-> kernel point
// 2*atom + kk/32 (kernel point 15 is zero padding), input channel pass*32 + kk%32.  One thread per 16-byte chunk.
// ---------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) k_weight_image(const float* __restrict__ w,
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
    const int cin = pass * 32 + (j & 3) * 8 + e;
    float v = 0.f;
    if (k < KP) v = __ldg(w + ((size_t)k * C + cin) * C + o) * tscale;
    const __half hi = __float2half_rn(v);
    h[e] = lo_part ? __float2half_rn(v - __half2float(hi)) : hi;
  }
  *reinterpret_cast<uint4*>(img + (size_t)blk * (NS * 128) + sw128_offset(r, j)) = *reinterpret_cast<const uint4*>(h);
}

// ---------------------------------------------------------------------------------------------
// main kernel
// ---------------------------------------------------------------------------------------------
template <int C>
struct TcCfg {
  static constexpr int PASSES = C / 32;
  static constexpr int NCOL = 2 * C;
  static constexpr int NS = NCOL < 128 ? NCOL : 128;  // N of one MMA = rows of one ring stage
  static constexpr int NSUB = NCOL / NS;
  static constexpr int STAGE_BYTES = NS * 128;
  static constexpr int NSTAGES = 3;
  static constexpr int BLOCKS_PER_PASS = 8 * NSUB;    // 8 K atoms of 64 fp16 per pass
  static constexpr int TQ = 64;
  static constexpr int WORKERS = 18;
  static constexpr int THREADS = (WORKERS + 2) * 32;
  static constexpr int A_ATOM_BYTES = 128 * 128;
  static constexpr int A_BYTES = 8 * A_ATOM_BYTES;
  static constexpr int WBUF_BYTES = 0;
  static constexpr int TMEM_COLS = NCOL < 32 ? 32 : NCOL;
  static constexpr int OFF_RING = A_BYTES;
  static constexpr int OFF_WBUF = OFF_RING + NSTAGES * STAGE_BYTES;
  static constexpr int OFF_MISC = OFF_WBUF + WBUF_BYTES;
  static constexpr int MISC_BYTES = 16 * 8 + 32 + 2 * TQ * 4 + 48 * 4 + TQ * 4;
  static constexpr size_t SMEM = 1024 + OFF_MISC + MISC_BYTES;
  static constexpr size_t IMG_BYTES = (size_t)PASSES * BLOCKS_PER_PASS * STAGE_BYTES;
};

__device__ __forceinline__ float influence_fast(float cx, float cy, float cz, float kx, float ky, float kz,
                                                float inv_extent) {
  const float dx = cx - kx, dy = cy - ky, dz = cz - kz;
  const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
  float d;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(d) : "f"(d2));  // MUFU.SQRT, rel. error ~2^-23, sqrt(0) = 0
  return fmaxf(fmaf(-d, inv_extent, 1.f), 0.f);
}

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t h2_bits(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }
// d[16x8] += a[16x8] * b[8x8], fp16 operands, fp32 accumulate (warp-level tensor path)
__device__ __forceinline__ void mma_f16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}

template <typename IdxT>
__device__ __forceinline__ int load_idx(const IdxT* __restrict__ p) {
  return (int)__ldg(p);
}

template <int C, typename IdxT, int HR>  // HR = ceil(H / 32): 32-slot rounds of a neighbour row
__global__ void __launch_bounds__(TcCfg<C>::THREADS, 1)
    k_kpconv_tc(const float* __restrict__ q, const float* __restrict__ s, const IdxT* __restrict__ idx, int row_stride,
                int H, const uint32_t* __restrict__ x16, const unsigned char* __restrict__ wimg,
                const float* __restrict__ kp, const float4* __restrict__ pts4,
                const unsigned int* __restrict__ amax_x_bits, const unsigned int* __restrict__ amax_w_bits, float extent, float* __restrict__ out, int nq, int ns, int tq,
                int n_tiles, uint4* __restrict__ frag_scratch, const int* __restrict__ order) {
  using K = TcCfg<C>;
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
  volatile int* sFlag = reinterpret_cast<volatile int*>(sKp + 48);  // [TQ] tile number whose influence fragments are cached
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
    for (int i = 0; i < K::TQ; ++i) sFlag[i] = 0;
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
    // Phase 1 of a query is itself a small matrix product, wf[16 x 32] = Infl^T[16 x H] * X[H x 32], done with
    // warp-level mma.sync.m16n8k8 on fp16 (hi, lo) pairs (3 MMAs, fp32 accumulate): rows = kernel points,
    // K = neighbours in blocks of 8, N = 4 tiles of 8 channels.  With g = lane / 4, t = lane % 4 a lane evaluates
    // the influences of kernel points g and g+8 on neighbours 2t and 2t+1 of the block (exactly its A fragment)
    // and loads channels 4g..4g+3 of those two neighbours' pre-split feature rows as one 16-byte vector (its B
    // fragments for the four channel tiles: tile i pairs column g with channel 4g+i).  Nothing is staged in
    // shared memory and nothing is broadcast.
    const float inv_extent = 1.0f / extent;
    const float a_scale = pow2i(es);
    const float o_scale = pow2i(-(es + et));
    const int g = lane >> 2, t = lane & 3;
    const float k0x = sKp[3 * g], k0y = sKp[3 * g + 1], k0z = sKp[3 * g + 2];
    const float k1x = g < 7 ? sKp[3 * (g + 8)] : 0.f, k1y = g < 7 ? sKp[3 * (g + 8) + 1] : 0.f,
                k1z = g < 7 ? sKp[3 * (g + 8) + 2] : 0.f;
    const float k1_on = g < 7 ? 1.f : 0.f;  // row 15 of the fragment is padding
    uint32_t seq = 0;
    int titer = 0;
    // One "item" = one block of 8 neighbours of one query.  The loads of item i+1 (two packed support points and
    // two 16-byte pieces of pre-split feature rows per lane) are in flight while item i is multiplied; the
    // neighbour row of the warp's next query is requested when the current query starts.
    struct Item {
      float4 pa, pb;   // packed support points of neighbours 2t, 2t+1: x, y, z, w = +-2^-e (sign = rowsum flag)
      uint4 xa, xb;    // their feature pieces: channels 4g..4g+3 as (hi, lo) fp16 pairs scaled by 2^e
      int b;           // block index within the row
    };  // an absent neighbour has pa/pb = 0: w = 0 zeroes its influences
    // Influence fragments depend on the geometry only: pass 0 of a tile stores them (one 16-byte vector per lane and
    // block) in an L2-resident scratch, passes 1.. of the same tile reload them instead of re-evaluating 4 influences
    // and re-fetching two packed points per block.  sFlag[ql] publishes "fragments of query ql are written".
    const int nb_max = (H + 7) >> 3;
    // two buffers, by tile parity: a warp already in pass 0 of tile t+1 must not overwrite fragments that a slower
    // warp still reads in the last pass of tile t (warps are at most one pass apart, see bar_done)
    uint4* frag0 = frag_scratch + (size_t)blockIdx.x * 2 * K::TQ * nb_max * 32;
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
        const uint32_t* xcol = x16 + pass * 32 + 4 * g;
        bool first = true;
        // queries are dealt to the warps dynamically (neighbourhood sizes vary); the dispenser of pass seq+2 is
        // reset by whoever draws query 0 of pass seq (no warp can still be in pass seq-2, see bar_done)
        int* ctr = s_ctr + (seq & 3);
        auto grab = [&]() {
          int v = 0;
          if (lane == 0) v = atomicAdd(ctr, 1);
          return __shfl_sync(kFull, v, 0);
        };
        int ql = grab();
        if (ql == 0 && lane == 0) s_ctr[(seq + 2) & 3] = 0;
        int ql_next = 0;
        if (ql < cnt) {
          int jr[HR];           // neighbour row of the query being issued (lanes = slots), -1 = padding
          int jrn[HR];          // raw row of the next query (loads in flight)
          float qx, qy, qz;     // query point, issue side
          float qnx = 0.f, qny = 0.f, qnz = 0.f;
          unsigned bm;          // blocks of the row still to issue
          int ql_iss = ql;      // query whose row is being issued
          bool row_pending = false, new_query = true;
          float d[4][4];
          float fcount = 0.f;
          float cqx, cqy, cqz;  // query point of the item being multiplied
          auto issue_row = [&](int qq) {
            // `order` (optional) walks the queries in cell order: the queries in flight on an SM are spatial
            // neighbours, so their neighbourhoods overlap and the gathers hit L1 / L2
            const int n = order ? __ldg(order + q0 + qq) : q0 + qq;
#pragma unroll
            for (int i = 0; i < HR; ++i) {
              const int h = 32 * i + lane;
              jrn[i] = -1;
              if (h < H) jrn[i] = load_idx(idx + (size_t)n * row_stride + h);
            }
            qnx = __ldg(q + 3 * (size_t)n);
            qny = __ldg(q + 3 * (size_t)n + 1);
            qnz = __ldg(q + 3 * (size_t)n + 2);
          };
          auto consume_row = [&]() {
            bm = 0;
#pragma unroll
            for (int i = 0; i < HR; ++i) {
              jr[i] = -1;
              if (32 * i < H) {
                const bool valid = jrn[i] >= 0 && jrn[i] < ns;
                jr[i] = valid ? jrn[i] : -1;
                if (valid) {  // lanes = slots: two instructions pull the whole neighbourhood towards L1
                  prefetch_l1(pts4 + jrn[i]);
                  prefetch_l1(xcol - 4 * g + (size_t)jrn[i] * C);
                }
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
            if (K::PASSES > 1 && pass > 0) {
              // (>=: a faster warp may already have published this slot for the NEXT tile, whose fragments live in
              // the other buffer; the flags only grow)
              while (sFlag[ql_iss] < titer + 1) {
              }
              __threadfence_block();  // acquire: the fragment loads below stay behind the flag
            }
          };
          auto issue_item = [&](Item& it) {
            const int b = __ffs(bm) - 1;
            bm &= bm - 1;
            const int src = (b & 3) * 8 + 2 * t;
            int jsel = jr[0];
#pragma unroll
            for (int i = 1; i < HR; ++i)
              if ((b >> 2) == i) jsel = jr[i];
            const int ja = __shfl_sync(kFull, jsel, src);
            const int jb = __shfl_sync(kFull, jsel, src + 1);
            const size_t ra = ja >= 0 ? (size_t)ja : 0, rb = jb >= 0 ? (size_t)jb : 0;
            it.b = b;
            it.pa = make_float4(0.f, 0.f, 0.f, 0.f);
            it.pb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (K::PASSES > 1 && pass > 0) {
              // ordinary (L1-coherent within the SM) load: ld.cg would read L2, where a CTA-scope release does not
              // promise the data yet
              const uint4 f = *(frag0 + (((size_t)(titer & 1) * K::TQ + ql_iss) * nb_max + b) * 32 + lane);
              it.pa = make_float4(__uint_as_float(f.x), __uint_as_float(f.y), __uint_as_float(f.z), __uint_as_float(f.w));
            } else {
              if (ja >= 0) it.pa = __ldg(pts4 + ra);
              if (jb >= 0) it.pb = __ldg(pts4 + rb);
            }
            it.xa = __ldg(reinterpret_cast<const uint4*>(xcol + ra * C));
            it.xb = __ldg(reinterpret_cast<const uint4*>(xcol + rb * C));
          };
          // multiply `cur` while `nxt` loads; returns true when the warp has finished its queries of this pass
          auto step = [&](const Item& cur, Item& nxt) -> bool {
            if (new_query) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int e = 0; e < 4; ++e) d[i][e] = 0.f;
              fcount = 0.f;
              ql_next = grab();
              row_pending = ql_next < cnt;
              if (row_pending) issue_row(ql_next);
              new_query = false;
            }
            const bool last = bm == 0;  // cur is the last block of its query
            if (!last) {
              issue_item(nxt);
            } else if (row_pending) {
              ql_iss = ql_next;
              consume_row();
              issue_item(nxt);
            }
            {
              uint32_t ah0, ah1, al0, al1;
              if (K::PASSES > 1 && pass > 0) {
                ah0 = __float_as_uint(cur.pa.x);
                ah1 = __float_as_uint(cur.pa.y);
                al0 = __float_as_uint(cur.pa.z);
                al1 = __float_as_uint(cur.pa.w);
              } else {
                const float ax = cur.pa.x - cqx, ay = cur.pa.y - cqy, az = cur.pa.z - cqz;
                const float bx = cur.pb.x - cqx, by = cur.pb.y - cqy, bz = cur.pb.z - cqz;
                const float sa = fabsf(cur.pa.w) * a_scale;
                const float sb = fabsf(cur.pb.w) * a_scale;
                fcount += (cur.pa.w > 0.f ? 1.f : 0.f) + (cur.pb.w > 0.f ? 1.f : 0.f);
                // A fragment: a0 = (k = g; h = 2t, 2t+1), a1 = (k = g+8; h = 2t, 2t+1), fp16 hi + lo
                const float f00 = influence_fast(ax, ay, az, k0x, k0y, k0z, inv_extent) * sa;
                const float f01 = influence_fast(bx, by, bz, k0x, k0y, k0z, inv_extent) * sb;
                const float f10 = influence_fast(ax, ay, az, k1x, k1y, k1z, inv_extent) * (sa * k1_on);
                const float f11 = influence_fast(bx, by, bz, k1x, k1y, k1z, inv_extent) * (sb * k1_on);
                const __half2 h0 = __floats2half2_rn(f00, f01), h1 = __floats2half2_rn(f10, f11);
                const float2 h0f = __half22float2(h0), h1f = __half22float2(h1);
                const __half2 l0 = __floats2half2_rn(f00 - h0f.x, f01 - h0f.y), l1 = __floats2half2_rn(f10 - h1f.x, f11 - h1f.y);
                ah0 = h2_bits(h0);
                ah1 = h2_bits(h1);
                al0 = h2_bits(l0);
                al1 = h2_bits(l1);
                if (K::PASSES > 1)
                  frag0[(((size_t)(titer & 1) * K::TQ + ql) * nb_max + cur.b) * 32 + lane] = make_uint4(ah0, ah1, al0, al1);
              }
              const uint32_t xa[4] = {cur.xa.x, cur.xa.y, cur.xa.z, cur.xa.w};
              const uint32_t xb[4] = {cur.xb.x, cur.xb.y, cur.xb.z, cur.xb.w};
              // B fragment of channel tile i: (h = 2t, 2t+1; channel 4g+i).  Product-major order: consecutive
              // MMAs write different accumulators, so they pipeline instead of waiting on each other.
              uint32_t bh[4], bl[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                bh[i] = __byte_perm(xa[i], xb[i], 0x5410);  // (hi_a, hi_b)
                bl[i] = __byte_perm(xa[i], xb[i], 0x7632);  // (lo_a, lo_b)
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) mma_f16(d[i], al0, al1, bh[i]);
#pragma unroll
              for (int i = 0; i < 4; ++i) mma_f16(d[i], ah0, ah1, bl[i]);
#pragma unroll
              for (int i = 0; i < 4; ++i) mma_f16(d[i], ah0, ah1, bh[i]);
            }
            if (!last) return false;
            // ---- the query is complete: split to fp16 pairs and store its two A rows ----
            // the A tile still feeds the MMAs of the previous pass until bar_done completes
            if (first) {
              if (seq > 0) mbar_wait_park(bar_done, (seq - 1) & 1);
              if (pass == 0 && titer > 0) epilogue(prev_q0, prev_cnt, sInv + ((titer - 1) & 1) * K::TQ);
              first = false;
            }
            // this lane holds wf[k][8t..8t+7] for k = g (d[i][0], d[i][1]) and k = g+8 (d[i][2], d[i][3]):
            // one 16-byte chunk of fp16 each, K element = k * 32 + channel  ->  atom k / 2, chunk (k % 2) * 4 + t
            const uint32_t r0 = 2 * ql, r1 = 2 * ql + 1;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int e0 = 2 * half;  // d[i][e0] = channel 8t+i, d[i][e0+1] = channel 8t+4+i
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int p2 = 0; p2 < 4; ++p2) {
                const float v0 = d[(2 * p2) & 3][e0 + (p2 >> 1)], v1 = d[(2 * p2 + 1) & 3][e0 + (p2 >> 1)];
                const __half2 hh = __floats2half2_rn(v0, v1);
                const float2 hf = __half22float2(hh);
                const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
                hi[p2] = h2_bits(hh);
                lo[p2] = h2_bits(ll);
              }
              const int k = g + 8 * half;
              unsigned char* atom = sA + (k >> 1) * K::A_ATOM_BYTES;
              const uint32_t j = (k & 1) * 4 + t;
              *reinterpret_cast<uint4*>(atom + sw128_offset(r0, j)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(atom + sw128_offset(r1, j)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (K::PASSES > 1 && pass == 0) {
              // release at CTA scope: the fragments are read back by other warps of this CTA with ordinary loads
              // (same SM, same L1), after an acquire fence on the flag
              __threadfence_block();
              __syncwarp();
              if (lane == 0) sFlag[ql] = titer + 1;
            }
            if (pass == 0) {
              // a neighbour is replicated over g: count the g == 0 copies (lanes 0..3)
              float c = g == 0 ? fcount : 0.f;
              c += __shfl_xor_sync(kFull, c, 1);
              c += __shfl_xor_sync(kFull, c, 2);
              if (lane == 0) inv_buf[ql] = 1.f / fmaxf(c, 1.f);
            }
            ql = ql_next;
            new_query = true;
            cqx = qx;  // consume_row has already moved the issue side to the next query
            cqy = qy;
            cqz = qz;
            return ql >= cnt;
          };

          issue_row(ql);
          consume_row();
          cqx = qx;
          cqy = qy;
          cqz = qz;
          Item ia, ib;
          issue_item(ia);
          while (true) {
            if (step(ia, ib)) break;
            if (step(ib, ia)) break;
          }
        }
        if (first) {  // warp without a query in this pass
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

// max|W| -> *amax_w_bits (atomicMax on float bits; the word must be zero before the launch)
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ w, int n, unsigned int* __restrict__ amax_bits) {
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
  m = warp_maxf(m);
  if ((threadIdx.x & 31) == 0 && __float_as_uint(m) > *reinterpret_cast<volatile unsigned int*>(amax_bits))
    atomicMax(amax_bits, __float_as_uint(m));
}

template <int C>
int prepare_weights(const float* w, unsigned char* img, unsigned int* amax_w_bits, cudaStream_t stream) {
  using K = TcCfg<C>;
  SPR_CUDA(cudaMemsetAsync(amax_w_bits, 0, sizeof(unsigned int), stream));
  k_absmax<<<(KP * C * C + 1023) / 1024, 256, 0, stream>>>(w, KP * C * C, amax_w_bits);
  SPR_LAUNCH_CHECK("k_absmax");
  constexpr int chunks = (int)(K::IMG_BYTES / 16);
  k_weight_image<C><<<(chunks + 255) / 256, 256, 0, stream>>>(w, amax_w_bits, img);
  SPR_LAUNCH_CHECK("k_weight_image");
  return SPR_OK;
}

template <int C, typename IdxT>
int launch_main(const float* q, const void* idx, int row_stride, int H, const uint32_t* x16, const unsigned char* img,
                const float* kp, const float4* pts4, const unsigned int* amax_x_bits, const unsigned int* amax_w_bits,
                float extent, float* out, int nq, int ns, void* scratch, const int* order, cudaStream_t stream) {
  using K = TcCfg<C>;
  SPR_CHECK_ARG(K::PASSES == 1 || scratch, "kpconv_forward(mode 1): scratch buffer missing");
  uint4* frag = static_cast<uint4*>(scratch);
  SPR_CHECK_ARG(H <= 96, "kpconv_forward(mode 1): at most 96 neighbour columns are supported (got %d)", H);
  // Tile size: the largest tq <= 64 that deals every SM the same number of tiles (a tile is the M extent of
  // one MMA; short tiles only leave MMA rows unused, which costs nothing on the critical path).
  int tq = K::TQ;
  {
    const int waves = (nq + kNumSMs * K::TQ - 1) / (kNumSMs * K::TQ);
    const int per = (nq + kNumSMs * waves - 1) / (kNumSMs * waves);
    tq = per < 16 ? 16 : (per > K::TQ ? K::TQ : per);
  }
  const int n_tiles = (nq + tq - 1) / tq;
  SPR_CUDA(ensure_max_dynamic_smem(reinterpret_cast<const void*>(k_kpconv_tc<C, IdxT, 1>), K::SMEM));
  SPR_CUDA(ensure_max_dynamic_smem(reinterpret_cast<const void*>(k_kpconv_tc<C, IdxT, 2>), K::SMEM));
  SPR_CUDA(ensure_max_dynamic_smem(reinterpret_cast<const void*>(k_kpconv_tc<C, IdxT, 3>), K::SMEM));
  const int grid = n_tiles < kNumSMs ? n_tiles : kNumSMs;
  const IdxT* idx_t = static_cast<const IdxT*>(idx);
  const float* s_unused = nullptr;
  if (H <= 32)
    k_kpconv_tc<C, IdxT, 1><<<grid, K::THREADS, K::SMEM, stream>>>(q, s_unused, idx_t, row_stride, H, x16, img, kp, pts4,
                                                                   amax_x_bits, amax_w_bits, extent, out, nq, ns, tq,
                                                                   n_tiles, frag, order);
  else if (H <= 64)
    k_kpconv_tc<C, IdxT, 2><<<grid, K::THREADS, K::SMEM, stream>>>(q, s_unused, idx_t, row_stride, H, x16, img, kp, pts4,
                                                                   amax_x_bits, amax_w_bits, extent, out, nq, ns, tq,
                                                                   n_tiles, frag, order);
  else
    k_kpconv_tc<C, IdxT, 3><<<grid, K::THREADS, K::SMEM, stream>>>(q, s_unused, idx_t, row_stride, H, x16, img, kp, pts4,
                                                                   amax_x_bits, amax_w_bits, extent, out, nq, ns, tq,
                                                                   n_tiles, frag, order);
  SPR_LAUNCH_CHECK("k_kpconv_tc");
  return SPR_OK;
}

template <int C, typename IdxT>
int launch_tc(const float* q, const float* s, const void* idx, int row_stride, int H, const float* x, const float* w,
              const float* kp, float extent, float* out, int nq, int ns, void* workspace, cudaStream_t stream) {
  using K = TcCfg<C>;
  Carver cv(workspace, (size_t)-1);
  TcScales* sc = cv.take<TcScales>(1);
  float4* pts4 = cv.take<float4>((size_t)ns);
  uint32_t* x16 = cv.take<uint32_t>((size_t)ns * C);
  unsigned char* img = cv.take<unsigned char>(K::IMG_BYTES);
  void* scratch = cv.take<unsigned char>(spr_kpconv_scratch_bytes(H, C));

  SPR_CUDA(cudaMemsetAsync(sc, 0, sizeof(TcScales), stream));
  constexpr int rows_per_block = 8 * (C / 4 < 32 ? 32 / (C / 4) : 1);
  const int n_xblocks = (ns + rows_per_block - 1) / rows_per_block;
  const int n_wblocks = (KP * C * C + 1023) / 1024;
  k_flags_absmax<C><<<n_xblocks + n_wblocks, 256, 0, stream>>>(x, s, ns, pts4, x16, w, KP * C * C, n_xblocks, sc);
  SPR_LAUNCH_CHECK("k_flags_absmax");
  constexpr int chunks = (int)(K::IMG_BYTES / 16);
  k_weight_image<C><<<(chunks + 255) / 256, 256, 0, stream>>>(w, &sc->amax_w_bits, img);
  SPR_LAUNCH_CHECK("k_weight_image");
  return launch_main<C, IdxT>(q, idx, row_stride, H, x16, img, kp, pts4, &sc->amax_x_bits, &sc->amax_w_bits, extent, out,
                              nq, ns, scratch, nullptr, stream);
}

template <int C>
size_t tc_ws(int ns) {
  return 256 + align_up((size_t)ns * 16, 256) + align_up((size_t)ns * C * 4, 256) + 256 + TcCfg<C>::IMG_BYTES + 256 +
         spr_kpconv_scratch_bytes(96, C) + 256;
}

}  // namespace

size_t kpconv_tc_workspace_bytes(int ns, int c) {
  switch (c) {
    case 32: return tc_ws<32>(ns);
    case 64: return tc_ws<64>(ns);
    case 128: return tc_ws<128>(ns);
    case 256: return tc_ws<256>(ns);
  }
  return 0;
}

int kpconv_tc_forward(const float* q, const float* s, const void* idx, int idx_is_64, int row_stride, int H,
                      const float* x, int c, const float* w, const float* kp, float extent, float* out, int nq, int ns,
                      void* workspace, cudaStream_t stream) {
#define SPR_TC(CC)                                                                                                  \
  case CC:                                                                                                          \
    return idx_is_64 ? launch_tc<CC, long long>(q, s, idx, row_stride, H, x, w, kp, extent, out, nq, ns, workspace, \
                                                stream)                                                             \
                     : launch_tc<CC, int>(q, s, idx, row_stride, H, x, w, kp, extent, out, nq, ns, workspace, stream);
  switch (c) {
    SPR_TC(32)
    SPR_TC(64)
    SPR_TC(128)
    SPR_TC(256)
  }
#undef SPR_TC
  set_error("kpconv_forward(mode 1): unsupported channel count %d", c);
  return SPR_EUNSUPPORTED;
}

}  // namespace spr

using namespace spr;

extern "C" size_t spr_kpconv_weight_image_bytes(int c) {
  switch (c) {
    case 32: return TcCfg<32>::IMG_BYTES;
    case 64: return TcCfg<64>::IMG_BYTES;
    case 128: return TcCfg<128>::IMG_BYTES;
    case 256: return TcCfg<256>::IMG_BYTES;
  }
  return 0;
}

extern "C" int spr_kpconv_prepare_weights(const float* d_w, int c, void* d_img, void* d_amax_w, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(d_w && d_img && d_amax_w, "kpconv_prepare_weights: null pointer");
  unsigned char* img = static_cast<unsigned char*>(d_img);
  unsigned int* am = static_cast<unsigned int*>(d_amax_w);
  switch (c) {
    case 32: return prepare_weights<32>(d_w, img, am, stream);
    case 64: return prepare_weights<64>(d_w, img, am, stream);
    case 128: return prepare_weights<128>(d_w, img, am, stream);
    case 256: return prepare_weights<256>(d_w, img, am, stream);
  }
  set_error("kpconv_prepare_weights: unsupported channel count %d", c);
  return SPR_EUNSUPPORTED;
}

extern "C" int spr_kpconv_forward_prepared(const float* d_q, const void* d_idx, int idx_is_64, int row_stride, int H,
                                           const void* d_pts4, const void* d_x16, const void* d_amax_x, int c,
                                           const void* d_wimg, const void* d_amax_w, const float* d_kp, float extent,
                                           float* d_out, int nq, int ns, void* d_scratch, const int32_t* d_order,
                                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(nq > 0 && ns > 0 && H > 0 && row_stride >= H, "kpconv_forward_prepared: bad shape");
  SPR_CHECK_ARG(extent > 0.f, "kpconv_forward_prepared: extent must be > 0");
  SPR_CHECK_ARG(d_q && d_idx && d_pts4 && d_x16 && d_amax_x && d_wimg && d_amax_w && d_kp && d_out,
                "kpconv_forward_prepared: null pointer");
  const float4* pts4 = static_cast<const float4*>(d_pts4);
  const uint32_t* x16 = static_cast<const uint32_t*>(d_x16);
  const unsigned char* img = static_cast<const unsigned char*>(d_wimg);
  const unsigned int* ax = static_cast<const unsigned int*>(d_amax_x);
  const unsigned int* aw = static_cast<const unsigned int*>(d_amax_w);
#define SPR_TCP(CC)                                                                                                  \
  case CC:                                                                                                           \
    return idx_is_64 ? launch_main<CC, long long>(d_q, d_idx, row_stride, H, x16, img, d_kp, pts4, ax, aw, extent, d_out, \
                                                  nq, ns, d_scratch, d_order, stream)                                \
                     : launch_main<CC, int>(d_q, d_idx, row_stride, H, x16, img, d_kp, pts4, ax, aw, extent, d_out, nq, \
                                            ns, d_scratch, d_order, stream);
  switch (c) {
    SPR_TCP(32)
    SPR_TCP(64)
    SPR_TCP(128)
    SPR_TCP(256)
  }
#undef SPR_TCP
  set_error("kpconv_forward_prepared: unsupported channel count %d", c);
  return SPR_EUNSUPPORTED;
}
