// attention_tc.cu -- variable-length multi-head attention on the Blackwell tensor cores (tcgen05 + TMEM); same operator,
// operands and results as attention.cu (reference: models/transformer/transformers.py:184-245 via nn.MultiheadAttention,
// head_dim = 32), second generation.
//
// What attention.cu measured (profiles/README.md, round 2): the warp-level tensor path is 58 % busy at the 3DMatch shape
// and the kernel sits at `math pipe throttle` -- three split-precision products (hi*hi + lo*hi + hi*lo) of mma.sync per
// score and per value are the bound.  Here both products of a (128 queries x 64 keys) tile are tcgen05.mma instructions
// issued by one thread, the scores and the per-tile P V product live in TMEM, and a query row belongs to the NPART
// soft-max threads that own its TMEM lane (no shuffles; one shared-memory exchange of the partial row maximum per tile).
//
//   CTA = (tile of <= 128 queries, head); warps 0-7 soft-max (TMEM lane quadrant w % 4, key columns / output channels
//   half w / 4: a query row is shared by two threads that exchange their partial row maximum through shared memory),
//   warp 8 loads key / value tiles (16-byte cp.async, mbarrier completion), warp 9 issues the MMAs.
//   (The first version had one thread per row: 925 instructions per warp and tile in long dependent chains, two such
//   warps per scheduler -- 27 % issue utilisation.  Half rows, packed fp32 arithmetic and twice the warps halved it.)
//   NPART = 4 (a quarter row per thread, 16 soft-max warps, 56 registers) was measured too: 7-15 % SLOWER than NPART = 2
//   at all three bench shapes -- the per-tile fixed cost (barriers, exchange, TMEM loads) is paid by twice the warps.
//   shared memory (SWIZZLE_128B, 128-byte rows):
//     Q  [128 rows]: row = [Q_hi (32 halves) | Q_lo (32)]                        A operand of the score product, K-major
//     K  [64 rows] x 3 stages: row = [K_hi | K_lo]                               B operand, K-major
//     V  [64 rows] x 3 stages: row = [V_hi | V_lo]                               B operand of P V, MN-major (rows = keys)
//     P  2 atoms of [128 rows]: P_hi (64 keys), P_lo (64 keys)                   A operand of P V, K-major
//   S = Q K^T is SIX K = 16 steps whose descriptors walk the column blocks (Q_hi, K_hi), (Q_lo, K_hi), (Q_hi, K_lo): the
//   three products accumulate in one TMEM tile, nothing is stacked or summed afterwards.  D = P V is eight steps
//   (P_hi, P_lo) x [V_hi | V_lo] with N = 64: columns 0-31 + columns 32-63 of D are the tile's contribution to O
//   (the P_lo V_lo term is noise far below fp32 rounding).
//   Soft-max thread, per tile j:  S_j -> running maximum m_j, p = exp2(s - m_j) (Q arrives pre-scaled by
//   log2(e)/sqrt(32)), (hi, lo) split of p into the P tile;  O += D_(j-1);  O *= exp2(m_(j-1) - m_j).  D alternates
//   between two TMEM buffers, so the read-back of D_(j-1) runs while P_j V_j is being multiplied.
//   The MMA thread issues S_(j+1) before it waits for P_j, so the score product of the next tile and the P V product of
//   the previous one overlap the soft-max of the current one (two score buffers in TMEM).
#include "spr_common.cuh"
#include "tc05.cuh"

#include <cuda_fp16.h>

namespace spr {
namespace {

using namespace tc;

constexpr int HD = 32;
constexpr int BQ = 128;
constexpr int BK = 64;
constexpr int NSTAGE = 3;
constexpr int Q_BYTES = BQ * 128;
constexpr int KV_BYTES = BK * 128;
constexpr int P_ATOM = BQ * 128;
constexpr int OFF_P = Q_BYTES;
constexpr int OFF_K = OFF_P + 2 * P_ATOM;
constexpr int OFF_V = OFF_K + NSTAGE * KV_BYTES;
constexpr int OFF_BAR = OFF_V + NSTAGE * KV_BYTES;
#ifndef SPR_ATTN_NPART
#define SPR_ATTN_NPART 2
#endif
constexpr int NPART = SPR_ATTN_NPART;             // threads per query row (2 or 4)
constexpr int PCOLS = BK / NPART;                 // key columns of a tile per thread
constexpr int PCH = HD / NPART;                   // output channels per thread
constexpr int OFF_MAX = OFF_BAR + 256;             // [2 tiles][NPART][128 rows] partial row maxima
constexpr size_t ATTN_SMEM = 1024 + OFF_MAX + 2 * NPART * BQ * 4;
constexpr int SM_WARPS = 4 * NPART;
constexpr int ATTN_THREADS = (SM_WARPS + 2) * 32;
constexpr uint32_t TMEM_COLS = 256;  // S0 @ 0, S1 @ 64, D0 @ 128, D1 @ 192; two CTAs per SM share the 512 columns
static_assert(2 * ATTN_SMEM <= 227 * 1024, "two CTAs per SM");

__device__ unsigned int g_attention_tc_flags;

struct AttnTile {
  int q_row0;   // first query row (global row index into the planes)
  int q_rows;   // valid query rows in this tile (1..128)
  int kv_row0;  // first key/value row of the segment
  int kv_len;   // key/value rows of the segment
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// SWIZZLE_128B descriptor of an MN-major operand: rows are K (keys), 64 contiguous N elements (128 B) per row, 8 rows per
// 1024-byte group; the leading byte offset (a second block of 64 N elements) is not used at N = 64
__device__ __forceinline__ uint64_t desc_sw128_mnmajor(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(KV_BYTES >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16_bmn(int m, int n) {  // A K-major, B MN-major
  return (1u << 4) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8w(uint32_t taddr, float (&v)[8]) {
  tmem_ld8(taddr, v);
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ldn<32>(uint32_t taddr, float (&v)[32]) { tmem_ld32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ldn<16>(uint32_t taddr, float (&v)[16]) { tmem_ld16(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ldn<8>(uint32_t taddr, float (&v)[8]) { tmem_ld8w(taddr, v); }
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// named barrier of the NPART warps that share a TMEM lane quadrant
__device__ __forceinline__ void row_barrier(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(32 * NPART) : "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t h2u(__half2 v) { return *reinterpret_cast<const uint32_t*>(&v); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(ATTN_THREADS, 2)
    k_attention_tc(const __half* __restrict__ hi, const __half* __restrict__ lo, int ld, int q_col, int k_col, int v_col,
                   const AttnTile* __restrict__ tiles, float* __restrict__ out, int out_ld,
                   unsigned char* __restrict__ out_img, int img_katoms, float img_scale) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = bars;        // [NSTAGE] 32 copy arrivals
  uint64_t* kv_empty = bars + 4;   // [NSTAGE] tcgen05.commit
  uint64_t* s_full = bars + 8;     // [2] tcgen05.commit
  uint64_t* s_empty = bars + 10;   // [2] one arrival per soft-max warp
  uint64_t* p_full = bars + 12;    // one arrival per soft-max warp
  float* s_max = reinterpret_cast<float*>(smem + OFF_MAX);
  uint64_t* pv_done = bars + 13;   // tcgen05.commit
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 16);
  const AttnTile tl = tiles[blockIdx.x];
  if (tl.q_rows <= 0) return;  // padding entry of a bucketed tile list (CUDA-graph replays keep the grid fixed)
  const int head = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_kv = (tl.kv_len + BK - 1) / BK;
  const uint32_t sQ = smem_u32(smem), sP = sQ + OFF_P, sK = sQ + OFF_K, sV = sQ + OFF_V;

  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(&kv_full[i], 32);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], SM_WARPS);
    }
    mbar_init(p_full, SM_WARPS);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == SM_WARPS + 1) tmem_alloc(s_tmem, TMEM_COLS);
  __syncthreads();  // the barriers exist
  // ---- key / value loader (warp SM_WARPS) ----
  // ONE warp feeds the CTA: everything that does not change from copy to copy is a per-lane constant (the first
  // version recomputed row, chunk, plane and swizzle per copy -- ~800 dependent instructions per tile, and the
  // whole CTA ran at the speed of this warp).  Lane -> 16-byte chunk c = lane % 8 (0-3 hi plane, 4-7 lo plane) of rows
  // lane / 8 + 4 it, it = 0..15; rows r and r + 4 differ in their swizzle phase, rows r and r + 8 by 1024 bytes.
  const int c = lane & 7, r0 = lane >> 3;
  const __half* plane = (c & 4) ? lo : hi;
  const size_t col = (size_t)head * HD + (c & 3) * 8;
  const __half* kbase = plane + (size_t)tl.kv_row0 * ld + k_col + col;
  const __half* vbase = plane + (size_t)tl.kv_row0 * ld + v_col + col;
  const uint32_t off_e = r0 * 128 + ((c ^ r0) << 4), off_o = (r0 + 4) * 128 + ((c ^ (r0 + 4)) << 4);
  const size_t step = (size_t)4 * ld;
  auto load_tile = [&](int j) {
    const int st = j % NSTAGE;
    mbar_wait_park(&kv_empty[st], ((j / NSTAGE) & 1) ^ 1);
    const int kv0 = j * BK + r0;
    const uint32_t dK = sK + st * KV_BYTES, dV = sV + st * KV_BYTES;
    if (kv0 + 60 < tl.kv_len) {  // every row of this lane exists
      const __half* ks = kbase + (size_t)kv0 * ld;
      const __half* vs = vbase + (size_t)kv0 * ld;
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const uint32_t d = ((it & 1) ? off_o : off_e) + (it >> 1) * 1024;
        cp_async16(dK + d, ks, 16u);
        cp_async16(dV + d, vs, 16u);
        ks += step;
        vs += step;
      }
    } else {  // the segment's last tile: rows past its end are zero-filled
#pragma unroll 4
      for (int it = 0; it < 16; ++it) {
        const uint32_t d = ((it & 1) ? off_o : off_e) + (it >> 1) * 1024;
        const int kv = kv0 + 4 * it;
        const bool ok = kv < tl.kv_len;
        const size_t ro = (size_t)(ok ? kv : 0) * ld;
        cp_async16(dK + d, kbase + ro, ok ? 16u : 0u);
        cp_async16(dV + d, vbase + ro, ok ? 16u : 0u);
      }
    }
    cp_async_arrive_noinc(&kv_full[st]);
  };
  // The loader requests its first tiles while the other warps fetch the Q tile: the two latencies overlap (a CTA of the
  // 3DMatch shape lives for five key tiles only, so its start-up is a third of its life).
  int j_loaded = 0;
  if (warp == SM_WARPS) {
    for (; j_loaded < NSTAGE && j_loaded < n_kv; ++j_loaded) load_tile(j_loaded);
  } else {
    // Q tile: 128 rows x 8 chunks (4 hi, 4 lo); rows past the end of the segment are zero
    const int qt = tid < SM_WARPS * 32 ? tid : tid - 32;
    for (int i = qt; i < BQ * 8; i += ATTN_THREADS - 32) {
      const int r = i >> 3, c = i & 7;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (r < tl.q_rows)
        v = *reinterpret_cast<const uint4*>(((c & 4) ? lo : hi) + (size_t)(tl.q_row0 + r) * ld + q_col + head * HD + (c & 3) * 8);
      sts128(sQ + sw128_offset(r, c), v.x, v.y, v.z, v.w);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp < SM_WARPS) {
    // ============================================ soft-max ============================================
    // thread = (query row, part): key columns PCOLS * part .. of every tile, output channels PCH * part ..
    const int quad = warp & 3, part = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
    f2_t o2[PCH / 2];  // (o[2i], o[2i+1]) of this thread's channels
#pragma unroll
    for (int d = 0; d < PCH / 2; ++d) o2[d] = 0ull;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      mbar_wait_park(&s_full[sb], (j >> 1) & 1, 2000u);
      tc_fence_after();
      float s[PCOLS];
      tmem_ldn<PCOLS>(trow + sb * 64 + part * PCOLS, s);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);  // the score buffer may take tile j + 2
      const int valid = tl.kv_len - j * BK - part * PCOLS;  // columns >= valid are past the end of the segment
      if (valid < PCOLS) {
#pragma unroll
        for (int i = 0; i < PCOLS; ++i)
          if (i >= valid) s[i] = -INFINITY;
      }
      // row maximum: this thread's columns, then the other parts' through shared memory
      float mh[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) mh[u] = fmaxf(s[u], s[u + 4]);
#pragma unroll
      for (int i = 8; i < PCOLS; i += 8)
#pragma unroll
        for (int u = 0; u < 4; ++u) mh[u] = max3(mh[u], s[i + u], s[i + u + 4]);
      const float mine = fmaxf(fmaxf(mh[0], mh[1]), fmaxf(mh[2], mh[3]));
      float* mx_buf = s_max + sb * (NPART * BQ);
      mx_buf[part * BQ + row] = mine;
      row_barrier(1 + quad);
      float mx = fmaxf(m, mine);  // (never -inf: column 0 of tile 0 is valid)
#pragma unroll
      for (int o = 1; o < NPART; ++o) mx = fmaxf(mx, mx_buf[((part + o) % NPART) * BQ + row]);
      const float alpha = ex2(m - mx);  // first tile: exp2(-inf) = 0
      m = mx;
      const f2_t nmx = f2_pack(-mx, -mx);
      f2_t rs0 = 0ull, rs1 = 0ull;
#pragma unroll
      for (int i = 0; i < PCOLS / 2; i += 2) {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(f2_pack(s[2 * i], s[2 * i + 1]), nmx), a0, a1);
        f2_unpack(f2_add(f2_pack(s[2 * i + 2], s[2 * i + 3]), nmx), b0, b1);
        s[2 * i] = ex2(a0);
        s[2 * i + 1] = ex2(a1);
        s[2 * i + 2] = ex2(b0);
        s[2 * i + 3] = ex2(b1);
        rs0 = f2_add(rs0, f2_pack(s[2 * i], s[2 * i + 1]));
        rs1 = f2_add(rs1, f2_pack(s[2 * i + 2], s[2 * i + 3]));
      }
      {
        float r0, r1;
        f2_unpack(f2_add(rs0, rs1), r0, r1);
        l = fmaf(l, alpha, r0 + r1);
      }
      // (hi, lo) split of this thread's probabilities: registers only, BEFORE the wait for the P tile
      const f2_t neg1 = f2_pack(-1.f, -1.f);
      uint32_t ph[PCOLS / 2], pl[PCOLS / 2];
#pragma unroll
      for (int e = 0; e < PCOLS / 2; ++e) {
        const float v0 = s[2 * e], v1 = s[2 * e + 1];
        const __half2 hh = __floats2half2_rn(v0, v1);
        const float2 hf2 = __half22float2(hh);
        float l0, l1;
        f2_unpack(f2_fma(f2_pack(hf2.x, hf2.y), neg1, f2_pack(v0, v1)), l0, l1);  // v - hi, exact
        ph[e] = h2u(hh);
        pl[e] = h2u(__floats2half2_rn(l0, l1));
      }
      // the P tile is free once the previous tile's P V product has completed.  From here to the arrival on p_full
      // is the critical hand-over (soft-max -> P V -> soft-max): a few stores and a fence, nothing else
      if (j > 0) mbar_wait(pv_done, (j - 1) & 1);
#pragma unroll
      for (int c = 0; c < PCOLS / 8; ++c) {  // chunk = keys PCOLS part + 8 c .. + 7 of this row, hi atom and lo atom
        const uint32_t off = sw128_offset(row, (PCOLS / 8) * part + c);
        sts128(sP + off, ph[4 * c], ph[4 * c + 1], ph[4 * c + 2], ph[4 * c + 3]);
        sts128(sP + P_ATOM + off, pl[4 * c], pl[4 * c + 1], pl[4 * c + 2], pl[4 * c + 3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (j > 0) {  // O += D of the previous tile (its own TMEM buffer; relative to the previous maximum), then rescale
        tc_fence_after();
        float a[PCH], b[PCH];
        const uint32_t dcol = 128 + ((j - 1) & 1) * 64 + part * PCH;
        tmem_ldn<PCH>(trow + dcol, a);
        tmem_ldn<PCH>(trow + dcol + 32, b);
        const f2_t al2 = f2_pack(alpha, alpha);
#pragma unroll
        for (int d = 0; d < PCH / 2; ++d)
          o2[d] = f2_mul(f2_add(o2[d], f2_add(f2_pack(a[2 * d], a[2 * d + 1]), f2_pack(b[2 * d], b[2 * d + 1]))), al2);
        tc_fence_before();
      }
    }
    // the threads of a row have summed different keys: exchange the partial sums
    float* l_buf = s_max + (n_kv & 1) * (NPART * BQ);  // (the buffer the last tile did not use)
    l_buf[part * BQ + row] = l;
    row_barrier(1 + quad);
#pragma unroll
    for (int o = 1; o < NPART; ++o) l += l_buf[((part + o) % NPART) * BQ + row];
    // last tile's product, normalisation, store
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    float o[PCH];
    {
      float a[PCH], b[PCH];
      const uint32_t dcol = 128 + ((n_kv - 1) & 1) * 64 + part * PCH;
      tmem_ldn<PCH>(trow + dcol, a);
      tmem_ldn<PCH>(trow + dcol + 32, b);
      const float inv = 1.f / l;
#pragma unroll
      for (int d = 0; d < PCH / 2; ++d) {
        float x0, x1;
        f2_unpack(o2[d], x0, x1);
        o[2 * d] = (x0 + (a[2 * d] + b[2 * d])) * inv;
        o[2 * d + 1] = (x1 + (a[2 * d + 1] + b[2 * d + 1])) * inv;
      }
    }
    tc_fence_before();
    if (row < tl.q_rows) {
      const int token = tl.q_row0 + row;
      const int c0 = head * HD + PCH * part;  // first output column of this thread
      if (out) {
        float4* dst = reinterpret_cast<float4*>(out + (size_t)token * out_ld + c0);
#pragma unroll
        for (int i = 0; i < PCH / 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
      if (out_img) {
        // A image of the output projection (gemm_tc.cu): token -> tile token / 64, stacked rows 2r (hi), 2r + 1 (lo);
        // column c -> K atom c / 64, 16-byte chunk (c % 64) / 8, SWIZZLE_128B
        float amax16 = 0.f;
        const uint32_t r2 = 2 * (token & 63);
#pragma unroll
        for (int i = 0; i < PCH / 8; ++i) {
          const int c = c0 + 8 * i;
          uint32_t qh[4], ql[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float v0 = o[8 * i + 2 * e] * img_scale, v1 = o[8 * i + 2 * e + 1] * img_scale;
            amax16 = fmaxf(amax16, fmaxf(fabsf(v0), fabsf(v1)));
            const __half2 hh = __floats2half2_rn(v0, v1);
            const float2 hf2 = __half22float2(hh);
            qh[e] = h2u(hh);
            ql[e] = h2u(__floats2half2_rn(v0 - hf2.x, v1 - hf2.y));
          }
          unsigned char* blk = out_img + ((size_t)(token >> 6) * img_katoms + (c >> 6)) * 16384;
          const uint32_t chunk = (c & 63) >> 3;
          *reinterpret_cast<uint4*>(blk + sw128_offset(r2, chunk)) = make_uint4(qh[0], qh[1], qh[2], qh[3]);
          *reinterpret_cast<uint4*>(blk + sw128_offset(r2 + 1, chunk)) = make_uint4(ql[0], ql[1], ql[2], ql[3]);
        }
        if (!(amax16 <= 65504.f)) atomicOr(&g_attention_tc_flags, SPR_FLAG_FP16_OVERFLOW);  // false for NaN as well
      }
    }
  } else if (warp == SM_WARPS) {
    // ============================================ key / value loader ============================================
    for (int j = j_loaded; j < n_kv; ++j) load_tile(j);
  } else {
    // ============================================ MMA issuer ============================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_f16_f32(BQ, BK);
      constexpr uint32_t idesc_pv = idesc_f16_bmn(BQ, 2 * HD);
      auto issue_s = [&](int j) {
        const int st = j % NSTAGE, sb = j & 1;
        mbar_wait_park(&kv_full[st], (j / NSTAGE) & 1);
        mbar_wait_park(&s_empty[sb], ((j >> 1) & 1) ^ 1);
        fence_proxy_async_smem();  // the tiles were written through the generic proxy (cp.async)
        tc_fence_after();
        const uint32_t kt = sK + st * KV_BYTES;
#pragma unroll
        for (int ks = 0; ks < 6; ++ks) {
          // (Q_hi, K_hi), (Q_lo, K_hi), (Q_hi, K_lo): two K = 16 steps of the 32-wide head each
          const uint32_t a_off = (ks >= 2 && ks < 4 ? 64u : 0u) + (ks & 1) * 32u;
          const uint32_t b_off = (ks >= 4 ? 64u : 0u) + (ks & 1) * 32u;
          umma_f16(tmem + sb * 64, desc_sw128_kmajor(sQ + a_off), desc_sw128_kmajor(kt + b_off), idesc_s, ks != 0);
        }
        umma_commit(&s_full[sb]);
      };
      issue_s(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_s(j + 1);
        const int st = j % NSTAGE;
        mbar_wait(p_full, j & 1);  // (spinning: the soft-max -> P V -> soft-max hand-over is the critical loop)
        tc_fence_after();
        const uint32_t vt = sV + st * KV_BYTES;
#pragma unroll
        for (int part = 0; part < 2; ++part)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_f16(tmem + 128 + (j & 1) * 64, desc_sw128_kmajor(sP + part * P_ATOM + ks * 32), desc_sw128_mnmajor(vt + ks * 2048),
                     idesc_pv, (part | ks) != 0);
        umma_commit(pv_done);
        umma_commit(&kv_empty[st]);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SM_WARPS + 1) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

unsigned int attention_tc_numeric_flags(bool reset) {
  unsigned int v = 0;
  if (cudaMemcpyFromSymbol(&v, g_attention_tc_flags, sizeof(v)) != cudaSuccess) return 0;
  if (reset && v) {
    const unsigned int zero = 0;
    cudaMemcpyToSymbol(g_attention_tc_flags, &zero, sizeof(zero));
  }
  return v;
}

}  // namespace spr

using namespace spr;

extern "C" int spr_attention_varlen_tc(const void* d_hi, const void* d_lo, int ld, int q_col, int k_col, int v_col,
                                       int n_heads, int head_dim, const int32_t* d_tiles, int n_tiles, float* d_out,
                                       int out_ld, void* d_out_img, float img_scale, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(head_dim == HD, "attention_varlen_tc: head_dim must be %d (got %d)", HD, head_dim);
  SPR_CHECK_ARG(n_tiles > 0 && n_heads > 0, "attention_varlen_tc: empty launch");
  SPR_CHECK_ARG((ld & 7) == 0 && (q_col & 7) == 0 && (k_col & 7) == 0 && (v_col & 7) == 0 && (out_ld & 3) == 0,
                "attention_varlen_tc: row strides / column offsets must keep 16-byte alignment");
  SPR_CHECK_ARG(d_hi && d_lo && d_tiles && (d_out || d_out_img), "attention_varlen_tc: null pointer");
  SPR_CUDA(ensure_max_dynamic_smem(reinterpret_cast<const void*>(k_attention_tc), ATTN_SMEM));
  dim3 grid(n_tiles, n_heads);
  k_attention_tc<<<grid, ATTN_THREADS, ATTN_SMEM, stream>>>(
      static_cast<const __half*>(d_hi), static_cast<const __half*>(d_lo), ld, q_col, k_col, v_col,
      reinterpret_cast<const AttnTile*>(d_tiles), d_out, out_ld, static_cast<unsigned char*>(d_out_img),
      (n_heads * HD + 63) / 64, img_scale);
  SPR_LAUNCH_CHECK("k_attention_tc");
  return SPR_OK;
}
