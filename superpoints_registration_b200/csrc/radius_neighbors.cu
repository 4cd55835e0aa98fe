// radius_neighbors.cu -- batched fixed-radius neighbour search on sm_100a (uniform cell list).
//
// Replaces batch_nanoflann_neighbors (reference: models/backbone_kpconv/cpp_wrappers/cpp_neighbors/neighbors/
// neighbors.cpp:211-332) plus the [:, :max_neighbors] truncation of kpconv.py:259-260.
//
// Build (spr_cell_grid_build): per-cloud bounding box of the supports -> per-cloud uniform grid with
// cell edge c >= radius*(1+1/256) (doubled until the cloud's cells fit its share of the cell table) ->
// counting sort of the supports by (cloud, z, y, x) cell: histogram, exclusive scan, scatter.  Sorted
// supports are stored as float4 (x, y, z, original global index) so a candidate costs one 16-byte load.
//
// Query (spr_radius_query): ONE WARP PER QUERY.  The 27 neighbouring cells are 9 contiguous runs in the
// sorted array (the 3 x-adjacent cells of a (z,y) row are adjacent in memory); 9 lanes fetch the run
// bounds, the warp then streams the concatenated runs 32 candidates at a time (coalesced float4 loads),
// tests d2 < r2 with the reference's exact fp32 rounding sequence, and stages the hits in shared memory.
// The row is finished by ranking the staged hits by (d2, index) -- the reference sorts by distance
// (nanoflann.hpp:1286-1287) and Python keeps the first `limit` -- and writing hit e to column rank(e).
// If more than kStage hits accumulate, the stage is compacted to the best `limit` and a (d2, index)
// admission threshold is kept, so arbitrarily dense neighbourhoods are handled in bounded memory.
#include "spr_common.cuh"

namespace spr {
namespace {

constexpr int kThreads = 256;
constexpr int kWarpsPerBlock = 8;
constexpr int kStage = 320;  // staged hits per warp (>= 2*SPR_MAX_NEIGHBOR_LIMIT + 32)
constexpr int kMaxDim = 1024;

struct CellGrid {  // per cloud
  float lox, loy, loz, inv_cell;  // cell edge stored as its reciprocal: coordinates are one subtract + one multiply
  int dx, dy, dz;
  int base;  // first cell of this cloud in the global cell table
};

// Workspace header (device): lives at the start of the grid workspace.
struct GridHeader {
  int total_cells;
  int pad[3];
};


// Cells a cloud of ns points may use.  Must match between workspace sizing and planning.
__host__ __device__ inline long long cell_budget(int ns) { return 16ll * ns + 512ll; }

// One thread per cloud chooses the cell edge; a serial prefix over clouds assigns table bases.
__global__ void k_plan_grids(const uint32_t* __restrict__ bb, const int* __restrict__ offs, int B, float radius,
                             CellGrid* __restrict__ grids, GridHeader* __restrict__ hdr) {
  __shared__ int s_cells[1024];
  __shared__ int s_total;
  int carry = 0;
  for (int base = 0; base < B; base += blockDim.x) {
    const int b = base + threadIdx.x;
    int cells = 0;
    CellGrid g;
    if (b < B) {
      const int ns = offs[b + 1] - offs[b];
      float lo[3], hi[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        lo[a] = ns > 0 ? ord2f(bb[3 * b + a]) : 0.f;
        hi[a] = ns > 0 ? ord2f(bb[3 * (B + b) + a]) : 0.f;
      }
      // margin: a pair accepted by the fp32 test d2 < r2 can never sit two cells apart
      float cell = radius * 1.00390625f;
      int d[3];
      const long long budget = cell_budget(ns);
      for (int it = 0; it < 64; ++it) {
        bool ok = true;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          double v = floor(((double)hi[a] - (double)lo[a]) / (double)cell) + 1.0;
          if (!(v <= (double)kMaxDim)) ok = false;
          d[a] = v < 1.0 ? 1 : (v > (double)kMaxDim ? kMaxDim : (int)v);
        }
        if (ok && (long long)d[0] * d[1] * d[2] <= budget) break;
        cell *= 2.f;
      }
      g.lox = lo[0];
      g.loy = lo[1];
      g.loz = lo[2];
      g.inv_cell = 1.f / cell;
      g.dx = d[0];
      g.dy = d[1];
      g.dz = d[2];
      cells = d[0] * d[1] * d[2];
    }
    s_cells[threadIdx.x] = cells;
    __syncthreads();
    if (threadIdx.x == 0) {  // B is small (2 x pairs): a serial prefix is fine
      int run = carry;
      for (int t = 0; t < blockDim.x && base + t < B; ++t) {
        int c = s_cells[t];
        s_cells[t] = run;
        run += c;
      }
      s_total = run;
    }
    __syncthreads();
    if (b < B) {
      g.base = s_cells[threadIdx.x];
      grids[b] = g;
    }
    carry = s_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) hdr->total_cells = carry;
}

__device__ __forceinline__ int cell_coord(float p, float lo, float inv_cell, int dim) {
  // fp32, same expression at build and query time.  The quotient is off by < 1e-6 relative, the cell edge exceeds
  // the radius by 1/256: two points within the radius still land at most one cell apart (floor is monotonic).
  float u = floorf(__fmul_rn(__fsub_rn(p, lo), inv_cell));
  u = fminf(fmaxf(u, -2.f), (float)dim + 1.f);
  return (int)u;
}

__global__ void __launch_bounds__(kThreads)
    k_cell_count(const float* __restrict__ s, const int* __restrict__ offs, int B, int n,
                 const CellGrid* __restrict__ grids, int* __restrict__ cell_cnt, int* __restrict__ point_cell) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = find_cloud(offs, B, i);
  const CellGrid g = grids[b];
  int cx = min(max(cell_coord(s[3 * (size_t)i + 0], g.lox, g.inv_cell, g.dx), 0), g.dx - 1);
  int cy = min(max(cell_coord(s[3 * (size_t)i + 1], g.loy, g.inv_cell, g.dy), 0), g.dy - 1);
  int cz = min(max(cell_coord(s[3 * (size_t)i + 2], g.loz, g.inv_cell, g.dz), 0), g.dz - 1);
  const int c = g.base + (cz * g.dy + cy) * g.dx + cx;
  point_cell[i] = c;
  atomicAdd(cell_cnt + c, 1);
}

__global__ void __launch_bounds__(kThreads)
    k_cell_scatter(const float* __restrict__ s, int n, const int* __restrict__ point_cell,
                   const int* __restrict__ cell_start, int* __restrict__ cell_fill, float4* __restrict__ sorted) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = point_cell[i];
  const int pos = cell_start[c] + atomicAdd(cell_fill + c, 1);
  sorted[pos] = make_float4(s[3 * (size_t)i], s[3 * (size_t)i + 1], s[3 * (size_t)i + 2], __int_as_float(i));
}

// A hit is one 64-bit key: fp32 bits of d2 (non-negative, so the bit pattern orders like the value) in the high
// word, support index in the low word.  key_a < key_b  <=>  (d2, idx) of a comes before b: the reference's order
// by distance (neighbors.cpp:296-310 via nanoflann's sorted result set) with ties by lower index.
__device__ __forceinline__ unsigned long long make_key(float d2, int id) {
  return ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)id;
}

// Keep the `limit` best staged hits, packed at the front in rank order (rare path: more than kStage hits).
__device__ __forceinline__ int compact_stage(unsigned long long* __restrict__ sk, int cnt, int limit, int lane) {
  unsigned long long kk[(kStage + 31) / 32];
  int kr[(kStage + 31) / 32];
#pragma unroll
  for (int t = 0; t < (kStage + 31) / 32; ++t) {
    const int e = t * 32 + lane;
    kr[t] = 0x7fffffff;
    if (e < cnt) {
      const unsigned long long k = sk[e];
      int r = 0;
      for (int f = 0; f < cnt; ++f) r += sk[f] < k ? 1 : 0;
      kk[t] = k;
      kr[t] = r;
    }
  }
  __syncwarp();
#pragma unroll
  for (int t = 0; t < (kStage + 31) / 32; ++t)
    if (kr[t] < limit) sk[kr[t]] = kk[t];
  __syncwarp();
  return min(cnt, limit);
}

// BY_INDEX: rows hold the first `limit` in-radius supports in INDEX order (pytorch3d ball_query as the reference's
// PreprocessorGPU calls it, kpconv.py:280-286) instead of the `limit` nearest in distance order: the same staging
// and ranking with the distance word of the key left at zero.
template <typename IdxT, bool BY_INDEX>
__global__ void __launch_bounds__(kThreads, 4)
    k_radius_query(const float* __restrict__ q, const int* __restrict__ q_offs, int B, int nq,
                   const CellGrid* __restrict__ grids, const int* __restrict__ cell_start,
                   const float4* __restrict__ sorted, int ns_total, float r2, int limit, IdxT* __restrict__ out,
                   int row_stride, int* __restrict__ max_count) {
  __shared__ __align__(16) unsigned long long s_k[kWarpsPerBlock][kStage];
  __shared__ int s_beg[kWarpsPerBlock][9];
  __shared__ int s_out[kWarpsPerBlock][SPR_MAX_NEIGHBOR_LIMIT];  // the row in rank order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long* sk = s_k[warp];
  int block_max = 0;

  for (int qi = blockIdx.x * kWarpsPerBlock + warp; qi < nq; qi += gridDim.x * kWarpsPerBlock) {
    // cloud of the query: the lanes compare the end offsets of 32 clouds at a time
    int b = 0;
    for (int c0 = 0; c0 < B; c0 += 32) {
      const int end = c0 + lane < B ? __ldg(q_offs + c0 + lane + 1) : 0x7fffffff;
      b += __popc(__ballot_sync(kFull, end <= qi));
    }
    const CellGrid g = grids[b];
    const float px = __ldg(q + 3 * (size_t)qi), py = __ldg(q + 3 * (size_t)qi + 1), pz = __ldg(q + 3 * (size_t)qi + 2);
    const int cx = cell_coord(px, g.lox, g.inv_cell, g.dx);
    const int cy = cell_coord(py, g.loy, g.inv_cell, g.dy);
    const int cz = cell_coord(pz, g.loz, g.inv_cell, g.dz);

    // 9 runs: lane r -> (dz, dy) = (r/3-1, r%3-1)
    int beg = 0, len = 0;
    if (lane < 9) {
      const int z = cz + lane / 3 - 1, y = cy + lane % 3 - 1;
      const int xlo = max(cx - 1, 0), xhi = min(cx + 1, g.dx - 1);
      if (z >= 0 && z < g.dz && y >= 0 && y < g.dy && xlo <= xhi) {
        const int row = g.base + (z * g.dy + y) * g.dx;
        beg = __ldg(cell_start + row + xlo);
        len = __ldg(cell_start + row + xhi + 1) - beg;
      }
    }
    int inc = len;  // inclusive prefix of the run lengths over lanes 0..8
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
      int t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane < 9) s_beg[warp][lane] = beg - (inc - len);  // candidate t of run r lives at sorted[s_beg[r] + t]
    __syncwarp();
    int pre[8];  // pre[k] = first flat candidate index of run k + 1, in registers for the run lookup of every pass
#pragma unroll
    for (int k = 0; k < 8; ++k) pre[k] = __shfl_sync(kFull, inc, k);
    const int total = __shfl_sync(kFull, inc, 8);

    int cnt = 0;         // staged hits
    int in_radius = 0;   // all hits (for max_count)
    unsigned long long thr = ~0ull;  // admission threshold: the limit-th best key once the stage was compacted
    for (int t0 = 0; t0 < total; t0 += 32) {
      const int t = t0 + lane;
      bool hit = false;
      unsigned long long key = ~0ull;
      if (t < total) {
        int r = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) r += (t >= pre[k]) ? 1 : 0;
        const float4 c = __ldg(sorted + s_beg[warp][r] + t);
        const float d2 = sqdist_exact(px, py, pz, c.x, c.y, c.z);
        hit = d2 < r2;
        key = BY_INDEX ? (unsigned long long)(unsigned int)__float_as_int(c.w) : make_key(d2, __float_as_int(c.w));
      }
      const unsigned hm = __ballot_sync(kFull, hit);
      in_radius += __popc(hm);
      const bool admit = hit && key < thr;
      const unsigned am = __ballot_sync(kFull, admit);
      if (am) {
        if (cnt + __popc(am) > kStage) {  // warp-uniform
          cnt = compact_stage(sk, cnt, limit, lane);
          if (cnt == limit) thr = sk[limit - 1];
          // re-test this batch against the new threshold
          const bool admit2 = admit && key < thr;
          const unsigned am2 = __ballot_sync(kFull, admit2);
          if (admit2) sk[cnt + __popc(am2 & ((1u << lane) - 1u))] = key;
          cnt += __popc(am2);
        } else {
          if (admit) sk[cnt + __popc(am & ((1u << lane) - 1u))] = key;
          cnt += __popc(am);
        }
        __syncwarp();
      }
    }
    block_max = max(block_max, in_radius);

    // rank the staged hits and emit the row
    IdxT* __restrict__ row = out + (size_t)qi * row_stride;
    int* so = s_out[warp];
    if (cnt <= 32) {
      // most rows: one staged hit per lane, ranked by counting the smaller keys (two keys per 16-byte load)
      const unsigned long long k0 = lane < cnt ? sk[lane] : ~0ull;
      if (lane == 0) sk[cnt] = ~0ull;  // sentinel for the odd tail (kStage > 32, slot cnt is free)
      __syncwarp();
      int r0 = 0;
      const ulonglong2* sk2 = reinterpret_cast<const ulonglong2*>(sk);
      for (int f = 0; f < cnt; f += 2) {
        const ulonglong2 kf = sk2[f >> 1];
        r0 += kf.x < k0 ? 1 : 0;
        r0 += kf.y < k0 ? 1 : 0;
      }
      if (lane < cnt && r0 < limit) so[r0] = (int)(unsigned int)k0;
    } else if (cnt <= 64) {
      // a lane owns staged hits `lane` and `lane + 32`; ONE pass over the stage ranks both
      const unsigned long long k0 = lane < cnt ? sk[lane] : ~0ull, k1 = lane + 32 < cnt ? sk[lane + 32] : ~0ull;
      if (lane == 0) sk[cnt] = ~0ull;  // sentinel for the odd tail
      __syncwarp();
      int r0 = 0, r1 = 0;
      const ulonglong2* sk2 = reinterpret_cast<const ulonglong2*>(sk);
      for (int f = 0; f < cnt; f += 2) {
        const ulonglong2 kf = sk2[f >> 1];
        r0 += kf.x < k0 ? 1 : 0;
        r1 += kf.x < k1 ? 1 : 0;
        r0 += kf.y < k0 ? 1 : 0;
        r1 += kf.y < k1 ? 1 : 0;
      }
      if (lane < cnt && r0 < limit) so[r0] = (int)(unsigned int)k0;
      if (lane + 32 < cnt && r1 < limit) so[r1] = (int)(unsigned int)k1;
    } else {
      for (int e = lane; e < cnt; e += 32) {
        const unsigned long long k = sk[e];
        int r = 0;
        for (int f = 0; f < cnt; ++f) r += sk[f] < k ? 1 : 0;
        if (r < limit) so[r] = (int)(unsigned int)k;
      }
    }
    // the ranks are a permutation of 0..cnt-1 (keys are distinct): the ordered row sits in shared memory and goes out
    // as contiguous stores, shadow-padded
    __syncwarp();
    const int nvalid = min(cnt, limit);
    for (int j = lane; j < limit; j += 32) row[j] = (IdxT)(j < nvalid ? so[j] : ns_total);
    __syncwarp();
  }
  if (lane == 0 && block_max > 0) atomicMax(max_count, block_max);
}

}  // namespace
}  // namespace spr

using namespace spr;

namespace {
struct GridLayout {
  GridHeader* hdr;
  int* offs;
  uint32_t* bb;
  CellGrid* grids;
  int* cell_start;  // [cap+1]
  int* cell_fill;   // [cap]
  int* point_cell;  // [ns]
  float4* sorted;   // [ns]
  int* scan_tmp;
  int* q_offs;      // [B+1] scratch for the query-side offsets (queries on one grid must be stream-ordered)
  size_t cap;
  size_t bytes;
};

GridLayout carve_grid(void* ws, size_t ws_bytes, int ns, int B) {
  GridLayout L;
  L.cap = (size_t)(cell_budget(0) * (long long)B + 16ll * ns);
  Carver c(ws, ws_bytes);
  L.hdr = c.take<GridHeader>(1);
  L.offs = c.take<int>((size_t)B + 1);
  L.bb = c.take<uint32_t>((size_t)B * 6);
  L.grids = c.take<CellGrid>(B);
  L.cell_start = c.take<int>(L.cap + 1);
  L.cell_fill = c.take<int>(L.cap);
  L.point_cell = c.take<int>(ns);
  L.sorted = c.take<float4>(ns);
  L.scan_tmp = c.take<int>(scan_tmp_ints(L.cap + 1));
  L.q_offs = c.take<int>((size_t)B + 1);
  L.bytes = c.off;
  return L;
}
}  // namespace

namespace spr {
namespace {
__global__ void __launch_bounds__(256) k_grid_order(const float4* __restrict__ sorted, int n, int32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float_as_int(sorted[i].w);
}
}  // namespace
}  // namespace spr

// order[i] = index of the i-th support in (cloud, z, y, x) cell order: a spatially coherent traversal of the cloud
extern "C" int spr_cell_grid_order(const void* d_grid_workspace, int n_supports, int n_clouds, int32_t* d_order,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(d_grid_workspace && d_order && n_supports > 0 && n_clouds > 0, "cell_grid_order: bad arguments");
  GridLayout L = carve_grid(const_cast<void*>(d_grid_workspace), (size_t)-1, n_supports, n_clouds);
  k_grid_order<<<(n_supports + 255) / 256, 256, 0, stream>>>(L.sorted, n_supports, d_order);
  SPR_LAUNCH_CHECK("k_grid_order");
  return SPR_OK;
}

extern "C" size_t spr_cell_grid_workspace_bytes(int n_supports, int n_clouds) {
  if (n_supports < 0 || n_clouds < 0) return 0;
  GridLayout L = carve_grid(nullptr, 0, n_supports, n_clouds);
  return L.bytes + 1024;
}

extern "C" int spr_cell_grid_build(const float* d_supports, const int32_t* d_s_lengths, int n_supports, int n_clouds,
                                   float radius, void* d_grid_workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(n_supports > 0 && n_clouds > 0, "cell_grid_build: empty input (n_supports=%d, n_clouds=%d)", n_supports,
                n_clouds);
  SPR_CHECK_ARG(radius > 0.f, "cell_grid_build: radius must be > 0");
  SPR_CHECK_ARG(d_supports && d_s_lengths && d_grid_workspace, "cell_grid_build: null pointer");
  if (workspace_bytes < spr_cell_grid_workspace_bytes(n_supports, n_clouds)) {
    set_error("cell_grid_build: workspace too small");
    return SPR_ENOSPACE;
  }
  const int ns = n_supports, B = n_clouds;
  GridLayout L = carve_grid(d_grid_workspace, workspace_bytes, ns, B);
  const int gp = (ns + kThreads - 1) / kThreads;
  int rc = cloud_offsets(d_s_lengths, B, L.offs, stream);
  if (rc) return rc;
  rc = cloud_bboxes(d_supports, L.offs, B, ns, L.bb, stream);
  if (rc) return rc;
  k_plan_grids<<<1, 1024, 0, stream>>>(L.bb, L.offs, B, radius, L.grids, L.hdr);
  SPR_LAUNCH_CHECK("k_plan_grids");
  SPR_CUDA(cudaMemsetAsync(L.cell_start, 0, (L.cap + 1) * 4, stream));
  SPR_CUDA(cudaMemsetAsync(L.cell_fill, 0, L.cap * 4, stream));
  k_cell_count<<<gp, kThreads, 0, stream>>>(d_supports, L.offs, B, ns, L.grids, L.cell_start, L.point_cell);
  SPR_LAUNCH_CHECK("k_cell_count");
  rc = exclusive_scan_i32(L.cell_start, L.cell_start, L.cap + 1, nullptr, L.scan_tmp, stream);
  if (rc) return rc;
  k_cell_scatter<<<gp, kThreads, 0, stream>>>(d_supports, ns, L.point_cell, L.cell_start, L.cell_fill, L.sorted);
  SPR_LAUNCH_CHECK("k_cell_scatter");
  return SPR_OK;
}

extern "C" int spr_radius_query_ex(const float* d_queries, const int32_t* d_q_lengths, int n_queries, int n_clouds,
                                   const void* d_grid_workspace, int n_supports, float radius, int limit, int order,
                                   void* d_out_idx, int idx_is_64, int row_stride, int32_t* d_out_max_count,
                                   void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(n_queries > 0 && n_clouds > 0 && n_supports > 0, "radius_query: empty input (nq=%d, ns=%d, B=%d)",
                n_queries, n_supports, n_clouds);
  SPR_CHECK_ARG(limit > 0 && limit <= SPR_MAX_NEIGHBOR_LIMIT, "radius_query: limit %d outside [1, %d]", limit,
                SPR_MAX_NEIGHBOR_LIMIT);
  SPR_CHECK_ARG(row_stride >= limit, "radius_query: row_stride < limit");
  SPR_CHECK_ARG(radius > 0.f, "radius_query: radius must be > 0");
  SPR_CHECK_ARG(order == SPR_ORDER_NEAREST || order == SPR_ORDER_INDEX, "radius_query: unknown row order %d", order);
  SPR_CHECK_ARG(d_queries && d_q_lengths && d_grid_workspace && d_out_idx && d_out_max_count, "radius_query: null pointer");
  GridLayout L = carve_grid(const_cast<void*>(d_grid_workspace), (size_t)-1, n_supports, n_clouds);
  int* q_offs = L.q_offs;
  int rc = cloud_offsets(d_q_lengths, n_clouds, q_offs, stream);
  if (rc) return rc;
  SPR_CUDA(cudaMemsetAsync(d_out_max_count, 0, sizeof(int32_t), stream));
  const float r2 = radius * radius;  // fp32 product, as neighbors.cpp:226
  int blocks = (n_queries + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int max_blocks = kNumSMs * 8 * 4;
  if (blocks > max_blocks) blocks = max_blocks;
#define SPR_RQ(T, BYI)                                                                                               \
  k_radius_query<T, BYI><<<blocks, kThreads, 0, stream>>>(d_queries, q_offs, n_clouds, n_queries, L.grids, L.cell_start, \
                                                          L.sorted, n_supports, r2, limit, static_cast<T*>(d_out_idx),   \
                                                          row_stride, d_out_max_count)
  if (order == SPR_ORDER_INDEX) {
    if (idx_is_64) SPR_RQ(long long, true); else SPR_RQ(int, true);
  } else {
    if (idx_is_64) SPR_RQ(long long, false); else SPR_RQ(int, false);
  }
#undef SPR_RQ
  SPR_LAUNCH_CHECK("k_radius_query");
  return SPR_OK;
}

extern "C" int spr_radius_query(const float* d_queries, const int32_t* d_q_lengths, int n_queries, int n_clouds,
                                const void* d_grid_workspace, int n_supports, float radius, int limit, void* d_out_idx,
                                int idx_is_64, int row_stride, int32_t* d_out_max_count, void* stream_) {
  return spr_radius_query_ex(d_queries, d_q_lengths, n_queries, n_clouds, d_grid_workspace, n_supports, radius, limit,
                             SPR_ORDER_NEAREST, d_out_idx, idx_is_64, row_stride, d_out_max_count, stream_);
}
