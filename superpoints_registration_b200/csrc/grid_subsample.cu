// grid_subsample.cu -- barycentre voxel-grid subsampling of stacked clouds on sm_100a.
//
// Replaces batch_grid_subsampling (reference: models/backbone_kpconv/cpp_wrappers/cpp_subsampling/
// grid_subsampling/grid_subsampling.cpp:109-211 -> grid_subsampling :5-106).  Algorithm here:
//   1. per-cloud bounding box (ordered-int atomics)                          k_bbox
//   2. voxel key of every point with the reference's fp32 recipe, hashed into an open-addressing table
//      whose slots are claimed by a representative point (atomicCAS); each slot tracks the lowest
//      member index and the member count                                     k_key_insert
//   3. a point is a voxel "head" iff it is the lowest index of its slot; an exclusive scan of the head
//      flags numbers the voxels in first-occurrence order                    k_heads + scan
//   4. members are binned per voxel (scan of counts + atomic cursor)         k_vox_counts + scan + k_scatter
//   5. one thread per voxel walks ITS members in ascending input index and accumulates x,y,z
//      sequentially in fp32 -- the reference's summation order (grid_subsampling.h:74-78) -- then scales
//      by (float)(1.0/count) (grid_subsampling.cpp:87)                        k_barycentre
// Steps 4-5 are the order-preserving form of a segmented reduction: a tree-shaped warp reduction would
// change the last ulp of the barycentre and with it the radius tests of every later level.
#include "spr_common.cuh"

namespace spr {
namespace {

constexpr int kThreads = 256;

struct CloudGrid {  // per cloud, derived from the bounding box exactly as grid_subsampling.cpp:25-31
  float ox, oy, oz;
  unsigned long long nx, ny;
};

__device__ __forceinline__ unsigned long long f2u64(float v) {
  // (size_t)floor(v) as compiled for x86-64: through a signed 64-bit conversion (negative wraps).
  return (unsigned long long)(long long)v;
}


// mode SPR_SUBSAMPLE_REFERENCE: the CPU reference's per-cloud origin and (p - origin) / dl voxels.
// other modes (MinkowskiEngine-style mean, first point): voxel = floor(p / dl) on a global lattice; (ox, oy, oz) then
// hold the cloud's lowest voxel coordinate (an integer-valued float) so that the local key stays non-negative.
__global__ void k_cloud_grid(const uint32_t* __restrict__ bb, int B, float dl, int mode, CloudGrid* __restrict__ g) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  if (mode != SPR_SUBSAMPLE_REFERENCE) {
    CloudGrid cg;
    float lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = floorf(__fdiv_rn(ord2f(bb[3 * b + a]), dl));
      hi[a] = floorf(__fdiv_rn(ord2f(bb[3 * (B + b) + a]), dl));
    }
    cg.ox = lo[0];
    cg.oy = lo[1];
    cg.oz = lo[2];
    cg.nx = f2u64(__fsub_rn(hi[0], lo[0])) + 1ull;
    cg.ny = f2u64(__fsub_rn(hi[1], lo[1])) + 1ull;
    g[b] = cg;
    return;
  }
  // grid_subsampling.cpp:27  originCorner = floor(minCorner * (1/sampleDl)) * sampleDl   (all fp32)
  const float inv = __fdiv_rn(1.0f, dl);
  float o[3], mx[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float mn = ord2f(bb[3 * b + a]);
    mx[a] = ord2f(bb[3 * (B + b) + a]);
    o[a] = __fmul_rn(floorf(__fmul_rn(mn, inv)), dl);
  }
  CloudGrid cg;
  cg.ox = o[0];
  cg.oy = o[1];
  cg.oz = o[2];
  // :30-31
  cg.nx = f2u64(floorf(__fdiv_rn(__fsub_rn(mx[0], o[0]), dl))) + 1ull;
  cg.ny = f2u64(floorf(__fdiv_rn(__fsub_rn(mx[1], o[1]), dl))) + 1ull;
  g[b] = cg;
}

__device__ __forceinline__ uint32_t hash_mix(unsigned long long key, int cloud) {
  unsigned long long h = key * 0x9E3779B97F4A7C15ull + (unsigned long long)(cloud + 1) * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  h *= 0xBF58476D1CE4E5B9ull;
  h ^= h >> 32;
  return (uint32_t)h;
}

// slot_rep: -1 = empty, otherwise the index of the point that claimed the slot.
__global__ void __launch_bounds__(kThreads)
    k_key_insert(const float* __restrict__ pts, const int* __restrict__ offs, int B, int n, float dl, int mode,
                 const CloudGrid* __restrict__ grids, unsigned long long* __restrict__ keys, int* __restrict__ slot_rep,
                 int* __restrict__ slot_min, int* __restrict__ slot_cnt, int* __restrict__ point_slot,
                 uint32_t table_mask) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = find_cloud(offs, B, i);
  const CloudGrid g = grids[b];
  unsigned long long ix, iy, iz;
  if (mode == SPR_SUBSAMPLE_REFERENCE) {
    // grid_subsampling.cpp:53-56 (fp32 subtract, fp32 divide, floor)
    ix = f2u64(floorf(__fdiv_rn(__fsub_rn(pts[3 * (size_t)i + 0], g.ox), dl)));
    iy = f2u64(floorf(__fdiv_rn(__fsub_rn(pts[3 * (size_t)i + 1], g.oy), dl)));
    iz = f2u64(floorf(__fdiv_rn(__fsub_rn(pts[3 * (size_t)i + 2], g.oz), dl)));
  } else {
    // floor(p / dl) on the global lattice, relative to the cloud's lowest voxel (integer-valued floats: exact)
    ix = f2u64(__fsub_rn(floorf(__fdiv_rn(pts[3 * (size_t)i + 0], dl)), g.ox));
    iy = f2u64(__fsub_rn(floorf(__fdiv_rn(pts[3 * (size_t)i + 1], dl)), g.oy));
    iz = f2u64(__fsub_rn(floorf(__fdiv_rn(pts[3 * (size_t)i + 2], dl)), g.oz));
  }
  const unsigned long long key = ix + g.nx * iy + g.nx * g.ny * iz;
  keys[i] = key;
  __threadfence();  // the key must be visible before this point can become a slot representative
  uint32_t h = hash_mix(key, b) & table_mask;
  const int lo = offs[b], hi = offs[b + 1];
  for (;;) {
    int rep = atomicCAS(slot_rep + h, -1, i);
    if (rep == -1) rep = i;
    bool same = false;
    if (rep == i) {
      same = true;
    } else if (rep >= lo && rep < hi) {
      // The representative published its key before claiming the slot (fence above); read it through L2.
      unsigned long long rk = *((volatile unsigned long long*)(keys + rep));
      same = (rk == key);
    }
    if (same) break;
    h = (h + 1) & table_mask;
  }
  atomicMin(slot_min + h, i);
  atomicAdd(slot_cnt + h, 1);
  point_slot[i] = (int)h;
}

__global__ void __launch_bounds__(kThreads) k_heads(const int* __restrict__ point_slot, const int* __restrict__ slot_min,
                                                    int n, int* __restrict__ head_flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) head_flag[i] = (slot_min[point_slot[i]] == i) ? 1 : 0;
}

// For head points: voxel id = scan[i]; publish it on the slot, emit the voxel's member count, and count the
// voxel for its cloud.
__global__ void __launch_bounds__(kThreads)
    k_vox_counts(const int* __restrict__ point_slot, const int* __restrict__ slot_min, const int* __restrict__ slot_cnt,
                 const int* __restrict__ head_scan, int n, int* __restrict__ slot_vox, int* __restrict__ vox_cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = point_slot[i];
  if (slot_min[s] == i) {
    const int v = head_scan[i];
    slot_vox[s] = v;
    vox_cnt[v] = slot_cnt[s];
  }
}

__global__ void k_out_lengths(const int* __restrict__ head_scan, const int* __restrict__ offs, int B, int n,
                              const int* __restrict__ total, int* __restrict__ out_lens) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int lo = offs[b], hi = offs[b + 1];
  const int s_lo = lo < n ? head_scan[lo] : *total;
  const int s_hi = hi < n ? head_scan[hi] : *total;
  out_lens[b] = s_hi - s_lo;
}

__global__ void __launch_bounds__(kThreads)
    k_scatter_members(const int* __restrict__ point_slot, const int* __restrict__ slot_vox,
                      const int* __restrict__ vox_off, int* __restrict__ vox_fill, int n, int* __restrict__ members) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int v = slot_vox[point_slot[i]];
  const int pos = vox_off[v] + atomicAdd(vox_fill + v, 1);
  members[pos] = i;
}

__global__ void __launch_bounds__(kThreads)
    k_barycentre(const float* __restrict__ pts, const int* __restrict__ members, const int* __restrict__ vox_off,
                 const int* __restrict__ vox_cnt, const int* __restrict__ total, int mode, float* __restrict__ out) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= *total) return;
  const int beg = vox_off[v], cnt = vox_cnt[v];
  const int* __restrict__ m = members + beg;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  if (cnt <= 16) {
    // small voxel: pull the member list into registers, sort it (ascending input index), then add
    int idx[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) idx[j] = j < cnt ? m[j] : 0x7fffffff;
#pragma unroll
    for (int a = 1; a < 16; ++a) {  // insertion sort network on a fixed-size array (stays in registers)
#pragma unroll
      for (int c = a; c > 0; --c) {
        int lo = min(idx[c - 1], idx[c]), hi = max(idx[c - 1], idx[c]);
        idx[c - 1] = lo;
        idx[c] = hi;
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < cnt) {
        const float* p = pts + 3 * (size_t)idx[j];
        sx = __fadd_rn(sx, p[0]);
        sy = __fadd_rn(sy, p[1]);
        sz = __fadd_rn(sz, p[2]);
      }
    }
  } else {
    // large voxel: repeated selection of the next-lowest member index (no extra memory)
    int last = -1;
    for (int t = 0; t < cnt; ++t) {
      int best = 0x7fffffff;
      for (int j = 0; j < cnt; ++j) {
        int c = m[j];
        if (c > last && c < best) best = c;
      }
      const float* p = pts + 3 * (size_t)best;
      sx = __fadd_rn(sx, p[0]);
      sy = __fadd_rn(sy, p[1]);
      sz = __fadd_rn(sz, p[2]);
      last = best;
    }
  }
  if (mode == SPR_SUBSAMPLE_REFERENCE) {
    // grid_subsampling.cpp:87 -- the double 1.0/count is narrowed to float by operator*(PointXYZ, float)
    const float w = (float)(1.0 / (double)cnt);
    out[3 * (size_t)v + 0] = __fmul_rn(sx, w);
    out[3 * (size_t)v + 1] = __fmul_rn(sy, w);
    out[3 * (size_t)v + 2] = __fmul_rn(sz, w);
  } else {  // unweighted average: sum / count
    const float c = (float)cnt;
    out[3 * (size_t)v + 0] = __fdiv_rn(sx, c);
    out[3 * (size_t)v + 1] = __fdiv_rn(sy, c);
    out[3 * (size_t)v + 2] = __fdiv_rn(sz, c);
  }
}

// first-point mode: a voxel is represented by its lowest-index member (its head), copied unchanged
__global__ void __launch_bounds__(kThreads)
    k_first_point(const float* __restrict__ pts, const int* __restrict__ point_slot, const int* __restrict__ slot_min,
                  const int* __restrict__ head_scan, int n, float* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || slot_min[point_slot[i]] != i) return;
  const int v = head_scan[i];
  out[3 * (size_t)v + 0] = pts[3 * (size_t)i + 0];
  out[3 * (size_t)v + 1] = pts[3 * (size_t)i + 1];
  out[3 * (size_t)v + 2] = pts[3 * (size_t)i + 2];
}

uint32_t table_size_for(int n) {
  uint32_t t = 1024;
  while (t < 2u * (uint32_t)n) t <<= 1;
  return t;
}

}  // namespace
}  // namespace spr

using namespace spr;

extern "C" size_t spr_grid_subsample_workspace_bytes(int n_points, int n_clouds) {
  if (n_points < 0 || n_clouds < 0) return 0;
  const size_t n = (size_t)n_points, B = (size_t)n_clouds, T = table_size_for(n_points);
  size_t bytes = 0;
  auto add = [&](size_t b) { bytes = align_up(bytes, 256) + b; };
  add((B + 1) * 4);            // offs
  add(B * 6 * 4);              // bbox
  add(B * sizeof(CloudGrid));  // grids
  add(n * 8);                  // keys
  add(T * 4 * 3);              // slot_rep, slot_min, slot_cnt
  add(T * 4);                  // slot_vox
  add(n * 4);                  // point_slot
  add((n + 1) * 4);            // head flags / scan
  add(n * 4 * 3);              // vox_cnt, vox_off, vox_fill
  add(n * 4);                  // members
  add(scan_tmp_ints(n + 1) * 4);
  add(16);
  return bytes + 1024;
}

extern "C" int spr_grid_subsample_batch(const float* d_points, const int32_t* d_lengths, int n_points, int n_clouds,
                                        float sample_dl, float* d_out_points, int32_t* d_out_lengths,
                                        int32_t* d_out_total, void* d_workspace, size_t workspace_bytes, void* stream_) {
  return spr_grid_subsample_batch_ex(d_points, d_lengths, n_points, n_clouds, sample_dl, SPR_SUBSAMPLE_REFERENCE,
                                     d_out_points, d_out_lengths, d_out_total, d_workspace, workspace_bytes, stream_);
}

extern "C" int spr_grid_subsample_batch_ex(const float* d_points, const int32_t* d_lengths, int n_points, int n_clouds,
                                           float sample_dl, int mode, float* d_out_points, int32_t* d_out_lengths,
                                           int32_t* d_out_total, void* d_workspace, size_t workspace_bytes,
                                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SPR_CHECK_ARG(mode == SPR_SUBSAMPLE_REFERENCE || mode == SPR_SUBSAMPLE_MEAN || mode == SPR_SUBSAMPLE_FIRST,
                "grid_subsample: unknown mode %d", mode);
  SPR_CHECK_ARG(n_points > 0 && n_clouds > 0, "grid_subsample: empty input (n_points=%d, n_clouds=%d)", n_points,
                n_clouds);
  SPR_CHECK_ARG(sample_dl > 0.f, "grid_subsample: sample_dl must be > 0");
  SPR_CHECK_ARG(d_points && d_lengths && d_out_points && d_out_lengths && d_out_total && d_workspace,
                "grid_subsample: null pointer");
  if (workspace_bytes < spr_grid_subsample_workspace_bytes(n_points, n_clouds)) {
    set_error("grid_subsample: workspace too small");
    return SPR_ENOSPACE;
  }
  const int n = n_points, B = n_clouds;
  const uint32_t T = table_size_for(n);
  Carver ws(d_workspace, workspace_bytes);
  int* offs = ws.take<int>(B + 1);
  uint32_t* bb = ws.take<uint32_t>((size_t)B * 6);
  CloudGrid* grids = ws.take<CloudGrid>(B);
  unsigned long long* keys = ws.take<unsigned long long>(n);
  int* slot_rep = ws.take<int>((size_t)T * 3);
  int* slot_min = slot_rep + T;
  int* slot_cnt = slot_min + T;
  int* slot_vox = ws.take<int>(T);
  int* point_slot = ws.take<int>(n);
  int* head = ws.take<int>((size_t)n + 1);
  int* vox_cnt = ws.take<int>((size_t)n * 3);
  int* vox_off = vox_cnt + n;
  int* vox_fill = vox_off + n;
  int* members = ws.take<int>(n);
  int* scan_tmp = ws.take<int>(scan_tmp_ints((size_t)n + 1));
  if (!ws.ok()) {
    set_error("grid_subsample: workspace carve overflow");
    return SPR_ENOSPACE;
  }
  const int gp = (n + kThreads - 1) / kThreads;

  int rc = cloud_offsets(d_lengths, B, offs, stream);
  if (rc) return rc;
  rc = cloud_bboxes(d_points, offs, B, n, bb, stream);
  if (rc) return rc;
  k_cloud_grid<<<(B + 127) / 128, 128, 0, stream>>>(bb, B, sample_dl, mode, grids);
  SPR_LAUNCH_CHECK("k_cloud_grid");
  SPR_CUDA(cudaMemsetAsync(slot_rep, 0xff, (size_t)T * 4, stream));  // -1
  SPR_CUDA(cudaMemsetAsync(slot_min, 0x7f, (size_t)T * 4, stream));  // 0x7f7f7f7f > any index
  SPR_CUDA(cudaMemsetAsync(slot_cnt, 0, (size_t)T * 4, stream));
  SPR_CUDA(cudaMemsetAsync(vox_cnt, 0, (size_t)n * 4 * 3, stream));  // vox_cnt, vox_off, vox_fill
  k_key_insert<<<gp, kThreads, 0, stream>>>(d_points, offs, B, n, sample_dl, mode, grids, keys, slot_rep, slot_min, slot_cnt,
                                            point_slot, T - 1);
  SPR_LAUNCH_CHECK("k_key_insert");
  k_heads<<<gp, kThreads, 0, stream>>>(point_slot, slot_min, n, head);
  SPR_LAUNCH_CHECK("k_heads");
  rc = exclusive_scan_i32(head, head, (size_t)n, d_out_total, scan_tmp, stream);
  if (rc) return rc;
  k_vox_counts<<<gp, kThreads, 0, stream>>>(point_slot, slot_min, slot_cnt, head, n, slot_vox, vox_cnt);
  SPR_LAUNCH_CHECK("k_vox_counts");
  k_out_lengths<<<(B + 127) / 128, 128, 0, stream>>>(head, offs, B, n, d_out_total, d_out_lengths);
  SPR_LAUNCH_CHECK("k_out_lengths");
  if (mode == SPR_SUBSAMPLE_FIRST) {
    k_first_point<<<gp, kThreads, 0, stream>>>(d_points, point_slot, slot_min, head, n, d_out_points);
    SPR_LAUNCH_CHECK("k_first_point");
    return SPR_OK;
  }
  rc = exclusive_scan_i32(vox_cnt, vox_off, (size_t)n, nullptr, scan_tmp, stream);  // entries >= M are zero
  if (rc) return rc;
  k_scatter_members<<<gp, kThreads, 0, stream>>>(point_slot, slot_vox, vox_off, vox_fill, n, members);
  SPR_LAUNCH_CHECK("k_scatter_members");
  k_barycentre<<<gp, kThreads, 0, stream>>>(d_points, members, vox_off, vox_cnt, d_out_total, mode, d_out_points);
  SPR_LAUNCH_CHECK("k_barycentre");
  return SPR_OK;
}

