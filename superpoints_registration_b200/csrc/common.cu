// common.cu -- error state, launch counter, device-wide exclusive scan, cloud offsets.
#include "spr_common.cuh"

#include <cstring>
#include <mutex>
#include <set>
#include <utility>

namespace spr {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

cudaError_t ensure_max_dynamic_smem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::set<std::pair<const void*, int>> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({func, dev})) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done.insert({func, dev});
  return e;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// exclusive scan: reduce tiles -> scan tile sums (one block) -> rescan tiles with carry-in
// ---------------------------------------------------------------------------------------------
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

size_t scan_tmp_ints(size_t n) { return (n + kScanTile - 1) / kScanTile + 8; }

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem_warp /*[32]*/, int& block_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (blockDim.x >> 5) ? smem_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(kFull, winc, o);
      if (lane >= o) winc += t;
    }
    smem_warp[lane] = winc - w;  // exclusive prefix of warp sums
    if (lane == 31) smem_warp[32] = winc;
  }
  __syncthreads();
  block_total = smem_warp[32];
  int r = inc - v + smem_warp[warp];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_reduce(const int32_t* __restrict__ in, size_t n,
                                                              int32_t* __restrict__ tile_sums) {
  __shared__ int sw[33];
  const size_t base = (size_t)blockIdx.x * kScanTile;
  int acc = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    size_t i = base + (size_t)j * kScanThreads + threadIdx.x;
    if (i < n) acc += in[i];
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = threadIdx.x < (kScanThreads >> 5) ? sw[threadIdx.x] : 0;
    v = warp_sum(v);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(1024) k_scan_tile_sums(int32_t* __restrict__ tile_sums, int n_tiles,
                                                         int32_t* __restrict__ total) {
  __shared__ int sw[33];
  int carry = 0;
  for (int base = 0; base < n_tiles; base += 1024) {
    int i = base + threadIdx.x;
    int v = i < n_tiles ? tile_sums[i] : 0;
    int bt;
    int ex = block_exclusive_scan(v, sw, bt);
    if (i < n_tiles) tile_sums[i] = ex + carry;
    carry += bt;
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

__global__ void __launch_bounds__(kScanThreads) k_scan_apply(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                             size_t n, const int32_t* __restrict__ tile_offs) {
  __shared__ int sw[33];
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  int v[kScanItems];
  int s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    size_t i = base + j;
    v[j] = i < n ? in[i] : 0;
    s += v[j];
  }
  int bt;
  int ex = block_exclusive_scan(s, sw, bt) + tile_offs[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    size_t i = base + j;
    if (i < n) out[i] = ex;
    ex += v[j];
  }
}

// Short inputs (the deeper pyramid levels, every level of a single pair): one block walks the array with a running
// carry -- one launch instead of three.  Integer sums: identical to the three-kernel scan.
constexpr size_t kScanSmall = 32768;
__global__ void __launch_bounds__(1024) k_scan_small(const int32_t* in, int32_t* out, size_t n, int32_t* total) {
  __shared__ int sw[33];
  int carry = 0;
  for (size_t base0 = 0; base0 < n; base0 += (size_t)1024 * kScanItems) {
    const size_t base = base0 + (size_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int s = 0;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      const size_t i = base + j;
      v[j] = i < n ? in[i] : 0;
      s += v[j];
    }
    int bt;
    int ex = block_exclusive_scan(s, sw, bt) + carry;
#pragma unroll
    for (int j = 0; j < kScanItems; ++j) {
      const size_t i = base + j;
      if (i < n) out[i] = ex;
      ex += v[j];
    }
    carry += bt;
  }
  if (threadIdx.x == 0 && total) *total = carry;
}

int exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, size_t n, int32_t* d_total, int32_t* d_tmp,
                       cudaStream_t stream) {
  if (n == 0) {
    if (d_total) SPR_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int32_t), stream));
    return SPR_OK;
  }
  if (n <= kScanSmall) {
    k_scan_small<<<1, 1024, 0, stream>>>(d_in, d_out, n, d_total);
    SPR_LAUNCH_CHECK("k_scan_small");
    return SPR_OK;
  }
  const int n_tiles = (int)((n + kScanTile - 1) / kScanTile);
  k_scan_reduce<<<n_tiles, kScanThreads, 0, stream>>>(d_in, n, d_tmp);
  SPR_LAUNCH_CHECK("k_scan_reduce");
  k_scan_tile_sums<<<1, 1024, 0, stream>>>(d_tmp, n_tiles, d_total);
  SPR_LAUNCH_CHECK("k_scan_tile_sums");
  k_scan_apply<<<n_tiles, kScanThreads, 0, stream>>>(d_in, d_out, n, d_tmp);
  SPR_LAUNCH_CHECK("k_scan_apply");
  return SPR_OK;
}

__global__ void k_cloud_offsets(const int32_t* __restrict__ lens, int B, int32_t* __restrict__ offs) {
  __shared__ int sw[33];
  int carry = 0;
  for (int base = 0; base < B; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = i < B ? lens[i] : 0;
    int bt;
    int ex = block_exclusive_scan(v, sw, bt);
    if (i < B) offs[i] = ex + carry;
    carry += bt;
  }
  if (threadIdx.x == 0) offs[B] = carry;
}

namespace {
__global__ void k_bbox_init(uint32_t* __restrict__ bb, int B) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * 3) {
    bb[i] = 0xffffffffu;  // min
    bb[B * 3 + i] = 0u;   // max
  }
}

constexpr int kBboxChunk = 256;  // points per warp

// A warp folds a contiguous chunk of points in registers and issues six atomics when the chunk lies in one cloud
// (all but 2 B of the chunks); a chunk that straddles a cloud boundary falls back to per-point atomics.
__global__ void __launch_bounds__(256) k_bbox(const float* __restrict__ pts, const int* __restrict__ offs, int B, int n,
                                              uint32_t* __restrict__ bb) {
  const int lane = threadIdx.x & 31;
  const int base = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kBboxChunk;
  if (base >= n) return;
  const int end = min(base + kBboxChunk, n);
  const int b0 = find_cloud(offs, B, base), b1 = find_cloud(offs, B, end - 1);
  if (b0 == b1) {
    uint32_t mnx = 0xffffffffu, mny = mnx, mnz = mnx, mxx = 0u, mxy = 0u, mxz = 0u;
    for (int i = base + lane; i < end; i += 32) {
      const uint32_t x = f2ord(pts[3 * (size_t)i]), y = f2ord(pts[3 * (size_t)i + 1]), z = f2ord(pts[3 * (size_t)i + 2]);
      mnx = min(mnx, x); mny = min(mny, y); mnz = min(mnz, z);
      mxx = max(mxx, x); mxy = max(mxy, y); mxz = max(mxz, z);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
      mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
      mnz = min(mnz, __shfl_xor_sync(0xffffffffu, mnz, o));
      mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
      mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
      mxz = max(mxz, __shfl_xor_sync(0xffffffffu, mxz, o));
    }
    if (lane == 0) {
      atomicMin(bb + 3 * b0 + 0, mnx);
      atomicMin(bb + 3 * b0 + 1, mny);
      atomicMin(bb + 3 * b0 + 2, mnz);
      atomicMax(bb + 3 * (B + b0) + 0, mxx);
      atomicMax(bb + 3 * (B + b0) + 1, mxy);
      atomicMax(bb + 3 * (B + b0) + 2, mxz);
    }
  } else {
    for (int i = base + lane; i < end; i += 32) {
      const int b = find_cloud(offs, B, i);
      const uint32_t x = f2ord(pts[3 * (size_t)i]), y = f2ord(pts[3 * (size_t)i + 1]), z = f2ord(pts[3 * (size_t)i + 2]);
      atomicMin(bb + 3 * b + 0, x);
      atomicMin(bb + 3 * b + 1, y);
      atomicMin(bb + 3 * b + 2, z);
      atomicMax(bb + 3 * (B + b) + 0, x);
      atomicMax(bb + 3 * (B + b) + 1, y);
      atomicMax(bb + 3 * (B + b) + 2, z);
    }
  }
}
}  // namespace

// bb[3 b + a] / bb[3 (B + b) + a] = ordered-int min / max of coordinate a over cloud b (f2ord / ord2f)
int cloud_bboxes(const float* d_pts, const int32_t* d_offs, int B, int n, uint32_t* d_bb, cudaStream_t stream) {
  k_bbox_init<<<(B * 3 + 255) / 256, 256, 0, stream>>>(d_bb, B);
  SPR_LAUNCH_CHECK("k_bbox_init");
  const int warps = (n + kBboxChunk - 1) / kBboxChunk;
  k_bbox<<<(warps + 7) / 8, 256, 0, stream>>>(d_pts, d_offs, B, n, d_bb);
  SPR_LAUNCH_CHECK("k_bbox");
  return SPR_OK;
}

int cloud_offsets(const int32_t* d_lens, int B, int32_t* d_offs, cudaStream_t stream) {
  k_cloud_offsets<<<1, 256, 0, stream>>>(d_lens, B, d_offs);
  SPR_LAUNCH_CHECK("k_cloud_offsets");
  return SPR_OK;
}

}  // namespace spr

extern "C" {
int spr_version(void) { return 100; }
const char* spr_last_error(void) { return spr::g_err; }
unsigned long long spr_launch_count(void) { return spr::g_launches.load(); }
unsigned int spr_numeric_flags(int reset) {
  return spr::gemm_numeric_flags(reset != 0) | spr::blocks_numeric_flags(reset != 0) |
         spr::attention_numeric_flags(reset != 0) | spr::attention_tc_numeric_flags(reset != 0);
}
}
