// common.cu -- error state, launch counter, device-wide exclusive scan, cloud offsets.
#include "spr_common.cuh"

#include <cstring>

namespace spr {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

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

int exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, size_t n, int32_t* d_total, int32_t* d_tmp,
                       cudaStream_t stream) {
  if (n == 0) {
    if (d_total) SPR_CUDA(cudaMemsetAsync(d_total, 0, sizeof(int32_t), stream));
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
}
