// Bring-up test of the two primitives the gather-by-TMA KPConv kernel rests on:
//   (1) cp.async.bulk.tensor.2d ... tile::gather4 : four rows of a [Ns, 2C] fp16 matrix, picked by index, land as four
//       consecutive 128-byte rows of a SWIZZLE_128B shared-memory tile (out-of-range rows are zero-filled);
//   (2) tcgen05.mma with an MN-MAJOR A operand: the gathered rows (K index = neighbour, 64 contiguous M elements per
//       row) are consumed as they lie, D[m][n] = sum_h X[idx[h]][col0 + m] * B[n][h], B K-major SWIZZLE_128B.
// D is read back from TMEM lanes 0..63 by the two warps that may address them and compared with a CPU reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_mn_test tools/umma_mn_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../superpoints_registration_b200/csrc/tc05.cuh"

using namespace spr::tc;

constexpr int KH = 48;      // neighbours (K of the product), 3 MMA K steps
constexpr int MM = 64;      // M = 32 channels x (hi, lo)
constexpr int NN = 32;      // N = 16 kernel points x (hi, lo)
constexpr int ROW_ELEMS = 128;  // C = 64: 128 fp16 per support row

__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* tmap, uint64_t* bar, int col, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

// SWIZZLE_128B descriptor with explicit leading / stride byte offsets (MN-major: LBO = stride between 64-element M blocks,
// SBO = stride between groups of 8 K rows)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(192) k_test(const __grid_constant__ CUtensorMap tmap, const int* __restrict__ idx,
                                              const __half* __restrict__ b, int col0, float* __restrict__ d,
                                              unsigned char* __restrict__ dump) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;              // 48 rows x 128 B (6 KB), written by the TMA engine
  unsigned char* sB = smem + 6144;       // 32 rows x 128 B (4 KB), K-major SWIZZLE_128B, written by threads
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&full, 1);
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&tmem_base, 32);
  // B: thread t < 32 writes row t (48 values + 16 zero)
  if (tid < NN) {
    for (int j = 0; j < 8; ++j) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (j < 6) v = *reinterpret_cast<const uint4*>(b + (size_t)tid * KH + j * 8);
      *reinterpret_cast<uint4*>(sB + sw128_offset(tid, j)) = v;
      *reinterpret_cast<uint4*>(sB + 4096 + sw128_offset(NN - 1 - tid, j)) = v;   // sB2: rows reversed
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 5 && lane == 0) {  // gather issue
    mbar_arrive_expect_tx(&full, KH * 128);
    for (int i = 0; i < KH / 4; ++i)
      tma_gather4(sA + i * 512, &tmap, &full, col0, idx[4 * i], idx[4 * i + 1], idx[4 * i + 2], idx[4 * i + 3]);
  } else if (warp == 4 && lane == 0) {  // MMA issuer
    constexpr uint32_t idesc = idesc_f16(MM, NN, 1, 0);
    mbar_wait(&full, 0);
    tc_fence_after();
    for (int ks = 0; ks < KH / 16; ++ks) {
      const uint64_t ad = desc_sw128(smem_u32(sA) + ks * 2048, 6144, 1024);
      const uint64_t bd = desc_sw128_kmajor(smem_u32(sB) + ks * 32);
      umma_f16(tb, ad, bd, idesc, ks != 0);
    }
    // a second M = 64 accumulator in the SAME columns at lane offset 16 (the other half of every lane quadrant):
    // same A, B rows in reverse order (sB2)
    for (int ks = 0; ks < KH / 16; ++ks) {
      const uint64_t ad = desc_sw128(smem_u32(sA) + ks * 2048, 6144, 1024);
      const uint64_t bd = desc_sw128_kmajor(smem_u32(sB + 4096) + ks * 32);
      umma_f16(tb + (16u << 16), ad, bd, idesc, ks != 0);
    }
    umma_commit(&done);
  }
  if (warp < 4) {  // every TMEM lane quadrant is dumped: the host works out which lanes hold which rows of D
    mbar_wait(&done, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < NN; c0 += 8) {
      float v[8];
      tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 8; ++i) d[(size_t)(warp * 32 + lane) * NN + c0 + i] = v[i];
    }
    tc_fence_before();
  }
  __syncthreads();
  for (int i = tid; i < 6144; i += blockDim.x) dump[i] = sA[i];
  if (warp == 4) tmem_dealloc(tb, 32);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int NS = 500;
  std::vector<__half> hx((size_t)NS * ROW_ELEMS), hb((size_t)NN * KH);
  srand(3);
  for (auto& v : hx) v = __float2half((rand() % 2001 - 1000) / 500.f);
  for (auto& v : hb) v = __float2half((rand() % 2001 - 1000) / 900.f);
  std::vector<int> idx(KH);
  for (int h = 0; h < KH; ++h) idx[h] = h < 40 ? rand() % NS : NS;   // rows 40..47: padding -> out of range
  idx[7] = NS;                                                        // a shadow entry in the middle
  idx[13] = NS + 5;
  const int col0 = 64;  // second 32-channel pass

  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || !fn) {
    printf("cuTensorMapEncodeTiled not available: %s\n", cudaGetErrorString(e));
    return 1;
  }
  __half* dx;
  cudaMalloc(&dx, hx.size() * 2);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmap;
  const cuuint64_t dims[2] = {ROW_ELEMS, (cuuint64_t)NS};
  const cuuint64_t strides[1] = {ROW_ELEMS * 2};
  const cuuint32_t box[2] = {64, 1};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dx, dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    printf("cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return 1;
  }
  int* didx;
  __half* db;
  float* dd;
  unsigned char* ddump;
  cudaMalloc(&didx, KH * 4);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dd, 128 * NN * 4);
  cudaMalloc(&ddump, 6144);
  cudaMemcpy(didx, idx.data(), KH * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dd, 0, 128 * NN * 4);
  const int smem = 6144 + 8192 + 1024;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_test<<<1, 192, smem>>>(tmap, didx, db, col0, dd, ddump);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> hd(128 * NN);
  std::vector<unsigned char> dump(6144);
  cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(dump.data(), ddump, 6144, cudaMemcpyDeviceToHost);
  // (1) the gathered tile: row h, 16-byte chunk j at sw128_offset(h, j)
  int bad_rows = 0;
  for (int h = 0; h < KH; ++h)
    for (int j = 0; j < 8; ++j) {
      unsigned char want[16];
      memset(want, 0, 16);
      if (idx[h] < NS) memcpy(want, &hx[(size_t)idx[h] * ROW_ELEMS + col0 + j * 8], 16);
      if (memcmp(want, &dump[sw128_offset(h, j)], 16) != 0) {
        if (bad_rows < 5) printf("gather mismatch at row %d chunk %d (idx %d)\n", h, j, idx[h]);
        ++bad_rows;
      }
    }
  printf("gather4: %s (%d mismatching chunks of %d)\n", bad_rows ? "MISMATCH" : "OK", bad_rows, KH * 8);
  // (2) the product: reference rows, then for every row the TMEM lane that holds it
  std::vector<double> ref(MM * NN);
  double maxref = 0;
  for (int m = 0; m < MM; ++m)
    for (int n = 0; n < NN; ++n) {
      double acc = 0;
      for (int h = 0; h < KH; ++h)
        if (idx[h] < NS)
          acc += (double)__half2float(hx[(size_t)idx[h] * ROW_ELEMS + col0 + m]) * __half2float(hb[(size_t)n * KH + h]);
      ref[m * NN + n] = acc;
      maxref = fmax(maxref, fabs(acc));
    }
  double maxerr = 0;
  bool identity = true;
  printf("row -> TMEM lane:");
  for (int m = 0; m < MM; ++m) {
    int best = -1;
    double best_err = 1e30;
    for (int l = 0; l < 128; ++l) {
      double err = 0;
      for (int n = 0; n < NN; ++n) err = fmax(err, fabs(ref[m * NN + n] - hd[l * NN + n]));
      if (err < best_err) {
        best_err = err;
        best = l;
      }
    }
    if (m % 16 == 0) printf(" [%d]=%d", m, best);
    identity = identity && best == m;
    maxerr = fmax(maxerr, best_err);
  }
  printf("  (%s)\n", identity ? "lane = row" : "NOT the identity: see the map");
  // the second accumulator: D2[m][n] = D[m][NN-1-n], expected at lane 16 + (m % 16) + 32 (m / 16)
  double maxerr2 = 0;
  for (int m = 0; m < MM; ++m) {
    const int l = 16 + (m % 16) + 32 * (m / 16);
    for (int n = 0; n < NN; ++n) maxerr2 = fmax(maxerr2, fabs(ref[m * NN + (NN - 1 - n)] - hd[l * NN + n]));
  }
  printf("second accumulator at lane offset 16: max abs err %.3e -> %s\n", maxerr2, maxerr2 < 1e-4 * maxref ? "OK" : "MISMATCH");
  double maxerr1 = 0;
  for (int m = 0; m < MM; ++m) {
    const int l = (m % 16) + 32 * (m / 16);
    for (int n = 0; n < NN; ++n) maxerr1 = fmax(maxerr1, fabs(ref[m * NN + n] - hd[l * NN + n]));
  }
  printf("first accumulator at lanes (m %% 16) + 32 (m / 16): max abs err %.3e -> %s\n", maxerr1, maxerr1 < 1e-4 * maxref ? "OK" : "MISMATCH");
  const bool ok = maxerr < 1e-4 * maxref;
  printf("MN-major UMMA M=%d N=%d K=%d: max abs err %.3e (max |ref| %.3e) -> %s\n", MM, NN, KH, maxerr, maxref,
         ok ? "OK" : "MISMATCH");
  printf((ok && !bad_rows) ? "ALL OK\n" : "FAILED\n");
  return (ok && !bad_rows) ? 0 : 1;
}
