// Bring-up test of the MN-major SWIZZLE_64B A operand used by the gather KPConv kernel (csrc/kpconv_g.cu):
// the A tile holds 96 rows of 64 bytes -- rows 0..47 the fp16 hi halves of 32 channels of 48 neighbours, rows 48..95 the lo
// halves -- so that  D[c][n] = sum_h (X_hi[h][c] + X_lo[h][c]) * B[n][h]  is accumulated by two chains of MMAs over the
// same B operand (K-major SWIZZLE_128B).  M = 64 with only 32 useful rows: the leading byte offset of the descriptor is 0,
// rows 32..63 of D repeat rows 0..31.  D is dumped from all four TMEM lane quadrants and compared with a CPU reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_sw64_test tools/umma_sw64_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../superpoints_registration_b200/csrc/tc05.cuh"

using namespace spr::tc;

constexpr int KH = 48, NN = 32, MC = 32;

// byte offset of 16-byte chunk j (0..3) of row r inside a (rows x 64 B) SWIZZLE_64B tile: 8 rows per 512-byte group
__host__ __device__ constexpr uint32_t sw64_offset(uint32_t r, uint32_t j) {
  return (r >> 3) * 512u + (r & 7u) * 64u + ((j ^ ((r >> 1) & 3u)) << 4);
}
__device__ __forceinline__ uint64_t desc_sw64_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n, int a_mn_major) {
  return (1u << 4) | ((uint32_t)a_mn_major << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__global__ void __launch_bounds__(192) k_test(const __half* __restrict__ xh, const __half* __restrict__ xl,
                                              const __half* __restrict__ b, float* __restrict__ d, int lbo) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;          // 96 rows x 64 B = 6 KB
  unsigned char* sB = smem + 6144;   // 32 rows x 128 B
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&tmem_base, 32);
  for (int i = tid; i < 96 * 4; i += blockDim.x) {
    const int r = i >> 2, j = i & 3;
    const __half* src = r < KH ? xh + (size_t)r * MC : xl + (size_t)(r - KH) * MC;
    *reinterpret_cast<uint4*>(sA + sw64_offset(r, j)) = *reinterpret_cast<const uint4*>(src + j * 8);
  }
  if (tid < NN)
    for (int j = 0; j < 8; ++j) {
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (j < 6) v = *reinterpret_cast<const uint4*>(b + (size_t)tid * KH + j * 8);
      *reinterpret_cast<uint4*>(sB + sw128_offset(tid, j)) = v;
    }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 4 && lane == 0) {
    constexpr uint32_t idesc = idesc_f16(64, NN, 1);
    for (int part = 0; part < 2; ++part)
      for (int ks = 0; ks < KH / 16; ++ks) {
        const uint64_t ad = desc_sw64_mnmajor(smem_u32(sA) + part * (KH * 64) + ks * 1024, lbo);
        const uint64_t bd = desc_sw128_kmajor(smem_u32(sB) + ks * 32);
        umma_f16(tb, ad, bd, idesc, (part | ks) != 0);
      }
    umma_commit(&done);
  }
  if (warp < 4) {
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
  if (warp == 4) tmem_dealloc(tb, 32);
}

int main() {
  std::vector<__half> xh((size_t)KH * MC), xl((size_t)KH * MC), hb((size_t)NN * KH);
  srand(5);
  for (auto& v : xh) v = __float2half((rand() % 2001 - 1000) / 500.f);
  for (auto& v : xl) v = __float2half((rand() % 2001 - 1000) / 500000.f);
  for (auto& v : hb) v = __float2half((rand() % 2001 - 1000) / 900.f);
  __half *dxh, *dxl, *db;
  float* dd;
  cudaMalloc(&dxh, xh.size() * 2);
  cudaMalloc(&dxl, xl.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dd, 128 * NN * 4);
  cudaMemcpy(dxh, xh.data(), xh.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dxl, xl.data(), xl.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 6144 + 4096 + 1024;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int bad = 0;
  for (int lbo : {0, 16, 512}) {
    cudaMemset(dd, 0, 128 * NN * 4);
    k_test<<<1, 192, smem>>>(dxh, dxl, db, dd, lbo);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("lbo=%d: CUDA error %s\n", lbo, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> hd(128 * NN);
    cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int c = 0; c < MC; ++c) {
      const int l = (c % 16) + 32 * (c / 16);
      for (int n = 0; n < NN; ++n) {
        double acc = 0;
        for (int h = 0; h < KH; ++h)
          acc += ((double)__half2float(xh[(size_t)h * MC + c]) + __half2float(xl[(size_t)h * MC + c])) * __half2float(hb[(size_t)n * KH + h]);
        maxerr = fmax(maxerr, fabs(acc - hd[l * NN + n]));
        maxref = fmax(maxref, fabs(acc));
      }
    }
    const bool ok = maxerr < 1e-4 * maxref;
    printf("SWIZZLE_64B MN-major A, LBO=%d: rows 0..31 at lanes (c %% 16) + 32 (c / 16): max abs err %.3e (max |ref| %.3e) -> %s\n", lbo,
           maxerr, maxref, ok ? "OK" : "MISMATCH");
    bad += !ok;
  }
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
