// Issue / completion rate of small tcgen05.mma instructions (kind::f16, operands in shared memory), one issuing thread
// on one SM: what does a chain of M x N x 16 products cost when N is small, when M = 64, when A is MN-major, and when
// consecutive instructions accumulate into the SAME TMEM columns?  (Sizing data for csrc/kpconv_g.cu.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_rate_test tools/umma_rate_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../superpoints_registration_b200/csrc/tc05.cuh"

using namespace spr::tc;

__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, int sw128) {
  // MN-major, leading byte offset 0 (the M atoms repeat), stride between 8-row K groups = 512 B (SW64) / 1024 B (SW128)
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((sw128 ? 1024 : 512) >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)(sw128 ? 2 : 4) << 61);
}

// a_kind: 0 = K-major SWIZZLE_128B, 1 = MN-major SWIZZLE_64B, 2 = MN-major SWIZZLE_128B
template <int N_ACC>
__global__ void __launch_bounds__(64) k_rate(int m, int n, int a_kind, int n_mma, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 0 && lane == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(a_kind != 0) << 15) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 64 * 1024;
    // descriptors of the 4 K steps are prepared; the loop body is 16 fully unrolled instructions
    uint64_t ad[4], bd[4];
    for (int ks = 0; ks < 4; ++ks) {
      ad[ks] = a_kind == 0 ? desc_sw128_kmajor(a0 + ks * 32) : desc_mn(a0 + ks * (a_kind == 2 ? 2048 : 1024), a_kind == 2);
      bd[ks] = desc_sw128_kmajor(b0 + ks * 32);
    }
    uint32_t dcol[N_ACC];
    for (int a = 0; a < N_ACC; ++a) dcol[a] = tb + a * n;
    const long long t0 = clock64();
    for (int it = 0; it < n_mma / 16; ++it) {
#pragma unroll
      for (int u = 0; u < 16; ++u) umma_f16(dcol[u % N_ACC], ad[(u / N_ACC) % 4], bd[(u / N_ACC) % 4], idesc, true);
    }
    const long long t1 = clock64();
    umma_commit(&done);
    mbar_wait(&done, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = 97 * 1024;
  struct Case { int m, n, a_kind, n_acc, k_steps; };
  cudaFuncSetAttribute(k_rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_rate<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_rate<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k_rate<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const Case cases[] = {
      {128, 32, 0, 1, 4}, {128, 32, 0, 2, 4}, {128, 32, 0, 4, 4}, {128, 32, 0, 8, 4},
      {64, 32, 0, 1, 4},  {64, 32, 0, 4, 4},
      {64, 32, 1, 1, 3},  {64, 32, 1, 2, 3},  {64, 32, 1, 4, 3},  {64, 32, 1, 8, 3},
      {128, 32, 1, 1, 3}, {128, 32, 1, 4, 3}, {128, 32, 1, 8, 3},
      {64, 32, 2, 1, 3},  {64, 32, 2, 4, 3},  {128, 32, 2, 1, 3}, {128, 32, 2, 4, 3},
      {128, 64, 0, 1, 4}, {128, 64, 0, 4, 4}, {128, 128, 0, 1, 4}, {128, 128, 0, 2, 4}, {128, 256, 0, 1, 4},
      {64, 64, 1, 1, 3},  {64, 64, 1, 4, 3},  {64, 128, 1, 1, 3}, {64, 128, 1, 2, 3}, {64, 256, 1, 1, 3},
      {128, 64, 1, 1, 3}, {128, 64, 1, 4, 3}, {128, 128, 1, 1, 3}, {128, 256, 1, 1, 3},
  };
  const char* kinds[] = {"K-major SW128", "MN-major SW64 ", "MN-major SW128"};
  printf("  M    N  A layout        accumulators  clk/MMA issue  clk/MMA complete\n");
  for (const Case& c : cases) {
    const int n_mma = 1600;
    long long best[2] = {1LL << 60, 1LL << 60};
    for (int rep = 0; rep < 3; ++rep) {
      switch (c.n_acc) {
        case 1: k_rate<1><<<1, 64, smem>>>(c.m, c.n, c.a_kind, n_mma, d); break;
        case 2: k_rate<2><<<1, 64, smem>>>(c.m, c.n, c.a_kind, n_mma, d); break;
        case 4: k_rate<4><<<1, 64, smem>>>(c.m, c.n, c.a_kind, n_mma, d); break;
        default: k_rate<8><<<1, 64, smem>>>(c.m, c.n, c.a_kind, n_mma, d); break;
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("M=%d N=%d kind=%d: CUDA error %s\n", c.m, c.n, c.a_kind, cudaGetErrorString(e));
        return 1;
      }
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      if (h[1] < best[1]) best[0] = h[0], best[1] = h[1];
    }
    printf("%4d %4d  %s  %12d  %13.1f  %16.1f\n", c.m, c.n, kinds[c.a_kind], c.n_acc, (double)best[0] / n_mma,
           (double)best[1] / n_mma);
  }
  return 0;
}
