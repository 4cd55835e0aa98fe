// Bring-up test of the tcgen05 plumbing in csrc/tc05.cuh: D[128 x N] = A[128 x K] * B[N x K]^T with fp16 operands,
// fp32 accumulation in TMEM.  A is written into the swizzled shared-memory layout by ordinary threads (as the
// KPConv producer does), B arrives as pre-swizzled stage images through bulk async copies (as the KPConv weight
// ring does), D is read back with tcgen05.ld.  Compared with a CPU reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_test tools/umma_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../superpoints_registration_b200/csrc/tc05.cuh"

using namespace spr::tc;

constexpr int M = 128;
constexpr int KATOMS = 4;  // K = 256
constexpr int K = KATOMS * 64;

// NS = rows per B stage (N of one MMA), NSUB = number of N sub-blocks
template <int NS, int NSUB>
__global__ void __launch_bounds__(192) k_test(const __half* __restrict__ a, const unsigned char* __restrict__ bimg,
                                              float* __restrict__ d) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* sA = smem;                              // KATOMS tiles of 128 rows x 128 B
  unsigned char* sB = smem + KATOMS * M * 128;           // 2 ring stages of NS x 128 B
  __shared__ uint64_t full[2], empty[2], done;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int STAGE = NS * 128;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(&done, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc(&tmem_base, NS * NSUB < 32 ? 32 : NS * NSUB);
  // A: thread t < 128 writes row t
  if (tid < M) {
    for (int atom = 0; atom < KATOMS; ++atom)
      for (int j = 0; j < 8; ++j) {
        const uint4 v = *reinterpret_cast<const uint4*>(a + (size_t)tid * K + atom * 64 + j * 8);
        *reinterpret_cast<uint4*>(sA + atom * M * 128 + sw128_offset(tid, j)) = v;
      }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base;
  if (warp == 5 && lane == 0) {  // producer
    int it = 0;
    for (int atom = 0; atom < KATOMS; ++atom)
      for (int s = 0; s < NSUB; ++s, ++it) {
        const int st = it & 1;
        mbar_wait(&empty[st], ((it >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[st], STAGE);
        bulk_g2s(sB + st * STAGE, bimg + (size_t)it * STAGE, STAGE, &full[st]);
      }
  } else if (warp == 4 && lane == 0) {  // MMA issuer
    constexpr uint32_t idesc = idesc_f16_f32(M, NS);
    int it = 0;
    for (int atom = 0; atom < KATOMS; ++atom)
      for (int s = 0; s < NSUB; ++s, ++it) {
        const int st = it & 1;
        mbar_wait(&full[st], (it >> 1) & 1);
        tc_fence_after();
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t ad = desc_sw128_kmajor(smem_u32(sA + atom * M * 128) + kk * 32);
          const uint64_t bd = desc_sw128_kmajor(smem_u32(sB + st * STAGE) + kk * 32);
          umma_f16(tb + s * NS, ad, bd, idesc, (atom | kk) != 0);
        }
        umma_commit(&empty[st]);
      }
    umma_commit(&done);
  }
  if (warp < 4) {
    mbar_wait(&done, 0);
    tc_fence_after();
    for (int c0 = 0; c0 < NS * NSUB; c0 += 8) {
      float v[8];
      tmem_ld8(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 8; ++i) d[(size_t)(warp * 32 + lane) * (NS * NSUB) + c0 + i] = v[i];
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) tmem_dealloc(tb, NS * NSUB < 32 ? 32 : NS * NSUB);
}

template <int NS, int NSUB>
int run() {
  constexpr int N = NS * NSUB;
  std::vector<__half> ha((size_t)M * K), hb((size_t)N * K);
  std::vector<float> fa(ha.size()), fb(hb.size());
  srand(1);
  for (size_t i = 0; i < ha.size(); ++i) {
    ha[i] = __float2half((rand() % 2001 - 1000) / 500.f);
    fa[i] = __half2float(ha[i]);
  }
  for (size_t i = 0; i < hb.size(); ++i) {
    hb[i] = __float2half((rand() % 2001 - 1000) / 700.f);
    fb[i] = __half2float(hb[i]);
  }
  // stage images: order (atom, sub), each NS rows x 128 B swizzled
  std::vector<unsigned char> img((size_t)KATOMS * NSUB * NS * 128);
  size_t it = 0;
  for (int atom = 0; atom < KATOMS; ++atom)
    for (int s = 0; s < NSUB; ++s, ++it)
      for (int r = 0; r < NS; ++r)
        for (int j = 0; j < 8; ++j)
          memcpy(&img[it * NS * 128 + sw128_offset(r, j)], &hb[(size_t)(s * NS + r) * K + atom * 64 + j * 8], 16);
  __half* da;
  unsigned char* dimg;
  float* dd;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&dimg, img.size());
  cudaMalloc(&dd, (size_t)M * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dimg, img.data(), img.size(), cudaMemcpyHostToDevice);
  cudaMemset(dd, 0, (size_t)M * N * 4);
  const int smem = KATOMS * M * 128 + 2 * NS * 128 + 1024;
  cudaFuncSetAttribute(k_test<NS, NSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_test<NS, NSUB><<<1, 192, smem>>>(da, dimg, dd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("NS=%d NSUB=%d: CUDA error %s\n", NS, NSUB, cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> hd((size_t)M * N);
  cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < K; ++k) acc += (double)fa[(size_t)m * K + k] * fb[(size_t)n * K + k];
      maxerr = fmax(maxerr, fabs(acc - hd[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(acc));
    }
  printf("NS=%d NSUB=%d (N=%d, K=%d): max abs err %.3e (max |ref| %.3e) -> %s\n", NS, NSUB, N, K, maxerr, maxref,
         maxerr < 1e-4 * maxref ? "OK" : "MISMATCH");
  cudaFree(da);
  cudaFree(dimg);
  cudaFree(dd);
  return maxerr < 1e-4 * maxref ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<64, 1>();
  bad += run<128, 1>();
  bad += run<128, 2>();
  bad += run<128, 4>();
  printf(bad ? "FAILED\n" : "ALL OK\n");
  return bad;
}
