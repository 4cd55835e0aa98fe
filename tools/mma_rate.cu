// Throughput probe for the legacy warp-level tensor path (mma.sync) on sm_100a:
//   m16n8k8 tf32 and m16n8k16 bf16, fp32 accumulate, 8 independent accumulator tiles per warp.
// Prints MAC/clk/SM so the 3xTF32 register-split contraction can be costed against the fp32 FMA pipe
// (128 FMA/clk/SM).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate tools/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) k_mma(float* out, int iters) {
  float c[8][4];
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
  unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
  unsigned b[2] = {threadIdx.x * 11u, threadIdx.x * 13u};
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      if (KIND == 0) {
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      } else {
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < 8; ++t)
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_ffma(float* out, int iters) {
  float c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 0.001f + i;
  float a = 1.0001f, b = 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fmaf(c[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed fp32 FMA (Blackwell fma.rn.f32x2 -> FFMA2): two FMAs per lane per instruction
__global__ void __launch_bounds__(256) k_ffma2(float* out, int iters) {
  unsigned long long c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo = threadIdx.x * 0.001f + i, hi = lo + 0.5f;
    c[i] = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
  }
  const float af = 1.0001f, bf = 0.5f;
  unsigned long long a = ((unsigned long long)__float_as_uint(af) << 32) | __float_as_uint(af);
  unsigned long long b = ((unsigned long long)__float_as_uint(bf) << 32) | __float_as_uint(bf);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(c[i]) : "l"(a), "l"(b));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)(c[i] & 0xffffffffu)) + __uint_as_float((unsigned)(c[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int blocks_per_sm = 1; blocks_per_sm <= 4; blocks_per_sm *= 2) {
    for (int kind = 0; kind < 4; ++kind) {
      float best = 1e9f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (kind == 0) k_mma<0><<<sms * blocks_per_sm, 256>>>(out, iters);
        else if (kind == 1) k_mma<1><<<sms * blocks_per_sm, 256>>>(out, iters);
        else if (kind == 2) k_ffma<<<sms * blocks_per_sm, 256>>>(out, iters);
        else k_ffma2<<<sms * blocks_per_sm, 256>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double warps = (double)sms * blocks_per_sm * 8;
      const double macs = kind == 0 ? warps * iters * 8.0 * 16 * 8 * 8 : kind == 1 ? warps * iters * 8.0 * 16 * 8 * 16
                                                                                   : warps * iters * 16.0 * 32;
      const double per_s = macs / (best * 1e-3);
      printf("%s blocks/SM=%d: %.3f ms  %.1f TMAC/s  (%.0f MAC/clk/SM at the max clock %d MHz)\n",
             kind == 0 ? "mma.m16n8k8.tf32 " : kind == 1 ? "mma.m16n8k16.bf16" : kind == 2 ? "ffma             " : "ffma2 (f32x2)    ", blocks_per_sm, best,
             per_s / 1e12, per_s / sms / (clk * 1e3), clk / 1000);
    }
  }
  return 0;
}
