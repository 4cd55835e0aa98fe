// spr_common.cuh -- shared helpers for the sm_100a kernels behind include/spr_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/spr_b200.h"

namespace spr {

// ---- host-side error plumbing -------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

#define SPR_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      ::spr::set_error(__VA_ARGS__);    \
      return SPR_EINVAL;                \
    }                                   \
  } while (0)

#define SPR_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      ::spr::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return SPR_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)

#define SPR_LAUNCH_CHECK(name)                                                             \
  do {                                                                                     \
    ::spr::count_launch();                                                                 \
    cudaError_t e__ = cudaGetLastError();                                                  \
    if (e__ != cudaSuccess) {                                                              \
      ::spr::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));          \
      return SPR_ECUDA;                                                                    \
    }                                                                                      \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// cudaFuncSetAttribute(func, MaxDynamicSharedMemorySize, bytes), done once per (kernel, device) and safe to call
// from several host threads (a process-wide `static bool` would leave the second GPU of a process unconfigured).
cudaError_t ensure_max_dynamic_smem(const void* func, size_t bytes);

// sticky numeric flags raised by kernels that write fp16 operand images (one word per translation unit, read and
// combined by spr_numeric_flags)
unsigned int gemm_numeric_flags(bool reset);
unsigned int blocks_numeric_flags(bool reset);
unsigned int attention_numeric_flags(bool reset);
unsigned int attention_tc_numeric_flags(bool reset);

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Carver {
  char* base;
  size_t off;
  size_t cap;
  Carver(void* p, size_t bytes) : base(static_cast<char*>(p)), off(0), cap(bytes) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

// ---- device helpers --------------------------------------------------------------------------------
#ifdef __CUDACC__

constexpr unsigned kFull = 0xffffffffu;

// Packed fp32 pairs (sm_100 add/mul/fma.f32x2: one issue slot for two lanes of the FMA pipe).  The kernel is bound by
// instruction issue, not by the pipe: the KPConv kernels evaluate their influences two at a time.
typedef unsigned long long f2_t;
__device__ __forceinline__ f2_t f2_pack(float lo, float hi) {
  f2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f2_t f2_add(f2_t a, f2_t b) {
  f2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2_t f2_mul(f2_t a, f2_t b) {
  f2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f2_t f2_fma(f2_t a, f2_t b, f2_t c) {
  f2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}


// Monotone map float -> uint32 so that atomicMin/atomicMax on the image order floats.
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Cloud that owns stacked row i, given the exclusive prefix offs[0..B] (offs[B] = total).
__device__ __forceinline__ int find_cloud(const int* __restrict__ offs, int B, int i) {
  int lo = 0, hi = B;  // invariant: offs[lo] <= i < offs[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(offs + mid) <= i)
      lo = mid;
    else
      hi = mid;
  }
  return lo;
}

// fp32 squared distance with the reference's rounding sequence: three rounded products, two rounded
// sums, left to right, NO fused multiply-add (cloud.h:63-66 / nanoflann.hpp:434-441; the reference
// build has no -march flag, hence no FMA contraction).
__device__ __forceinline__ float sqdist_exact(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
  float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  return __fadd_rn(s, __fmul_rn(dz, dz));
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int warp_max(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float warp_maxf(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// power-of-two exponent e such that v * 2^e <= 2^target (v > 0, finite, normal); 0 otherwise.  |e| <= 60.
__device__ __forceinline__ int scale_exp(float v, int target) {
  const int ef = (int)((__float_as_uint(v) >> 23) & 0xffu);
  if (!(v > 0.f) || ef == 0 || ef == 255) return 0;  // zero, denormal, inf, nan
  const int e = target - (ef - 126);                  // v = m * 2^(ef-126), m in [0.5, 1)
  return e < -60 ? -60 : (e > 60 ? 60 : e);
}
__device__ __forceinline__ float pow2i(int e) { return __uint_as_float((uint32_t)(e + 127) << 23); }  // |e| <= 126

#endif  // __CUDACC__

// Exclusive scan of int32 (device), n up to 2^31; out may alias in.  d_total (optional) gets the sum.
// tmp: at least scan_tmp_ints(n) int32.
size_t scan_tmp_ints(size_t n);
int exclusive_scan_i32(const int32_t* d_in, int32_t* d_out, size_t n, int32_t* d_total, int32_t* d_tmp,
                       cudaStream_t stream);
// offs[0..B] = exclusive prefix of lens[0..B-1] (single small kernel).
int cloud_offsets(const int32_t* d_lens, int B, int32_t* d_offs, cudaStream_t stream);
// per-cloud bounding boxes as ordered ints: bb[3 b + a] = min, bb[3 (B + b) + a] = max of coordinate a
int cloud_bboxes(const float* d_pts, const int32_t* d_offs, int B, int n, uint32_t* d_bb, cudaStream_t stream);


// kpconv_tc.cu: tensor-core (tcgen05) KPConv path, Cin = Cout = c in {32, 64, 128, 256}
size_t kpconv_tc_workspace_bytes(int ns, int c);
int kpconv_tc_forward(const float* q, const float* s, const void* idx, int idx_is_64, int row_stride, int H,
                      const float* x, int c, const float* w, const float* kp, float extent, float* out, int nq, int ns,
                      void* workspace, cudaStream_t stream);

}  // namespace spr
