// Shared helpers for libedgeconv_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "edgeconv_b200.h"

namespace ecb200 {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char* fmt, ...);

constexpr int kMaxDevices = 64;
// true the first time the calling host thread sees the current device: used to opt a
// kernel into >48 KB of dynamic shared memory once per (thread, device)
inline bool first_use_on_device(bool (&seen)[kMaxDevices]) {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) return true;
  if (seen[d]) return false;
  seen[d] = true;
  return true;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div64(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace ecb200

#define ECB_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      ecb200::set_error(__VA_ARGS__);       \
      return ECB200_ERR_ARG;                \
    }                                       \
  } while (0)

#define ECB_CUDA(expr)                                                               \
  do {                                                                               \
    cudaError_t e__ = (expr);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      ecb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),      \
                        __FILE__, __LINE__);                                         \
      return ECB200_ERR_CUDA;                                                        \
    }                                                                                \
  } while (0)

// after a <<<>>> launch: report launch-configuration errors without synchronising
#define ECB_LAUNCH_CHECK(name)                                                       \
  do {                                                                               \
    cudaError_t e__ = cudaPeekAtLastError();                                         \
    if (e__ != cudaSuccess) {                                                        \
      (void)cudaGetLastError();                                                      \
      ecb200::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__));    \
      return ECB200_ERR_CUDA;                                                        \
    }                                                                                \
  } while (0)

namespace ecb200 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// max over the block (any shape up to 1024 threads); the result is valid in thread 0
__device__ __forceinline__ unsigned block_max_u32(unsigned m) {
  __shared__ unsigned wm[32];
  const int t = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((t & 31) == 0) wm[t >> 5] = m;
  __syncthreads();
  if (t < 32) {
    m = t < nw ? wm[t] : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  }
  return m;
}

__device__ __forceinline__ float leaky(float y, float slope) { return y > 0.f ? y : y * slope; }

// round to tf32 (nearest, ties away): hi = tf32_rna(v), lo = tf32_rna(v - hi) is the operand pair of
// the 3xTF32 tensor-core products
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// streaming (read-once) 128-bit load that does not pollute L1
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

}  // namespace ecb200
