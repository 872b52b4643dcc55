// Shared host/device helpers for libitr_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <string>

#include "../../include/itr_b200.h"

namespace itr {

// thread-local last-error string behind itr_last_error()
std::string& last_error();
int fail(int code, const char* fmt, ...);

#define ITR_CHECK_CUDA(expr)                                                                          \
  do {                                                                                                \
    cudaError_t err__ = (expr);                                                                       \
    if (err__ != cudaSuccess)                                                                         \
      return ::itr::fail(ITR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__),     \
                         __FILE__, __LINE__);                                                         \
  } while (0)

#define ITR_REQUIRE(cond, ...)                                    \
  do {                                                            \
    if (!(cond)) return ::itr::fail(ITR_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define ITR_CHECK_LAUNCH() ITR_CHECK_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- device helpers --------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Monotone map float -> uint32 (larger float <=> larger key); NaN sorts above +inf.
__device__ __forceinline__ uint32_t orderable(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__host__ __device__ __forceinline__ uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } cv; cv.f = f; u = cv.u;
#endif
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);   // quiet NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
__device__ __forceinline__ float bf16_to_f32(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }

}  // namespace itr
