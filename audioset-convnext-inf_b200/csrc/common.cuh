// Shared host/device helpers for libacx (B200 / sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/acx.h"

namespace acx {

// ---- error plumbing (thread-local last error; no exceptions cross the C ABI) -------------
void set_error(const char* fmt, ...);
const char* last_error();

#define ACX_CHECK(cond, code, ...)                  \
  do {                                              \
    if (!(cond)) {                                  \
      ::acx::set_error(__VA_ARGS__);                \
      return (code);                                \
    }                                               \
  } while (0)

#define ACX_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::acx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                       __FILE__, __LINE__);                                         \
      return ACX_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers ---------------------------------------------------------------------
using bf16 = __nv_bfloat16;

template <typename T> struct Pair;          // two consecutive channels
template <> struct Pair<float> {
  using type = float2;
  __device__ static __forceinline__ float2 unpack(float2 v) { return v; }
  __device__ static __forceinline__ float2 pack(float a, float b) { return make_float2(a, b); }
};
template <> struct Pair<bf16> {
  using type = uint32_t;
  // bf16 -> fp32 is a 16-bit shift: two ALU ops per pair, no conversion pipe.
  __device__ static __forceinline__ float2 unpack(uint32_t v) {
    return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
  }
  __device__ static __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_float<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Exact (erf) GELU -- nn.GELU() default, reference convnext.py:65.
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

}  // namespace acx
