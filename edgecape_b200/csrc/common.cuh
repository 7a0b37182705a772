// Shared helpers for the edgecape_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/edgecape_b200.h"

namespace ec {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return EC_ERR_CUDA;
  }
  count_launch();
  return EC_OK;
}

#define EC_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      ec::set_error(__VA_ARGS__);      \
      return EC_ERR_INVALID;           \
    }                                  \
  } while (0)

#define EC_CUDA(call)                                                          \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      ec::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
      return EC_ERR_CUDA;                                                      \
    }                                                                          \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

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

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// erff / tanhf expand to ~50-100 instructions each: kept out of line so that heavily unrolled epilogues
// (32+ call sites per loop body) stay inside the instruction cache
static __device__ __noinline__ float apply_act_transcendental(float x, int act) {
  return act == EC_ACT_GELU ? gelu_erf(x) : tanhf(x);
}
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == EC_ACT_NONE) return x;
  if (act == EC_ACT_RELU) return fmaxf(x, 0.0f);
  return apply_act_transcendental(x, act);
}

}  // namespace ec
