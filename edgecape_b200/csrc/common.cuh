// Shared helpers for the edgecape_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/edgecape_b200.h"

namespace ec {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// device counters {values beyond the e4m3 range, values beyond the fp16 range} seen by F16F8 split producers on the
// current device; allocated on first use (which must be outside a stream capture).  nullptr + error on failure.
unsigned long long* overflow_counters();
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: raises it to `bytes` for `func` on the current
// device the first time (and whenever a larger value is asked for), under a mutex.  Returns a cudaError_t as int.
int ensure_dynamic_smem_impl(const void* func, int bytes);
template <typename F>
inline int ensure_dynamic_smem(F* func, int bytes) { return ensure_dynamic_smem_impl((const void*)func, bytes); }

// Programmatic dependent launch.  Kernels launched through launch_pdl() may start while their predecessor in the
// stream is still running; each of them calls pdl_wait() -- which returns once the predecessor grid has completed
// and its writes are visible -- BEFORE its first global-memory access (read or write), and
// pdl_launch_dependents() as early as it safely can (after TMEM allocation in the kernels that allocate TMEM: a
// dependent CTA that grabbed TMEM first would wait for us while we wait for TMEM).
// Measured (profiles/r01_n_attention_and_issue.md): it shortens eager launch chains slightly, does nothing for a
// CUDA graph running alone, and costs ~0.7 ms per step when two graphs share the GPU (pre-launched dependents sit
// on SMs while they wait), so the detector records its graphs with it switched off (detector.py, test_cfg['pdl']).
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in check_launch()
}

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

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

// (a, b) -> fp16 pairs hi = rn(a, b), lo = rn((a, b) - hi), so that a ~= hi.x + lo.x to 2^-22.  Uses the packed
// conversion (cvt.rn.f16x2.f32 -> F2FP.PACK_AB, an ALU op); the scalar __float2half_rn form compiles to F2F,
// which shares the quarter-rate unit with MUFU and was the limiter of the softmax / split epilogues.
// (a, b) minus the fp16 pair h2, in fp32 (exact): mixed-precision adds with a negated fp16 source -- ptxas folds each into
// ONE FHADD (`FHADD d, -h.H0/H1, a`), where converting the halves and subtracting took two instructions per element.
__device__ __forceinline__ void sub_h2(float a, float b, uint32_t h2, float& d0, float& d1) {
  asm("{\n\t.reg .b16 l, h, nl, nh;\n\t"
      "mov.b32 {l, h}, %2;\n\t"
      "neg.f16 nl, l;\n\t"
      "neg.f16 nh, h;\n\t"
      "add.rn.f32.f16 %0, nl, %3;\n\t"
      "add.rn.f32.f16 %1, nh, %4;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "r"(h2), "f"(a), "f"(b));
}
__device__ __forceinline__ void split_pair(float a, float b, __half2& hi, __half2& lo) {
  hi = __floats2half2_rn(a, b);
  float d0, d1;
  sub_h2(a, b, *reinterpret_cast<const uint32_t*>(&hi), d0, d1);
  lo = __floats2half2_rn(d0, d1);
}
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  __half2 h, l;
  split_pair(a, b, h, l);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---- split operand rows.  A matrix [rows, K] feeds the tensor-core GEMMs as rows of 4*Kp bytes in one of two formats
// (Kp = K rounded up to 64, zero padded; `fmt` = EC_SPLIT_F16X2 / EC_SPLIT_F16F8, include/edgecape_b200.h):
//   F16X2: [ hi16 : Kp halves | lo16 : Kp halves ]                 three fp16 products (ec_gemm_f16x3)
//   F16F8: [ hi16 : Kp halves | per 64-column block: hi8 x 64, lo8 x 64 ]  fp16 hi.hi + two e4m3 cross terms (ec_gemm_f16f8);
//          the two e4m3 planes are interleaved per 64 columns so that ONE TMA box with 128-byte rows (whole L2 lines) holds
//          both planes of a 64-deep k-block -- boxes with 64-byte rows moved measurably slower (qkv 65.9 -> 63.8 us, fc2
//          94.3 -> 90.0 us in the timing experiment, profiles/r03_w_gemm_e4m3_box_rows.log)
// F16F8 plane scales are static powers of two, so every product carries the same scale and all three accumulate into
// one fp32 accumulator:  A role (activations): hi8 = e4m3(hi16), lo8 = e4m3((a - hi16) 2^11);
//                        B role (weights, pre-scaled by s_w): hi8 = e4m3(hi16 2^-11), lo8 = e4m3(b s_w - hi16).
// e4m3 saturates at 448: an activation beyond that only degrades ITS cross terms to plain-fp16 accuracy (the hi16 plane
// carries the value); beyond 65504 hi16 itself overflows.  Both events are counted (ec_overflow_count).
// byte offset of the hi8 entry of column c in an F16F8 row of kp columns; the lo8 entry sits 64 bytes further
__host__ __device__ inline int f8_off(int kp, int c) { return 2 * kp + 128 * (c >> 6) + (c & 63); }
__device__ __forceinline__ uint32_t e4m3x2(float a, float b) {
  return (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);   // a in the low byte
}
__device__ __forceinline__ uint32_t e4m3x2_h2(uint32_t h2) {       // the same from a packed fp16 pair (one instruction)
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(r) : "r"(h2));
  return r;
}
// four consecutive elements (column c, a multiple of 4) of an A-role row; `row` = the row's first byte
__device__ __forceinline__ void store_split4(uint8_t* row, int kp, int c, float v0, float v1, float v2, float v3,
                                             int fmt, uint32_t& flags) {
  const __half2 h01 = __floats2half2_rn(v0, v1), h23 = __floats2half2_rn(v2, v3);
  const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
  float d0, d1, d2, d3;                                    // the low parts v - hi16 (exact)
  sub_h2(v0, v1, u01, d0, d1);
  sub_h2(v2, v3, u23, d2, d3);
  *reinterpret_cast<uint2*>(row + 2 * c) = make_uint2(u01, u23);
  if (fmt == EC_SPLIT_F16X2) {
    const __half2 l01 = __floats2half2_rn(d0, d1), l23 = __floats2half2_rn(d2, d3);
    *reinterpret_cast<uint2*>(row + 2 * kp + 2 * c) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  } else {
    *reinterpret_cast<uint32_t*>(row + f8_off(kp, c)) = e4m3x2_h2(u01) | (e4m3x2_h2(u23) << 16);
    *reinterpret_cast<uint32_t*>(row + f8_off(kp, c) + 64) = e4m3x2(d0 * 2048.f, d1 * 2048.f) | (e4m3x2(d2 * 2048.f, d3 * 2048.f) << 16);
    const float m = fmaxf(fmaxf(fabsf(v0), fabsf(v1)), fmaxf(fabsf(v2), fabsf(v3)));
    flags |= (m > 448.f ? 1u : 0u) | (m > 65504.f ? 2u : 0u);
  }
}
// eight consecutive elements (column c, a multiple of 8): 16-byte hi16 store, 16-byte lo16 store or two 8-byte e4m3
// stores -- with four lanes per row every 32-byte sector of the row is written whole by one instruction
__device__ __forceinline__ void store_split8(uint8_t* row, int kp, int c, const float* v, int fmt, uint32_t& flags) {
  uint32_t h[4];
  float dl[8];                                             // the low parts v - hi16 (exact)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    sub_h2(v[2 * i], v[2 * i + 1], h[i], dl[2 * i], dl[2 * i + 1]);
  }
  *reinterpret_cast<uint4*>(row + 2 * c) = make_uint4(h[0], h[1], h[2], h[3]);
  if (fmt == EC_SPLIT_F16X2) {
    uint32_t l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 ll = __floats2half2_rn(dl[2 * i], dl[2 * i + 1]);
      l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(row + 2 * kp + 2 * c) = make_uint4(l[0], l[1], l[2], l[3]);
  } else {
    *reinterpret_cast<uint2*>(row + f8_off(kp, c)) =
        make_uint2(e4m3x2_h2(h[0]) | (e4m3x2_h2(h[1]) << 16), e4m3x2_h2(h[2]) | (e4m3x2_h2(h[3]) << 16));
    uint32_t l8[4];
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      l8[i] = e4m3x2(dl[2 * i] * 2048.f, dl[2 * i + 1] * 2048.f);
      m = fmaxf(m, fmaxf(fabsf(v[2 * i]), fabsf(v[2 * i + 1])));
    }
    *reinterpret_cast<uint2*>(row + f8_off(kp, c) + 64) = make_uint2(l8[0] | (l8[1] << 16), l8[2] | (l8[3] << 16));
    flags |= (m > 448.f ? 1u : 0u) | (m > 65504.f ? 2u : 0u);
  }
}
// end of a thread's stores: one atomic per event class at most (flags is zero on every healthy run)
__device__ __forceinline__ void report_overflow(unsigned long long* counters, uint32_t flags) {
  if (flags && counters) {
    if (flags & 1u) atomicAdd(counters, 1ULL);
    if (flags & 2u) atomicAdd(counters + 1, 1ULL);
  }
}

// One lane of a converged warp.  Code that issues tcgen05.mma / TMA / tcgen05.commit must be guarded by this
// (inside a warp-uniform branch), not by `lane == 0`: those are uniform-datapath instructions, and under a
// lane-id branch ptxas wraps EVERY one of them in an ELECT / BRA.U.ANY serialisation loop with R2UR operand
// moves -- ~98 clk per tcgen05.mma whatever its shape (scripts/probes/umma_probe.cu), against 32..128 clk of
// tensor-pipe time.  After elect.sync ptxas knows a single thread is active and emits them back to back.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

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

// Exact-erf GELU for the tensor-core GEMM epilogue, in 13 instructions (one MUFU) instead of erff's ~30 with a
// divergent branch: erfc(t) = 2^(-t q(t)) with q a degree-7 polynomial (weighted least-squares fit of
// -log2(erfc(t)) / t on [0, 4.6], scripts/probes/gelu_fit.py: |erf error| <= 1.7e-8 in exact arithmetic), then
//   gelu(x) = x - x erfc(t) / 2   (x >= 0),   x erfc(t) / 2   (x < 0),   t = |x| / sqrt 2.
// Measured against the fp64 definition on [-12, 12]: max abs error 5.1e-7 (1 ulp of the result around x = 4);
// torch's own fp32 CPU GELU: 1.2e-6.  Beyond t = 4.6 erfc < 1e-10 and the clamp keeps q inside its fitted range.
__device__ __forceinline__ float gelu_fast(float x) {
  const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.6f);
  float q = 4.5224536734167486e-05f;
  q = fmaf(q, t, -0.00044406522647477686f);
  q = fmaf(q, t, 0.0014838487841188908f);
  q = fmaf(q, t, 0.0007849961402826011f);
  q = fmaf(q, t, -0.028263606131076813f);
  q = fmaf(q, t, 0.1484864503145218f);
  q = fmaf(q, t, 0.9184153079986572f);
  q = fmaf(q, t, 1.627908706665039f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-t * q));
  const float h = 0.5f * x * e;
  return x >= 0.f ? x - h : h;
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
