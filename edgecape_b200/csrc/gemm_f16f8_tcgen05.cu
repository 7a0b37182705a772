// EXPERIMENTAL (opt-in, not on the default path; written at the end of round 1 after the GPU budget was spent --
// it compiles for sm_100a but has NOT run on hardware yet; tests/test_experimental_gpu.py is skipped unless
// EDGECAPE_TEST_EXPERIMENTAL=1).  The plan and the evidence behind it: DESIGN.md section 9 item 1,
// scripts/probes/f8_mma_probe.cu (hardware facts), scripts/probes/fp8_split_e2e.py (end-to-end accuracy emulation).
//
// fp32-grade GEMM in 2 instead of 3 units of tensor time:  C = epi(A . B^T) with
//     a.b ~= a_hi.b_hi (fp16 UMMAs)  +  a_lo.b_hi + a_hi.b_lo (e4m3 UMMAs at twice the contraction depth per clock)
// The cross terms only need ~11 bits of relative accuracy, and with STATIC power-of-two plane scales every product
// carries the same scale (the weight scale s_w), so all three accumulate into ONE fp32 TMEM accumulator:
//     operand rows  [ hi16 : Kp halves | hi8 : Kp bytes | lo8 : Kp bytes ]      (4 Kp bytes, as [hi16 | lo16] today)
//     role A (activations):  hi16 = fp16(a)       hi8 = e4m3(hi16)           lo8 = e4m3((a - hi16) 2^11)
//     role B (weights):      hi16 = fp16(b s_w)   hi8 = e4m3(hi16 2^-11)     lo8 = e4m3(b s_w - hi16)
//     acc = A.hi16 . B.hi16  +  A.lo8 . B.hi8  +  A.hi8 . B.lo8               (every term = s_w x the exact product term)
// One 64-deep k-block stage: fp16 tiles in 128B swizzle (as gemm_tcgen05.cu), fp8 tiles in 64B swizzle (64 B rows,
// 8-row atoms of 512 B, descriptor layout type 4) through a second, UINT8 tensor map on the same buffer;
// 4 kind::f16 UMMAs (K = 16) + 2 x 2 kind::f8f6f4 UMMAs (K = 32) per k-block instead of 12 kind::f16.
// Persistent 128 x 256 tiles on one CTA, two accumulator stages, warp-specialised (8 epilogue warps, MMA warp, TMA warp).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include <mutex>

#include "common.cuh"

namespace ec {
namespace tc {
int get_tensor_map(const void* ptr, int rows, int kp, int box_rows, CUtensorMap* out);   // gemm_tcgen05.cu (fp16 view)
}
namespace tc8 {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int EPI_WARPS = 8;
constexpr int THREADS = EPI_WARPS * 32 + 64;
constexpr int EPI_LD = 20;
constexpr int A16 = BM * BK * 2, B16 = BN * BK * 2, A8 = BM * BK, B8 = BN * BK;   // tile bytes
constexpr int STAGE_BYTES = A16 + B16 + 2 * A8 + 2 * B8;                           // 96 KB
constexpr int STAGES = 2;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + EPI_WARPS * 32 * EPI_LD * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major tile descriptors: 128B swizzle (8-row atoms of 1024 B, layout type 2) / 64B swizzle (atoms of 512 B, type 4)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Params {
  float* C;
  int M, N, num_kb, Kp, ldc;
  float out_scale;
  const float* bias;
  int act;
};

__global__ void __launch_bounds__(THREADS, 1)
gemm_f16f8_kernel(const __grid_constant__ CUtensorMap tmA16, const __grid_constant__ CUtensorMap tmB16,
                  const __grid_constant__ CUtensorMap tmA8, const __grid_constant__ CUtensorMap tmB8, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  const uint32_t epi_base = bar_base + 256u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BM - 1) / BM, n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();
  pdl_wait();

  // stage layout: [A16 | B16 | A.hi8 | A.lo8 | B.hi8 | B.lo8]
  constexpr uint32_t O_B16 = A16, O_AH8 = A16 + B16, O_AL8 = O_AH8 + A8, O_BH8 = O_AL8 + A8, O_BL8 = O_BH8 + B8;

  if (warp == EPI_WARPS + 1) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tm = tile / n_tiles, tn = tile % n_tiles;          // N fastest: an A row block is read once
        const int m0 = tm * BM, n0 = tn * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sb = base + stage * STAGE_BYTES;
          mbar_expect_tx(full_bar(stage), STAGE_BYTES);
          tma_load_2d(sb, &tmA16, full_bar(stage), kb * BK, m0);                       // halves: hi16 plane
          tma_load_2d(sb + O_B16, &tmB16, full_bar(stage), kb * BK, n0);
          tma_load_2d(sb + O_AH8, &tmA8, full_bar(stage), 2 * p.Kp + kb * BK, m0);     // bytes: hi8 plane
          tma_load_2d(sb + O_AL8, &tmA8, full_bar(stage), 3 * p.Kp + kb * BK, m0);     //        lo8 plane
          tma_load_2d(sb + O_BH8, &tmB8, full_bar(stage), 2 * p.Kp + kb * BK, n0);
          tma_load_2d(sb + O_BL8, &tmB8, full_bar(stage), 3 * p.Kp + kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == EPI_WARPS) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      // f32 accumulator, M = 128, N = 256, K-major operands; a / b format 0 = F16 (kind::f16) or E4M3 (kind::f8f6f4)
      constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const int acc = t & 1;
        mbar_wait(tempty_bar(acc), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sb = base + stage * STAGE_BYTES;
          const uint64_t a16 = desc_sw128(sb), b16 = desc_sw128(sb + O_B16);
          const uint64_t ah8 = desc_sw64(sb + O_AH8), al8 = desc_sw64(sb + O_AL8);
          const uint64_t bh8 = desc_sw64(sb + O_BH8), bl8 = desc_sw64(sb + O_BL8);
          // small terms first, then hi.hi
#pragma unroll
          for (int k = 0; k < BK / 32; ++k) umma_f8(tmem_d, al8 + 2 * k, bh8 + 2 * k, IDESC, (kb | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < BK / 32; ++k) umma_f8(tmem_d, ah8 + 2 * k, bl8 + 2 * k, IDESC, 1u);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_d, a16 + 2 * k, b16 + 2 * k, IDESC, 1u);
          umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int quarter = warp & 3, half = warp >> 2;          // TMEM lane quarter; which half of the 256 columns
    float* stg = reinterpret_cast<float*>(smem_raw + (epi_base - smem_u32(smem_raw))) + warp * (32 * EPI_LD);
    const int sub_row = lane >> 2, c4 = (lane & 3) * 4;
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      const int tm = tile / n_tiles, tn = tile % n_tiles;
      const int m0 = tm * BM, n0 = tn * BN;
      mbar_wait(tfull_bar(acc), (t >> 1) & 1);
      tc_fence_after();
      for (int sc = 0; sc < BN / 32; ++sc) {
        const int col0 = half * (BN / 2) + sc * 16;
        float r[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + col0), r);
        if (sc == BN / 32 - 1) {                             // this warp has drained its share of the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(stg + lane * EPI_LD + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        __syncwarp();
        const int gcol = n0 + col0 + c4;
        if (gcol < p.N) {
          float bv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (p.bias && gcol + u < p.N) bv[u] = __ldg(p.bias + gcol + u);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + sub_row, row = m0 + quarter * 32 + rr;
            if (row >= p.M) continue;
            const float4 a4 = *reinterpret_cast<const float4*>(stg + rr * EPI_LD + c4);
            float y[4] = {fmaf(a4.x, p.out_scale, bv[0]), fmaf(a4.y, p.out_scale, bv[1]), fmaf(a4.z, p.out_scale, bv[2]),
                          fmaf(a4.w, p.out_scale, bv[3])};
#pragma unroll
            for (int u = 0; u < 4; ++u) y[u] = apply_act(y[u], p.act);
            float* cp = p.C + (long long)row * p.ldc + gcol;
            if (gcol + 3 < p.N && (p.ldc & 3) == 0) {
              *reinterpret_cast<float4*>(cp) = make_float4(y[0], y[1], y[2], y[3]);
            } else {
#pragma unroll
              for (int u = 0; u < 4; ++u)
                if (gcol + u < p.N) cp[u] = y[u];
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == EPI_WARPS) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

// X [M, K] fp32 (row stride ldx) -> rows [hi16 : Kp halves | hi8 : Kp bytes | lo8 : Kp bytes] of X * scale, zero padded.
// role 0 = A operand (activations), 1 = B operand (weights): see the plane scales at the top of the file.
__global__ void split_f16f8_kernel(const float* __restrict__ X, uint8_t* __restrict__ out, int M, int K, int ldx, int Kp,
                                   float scale, int role) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 2 elements
  const int half_kp = Kp >> 1;
  if (i >= (long long)M * half_kp) return;
  const int m = (int)(i / half_kp), k = (int)(i % half_kp) * 2;
  const float* x = X + (long long)m * ldx;
  float v0 = 0.f, v1 = 0.f;
  if (k < K) v0 = x[k] * scale;
  if (k + 1 < K) v1 = x[k + 1] * scale;
  const __half2 hi = __floats2half2_rn(v0, v1);
  const float2 hf = __half22float2(hi);
  const float s_hi = role ? 0.00048828125f : 1.0f;          // 2^-11 for the weights' hi8 plane
  const float s_lo = role ? 1.0f : 2048.0f;                 // 2^11 for the activations' lo8 plane
  const __nv_fp8x2_storage_t h8 = __nv_cvt_float2_to_fp8x2(make_float2(hf.x * s_hi, hf.y * s_hi), __NV_SATFINITE, __NV_E4M3);
  const __nv_fp8x2_storage_t l8 =
      __nv_cvt_float2_to_fp8x2(make_float2((v0 - hf.x) * s_lo, (v1 - hf.y) * s_lo), __NV_SATFINITE, __NV_E4M3);
  uint8_t* row = out + (long long)m * 4 * Kp;
  *reinterpret_cast<__half2*>(row + 2 * k) = hi;
  *reinterpret_cast<__nv_fp8x2_storage_t*>(row + 2 * Kp + k) = h8;
  *reinterpret_cast<__nv_fp8x2_storage_t*>(row + 3 * Kp + k) = l8;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

// byte view of a three-plane operand: dims {4 Kp bytes, rows}, 64-byte x box_rows boxes, 64B swizzle
static int byte_map(const void* ptr, int rows, int kp, int box_rows, CUtensorMap* out) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return EC_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)(4 * kp), (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(4 * kp)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (uint8 view) failed with CUresult %d (ptr %p rows %d kp %d)", (int)r, ptr, rows, kp);
    return EC_ERR_CUDA;
  }
  return EC_OK;
}

}  // namespace tc8
}  // namespace ec

using namespace ec;

extern "C" int ec_split_f16f8(const float* X, void* out, int M, int K, int ldx, int Kp, float scale, int role,
                              void* stream) {
  EC_REQUIRE(X && out, "ec_split_f16f8: null pointer");
  EC_REQUIRE(Kp % tc8::BK == 0 && Kp >= K && K > 0, "ec_split_f16f8: Kp must be a multiple of 64 and >= K");
  EC_REQUIRE(role == 0 || role == 1, "ec_split_f16f8: role is 0 (A operand) or 1 (B operand)");
  if (M == 0) return EC_OK;
  const long long total = (long long)M * (Kp / 2);
  launch_pdl(tc8::split_f16f8_kernel, dim3(cdiv(total, 256)), dim3(256), 0, (cudaStream_t)stream, X, (uint8_t*)out, M, K,
             ldx, Kp, scale, role);
  return check_launch("ec_split_f16f8");
}

extern "C" int ec_gemm_f16f8(const void* A3, const void* B3, float* C, int M, int N, int Kp, int ldc, float out_scale,
                             const float* bias, int act, void* stream) {
  EC_REQUIRE(A3 && B3 && C, "ec_gemm_f16f8: null operand");
  EC_REQUIRE(Kp > 0 && Kp % tc8::BK == 0, "ec_gemm_f16f8: Kp must be a positive multiple of 64");
  EC_REQUIRE(aligned16(A3) && aligned16(B3), "ec_gemm_f16f8: operands must be 16-byte aligned");
  if (M == 0 || N == 0) return EC_OK;
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    EC_CUDA(cudaGetDevice(&dev));
    EC_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    EC_CUDA(cudaFuncSetAttribute(tc8::gemm_f16f8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc8::SMEM_BYTES));
  }
  CUtensorMap tmA16, tmB16, tmA8, tmB8;
  int rc = tc::get_tensor_map(A3, M, Kp, tc8::BM, &tmA16);       // halves view: row pitch 2 Kp halves = 4 Kp bytes
  if (rc) return rc;
  rc = tc::get_tensor_map(B3, N, Kp, tc8::BN, &tmB16);
  if (rc) return rc;
  rc = tc8::byte_map(A3, M, Kp, tc8::BM, &tmA8);
  if (rc) return rc;
  rc = tc8::byte_map(B3, N, Kp, tc8::BN, &tmB8);
  if (rc) return rc;
  tc8::Params p;
  p.C = C; p.M = M; p.N = N; p.num_kb = Kp / tc8::BK; p.Kp = Kp; p.ldc = ldc; p.out_scale = out_scale; p.bias = bias; p.act = act;
  const int tiles = cdiv(M, tc8::BM) * cdiv(N, tc8::BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  launch_pdl(tc8::gemm_f16f8_kernel, dim3(grid), dim3(tc8::THREADS), (size_t)tc8::SMEM_BYTES, (cudaStream_t)stream, tmA16,
             tmB16, tmA8, tmB8, p);
  return check_launch("ec_gemm_f16f8");
}
