// Micro-probe for the next GEMM step: does tcgen05.mma kind::f8f6f4 (e4m3 x e4m3 -> f32)
// produce the exact product from K-major shared-memory tiles laid out the way this library would lay them out, and at
// what rate?  Variants (M = 128, N = 128):
//   0  kind::f16,    128B swizzle, 64 fp16 per row (the layout gemm_tcgen05.cu uses today; sanity check of the probe)
//   1  kind::f8f6f4, 128B swizzle, 128 fp8 per row, K = 32 per UMMA (descriptor + 32 B per k-step)
//   2  kind::f8f6f4,  64B swizzle,  64 fp8 per row (SBO 512, layout type 4), chunk ^= (row >> 1) & 3
//   3  as 2 with chunk ^= row & 3 (in case the 64B pattern is not the canonical Swizzle<2,4,3>)
// The host builds the swizzled shared-memory images; values are small multiples of 0.5, so every product is exact.
// Then 64 back-to-back UMMAs (N = 256) are timed for kind::f16 (K = 16) and kind::f8f6f4 (K = 32).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f8_mma_probe f8_mma_probe.cu ; run on a B200.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(sbo >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* u) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
        "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
        "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// A image at smem + 0, B image at smem + b_off.  ksteps UMMAs, descriptors advance by 2 (32 B) per k-step; `reps` > 1
// repeats the whole sequence (timing).  out: [128][N] floats (nullptr when timing), clk: elapsed clocks of the issue loop.
__global__ void __launch_bounds__(128) probe(const uint8_t* imgA, int bytesA, const uint8_t* imgB, int bytesB, int b_off,
                                             int sbo, int layout, int f8, int N, int ksteps, int reps, float* out,
                                             long long* clk) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* g = raw + (base - smem_u32(raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < bytesA; i += blockDim.x) g[i] = imgA[i];
  for (int i = threadIdx.x; i < bytesB; i += blockDim.x) g[b_off + i] = imgB[i];
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  // f32 accumulator, M = 128, N; a / b format 0 = F16 (kind::f16) or E4M3 (kind::f8f6f4); both operands K-major
  const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x < 32) {
    if (elect_one()) {
      const uint64_t a = make_desc(base, sbo, layout), b = make_desc(base + b_off, sbo, layout);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r)
        for (int k = 0; k < ksteps; ++k) {
          if (f8) umma_f8(tmem, a + 2 * k, b + 2 * k, idesc, (r | k) ? 1u : 0u);
          else umma_f16(tmem, a + 2 * k, b + 2 * k, idesc, (r | k) ? 1u : 0u);
        }
      commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), 0);
      if (clk) *clk = clock64() - t0;
    }
    __syncwarp();
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = 0; c < N; c += 32) {
      uint32_t u[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, u);
      for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c + j] = __uint_as_float(u[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

static const float VALS[8] = {0.f, 0.5f, 1.f, 1.5f, 2.f, 3.f, -1.f, -2.f};
static const uint8_t E4M3[8] = {0x00, 0x30, 0x38, 0x3C, 0x40, 0x44, 0xB8, 0xC0};

int main() {
  const int M = 128, N = 128;
  int fails = 0;
  for (int variant = 0; variant < 4; ++variant) {
    const bool f8 = variant > 0;
    const int row_bytes = variant >= 2 ? 64 : 128;
    const int esz = f8 ? 1 : 2;
    const int K = row_bytes / esz;                       // one swizzle row of K elements
    const int kstep_elems = f8 ? 32 : 16;
    const int ksteps = K / kstep_elems;
    const int layout = variant >= 2 ? 4 : 2, sbo = 8 * row_bytes;
    std::vector<int> ia(M * K), ib(N * K);
    srand(1234 + variant);
    for (auto& v : ia) v = rand() % 8;
    for (auto& v : ib) v = rand() % 8;
    auto image = [&](const std::vector<int>& idx, int rows) {
      std::vector<uint8_t> img((size_t)rows * row_bytes, 0);
      const int chunks = row_bytes / 16;
      for (int r = 0; r < rows; ++r)
        for (int k = 0; k < K; ++k) {
          const int byte = k * esz, c = byte / 16, within = byte % 16;
          int cs = c;
          if (variant <= 1) cs = c ^ (r & 7);
          else if (variant == 2) cs = c ^ ((r >> 1) & 3);
          else cs = c ^ (r & 3);
          (void)chunks;
          uint8_t* dst = &img[(size_t)r * row_bytes + cs * 16 + within];
          if (f8) dst[0] = E4M3[idx[r * K + k]];
          else {
            const __half h = __float2half(VALS[idx[r * K + k]]);
            memcpy(dst, &h, 2);
          }
        }
      return img;
    };
    std::vector<uint8_t> A = image(ia, M), B = image(ib, N);
    uint8_t *dA, *dB;
    float* dO;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dO, M * N * 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, M * N * 4);
    const int b_off = 32768, smem = 65536 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    probe<<<1, 128, smem>>>(dA, (int)A.size(), dB, (int)B.size(), b_off, sbo, layout, f8 ? 1 : 0, N, ksteps, 1, dO, nullptr);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> O(M * N);
    cudaMemcpy(O.data(), dO, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    int bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < K; ++k) ref += VALS[ia[m * K + k]] * VALS[ib[n * K + k]];
        const double d = fabs((double)O[m * N + n] - ref);
        if (d > maxerr) maxerr = d;
        bad += d > 1e-3;
      }
    printf("variant %d (%s, %dB swizzle, K=%d, %d UMMAs): %s  max|err| %.3g  wrong %d / %d\n", variant, f8 ? "kind::f8f6f4 e4m3" : "kind::f16",
           row_bytes, K, ksteps, e == cudaSuccess ? "ran" : cudaGetErrorString(e), maxerr, bad, M * N);
    if (e != cudaSuccess) { fails++; cudaDeviceReset(); continue; }
    cudaFree(dA); cudaFree(dB); cudaFree(dO);
  }
  // ---- rate: 64 x 4 back-to-back UMMAs, M = 128, N = 256, one issuing thread
  for (int f8 = 0; f8 < 2; ++f8) {
    uint8_t *dA, *dB;
    long long* dC;
    cudaMalloc(&dA, 16384); cudaMalloc(&dB, 32768); cudaMalloc(&dC, 8);
    cudaMemset(dA, 0, 16384); cudaMemset(dB, 0, 32768);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; ++rep)
      probe<<<1, 128, 65536 + 1024>>>(dA, 16384, dB, 32768, 16384, 1024, 2, f8, 256, 4, 64, nullptr, dC);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
    printf("rate %s: N=256, K=%d per UMMA: %.1f clk per UMMA over 256 (%s)\n", f8 ? "kind::f8f6f4" : "kind::f16   ", f8 ? 32 : 16,
           c / 256.0, e == cudaSuccess ? "ok" : cudaGetErrorString(e));
  }
  return fails;
}
