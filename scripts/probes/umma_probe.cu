// Micro-probe: issue rate of tcgen05.mma kind::f16 (M = 128, K = 16) as a function of N, operand source and
// accumulator dependence.  One CTA, one issuing thread, clock64 around [first issue .. commit observed].
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu ; run on a B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// mode bits: 1 = A from TMEM, 2 = B MN-major, 4 = alternate between two accumulators, 8 = four accumulators
__global__ void __launch_bounds__(128) probe(int M, int N, int mode, int count, int tmem_cols, int smem_kb, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < smem_kb * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  if (threadIdx.x < 32 && (((mode & 16) && elect_one()) || (!(mode & 16) && threadIdx.x == 0))) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24) | ((mode & 2) ? (1u << 16) : 0u);
    const uint32_t a_s = base, b_s = base + 16384;
    const int nacc = (mode & 8) ? 4 : ((mode & 4) ? 2 : 1);
    for (int rep = 0; rep < 2; ++rep) {          // rep 0 warms up
      const long long t0 = clock64();
      for (int i = 0; i < count; ++i) {
        const int k = i & 3;                       // 4 k-steps inside a 64-wide swizzle span, then wrap
        const uint32_t d = tmem + (uint32_t)((i % nacc) * 64);
        const uint64_t bd = (mode & 2) ? desc_mn(b_s) + 128 * k : desc_k(b_s) + 2 * k;
        if (mode & 1) umma_ts(d, tmem + (tmem_cols - 32) + 8 * k, bd, idesc, 1u);
        else umma_ss(d, desc_k(a_s) + 2 * k, bd, idesc, 1u);
      }
      const long long t1 = clock64();
      commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), rep & 1);
      const long long t2 = clock64();
      if (rep == 1 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
}

static void run(int M, int N, int mode, int grid, int tmem_cols, int smem_kb, const char* what, long long* out) {
  const int count = 96;
  probe<<<grid, 128, smem_kb * 1024 + 1024>>>(M, N, mode, count, tmem_cols, smem_kb, out);
  long long h[2];
  cudaError_t e = cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("M=%d N=%d mode=%d: %s\n", M, N, mode, cudaGetErrorString(e)); exit(1); }
  printf("M=%-4d N=%-4d %-44s issue/mma %7.1f  total/mma %7.1f\n", M, N, what, (double)h[0] / count, (double)h[1] / count);
}

int main() {
  long long* out;
  cudaMalloc(&out, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 162 * 1024);
  const char* names[] = {"SS K-major B", "TS K-major B", "SS MN-major B", "TS MN-major B"};
  // (a) one CTA on the GPU: N sweep, operand source, accumulator dependence
  for (int N : {16, 64, 128, 192, 256})
    for (int m = 0; m < 4; ++m)
      for (int accs : {0, 4}) {
        if (accs && N > 64) continue;
        char nm[96];
        snprintf(nm, sizeof nm, "%s, %d acc, 1 CTA", names[m], accs ? 2 : 1);
        run(128, N, m | accs, 1, 512, 160, nm, out);
      }
  // (a2) the same issued under elect.sync instead of threadIdx.x == 0
  for (int N : {16, 64, 128, 192, 256})
    for (int m : {0, 1, 3}) {
      char nm[96];
      snprintf(nm, sizeof nm, "%s, 1 acc, 1 CTA, elect.sync", names[m]);
      run(128, N, m | 16, 1, 512, 160, nm, out);
    }
  run(128, 64, 16 | 4, 1, 512, 160, "SS K-major B, 2 acc, 1 CTA, elect.sync", out);
  run(64, 64, 16, 1, 512, 160, "SS K-major B, 1 acc, 1 CTA, elect.sync", out);
  run(64, 256, 16, 1, 512, 160, "SS K-major B, 1 acc, 1 CTA, elect.sync", out);
  run(128, 64, 16, 296, 256, 64, "SS, 296 CTAs (2 per SM), elect.sync", out);
  run(128, 64, 17, 296, 256, 64, "TS, 296 CTAs (2 per SM), elect.sync", out);
  // (b) M = 64
  for (int N : {64, 128, 256}) run(64, N, 0, 1, 512, 160, "SS K-major B, 1 acc, 1 CTA", out);
  // (c) two co-resident CTAs per SM (256 TMEM columns, 64 KB shared memory each), 2 x 148 CTAs
  for (int N : {64, 128}) {
    run(128, N, 0, 148, 256, 64, "SS, 148 CTAs (1 per SM)", out);
    run(128, N, 0, 296, 256, 64, "SS, 296 CTAs (2 per SM)", out);
    run(128, N, 1, 296, 256, 64, "TS, 296 CTAs (2 per SM)", out);
  }
  return 0;
}
