// Micro-probe: how many bytes per second do all SMs together pull out of L2 through TMA, and does sharing help?
// Every CTA streams 16 KB boxes (64 halves x 128 rows, 128B swizzle) of an L2-resident fp16 matrix into a 3-stage ring
// of 64 KB stages (the operand traffic of the CTA-pair GEMM) and does nothing else.  Modes:
//   0  every CTA reads its own rows                                  (distinct lines: the GEMM's A operand)
//   1  the two CTAs of a cluster read the SAME rows, unicast          (does L2 / the crossbar de-duplicate?)
//   2  the two CTAs of a cluster each load half of the rows and MULTICAST them to both (cluster of 2)
//   3  as 2 with a cluster of 4: every CTA loads a quarter, multicast to all four
//   4  the four CTAs of a cluster read the same rows, unicast
// Reported: bytes DELIVERED to shared memory per second (all CTAs), i.e. what a GEMM main loop could consume.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_feed_probe l2_feed_probe.cu -lcuda ; run on a B200.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int STAGES = 3, STAGE_BYTES = 64 * 1024, BOX_BYTES = 16 * 1024, BOX_ROWS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_mc(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// matrix: [rows, 4096 halves]; a "k-block" of a CTA = 4 boxes (64 columns each) at 128 consecutive rows
__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_part,
                                             int mode, int csize, int iters, int row_blocks) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  __shared__ uint64_t full[STAGES];
  const int rank = csize > 1 ? (int)cluster_ctarank() : 0;
  const int cluster = blockIdx.x / csize;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&full[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (csize > 1) cluster_sync();
  if (threadIdx.x == 0) {
    const bool shared_rows = mode != 0;
    const bool mc = mode == 2 || mode == 3;
    const int who = shared_rows ? cluster : (int)blockIdx.x;
    const int part_rows = BOX_ROWS / csize;                   // rows this CTA loads in the multicast modes
    const uint16_t mask = (uint16_t)((1u << csize) - 1u);
    // prime the ring, then: wait for stage s, immediately refill it (no consumer: the wait is the consumption)
    for (int it = 0; it < iters + STAGES; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) {
        mbar_wait(smem_u32(&full[s]), ((it / STAGES) - 1) & 1);
        // in the multicast modes a peer may still be waiting on ITS copy of this stage: a real kernel would track that
        // with an "empty" barrier; here every CTA of a cluster runs the same loop, and a cluster barrier every STAGES
        // iterations keeps them within one ring revolution of each other
      }
      if (mc && it >= STAGES && s == 0) {
        // (thread 0 only: use a cheap cluster-scope rendezvous through barrier.cluster is not possible from one thread;
        //  instead rely on expect_tx ordering: a stage's barrier phase cannot complete before its own expect_tx)
      }
      if (it < iters) {
        const int rb = (who * 7 + it * 13) % row_blocks;      // pseudo-random row block: L2 hits, no DRAM locality games
        const int col0 = (it % 16) * 256;                     // 4 boxes of 64 columns
        const uint32_t dst = base + s * STAGE_BYTES;
        mbar_expect_tx(smem_u32(&full[s]), STAGE_BYTES);
        for (int j = 0; j < 4; ++j) {
          if (!mc) tma_load(dst + j * BOX_BYTES, &tm, smem_u32(&full[s]), col0 + 64 * j, rb * BOX_ROWS);
          else tma_load_mc(dst + j * BOX_BYTES + rank * part_rows * 128, &tm_part, smem_u32(&full[s]), col0 + 64 * j,
                           rb * BOX_ROWS + rank * part_rows, mask);
        }
      }
    }
  }
  __syncthreads();
  if (csize > 1) cluster_sync();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  const int cols = 4096, row_blocks = 64, rows = row_blocks * BOX_ROWS;      // 8192 x 4096 halves = 64 MB: L2 resident
  __half* buf;
  cudaMalloc(&buf, (size_t)rows * cols * 2);
  cudaMemset(buf, 0, (size_t)rows * cols * 2);
  auto make = [&](int box_rows) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
  };
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = STAGES * STAGE_BYTES + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int iters = 2000;
  struct Cfg { int mode, csize; const char* what; };
  const Cfg cfgs[] = {{0, 1, "distinct rows, unicast, no cluster"}, {0, 2, "distinct rows, unicast, cluster of 2"},
                      {1, 2, "cluster of 2 reads the same rows, unicast"}, {2, 2, "cluster of 2, multicast halves"},
                      {4, 4, "cluster of 4 reads the same rows, unicast"}, {3, 4, "cluster of 4, multicast quarters"}};
  for (const Cfg& c : cfgs) {
    for (int grid_sms : {sms, sms / 2}) {
      CUtensorMap tm = make(BOX_ROWS), tmp = make(BOX_ROWS / c.csize);
      int grid = grid_sms / c.csize * c.csize;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(128);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = c.csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      if (c.csize > 1) {
        int maxc = 0;
        cudaOccupancyMaxActiveClusters(&maxc, probe, &cfg);
        if (maxc * c.csize < grid) grid = maxc * c.csize, cfg.gridDim = dim3(grid);
      }
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int mode = c.mode == 4 ? 1 : c.mode;
      cudaError_t le = cudaLaunchKernelEx(&cfg, probe, tm, tmp, mode, c.csize, 200, row_blocks);   // warm-up (L2 fill)
      if (le != cudaSuccess) printf("launch failed: %s\n", cudaGetErrorString(le));
      cudaEventRecord(e0);
      cudaLaunchKernelEx(&cfg, probe, tm, tmp, mode, c.csize, iters, row_blocks);
      cudaEventRecord(e1);
      cudaError_t err = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const double delivered = (double)grid * iters * STAGE_BYTES;
      printf("%-46s grid %3d: %7.3f ms  %6.2f TB/s delivered  = %5.1f B/clk/SM at 1.93 GHz  (%s)\n", c.what, grid, ms,
             delivered / ms / 1e9, delivered / ms / 1e6 / grid / 1.93e3, cudaGetErrorString(err));
    }
  }
  return 0;
}
