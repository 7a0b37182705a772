// Fused GCN feed-forward for sm_100a (encoder_decoder.py:508-524 of the reference, aggregate-first form):
//     Y[b,w,:] = relu( a0[b,w] * (X[b,w,:] W0^T + b0) + sum_v A1[b,w,v] (X[b,v,:] W1^T + b1) )
// as ONE kernel: no Z operand round trip through global memory, no second launch.
//
// CTA = (sample b, slice of NS output channels).  Everything runs on the tensor cores at fp32-grade accuracy
// (three fp16 products per fp32 product, see gemm_tcgen05.cu):
//   fill    16 worker warps read A1[b] and X[b] once (coalesced float4, every load issued ahead of the CTA setup
//           barrier), split them into fp16 hi | lo and write 128B-swizzled shared-memory tiles: A1 as a K-major
//           A operand, X as [v][64 columns] row tiles; row sums of A1 by 16-lane shuffles on the way.
//   GEMM 1  D1[128 x d] = A1 . X in TMEM; the X tiles are consumed as an MN-major B operand (N = d in one UMMA).
//   scale   X tiles are rescaled in place by a0[w] (a no-op for the usual a0 = 1): the very same bytes are now
//           the K-major A operand a0*X of GEMM 2 (an MN-major [v][c] tile and a K-major [w][c] tile coincide).
//   GEMM 2  acc[128 x NS] = (a0 X) . W0^T  (k-blocks 0..d/64-1)  +  (A1 X) . W1^T  (k-blocks d/64..2d/64-1);
//           W streams through a two-stage TMA ring from the pre-split packed weights; while the first half runs
//           the workers drain D1 (tcgen05.ld), split it and write it over the X tile whose k-block has retired.
//   epilogue  relu(acc / w_scale + a0[w] b0[n] + rowsum(A1[w,:]) b1[n]) in fp32 through a per-warp staging tile
//           (64 B row segments), written as fp32 rows and / or as the split-fp16 A operand of the ffn2 GEMM.
// Measured (profiles/r01_ah_gcn_fused.md): 15.5 us at B = 64, K = 100, d = 256, dff = 384 (two-launch path 52.8 us);
// tensor pipe 13.3 K of the 24 K clocks of a CTA, then L2 write ingest (epilogue) and the fill.
// Shared memory (K = 100, d = 256, NS = 192): X tiles 8 x 14 KB, A1 tiles 4 x 14 KB (recycled as W stage 1 once
// GEMM 1 has retired), W stage 0 48 KB.  Tiles are k16 = ceil16(K) rows tall; an A operand always spans 128 rows,
// so the MMA reads past a tile into its neighbour: that only produces garbage in accumulator rows >= K, which
// are never stored.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace ec {
namespace tc {
int get_tensor_map(const void* ptr, int rows, int kp, int box_rows, CUtensorMap* out);   // gemm_tcgen05.cu
}
namespace gf {

constexpr int WORKERS = 512;             // 16 worker warps (4 per TMEM lane quarter)
constexpr int EPI_LD = 20;               // epilogue staging row stride (floats): conflict-free float4 writes
constexpr int THREADS = WORKERS + 64;    // + MMA warp + TMA warp
constexpr int ACC_COL = 256;             // TMEM: D1 in columns [0, d), the output accumulator in [256, 256 + NS)
constexpr int MAX_G = 4;                 // d <= 256: at most four 64-column groups

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
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major operand tile, 128B swizzle: rows of 128 B, 8-row atoms of 1024 B (SBO)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major B operand, 128B swizzle: rows (= k index) of 128 B holding 64 contiguous N elements, 8-row atoms of
// 1024 B (SBO); the next 64 N elements live `lbo` bytes further (one tile)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {           // f16 x f16 -> f32, M = 128, N = n, K-major A and B
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
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

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float* r) {   // caller issues tcgen05.wait::ld
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr));
}

// byte offset of the 16-byte chunk `c` (8 fp16) of row `row` inside a 128B-swizzled tile
__device__ __forceinline__ uint32_t swz(int row, int c) { return (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4)); }

__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair(v[2 * i], v[2 * i + 1], h[i], l[i]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct Params {
  const float* X;      // [B, K, d]
  const float* adj;    // [B, 2, K, K], plane 0 diagonal
  const float* Wp;     // packed fp32 weights [dff, 2d + 4]: the two bias columns are read from here
  float* Y;            // [B, K, dff] or NULL
  __half* split_out;   // [B*K, 2*split_kp] = [hi | lo] of Y, or NULL
  int split_kp;
  int K, d, dff, NS, k16, Kp;
  float out_scale;     // 1 / (power-of-two scale of the split weights)
  long long* trace;    // optional [trace_n][32] clock64 stamps of CTA phases (ec_gcn_fused_set_trace)
  int trace_n;
  int dbg;             // experiments (ec_gcn_fused_set_debug; results are wrong when != 0): 1 = no global loads of A1 / X, 2 = no stores
};

// barrier indices (8 bytes each)
enum { W_FULL = 0, W_EMPTY = 2, B_FILL = 4, B_D1 = 5, B_ZA = 6, B_G2A = 7, B_ZB = 11, B_ACC = 15, NUM_BARS = 16 };

struct Layout {
  uint32_t ts, xb_bytes, a1_bytes, wst, total;
};
__host__ __device__ inline Layout make_layout(int k16, int d, int NS) {
  Layout L;
  L.ts = (uint32_t)k16 * 128u;
  const int ng = d / 64, kbs = (k16 + 63) / 64;
  L.xb_bytes = 2u * ng * L.ts;
  L.wst = (uint32_t)NS * 256u;                    // hi + lo tiles of NS rows x 128 B
  const uint32_t a1 = 2u * kbs * L.ts;
  L.a1_bytes = a1 > L.wst ? a1 : L.wst;            // the A1 region is recycled as W stage 1
  const uint32_t epi = (WORKERS / 32) * 32 * EPI_LD * 4;   // the epilogue staging tiles reuse the operand regions
  if (L.xb_bytes + L.a1_bytes + L.wst < epi) L.a1_bytes = epi - L.xb_bytes - L.wst;
  // + barriers / TMEM slot (256 B) + rs, a0 (128 floats each) + b0, b1 (256 floats each)
  L.total = L.xb_bytes + L.a1_bytes + L.wst + 256u + (128u + 128u + 256u + 256u) * 4u;
  return L;
}

__global__ void __launch_bounds__(THREADS, 1) gcn_fused_kernel(const __grid_constant__ CUtensorMap tmW, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int K = p.K, d = p.d, NS = p.NS, k16 = p.k16;
  const int ng = d / 64, kbs = (k16 + 63) / 64, ksteps1 = k16 / 16;
  const Layout L = make_layout(k16, d, NS);
  const uint32_t TS = L.ts;
  const uint32_t xb = base, a1 = xb + L.xb_bytes, w0 = a1 + L.a1_bytes, misc = w0 + L.wst;
  auto bar = [&](int i) { return misc + 8u * i; };
  const uint32_t tmem_slot = misc + 8u * NUM_BARS;
  int* flag = reinterpret_cast<int*>(gbase + (tmem_slot - base) + 8);     // "some a0 != 1": the rescale pass is needed
  float* rs = reinterpret_cast<float*>(gbase + (misc - base) + 256);
  float* a0s = rs + 128;
  float* b0s = a0s + 128;
  float* b1s = b0s + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, n0 = blockIdx.x * NS;
  const int cta = blockIdx.y * gridDim.x + blockIdx.x;
  auto stamp = [&](int i) {
    if (p.trace && cta < p.trace_n) p.trace[cta * 32 + i] = clock64();
  };

  if (threadIdx.x == 0) {
    *flag = 0;
    mbar_init(bar(W_FULL + 0), 1); mbar_init(bar(W_FULL + 1), 1);
    mbar_init(bar(W_EMPTY + 0), 1); mbar_init(bar(W_EMPTY + 1), 1);
    mbar_init(bar(B_FILL), WORKERS / 32);
    mbar_init(bar(B_D1), 1);
    mbar_init(bar(B_ZA), WORKERS / 32);
    for (int g = 0; g < MAX_G; ++g) { mbar_init(bar(B_G2A + g), 1); mbar_init(bar(B_ZB + g), WORKERS / 32); }
    mbar_init(bar(B_ACC), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WORKERS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- workers: every global load of the fill is issued BEFORE the CTA-wide setup barrier, so the first-touch
  // latency overlaps the TMEM allocation and the barrier (values are consumed in the worker branch below)
  const int tid = threadIdx.x;
  const bool worker = warp < WORKERS / 32;
  const long long KK = (long long)K * K;
  const float* a0p = p.adj + (long long)b * 2 * KK;
  const float* a1p = a0p + KK;
  const float* Xb = p.X + (long long)b * K * d;
  // Work items are (row, 8-column chunk) pairs.  A thread keeps its chunk and steps over rows by a multiple of 8,
  // so the swizzle term (row & 7) is loop invariant: one base address per thread, constant strides after that.
  const int ncx = d / 8;                                           // chunks per X row: 8, 16 or 32
  const int xs = 31 - __clz(ncx);
  const int xv0 = tid >> xs, xc = tid & (ncx - 1), xstep = WORKERS >> xs;   // first row, chunk, rows per step
  constexpr int X_IT = 8;                                          // 128 rows x 32 chunks / 512 threads
  const float* xsrc = Xb + (long long)xv0 * d + xc * 8;
  const long long xstride = (long long)xstep * d;
  float4 q[X_IT][2];
  auto load_x = [&](int u0, int u1) {
#pragma unroll
    for (int u = u0; u < u1; ++u) {
      q[u][0] = q[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (xv0 + u * xstep < K && !(p.dbg & 1)) {
        const float4* src = reinterpret_cast<const float4*>(xsrc + u * xstride);
        q[u][0] = __ldg(src);
        q[u][1] = __ldg(src + 1);
      }
    }
  };
  // A1: 16 lanes per row (one 8-column chunk each), 32 rows per step
  const int ac = tid & 15, aw0 = tid >> 4;
  constexpr int A1_IT = (128 * 16 + WORKERS - 1) / WORKERS;       // K <= 128
  float av[A1_IT][8];
  if (worker) {
    pdl_wait();
    const bool vec = (K & 3) == 0;
    const float* asrc = a1p + (long long)aw0 * K + ac * 8;
#pragma unroll
    for (int i = 0; i < A1_IT; ++i) {
#pragma unroll
      for (int u = 0; u < 8; ++u) av[i][u] = 0.f;
      if (aw0 + 32 * i < K && !(p.dbg & 1)) {
        const float* src = asrc + (long long)(32 * i) * K;
        if (vec) {
          if (ac * 8 < K) {
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(src));
            av[i][0] = q0.x; av[i][1] = q0.y; av[i][2] = q0.z; av[i][3] = q0.w;
          }
          if (ac * 8 + 4 < K) {
            const float4 q1 = __ldg(reinterpret_cast<const float4*>(src + 4));
            av[i][4] = q1.x; av[i][5] = q1.y; av[i][6] = q1.z; av[i][7] = q1.w;
          }
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (ac * 8 + u < K) av[i][u] = __ldg(src + u);
        }
      }
    }
    load_x(0, X_IT / 2);          // first half of the X loads, behind the A1 loads and ahead of any use
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  pdl_launch_dependents();
  if (!worker) pdl_wait();

  if (warp == WORKERS / 32 + 1) {
    // ------------------------------------------------------------------ TMA producer: the W ring
    if (elect_one()) {
      const int nkb = 2 * ng;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        if (kb == 1) mbar_wait(bar(B_D1), 0);                              // stage 1 lives where A1 was
        if (kb >= 2) mbar_wait(bar(W_EMPTY + s), (uint32_t)((kb >> 1) - 1) & 1u);
        const uint32_t dst = s ? a1 : w0;
        mbar_expect_tx(bar(W_FULL + s), L.wst);
        tma_load_2d(dst, &tmW, bar(W_FULL + s), kb * 64, n0);
        tma_load_2d(dst + (uint32_t)NS * 128u, &tmW, bar(W_FULL + s), p.Kp + kb * 64, n0);
      }
    }
  } else if (warp == WORKERS / 32) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      mbar_wait(bar(B_FILL), 0);
      tc_fence_after();
      stamp(12);
      {
        // GEMM 1: D1[128 x d] = A1 (K-major, k16 deep) . X (MN-major rows v, d columns in ng tiles TS apart)
        const uint32_t idesc1 = make_idesc(d) | (1u << 16);
        const uint64_t x_hi = make_desc_mn(xb, TS), x_lo = make_desc_mn(xb + (uint32_t)ng * TS, TS);
        for (int k = 0; k < ksteps1; ++k)
          umma(tmem_base, make_desc(a1 + (uint32_t)(kbs + (k >> 2)) * TS) + 2 * (k & 3), x_hi + 128 * k, idesc1, k ? 1u : 0u);
        for (int k = 0; k < ksteps1; ++k)
          umma(tmem_base, make_desc(a1 + (uint32_t)(k >> 2) * TS) + 2 * (k & 3), x_lo + 128 * k, idesc1, 1u);
        for (int k = 0; k < ksteps1; ++k)
          umma(tmem_base, make_desc(a1 + (uint32_t)(k >> 2) * TS) + 2 * (k & 3), x_hi + 128 * k, idesc1, 1u);
        umma_commit(bar(B_D1));
      }
      stamp(13);
      mbar_wait(bar(B_ZA), 0);
      tc_fence_after();
      stamp(14);
      const uint32_t idesc2 = make_idesc(NS);
      const int nkb = 2 * ng;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1, g = kb < ng ? kb : kb - ng;
        mbar_wait(bar(W_FULL + s), (uint32_t)(kb >> 1) & 1u);
        if (kb >= ng) mbar_wait(bar(B_ZB + g), 0);
        tc_fence_after();
        stamp(15 + 2 * kb);
        const uint64_t a_hi = make_desc(xb + (uint32_t)g * TS), a_lo = make_desc(xb + (uint32_t)(ng + g) * TS);
        const uint32_t wb = s ? a1 : w0;
        const uint64_t b_hi = make_desc(wb), b_lo = make_desc(wb + (uint32_t)NS * 128u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma(tmem_base + ACC_COL, a_lo + 2 * k, b_hi + 2 * k, idesc2, (kb | k) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma(tmem_base + ACC_COL, a_hi + 2 * k, b_lo + 2 * k, idesc2, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma(tmem_base + ACC_COL, a_hi + 2 * k, b_hi + 2 * k, idesc2, 1u);
        umma_commit(bar(W_EMPTY + s));
        if (kb < ng) umma_commit(bar(B_G2A + g));          // X tile g may be overwritten by (A1 X) tile g
        stamp(16 + 2 * kb);
      }
      umma_commit(bar(B_ACC));
    }
  } else {
    // ------------------------------------------------------------------ workers (16 warps)
    const int ldw = 2 * d + 4;
    if (tid == 0) stamp(0);
    // A1 -> K-major tiles (hi tiles [0, kbs), lo tiles [kbs, 2 kbs)), columns >= K zero; row sums by shuffles
    {
      const int nc = k16 / 8;
      const int c = ac, w0 = aw0;
      float (&v)[A1_IT][8] = av;
      uint8_t* adst = gbase + (a1 - base) + (uint32_t)(c >> 3) * TS + swz(w0, c & 7);
#pragma unroll
      for (int i = 0; i < A1_IT; ++i) {
        const int w = w0 + 32 * i;
        float sacc = ((v[i][0] + v[i][1]) + (v[i][2] + v[i][3])) + ((v[i][4] + v[i][5]) + (v[i][6] + v[i][7]));
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
        if (w < K) {
          if (c == 0) rs[w] = sacc;
          if (c < nc) {
            uint4 hi, lo;
            split8(v[i], hi, lo);
            *reinterpret_cast<uint4*>(adst + i * (32 * 128)) = hi;
            *reinterpret_cast<uint4*>(adst + i * (32 * 128) + (uint32_t)kbs * TS) = lo;
          }
        }
      }
    }
    if (tid == 0) stamp(2);
    // X -> [v][64-column] tiles (hi tiles [0, ng), lo tiles [ng, 2 ng)), rows [K, k16) zero; up to 8 items per
    // thread, the second half of the loads in flight while the first half is converted
    load_x(X_IT / 2, X_IT);
    {
      uint8_t* xdst = gbase + (xb - base) + (uint32_t)(xc >> 3) * TS + swz(xv0, xc & 7);
      const uint32_t lo_off = (uint32_t)ng * TS;
#pragma unroll
      for (int u = 0; u < X_IT; ++u) {
        if (xv0 + u * xstep >= k16) continue;
        const float f[8] = {q[u][0].x, q[u][0].y, q[u][0].z, q[u][0].w, q[u][1].x, q[u][1].y, q[u][1].z, q[u][1].w};
        uint4 hi, lo;
        split8(f, hi, lo);
        *reinterpret_cast<uint4*>(xdst + (uint32_t)(u * xstep) * 128u) = hi;
        *reinterpret_cast<uint4*>(xdst + (uint32_t)(u * xstep) * 128u + lo_off) = lo;
      }
    }
    proxy_fence();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(B_FILL));
    if (tid == 0) stamp(3);
    // strided gathers (bias columns of Wp, diagonal of plane 0): their latency hides behind GEMM 1
    for (int i = tid; i < NS; i += WORKERS) {
      b0s[i] = __ldg(p.Wp + (long long)(n0 + i) * ldw + 2 * d);
      b1s[i] = __ldg(p.Wp + (long long)(n0 + i) * ldw + 2 * d + 1);
    }
    if (tid < K) {                                   // does any row need the rescale pass?
      const float a = __ldg(a0p + (long long)tid * K + tid);
      a0s[tid] = a;
      if (a != 1.0f) atomicOr(flag, 1);
    }

    // ---- GEMM 1 done: rescale the X tiles by a0[w] in place (skipped when every a0 is 1, the usual case)
    mbar_wait(bar(B_D1), 0);
    tc_fence_after();
    if (tid == 0) stamp(4);
    asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory");     // rs / a0s / flag of every warp are visible
    if (*reinterpret_cast<volatile int*>(flag)) {
      for (int it = tid; it < K * ncx; it += WORKERS) {
        const int v = it / ncx, c = it - v * ncx;
        const float a0v = a0s[v];
        if (a0v == 1.0f) continue;
        const uint32_t off = (uint32_t)(c >> 3) * TS + swz(v, c & 7);
        uint4* ph = reinterpret_cast<uint4*>(gbase + (xb - base) + off);
        uint4* pl = reinterpret_cast<uint4*>(gbase + (xb - base) + (uint32_t)ng * TS + off);
        const uint4 h = *ph, l = *pl;
        const __half2* hh = reinterpret_cast<const __half2*>(&h);
        const __half2* ll = reinterpret_cast<const __half2*>(&l);
        float f[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 fh = __half22float2(hh[u]), fl = __half22float2(ll[u]);
          f[2 * u] = a0v * (fh.x + fl.x);
          f[2 * u + 1] = a0v * (fh.y + fl.y);
        }
        uint4 hi, lo;
        split8(f, hi, lo);
        *ph = hi;
        *pl = lo;
      }
      proxy_fence();
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(B_ZA));
    if (tid == 0) stamp(5);

    // ---- drain D1 = A1 X group by group into the X tile whose k-block of GEMM 2 has retired
    const int quarter = warp & 3, part = warp >> 2;                 // TMEM lane quarter; 16-column part of a group
    const int row = quarter * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int g = 0; g < ng; ++g) {
      float r[16];
      tmem_ld16(t_row + (uint32_t)(g * 64 + part * 16), r);
      mbar_wait(bar(B_G2A + g), 0);
      if (row < K) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          uint4 hi, lo;
          split8(r + 8 * i, hi, lo);
          const uint32_t off = (uint32_t)g * TS + swz(row, part * 2 + i);
          *reinterpret_cast<uint4*>(gbase + (xb - base) + off) = hi;
          *reinterpret_cast<uint4*>(gbase + (xb - base) + (uint32_t)ng * TS + off) = lo;
        }
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_ZB + g));
      if (tid == 0) stamp(6 + g);
    }

    // ---- epilogue: this warp owns rows [32 quarter, +32) and columns [part NS/4, +NS/4) of the slice.  Each
    // 16-column sub-chunk is transposed through a per-warp staging tile (the X tiles are dead by now) so that the
    // global stores are row-contiguous: 4 lanes cover a 64 B row segment, a warp 8 rows per instruction.
    mbar_wait(bar(B_ACC), 0);
    tc_fence_after();
    if (tid == 0) stamp(10);
    float* stg = reinterpret_cast<float*>(gbase + (xb - base)) + warp * (32 * EPI_LD);
    const int sub_row = lane >> 2, c4 = (lane & 3) * 4;
    const int cw = NS / 4, nsub = cw / 16;           // nsub <= 3 (NS <= 192)
    float r[3][16];
#pragma unroll
    for (int sc = 0; sc < 3; ++sc)
      if (sc < nsub) tmem_ld16_nowait(t_row + (uint32_t)(ACC_COL + part * cw + sc * 16), r[sc]);
    float a0r[4], rsr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int orow = quarter * 32 + i * 8 + sub_row;
      a0r[i] = orow < K ? a0s[orow] : 0.f;
      rsr[i] = orow < K ? rs[orow] : 0.f;
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int sc = 0; sc < 3; ++sc) {
      if (sc >= nsub) break;
      const int col0 = part * cw + sc * 16;
#pragma unroll
      for (int j = 0; j < 16; j += 4)
        *reinterpret_cast<float4*>(stg + lane * EPI_LD + j) = make_float4(r[sc][j], r[sc][j + 1], r[sc][j + 2], r[sc][j + 3]);
      __syncwarp();
      const int gcol = col0 + c4;
      const float4 bb0 = *reinterpret_cast<const float4*>(b0s + gcol), bb1 = *reinterpret_cast<const float4*>(b1s + gcol);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = i * 8 + sub_row, orow = quarter * 32 + rr;
        if (orow >= K) continue;
        const float4 a4 = *reinterpret_cast<const float4*>(stg + rr * EPI_LD + c4);
        const float a0v = a0r[i], rsv = rsr[i];
        float y[4];
        y[0] = fmaxf(fmaf(a4.x, p.out_scale, fmaf(a0v, bb0.x, rsv * bb1.x)), 0.f);
        y[1] = fmaxf(fmaf(a4.y, p.out_scale, fmaf(a0v, bb0.y, rsv * bb1.y)), 0.f);
        y[2] = fmaxf(fmaf(a4.z, p.out_scale, fmaf(a0v, bb0.z, rsv * bb1.z)), 0.f);
        y[3] = fmaxf(fmaf(a4.w, p.out_scale, fmaf(a0v, bb0.w, rsv * bb1.w)), 0.f);
        const long long grow = (long long)b * K + orow;
        if (p.dbg & 2) continue;
        if (p.Y) *reinterpret_cast<float4*>(p.Y + grow * p.dff + n0 + gcol) = make_float4(y[0], y[1], y[2], y[3]);
        if (p.split_out) {
          uint32_t h01, l01, h23, l23;
          split_pair(y[0], y[1], h01, l01);
          split_pair(y[2], y[3], h23, l23);
          __half* sp = p.split_out + grow * (2LL * p.split_kp) + n0 + gcol;
          *reinterpret_cast<uint2*>(sp) = make_uint2(h01, h23);
          *reinterpret_cast<uint2*>(sp + p.split_kp) = make_uint2(l01, l23);
        }
      }
      __syncwarp();
    }
  }
  if (threadIdx.x == 0) stamp(11);
  tc_fence_before();
  __syncthreads();
  if (warp == WORKERS / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

constexpr uint32_t SMEM_LIMIT = 227u * 1024u;

// slice width for (K, d, dff), or 0 when the fused kernel cannot take the shape
inline int pick_slice(int K, int d, int dff) {
  if (K < 1 || K > 128 || (d != 64 && d != 128 && d != 256) || dff % 64) return 0;
  const int k16 = (K + 15) / 16 * 16;
  const int cand[3] = {192, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int NS = cand[i];
    if (dff % NS) continue;
    if (make_layout(k16, d, NS).total + 1024u <= SMEM_LIMIT) return NS;
  }
  return 0;
}

}  // namespace gf
}  // namespace ec

using namespace ec;

static long long* gf_trace = nullptr;
static int gf_trace_n = 0;
extern "C" int ec_gcn_fused_set_trace(void* buf, int n_ctas) {   // profiling hook: [n_ctas][32] int64 device buffer, or NULL
  gf_trace = (long long*)buf;
  gf_trace_n = buf ? n_ctas : 0;
  return EC_OK;
}

static int gf_debug = 0;
extern "C" int ec_gcn_fused_set_debug(int flags) {   // bring-up / profiling experiments only (results are wrong when != 0)
  gf_debug = flags;
  return EC_OK;
}

extern "C" int ec_gcn_fused_slice(int K, int d, int dff) { return gf::pick_slice(K, d, dff); }

extern "C" int ec_gcn_fused(const float* X, const float* adj, const float* Wp, const void* W2, int Kp, float w_scale,
                            float* Y, void* split_out, int split_kp, int B, int K, int d, int dff, void* stream) {
  EC_REQUIRE(X && adj && Wp && W2 && (Y || split_out), "ec_gcn_fused: null pointer");
  const int NS = gf::pick_slice(K, d, dff);
  EC_REQUIRE(NS > 0, "ec_gcn_fused: unsupported shape (K <= 128, d in {64,128,256}, dff %% 64 == 0, shared memory)");
  EC_REQUIRE(Kp % 64 == 0 && Kp >= 2 * d, "ec_gcn_fused: Kp must be a multiple of 64 and >= 2d");
  EC_REQUIRE(w_scale > 0.f, "ec_gcn_fused: bad weight scale");
  EC_REQUIRE(aligned16(X) && aligned16(adj) && aligned16(W2) && (!Y || aligned16(Y)) && (!split_out || aligned16(split_out)),
             "ec_gcn_fused: operands must be 16-byte aligned");
  EC_REQUIRE(!split_out || (split_kp % 8 == 0 && split_kp >= dff), "ec_gcn_fused: bad split_kp");
  if (B == 0) return EC_OK;
  const int k16 = (K + 15) / 16 * 16;
  const uint32_t smem = gf::make_layout(k16, d, NS).total + 1024u;
  EC_CUDA((cudaError_t)ensure_dynamic_smem(gf::gcn_fused_kernel, (int)gf::SMEM_LIMIT));
  CUtensorMap tmW;
  int rc = tc::get_tensor_map(W2, dff, Kp, NS, &tmW);
  if (rc) return rc;
  gf::Params p;
  p.X = X; p.adj = adj; p.Wp = Wp; p.Y = Y; p.split_out = (__half*)split_out; p.split_kp = split_kp;
  p.K = K; p.d = d; p.dff = dff; p.NS = NS; p.k16 = k16; p.Kp = Kp; p.out_scale = 1.0f / w_scale;
  p.trace = gf_trace; p.trace_n = gf_trace_n; p.dbg = gf_debug;
  launch_pdl(gf::gcn_fused_kernel, dim3(dff / NS, B), dim3(gf::THREADS), (size_t)smem, (cudaStream_t)stream, tmW, p);
  return check_launch("ec_gcn_fused");
}
