// Fused GCN feed-forward for sm_100a (encoder_decoder.py:508-524 of the reference, aggregate-first form):
//     Y[b,w,:] = relu( a0[b,w] * (X[b,w,:] W0^T + b0) + sum_v A1[b,w,v] (X[b,v,:] W1^T + b1) )
// as ONE kernel: no Z operand round trip through global memory, no second launch.
//
// CTA = (sample b, slice of NS output channels).  Everything runs on the tensor cores at fp32-grade accuracy
// (three fp16 products per fp32 product, see gemm_tcgen05.cu):
//   fill    8 worker warps read A1[b] and X[b] once (coalesced float4), split them into fp16 hi | lo and write
//           128B-swizzled shared-memory tiles: A1 as a K-major A operand, X as [v][64 columns] row tiles.
//   GEMM 1  D1[128 x d] = A1 . X in TMEM; the X tiles are consumed as an MN-major B operand (N = d in one UMMA).
//   scale   X tiles are rescaled in place by a0[w] (a no-op for the usual a0 = 1): the very same bytes are now
//           the K-major A operand a0*X of GEMM 2 (an MN-major [v][c] tile and a K-major [w][c] tile coincide).
//   GEMM 2  acc[128 x NS] = (a0 X) . W0^T  (k-blocks 0..d/64-1)  +  (A1 X) . W1^T  (k-blocks d/64..2d/64-1);
//           W streams through a two-stage TMA ring from the pre-split packed weights; while the first half runs
//           the workers drain D1 (tcgen05.ld), split it and write it over the X tile whose k-block has retired.
//   epilogue  relu(acc / w_scale + a0[w] b0[n] + rowsum(A1[w,:]) b1[n]) in fp32, written as fp32 rows and / or
//           as the split-fp16 A operand of the ffn2 GEMM that follows.
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

constexpr int WORKERS = 256;             // 8 worker warps
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
  float* rs = reinterpret_cast<float*>(gbase + (misc - base) + 256);
  float* a0s = rs + 128;
  float* b0s = a0s + 128;
  float* b1s = b0s + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, n0 = blockIdx.x * NS;

  if (threadIdx.x == 0) {
    mbar_init(bar(W_FULL + 0), 1); mbar_init(bar(W_FULL + 1), 1);
    mbar_init(bar(W_EMPTY + 0), 1); mbar_init(bar(W_EMPTY + 1), 1);
    mbar_init(bar(B_FILL), WORKERS / 32);
    mbar_init(bar(B_D1), 1);
    mbar_init(bar(B_ZA), WORKERS / 32);
    for (int g = 0; g < MAX_G; ++g) { mbar_init(bar(B_G2A + g), 1); mbar_init(bar(B_ZB + g), WORKERS / 32); }
    mbar_init(bar(B_ACC), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 9) {
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
  } else if (warp == 8) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      mbar_wait(bar(B_FILL), 0);
      tc_fence_after();
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
      mbar_wait(bar(B_ZA), 0);
      tc_fence_after();
      const uint32_t idesc2 = make_idesc(NS);
      const int nkb = 2 * ng;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1, g = kb < ng ? kb : kb - ng;
        mbar_wait(bar(W_FULL + s), (uint32_t)(kb >> 1) & 1u);
        if (kb >= ng) mbar_wait(bar(B_ZB + g), 0);
        tc_fence_after();
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
      }
      umma_commit(bar(B_ACC));
    }
  } else {
    // ------------------------------------------------------------------ workers (8 warps)
    const int tid = threadIdx.x;
    const long long KK = (long long)K * K;
    const float* a0p = p.adj + (long long)b * 2 * KK;
    const float* a1p = a0p + KK;
    const float* Xb = p.X + (long long)b * K * d;
    const int ldw = 2 * d + 4;
    for (int i = tid; i < NS; i += WORKERS) {
      b0s[i] = __ldg(p.Wp + (long long)(n0 + i) * ldw + 2 * d);
      b1s[i] = __ldg(p.Wp + (long long)(n0 + i) * ldw + 2 * d + 1);
    }
    // row sums of A1 (fp32, same lane order as the unfused kernel) and the diagonal of plane 0
    for (int w = warp; w < K; w += WORKERS / 32) {
      float sacc = 0.f;
      for (int v = lane; v < K; v += 32) sacc += __ldg(a1p + (long long)w * K + v);
      sacc = warp_sum(sacc);
      if (lane == 0) {
        rs[w] = sacc;
        a0s[w] = __ldg(a0p + (long long)w * K + w);
      }
    }
    // A1 -> K-major tiles (hi tiles [0, kbs), lo tiles [kbs, 2 kbs)), columns >= K zero
    {
      const int nc = k16 / 8;
      const bool vec = (K & 3) == 0;
      for (int it = tid; it < K * nc; it += WORKERS) {
        const int w = it / nc, c = it - w * nc;
        float v[8];
        const float* src = a1p + (long long)w * K + c * 8;
        if (vec) {
          const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 q0 = (c * 8 < K) ? __ldg(reinterpret_cast<const float4*>(src)) : z;
          const float4 q1 = (c * 8 + 4 < K) ? __ldg(reinterpret_cast<const float4*>(src + 4)) : z;
          v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
        } else {
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = (c * 8 + u < K) ? __ldg(src + u) : 0.f;
        }
        uint4 hi, lo;
        split8(v, hi, lo);
        const uint32_t off = (uint32_t)(c >> 3) * TS + swz(w, c & 7);
        *reinterpret_cast<uint4*>(gbase + (a1 - base) + off) = hi;
        *reinterpret_cast<uint4*>(gbase + (a1 - base) + (uint32_t)kbs * TS + off) = lo;
      }
    }
    // X -> [v][64-column] tiles (hi tiles [0, ng), lo tiles [ng, 2 ng)), rows [K, k16) zero; 4 items in flight
    const int ncx = d / 8;
    {
      const int items = k16 * ncx;
      for (int it0 = tid; it0 < items; it0 += 4 * WORKERS) {
        float4 q[4][2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int it = it0 + u * WORKERS;
          q[u][0] = q[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (it < items) {
            const int v = it / ncx, c = it - v * ncx;
            if (v < K) {
              const float4* src = reinterpret_cast<const float4*>(Xb + (long long)v * d + c * 8);
              q[u][0] = __ldg(src);
              q[u][1] = __ldg(src + 1);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int it = it0 + u * WORKERS;
          if (it >= items) continue;
          const int v = it / ncx, c = it - v * ncx;
          const float f[8] = {q[u][0].x, q[u][0].y, q[u][0].z, q[u][0].w, q[u][1].x, q[u][1].y, q[u][1].z, q[u][1].w};
          uint4 hi, lo;
          split8(f, hi, lo);
          const uint32_t off = (uint32_t)(c >> 3) * TS + swz(v, c & 7);
          *reinterpret_cast<uint4*>(gbase + (xb - base) + off) = hi;
          *reinterpret_cast<uint4*>(gbase + (xb - base) + (uint32_t)ng * TS + off) = lo;
        }
      }
    }
    proxy_fence();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(B_FILL));

    // ---- GEMM 1 done: rescale the X tiles by a0[w] in place (rows with a0 == 1 are left alone)
    mbar_wait(bar(B_D1), 0);
    tc_fence_after();
    asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory");     // rs / a0s of every warp are visible
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
    __syncwarp();
    if (lane == 0) mbar_arrive(bar(B_ZA));

    // ---- drain D1 = A1 X group by group into the X tile whose k-block of GEMM 2 has retired
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    for (int g = 0; g < ng; ++g) {
      float r[32];
      tmem_ld32(t_row + (uint32_t)(g * 64 + half * 32), r);
      mbar_wait(bar(B_G2A + g), 0);
      if (row < K) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 hi, lo;
          split8(r + 8 * i, hi, lo);
          const uint32_t off = (uint32_t)g * TS + swz(row, half * 4 + i);
          *reinterpret_cast<uint4*>(gbase + (xb - base) + off) = hi;
          *reinterpret_cast<uint4*>(gbase + (xb - base) + (uint32_t)ng * TS + off) = lo;
        }
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_ZB + g));
    }

    // ---- epilogue: this warp owns rows [32 quarter, +32) and columns [half NS/2, +NS/2) of the slice
    mbar_wait(bar(B_ACC), 0);
    tc_fence_after();
    const float a0v = row < K ? a0s[row] : 0.f, rsv = row < K ? rs[row] : 0.f;
    const int hw = NS / 2;
    for (int ch = 0; ch < hw / 32; ++ch) {
      const int col0 = half * hw + ch * 32;
      float r[32];
      tmem_ld32(t_row + (uint32_t)(ACC_COL + col0), r);
      if (row < K) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = fmaxf(fmaf(r[j], p.out_scale, fmaf(a0v, b0s[col0 + j], rsv * b1s[col0 + j])), 0.f);
        const long long grow = (long long)b * K + row;
        if (p.Y) {
          float4* yp = reinterpret_cast<float4*>(p.Y + grow * p.dff + n0 + col0);
#pragma unroll
          for (int j = 0; j < 8; ++j) yp[j] = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
        }
        if (p.split_out) {
          __half* sp = p.split_out + grow * (2LL * p.split_kp) + n0 + col0;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 hi, lo;
            split8(r + 8 * j, hi, lo);
            *reinterpret_cast<uint4*>(sp + 8 * j) = hi;
            *reinterpret_cast<uint4*>(sp + p.split_kp + 8 * j) = lo;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

constexpr uint32_t SMEM_LIMIT = 227u * 1024u;

// slice width for (K, d, dff), or 0 when the fused kernel cannot take the shape
inline int pick_slice(int K, int d, int dff) {
  if (K < 1 || K > 128 || d < 64 || d > 256 || d % 64 || dff % 64) return 0;
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

extern "C" int ec_gcn_fused_slice(int K, int d, int dff) { return gf::pick_slice(K, d, dff); }

extern "C" int ec_gcn_fused(const float* X, const float* adj, const float* Wp, const void* W2, int Kp, float w_scale,
                            float* Y, void* split_out, int split_kp, int B, int K, int d, int dff, void* stream) {
  EC_REQUIRE(X && adj && Wp && W2 && (Y || split_out), "ec_gcn_fused: null pointer");
  const int NS = gf::pick_slice(K, d, dff);
  EC_REQUIRE(NS > 0, "ec_gcn_fused: unsupported shape (K <= 128, d in {64,128,192,256}, dff %% 64 == 0, shared memory)");
  EC_REQUIRE(Kp % 64 == 0 && Kp >= 2 * d, "ec_gcn_fused: Kp must be a multiple of 64 and >= 2d");
  EC_REQUIRE(w_scale > 0.f, "ec_gcn_fused: bad weight scale");
  EC_REQUIRE(aligned16(X) && aligned16(adj) && aligned16(W2) && (!Y || aligned16(Y)) && (!split_out || aligned16(split_out)),
             "ec_gcn_fused: operands must be 16-byte aligned");
  EC_REQUIRE(!split_out || (split_kp % 8 == 0 && split_kp >= dff), "ec_gcn_fused: bad split_kp");
  if (B == 0) return EC_OK;
  const int k16 = (K + 15) / 16 * 16;
  const uint32_t smem = gf::make_layout(k16, d, NS).total + 1024u;
  static bool attr_set = false;
  if (!attr_set) {
    EC_CUDA(cudaFuncSetAttribute(gf::gcn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gf::SMEM_LIMIT));
    attr_set = true;
  }
  CUtensorMap tmW;
  int rc = tc::get_tensor_map(W2, dff, Kp, NS, &tmW);
  if (rc) return rc;
  gf::Params p;
  p.X = X; p.adj = adj; p.Wp = Wp; p.Y = Y; p.split_out = (__half*)split_out; p.split_kp = split_kp;
  p.K = K; p.d = d; p.dff = dff; p.NS = NS; p.k16 = k16; p.Kp = Kp; p.out_scale = 1.0f / w_scale;
  launch_pdl(gf::gcn_fused_kernel, dim3(dff / NS, B), dim3(gf::THREADS), (size_t)smem, (cudaStream_t)stream, tmW, p);
  return check_launch("ec_gcn_fused");
}
