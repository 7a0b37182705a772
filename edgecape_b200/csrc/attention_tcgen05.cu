// Tensor-core attention for sm_100a (head dim 64, no mask / bias: the DINOv2 block attention and the
// decoder / skeleton cross-attentions).  One CTA = one (batch, head, 128-query tile):
//
//   stage   Q tile and all keys as split-fp16 (hi, lo) tiles in 128B-swizzled shared memory
//   S       = Q K^T on tcgen05 (q_lo k_hi + q_hi k_lo + q_hi k_hi), fp32, the whole row block [128 x Lk]
//           stays in TMEM (Lk <= 448 columns) -- no online-softmax rescaling is needed
//   softmax row max from TMEM (tcgen05.ld, one lane = one row), then per 64-key chunk p = exp(s - max)
//           is written back to shared memory as split-fp16 P tiles while V^T of the same chunk is staged
//   O      += P V on tcgen05 (p_lo v_hi + p_hi v_lo + p_hi v_hi) into TMEM; chunks are double buffered so
//           the exponentials of chunk i+1 overlap the MMAs of chunk i
//   out     O / rowsum, as fp32 and / or directly in the split-fp16 form the next GEMM consumes.
//
// fp32-grade by construction: every operand is represented to 2^-22 and accumulation is fp32.
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"
#include "tma.cuh"

namespace ec {
namespace atc {

constexpr int D = 64;
constexpr int BM = 128;                 // query rows per CTA (UMMA M)
constexpr int KC = 64;                  // keys per P/V chunk (one 128-byte swizzle span of fp16)
constexpr int THREADS = 512;            // 16 warps: 4 per TMEM lane quarter (latency hiding for the softmax / staging work)
constexpr int NPART = THREADS / 128;    // warps sharing one lane quarter split the keys NPART ways
constexpr int PW = KC / NPART;          // keys of a chunk handled by one warp (16)
constexpr int MAX_LKP = 448;            // S columns in TMEM; O uses columns [448, 512)
constexpr int O_COL = 448;
constexpr int Q_BYTES = 2 * BM * 128;   // Q_hi + Q_lo
constexpr int PBUF_BYTES = 2 * BM * 128;      // P_hi + P_lo of one chunk (32 KB)
constexpr int VBUF_BYTES = 2 * D * 128;       // V_hi^T + V_lo^T of one chunk (16 KB)
constexpr int REUSE_BYTES = 2 * PBUF_BYTES + 2 * VBUF_BYTES;   // 96 KB, aliases Q and K after S is done
constexpr int MISC_BYTES = 64 + 2 * NPART * BM * 4 + 512;   // barriers + tmem slot, row max / row sum exchange, key mask

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {   // K-major, SWIZZLE_128B, SBO = 1024 B
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {           // f16 x f16 -> f32, M = 128, N = n
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
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

// byte offset of the 16-byte chunk `c` (8 fp16) of row `row` inside a 128B-swizzled K-major tile
__device__ __forceinline__ uint32_t swz(int row, int c) { return (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4)); }

// 8 fp32 -> 8 hi + 8 lo fp16, packed as two 16-byte vectors
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_pair(v[2 * i], v[2 * i + 1], h[i], l[i]);
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

struct Params {
  const float *Q, *K, *V;
  float* O;
  int B, H, Lq, Lk, LKP;
  int ldq, ldk, ldv, ldo;
  long long sq, sk, sv, so;
  float scale;
  __half* split_out;
  int split_kp;
  int dv;                      // valid head dim (32 or 64); the tiles are zero padded to 64
  const uint8_t* key_mask;     // [B, Lk], 1 = ignore key, or NULL
  const float* bias;           // [B, H, Lq, Lk] additive, or NULL
};

constexpr int THREADS_T = THREADS + 32;   // 16 softmax / staging warps + one control warp (MMA issue; TMA in the split variant)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void softmax_sync() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {   // MN-major B operand, SWIZZLE_128B
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc_bmn(int n) {       // as make_idesc, B operand MN-major
  return make_idesc(n) | (1u << 16);
}

// Warps 0..15 stage operands and do the softmax; lane 0 of warp 16 issues every tcgen05.mma and its commits
// (bar_pr*: "P and V of this chunk are in shared memory", 16 warp arrivals; bar_pv*: "the MMAs reading that
// buffer are done"), so MMA issue never sits on the softmax warps' critical path.
__global__ void __launch_bounds__(THREADS_T, 1) attention_tc_kernel(Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int data_bytes = max(Q_BYTES + 2 * p.LKP * 128, REUSE_BYTES);
  const uint32_t misc = base + data_bytes;
  const uint32_t bar_s = misc, bar_pv0 = misc + 8, bar_pv1 = misc + 16, bar_pr0 = misc + 24, bar_pr1 = misc + 32,
                 tmem_slot = misc + 40;
  float* xmax = reinterpret_cast<float*>(gbase + data_bytes + 64);   // [NPART][BM]
  float* xsum = xmax + NPART * BM;                                    // [NPART][BM]
  uint8_t* msk = reinterpret_cast<uint8_t*>(xsum + NPART * BM);       // [MAX_LKP] key-padding mask of this batch row

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, part = (warp >> 2) & 3;
  const bool control = warp == THREADS / 32;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BM;
  const int dsteps = (p.dv + 15) / 16;                  // k-steps of S = Q K^T that carry data (head dim 32: 2 of 4)

  if (warp == THREADS / 32 && elect_one()) {
    mbar_init(bar_s, 1);
    mbar_init(bar_pv0, 1);
    mbar_init(bar_pv1, 1);
    mbar_init(bar_pr0, THREADS / 32);
    mbar_init(bar_pr1, THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }

  // ------------------------------------------------------------------ stage Q (scaled) and K
  const uint32_t q_hi = base, q_lo = base + BM * 128, k_hi = base + Q_BYTES, k_lo = k_hi + p.LKP * 128;
  if (!control) {
    for (int j = tid; j < p.LKP; j += THREADS) msk[j] = (j < p.Lk && p.key_mask) ? p.key_mask[(long long)b * p.Lk + j] : 0;
    const float* Qb = p.Q + (long long)b * p.sq + h * p.dv;
    for (int it = tid; it < BM * 8; it += THREADS) {
      const int row = it >> 3, c = it & 7;
      float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (q0 + row < p.Lq && c * 8 < p.dv) {
        const float4* src = reinterpret_cast<const float4*>(Qb + (long long)(q0 + row) * p.ldq + c * 8);
        const float4 a = __ldg(src), bb = __ldg(src + 1);
        v[0] = a.x * p.scale; v[1] = a.y * p.scale; v[2] = a.z * p.scale; v[3] = a.w * p.scale;
        v[4] = bb.x * p.scale; v[5] = bb.y * p.scale; v[6] = bb.z * p.scale; v[7] = bb.w * p.scale;
      }
      uint4 hi, lo;
      split8(v, hi, lo);
      const uint32_t off = swz(row, c);
      *reinterpret_cast<uint4*>(gbase + (q_hi - base) + off) = hi;
      *reinterpret_cast<uint4*>(gbase + (q_lo - base) + off) = lo;
    }
    const float* Kb = p.K + (long long)b * p.sk + h * p.dv;
    // 4 items (= 8 float4 loads) in flight per thread before any conversion: hides the global latency
    for (int it0 = tid; it0 < p.LKP * 8; it0 += 4 * THREADS) {
      float4 ra[4], rb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int it = it0 + u * THREADS;
        const int row = it >> 3, c = it & 7;
        ra[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        rb[u] = ra[u];
        if (it < p.LKP * 8 && row < p.Lk && c * 8 < p.dv) {
          const float4* src = reinterpret_cast<const float4*>(Kb + (long long)row * p.ldk + c * 8);
          ra[u] = __ldg(src);
          rb[u] = __ldg(src + 1);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int it = it0 + u * THREADS;
        if (it >= p.LKP * 8) continue;
        const int row = it >> 3, c = it & 7;
        const float v[8] = {ra[u].x, ra[u].y, ra[u].z, ra[u].w, rb[u].x, rb[u].y, rb[u].z, rb[u].w};
        uint4 hi, lo;
        split8(v, hi, lo);
        const uint32_t off = swz(row, c);
        *reinterpret_cast<uint4*>(gbase + (k_hi - base) + off) = hi;
        *reinterpret_cast<uint4*>(gbase + (k_lo - base) + off) = lo;
      }
    }
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  const int nchunks = (p.Lk + KC - 1) / KC;

  if (control) {
    // ================================================================== control warp
    if (elect_one()) {
      // S = Q K^T  (all keys, into TMEM)
      const int n0 = p.LKP <= 256 ? p.LKP : ((p.LKP / 2 + 15) / 16) * 16;
      for (int noff = 0; noff < p.LKP; noff += n0) {
        const int n = min(n0, p.LKP - noff);
        const uint32_t idesc = make_idesc(n);
        const uint64_t aq_hi = make_desc(q_hi), aq_lo = make_desc(q_lo);
        const uint64_t bk_hi = make_desc(k_hi + noff * 128), bk_lo = make_desc(k_lo + noff * 128);
        for (int k = 0; k < dsteps; ++k) umma(tmem_base + noff, aq_lo + 2 * k, bk_hi + 2 * k, idesc, k ? 1u : 0u);
        for (int k = 0; k < dsteps; ++k) umma(tmem_base + noff, aq_hi + 2 * k, bk_lo + 2 * k, idesc, 1u);
        for (int k = 0; k < dsteps; ++k) umma(tmem_base + noff, aq_hi + 2 * k, bk_hi + 2 * k, idesc, 1u);
      }
      umma_commit(bar_s);
      // O += P V per chunk; N = the head dim actually used (32 or 64)
      const uint32_t idesc_o = make_idesc_bmn(dsteps * 16);
      for (int i = 0; i < nchunks; ++i) {
        const int buf = i & 1;
        const uint32_t p_hi = base + buf * PBUF_BYTES, p_lo = p_hi + BM * 128;
        const uint32_t v_hi = base + 2 * PBUF_BYTES + buf * VBUF_BYTES, v_lo = v_hi + KC * 128;
        mbar_wait(buf ? bar_pr1 : bar_pr0, (i >> 1) & 1);
        tc_fence_after();
        const int valid = min(KC, p.Lk - i * KC);
        const int ksteps = (valid + 15) / 16;
        const uint64_t ap_hi = make_desc(p_hi), ap_lo = make_desc(p_lo);
        const uint64_t bv_hi = make_desc_mn(v_hi), bv_lo = make_desc_mn(v_lo);
        // P: K-major, +32 B per 16-key step; V: MN-major [key][d] rows, 16 keys = two 1024 B atoms per step
        for (int k = 0; k < ksteps; ++k) umma(tmem_base + O_COL, ap_lo + 2 * k, bv_hi + 128 * k, idesc_o, (i | k) ? 1u : 0u);
        for (int k = 0; k < ksteps; ++k) umma(tmem_base + O_COL, ap_hi + 2 * k, bv_lo + 128 * k, idesc_o, 1u);
        for (int k = 0; k < ksteps; ++k) umma(tmem_base + O_COL, ap_hi + 2 * k, bv_hi + 128 * k, idesc_o, 1u);
        umma_commit(buf ? bar_pv1 : bar_pv0);
      }
    }
    __syncwarp();
  } else {
    // ================================================================== softmax / staging warps
    mbar_wait(bar_s, 0);
    tc_fence_after();
    // ---------------------------------------------------------------- row max (four warps per lane quarter)
    const int row = quarter * 32 + lane;               // TMEM lane == row of the query tile
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int nchunk32 = (p.Lk + 31) / 32;
    const float* bias_row = (p.bias && q0 + row < p.Lq)
                                ? p.bias + (((long long)b * p.H + h) * p.Lq + (q0 + row)) * p.Lk : nullptr;
    float mymax = -INFINITY;
    for (int j = part; j < nchunk32; j += NPART) {
      float s[32];
      tmem_ld32(t_row + j * 32, s);
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const int key = j * 32 + u;
        if (key < p.Lk && !msk[key]) mymax = fmaxf(mymax, bias_row ? s[u] + __ldg(bias_row + key) : s[u]);
      }
    }
    xmax[part * BM + row] = mymax;
    softmax_sync();
    float rmax = xmax[row];
#pragma unroll
    for (int q = 1; q < NPART; ++q) rmax = fmaxf(rmax, xmax[q * BM + row]);
    if (rmax == -INFINITY) rmax = 0.f;                 // fully masked row: every p below is 0

    // ---------------------------------------------------------------- P = exp(S - max) and V, 64 keys per chunk
    const float* Vb = p.V + (long long)b * p.sv + h * p.dv;
    const float L2E = 1.4426950408889634f;
    const float nb = -rmax * L2E;
    float rsum = 0.f;
    constexpr int VI = KC * 8 / THREADS;               // V items (8 floats of one key) per thread and chunk
    float4 va[VI], vb[VI];                             // V rows of the next chunk, prefetched
    auto load_v = [&](int chunk) {
#pragma unroll
      for (int w = 0; w < VI; ++w) {
        const int it = tid + w * THREADS;
        const int kl = it >> 3, dc = it & 7;
        const int key = chunk * KC + kl;
        va[w] = make_float4(0.f, 0.f, 0.f, 0.f);
        vb[w] = va[w];
        if (key < p.Lk && dc * 8 < p.dv) {
          const float4* src = reinterpret_cast<const float4*>(Vb + (long long)key * p.ldv + dc * 8);
          va[w] = __ldg(src);
          vb[w] = __ldg(src + 1);
        }
      }
    };
    load_v(0);
    for (int i = 0; i < nchunks; ++i) {
      const int buf = i & 1;
      const uint32_t p_off = buf * PBUF_BYTES;
      const uint32_t v_off = 2 * PBUF_BYTES + buf * VBUF_BYTES;
      // P of this chunk: this warp covers keys [i*64 + PW*part, +PW) of its 32 rows
      float s[PW];
      const int kbase = i * KC + PW * part;
      tmem_ld16(t_row + kbase, s);
#pragma unroll
      for (int u = 0; u < PW; ++u) {
        const int key = kbase + u;
        const bool ok = key < p.Lk && !msk[key];
        const float sv_ = (ok && bias_row) ? s[u] + __ldg(bias_row + key) : s[u];
        const float e = ok ? ex2(fmaf(sv_, L2E, nb)) : 0.f;
        s[u] = e;
        rsum += e;
      }
      if (i >= 2) mbar_wait(buf ? bar_pv1 : bar_pv0, ((i >> 1) - 1) & 1);   // MMAs of chunk i-2 released this buffer
      // V of this chunk as an MN-major tile ([key][d] rows, 128B swizzle -- the layout K uses), from the
      // registers prefetched one iteration ago; then the loads of the next chunk are issued
#pragma unroll
      for (int w = 0; w < VI; ++w) {
        const int it = tid + w * THREADS;
        const int kl = it >> 3, dc = it & 7;
        const float v[8] = {va[w].x, va[w].y, va[w].z, va[w].w, vb[w].x, vb[w].y, vb[w].z, vb[w].w};
        uint4 hi, lo;
        split8(v, hi, lo);
        const uint32_t off = swz(kl, dc);
        *reinterpret_cast<uint4*>(gbase + v_off + off) = hi;
        *reinterpret_cast<uint4*>(gbase + v_off + KC * 128 + off) = lo;
      }
      load_v(i + 1);
#pragma unroll
      for (int j = 0; j < PW / 8; ++j) {
        uint4 hi, lo;
        split8(s + 8 * j, hi, lo);
        const uint32_t off = swz(row, (PW / 8) * part + j);
        *reinterpret_cast<uint4*>(gbase + p_off + off) = hi;
        *reinterpret_cast<uint4*>(gbase + p_off + BM * 128 + off) = lo;
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(buf ? bar_pr1 : bar_pr0);
    }
    // drain: the last commit on each buffer
    {
      const int c0 = (nchunks + 1) / 2, c1 = nchunks / 2;
      if (c0 > 0) mbar_wait(bar_pv0, (c0 - 1) & 1);
      if (c1 > 0) mbar_wait(bar_pv1, (c1 - 1) & 1);
    }
    tc_fence_after();
    xsum[part * BM + row] = rsum;
    softmax_sync();
    float tot = 0.f;
#pragma unroll
    for (int q = 0; q < NPART; ++q) tot += xsum[q * BM + row];
    const float inv = tot > 0.f ? 1.0f / tot : 0.f;    // fully masked row -> 0 (as the fp32 kernel)

    // ---------------------------------------------------------------- epilogue: O / rowsum
    constexpr int OW = D / NPART;                      // output columns per warp (16)
    const int grow = q0 + row;
    if (OW * part < p.dv) {                            // warp-uniform
      float o[OW];
      tmem_ld16(t_row + O_COL + OW * part, o);
      if (grow < p.Lq) {
#pragma unroll
        for (int u = 0; u < OW; ++u) o[u] *= inv;
        if (p.O) {
          float4* dst = reinterpret_cast<float4*>(p.O + (long long)b * p.so + (long long)grow * p.ldo + h * p.dv + OW * part);
#pragma unroll
          for (int u = 0; u < OW / 4; ++u) dst[u] = make_float4(o[4 * u], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]);
        }
        if (p.split_out) {
          __half* sp = p.split_out + ((long long)b * p.Lq + grow) * (2 * p.split_kp) + h * p.dv + OW * part;
#pragma unroll
          for (int j = 0; j < OW / 8; ++j) {
            uint4 hi, lo;
            split8(o + 8 * j, hi, lo);
            *reinterpret_cast<uint4*>(sp + 8 * j) = hi;
            *reinterpret_cast<uint4*>(sp + p.split_kp + 8 * j) = lo;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// =====================================================================================================
// TMA-fed variant: Q, K, V arrive already split into fp16 (hi | lo) halves -- the QKV GEMM epilogue writes
// them -- so the kernel stages nothing by hand: one thread issues TMA loads (128B-swizzled 64 x 64 boxes)
// and the UMMAs; all 16 warps do only the softmax.  V is consumed as an MN-major B operand straight from
// its [keys][head-dim] tile (MN-major, SWIZZLE_128B: rows (= k index) of 128 B holding 64 contiguous MN
// elements, 8-row atoms of 1024 B), so there is no transposition either.
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
struct ParamsT {
  float* O;
  int B, H, Lq, Lk;
  int NB, LB;                       // key blocks (1 or 2) of LB keys each (LB a multiple of 64, <= 448 / 384)
  int ldo;
  long long so;
  float scale;
  __half* split_out;
  int split_kp;
  int q_col, k_col, v_col;          // column of head 0 in the hi half of each buffer
  int q_kp, k_kp, v_kp;             // columns per half
  int q_rows, k_rows;               // rows per batch element in the Q and K/V buffers
  long long* trace;                 // optional [trace_n][10] clock64 stamps of CTA phases (ec_attention_tc_set_trace)
  int trace_n;
  int wide;                         // P-in-TMEM kernel: P_hi [V_hi | V_lo] as one N = 128 MMA (one key block <= 384 keys)
  int dv;                           // head dim, 32 or 64 (P-in-TMEM kernel; boxes are 64 columns wide either way)
  const uint8_t* key_mask;          // [B, Lk], 1 = ignore key, or NULL   (P-in-TMEM kernel, one key block)
  const float* bias;                // [B, H, Lq, Lk] additive, or NULL   (P-in-TMEM kernel, one key block)
  // structural bias computed in the kernel (BiasedMultiheadAttention, utils/bias_attn.py:188-191): bias[b,h,i,j] =
  // W1[h,:] . relu(W0 . hops[:,b,i,j] + b0) + b1[h] -- instead of a [B,H,Lq,Lk] tensor written by a kernel in front
  const float* hops;                // [n_hops, B, Lq, Lk] Markov hop matrices, or NULL
  const float *hw0, *hb0, *hw1, *hb1;   // Linear(n_hops, hidden), Linear(hidden, H)
  int n_hops, hidden;
  int split_fmt;                    // row format of split_out: EC_SPLIT_F16X2, or EC_SPLIT_F16F8 (head dim 64: the consumer is
                                    // ec_gemm_f16f8, e.g. the ViT proj GEMM) -- ec_attention_split_fmt_next
  unsigned long long* overflow;     // F16F8 output: {beyond e4m3, beyond fp16} event counters
};

constexpr int BOX_BYTES = 64 * 128;  // one 64-row x 64-column fp16 box

// Warp-specialised: warps 0..15 do the softmax (4 warps per TMEM lane quarter, 16 keys of every 64-key chunk
// each) and never issue TMA / MMA; lane 0 of warp 16 owns every TMA load, every tcgen05.mma and its commits,
// so descriptor building and instruction issue are off the softmax warps' critical path.  Hand-offs:
//   bar_qk  TMA -> control      Q tile + keys of the block landed
//   bar_s   MMA -> everyone     S = Q K^T of the block is complete (Q / K shared memory is dead)
//   bar_v   TMA -> control      V chunks of the block landed
//   bar_pr* softmax -> control  P chunk written to its buffer (16 warp arrivals)
//   bar_pv* MMA -> softmax      the MMAs reading that P buffer are done (buffer free / O complete)
//   bar_sf  softmax -> control  S of the block consumed and its PV drained (two-block case: next block may start)
//
// Up to 448 keys: one key block, S [128 x Lk] in TMEM columns [0, 448), O in [448, 512).
// Up to 768 keys (ViT-L at 384^2: 730): two key blocks of <= 384 keys processed one after the other, each with
// its own row max / row sum and its own O accumulator (TMEM columns [384, 448) and [448, 512)); the two partial
// results are merged in registers at the end (no rescaling of an accumulator in TMEM).
__global__ void __launch_bounds__(THREADS_T, 1)
attention_tc_tma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                        const __grid_constant__ CUtensorMap tmV, ParamsT p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int NKB = p.LB / 64;                            // 64-key boxes / chunks per block
  const int data_bytes = max(Q_BYTES + 2 * NKB * BOX_BYTES, 2 * PBUF_BYTES + NKB * VBUF_BYTES);
  const uint32_t misc = base + data_bytes;
  const uint32_t bar_qk = misc, bar_s = misc + 8, bar_pv0 = misc + 16, bar_pv1 = misc + 24, bar_v = misc + 32,
                 bar_pr0 = misc + 40, bar_pr1 = misc + 48, bar_sf = misc + 56, tmem_slot = misc + 64;
  float* xmax = reinterpret_cast<float*>(gbase + data_bytes + 128);   // [NPART][BM]
  float* xsum = xmax + NPART * BM;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, part = (warp >> 2) & 3;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BM;
  const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  long long* trace = (p.trace && tid == 0 && cta < p.trace_n) ? p.trace + 10 * cta : nullptr;
  auto stamp = [&](int i) {
    if (trace) trace[i] = clock64();
  };
  stamp(0);
  const uint32_t q_hi = base, q_lo = base + BM * 128, k_hi = base + Q_BYTES, k_lo = k_hi + NKB * BOX_BYTES;
  const int qrow = b * p.q_rows + q0, krow = b * p.k_rows;

  auto load_qk = [&](int blk) {   // Q tile (2 boxes) and the keys of block `blk` (NKB boxes), hi and lo halves
    mbar_expect_tx(bar_qk, (uint32_t)((2 + NKB) * 2 * BOX_BYTES));
    for (int j = 0; j < 2; ++j) {
      tma_load_2d(q_hi + j * BOX_BYTES, &tmQ, bar_qk, p.q_col + h * D, qrow + 64 * j);
      tma_load_2d(q_lo + j * BOX_BYTES, &tmQ, bar_qk, p.q_kp + p.q_col + h * D, qrow + 64 * j);
    }
    for (int j = 0; j < NKB; ++j) {
      tma_load_2d(k_hi + j * BOX_BYTES, &tmK, bar_qk, p.k_col + h * D, krow + blk * p.LB + 64 * j);
      tma_load_2d(k_lo + j * BOX_BYTES, &tmK, bar_qk, p.k_kp + p.k_col + h * D, krow + blk * p.LB + 64 * j);
    }
  };

  if (warp == THREADS / 32 && elect_one()) {
    mbar_init(bar_qk, 1); mbar_init(bar_s, 1);
    mbar_init(bar_pv0, 1); mbar_init(bar_pv1, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_pr0, THREADS / 32); mbar_init(bar_pr1, THREADS / 32);
    mbar_init(bar_sf, THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    load_qk(0);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  stamp(1);

  if (warp == THREADS / 32) {
    // ================================================================== control warp: TMA + MMA issue
    if (elect_one()) {
      const uint32_t idesc_o = make_idesc_bmn(D);
      int g = 0;
      for (int blk = 0; blk < p.NB; ++blk) {
        const int key0 = blk * p.LB;
        const int nkeys = min(p.LB, p.Lk - key0);
        const int lkp = (nkeys + 15) / 16 * 16;
        const int nchunks = (nkeys + KC - 1) / KC;
        const uint32_t o_col = 512 - 64 * (p.NB - blk);
        if (blk > 0) {                                   // S of the previous block consumed, its P / V buffers drained
          mbar_wait(bar_sf, (blk - 1) & 1);
          tc_fence_after();
          load_qk(blk);
        }
        mbar_wait(bar_qk, blk & 1);
        tc_fence_after();
        {
          const int n0 = lkp <= 256 ? lkp : ((lkp / 2 + 15) / 16) * 16;
          for (int noff = 0; noff < lkp; noff += n0) {
            const int n = min(n0, lkp - noff);
            const uint32_t idesc = make_idesc(n);
            const uint64_t aq_hi = make_desc(q_hi), aq_lo = make_desc(q_lo);
            const uint64_t bk_hi = make_desc(k_hi + noff * 128), bk_lo = make_desc(k_lo + noff * 128);
#pragma unroll
            for (int k = 0; k < D / 16; ++k) umma(tmem_base + noff, aq_lo + 2 * k, bk_hi + 2 * k, idesc, k ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < D / 16; ++k) umma(tmem_base + noff, aq_hi + 2 * k, bk_lo + 2 * k, idesc, 1u);
#pragma unroll
            for (int k = 0; k < D / 16; ++k) umma(tmem_base + noff, aq_hi + 2 * k, bk_hi + 2 * k, idesc, 1u);
          }
          umma_commit(bar_s);
        }
        // Q / K shared memory is dead once S is complete: every V chunk of the block ([64 keys][64 d], hi and
        // lo) is loaded behind the P double buffer while the softmax warps take the row max
        mbar_wait(bar_s, blk & 1);
        mbar_expect_tx(bar_v, (uint32_t)(nchunks * 2 * BOX_BYTES));
        for (int i = 0; i < nchunks; ++i) {
          const uint32_t v_hi = base + 2 * PBUF_BYTES + i * VBUF_BYTES;
          tma_load_2d(v_hi, &tmV, bar_v, p.v_col + h * D, krow + key0 + i * KC);
          tma_load_2d(v_hi + BOX_BYTES, &tmV, bar_v, p.v_kp + p.v_col + h * D, krow + key0 + i * KC);
        }
        mbar_wait(bar_v, blk & 1);
        for (int i = 0; i < nchunks; ++i, ++g) {
          const int buf = g & 1;
          const uint32_t p_hi = base + buf * PBUF_BYTES, p_lo = p_hi + BM * 128;
          const uint32_t v_hi = base + 2 * PBUF_BYTES + i * VBUF_BYTES, v_lo = v_hi + BOX_BYTES;
          mbar_wait(buf ? bar_pr1 : bar_pr0, (g >> 1) & 1);
          tc_fence_after();
          const int valid = min(KC, nkeys - i * KC);
          const int ksteps = (valid + 15) / 16;
          const uint64_t ap_hi = make_desc(p_hi), ap_lo = make_desc(p_lo);
          const uint64_t bv_hi = make_desc_mn(v_hi), bv_lo = make_desc_mn(v_lo);
          // P: K-major, +32 B per 16-key step; V: MN-major, 16 key rows = two 1024 B atoms per step
          for (int k = 0; k < ksteps; ++k) umma(tmem_base + o_col, ap_lo + 2 * k, bv_hi + 128 * k, idesc_o, (i | k) ? 1u : 0u);
          for (int k = 0; k < ksteps; ++k) umma(tmem_base + o_col, ap_hi + 2 * k, bv_lo + 128 * k, idesc_o, 1u);
          for (int k = 0; k < ksteps; ++k) umma(tmem_base + o_col, ap_hi + 2 * k, bv_hi + 128 * k, idesc_o, 1u);
          umma_commit(buf ? bar_pv1 : bar_pv0);
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================== softmax warps
    const int row = quarter * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;   // exp(scale * (s - max)) = exp2((s - max) * scale * log2 e)
    float bmax0 = -INFINITY, bmax1 = -INFINITY, bsum0 = 0.f, bsum1 = 0.f;
    int g = 0;                                         // running chunk index: selects the P buffer / barrier parity
    for (int blk = 0; blk < p.NB; ++blk) {
      const int key0 = blk * p.LB;
      const int nkeys = min(p.LB, p.Lk - key0);        // valid keys of this block
      const int nchunks = (nkeys + KC - 1) / KC;
      mbar_wait(bar_s, blk & 1);
      tc_fence_after();
      if (blk == 0) stamp(2);
      // -------------------------------------------------------------- row max over the block
      const int nchunk32 = (nkeys + 31) / 32;
      float mymax = -INFINITY;
      for (int j = part; j < nchunk32; j += NPART) {
        float s[32];
        tmem_ld32(t_row + j * 32, s);
#pragma unroll
        for (int u = 0; u < 32; ++u)
          if (j * 32 + u < nkeys) mymax = fmaxf(mymax, s[u]);
      }
      if (blk > 0) softmax_sync();                     // xmax of the previous block has been consumed
      xmax[part * BM + row] = mymax;
      softmax_sync();
      float rmax = xmax[row];
#pragma unroll
      for (int q = 1; q < NPART; ++q) rmax = fmaxf(rmax, xmax[q * BM + row]);
      if (blk == 0) bmax0 = rmax; else bmax1 = rmax;
      if (blk == 0) stamp(3);
      // -------------------------------------------------------------- P = exp(S - max), 64 keys per chunk
      const float nb = -rmax * sl2;
      float rsum = 0.f;
      for (int i = 0; i < nchunks; ++i, ++g) {
        const int buf = g & 1;
        const uint32_t p_off = buf * PBUF_BYTES;
        float s[PW];
        const int kbase = i * KC + PW * part;
        tmem_ld16(t_row + kbase, s);
#pragma unroll
        for (int u = 0; u < PW; ++u) {
          const float e = (kbase + u < nkeys) ? ex2(fmaf(s[u], sl2, nb)) : 0.f;
          s[u] = e;
          rsum += e;
        }
        // the MMAs two chunks back released this P buffer (buffers are drained between key blocks: there the
        // wait is on an already completed phase)
        if (g >= 2) mbar_wait(buf ? bar_pv1 : bar_pv0, ((g >> 1) - 1) & 1);
#pragma unroll
        for (int j = 0; j < PW / 8; ++j) {
          uint4 hi, lo;
          split8(s + 8 * j, hi, lo);
          const uint32_t off = swz(row, (PW / 8) * part + j);
          *reinterpret_cast<uint4*>(gbase + p_off + off) = hi;
          *reinterpret_cast<uint4*>(gbase + p_off + BM * 128 + off) = lo;
        }
        proxy_fence();
        __syncwarp();
        if (lane == 0) mbar_arrive(buf ? bar_pr1 : bar_pr0);
        if (blk == 0 && i == 0) stamp(4);
      }
      if (blk == 0) bsum0 = rsum; else bsum1 = rsum;
      if (blk == 0) stamp(5);
      // drain the P / V consumers of this block (the next block's TMA loads overwrite the same shared memory,
      // and the epilogue reads the accumulators)
      {
        const int c0 = (g + 1) / 2, c1 = g / 2;        // commits issued so far on bar_pv0 / bar_pv1
        if (c0 > 0) mbar_wait(bar_pv0, (c0 - 1) & 1);
        if (c1 > 0) mbar_wait(bar_pv1, (c1 - 1) & 1);
      }
      tc_fence_after();
      if (blk == 0) stamp(6);
      if (blk + 1 < p.NB) {
        tc_fence_before();                             // this warp's tcgen05.ld of S are complete
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sf);
      }
    }

    // ---------------------------------------------------------------- merge the key blocks, normalise, store
    const float m = fmaxf(bmax0, bmax1);
    const float w0 = ex2((bmax0 - m) * sl2), w1 = p.NB > 1 ? ex2((bmax1 - m) * sl2) : 0.f;
    xsum[part * BM + row] = bsum0 * w0 + bsum1 * w1;
    softmax_sync();
    float tot = 0.f;
#pragma unroll
    for (int q = 0; q < NPART; ++q) tot += xsum[q * BM + row];
    const float inv = 1.0f / tot;
    {
      constexpr int OW = D / NPART;
      float o[OW];
      tmem_ld16(t_row + (512 - 64 * p.NB) + OW * part, o);
      if (p.NB > 1) {
        float o1[OW];
        tmem_ld16(t_row + 448 + OW * part, o1);
#pragma unroll
        for (int u = 0; u < OW; ++u) o[u] = o[u] * w0 + o1[u] * w1;
      }
      const int grow = q0 + row;
      if (grow < p.Lq) {
#pragma unroll
        for (int u = 0; u < OW; ++u) o[u] *= inv;
        if (p.O) {
          float4* dst = reinterpret_cast<float4*>(p.O + (long long)b * p.so + (long long)grow * p.ldo + h * D + OW * part);
#pragma unroll
          for (int u = 0; u < OW / 4; ++u) dst[u] = make_float4(o[4 * u], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]);
        }
        if (p.split_out) {
          __half* sp = p.split_out + ((long long)b * p.Lq + grow) * (2 * p.split_kp) + h * D + OW * part;
#pragma unroll
          for (int j = 0; j < OW / 8; ++j) {
            uint4 hi, lo;
            split8(o + 8 * j, hi, lo);
            *reinterpret_cast<uint4*>(sp + 8 * j) = hi;
            *reinterpret_cast<uint4*>(sp + p.split_kp + 8 * j) = lo;
          }
        }
      }
    }
    stamp(7);
  }
  tc_fence_before();
  __syncthreads();
  stamp(8);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
  stamp(9);
}


// One thread's 16 output columns (part `part` of a 64-wide head) -> 16-byte chunks of its row of the [row][256 B] staging
// tile.  F16X2: chunks [0, 8) = hi16, [8, 16) = lo16.  F16F8 (A role, common.cuh): chunks [0, 8) = hi16, [8, 12) = hi8,
// [12, 16) = lo8.  Chunks are XOR-swizzled by the row.
__device__ __forceinline__ void stage_split16(uint8_t* stg, int row, int part, const float* o, int fmt, uint32_t& ovf) {
  uint32_t h[8];
  float dl[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __half2 hh = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    sub_h2(o[2 * i], o[2 * i + 1], h[i], dl[2 * i], dl[2 * i + 1]);
  }
  uint8_t* r = stg + row * 256;
  const int sw = row & 15;
  *reinterpret_cast<uint4*>(r + (((2 * part) ^ sw) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(r + (((2 * part + 1) ^ sw) << 4)) = make_uint4(h[4], h[5], h[6], h[7]);
  if (fmt == EC_SPLIT_F16X2) {
    uint32_t l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __half2 ll = __floats2half2_rn(dl[2 * i], dl[2 * i + 1]);
      l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(r + (((8 + 2 * part) ^ sw) << 4)) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(r + (((9 + 2 * part) ^ sw) << 4)) = make_uint4(l[4], l[5], l[6], l[7]);
  } else {
    uint32_t h8[4], l8[4];
    float m = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      h8[i] = e4m3x2_h2(h[2 * i]) | (e4m3x2_h2(h[2 * i + 1]) << 16);
      l8[i] = e4m3x2(dl[4 * i] * 2048.f, dl[4 * i + 1] * 2048.f) | (e4m3x2(dl[4 * i + 2] * 2048.f, dl[4 * i + 3] * 2048.f) << 16);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(o[4 * i]), fabsf(o[4 * i + 1])), fmaxf(fabsf(o[4 * i + 2]), fabsf(o[4 * i + 3]))));
    }
    ovf |= (m > 448.f ? 1u : 0u) | (m > 65504.f ? 2u : 0u);
    *reinterpret_cast<uint4*>(r + (((8 + part) ^ sw) << 4)) = make_uint4(h8[0], h8[1], h8[2], h8[3]);
    *reinterpret_cast<uint4*>(r + (((12 + part) ^ sw) << 4)) = make_uint4(l8[0], l8[1], l8[2], l8[3]);
  }
}
// chunk c of the staged row -> its place in the split row of `kp` columns (head h of width 64): a pointer in halves
__device__ __forceinline__ __half* split_chunk_dst(__half* row, int kp, int h, int c, int fmt) {
  if (fmt == EC_SPLIT_F16X2) return row + h * 64 + (c < 8 ? c * 8 : kp + (c - 8) * 8);
  if (c < 8) return row + h * 64 + c * 8;                                    // hi16: halves [0, kp)
  // e4m3 planes: the head's 64 columns are one block of 128 bytes at byte 2 kp + 128 h: hi8 chunks [0, 4), lo8 chunks [4, 8)
  return row + kp + h * 64 + (c - 8) * 8;
}
// ---------------------------------------------------------------------------------------------------------
// P-in-TMEM variant (the default).  The shared-memory P path above is bound by shared-memory bandwidth in the
// P V phase: per 64-key chunk the MMAs re-read P three times (48 KB) and V three times (24 KB) and the softmax
// warps write P once (32 KB) -- 104 KB at 128 B/clk is ~810 clk, more than the exponentials cost.  Here each
// softmax warp writes its 16 probabilities back into the 16 TMEM columns it read them from, as packed fp16:
// columns [c, c+8) = hi(p[2j], p[2j+1]), columns [c+8, c+16) = lo -- exactly one K = 16 A-operand slice each --
// and the MMAs take A from TMEM (tcgen05.mma [d], [a], b_desc).  No P buffers, no buffer hand-back, the only
// shared-memory traffic of the phase is V (24 KB per chunk), and every chunk has its own "P ready" barrier so
// the softmax warps run ahead freely.
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
        "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
        "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]),
        "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]),
        "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
constexpr int MAX_CHUNKS = 8;

constexpr int THREADS_TS = THREADS + 64;   // 16 softmax warps, control warp, second S-issuing warp

// HOP = true: the instance that evaluates the hop-bias MLP per logit (p.hops != NULL); kept apart so that its register
// budget does not touch the plain instance.
template <bool HOP>
__global__ void __launch_bounds__(THREADS_TS, 1)
attention_tc_ts_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, ParamsT p) {
  if (threadIdx.x == 64) {   // the TMA unit fetches a descriptor on its first use: start those fetches now
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int NKB = p.LB / 64;                            // 64-key boxes / chunks per block
  const int data_bytes = Q_BYTES + 2 * NKB * BOX_BYTES; // V (NKB * 16 KB) aliases Q / K once S is complete
  const uint32_t misc = base + data_bytes;
  const uint32_t bar_qk = misc, bar_s = misc + 8, bar_v = misc + 16, bar_o = misc + 24, bar_sf = misc + 32,
                 tmem_slot = misc + 40, bar_pr = misc + 48;            // bar_pr: MAX_CHUNKS barriers, one per chunk
  float* xmax = reinterpret_cast<float*>(gbase + data_bytes + 128);   // [NPART][BM]
  float* xsum = xmax + NPART * BM;
  // key-padding mask of this batch row as a bit per key (1 = ignore; keys >= Lk are set too): one shared-memory word
  // per 32 keys, tested with shifts.  (A byte per key read inside the softmax loops cost a dependent generic load
  // and a branch per element: 19K clk for the row max of 424 keys instead of 1.3K.)
  uint32_t* mskw = reinterpret_cast<uint32_t*>(xsum + NPART * BM);    // [16] words = 512 keys

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, part = (warp >> 2) & 3;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BM;
  const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  long long* trace = (p.trace && tid == 0 && cta < p.trace_n) ? p.trace + 10 * cta : nullptr;
  auto stamp = [&](int i) {
    if (trace) trace[i] = clock64();
  };
  stamp(0);
  const uint32_t q_hi = base, q_lo = base + BM * 128, k_hi = base + Q_BYTES, k_lo = k_hi + NKB * BOX_BYTES;
  const int qrow = b * p.q_rows + q0, krow = b * p.k_rows;

  auto load_qk = [&](int blk) {   // Q tile (2 boxes) and the keys of block `blk` (NKB boxes), hi and lo halves
    mbar_expect_tx(bar_qk, (uint32_t)((2 + NKB) * 2 * BOX_BYTES));
    for (int j = 0; j < 2; ++j) {
      tma_load_2d(q_hi + j * BOX_BYTES, &tmQ, bar_qk, p.q_col + h * p.dv, qrow + 64 * j);
      tma_load_2d(q_lo + j * BOX_BYTES, &tmQ, bar_qk, p.q_kp + p.q_col + h * p.dv, qrow + 64 * j);
    }
    for (int j = 0; j < NKB; ++j) {
      tma_load_2d(k_hi + j * BOX_BYTES, &tmK, bar_qk, p.k_col + h * p.dv, krow + blk * p.LB + 64 * j);
      tma_load_2d(k_lo + j * BOX_BYTES, &tmK, bar_qk, p.k_kp + p.k_col + h * p.dv, krow + blk * p.LB + 64 * j);
    }
  };

  if (warp == THREADS / 32 && elect_one()) {
    mbar_init(bar_qk, 1); mbar_init(bar_s, 2); mbar_init(bar_v, 1); mbar_init(bar_o, 1);
    mbar_init(bar_sf, THREADS / 32);
    for (int i = 0; i < MAX_CHUNKS; ++i) mbar_init(bar_pr + 8 * i, THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  pdl_launch_dependents();     // TMEM is allocated: the next kernel in the stream may start its prologue
  pdl_wait();                  // Q / K / V written by the previous kernel are visible from here on
  stamp(1);

  // S = Q K^T of one key block.  A tcgen05.mma costs its issuing thread >= ~98 clk whatever its shape, so the
  // two halves of the key range are issued by two warps concurrently (disjoint TMEM columns); each commits
  // once to bar_s (count 2).
  auto issue_s = [&](int half, int lkp) {
    const int n0 = ((lkp / 2 + 15) / 16) * 16;
    const int noff = half ? n0 : 0, n = half ? lkp - n0 : n0;
    if (n > 0) {
      const uint32_t idesc = make_idesc(n);
      const uint64_t aq_hi = make_desc(q_hi), aq_lo = make_desc(q_lo);
      const uint64_t bk_hi = make_desc(k_hi + noff * 128), bk_lo = make_desc(k_lo + noff * 128);
      const int dsteps = p.dv / 16;                   // head dim 32: the upper half of the box is the next head
      for (int k = 0; k < dsteps; ++k) umma(tmem_base + noff, aq_lo + 2 * k, bk_hi + 2 * k, idesc, k ? 1u : 0u);
      for (int k = 0; k < dsteps; ++k) umma(tmem_base + noff, aq_hi + 2 * k, bk_lo + 2 * k, idesc, 1u);
      for (int k = 0; k < dsteps; ++k) umma(tmem_base + noff, aq_hi + 2 * k, bk_hi + 2 * k, idesc, 1u);
    }
    umma_commit(bar_s);
  };

  if (warp == THREADS / 32 + 1) {
    // ================================================================== second S-issuing warp
    if (elect_one()) {
      for (int blk = 0; blk < p.NB; ++blk) {
        const int nkeys = min(p.LB, p.Lk - blk * p.LB);
        mbar_wait(bar_qk, blk & 1);
        tc_fence_after();
        issue_s(1, (nkeys + 15) / 16 * 16);
      }
    }
    __syncwarp();
  } else if (warp == THREADS / 32) {
    // ================================================================== control warp: TMA + MMA issue
    if (elect_one()) {
      const uint32_t idesc_o = make_idesc_bmn(D), idesc_w = make_idesc_bmn(2 * D);
      for (int blk = 0; blk < p.NB; ++blk) {
        const int key0 = blk * p.LB;
        const int nkeys = min(p.LB, p.Lk - key0);
        const int lkp = (nkeys + 15) / 16 * 16;
        const int nchunks = (nkeys + KC - 1) / KC;
        const uint32_t o_col = 512 - 64 * (p.NB - blk);
        if (blk > 0) {                                   // P of the previous block consumed by its MMAs, S consumed
          mbar_wait(bar_sf, (blk - 1) & 1);
          tc_fence_after();
        }
        load_qk(blk);
        mbar_wait(bar_qk, blk & 1);
        tc_fence_after();
        issue_s(0, lkp);                                 // key columns [0, n0); warp 17 issues [n0, lkp)
        // Q / K shared memory is dead once S is complete: the V chunks of the block ([64 keys][64 d], hi and lo)
        // are loaded over it while the softmax warps take the row max
        mbar_wait(bar_s, blk & 1);
        mbar_expect_tx(bar_v, (uint32_t)(nchunks * 2 * BOX_BYTES));
        for (int i = 0; i < nchunks; ++i) {
          const uint32_t v_hi = base + i * VBUF_BYTES;
          tma_load_2d(v_hi, &tmV, bar_v, p.v_col + h * p.dv, krow + key0 + i * KC);
          tma_load_2d(v_hi + BOX_BYTES, &tmV, bar_v, p.v_kp + p.v_col + h * p.dv, krow + key0 + i * KC);
        }
        mbar_wait(bar_v, blk & 1);
        for (int i = 0; i < nchunks; ++i) {
          const uint32_t v_hi = base + i * VBUF_BYTES, v_lo = v_hi + BOX_BYTES;
          mbar_wait(bar_pr + 8 * i, blk & 1);
          tc_fence_after();
          const int valid = min(KC, nkeys - i * KC);
          const int ksteps = (valid + 15) / 16;
          const uint64_t bv_hi = make_desc_mn(v_hi), bv_lo = make_desc_mn(v_lo);
          const uint32_t a0 = tmem_base + i * KC;        // P of keys [16k, 16k+16): hi at +16k, lo at +16k+8
          // V: MN-major, 16 key rows = two 1024 B atoms per step
          if (p.wide) {
            // every tcgen05.mma costs >= ~99 clk whatever its N (scripts/probes/umma_probe.cu), so the MMA count is
            // what matters: V_lo sits 8 KB after V_hi, exactly the MN-atom stride of the descriptor, and one N = 128
            // MMA yields P_hi V_hi (columns 384..447) and P_hi V_lo (448..511); the epilogue adds the halves.
            for (int k = 0; k < ksteps; ++k) umma_ts(tmem_base + 384, a0 + 16 * k, bv_hi + 128 * k, idesc_w, (i | k) ? 1u : 0u);
            for (int k = 0; k < ksteps; ++k) umma_ts(tmem_base + 384, a0 + 16 * k + 8, bv_hi + 128 * k, idesc_o, 1u);
          } else {
            for (int k = 0; k < ksteps; ++k) umma_ts(tmem_base + o_col, a0 + 16 * k + 8, bv_hi + 128 * k, idesc_o, (i | k) ? 1u : 0u);
            for (int k = 0; k < ksteps; ++k) umma_ts(tmem_base + o_col, a0 + 16 * k, bv_lo + 128 * k, idesc_o, 1u);
            for (int k = 0; k < ksteps; ++k) umma_ts(tmem_base + o_col, a0 + 16 * k, bv_hi + 128 * k, idesc_o, 1u);
          }
        }
        umma_commit(bar_o);
      }
    }
    __syncwarp();
  } else {
    // ================================================================== softmax warps
    const int row = quarter * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;   // exp(scale * (s - max)) = exp2((s - max) * scale * log2 e)
    float bmax0 = -INFINITY, bmax1 = -INFINITY, bsum0 = 0.f, bsum1 = 0.f;
    // key mask / additive bias (the decoder and encoder attentions of the head; always one key block): the
    // logits become x = s * scale + bias over the unmasked keys
    const bool gen = p.key_mask != nullptr || p.bias != nullptr || HOP;
    const float* bias_row = (p.bias && q0 + row < p.Lq)
                                ? p.bias + (((long long)b * p.H + h) * p.Lq + (q0 + row)) * p.Lk : nullptr;
    // hop-MLP weights of this head behind the mask words: W0 [hidden][n_hops], b0 [hidden], W1[h] [hidden], b1[h]
    float* hopw = reinterpret_cast<float*>(mskw + 16);
    if (HOP) {
      const int nw0 = p.hidden * p.n_hops;
      for (int i = tid; i < nw0 + 2 * p.hidden + 1; i += THREADS)
        hopw[i] = i < nw0 ? p.hw0[i]
                          : (i < nw0 + p.hidden ? p.hb0[i - nw0]
                                                : (i < nw0 + 2 * p.hidden ? p.hw1[h * p.hidden + (i - nw0 - p.hidden)] : p.hb1[h]));
    }                                                   // (published by the barrier of the mask-word block below)
    if (gen) {
      const int key = tid;                              // 16 warps x 32 lanes cover 512 keys >= LB
      const bool ignore = key >= p.Lk || (p.key_mask && p.key_mask[(long long)b * p.Lk + key] != 0);
      const uint32_t word = __ballot_sync(0xffffffffu, ignore);
      if (lane == 0) mskw[warp] = word;
      softmax_sync();
    }
    // x = s * scale + hop-MLP bias for 32 keys of this lane's query row, -inf where masked; written back over S so
    // that the exponential pass (other warps of the lane quarter) reads finished logits
    auto hop_logits = [&](float* s32, int j) {
      const uint32_t mw = mskw[j];
      const int qi = q0 + row;
      const long long plane = (long long)p.B * p.Lq * p.Lk;
      const float* hp = p.hops + ((long long)b * p.Lq + (qi < p.Lq ? qi : 0)) * p.Lk + j * 32;
      const int nh = p.n_hops, hd = p.hidden;
      const float* w0 = hopw;
      const float* b0 = hopw + hd * nh;
      const float* w1 = b0 + hd;
      const float b1 = w1[hd];
#pragma unroll
      for (int u = 0; u < 32; ++u) {                     // (fully unrolled: s32 stays in registers)
        float x = -INFINITY;
        if (!((mw >> u) & 1u)) {                         // warp-uniform (the mask word is shared by the lanes)
          float hv[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) hv[t] = t < nh ? __ldg(hp + t * plane + u) : 0.f;
          float acc = b1;
          for (int c = 0; c < hd; ++c) {
            float a = b0[c];
#pragma unroll
            for (int t = 0; t < 8; ++t)
              if (t < nh) a = fmaf(w0[c * nh + t], hv[t], a);
            acc = fmaf(w1[c], fmaxf(a, 0.f), acc);
          }
          x = fmaf(s32[u], p.scale, acc);
        }
        s32[u] = x;
      }
    };
    for (int blk = 0; blk < p.NB; ++blk) {
      const int key0 = blk * p.LB;
      const int nkeys = min(p.LB, p.Lk - key0);        // valid keys of this block
      const int nchunks = (nkeys + KC - 1) / KC;
      mbar_wait(bar_s, blk & 1);
      tc_fence_after();
      if (blk == 0) stamp(2);
      // -------------------------------------------------------------- row max over the block
      const int nchunk32 = (nkeys + 31) / 32;
      float mymax = -INFINITY;
      for (int j = part; j < nchunk32; j += NPART) {
        float s[32];
        tmem_ld32(t_row + j * 32, s);
        if (HOP) {
          hop_logits(s, j);
#pragma unroll
          for (int u = 0; u < 32; ++u) mymax = fmaxf(mymax, s[u]);
          tmem_st32(t_row + j * 32, reinterpret_cast<const uint32_t*>(s));
        } else if (gen && !bias_row && mskw[j] == 0u) {
          // the key mask is one word per 32 keys, the same for every lane: a word with no masked key (all but the
          // tail of a padded keypoint list) takes the plain path -- the per-element tests doubled the cost of the
          // encoder self-attention (424 keys, 24 of them padding)
#pragma unroll
          for (int u = 0; u < 32; ++u) mymax = fmaxf(mymax, s[u] * p.scale);
        } else if (gen) {
          const uint32_t mw = mskw[j];
          if (bias_row) {
#pragma unroll
            for (int g8 = 0; g8 < 32; g8 += 8) {          // 8 bias loads in flight at a time (register budget)
#pragma unroll
              for (int u = g8; u < g8 + 8; ++u) {
                const float x = fmaf(s[u], p.scale, (mw >> u) & 1u ? 0.f : __ldg(bias_row + j * 32 + u));
                mymax = (mw >> u) & 1u ? mymax : fmaxf(mymax, x);
              }
              asm volatile("" ::: "memory");
            }
          } else {
#pragma unroll
            for (int u = 0; u < 32; ++u) mymax = (mw >> u) & 1u ? mymax : fmaxf(mymax, s[u] * p.scale);
          }
        } else if (j * 32 + 32 <= nkeys) {
#pragma unroll
          for (int u = 0; u < 32; ++u) mymax = fmaxf(mymax, s[u]);
        } else {
#pragma unroll
          for (int u = 0; u < 32; ++u)
            if (j * 32 + u < nkeys) mymax = fmaxf(mymax, s[u]);
        }
      }
      if (blk > 0) softmax_sync();                     // xmax of the previous block has been consumed
      xmax[part * BM + row] = mymax;
      if (HOP) tc_fence_before();                   // the logits written back over S are read by other warps below
      softmax_sync();                                  // also: every warp is done reading S for the max pass
      if (HOP) tc_fence_after();
      float rmax = xmax[row];
#pragma unroll
      for (int q = 1; q < NPART; ++q) rmax = fmaxf(rmax, xmax[q * BM + row]);
      if (rmax == -INFINITY) rmax = 0.f;               // fully masked row: every p below is 0
      if (blk == 0) bmax0 = rmax; else bmax1 = rmax;
      if (blk == 0) stamp(3);
      // -------------------------------------------------------------- P = exp(S - max), in place in TMEM
      const float nb = -rmax * sl2;
      float rsum = 0.f;
      for (int i = 0; i < nchunks; ++i) {
        float s[PW];
        const int kbase = i * KC + PW * part;
        tmem_ld16(t_row + kbase, s);
        if (HOP) {
          // finished logits (scale and bias applied, masked keys -inf -> probability 0)
          const float nbg = -rmax * 1.4426950408889634f;
#pragma unroll
          for (int u = 0; u < PW; ++u) {                 // (columns past the last 32-key word were never rewritten)
            s[u] = (kbase + u < nkeys) ? ex2(fmaf(s[u], 1.4426950408889634f, nbg)) : 0.f;
            rsum += s[u];
          }
        } else if (gen && !bias_row && ((mskw[kbase >> 5] >> (kbase & 31)) & 0xffffu) == 0u) {
          // no masked key among these 16 (warp-uniform): x = s * scale, p = exp2(x * log2 e - max * log2 e) in one FMA
          const float nbg = -rmax * 1.4426950408889634f;
#pragma unroll
          for (int u = 0; u < PW; ++u) {
            s[u] = ex2(fmaf(s[u], sl2, nbg));
            rsum += s[u];
          }
        } else if (gen) {
          const float L2E = 1.4426950408889634f;
          const uint32_t mw = mskw[kbase >> 5] >> (kbase & 31);        // PW = 16 keys: half a word
          const float nbg = -rmax * L2E;
#pragma unroll
          for (int u = 0; u < PW; ++u) {
            const bool ok = !((mw >> u) & 1u);
            const float x = (ok && bias_row) ? fmaf(s[u], p.scale, __ldg(bias_row + kbase + u)) : s[u] * p.scale;
            s[u] = ok ? ex2(fmaf(x, L2E, nbg)) : 0.f;
            rsum += s[u];
          }
        } else if (kbase + PW <= nkeys) {
#pragma unroll
          for (int u = 0; u < PW; ++u) {
            s[u] = ex2(fmaf(s[u], sl2, nb));
            rsum += s[u];
          }
        } else {
#pragma unroll
          for (int u = 0; u < PW; ++u) {
            s[u] = (kbase + u < nkeys) ? ex2(fmaf(s[u], sl2, nb)) : 0.f;
            rsum += s[u];
          }
        }
        uint32_t pk[PW];
#pragma unroll
        for (int j = 0; j < PW / 2; ++j) split_pair(s[2 * j], s[2 * j + 1], pk[j], pk[PW / 2 + j]);
        tmem_st16(t_row + kbase, pk);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pr + 8 * i);
        if (blk == 0 && i == 0) stamp(4);
      }
      if (blk == 0) bsum0 = rsum; else bsum1 = rsum;
      if (blk == 0) stamp(5);
      mbar_wait(bar_o, blk & 1);                       // every P V MMA of the block is complete
      tc_fence_after();
      if (blk == 0) stamp(6);
      if (blk + 1 < p.NB) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sf);
      }
    }

    // ---------------------------------------------------------------- merge the key blocks, normalise, store
    const float m = fmaxf(bmax0, bmax1);
    const float w0 = p.NB > 1 ? ex2((bmax0 - m) * sl2) : 1.f, w1 = p.NB > 1 ? ex2((bmax1 - m) * sl2) : 0.f;
    xsum[part * BM + row] = bsum0 * w0 + bsum1 * w1;
    softmax_sync();
    float tot = 0.f;
#pragma unroll
    for (int q = 0; q < NPART; ++q) tot += xsum[q * BM + row];
    const float inv = tot > 0.f ? 1.0f / tot : 0.f;    // fully masked row -> 0 (as the fp32 kernels)
    constexpr int OW = D / NPART;
    uint32_t ovf = 0;
    if (OW * part < p.dv) {                            // warp-uniform: head dim 32 keeps parts 0 and 1
      float o[OW];
      tmem_ld16(t_row + (p.wide ? 384 : 512 - 64 * p.NB) + OW * part, o);
      if (p.NB > 1 || p.wide) {
        float o1[OW];
        tmem_ld16(t_row + 448 + OW * part, o1);
        const float wa = p.wide ? 1.f : w0, wb = p.wide ? 1.f : w1;
#pragma unroll
        for (int u = 0; u < OW; ++u) o[u] = o[u] * wa + o1[u] * wb;
      }
      const int grow = q0 + row;
#pragma unroll
      for (int u = 0; u < OW; ++u) o[u] *= inv;
      if (grow < p.Lq) {
        if (p.O) {
          float4* dst = reinterpret_cast<float4*>(p.O + (long long)b * p.so + (long long)grow * p.ldo + h * p.dv + OW * part);
#pragma unroll
          for (int u = 0; u < OW / 4; ++u) dst[u] = make_float4(o[4 * u], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]);
        }
      }
      if (p.split_out) {
        // Written straight from the TMEM layout every lane owns a different row (16-byte pieces 3 KB apart: 2048
        // store wavefronts per CTA, ~2K clk of LSU time).  Instead the tile goes through shared memory (dead by
        // now: every MMA has retired) as [row][hi 128 B | lo 128 B], 16-byte chunks XOR-swizzled by the row, and
        // is written out as whole 128-byte lines: 16 lanes per row.
        if (p.split_fmt == EC_SPLIT_F16F8) {             // (head dim 64) [hi16 | hi8 | lo8] chunks
          stage_split16(gbase, row, part, o, EC_SPLIT_F16F8, ovf);
        } else {
#pragma unroll
          for (int j = 0; j < OW / 8; ++j) {
            uint4 hi, lo;
            split8(o + 8 * j, hi, lo);
            const int c = (OW / 8) * part + j;           // 16-byte chunk of the hi half; lo is chunk 8 + c
            *reinterpret_cast<uint4*>(gbase + row * 256 + ((c ^ (row & 15)) << 4)) = hi;
            *reinterpret_cast<uint4*>(gbase + row * 256 + (((8 + c) ^ (row & 15)) << 4)) = lo;
          }
        }
      }
    }
    if (p.split_out) {
      softmax_sync();
      const int w16 = warp;                              // 0..15: rows [8 w16, 8 w16 + 8)
      const int nch = p.dv / 8;                          // 16-byte chunks per half
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int r = 8 * w16 + 2 * it + (lane >> 4), c = lane & 15;
        if (q0 + r < p.Lq && ((c & 7) < nch || p.split_fmt == EC_SPLIT_F16F8)) {
          const uint4 v = *reinterpret_cast<const uint4*>(gbase + r * 256 + ((c ^ (r & 15)) << 4));
          __half* srow = p.split_out + ((long long)b * p.Lq + q0 + r) * (2 * p.split_kp);
          __half* sp = p.split_fmt == EC_SPLIT_F16F8 ? split_chunk_dst(srow, p.split_kp, h, c, EC_SPLIT_F16F8)
                                                     : srow + h * p.dv + (c < 8 ? c * 8 : p.split_kp + (c - 8) * 8);
          *reinterpret_cast<uint4*>(sp) = v;
        }
      }
      report_overflow(p.overflow, ovf);
    }
    stamp(7);
  }
  tc_fence_before();
  __syncthreads();
  stamp(8);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
  stamp(9);
}

// ---------------------------------------------------------------------------------------------------------
// Persistent, software-pipelined form of the P-in-TMEM kernel for the ViT block attention (head dim 64, no mask,
// no bias, one key block of <= 368 keys).  In the kernel above a CTA lives for one (image, head, 128-query) tile
// and its phases are strictly serial: setup + TMEM allocation, Q / K loads + S MMAs (7.0 K of 17.0 K clk in the
// pass-E trace, profiles/r02_e_attention_trace.log), row max, exponentials under the P V MMAs, epilogue.  Here a CTA
// walks a list of tiles and keeps separate shared-memory regions for Q, K and V, so that
//   * the Q / K tiles of tile t+1 are loaded while tile t is in its exponentials (issued as soon as S(t) is complete),
//   * S(t+1) = Q K^T is issued as soon as the P V MMAs of tile t have retired, i.e. it runs under the epilogue of
//     tile t (S lives in TMEM columns [0, 384), O in [384, 512): disjoint),
//   * V(t+1) is loaded under the row max of tile t+1 (its region doubles as the epilogue's staging tile),
//   * barrier set-up and the TMEM allocation happen once per CTA.
// What stays on the softmax warps' critical path per tile is the row max, the exponentials and the epilogue.
// K is loaded with a partial last box (16-row granularity, second tensor map) so that Q + K + V fit 227 KB.
// Hand-offs (phase = tile parity): bar_qk, bar_s, bar_v, bar_pr[i], bar_o as above, plus
//   bar_oe  softmax -> control   the epilogue of the tile has read O out of TMEM and drained its staging tile
//                                (16 warp arrivals): the next tile's P V MMAs / V loads may start.
struct ParamsP {
  ParamsT t;
  int rows_k;        // key rows staged per tile: 64 nfull + r16 (= the S columns used)
  int nfull, r16;    // full 64-key boxes and the 16-row-granular remainder box
  int QT;            // query tiles per (image, head)
  int n_tiles;
};

__global__ void __launch_bounds__(THREADS_TS, 1)
attention_tc_ps_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmKp,
                       const __grid_constant__ CUtensorMap tmVp, ParamsP pp) {
  const ParamsT& p = pp.t;
  if (threadIdx.x == 64) {   // the TMA unit fetches a descriptor on its first use: start those fetches now
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmKp) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmVp) : "memory");
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int nchunks = pp.nfull + (pp.r16 ? 1 : 0);
  const uint32_t q_hi = base, q_lo = base + BM * 128;
  const uint32_t k_hi = base + Q_BYTES, k_lo = k_hi + pp.rows_k * 128;
  const uint32_t v_base = (k_lo + pp.rows_k * 128 + 1023u) & ~1023u;           // nchunks x [hi 8 KB | lo 8 KB]
  uint8_t* stg = gbase + (v_base - base);                                      // epilogue staging tile (32 KB) = start of V
  const uint32_t misc = v_base + max(nchunks * VBUF_BYTES, 2 * BM * 128);     // the staging tile needs 32 KB whatever Lk
  const uint32_t bar_qk = misc, bar_s = misc + 8, bar_v = misc + 16, bar_o = misc + 24, bar_oe = misc + 32,
                 tmem_slot = misc + 40, bar_pr = misc + 48;                    // bar_pr: MAX_CHUNKS barriers
  float* xmax = reinterpret_cast<float*>(gbase + (misc - base) + 128);         // [NPART][BM]
  float* xsum = xmax + NPART * BM;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, part = (warp >> 2) & 3;
  long long* trace = (p.trace && tid == 0 && (int)blockIdx.x < p.trace_n) ? p.trace + 10 * blockIdx.x : nullptr;
  auto stamp = [&](int i) {
    if (trace) trace[i] = clock64();
  };
  stamp(0);
  const int n_it = ((int)blockIdx.x < pp.n_tiles) ? (pp.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto tile_of = [&](int it, int& b, int& h, int& q0) {
    const int w = (int)blockIdx.x + it * (int)gridDim.x;
    const int qt = w % pp.QT, bh = w / pp.QT;
    q0 = qt * BM;
    h = bh % p.H;
    b = bh / p.H;
  };

  if (warp == THREADS / 32 && elect_one()) {
    mbar_init(bar_qk, 1); mbar_init(bar_s, 2); mbar_init(bar_v, 1); mbar_init(bar_o, 2);   // two P V issuing threads
    mbar_init(bar_oe, THREADS / 32);
    for (int i = 0; i < MAX_CHUNKS; ++i) mbar_init(bar_pr + 8 * i, THREADS / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));
  pdl_launch_dependents();
  pdl_wait();
  stamp(1);

  const int lkp = pp.rows_k;                            // S columns: a multiple of 16
  auto issue_s = [&](int half) {                        // two warps issue the two halves of the key range
    const int n0 = ((lkp / 2 + 15) / 16) * 16;
    const int noff = half ? n0 : 0, n = half ? lkp - n0 : n0;
    if (n > 0) {
      const uint32_t idesc = make_idesc(n);
      const uint64_t aq_hi = make_desc(q_hi), aq_lo = make_desc(q_lo);
      const uint64_t bk_hi = make_desc(k_hi + noff * 128), bk_lo = make_desc(k_lo + noff * 128);
      for (int k = 0; k < D / 16; ++k) umma(tmem_base + noff, aq_lo + 2 * k, bk_hi + 2 * k, idesc, k ? 1u : 0u);
      for (int k = 0; k < D / 16; ++k) umma(tmem_base + noff, aq_hi + 2 * k, bk_lo + 2 * k, idesc, 1u);
      for (int k = 0; k < D / 16; ++k) umma(tmem_base + noff, aq_hi + 2 * k, bk_hi + 2 * k, idesc, 1u);
    }
    umma_commit(bar_s);
  };

  // O = P V, issued by TWO threads: the control warp takes the even 64-key chunks and accumulates them in TMEM columns
  // [384, 448), the second issuing warp the odd chunks in [448, 512); the epilogue adds the two partial sums.
  // A tcgen05.mma costs its issuing thread ~100 clk whatever its shape, so one thread issuing the whole P V (48 MMAs in
  // the "wide" form, started late behind the loads) was the tail of every tile: the softmax warps waited ~3 K clk for
  // bar_o after their last chunk (profiles/r02_v_attention_trace_persistent.log).  Each accumulator has ONE issuing
  // thread and a fixed order, so the result is bit-reproducible (both threads feeding one accumulator was 2 % faster
  // and not: the order in which the tensor pipe retires them varies from run to run).
  const uint32_t idesc_o = make_idesc_bmn(D);
  auto issue_pv = [&](int first, int it) {
    const uint32_t o_col = tmem_base + 384 + 64 * first;
    for (int i = first; i < nchunks; i += 2) {
      const uint32_t v_hi = v_base + i * VBUF_BYTES, v_lo = v_hi + BOX_BYTES;
      mbar_wait(bar_pr + 8 * i, it & 1);
      tc_fence_after();
      const int valid = min(KC, p.Lk - i * KC);
      const int ksteps = (valid + 15) / 16;
      const uint64_t bv_hi = make_desc_mn(v_hi), bv_lo = make_desc_mn(v_lo);
      const uint32_t a0 = tmem_base + i * KC;       // P of keys [16k, 16k+16): hi at +16k, lo at +16k+8
      for (int k = 0; k < ksteps; ++k) umma_ts(o_col, a0 + 16 * k + 8, bv_hi + 128 * k, idesc_o, (i != first || k) ? 1u : 0u);
      for (int k = 0; k < ksteps; ++k) umma_ts(o_col, a0 + 16 * k, bv_lo + 128 * k, idesc_o, 1u);
      for (int k = 0; k < ksteps; ++k) umma_ts(o_col, a0 + 16 * k, bv_hi + 128 * k, idesc_o, 1u);
    }
    umma_commit(bar_o);
  };

  if (warp == THREADS / 32 + 1) {
    // ================================================================== second issuing warp: half of S, odd P V chunks
    if (elect_one()) {
      for (int it = 0; it < n_it; ++it) {
        mbar_wait(bar_qk, it & 1);
        if (it > 0) mbar_wait(bar_o, (it - 1) & 1);     // the P V MMAs of the previous tile have consumed P (= the S columns)
        tc_fence_after();
        issue_s(1);
        mbar_wait(bar_oe, it & 1);                      // O read out by the epilogue of the previous tile (phase 0: start-up)
        mbar_wait(bar_v, it & 1);
        tc_fence_after();
        issue_pv(1, it);
      }
    }
    __syncwarp();
  } else if (warp == THREADS / 32) {
    // ================================================================== control warp: TMA + MMA issue
    if (elect_one()) {
      auto load_qk = [&](int it) {
        int b, h, q0;
        tile_of(it, b, h, q0);
        const int qrow = b * p.q_rows + q0, krow = b * p.k_rows;
        mbar_expect_tx(bar_qk, (uint32_t)(Q_BYTES + 2 * pp.rows_k * 128));
        for (int j = 0; j < 2; ++j) {
          tma_load_2d(q_hi + j * BOX_BYTES, &tmQ, bar_qk, p.q_col + h * D, qrow + 64 * j);
          tma_load_2d(q_lo + j * BOX_BYTES, &tmQ, bar_qk, p.q_kp + p.q_col + h * D, qrow + 64 * j);
        }
        for (int j = 0; j < pp.nfull; ++j) {
          tma_load_2d(k_hi + j * BOX_BYTES, &tmK, bar_qk, p.k_col + h * D, krow + 64 * j);
          tma_load_2d(k_lo + j * BOX_BYTES, &tmK, bar_qk, p.k_kp + p.k_col + h * D, krow + 64 * j);
        }
        if (pp.r16) {
          tma_load_2d(k_hi + pp.nfull * BOX_BYTES, &tmKp, bar_qk, p.k_col + h * D, krow + 64 * pp.nfull);
          tma_load_2d(k_lo + pp.nfull * BOX_BYTES, &tmKp, bar_qk, p.k_kp + p.k_col + h * D, krow + 64 * pp.nfull);
        }
      };
      if (n_it > 0) load_qk(0);
      for (int it = 0; it < n_it; ++it) {
        int b, h, q0;
        tile_of(it, b, h, q0);
        const int krow = b * p.k_rows;
        mbar_wait(bar_qk, it & 1);
        if (it > 0) mbar_wait(bar_o, (it - 1) & 1);
        tc_fence_after();
        issue_s(0);
        // V first (the P V MMAs of chunk 0 need it ~2 K clk from now, the next tile's Q / K much later): its region
        // doubles as the epilogue's staging tile, free once the previous tile's epilogue has drained it
        mbar_wait(bar_oe, it & 1);                      // staging tile drained, O read out of TMEM (phase 0: start-up)
        tc_fence_after();
        mbar_expect_tx(bar_v, (uint32_t)(pp.nfull * 2 * BOX_BYTES + 2 * pp.r16 * 128));
        for (int i = 0; i < pp.nfull; ++i) {
          const uint32_t v_hi = v_base + i * VBUF_BYTES;
          tma_load_2d(v_hi, &tmV, bar_v, p.v_col + h * D, krow + i * KC);
          tma_load_2d(v_hi + BOX_BYTES, &tmV, bar_v, p.v_kp + p.v_col + h * D, krow + i * KC);
        }
        if (pp.r16) {
          const uint32_t v_hi = v_base + pp.nfull * VBUF_BYTES;
          tma_load_2d(v_hi, &tmVp, bar_v, p.v_col + h * D, krow + pp.nfull * KC);
          tma_load_2d(v_hi + BOX_BYTES, &tmVp, bar_v, p.v_kp + p.v_col + h * D, krow + pp.nfull * KC);
        }
        mbar_wait(bar_s, it & 1);                       // Q / K shared memory is dead: prefetch the next tile's
        if (it + 1 < n_it) load_qk(it + 1);
        mbar_wait(bar_v, it & 1);
        tc_fence_after();
        issue_pv(0, it);
      }
    }
    __syncwarp();
  } else {
    // ================================================================== softmax warps
    const int row = quarter * 32 + lane;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sl2 = p.scale * 1.4426950408889634f;
    const int nkeys = p.Lk;
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_oe);                 // phase 0 of bar_oe ("nothing to drain"); the epilogue of tile t completes phase t + 1
    const int tr_it = n_it > 2 ? 2 : n_it - 1;        // traced tile: the third one (steady state of the pipeline)
    uint32_t ovf = 0;                                 // F16F8 output: values beyond the e4m3 / fp16 range seen by this thread
    for (int it = 0; it < n_it; ++it) {
      int b, h, q0;
      tile_of(it, b, h, q0);
      if (it == tr_it) stamp(1);                      // (overwrites the set-up stamp: [1] -> [2] = waiting for S of this tile)
      mbar_wait(bar_s, it & 1);
      tc_fence_after();
      if (it == tr_it) stamp(2);
      // -------------------------------------------------------------- row max
      const int nchunk32 = (nkeys + 31) / 32;
      float mymax = -INFINITY;
      for (int j = part; j < nchunk32; j += NPART) {
        float s[32];
        if (j * 32 + 32 <= lkp) {
          tmem_ld32(t_row + j * 32, s);
        } else {                                        // the last 16 columns of S (lkp is a multiple of 16, not of 32)
          tmem_ld16(t_row + j * 32, s);
#pragma unroll
          for (int u = 16; u < 32; ++u) s[u] = -INFINITY;
        }
        if (j * 32 + 32 <= nkeys) {
#pragma unroll
          for (int u = 0; u < 32; ++u) mymax = fmaxf(mymax, s[u]);
        } else {
#pragma unroll
          for (int u = 0; u < 32; ++u)
            if (j * 32 + u < nkeys) mymax = fmaxf(mymax, s[u]);
        }
      }
      xmax[part * BM + row] = mymax;
      softmax_sync();
      float rmax = xmax[row];
#pragma unroll
      for (int q = 1; q < NPART; ++q) rmax = fmaxf(rmax, xmax[q * BM + row]);
      if (it == tr_it) stamp(3);
      // -------------------------------------------------------------- P = exp(S - max), in place in TMEM
      const float nb = -rmax * sl2;
      float rsum = 0.f;
      for (int i = 0; i < nchunks; ++i) {
        const int kbase = i * KC + PW * part;
        if (kbase < lkp) {                               // warp-uniform: the partial chunk has fewer 16-key slices
          float s[PW];
          tmem_ld16(t_row + kbase, s);
          if (kbase + PW <= nkeys) {
#pragma unroll
            for (int u = 0; u < PW; ++u) {
              s[u] = ex2(fmaf(s[u], sl2, nb));
              rsum += s[u];
            }
          } else {
#pragma unroll
            for (int u = 0; u < PW; ++u) {
              s[u] = (kbase + u < nkeys) ? ex2(fmaf(s[u], sl2, nb)) : 0.f;
              rsum += s[u];
            }
          }
          uint32_t pk[PW];
#pragma unroll
          for (int j = 0; j < PW / 2; ++j) split_pair(s[2 * j], s[2 * j + 1], pk[j], pk[PW / 2 + j]);
          tmem_st16(t_row + kbase, pk);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_pr + 8 * i);
        if (it == tr_it && i == 0) stamp(4);
      }
      if (it == tr_it) stamp(5);
      xsum[part * BM + row] = rsum;
      softmax_sync();                                   // (also: every warp has read xmax of this tile)
      float tot = 0.f;
#pragma unroll
      for (int q = 0; q < NPART; ++q) tot += xsum[q * BM + row];
      const float inv = tot > 0.f ? 1.0f / tot : 0.f;
      mbar_wait(bar_o, it & 1);                         // every P V MMA of the tile is complete
      tc_fence_after();
      if (it == tr_it) stamp(6);
      // -------------------------------------------------------------- normalise, store
      constexpr int OW = D / NPART;
      float o[OW], o1[OW];
      tmem_ld16(t_row + 384 + OW * part, o);
      tmem_ld16(t_row + 448 + OW * part, o1);
      if (nchunks < 2) {                                // a single chunk: the odd accumulator was never written
#pragma unroll
        for (int u = 0; u < OW; ++u) o1[u] = 0.f;
      }
#pragma unroll
      for (int u = 0; u < OW; ++u) o[u] = (o[u] + o1[u]) * inv;
      const int grow = q0 + row;
      if (grow < p.Lq && p.O) {
        float4* dst = reinterpret_cast<float4*>(p.O + (long long)b * p.so + (long long)grow * p.ldo + h * D + OW * part);
#pragma unroll
        for (int u = 0; u < OW / 4; ++u) dst[u] = make_float4(o[4 * u], o[4 * u + 1], o[4 * u + 2], o[4 * u + 3]);
      }
      if (p.split_out) {
        // through the staging tile (the V region: every MMA reading V has retired) as [row][hi 128 B | lo 128 B],
        // 16-byte chunks XOR-swizzled by the row, written out as whole 128-byte lines: 16 lanes per row
        stage_split16(stg, row, part, o, p.split_fmt, ovf);
        softmax_sync();
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int r = 8 * warp + 2 * k4 + (lane >> 4), c = lane & 15;
          if (q0 + r < p.Lq) {
            const uint4 v = *reinterpret_cast<const uint4*>(stg + r * 256 + ((c ^ (r & 15)) << 4));
            __half* srow = p.split_out + ((long long)b * p.Lq + q0 + r) * (2 * p.split_kp);
            *reinterpret_cast<uint4*>(split_chunk_dst(srow, p.split_kp, h, c, p.split_fmt)) = v;
          }
        }
      }
      // O has been read out of TMEM and this warp is done with the staging tile
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_oe);
      if (it == tr_it) stamp(7);
    }
    report_overflow(p.overflow, ovf);
  }
  tc_fence_before();
  __syncthreads();
  stamp(8);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
  stamp(9);
}

struct HopArm {   // armed by ec_attention_hop_bias_next, consumed by the next ec_attention_tc_split on this thread
  const float *hops = nullptr, *w0 = nullptr, *b0 = nullptr, *w1 = nullptr, *b1 = nullptr;
  int n_hops = 0, hidden = 0;
};
static thread_local HopArm g_hop;
static thread_local int g_out_fmt = EC_SPLIT_F16X2;   // armed by ec_attention_split_fmt_next, consumed by the next ec_attention_tc_split
static int g_variant = 0;   // 0: persistent pipelined kernel where it applies, else P in TMEM + wide P V MMAs; 3: never the persistent kernel; 2: P in TMEM, three N = 64 MMAs per k-step; 1: P through shared memory
static int g_cta_limit = 0;   // 0 = one persistent CTA per SM; else at most this many (ec_attention_set_cta_limit)
static long long* g_trace = nullptr;
static int g_trace_n = 0;

}  // namespace atc
}  // namespace ec

using namespace ec;

extern "C" int ec_attention_tc(const float* Q, const float* K, const float* V, float* O, int B, int H, int Lq, int Lk,
                               int D, int ldq, int ldk, int ldv, int ldo, long long sq, long long sk, long long sv,
                               long long so, float scale, const uint8_t* key_mask, const float* bias, void* split_out,
                               int split_kp, void* stream) {
  EC_REQUIRE(Q && K && V && (O || split_out), "ec_attention_tc: null pointer");
  EC_REQUIRE(D == 64 || D == 32, "ec_attention_tc: head dim must be 32 or 64");
  EC_REQUIRE(B >= 0 && H > 0 && Lq >= 0 && Lk > 0, "ec_attention_tc: bad shape");
  const int LKP = (Lk + 15) / 16 * 16;
  if (LKP > atc::MAX_LKP) {
    set_error("ec_attention_tc: %d keys exceed the %d S columns that fit in TMEM", Lk, atc::MAX_LKP);
    return EC_ERR_UNSUPPORTED;
  }
  EC_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(V) && (!O || aligned16(O)) && ldq % 4 == 0 && ldk % 4 == 0 &&
                 ldv % 4 == 0 && ldo % 4 == 0 && sq % 4 == 0 && sk % 4 == 0 && sv % 4 == 0 && so % 4 == 0,
             "ec_attention_tc: operands must be 16-byte aligned with strides that are multiples of 4");
  EC_REQUIRE(!split_out || (split_kp == H * D && split_kp % 64 == 0 && (((uintptr_t)split_out) & 15) == 0),
             "ec_attention_tc: split_out needs split_kp == H*D (a multiple of 64) and 16-byte alignment");
  if (B == 0 || Lq == 0) return EC_OK;
  EC_REQUIRE(B <= 65535 && H <= 65535, "ec_attention_tc: grid too large");
  const int data_bytes = atc::Q_BYTES + 2 * LKP * 128 > atc::REUSE_BYTES ? atc::Q_BYTES + 2 * LKP * 128 : atc::REUSE_BYTES;
  const int smem = data_bytes + atc::MISC_BYTES + 1024;
  EC_CUDA((cudaError_t)ensure_dynamic_smem(atc::attention_tc_kernel, atc::Q_BYTES + 2 * atc::MAX_LKP * 128 + atc::MISC_BYTES + 1024));
  atc::Params p{Q, K, V, O, B, H, Lq, Lk, LKP, ldq, ldk, ldv, ldo, sq, sk, sv, so, scale, (__half*)split_out, split_kp,
                D, key_mask, bias};
  dim3 grid(cdiv(Lq, atc::BM), H, B);
  atc::attention_tc_kernel<<<grid, atc::THREADS_T, smem, (cudaStream_t)stream>>>(p);
  return check_launch("ec_attention_tc");
}

/* Q2 / K2 / V2: split-fp16 buffers [rows, 2*kp] (hi | lo); head h of Q lives in columns q_col + 64 h of each
 * half, rows b * q_rows + i; likewise K, V (k_rows rows per batch element).  The three may be one buffer. */
extern "C" int ec_attention_tc_split(const void* Q2, int q_total_rows, int q_kp, int q_col, int q_rows,
                                     const void* K2, int k_total_rows, int k_kp, int k_col, const void* V2,
                                     int v_total_rows, int v_kp, int v_col, int k_rows, float* O, int B, int H, int Lq,
                                     int Lk, int ldo, long long so, float scale, int dv, const uint8_t* key_mask,
                                     const float* bias, void* split_out, int split_kp, void* stream) {
  EC_REQUIRE(Q2 && K2 && V2 && (O || split_out), "ec_attention_tc_split: null pointer");
  EC_REQUIRE(B >= 0 && H > 0 && Lq >= 0 && Lk > 0, "ec_attention_tc_split: bad shape");
  EC_REQUIRE(dv == 32 || dv == 64, "ec_attention_tc_split: head dim must be 32 or 64");
  EC_REQUIRE(q_kp % 64 == 0 && k_kp % 64 == 0 && v_kp % 64 == 0 && q_col % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0,
             "ec_attention_tc_split: halves must be multiples of 64 columns, head offsets multiples of 8");
  // one key block up to 448 keys, two blocks (each a multiple of 64, <= 384) up to 768
  int NB = 1, LB = (Lk + 63) / 64 * 64;
  if (LB > atc::MAX_LKP) {
    NB = 2;
    LB = ((Lk + 1) / 2 + 63) / 64 * 64;
    if (LB > 384) {
      set_error("ec_attention_tc_split: %d keys exceed the two 384-key blocks that fit in TMEM", Lk);
      return EC_ERR_UNSUPPORTED;
    }
  }
  const atc::HopArm hop = atc::g_hop;                  // (consumed: applies to this call only)
  atc::g_hop = atc::HopArm();
  const int out_fmt = atc::g_out_fmt;
  atc::g_out_fmt = EC_SPLIT_F16X2;
  EC_REQUIRE(out_fmt == EC_SPLIT_F16X2 || (split_out && dv == 64), "ec_attention_tc_split: F16F8 output rows need head dim 64 and split_out");
  unsigned long long* ovf_counters = nullptr;
  if (out_fmt == EC_SPLIT_F16F8) {
    ovf_counters = overflow_counters();
    if (!ovf_counters) return EC_ERR_CUDA;
  }
  EC_REQUIRE(!hop.hops || !bias, "ec_attention_tc_split: an additive bias tensor and the fused hop-bias MLP exclude each other");
  EC_REQUIRE(!hop.hops || (hop.n_hops <= 8 && hop.hidden * (hop.n_hops + 2) + 1 <= 112),
             "ec_attention_tc_split: hop-bias MLP too large for the fused form (n_hops <= 8, hidden * (n_hops + 2) < 112)");
  const bool general = dv != 64 || key_mask || bias || hop.hops;   // needs the P-in-TMEM kernel, one key block
  if (general && NB != 1) {
    set_error("ec_attention_tc_split: head dim 32 / key mask / bias need <= 448 keys (got %d)", Lk);
    return EC_ERR_UNSUPPORTED;
  }
  EC_REQUIRE(!split_out || (split_kp == H * dv && split_kp % 64 == 0 && (((uintptr_t)split_out) & 15) == 0),
             "ec_attention_tc_split: split_out needs split_kp == H*dv (a multiple of 64) and 16-byte alignment");
  EC_REQUIRE(!O || (aligned16(O) && ldo % 4 == 0 && so % 4 == 0), "ec_attention_tc_split: O must be 16-byte aligned");
  if (B == 0 || Lq == 0) return EC_OK;
  EC_REQUIRE(B <= 65535 && H <= 65535, "ec_attention_tc_split: grid too large");
  const int NKB = LB / 64;
  const int kq = atc::Q_BYTES + 2 * NKB * atc::BOX_BYTES;
  const int pv = 2 * atc::PBUF_BYTES + NKB * atc::VBUF_BYTES;
  const int data_bytes = kq > pv ? kq : pv;
  const int smem = data_bytes + atc::MISC_BYTES + 1024;
  EC_CUDA((cudaError_t)ensure_dynamic_smem(atc::attention_tc_tma_kernel,
                                           2 * atc::PBUF_BYTES + 7 * atc::VBUF_BYTES + atc::MISC_BYTES + 1024));
  CUtensorMap tmQ, tmK, tmV;
  int rc = tc::get_tensor_map(Q2, q_total_rows, q_kp, 64, &tmQ);
  if (rc) return rc;
  rc = tc::get_tensor_map(K2, k_total_rows, k_kp, 64, &tmK);
  if (rc) return rc;
  rc = tc::get_tensor_map(V2, v_total_rows, v_kp, 64, &tmV);
  if (rc) return rc;
  atc::ParamsT p{O, B, H, Lq, Lk, NB, LB, ldo, so, scale, (__half*)split_out, split_kp, q_col, k_col, v_col,
                 q_kp, k_kp, v_kp, q_rows, k_rows, atc::g_trace, atc::g_trace_n,
                 (atc::g_variant != 2 && NB == 1 && LB <= 384) ? 1 : 0, dv, key_mask, bias,
                 hop.hops, hop.w0, hop.b0, hop.w1, hop.b1, hop.n_hops, hop.hidden, out_fmt, ovf_counters};
  dim3 grid(cdiv(Lq, atc::BM), H, B);
  // the persistent, pipelined kernel takes the plain head-dim-64 attentions whose Q + K + V tiles fit shared memory
  // together (the ViT blocks up to 368 tokens; variant 3 = force off for A/B measurements)
  {
    const int nfull = Lk / 64, r16 = ((Lk % 64) + 15) / 16 * 16, rows_k = nfull * 64 + r16;
    const int nchunks = nfull + (r16 ? 1 : 0);
    const size_t v_bytes = (size_t)nchunks * atc::VBUF_BYTES > 2 * atc::BM * 128 ? (size_t)nchunks * atc::VBUF_BYTES : 2 * atc::BM * 128;
    const size_t ps_smem = (size_t)atc::Q_BYTES + 2 * (size_t)rows_k * 128 + 1024 + v_bytes + atc::MISC_BYTES + 1024;
    if (atc::g_variant == 0 && !general && NB == 1 && rows_k <= 384 && nchunks <= atc::MAX_CHUNKS && ps_smem <= 232448) {
      static int num_sms_dev[64] = {};
      int dev = 0;
      EC_CUDA(cudaGetDevice(&dev));
      if (dev >= 0 && dev < 64 && !num_sms_dev[dev])
        EC_CUDA(cudaDeviceGetAttribute(&num_sms_dev[dev], cudaDevAttrMultiProcessorCount, dev));
      const int num_sms = (dev >= 0 && dev < 64) ? num_sms_dev[dev] : 148;
      EC_CUDA((cudaError_t)ensure_dynamic_smem(atc::attention_tc_ps_kernel, 232448));
      CUtensorMap tmKp = tmK, tmVp = tmV;
      if (r16) {
        rc = tc::get_tensor_map(K2, k_total_rows, k_kp, r16, &tmKp);
        if (rc) return rc;
        rc = tc::get_tensor_map(V2, v_total_rows, v_kp, r16, &tmVp);
        if (rc) return rc;
      }
      atc::ParamsP pp{p, rows_k, nfull, r16, cdiv(Lq, atc::BM), cdiv(Lq, atc::BM) * H * B};
      pp.t.wide = 1;
      int ctas = pp.n_tiles < num_sms ? pp.n_tiles : num_sms;
      if (atc::g_cta_limit > 0 && ctas > atc::g_cta_limit) ctas = atc::g_cta_limit;   // (ec_attention_set_cta_limit)
      launch_pdl(atc::attention_tc_ps_kernel, dim3(ctas), dim3(atc::THREADS_TS), ps_smem, (cudaStream_t)stream, tmQ, tmK,
                 tmV, tmKp, tmVp, pp);
      return check_launch("ec_attention_tc_split");
    }
  }
  if (atc::g_variant != 1 || general) {
    const int ts_smem_max = atc::Q_BYTES + 2 * 7 * atc::BOX_BYTES + atc::MISC_BYTES + 1024;
    if (hop.hops) {
      EC_CUDA((cudaError_t)ensure_dynamic_smem(atc::attention_tc_ts_kernel<true>, ts_smem_max));
      launch_pdl(atc::attention_tc_ts_kernel<true>, grid, dim3(atc::THREADS_TS), (size_t)(kq + atc::MISC_BYTES + 1024),
                 (cudaStream_t)stream, tmQ, tmK, tmV, p);
    } else {
      EC_CUDA((cudaError_t)ensure_dynamic_smem(atc::attention_tc_ts_kernel<false>, ts_smem_max));
      launch_pdl(atc::attention_tc_ts_kernel<false>, grid, dim3(atc::THREADS_TS), (size_t)(kq + atc::MISC_BYTES + 1024),
                 (cudaStream_t)stream, tmQ, tmK, tmV, p);
    }
  } else {
    EC_REQUIRE(out_fmt == EC_SPLIT_F16X2, "ec_attention_tc_split: variant 1 (A/B only) writes F16X2 rows only");
    atc::attention_tc_tma_kernel<<<grid, atc::THREADS_T, smem, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
  }
  return check_launch("ec_attention_tc_split");
}

extern "C" int ec_attention_hop_bias_next(const float* hops, int n_hops, int hidden, const float* w0, const float* b0,
                                          const float* w1, const float* b1) {
  EC_REQUIRE(hops && w0 && b0 && w1 && b1 && n_hops > 0 && hidden > 0, "ec_attention_hop_bias_next: bad arguments");
  atc::g_hop.hops = hops; atc::g_hop.n_hops = n_hops; atc::g_hop.hidden = hidden;
  atc::g_hop.w0 = w0; atc::g_hop.b0 = b0; atc::g_hop.w1 = w1; atc::g_hop.b1 = b1;
  return EC_OK;
}

extern "C" int ec_attention_split_fmt_next(int fmt) {
  EC_REQUIRE(fmt == EC_SPLIT_F16X2 || fmt == EC_SPLIT_F16F8, "ec_attention_split_fmt_next: EC_SPLIT_F16X2 or EC_SPLIT_F16F8");
  atc::g_out_fmt = fmt;
  return EC_OK;
}

extern "C" int ec_attention_set_cta_limit(int ctas) {
  EC_REQUIRE(ctas >= 0, "ec_attention_set_cta_limit: negative");
  atc::g_cta_limit = ctas;
  return EC_OK;
}

extern "C" int ec_attention_tc_set_trace(void* buf, int n_ctas) {
  atc::g_trace = (long long*)buf;
  atc::g_trace_n = buf ? n_ctas : 0;
  return EC_OK;
}

extern "C" int ec_attention_tc_set_variant(int variant) {
  EC_REQUIRE(variant >= 0 && variant <= 3, "ec_attention_tc_set_variant: 0 .. 3");
  atc::g_variant = variant;
  return EC_OK;
}
