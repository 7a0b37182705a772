// Fused GCN feed-forward for sm_100a, second formulation (encoder_decoder.py:508-524 of the reference):
//     Y[b,w,:] = relu( a0[b,w] * (X[b,w,:] W0^T + b0) + sum_v A1[b,w,v] (X[b,v,:] W1^T + b1) )
// evaluated PROJECT-FIRST:   T0 = X W0^T + b0,  T1 = X W1^T + b1,  D2 = A1 T1 (= A1 X W1^T + rowsum(A1) b1),
//     Y = relu( a0 T0 + D2 ).
// The first formulation (gcn_fused_tcgen05.cu: aggregate first) is a chain  fill -> A1 X -> [a0 X | A1 X] W^T  in which
// nothing overlaps: at batch 64 the launch is one wave and its time is that chain (fill 6.1 K clk, GEMM 1 3.0 K, GEMM 2
// 10.3 K, epilogue 4.1 K).  Here the large product (X W^T, 2 d NS MACs per row) depends on X alone, so it runs WHILE X is
// being filled, one 64-channel block behind the workers; X is only ever a K-major A operand, which lets the cross terms
// of the fp32-grade split run on e4m3 tensor cores exactly as in gemm_tcgen05.cu (hi16.hi16 on kind::f16, lo8.hi8 and
// hi8.lo8 on kind::f8f6f4: 2 instead of 3 units of tensor time); the small product A1 T1 takes A1 from TENSOR MEMORY
// (packed fp16 hi / lo written by the workers with tcgen05.st) and T1 as an MN-major B operand the workers drain out of
// TMEM, rescale and split into the shared memory the retired X blocks occupied.  The biases are preloaded into the
// accumulators and a0 enters in the epilogue (two accumulators): no rescale pass, no row sums, no "a0 != 1" special case.
//
// CTA = persistent worker over items (sample b, slice of NS output channels).  Warps 0..15: workers (fill, drain,
// epilogue); warp 16: MMA issue; warp 17: TMA.  Per item:
//   TMA       raw X[b] as ng blocks [K rows x 64 channels] INTO the regions of their own tiles; the biases of the slice (two
//             bulk copies); W in 32-deep slices through a ring of four (one box of 128-byte rows [hi16 | hi8 | lo8] each:
//             the weights are packed with their planes interleaved per 32 columns, ec_split_f16f8 role 2); A1[b] as one
//             bulk copy into X blocks 0 / 1 once GEMM A has consumed them; the NEXT item's X as soon as GEMM B has retired
//   init      T0 = b0 w_scale, T1 = b1 w_scale (tcgen05.st)
//   fill      per block: raw rows -> registers -> CTA barrier -> [hi16 (128B swizzle) | hi8 | lo8 (64B swizzle)] K-major
//             tiles over the same bytes, one mbarrier per block
//   GEMM A    per block kb: T0 += X_kb W0_kb^T, T1 += X_kb W1_kb^T; in the last block T1 goes first, so its drain overlaps
//             the last T0 step
//   A1        raw rows (shared memory) -> packed fp16 hi | lo -> TMEM columns (row w = lane, k = v two per column)
//   drain     T1 / w_scale -> fp16 hi | lo, [v][64 n] MN-major tiles over X blocks 0 .. NS/64-1
//   GEMM B    D2 (the T1 columns) = A1 . T1, three fp16 products, A from TMEM
//   epilogue  relu(a0 T0 / w_scale + D2), 64 columns per round into 128B-swizzled tiles in the W ring (idle by then), each
//             round's tiles leave by TMA tensor stores while the next round is computed
// Shared memory (K = 100, d = 256, NS = 192): X 4 x 28 KB, W ring 4 x 24 KB.  TMEM: T0 [0, NS), T1 / D2 [NS, 2 NS), A1 hi
// [2 NS, +64), A1 lo [2 NS + 64, +64).  The worker branch runs at the 96-register limit of an 18-warp CTA and shared memory
// takes nearly all of the L1: a spill in a hot path costs an L2 round trip (profiles/r03_gcn2_development.md, steps 2 / 5),
// which is why the fill is single buffered and shared memory is addressed by 32-bit shared addresses.
// Measured: 12.3 us at B = 64 (aggregate-first kernel 15.3), 298 us at B = 2048 (455): 0.24 / 0.31 of the HBM roofline.
#include <cuda.h>
#include <cuda_fp16.h>

#include <mutex>

#include "common.cuh"

namespace ec {
namespace tc {
int get_tensor_map(const void* ptr, int rows, int kp, int box_rows, CUtensorMap* out);         // gemm_tcgen05.cu
int get_tensor_map_slice32(const void* ptr, int rows, int kp, int box_rows, CUtensorMap* out); // (128-byte boxes = one 32-deep slice of interleaved planes)
int get_tensor_map_f32(const void* ptr, long long rows, int cols, int box_rows, int box_cols, bool swizzle128, CUtensorMap* out);
}
namespace gf2 {

constexpr int WORKERS = 512;             // 16 worker warps (4 per TMEM lane quarter)
constexpr int THREADS = WORKERS + 64;    // + MMA warp + TMA warp
constexpr int MAX_G = 4;                 // d <= 256: at most four 64-channel blocks

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
// bulk copy of `bytes` (a multiple of 16) from shared to global memory, bulk async-group bookkeeping
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// ---- two CTAs per sample (TWO): distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void sts128_cluster(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {   // acquires the peer CTA's writes
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// K-major operand tile, 128B swizzle: rows of 128 B (64 fp16), 8-row atoms of 1024 B (SBO)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// K-major tile of e4m3 bytes, 64B swizzle: rows of 64 B, 8-row atoms of 512 B (layout type 4; f8_mma_probe.cu)
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// MN-major B operand, 128B swizzle: rows (= k index) of 128 B holding 64 contiguous N elements, 8-row atoms of
// 1024 B (SBO); the next 64 N elements live `lbo` bytes further (one tile)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {           // D = f32, M = 128, N = n, A and B K-major, formats 0
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
        "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8_nowait(uint32_t taddr, const uint32_t* u) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// byte offset of the 16-byte chunk `c` of row `row`: 128-byte rows / 128B swizzle, 64-byte rows / 64B swizzle
__device__ __forceinline__ uint32_t swz128(int row, int c) { return (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4)); }
__device__ __forceinline__ uint32_t swz64(int row, int c) { return (uint32_t)(row * 64 + ((c ^ ((row >> 1) & 3)) << 4)); }

struct Params {
  const float* X;      // [B, K, d]
  const float* adj;    // [B, 2, K, K], plane 0 diagonal
  const float* bias2;  // [2][dff] fp32: b0 then b1 (the bias columns of the packed weights, contiguous)
  float* Y;            // [B, K, dff] or NULL
  __half* split_out;   // [B*K, 2*split_kp] = [hi | lo] of Y, or NULL
  int split_kp;
  int B, K, d, dff, NS, k16, Kp;   // k16 = K rounded up to 16: the depth of GEMM B
  int Kt, kt16;        // rows of one CTA's tile (K, or K / 2 when two CTAs share a sample) and Kt rounded up to 16
  int kb_ret;          // after this k-block of GEMM A the first X blocks are dead: A1 may land there
  float out_scale;     // 1 / (power-of-two scale of the split weights)
  float w_scale;       // that scale
  int a1_tma;          // A1[b] arrives as one bulk copy into the retired X blocks 0 and 1 (else: per-thread loads)
  unsigned long long* overflow;   // {beyond e4m3, beyond fp16} event counters of the X split
  long long* trace;    // optional [trace_n][32] clock64 stamps of the first item of a CTA (ec_gcn_fused_set_trace)
  int trace_n;
  int dbg;             // experiments (ec_gcn_fused_set_debug; results are wrong when != 0): 1 = no global loads of A1 / X, 2 = no stores, 4 = no GEMM A UMMAs, 8 = no W loads
};

// barrier indices (8 bytes each)
constexpr int WRING = 4;                 // W ring: four slices of 32 k (see the TMA producer)
enum { W_FULL = 0, W_EMPTY = 4, B_XRAW = 8, B_XF = 16, B_T1 = 20, B_T1S = 21, B_ACC = 22, B_OFREE = 23, B_XRET = 24,
       B_A1RAW = 25, B_BIAS = 26, B_PEER = 27, NUM_BARS = 32 };

struct Layout {
  uint32_t blk, x_bytes, wst, total;
};
// kt16: rows of the CTA's X tile (rounded up to 16); k16: depth of GEMM B = rows of the T1 tiles (== kt16 unless two CTAs
// share a sample)
__host__ __device__ inline Layout make_layout(int kt16, int k16, int d, int NS) {
  Layout L;
  L.blk = (uint32_t)kt16 * 256u;                                  // one X block: hi16 (kt16 x 128 B) | hi8 | lo8 (kt16 x 64 B each)
  L.x_bytes = (uint32_t)(d / 64) * L.blk;
  const uint32_t t1 = (uint32_t)(NS / 64) * (uint32_t)k16 * 256u; // T1 tiles: NS / 64 groups of hi | lo [k16 x 128 B]
  if (L.x_bytes < t1) L.x_bytes = t1;
  // ... the region later holds A1 (raw) and the T1 tiles (NS <= d: they fit); the output tiles (4 NS k16 bytes) are built in
  // the W ring
  L.wst = (uint32_t)NS * 128u;                                    // one W slice (32 k): NS rows of [hi16 64 B | hi8 32 B | lo8 32 B]
  // + barriers / TMEM slot (512 B) + a0 [2][128] floats + the biases of the slice, b0 [256] and b1 [256] floats
  L.total = L.x_bytes + WRING * L.wst + 512u + 2u * 128u * 4u + 2u * 256u * 4u;
  return L;
}

// shared-memory accesses by 32-bit shared address (no 64-bit generic pointers to keep live: the worker branch runs at the
// 96-register limit of an 18-warp CTA, and a spill is expensive here -- shared memory takes nearly all of the L1)
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts32f(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
__device__ __forceinline__ float4 lds128f(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ float lds32f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
  return v;
}

// eight channels (chunk xc of the block) of row r -> hi16 (128B swizzle) | hi8 | lo8 (64B swizzle) tiles of the block.
// A role of the F16F8 split (common.cuh): hi8 = e4m3(hi16), lo8 = e4m3((x - hi16) 2^11); the fill is bound by the issue
// slots of this conversion (~830 clk per block with every scheduler busy), hence the mixed-precision forms.
__device__ __forceinline__ void conv_x_row(uint32_t blk, uint32_t TS, int r, int xc, float4 v0, float4 v1, uint32_t& ovf) {
  const float f[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  uint32_t h[4], h8[2], l8[2];
  __half2 m2 = __float2half2_rn(0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 hh = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&hh);
    m2 = __hmax2(m2, __habs2(hh));
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float d0, d1, d2, d3;                                           // the low parts x - hi16 (exact; one FHADD each)
    sub_h2(f[4 * i], f[4 * i + 1], h[2 * i], d0, d1);
    sub_h2(f[4 * i + 2], f[4 * i + 3], h[2 * i + 1], d2, d3);
    h8[i] = e4m3x2_h2(h[2 * i]) | (e4m3x2_h2(h[2 * i + 1]) << 16);
    l8[i] = e4m3x2(d0 * 2048.f, d1 * 2048.f) | (e4m3x2(d2 * 2048.f, d3 * 2048.f) << 16);
  }
  const float m = fmaxf(__low2float(m2), __high2float(m2));      // (inf when a value is beyond the fp16 range)
  ovf |= (m > 448.f ? 1u : 0u) | (m > 65504.f ? 2u : 0u);
  sts128(blk + swz128(r, xc), h[0], h[1], h[2], h[3]);
  const uint32_t o8 = blk + TS + swz64(r, xc >> 1) + (uint32_t)(xc & 1) * 8u;
  sts64(o8, h8[0], h8[1]);
  sts64(o8 + TS / 2, l8[0], l8[1]);
}

// TWO = false: one CTA per (sample, slice) item, K <= 128.
// TWO = true : a cluster of two CTAs per item, K in (128, 256] (configs[4]: K = 200): CTA r owns rows [r K/2, (r+1) K/2) of
//              the sample -- its X rows, its rows of A1 (all K columns, 2 x k16/2 TMEM columns), its rows of T0 / T1 / Y.
//              GEMM B needs T1 of ALL rows: each CTA drains its T1 rows into BOTH CTAs' T1 tiles (st.shared::cluster), after
//              the peer has signalled that its X region is dead (B_PEER), and GEMM B waits for both halves (B_T1S counts the
//              sixteen local warps and the peer's sixteen).
template <bool TWO>
__global__ void __launch_bounds__(THREADS, 1)
gcn_fused2_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmS, Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const int K = p.K, d = p.d, NS = p.NS, k16 = p.k16, Kt = p.Kt, kt16 = p.kt16;
  const int ng = d / 64, nsg = NS / 64, ksteps = k16 / 16;
  const Layout L = make_layout(kt16, k16, d, NS);
  const uint32_t TS = (uint32_t)kt16 * 128u;                       // one [kt16 rows x 128 B] tile (X, output)
  const uint32_t TS2 = (uint32_t)k16 * 128u;                       // one [k16 rows x 128 B] T1 tile (== TS unless TWO)
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;              // which half of the sample's rows
  const int row0 = (int)rank * Kt;                                 // first row of this CTA's tile within the sample
  const uint32_t xb = base, w0 = xb + L.x_bytes, misc = w0 + WRING * L.wst;
  auto bar = [&](int i) { return misc + 8u * i; };
  const uint32_t tmem_slot = misc + 8u * NUM_BARS;
  const uint32_t a0s_u = misc + 512u, bias_u = a0s_u + 1024u;        // a0 [2][128] floats; b0 [256], b1 [256] floats
  // TMEM columns
  const uint32_t T0C = 0, T1C = (uint32_t)NS, A1H = 2u * NS, A1L = 2u * NS + (uint32_t)k16 / 2u;
  // the T1 tiles fit the X blocks that are dead before the last T0 step retires (never with a peer writing into them)
  const bool early = !TWO && nsg <= ng - 1;
  const int kb_ret = p.kb_ret;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int nslice = p.dff / NS, items = p.B * nslice;
  const int cta = TWO ? (int)blockIdx.x / 2 : (int)blockIdx.x;     // worker (CTA or CTA pair) index
  const int nworkers = TWO ? (int)gridDim.x / 2 : (int)gridDim.x;
  auto stamp = [&](int i) {
    if (p.trace && (int)blockIdx.x < p.trace_n) p.trace[blockIdx.x * 32 + i] = clock64();
  };

  if (tid == 0) {
    for (int g = 0; g < MAX_G; ++g) mbar_init(bar(B_XF + g), WORKERS / 32);
    mbar_init(bar(B_T1), 1);
    mbar_init(bar(B_T1S), (TWO ? 2 : 1) * (WORKERS / 32));
    mbar_init(bar(B_PEER), WORKERS / 32);
    mbar_init(bar(B_ACC), 1);
    mbar_init(bar(B_OFREE), 1);
    mbar_init(bar(B_XRET), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid == 32) {
    // the TMA unit fetches a descriptor on its first use: start those fetches now
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    if (p.Y) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    if (p.split_out) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmS) : "memory");
  }
  if (warp == WORKERS / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ------------------------------------------------------------------ TMA producer (warp 17): raw X blocks, A1, the W ring
  // X[b] arrives as ng raw fp32 blocks [K rows x 64 channels] (256-byte rows), each INTO the region its converted
  // tiles will occupy (4 bytes per element either way); the workers read a block into registers, synchronise and
  // write the tiles over it.  No thread holds X in registers across a global-memory latency, and all of X is in
  // flight at once.  (Loaded through registers with a two-block ring the values were spilled, and each spill store
  // waited for its load: the blocks arrived one latency apart, profiles/r03_gcn2_development.md.)
  bool leader = false;                     // (TMA warp) the elected lane
  uint32_t gs = 0;                         // (TMA warp) W slices requested so far
  // W streams in slices of 32 k through a ring of four (24 KB each at NS = 192): with two 48 KB stages of 64 k a
  // stage could only be re-requested when its UMMAs had retired, i.e. one request in flight per SM against an L2
  // latency of ~1.4 K clk -- 1.2 K clk per 784 clk of math (profiles/r03_gcn2_development.md).  A slice is ONE box with
  // 128-byte rows [hi16 x 32 | hi8 x 32 | lo8 x 32] (the weights are packed with the planes interleaved per 32 columns:
  // ec_split_f16f8, role 2).
  auto w_step = [&](int s, int n0) {
      const int kb = s >> 2, hh = (s >> 1) & 1, sub = s & 1, half = (kb == ng - 1) ? 1 - hh : hh;
      const int st = (int)(gs & (WRING - 1));
      if (gs >= WRING) mbar_wait(bar(W_EMPTY + st), ((gs / WRING) - 1u) & 1u);
      const uint32_t dst = w0 + (uint32_t)st * L.wst;
      const int k0 = half * d + kb * 64 + sub * 32;
      if (p.dbg & 8) {
        mbar_arrive(bar(W_FULL + st));
      } else {
        mbar_expect_tx(bar(W_FULL + st), L.wst);
        tma_load_2d(dst, &tmW, bar(W_FULL + st), 4 * k0, n0);         // byte column 128 (k0 / 32)
      }
      ++gs;
    };
    auto x_loads = [&](int it, int b, int n0) {
      if (it > 0) mbar_wait(bar(B_ACC), (uint32_t)(it - 1) & 1u);     // GEMM B of the previous item has retired: the region is dead
      // biases of the slice: b0[n0 .. n0 + NS) and b1[n0 .. n0 + NS), two bulk copies
      mbar_expect_tx(bar(B_BIAS), (uint32_t)NS * 8u);
      bulk_load(bias_u, p.bias2 + n0, (uint32_t)NS * 4u, bar(B_BIAS));
      bulk_load(bias_u + 1024u, p.bias2 + p.dff + n0, (uint32_t)NS * 4u, bar(B_BIAS));
      for (int g = 0; g < ng; ++g) {
        if (p.dbg & 1) { mbar_arrive(bar(B_XRAW + g)); continue; }
        mbar_expect_tx(bar(B_XRAW + g), (uint32_t)Kt * 256u);
        tma_load_2d(xb + (uint32_t)g * L.blk, &tmX, bar(B_XRAW + g), g * 64, b * K + row0);
      }
    };
  if (warp == WORKERS / 32 + 1) {
    // before the CTA-wide setup barrier: the TMA thread initialises its own barriers and requests the first item's
    // operands, so their latency overlaps the TMEM allocation and the barrier
    leader = elect_one();
    if (leader) {
      for (int s = 0; s < WRING; ++s) { mbar_init(bar(W_FULL + s), 1); mbar_init(bar(W_EMPTY + s), 1); }
      for (int g = 0; g < MAX_G; ++g) mbar_init(bar(B_XRAW + g), 1);
      mbar_init(bar(B_A1RAW), 1);
      mbar_init(bar(B_BIAS), 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      pdl_wait();
      if (cta < items) {                   // the first item's loads go out before the CTA has finished setting up:
        const int n0 = (cta % nslice) * NS;   // X (from HBM, and converted before it is of use) ahead of W (from L2)
        x_loads(0, cta / nslice, n0);
        for (int s = 0; s < WRING; ++s) w_step(s, n0);
      }
    }
    __syncwarp();
  }
  // CTA-wide setup barrier
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();             // the peer's barriers are initialised before any remote arrive
  tc_fence_after();
  pdl_launch_dependents();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + (tmem_slot - base));

  if (warp == WORKERS / 32 + 1) {
    if (leader) {
      int it = 0;
      const int s_a1 = min(4 * (kb_ret + 1) + WRING, 4 * ng);      // slices that can go out before X blocks 0 / 1 retire
      for (int item = cta; item < items; item += nworkers, ++it) {
        const int n0 = (item % nslice) * NS, b = item / nslice;
        if (it > 0) {
          // the next item's X is requested as soon as the region is dead -- while the workers are still in the previous
          // epilogue; the W ring holds that epilogue's output tile until its stores have read it
          x_loads(it, b, n0);
          mbar_wait(bar(B_OFREE), (uint32_t)(it - 1) & 1u);
          for (int s = 0; s < WRING; ++s) w_step(s, n0);
        }
        for (int s = WRING; s < s_a1; ++s) w_step(s, n0);
        if (p.a1_tma) {
          // A1[b] (K x K fp32, contiguous) -> X blocks 0 / 1 once GEMM A has consumed them; the workers move it on to
          // TMEM.  (Read by the workers straight from global memory -- one row per lane -- it took ~6 K clk.)
          mbar_wait(bar(B_XRET), (uint32_t)it & 1u);
          if (p.dbg & 1) {
            mbar_arrive(bar(B_A1RAW));
          } else {
            mbar_expect_tx(bar(B_A1RAW), (uint32_t)(Kt * K) * 4u);    // this CTA's rows of A1, all K columns
            bulk_load(xb, p.adj + (((long long)b * 2 + 1) * K + row0) * K, (uint32_t)(Kt * K) * 4u, bar(B_A1RAW));
          }
        }
        for (int s = s_a1; s < 4 * ng; ++s) w_step(s, n0);
      }
    }
  } else if (warp == WORKERS / 32) {
    // ------------------------------------------------------------------ MMA issuer
    pdl_wait();
    if (elect_one()) {
      uint32_t gs = 0;
      int it = 0;
      const uint32_t idesc_a = make_idesc(NS), idesc_b = make_idesc(NS) | (1u << 16);
      for (int item = cta; item < items; item += nworkers, ++it) {
        const uint32_t par = (uint32_t)it & 1u;
        for (int kb = 0; kb < ng; ++kb) {
          mbar_wait(bar(B_XF + kb), par);
          tc_fence_after();
          if (it == 0) stamp(12 + kb);
          const uint32_t xblk = xb + (uint32_t)kb * L.blk;
          const uint64_t a_hi = make_desc(xblk), a_h8 = make_desc_sw64(xblk + TS), a_l8 = make_desc_sw64(xblk + TS + TS / 2);
          for (int hh = 0; hh < 2; ++hh) {
            const int half = (kb == ng - 1) ? 1 - hh : hh;
            const uint32_t acc = tmem_base + (half ? T1C : T0C);     // (the accumulators start out holding the biases)
#pragma unroll
            for (int sub = 0; sub < 2; ++sub, ++gs) {
              const int st = (int)(gs & (WRING - 1));
              mbar_wait(bar(W_FULL + st), (gs / WRING) & 1u);
              tc_fence_after();
              const uint32_t wb = w0 + (uint32_t)st * L.wst;
              // slice rows of 128 B (128B swizzle): hi16 in bytes [0, 64), hi8 in [64, 96), lo8 in [96, 128)
              const uint64_t b_hi = make_desc(wb), b_h8 = b_hi + 4, b_l8 = b_hi + 6;
              if (!(p.dbg & 4)) {
                umma_f8(acc, a_l8 + 2 * sub, b_h8, idesc_a, 1u);
                umma_f8(acc, a_h8 + 2 * sub, b_l8, idesc_a, 1u);
                umma_f16(acc, a_hi + 4 * sub, b_hi, idesc_a, 1u);
                umma_f16(acc, a_hi + 4 * sub + 2, b_hi + 2, idesc_a, 1u);
              }
              umma_commit(bar(W_EMPTY + st));
            }
            if (kb == ng - 1 && hh == (early ? 0 : 1)) umma_commit(bar(B_T1));   // T1 complete (and, late, X dead)
          }
          if (kb == kb_ret) umma_commit(bar(B_XRET));                // X blocks 0 .. kb_ret are dead
        }
        if (it == 0) stamp(16);
        if (TWO) mbar_wait_cluster(bar(B_T1S), par); else mbar_wait(bar(B_T1S), par);
        tc_fence_after();
        if (it == 0) stamp(17);
        {
          // GEMM B: D2[128 x NS] = A1 (TMEM, k16 deep) . T1 (MN-major rows v, NS columns in nsg tiles TS apart)
          const uint64_t t_hi = make_desc_mn(xb, TS2), t_lo = make_desc_mn(xb + (uint32_t)nsg * TS2, TS2);
          const uint32_t d2 = tmem_base + T1C, ah = tmem_base + A1H, al = tmem_base + A1L;
          for (int k = 0; k < ksteps; ++k) umma_ts(d2, al + 8 * k, t_hi + 128 * k, idesc_b, k ? 1u : 0u);
          for (int k = 0; k < ksteps; ++k) umma_ts(d2, ah + 8 * k, t_lo + 128 * k, idesc_b, 1u);
          for (int k = 0; k < ksteps; ++k) umma_ts(d2, ah + 8 * k, t_hi + 128 * k, idesc_b, 1u);
          umma_commit(bar(B_ACC));
        }
        if (it == 0) stamp(18);
      }
    }
  } else {
    // ------------------------------------------------------------------ workers (16 warps)
    pdl_wait();
    const int quarter = warp & 3, part = warp >> 2;                 // TMEM lane quarter; column part
    const int row = quarter * 32 + lane;                             // this thread's TMEM lane = matrix row
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int xr0 = tid >> 3, xc = tid & 7;                          // fill: rows xr0 and xr0 + 64, 8-channel chunk xc
    const int cw = NS / 4, nsub = cw / 16;                           // columns per part, 16-column sub-chunks (<= 3)
    uint32_t ovf = 0;
    int it = 0;
    for (int item = cta; item < items; item += nworkers, ++it) {
      const uint32_t par = (uint32_t)it & 1u;
      const int b = item / nslice, n0 = (item % nslice) * NS;
      if (it == 0 && tid == 0) stamp(0);
      // ---- the accumulators start out holding the biases: T0 = b0 w_scale, T1 = b1 w_scale, so that a0 (X W0^T + b0)
      // and A1 (X W1^T + b1) = A1 X W1^T + rowsum(A1) b1 need no per-element bias arithmetic (and no row sums) in the
      // epilogue.  The bias columns of the slice arrive by TMA with the first X block.
      mbar_wait(bar(B_BIAS), par);
#pragma unroll 1
      for (int sc = 0; sc < nsub; ++sc) {
        const int col0 = sc * 64 + part * 16;      // the columns this thread reads back in the epilogue
        uint32_t u0[16], u1[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                                         // broadcast reads
          const float4 q0 = lds128f(bias_u + (uint32_t)(col0 + 4 * j) * 4u), q1 = lds128f(bias_u + 1024u + (uint32_t)(col0 + 4 * j) * 4u);
          u0[4 * j] = __float_as_uint(q0.x * p.w_scale); u0[4 * j + 1] = __float_as_uint(q0.y * p.w_scale);
          u0[4 * j + 2] = __float_as_uint(q0.z * p.w_scale); u0[4 * j + 3] = __float_as_uint(q0.w * p.w_scale);
          u1[4 * j] = __float_as_uint(q1.x * p.w_scale); u1[4 * j + 1] = __float_as_uint(q1.y * p.w_scale);
          u1[4 * j + 2] = __float_as_uint(q1.z * p.w_scale); u1[4 * j + 3] = __float_as_uint(q1.w * p.w_scale);
        }
        tmem_st16_nowait(t_row + T0C + (uint32_t)col0, u0);
        tmem_st16_nowait(t_row + T1C + (uint32_t)col0, u1);
      }
      tmem_st_wait();
      tc_fence_before();
      // ---- X blocks: raw fp32 rows -> registers -> ("every warp has read the block") -> K-major tiles [hi16 | hi8 | lo8]
      // over the same bytes
#pragma unroll 1
      for (int g = 0; g < ng; ++g) {
        const uint32_t blk = xb + (uint32_t)g * L.blk;
        mbar_wait(bar(B_XRAW + g), par);
        if (it == 0 && tid == 0 && g == 0) stamp(19);
        // (the eight lanes of a row read 32 B each at a pitch of 32 B: lanes 4..7 take their upper 16 bytes first, so
        // that each 128-bit load of a quarter-warp covers all 32 banks once)
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0, v2 = v0, v3 = v0;   // rows >= K: zero
        const uint32_t sw = (uint32_t)(xc >> 2) * 16u;
        if (xr0 < Kt) {
          const float4 t0 = lds128f(blk + xr0 * 256 + xc * 32 + sw), t1 = lds128f(blk + xr0 * 256 + xc * 32 + (sw ^ 16u));
          v0 = sw ? t1 : t0;
          v1 = sw ? t0 : t1;
        }
        if (xr0 + 64 < Kt) {
          const float4 t0 = lds128f(blk + (xr0 + 64) * 256 + xc * 32 + sw), t1 = lds128f(blk + (xr0 + 64) * 256 + xc * 32 + (sw ^ 16u));
          v2 = sw ? t1 : t0;
          v3 = sw ? t0 : t1;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory");   // every worker holds its part of the block
        if (it == 0 && tid == 0 && g == 0) stamp(20);
        if (xr0 < kt16) conv_x_row(blk, TS, xr0, xc, v0, v1, ovf);
        if (xr0 + 64 < kt16) conv_x_row(blk, TS, xr0 + 64, xc, v2, v3, ovf);
        if (it == 0 && tid == 0 && g == 0) stamp(21);
        proxy_fence();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_XF + g));
        if (it == 0 && tid == 0) stamp(1 + g);
      }
      // diagonal of plane 0 (used by the epilogue; the load latency hides behind the wait for A1)
      if (tid < 128)
        sts32f(a0s_u + par * 512u + (uint32_t)tid * 4u,
               tid < Kt ? __ldg(p.adj + (long long)b * 2 * K * K + (long long)(row0 + tid) * K + row0 + tid) : 0.f);
      // ---- A1 row of this thread -> TMEM, packed fp16 pairs (k = 2j, 2j+1 in column j of the hi / lo ranges), zero beyond
      // K.  A part covers 32 k (K <= 128) or 64 k in two rounds (TWO: K <= 256).
      if (p.a1_tma) mbar_wait(bar(B_A1RAW), par);
#pragma unroll 1
      for (int rd = 0; rd < (TWO ? 2 : 1); ++rd) {
        const int kq = (TWO ? 64 : 32) * part + 32 * rd;              // first k of this round
        float av[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) av[j] = 0.f;
        if (p.a1_tma) {
          if (row < Kt && !(p.dbg & 1)) {
            const uint32_t src = xb + (uint32_t)(row * K + kq) * 4u;   // raw rows of K floats
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (kq + 4 * j < K) {
                const float4 t = lds128f(src + 16u * j);
                av[4 * j] = t.x; av[4 * j + 1] = t.y; av[4 * j + 2] = t.z; av[4 * j + 3] = t.w;
              }
          }
        } else if (row < Kt && !(p.dbg & 1)) {
          const float* src = p.adj + (((long long)b * 2 + 1) * K + row0 + row) * K + kq;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (kq + j < K) av[j] = __ldg(src + j);
        }
        if (kq < k16) {                                                // (warp-uniform) columns beyond k16 / 2 do not exist
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {        // 16 k values = one k-step = 8 columns at a time
            if (kq + 16 * hf >= k16) break;
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) split_pair(av[16 * hf + 2 * j], av[16 * hf + 2 * j + 1], hi[j], lo[j]);
            tmem_st8_nowait(t_row + A1H + (uint32_t)(kq / 2 + 8 * hf), hi);
            tmem_st8_nowait(t_row + A1L + (uint32_t)(kq / 2 + 8 * hf), lo);
          }
        }
      }
      tmem_st_wait();
      // every warp has read its part of the raw A1 before anybody writes T1 tiles over it
      asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory");
      if (it == 0 && tid == 0) stamp(5);
      // ---- drain T1 (scaled back by 1 / w_scale) into MN-major tiles [v][64 n]: hi tiles [0, nsg), lo tiles [nsg, 2 nsg),
      // k16 rows each; this CTA's rows are v = row0 + row (TWO: written into the peer's tiles as well); rows [K, k16) zero
      mbar_wait(bar(B_T1), par);
      tc_fence_after();
      if (it == 0 && tid == 0) stamp(6);
      if (TWO) {
        // the peer's X region (raw A1, X tiles) is dead once ITS sixteen warps have got here
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(mapa(bar(B_PEER), rank ^ 1u));
        mbar_wait_cluster(bar(B_PEER), par);
      }
      {
        float r[3][16];
#pragma unroll
        for (int sc = 0; sc < 3; ++sc)
          if (sc < nsub) tmem_ld16_nowait(t_row + T1C + (uint32_t)(part * cw + sc * 16), r[sc]);
        tmem_ld_wait();
        // rows of data, plus (last CTA of the sample) the zero rows that pad K to k16
        const bool data = row < Kt, pad = (!TWO || rank == 1u) && row >= Kt && row0 + row < k16;
        if (data || pad) {
#pragma unroll
          for (int sc = 0; sc < 3; ++sc) {
            if (sc >= nsub) break;
            const int col = part * cw + sc * 16;
            const uint32_t tile = xb + (uint32_t)(col >> 6) * TS2;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              uint32_t h[4], l[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float v0 = data ? r[sc][8 * i + 2 * j] * p.out_scale : 0.f;
                const float v1 = data ? r[sc][8 * i + 2 * j + 1] * p.out_scale : 0.f;
                split_pair(v0, v1, h[j], l[j]);
              }
              const uint32_t off = tile + swz128(row0 + row, ((col & 63) >> 3) + i);
              sts128(off, h[0], h[1], h[2], h[3]);
              sts128(off + (uint32_t)nsg * TS2, l[0], l[1], l[2], l[3]);
              if (TWO) {
                const uint32_t roff = mapa(off, rank ^ 1u);
                sts128_cluster(roff, h[0], h[1], h[2], h[3]);
                sts128_cluster(roff + (uint32_t)nsg * TS2, l[0], l[1], l[2], l[3]);
              }
            }
          }
        }
      }
      if (TWO) asm volatile("fence.proxy.async;" ::: "memory"); else proxy_fence();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(B_T1S));
        if (TWO) mbar_arrive_remote(mapa(bar(B_T1S), rank ^ 1u));
      }
      if (it == 0 && tid == 0) stamp(7);
      // ---- epilogue: this thread owns row `row` and columns [part NS/4, +NS/4) of the slice: y = relu(a0 T0 / w_scale + D2).
      // Results are assembled in shared memory (the X / T1 region) as 128-byte-row, 128B-swizzled tiles -- 32 fp32 or 64
      // fp16 columns wide: each lane writes 16-byte pieces of its own row, conflict free -- and leave through a few TMA
      // tensor stores.  (From registers, 77 KB of 16-byte stores per CTA cost 3.7 K clk at ~26 B/clk per SM and could not
      // overlap the next item; one bulk copy per row cost as much in issue time.  profiles/r03_gcn2_development.md)
      mbar_wait(bar(B_ACC), par);
      tc_fence_after();
      if (it == 0 && tid == 0) stamp(8);
      const float sa = p.out_scale * lds32f(a0s_u + par * 512u + (uint32_t)row * 4u);
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {                         // pass 0: fp32 rows of Y, pass 1: split rows
        if (pass == 0 ? p.Y == nullptr : p.split_out == nullptr) continue;
        if (pass == 1 && p.Y) {                                      // (both outputs: the tile region is used twice)
          if (warp == 0 && elect_one()) bulk_wait_read();
          asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory");
        }
        // 64 columns per round (16 per part): each round completes two fp32 tiles / one hi and one lo tile, whose stores go
        // out while the next round is computed -- only the last round's stores are exposed at the end of the kernel
#pragma unroll 1
        for (int sc = 0; sc < nsub; ++sc) {
          // (requesting the next round's accumulators before the barrier keeps 32 more registers live across it: the
          // spills that caused turned a 3.3 K clk epilogue into 9 K -- shared memory leaves almost no L1 for local memory)
          const int col0 = sc * 64 + part * 16;
          float t0[16], d2[16];
          tmem_ld16_nowait(t_row + T0C + (uint32_t)col0, t0);
          tmem_ld16_nowait(t_row + T1C + (uint32_t)col0, d2);
          tmem_ld_wait();
          if (row < Kt) {
            float y[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] = fmaxf(fmaf(t0[j], sa, d2[j]), 0.f);
            if (pass == 0) {
              const uint32_t tile = w0 + (uint32_t)(col0 >> 5) * TS;   // 32-column group
#pragma unroll
              for (int j = 0; j < 4; ++j)
                sts128(tile + swz128(row, ((col0 & 31) >> 2) + j), __float_as_uint(y[4 * j]), __float_as_uint(y[4 * j + 1]),
                       __float_as_uint(y[4 * j + 2]), __float_as_uint(y[4 * j + 3]));
            } else {
              uint32_t h[8], l[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) split_pair(y[2 * j], y[2 * j + 1], h[j], l[j]);
              const uint32_t tile = w0 + (uint32_t)sc * TS;            // 64-column group: hi tiles [0, nsg), lo tiles [nsg, 2 nsg)
              const uint32_t o0 = tile + swz128(row, part * 2), o1 = tile + swz128(row, part * 2 + 1);
              sts128(o0, h[0], h[1], h[2], h[3]);
              sts128(o1, h[4], h[5], h[6], h[7]);
              sts128(o0 + (uint32_t)nsg * TS, l[0], l[1], l[2], l[3]);
              sts128(o1 + (uint32_t)nsg * TS, l[4], l[5], l[6], l[7]);
            }
          }
          proxy_fence();                                             // generic-proxy writes -> visible to the TMA engine
          asm volatile("bar.sync 1, %0;" ::"n"(WORKERS) : "memory"); // the round's tiles are complete
          if (warp == 0 && elect_one()) {
            if (!(p.dbg & 2)) {
              if (pass == 0) {
                tma_store_2d(&tmY, w0 + (uint32_t)(2 * sc) * TS, n0 + 64 * sc, b * K + row0);
                tma_store_2d(&tmY, w0 + (uint32_t)(2 * sc + 1) * TS, n0 + 64 * sc + 32, b * K + row0);
              } else {
                tma_store_2d(&tmS, w0 + (uint32_t)sc * TS, n0 + 64 * sc, b * K + row0);
                tma_store_2d(&tmS, w0 + (uint32_t)(nsg + sc) * TS, p.split_kp + n0 + 64 * sc, b * K + row0);
              }
            }
            bulk_commit();
          }
          __syncwarp();
        }
      }
      if (it == 0 && tid == 0) stamp(9);
      if (warp == 0) {
        // the tiles have been read out: the TMA thread may load the next item's W slices over them
        if (elect_one()) {
          bulk_wait_read();
          mbar_arrive(bar(B_OFREE));
        }
        __syncwarp();
      }
      if (it == 0 && tid == 0) stamp(10);
    }
    // (the last stores have been waited for as far as their READS of shared memory go -- B_OFREE above; their writes are
    // complete when the grid is)
    __syncwarp();
    report_overflow(p.overflow, ovf);
  }
  if (threadIdx.x == 0) stamp(11);
  tc_fence_before();
  __syncthreads();
  if (TWO) cluster_sync_all();             // the peer's last remote stores / arrives have landed before this CTA's memory goes
  if (warp == WORKERS / 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

constexpr uint32_t SMEM_LIMIT = 227u * 1024u;

// Shape plan: slice width NS (0 = this kernel cannot take the shape), one or two CTAs per (sample, slice) item and the
// row / depth geometry that follows.
struct Plan {
  int NS, two, Kt, kt16, k16, kb_ret, a1_tma;
};
inline Plan make_plan(int K, int d, int dff) {
  Plan pl = {0, 0, 0, 0, 0, 0, 0};
  if (K < 1 || K > 256 || (d != 64 && d != 128 && d != 256) || dff % 64) return pl;
  const int two = K > 128;
  if (two && (K % 8 != 0)) return pl;                               // two equal row tiles whose rows start 16-byte aligned
  const int Kt = two ? K / 2 : K;
  const int kt16 = (Kt + 15) / 16 * 16, k16 = (K + 15) / 16 * 16;
  const int cand[3] = {192, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int NS = cand[i];
    if (dff % NS || NS > d) continue;                               // (single CTA: the T1 tiles live in the X region, NS <= d)
    if (2 * NS + k16 > 512) continue;                               // TMEM: T0 | T1 | A1 hi | A1 lo
    const Layout L = make_layout(kt16, k16, d, NS);
    if (L.total + 1024u > SMEM_LIMIT) continue;
    pl.NS = NS; pl.two = two; pl.Kt = Kt; pl.kt16 = kt16; pl.k16 = k16;
    // A1 (this CTA's Kt rows x K columns, raw fp32) lands in the first X blocks once GEMM A has consumed them
    const int ng = d / 64;
    const uint32_t a1_bytes = (uint32_t)Kt * (uint32_t)K * 4u;
    int kb = (int)((a1_bytes + L.blk - 1) / L.blk) - 1;
    pl.a1_tma = (K % 4 == 0) && kb <= ng - 1;
    pl.kb_ret = pl.a1_tma ? (kb < 0 ? 0 : kb) : 0;
    return pl;
  }
  return pl;
}
inline int pick_slice(int K, int d, int dff) { return make_plan(K, d, dff).NS; }

static int sm_count() {
  constexpr int MAX_DEV = 64;
  static int counts[MAX_DEV] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) return 0;
  std::lock_guard<std::mutex> lock(mu);
  if (!counts[dev]) cudaDeviceGetAttribute(&counts[dev], cudaDevAttrMultiProcessorCount, dev);
  return counts[dev];
}

}  // namespace gf2
}  // namespace ec

using namespace ec;

static long long* gf2_trace = nullptr;
static int gf2_trace_n = 0;
extern "C" int ec_gcn_fused2_set_trace(void* buf, int n_ctas) {   // profiling hook: [n_ctas][32] int64 device buffer, or NULL
  gf2_trace = (long long*)buf;
  gf2_trace_n = buf ? n_ctas : 0;
  return EC_OK;
}
static int gf2_debug = 0;
extern "C" int ec_gcn_fused2_set_debug(int flags) {   // bring-up / profiling experiments only (results are wrong when != 0)
  gf2_debug = flags;
  return EC_OK;
}
static int gf2_cta_limit = 0;
extern "C" int ec_gcn_fused2_set_cta_limit(int ctas) {   // 0 = one CTA per SM; else at most this many persistent CTAs
  EC_REQUIRE(ctas >= 0, "ec_gcn_fused2_set_cta_limit: negative");
  gf2_cta_limit = ctas;
  return EC_OK;
}

extern "C" int ec_gcn_fused2_slice(int K, int d, int dff) { return gf2::pick_slice(K, d, dff); }

extern "C" int ec_gcn_fused2(const float* X, const float* adj, const float* bias2, const void* W3, int Kp, float w_scale,
                             float* Y, void* split_out, int split_kp, int B, int K, int d, int dff, void* stream) {
  EC_REQUIRE(X && adj && bias2 && W3 && (Y || split_out), "ec_gcn_fused2: null pointer");
  const gf2::Plan pl = gf2::make_plan(K, d, dff);
  const int NS = pl.NS;
  EC_REQUIRE(NS > 0, "ec_gcn_fused2: unsupported shape (K <= 128, or K <= 256 and a multiple of 8; d in {64,128,256}; dff %% 64 == 0; shared memory / TMEM)");
  EC_REQUIRE(Kp % 64 == 0 && Kp >= 2 * d, "ec_gcn_fused2: Kp must be a multiple of 64 and >= 2d");
  EC_REQUIRE(w_scale > 0.f, "ec_gcn_fused2: bad weight scale");
  EC_REQUIRE(aligned16(X) && aligned16(adj) && aligned16(W3) && aligned16(bias2) && (!Y || aligned16(Y)) && (!split_out || aligned16(split_out)),
             "ec_gcn_fused2: operands must be 16-byte aligned");
  EC_REQUIRE(!split_out || (split_kp % 8 == 0 && split_kp >= dff), "ec_gcn_fused2: bad split_kp");
  if (B == 0) return EC_OK;
  const uint32_t smem = gf2::make_layout(pl.kt16, pl.k16, d, NS).total + 1024u;
  EC_CUDA((cudaError_t)ensure_dynamic_smem(gf2::gcn_fused2_kernel<false>, (int)gf2::SMEM_LIMIT));
  EC_CUDA((cudaError_t)ensure_dynamic_smem(gf2::gcn_fused2_kernel<true>, (int)gf2::SMEM_LIMIT));
  CUtensorMap tmW, tmX, tmY, tmS;
  int rc = tc::get_tensor_map_slice32(W3, dff, Kp, NS, &tmW);
  if (rc) return rc;
  rc = tc::get_tensor_map_f32(X, (long long)B * K, d, pl.Kt, 64, false, &tmX);   // one box = a CTA's rows of a sample x 64 channels
  if (rc) return rc;
  tmY = tmS = tmX;
  if (Y) {
    rc = tc::get_tensor_map_f32(Y, (long long)B * K, dff, pl.Kt, 32, true, &tmY);
    if (rc) return rc;
  }
  if (split_out) {
    EC_REQUIRE(split_kp % 64 == 0, "ec_gcn_fused2: split_kp must be a multiple of 64");
    rc = tc::get_tensor_map(split_out, B * K, split_kp, pl.Kt, &tmS);
    if (rc) return rc;
  }
  gf2::Params p;
  p.X = X; p.adj = adj; p.bias2 = bias2; p.Y = Y; p.split_out = (__half*)split_out; p.split_kp = split_kp;
  p.B = B; p.K = K; p.d = d; p.dff = dff; p.NS = NS; p.k16 = pl.k16; p.Kp = Kp; p.out_scale = 1.0f / w_scale; p.w_scale = w_scale;
  p.Kt = pl.Kt; p.kt16 = pl.kt16; p.kb_ret = pl.kb_ret; p.a1_tma = pl.a1_tma;
  p.overflow = overflow_counters();
  if (!p.overflow) return EC_ERR_CUDA;
  p.trace = gf2_trace; p.trace_n = gf2_trace_n; p.dbg = gf2_debug;
  const int sms = gf2::sm_count();
  EC_REQUIRE(sms > 0, "ec_gcn_fused2: no current CUDA device");
  const int items = B * (dff / NS);
  if (!pl.two) {
    int grid = items < sms ? items : sms;
    if (gf2_cta_limit > 0 && grid > gf2_cta_limit) grid = gf2_cta_limit;
    launch_pdl(gf2::gcn_fused2_kernel<false>, dim3(grid), dim3(gf2::THREADS), (size_t)smem, (cudaStream_t)stream, tmW, tmX, tmY, tmS, p);
  } else {
    int pairs = items < sms / 2 ? items : sms / 2;
    if (gf2_cta_limit > 1 && pairs > gf2_cta_limit / 2) pairs = gf2_cta_limit / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(gf2::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    EC_CUDA(cudaLaunchKernelEx(&cfg, gf2::gcn_fused2_kernel<true>, tmW, tmX, tmY, tmS, p));
  }
  return check_launch("ec_gcn_fused2");
}
