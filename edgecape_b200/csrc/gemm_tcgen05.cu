// Tensor-core GEMM for sm_100a: TMA -> shared memory (128B swizzle) -> tcgen05.mma -> TMEM -> fused
// epilogue.  fp32-grade accuracy from fp16 tensor cores by operand splitting:
//     a = a_hi + a_lo,  b = b_hi + b_lo   (fp16 pairs; |a - a_hi - a_lo| <= max(2^-22 |a|, 2^-25))
//     a.b ~= a_hi.b_hi + a_lo.b_hi + a_hi.b_lo      (fp32 accumulation in TMEM)
// Each pipeline stage holds the four 128x64 fp16 tiles {A_hi, A_lo, B_hi, B_lo} of one 64-wide k-block
// and feeds 3 x 4 UMMA 128x128x16 instructions, i.e. 4 tile loads per 3 products instead of 6.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected
// lane), warps 2..17 = epilogue (TMEM -> registers -> bias/activation/LayerScale/residual -> global, and
// optionally the split form of the result for the next GEMM).  Two TMEM accumulator stages overlap
// the epilogue of tile i with the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_fp16.h>

#include <mutex>
#include <unordered_map>

#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace ec {
namespace tc {

constexpr int BM = 128, BK = 64;                     // fp16 elements; BK * 2 B = 128 B = one swizzle row
constexpr int TILE_BYTES = BM * BK * 2;               // one 128-row operand tile: 16 KB
// Tile shapes: 128x128 (3 stages of 64 KB), 128x256 (2 stages of 96 KB) on one CTA, and 256x256 on a CTA pair
// (cta_group::2, 3 stages of 64 KB per CTA).  The main loop is bound by how many operand bytes can be in flight
// in shared memory per unit of math (profiles/r01_d): the pair tile feeds twice the math per staged byte.
constexpr int NUM_ACC = 2;
constexpr int EPI_WARPS = 16;                        // four per TMEM lane quarter (see the epilogue)
constexpr int THREADS = 64 + 32 * EPI_WARPS;         // TMA warp + MMA warp + the epilogue warps
constexpr int EPI_TILE = 32 * 16;                    // floats of one 32 x 16 staging tile (rows of 64 B, chunks XOR-swizzled)
constexpr int EPI_BYTES = EPI_WARPS * EPI_TILE * 4;  // one staging tile per epilogue warp: 32 KB
constexpr int SMEM_BYTES = 192 * 1024 + 1024 /*alignment slack*/ + 256 /*barriers*/ + EPI_BYTES;   // all variants

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
// TMA store of one box from shared memory (bulk async-group form) and its group bookkeeping
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(tmap), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * 16) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO), descriptor version 1.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                            // layout type: SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor (Cfg<BN>::IDESC): D = f32, A = B = f16, both K-major, M = 128, N = BN.

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f8f6f4 with the same instruction descriptor: a / b format 0 = E4M3, fp32 accumulate, K = 32 per instruction
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct TcParams {
  float* C;
  int M, N, num_kb, Kp, ldc, seg_c;
  long long seg_stride_c;
  int vec_c, vec_r;
  int n_fastest;   // tile order: consecutive tiles walk N first (few N tiles, large A: every A tile is read once)
  int dbg;   // experiment flags (ec_tc_set_debug): 1 = no TMA after the pipeline is primed, 2 = hi*hi only (single-CTA F16X2), 4 = no epilogue stores, 8 = no epilogue, 16 = no split_out stores, 32 = no GELU
  float out_scale;
  const float* bias;
  const float* colscale;
  const float* R;
  int ldr, act, res_mode;
  int res_rows;        // > 0: residual row = row % res_rows (broadcast over a batch of res_rows-row blocks)
  __half* split_out;   // optional [M, 2*split_kp] halves = the split rows of the result (next GEMM's A operand)
  int split_kp;
  int split_fmt;       // EC_SPLIT_F16X2 = [hi16 | lo16], EC_SPLIT_F16F8 = [hi16 | hi8 | lo8] (common.cuh)
  float split_scale;
  unsigned long long* overflow;   // F16F8 producers: {beyond e4m3, beyond fp16} event counters
  long long* trace;    // profiling (ec_tc_set_trace): per leader CTA {MMA thread total clk, clk waiting for an accumulator, clk waiting for operands, tiles}
  int split_tma;       // split_out is the only output: the epilogue assembles 128 x 64 blocks in shared memory and TMA-stores them
  int* sched;          // dynamic tile scheduler: {next tile, finished workers} of this launch (zero on entry), or NULL
  // K-split of the TAIL tiles (ksplit > 1; fp32-output epilogue without activation only).  Work units [0, whole_units) are
  // whole tiles; every later tile is ksplit units of num_kb / ksplit k-blocks each.  Unit p of a tile accumulates onto
  // what unit p - 1 wrote -- C += colscale (acc out_scale), in that fixed order, so the result does not depend on which
  // worker ran what -- and waits for it through tile_flags[tile - whole_units] (one arrival per epilogue warp).  With
  // the tail cut in thirds a 1.66-wave GEMM (fc2) costs 1.67 tile times instead of 2.
  int ksplit, whole_units, tail_tiles;
  int* tile_flags;
};

// ---- cluster / 2-CTA helpers (cta_group::2: one UMMA spans the tensor cores and shared memory of an SM pair)
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
__device__ __forceinline__ void cluster_sync() {
  __syncwarp();   // .aligned: the whole warp must arrive together (role lanes re-converge here)
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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
__device__ __forceinline__ void st_shared_cluster(uint32_t cluster_addr, int v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Relaxed form for hand-backs that publish no memory: "this warp has finished READING" a TMEM accumulator
// (tcgen05.wait::ld has completed the loads) or a ring slot.  The release form above compiles to MEMBAR.ALL.GPU +
// ERRBAR in front of the arrive, i.e. it waits for every global store the warp has in flight -- once per tile and
// epilogue warp, exactly where the epilogue is the bound.  `dep` is a value the arrive must not overtake (the word
// read from the ring slot): it is folded into the address so the instruction cannot issue before the load returns.
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t cluster_addr, int dep = 0) {
  asm volatile(
      "{\n\t.reg .b32 t;\n\t"
      "and.b32 t, %1, 0;\n\t"
      "add.u32 t, t, %0;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [t];\n\t}"
      ::"r"(cluster_addr), "r"(dep)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const void* tmap, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {   // arrives on `bar` of BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

struct Unit {
  int tile, kb0, nkb, part;   // part: -1 = a whole tile, else the K-part of a split tile
};
// Units are numbered part-major -- part 0 of every tail tile, then part 1 of every tail tile, ... -- so that a part's
// predecessor was handed out a whole round of tail tiles earlier and has (nearly always) stored its result by the time the
// part's own main loop ends.  (Tile-major numbering put the three parts of a tile on three workers at the same moment: each
// epilogue then waited for another worker's, which waited for a third -- chains through the in-order epilogues that made
// fc2 38 % SLOWER than without the split.)
__device__ __forceinline__ Unit decode_unit(int u, const TcParams& p) {
  if (p.ksplit <= 1 || u < p.whole_units) return {u, 0, p.num_kb, -1};
  const int v = u - p.whole_units, n = p.num_kb / p.ksplit, tail = p.tail_tiles;
  const int part = v / tail;
  return {p.whole_units + (v - part * tail), part * n, n, part};
}

// TWO = false: one CTA per 128 x BN tile (cta_group::1).
// TWO = true : a cluster of two CTAs per 256 x 256 tile (cta_group::2, BN must be 256): each CTA stages its own
//              128 rows of A and 128 rows of B (64 KB per k-block, 3 stages), the leader issues M = 256 UMMAs that
//              read both CTAs' shared memory, so every staged byte feeds twice the math of the single-CTA tile.
// F8 = false: operands in F16X2 rows, three kind::f16 products per k-block (12 UMMAs of K = 16).
// F8 = true : operands in F16F8 rows; a stage holds {A.hi16, A.hi8, A.lo8, B.hi16, B.hi8, B.lo8} (the same bytes), the
//             fp16 tiles in 128B swizzle and the e4m3 tiles in 64B swizzle (64-byte rows, loaded through the byte
//             view tmA8 / tmB8 of the same buffers); a k-block is 2 + 2 kind::f8f6f4 UMMAs (K = 32) for the cross terms
//             a_lo.b_hi + a_hi.b_lo and 4 kind::f16 UMMAs for a_hi.b_hi: 2 instead of 3 units of tensor time.
template <int BN, bool TWO, bool F8>
__global__ void __launch_bounds__(THREADS, 1)
gemm_f16x3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmA8, const __grid_constant__ CUtensorMap tmB8,
                  const __grid_constant__ CUtensorMap tmO16, const __grid_constant__ CUtensorMap tmO8, TcParams p) {
  constexpr int B_ROWS = TWO ? BN / 2 : BN;                      // B rows staged by one CTA
  constexpr int B_TILE_BYTES = B_ROWS * BK * 2;
  constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * B_TILE_BYTES;
  constexpr int STAGES = STAGE_BYTES <= 64 * 1024 ? 3 : 2;
  constexpr int TMEM_COLS = 2 * BN;
  constexpr int TM = TWO ? 2 * BM : BM;                           // rows of the (pair) tile
  constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  static_assert(!TWO || BN == 256, "the 2-CTA kernel uses 256 x 256 tiles");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  // after the pipeline stages: the epilogue staging area (1024-byte aligned: the split-only epilogue builds 128B- /
  // 64B-swizzled tiles in it for TMA stores), then the barriers
  const uint32_t epi_base = base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  // barriers: full[STAGES], empty[STAGES], tmem_full[NUM_ACC], tmem_empty[NUM_ACC]; then the TMEM address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + NUM_ACC + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * NUM_ACC);
  // tile ring: the TMA thread of the (leader) CTA draws tile indices -- from a global counter when p.sched is set, so a
  // CTA that starts late (its SM was busy with another stream's kernel) simply takes fewer tiles -- and hands them to
  // the MMA thread, the epilogue warps and the peer CTA through RING slots: ring_full (1 producer arrival) /
  // ring_empty on the leader (every consumer of both CTAs arrives once it has read the slot).  -1 ends the kernel.
  constexpr int RING = 4;
  auto ring_full = [&](int r) { return bar_base + 96u + 8u * r; };
  auto ring_empty = [&](int r) { return bar_base + 128u + 8u * r; };
  const uint32_t ring_tile = bar_base + 160u;                       // RING x int32
  volatile int* ring_tile_ptr = reinterpret_cast<volatile int*>(smem_raw + (ring_tile - smem_u32(smem_raw)));
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = TWO ? (int)cluster_ctarank() : 0;              // 0 = leader of the pair
  const int worker = TWO ? (int)blockIdx.x / 2 : (int)blockIdx.x;
  const int num_workers = TWO ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int m_tiles = (p.M + TM - 1) / TM, n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_units = p.ksplit > 1 ? p.whole_units + (num_tiles - p.whole_units) * p.ksplit : num_tiles;

  if (warp == 0 && lane == 1) {
    // the TMA unit fetches a descriptor on its first use: start those fetches now, under the barrier set-up
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (F8) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA8) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB8) : "memory");
    }
    if (p.split_tma) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO16) : "memory");
      if (p.split_fmt == EC_SPLIT_F16F8) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO8) : "memory");
    }
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < NUM_ACC; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), (TWO ? 2 : 1) * EPI_WARPS); }
    // ring consumers: the epilogue warps and the MMA thread of every CTA, plus the peer's TMA thread
    for (int r = 0; r < RING; ++r) { mbar_init(ring_full(r), 1); mbar_init(ring_empty(r), TWO ? 2 * EPI_WARPS + 2 : EPI_WARPS + 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (TWO) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  if (TWO) cluster_sync(); else __syncthreads();                  // peer barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_launch_dependents();     // TMEM is ours: the next kernel in the stream may start its prologue
  pdl_wait();                  // ... and ours ends here: operands written by the previous kernel are visible

  // consumer side of the tile ring: j-th tile of this CTA (pair); one arrival per consumer on the leader's ring_empty
  auto ring_get = [&](int j) -> int {
    const int r = j % RING;
    const uint32_t ph = (uint32_t)(j / RING) & 1u;
    if (TWO && rank == 1) mbar_wait_cluster(ring_full(r), ph); else mbar_wait(ring_full(r), ph);
    return ring_tile_ptr[r];
  };
  auto ring_release = [&](int j, int tile_read) {   // tile_read: the slot's value (the arrive must follow the read)
    const int r = j % RING;
    if (TWO && rank == 1) mbar_arrive_remote_relaxed(mapa(ring_empty(r), 0), tile_read); else mbar_arrive(ring_empty(r));
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one per CTA)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0;; ++j) {
        int tile;
        if (rank == 0) {
          // tile scheduler (leader): draw the next tile, publish it to this CTA and to the peer
          const int r = j % RING;
          if (j >= RING) {
            const uint32_t ph = (uint32_t)(j / RING - 1) & 1u;
            if (TWO) mbar_wait_cluster(ring_empty(r), ph); else mbar_wait(ring_empty(r), ph);
          }
          // dynamic scheduling from the SECOND tile on: the first one is the worker's own index (a launch with dynamic tiles has
          // more tiles than workers), so no global atomic round trip sits in front of the first TMA load
          tile = (p.sched && j > 0) ? num_workers + atomicAdd(p.sched, 1) : worker + j * num_workers;
          if (tile >= num_units) tile = -1;
          ring_tile_ptr[r] = tile;
          if (TWO) st_shared_cluster(mapa(ring_tile + 4u * r, 1), tile);
          mbar_arrive(ring_full(r));
          if (TWO) mbar_arrive_remote(mapa(ring_full(r), 1));
        } else {
          tile = ring_get(j);
          ring_release(j, tile);
        }
        if (tile < 0) break;
        const Unit un = decode_unit(tile, p);
        const int tm = p.n_fastest ? un.tile / n_tiles : un.tile % m_tiles, tn = p.n_fastest ? un.tile % n_tiles : un.tile / m_tiles;
        const int m0 = tm * TM + rank * BM, n0 = tn * BN + rank * B_ROWS;
        for (int kb = un.kb0; kb < un.kb0 + un.nkb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sb = base + stage * STAGE_BYTES;
          if ((p.dbg & 1) && (j > 0 || kb - un.kb0 >= STAGES)) {   // experiment: operands stay resident after the first fill
            if (rank == 0) mbar_arrive(full_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          // stage: [A.hi16 16 KB | A second plane(s) 16 KB | B.hi16 | B second plane(s)]; the second plane is lo16
          // (one box of halves) or hi8 | lo8 (two boxes of bytes at byte columns 2 Kp / 3 Kp of the same rows)
          if (TWO) {
            // both CTAs' bytes are accounted on the leader's barrier: it expects 2 x STAGE_BYTES
            const uint32_t lbar = mapa(full_bar(stage), 0);
            if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * STAGE_BYTES);
            tma_load_2d_2sm(sb + 0 * TILE_BYTES, &tmA, lbar, kb * BK, m0);
            tma_load_2d_2sm(sb + 2 * TILE_BYTES, &tmB, lbar, kb * BK, n0);
            if (F8) {     // both e4m3 planes of the k-block: ONE box of 128-byte rows [hi8 x 64 | lo8 x 64] per operand
              tma_load_2d_2sm(sb + TILE_BYTES, &tmA8, lbar, 2 * p.Kp + 2 * kb * BK, m0);
              tma_load_2d_2sm(sb + 2 * TILE_BYTES + B_TILE_BYTES, &tmB8, lbar, 2 * p.Kp + 2 * kb * BK, n0);
            } else {
              tma_load_2d_2sm(sb + 1 * TILE_BYTES, &tmA, lbar, p.Kp + kb * BK, m0);
              tma_load_2d_2sm(sb + 2 * TILE_BYTES + B_TILE_BYTES, &tmB, lbar, p.Kp + kb * BK, n0);
            }
          } else {
            mbar_expect_tx(full_bar(stage), STAGE_BYTES);
            tma_load_2d(sb + 0 * TILE_BYTES, &tmA, full_bar(stage), kb * BK, m0);
            tma_load_2d(sb + 2 * TILE_BYTES, &tmB, full_bar(stage), kb * BK, n0);
            if (F8) {
              tma_load_2d(sb + TILE_BYTES, &tmA8, full_bar(stage), 2 * p.Kp + 2 * kb * BK, m0);
              tma_load_2d(sb + 2 * TILE_BYTES + B_TILE_BYTES, &tmB8, full_bar(stage), 2 * p.Kp + 2 * kb * BK, n0);
            } else {
              tma_load_2d(sb + 1 * TILE_BYTES, &tmA, full_bar(stage), p.Kp + kb * BK, m0);
              tma_load_2d(sb + 2 * TILE_BYTES + B_TILE_BYTES, &tmB, full_bar(stage), p.Kp + kb * BK, n0);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      long long tr_t0 = 0, tr_acc = 0, tr_full = 0;
      int tr_tiles = 0;
      if (p.trace) tr_t0 = clock64();
      for (int t = 0;; ++t) {
        const int tile = ring_get(t);
        ring_release(t, tile);
        if (tile < 0) break;
        const int acc = t & 1;
        long long c0 = 0;
        if (p.trace) c0 = clock64();
        mbar_wait(tempty_bar(acc), ((t >> 1) & 1) ^ 1);
        if (p.trace) { tr_acc += clock64() - c0; ++tr_tiles; }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        const int nkb = decode_unit(tile, p).nkb;                  // (a K-part of a split tail tile: fewer k-blocks)
        for (int kb = 0; kb < nkb; ++kb) {
          if (p.trace) c0 = clock64();
          mbar_wait(full_bar(stage), phase);
          if (p.trace) tr_full += clock64() - c0;
          tc_fence_after();
          const uint32_t sb = base + stage * STAGE_BYTES;
          const uint64_t a_hi = make_smem_desc(sb + 0 * TILE_BYTES), b_hi = make_smem_desc(sb + 2 * TILE_BYTES);
          if (F8) {
            // cross terms first (e4m3, K = 32 per UMMA, descriptors advance 32 B), then hi16 . hi16
            // e4m3 tiles: rows of 128 B in 128B swizzle, hi8 in bytes [0, 64) and lo8 in [64, 128) of each row -- the lo8 operand
            // is the same tile entered 64 bytes (4 x 16 B) into the row, like a K advance
            const uint64_t a_h8 = make_smem_desc(sb + TILE_BYTES), a_l8 = a_h8 + 4;
            const uint64_t b_h8 = make_smem_desc(sb + 2 * TILE_BYTES + B_TILE_BYTES), b_l8 = b_h8 + 4;
            if (TWO) {
#pragma unroll
              for (int k = 0; k < BK / 32; ++k) umma_f8_2sm(tmem_d, a_l8 + 2 * k, b_h8 + 2 * k, IDESC, (kb | k) ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < BK / 32; ++k) umma_f8_2sm(tmem_d, a_h8 + 2 * k, b_l8 + 2 * k, IDESC, 1u);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_f16_2sm(tmem_d, a_hi + 2 * k, b_hi + 2 * k, IDESC, 1u);
              umma_commit_2sm(empty_bar(stage));
            } else {
#pragma unroll
              for (int k = 0; k < BK / 32; ++k) umma_f8(tmem_d, a_l8 + 2 * k, b_h8 + 2 * k, IDESC, (kb | k) ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < BK / 32; ++k) umma_f8(tmem_d, a_h8 + 2 * k, b_l8 + 2 * k, IDESC, 1u);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_d, a_hi + 2 * k, b_hi + 2 * k, IDESC, 1u);
              umma_commit(empty_bar(stage));
            }
          } else {
          const uint64_t a_lo = make_smem_desc(sb + 1 * TILE_BYTES);
          const uint64_t b_lo = make_smem_desc(sb + 2 * TILE_BYTES + B_TILE_BYTES);
          // small terms first, then hi*hi
          if (TWO) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_f16_2sm(tmem_d, a_lo + 2 * k, b_hi + 2 * k, IDESC, (kb | k) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_f16_2sm(tmem_d, a_hi + 2 * k, b_lo + 2 * k, IDESC, 1u);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_f16_2sm(tmem_d, a_hi + 2 * k, b_hi + 2 * k, IDESC, 1u);
            umma_commit_2sm(empty_bar(stage));       // frees the stage in both CTAs when these MMAs retire
          } else {
            if (!(p.dbg & 2)) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_d, a_lo + 2 * k, b_hi + 2 * k, IDESC, (kb | k) ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_d, a_hi + 2 * k, b_lo + 2 * k, IDESC, 1u);
            }
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_f16(tmem_d, a_hi + 2 * k, b_hi + 2 * k, IDESC, ((p.dbg & 2) && !(kb | k)) ? 0u : 1u);
            umma_commit(empty_bar(stage));           // frees the smem stage when these MMAs retire
          }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (TWO) umma_commit_2sm(tfull_bar(acc)); else umma_commit(tfull_bar(acc));   // accumulator complete
      }
      if (p.trace) {
        long long* o = p.trace + 4 * worker;
        o[0] = clock64() - tr_t0; o[1] = tr_acc; o[2] = tr_full; o[3] = tr_tiles;
      }
    }
  } else {
    // ---------------------------------------------------------------------- epilogue (16 warps)
    // TMEM gives each lane one accumulator row.  Four warps share a lane quarter and take every fourth 16-column
    // sub-chunk; each sub-chunk is transposed through a per-warp shared-memory tile (32 rows of 64 B, the 16-byte
    // chunk c of row r stored at chunk c ^ ((r >> 1) & 3): conflict-free float4 writes by row and reads by
    // (8 rows x 4 chunks)) so that global traffic is row-contiguous: 4 lanes cover one 16-column row segment
    // (64 B), a warp covers 8 rows per step.  With the cross terms on e4m3 the main loop of a K = 768 tile is
    // ~15 K clk and the epilogue -- a chain of TMEM / shared / global latencies per sub-chunk, ~60 instructions
    // per element with erf -- had become the bound of the qkv and fc1 GEMMs with 8 warps (fc1: 132 us against
    // 68 us with the epilogue switched off, profiles/r02_d_gemm_f16f8_experiment_flags.log): 16 warps halve the
    // chain every warp walks per tile and give each scheduler four warps to overlap.
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int group = (warp - 2) >> 2;             // which of the four warps of the quarter
    constexpr int NGROUP = EPI_WARPS / 4;
    float* stage = reinterpret_cast<float*>(smem_raw + (epi_base - smem_u32(smem_raw))) + (warp - 2) * EPI_TILE;
    const int sub_row = lane >> 2, c4 = (lane & 3) * 4;
    const int st_swz = (lane >> 1) & 3;            // this lane's row as a writer: chunk j -> j ^ st_swz
    constexpr int NSUB = BN / 16;
    int t = 0;
    uint32_t ovf = 0;                              // F16F8 split_out: values beyond the e4m3 / fp16 range seen by this thread
    const uint32_t tempty_leader0 = TWO ? mapa(tempty_bar(0), 0) : 0u, tempty_leader1 = TWO ? mapa(tempty_bar(1), 0) : 0u;
    for (;; ++t) {
      const int tile = ring_get(t);
      __syncwarp();                                  // every lane has read the slot
      if (lane == 0) ring_release(t, tile);
      if (tile < 0) break;
      const int acc = t & 1;
      const Unit un = decode_unit(tile, p);
      const int tm = p.n_fastest ? un.tile / n_tiles : un.tile % m_tiles, tn = p.n_fastest ? un.tile % n_tiles : un.tile / m_tiles;
      const int m0 = tm * TM + rank * BM, n0 = tn * BN;
      if (p.split_tma) {
        // ------------------------------------------------------------ split-only output through TMA stores
        // The qkv / fc1 GEMMs write nothing but the split form of their result.  Scattered from the lanes (8- and
        // 4-byte pieces, eight rows per store instruction) those stores were 32 of fc1's 110 us against 69 us with
        // no epilogue (profiles/r02_f_gemm_epilogue_decomposition.log).  Here the 16 warps assemble the rows of a
        // 128-row x 64-column block in shared memory in the layout TMA expects -- the very layout the consumer's
        // A-operand tiles have: hi16 rows of 128 B in 128B swizzle, lo16 likewise, e4m3 planes rows of 64 B in 64B
        // swizzle -- each lane converting the 16 columns of its own accumulator row (no transposition: the tile IS
        // the transposition, and consecutive rows hit distinct bank groups by construction of the swizzle), and
        // one thread stores the block with two or three bulk tensor copies (whole 128-byte lines; rows >= M are
        // clipped by the tensor map, columns >= N inside the padded width are the zero padding of the operand).
        uint8_t* blk = smem_raw + (epi_base - smem_u32(smem_raw));      // [hi16 16 KB | lo16 16 KB] or [hi16 16 KB | hi8 + lo8 16 KB]
        const uint32_t blk_u32 = epi_base;
        const int row_l = quarter * 32 + lane;                          // row of the 128-row block = TMEM lane
        // bias / LayerScale of a block: ONE coalesced load per warp (lane u holds column u of the 16), fetched a block
        // ahead of its use -- shared memory takes nearly all of the L1 carve-out, so these loads come from L2
        // (~700 clk); read at the point of use (16 loads per lane and round) they were the largest single stall of
        // the epilogue warps (ncu source page, profiles/r02_q_ncu_fc1_stalls.md)
        auto load_vec = [&](const float* v, int cb, float dflt) {
          const int c = n0 + cb * 64 + group * 16 + (lane & 15);
          return (v && c < p.N) ? __ldg(v + c) : dflt;
        };
        float bias_nxt = load_vec(p.bias, 0, 0.f), cs_nxt = load_vec(p.colscale, 0, 1.f);
        mbar_wait(tfull_bar(acc), (t >> 1) & 1);
        tc_fence_after();
        if (p.dbg & 8) {   // experiment: no epilogue at all (the accumulator is handed straight back)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (TWO && rank == 1) mbar_arrive_remote_relaxed(acc ? tempty_leader1 : tempty_leader0);
            else mbar_arrive(tempty_bar(acc));
          }
          continue;
        }
        constexpr int NBLK = BN / 64;
#pragma unroll 1
        for (int cb = 0; cb < NBLK; ++cb) {
          const int col0 = n0 + cb * 64;                                // first column of the block
          if (col0 >= p.split_kp) break;                                // warp-uniform: nothing of the operand lies here
          uint32_t r[16];
          tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + cb * 64 + group * 16), r);
          if (cb + 1 == NBLK || col0 + 64 >= p.split_kp) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (TWO && rank == 1) mbar_arrive_remote_relaxed(acc ? tempty_leader1 : tempty_leader0);
              else mbar_arrive(tempty_bar(acc));
            }
          }
          const int gcol = col0 + group * 16;
          const float bias_cur = bias_nxt, cs_cur = cs_nxt;
          if (cb + 1 < NBLK) {
            bias_nxt = load_vec(p.bias, cb + 1, 0.f);
            cs_nxt = load_vec(p.colscale, cb + 1, 1.f);
          }
          float y[16];
#pragma unroll
          for (int u = 0; u < 16; ++u)
            y[u] = fmaf(__uint_as_float(r[u]), p.out_scale, __shfl_sync(0xffffffffu, bias_cur, u));
          if (p.act == EC_ACT_RELU) {
#pragma unroll
            for (int u = 0; u < 16; ++u) y[u] = fmaxf(y[u], 0.f);
          } else if (p.act == EC_ACT_GELU && !(p.dbg & 32)) {
#pragma unroll
            for (int u = 0; u < 16; ++u) y[u] = gelu_fast(y[u]);
          } else if (p.act == EC_ACT_TANH) {
#pragma unroll
            for (int u = 0; u < 16; ++u) y[u] = tanhf(y[u]);
          }
          if (p.colscale) {
#pragma unroll
            for (int u = 0; u < 16; ++u) y[u] *= __shfl_sync(0xffffffffu, cs_cur, u);
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) y[u] = (gcol + u < p.N) ? y[u] * p.split_scale : 0.f;
          uint32_t h[8];
          float2 f[8];                                                  // the low parts y - hi16 (exact; one FHADD each)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const __half2 hh = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
            h[i] = *reinterpret_cast<const uint32_t*>(&hh);
            sub_h2(y[2 * i], y[2 * i + 1], h[i], f[i].x, f[i].y);
          }
          // the previous block has been read out of shared memory: the issuing thread waits for its bulk stores' reads
          // only here, after its own conversions, so the TMA engine drains the tile under everybody's arithmetic
          if (p.split_tma == 2) {
            // experiment / alternative: every lane stores the pieces of its own row straight from registers -- no
            // shared-memory traffic at all (the F16F8 main loop already moves ~106 of the 128 B/clk of shared-memory
            // bandwidth: 64 KB of TMA fill + 64 KB of UMMA operand reads per 1208-clk k-block), no CTA-wide barriers;
            // 32 rows x 16 B per store instruction
            const int grow = m0 + row_l;
            if (grow < p.M && !(p.dbg & 16)) {
              uint8_t* orow = reinterpret_cast<uint8_t*>(p.split_out) + (long long)grow * (4 * p.split_kp);
              *reinterpret_cast<uint4*>(orow + 2 * gcol) = make_uint4(h[0], h[1], h[2], h[3]);
              *reinterpret_cast<uint4*>(orow + 2 * gcol + 16) = make_uint4(h[4], h[5], h[6], h[7]);
              if (p.split_fmt == EC_SPLIT_F16X2) {
                uint32_t l[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const __half2 ll = __floats2half2_rn(f[i].x, f[i].y);
                  l[i] = *reinterpret_cast<const uint32_t*>(&ll);
                }
                *reinterpret_cast<uint4*>(orow + 2 * p.split_kp + 2 * gcol) = make_uint4(l[0], l[1], l[2], l[3]);
                *reinterpret_cast<uint4*>(orow + 2 * p.split_kp + 2 * gcol + 16) = make_uint4(l[4], l[5], l[6], l[7]);
              } else {
                uint32_t h8[4], l8[4];
                float m = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  h8[i] = e4m3x2_h2(h[2 * i]) | (e4m3x2_h2(h[2 * i + 1]) << 16);
                  l8[i] = e4m3x2(f[2 * i].x * 2048.f, f[2 * i].y * 2048.f) | (e4m3x2(f[2 * i + 1].x * 2048.f, f[2 * i + 1].y * 2048.f) << 16);
                  m = fmaxf(m, fmaxf(fmaxf(fabsf(y[4 * i]), fabsf(y[4 * i + 1])), fmaxf(fabsf(y[4 * i + 2]), fabsf(y[4 * i + 3]))));
                }
                ovf |= (m > 448.f ? 1u : 0u) | (m > 65504.f ? 2u : 0u);
                *reinterpret_cast<uint4*>(orow + f8_off(p.split_kp, gcol)) = make_uint4(h8[0], h8[1], h8[2], h8[3]);
                *reinterpret_cast<uint4*>(orow + f8_off(p.split_kp, gcol) + 64) = make_uint4(l8[0], l8[1], l8[2], l8[3]);
              }
            }
            continue;
          }
          if (cb > 0 || t > 0) {
            if (warp == 2) tma_store_wait_read();
            epi_bar_sync();
          }
          // hi16: 16 columns = two 16-byte chunks (2 group, 2 group + 1) of the row's 128 bytes
          uint8_t* hrow = blk + row_l * 128;
          *reinterpret_cast<uint4*>(hrow + (((2 * group) ^ (row_l & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(hrow + (((2 * group + 1) ^ (row_l & 7)) << 4)) = make_uint4(h[4], h[5], h[6], h[7]);
          if (p.split_fmt == EC_SPLIT_F16X2) {
            uint32_t l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const __half2 ll = __floats2half2_rn(f[i].x, f[i].y);
              l[i] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            uint8_t* lrow = blk + 16384 + row_l * 128;
            *reinterpret_cast<uint4*>(lrow + (((2 * group) ^ (row_l & 7)) << 4)) = make_uint4(l[0], l[1], l[2], l[3]);
            *reinterpret_cast<uint4*>(lrow + (((2 * group + 1) ^ (row_l & 7)) << 4)) = make_uint4(l[4], l[5], l[6], l[7]);
          } else {
            uint32_t h8[4], l8[4];
            float m = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              h8[i] = e4m3x2_h2(h[2 * i]) | (e4m3x2_h2(h[2 * i + 1]) << 16);
              l8[i] = e4m3x2(f[2 * i].x * 2048.f, f[2 * i].y * 2048.f) | (e4m3x2(f[2 * i + 1].x * 2048.f, f[2 * i + 1].y * 2048.f) << 16);
              m = fmaxf(m, fmaxf(fmaxf(fabsf(y[4 * i]), fabsf(y[4 * i + 1])), fmaxf(fabsf(y[4 * i + 2]), fabsf(y[4 * i + 3]))));
            }
            ovf |= (m > 448.f ? 1u : 0u) | (m > 65504.f ? 2u : 0u);
            // e4m3 planes: ONE tile of 128-byte rows (128B swizzle), hi8 of the block's 64 columns in chunks [0, 4), lo8 in
            // chunks [4, 8): this lane's 16 columns are chunk `group` of each
            *reinterpret_cast<uint4*>(blk + 16384 + row_l * 128 + ((group ^ (row_l & 7)) << 4)) = make_uint4(h8[0], h8[1], h8[2], h8[3]);
            *reinterpret_cast<uint4*>(blk + 16384 + row_l * 128 + (((4 + group) ^ (row_l & 7)) << 4)) = make_uint4(l8[0], l8[1], l8[2], l8[3]);
          }
          fence_proxy_async_smem();                                     // generic-proxy writes -> visible to the TMA engine
          epi_bar_sync();                                               // the block is complete
          if (warp == 2 && elect_one()) {
            if (!(p.dbg & 16)) {
              tma_store_2d(&tmO16, blk_u32, col0, m0);
              if (p.split_fmt == EC_SPLIT_F16X2) {
                tma_store_2d(&tmO16, blk_u32 + 16384, p.split_kp + col0, m0);
              } else {
                tma_store_2d(&tmO8, blk_u32 + 16384, 2 * p.split_kp + 2 * col0, m0);   // byte column 2 kp + 128 (col0 / 64)
              }
              tma_store_commit();                                       // (reads awaited at the top of the next block)
            }
          }
          __syncwarp();
        }
        continue;
      }
      // residual rows of one sub-chunk (4 rows x 4 columns per lane).  They do not depend on the accumulator, so
      // the first sub-chunk's are requested before waiting for the MMAs and every later one a sub-chunk ahead:
      // loaded at the point of use, the four dependent L2 round trips per sub-chunk (the in-place store to C
      // may alias R, so the compiler cannot hoist them) were the exposed tail of the narrow GEMMs.
      // K-parts of a split tail tile: part 0 is the ordinary epilogue; a later part adds colscale (acc out_scale) to what the
      // previous part stored -- its "residual" is C, without bias -- once that part's stores are visible
      const bool later_part = un.part > 0;
      const bool has_res = later_part || p.R != nullptr;
      if (later_part) {
        const int need = un.part * EPI_WARPS * (TWO ? 2 : 1);      // one arrival per epilogue warp of every earlier part
        if (lane == 0) {
          const int* f = p.tile_flags + (un.tile - p.whole_units);
          int v;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if (v < need) __nanosleep(64);
          } while (v < need);
        }
        __syncwarp();
      }
      auto load_res = [&](int sc, float4* out) {
        const int gcol = n0 + sc * 16 + c4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          out[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          const int row = m0 + quarter * 32 + i * 8 + sub_row;
          if (sc >= NSUB || gcol >= p.N || row >= p.M) continue;
          const float* rp = later_part
                                ? (p.seg_c > 0 ? p.C + (long long)(row / p.seg_c) * p.seg_stride_c + (long long)(row % p.seg_c) * p.ldc
                                               : p.C + (long long)row * p.ldc) + gcol
                                : p.R + (long long)(p.res_rows > 0 ? row % p.res_rows : row) * p.ldr + gcol;
          if (gcol + 3 < p.N && (later_part ? p.vec_c : p.vec_r)) {
            out[i] = *reinterpret_cast<const float4*>(rp);
          } else {
            out[i].x = rp[0];
            if (gcol + 1 < p.N) out[i].y = rp[1];
            if (gcol + 2 < p.N) out[i].z = rp[2];
            if (gcol + 3 < p.N) out[i].w = rp[3];
          }
        }
      };
      // bias / LayerScale of a sub-chunk (4 columns per lane), fetched a sub-chunk ahead like the residual rows: with
      // shared memory taking the L1 carve-out these are L2 round trips (~700 clk), exposed once per sub-chunk when
      // read at the point of use
      auto load_bs = [&](int sc, float4& b4, float4& s4) {
        const int gc = n0 + sc * 16 + c4;
        float bb[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {1.f, 1.f, 1.f, 1.f};
        if (sc < NSUB) {
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (gc + u < p.N) {
              if (p.bias && !later_part) bb[u] = __ldg(p.bias + gc + u);
              if (p.colscale) ss[u] = __ldg(p.colscale + gc + u);
            }
        }
        b4 = make_float4(bb[0], bb[1], bb[2], bb[3]);
        s4 = make_float4(ss[0], ss[1], ss[2], ss[3]);
      };
      float4 rcur[4], rnext[4], bcur, scur, bnext, snext;
      load_bs(group, bcur, scur);
      if (has_res) load_res(group, rcur);
      mbar_wait(tfull_bar(acc), (t >> 1) & 1);
      tc_fence_after();
      if (p.dbg & 8) {   // experiment: no epilogue at all (the accumulator is handed straight back)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (TWO && rank == 1) mbar_arrive_remote_relaxed(acc ? tempty_leader1 : tempty_leader0);
          else mbar_arrive(tempty_bar(acc));
        }
        continue;
      }
#pragma unroll 1
      for (int sc = group; sc < NSUB; sc += NGROUP) {
        uint32_t r[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + sc * 16), r);
        if (has_res) load_res(sc + NGROUP, rnext);
        load_bs(sc + NGROUP, bnext, snext);
        if (sc + NGROUP >= NSUB) {
          // this warp has drained its share of the accumulator: hand it back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {   // the leader's MMA thread waits for every epilogue warp of the pair
            if (TWO && rank == 1) mbar_arrive_remote_relaxed(acc ? tempty_leader1 : tempty_leader0);
            else mbar_arrive(tempty_bar(acc));
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<float4*>(stage + lane * 16 + ((j ^ st_swz) << 2)) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
        __syncwarp();
        const int gcol = n0 + sc * 16 + c4;
        if (gcol < p.N) {
          const bool full = gcol + 3 < p.N;
          const float bv[4] = {bcur.x, bcur.y, bcur.z, bcur.w}, sv[4] = {scur.x, scur.y, scur.z, scur.w};
          float y[4][4];
          // accumulator * out_scale + bias, activation (transcendental ones in a rolled loop: code size)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = i * 8 + sub_row;
            const float4 a4 = *reinterpret_cast<const float4*>(stage + rr * 16 + (((c4 >> 2) ^ ((rr >> 1) & 3)) << 2));
            y[i][0] = fmaf(a4.x, p.out_scale, bv[0]); y[i][1] = fmaf(a4.y, p.out_scale, bv[1]);
            y[i][2] = fmaf(a4.z, p.out_scale, bv[2]); y[i][3] = fmaf(a4.w, p.out_scale, bv[3]);
          }
          if (p.act == EC_ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int u = 0; u < 4; ++u) y[i][u] = fmaxf(y[i][u], 0.f);
          } else if (p.act == EC_ACT_GELU && !(p.dbg & 32)) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int u = 0; u < 4; ++u) y[i][u] = gelu_fast(y[i][u]);
          } else if (p.act == EC_ACT_TANH) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int u = 0; u < 4; ++u) y[i][u] = tanhf(y[i][u]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int row = m0 + quarter * 32 + i * 8 + sub_row;
            if (row >= p.M) continue;
            if (p.colscale) {
#pragma unroll
              for (int u = 0; u < 4; ++u) y[i][u] *= sv[u];
            }
            if (has_res) {
              const float rr[4] = {rcur[i].x, rcur[i].y, rcur[i].z, rcur[i].w};
#pragma unroll
              for (int u = 0; u < 4; ++u)
                y[i][u] = (p.res_mode == EC_RES_GATE && !later_part) ? (y[i][u] + 1.0f) * rr[u] : rr[u] + y[i][u];
            }
            if (p.C && !(p.dbg & 4)) {
              float* cp = (p.seg_c > 0 ? p.C + (long long)(row / p.seg_c) * p.seg_stride_c + (long long)(row % p.seg_c) * p.ldc
                                       : p.C + (long long)row * p.ldc) + gcol;
              if (full && p.vec_c) {
                *reinterpret_cast<float4*>(cp) = make_float4(y[i][0], y[i][1], y[i][2], y[i][3]);
              } else {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                  if (gcol + u < p.N) cp[u] = y[i][u];
              }
            }
            if (p.split_out && !(p.dbg & 16)) {
              float sc_[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) sc_[u] = (gcol + u < p.N) ? y[i][u] * p.split_scale : 0.f;
              // gcol % 4 == 0: 8-byte aligned hi16 / lo16 stores, 4-byte aligned e4m3 stores
              store_split4(reinterpret_cast<uint8_t*>(p.split_out) + (long long)row * (4 * p.split_kp), p.split_kp, gcol,
                           sc_[0], sc_[1], sc_[2], sc_[3], p.split_fmt, ovf);
            }
          }
        }
        __syncwarp();
        if (has_res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) rcur[i] = rnext[i];
        }
        bcur = bnext;
        scur = snext;
      }
      if (un.part >= 0 && un.part + 1 < p.ksplit) {
        // this warp's share of the part is stored: publish it to the next part of the tile
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(p.tile_flags + (un.tile - p.whole_units), 1);
      }
    }
    report_overflow(p.overflow, ovf);
    // the issuing thread's bulk stores have READ their shared-memory tiles (their writes are complete when the grid is)
    if (p.split_tma && warp == 2) tma_store_wait_read();
  }
  tc_fence_before();
  if (TWO) cluster_sync(); else __syncthreads();   // the peer's shared memory / barriers stay alive until both are done
  if (p.sched && rank == 0 && threadIdx.x == 0) {
    // the last worker to finish re-arms the counters for the next use of this slot (every worker has drawn its -1)
    if (atomicAdd(p.sched + 1, 1) == num_workers - 1) {
      atomicExch(p.sched, 0);
      atomicExch(p.sched + 1, 0);
      if (p.ksplit > 1)      // (plain stores: a chain of atomics here cost a round trip each, ~40 us for fc2's 49 flags)
        for (int i = 0; i < num_tiles - p.whole_units; ++i) reinterpret_cast<volatile int*>(p.tile_flags)[i] = 0;
    }
  }
  if (warp == 1) {
    if (TWO) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

// X [M, K] fp32 (row stride ldx) -> X2 [M, 2*Kp] fp16 = [hi | lo] of X*scale, zero padded to Kp.
__global__ void split_f16_kernel(const float* __restrict__ X, __half* __restrict__ X2, int M, int K, int ldx, int seg,
                                 long long seg_stride, int Kp, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 2 elements
  const int half_kp = Kp >> 1;
  if (i >= (long long)M * half_kp) return;
  const int m = (int)(i / half_kp), k = (int)(i % half_kp) * 2;
  const float* x = seg > 0 ? X + (long long)(m / seg) * seg_stride + (long long)(m % seg) * ldx : X + (long long)m * ldx;
  float v0 = 0.f, v1 = 0.f;
  if (k < K) v0 = x[k] * scale;
  if (k + 1 < K) v1 = x[k + 1] * scale;
  __half2 hi, lo;
  split_pair(v0, v1, hi, lo);
  __half2* row = reinterpret_cast<__half2*>(X2 + (long long)m * 2 * Kp);
  row[k >> 1] = hi;
  row[(Kp + k) >> 1] = lo;
}

// X [M, K] fp32 -> F16F8 rows [hi16 : Kp halves | per 64 columns: hi8 x 64, lo8 x 64] of X*scale, zero padded to Kp.
// role 0 = A operand (activations), 1 = B operand (weights): the plane scales at the top of common.cuh's split section.
__global__ void split_f16f8_kernel(const float* __restrict__ X, uint8_t* __restrict__ out, int M, int K, int ldx, int seg,
                                   long long seg_stride, int Kp, float scale, int role, unsigned long long* overflow) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 4 elements
  const int quarter_kp = Kp >> 2;
  if (i >= (long long)M * quarter_kp) return;
  const int m = (int)(i / quarter_kp), k = (int)(i % quarter_kp) * 4;
  const float* x = seg > 0 ? X + (long long)(m / seg) * seg_stride + (long long)(m % seg) * ldx : X + (long long)m * ldx;
  float v[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) v[u] = (k + u < K) ? x[k + u] * scale : 0.f;
  uint8_t* row = out + (long long)m * 4 * Kp;
  if (role == 0) {
    uint32_t flags = 0;
    store_split4(row, Kp, k, v[0], v[1], v[2], v[3], EC_SPLIT_F16F8, flags);
    report_overflow(overflow, flags);
  } else {
    const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const float s = 0.00048828125f;                     // 2^-11
    if (role == 2) {
      // role 2: all three planes interleaved per 32 columns -- 128 bytes per (row, 32-column slice j) at byte 128 j:
      // [hi16 x 32 (64 B) | hi8 x 32 | lo8 x 32] -- so that ONE TMA box with 128-byte rows (whole L2 lines) holds a 32-deep
      // k-slice of the operand (gcn_fused2_tcgen05.cu; 64-byte box rows move at about half the rate)
      uint8_t* sl = row + 128 * (k >> 5);
      const int o = k & 31;
      *reinterpret_cast<uint2*>(sl + 2 * o) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
      *reinterpret_cast<uint32_t*>(sl + 64 + o) = e4m3x2(f01.x * s, f01.y * s) | (e4m3x2(f23.x * s, f23.y * s) << 16);
      *reinterpret_cast<uint32_t*>(sl + 96 + o) =
          e4m3x2(v[0] - f01.x, v[1] - f01.y) | (e4m3x2(v[2] - f23.x, v[3] - f23.y) << 16);
      return;
    }
    *reinterpret_cast<uint2*>(row + 2 * k) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    *reinterpret_cast<uint32_t*>(row + f8_off(Kp, k)) = e4m3x2(f01.x * s, f01.y * s) | (e4m3x2(f23.x * s, f23.y * s) << 16);
    *reinterpret_cast<uint32_t*>(row + f8_off(Kp, k) + 64) =
        e4m3x2(v[0] - f01.x, v[1] - f01.y) | (e4m3x2(v[2] - f23.x, v[3] - f23.y) << 16);
  }
}

// ------------------------------------------------------------------------- host side
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

struct MapKey {
  const void* ptr;
  int rows, kp, box_rows;   // box_rows < 0: the byte view of an F16F8 operand (128-byte boxes)
  bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && kp == o.kp && box_rows == o.box_rows; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ (std::hash<int>()(k.rows) * 1000003u) ^ (std::hash<int>()(k.kp) * 7919u) ^
           (std::hash<int>()(k.box_rows) * 104729u);
  }
};

// halves view (bytes == false): dims {2 Kp halves, rows}, 64-half x box_rows boxes, 128B swizzle -- the hi16 | lo16 planes.
// byte view (kind 1): dims {4 Kp bytes, rows}, 128-byte x box_rows boxes, 128B swizzle -- both e4m3 planes of a 64-column
// block of an F16F8 operand ([hi8 x 64 | lo8 x 64] at byte column 2 Kp + 128 (k / 64) of the same rows).
// slice view (kind 2): dims {4 Kp bytes, rows}, 128-byte x box_rows boxes, 128B swizzle -- one 32-deep k-slice of an operand
// whose planes are interleaved per 32 columns (ec_split_f16f8 role 2; gcn_fused2_tcgen05.cu streams its weights so).
static int get_tensor_map_any(const void* ptr, int rows, int kp, int box_rows, CUtensorMap* out, int kind) {
  const bool bytes = kind != 0;
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  std::lock_guard<std::mutex> lock(mu);
  MapKey key{ptr, rows, kp, kind == 1 ? -box_rows : (kind == 2 ? box_rows + 100000 : box_rows)};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return EC_OK;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return EC_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)(bytes ? 4 * kp : 2 * kp), (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)(4 * kp)};
  cuuint32_t box[2] = {(cuuint32_t)(kind == 0 ? BK : 2 * BK), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, bytes ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (ptr %p rows %d kp %d)", (int)r, ptr, rows, kp);
    return EC_ERR_CUDA;
  }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return EC_OK;
}

// Per-device state of ec_gemm_f16x3: SM count, the opt-in shared-memory attribute of the three kernel instances (a
// per-device function attribute) and the tile counters of the dynamic scheduler.  Created under a mutex on the first
// call made with that device current -- which therefore must not be inside a stream capture (cudaMalloc / cudaMemset).
constexpr unsigned EAGER_SLOTS = 4096, GRAPH_SLOTS = 61440;
// flag regions of the K-split tail (TcParams::tile_flags): FLAGS_PER_REGION ints each; eager launches cycle through a small
// ring, launches recorded into a CUDA graph keep theirs (when the pool runs out a launch simply does not split)
constexpr unsigned FLAGS_PER_REGION = 160, EAGER_FLAG_REGIONS = 256, GRAPH_FLAG_REGIONS = 4096;
struct DevState {
  int num_sms = 0;
  int* sched_base = nullptr;
  int* flag_base = nullptr;
  std::atomic<unsigned> eager_seq{0}, graph_seq{0}, eager_flag_seq{0}, graph_flag_seq{0};
};
std::atomic<long long> g_mode_launches[6];   // launches per tile mode: 128x128, 128x256, CTA-pair 256x256; [3..5] = the F16F8 kernels

static DevState* dev_state() {
  constexpr int MAX_DEV = 64;
  static DevState* states[MAX_DEV] = {};
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) {
    set_error("ec_gemm_f16x3: no current CUDA device");
    return nullptr;
  }
  std::lock_guard<std::mutex> lock(mu);
  if (states[dev]) return states[dev];
  DevState* d = new DevState();
  cudaError_t e = cudaDeviceGetAttribute(&d->num_sms, cudaDevAttrMultiProcessorCount, dev);
  const void* kernels[6] = {(const void*)gemm_f16x3_kernel<128, false, false>, (const void*)gemm_f16x3_kernel<256, false, false>,
                            (const void*)gemm_f16x3_kernel<256, true, false>,  (const void*)gemm_f16x3_kernel<128, false, true>,
                            (const void*)gemm_f16x3_kernel<256, false, true>,  (const void*)gemm_f16x3_kernel<256, true, true>};
  for (int i = 0; i < 6 && e == cudaSuccess; ++i)
    e = cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  const size_t bytes = (size_t)(EAGER_SLOTS + GRAPH_SLOTS) * 2 * sizeof(int);
  if (e == cudaSuccess) e = cudaMalloc(&d->sched_base, bytes);
  if (e == cudaSuccess) e = cudaMemset(d->sched_base, 0, bytes);
  const size_t fbytes = (size_t)(EAGER_FLAG_REGIONS + GRAPH_FLAG_REGIONS) * FLAGS_PER_REGION * sizeof(int);
  if (e == cudaSuccess) e = cudaMalloc(&d->flag_base, fbytes);
  if (e == cudaSuccess) e = cudaMemset(d->flag_base, 0, fbytes);
  if (e != cudaSuccess) {
    set_error("ec_gemm_f16x3: per-device setup failed on device %d: %s (the first call on a device allocates the tile "
              "counters and must not be inside a stream capture)", dev, cudaGetErrorString(e));
    cudaGetLastError();
    delete d;
    return nullptr;
  }
  states[dev] = d;
  return d;
}

int get_tensor_map(const void* ptr, int rows, int kp, int box_rows, CUtensorMap* out) {   // also used by other kernels
  return get_tensor_map_any(ptr, rows, kp, box_rows, out, 0);
}
int get_tensor_map_slice32(const void* ptr, int rows, int kp, int box_rows, CUtensorMap* out) {   // 128-byte slices, planes interleaved per 32
  return get_tensor_map_any(ptr, rows, kp, box_rows, out, 2);
}
// plain fp32 matrix [rows, cols] (contiguous rows), box_rows x box_cols boxes; no swizzle (raw activations that a kernel
// converts itself) or, with box_cols = 32, 128B swizzle (result tiles a kernel assembles for TMA stores)
int get_tensor_map_f32(const void* ptr, long long rows, int cols, int box_rows, int box_cols, bool swizzle128, CUtensorMap* out) {
  struct Key {
    const void* ptr; long long rows; int cols, br, bc;
    bool operator==(const Key& o) const { return ptr == o.ptr && rows == o.rows && cols == o.cols && br == o.br && bc == o.bc; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      return std::hash<const void*>()(k.ptr) ^ (std::hash<long long>()(k.rows) * 1000003u) ^ (std::hash<int>()(k.cols) * 7919u) ^
             (std::hash<int>()(k.br) * 104729u) ^ (std::hash<int>()(k.bc) * 15485863u);
    }
  };
  static std::mutex mu;
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  std::lock_guard<std::mutex> lock(mu);
  Key key{ptr, rows, cols, box_rows, swizzle128 ? -box_cols : box_cols};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return EC_OK;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return EC_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4u};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d (ptr %p rows %lld cols %d box %d x %d)", (int)r, ptr, rows,
              cols, box_rows, box_cols);
    return EC_ERR_CUDA;
  }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return EC_OK;
}
int get_tensor_map_bytes(const void* ptr, int rows, int kp, int box_rows, CUtensorMap* out) {   // e4m3 planes of an F16F8 operand
  return get_tensor_map_any(ptr, rows, kp, box_rows, out, 1);
}

}  // namespace tc
}  // namespace ec

using namespace ec;

static int ec_tc_debug = 0;
extern "C" int ec_tc_set_debug(int flags) {   // bring-up / profiling experiments only (results are wrong when != 0)
  ec_tc_debug = flags;
  return EC_OK;
}
static int ec_tc_dynamic = -1;   // dynamic tile scheduling (default on; EDGECAPE_GEMM_DYNAMIC=0 / ec_tc_set_dynamic(0) = static)
extern "C" int ec_tc_set_dynamic(int on) {
  ec_tc_dynamic = on ? 1 : 0;
  return EC_OK;
}
// K-split of the tail tiles (TcParams::ksplit): EDGECAPE_GEMM_KSPLIT=0 / ec_tc_set_ksplit(0) switches it off
static int ec_tc_ksplit = -1;
static std::atomic<long long> g_ksplit_launches{0};
extern "C" int ec_tc_set_ksplit(int on) {
  ec_tc_ksplit = on ? 1 : 0;
  return EC_OK;
}
extern "C" long long ec_tc_ksplit_launches() { return g_ksplit_launches.load(std::memory_order_relaxed); }
static int ec_tc_cta_limit = 0;  // 0 = every SM; else the persistent grids use at most this many CTAs
extern "C" int ec_tc_set_cta_limit(int ctas) {
  EC_REQUIRE(ctas >= 0, "ec_tc_set_cta_limit: negative");
  ec_tc_cta_limit = ctas & ~1;
  return EC_OK;
}
static int ec_tc_split_tma = 1;  // split-only epilogues through TMA stores (0: scattered stores, for A/B measurements)
extern "C" int ec_tc_set_split_tma(int on) {
  EC_REQUIRE(on >= 0 && on <= 2, "ec_tc_set_split_tma: 0, 1 or 2");
  ec_tc_split_tma = on;
  return EC_OK;
}
static long long* ec_tc_trace = nullptr;   // profiling: [workers][4] int64, see TcParams::trace
extern "C" int ec_tc_set_trace(void* buf) {
  ec_tc_trace = (long long*)buf;
  return EC_OK;
}
static int ec_tc_force_bn = 0;   // 0 = heuristic; 128 / 256 force a tile width (tuning / tests)
extern "C" int ec_tc_set_tile_n(int bn) {
  EC_REQUIRE(bn == 0 || bn == 128 || bn == 256 || bn == 512, "ec_tc_set_tile_n: 0, 128, 256 or 512 (CTA pair)");
  ec_tc_force_bn = bn;
  return EC_OK;
}

extern "C" long long ec_tc_mode_launches(int mode) {   // mode + 1: the same tile mode on the F16F8 kernels
  const int f8 = mode & 1, m = mode & ~1;
  if (m != 128 && m != 256 && m != 512) return -1;
  return tc::g_mode_launches[3 * f8 + (m == 512 ? 2 : (m == 256 ? 1 : 0))].load(std::memory_order_relaxed);
}

extern "C" int ec_split_f16(const float* X, void* X2, int M, int K, int ldx, int seg, long long seg_stride, int Kp,
                            float scale, void* stream) {
  EC_REQUIRE(X && X2, "ec_split_f16: null pointer");
  EC_REQUIRE(Kp % tc::BK == 0 && Kp >= K && K > 0, "ec_split_f16: Kp must be a multiple of 64 and >= K");
  if (M == 0) return EC_OK;
  long long total = (long long)M * (Kp / 2);
  launch_pdl(tc::split_f16_kernel, dim3(cdiv(total, 256)), dim3(256), 0, (cudaStream_t)stream, X, (__half*)X2, M, K, ldx, seg, seg_stride, Kp,
                                                                              scale);
  return check_launch("ec_split_f16");
}

extern "C" int ec_split_f16f8(const float* X, void* out, int M, int K, int ldx, int seg, long long seg_stride, int Kp,
                              float scale, int role, void* stream) {
  EC_REQUIRE(X && out, "ec_split_f16f8: null pointer");
  EC_REQUIRE(Kp % tc::BK == 0 && Kp >= K && K > 0, "ec_split_f16f8: Kp must be a multiple of 64 and >= K");
  EC_REQUIRE(role >= 0 && role <= 2, "ec_split_f16f8: role is 0 (A operand), 1 (B operand) or 2 (B operand, planes interleaved per 32 columns)");
  EC_REQUIRE(aligned16(out), "ec_split_f16f8: the operand buffer must be 16-byte aligned");
  if (M == 0) return EC_OK;
  unsigned long long* ovf = overflow_counters();
  if (!ovf) return EC_ERR_CUDA;
  const long long total = (long long)M * (Kp / 4);
  launch_pdl(tc::split_f16f8_kernel, dim3(cdiv(total, 256)), dim3(256), 0, (cudaStream_t)stream, X, (uint8_t*)out, M, K,
             ldx, seg, seg_stride, Kp, scale, role, ovf);
  return check_launch("ec_split_f16f8");
}

static int gemm_split_launch(const char* what, bool f8, const void* A2, const void* B2, float* C, int M, int N, int Kp,
                             int ldc, int seg_c, long long seg_stride_c, float out_scale, const float* bias, int act,
                             const float* colscale, const float* R, int ldr, int res_mode, int res_rows,
                             void* split_out, int split_kp, float split_scale, int split_fmt, void* stream) {
  EC_REQUIRE(A2 && B2 && (C || split_out), "%s: null operand", what);
  EC_REQUIRE(Kp > 0 && Kp % tc::BK == 0, "%s: Kp must be a positive multiple of 64", what);
  EC_REQUIRE(aligned16(A2) && aligned16(B2), "%s: split operands must be 16-byte aligned", what);
  EC_REQUIRE(!split_out || (((uintptr_t)split_out & 7) == 0), "%s: split_out must be 8-byte aligned", what);
  EC_REQUIRE((res_mode == EC_RES_NONE) == (R == nullptr), "%s: residual pointer/mode mismatch", what);
  EC_REQUIRE(!split_out || (split_kp % tc::BK == 0 && split_kp >= N), "%s: bad split_kp", what);
  EC_REQUIRE(split_fmt == EC_SPLIT_F16X2 || split_fmt == EC_SPLIT_F16F8, "%s: split_fmt is EC_SPLIT_F16X2 or EC_SPLIT_F16F8", what);
  if (M == 0 || N == 0) return EC_OK;
  tc::DevState* ds = tc::dev_state();
  if (!ds) return EC_ERR_CUDA;
  const int num_sms = ds->num_sms;
  // tile selection: CTA-pair 256x256 tiles (half the operand traffic per unit of math) for problems that give every
  // SM pair >= 1.5 tiles; 128x128 single-CTA tiles otherwise.  (A waves x tile-cost model that also moved mid-size
  // problems to pair tiles measured 1.5% slower end to end -- pass L -- the cluster launch has a higher fixed cost.)
  const long long pair_tiles = (long long)cdiv(M, 256) * cdiv(N, 256);
  int mode = ec_tc_force_bn;                     // 0 heuristic, 128, 256, 512 (= pair)
  // (with ~1.7 waves of pair tiles and a short K loop -- the ViT proj GEMM -- the finer 128x128 tiles measure 10 % faster)
  if (mode == 0)
    mode = (N >= 256 && pair_tiles * 4 >= 3LL * num_sms && (pair_tiles >= num_sms || Kp >= 1024)) ? 512 : 128;
  const int BN = mode == 128 ? 128 : 256;
  CUtensorMap tmA, tmB, tmA8, tmB8;
  int rc = tc::get_tensor_map(A2, M, Kp, tc::BM, &tmA);
  if (rc) return rc;
  rc = tc::get_tensor_map(B2, N, Kp, mode == 256 ? 256 : 128, &tmB);
  if (rc) return rc;
  tmA8 = tmA;
  tmB8 = tmB;
  if (f8) {
    rc = tc::get_tensor_map_any(A2, M, Kp, tc::BM, &tmA8, 1);
    if (rc) return rc;
    rc = tc::get_tensor_map_any(B2, N, Kp, mode == 256 ? 256 : 128, &tmB8, 1);
    if (rc) return rc;
  }
  // split-only outputs leave through TMA stores (128-row x 64-column blocks in the layout of the consumer's A tiles)
  CUtensorMap tmO16 = tmA, tmO8 = tmA;
  const bool split_tma = split_out && !C && aligned16(split_out) && ec_tc_split_tma;
  if (split_tma) {
    rc = tc::get_tensor_map_any(split_out, M, split_kp, tc::BM, &tmO16, 0);
    if (rc) return rc;
    if (split_fmt == EC_SPLIT_F16F8) {
      rc = tc::get_tensor_map_any(split_out, M, split_kp, tc::BM, &tmO8, 1);
      if (rc) return rc;
    }
  }
  tc::TcParams p{};
  p.split_tma = split_tma ? ec_tc_split_tma : 0;
  p.split_fmt = split_fmt;
  if (split_out && split_fmt == EC_SPLIT_F16F8) {
    p.overflow = overflow_counters();
    if (!p.overflow) return EC_ERR_CUDA;
  }
  // consecutive tiles walk N first when there are few N tiles: the CTAs working at the same time then share their A
  // row block (read from DRAM once, L2 hits for the other N tiles) and the whole of B stays in L2
  p.n_fastest = cdiv(N, BN) <= 16 ? 1 : 0;
  p.C = C; p.M = M; p.N = N; p.num_kb = Kp / tc::BK; p.Kp = Kp; p.ldc = ldc;
  p.seg_c = seg_c; p.seg_stride_c = seg_stride_c;
  p.vec_c = C && aligned16(C) && (ldc % 4 == 0) && (seg_stride_c % 4 == 0);
  p.vec_r = R && aligned16(R) && (ldr % 4 == 0);
  p.out_scale = out_scale; p.bias = bias; p.colscale = colscale; p.R = R; p.ldr = ldr; p.act = act;
  p.res_mode = res_mode; p.res_rows = res_rows; p.split_out = (__half*)split_out; p.split_kp = split_kp; p.split_scale = split_scale;
  p.dbg = ec_tc_debug;
  p.trace = ec_tc_trace;
  // one {next tile, finished workers} pair per launch, re-armed by the launch's last worker.  Eager launches cycle
  // through a ring of slots (a slot is reused 4096 launches later); launches recorded into a CUDA graph keep their slot
  // for the life of the process (the node replays with it, possibly concurrently with eager work on another stream), so
  // they come from a separate pool that never wraps -- when it runs out, further captured launches use static tiles.
  if (ec_tc_dynamic < 0) {
    const char* e = getenv("EDGECAPE_GEMM_DYNAMIC");
    ec_tc_dynamic = (e && e[0] == '0') ? 0 : 1;
  }
  p.sched = nullptr;
  // dynamic tiles only pay when a CTA can take more than one: with one tile per CTA (the head's many small GEMMs) the
  // two global atomics -- the tile and the terminator, a round trip each before the first TMA load can be issued --
  // are pure latency, and the static assignment gives the same result
  const long long single_tiles = (long long)cdiv(M, tc::BM) * cdiv(N, BN);
  const int sched_ctas = (ec_tc_cta_limit > 0 && ec_tc_cta_limit < num_sms) ? ec_tc_cta_limit : num_sms;
  const bool one_wave = mode == 512 ? pair_tiles <= sched_ctas / 2 : single_tiles <= sched_ctas;
  if (ec_tc_dynamic && !one_wave) {
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    EC_CUDA(cudaStreamIsCapturing((cudaStream_t)stream, &cap));
    if (cap == cudaStreamCaptureStatusNone) {
      p.sched = ds->sched_base + 2 * (ds->eager_seq.fetch_add(1) % tc::EAGER_SLOTS);
    } else {
      const unsigned g = ds->graph_seq.fetch_add(1);
      if (g < tc::GRAPH_SLOTS) p.sched = ds->sched_base + 2 * (tc::EAGER_SLOTS + g);
    }
  }
  // K-split of the tail: when the last wave is partly empty, cut its tiles into three K-parts so that they fill it (fc2:
  // 123 pair tiles on 74 pairs = 2 tile times -> 74 whole tiles + 49 x 3 parts = 1.67; proj: 492 tiles on 148 CTAs = 4 -> 3.33)
  p.ksplit = 1; p.whole_units = 0; p.tail_tiles = 0; p.tile_flags = nullptr;
  if (ec_tc_ksplit < 0) {
    const char* e = getenv("EDGECAPE_GEMM_KSPLIT");
    ec_tc_ksplit = (e && e[0] == '1') ? 1 : 0;
  }
  if (ec_tc_ksplit && p.sched && C && !split_out && act == EC_ACT_NONE && res_mode != EC_RES_GATE && p.num_kb % 3 == 0 &&
      p.num_kb >= 12) {
    const long long tiles = mode == 512 ? pair_tiles : (long long)cdiv(M, tc::BM) * cdiv(N, BN);
    const long long workers = mode == 512 ? sched_ctas / 2 : sched_ctas;
    const long long whole = tiles / workers * workers, tail = tiles - whole;
    if (whole > 0 && tail > 0 && 3 * tail <= 2 * workers && tail <= (long long)tc::FLAGS_PER_REGION) {
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      EC_CUDA(cudaStreamIsCapturing((cudaStream_t)stream, &cap));
      unsigned region = ~0u;
      if (cap == cudaStreamCaptureStatusNone) {
        region = ds->eager_flag_seq.fetch_add(1) % tc::EAGER_FLAG_REGIONS;
      } else {
        const unsigned g = ds->graph_flag_seq.fetch_add(1);
        if (g < tc::GRAPH_FLAG_REGIONS) region = tc::EAGER_FLAG_REGIONS + g;
      }
      if (region != ~0u) {
        p.ksplit = 3;
        g_ksplit_launches.fetch_add(1, std::memory_order_relaxed);
        p.whole_units = (int)whole;
        p.tail_tiles = (int)tail;
        p.tile_flags = ds->flag_base + (size_t)region * tc::FLAGS_PER_REGION;
      }
    }
  }
  tc::g_mode_launches[(f8 ? 3 : 0) + (mode == 512 ? 2 : (mode == 256 ? 1 : 0))].fetch_add(1, std::memory_order_relaxed);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 512) {
    const int max_ctas = (ec_tc_cta_limit > 0 && ec_tc_cta_limit < num_sms) ? ec_tc_cta_limit : num_sms;
    const int pairs = (int)(pair_tiles < max_ctas / 2 ? pair_tiles : max_ctas / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(tc::THREADS);
    cfg.dynamicSmemBytes = tc::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (f8) EC_CUDA(cudaLaunchKernelEx(&cfg, tc::gemm_f16x3_kernel<256, true, true>, tmA, tmB, tmA8, tmB8, tmO16, tmO8, p));
    else EC_CUDA(cudaLaunchKernelEx(&cfg, tc::gemm_f16x3_kernel<256, true, false>, tmA, tmB, tmA8, tmB8, tmO16, tmO8, p));
  } else {
    const int tiles = cdiv(M, tc::BM) * cdiv(N, BN);
    const int max_ctas = (ec_tc_cta_limit > 0 && ec_tc_cta_limit < num_sms) ? ec_tc_cta_limit : num_sms;
    const int grid = tiles < max_ctas ? tiles : max_ctas;
    if (BN == 256 && f8)
      launch_pdl(tc::gemm_f16x3_kernel<256, false, true>, dim3(grid), dim3(tc::THREADS), tc::SMEM_BYTES, st, tmA, tmB, tmA8, tmB8, tmO16, tmO8, p);
    else if (BN == 256)
      launch_pdl(tc::gemm_f16x3_kernel<256, false, false>, dim3(grid), dim3(tc::THREADS), tc::SMEM_BYTES, st, tmA, tmB, tmA8, tmB8, tmO16, tmO8, p);
    else if (f8)
      launch_pdl(tc::gemm_f16x3_kernel<128, false, true>, dim3(grid), dim3(tc::THREADS), tc::SMEM_BYTES, st, tmA, tmB, tmA8, tmB8, tmO16, tmO8, p);
    else
      launch_pdl(tc::gemm_f16x3_kernel<128, false, false>, dim3(grid), dim3(tc::THREADS), tc::SMEM_BYTES, st, tmA, tmB, tmA8, tmB8, tmO16, tmO8, p);
  }
  return check_launch(what);
}

extern "C" int ec_gemm_f16x3(const void* A2, const void* B2, float* C, int M, int N, int Kp, int ldc, int seg_c,
                             long long seg_stride_c, float out_scale, const float* bias, int act, const float* colscale,
                             const float* R, int ldr, int res_mode, int res_rows, void* split_out, int split_kp,
                             float split_scale, void* stream) {
  return gemm_split_launch("ec_gemm_f16x3", false, A2, B2, C, M, N, Kp, ldc, seg_c, seg_stride_c, out_scale, bias, act,
                           colscale, R, ldr, res_mode, res_rows, split_out, split_kp, split_scale, EC_SPLIT_F16X2, stream);
}

extern "C" int ec_gemm_f16f8(const void* A3, const void* B3, float* C, int M, int N, int Kp, int ldc, int seg_c,
                             long long seg_stride_c, float out_scale, const float* bias, int act, const float* colscale,
                             const float* R, int ldr, int res_mode, int res_rows, void* split_out, int split_kp,
                             float split_scale, int split_fmt, void* stream) {
  return gemm_split_launch("ec_gemm_f16f8", true, A3, B3, C, M, N, Kp, ldc, seg_c, seg_stride_c, out_scale, bias, act,
                           colscale, R, ldr, res_mode, res_rows, split_out, split_kp, split_scale, split_fmt, stream);
}
