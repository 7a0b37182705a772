// Skeleton / graph kernels: adjacency from edge lists, soft normalisation, edge-weight
// prediction with Markov hop matrices, and the GCN feed-forward (aggregate + packed GEMM).
#include <cuda_fp16.h>

#include "common.cuh"

namespace ec {

// One CTA per batch element.  `binary` (global, [K,K]) doubles as the scatter target.
__global__ void __launch_bounds__(1024) adj_from_edges_kernel(const int32_t* __restrict__ edges,
                                                             const int32_t* __restrict__ offsets,
                                                             const uint8_t* __restrict__ kp_mask,
                                                             float* __restrict__ adj, float* __restrict__ binary,
                                                             int K) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  float* bin = binary + (long long)b * K * K;
  const uint8_t* mk = kp_mask + (long long)b * K;
  const int KK = K * K;
  for (int i = threadIdx.x; i < KK; i += blockDim.x) bin[i] = 0.f;
  __syncthreads();
  const int e0 = offsets[b], e1 = offsets[b + 1];
  for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    int i = edges[2 * e], j = edges[2 * e + 1];
    if (i < 0) i += K;   // torch advanced indexing wraps negative indices
    if (j < 0) j += K;
    if (i < 0 || j < 0 || i >= K || j >= K) continue;
    if (mk[i] || mk[j]) continue;          // masked rows / columns are zeroed afterwards anyway
    bin[i * K + j] = 1.f;
    bin[j * K + i] = 1.f;
  }
  __syncthreads();
  float* a0 = adj + (long long)b * 2 * KK;
  float* a1 = a0 + KK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < K; i += nw) {
    float s = 0.f;
    for (int j = lane; j < K; j += 32) s += bin[i * K + j];
    s = warp_sum(s);
    const float valid = mk[i] ? 0.f : 1.f;
    for (int j = lane; j < K; j += 32) {
      // nan_to_num(0/0) = 0; a non-empty row has s >= 1
      a1[i * K + j] = s > 0.f ? bin[i * K + j] / s : 0.f;
      a0[i * K + j] = (i == j) ? valid : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) soft_normalize_kernel(const float* __restrict__ U,
                                                             const uint8_t* __restrict__ kp_mask,
                                                             float* __restrict__ adj, int K) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int KK = K * K;
  const float* u = U + (long long)b * KK;
  const uint8_t* mk = kp_mask + (long long)b * K;
  float* a0 = adj + (long long)b * 2 * KK;
  float* a1 = a0 + KK;
  // rows are independent: gridDim.y CTAs share the rows of a sample, one warp per row at a time
  const int lane = threadIdx.x & 31, nw = (blockDim.x >> 5) * gridDim.y;
  const int warp = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int i = warp; i < K; i += nw) {
    const float vi = mk[i] ? 0.f : 1.f;
    float s = 0.f;
    for (int j = lane; j < K; j += 32) s += u[i * K + j] * (vi * (mk[j] ? 0.f : 1.f));
    s = warp_sum(s);
    for (int j = lane; j < K; j += 32) {
      a1[i * K + j] = u[i * K + j] * (vi * (mk[j] ? 0.f : 1.f)) / (s + 1e-8f);
      a0[i * K + j] = (i == j) ? vi : 0.f;
    }
  }
}

__global__ void __launch_bounds__(256) edge_weights_kernel(const float* __restrict__ S,
                                                           const float* __restrict__ binary,
                                                           const uint8_t* __restrict__ kp_mask, float zc_w,
                                                           float zc_b, int use_zc, float* __restrict__ adj,
                                                           float* __restrict__ unnorm, float* __restrict__ hop0,
                                                           float* __restrict__ hop1, int K) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x;
  const int KK = K * K;
  const float* s = S + (long long)b * KK;
  const float* bin = binary + (long long)b * KK;
  const uint8_t* mk = kp_mask + (long long)b * K;
  float* a0 = adj + (long long)b * 2 * KK;
  float* a1 = a0 + KK;
  float* un = unnorm ? unnorm + (long long)b * KK : nullptr;
  float* h0 = hop0 ? hop0 + (long long)b * KK : nullptr;
  float* h1 = hop1 ? hop1 + (long long)b * KK : nullptr;
  const int lane = threadIdx.x & 31, nw = (blockDim.x >> 5) * gridDim.y;
  const int warp = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int i = warp; i < K; i += nw) {
    const float vi = mk[i] ? 0.f : 1.f;
    auto value = [&](int j) {
      float v = (s[i * K + j] + s[j * K + i]) / 2.0f;
      if (use_zc) v = v * zc_w + zc_b;
      v = fmaxf(bin[i * K + j] + v, 0.f);
      return v * (vi * (mk[j] ? 0.f : 1.f));
    };
    float sum1 = 0.f;
    for (int j = lane; j < K; j += 32) sum1 += value(j);
    sum1 = warp_sum(sum1);
    float sum2 = 0.f;
    for (int j = lane; j < K; j += 32) sum2 += value(j) / (sum1 + 1e-8f);
    sum2 = warp_sum(sum2);
    for (int j = lane; j < K; j += 32) {
      const float um = value(j);
      const float a = um / (sum1 + 1e-8f);
      a1[i * K + j] = a;
      a0[i * K + j] = (i == j) ? vi : 0.f;
      if (un) un[i * K + j] = um;
      if (h0) h0[i * K + j] = (i == j) ? 1.f : 0.f;
      if (h1) h1[i * K + j] = a / (sum2 + 1e-8f);
    }
  }
}

// Z[b,w,:] = [ a0[w] * X[b,w,:] | (A1 X)[b,w,:] (written by the batched GEMM) | a0[w] | rowsum(A1[w,:]) | 0 0 ]
__global__ void __launch_bounds__(256) gcn_fill_kernel(const float* __restrict__ X,
                                                       const float* __restrict__ adj, float* __restrict__ Z,
                                                       int K, int d, int ldz) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  if (warp >= K) return;
  const int w = warp;
  const int KK = K * K;
  const float* a0p = adj + (long long)b * 2 * KK;
  const float* a1p = a0p + KK;
  const float a0 = a0p[w * K + w];
  float rs = 0.f;
  for (int v = lane; v < K; v += 32) rs += a1p[w * K + v];
  rs = warp_sum(rs);
  const float* x = X + ((long long)b * K + w) * d;
  float* z = Z + ((long long)b * K + w) * ldz;
  for (int c = lane; c < d; c += 32) z[c] = a0 * x[c];
  if (lane == 0) {
    z[2 * d + 0] = a0;
    z[2 * d + 1] = rs;
    z[2 * d + 2] = 0.f;
    z[2 * d + 3] = 0.f;
  }
}

// Fused aggregate for the tensor-core GCN: one pass produces the GEMM A operand
//   Z[b,w,:] = [ a0[w] X[b,w,:] | sum_v A1[b,w,v] X[b,v,:] | a0[w] | rowsum(A1[b,w,:]) | 0 ... ]
// directly in split-fp16 form Z2 [B*K, 2*Kp] (hi | lo).  CTA = (8 rows of one sample); the 8 x K block of
// A1 sits in shared memory and is read as float4 broadcasts, X rows stream through coalesced loads.
constexpr int GA_ROWS = 8;     // rows of one sample per CTA: 13 x B CTAs at K = 100 (several co-resident per SM)
__global__ void __launch_bounds__(256, 3) gcn_aggregate_split_kernel(const float* __restrict__ X,
                                                                  const float* __restrict__ adj,
                                                                  __half* __restrict__ Z2, int K, int d, int Kp) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float a1s[];                 // [GA_ROWS][KP4] (KP4 = K rounded up to 4, zero padded)
  __shared__ float rs[GA_ROWS], a0s[GA_ROWS];
  const int b = blockIdx.y, w0 = blockIdx.x * GA_ROWS;
  const int KP4 = (K + 3) & ~3;
  const long long KK = (long long)K * K;
  const float* a0p = adj + (long long)b * 2 * KK;
  const float* a1p = a0p + KK;
  for (int i = threadIdx.x; i < GA_ROWS * KP4; i += blockDim.x) {
    const int r = i / KP4, v = i % KP4;
    a1s[i] = (w0 + r < K && v < K) ? a1p[(long long)(w0 + r) * K + v] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < GA_ROWS; r += 8) {
    float sacc = 0.f;
    for (int v = lane; v < KP4; v += 32) sacc += a1s[r * KP4 + v];
    sacc = warp_sum(sacc);
    if (lane == 0) {
      rs[r] = sacc;
      a0s[r] = (w0 + r < K) ? a0p[(long long)(w0 + r) * K + (w0 + r)] : 0.f;
    }
  }
  __syncthreads();
  const float* Xb = X + (long long)b * K * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float acc[GA_ROWS];
#pragma unroll
    for (int r = 0; r < GA_ROWS; ++r) acc[r] = 0.f;
    // 16 X rows in flight per thread before the FMAs (the loop is latency bound otherwise)
    for (int v0 = 0; v0 < KP4; v0 += 16) {
      float x[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) x[u] = (v0 + u < K) ? __ldg(Xb + (long long)(v0 + u) * d + c) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int v = v0 + 4 * q;
        if (v >= KP4) break;
#pragma unroll
        for (int r = 0; r < GA_ROWS; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(&a1s[r * KP4 + v]);
          acc[r] = fmaf(a.x, x[4 * q + 0], acc[r]);
          acc[r] = fmaf(a.y, x[4 * q + 1], acc[r]);
          acc[r] = fmaf(a.z, x[4 * q + 2], acc[r]);
          acc[r] = fmaf(a.w, x[4 * q + 3], acc[r]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < GA_ROWS; ++r) {
      const int w = w0 + r;
      if (w >= K) continue;
      __half* z = Z2 + ((long long)b * K + w) * 2 * Kp;
      const float y0 = a0s[r] * __ldg(Xb + (long long)w * d + c), y1 = acc[r];
      __half2 hi, lo;
      split_pair(y0, y1, hi, lo);
      z[c] = __low2half(hi);
      z[Kp + c] = __low2half(lo);
      z[d + c] = __high2half(hi);
      z[Kp + d + c] = __high2half(lo);
    }
  }
  // tail columns: a0, rowsum, zero padding
  for (int i = threadIdx.x; i < GA_ROWS * (Kp - 2 * d); i += blockDim.x) {
    const int r = i / (Kp - 2 * d), c = 2 * d + i % (Kp - 2 * d);
    const int w = w0 + r;
    if (w >= K) continue;
    const float y = c == 2 * d ? a0s[r] : (c == 2 * d + 1 ? rs[r] : 0.f);
    __half* z = Z2 + ((long long)b * K + w) * 2 * Kp;
    const __half hi = __float2half_rn(y);
    z[c] = hi;
    z[Kp + c] = __float2half_rn(y - __half2float(hi));
  }
}

// Wp[c, :] = [ W[c, 0:d] | W[dff + c, 0:d] | bias[c] | bias[dff + c] | 0 0 ]
__global__ void gcn_pack_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                float* __restrict__ Wp, int d, int dff) {
  const int ld = 2 * d + 4;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)dff * ld) return;
  int c = (int)(i / ld), k = (int)(i % ld);
  float v = 0.f;
  if (k < d) v = W[(long long)c * d + k];
  else if (k < 2 * d) v = W[(long long)(dff + c) * d + (k - d)];
  else if (k == 2 * d) v = bias[c];
  else if (k == 2 * d + 1) v = bias[dff + c];
  Wp[i] = v;
}

}  // namespace ec

using namespace ec;

extern "C" int ec_adj_from_edges(const int32_t* edges, const int32_t* offsets, const uint8_t* kp_mask,
                                 float* adj, float* binary, int B, int K, void* stream) {
  EC_REQUIRE(offsets && kp_mask && adj && binary, "ec_adj_from_edges: null pointer");
  if (B == 0) return EC_OK;
  // one CTA per sample (the scatter needs the whole matrix), 32 warps: the row passes are latency chains
  launch_pdl(adj_from_edges_kernel, dim3(B), dim3(1024), (size_t)(0), (cudaStream_t)stream, edges, offsets, kp_mask, adj, binary, K);
  return check_launch("ec_adj_from_edges");
}

extern "C" int ec_soft_normalize_adj(const float* U, const uint8_t* kp_mask, float* adj, int B, int K,
                                     void* stream) {
  EC_REQUIRE(U && kp_mask && adj, "ec_soft_normalize_adj: null pointer");
  if (B == 0) return EC_OK;
  launch_pdl(soft_normalize_kernel, dim3(B, cdiv(K, 8)), dim3(256), (size_t)(0), (cudaStream_t)stream, U, kp_mask, adj, K);
  return check_launch("ec_soft_normalize_adj");
}

extern "C" int ec_edge_weights(const float* S, const float* binary, const uint8_t* kp_mask, float zc_w,
                               float zc_b, int use_zero_conv, float* adj, float* unnorm, float* hop0,
                               float* hop1, int B, int K, void* stream) {
  EC_REQUIRE(S && binary && kp_mask && adj, "ec_edge_weights: null pointer");
  if (B == 0) return EC_OK;
  launch_pdl(edge_weights_kernel, dim3(B, cdiv(K, 8)), dim3(256), (size_t)(0), (cudaStream_t)stream, S, binary, kp_mask, zc_w, zc_b, use_zero_conv, adj,
                                                           unnorm, hop0, hop1, K);
  return check_launch("ec_edge_weights");
}

// Markov hop matrices P^2 .. P^H (skeleton.py:152-161, torch.matrix_power).  P^h = P^(h-1) . P^1, and row i of P^h only
// needs row i of P^(h-1): a warp owns 4 rows of one sample and walks them through every power on its own -- no
// synchronisation between steps.  CTA = (sample, block of 32 rows; 8 warps): P^1 in shared memory as the right operand
// (rows padded by one float), the warp's 4 current rows transposed in a private buffer so that one float4 broadcasts
// them; 4 x NC accumulators per lane (columns lane + 32 c; NC = ceil(K / 32) is a template parameter: no predicated
// or branching inner loop -- with two warps per scheduler every exposed latency counts).
constexpr int MK_ROWS = 4, MK_WARPS = 8;
template <int NC>
__global__ void __launch_bounds__(32 * MK_WARPS) markov_powers_kernel(float* __restrict__ hops, int H, int B, int K) {
  extern __shared__ float mk_smem[];
  pdl_launch_dependents();
  pdl_wait();
  const int ldp = K + 1;
  float* P1 = mk_smem;                                         // [K][K+1]
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* rowsT = P1 + (((size_t)K * ldp + 3) & ~(size_t)3) + (size_t)warp * K * MK_ROWS;   // [K][4] (16-byte aligned)
  const size_t plane = (size_t)B * K * K;
  const float* p1g = hops + plane + (size_t)b * K * K;
  // 8 independent loads in flight per thread (a one-load-per-iteration loop is a chain of L2 round trips)
  for (int base0 = 0; base0 < K * K; base0 += 8 * (int)blockDim.x) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base0 + u * (int)blockDim.x + (int)threadIdx.x;
      v[u] = i < K * K ? p1g[i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = base0 + u * (int)blockDim.x + (int)threadIdx.x;
      if (i < K * K) P1[(i / K) * ldp + (i % K)] = v[u];
    }
  }
  __syncthreads();
  const int row0 = (blockIdx.y * MK_WARPS + warp) * MK_ROWS;
  if (row0 >= K) return;
  for (int v = lane; v < K; v += 32)
#pragma unroll
    for (int r = 0; r < MK_ROWS; ++r) rowsT[v * MK_ROWS + r] = (row0 + r < K) ? P1[(row0 + r) * ldp + v] : 0.f;
  __syncwarp();
  // columns >= K of the last group read the padding / the next row of P1 (finite values) and are never stored
  const float* pcol = P1 + lane;
  for (int h = 2; h <= H; ++h) {
    float acc[MK_ROWS][NC];
#pragma unroll
    for (int r = 0; r < MK_ROWS; ++r)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[r][c] = 0.f;
#pragma unroll 4
    for (int v = 0; v < K; ++v) {
      const float4 l = *reinterpret_cast<const float4*>(rowsT + v * MK_ROWS);
      const float* pr = pcol + v * ldp;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float pv = pr[32 * c];
        acc[0][c] = fmaf(l.x, pv, acc[0][c]);
        acc[1][c] = fmaf(l.y, pv, acc[1][c]);
        acc[2][c] = fmaf(l.z, pv, acc[2][c]);
        acc[3][c] = fmaf(l.w, pv, acc[3][c]);
      }
    }
    __syncwarp();                                              // every lane is done reading the rows of P^(h-1)
    float* out = hops + (size_t)h * plane + (size_t)b * K * K;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int j = lane + 32 * c;
      if (j < K) {
#pragma unroll
        for (int r = 0; r < MK_ROWS; ++r)
          if (row0 + r < K) out[(size_t)(row0 + r) * K + j] = acc[r][c];
        *reinterpret_cast<float4*>(rowsT + j * MK_ROWS) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
      }
    }
    __syncwarp();
  }
}

extern "C" int ec_markov_powers(float* hops, int max_hop, int B, int K, void* stream) {
  EC_REQUIRE(hops, "ec_markov_powers: null pointer");
  EC_REQUIRE(K > 0 && K <= 256, "ec_markov_powers: K must be in [1, 256]");
  if (B == 0 || max_hop < 2) return EC_OK;
  // P1 (+ one spare row: the last column group of the last rows reads up to 31 floats past the matrix) + row buffers
  const size_t smem = ((size_t)(K + 1) * (K + 1) + 4 + (size_t)MK_WARPS * MK_ROWS * K) * sizeof(float);
  EC_REQUIRE(smem <= 227 * 1024, "ec_markov_powers: K too large for shared memory");
  const dim3 grid(B, cdiv(K, MK_WARPS * MK_ROWS)), block(32 * MK_WARPS);
  cudaStream_t st = (cudaStream_t)stream;
#define EC_MK(NC)                                                                          \
  case NC:                                                                                 \
    EC_CUDA((cudaError_t)ensure_dynamic_smem(markov_powers_kernel<NC>, (int)smem));        \
    launch_pdl(markov_powers_kernel<NC>, grid, block, smem, st, hops, max_hop, B, K);      \
    break;
  switch (cdiv(K, 32)) {
    EC_MK(1) EC_MK(2) EC_MK(3) EC_MK(4) EC_MK(5) EC_MK(6) EC_MK(7) EC_MK(8)
    default: break;
  }
#undef EC_MK
  return check_launch("ec_markov_powers");
}

extern "C" size_t ec_workspace_bytes_gcn(int B, int K, int d, int dff) {
  (void)dff;
  return (size_t)B * K * (2 * d + 4) * sizeof(float);
}

extern "C" int ec_gcn_pack_weights(const float* W, const float* bias, float* Wp, int d, int dff, void* stream) {
  EC_REQUIRE(W && bias && Wp, "ec_gcn_pack_weights: null pointer");
  long long total = (long long)dff * (2 * d + 4);
  gcn_pack_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(W, bias, Wp, d, dff);
  return check_launch("ec_gcn_pack_weights");
}

extern "C" int ec_gcn_aggregate_split(const float* X, const float* adj, void* Z2, int B, int K, int d, int Kp,
                                      void* stream) {
  EC_REQUIRE(X && adj && Z2, "ec_gcn_aggregate_split: null pointer");
  EC_REQUIRE(Kp % 64 == 0 && Kp >= 2 * d + 2, "ec_gcn_aggregate_split: Kp must be a multiple of 64 and >= 2d+2");
  if (B == 0 || K == 0) return EC_OK;
  const size_t smem = (size_t)GA_ROWS * ((K + 3) & ~3) * sizeof(float);
  EC_REQUIRE(smem <= 48 * 1024, "ec_gcn_aggregate_split: K too large for the shared-memory tile");
  dim3 grid(cdiv(K, GA_ROWS), B);
  launch_pdl(gcn_aggregate_split_kernel, dim3(grid), dim3(256), (size_t)(smem), (cudaStream_t)stream, X, adj, (__half*)Z2, K, d, Kp);
  return check_launch("ec_gcn_aggregate_split");
}

extern "C" int ec_gcn(const float* X, const float* adj, const float* Wp, float* Y, int B, int K, int d, int dff,
                      float* workspace, size_t workspace_bytes, void* stream) {
  EC_REQUIRE(X && adj && Wp && Y && workspace, "ec_gcn: null pointer");
  EC_REQUIRE(workspace_bytes >= ec_workspace_bytes_gcn(B, K, d, dff), "ec_gcn: workspace too small");
  EC_REQUIRE(d % 4 == 0, "ec_gcn: d must be a multiple of 4");
  if (B == 0 || K == 0) return EC_OK;
  const int ldz = 2 * d + 4;
  const long long KK = (long long)K * K;
  // Z[:, d:2d] = A1 X   (batched [K,K] x [K,d])
  int rc = ec_gemm(adj + KK, X, workspace + d, K, d, K, K, d, ldz, /*b_kmajor=*/0, B, 2 * KK, (long long)K * d,
                   (long long)K * ldz, nullptr, EC_ACT_NONE, nullptr, nullptr, 0, 0, EC_RES_NONE, stream);
  if (rc) return rc;
  dim3 grid(cdiv(K, 8), B);
  gcn_fill_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, adj, workspace, K, d, ldz);
  rc = check_launch("ec_gcn(fill)");
  if (rc) return rc;
  // Y = relu(Z Wp^T): biases ride along as the two extra K columns
  return ec_gemm(workspace, Wp, Y, B * K, dff, ldz, ldz, ldz, dff, /*b_kmajor=*/1, 1, 0, 0, 0, nullptr,
                 EC_ACT_RELU, nullptr, nullptr, 0, 0, EC_RES_NONE, stream);
}
