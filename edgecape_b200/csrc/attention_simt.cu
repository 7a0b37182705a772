// fp32 fused attention (ec_attention) and the structural hop-bias MLP (ec_hop_bias).
//
// One thread owns one query row: q[D] and the output accumulator o[D] live in registers, key and
// value tiles are staged in shared memory with coalesced float4 loads and read back as warp-wide
// broadcasts (every lane reads the same key), scores are processed in chunks of 8 keys with one
// online-softmax rescale per chunk.  Exact fp32 arithmetic (expf, no fast-math): this kernel is
// the parity path for all attentions of the head and, until the tcgen05 attention is enabled,
// the ViT.
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace ec {

constexpr int ATT_ROWS = 128;  // query rows (threads) per CTA
constexpr int ATT_TK = 64;     // keys per shared-memory tile
constexpr int ATT_CH = 8;      // keys per online-softmax chunk

struct AttnParams {
  const float *Q, *K, *V;
  float* O;
  int B, H, Lq, Lk;
  int ldq, ldk, ldv, ldo;
  long long sq, sk, sv, so;
  float scale;
  const uint8_t* key_mask;
  const float* bias;
  __half* split_out;   // optional [B*Lq, 2*split_kp] split-fp16 form of O
  int split_kp;
};

template <int D>
__global__ void __launch_bounds__(ATT_ROWS) attention_kernel(AttnParams p) {
  __shared__ __align__(16) float Ks[ATT_TK][D];
  __shared__ __align__(16) float Vs[ATT_TK][D];
  __shared__ uint8_t Ms[ATT_TK];

  const int b = blockIdx.z, h = blockIdx.y;
  const int row = blockIdx.x * ATT_ROWS + threadIdx.x;
  const bool active = row < p.Lq;

  float q[D], o[D];
#pragma unroll
  for (int d = 0; d < D; ++d) o[d] = 0.f;
  if (active) {
    const float4* qp = reinterpret_cast<const float4*>(p.Q + (long long)b * p.sq + (long long)row * p.ldq + h * D);
#pragma unroll
    for (int d = 0; d < D / 4; ++d) {
      float4 v = __ldg(qp + d);
      q[4 * d + 0] = v.x * p.scale; q[4 * d + 1] = v.y * p.scale;
      q[4 * d + 2] = v.z * p.scale; q[4 * d + 3] = v.w * p.scale;
    }
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = 0.f;
  }
  float m_run = -INFINITY, l_run = 0.f;
  const float* bias_row =
      (p.bias && active) ? p.bias + (((long long)b * p.H + h) * p.Lq + row) * p.Lk : nullptr;

  const float* Kb = p.K + (long long)b * p.sk + h * D;
  const float* Vb = p.V + (long long)b * p.sv + h * D;

  for (int j0 = 0; j0 < p.Lk; j0 += ATT_TK) {
    const int nj = min(ATT_TK, p.Lk - j0);
    __syncthreads();
    // cooperative, coalesced tile load (float4 along the head dimension)
    for (int idx = threadIdx.x; idx < ATT_TK * (D / 4); idx += ATT_ROWS) {
      int j = idx / (D / 4), dq = idx % (D / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j < nj) {
        kv = __ldg(reinterpret_cast<const float4*>(Kb + (long long)(j0 + j) * p.ldk) + dq);
        vv = __ldg(reinterpret_cast<const float4*>(Vb + (long long)(j0 + j) * p.ldv) + dq);
      }
      *reinterpret_cast<float4*>(&Ks[j][dq * 4]) = kv;
      *reinterpret_cast<float4*>(&Vs[j][dq * 4]) = vv;
    }
    if (threadIdx.x < ATT_TK) {
      int j = threadIdx.x;
      uint8_t mk = 1;
      if (j < nj) mk = p.key_mask ? p.key_mask[(long long)b * p.Lk + j0 + j] : 0;
      Ms[j] = mk;
    }
    __syncthreads();

    for (int c0 = 0; c0 < nj; c0 += ATT_CH) {
      float s[ATT_CH];
      float cmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < ATT_CH; ++c) {
        const int j = c0 + c;   // j < ATT_TK always (ATT_TK % ATT_CH == 0); padded keys are masked
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < D / 4; ++d) {
          float4 kv = *reinterpret_cast<const float4*>(&Ks[j][d * 4]);
          acc = fmaf(q[4 * d + 0], kv.x, acc);
          acc = fmaf(q[4 * d + 1], kv.y, acc);
          acc = fmaf(q[4 * d + 2], kv.z, acc);
          acc = fmaf(q[4 * d + 3], kv.w, acc);
        }
        if (bias_row && j < nj) acc += __ldg(bias_row + j0 + j);
        if (Ms[j]) acc = -INFINITY;
        s[c] = acc;
        cmax = fmaxf(cmax, acc);
      }
      const float m_new = fmaxf(m_run, cmax);
      if (m_new == -INFINITY) continue;        // everything so far masked
      const float corr = expf(m_run - m_new);  // m_run = -inf -> 0
      l_run *= corr;
#pragma unroll
      for (int d = 0; d < D; ++d) o[d] *= corr;
#pragma unroll
      for (int c = 0; c < ATT_CH; ++c) {
        const int j = c0 + c;
        const float pj = expf(s[c] - m_new);   // masked -> exp(-inf) = 0
        l_run += pj;
#pragma unroll
        for (int d = 0; d < D / 4; ++d) {
          float4 vv = *reinterpret_cast<const float4*>(&Vs[j][d * 4]);
          o[4 * d + 0] = fmaf(pj, vv.x, o[4 * d + 0]);
          o[4 * d + 1] = fmaf(pj, vv.y, o[4 * d + 1]);
          o[4 * d + 2] = fmaf(pj, vv.z, o[4 * d + 2]);
          o[4 * d + 3] = fmaf(pj, vv.w, o[4 * d + 3]);
        }
      }
      m_run = m_new;
    }
  }
  if (active) {
    // a fully masked row is NaN in the reference (softmax of all -inf); the reference guards
    // against it (encoder_decoder.py:359-360) so it never occurs on the path.  We return 0.
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
    if (p.O) {
      float4* op = reinterpret_cast<float4*>(p.O + (long long)b * p.so + (long long)row * p.ldo + h * D);
#pragma unroll
      for (int d = 0; d < D / 4; ++d)
        op[d] = make_float4(o[4 * d + 0] * inv, o[4 * d + 1] * inv, o[4 * d + 2] * inv, o[4 * d + 3] * inv);
    }
    if (p.split_out) {
      __half* sp = p.split_out + ((long long)b * p.Lq + row) * (2 * p.split_kp) + h * D;
#pragma unroll
      for (int d = 0; d < D; d += 2) {
        const float v0 = o[d] * inv, v1 = o[d + 1] * inv;
        __half2 hi, lo;
        split_pair(v0, v1, hi, lo);
        *reinterpret_cast<__half2*>(sp + d) = hi;
        *reinterpret_cast<__half2*>(sp + p.split_kp + d) = lo;
      }
    }
  }
}

// bias[b,h,i,j] = b1[h] + sum_u w1[h,u] relu(b0[u] + sum_t w0[u,t] hops[t,b,i,j])
__global__ void __launch_bounds__(256) hop_bias_kernel(const float* __restrict__ hops,
                                                       const float* __restrict__ w0,
                                                       const float* __restrict__ b0,
                                                       const float* __restrict__ w1,
                                                       const float* __restrict__ b1, float* __restrict__ bias,
                                                       int B, int K, int n_hops, int hidden, int H) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* sw0 = sm;                       // hidden * n_hops
  float* sb0 = sw0 + hidden * n_hops;    // hidden
  float* sw1 = sb0 + hidden;             // H * hidden
  float* sb1 = sw1 + H * hidden;         // H
  for (int i = threadIdx.x; i < hidden * n_hops; i += blockDim.x) sw0[i] = w0[i];
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) sb0[i] = b0[i];
  for (int i = threadIdx.x; i < H * hidden; i += blockDim.x) sw1[i] = w1[i];
  for (int i = threadIdx.x; i < H; i += blockDim.x) sb1[i] = b1[i];
  __syncthreads();
  const long long KK = (long long)K * K;
  const long long total = (long long)B * KK;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / KK, ij = e % KK;
    float x[8], hid[32];
    for (int t = 0; t < n_hops; ++t) x[t] = hops[(long long)t * total + e];
    for (int u = 0; u < hidden; ++u) {
      float a = sb0[u];
      for (int t = 0; t < n_hops; ++t) a = fmaf(sw0[u * n_hops + t], x[t], a);
      hid[u] = fmaxf(a, 0.f);
    }
    for (int h = 0; h < H; ++h) {
      float a = sb1[h];
      for (int u = 0; u < hidden; ++u) a = fmaf(sw1[h * hidden + u], hid[u], a);
      bias[(b * H + h) * KK + ij] = a;
    }
  }
}

}  // namespace ec

using namespace ec;

extern "C" int ec_attention(const float* Q, const float* K, const float* V, float* O, int B, int H, int Lq,
                            int Lk, int D, int ldq, int ldk, int ldv, int ldo, long long sq, long long sk,
                            long long sv, long long so, float scale, const uint8_t* key_mask,
                            const float* bias, void* split_out, int split_kp, void* stream) {
  EC_REQUIRE(Q && K && V && (O || split_out), "ec_attention: null pointer");
  EC_REQUIRE(!split_out || split_kp == H * D, "ec_attention: split_out needs split_kp == H*D (a multiple of 64)");
  EC_REQUIRE(!split_out || (H * D) % 64 == 0, "ec_attention: split_out needs H*D to be a multiple of 64");
  EC_REQUIRE(B >= 0 && H > 0 && Lq >= 0 && Lk > 0, "ec_attention: bad shape");
  EC_REQUIRE(aligned16(Q) && aligned16(K) && aligned16(V) && (!O || aligned16(O)) && ldq % 4 == 0 && ldk % 4 == 0 &&
                 ldv % 4 == 0 && ldo % 4 == 0 && sq % 4 == 0 && sk % 4 == 0 && sv % 4 == 0 && so % 4 == 0,
             "ec_attention: operands must be 16-byte aligned with strides that are multiples of 4");
  if (B == 0 || Lq == 0) return EC_OK;
  EC_REQUIRE(B <= 65535 && H <= 65535, "ec_attention: grid too large");
  AttnParams p{Q, K, V, O, B, H, Lq, Lk, ldq, ldk, ldv, ldo, sq, sk, sv, so, scale, key_mask, bias,
               (__half*)split_out, split_kp};
  dim3 grid(cdiv(Lq, ATT_ROWS), H, B);
  cudaStream_t st = (cudaStream_t)stream;
  switch (D) {
    case 16: attention_kernel<16><<<grid, ATT_ROWS, 0, st>>>(p); break;
    case 32: attention_kernel<32><<<grid, ATT_ROWS, 0, st>>>(p); break;
    case 64: attention_kernel<64><<<grid, ATT_ROWS, 0, st>>>(p); break;
    default:
      set_error("ec_attention: head dim %d not instantiated (16, 32, 64)", D);
      return EC_ERR_UNSUPPORTED;
  }
  return check_launch("ec_attention");
}

extern "C" int ec_hop_bias(const float* attn_adj, const float* w0, const float* b0, const float* w1,
                           const float* b1, float* bias, int B, int K, int n_hops, int hidden, int H,
                           void* stream) {
  EC_REQUIRE(attn_adj && w0 && b0 && w1 && b1 && bias, "ec_hop_bias: null pointer");
  EC_REQUIRE(n_hops <= 8 && hidden <= 32 && H <= 64, "ec_hop_bias: MLP larger than instantiated (8, 32, 64)");
  long long total = (long long)B * K * K;
  if (total == 0) return EC_OK;
  size_t smem = sizeof(float) * (hidden * n_hops + hidden + H * hidden + H);
  int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
  launch_pdl(hop_bias_kernel, dim3(blocks), dim3(256), (size_t)(smem), (cudaStream_t)stream, attn_adj, w0, b0, w1, b1, bias, B, K, n_hops,
                                                               hidden, H);
  return check_launch("ec_hop_bias");
}
