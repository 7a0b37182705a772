// Row-wise kernels: LayerNorm (+ fused residual), positional adds, row copies, L2 normalisation,
// mask bookkeeping and the sigmoid-space point update.  One warp per row, shuffle reductions.
#include <cuda_fp16.h>

#include "common.cuh"

namespace ec {

// ------------------------------------------------------------------------------ LayerNorm
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ X, int ldx, int seg,
                                                        long long seg_stride, const float* __restrict__ R,
                                                        int ldr, float* __restrict__ sum_out, int ld_sum,
                                                        float* __restrict__ Y, int ldy,
                                                        const float* __restrict__ w,
                                                        const float* __restrict__ b, float eps, int M, int C,
                                                        __half* __restrict__ split_out, int split_kp, int split_fmt,
                                                        unsigned long long* overflow) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int m = warp;
  const float* x = seg > 0 ? X + (long long)(m / seg) * seg_stride + (long long)(m % seg) * ldx
                           : X + (long long)m * ldx;
  const float* r = R ? R + (long long)m * ldr : nullptr;
  float* so = sum_out ? sum_out + (long long)m * ld_sum : nullptr;
  // pass 1: mean (and optional materialisation of x + r)
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    float v = x[c];
    if (r) v += r[c];
    if (so) so[c] = v;
    s += v;
  }
  const float mean = warp_sum(s) / (float)C;
  // pass 2: centred second moment (two-pass variance, as torch's CPU/CUDA kernels)
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    float v = x[c];
    if (r) v += r[c];
    float d = v - mean;
    q = fmaf(d, d, q);
  }
  const float var = warp_sum(q) / (float)C;
  const float rstd = 1.0f / sqrtf(var + eps);
  float* y = Y ? Y + (long long)m * ldy : nullptr;
  __half* sp = split_out ? split_out + (long long)m * 2 * split_kp : nullptr;
  uint32_t ovf = 0;
  for (int c = lane; c < C; c += 32) {
    float v = x[c];
    if (r) v += r[c];
    const float o = (v - mean) * rstd * w[c] + b[c];
    if (y) y[c] = o;
    if (sp) {   // split form for the tensor-core GEMM that consumes this row (scalar stores: odd widths only)
      const __half hi = __float2half_rn(o);
      const float hf = __half2float(hi);
      sp[c] = hi;
      if (split_fmt == EC_SPLIT_F16X2) {
        sp[split_kp + c] = __float2half_rn(o - hf);
      } else {
        uint8_t* b8 = reinterpret_cast<uint8_t*>(sp);
        b8[f8_off(split_kp, c)] = (uint8_t)(e4m3x2(hf, 0.f) & 0xff);
        b8[f8_off(split_kp, c) + 64] = (uint8_t)(e4m3x2((o - hf) * 2048.f, 0.f) & 0xff);
        ovf |= (fabsf(o) > 448.f ? 1u : 0u) | (fabsf(o) > 65504.f ? 2u : 0u);
      }
    }
  }
  if (sp)   // zero padding of columns [C, split_kp): all-zero bytes are +0 in fp16 and in e4m3
    for (int c = C + lane; c < split_kp; c += 32) {
      sp[c] = __float2half_rn(0.f);
      if (split_fmt == EC_SPLIT_F16X2) {
        sp[split_kp + c] = __float2half_rn(0.f);
      } else {
        uint8_t* b8 = reinterpret_cast<uint8_t*>(sp);
        b8[f8_off(split_kp, c)] = 0;
        b8[f8_off(split_kp, c) + 64] = 0;
      }
    }
  report_overflow(overflow, ovf);
}

// Vector form for C = NV * 128 with 16-byte aligned rows (every LayerNorm of the ViT and the head): the row
// lives in registers (NV float4 per lane), so memory is read once; outputs are float4 / packed-fp16 stores.
template <int NV>
__global__ void __launch_bounds__(256) layernorm_vec_kernel(const float* __restrict__ X, int ldx, int seg,
                                                            long long seg_stride, const float* __restrict__ R,
                                                            int ldr, float* __restrict__ sum_out, int ld_sum,
                                                            float* __restrict__ Y, int ldy,
                                                            const float* __restrict__ w,
                                                            const float* __restrict__ b, float eps, int M,
                                                            __half* __restrict__ split_out, int split_kp, int split_fmt,
                                                            unsigned long long* overflow) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int C = NV * 128;
  const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  const float* x = seg > 0 ? X + (long long)(m / seg) * seg_stride + (long long)(m % seg) * ldx
                           : X + (long long)m * ldx;
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = *reinterpret_cast<const float4*>(x + (j * 32 + lane) * 4);
  if (R) {
    const float* r = R + (long long)m * ldr;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 t = *reinterpret_cast<const float4*>(r + (j * 32 + lane) * 4);
      v[j].x += t.x; v[j].y += t.y; v[j].z += t.z; v[j].w += t.w;
    }
  }
  if (sum_out) {
    float* so = sum_out + (long long)m * ld_sum;
#pragma unroll
    for (int j = 0; j < NV; ++j) *reinterpret_cast<float4*>(so + (j * 32 + lane) * 4) = v[j];
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
    q = fmaf(v[j].x, v[j].x, q); q = fmaf(v[j].y, v[j].y, q); q = fmaf(v[j].z, v[j].z, q); q = fmaf(v[j].w, v[j].w, q);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + eps);
  float* y = Y ? Y + (long long)m * ldy : nullptr;
  __half* sp = split_out ? split_out + (long long)m * 2 * split_kp : nullptr;
  uint32_t ovf = 0;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = (j * 32 + lane) * 4;
    const float4 ww = __ldg(reinterpret_cast<const float4*>(w + c)), bb = __ldg(reinterpret_cast<const float4*>(b + c));
    float4 o;
    o.x = v[j].x * rstd * ww.x + bb.x;
    o.y = v[j].y * rstd * ww.y + bb.y;
    o.z = v[j].z * rstd * ww.z + bb.z;
    o.w = v[j].w * rstd * ww.w + bb.w;
    if (y) *reinterpret_cast<float4*>(y + c) = o;
    if (sp) store_split4(reinterpret_cast<uint8_t*>(sp), split_kp, c, o.x, o.y, o.z, o.w, split_fmt, ovf);
  }
  if (sp)
    for (int c = C + lane; c < split_kp; c += 32) { sp[c] = __float2half_rn(0.f); sp[split_kp + c] = __float2half_rn(0.f); }
  report_overflow(overflow, ovf);
}

__global__ void add_rows_kernel(float* __restrict__ X, const float* __restrict__ P, int T, int S, int C,
                                long long total) {
  pdl_launch_dependents();
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long row = i / C;
  int s = (int)(row % S);
  long long b = row / S;
  X[(b * T + s) * C + c] += P[(long long)s * C + c];
}

// dst block r = src block idx[r]  (blocks of `n` contiguous floats; support de-duplication expands the features of the
// unique support images back to one block per batch row)
__global__ void gather_blocks_kernel(const float* __restrict__ src, long long src_stride, const int* __restrict__ idx,
                                     float* __restrict__ dst, long long dst_stride, long long n, int vec) {
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.y;
  const float* s = src + (long long)idx[r] * src_stride;
  float* d = dst + (long long)r * dst_stride;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    if (i * 4 < n) reinterpret_cast<float4*>(d)[i] = __ldg(reinterpret_cast<const float4*>(s) + i);
  } else if (i < n) {
    d[i] = s[i];
  }
}

__global__ void copy_rows_kernel(const float* __restrict__ X, int ldx, int segx, long long sstridex,
                                 float* __restrict__ Y, int ldy, int segy, long long sstridey, int C,
                                 int bcast_rows, long long total) {
  pdl_launch_dependents();
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int c = (int)(i % C);
  long long m = i / C;
  long long ms = bcast_rows > 0 ? (m % bcast_rows) : m;
  const float* x = segx > 0 ? X + (ms / segx) * sstridex + (ms % segx) * ldx : X + ms * ldx;
  float* y = segy > 0 ? Y + (m / segy) * sstridey + (m % segy) * ldy : Y + m * ldy;
  y[c] = x[c];
}

__global__ void __launch_bounds__(256) l2_normalize_kernel(const float* __restrict__ X, float* __restrict__ Y,
                                                           int M, int C, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* x = X + (long long)warp * C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) q = fmaf(x[c], x[c], q);
  const float nrm = sqrtf(warp_sum(q)) + eps;
  float* y = Y + (long long)warp * C;
  for (int c = lane; c < C; c += 32) y[c] = x[c] / nrm;
}

__global__ void mask_accumulate_kernel(const float* __restrict__ tw, float* __restrict__ mask_s, int n,
                                       int first) {
  pdl_launch_dependents();
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = tw[i];
  mask_s[i] = first ? t * t : mask_s[i] * t;
}

__global__ void kp_masks_kernel(const float* __restrict__ mask_s, uint8_t* __restrict__ kp_mask,
                                uint8_t* __restrict__ kp_fixed, int K) {
  pdl_launch_dependents();
  pdl_wait();
  // one block per batch row
  const int b = blockIdx.x;
  __shared__ int any_valid;
  if (threadIdx.x == 0) any_valid = 0;
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    // reference: kp_mask = ~mask_s.to(bool): masked iff mask_s == 0
    uint8_t m = (mask_s[(long long)b * K + k] == 0.0f) ? 1 : 0;
    kp_mask[(long long)b * K + k] = m;
    kp_fixed[(long long)b * K + k] = m;
    if (!m) any_valid = 1;
  }
  __syncthreads();
  if (threadIdx.x == 0 && !any_valid) kp_fixed[(long long)b * K] = 0;
}

__device__ __forceinline__ float inv_sigmoid(float x) {
  const float eps = 1e-3f;
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  float x1 = fmaxf(x, eps);
  float x2 = fmaxf(1.0f - x, eps);
  return logf(x1 / x2);
}

__global__ void point_update_kernel(const float* __restrict__ bi, const float* __restrict__ delta, int ldd,
                                    float* __restrict__ out, int M) {
  pdl_launch_dependents();
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * M) return;
  int m = i >> 1, c = i & 1;
  float z = inv_sigmoid(bi[i]) + delta[(long long)m * ldd + c];
  out[i] = 1.0f / (1.0f + expf(-z));
}

__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                             float a, float b, float div, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = a * x[i];
  if (b != 0.f) v += b * y[i];
  out[i] = div != 1.0f ? v / div : v;
}

}  // namespace ec

using namespace ec;

extern "C" int ec_axpby(const float* x, const float* y, float* out, float a, float b, float div, long long n,
                        void* stream) {
  EC_REQUIRE(x && y && out, "ec_axpby: null pointer");
  if (n == 0) return EC_OK;
  launch_pdl(axpby_kernel, dim3(cdiv(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, x, y, out, a, b, div, n);
  return check_launch("ec_axpby");
}

extern "C" int ec_layernorm(const float* X, int ldx, int seg, long long seg_stride, const float* R, int ldr,
                            float* sum_out, int ld_sum, float* Y, int ldy, const float* w, const float* b,
                            float eps, int M, int C, void* split_out, int split_kp, int split_fmt, void* stream) {
  EC_REQUIRE(X && (Y || split_out) && w && b, "ec_layernorm: null pointer");
  EC_REQUIRE(split_fmt == EC_SPLIT_F16X2 || split_fmt == EC_SPLIT_F16F8, "ec_layernorm: bad split_fmt");
  unsigned long long* ovf = nullptr;
  if (split_out && split_fmt == EC_SPLIT_F16F8) {
    ovf = overflow_counters();
    if (!ovf) return EC_ERR_CUDA;
  }
  EC_REQUIRE(!split_out || (split_kp >= C && split_kp % 64 == 0), "ec_layernorm: bad split_kp");
  EC_REQUIRE(M >= 0 && C > 0, "ec_layernorm: bad shape");
  if (M == 0) return EC_OK;
  const int warps_per_block = 8;
  const dim3 grid(cdiv(M, warps_per_block)), block(warps_per_block * 32);
  cudaStream_t st = (cudaStream_t)stream;
  auto al16 = [](const void* p) { return (((uintptr_t)p) & 15) == 0; };
  const bool vec = C % 128 == 0 && al16(X) && ldx % 4 == 0 && (seg <= 0 || seg_stride % 4 == 0) &&
                   (!R || (al16(R) && ldr % 4 == 0)) && (!sum_out || (al16(sum_out) && ld_sum % 4 == 0)) &&
                   (!Y || (al16(Y) && ldy % 4 == 0)) && al16(w) && al16(b) && (!split_out || al16(split_out));
#define EC_LN_VEC(NV)                                                                                             \
  launch_pdl(layernorm_vec_kernel<NV>, grid, block, 0, st, X, ldx, seg, seg_stride, R, ldr, sum_out, ld_sum, Y, ldy, w, \
             b, eps, M, (__half*)split_out, split_kp, split_fmt, ovf)
  if (vec && C == 256) EC_LN_VEC(2);
  else if (vec && C == 384) EC_LN_VEC(3);
  else if (vec && C == 768) EC_LN_VEC(6);
  else if (vec && C == 1024) EC_LN_VEC(8);
  else
    launch_pdl(layernorm_kernel, grid, block, 0, st, X, ldx, seg, seg_stride, R, ldr, sum_out, ld_sum, Y, ldy, w, b, eps,
               M, C, (__half*)split_out, split_kp, split_fmt, ovf);
#undef EC_LN_VEC
  return check_launch("ec_layernorm");
}

extern "C" int ec_add_rows(float* X, const float* P, int batch, int T, int S, int C, void* stream) {
  EC_REQUIRE(X && P && S <= T, "ec_add_rows: bad arguments");
  long long total = (long long)batch * S * C;
  if (total == 0) return EC_OK;
  launch_pdl(add_rows_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, X, P, T, S, C, total);
  return check_launch("ec_add_rows");
}

extern "C" int ec_copy_rows(const float* X, int ldx, int seg_x, long long seg_stride_x, float* Y, int ldy,
                            int seg_y, long long seg_stride_y, int M, int C, int bcast_rows, void* stream) {
  EC_REQUIRE(X && Y, "ec_copy_rows: null pointer");
  long long total = (long long)M * C;
  if (total == 0) return EC_OK;
  launch_pdl(copy_rows_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, X, ldx, seg_x, seg_stride_x, Y, ldy, seg_y,
                                                                       seg_stride_y, C, bcast_rows, total);
  return check_launch("ec_copy_rows");
}

extern "C" int ec_gather_blocks(const float* src, long long src_stride, const int32_t* idx, float* dst,
                                long long dst_stride, int n_out, long long block_elems, void* stream) {
  EC_REQUIRE(src && idx && dst && block_elems > 0 && n_out >= 0, "ec_gather_blocks: bad arguments");
  if (n_out == 0) return EC_OK;
  const int vec = (block_elems % 4 == 0 && src_stride % 4 == 0 && dst_stride % 4 == 0 && (((uintptr_t)src) & 15) == 0 &&
                   (((uintptr_t)dst) & 15) == 0) ? 1 : 0;
  const long long work = vec ? block_elems / 4 : block_elems;
  dim3 grid((unsigned)cdiv(work, 256), n_out);
  launch_pdl(gather_blocks_kernel, grid, dim3(256), (size_t)0, (cudaStream_t)stream, src, src_stride, (const int*)idx, dst,
             dst_stride, block_elems, vec);
  return check_launch("ec_gather_blocks");
}

extern "C" int ec_l2_normalize(const float* X, float* Y, int M, int C, float eps, void* stream) {
  EC_REQUIRE(X && Y, "ec_l2_normalize: null pointer");
  if (M == 0) return EC_OK;
  launch_pdl(l2_normalize_kernel, dim3(cdiv(M, 8)), dim3(256), (size_t)(0), (cudaStream_t)stream, X, Y, M, C, eps);
  return check_launch("ec_l2_normalize");
}

extern "C" int ec_mask_accumulate(const float* tw, float* mask_s, int n, int first, void* stream) {
  EC_REQUIRE(tw && mask_s, "ec_mask_accumulate: null pointer");
  if (n == 0) return EC_OK;
  launch_pdl(mask_accumulate_kernel, dim3(cdiv(n, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, tw, mask_s, n, first);
  return check_launch("ec_mask_accumulate");
}

extern "C" int ec_kp_masks(const float* mask_s, uint8_t* kp_mask, uint8_t* kp_mask_fixed, int B, int K,
                           void* stream) {
  EC_REQUIRE(mask_s && kp_mask && kp_mask_fixed, "ec_kp_masks: null pointer");
  if (B == 0) return EC_OK;
  launch_pdl(kp_masks_kernel, dim3(B), dim3(128), (size_t)(0), (cudaStream_t)stream, mask_s, kp_mask, kp_mask_fixed, K);
  return check_launch("ec_kp_masks");
}

extern "C" int ec_point_update(const float* bi, const float* delta, int ldd, float* out, int M, void* stream) {
  EC_REQUIRE(bi && delta && out, "ec_point_update: null pointer");
  if (M == 0) return EC_OK;
  launch_pdl(point_update_kernel, dim3(cdiv(2LL * M, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, bi, delta, ldd, out, M);
  return check_launch("ec_point_update");
}
