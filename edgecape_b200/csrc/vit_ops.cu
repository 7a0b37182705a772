// ViT front-end kernels: patch im2col, bicubic position-table resampling, cls row.
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace ec {

__global__ void im2col_kernel(const float* __restrict__ img, float* __restrict__ cols, int H, int W, int P,
                              int h0, int w0, int ldc, long long total, __half* __restrict__ split_out, int split_kp) {
  pdl_launch_dependents();
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int col = (int)(i % ldc);
  const long long r = i / ldc;
  const int px = (int)(r % w0), py = (int)((r / w0) % h0);
  const long long b = r / ((long long)w0 * h0);
  float v = 0.f;
  if (col < 3 * P * P) {
    const int c = col / (P * P), iy = (col / P) % P, ix = col % P;
    v = img[((b * 3 + c) * H + (py * P + iy)) * W + (px * P + ix)];
  }
  if (cols) cols[i] = v;
  if (split_out) {
    const __half hi = __float2half_rn(v);
    split_out[r * 2 * split_kp + col] = hi;
    split_out[r * 2 * split_kp + split_kp + col] = __float2half_rn(v - __half2float(hi));
  }
}

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__global__ void interp_pos_kernel(const float* __restrict__ pos, float* __restrict__ out, int Mg, int h0, int w0,
                                  int C, float ratio_h, float ratio_w, int identity) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)(1 + h0 * w0) * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int tkn = (int)(i / C);
  if (tkn == 0 || identity) {
    out[i] = pos[i];
    return;
  }
  const int oy = (tkn - 1) / w0, ox = (tkn - 1) % w0;
  const float A = -0.75f;
  const float ry = ratio_h * ((float)oy + 0.5f) - 0.5f;
  const float rx = ratio_w * ((float)ox + 0.5f) - 0.5f;
  const int iy = (int)floorf(ry), ix = (int)floorf(rx);
  const float ty = ry - (float)iy, tx = rx - (float)ix;
  float wy[4] = {cubic2(ty + 1.f, A), cubic1(ty, A), cubic1(1.f - ty, A), cubic2(2.f - ty, A)};
  float wx[4] = {cubic2(tx + 1.f, A), cubic1(tx, A), cubic1(1.f - tx, A), cubic2(2.f - tx, A)};
  float acc = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int yy = min(max(iy - 1 + a, 0), Mg - 1);
    float rowv = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int xx = min(max(ix - 1 + b, 0), Mg - 1);
      rowv += wx[b] * pos[(long long)(1 + yy * Mg + xx) * C + c];
    }
    acc += wy[a] * rowv;
  }
  out[i] = acc;
}

__global__ void write_cls_kernel(const float* __restrict__ cls, const float* __restrict__ pos0,
                                 float* __restrict__ tokens, long long stride, int C, int B) {
  pdl_launch_dependents();
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int b = i / C, c = i % C;
  tokens[(long long)b * stride + c] = cls[c] + pos0[c];
}

}  // namespace ec

using namespace ec;

extern "C" int ec_im2col_patches(const float* img, float* cols, int B, int H, int W, int P, int ldc,
                                 void* split_out, int split_kp, void* stream) {
  EC_REQUIRE(img && (cols || split_out) && P > 0 && ldc >= 3 * P * P, "ec_im2col_patches: bad arguments");
  EC_REQUIRE(!split_out || (split_kp == ldc && split_kp % 64 == 0), "ec_im2col_patches: split_out needs ldc == split_kp (multiple of 64)");
  const int h0 = H / P, w0 = W / P;
  long long total = (long long)B * h0 * w0 * ldc;
  if (total == 0) return EC_OK;
  launch_pdl(im2col_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, img, cols, H, W, P, h0, w0, ldc, total,
                                                                    (__half*)split_out, split_kp);
  return check_launch("ec_im2col_patches");
}

extern "C" int ec_interp_pos_embed(const float* pos_embed, float* pos_out, int Mgrid, int h0, int w0, int C,
                                   double offset, void* stream) {
  EC_REQUIRE(pos_embed && pos_out && Mgrid > 0 && h0 > 0 && w0 > 0, "ec_interp_pos_embed: bad arguments");
  // torch: scale_factor given -> ratio = float(1 / scale_factor); size given -> in / out
  float rh, rw;
  if (offset != 0.0) {
    rh = (float)(1.0 / ((h0 + offset) / (double)Mgrid));
    rw = (float)(1.0 / ((w0 + offset) / (double)Mgrid));
  } else {
    rh = (float)Mgrid / (float)h0;
    rw = (float)Mgrid / (float)w0;
  }
  const int identity = (h0 == Mgrid && w0 == Mgrid) ? 1 : 0;
  long long total = (long long)(1 + h0 * w0) * C;
  interp_pos_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(pos_embed, pos_out, Mgrid, h0, w0, C, rh,
                                                                        rw, identity);
  return check_launch("ec_interp_pos_embed");
}

extern "C" int ec_write_cls(const float* cls, const float* pos0, float* tokens, int B, long long stride, int C,
                            void* stream) {
  EC_REQUIRE(cls && pos0 && tokens, "ec_write_cls: null pointer");
  if (B == 0) return EC_OK;
  launch_pdl(write_cls_kernel, dim3(cdiv((long long)B * C, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, cls, pos0, tokens, stride, C, B);
  return check_launch("ec_write_cls");
}
