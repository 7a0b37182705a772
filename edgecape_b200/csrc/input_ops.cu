// Input stage on the device: the reference's TopDownAffineFewShot crop (cv2.warpAffine, INTER_LINEAR,
// uint8) + ToTensor + NormalizeTensor, and TopDownGenerateTargetFewShot's MSRA heat-maps.
#include <math.h>

#include "common.cuh"

namespace ec {

struct WarpParams {
  double m00, m01, m02, m10, m11, m12;   // inverse matrix (crop pixel -> source pixel), double as OpenCV keeps it
  float mean[3], stdv[3];
};

// One thread per crop pixel.  OpenCV's 8-bit bilinear path, integer for integer: source coordinates in 10-bit fixed
// point (row term and column term rounded separately, + half of 1/32 pixel), truncated to 1/32 pixel; 15-bit weights
// 32 (32-fx)(32-fy) ...; (sum + 2^14) >> 15; pixels outside the source are 0.  Then x/255, (x - mean) / std in fp32
// with individually rounded operations (torchvision to_tensor / normalize).
__global__ void __launch_bounds__(256) warp_affine_normalize_kernel(const uint8_t* __restrict__ src, int Hs, int Ws,
                                                                   long long row_stride, float* __restrict__ out,
                                                                   int H, int W, WarpParams p) {
  pdl_launch_dependents();
  pdl_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const long long X0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(p.m01, (double)y), p.m02), 1024.0)) + 16;
  const long long Y0 = __double2ll_rn(__dmul_rn(__dadd_rn(__dmul_rn(p.m11, (double)y), p.m12), 1024.0)) + 16;
  const long long ad = __double2ll_rn(__dmul_rn(__dmul_rn(p.m00, (double)x), 1024.0));
  const long long bd = __double2ll_rn(__dmul_rn(__dmul_rn(p.m10, (double)x), 1024.0));
  const long long X = (X0 + ad) >> 5, Y = (Y0 + bd) >> 5;
  const long long sx = X >> 5, sy = Y >> 5;
  const int fx = (int)(X & 31), fy = (int)(Y & 31);
  const int w00 = 32 * (32 - fx) * (32 - fy), w01 = 32 * fx * (32 - fy), w10 = 32 * (32 - fx) * fy, w11 = 32 * fx * fy;
  int acc[3] = {0, 0, 0};
  auto add = [&](long long yy, long long xx, int w) {
    if (w == 0 || yy < 0 || yy >= Hs || xx < 0 || xx >= Ws) return;
    const uint8_t* q = src + yy * row_stride + xx * 3;
    acc[0] += w * q[0]; acc[1] += w * q[1]; acc[2] += w * q[2];
  };
  add(sy, sx, w00); add(sy, sx + 1, w01); add(sy + 1, sx, w10); add(sy + 1, sx + 1, w11);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    int v = (acc[c] + (1 << 14)) >> 15;
    v = min(max(v, 0), 255);
    const float t = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.0f), p.mean[c]), p.stdv[c]);
    out[((long long)c * H + y) * W + x] = t;
  }
}

// One CTA per keypoint: zero the H x W map, then the (6 sigma + 1)^2 un-normalised Gaussian patch around
// (int(x / stride_x + .5), int(y / stride_y + .5)), clipped at the border; weight = visibility, 0 when the patch
// lies completely outside (top_down_transform.py:170-196).
__global__ void __launch_bounds__(256) msra_targets_kernel(const float* __restrict__ joints, int ldj,
                                                          const float* __restrict__ visible, int ldv,
                                                          float* __restrict__ target, float* __restrict__ weight,
                                                          double stride_x, double stride_y, int W, int H, int sigma) {
  pdl_launch_dependents();
  pdl_wait();
  const long long k = blockIdx.x;
  float* t = target + k * H * W;
  const int tmp = 3 * sigma;
  const int mx = (int)((double)joints[k * ldj] / stride_x + 0.5), my = (int)((double)joints[k * ldj + 1] / stride_y + 0.5);
  const int ulx = mx - tmp, uly = my - tmp, brx = mx + tmp + 1, bry = my + tmp + 1;
  float wgt = visible[k * ldv];
  if (ulx >= W || uly >= H || brx < 0 || bry < 0) wgt = 0.f;
  if (threadIdx.x == 0) weight[k] = wgt;
  const bool on = wgt > 0.5f;
  const double inv = 1.0 / (2.0 * sigma * sigma);
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
    const int y = i / W, x = i % W;
    float v = 0.f;
    if (on && x >= ulx && x < brx && y >= uly && y < bry) {
      const int dx = x - mx, dy = y - my;
      v = (float)exp(-(double)(dx * dx + dy * dy) * inv);
    }
    t[i] = v;
  }
}

}  // namespace ec

using namespace ec;

extern "C" int ec_warp_affine_normalize_u8(const uint8_t* src, int Hs, int Ws, long long row_stride, const double* M,
                                           float* out, int H, int W, const float* mean, const float* stdv,
                                           void* stream) {
  EC_REQUIRE(src && M && out && mean && stdv, "ec_warp_affine_normalize_u8: null pointer");
  EC_REQUIRE(Hs > 0 && Ws > 0 && H > 0 && W > 0 && row_stride >= 3LL * Ws, "ec_warp_affine_normalize_u8: bad shape");
  // cv::invertAffineTransform, in double
  double D = M[0] * M[4] - M[1] * M[3];
  D = D != 0 ? 1.0 / D : 0.0;
  WarpParams p;
  p.m00 = M[4] * D; p.m11 = M[0] * D; p.m01 = M[1] * (-D); p.m10 = M[3] * (-D);
  p.m02 = -p.m00 * M[2] - p.m01 * M[5];
  p.m12 = -p.m10 * M[2] - p.m11 * M[5];
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean[c]; p.stdv[c] = stdv[c]; }
  dim3 grid(cdiv(W, 256), H);
  launch_pdl(warp_affine_normalize_kernel, grid, dim3(256), (size_t)0, (cudaStream_t)stream, src, Hs, Ws, row_stride, out,
             H, W, p);
  return check_launch("ec_warp_affine_normalize_u8");
}

extern "C" int ec_msra_targets(const float* joints, int ldj, const float* visible, int ldv, float* target, float* weight,
                               int n, int img_w, int img_h, int W, int H, float sigma, void* stream) {
  EC_REQUIRE(joints && visible && target && weight, "ec_msra_targets: null pointer");
  EC_REQUIRE(ldj >= 2 && ldv >= 1 && W > 0 && H > 0 && img_w > 0 && img_h > 0, "ec_msra_targets: bad shape");
  EC_REQUIRE(sigma >= 1.f && sigma == floorf(sigma), "ec_msra_targets: sigma must be a positive integer (MSRA encoding)");
  if (n == 0) return EC_OK;
  launch_pdl(msra_targets_kernel, dim3(n), dim3(256), (size_t)0, (cudaStream_t)stream, joints, ldj, visible, ldv, target,
             weight, (double)img_w / (double)W, (double)img_h / (double)H, W, H, (int)sigma);
  return check_launch("ec_msra_targets");
}
