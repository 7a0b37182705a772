// Head-specific kernels: support-keypoint pooling weights, sine positional encoding of
// coordinates, the ProposalGenerator tail (softmax / argmax / local soft-argmax) and PCK counters.
#include <math.h>

#include "common.cuh"

namespace ec {

// Bilinear (align_corners=False) source index / weights of F.interpolate(size=...).
__device__ __forceinline__ void bilinear_src(int dst, int in_size, int out_size, int& i0, int& i1, float& l0,
                                             float& l1) {
  const float scale = (float)in_size / (float)out_size;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.0f - l1;
}

// One CTA per (b,k).  Tw[sy,sx] = scale/(sum t + 1e-8) * sum_{py,px} t[py,px] Wy[py,sy] Wx[px,sx].
__global__ void __launch_bounds__(256) support_weights_kernel(const float* __restrict__ target,
                                                              const float* __restrict__ rowscale,
                                                              float* __restrict__ Tw, int ldtw, int hm_h, int hm_w,
                                                              int h, int w) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];
  float* t = sm;                    // hm_h * hm_w
  float* Wx = t + hm_h * hm_w;      // hm_w * w
  float* Wy = Wx + hm_w * w;        // hm_h * h
  float* Rr = Wy + hm_h * h;        // hm_h * w
  __shared__ float red[8];
  const int bk = blockIdx.x;
  const float* tg = target + (long long)bk * hm_h * hm_w;
  float part = 0.f;
  for (int i = threadIdx.x; i < hm_h * hm_w; i += blockDim.x) {
    float v = tg[i];
    t[i] = v;
    part += v;
  }
  for (int i = threadIdx.x; i < hm_w * w; i += blockDim.x) Wx[i] = 0.f;
  for (int i = threadIdx.x; i < hm_h * h; i += blockDim.x) Wy[i] = 0.f;
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
  __syncthreads();
  // each px / py is owned by exactly one thread, so the += below never races
  for (int px = threadIdx.x; px < hm_w; px += blockDim.x) {
    int i0, i1; float l0, l1;
    bilinear_src(px, w, hm_w, i0, i1, l0, l1);
    Wx[px * w + i0] += l0;
    Wx[px * w + i1] += l1;
  }
  for (int py = threadIdx.x; py < hm_h; py += blockDim.x) {
    int i0, i1; float l0, l1;
    bilinear_src(py, h, hm_h, i0, i1, l0, l1);
    Wy[py * h + i0] += l0;
    Wy[py * h + i1] += l1;
  }
  __syncthreads();
  // the interpolation matrices are banded (every heat-map pixel feeds at most two grid cells): per grid column / row
  // the range of heat-map pixels with a non-zero weight.  Skipping the exact zeros leaves the sums bit-identical and
  // cuts the two contractions from 64 to ~8 terms per output (the dense form was shared-memory bound: 74 us).
  __shared__ int xlo[128], xhi[128], ylo[128], yhi[128];
  for (int sx = threadIdx.x; sx < w; sx += blockDim.x) {
    int lo = hm_w, hi = -1;
    for (int px = 0; px < hm_w; ++px)
      if (Wx[px * w + sx] != 0.f) { lo = min(lo, px); hi = px; }
    xlo[sx] = lo; xhi[sx] = hi;
  }
  for (int sy = threadIdx.x; sy < h; sy += blockDim.x) {
    int lo = hm_h, hi = -1;
    for (int py = 0; py < hm_h; ++py)
      if (Wy[py * h + sy] != 0.f) { lo = min(lo, py); hi = py; }
    ylo[sy] = lo; yhi[sy] = hi;
  }
  __syncthreads();
  float total = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) total += red[i];
  const float scale = (rowscale ? rowscale[bk] : 1.0f) / (total + 1e-8f);
  for (int i = threadIdx.x; i < hm_h * w; i += blockDim.x) {
    const int py = i / w, sx = i % w;
    float a = 0.f;
    for (int px = xlo[sx]; px <= xhi[sx]; ++px) a = fmaf(t[py * hm_w + px], Wx[px * w + sx], a);
    Rr[i] = a;
  }
  __syncthreads();
  float* out = Tw + (long long)bk * ldtw;
  for (int i = threadIdx.x; i < h * w; i += blockDim.x) {
    const int sy = i / w, sx = i % w;
    float a = 0.f;
    for (int py = ylo[sy]; py <= yhi[sy]; ++py) a = fmaf(Wy[py * h + sy], Rr[py * w + sx], a);
    out[i] = a * scale;
  }
}

__global__ void sine_pe_kernel(const float* __restrict__ coord, float* __restrict__ out, int ldo, int M,
                               int num_feats, float temperature, float scale) {
  pdl_launch_dependents();
  pdl_wait();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int C = 2 * num_feats;
  if (i >= (long long)M * C) return;
  const int m = (int)(i / C), j = (int)(i % C);
  const int half = j / num_feats, jj = j % num_feats;
  // channel order [y-half | x-half]; coord = (x, y)
  const float v = coord[2 * m + (half == 0 ? 1 : 0)] * scale;
  const float dim_t = powf(temperature, (float)(2 * (jj / 2)) / (float)num_feats);
  const float a = v / dim_t;
  out[(long long)m * ldo + j] = (jj & 1) ? cosf(a) : sinf(a);
}

// One warp per (b,k) row of the similarity map.
__global__ void __launch_bounds__(256) proposal_kernel(const float* __restrict__ sim,
                                                       float* __restrict__ prop_loss, float* __restrict__ prop,
                                                       long long* __restrict__ argmax, int BK, int h, int w) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= BK) return;
  const int S = h * w;
  const float* s = sim + (long long)row * S;
  // max / first argmax
  float best = -INFINITY;
  int bidx = 0x7fffffff;
  for (int i = lane; i < S; i += 32) {
    float v = s[i];
    if (v > best || (v == best && i < bidx)) { best = v; bidx = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
  }
  if (bidx == 0x7fffffff) bidx = 0;   // all-NaN row: torch returns the first NaN; not on the path
  float den = 0.f;
  for (int i = lane; i < S; i += 32) den += expf(s[i] - best);
  den = warp_sum(den);
  // the reference reshapes the one-hot as (w, h) before the 3x3 max-pool (encoder_decoder.py:93)
  const int ra = bidx / h, ca = bidx % h;
  float gx = 0.f, gy = 0.f, lsum = 0.f, lx = 0.f, ly = 0.f;
  for (int i = lane; i < S; i += 32) {
    const float p = expf(s[i] - best) / den;
    const float cx = (float)(i % w) + 0.5f, cy = (float)(i / w) + 0.5f;
    gx = fmaf(p, cx, gx);
    gy = fmaf(p, cy, gy);
    const int r = i / h, c = i % h;
    if (abs(r - ra) <= 1 && abs(c - ca) <= 1) {
      lsum += p;
      lx = fmaf(p, cx, lx);
      ly = fmaf(p, cy, ly);
    }
  }
  gx = warp_sum(gx); gy = warp_sum(gy); lsum = warp_sum(lsum); lx = warp_sum(lx); ly = warp_sum(ly);
  if (lane == 0) {
    prop_loss[2 * row + 0] = gx / (float)w;
    prop_loss[2 * row + 1] = gy / (float)h;
    const float inv = 1.0f / (lsum + 1e-10f);
    prop[2 * row + 0] = lx * inv / (float)w;
    prop[2 * row + 1] = ly * inv / (float)h;
    argmax[row] = bidx;
  }
}

// per-sample PCK (mmpose keypoint_pck_accuracy with N = 1): fraction of valid keypoints whose
// normalised distance is below thr; samples without valid keypoints contribute 0.
__global__ void __launch_bounds__(128) pck_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                  const uint8_t* __restrict__ valid, const float* __restrict__ norm,
                                                  const float* __restrict__ thr, int T, double* counters, int K) {
  const int b = blockIdx.x;
  __shared__ int hits[16];
  __shared__ int nvalid;
  if (threadIdx.x < 16) hits[threadIdx.x] = 0;
  if (threadIdx.x == 0) nvalid = 0;
  __syncthreads();
  const float nx = norm[2 * b], ny = norm[2 * b + 1];
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    if (!valid[(long long)b * K + k]) continue;
    const float dx = (pred[((long long)b * K + k) * 2] - gt[((long long)b * K + k) * 2]) / nx;
    const float dy = (pred[((long long)b * K + k) * 2 + 1] - gt[((long long)b * K + k) * 2 + 1]) / ny;
    const float dist = sqrtf(dx * dx + dy * dy);
    atomicAdd(&nvalid, 1);
    for (int t = 0; t < T; ++t)
      if (dist < thr[t]) atomicAdd(&hits[t], 1);
  }
  __syncthreads();
  if (threadIdx.x < T) {
    double v = nvalid > 0 ? (double)hits[threadIdx.x] / (double)nvalid : 0.0;
    atomicAdd(&counters[threadIdx.x], v);
  }
  if (threadIdx.x == 0) atomicAdd(&counters[T], 1.0);
}

// All four metrics of the reference's test loop for one sample per CTA (test_base_dataset.py:119-154 on mmpose 0.29
// keypoint_pck_accuracy / keypoint_nme / keypoint_auc / keypoint_epe with N = 1):
//   counters[0..T-1] += PCK@thr[t]   (valid keypoints with normalised distance < thr) / valid
//   counters[T]      += NME          mean normalised distance over valid keypoints
//   counters[T+1]    += AUC          mean over i < auc_steps of PCK@(i / auc_steps), normaliser norm[b,0] on both axes
//   counters[T+2]    += EPE          mean un-normalised distance over valid keypoints
//   counters[T+3]    += 1
// A zero in the normaliser masks the whole sample for PCK / NME / AUC; a negative one becomes 1e6 (_calc_distances).
// Distances are formed in fp64 and rounded to fp32 as mmpose does (float64 inputs, float32 distance table).
__global__ void __launch_bounds__(128) metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                      const uint8_t* __restrict__ valid, const float* __restrict__ norm,
                                                      const float* __restrict__ thr, int T, int auc_steps,
                                                      double* counters, int K) {
  const int b = blockIdx.x;
  __shared__ int hits[16], auc_hits[64];
  __shared__ int nvalid, nvalid_n;
  __shared__ double sum_n, sum_e;
  if (threadIdx.x < 16) hits[threadIdx.x] = 0;
  if (threadIdx.x < 64) auc_hits[threadIdx.x] = 0;
  if (threadIdx.x == 0) { nvalid = 0; nvalid_n = 0; sum_n = 0.0; sum_e = 0.0; }
  __syncthreads();
  float nx = norm[2 * b], ny = norm[2 * b + 1];
  const bool norm_ok = nx != 0.f && ny != 0.f;          // a zero normaliser masks the sample (not for EPE)
  if (nx <= 0.f) nx = 1e6f;
  if (ny <= 0.f) ny = 1e6f;
  const float ax = norm[2 * b] <= 0.f ? 1e6f : norm[2 * b];   // keypoint_auc tiles norm[b,0] on both axes
  const bool auc_ok = norm[2 * b] != 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    if (!valid[(long long)b * K + k]) continue;
    const double ex = (double)pred[((long long)b * K + k) * 2] - (double)gt[((long long)b * K + k) * 2];
    const double ey = (double)pred[((long long)b * K + k) * 2 + 1] - (double)gt[((long long)b * K + k) * 2 + 1];
    const float d_e = (float)sqrt(ex * ex + ey * ey);
    atomicAdd(&nvalid, 1);
    atomicAdd(&sum_e, (double)d_e);
    if (norm_ok) {
      const double dx = ex / (double)nx, dy = ey / (double)ny;
      const float d_n = (float)sqrt(dx * dx + dy * dy);
      atomicAdd(&nvalid_n, 1);
      atomicAdd(&sum_n, (double)d_n);
      for (int t = 0; t < T; ++t)
        if ((double)d_n < (double)thr[t]) atomicAdd(&hits[t], 1);
    }
    if (auc_ok) {
      const double dx = ex / (double)ax, dy = ey / (double)ax;
      const float d_a = (float)sqrt(dx * dx + dy * dy);
      for (int i = 0; i < auc_steps; ++i)
        if ((double)d_a < (double)i / (double)auc_steps) atomicAdd(&auc_hits[i], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < T) atomicAdd(&counters[threadIdx.x], nvalid_n > 0 ? (double)hits[threadIdx.x] / (double)nvalid_n : 0.0);
  if (threadIdx.x == 0) {
    atomicAdd(&counters[T], sum_n / (double)max(1, nvalid_n));
    double auc = 0.0;
    if (auc_ok && nvalid > 0)
      for (int i = 0; i < auc_steps; ++i) auc += (1.0 / auc_steps) * ((double)auc_hits[i] / (double)nvalid);
    atomicAdd(&counters[T + 1], auc);
    atomicAdd(&counters[T + 2], sum_e / (double)max(1, nvalid));
    atomicAdd(&counters[T + 3], 1.0);
  }
}

// preds[b,k,:] = (transform_preds(points[b,k] * [W,H], center[b], scale[b], [W,H]), 1): the heat-map-space ->
// image-space affine of TwoStageHead.decode, cs = [cx, cy, sx, sy] per sample
__global__ void decode_preds_kernel(const float* __restrict__ points, const float* __restrict__ cs,
                                    float* __restrict__ preds, int K, float W, float H, int use_udp, int total) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int b = i / K;
  const float cx = cs[4 * b], cy = cs[4 * b + 1], sx = cs[4 * b + 2] * 200.0f, sy = cs[4 * b + 3] * 200.0f;
  const float kx = use_udp ? sx / (W - 1.0f) : sx / W, ky = use_udp ? sy / (H - 1.0f) : sy / H;
  preds[3 * i + 0] = points[2 * i] * W * kx + cx - sx * 0.5f;
  preds[3 * i + 1] = points[2 * i + 1] * H * ky + cy - sy * 0.5f;
  preds[3 * i + 2] = 1.0f;
}

}  // namespace ec

using namespace ec;

extern "C" int ec_decode_preds(const float* points, const float* center_scale, float* preds, int B, int K, float W,
                               float H, int use_udp, void* stream) {
  EC_REQUIRE(points && center_scale && preds, "ec_decode_preds: null pointer");
  if (B * K == 0) return EC_OK;
  launch_pdl(decode_preds_kernel, dim3(cdiv((long long)B * K, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, points, center_scale, preds, K, W, H,
                                                                                     use_udp, B * K);
  return check_launch("ec_decode_preds");
}

extern "C" int ec_support_weights(const float* target, const float* rowscale, float* Tw, int ldtw, int BK,
                                  int hm_h, int hm_w, int h, int w, void* stream) {
  EC_REQUIRE(target && Tw, "ec_support_weights: null pointer");
  EC_REQUIRE(ldtw >= h * w, "ec_support_weights: ldtw too small");
  EC_REQUIRE(h <= 128 && w <= 128, "ec_support_weights: feature grid larger than 128 x 128");
  if (BK == 0) return EC_OK;
  size_t smem = sizeof(float) * ((size_t)hm_h * hm_w + (size_t)hm_w * w + (size_t)hm_h * h + (size_t)hm_h * w);
  EC_REQUIRE(smem <= 200 * 1024, "ec_support_weights: heat-map / grid too large for shared memory");
  if (smem > 48 * 1024)
    EC_CUDA(cudaFuncSetAttribute(support_weights_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  launch_pdl(support_weights_kernel, dim3(BK), dim3(256), (size_t)(smem), (cudaStream_t)stream, target, rowscale, Tw, ldtw, hm_h, hm_w, h, w);
  return check_launch("ec_support_weights");
}

extern "C" int ec_sine_pe_coords(const float* coord, float* out, int ldo, int M, int num_feats,
                                 float temperature, float scale, void* stream) {
  EC_REQUIRE(coord && out && ldo >= 2 * num_feats, "ec_sine_pe_coords: bad arguments");
  long long total = (long long)M * 2 * num_feats;
  if (total == 0) return EC_OK;
  launch_pdl(sine_pe_kernel, dim3(cdiv(total, 256)), dim3(256), (size_t)(0), (cudaStream_t)stream, coord, out, ldo, M, num_feats, temperature,
                                                                      scale);
  return check_launch("ec_sine_pe_coords");
}

extern "C" int ec_proposal(const float* sim, float* prop_loss, float* prop, int64_t* argmax, int BK, int h, int w,
                           void* stream) {
  EC_REQUIRE(sim && prop_loss && prop && argmax, "ec_proposal: null pointer");
  EC_REQUIRE(h > 0 && w > 0, "ec_proposal: empty map");
  if (BK == 0) return EC_OK;
  launch_pdl(proposal_kernel, dim3(cdiv(BK, 8)), dim3(256), (size_t)(0), (cudaStream_t)stream, sim, prop_loss, prop, (long long*)argmax, BK, h, w);
  return check_launch("ec_proposal");
}

extern "C" int ec_pck_accumulate(const float* pred, const float* gt, const uint8_t* valid, const float* norm,
                                 const float* thr, int T, double* counters, int B, int K, void* stream) {
  EC_REQUIRE(pred && gt && valid && norm && thr && counters, "ec_pck_accumulate: null pointer");
  EC_REQUIRE(T >= 1 && T <= 16, "ec_pck_accumulate: 1..16 thresholds");
  if (B == 0) return EC_OK;
  pck_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(pred, gt, valid, norm, thr, T, counters, K);
  return check_launch("ec_pck_accumulate");
}

extern "C" int ec_metrics_accumulate(const float* pred, const float* gt, const uint8_t* valid, const float* norm,
                                     const float* thr, int T, int auc_steps, double* counters, int B, int K,
                                     void* stream) {
  if (B == 0) return EC_OK;                       // an empty shard: nothing to add (its tensors may be null)
  EC_REQUIRE(pred && gt && valid && norm && thr && counters, "ec_metrics_accumulate: null pointer");
  EC_REQUIRE(T >= 1 && T <= 16 && auc_steps >= 1 && auc_steps <= 64, "ec_metrics_accumulate: 1..16 thresholds, 1..64 AUC steps");
  if (B == 0) return EC_OK;
  metrics_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(pred, gt, valid, norm, thr, T, auc_steps, counters, K);
  return check_launch("ec_metrics_accumulate");
}
