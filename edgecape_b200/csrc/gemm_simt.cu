// fp32 SIMT GEMM with fused epilogue (ec_gemm).  Exact fp32 FFMA arithmetic: this is the
// contraction used wherever fp32 bit-level fidelity matters (similarity heat-map / argmax) and
// for the small batched products of the skeleton head; the ViT/transformer linears go through
// the tcgen05 kernel (gemm_tcgen05.cu) when the tensor-core path is enabled.
//
// Tiling: BM x BN x 16 block tile, 256 threads, (BM/16) x (BN/16) register tile per thread,
// global -> register -> shared double buffering, float4 global loads along the contiguous
// dimension, operands stored k-major in shared memory so the inner product reads are
// conflict-free float4 broadcasts.
#include "common.cuh"

namespace ec {

struct GemmParams {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  int lda, ldb, ldc;
  long long sA, sB, sC;
  const float* bias;
  const float* colscale;
  const float* R;
  int ldr;
  long long sR;
  int act, res_mode;
  int vecA, vecB, vecC, vecR;
};

constexpr int BK = 16;

// 4 consecutive elements along the contiguous dimension, zero filled out of range.
__device__ __forceinline__ float4 load4(const float* base, long long off, int c, int cmax, bool row_ok,
                                        bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!row_ok) return v;
  if (vec && c + 3 < cmax) {
    v = __ldg(reinterpret_cast<const float4*>(base + off + c));
  } else {
    if (c + 0 < cmax) v.x = __ldg(base + off + c + 0);
    if (c + 1 < cmax) v.y = __ldg(base + off + c + 1);
    if (c + 2 < cmax) v.z = __ldg(base + off + c + 2);
    if (c + 3 < cmax) v.w = __ldg(base + off + c + 3);
  }
  return v;
}

template <int BM, int BN, bool B_KMAJOR>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmParams p) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int CM = TM / 4, CN = TN / 4;       // float4 chunks per thread
  constexpr int PAD = 4;
  constexpr int A_V4 = BM * BK / 4 / 256;       // float4 loads per thread per tile
  constexpr int B_V4 = BN * BK / 4 / 256;
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int bz = blockIdx.z;
  const float* A = p.A + (long long)bz * p.sA;
  const float* B = p.B + (long long)bz * p.sB;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_V4], rb[B_V4];
  const int nk = (p.K + BK - 1) / BK;

  auto gload = [&](int kt) {
    const int k0 = kt * BK;
#pragma unroll
    for (int i = 0; i < A_V4; ++i) {
      int idx = tid + i * 256;
      int row = idx >> 2, kq = (idx & 3) * 4;
      int gm = m0 + row;
      ra[i] = load4(A, (long long)gm * p.lda, k0 + kq, p.K, gm < p.M, p.vecA);
    }
#pragma unroll
    for (int i = 0; i < B_V4; ++i) {
      int idx = tid + i * 256;
      if (B_KMAJOR) {
        int row = idx >> 2, kq = (idx & 3) * 4;
        int gn = n0 + row;
        rb[i] = load4(B, (long long)gn * p.ldb, k0 + kq, p.K, gn < p.N, p.vecB);
      } else {
        int krow = idx / (BN / 4), nq = (idx % (BN / 4)) * 4;
        int gk = k0 + krow;
        rb[i] = load4(B, (long long)gk * p.ldb, n0 + nq, p.N, gk < p.K, p.vecB);
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_V4; ++i) {
      int idx = tid + i * 256;
      int row = idx >> 2, kq = (idx & 3) * 4;
      As[buf][kq + 0][row] = ra[i].x;
      As[buf][kq + 1][row] = ra[i].y;
      As[buf][kq + 2][row] = ra[i].z;
      As[buf][kq + 3][row] = ra[i].w;
    }
#pragma unroll
    for (int i = 0; i < B_V4; ++i) {
      int idx = tid + i * 256;
      if (B_KMAJOR) {
        int row = idx >> 2, kq = (idx & 3) * 4;
        Bs[buf][kq + 0][row] = rb[i].x;
        Bs[buf][kq + 1][row] = rb[i].y;
        Bs[buf][kq + 2][row] = rb[i].z;
        Bs[buf][kq + 3][row] = rb[i].w;
      } else {
        int krow = idx / (BN / 4), nq = (idx % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Bs[buf][krow][nq]) = rb[i];
      }
    }
  };

  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int c = 0; c < CM; ++c) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][k][c * (BM / CM) + ty * 4]);
        a[c * 4 + 0] = v.x; a[c * 4 + 1] = v.y; a[c * 4 + 2] = v.z; a[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < CN; ++c) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][c * (BN / CN) + tx * 4]);
        b[c * 4 + 0] = v.x; b[c * 4 + 1] = v.y; b[c * 4 + 2] = v.z; b[c * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);   // the other buffer was last read before the previous barrier
      __syncthreads();
    }
  }

  // ---- epilogue
  float* C = p.C + (long long)bz * p.sC;
  const float* R = p.R ? p.R + (long long)bz * p.sR : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + (i / 4) * (BM / CM) + ty * 4 + (i & 3);
    if (gm >= p.M) continue;
#pragma unroll
    for (int c = 0; c < CN; ++c) {
      const int gn = n0 + c * (BN / CN) + tx * 4;
      if (gn >= p.N) continue;
      float y[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[i][c * 4 + j];
        int n = gn + j;
        if (n < p.N) {
          if (p.bias) v += __ldg(p.bias + n);
          v = apply_act(v, p.act);
          if (p.colscale) v *= __ldg(p.colscale + n);
        }
        y[j] = v;
      }
      const bool full = gn + 3 < p.N;
      if (R) {
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        const float* rp = R + (long long)gm * p.ldr + gn;
        if (full && p.vecR) {
          float4 v = *reinterpret_cast<const float4*>(rp);
          r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (gn + j < p.N) r[j] = rp[j];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = (p.res_mode == EC_RES_GATE) ? (y[j] + 1.0f) * r[j] : r[j] + y[j];
      }
      float* cp = C + (long long)gm * p.ldc + gn;
      if (full && p.vecC) {
        *reinterpret_cast<float4*>(cp) = make_float4(y[0], y[1], y[2], y[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < p.N) cp[j] = y[j];
      }
    }
  }
}

template <int BM, int BN>
static int launch_gemm(const GemmParams& p, int b_kmajor, int batch, cudaStream_t st) {
  dim3 grid(cdiv(p.N, BN), cdiv(p.M, BM), batch);
  if (b_kmajor)
    launch_pdl(gemm_simt_kernel<BM, BN, true>, dim3(grid), dim3(256), (size_t)(0), st, p);
  else
    launch_pdl(gemm_simt_kernel<BM, BN, false>, dim3(grid), dim3(256), (size_t)(0), st, p);
  return check_launch("ec_gemm");
}

}  // namespace ec

extern "C" int ec_gemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb,
                       int ldc, int b_kmajor, int batch, long long strideA, long long strideB,
                       long long strideC, const float* bias, int act, const float* colscale,
                       const float* R, int ldr, long long strideR, int res_mode, void* stream) {
  using namespace ec;
  EC_REQUIRE(A && B && C, "ec_gemm: null operand");
  EC_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "ec_gemm: negative size");
  EC_REQUIRE(act >= EC_ACT_NONE && act <= EC_ACT_TANH, "ec_gemm: bad activation %d", act);
  EC_REQUIRE(res_mode >= EC_RES_NONE && res_mode <= EC_RES_GATE, "ec_gemm: bad residual mode %d", res_mode);
  EC_REQUIRE((res_mode == EC_RES_NONE) == (R == nullptr), "ec_gemm: residual pointer/mode mismatch");
  EC_REQUIRE(lda >= K && ldc >= N && ldb >= (b_kmajor ? K : N), "ec_gemm: leading dimension too small");
  if (M == 0 || N == 0 || batch == 0) return EC_OK;
  EC_REQUIRE(batch <= 65535, "ec_gemm: batch %d > 65535", batch);
  GemmParams p;
  p.A = A; p.B = B; p.C = C;
  p.M = M; p.N = N; p.K = K;
  p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  p.sA = strideA; p.sB = strideB; p.sC = strideC;
  p.bias = bias; p.colscale = colscale; p.R = R; p.ldr = ldr; p.sR = strideR;
  p.act = act; p.res_mode = res_mode;
  p.vecA = aligned16(A) && (lda % 4 == 0) && (strideA % 4 == 0);
  p.vecB = aligned16(B) && (ldb % 4 == 0) && (strideB % 4 == 0);
  p.vecC = aligned16(C) && (ldc % 4 == 0) && (strideC % 4 == 0);
  p.vecR = R && aligned16(R) && (ldr % 4 == 0) && (strideR % 4 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  long long tiles128 = (long long)cdiv(M, 128) * cdiv(N, 128) * batch;
  if (tiles128 >= 148 * 2) return launch_gemm<128, 128>(p, b_kmajor, batch, st);
  return launch_gemm<64, 64>(p, b_kmajor, batch, st);
}
