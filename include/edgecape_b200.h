/*
 * edgecape_b200 -- C ABI of the B200-native EdgeCape inference hot path.
 *
 * This header is the drop-in boundary below the reference's Python operator layer
 * (mmpose registry classes EdgeCape / TwoStageHead / SkeletonPredictor /
 * TwoStageSupportRefineTransformer).  The reference has no native code, so every entry
 * point replaces a block of PyTorch-eager calls; the replaced reference lines are cited
 * per function (paths relative to /root/reference/EdgeCape/models/).  INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions:
 *   - every function returns 0 on success or a negative EC_ERR_* code; ec_last_error_string()
 *     describes the last failure on the calling thread.  Nothing throws.
 *   - nothing allocates and nothing synchronises: all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*); the caller owns buffers and synchronisation.
 *   - all tensors are device pointers to dense row-major fp32 unless stated; `ld*` are row
 *     strides in elements.  Masks are uint8 (1 = masked / padded), indices int32/int64.
 *   - the device is the caller's current CUDA device; per-device state (tile counters, function
 *     attributes) is created under a mutex on the first call made with a device current, so
 *     several devices per process and concurrent host threads are fine.
 */
#ifndef EDGECAPE_B200_H
#define EDGECAPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EC_OK 0
#define EC_ERR_INVALID -1      /* bad argument (shape, alignment, null pointer) */
#define EC_ERR_CUDA -2         /* a CUDA runtime/driver call failed */
#define EC_ERR_UNSUPPORTED -3  /* shape outside what the kernels are instantiated for */

/* activation codes for ec_gemm / ec_gemm_f16x3 epilogues */
#define EC_ACT_NONE 0
#define EC_ACT_RELU 1
#define EC_ACT_GELU 2 /* exact erf GELU (torch.nn.GELU default) */
#define EC_ACT_TANH 3

/* residual modes */
#define EC_RES_NONE 0
#define EC_RES_ADD 1  /* y = R + colscale * act(acc + bias) */
#define EC_RES_GATE 2 /* y = (act(acc + bias) + 1) * R   (ProposalGenerator gate) */

int ec_version(void);
const char* ec_last_error_string(void);
/* number of kernel launches enqueued by this library in this process so far */
long long ec_launch_count(void);

/* ---------------------------------------------------------------- dense contractions
 * C[b] = epilogue(A[b] (MxK, row stride lda) * op(B[b])) for b < batch.
 * b_kmajor = 1: B is [N,K] row-major (an nn.Linear / 1x1-conv weight), C = A * B^T.
 * b_kmajor = 0: B is [K,N] row-major, C = A * B.
 * epilogue: y = act(acc + bias[n]); y *= colscale[n]; then the residual mode with R (row stride
 * ldr, batch stride strideR; strideR = 0 broadcasts R over the batch).  C may alias R.
 * Replaces every nn.Linear / Conv1d(k=1) / Conv2d(1x1) / torch.bmm on the path, e.g.
 * keypoint_heads/head.py:169,188, encoder_decoder.py:56-63,471-482,508-524, skeleton.py:92.
 * fp32 SIMT kernel (FFMA); exact fp32 arithmetic. */
int ec_gemm(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc,
            int b_kmajor, int batch, long long strideA, long long strideB, long long strideC,
            const float* bias, int act, const float* colscale, const float* R, int ldr,
            long long strideR, int res_mode, void* stream);

/* ------------------------------------------------- tensor-core contraction (tcgen05, sm_100a)
 * C = epilogue(out_scale * A * B^T) with the epilogue of ec_gemm (b_kmajor = 1, batch = 1),
 * computed on the 5th-gen tensor cores with fp32-grade accuracy: both operands are pre-split
 * into fp16 (hi, lo) pairs, A2 = [M, 2*Kp] and B2 = [N, 2*Kp] halves laid out [hi | lo] per row
 * (Kp = K rounded up to a multiple of 64, zero padded), and the kernel accumulates
 * lo*hi + hi*lo + hi*hi into one fp32 TMEM accumulator (TMA-fed, 128B-swizzled shared-memory
 * tiles, 128x128x16 UMMA).  ec_split_f16 produces the split layout from fp32 rows
 * (row m at X + (m / seg) * seg_stride + (m % seg) * ldx; seg = 0: m * ldx), multiplying by
 * `scale` first (a power of two chosen per weight tensor keeps small weights out of the fp16
 * subnormal range; out_scale undoes it).  If split_out != NULL the epilogue also stores the
 * split form of the result, [M, 2*split_kp], ready to be the next GEMM's A operand.  Row m of C
 * lives at C + (m / seg_c) * seg_stride_c + (m % seg_c) * ldc (seg_c = 0: m * ldc); res_rows > 0 reads
 * residual row m % res_rows (a [res_rows, N] table broadcast over the batch, e.g. the position embedding). */
int ec_split_f16(const float* X, void* X2, int M, int K, int ldx, int seg, long long seg_stride,
                 int Kp, float scale, void* stream);
int ec_gemm_f16x3(const void* A2, const void* B2, float* C, int M, int N, int Kp, int ldc,
                  int seg_c, long long seg_stride_c, float out_scale, const float* bias, int act, const float* colscale,
                  const float* R, int ldr, int res_mode, int res_rows, void* split_out, int split_kp,
                  float split_scale, void* stream);

/* The same contraction in 2 instead of 3 units of tensor time: a_hi.b_hi on fp16 UMMAs, the two cross terms
 * a_lo.b_hi + a_hi.b_lo -- which only need ~11 bits of relative accuracy -- on e4m3 UMMAs (kind::f8f6f4, twice the
 * contraction depth per clock).  Operand rows are EC_SPLIT_F16F8: [hi16 : Kp halves | per 64 columns:
 * hi8 x 64, lo8 x 64] (hi8 of column c at byte 2 Kp + 128 (c / 64) + c % 64, lo8 64 bytes behind it; Kp % 64 == 0), 4*Kp
 * bytes as EC_SPLIT_F16X2's [hi16 | lo16] -- so both e4m3 planes of a 64-deep k-block arrive in ONE TMA box of
 * 128-byte rows.  The plane scales are static powers of two --
 *   A role (activations): hi8 = e4m3(hi16),          lo8 = e4m3((a - hi16) * 2^11)
 *   B role (weights * s_w): hi8 = e4m3(hi16 * 2^-11), lo8 = e4m3(b * s_w - hi16)
 * -- so every product carries s_w and all three accumulate into ONE fp32 TMEM accumulator (out_scale = 1 / s_w undoes
 * it).  Measured error vs fp64 on ViT-B shapes: 1.1e-5 of max|C| (three fp16 products: 3.4e-6; bar 1e-3).
 * ec_split_f16f8 produces the rows from fp32 (role 0 = A, 1 = B, 2 = B with the three planes interleaved per 32
 * columns -- [hi16 x 32 | hi8 x 32 | lo8 x 32], 128 bytes at byte column 128 (k / 32), the weight format of ec_gcn_fused2;
 * addressing as ec_split_f16).  ec_gemm_f16f8 has
 * ec_gemm_f16x3's arguments plus split_fmt, the format of split_out (EC_SPLIT_F16X2 when an attention kernel or an
 * ec_gemm_f16x3 consumes it, EC_SPLIT_F16F8 when an ec_gemm_f16f8 does).
 * Range: e4m3 saturates at 448.  An activation beyond that only loses ITS cross terms (the element degrades to
 * plain-fp16 accuracy, ~2^-11 relative); beyond 65504 the hi16 plane overflows.  Every EC_SPLIT_F16F8 producer counts
 * both events on the device: ec_overflow_count(out2, reset) synchronises the device and returns
 * {values beyond 448, values beyond 65504} seen so far (a checkpoint with outlier activations reports itself instead
 * of degrading silently). */
#define EC_SPLIT_F16X2 0
#define EC_SPLIT_F16F8 1
int ec_split_f16f8(const float* X, void* out, int M, int K, int ldx, int seg, long long seg_stride, int Kp,
                   float scale, int role, void* stream);
int ec_gemm_f16f8(const void* A3, const void* B3, float* C, int M, int N, int Kp, int ldc,
                  int seg_c, long long seg_stride_c, float out_scale, const float* bias, int act, const float* colscale,
                  const float* R, int ldr, int res_mode, int res_rows, void* split_out, int split_kp,
                  float split_scale, int split_fmt, void* stream);
int ec_overflow_count(unsigned long long* out2, int reset);

/* tuning knob for ec_gemm_f16x3: 0 = pick the tile width per shape (128x256 tiles for wide, large
 * problems, 128x128 otherwise), 128 / 256 = force it. */
int ec_tc_set_tile_n(int bn);
/* number of ec_gemm_f16x3 launches so far in this process that took the tile mode `mode`: 128 (128x128, one CTA),
 * 256 (128x256, one CTA) or 512 (256x256 on a CTA pair, cta_group::2); -1 for any other argument.  Tests use the
 * deltas to prove which kernel instance a given configuration runs.  mode + 1 (129 / 257 / 513) counts the same tile
 * mode of ec_gemm_f16f8. */
long long ec_tc_mode_launches(int mode);
/* epilogue of the GEMMs whose only output is split_out (default 1): 1 = the CTA assembles 128-row x 64-column blocks of
 * the split rows in shared memory and writes them with TMA bulk tensor stores; 2 = every lane converts the 16 columns of
 * its own accumulator row and stores them straight from registers (no shared-memory traffic); 0 = the transposing
 * epilogue of the fp32 outputs (kept for A/B measurements). */
int ec_tc_set_split_tma(int on);
/* profiling: buf = device array [148][4] of int64, or NULL (default).  Every leader CTA's MMA thread then writes {total
 * clocks, clocks waiting for a free accumulator (= for the epilogue), clocks waiting for operands (= for TMA), tiles}. */
int ec_tc_set_trace(void* buf);
/* cap on the CTAs of the persistent GEMM grids (0 = one per SM).  With consecutive batches pipelined (backbone of
 * batch i+1 beside the head of batch i) a cap below the SM count leaves SMs to the other stream's small kernels. */
int ec_tc_set_cta_limit(int ctas);
/* tile scheduling of ec_gemm_f16x3's persistent CTAs: 1 (default) = tiles drawn from a per-launch global counter, so a
 * CTA whose SM is busy with another stream's kernel takes fewer tiles instead of stalling the launch; 0 = static
 * round-robin.  EDGECAPE_GEMM_DYNAMIC=0 selects static at start-up. */
int ec_tc_set_dynamic(int on);
int ec_tc_set_ksplit(int on);      /* K-split of the tail tiles of fp32-output GEMMs whose last wave is partly empty (EDGECAPE_GEMM_KSPLIT) */
long long ec_tc_ksplit_launches(void);   /* launches that used it so far (tests) */
/* programmatic dependent launch of the library's kernels (default on; EDGECAPE_PDL=0 or ec_set_pdl(0) = plain
 * stream-ordered launches, for A/B measurements) */
int ec_set_pdl(int on);
/* profiling experiments on ec_gemm_f16x3 (results are WRONG when flags != 0): 1 = operands stay resident
 * (no TMA after the pipeline is primed), 2 = hi*hi product only, 4 = no epilogue stores, 8 = no epilogue,
 * 16 = no split_out stores, 32 = no GELU. */
int ec_tc_set_debug(int flags);
/* profiling aid for ec_attention_tc_split: when buf != NULL the first n_ctas CTAs of every launch write ten
 * clock64() stamps (int64) of their phases to buf[cta][10] (device memory); NULL switches it off. */
int ec_attention_tc_set_trace(void* buf, int n_ctas);
/* ec_attention_tc_split kernel variant: 0 (default) keeps the probabilities in TMEM as the A operand of the
 * P V MMAs and folds P_hi V_hi, P_hi V_lo into one N = 128 MMA; 2 = P in TMEM, three N = 64 MMAs per k-step;
 * 1 = probabilities through shared memory.  1 and 2 are kept for A/B measurements. */
int ec_attention_tc_set_variant(int variant);

/* ------------------------------------------------------------------------- normalisation
 * Y[m,:] = LayerNorm(X[m,:] (+ R[m,:])) * w + b  (biased variance, eps inside the sqrt).
 * Row m of X lives at X + (m / seg) * seg_stride + (m % seg) * ldx  (seg = 0: plain m * ldx) so a
 * ViT token matrix can be read with its cls row dropped.  If sum_out != NULL the pre-norm sum
 * X+R is stored there (row stride ld_sum).  torch.nn.LayerNorm on the path:
 * encoder_decoder.py:477-482,612-637, ViT norm1/norm2/norm (DINOv2, eps 1e-6).
 * split_out (optional, [M, 4*split_kp bytes]) receives the split form of the result for a following
 * tensor-core GEMM, in the row format split_fmt (EC_SPLIT_F16X2 for ec_gemm_f16x3, EC_SPLIT_F16F8 for
 * ec_gemm_f16f8); Y may then be NULL. */
int ec_layernorm(const float* X, int ldx, int seg, long long seg_stride, const float* R, int ldr,
                 float* sum_out, int ld_sum, float* Y, int ldy, const float* w, const float* b,
                 float eps, int M, int C, void* split_out, int split_kp, int split_fmt, void* stream);

/* X[b, t, :] += P[t, :] for t < S (rows S..T-1 untouched): the encoder adds the grid positional
 * encoding to the residual stream every layer (encoder_decoder.py:467). */
int ec_add_rows(float* X, const float* P, int batch, int T, int S, int C, void* stream);

/* Y[row(m), 0:C] = X[row(m % bcast_rows if bcast_rows else m), 0:C] with independent strides:
 * row(m) = (m / seg) * seg_stride + (m % seg) * ld  (seg = 0: m * ld).  Concatenation / slicing
 * helper (torch.cat at encoder_decoder.py:198-203,620-622; the token split at :307-308). */
int ec_copy_rows(const float* X, int ldx, int seg_x, long long seg_stride_x, float* Y, int ldy,
                 int seg_y, long long seg_stride_y, int M, int C, int bcast_rows, void* stream);
/* dst[r, 0:block_elems] = src[idx[r], 0:block_elems] for r < n_out (strides in floats between blocks).  Expands the ViT
 * features of de-duplicated support images to one block per batch row (test_dataset.py:93-97: the queries of an episode
 * share their support sample; the reference runs the backbone on every copy). */
int ec_gather_blocks(const float* src, long long src_stride, const int32_t* idx, float* dst, long long dst_stride,
                     int n_out, long long block_elems, void* stream);

/* out = (a*x + b*y) / div elementwise: the mean over shots (head.py:186, skeleton.py:113). */
int ec_axpby(const float* x, const float* y, float* out, float a, float b, float div, long long n,
             void* stream);

/* ----------------------------------------------------------------------------- attention
 * O[b, i, h*DV:(h+1)*DV] = softmax_j( scale * <Q[b,i,h], K[b,j,h]> + bias[b,h,i,j] , masked ) V[b,j,h]
 * Q/K/V are [B, L, H*D] views with row strides ldq/ldk/ldv and batch strides in elements
 * (they may point into one packed QKV buffer).  key_mask: uint8 [B, Lk], 1 = ignore key
 * (key_padding_mask), may be NULL.  bias: fp32 [B, H, Lq, Lk] or NULL.  D in {16, 32, 64}.
 * Replaces F.multi_head_attention_forward / SDPA (encoder_decoder.py:471-477, 606-631,
 * 638-649), BiasedMultiheadAttention's bmm/softmax/bmm (utils/bias_attn.py:176-222) and the
 * DINOv2 block attention.  split_out (optional, fp16 [B*Lq, 2*split_kp], split_kp == H*D)
 * receives the split-fp16 form of O for a following ec_gemm_f16x3; O may then be NULL. */
int ec_attention(const float* Q, const float* K, const float* V, float* O, int B, int H, int Lq,
                 int Lk, int D, int ldq, int ldk, int ldv, int ldo, long long sq, long long sk,
                 long long sv, long long so, float scale, const uint8_t* key_mask,
                 const float* bias, void* split_out, int split_kp, void* stream);

/* Tensor-core attention (tcgen05, sm_100a): same contract as ec_attention (key_mask / bias included),
 * head dim 64 or 32 (tiles zero padded to 64), Lk <= 448 (the whole [128 x Lk] score block lives in TMEM).  Q, K, V, P are split into
 * fp16 (hi, lo) pairs in shared memory and both contractions run as 3-product UMMAs with fp32 TMEM
 * accumulation (fp32-grade results).  Used for the DINOv2 block attention and the 8 x 64 cross
 * attentions (encoder_decoder.py:620-631, 638-649). */
int ec_attention_tc(const float* Q, const float* K, const float* V, float* O, int B, int H, int Lq,
                    int Lk, int D, int ldq, int ldk, int ldv, int ldo, long long sq, long long sk,
                    long long sv, long long so, float scale, const uint8_t* key_mask,
                    const float* bias, void* split_out, int split_kp, void* stream);

/* TMA-fed form of ec_attention_tc: Q, K, V are already split-fp16 buffers [rows, 2*kp] (hi | lo halves),
 * e.g. the split output of the QKV GEMM (one buffer, q_col = 0, k_col = C, v_col = 2C).  Head h of Q lives in
 * columns q_col + dv h of each half and rows b * q_rows + i; K / V likewise with k_rows rows per batch
 * element.  Nothing is staged by hand: 64 x 64 boxes are loaded by TMA, V is consumed as an MN-major UMMA
 * operand, the probabilities stay in TMEM.  dv = 64 handles up to 768 keys; dv = 32, key_mask ([B, Lk], 1 =
 * ignore) and bias ([B, H, Lq, Lk], added to the scaled logits) need Lk <= 448.  Same output contract as
 * ec_attention_tc (O fp32 and / or split_out with split_kp == H * dv).  Every attention of the path goes
 * through this entry in tensor-core mode (models/.../encoder_decoder.py:461-483, 584-651; DINOv2 blocks). */
/* Structural attention bias computed INSIDE the attention kernel (BiasedMultiheadAttention, utils/bias_attn.py:188-191;
 * SURVEY K11): arms the NEXT ec_attention_tc_split call made by this host thread -- and only that one -- to add
 *   bias[b,h,i,j] = W1[h,:] . relu(W0 . hops[:,b,i,j] + b0) + b1[h]
 * to its scaled logits, from the Markov hop tensor hops [n_hops, B, Lq, Lk] (ec_edge_weights + ec_markov_powers) and the
 * markov_structural_mlp weights W0 [hidden, n_hops], b0 [hidden], W1 [H, hidden], b1 [H]; that call must pass
 * bias = NULL.  The row-max pass evaluates the MLP per logit and writes the finished logits back over S in TMEM, so
 * the [B,H,Lq,Lk] tensor ec_hop_bias writes (and the attention kernel reads back, twice) never exists.
 * Limits: n_hops <= 8, hidden * (n_hops + 2) < 112 (else use ec_hop_bias + the bias argument). */
int ec_attention_hop_bias_next(const float* hops, int n_hops, int hidden, const float* w0, const float* b0,
                               const float* w1, const float* b1);
/* Arms the NEXT ec_attention_tc_split call on this thread to write its split_out rows in the format `fmt`:
 * EC_SPLIT_F16X2 (the default) or EC_SPLIT_F16F8 (A role; head dim 64 only) -- the format ec_gemm_f16f8 consumes, so that
 * the projection behind a ViT attention (dino.py Attention.proj) runs its cross terms on e4m3 like qkv / fc1 / fc2. */
int ec_attention_split_fmt_next(int fmt);
/* Experiment switch (like ec_tc_set_cta_limit / ec_gcn_fused2_set_cta_limit): the persistent attention kernel launches at most
 * `ctas` CTAs (0 = one per SM).  Grid sizes are fixed when a CUDA graph is captured: the detector's EDGECAPE_HEAD_CTAS narrows the
 * head graph's persistent kernels so that they run beside the next batch's backbone instead of in front of it. */
int ec_attention_set_cta_limit(int ctas);
int ec_attention_tc_split(const void* Q2, int q_total_rows, int q_kp, int q_col, int q_rows,
                          const void* K2, int k_total_rows, int k_kp, int k_col, const void* V2,
                          int v_total_rows, int v_kp, int v_col, int k_rows, float* O, int B, int H,
                          int Lq, int Lk, int ldo, long long so, float scale, int dv,
                          const uint8_t* key_mask, const float* bias, void* split_out, int split_kp,
                          void* stream);

/* bias[b,h,i,j] = W1 relu(W0 hops[:,b,i,j] + b0) + b1 with hops = attn_adj [n_hops, B, K, K]:
 * the Graphormer-style structural bias MLP (utils/bias_attn.py:82-83,188-191). */
int ec_hop_bias(const float* attn_adj, const float* w0, const float* b0, const float* w1,
                const float* b1, float* bias, int B, int K, int n_hops, int hidden, int H,
                void* stream);

/* ------------------------------------------------------------------------------ skeleton
 * mask bookkeeping (detectors/EdgeCape.py:175-177, head.py:189, encoder_decoder.py:359-360):
 * mask_s[b,k] = (first ? tw*tw : mask_s*tw); ec_kp_masks derives kp_mask = (mask_s == 0) and
 * kp_mask_fixed (column 0 un-masked for rows whose keypoints are all masked). */
int ec_mask_accumulate(const float* tw, float* mask_s, int n, int first, void* stream);
int ec_kp_masks(const float* mask_s, uint8_t* kp_mask, uint8_t* kp_mask_fixed, int B, int K,
                void* stream);

/* adjacency from edge lists (keypoint_heads/skeleton.py:171-194): edges int32 [E_total,2],
 * offsets int32 [B+1] (CSR).  adj [B,2,K,K] = [diag(valid), nan_to_num(A / rowsum)],
 * binary [B,K,K] = (masked symmetric 0/1 adjacency).  Edges with an index outside [0,K) are
 * an error in the reference (IndexError); here they are ignored. */
int ec_adj_from_edges(const int32_t* edges, const int32_t* offsets, const uint8_t* kp_mask,
                      float* adj, float* binary, int B, int K, void* stream);

/* soft_normalize_adj (skeleton.py:196-205): adj [B,2,K,K] = [diag(valid), (U*vv)/(rowsum+1e-8)]. */
int ec_soft_normalize_adj(const float* U, const uint8_t* kp_mask, float* adj, int B, int K,
                          void* stream);

/* edge-weight prediction (skeleton.py:134-150) from the Gram matrix S = Fn Fn^T of the
 * L2-normalised refined keypoint tokens (ec_l2_normalize + ec_gemm):
 * U = relu(binary + w*(S+S^T)/2 + b); adj = soft_normalize(U); unnorm = U * valid x valid;
 * hops[0..n_hops-1] (skeleton.py:152-161) get P^0 = I and P^1 = adj1/(rowsum+1e-8); higher
 * powers: ec_markov_powers. */
int ec_l2_normalize(const float* X, float* Y, int M, int C, float eps, void* stream);
int ec_edge_weights(const float* S, const float* binary, const uint8_t* kp_mask, float zc_w,
                    float zc_b, int use_zero_conv, float* adj, float* unnorm, float* hop0,
                    float* hop1, int B, int K, void* stream);
/* markov_transition_matrix's higher powers (skeleton.py:152-161: torch.matrix_power(P, h), h = 2 .. max_hop) in ONE
 * launch: hops is [max_hop + 1, B, K, K] with planes 0 and 1 filled by ec_edge_weights; plane h = plane (h-1) . plane 1,
 * row i of a power only needs row i of the one before, so a warp walks 8 rows through every power on its own (CTA =
 * sample x 32 rows, P^1 resident in shared memory; K <= 224; exact fp32 FFMA). */
int ec_markov_powers(float* hops, int max_hop, int B, int K, void* stream);

/* GCN feed-forward (encoder_decoder.py:508-524), aggregate-first form:
 * Y[b,w,:] = relu( a0[b,w] * (X[b,w,:] W0^T + b0) + sum_v A1[b,w,v] (X[b,v,:] W1^T + b1) )
 *          = relu( Z[b,w,:] Wp^T ),  Z = [ a0*X | A1 X | a0 | rowsum(A1) | 0 0 ]  (K' = 2d+4)
 * adj [B,2,K,K]; plane 0 must be diagonal (it always is: skeleton.py:193,204 build it with
 * diag_embed), a0 is its diagonal.  W [2*dff, d] is the Conv1d(k=1) weight, bias [2*dff];
 * ec_gcn_pack_weights builds Wp [dff, 2d+4] = [W0 | W1 | b0 | b1 | 0 0] once per checkpoint.
 * X [B,K,d] -> Y [B,K,dff].  workspace holds Z (ec_workspace_bytes_gcn). */
int ec_gcn_pack_weights(const float* W, const float* bias, float* Wp, int d, int dff, void* stream);
int ec_gcn(const float* X, const float* adj, const float* Wp, float* Y, int B, int K, int d, int dff,
           float* workspace, size_t workspace_bytes, void* stream);
size_t ec_workspace_bytes_gcn(int B, int K, int d, int dff);
/* Tensor-core form of the same layer: one fused kernel writes Z directly as the split-fp16 A operand
 * Z2 [B*K, 2*Kp] (Kp = 2d+4 rounded up to 64), then ec_gemm_f16x3(Z2, split(Wp), act = ReLU) finishes it. */
int ec_gcn_aggregate_split(const float* X, const float* adj, void* Z2, int B, int K, int d, int Kp,
                           void* stream);
/* The same layer as ONE kernel (csrc/gcn_fused_tcgen05.cu): per (sample, slice of output channels) CTA the
 * aggregate A1 X and both weight products run on the tensor cores out of shared memory / TMEM; nothing but X,
 * adj and the weights is read and nothing but the result is written.  W2 [dff, 2*Kp] is the split-fp16 form of
 * Wp * w_scale (ec_split_f16; Kp = 2d+4 rounded up to 64, w_scale a power of two); the bias columns are taken
 * from the fp32 Wp.  Y [B,K,dff] and / or split_out [B*K, 2*split_kp] (hi | lo of Y) are written.
 * ec_gcn_fused_slice returns the channel-slice width the kernel would use, or 0 when it cannot take the shape
 * (it needs K <= 128, d in {64, 128, 256}, dff a multiple of 64, and the tiles to fit 227 KB). */
int ec_gcn_fused_slice(int K, int d, int dff);
int ec_gcn_fused_set_debug(int flags);               /* profiling experiments only: 1 = skip the A1 / X loads, 2 = skip the stores */
int ec_gcn_fused_set_trace(void* buf, int n_ctas);   /* profiling: [n_ctas][32] int64 clock stamps per CTA, NULL = off */
int ec_gcn_fused(const float* X, const float* adj, const float* Wp, const void* W2, int Kp, float w_scale,
                 float* Y, void* split_out, int split_kp, int B, int K, int d, int dff, void* stream);
/* Second formulation of the same layer (csrc/gcn_fused2_tcgen05.cu, the default): project first,
 *   T0 = X W0^T, T1 = X W1^T (running while X is still being read; cross terms of the fp32-grade split on e4m3
 *   tensor cores), D2 = A1 T1 with A1 in tensor memory, Y = relu(a0 (T0 + b0) + D2 + rowsum(A1) b1);
 * persistent CTAs over (sample, channel slice) items.  Arguments as ec_gcn_fused except that bias2 [2][dff] holds the
 * bias columns of Wp contiguously (b0 then b1; they are preloaded into the accumulators) and W3 [dff, 4*Kp bytes]
 * is the F16F8 B-role form of Wp * w_scale with its planes interleaved per 32 columns (ec_split_f16f8 with
 * role 2): the weights stream in 32-deep k-slices.  Shape gate ec_gcn_fused2_slice: K <= 128 (one CTA per sample and
 * channel slice), or K <= 256 and a multiple of 8 (a cluster of two CTAs, each owning half of the sample's rows: configs[4],
 * K = 200); d in {64, 128, 256}; dff a multiple of 64. */
int ec_gcn_fused2_slice(int K, int d, int dff);
int ec_gcn_fused2_set_debug(int flags);               /* profiling experiments only: 1 = skip the A1 / X loads, 2 = skip the stores */
int ec_gcn_fused2_set_trace(void* buf, int n_ctas);   /* profiling: [n_ctas][32] int64 clock stamps of each CTA's first item */
int ec_gcn_fused2_set_cta_limit(int ctas);            /* profiling: at most this many persistent CTAs (0 = one per SM) */
int ec_gcn_fused2(const float* X, const float* adj, const float* bias2, const void* W3, int Kp, float w_scale,
                  float* Y, void* split_out, int split_kp, int B, int K, int d, int dff, void* stream);

/* ----------------------------------------------------------------------------- head ops
 * support-keypoint pooling weights (head.py:175-184, exact by linearity):
 * Tw[b,k,s] = scale[b,k] / (sum(t[b,k]) + 1e-8) * sum_p t[b,k,p] U[p,s], U = bilinear
 * (align_corners=False) interpolation matrix from the h x w feature grid to the hm x hm map.
 * pooled = Tw @ feat is then one batched ec_gemm.  rowscale [B,K] (mask / shots) may be NULL. */
int ec_support_weights(const float* target, const float* rowscale, float* Tw, int ldtw, int BK,
                       int hm_h, int hm_w, int h, int w, void* stream);

/* DETR sine positional encoding of continuous coordinates (utils/positional_encoding.py:96-122):
 * coord [M,2] (x,y) in [0,1] -> out[m, :] = [PE(y) | PE(x)], 2*num_feats channels. */
int ec_sine_pe_coords(const float* coord, float* out, int ldo, int M, int num_feats,
                      float temperature, float scale, void* stream);

/* ProposalGenerator tail (encoder_decoder.py:66-112) on sim [BK, S] (S = h*w):
 * proposal_for_loss (global soft-argmax), argmax (first max), proposals (3x3 local soft-argmax
 * with the reference's (w,h) reshape). */
int ec_proposal(const float* sim, float* prop_loss, float* prop, int64_t* argmax, int BK, int h,
                int w, void* stream);

/* bi' = sigmoid(inverse_sigmoid(bi) + delta), eps = 1e-3 (encoder_decoder.py:395-403,427-431,
 * head.py:27-31,216-220).  Also used for the final per-layer decode. */
int ec_point_update(const float* bi, const float* delta, int ldd, float* out, int M, void* stream);

/* TwoStageHead.decode on the device (head.py:324-387, mmpose transform_preds): points [B,K,2] in [0,1],
 * center_scale [B,4] = (cx, cy, sx, sy) -> preds [B,K,3] = (x, y, 1) in image coordinates. */
int ec_decode_preds(const float* points, const float* center_scale, float* preds, int B, int K, float W,
                    float H, int use_udp, void* stream);

/* ------------------------------------------------------------------------------ ViT ops
 * im2col for the stride-P patch embedding (floor semantics): img [B,3,H,W] ->
 * cols [B*h0*w0, ldc] with (c,py,px) ordering = Conv2d weight flattening; columns 3*P*P..ldc-1
 * are zero-filled.  split_out (optional, fp16 [B*h0*w0, 2*split_kp]) receives the split-fp16 form for
 * ec_gemm_f16x3; cols may then be NULL. */
int ec_im2col_patches(const float* img, float* cols, int B, int H, int W, int P, int ldc,
                      void* split_out, int split_kp, void* stream);
/* bicubic (A=-0.75, align_corners=False, scale_factor=(n+offset)/M) resampling of the M x M
 * patch position table to h0 x w0 (DINOv2 interpolate_pos_encoding); pos_out [1+h0*w0, C]. */
int ec_interp_pos_embed(const float* pos_embed, float* pos_out, int Mgrid, int h0, int w0, int C,
                        double offset, void* stream);
/* tokens[b, 0, :] = cls + pos[0]  (cls row of every image) */
int ec_write_cls(const float* cls, const float* pos0, float* tokens, int B, long long stride, int C,
                 void* stream);

/* -------------------------------------------------------------------------- input stage
 * TopDownAffineFewShot + ToTensor + NormalizeTensor (datasets/pipelines/top_down_transform.py:35-67, configs/test/
 * 1shot_split1.py:101-108): src = uint8 HxWx3 image in device memory (row_stride bytes per row), M = the 2x3 double matrix
 * of get_affine_transform (HOST pointer, source -> crop), out = fp32 [3,H,W] = (warpAffine(src)/255 - mean) / std.
 * The crop reproduces cv2.warpAffine(INTER_LINEAR) on uint8 bit for bit (fixed-point coordinates and weights).
 * mean / stdv: 3 floats each, HOST pointers. */
int ec_warp_affine_normalize_u8(const uint8_t* src, int Hs, int Ws, long long row_stride, const double* M,
                                float* out, int H, int W, const float* mean, const float* stdv, void* stream);
/* TopDownGenerateTargetFewShot._msra_generate_target (top_down_transform.py:113-199, unbiased_encoding=False):
 * joints [n, ldj] (x, y in crop pixels), visible [n, ldv] -> target [n,H,W] un-normalised Gaussian patches (sigma an
 * integer), weight [n] (visibility, 0 when the patch falls outside the map). */
int ec_msra_targets(const float* joints, int ldj, const float* visible, int ldv, float* target, float* weight,
                    int n, int img_w, int img_h, int W, int H, float sigma, void* stream);

/* -------------------------------------------------------------------------- evaluation
 * per-sample PCK (mmpose keypoint_pck_accuracy semantics, datasets/.../test_base_dataset.py:
 * 104-133): pred/gt [B,K,2], valid uint8 [B,K], norm [B,2]; counters[0..T-1] += per-sample
 * PCK@thr[t]; counters[T] += number of samples.  Accumulated in fp64 for one ncclAllReduce. */
int ec_pck_accumulate(const float* pred, const float* gt, const uint8_t* valid, const float* norm,
                      const float* thr, int T, double* counters, int B, int K, void* stream);
/* every metric of `_report_metric` (test_base_dataset.py:119-154; mmpose 0.29 keypoint_pck_accuracy / keypoint_nme /
 * keypoint_auc / keypoint_epe, one sample at a time): counters[0..T-1] += PCK@thr[t]; counters[T] += NME;
 * counters[T+1] += AUC over auc_steps thresholds i/auc_steps (mmpose: 20) normalised by norm[b,0];
 * counters[T+2] += EPE (pixels); counters[T+3] += 1.  One fp64 vector for the final all-reduce. */
int ec_metrics_accumulate(const float* pred, const float* gt, const uint8_t* valid, const float* norm,
                          const float* thr, int T, int auc_steps, double* counters, int B, int K,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif
