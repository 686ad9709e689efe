/*
 * pose_b200.h -- C ABI of libpose_b200.so: the B200 (sm_100a) implementation of the per-sample
 * training / inference hot path of AliEmreSenel/3DHumanPoseEstimation.
 *
 * The reference has no FFI of its own: its boundary is a Python module API.  Each
 * entry point below is what the body of one reference callable binds to; the reference symbol it
 * replaces is cited as file:line into the reference tree.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - plain pointers and sizes only; every DEVICE pointer is owned by the caller (the library never
 *     allocates, frees or retains device memory); HOST pointers are marked [host].
 *   - all device work is enqueued asynchronously on `stream` (a CUstream / cudaStream_t); no hidden
 *     synchronisation.
 *   - return value: 0 = ok; > 0 = cudaError_t from a launch; < 0 = POSE_E_* argument error.
 *     pose_b200_error_string() maps any of them to text.  No exceptions cross the ABI.
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef POSE_B200_H
#define POSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *pose_stream_t; /* cudaStream_t */

enum {
    POSE_OK = 0,
    POSE_E_NULL = -1,        /* required pointer is NULL */
    POSE_E_SHAPE = -2,       /* unsupported / inconsistent shape */
    POSE_E_WORKSPACE = -3,   /* workspace too small */
    POSE_E_UNSUPPORTED = -4, /* valid request the sm_100a kernels do not cover (documented per call) */
    POSE_E_ALIGN = -5        /* pointer not aligned as required */
};

int pose_b200_abi_version(void);
const char *pose_b200_error_string(int code);

/* ---------------------------------------------------------------------------------------------
 * F. ComprehensivePoseLoss.forward                      reference: src/loss.py:57-85 (+ :29-55)
 *    One fused forward + backward pass.
 *    pred, gt   [B, J, 3] fp32 contiguous
 *    weights    [host] {mse_weight, l1_weight, inter_joint_loss_weight, abs_root_loss_weight}
 *    out5       [5] fp32: {mse, l1, inter_joint, abs_root, total}   (the reference's dict values)
 *    grad       [B, J, 3] fp32 or NULL: grad_scale * d(total)/d(pred)
 *    workspace  pose_loss_workspace_bytes(B, J) bytes, zero-filled before its FIRST use; the
 *               kernel leaves it zero-filled again, so it can be reused without clearing.
 * ------------------------------------------------------------------------------------------- */
size_t pose_loss_workspace_bytes(int B, int J);
int pose_loss_fwd_bwd(const float *pred, const float *gt, int B, int J, const float *weights, float *out5,
                      float *grad, float grad_scale, void *workspace, size_t workspace_bytes,
                      pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * B. GaussianHeatmapGenerator.forward                    reference: src/models/common.py:23-51
 *    kp         [B, J, 2] fp32 normalised key-points
 *    out        out_layout 0: [B, J, hs, hs] planes (the reference's layout)
 *               out_layout 1: [B, hs, hs, c_stride] channels-last, plane j at channel c_offset + j
 *                             (the CNN's 21-channel conv1 operand: src/models/cnn.py:644-648)
 *    out_dtype  0 = fp32, 1 = bf16
 * ------------------------------------------------------------------------------------------- */
int pose_heatmap_render(const float *kp, int B, int J, int hs, float sigma, void *out, int out_dtype,
                        int out_layout, int c_stride, int c_offset, pose_stream_t stream);
/* The same heat-maps rendered straight into the bf16 operand of the ViT's heat-map patch embedding
 * (src/models/transformers.py:41-46, :348-350): out [B * (hs/P)^2, J * P * P], row = patch, column = (j, ky, kx) = the
 * flattened Conv2d(J, E, P, P) weight; the fp32 planes are never written.  P % 8 == 0, hs % P == 0. */
int pose_heatmap_patchify_bf16(const float *kp, int B, int J, int hs, float sigma, int P, void *out, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * A. PoseAugmentor.__call__ (batched)            reference: src/dataset/augmentation.py:182-351
 *    The random draws are made by the caller in the reference's order (np.random: flip, angle,
 *    scale, tx, ty, brightness, contrast) and handed over explicitly; the kernels are RNG-free.
 *
 *    params     [host] [B, 8] fp64: {flip (0/1, i.e. random() < flip_prob), angle_deg, scale,
 *               tx_fraction, ty_fraction, brightness, contrast, unused}
 *    flags      bit mask of POSE_AUG_* = the reference's enable_* constructor switches
 *
 *    pose_augment_plan   [host-only, no GPU work] derives, per sample, everything that needs libm or
 *                        decimal rounding (PIL's round(cos, 15) matrix, R_y, output size) and the
 *                        launch geometry.  plan_out: [host] B * POSE_AUG_PLAN_BYTES, to be copied to
 *                        the device by the caller.  launch_out: [host] pose_aug_launch.
 *    pose_augment_batch  [device] quantise+pack -> per-sample tables -> fused cluster kernel.
 *      image     [B,3,H,W], depth [B,1,H,W]; in_dtype 0 = fp32 in [0,1] (the reference's sample
 *                schema, chunked_dataset.py:219-231), 1 = uint8 decoded pixels p, treated exactly as the
 *                reference treats the fp32 sample p/255 (its fp32 round trip (p/255)*255 -> byte is reproduced)
 *      kp [B,J,2] fp32, joints [B,J,3] fp32, cam [B,4] fp64 {fx,fy,cx,cy}
 *      plan      device copy of plan_out
 *      image_out [B,3,PH,PW] fp32, depth_out [B,1,PH,PW] fp32: sample i occupies the top-left
 *                out_hw[i] = {H'_i, W'_i} corner, zero elsewhere (Human36MCollator padding,
 *                src/dataset/collator.py:20-44).  PW % 4 == 0, PH >= max H', PW >= max W'.
 *      kp_out [B,J,2], joints_out [B,J,3] fp32, cam_out [B,4] fp64, out_hw [B,2] int32
 *      workspace pose_augment_workspace_bytes(...) bytes
 * ------------------------------------------------------------------------------------------- */
enum { POSE_AUG_FLIP = 1, POSE_AUG_ROTATE = 2, POSE_AUG_SCALE = 4, POSE_AUG_TRANSLATE = 8, POSE_AUG_COLOR = 16 };
#define POSE_AUG_PLAN_BYTES 256

typedef struct pose_aug_launch {
    int32_t max_out_h, max_out_w; /* over the batch: smallest legal PH / PW (before PW % 4 rounding) */
    int32_t max_rot_rows;         /* rows of the rotated band a CTA stages in shared memory */
    int32_t max_band_rows;        /* output rows per CTA */
    int32_t max_ksize;            /* taps of the antialiased resize */
    int32_t smem_bytes;           /* dynamic shared memory of the fused kernel */
    int32_t cluster;              /* CTAs per sample */
    int32_t reserved;
} pose_aug_launch;

int pose_augment_plan(const double *params, int B, int H, int W, int flags, void *plan_out,
                      pose_aug_launch *launch_out);
size_t pose_augment_workspace_bytes(int B, int H, int W, const pose_aug_launch *launch);
int pose_augment_batch(const void *image, const void *depth, int in_dtype, const float *kp, const float *joints,
                       const double *cam, const void *plan, const pose_aug_launch *launch, int B, int H, int W,
                       int J, int flags, float *image_out, float *depth_out, int PH, int PW, float *kp_out,
                       float *joints_out, double *cam_out, int32_t *out_hw, void *workspace,
                       size_t workspace_bytes, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * C. nn.Linear / 1x1 convolution contraction      reference: src/models/common.py:73-89 (head),
 *    src/models/cnn.py:16-18 (SE), :122-131 (1x1 convs in NHWC), src/models/transformers.py:20-26
 *    C[M,N] = act(A[M,K] . W[N,K]^T + bias[N]) on tcgen05 tensor cores (bf16 in, fp32 accumulate).
 *    A [M,K] bf16 row-major (pitch lda elements), W [N,K] bf16 row-major (torch Linear layout, pitch
 *    ldw), bias [N] fp32 or NULL, C [M,N] fp32 (out_dtype 0) or bf16 (1), pitch ldc.
 *    lda, ldw multiples of 8; A, W 16-byte aligned.  act: 0 none, 1 relu, 2 silu, 3 gelu(erf).
 *    pose_cast_f32_bf16 prepares operands (n elements, 16-byte aligned pointers).
 * ------------------------------------------------------------------------------------------- */
int pose_gemm_bf16(const void *A, int lda, const void *W, int ldw, const float *bias, void *C, int ldc, int M,
                   int N, int K, int act, int out_dtype, pose_stream_t stream);
int pose_cast_f32_bf16(const float *in, void *out, long n, pose_stream_t stream);

/* Extended epilogue shared by the GEMM and the implicit-GEMM convolution:
 *    C = act(acc + bias) * out_scale + residual * res_scale
 * bias [N] fp32 or NULL (a folded BatchNorm leaves only this), residual [M, ldr] bf16 or NULL, C [M, ldc] fp32
 * (out_dtype 0) or bf16 (1).  Writing with ldc > N into a column slice implements torch.cat along channels.
 * act: 0 none, 1 relu, 2 silu, 3 gelu(erf), 4 sigmoid. */
struct pose_bn_fuse;
typedef struct pose_gemm_epilogue {
    const float *bias;
    const void *residual;
    void *C;
    int32_t ldc, ldr, act, out_dtype;
    float out_scale, res_scale;
    /* training: */
    void *preact;       /* [M, ldc] bf16 or NULL: act'(acc + bias), the activation's derivative, saved for the backward pass */
    int32_t accumulate; /* 1: fp32 C += acc * out_scale (atomic adds; split-K weight gradients accumulate into .grad) */
    int32_t reserved;
    /* nn.Dropout fused after the activation and before the residual add: element (row, col) is kept iff
     * hash(drop_seed, row * ldc + col) >= drop_p * 2^32 and scaled by 1 / (1 - drop_p); with act 5..7 the same mask gates
     * the gradient.  pose_dropout_bf16 over the compact [M, ldc] tensor with the same seed reproduces the mask.
     * hash(seed, i) = lowbias32(lo32(i) ^ key(seed, hi32(i))), key = a 64-bit murmur finaliser (csrc/common.cuh); the
     * fused kernels use a 32-bit counter: M * ldc (attention: B * heads * Nq * Nk) must stay below 2^32
     * (POSE_E_UNSUPPORTED otherwise). */
    uint64_t drop_seed;
    float drop_p;
    int32_t reserved2;
    /* training-mode BatchNorm fused behind the contraction (ConvBnAct, src/models/cnn.py:135-139; nn.BatchNorm2d of
     * src/utils.py:186-187): NULL = off.  See pose_bn_fuse below. */
    const struct pose_bn_fuse *bn;
} pose_gemm_epilogue;

/* Batch statistics of a convolution's output taken from the fp32 ACCUMULATORS in the epilogue of the producing GEMM /
 * implicit-GEMM convolution (act 0, bf16 or fp32 output, N a multiple of 32): every persistent CTA keeps per-column sums
 * and sums of squares over its tiles (fixed order, no atomics), writes one partial row per TMEM lane quarter into
 * `partials` [parts, 2, N], and the same call launches the fold: mean / rstd, the affine coefficients
 * z = y * scale + shift, and nn.BatchNorm2d's running-statistics update (momentum, unbiased variance) -- the
 * pose_bn_stats_bf16 pass over the bf16 output is not needed.  `count` = rows of the output (N * Ho * Wo). */
typedef struct pose_bn_fuse {
    float *partials;            /* scratch, cap_floats >= 4 * 148 * 2 * N */
    int64_t cap_floats;
    const float *gamma, *beta;  /* [N] */
    float eps, momentum;
    int64_t count;
    float *mean_rstd;           /* [2, N] out */
    float *scale_shift;         /* [2, N] out */
    float *running_mean, *running_var;   /* [N] in/out or NULL */
} pose_bn_fuse;

int pose_gemm_bf16_ex(const void *A, int lda, const void *W, int ldw, int M, int N, int K,
                      const pose_gemm_epilogue *epilogue, pose_stream_t stream);

/* Backward contractions of nn.Linear / 1x1 convolutions (autograd of src/train.py:89-92), reading the saved
 * activations and the weights IN PLACE through MN-major tcgen05 operands (no transposed copies):
 *    a_mn = 0: A is [M, K] row-major;  a_mn = 1: A's memory image is [K, M] row-major (pitch lda)
 *    b_mn = 0: W is [N, K] row-major;  b_mn = 1: W's memory image is [K, N] row-major (pitch ldw)
 *    data gradient    dX[M, Kin] = dY[M, Nout] . W[Nout, Kin]     -> (A = dY, a_mn 0; W = weight, b_mn 1)
 *    weight gradient  dW[Nout, Kin] = dY^T . X                    -> (A = dY, a_mn 1; W = X, b_mn 1), contraction over
 *                     the rows, split k_splits ways across CTAs, epilogue->accumulate = 1
 *    act 5 in the epilogue: C = acc * g * out_scale with g = `residual` = the derivative act'(u) the forward GEMM saved
 *    through `preact` (bf16); act 6 | 7: C = acc * silu'(u) | relu'(u) with u = `residual` a saved PRE-activation. */
int pose_gemm_bf16_tr(const void *A, long lda, int a_mn, const void *W, long ldw, int b_mn, int M, int N, int K,
                      int k_splits, const pose_gemm_epilogue *epilogue, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * D. dense nn.Conv2d as implicit GEMM on tcgen05      reference: ConvBnAct, src/models/cnn.py:101-139
 *    X  [Nimg, H, W, Cin] bf16 channels-last (Cin a multiple of 32; 21-channel input padded to 32)
 *    Wt [Cout, KH, KW, Cin] bf16 (KRSC), square stride / dilation, zero padding `pad`
 *    output rows are the Nimg*Ho*Wo output pixels in NHWC order, columns the Cout channels (epilogue).
 *    Ho, Wo must tile into 128-pixel patches (TW = largest power of two dividing Wo, TH = 128 / TW).
 * ------------------------------------------------------------------------------------------- */
int pose_conv2d_bf16(const void *X, int Nimg, int H, int W, int Cin, const void *Wt, int Cout, int KH, int KW,
                     int stride, int dil, int pad, const pose_gemm_epilogue *epilogue, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * D. bandwidth-bound pieces of CNNPoseEstimation.forward, channels-last bf16
 *    (reference: src/models/cnn.py; each activation crosses HBM once per call)
 *  pose_cnn_input_pack   cat([image, depth, heatmaps], 1) (cnn.py:644-648) written as the conv1 operand
 *                        [B,S,S,32] bf16 = {R,G,B,depth, J heat-maps, zero pad}; heat-maps rendered in flight
 *  pose_dwconv3x3_bf16   depthwise 3x3, pad 1, stride 1|2 + per-channel bias (folded BN) + act; Wd [3,3,C]
 *                        fp32; pool_sum [B,pool_parts,C] fp32 (optional): partial sums of the OUTPUT per output
 *                        tile (SE/ECA squeeze), pool_parts = pose_dwconv3x3_pool_parts(H, W, stride); written,
 *                        not accumulated -- consumers add the parts in order (deterministic)
 *  pose_pool_sum_bf16    sums [B,parts,C] fp32 = partial sums over pixel chunks of X [B,HW,C]
 *  pose_se_gate          SEBlock (cnn.py:9-26): gate = sigmoid(W2 . act(W1 . (pool_sum*inv_hw)))
 *  pose_eca_gate         ECABlock (cnn.py:29-45): gate = sigmoid(conv1d_k(mean)); feat_out (optional,
 *                        bf16 [B,C]) = mean * gate (= ECA followed by global average pooling, cnn.py:612-613)
 *  pose_channel_affine_bf16  Y[b,p,c] = X[b,p,c] * mul[b,c] (fp32, optional) + add[b,c] (bf16, optional)
 *  pose_coord_pool_bf16 / pose_coord_apply_bf16   CoordAttention (cnn.py:48-98) directional means
 *                        P [B,H+W,C] and out = x * G[b,h,c] * G[b,H+w,C+c] with G [B,H+W,2C] bf16
 *  pose_avgpool2x2_bf16  AdaptiveAvgPool2d(8) on a 16x16 map (cnn.py:602)
 *  pose_sums_to_bf16     out [B,C] bf16 = (sum of the parts) * scale
 * ------------------------------------------------------------------------------------------- */
int pose_cnn_input_pack(const float *image, const float *depth, const float *kp, int B, int S, int J, float sigma,
                        void *out, pose_stream_t stream);
/* same, pixel pitch c_stride >= 32 channels (channels 32.. are left untouched: zero them once); the training step pads
 * the conv1 operand to 64 channels so that its weight gradient reads 128-byte channel chunks */
int pose_cnn_input_pack_ex(const float *image, const float *depth, const float *kp, int B, int S, int J, float sigma,
                           int c_stride, void *out, pose_stream_t stream);
int pose_dwconv3x3_pool_parts(int H, int W, int stride); /* partial-sum slots per image the kernel writes */
int pose_dwconv3x3_bf16(const void *X, int B, int H, int W, int C, const float *Wd, const float *bias, int stride,
                        int act, void *Y, float *pool_sum, int pool_parts, pose_stream_t stream);
int pose_pool_sum_bf16(const void *X, int B, int HW, int C, float *sums, int parts, pose_stream_t stream);
int pose_se_gate(const float *pool_sum, int parts, float inv_hw, const float *W1, const float *W2, int B, int C, int Cr,
                 int act, float *gate, pose_stream_t stream);
int pose_eca_gate(const float *pool_sum, int parts, float inv_hw, const float *w, int k, int B, int C, float *gate,
                  void *feat_out, pose_stream_t stream);
int pose_channel_affine_bf16(const void *X, const float *mul, const void *add, int B, long HW, int C, void *Y,
                             pose_stream_t stream);
int pose_coord_pool_bf16(const void *X, int B, int H, int W, int C, void *P, pose_stream_t stream);
int pose_coord_apply_bf16(const void *X, const void *G, int B, int H, int W, int C, void *Y, pose_stream_t stream);
int pose_avgpool2x2_bf16(const void *X, int B, int H, int W, int C, void *Y, pose_stream_t stream);
/* nn.AdaptiveAvgPool2d((OH, OW)) for any H >= OH, W >= OW (the reference default 500 x 500 input reaches the pool as a
 * 32 x 32 map): window i = [floor(i H / OH), ceil((i + 1) H / OH)) */
int pose_adaptive_avgpool_bf16(const void *X, int B, int H, int W, int C, int OH, int OW, void *Y, pose_stream_t stream);
int pose_sums_to_bf16(const float *sums, int parts, int B, int C, float scale, void *out, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * E. non-GEMM pieces of TransformerPoseEstimation.forward     reference: src/models/transformers.py:326-373
 *    (+ the timm ViT-B/16 backbone it wraps); all activations bf16 [rows, D], the Linear layers and the
 *    k16/s16 patch-embedding convolutions run on pose_gemm_bf16_ex.
 *  pose_layernorm_bf16    nn.LayerNorm over D (256/512/768/1024), fp32 statistics.  Row r = g*rows + i reads input
 *                         row g*in_group + in_off + i and writes output row g*out_group + out_off + i (drops the cls
 *                         token of every sample, or normalises only token 0: transformers.py:346, :371)
 *  pose_token_concat_bf16 dst[b] = [cls (fp32 [D], optional)] ++ src1[b] ++ src2[b] (optional) (+ pos fp32 [T,D]):
 *                         timm `_pos_embed` and transformers.py:359-366
 *  pose_patchify_bf16     fp32 NCHW planes (two channel-concatenated sources) -> bf16 [B*(H/P)*(W/P), C*P*P] in
 *                         flattened-conv-weight column order: Conv2d(kernel = stride = P) becomes one GEMM
 *  pose_attention_bf16    softmax(Q K^T * scale) V per (batch, head); Q/K/V/O rows with pitches ld* and batch
 *                         strides bs* (elements), head h in columns [h*head_dim, (h+1)*head_dim); head_dim 48 | 64,
 *                         any sequence length (key chunks of 64 stream through shared memory, online softmax with the
 *                         O accumulator in TMEM; tcgen05 kernels in csrc/attention_tc.cu).
 *                         nn.MultiheadAttention (transformers.py:61-63, :98-106) and timm Attention.
 * ------------------------------------------------------------------------------------------- */
int pose_layernorm_bf16(const void *X, const float *gamma, const float *beta, float eps, long M, int rows, long in_group,
                        long in_off, long out_group, long out_off, int D, void *Y, pose_stream_t stream);
int pose_token_concat_bf16(void *dst, int B, int T, int D, const float *cls, const void *src1, int N1, const void *src2,
                           int N2, const float *pos, pose_stream_t stream);
int pose_patchify_bf16(const float *src0, int C0, const float *src1, int C1, int B, int H, int W, int P, void *out,
                       pose_stream_t stream);
int pose_attention_bf16(const void *Q, const void *K, const void *V, void *O, int B, int heads, int Nq, int Nk, int head_dim,
                        long ldq, long ldk, long ldv, long ldo, long bsq, long bsk, long bsv, long bso, float scale,
                        float *lse /* [B, heads, Nq] fp32 or NULL: log-sum-exp of the scaled scores, saved for backward */,
                        float drop_p, uint64_t drop_seed /* attention-weight dropout (MultiheadAttention(dropout=p)); 0 = off */,
                        pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * E/G. backward of the non-GEMM transformer pieces           reference: autograd of src/train.py:89-92 over
 *      src/models/transformers.py:33-137, 326-373.  Parameter gradients are ACCUMULATED (+=, fp32) like torch's .grad.
 *  pose_layernorm_bwd_bf16  dX = dRes + LayerNorm'(dY); dgamma += , dbeta += ; X / dX / dRes use the forward's input row
 *                           addressing, dY the output addressing; statistics are recomputed from the saved input X
 *  pose_colsum_bf16         out[c] += sum_r X[r, c]                      (bias gradients)
 *  pose_batch_rowsum_bf16   out[t, :] += sum_b X[b, t_off + t, :]        (positional-embedding / class-token gradients)
 *  pose_token_slice_bf16    dst[b, i, :] = src[b, t_off + i, :], i < n   (gradient of the token concatenation)
 *  pose_attention_bwd_bf16  dQ, dK, dV from Q, K, V, O, dO and the forward's lse; Dws [B, heads, Nq] fp32 scratch
 * ------------------------------------------------------------------------------------------- */
int pose_layernorm_bwd_bf16(const void *X, const void *dY, const float *gamma, float eps, long M, int rows, long in_group,
                            long in_off, long out_group, long out_off, int D, const void *dRes, void *dX, float *dgamma,
                            float *dbeta, pose_stream_t stream);
int pose_colsum_bf16(const void *X, long M, int N, long ld, float *out, pose_stream_t stream);
/* out[r, c] (bf16, pitch ld_out) = in[r, c] (fp32, pitch ld_in) for c < cols, 0 for cols <= c < ld_out */
int pose_cast_f32_bf16_2d(const float *in, long ld_in, long rows, int cols, void *out, long ld_out, pose_stream_t stream);
int pose_batch_rowsum_bf16(const void *X, int B, long T_in, long t_off, int T_out, int D, float *out, pose_stream_t stream);
int pose_token_slice_bf16(const void *src, int B, long T, long t_off, int n, int D, void *dst, pose_stream_t stream);
int pose_attention_bwd_bf16(const void *Q, const void *K, const void *V, const void *O, const void *dO, const float *lse,
                            void *dQ, void *dK, void *dV, float *Dws, int B, int heads, int Nq, int Nk, int head_dim,
                            long ldq, long ldk, long ldv, long ldo, long lddo, long lddq, long lddk, long lddv, long bsq,
                            long bsk, long bsv, long bso, long bsdo, long bsdq, long bsdk, long bsdv, float scale,
                            float drop_p, uint64_t drop_seed, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * 8f-2. transforms.Resize(image_size) of decoded frames     reference: main.py:171-173, applied per frame in
 *    src/dataset/chunked_dataset.py:100-129; depth rescale :159-164.
 *    = F.interpolate(mode="bilinear", align_corners=False, antialias=True) of float frames: ATen's separable anti-aliased
 *    resampling, bit-equal to the CPU tensor the reference's DataLoader worker produces (weights, promotions and tap order
 *    restated, see csrc/resize.cu).  src [B, C, H, W] fp32 in [0, 1] (in_dtype 0) or uint8 (1: converted as `.float() /
 *    255.0`); dst [B, C, OH, OW] fp32; mul / add [B] fp32 or NULL: dst = dst * mul[b] + add[b] (depth * (max - min) + min).
 *    workspace: pose_resize_workspace_bytes(H, W, OH, OW), 16-byte aligned (weight tables, rebuilt by every call).
 * ------------------------------------------------------------------------------------------- */
size_t pose_resize_workspace_bytes(int H, int W, int OH, int OW);
int pose_resize_bilinear_aa(const void *src, int in_dtype, int B, int C, int H, int W, int OH, int OW, const float *mul,
                            const float *add, void *workspace, size_t workspace_bytes, float *dst, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * G'. device-resident per-step state: what changes from one training step to the next -- the dropout stream and AdamW's
 *    step count (bias correction) -- read from device memory, so that the whole step (src/train.py:76-119) can be captured
 *    ONCE as a CUDA graph and replayed.  pose_step_tick (one thread) advances it at the start of a step: ++counter,
 *    drop_key = hash(counter), ++adam_step.  While a state is bound (host-side, process-wide; NULL unbinds), every dropout
 *    mask key becomes key(seed) ^ drop_key and pose_adamw_step[_g16] with step == 0 takes the step count from adam_step.
 * ------------------------------------------------------------------------------------------- */
typedef struct pose_step_state {
    uint32_t drop_key;
    int32_t adam_step;
    uint32_t counter;
    uint32_t reserved;
} pose_step_state;
int pose_step_state_bind(pose_step_state *dev_state);
int pose_step_tick(pose_step_state *dev_state, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * G. optimizer step                       reference: torch.optim.AdamW(lr 1e-3, weight_decay 0.01), main.py:154-156,
 *    src/train.py:117-119.  One launch over a flat fp32 parameter buffer (n % 4 == 0): p, exp_avg, exp_avg_sq updated
 *    in place from grad * grad_scale; shadow_bf16 (optional) receives the bf16 copy of the new parameters (the GEMM
 *    operands); zero_grad = 1 clears grad for the next accumulation window (optimizer.zero_grad()).  step >= 1, or 0 =
 *    the bound pose_step_state's adam_step (G').
 * ------------------------------------------------------------------------------------------- */
int pose_adamw_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, void *shadow_bf16, long n, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, int zero_grad,
                    pose_stream_t stream);
/* Data-parallel variant: the gradient is read from `grad_bf16` [n] (the bf16 copy of the flat gradient that went through
 * the all-reduce: half the NVLink bytes of an fp32 exchange); `grad` (fp32, the buffer the backward pass accumulates
 * into) is only cleared. */
int pose_adamw_step_g16(float *param, float *grad, const void *grad_bf16, float *exp_avg, float *exp_avg_sq,
                        void *shadow_bf16, long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                        int step, float grad_scale, int zero_grad, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * D/G. CNN training step                reference: loss.backward() over src/models/cnn.py (src/train.py:83-92)
 *  pose_conv2d_wgrad_bf16   weight gradient of a dense convolution as an implicit GEMM on tcgen05: dWk[Cout, KH*KW*Cin]
 *                           (fp32, KRSC, ACCUMULATED) += sum over output pixels of dY[pixel, co] * X[pixel @ tap, ci];
 *                           the contraction runs over patches of 64 output pixels fetched by 4-D TMA boxes (zero fill =
 *                           padding), split k_splits ways.  Cin % 64 == 0.  The data gradient of a stride-1 convolution is
 *                           pose_conv2d_bf16 over dY with the flipped / transposed weights (pose_param_repack kind 2).
 *  BatchNorm2d in training mode (batch statistics, cnn.py:135-139), fused with activation, residual add and concat:
 *   pose_bn_stats_bf16      per-block partial sums of y and y^2 into `partials` (scratch of cap_floats floats, shared by
 *                           all layers: consumers run before the next producer on the stream); no atomics
 *   pose_bn_finalize        folds the partials in a fixed order (deterministic statistics): mean / rstd, the affine
 *                           (scale, shift) that normalises, running statistics (momentum, unbiased variance) --
 *                           nn.BatchNorm2d semantics.  Pass the same (count = M, C, cap_floats) as to pose_bn_stats_bf16
 *   pose_bn_apply_bf16      out[r, :ld_out] = residual + out_scale * act(y * scale + shift)
 *   pose_bn_bwd_bf16        dY from dA (pitch ld_da: a column slice of a concatenation) -- per-channel partial sums, a
 *                           fixed-order fold (dgamma, dbeta accumulated; coef [2, C] scratch), then the gradient
 *  pose_dwconv3x3_bwd_bf16  depthwise 3x3 backward: dX (+ add) and / or dW (parameter layout [C,1,3,3], accumulated)
 *  pose_gate_bwd_*          x * gate[b, c] (SE / ECA): dgate[b,c] += sum_p dOut * x;  dX = add + dOut * gate + dmean / HW
 *  pose_sigmoid_bwd, pose_eca_bwd, pose_coord_bwd_*, pose_wasp_mix_*: see csrc/cnn_train.cu
 *  pose_avgpool2x2_bwd_bf16, pose_scatter_strided_add_bf16 (data gradient of a strided 1x1 conv), pose_add_bf16
 *  pose_dropout_bf16        counter-based mask hash(seed, i) -- the same call on the gradient is the backward
 *  pose_param_repack        one launch re-lays-out every parameter the kernels read in another layout / precision
 *                           (table of pose_repack_entry on the device; kinds documented in csrc/cnn_train.cu)
 * ------------------------------------------------------------------------------------------- */
typedef struct pose_repack_entry {
    int64_t src, dst;
    int32_t kind, d0, d1, d2, d3, pad;
} pose_repack_entry;

int pose_conv2d_wgrad_bf16(const void *dY, const void *X, int Nimg, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                           int dil, int pad, float *dWk, int k_splits, pose_stream_t stream);
int pose_bn_stats_bf16(const void *Y, long M, int C, long ld, float *partials, long cap_floats, pose_stream_t stream);
/* depthwise 3x3 (no bias, no activation) that also emits the first stage of the BatchNorm batch statistics of its (bf16-
 * rounded) output: partials [B * pose_dwconv3x3_pool_parts(H, W, stride), 2, C]; fold with pose_bn_finalize_parts.  The
 * separate statistics pass over the convolution output disappears (cnn.py:135-139 ConvBnAct with groups = channels). */
int pose_dwconv3x3_bn_stats_bf16(const void *X, int B, int H, int W, int C, const float *Wd, int stride, void *Y, float *partials,
                                 long cap_floats, pose_stream_t stream);
int pose_bn_finalize_parts(const float *partials, int parts, long count, const float *gamma, const float *beta, float eps,
                           float momentum, int C, float *mean_rstd, float *scale_shift, float *running_mean, float *running_var,
                           pose_stream_t stream);
int pose_bn_finalize(const float *partials, long cap_floats, long count, const float *gamma, const float *beta, float eps,
                     float momentum, int C, float *mean_rstd, float *scale_shift, float *running_mean, float *running_var,
                     pose_stream_t stream);
int pose_bn_apply_bf16(const void *Y, long M, int C, const float *scale_shift, int act, float out_scale, const void *residual,
                       long ld_res, void *out, long ld_out, pose_stream_t stream);
/* pose_bn_apply_bf16 (no residual, compact output) for the layer in front of an SE / ECA block, fused with that block's
 * squeeze (cnn.py:22-23, 40-41): pool [B, parts, C] receives per-(image, row block) channel sums of the bf16 outputs. */
int pose_bn_apply_pool_bf16(const void *Y, int B, long HW, int C, const float *scale_shift, int act, void *out, float *pool,
                            int parts, pose_stream_t stream);
int pose_bn_bwd_bf16(const void *dA, long ld_da, const void *Y, long M, int C, const float *scale_shift, const float *mean_rstd,
                     int act, float out_scale, float *partials, long cap_floats, float *coef, void *dY, float *dgamma,
                     float *dbeta, pose_stream_t stream);
int pose_dwconv3x3_bwd_bf16(const void *dY, const void *X, const float *Wd, int B, int H, int W, int C, int stride,
                            const void *add, void *dX, float *dW, pose_stream_t stream);
/* BatchNorm backward with its reduction pass folded into the PRODUCER of the incoming gradient (src/models/cnn.py:135-139
 * under loss.backward()): the producer multiplies by act'(z) of the layer whose output it differentiates -- reading that
 * layer's saved conv output `Yprev` and (scale, shift) -- writes dz instead of dA and emits [parts, 2, C] partial sums of dz
 * and dz * y; pose_bn_bwd_from_dz_bf16 folds them (dgamma, dbeta accumulated, coef [2, C] scratch) and writes dY.
 *   pose_dwconv3x3_bnbwd_bf16    producer = data gradient of a stride-1 depthwise 3x3 (taps flipped: repack kind 6);
 *                                parts = B * pose_dwconv3x3_pool_parts(H, W, 1)
 *   pose_gate_bwd_apply_bn_bf16  producer = backward of x * gate[b, c] (SE / ECA); *parts_out receives the partial count
 * act_prev: 1 relu, 2 silu. */
int pose_dwconv3x3_bnbwd_bf16(const void *dY, int B, int H, int W, int C, const float *Wflip, const void *Yprev,
                              const float *scale_shift_prev, int act_prev, void *dZ, float *partials, long cap_floats,
                              pose_stream_t stream);
int pose_gate_bwd_apply_bn_bf16(const void *dOut, const float *gate, const void *dmean, float inv_hw, int B, long HW, int C,
                                const void *Yprev, const float *scale_shift_prev, int act_prev, void *dZ, float *partials,
                                long cap_floats, int *parts_out, pose_stream_t stream);
int pose_bn_bwd_from_dz_bf16(const void *dZ, long ld_dz, const void *Y, long M, int C, const float *scale_shift,
                             const float *mean_rstd, const float *partials, int parts, float *coef, void *dY, float *dgamma,
                             float *dbeta, pose_stream_t stream);
int pose_gate_bwd_reduce_bf16(const void *dOut, const void *X, int B, long HW, int C, float *dgate, pose_stream_t stream);
int pose_gate_bwd_apply_bf16(const void *dOut, const float *gate, const void *dmean, float inv_hw, int B, long HW, int C,
                             const void *add, void *dX, pose_stream_t stream);
int pose_sigmoid_bwd(const float *dgate, const float *gate, long n, void *dz, pose_stream_t stream);
int pose_eca_bwd(const float *dgate, const void *dfeat, const float *gate, const float *pool_sum, int parts, float inv_hw,
                 const float *w, int k, int B, int C, int mode, void *dmean, float *dw, pose_stream_t stream);
int pose_coord_bwd_reduce_bf16(const void *dOut, const void *X, const void *G, int B, int H, int W, int C, void *dZ,
                               pose_stream_t stream);
int pose_coord_bwd_apply_bf16(const void *dOut, const void *G, const void *dP, int B, int H, int W, int C, void *dX,
                              pose_stream_t stream);
int pose_wasp_mix_bf16(const void *branches, int nb, const void *glob, const float *raw_weights, int B, long HW, int C, void *out,
                       pose_stream_t stream);
int pose_wasp_mix_bwd_bf16(const void *dOut, const void *branches, int nb, const void *glob, const float *raw_weights, int B,
                           long HW, int C, void *dbranches, float *dglob, float *dots, float *draw, pose_stream_t stream);
int pose_avgpool2x2_bwd_bf16(const void *dY, int B, int H, int W, int C, void *dX, pose_stream_t stream);
int pose_adaptive_avgpool_bwd_bf16(const void *dY, int B, int H, int W, int C, int OH, int OW, void *dX, pose_stream_t stream);
int pose_scatter_strided_add_bf16(const void *dXs, int B, int Ho, int Wo, int H, int W, int C, int stride, void *dX,
                                  pose_stream_t stream);
int pose_add_bf16(const void *a, const void *b, long n, void *out, pose_stream_t stream);
int pose_dropout_bf16(const void *x, long n, float p, uint64_t seed, void *out, pose_stream_t stream);
int pose_param_repack(const void *table, int n_entries, const float *src_f32, float *dst_f32, void *dst_bf16,
                      pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * next row: validation metrics     reference: src/utils.py:55-69 compute_mpjpe, :72-165
 *    compute_pa_mpjpe (per-sample Python loop + 3x3 SVD in the reference; src/train.py:249-254).
 *    pred, gt [B, J, 3] fp32; per_sample [2, B] fp32 = {MPJPE_b}, {PA-MPJPE_b}; means [2] fp32 = batch means
 *    (what the reference functions return).  The reference's rotation convention (Pc @ V U^T) is reproduced.
 * ------------------------------------------------------------------------------------------- */
int pose_eval_metrics(const float *pred, const float *gt, int B, int J, float *per_sample, float *means, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * next row: model-input preparation of the inference path     reference: infer.py:319-380
 *    depth [B, h, w] fp32 -> depth_out [B, H, W] = F.interpolate(mode="bilinear", align_corners=False) (infer.py:362-367,
 *    ATen's source-index rule, evaluated without contraction: equal to the CPU path bit for bit);
 *    kpts_px_conf [B, K, 3] = (x_pixel, y_pixel, confidence) -> kp_norm [B, K, 2] = (x / img_w, y / img_h) (infer.py:217-221)
 *    and, optionally, kp_norm_conf [B, K, 3] (the visualisation copy).  kpts_px_conf may be NULL (depth only).
 * ------------------------------------------------------------------------------------------- */
int pose_infer_prep(const float *depth, int B, int h, int w, int H, int W, float *depth_out, const float *kpts_px_conf, int K,
                    float img_w, float img_h, float *kp_norm, float *kp_norm_conf, pose_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * next row: collate on the device     reference: src/dataset/collator.py:10-61 (Human36MCollator)
 *    table: B entries of 32 bytes on the device {const float *image [3,h,w]; const float *depth [1,h,w]; int h, w;
 *    float depth_scale, depth_shift}; image [B,3,Hm,Wm] / depth [B,1,Hm,Wm] fp32 = every sample zero padded on the right /
 *    bottom to the batch maximum (F.pad + torch.stack), depth optionally rescaled depth * scale + shift on the way
 *    (src/dataset/chunked_dataset.py:159-164; pass scale 1, shift 0 for a plain copy).
 * ------------------------------------------------------------------------------------------- */
int pose_collate_pad(const void *table, int B, int Hm, int Wm, float *image, float *depth, pose_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* POSE_B200_H */
