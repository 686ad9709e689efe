// gemm_tcgen05.cu -- bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM),
// operands staged by TMA, fused bias + activation epilogue.
//
//   C[M, N] = act(A[M, K] . W[N, K]^T + bias[N])         A, W bf16 row-major (K contiguous)
//
// This is the contraction behind every nn.Linear on the hot path (PoseRegressionHead:
// src/models/common.py:73-81, SE gates: src/models/cnn.py:16-18, ViT qkv/proj/fc1/fc2) and every 1x1
// convolution once activations are channels-last (NHWC makes a 1x1 conv exactly this GEMM with
// M = B*H*W: src/models/cnn.py:122-131).
//
// Structure (PERSISTENT: grid = min(tiles, 148), each CTA walks 128 x BN output tiles; 320 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor loads of A (128 x 64) and W (BN x 64) tiles into a kStages-deep
//            128B-swizzled shared-memory ring (`full` / `empty` mbarriers); runs ahead across tile boundaries
//   warp 1   TMEM allocator + MMA issuer: one elected lane issues tcgen05.mma (M=128, N=BN, K=16) four times
//            per stage; tcgen05.commit releases the stage and, after the last k-block, signals `acc_full[s]`
//   warps 2-17 epilogue: the accumulator is DOUBLE BUFFERED in TMEM (2 x BN columns): tcgen05.ld the finished
//            stage (four warps per 32-lane quarter split the column chunks), hand it back (`acc_empty[s]`) and do
//            bias + activation + residual + convert + 16-byte stores while the next tile's MMAs already run
// The same kernel is the implicit-GEMM convolution (MODE 1): the A tile of a k-block is one filter tap
// x one channel chunk, fetched by a 4-D tiled TMA box over the NHWC activation (box = channels x TW x TH
// output-pixel patch, element strides = conv stride, start = tap offset - padding; out-of-bounds
// elements are zero-filled by the TMA unit, which is exactly the conv padding).  Weights are KRSC
// ([Cout][kh][kw][Cin]) so the B operand stays a plain K-major 2-D tile.  Dense 3x3 / 5x5 / dilated /
// strided convolutions of src/models/cnn.py:122-131 never materialise an im2col matrix.
// All mbarrier waits are bounded (trap instead of hanging the GPU if a descriptor is wrong).
#include <cuda.h>
#include <cstdlib>
#include "tc_common.cuh"

namespace pose {

constexpr int BM = 128;   // UMMA M (cta_group::1)
constexpr int BK = 64;    // 64 bf16 = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case 1: return v > 0.f ? v : 0.f;                                   // relu
        case 2: return v / (1.0f + __expf(-v));                             // silu
        case 3: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));  // gelu (erf form, nn.GELU default)
        case 4: return 1.0f / (1.0f + __expf(-v));                          // sigmoid
        default: return v;
    }
}

// ------------------------------------------------------------------------------------------ kernel
struct Epilogue {
    const float *bias;               // [N] fp32 or null   (BatchNorm folded into weights leaves only this)
    const __nv_bfloat16 *residual;   // [M, ldr] bf16 or null
    void *C;
    int ldc, ldr, act, out_bf16;
    int vec;                         // 16-byte vector path allowed (bases and row pitches aligned; host-checked)
    float out_scale, res_scale;      // C = act(acc + bias) * out_scale + residual * res_scale
    __nv_bfloat16 *preact;           // [M, ldc] bf16 or null: acc + bias before the activation (saved for backward)
    int accumulate;                  // fp32 C += acc * out_scale with red.global (split-K weight gradients)
    int tma;                         // bf16 output tiles leave through TMA stores (row-per-lane epilogue, no transpose)
    uint32_t drop_thresh;            // dropout after the activation (before the residual): keep iff hash >= thresh
    float drop_scale;                // 1 / (1 - p)
    DropSeed drop_seed;              // element index = row * ldc + col (common.cuh: drop_keep)
    float *stats;                    // [parts, 2, N] fp32 or null: per-CTA column sums / sums of squares of the fp32 result
                                     // (training-mode BatchNorm statistics taken from the accumulators, fixed order)
    const pose_bn_fuse *bn;          // host side only: the fold launched behind the kernel
};

struct ConvGeom {          // MODE 1 only
    int Ho, Wo;            // output height / width
    int TH, TW, TN;        // output patch of one M tile: TN images x TH x TW pixels (= 128); edge patches may be ragged (rows
                           // past Ho / Wo are loaded as whatever the TMA box covers -- zeros outside the tensor -- and masked
                           // in the epilogue; in the weight gradient their dY is the TMA zero fill)
    int Nimg;
    int KW, taps;          // filter width, KH * KW
    int stride, dil, pad;
    int cchunks;           // Cin_pad / BKC
    int tiles_w, tiles_h;  // ceil(Wo / TW), ceil(Ho / TH)
    int cw;                // MODE 2: channels per B block (64: 128-byte pixels, SWIZZLE_128B; 32: 64-byte pixels, SWIZZLE_64B)
    int linear;            // MODE 1, host side: tile row r of M tile mt IS output pixel mt * 128 + r (full-width or single-row
                           // patches that tile the map exactly), so the TMA-store epilogues of the plain GEMM apply
};

template <int BN, int kStages, int BKC, int TMA_EPI = 0, int CG2 = 0>
struct GemmSmem {
    // CG2 (CTA pair): each CTA of the pair keeps its 128 rows of A and HALF of the B tile
    static constexpr int kABytes = BM * BKC * 2, kBBytes = (BN >> CG2) * BKC * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarrierBytes = 1024;                   // keeps the staging buffers 1024 B aligned (TMA swizzle)
    static constexpr int kStagePitch = 144;                      // 32 fp32 + 16 B pad: conflict-free 16 B accesses
    // per epilogue warp: a 32 x 32 fp32 transpose buffer (4608 B), or two 32 x 32 bf16 TMA tiles (output at +0,
    // auxiliary at +2048); the smaller TMA staging leaves room for one more pipeline stage
    // TMA_EPI 2 (lean plain epilogue): one 2560-byte region (residual tile / column-reduction half tiles / output tile in
    // turn) at +0 and the running column statistics (512) at +3072
    static constexpr int kWarpStaging = TMA_EPI ? 4096 : 5120;
    static constexpr int kStagingBytes = 16 * kWarpStaging;
    static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + kStagingBytes + 1024;  // + alignment slack
    static_assert(kStageBytes % 1024 == 0, "stage bases must stay 1024 B aligned");
};

constexpr int kEpiWarps = 16;                             // four per TMEM lane quarter
constexpr int kGemmThreadsP = 64 + kEpiWarps * 32;        // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue

// explicit shared-space accesses (the staging pointer is carved out of the raw dynamic buffer, so the compiler
// would otherwise emit generic LD / ST)
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr) : "memory");
    return v;
}

// sigmoid(x) = 0.5 (1 + tanh(x / 2)): MUFU.TANH (rel. error ~2^-11, below bf16 resolution) instead of EX2 + RCP --
// the epilogue of the small-K layers is bound by the 16 MUFU/clk/SM special-function unit
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// erf(|x|) and exp(-x^2) in one go, branch-free (Abramowitz & Stegun 7.1.26, |error| <= 1.5e-7: two MUFU ops -- rcp and
// ex2 -- and six FMAs instead of the two-range libdevice erff; the exact-GELU epilogues were instruction bound on it).
// The same exponential is the Gaussian density of gelu'.
__device__ __forceinline__ float erf_abs_exp(float ax, float &e) {
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
    e = __expf(-ax * ax);
    const float p = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t, -0.284496736f), t, 0.254829592f) * t;
    return fmaf(-p, e, 1.0f);
}
__device__ __forceinline__ float gelu_erf(float v) {     // nn.GELU() (exact, erf form)
    float e;
    const float x = v * 0.70710678118654752f;
    const float er = copysignf(erf_abs_exp(fabsf(x), e), x);
    return 0.5f * v * (1.0f + er);
}

// Two exact-GELU evaluations on the packed fp32 pipe (FFMA2 / FMUL2: two operations per issue slot).  The epilogues of the
// MLP GEMMs were bound by the FMA pipe (~17 fp32 / integer-multiply operations per element against a K = 768 main loop);
// `g` receives gelu'(v) when WITH_GRAD.  Same Abramowitz-Stegun erf as erf_abs_exp.
template <bool WITH_GRAD>
__device__ __forceinline__ float2 gelu2(float2 v, float2 &g, float out_scale) {
    const float2 x = __fmul2_rn(v, make_float2(0.70710678118654752f, 0.70710678118654752f));
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 den = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), ax, make_float2(1.0f, 1.0f));
    float2 t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
    const float2 ee = __fmul2_rn(__fmul2_rn(ax, ax), make_float2(-1.4426950408889634f, -1.4426950408889634f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(ee.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(ee.y));
    float2 p = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
    p = __ffma2_rn(p, t, make_float2(1.421413741f, 1.421413741f));
    p = __ffma2_rn(p, t, make_float2(-0.284496736f, -0.284496736f));
    p = __ffma2_rn(p, t, make_float2(0.254829592f, 0.254829592f));
    p = __fmul2_rn(p, t);
    float2 h = __ffma2_rn(__fmul2_rn(p, e), make_float2(-0.5f, -0.5f), make_float2(0.5f, 0.5f));     // erf(|x|) / 2
    h.x = copysignf(h.x, x.x);
    h.y = copysignf(h.y, x.y);
    const float2 phi = __fadd2_rn(h, make_float2(0.5f, 0.5f));
    if (WITH_GRAD) g = __ffma2_rn(__fmul2_rn(v, make_float2(0.3989422804014327f, 0.3989422804014327f)), e, phi);
    return __fmul2_rn(__fmul2_rn(v, phi), make_float2(out_scale, out_scale));
}

template <int ACT>
__device__ __forceinline__ float fast_act(float v) {
    if (ACT == 1) return fmaxf(v, 0.f);
    if (ACT == 2) return 0.5f * v * (1.0f + tanh_approx(0.5f * v));   // silu(x) = x * sigmoid(x), one MUFU
    if (ACT == 3) return gelu_erf(v);
    if (ACT == 4) return 0.5f * (1.0f + tanh_approx(0.5f * v));       // sigmoid
    return v;
}
// activation and its derivative in one evaluation (the forward pass of a training step stores act'(u) for the backward
// data-gradient GEMM, whose epilogue then only multiplies: the erf / exp / tanh are paid once)
template <int ACT>
__device__ __forceinline__ float act_with_grad(float v, float &g) {
    if (ACT == 1) {
        g = v > 0.f ? 1.f : 0.f;
        return fmaxf(v, 0.f);
    }
    if (ACT == 2) {
        const float sg = 0.5f * (1.0f + tanh_approx(0.5f * v));
        g = sg * (1.0f + v * (1.0f - sg));
        return v * sg;
    }
    if (ACT == 3) {
        float e;
        const float x = v * 0.70710678118654752f;
        const float phi = 0.5f * (1.0f + copysignf(erf_abs_exp(fabsf(x), e), x));
        g = fmaf(v * 0.3989422804014327f, e, phi);
        return v * phi;
    }
    if (ACT == 4) {
        const float sg = 0.5f * (1.0f + tanh_approx(0.5f * v));
        g = sg * (1.0f - sg);
        return sg;
    }
    g = 1.f;
    return v;
}

// derivative of the activation at a saved pre-activation u (ACT 6 = silu', 7 = relu'; ACT 5 multiplies by a SAVED derivative)
template <int ACT>
__device__ __forceinline__ float act_grad(float u) {
    if (ACT == 5) return u;
    if (ACT == 6) {
        const float sg = 0.5f * (1.0f + tanh_approx(0.5f * u));
        return sg * (1.0f + u * (1.0f - sg));
    }
    if (ACT == 7) return u > 0.f ? 1.0f : 0.0f;
    return 1.0f;
}
__device__ __forceinline__ void red_add_f4(float *p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// Phase 2 of the epilogue for one 32 x 32 chunk sitting in the warp's staging buffer (fp32, pitch kPitch):
// 8 lanes per output row, 4 columns per lane.  The activation is a template parameter so that the loop body
// stays small (a run-time switch per element made the unrolled epilogue overflow the instruction cache).
template <int ACT, int kPitch>
__device__ __forceinline__ void epilogue_rows(uint32_t stg, const Epilogue &ep, long row, int lane, int col,
                                              int N, uint32_t wstat) {
    const int sub_r = lane >> 3, colq = (lane & 7) * 4;
    const uint32_t dkey = ep.drop_thresh ? drop_key0(ep.drop_seed) : 0u;
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = make_float4(0.f, 0.f, 0.f, 0.f);   // column statistics (ACT 0 only)
    float4 bz = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias != nullptr) bz = __ldg((const float4 *)(ep.bias + col));
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const int tr = g * 4 + sub_r;                       // tile row within the quarter
        const long grow = __shfl_sync(0xffffffffu, row, tr);
        float4 x = lds128(stg + tr * kPitch + colq * 4);
        if (grow < 0) continue;
        if (ACT >= 5) {
            // backward through an activation: C = acc * act'(u) * out_scale (saved derivative or pre-activation: `residual`)
            const uint2 pk = __ldg((const uint2 *)(ep.residual + grow * ep.ldr + col));
            const float2 u0 = __bfloat1622float2(*(const __nv_bfloat162 *)&pk.x);
            const float2 u1 = __bfloat1622float2(*(const __nv_bfloat162 *)&pk.y);
            x.x *= act_grad<ACT>(u0.x) * ep.out_scale; x.y *= act_grad<ACT>(u0.y) * ep.out_scale;
            x.z *= act_grad<ACT>(u1.x) * ep.out_scale; x.w *= act_grad<ACT>(u1.y) * ep.out_scale;
            if (ep.drop_thresh) {     // the forward dropped act(u): the same mask gates the gradient
                const uint32_t i0 = (uint32_t)(grow * ep.ldc + col);   // < 2^32 (host check)
                x.x = drop_keep32(dkey, i0, ep.drop_thresh) ? x.x * ep.drop_scale : 0.f;
                x.y = drop_keep32(dkey, i0 + 1, ep.drop_thresh) ? x.y * ep.drop_scale : 0.f;
                x.z = drop_keep32(dkey, i0 + 2, ep.drop_thresh) ? x.z * ep.drop_scale : 0.f;
                x.w = drop_keep32(dkey, i0 + 3, ep.drop_thresh) ? x.w * ep.drop_scale : 0.f;
            }
        } else {
            x.x += bz.x; x.y += bz.y; x.z += bz.z; x.w += bz.w;
            if (ep.preact != nullptr) {          // training: keep act'(u) (bf16) for the backward data-gradient GEMM
                float4 g;
                x.x = act_with_grad<ACT>(x.x, g.x) * ep.out_scale;
                x.y = act_with_grad<ACT>(x.y, g.y) * ep.out_scale;
                x.z = act_with_grad<ACT>(x.z, g.z) * ep.out_scale;
                x.w = act_with_grad<ACT>(x.w, g.w) * ep.out_scale;
                __nv_bfloat162 p0 = __floats2bfloat162_rn(g.x, g.y), p1 = __floats2bfloat162_rn(g.z, g.w);
                *(uint2 *)(ep.preact + grow * ep.ldc + col) = make_uint2(*(uint32_t *)&p0, *(uint32_t *)&p1);
            } else {
                x.x = fast_act<ACT>(x.x) * ep.out_scale;
                x.y = fast_act<ACT>(x.y) * ep.out_scale;
                x.z = fast_act<ACT>(x.z) * ep.out_scale;
                x.w = fast_act<ACT>(x.w) * ep.out_scale;
            }
            if (ep.drop_thresh) {
                const uint32_t i0 = (uint32_t)(grow * ep.ldc + col);   // < 2^32 (host check)
                x.x = drop_keep32(dkey, i0, ep.drop_thresh) ? x.x * ep.drop_scale : 0.f;
                x.y = drop_keep32(dkey, i0 + 1, ep.drop_thresh) ? x.y * ep.drop_scale : 0.f;
                x.z = drop_keep32(dkey, i0 + 2, ep.drop_thresh) ? x.z * ep.drop_scale : 0.f;
                x.w = drop_keep32(dkey, i0 + 3, ep.drop_thresh) ? x.w * ep.drop_scale : 0.f;
            }
            if (ep.residual != nullptr) {
                const uint2 pk = __ldg((const uint2 *)(ep.residual + grow * ep.ldr + col));
                const float2 f0 = __bfloat1622float2(*(const __nv_bfloat162 *)&pk.x);
                const float2 f1 = __bfloat1622float2(*(const __nv_bfloat162 *)&pk.y);
                x.x += f0.x * ep.res_scale; x.y += f0.y * ep.res_scale;
                x.z += f1.x * ep.res_scale; x.w += f1.y * ep.res_scale;
            }
        }
        if (ACT == 0 && ep.stats != nullptr) {
            s1.x += x.x; s1.y += x.y; s1.z += x.z; s1.w += x.w;
            s2.x = fmaf(x.x, x.x, s2.x); s2.y = fmaf(x.y, x.y, s2.y); s2.z = fmaf(x.z, x.z, s2.z); s2.w = fmaf(x.w, x.w, s2.w);
        }
        if (ep.accumulate) {
            red_add_f4((float *)ep.C + grow * ep.ldc + col, x);
        } else if (ep.out_bf16) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(x.x, x.y), p1 = __floats2bfloat162_rn(x.z, x.w);
            *(uint2 *)((__nv_bfloat16 *)ep.C + grow * ep.ldc + col) = make_uint2(*(uint32_t *)&p0, *(uint32_t *)&p1);
        } else {
            *(float4 *)((float *)ep.C + grow * ep.ldc + col) = x;
        }
    }
    if (ACT == 0 && ep.stats != nullptr) {
        // fold the four row groups of the warp (lanes l, l + 8, l + 16, l + 24 own the same columns) in a fixed order and add
        // the chunk's 32-row column sums to the warp's private running totals in shared memory
#define FOLD_(v_) v_ += __shfl_xor_sync(0xffffffffu, v_, 8); v_ += __shfl_xor_sync(0xffffffffu, v_, 16);
        FOLD_(s1.x) FOLD_(s1.y) FOLD_(s1.z) FOLD_(s1.w) FOLD_(s2.x) FOLD_(s2.y) FOLD_(s2.z) FOLD_(s2.w)
#undef FOLD_
        if (lane < 8) {
            const float4 t1 = lds128(wstat + colq * 4), t2 = lds128(wstat + 128 + colq * 4);
            sts128(wstat + colq * 4, __float_as_uint(t1.x + s1.x), __float_as_uint(t1.y + s1.y), __float_as_uint(t1.z + s1.z),
                   __float_as_uint(t1.w + s1.w));
            sts128(wstat + 128 + colq * 4, __float_as_uint(t2.x + s2.x), __float_as_uint(t2.y + s2.y), __float_as_uint(t2.z + s2.z),
                   __float_as_uint(t2.w + s2.w));
        }
    }
}

// Ragged / unaligned variant (N % 4 != 0 at the edge, or unaligned pitches): element-wise, kept out of line.
template <int kPitch>
__device__ __forceinline__ void epilogue_rows_slow(uint32_t stg, const Epilogue &ep, long row, int lane,
                                                   int col, int N) {
    // (inlined so that `ep` stays in the constant bank -- taking its address would spill it to local memory --
    //  but with rolled loops: this path only runs for ragged / unaligned edges)
    const int sub_r = lane >> 3, colq = (lane & 7) * 4;
    const uint32_t dkey = ep.drop_thresh ? drop_key0(ep.drop_seed) : 0u;
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
        const int tr = g * 4 + sub_r;
        const long grow = __shfl_sync(0xffffffffu, row, tr);
        if (grow < 0) continue;
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
            if (col + q >= N) break;
            float y = lds32(stg + tr * kPitch + (colq + q) * 4);
            if (ep.act >= 5) {
                const float u = __bfloat162float(ep.residual[grow * ep.ldr + col + q]);
                y *= (ep.act == 5 ? act_grad<5>(u) : (ep.act == 6 ? act_grad<6>(u) : act_grad<7>(u))) * ep.out_scale;
                if (ep.drop_thresh)
                    y = drop_keep32(dkey, (uint32_t)(grow * ep.ldc + col + q), ep.drop_thresh) ? y * ep.drop_scale : 0.f;
            } else {
                if (ep.bias != nullptr) y += __ldg(ep.bias + col + q);
                float gq = 1.f;
                switch (ep.act) {
                    case 1: y = act_with_grad<1>(y, gq); break;
                    case 2: y = act_with_grad<2>(y, gq); break;
                    case 3: y = act_with_grad<3>(y, gq); break;
                    case 4: y = act_with_grad<4>(y, gq); break;
                    default: break;
                }
                if (ep.preact != nullptr) ep.preact[grow * ep.ldc + col + q] = __float2bfloat16_rn(gq);
                y *= ep.out_scale;
                if (ep.drop_thresh)
                    y = drop_keep32(dkey, (uint32_t)(grow * ep.ldc + col + q), ep.drop_thresh) ? y * ep.drop_scale : 0.f;
                if (ep.residual != nullptr) y += __bfloat162float(ep.residual[grow * ep.ldr + col + q]) * ep.res_scale;
            }
            if (ep.accumulate) atomicAdd((float *)ep.C + grow * ep.ldc + col + q, y);
            else if (ep.out_bf16) ((__nv_bfloat16 *)ep.C)[grow * ep.ldc + col + q] = __float2bfloat16_rn(y);
            else ((float *)ep.C)[grow * ep.ldc + col + q] = y;
        }
    }
}

// ---- TMA epilogue --------------------------------------------------------------------------------------------------
// Row-per-lane epilogue for bf16 outputs (MODE 0): a lane keeps its accumulator row (32 columns) in registers, reads the
// auxiliary operand (residual / saved derivative) from a 32 x 32 bf16 tile the TMA unit prefetched while the main loop
// was still running, and writes its 64 output bytes into a 64-byte-swizzled shared-memory tile that one
// cp.async.bulk.tensor store sends to HBM: ~70 instructions per 32 x 32 chunk instead of ~190 (+ per-element math) for
// the transposing epilogue, which was issue bound on the K = 768 layers.  TMA clips rows >= M.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_src),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128u(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr) : "memory");
    return v;
}
// byte offset of 16-byte chunk c of row r in a [32 x 64 B] SWIZZLE_64B tile
__device__ __forceinline__ uint32_t sw64(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

// The accumulator chunk is read 8 columns at a time (a 576-thread CTA caps the kernel at 96 registers: with all 32
// accumulators live the exact-GELU / derivative epilogues kept ~58 words per thread in local memory); the TMEM stage is
// released after the last load.
template <int ACT>
__device__ __forceinline__ void epilogue_tma(uint32_t tmem_chunk, uint32_t acc_empty_bar, bool release, const Epilogue &ep,
                                             const CUtensorMap *map_c, const CUtensorMap *map_pre, uint32_t out_stg,
                                             uint32_t aux_stg, uint64_t *auxbar, uint32_t &aux_phase, long row0, int lane,
                                             int col0) {
    const long grow = row0 + lane;
    const uint32_t dkey = ep.drop_thresh ? drop_key0(ep.drop_seed) : 0u;
    if (ep.residual != nullptr) {
        mbar_wait(auxbar, aux_phase);
        aux_phase ^= 1;
    }
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {                      // 8 columns at a time, rolled: low register pressure, 4x less code
        uint32_t acc[8];
        tmem_ld8(tmem_chunk + c * 8, acc);
        if (c == 3 && release) {                       // the warp's last chunk has left TMEM: release the accumulator stage
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty_bar);
        }
        float x[8], g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(acc[j]);
        float a[8];
        if (ep.residual != nullptr) {
            const uint4 pk = lds128u(aux_stg + sw64(lane, c));
            const __nv_bfloat162 *h = (const __nv_bfloat162 *)&pk;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 t = __bfloat1622float2(h[q]);
                a[2 * q] = t.x;
                a[2 * q + 1] = t.y;
            }
        }
        if (ACT >= 5) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] *= act_grad<ACT>(a[j]) * ep.out_scale;
        } else {
            if (ep.bias != nullptr) {
                const float4 b0 = __ldg((const float4 *)(ep.bias + col0 + c * 8)), b1 = __ldg((const float4 *)(ep.bias + col0 + c * 8) + 1);
                x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
                x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
            }
            if (ACT == 3) {                             // exact GELU: packed pairs
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    float2 gg = make_float2(0.f, 0.f);
                    const float2 y = ep.preact != nullptr ? gelu2<true>(make_float2(x[j], x[j + 1]), gg, ep.out_scale)
                                                          : gelu2<false>(make_float2(x[j], x[j + 1]), gg, ep.out_scale);
                    x[j] = y.x; x[j + 1] = y.y;
                    g[j] = gg.x; g[j + 1] = gg.y;
                }
            } else if (ep.preact != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = act_with_grad<ACT>(x[j], g[j]) * ep.out_scale;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = fast_act<ACT>(x[j]) * ep.out_scale;
            }
        }
        if (ep.drop_thresh) {
            const uint32_t i0 = (uint32_t)(grow * ep.ldc + col0 + c * 8);   // < 2^32 (host check)
#pragma unroll
            for (int j = 0; j < 8; j += 2) {            // mask pair, one packed multiply
                const float2 mk = make_float2(drop_keep32(dkey, i0 + j, ep.drop_thresh) ? ep.drop_scale : 0.f,
                                              drop_keep32(dkey, i0 + j + 1, ep.drop_thresh) ? ep.drop_scale : 0.f);
                const float2 y = __fmul2_rn(make_float2(x[j], x[j + 1]), mk);
                x[j] = y.x; x[j + 1] = y.y;
            }
        }
        if (ACT < 5 && ep.residual != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = fmaf(a[j], ep.res_scale, x[j]);
        }
        __nv_bfloat162 p0 = __floats2bfloat162_rn(x[0], x[1]), p1 = __floats2bfloat162_rn(x[2], x[3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(x[4], x[5]), p3 = __floats2bfloat162_rn(x[6], x[7]);
        sts128(out_stg + sw64(lane, c), *(uint32_t *)&p0, *(uint32_t *)&p1, *(uint32_t *)&p2, *(uint32_t *)&p3);
        if (ACT < 5 && ep.preact != nullptr) {          // the derivative tile reuses this lane's own row of the aux tile
            p0 = __floats2bfloat162_rn(g[0], g[1]); p1 = __floats2bfloat162_rn(g[2], g[3]);
            p2 = __floats2bfloat162_rn(g[4], g[5]); p3 = __floats2bfloat162_rn(g[6], g[7]);
            sts128(aux_stg + sw64(lane, c), *(uint32_t *)&p0, *(uint32_t *)&p1, *(uint32_t *)&p2, *(uint32_t *)&p3);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(map_c, out_stg, col0, (int)row0);
        if (ACT < 5 && ep.preact != nullptr) tma_store_2d(map_pre, aux_stg, col0, (int)row0);
        tma_store_commit();
    }
}

// ---- lean TMA epilogue (TMA_EPI 2) -------------------------------------------------------------------------------------
// The 1x1 convolutions of the CNN are short contractions (K = 64..768) whose cost is the output write: the transposing
// epilogue above spends ~630 instructions per 32 x 32 chunk (ncu: profiles/r02_gemm_1x1_k256_plain_epilogue_full_raw.csv,
// issue bound at 2.3 TB/s).  This variant handles exactly what those layers need -- C = acc (+ residual), bf16, no bias /
// activation -- in ~40 instructions per chunk: one 32-column tcgen05.ld, 16 packed converts, four 16-byte stores into a
// 64-byte-swizzled tile and one TMA store.  Optional column statistics (training-mode BatchNorm): the fp32 chunk goes
// through a padded staging tile and every lane sums two columns over 16 rows with packed FADD2 / FFMA2 (~75 instructions).
__device__ __forceinline__ float2 lds64f(uint32_t saddr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts64f(uint32_t saddr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(saddr), "f"(v.x), "f"(v.y) : "memory");
}
// One staging region R per warp is reused in sequence: residual tile (TMA load) -> fp32 half tiles of the column reduction
// -> bf16 output tile (TMA store); 2560 B, so the kernel keeps the five-stage operand ring of the other TMA epilogue.
__device__ __forceinline__ void epilogue_lean(uint32_t tmem_chunk, uint32_t acc_empty_bar, bool release, const Epilogue &ep,
                                              const CUtensorMap *map_c, uint32_t R, float2 (&st)[2][2], uint64_t *auxbar,
                                              uint32_t &aux_phase, long row0, int lane, int col0) {
    uint32_t acc[32];
    tmem_ld32(tmem_chunk, acc);
    if (release) {                                      // the warp's last chunk has left TMEM: release the accumulator stage
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_bar);
    }
    if (ep.residual != nullptr) {                       // + residual tile (TMA load issued by the caller)
        mbar_wait(auxbar, aux_phase);
        aux_phase ^= 1;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint4 pk = lds128u(R + sw64(lane, c));
            const __nv_bfloat162 *h = (const __nv_bfloat162 *)&pk;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 t = __bfloat1622float2(h[q]);
                acc[c * 8 + 2 * q] = __float_as_uint(__uint_as_float(acc[c * 8 + 2 * q]) + t.x);
                acc[c * 8 + 2 * q + 1] = __float_as_uint(__uint_as_float(acc[c * 8 + 2 * q + 1]) + t.y);
            }
        }
        __syncwarp();                                   // every lane has read its row: R may be rewritten
    }
    if (ep.stats != nullptr) {
        // column sums / sums of squares of the fp32 chunk, 16 columns per pass through R (pitch 80 B: conflict-free 16-byte
        // stores); lane (g, jq) adds columns 2 jq, 2 jq + 1 of the eight rows 4 r + g to its running totals `st`, which stay
        // in registers for the whole kernel (folded across lanes once, after the last tile).  Rows past M are zero (TMA zero
        // fill of A and of the residual) and add nothing.
        const int g = lane >> 3, jq = lane & 7;
#pragma unroll
        for (int p = 0; p < 2; ++p) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                sts128(R + lane * 80 + j * 16, acc[16 * p + 4 * j], acc[16 * p + 4 * j + 1], acc[16 * p + 4 * j + 2], acc[16 * p + 4 * j + 3]);
            __syncwarp();
            float2 s1 = st[p][0], s2 = st[p][1];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float2 v = lds64f(R + (4 * r + g) * 80 + jq * 8);
                s1 = __fadd2_rn(s1, v);
                s2 = __ffma2_rn(v, v, s2);
            }
            st[p][0] = s1;
            st[p][1] = s2;
            __syncwarp();                               // R is rewritten by the next pass / the output tile
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        __nv_bfloat162 p[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
            p[q] = __floats2bfloat162_rn(__uint_as_float(acc[c * 8 + 2 * q]), __uint_as_float(acc[c * 8 + 2 * q + 1]));
        sts128(R + sw64(lane, c), *(uint32_t *)&p[0], *(uint32_t *)&p[1], *(uint32_t *)&p[2], *(uint32_t *)&p[3]);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(map_c, R, col0, (int)row0);
        tma_store_commit();
    }
}

// PERSISTENT kernel: grid = min(tiles, SMs); every CTA walks tiles t = blockIdx.x, += gridDim.x (N fastest so an A
// tile is reused from L2 by its N neighbours).  The TMA ring runs ahead across tile boundaries, the accumulator is
// double buffered in TMEM (2 x BN columns), so the epilogue of tile i overlaps the MMAs of tile i + 1.
// A_MN / B_MN: the operand is MN-major (its memory image is the transposed matrix [K, M] resp. [K, N], row-major) --
// backward GEMMs read activations and weights in place: dX = dY . W (B_MN) and dW = dY^T . X (A_MN and B_MN).
// Split-K: a work item is (tile, split); split s contracts k-blocks [s * kb_per, (s + 1) * kb_per) and the
// epilogue accumulates with red.global.add (ep.accumulate).
// TMA_EPI selects the epilogue at compile time (each instantiation carries only one of the two: the combined kernel was
// 28 K instructions and measurably slower on every path).
template <int BN, int kStages, int BKC, int MODE, int A_MN, int B_MN, int TMA_EPI, int CG2 = 0>
__global__ void __launch_bounds__(kGemmThreadsP, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_aux,
                    const __grid_constant__ CUtensorMap map_pre, int M, int N, int K, const Epilogue ep, const ConvGeom cg,
                    int m_tiles, int n_tiles, int kb_per, int k_splits) {
    using S = GemmSmem<BN, kStages, BKC, TMA_EPI, CG2>;
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);  // swizzle atoms: 1024 B
    unsigned char *bars = smem + kStages * S::kStageBytes;
    uint64_t *full = (uint64_t *)bars;
    uint64_t *empty = full + kStages;
    uint64_t *acc_full = empty + kStages;   // [2]
    uint64_t *acc_empty = acc_full + 2;     // [2]
    uint64_t *aux_bar = acc_empty + 2;      // [16]: auxiliary-tile TMA loads of the epilogue warps
    uint32_t *tmem_slot = (uint32_t *)(aux_bar + 16);
    unsigned char *staging = bars + S::kBarrierBytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = (K + BKC - 1) / BKC;
    const int mn_tiles = m_tiles * n_tiles;
    const int total_tiles = mn_tiles * k_splits;      // work items
    constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // power of two >= 32 (BN in {32, 64, 128})

    // CG2: the two CTAs of a cluster form a pair; rank 0 (the leader) issues the M = 256 MMAs for both.  `m_tiles` then counts
    // 256-row PAIR tiles, the pair walks them together and this CTA owns rows [mt2 * 256 + crank * 128, + 128) of each.
    const uint32_t crank = CG2 ? cluster_ctarank() : 0u;
    const int cta_id = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, cta_n = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full + s, CG2 ? 2 : 1);            // pair: both CTAs' producers arrive (with their byte counts) on the leader's
            mbar_init(empty + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(acc_full + s, 1);
            mbar_init(acc_empty + s, CG2 ? 2 * kEpiWarps : kEpiWarps);   // pair: the epilogue warps of both CTAs release the leader
        }
        for (int s = 0; s < 16; ++s) mbar_init(aux_bar + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CG2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                         : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG2) cluster_sync_all();                         // the peer's barriers are initialised before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // where the epilogue warps release an accumulator stage: the leader's acc_empty barriers
    const uint32_t acc_rel[2] = {CG2 ? mapa_u32(smem_u32(acc_empty), 0) : smem_u32(acc_empty),
                                 CG2 ? mapa_u32(smem_u32(acc_empty + 1), 0) : smem_u32(acc_empty + 1)};

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t it = 0;
            for (int item = cta_id; item < total_tiles; item += cta_n) {
                const int split = item / mn_tiles, tile = item - split * mn_tiles;
                const int kb0 = split * kb_per, kb1 = min(num_kb, kb0 + kb_per);
                const int mt2 = tile / n_tiles, n0 = (tile - mt2 * n_tiles) * BN;
                const int mt = CG2 ? mt2 * 2 + (int)crank : mt2;
                // Every index below advances incrementally: this single thread issues all loads of the CTA, and the runtime
                // integer divisions that used to decode (tap, channel chunk) / (patch column, row, image) from the block
                // number on every k-block were the critical path of the convolution kernels.
                int img = 0, oh0 = 0, ow0 = 0;
                int f_cc = 0, f_kw = 0, f_kh = 0;                       // MODE 1: channel chunk / filter tap of block kb
                int p_tw = 0, p_th = 0, p_img = 0;                      // MODE 2: 64-pixel patch of block kb
                int w_c0[4], w_dx[4], w_dy[4];                          // MODE 2: the B blocks of this column tile
                if (MODE == 1) {
                    const int tiles_w = cg.tiles_w, tiles_h = cg.tiles_h;
                    int t = mt;
                    const int tw = t % tiles_w;
                    t /= tiles_w;
                    const int th = t % tiles_h;
                    img = (t / tiles_h) * cg.TN;
                    oh0 = th * cg.TH * cg.stride - cg.pad;
                    ow0 = tw * cg.TW * cg.stride - cg.pad;
                    const int tap = kb0 / cg.cchunks;
                    f_cc = kb0 - tap * cg.cchunks;
                    f_kh = tap / cg.KW;
                    f_kw = tap - f_kh * cg.KW;
                }
                if (MODE == 2) {
                    int t = kb0;
                    p_tw = t % cg.tiles_w;
                    t /= cg.tiles_w;
                    p_th = t % cg.tiles_h;
                    p_img = (t / cg.tiles_h) * cg.TN;
                    const int cw = cg.cw;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int jb = n0 / cw + j;
                        const int tap = jb / cg.cchunks, cc = jb - tap * cg.cchunks;
                        const int kh = tap / cg.KW, kw = tap - kh * cg.KW;
                        // past the last tap: a box that is entirely out of bounds (zero filled, full byte count)
                        w_c0[j] = tap < cg.taps ? cc * cw : cg.cchunks * cw;
                        w_dx[j] = kw * cg.dil - cg.pad;
                        w_dy[j] = kh * cg.dil - cg.pad;
                    }
                }
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1;
                    mbar_wait(empty + s, ph ^ 1);
                    unsigned char *sa = smem + s * S::kStageBytes, *sb = sa + S::kABytes;
                    if (CG2) {
                        // CTA pair: this CTA's A rows and its half of the B tile; the bytes are counted on the leader's barrier
                        const uint32_t lf = mapa_u32(smem_u32(full + s), 0);
                        mbar_expect_tx_cluster(lf, S::kStageBytes);
                        if (A_MN) {
                            tma_load_2d_2sm(sa, &map_a, lf, mt * BM, kb * BKC);
                            tma_load_2d_2sm(sa + BKC * 128, &map_a, lf, mt * BM + 64, kb * BKC);
                        } else {
                            tma_load_2d_2sm(sa, &map_a, lf, kb * BKC, mt * BM);
                        }
                        if (B_MN) {
#pragma unroll
                            for (int j = 0; j < BN / 128; ++j)
                                tma_load_2d_2sm(sb + j * BKC * 128, &map_w, lf, n0 + ((int)crank * (BN / 128) + j) * 64, kb * BKC);
                        } else {
                            tma_load_2d_2sm(sb, &map_w, lf, kb * BKC, n0 + (int)crank * (BN / 2));
                        }
                        continue;
                    }
                    mbar_expect_tx(full + s, S::kStageBytes);
                    if (MODE == 2) {
                        // convolution weight gradient: the contraction block is a patch of 64 OUTPUT PIXELS; A = dY^T
                        // (two boxes of 64 output channels), B = the input activation at one filter tap per block of cw
                        // columns (4-D boxes, zero fill outside the image = the convolution padding)
                        const int poh = p_th * cg.TH, pow_ = p_tw * cg.TW;
                        tma_load_4d(sa, &map_a, full + s, mt * BM, pow_, poh, p_img);
                        tma_load_4d(sa + 64 * 128, &map_a, full + s, mt * BM + 64, pow_, poh, p_img);
                        const int bstride = 64 * cg.cw * 2, nblk = BN * 128 / bstride;    // 2 blocks of 64 channels or 4 of 32
                        const int px = pow_ * cg.stride, py = poh * cg.stride;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j < nblk) tma_load_4d(sb + j * bstride, &map_w, full + s, w_c0[j], px + w_dx[j], py + w_dy[j], p_img);
                        if (++p_tw == cg.tiles_w) {
                            p_tw = 0;
                            if (++p_th == cg.tiles_h) {
                                p_th = 0;
                                p_img += cg.TN;
                            }
                        }
                    } else if (MODE == 0 && A_MN) {
                        // two boxes of 64 MN elements x BKC contraction rows
                        tma_load_2d(sa, &map_a, full + s, mt * BM, kb * BKC);
                        tma_load_2d(sa + BKC * 128, &map_a, full + s, mt * BM + 64, kb * BKC);
                    } else if (MODE == 0) {
                        tma_load_2d(sa, &map_a, full + s, kb * BKC, mt * BM);
                    } else {
                        tma_load_4d(sa, &map_a, full + s, f_cc * BKC, ow0 + f_kw * cg.dil, oh0 + f_kh * cg.dil, img);
                        if (++f_cc == cg.cchunks) {
                            f_cc = 0;
                            if (++f_kw == cg.KW) {
                                f_kw = 0;
                                ++f_kh;
                            }
                        }
                    }
                    if (MODE == 2) {
                    } else if (B_MN) {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            tma_load_2d(sb + j * BKC * 128, &map_w, full + s, n0 + j * 64, kb * BKC);
                    } else {
                        tma_load_2d(sb, &map_w, full + s, kb * BKC, n0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = umma_idesc_bf16(CG2 ? 2 * BM : BM, BN, A_MN, B_MN);
        uint32_t it = 0, tile_iter = 0;
        for (int item = (CG2 && crank != 0) ? total_tiles : cta_id; item < total_tiles; item += cta_n, ++tile_iter) {   // pair: leader only
            const int split = item / mn_tiles;
            const int kb0 = split * kb_per, kb1 = min(num_kb, kb0 + kb_per);
            const uint32_t as = tile_iter & 1, aph = (tile_iter >> 1) & 1;
            mbar_wait(acc_empty + as, aph ^ 1);   // the epilogue has drained this accumulator stage
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + as * BN;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                mbar_wait(full + s, ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sa = smem_u32(smem + s * S::kStageBytes), sb = sa + S::kABytes;
                    const uint64_t da = A_MN ? umma_desc_mn(sa, BKC * 128) : umma_desc_k<BKC * 2>(sa);
                    const bool b32 = MODE == 2 && cg.cw == 32;          // 32-channel pixels: 64-byte rows, SWIZZLE_64B
                    const uint64_t db = b32 ? umma_desc_mn_sw64(sb, 64 * 64)
                                            : (B_MN ? umma_desc_mn(sb, BKC * 128) : umma_desc_k<BKC * 2>(sb));
                    // advancing K by 16: K-major = 32 B inside the swizzled row (+2 in the >>4 address field);
                    // MN-major = 16 contraction rows of 128 B (+128) or of 64 B (+64)
                    constexpr uint64_t ka = A_MN ? 128 : 2;
                    const uint64_t kbs = b32 ? 64 : (B_MN ? 128 : 2);
#pragma unroll
                    for (int k = 0; k < BKC / UMMA_K; ++k) {
                        if (CG2)
                            tc_mma_f16_2sm(tmem_acc, da + (uint64_t)k * ka, db + (uint64_t)k * kbs, idesc, ((kb - kb0) | k) ? 1u : 0u);
                        else
                            tc_mma_f16(tmem_acc, da + (uint64_t)k * ka, db + (uint64_t)k * kbs, idesc, ((kb - kb0) | k) ? 1u : 0u);
                    }
                    if (CG2) {
                        tc_commit_2sm(empty + s, 3);                          // frees this stage in BOTH CTAs
                        if (kb == kb1 - 1) tc_commit_2sm(acc_full + as, 3);   // accumulator complete (each CTA reads its own TMEM)
                    } else {
                        tc_commit(empty + s);                          // frees the smem stage when these MMAs retire
                        if (kb == kb1 - 1) tc_commit(acc_full + as);   // accumulator complete
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: warps 2..17; TMEM lane quarter = warp % 4, the two warps of a quarter split the columns =====
        // A warp's accumulator fragment is row-per-lane; storing that directly scatters every 16 B to a different
        // 128 B line.  It is transposed through a private shared-memory buffer instead, so that bias / activation /
        // residual / convert / store all run with 8 lanes per output row (4 full rows per instruction).
        const int quarter = warp & 3, cgrp = (warp - 2) >> 2;   // cgrp 0..3: which column chunks this warp owns
        const uint32_t stg = smem_u32(staging + (warp - 2) * S::kWarpStaging);
        const int colq = (lane & 7) * 4;  // phase-2 mapping: 8 lanes per row, 4 columns per lane
        uint32_t tile_iter = 0, aux_phase = 0;
        // column statistics (ep.stats): the warp's running totals live behind its transpose buffer (2 chunks x 2 x 32 fp32);
        // the host sizes the grid as a multiple of n_tiles, so every tile of this CTA covers the same columns
        const uint32_t wstat0 = stg + 32 * S::kStagePitch;
        float2 lst[(BN / 32 + 3) / 4][2][2];             // TMA_EPI 2: per-lane running column sums [chunk][pass][sum, sum sq]
#pragma unroll
        for (int i = 0; i < (BN / 32 + 3) / 4; ++i)
#pragma unroll
            for (int p = 0; p < 2; ++p) lst[i][p][0] = lst[i][p][1] = make_float2(0.f, 0.f);
        if (TMA_EPI == 0 && ep.stats != nullptr) {
            sts128(wstat0 + lane * 16, 0u, 0u, 0u, 0u);
            __syncwarp();
        }
        for (int item = cta_id; item < total_tiles; item += cta_n, ++tile_iter) {
            const int tile = item % mn_tiles;
            const uint32_t as = tile_iter & 1, aph = (tile_iter >> 1) & 1;
            const int mt2 = tile / n_tiles, n0 = (tile - mt2 * n_tiles) * BN;
            const int mt = CG2 ? mt2 * 2 + (int)crank : mt2;
            const int r = quarter * 32 + lane;
            if constexpr (TMA_EPI == 2) {
                // ---- lean plain epilogue (see epilogue_lean) ----
                constexpr int kMineT = (BN / 32 + 3) / 4;
                const long row0 = (long)mt * BM + quarter * 32;
                bool waited = false;
#pragma unroll
                for (int ci = 0; ci < kMineT; ++ci) {                   // unrolled: lst[ci] must stay in registers
                    const int c = cgrp + 4 * ci;
                    const int col0 = n0 + c * 32;
                    const bool mine = c < BN / 32 && col0 < N;          // warp-uniform
                    const bool last = ci == kMineT - 1;
                    if (lane == 0 && mine) {
                        tma_store_wait_read();                          // the previous store has left the output tile
                        if (ep.residual != nullptr) {
                            mbar_expect_tx(aux_bar + (warp - 2), 32 * 64);
                            tma_load_2d((void *)(staging + (warp - 2) * S::kWarpStaging), &map_aux, aux_bar + (warp - 2), col0, (int)row0);
                        }
                    }
                    __syncwarp();
                    if (!waited) {
                        mbar_wait(acc_full + as, aph);
                        tc_fence_after();
                        waited = true;
                    }
                    const uint32_t chunk = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + (uint32_t)(c * 32);
                    if (mine) {
                        epilogue_lean(chunk, acc_rel[as], last, ep, &map_c, stg, lst[ci], aux_bar + (warp - 2), aux_phase, row0, lane,
                                      col0);
                    } else if (last) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(acc_rel[as]);
                    }
                }
            } else if constexpr (TMA_EPI != 0) {
                // ---- row-per-lane epilogue with TMA stores: chunk c = cgrp + 4 * ci of the tile (one per warp at BN = 128, two at
                //      BN = 256; the warp's staging tiles are reused chunk after chunk) ----
                constexpr int kMineT = (BN / 32 + 3) / 4;
                const long row0 = (long)mt * BM + quarter * 32;
                bool waited = false;
#pragma unroll 1
                for (int ci = 0; ci < kMineT; ++ci) {
                    const int c = cgrp + 4 * ci;
                    const int col0 = n0 + c * 32;
                    const bool mine = c < BN / 32 && col0 < N;          // warp-uniform
                    const bool last = ci == kMineT - 1;
                    if (lane == 0 && mine) {
                        tma_store_wait_read();                          // the previous stores have left both staging tiles
                        if (ep.residual != nullptr) {
                            mbar_expect_tx(aux_bar + (warp - 2), 32 * 64);
                            tma_load_2d((void *)(staging + (warp - 2) * S::kWarpStaging + 2048), &map_aux, aux_bar + (warp - 2), col0,
                                        (int)row0);
                        }
                    }
                    __syncwarp();
                    if (!waited) {
                        mbar_wait(acc_full + as, aph);
                        tc_fence_after();
                        waited = true;
                    }
                    const uint32_t chunk = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + (uint32_t)(c * 32);
                    if (mine) {
                        switch (ep.act) {
#define EPI_TMA(A_) case A_: epilogue_tma<A_>(chunk, acc_rel[as], last, ep, &map_c, &map_pre, stg, stg + 2048, aux_bar + (warp - 2), aux_phase, row0, lane, col0); break;
                            EPI_TMA(0) EPI_TMA(1) EPI_TMA(2) EPI_TMA(3) EPI_TMA(4) EPI_TMA(5) EPI_TMA(6)
                            default: epilogue_tma<7>(chunk, acc_rel[as], last, ep, &map_c, &map_pre, stg, stg + 2048, aux_bar + (warp - 2), aux_phase, row0, lane, col0); break;
#undef EPI_TMA
                        }
                    } else if (last) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(acc_rel[as]);
                    }
                }
            } else {
            long row;       // global output row of tile row r (this lane's accumulator row)
            if (MODE != 1) {
                row = (long)mt * BM + r;
                if (row >= M) row = -1;
            } else {
                const int tiles_w = cg.tiles_w, tiles_h = cg.tiles_h;
                int t = mt;
                const int tw = t % tiles_w;
                t /= tiles_w;
                const int th = t % tiles_h;
                const int img = (t / tiles_h) * cg.TN;
                const int per_img = cg.TH * cg.TW;
                const int dn = r / per_img, rr = r - dn * per_img;
                const int dh = rr / cg.TW, dw = rr - dh * cg.TW;
                row = ((long)(img + dn) * cg.Ho + th * cg.TH + dh) * cg.Wo + tw * cg.TW + dw;
                if (img + dn >= cg.Nimg || th * cg.TH + dh >= cg.Ho || tw * cg.TW + dw >= cg.Wo) row = -1;
            }
            mbar_wait(acc_full + as, aph);
            tc_fence_after();
            constexpr int kChunks = BN / 32;
            constexpr int kMine = (kChunks + 3) / 4;          // chunks per warp: c = cgrp + 4 * ci
            uint32_t acc[kMine][32];
#pragma unroll
            for (int ci = 0; ci < kMine; ++ci) {
                const int c = cgrp + 4 * ci;
                if (c < kChunks) tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + (uint32_t)(c * 32), acc[ci]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_rel[as]);
#pragma unroll
            for (int ci = 0; ci < kMine; ++ci) {
                const int c = cgrp + 4 * ci;
                const int col0 = n0 + c * 32;
                if (c >= kChunks || col0 >= N) continue;   // warp-uniform
                // phase 1: raw fp32 accumulators, row-per-lane -> staging
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    sts128(stg + lane * S::kStagePitch + j * 4, acc[ci][j], acc[ci][j + 1], acc[ci][j + 2], acc[ci][j + 3]);
                __syncwarp();
                // phase 2: 8 lanes per row (warp-uniform dispatch on the activation)
                const int col = col0 + colq;
                const bool fast = ep.vec && col0 + 32 <= N;   // warp-uniform
                if (!fast) epilogue_rows_slow<S::kStagePitch>(stg, ep, row, lane, col, N);
                else switch (ep.act) {
                    case 0: epilogue_rows<0, S::kStagePitch>(stg, ep, row, lane, col, N, wstat0 + ci * 256); break;
                    case 1: epilogue_rows<1, S::kStagePitch>(stg, ep, row, lane, col, N, wstat0 + ci * 256); break;
                    case 2: epilogue_rows<2, S::kStagePitch>(stg, ep, row, lane, col, N, wstat0 + ci * 256); break;
                    case 3: epilogue_rows<3, S::kStagePitch>(stg, ep, row, lane, col, N, wstat0 + ci * 256); break;
                    case 4: epilogue_rows<4, S::kStagePitch>(stg, ep, row, lane, col, N, wstat0 + ci * 256); break;
                    case 5: epilogue_rows<5, S::kStagePitch>(stg, ep, row, lane, col, N, wstat0 + ci * 256); break;
                    case 6: epilogue_rows<6, S::kStagePitch>(stg, ep, row, lane, col, N, wstat0 + ci * 256); break;
                    default: epilogue_rows<7, S::kStagePitch>(stg, ep, row, lane, col, N, wstat0 + ci * 256); break;
                }
                __syncwarp();   // staging is rewritten by the next chunk
            }
            }   // !TMA_EPI
        }
        if (TMA_EPI != 1 && ep.stats != nullptr) {
            // partial row (blockIdx / n_tiles) * 4 + quarter of [parts, 2, N]: every (row, column) has exactly one writer
            __syncwarp();
            const int n0 = (blockIdx.x % n_tiles) * BN;
            float *dst = ep.stats + (long)((blockIdx.x / n_tiles) * 4 + quarter) * 2 * N;
#pragma unroll
            for (int ci = 0; ci < (BN / 32 + 3) / 4; ++ci) {
                const int c = cgrp + 4 * ci;
                const int col0 = n0 + c * 32;
                if (TMA_EPI == 2) {
                    // fold the four row groups (lanes jq, jq + 8, jq + 16, jq + 24) in a fixed order
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
#define FOLD_(v_) v_ += __shfl_xor_sync(0xffffffffu, v_, 8); v_ += __shfl_xor_sync(0xffffffffu, v_, 16);
                        FOLD_(lst[ci][p][0].x) FOLD_(lst[ci][p][0].y) FOLD_(lst[ci][p][1].x) FOLD_(lst[ci][p][1].y)
#undef FOLD_
                        if (lane < 8 && c < BN / 32 && col0 < N) {
                            *(float2 *)(dst + col0 + p * 16 + 2 * lane) = lst[ci][p][0];
                            *(float2 *)(dst + N + col0 + p * 16 + 2 * lane) = lst[ci][p][1];
                        }
                    }
                } else if (c < BN / 32 && col0 < N) {
                    dst[col0 + lane] = lds32(wstat0 + ci * 256 + lane * 4);
                    dst[N + col0 + lane] = lds32(wstat0 + ci * 256 + 128 + lane * 4);
                }
            }
        }
    }
    if (TMA_EPI != 0 && warp >= 2 && lane == 0) tma_store_wait_all();      // bulk stores of this thread have completed
    tc_fence_before();
    __syncthreads();
    if (CG2) cluster_sync_all();          // the leader's MMAs read the peer's shared memory and write its TMEM: leave together
    if (warp == 1) {
        if (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// fp32 -> bf16 (operand preparation for the GEMM); 8 elements per thread
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float *__restrict__ in, __nv_bfloat16 *__restrict__ out, long n) {
    const long n8 = n >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        const float4 a = __ldg((const float4 *)in + i * 2), b = __ldg((const float4 *)in + i * 2 + 1);
        __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
        ((uint4 *)out)[i] = make_uint4(*(uint32_t *)&p0, *(uint32_t *)&p1, *(uint32_t *)&p2, *(uint32_t *)&p3);
    }
    for (long i = n8 * 8 + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}

// -------------------------------------------------------------------------------------------- host
// 2-D bf16 row-major [rows, cols] tensor, box = box_rows x bkc columns, swizzle = row bytes, OOB reads give zeros
static int make_map_2d(CUtensorMap *map, const void *ptr, long rows, long cols, long ld_elems, int box_rows, int bkc) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return POSE_E_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)bkc, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bkc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? POSE_OK : POSE_E_SHAPE;
}

// 4-D NHWC bf16 activation [N, H, W, C]; box = bkc channels x (TW, TH) output pixels traversed with the conv stride
static int make_map_nhwc(CUtensorMap *map, const void *ptr, int Nimg, int H, int W, int C, int bkc, int TW, int TH,
                         int TN, int stride) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return POSE_E_UNSUPPORTED;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)bkc, (cuuint32_t)(TW * stride), (cuuint32_t)(TH * stride), (cuuint32_t)TN};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    if (box[1] > 256 || box[2] > 256) return POSE_E_UNSUPPORTED;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bkc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? POSE_OK : POSE_E_SHAPE;
}

// bf16 [rows, cols] tensor (pitch ld elements), box = 32 x 32, 64-byte swizzle: the epilogue's output / auxiliary tiles
static int make_map_tile32(CUtensorMap *map, const void *ptr, long rows, long cols, long ld) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return POSE_E_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? POSE_OK : POSE_E_SHAPE;
}

template <int BN, int kStages, int BKC, int MODE, int A_MN = 0, int B_MN = 0>
static int launch_gemm(const CUtensorMap &ma, const CUtensorMap &mw, int M, int N, int K, const Epilogue &ep_in,
                       const ConvGeom &cg, int m_tiles, cudaStream_t s, int k_splits = 1, bool pair = false) {
    Epilogue ep = ep_in;
    CUtensorMap mc = {}, maux = {}, mpre = {};
    // Epilogue choice.  Round 1 measured the row-per-lane TMA epilogue ahead only for epilogues with real per-element work
    // (exact GELU, saved derivative, multiply-by-derivative) and ~10 % behind on plain / SiLU / residual-only ones; with the
    // operand ring one stage deeper it now wins on every bf16 epilogue of the ViT step (25.6 -> 25.0 ms with it everywhere:
    // the transposing epilogue costs ~630 instructions per 32 x 32 chunk), so it is the default for bf16 outputs; the
    // transposing epilogue remains for fp32 / accumulating / ragged outputs and the convolution modes.
    ep.tma = 0;
    static const bool tma_off = getenv("POSE_NO_TMA_EPILOGUE") != nullptr;     // A/B switches for measurements
    static const bool tma_heavy_only = getenv("POSE_TMA_EPILOGUE_HEAVY_ONLY") != nullptr;
    const bool heavy = ep.act == 3 || ep.act >= 5 || ep.preact != nullptr || !tma_heavy_only;
    const bool lean_shape = ep.act == 0 && ep.bias == nullptr && ep.preact == nullptr && ep.drop_thresh == 0 &&
                            ep.out_scale == 1.0f && (ep.residual == nullptr || ep.res_scale == 1.0f) && k_splits <= 1;
    // the TMA-store epilogues address output rows linearly: the plain GEMM, and convolutions whose patches are linear
    static const bool conv_tma_off = getenv("POSE_NO_CONV_TMA_EPILOGUE") != nullptr;      // A/B switch for measurements
    const bool rows_linear = MODE == 0 || (MODE == 1 && cg.linear && !conv_tma_off);
    if (rows_linear && ep.out_bf16 && !ep.accumulate && ep.vec && N % 32 == 0 && heavy && !tma_off && !lean_shape &&
        ep.stats == nullptr) {
        int e = make_map_tile32(&mc, ep.C, M, N, ep.ldc);
        if (!e && ep.residual) e = make_map_tile32(&maux, ep.residual, M, N, ep.ldr);
        if (!e && ep.preact) e = make_map_tile32(&mpre, ep.preact, M, N, ep.ldc);
        ep.tma = e ? 0 : 1;
    }
    constexpr int kTma = (MODE == 0 || MODE == 1) ? 1 : 0;
    constexpr int kStagesT = (kTma && BN == 128) ? kStages + 1 : kStages;      // 5 x 32 KB stages fit beside the TMA staging
    // lean plain epilogue (TMA_EPI 2; MODE 0, tiles up to 128 columns): bf16 C = acc (+ residual), optional column
    // statistics (tools/bench_gemm_cnn.py measures it on the CNN's 1x1 shapes).
    constexpr bool kHasLean = MODE == 0 || MODE == 1;
    constexpr int kLean = kHasLean ? 2 : 0;
    constexpr int kStagesL = kStagesT;
    using S0 = GemmSmem<BN, kStages, BKC, 0>;
    using S1 = GemmSmem<BN, kStagesT, BKC, kTma>;
    using S2 = GemmSmem<BN, kStagesL, BKC, kLean>;
    static_assert(S0::kTotal <= 232448 && S1::kTotal <= 232448 && S2::kTotal <= 232448, "shared memory budget");
    auto kern_plain = gemm_bf16_tn_kernel<BN, kStages, BKC, MODE, A_MN, B_MN, 0>;
    auto kern_tma = gemm_bf16_tn_kernel<BN, kStagesT, BKC, MODE, A_MN, B_MN, kTma>;
    auto kern_lean = gemm_bf16_tn_kernel<BN, kStagesL, BKC, MODE, A_MN, B_MN, kLean>;
    static bool configured = false;
    if (!configured) {
        cudaError_t ce = cudaFuncSetAttribute(kern_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, S0::kTotal);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(kern_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, S1::kTotal);
        if (ce == cudaSuccess) ce = cudaFuncSetAttribute(kern_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, S2::kTotal);
        if (ce != cudaSuccess) return (int)ce;
        configured = true;
    }
    static const bool lean_off = getenv("POSE_NO_LEAN_EPILOGUE") != nullptr;   // A/B switch for measurements
    bool lean = false;
    if (kHasLean && rows_linear && !lean_off && !ep.tma && ep.out_bf16 && !ep.accumulate && ep.vec && N % 32 == 0 && lean_shape) {
        int e = make_map_tile32(&mc, ep.C, M, N, ep.ldc);
        if (!e && ep.residual) e = make_map_tile32(&maux, ep.residual, M, N, ep.ldr);
        lean = !e;
    }
    auto kern = lean ? kern_lean : (ep.tma ? kern_tma : kern_plain);
    const int smem_bytes = lean ? S2::kTotal : (ep.tma ? S1::kTotal : S0::kTotal);
    const int n_tiles = (N + BN - 1) / BN;
    const int num_kb = (K + BKC - 1) / BKC;
    if (k_splits < 1) k_splits = 1;
    if (k_splits > num_kb) k_splits = num_kb;
    const int kb_per = (num_kb + k_splits - 1) / k_splits;
    k_splits = (num_kb + kb_per - 1) / kb_per;            // no empty split
    const long total = (long)m_tiles * n_tiles * k_splits;
    int grid = (int)(total < kNumSMs ? total : kNumSMs);
    if (ep.stats != nullptr) {
        // column statistics: every CTA must see ONE column tile (grid a multiple of n_tiles; tiles are N-fastest)
        if (ep.tma || k_splits != 1 || n_tiles > grid || N % 32) return POSE_E_UNSUPPORTED;
        grid = grid / n_tiles * n_tiles;
        const int parts = grid / n_tiles * 4;
        if (ep.bn->cap_floats < (int64_t)parts * 2 * N) return POSE_E_WORKSPACE;
        kern<<<grid, kGemmThreadsP, smem_bytes, s>>>(ma, mw, mc, maux, mpre, M, N, K, ep, cg, m_tiles, n_tiles, kb_per, k_splits);
        int e = launch_status();
        if (e) return e;
        const pose_bn_fuse *bn = ep.bn;
        return launch_bn_finalize_parts(bn->partials, parts, (long)bn->count, bn->gamma, bn->beta, bn->eps, bn->momentum, N,
                                        bn->mean_rstd, bn->scale_shift, bn->running_mean, bn->running_var, s);
    }
    if constexpr (MODE == 0 && BN == 256) {
        if (pair) {
            // CTA pairs (cta_group::2): clusters of two CTAs, each pair walks 256-row tiles; B tile halved per CTA, so the
            // operand ring is 32 KB per stage (4 stages beside the transposing epilogue's staging, 5 beside the TMA ones)
            using P0 = GemmSmem<BN, 4, BKC, 0, 1>;
            using P1 = GemmSmem<BN, 5, BKC, 1, 1>;
            static_assert(P0::kTotal <= 232448 && P1::kTotal <= 232448, "shared memory budget");
            auto pk_plain = gemm_bf16_tn_kernel<BN, 4, BKC, MODE, A_MN, B_MN, 0, 1>;
            auto pk_tma = gemm_bf16_tn_kernel<BN, 5, BKC, MODE, A_MN, B_MN, 1, 1>;
            auto pk_lean = gemm_bf16_tn_kernel<BN, 5, BKC, MODE, A_MN, B_MN, 2, 1>;
            static bool pconf = false;
            if (!pconf) {
                cudaError_t ce = cudaFuncSetAttribute(pk_plain, cudaFuncAttributeMaxDynamicSharedMemorySize, P0::kTotal);
                if (ce == cudaSuccess) ce = cudaFuncSetAttribute(pk_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, P1::kTotal);
                if (ce == cudaSuccess) ce = cudaFuncSetAttribute(pk_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, P1::kTotal);
                if (ce != cudaSuccess) return (int)ce;
                pconf = true;
            }
            auto pk = lean ? pk_lean : (ep.tma ? pk_tma : pk_plain);
            const int m_pairs = (m_tiles + 1) / 2;
            const long ptotal = (long)m_pairs * n_tiles * k_splits;
            const int pairs = (int)(ptotal < kNumSMs / 2 ? ptotal : kNumSMs / 2);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * pairs);
            cfg.blockDim = dim3(kGemmThreadsP);
            cfg.dynamicSmemBytes = lean || ep.tma ? P1::kTotal : P0::kTotal;
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2;
            at[0].val.clusterDim.y = 1;
            at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            cudaError_t ce = cudaLaunchKernelEx(&cfg, pk, ma, mw, mc, maux, mpre, M, N, K, ep, cg, m_pairs, n_tiles, kb_per, k_splits);
            return ce == cudaSuccess ? launch_status() : (int)ce;
        }
    }
    kern<<<grid, kGemmThreadsP, smem_bytes, s>>>(ma, mw, mc, maux, mpre, M, N, K, ep, cg, m_tiles, n_tiles, kb_per, k_splits);
    return launch_status();
}

// 128 x 256 tiles (3 stages of 48 KB): the 128 x 128 main loop reads 32 KB of operands per 256 tensor-core cycles, which is
// the whole 128 B/clk shared-memory read bandwidth of the SM; the wide tile needs 96 B/clk.  Used for plain epilogues
// and for the TMA-store epilogue of the GELU / derivative GEMMs (two chunks per warp) when there are enough tiles to go round.
static bool wide_tile_ok(const Epilogue &ep, int N, int K, int m_tiles, int k_splits) {
    static const bool off = getenv("POSE_GEMM_NO_BN256") != nullptr;      // A/B switch for measurements
    // Shortest contraction that takes the wide tile.  Round 1 set 512 ("short contractions are output-write bound") with the
    // transposing epilogue; with the lean TMA-store epilogue the K = 256 layers are bound by the L2 -> shared-memory fills
    // instead, and a 128 x 256 tile needs 25 % fewer of them per output: 131072 x 768 x 256 with BatchNorm statistics
    // 136 -> 93 us, CNN step 22.39 -> 22.16 ms (128: no further change).  POSE_GEMM_WIDE_MIN_K overrides (A/B).
    static const int min_k = getenv("POSE_GEMM_WIDE_MIN_K") ? atoi(getenv("POSE_GEMM_WIDE_MIN_K")) : 256;
    if (off || N % 256 || K < min_k) return false;

    static const bool heavy_narrow = getenv("POSE_GEMM_HEAVY_BN128") != nullptr;      // A/B switch for measurements
    const bool heavy = ep.act == 3 || ep.act >= 5 || ep.preact != nullptr;
    if (heavy_narrow && heavy && ep.out_bf16 && !ep.accumulate && ep.vec) return false;
    return (long)m_tiles * (N / 256) * k_splits >= kNumSMs;
}
// short contractions with per-element epilogue work (activation, residual read) are epilogue bound: two 32-column chunks
// per epilogue warp lose to the 128-column tile there (measured: 37.0 vs 34.4 us at K = 768 with a residual)
static bool wide_tile_fwd_ok(const Epilogue &ep, int N, int K, int m_tiles) {
    const bool tma_heavy = (ep.act == 3 || ep.act >= 5 || ep.preact != nullptr) && ep.out_bf16 && !ep.accumulate && ep.vec;
    // (round 1 kept activation / residual epilogues on 128-column tiles below K = 1024; re-measured at the end of round 2
    // with the TMA-store epilogues: CNN eval forward at B = 512 22.24 -> 21.35 ms with the wide tile from K = 256, ViT step
    // 23.41 -> 23.32 ms.  POSE_GEMM_WIDE_ACT_MIN_K overrides, A/B)
    static const int act_min_k = getenv("POSE_GEMM_WIDE_ACT_MIN_K") ? atoi(getenv("POSE_GEMM_WIDE_ACT_MIN_K")) : 256;
    if (!tma_heavy && (ep.act != 0 || ep.residual != nullptr) && K < act_min_k) return false;
    return wide_tile_ok(ep, N, K, m_tiles, 1);
}

// CTA pairs pay off where the main loop is bound by operand traffic (L2 requests / shared-memory fill): large tiles counts,
// long contractions.  POSE_GEMM_PAIR=0 disables them (A/B switch for measurements).
static bool pair_ok(const Epilogue &ep, int M, int N, int K, int k_splits) {
    static const bool on = getenv("POSE_GEMM_PAIR") == nullptr || atoi(getenv("POSE_GEMM_PAIR")) != 0;
    if (!on || ep.stats != nullptr || N % 256 || M < 256) return false;
    const long pair_items = (long)((M + 255) / 256) * (N / 256) * (k_splits > 1 ? k_splits : 1);
    return pair_items >= kNumSMs / 4;
}

template <int BKC, int MODE>
static int dispatch_bn(const CUtensorMap &ma, const void *W, int ldw, int M, int N, int K, const Epilogue &ep,
                       const ConvGeom &cg, int m_tiles, cudaStream_t s) {
    CUtensorMap mw;
    int bn = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    // 128 x 256 tiles: the plain GEMM, and the implicit-GEMM convolutions with long contractions (the 512-channel dilated
    // 3x3 layers: K = 4608) -- POSE_CONV_NO_BN256 switches the latter off (A/B measurements)
    static const bool conv_wide_off = getenv("POSE_CONV_NO_BN256") != nullptr;
    if ((MODE == 0 || (MODE == 1 && !conv_wide_off && K >= 1024)) && BKC == 64 && wide_tile_fwd_ok(ep, N, K, m_tiles)) bn = 256;
    const bool pair = MODE == 0 && bn == 256 && pair_ok(ep, M, N, K, 1);
    int e = make_map_2d(&mw, W, N, K, ldw, pair ? bn / 2 : bn, BKC);        // a CTA of a pair loads half of the B tile
    if (e) return e;
    if constexpr ((MODE == 0 || MODE == 1) && BKC == 64) {
        if (bn == 256) return launch_gemm<256, 3, BKC, MODE>(ma, mw, M, N, K, ep, cg, m_tiles, s, 1, pair);
    }
    if (bn == 32) return launch_gemm<32, 6, BKC, MODE>(ma, mw, M, N, K, ep, cg, m_tiles, s);
    if (bn == 64) return launch_gemm<64, 6, BKC, MODE>(ma, mw, M, N, K, ep, cg, m_tiles, s);
    return launch_gemm<128, 4, BKC, MODE>(ma, mw, M, N, K, ep, cg, m_tiles, s);
}

static int check_epilogue(const pose_gemm_epilogue *e, long M, int N, Epilogue &ep) {
    if (!e || !e->C) return POSE_E_NULL;
    if (e->drop_p > 0.f && (double)M * (double)e->ldc >= 4294967296.0) return POSE_E_UNSUPPORTED;   // 32-bit mask counter
    if (e->ldc < N || (e->residual && e->ldr < N)) return POSE_E_SHAPE;
    if (e->act < 0 || e->act > 7 || (e->out_dtype != 0 && e->out_dtype != 1)) return POSE_E_UNSUPPORTED;
    if (e->act >= 5 && !e->residual) return POSE_E_NULL;        // act' needs the saved pre-activation
    if (e->accumulate && e->out_dtype != 0) return POSE_E_UNSUPPORTED;
    // the fast epilogue moves 16-byte vectors: needs aligned bases and row pitches, else element-wise stores
    const int celt = e->out_dtype ? 2 : 4;
    ep.vec = !((uintptr_t)e->C % 16 || ((long)e->ldc * celt) % 16 ||
               (e->residual && ((uintptr_t)e->residual % 16 || (e->ldr * 2) % 16)) || (e->bias && (uintptr_t)e->bias % 16) ||
               (e->preact && ((uintptr_t)e->preact % 16 || (e->ldc * 2) % 16)));
    ep.bias = e->bias;
    ep.residual = (const __nv_bfloat16 *)e->residual;
    ep.C = e->C;
    ep.ldc = e->ldc;
    ep.ldr = e->ldr;
    ep.act = e->act;
    ep.out_bf16 = e->out_dtype;
    ep.out_scale = e->out_scale;
    ep.res_scale = e->res_scale;
    ep.preact = (__nv_bfloat16 *)e->preact;
    ep.accumulate = e->accumulate;
    if (e->drop_p < 0.f || e->drop_p >= 1.f) return POSE_E_SHAPE;
    ep.drop_thresh = e->drop_p > 0.f ? drop_threshold(e->drop_p) : 0u;
    ep.drop_scale = e->drop_p > 0.f ? 1.0f / (1.0f - e->drop_p) : 1.0f;
    ep.drop_seed = make_drop_seed(e->drop_seed);
    ep.stats = nullptr;
    ep.bn = e->bn;
    if (e->bn != nullptr) {
        const pose_bn_fuse *bn = e->bn;
        if (!bn->partials || !bn->gamma || !bn->beta || !bn->mean_rstd || !bn->scale_shift) return POSE_E_NULL;
        if (bn->count <= 0) return POSE_E_SHAPE;
        // statistics of the raw contraction (+ bias): no activation, residual, dropout or accumulation in the same epilogue
        if (e->act != 0 || e->residual || e->accumulate || e->preact || e->drop_p > 0.f || !ep.vec || N % 32)
            return POSE_E_UNSUPPORTED;
        ep.stats = bn->partials;
    }
    return POSE_OK;
}

}  // namespace pose

POSE_API int pose_gemm_bf16_ex(const void *A, int lda, const void *W, int ldw, int M, int N, int K,
                               const pose_gemm_epilogue *epilogue, pose_stream_t stream) {
    using namespace pose;
    if (!A || !W) return POSE_E_NULL;
    if (M <= 0 || N <= 0 || K <= 0) return POSE_E_SHAPE;
    if (lda < K || ldw < K) return POSE_E_SHAPE;
    if (lda % 8 || ldw % 8) return POSE_E_SHAPE;  // TMA: row pitch must be a multiple of 16 bytes
    if ((uintptr_t)A % 16 || (uintptr_t)W % 16) return POSE_E_ALIGN;
    Epilogue ep;
    int e = check_epilogue(epilogue, M, N, ep);
    if (e) return e;
    CUtensorMap ma;
    e = make_map_2d(&ma, A, M, K, lda, BM, 64);
    if (e) return e;
    ConvGeom cg = {};
    return dispatch_bn<64, 0>(ma, W, ldw, M, N, K, ep, cg, (M + BM - 1) / BM, (cudaStream_t)stream);
}

// MN-major operand map: memory image [K rows, MN cols] row-major; box = 64 MN elements x bkc contraction rows
static int make_map_mn(CUtensorMap *map, const void *ptr, long k_rows, long mn_cols, long ld_elems, int bkc) {
    pose::EncodeTiledFn fn = pose::encode_tiled();
    if (!fn) return POSE_E_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)mn_cols, (cuuint64_t)k_rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld_elems * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bkc};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? POSE_OK : POSE_E_SHAPE;
}

POSE_API int pose_gemm_bf16_tr(const void *A, long lda, int a_mn, const void *W, long ldw, int b_mn, int M, int N, int K,
                               int k_splits, const pose_gemm_epilogue *epilogue, pose_stream_t stream) {
    using namespace pose;
    if (!A || !W) return POSE_E_NULL;
    if (M <= 0 || N <= 0 || K <= 0) return POSE_E_SHAPE;
    if (lda < (a_mn ? M : K) || ldw < (b_mn ? N : K) || lda % 8 || ldw % 8) return POSE_E_SHAPE;
    if ((uintptr_t)A % 16 || (uintptr_t)W % 16) return POSE_E_ALIGN;
    Epilogue ep;
    int e = check_epilogue(epilogue, M, N, ep);
    if (e) return e;
    if (k_splits > 1 && !ep.accumulate) return POSE_E_UNSUPPORTED;   // partial sums must be accumulated
    CUtensorMap ma, mw;
    e = a_mn ? make_map_mn(&ma, A, K, M, lda, 64) : make_map_2d(&ma, A, M, K, lda, BM, 64);
    if (e) return e;
    ConvGeom cg = {};
    const int m_tiles = (M + BM - 1) / BM;
    cudaStream_t s = (cudaStream_t)stream;
    if (!b_mn) {
        if (a_mn) return POSE_E_UNSUPPORTED;                          // (A^T, W) is not needed by any backward
        return dispatch_bn<64, 0>(ma, W, (int)ldw, M, N, K, ep, cg, m_tiles, s);
    }
    e = make_map_mn(&mw, W, K, N, ldw, 64);
    if (e) return e;
    if (N <= 64) {
        if (a_mn) return launch_gemm<64, 6, 64, 0, 1, 1>(ma, mw, M, N, K, ep, cg, m_tiles, s, k_splits);
        return launch_gemm<64, 6, 64, 0, 0, 1>(ma, mw, M, N, K, ep, cg, m_tiles, s, k_splits);
    }
    if (wide_tile_ok(ep, N, K / (k_splits > 1 ? k_splits : 1), m_tiles, k_splits > 1 ? 2 * k_splits : 1)) {
        if (k_splits > 1) k_splits *= 2;                 // the caller sized the splits for 128-column tiles
        const bool pair = pair_ok(ep, M, N, K, k_splits);
        if (a_mn) return launch_gemm<256, 3, 64, 0, 1, 1>(ma, mw, M, N, K, ep, cg, m_tiles, s, k_splits, pair);
        return launch_gemm<256, 3, 64, 0, 0, 1>(ma, mw, M, N, K, ep, cg, m_tiles, s, k_splits, pair);
    }
    if (a_mn) return launch_gemm<128, 4, 64, 0, 1, 1>(ma, mw, M, N, K, ep, cg, m_tiles, s, k_splits);
    return launch_gemm<128, 4, 64, 0, 0, 1>(ma, mw, M, N, K, ep, cg, m_tiles, s, k_splits);
}

POSE_API int pose_gemm_bf16(const void *A, int lda, const void *W, int ldw, const float *bias, void *C, int ldc,
                            int M, int N, int K, int act, int out_dtype, pose_stream_t stream) {
    pose_gemm_epilogue e = {bias, nullptr, C, ldc, 0, act, out_dtype, 1.0f, 0.0f, nullptr, 0, 0, 0ull, 0.0f, 0};
    return pose_gemm_bf16_ex(A, lda, W, ldw, M, N, K, &e, stream);
}

POSE_API int pose_conv2d_bf16(const void *X, int Nimg, int H, int Wd, int Cin, const void *Wt, int Cout, int KH, int KW,
                              int stride, int dil, int pad, const pose_gemm_epilogue *epilogue, pose_stream_t stream) {
    using namespace pose;
    if (!X || !Wt) return POSE_E_NULL;
    if (Nimg <= 0 || H <= 0 || Wd <= 0 || Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0 || stride <= 0 || dil <= 0 || pad < 0)
        return POSE_E_SHAPE;
    if ((uintptr_t)X % 16 || (uintptr_t)Wt % 16) return POSE_E_ALIGN;
    const int bkc = Cin % 64 == 0 ? 64 : (Cin % 32 == 0 ? 32 : 0);
    if (!bkc) return POSE_E_UNSUPPORTED;  // channels-last activations are padded to a multiple of 32 channels
    const int Ho = (H + 2 * pad - dil * (KH - 1) - 1) / stride + 1, Wo = (Wd + 2 * pad - dil * (KW - 1) - 1) / stride + 1;
    // output patch per M tile (128 pixels): TW x TH pixels of TN images.  Preferred: TW = largest power of two <= 128
    // dividing Wo, then rows, then whole images; when Ho / Wo do not tile exactly (250 x 250 from a 500 x 500 input) the
    // patch shape with the least padding is used and the edge patches are ragged
    int TW = 128;
    while (TW > 1 && (Wo % TW)) TW >>= 1;
    int TH = BM / TW, TN = 1;
    bool exact = true;
    if (TH > Ho) {
        if ((TH % Ho) || TW != Wo) exact = false;
        else { TN = TH / Ho; TH = Ho; }
    }
    if (exact && (TW * TH * TN != BM || Ho % TH || Wo % TW)) exact = false;
    if (!exact) {
        long best = -1;
        TN = 1;
        for (int tw = 128; tw >= 4; tw >>= 1) {
            const int th = BM / tw;
            if (tw * stride > 256 || th * stride > 256) continue;           // TMA box limits
            const long cover = (long)((Wo + tw - 1) / tw) * tw * ((Ho + th - 1) / th) * th;
            if (best < 0 || cover < best) { best = cover; TW = tw; TH = th; }
        }
        if (best < 0) return POSE_E_UNSUPPORTED;
    }
    const int K = KH * KW * Cin, N = Cout;
    Epilogue ep;
    int e = check_epilogue(epilogue, (long)Nimg * Ho * Wo, N, ep);
    if (e) return e;
    CUtensorMap ma;
    e = make_map_nhwc(&ma, X, Nimg, H, Wd, Cin, bkc, TW, TH, TN, stride);
    if (e) return e;
    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    ConvGeom cg = {Ho, Wo, TH, TW, TN, Nimg, KW, KH * KW, stride, dil, pad, Cin / bkc, tiles_w, tiles_h, 64,
                   exact && (TW == Wo || TH == 1) ? 1 : 0};
    const int m_tiles = ((Nimg + TN - 1) / TN) * tiles_h * tiles_w;
    const int M = Nimg * Ho * Wo;
    cudaStream_t s = (cudaStream_t)stream;
    if (bkc == 64) return dispatch_bn<64, 1>(ma, Wt, K, M, N, K, ep, cg, m_tiles, s);
    return dispatch_bn<32, 1>(ma, Wt, K, M, N, K, ep, cg, m_tiles, s);
}

POSE_API int pose_conv2d_wgrad_bf16(const void *dY, const void *X, int Nimg, int H, int Wd, int Cin, int Cout, int KH, int KW,
                                    int stride, int dil, int pad, float *dWk, int k_splits, pose_stream_t stream) {
    using namespace pose;
    if (!dY || !X || !dWk) return POSE_E_NULL;
    if (Nimg <= 0 || H <= 0 || Wd <= 0 || Cin <= 0 || Cout <= 0 || KH <= 0 || KW <= 0 || stride <= 0 || dil <= 0 || pad < 0)
        return POSE_E_SHAPE;
    // channel chunks of 128 bytes, or 32-channel pixels (the 21-channel network input padded to 32: 64-byte rows)
    const int cw = Cin % 64 == 0 ? 64 : (Cin == 32 ? 32 : 0);
    if (!cw || Cout % 8) return POSE_E_UNSUPPORTED;
    if ((uintptr_t)dY % 16 || (uintptr_t)X % 16 || (uintptr_t)dWk % 16) return POSE_E_ALIGN;
    const int Ho = (H + 2 * pad - dil * (KH - 1) - 1) / stride + 1, Wo = (Wd + 2 * pad - dil * (KW - 1) - 1) / stride + 1;
    int TW = 64;
    while (TW > 1 && (Wo % TW)) TW >>= 1;
    int TH = 64 / TW, TN = 1;
    bool exact = true;
    if (TH > Ho) {
        if ((TH % Ho) || TW != Wo) exact = false;
        else { TN = TH / Ho; TH = Ho; }
    }
    if (exact && (TW * TH * TN != 64 || Ho % TH || Wo % TW)) exact = false;
    if (!exact) {           // ragged edge patches: dY outside the map is the TMA zero fill, so they add nothing
        long best = -1;
        TN = 1;
        for (int tw = 64; tw >= 4; tw >>= 1) {
            const int th = 64 / tw;
            if (tw * stride > 256 || th * stride > 256) continue;
            const long cover = (long)((Wo + tw - 1) / tw) * tw * ((Ho + th - 1) / th) * th;
            if (best < 0 || cover < best) { best = cover; TW = tw; TH = th; }
        }
        if (best < 0) return POSE_E_UNSUPPORTED;
    }
    const int tiles_w = (Wo + TW - 1) / TW, tiles_h = (Ho + TH - 1) / TH;
    const int patches = ((Nimg + TN - 1) / TN) * tiles_h * tiles_w;
    const int M = Cout, N = KH * KW * Cin, K = patches * 64;
    pose_gemm_epilogue pe = {nullptr, nullptr, dWk, N, 0, 0, 0, 1.0f, 0.0f, nullptr, 1, 0, 0ull, 0.0f, 0};
    Epilogue ep;
    int e = check_epilogue(&pe, M, N, ep);
    if (e) return e;
    CUtensorMap ma, mw;
    e = make_map_nhwc(&ma, dY, Nimg, Ho, Wo, Cout, 64, TW, TH, TN, 1);
    if (e) return e;
    e = make_map_nhwc(&mw, X, Nimg, H, Wd, Cin, cw, TW, TH, TN, stride);
    if (e) return e;
    ConvGeom cg = {Ho, Wo, TH, TW, TN, Nimg, KW, KH * KW, stride, dil, pad, Cin / cw, tiles_w, tiles_h, cw};
    const int m_tiles = (M + BM - 1) / BM;
    // 128 x 256 tiles (four 64-channel blocks per column tile) for the wide layers: as in pose_gemm_bf16_tr the caller
    // sized the splits for 128-column tiles, so the factor doubles -- the 512-channel dilated 3x3 layers ran ONE wave of 144
    // narrow tiles.  POSE_CONV_WGRAD_NO_BN256 switches it off (A/B measurements)
    static const bool wide_off = getenv("POSE_CONV_WGRAD_NO_BN256") != nullptr;
    if (!wide_off && cw == 64 && N % 256 == 0 && M >= 256 && K >= 8192 && (long)m_tiles * (N / 256) * 2 * k_splits >= kNumSMs / 2)
        return launch_gemm<256, 3, 64, 2, 1, 1>(ma, mw, M, N, K, ep, cg, m_tiles, (cudaStream_t)stream, 2 * k_splits);
    return launch_gemm<128, 4, 64, 2, 1, 1>(ma, mw, M, N, K, ep, cg, m_tiles, (cudaStream_t)stream, k_splits);
}

POSE_API int pose_cast_f32_bf16(const float *in, void *out, long n, pose_stream_t stream) {
    using namespace pose;
    if (!in || !out) return POSE_E_NULL;
    if (n <= 0) return POSE_E_SHAPE;
    if ((uintptr_t)in % 16 || (uintptr_t)out % 16) return POSE_E_ALIGN;
    long blocks = (n / 8 + 255) / 256 + 1;
    int grid = (int)(blocks < (long)kNumSMs * 8 ? blocks : (long)kNumSMs * 8);
    cast_f32_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, (__nv_bfloat16 *)out, n);
    return launch_status();
}
