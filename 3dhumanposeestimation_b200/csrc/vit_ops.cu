// vit_ops.cu -- the non-GEMM pieces of TransformerPoseEstimation.forward (src/models/transformers.py:326-373 and the
// timm ViT-B/16 backbone it wraps): LayerNorm, token assembly (cls + tokens + positional embedding), patch
// extraction for the k16/s16 patch-embedding convolutions (which then run as plain tcgen05 GEMMs), and scaled
// dot-product attention for the short sequences of this model (<= 288 keys: 257 backbone, 273 final encoder,
// 16 / 256 cross-modal), with the whole score row resident in shared memory.
//
// Attention uses warp-level bf16 tensor-core MMAs (nvcuda::wmma, mma.sync) -- a first correct version; the
// tcgen05/TMEM flash-attention kernel is the planned replacement.  It is ~6 % of the model's FLOPs.
#include <mma.h>
#include "common.cuh"

namespace pose {

using namespace nvcuda;

__device__ __forceinline__ void unpack8v(const uint4 &p, float (&f)[8]) {
    const __nv_bfloat162 *h = (const __nv_bfloat162 *)&p;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 t = __bfloat1622float2(h[q]);
        f[2 * q] = t.x;
        f[2 * q + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8v(const float (&f)[8]) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
    return *(uint4 *)h;
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm over the last dimension, one warp per row, fp32 statistics (two passes over registers).
// Rows are addressed in groups so that a sub-range of every sample's tokens can be normalised into a compact
// buffer (e.g. dropping the backbone's cls token, or normalising only token 0):
//   in  row = (g * in_group  + in_off  + i) ,  out row = (g * out_group + out_off + i),  g = r / rows, i = r % rows
// ---------------------------------------------------------------------------------------------------------
template <int D8PL>  // uint4 (8 bf16) per lane: D = 256 * D8PL
__global__ void __launch_bounds__(256)
layernorm_kernel(const __nv_bfloat16 *__restrict__ X, const float *__restrict__ gamma, const float *__restrict__ beta,
                 float eps, long M, int rows, long in_group, long in_off, long out_group, long out_off, int D,
                 __nv_bfloat16 *__restrict__ Y) {
    const long r = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= M) return;
    const long g = r / rows, i = r - g * rows;
    const __nv_bfloat16 *x = X + (g * in_group + in_off + i) * D;
    __nv_bfloat16 *y = Y + (g * out_group + out_off + i) * D;
    float v[D8PL][8];
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < D8PL; ++q) {
        unpack8v(__ldg((const uint4 *)x + q * 32 + lane), v[q]);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += v[q][k];
    }
    const float mean = warp_sum(s) / (float)D;
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < D8PL; ++q)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float d = v[q][k] - mean;
            ss += d * d;
        }
    const float rstd = rsqrtf(warp_sum(ss) / (float)D + eps);
#pragma unroll
    for (int q = 0; q < D8PL; ++q) {
        const int c = (q * 32 + lane) * 8;
        const float4 g0 = __ldg((const float4 *)(gamma + c)), g1 = __ldg((const float4 *)(gamma + c) + 1);
        const float4 b0 = __ldg((const float4 *)(beta + c)), b1 = __ldg((const float4 *)(beta + c) + 1);
        float o[8];
        o[0] = (v[q][0] - mean) * rstd * g0.x + b0.x; o[1] = (v[q][1] - mean) * rstd * g0.y + b0.y;
        o[2] = (v[q][2] - mean) * rstd * g0.z + b0.z; o[3] = (v[q][3] - mean) * rstd * g0.w + b0.w;
        o[4] = (v[q][4] - mean) * rstd * g1.x + b1.x; o[5] = (v[q][5] - mean) * rstd * g1.y + b1.y;
        o[6] = (v[q][6] - mean) * rstd * g1.z + b1.z; o[7] = (v[q][7] - mean) * rstd * g1.w + b1.w;
        ((uint4 *)y)[q * 32 + lane] = pack8v(o);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Token assembly: dst[b] = [cls?] ++ src1[b] ++ src2[b]?, plus pos (fp32 [T, D]) added row-wise.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
token_concat_kernel(__nv_bfloat16 *__restrict__ dst, int T, int D, const float *__restrict__ cls,
                    const __nv_bfloat16 *__restrict__ src1, int N1, const __nv_bfloat16 *__restrict__ src2, int N2,
                    const float *__restrict__ pos, long total8) {
    const int D8 = D >> 3;
    const int has_cls = cls != nullptr;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % D8);
        const long row = i / D8;
        const int t = (int)(row % T);
        const long b = row / T;
        float f[8];
        if (has_cls && t == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = __ldg(cls + c8 * 8 + k);
        } else if (t - has_cls < N1) {
            unpack8v(__ldg((const uint4 *)(src1 + ((b * N1 + (t - has_cls)) * (long)D)) + c8), f);
        } else {
            unpack8v(__ldg((const uint4 *)(src2 + ((b * N2 + (t - has_cls - N1)) * (long)D)) + c8), f);
        }
        if (pos != nullptr) {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += __ldg(pos + (long)t * D + c8 * 8 + k);
        }
        ((uint4 *)dst)[i] = pack8v(f);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Patch extraction for Conv2d(kernel = stride = P) patch embeddings: fp32 NCHW planes (two sources concatenated
// along channels: image + depth, transformers.py:328-330) -> bf16 [B * (H/P) * (W/P), C * P * P], column order
// (c, ky, kx) = the flattened conv weight.  The convolution is then one GEMM.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
patchify_kernel(const float *__restrict__ src0, int C0, const float *__restrict__ src1, int C1, int H, int W, int P,
                long total, __nv_bfloat16 *__restrict__ out) {
    const int C = C0 + C1, PW = W / P, PH = H / P, K = C * P * P;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        long p = i / K;
        const int px = (int)(p % PW);
        p /= PW;
        const int py = (int)(p % PH);
        const long b = p / PH;
        const int kx = k % P, ky = (k / P) % P, c = k / (P * P);
        const int y = py * P + ky, x = px * P + kx;
        const float v = c < C0 ? __ldg(src0 + ((b * C0 + c) * H + y) * (long)W + x)
                               : __ldg(src1 + ((b * C1 + (c - C0)) * H + y) * (long)W + x);
        out[i] = __float2bfloat16_rn(v);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Attention: O[b, q, h*HD:(h+1)*HD] = softmax(Q K^T * scale) V for one (64-query tile, head, batch) per CTA.
// Q/K/V rows have arbitrary pitches (packed qkv or separate projections); the whole score strip (64 x Nk) lives
// in shared memory, so there is no online rescaling.  4 warps, 16 query rows each.
// ---------------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(128)
attention_kernel(const __nv_bfloat16 *__restrict__ Q, const __nv_bfloat16 *__restrict__ K, const __nv_bfloat16 *__restrict__ V,
                 __nv_bfloat16 *__restrict__ O, int Nq, int Nk, int Nkp, long ldq, long ldk, long ldv, long ldo, long bsq,
                 long bsk, long bsv, long bso, float scale) {
    extern __shared__ __align__(128) unsigned char smem[];
    __nv_bfloat16 *Qs = (__nv_bfloat16 *)smem;                 // [64][HD]
    __nv_bfloat16 *Ks = Qs + 64 * HD;                           // [Nkp][HD]
    __nv_bfloat16 *Vs = Ks + (size_t)Nkp * HD;                  // [Nkp][HD]
    float *S = (float *)(Vs + (size_t)Nkp * HD);                // [64][Nkp]
    __nv_bfloat16 *P = (__nv_bfloat16 *)(S + (size_t)64 * Nkp); // [64][Nkp]
    float *Os = (float *)(P + (size_t)64 * Nkp);                // [64][HD] output staging
    __shared__ float rowinv[64];
    const int q0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __nv_bfloat16 *qg = Q + b * bsq + (long)h * HD, *kg = K + b * bsk + (long)h * HD, *vg = V + b * bsv + (long)h * HD;
    constexpr int C8 = HD / 8;
    for (int i = threadIdx.x; i < 64 * C8; i += 128) {
        const int r = i / C8, c = i - r * C8;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (q0 + r < Nq) v = __ldg((const uint4 *)(qg + (long)(q0 + r) * ldq) + c);
        ((uint4 *)Qs)[i] = v;
    }
    for (int i = threadIdx.x; i < Nkp * C8; i += 128) {
        const int r = i / C8, c = i - r * C8;
        uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
        if (r < Nk) {
            kv = __ldg((const uint4 *)(kg + (long)r * ldk) + c);
            vv = __ldg((const uint4 *)(vg + (long)r * ldv) + c);
        }
        ((uint4 *)Ks)[i] = kv;
        ((uint4 *)Vs)[i] = vv;
    }
    __syncthreads();
    // S = Q K^T for this warp's 16 rows
    const int r0 = warp * 16;
    {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> a[HD / 16];
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) wmma::load_matrix_sync(a[k], Qs + r0 * HD + k * 16, HD);
        for (int n = 0; n < Nkp / 16; ++n) {
            wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
            wmma::fill_fragment(acc, 0.f);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
                wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::col_major> bf;  // B(k, n) = K[n][k]
                wmma::load_matrix_sync(bf, Ks + (size_t)n * 16 * HD + k * 16, HD);
                wmma::mma_sync(acc, a[k], bf, acc);
            }
            wmma::store_matrix_sync(S + (size_t)r0 * Nkp + n * 16, acc, Nkp, wmma::mem_row_major);
        }
    }
    __syncwarp();
    // softmax over the Nk real keys of each of the warp's rows (fp32), P unnormalised in bf16
    for (int r = r0; r < r0 + 16; ++r) {
        float *srow = S + (size_t)r * Nkp;
        float m = -INFINITY;
        for (int c = lane; c < Nk; c += 32) m = fmaxf(m, srow[c] * scale);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float sum = 0.f;
        for (int c = lane; c < Nkp; c += 32) {
            float e = 0.f;
            if (c < Nk) {
                e = __expf(srow[c] * scale - m);
                sum += e;
            }
            P[(size_t)r * Nkp + c] = __float2bfloat16_rn(e);
        }
        sum = warp_sum(sum);
        if (lane == 0) rowinv[r] = 1.0f / sum;
    }
    __syncwarp();
    // O = P V, normalised on the way out
#pragma unroll
    for (int n = 0; n < HD / 16; ++n) {
        wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
        wmma::fill_fragment(acc, 0.f);
        for (int k = 0; k < Nkp / 16; ++k) {
            wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> a;
            wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> bf;
            wmma::load_matrix_sync(a, P + (size_t)r0 * Nkp + k * 16, Nkp);
            wmma::load_matrix_sync(bf, Vs + (size_t)k * 16 * HD + n * 16, HD);
            wmma::mma_sync(acc, a, bf, acc);
        }
        wmma::store_matrix_sync(Os + r0 * HD + n * 16, acc, HD, wmma::mem_row_major);
    }
    __syncwarp();
    __nv_bfloat16 *og = O + b * bso + (long)h * HD;
    for (int i = lane; i < 16 * C8; i += 32) {
        const int r = r0 + i / C8, c = i % C8;
        if (q0 + r < Nq) {
            const float inv = rowinv[r];
            float f[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = Os[r * HD + c * 8 + k] * inv;
            *((uint4 *)(og + (long)(q0 + r) * ldo) + c) = pack8v(f);
        }
    }
}

template <int HD>
static size_t attn_smem(int Nkp) {
    return (size_t)64 * HD * 2 + (size_t)Nkp * HD * 2 * 2 + (size_t)64 * Nkp * 4 + (size_t)64 * Nkp * 2 + (size_t)64 * HD * 4;
}

static int grid_cap(long items, int per_block = 256) {
    long blocks = (items + per_block - 1) / per_block;
    long cap = (long)kNumSMs * 16;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace pose

using namespace pose;

POSE_API int pose_layernorm_bf16(const void *X, const float *gamma, const float *beta, float eps, long M, int rows,
                                 long in_group, long in_off, long out_group, long out_off, int D, void *Y,
                                 pose_stream_t stream) {
    if (!X || !gamma || !beta || !Y) return POSE_E_NULL;
    if (M <= 0 || rows <= 0 || D <= 0) return POSE_E_SHAPE;
    if ((uintptr_t)X % 16 || (uintptr_t)Y % 16 || (uintptr_t)gamma % 16 || (uintptr_t)beta % 16) return POSE_E_ALIGN;
    const int grid = (int)((M + 7) / 8);
    cudaStream_t s = (cudaStream_t)stream;
#define LN_LAUNCH(N_)                                                                                                  \
    layernorm_kernel<N_><<<grid, 256, 0, s>>>((const __nv_bfloat16 *)X, gamma, beta, eps, M, rows, in_group, in_off,   \
                                              out_group, out_off, D, (__nv_bfloat16 *)Y)
    if (D == 256) LN_LAUNCH(1);
    else if (D == 512) LN_LAUNCH(2);
    else if (D == 768) LN_LAUNCH(3);
    else if (D == 1024) LN_LAUNCH(4);
    else return POSE_E_UNSUPPORTED;  // embedding widths of the ViT family used here
#undef LN_LAUNCH
    return launch_status();
}

POSE_API int pose_token_concat_bf16(void *dst, int B, int T, int D, const float *cls, const void *src1, int N1,
                                    const void *src2, int N2, const float *pos, pose_stream_t stream) {
    if (!dst || !src1) return POSE_E_NULL;
    if (B <= 0 || T <= 0 || D <= 0 || D % 8 || N1 < 0 || N2 < 0) return POSE_E_SHAPE;
    if ((cls ? 1 : 0) + N1 + (src2 ? N2 : 0) != T) return POSE_E_SHAPE;
    const long total8 = (long)B * T * (D / 8);
    token_concat_kernel<<<grid_cap(total8), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16 *)dst, T, D, cls,
                                                                           (const __nv_bfloat16 *)src1, N1,
                                                                           (const __nv_bfloat16 *)src2, N2, pos, total8);
    return launch_status();
}

POSE_API int pose_patchify_bf16(const float *src0, int C0, const float *src1, int C1, int B, int H, int W, int P,
                                void *out, pose_stream_t stream) {
    if (!src0 || !out || (C1 > 0 && !src1)) return POSE_E_NULL;
    if (B <= 0 || C0 <= 0 || C1 < 0 || H <= 0 || W <= 0 || P <= 0 || H % P || W % P) return POSE_E_SHAPE;
    const long total = (long)B * (H / P) * (W / P) * (C0 + C1) * P * P;
    patchify_kernel<<<grid_cap(total), 256, 0, (cudaStream_t)stream>>>(src0, C0, src1, C1, H, W, P, total,
                                                                      (__nv_bfloat16 *)out);
    return launch_status();
}

POSE_API int pose_attention_bf16(const void *Q, const void *K, const void *V, void *O, int B, int heads, int Nq, int Nk,
                                 int head_dim, long ldq, long ldk, long ldv, long ldo, long bsq, long bsk, long bsv, long bso,
                                 float scale, pose_stream_t stream) {
    if (!Q || !K || !V || !O) return POSE_E_NULL;
    if (B <= 0 || heads <= 0 || Nq <= 0 || Nk <= 0) return POSE_E_SHAPE;
    if (head_dim != 48 && head_dim != 64) return POSE_E_UNSUPPORTED;
    if (ldq % 8 || ldk % 8 || ldv % 8 || ldo % 8 || bsq % 8 || bsk % 8 || bsv % 8 || bso % 8) return POSE_E_ALIGN;
    if ((uintptr_t)Q % 16 || (uintptr_t)K % 16 || (uintptr_t)V % 16 || (uintptr_t)O % 16) return POSE_E_ALIGN;
    const int Nkp = (Nk + 15) / 16 * 16;
    const size_t smem = head_dim == 64 ? attn_smem<64>(Nkp) : attn_smem<48>(Nkp);
    if (smem > 227 * 1024 - 1024) return POSE_E_UNSUPPORTED;  // whole-row scores: up to ~288 keys
    dim3 grid((Nq + 63) / 64, heads, B);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e;
    if (head_dim == 64) {
        e = cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attention_kernel<64><<<grid, 128, smem, s>>>((const __nv_bfloat16 *)Q, (const __nv_bfloat16 *)K, (const __nv_bfloat16 *)V,
                                                     (__nv_bfloat16 *)O, Nq, Nk, Nkp, ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso, scale);
    } else {
        e = cudaFuncSetAttribute(attention_kernel<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        attention_kernel<48><<<grid, 128, smem, s>>>((const __nv_bfloat16 *)Q, (const __nv_bfloat16 *)K, (const __nv_bfloat16 *)V,
                                                     (__nv_bfloat16 *)O, Nq, Nk, Nkp, ldq, ldk, ldv, ldo, bsq, bsk, bsv, bso, scale);
    }
    return launch_status();
}
