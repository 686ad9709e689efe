// vit_ops.cu -- the non-GEMM pieces of TransformerPoseEstimation.forward (src/models/transformers.py:326-373 and the
// timm ViT-B/16 backbone it wraps): LayerNorm, token assembly (cls + tokens + positional embedding), patch
// extraction for the k16/s16 patch-embedding convolutions (which then run as plain tcgen05 GEMMs), and scaled
// the backward of those pieces, and the fused AdamW step.
//
// Attention itself runs on the tcgen05 tensor cores: attention_tc.cu.
#include "common.cuh"
#include "rowpipe.cuh"

namespace pose {

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

// 8 bf16 -> four fp32 pairs in two instructions per pair (shift, mask)
__device__ __forceinline__ void unpack8p(const uint4 &p, float2 (&f)[4]) {
    const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) f[q] = make_float2(__uint_as_float(w[q] << 16), __uint_as_float(w[q] & 0xffff0000u));
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
// P % 8 == 0 and 16-byte aligned planes: a thread converts 8 consecutive kx (two float4 loads, one 16-byte store); the
// index of a 16-byte group is decoded with 32-bit arithmetic below the patch level (the element-wise version above spends
// seven 64-bit divisions per element: 0.7 TB/s)
__global__ void __launch_bounds__(256)
patchify8_kernel(const float *__restrict__ src0, int C0, const float *__restrict__ src1, int C1, int H, int W, int P,
                 long total8, __nv_bfloat16 *__restrict__ out) {
    const int C = C0 + C1, PW = W / P, PH = H / P, P8 = P >> 3, K8 = C * P * P8, patches = PW * PH;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int k8 = (int)(i % K8);
        const long pidx = i / K8;
        const int pp = (int)(pidx % patches);
        const long b = pidx / patches;
        const int py = pp / PW, px = pp - py * PW;
        const int kx8 = k8 % P8, t = k8 / P8;
        const int ky = t % P, c = t / P;
        const int y = py * P + ky, x = px * P + kx8 * 8;
        const float *src = c < C0 ? src0 + ((b * C0 + c) * H + y) * (long)W + x : src1 + ((b * C1 + (c - C0)) * H + y) * (long)W + x;
        const float4 lo = __ldg((const float4 *)src), hi = __ldg((const float4 *)src + 1);
        __nv_bfloat162 p0 = __floats2bfloat162_rn(lo.x, lo.y), p1 = __floats2bfloat162_rn(lo.z, lo.w);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(hi.x, hi.y), p3 = __floats2bfloat162_rn(hi.z, hi.w);
        ((uint4 *)out)[i] = make_uint4(*(uint32_t *)&p0, *(uint32_t *)&p1, *(uint32_t *)&p2, *(uint32_t *)&p3);
    }
}

// ---------------------------------------------------------------------------------------------------------
// LayerNorm backward.  dX = dRes + rstd * (g*dY - mean(g*dY) - xhat * mean(g*dY*xhat)); dgamma += sum dY*xhat,
// dbeta += sum dY.  Statistics are recomputed from the saved input (one warp per row, rows strided over the
// grid so that each thread keeps its own columns' dgamma / dbeta partials in registers); row addressing as in
// the forward: X / dX / dRes use the input addressing, dY the output addressing.
// ---------------------------------------------------------------------------------------------------------
template <int D8PL>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const __nv_bfloat16 *__restrict__ X, const __nv_bfloat16 *__restrict__ dY,
                     const float *__restrict__ gamma, float eps, long M, int rows, long in_group, long in_off,
                     long out_group, long out_off, int D, const __nv_bfloat16 *__restrict__ dRes,
                     __nv_bfloat16 *__restrict__ dX, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    // dgamma / dbeta partials live in shared memory, one private row pair per warp (a lane only ever touches its own
    // columns: no synchronisation until the final fold).  Keeping those 48 accumulators in registers held the kernel at
    // 168 registers = one CTA (8 warps) per SM, each warp walking its rows strictly one after the other.
    extern __shared__ __align__(16) float s_acc[];          // [8 warps][2][D]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *ga = s_acc + (size_t)warp * 2 * D, *ba = ga + D;
#pragma unroll
    for (int q = 0; q < D8PL; ++q) {
        const int c = (q * 32 + lane) * 8;
        *(float4 *)(ga + c) = make_float4(0.f, 0.f, 0.f, 0.f); *(float4 *)(ga + c + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        *(float4 *)(ba + c) = make_float4(0.f, 0.f, 0.f, 0.f); *(float4 *)(ba + c + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // Everything elementwise runs on packed fp32 pairs (FADD2 / FMUL2 / FFMA2: two lanes per issue slot): the scalar version
    // executed 31 instructions per element and was issue / latency bound at 2.3 TB/s on L2-resident tensors
    // (profiles/r02h_ln_bwd_full_raw.csv).
    const float inv_d = 1.0f / (float)D;
    // A warp walks its rows strictly one after the other and every row is a chain of dependent latencies (row loads, three
    // warp reductions, the residual load): 33 us for 76 MB with 24 warps per SM.  The NEXT row of the warp is therefore
    // copied into a lane-private shared-memory slot with cp.async while the current one is processed ([warp][tensor][chunk]
    // [lane] uint4 behind the accumulators; nobody else reads a lane's slots, so cp.async.wait_group is the only sync).
    const uint32_t slot = (uint32_t)__cvta_generic_to_shared(s_acc + (size_t)8 * 2 * D) + (uint32_t)(((warp * 3 * D8PL) * 32 + lane) * 16);
    const long rstep = (long)gridDim.x * 8;
    auto issue = [&](long rr_) {
        const long g = rr_ / rows, i = rr_ - g * rows;
        const long rin = g * in_group + in_off + i, rout = g * out_group + out_off + i;
#pragma unroll
        for (int q = 0; q < D8PL; ++q) {
            cp16(slot + (uint32_t)((0 * D8PL + q) * 32 * 16), (const uint4 *)(X + rin * D) + q * 32 + lane);
            cp16(slot + (uint32_t)((1 * D8PL + q) * 32 * 16), (const uint4 *)(dY + rout * D) + q * 32 + lane);
            if (dRes != nullptr) cp16(slot + (uint32_t)((2 * D8PL + q) * 32 * 16), (const uint4 *)(dRes + rin * D) + q * 32 + lane);
        }
        cp_async_commit();
    };
    long r = (long)blockIdx.x * 8 + warp;
    if (r < M) issue(r);
    for (; r < M; r += rstep) {
        const long g = r / rows, i = r - g * rows;
        const long rin = g * in_group + in_off + i;
        cp_async_wait<0>();
        float2 v[D8PL][4], d[D8PL][4];
        uint4 rres[D8PL];
        float2 sa = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < D8PL; ++q) {
            unpack8p(lds16(slot + (uint32_t)((0 * D8PL + q) * 32 * 16)), v[q]);
            unpack8p(lds16(slot + (uint32_t)((1 * D8PL + q) * 32 * 16)), d[q]);
            rres[q] = dRes != nullptr ? lds16(slot + (uint32_t)((2 * D8PL + q) * 32 * 16)) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int k = 0; k < 4; ++k) sa = __fadd2_rn(sa, v[q][k]);
        }
        if (r + rstep < M) issue(r + rstep);        // the slots were just read: the next row lands while this one is processed
        const float mean = warp_sum(sa.x + sa.y) * inv_d;
        const float2 nm = make_float2(-mean, -mean);
        float2 sq = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < D8PL; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[q][k] = __fadd2_rn(v[q][k], nm);
                sq = __ffma2_rn(v[q][k], v[q][k], sq);
            }
        const float rstd = rsqrtf(warp_sum(sq.x + sq.y) * inv_d + eps);
        const float2 rs2 = make_float2(rstd, rstd);
        float2 t1 = make_float2(0.f, 0.f), t2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < D8PL; ++q) {
            const int c = (q * 32 + lane) * 8;
            float2 ga8[4], ba8[4], gm[4];
            *(float4 *)&ga8[0] = *(const float4 *)(ga + c); *(float4 *)&ga8[2] = *(const float4 *)(ga + c + 4);
            *(float4 *)&ba8[0] = *(const float4 *)(ba + c); *(float4 *)&ba8[2] = *(const float4 *)(ba + c + 4);
            *(float4 *)&gm[0] = __ldg((const float4 *)(gamma + c)); *(float4 *)&gm[2] = __ldg((const float4 *)(gamma + c) + 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[q][k] = __fmul2_rn(v[q][k], rs2);                     // xhat
                ga8[k] = __ffma2_rn(d[q][k], v[q][k], ga8[k]);
                ba8[k] = __fadd2_rn(ba8[k], d[q][k]);
                d[q][k] = __fmul2_rn(d[q][k], gm[k]);                   // g * dY
                t1 = __fadd2_rn(t1, d[q][k]);
                t2 = __ffma2_rn(d[q][k], v[q][k], t2);
            }
            *(float4 *)(ga + c) = *(const float4 *)&ga8[0]; *(float4 *)(ga + c + 4) = *(const float4 *)&ga8[2];
            *(float4 *)(ba + c) = *(const float4 *)&ba8[0]; *(float4 *)(ba + c + 4) = *(const float4 *)&ba8[2];
        }
        // the two row sums travel through one butterfly
        float s1 = t1.x + t1.y, s2 = t2.x + t2.y;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        const float2 ns1 = make_float2(-s1 * inv_d, -s1 * inv_d), ns2 = make_float2(-s2 * inv_d, -s2 * inv_d);
#pragma unroll
        for (int q = 0; q < D8PL; ++q) {
            float2 o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                o[k] = __fmul2_rn(rs2, __fadd2_rn(__ffma2_rn(v[q][k], ns2, d[q][k]), ns1));   // rstd (g dY - s1 - xhat s2)
            if (dRes != nullptr) {
                float2 rv[4];
                unpack8p(rres[q], rv);
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = __fadd2_rn(o[k], rv[k]);
            }
            __nv_bfloat162 h[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(o[k].x, o[k].y);
            ((uint4 *)(dX + rin * D))[q * 32 + lane] = *(uint4 *)h;
        }
    }
    // fold the 8 warps' partials, then one atomic per column per CTA
    __syncthreads();
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        float *dst = pass == 0 ? dgamma : dbeta;
        if (dst != nullptr)
            for (int c = threadIdx.x; c < D; c += 256) {
                float t = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) t += s_acc[((size_t)w * 2 + pass) * D + c];
                atomicAdd(dst + c, t);
            }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Column sums (bias gradients): out[c] += sum_r X[r, c], X bf16 [M, ld].  Block = 32 column groups of 8 x 8 rows.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16 *__restrict__ X, long M, int N, long ld, float *__restrict__ out) {
    __shared__ float red[8][256];
    const int cg = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c0 = (blockIdx.x * 32 + cg) * 8;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (c0 < N) {
        const long step = (long)gridDim.y * 8;
        long r = (long)blockIdx.y * 8 + ry;
        if (c0 + 8 <= N) {
            for (; r + 3 * step < M; r += 4 * step) {          // four independent 16-byte loads in flight per thread
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg((const uint4 *)(X + (r + u * step) * ld + c0));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float f[8];
                    unpack8v(v[u], f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[k] += f[k];
                }
            }
            for (; r < M; r += step) {
                float f[8];
                unpack8v(__ldg((const uint4 *)(X + r * ld + c0)), f);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += f[k];
            }
        } else {
            for (; r < M; r += step)
                for (int k = 0; c0 + k < N; ++k) acc[k] += __bfloat162float(X[r * ld + c0 + k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) red[ry][cg * 8 + k] = acc[k];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < N) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

// out[t, c] (fp32, accumulated) += sum_b X[(b * T_in + t_off + t) * D + c]: gradients of the positional embeddings and
// of the class tokens (parameters broadcast over the batch).
__global__ void __launch_bounds__(256)
batch_rowsum_kernel(const __nv_bfloat16 *__restrict__ X, int B, long T_in, long t_off, int T_out, int D,
                    float *__restrict__ out) {
    const int D8 = D >> 3;
    const long total = (long)T_out * D8;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % D8);
        const long t = i / D8;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int b = 0; b < B; ++b) {
            float f[8];
            unpack8v(__ldg((const uint4 *)(X + ((long)b * T_in + t_off + t) * D) + c8), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] += f[k];
        }
        float *o = out + t * D + c8 * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += acc[k];
    }
}

// dst[b, i, :] = src[b, t_off + i, :] for i < n   (gradient of torch.cat along the token axis)
__global__ void __launch_bounds__(256)
token_slice_kernel(const __nv_bfloat16 *__restrict__ src, long T, long t_off, int n, int D, long total8,
                   __nv_bfloat16 *__restrict__ dst) {
    const int D8 = D >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % D8);
        const long row = i / D8;
        const long b = row / n, t = row - b * n;
        ((uint4 *)dst)[i] = __ldg((const uint4 *)(src + (b * T + t_off + t) * D) + c8);
    }
}

// fp32 [rows, cols] (pitch ld_in) -> bf16 [rows, ld_out], columns >= cols zero-filled: gradients whose width is not a
// multiple of 8 (the 51-wide pose output) become TMA-addressable GEMM operands
__global__ void __launch_bounds__(256)
cast_pad_kernel(const float *__restrict__ in, long ld_in, int cols, __nv_bfloat16 *__restrict__ out, long ld_out, long total) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / ld_out;
        const int c = (int)(i - r * ld_out);
        out[i] = __float2bfloat16_rn(c < cols ? in[r * ld_in + c] : 0.f);
    }
}

// ---------------------------------------------------------------------------------------------------------
// AdamW over one flat fp32 parameter buffer (torch.optim.AdamW semantics: decoupled weight decay, bias
// correction, eps added after the sqrt), fused with: gradient scaling (loss scale / world size), refresh of the
// bf16 shadow weights the GEMMs read, and zeroing of the gradient for the next accumulation window.
// 28 B/param of HBM traffic (+ 2 B for the shadow): 16 B loads, 16 + 2 B stores per 4 parameters x 4.
// ---------------------------------------------------------------------------------------------------------
template <bool G16>
__global__ void __launch_bounds__(256)
adamw_kernel(float *__restrict__ p, float *__restrict__ g, const __nv_bfloat16 *__restrict__ g16, float *__restrict__ m,
             float *__restrict__ v, __nv_bfloat16 *__restrict__ shadow, long n, float lr, float beta1, float beta2, float eps,
             float wd, float bc1, float bc2_sqrt, float grad_scale, int zero_grad, const int *__restrict__ step_dev) {
    if (step_dev != nullptr) {      // step count from the bound pose_step_state (CUDA-graph replay): bias corrections here
        const float st = (float)__ldg(step_dev);
        bc1 = 1.0f - powf(beta1, st);
        bc2_sqrt = sqrtf(1.0f - powf(beta2, st));
    }
    const long n4 = n >> 2;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        float4 pp = ((float4 *)p)[i], gg, mm = ((float4 *)m)[i], vv = ((float4 *)v)[i];
        if (G16) {          // the all-reduced gradient arrives as bf16 (data-parallel exchange at half the wire bytes)
            const uint2 pk = ((const uint2 *)g16)[i];
            const float2 a = __bfloat1622float2(*(const __nv_bfloat162 *)&pk.x), b = __bfloat1622float2(*(const __nv_bfloat162 *)&pk.y);
            gg = make_float4(a.x, a.y, b.x, b.y);
        } else {
            gg = ((const float4 *)g)[i];
        }
        float *pa = (float *)&pp, *ga = (float *)&gg, *ma = (float *)&mm, *va = (float *)&vv;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gr = ga[k] * grad_scale;
            pa[k] *= 1.0f - lr * wd;
            ma[k] = beta1 * ma[k] + (1.0f - beta1) * gr;
            va[k] = beta2 * va[k] + (1.0f - beta2) * gr * gr;
            const float denom = sqrtf(va[k]) / bc2_sqrt + eps;
            pa[k] -= (lr / bc1) * (ma[k] / denom);
        }
        ((float4 *)p)[i] = pp;
        ((float4 *)m)[i] = mm;
        ((float4 *)v)[i] = vv;
        if (zero_grad) ((float4 *)g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (shadow != nullptr) {
            __nv_bfloat162 s0 = __floats2bfloat162_rn(pa[0], pa[1]), s1 = __floats2bfloat162_rn(pa[2], pa[3]);
            ((uint2 *)shadow)[i] = make_uint2(*(uint32_t *)&s0, *(uint32_t *)&s1);
        }
    }
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
    if (P % 8 == 0 && W % 4 == 0 && (uintptr_t)src0 % 16 == 0 && (!src1 || (uintptr_t)src1 % 16 == 0) && (uintptr_t)out % 16 == 0)
        patchify8_kernel<<<grid_cap(total / 8), 256, 0, (cudaStream_t)stream>>>(src0, C0, src1, C1, H, W, P, total / 8,
                                                                               (__nv_bfloat16 *)out);
    else
        patchify_kernel<<<grid_cap(total), 256, 0, (cudaStream_t)stream>>>(src0, C0, src1, C1, H, W, P, total,
                                                                          (__nv_bfloat16 *)out);
    return launch_status();
}

POSE_API int pose_layernorm_bwd_bf16(const void *X, const void *dY, const float *gamma, float eps, long M, int rows,
                                     long in_group, long in_off, long out_group, long out_off, int D, const void *dRes,
                                     void *dX, float *dgamma, float *dbeta, pose_stream_t stream) {
    if (!X || !dY || !gamma || !dX) return POSE_E_NULL;
    if (M <= 0 || rows <= 0 || D <= 0) return POSE_E_SHAPE;
    if ((uintptr_t)X % 16 || (uintptr_t)dY % 16 || (uintptr_t)dX % 16 || (uintptr_t)gamma % 16 || (dRes && (uintptr_t)dRes % 16))
        return POSE_E_ALIGN;
    long blocks = (M + 7) / 8;
    const int smem = 8 * 2 * D * (int)sizeof(float) + 8 * 3 * (D / 256) * 32 * 16;     // accumulators + the next-row slots
    cudaStream_t s = (cudaStream_t)stream;
#define LNB_LAUNCH(N_)                                                                                                  \
    static bool cfg##N_ = false;                                                                                        \
    if (!cfg##N_) {                                                                                                     \
        if (cudaFuncSetAttribute(layernorm_bwd_kernel<N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) \
            return POSE_E_UNSUPPORTED;                                                                                  \
        cfg##N_ = true;                                                                                                 \
    }                                                                                                                   \
    int per_sm = 1;                                                                                                     \
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, layernorm_bwd_kernel<N_>, 256, smem) != cudaSuccess || per_sm < 1) \
        per_sm = 1;                                                                                                     \
    const int grid = (int)(blocks < (long)kNumSMs * per_sm ? blocks : (long)kNumSMs * per_sm);   /* one resident wave */   \
    layernorm_bwd_kernel<N_><<<grid, 256, smem, s>>>((const __nv_bfloat16 *)X, (const __nv_bfloat16 *)dY, gamma, eps, M,   \
                                                  rows, in_group, in_off, out_group, out_off, D,                        \
                                                  (const __nv_bfloat16 *)dRes, (__nv_bfloat16 *)dX, dgamma, dbeta)
    if (D == 256) { LNB_LAUNCH(1); }
    else if (D == 512) { LNB_LAUNCH(2); }
    else if (D == 768) { LNB_LAUNCH(3); }
    else if (D == 1024) { LNB_LAUNCH(4); }
    else return POSE_E_UNSUPPORTED;
#undef LNB_LAUNCH
    return launch_status();
}

POSE_API int pose_colsum_bf16(const void *X, long M, int N, long ld, float *out, pose_stream_t stream) {
    if (!X || !out) return POSE_E_NULL;
    if (M <= 0 || N <= 0 || ld < N) return POSE_E_SHAPE;
    if ((uintptr_t)X % 16 || ld % 8) return POSE_E_ALIGN;
    const int gx = (N + 255) / 256;
    long gy = (M + 63) / 64;                       // >= 8 rows per thread row
    const long cap = (kNumSMs * 6 + gx - 1) / gx;
    if (gy > cap) gy = cap;
    if (gy < 1) gy = 1;
    colsum_kernel<<<dim3(gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)X, M, N, ld, out);
    return launch_status();
}

POSE_API int pose_batch_rowsum_bf16(const void *X, int B, long T_in, long t_off, int T_out, int D, float *out,
                                    pose_stream_t stream) {
    if (!X || !out) return POSE_E_NULL;
    if (B <= 0 || T_in <= 0 || t_off < 0 || T_out <= 0 || t_off + T_out > T_in || D <= 0 || D % 8) return POSE_E_SHAPE;
    if ((uintptr_t)X % 16) return POSE_E_ALIGN;
    batch_rowsum_kernel<<<grid_cap((long)T_out * (D / 8)), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)X, B, T_in,
                                                                                         t_off, T_out, D, out);
    return launch_status();
}

POSE_API int pose_token_slice_bf16(const void *src, int B, long T, long t_off, int n, int D, void *dst, pose_stream_t stream) {
    if (!src || !dst) return POSE_E_NULL;
    if (B <= 0 || T <= 0 || t_off < 0 || n <= 0 || t_off + n > T || D <= 0 || D % 8) return POSE_E_SHAPE;
    if ((uintptr_t)src % 16 || (uintptr_t)dst % 16) return POSE_E_ALIGN;
    const long total8 = (long)B * n * (D / 8);
    token_slice_kernel<<<grid_cap(total8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)src, T, t_off, n, D, total8,
                                                                          (__nv_bfloat16 *)dst);
    return launch_status();
}

static int adamw_launch(float *param, float *grad, const void *grad_bf16, float *exp_avg, float *exp_avg_sq, void *shadow_bf16, long n,
                        float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale, int zero_grad,
                        pose_stream_t stream) {
    if (!param || !grad || !exp_avg || !exp_avg_sq) return POSE_E_NULL;
    if (n <= 0 || n % 4 || step < 0) return POSE_E_SHAPE;
    const int *step_dev = nullptr;
    if (step == 0) {                 // the step count lives in the bound per-step state
        if (!g_step_state) return POSE_E_SHAPE;
        step_dev = &g_step_state->adam_step;
        step = 1;
    }
    if ((uintptr_t)param % 16 || (uintptr_t)grad % 16 || (uintptr_t)exp_avg % 16 || (uintptr_t)exp_avg_sq % 16 ||
        (shadow_bf16 && (uintptr_t)shadow_bf16 % 8) || (grad_bf16 && (uintptr_t)grad_bf16 % 8))
        return POSE_E_ALIGN;
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2s = sqrtf(1.0f - powf(beta2, (float)step));
    cudaStream_t s = (cudaStream_t)stream;
    if (grad_bf16 != nullptr)
        adamw_kernel<true><<<grid_cap(n / 4, 256), 256, 0, s>>>(param, grad, (const __nv_bfloat16 *)grad_bf16, exp_avg, exp_avg_sq,
                                                                (__nv_bfloat16 *)shadow_bf16, n, lr, beta1, beta2, eps, weight_decay,
                                                                bc1, bc2s, grad_scale, zero_grad, step_dev);
    else
        adamw_kernel<false><<<grid_cap(n / 4, 256), 256, 0, s>>>(param, grad, nullptr, exp_avg, exp_avg_sq,
                                                                 (__nv_bfloat16 *)shadow_bf16, n, lr, beta1, beta2, eps, weight_decay,
                                                                 bc1, bc2s, grad_scale, zero_grad, step_dev);
    return launch_status();
}

POSE_API int pose_adamw_step(float *param, float *grad, float *exp_avg, float *exp_avg_sq, void *shadow_bf16, long n,
                             float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                             float grad_scale, int zero_grad, pose_stream_t stream) {
    return adamw_launch(param, grad, nullptr, exp_avg, exp_avg_sq, shadow_bf16, n, lr, beta1, beta2, eps, weight_decay, step,
                        grad_scale, zero_grad, stream);
}

POSE_API int pose_adamw_step_g16(float *param, float *grad, const void *grad_bf16, float *exp_avg, float *exp_avg_sq,
                                 void *shadow_bf16, long n, float lr, float beta1, float beta2, float eps, float weight_decay,
                                 int step, float grad_scale, int zero_grad, pose_stream_t stream) {
    if (!grad_bf16) return POSE_E_NULL;
    return adamw_launch(param, grad, grad_bf16, exp_avg, exp_avg_sq, shadow_bf16, n, lr, beta1, beta2, eps, weight_decay, step,
                        grad_scale, zero_grad, stream);
}

POSE_API int pose_cast_f32_bf16_2d(const float *in, long ld_in, long rows, int cols, void *out, long ld_out,
                                   pose_stream_t stream) {
    if (!in || !out) return POSE_E_NULL;
    if (rows <= 0 || cols <= 0 || ld_in < cols || ld_out < cols) return POSE_E_SHAPE;
    const long total = rows * ld_out;
    cast_pad_kernel<<<grid_cap(total), 256, 0, (cudaStream_t)stream>>>(in, ld_in, cols, (__nv_bfloat16 *)out, ld_out, total);
    return launch_status();
}
