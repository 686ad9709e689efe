// cnn_train.cu -- the bandwidth-bound kernels of the CNN TRAINING step (reference: loss.backward() over
// src/models/cnn.py; src/train.py:83-92): batch-statistics BatchNorm forward / backward fused with the activation,
// the residual add and the channel concatenation; depthwise 3x3 convolution backward (data and weights); the
// backward of the SE / ECA / CoordAttention gates, of the WASP branch mix and of the pooling layers; dropout; and
// the per-step re-layout of the parameters the tensor-core kernels read (one table-driven launch).
// Channels-last bf16 activations, fp32 statistics and parameter gradients (accumulated like torch's .grad).
#include "tc_common.cuh"
#include "rowpipe.cuh"

namespace pose {

__device__ __forceinline__ void up8(const uint4 &p, float (&f)[8]) {
    const __nv_bfloat162 *h = (const __nv_bfloat162 *)&p;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 t = __bfloat1622float2(h[q]);
        f[2 * q] = t.x;
        f[2 * q + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pk8(const float (&f)[8]) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
    return *(uint4 *)h;
}
__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + __expf(-v)); }
__device__ __forceinline__ float act_fwd(float z, int act) {
    switch (act) {
        case 1: return z > 0.f ? z : 0.f;
        case 2: return z * sigmoid_f(z);
        case 3: return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f));
        case 4: return sigmoid_f(z);
        default: return z;
    }
}
__device__ __forceinline__ float act_bwd(float z, int act) {   // d act / dz
    switch (act) {
        case 1: return z > 0.f ? 1.f : 0.f;
        case 2: { const float s = sigmoid_f(z); return s * (1.0f + z * (1.0f - s)); }
        case 3: return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
        case 4: { const float s = sigmoid_f(z); return s * (1.0f - s); }
        default: return 1.f;
    }
}

static int grid_for(long items, int per_block = 256, int max_waves = 16) {
    long blocks = (items + per_block - 1) / per_block;
    long cap = (long)kNumSMs * max_waves;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// ---- BatchNorm (training mode) ------------------------------------------------------------------------------
// Thread mapping of the row-streaming kernels: a thread owns ONE group of 8 channels for its whole life (its per-channel
// coefficients live in registers) and walks rows; block = G * k threads (G = C / 8 groups, k rows side by side), so
// consecutive threads read consecutive 16-byte chunks of a row.
struct RowMap {
    int g, rsub, rpb, c0;
    __device__ __forceinline__ RowMap(int C) {
        const int G = C >> 3;
        g = threadIdx.x % G;
        rsub = threadIdx.x / G;
        rpb = blockDim.x / G;
        c0 = g * 8;
    }
};
static int rowmap_threads(int C) {
    const int G = C / 8;
    const int k = 256 / G > 0 ? 256 / G : 1;
    return G * k;
}
static int rowmap_grid(long M, int C, int rows_per_thread = 8) {
    const int rpb = rowmap_threads(C) / (C / 8);
    long blocks = (M + (long)rpb * rows_per_thread - 1) / ((long)rpb * rows_per_thread);
    const long cap = (long)kNumSMs * 8;
    return (int)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}
__device__ __forceinline__ void ld8f(const float *p, float (&f)[8]) {
    const float4 a = __ldg((const float4 *)p), b = __ldg((const float4 *)p + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// activation and its derivative, template-selected (silu through one MUFU.TANH)
template <int ACT>
__device__ __forceinline__ float actf(float z) {
    if (ACT == 1) return fmaxf(z, 0.f);
    if (ACT == 2) return 0.5f * z * (1.0f + tanh_fast(0.5f * z));
    return z;
}
template <int ACT>
__device__ __forceinline__ float dactf(float z) {
    if (ACT == 1) return z > 0.f ? 1.f : 0.f;
    if (ACT == 2) {
        const float sg = 0.5f * (1.0f + tanh_fast(0.5f * z));
        return sg * (1.0f + z * (1.0f - sg));
    }
    return 1.f;
}
// The same on packed fp32 pairs (FFMA2 / FMUL2 / FADD2: two lanes per issue slot).  The BatchNorm passes are instruction
// bound as much as bandwidth bound (identical shapes: 121 us without an activation, 163 us with SiLU evaluated lane by
// lane), so their SiLU variants run the affine map, the activation and the reductions on pairs.
__device__ __forceinline__ float2 tanh2(float2 v) { return make_float2(tanh_fast(v.x), tanh_fast(v.y)); }
template <int ACT>
__device__ __forceinline__ float2 actf2(float2 z) {                 // act(z)
    if (ACT == 2) {
        const float2 h = __fmul2_rn(z, make_float2(0.5f, 0.5f));
        return __ffma2_rn(h, tanh2(h), h);                          // 0.5 z (1 + tanh(0.5 z))
    }
    if (ACT == 1) return make_float2(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f));
    return z;
}
template <int ACT>
__device__ __forceinline__ float2 dactf2(float2 z) {                // act'(z)
    if (ACT == 2) {
        const float2 t = tanh2(__fmul2_rn(z, make_float2(0.5f, 0.5f)));
        const float2 sg = __ffma2_rn(t, make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));
        const float2 om = __ffma2_rn(sg, make_float2(-1.0f, -1.0f), make_float2(1.0f, 1.0f));
        return __fmul2_rn(sg, __ffma2_rn(z, om, make_float2(1.0f, 1.0f)));       // sg (1 + z (1 - sg))
    }
    if (ACT == 1) return make_float2(z.x > 0.f ? 1.f : 0.f, z.y > 0.f ? 1.f : 0.f);
    return make_float2(1.f, 1.f);
}
__device__ __forceinline__ void up8p(const uint4 &p, float2 (&f)[4]) {
    const __nv_bfloat162 *h = (const __nv_bfloat162 *)&p;
#pragma unroll
    for (int q = 0; q < 4; ++q) f[q] = __bfloat1622float2(h[q]);
}
__device__ __forceinline__ uint4 pk8p(const float2 (&f)[4]) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(f[q].x, f[q].y);
    return *(uint4 *)h;
}
__device__ __forceinline__ void ld8p(const float *p, float2 (&f)[4]) {
    const float4 a = __ldg((const float4 *)p), b = __ldg((const float4 *)p + 1);
    f[0] = make_float2(a.x, a.y); f[1] = make_float2(a.z, a.w); f[2] = make_float2(b.x, b.y); f[3] = make_float2(b.z, b.w);
}

// block-level fold of K * 8 per-thread partials over the rsub lanes that share a channel group; the block's result is
// either WRITTEN to its own slot (deterministic two-stage reductions: out_off = blockIdx * K * C) or added atomically
template <int K, bool ATOMIC = true, bool DYN = false>
__device__ __forceinline__ void rowmap_fold(const RowMap &rm, int C, float (&acc)[K][8], float *__restrict__ out, long out_off) {
    __shared__ float red_static[DYN ? 1 : 384 * 8 * K];
    float *red = DYN ? (float *)g_rowpipe : red_static;      // DYN: the row pipeline's ring (drained) holds the scratch
    if (DYN) __syncthreads();
    const int G = C >> 3;
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) red[(k * 8 + j) * blockDim.x + threadIdx.x] = acc[k][j];
    __syncthreads();
    if (rm.rsub == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = 0.f;
                for (int q = 0; q < rm.rpb; ++q) t += red[(k * 8 + j) * blockDim.x + q * G + rm.g];
                if (ATOMIC) atomicAdd(out + out_off + (long)k * C + rm.c0 + j, t);
                else out[out_off + (long)k * C + rm.c0 + j] = t;
            }
    }
}

__global__ void __launch_bounds__(384)
bn_stats_kernel(const __nv_bfloat16 *__restrict__ Y, long M, int C, long ld, float *__restrict__ partials) {
    const RowMap rm(C);
    float2 s1[4], s2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) s1[j] = s2[j] = make_float2(0.f, 0.f);
    const void *const base[1] = {Y};
    const long pitch[1] = {ld * 2};
    row_stream<1>(base, pitch, rm.c0 * 2L, (long)blockIdx.x * rm.rpb + rm.rsub, (long)gridDim.x * rm.rpb, M,
                  [&](long, const uint4 (&v)[1]) {
        float2 y[4];
        up8q(v[0], y);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s1[j] = __fadd2_rn(s1[j], y[j]);
            s2[j] = __ffma2_rn(y[j], y[j], s2[j]);
        }
    });
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        acc[0][2 * j] = s1[j].x; acc[0][2 * j + 1] = s1[j].y;
        acc[1][2 * j] = s2[j].x; acc[1][2 * j + 1] = s2[j].y;
    }
    rowmap_fold<2, false, true>(rm, C, acc, partials, (long)blockIdx.x * 2 * C);
}

// Second stage of the two-stage reductions: block = 8 channels x 32 part lanes; every lane adds its share of the
// per-block partials in a fixed order (4 independent accumulators), shared memory folds the 32 lanes in a fixed order.
// Returns the two totals to the threads with part lane 0 (tid < 8).
__device__ __forceinline__ void fold_parts(const float *__restrict__ partials, int parts, int C, int c, float &s1, float &s2) {
    __shared__ float red[2][32][9];
    const int pl = threadIdx.x >> 3, cl = threadIdx.x & 7;      // 32 part lanes x 8 channels
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (c < C) {
        int q = pl;
        for (; q + 96 < parts; q += 128) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a[u] += partials[(long)(q + 32 * u) * 2 * C + c];
                b[u] += partials[(long)(q + 32 * u) * 2 * C + C + c];
            }
        }
        for (; q < parts; q += 32) {
            a[0] += partials[(long)q * 2 * C + c];
            b[0] += partials[(long)q * 2 * C + C + c];
        }
    }
    red[0][pl][cl] = (a[0] + a[1]) + (a[2] + a[3]);
    red[1][pl][cl] = (b[0] + b[1]) + (b[2] + b[3]);
    __syncthreads();
    s1 = s2 = 0.f;
    if (pl == 0)
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            s1 += red[0][q][cl];
            s2 += red[1][q][cl];
        }
}

// partials [parts, 2, C] -> mean_rstd [2, C], scale_shift [2, C] (z = y * scale + shift), running statistics (momentum,
// unbiased variance).  grid = ceil(C / 8), 256 threads.
__global__ void __launch_bounds__(256)
bn_finalize_kernel(const float *__restrict__ partials, int parts, float count, const float *__restrict__ gamma,
                   const float *__restrict__ beta, float eps, float momentum, int C, float *__restrict__ mean_rstd,
                   float *__restrict__ scale_shift, float *__restrict__ running_mean, float *__restrict__ running_var) {
    const int c = blockIdx.x * 8 + (threadIdx.x & 7);
    float s1, s2;
    fold_parts(partials, parts, C, c, s1, s2);
    if (threadIdx.x < 8 && c < C) {
        const float mean = s1 / count;
        const float var = fmaxf(s2 / count - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        mean_rstd[c] = mean;
        mean_rstd[C + c] = rstd;
        const float a = gamma[c] * rstd;
        scale_shift[c] = a;
        scale_shift[C + c] = beta[c] - mean * a;
        if (running_mean != nullptr) {
            running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.0f - momentum) * running_var[c] + momentum * var * (count / fmaxf(count - 1.0f, 1.0f));
        }
    }
}

// out[r, :] (pitch ld_out) = residual[r, :] + out_scale * act(y * scale + shift)
template <int ACT, bool RES>
__global__ void __launch_bounds__(384, RES ? 1 : 3)
bn_apply_kernel(const __nv_bfloat16 *__restrict__ Y, long M, int C, const float *__restrict__ scale_shift, float out_scale,
                const __nv_bfloat16 *__restrict__ residual, long ld_res, __nv_bfloat16 *__restrict__ out, long ld_out) {
    const RowMap rm(C);
    float2 a[4], b[4];
    ld8p(scale_shift + rm.c0, a);
    ld8p(scale_shift + C + rm.c0, b);
    const float2 os2 = make_float2(out_scale, out_scale);
    if (!RES) {
        // one input tensor and few live registers: ptxas keeps four plain loads in flight and 6 CTAs resident, which
        // measured slightly faster (70.6 vs 76.5 us on 131072 x 768) than the cp.async ring
        const long step = (long)gridDim.x * rm.rpb;
        auto body = [&](long r, const uint4 &vy) {
            float2 y[4];
            up8q(vy, y);
#pragma unroll
            for (int j = 0; j < 4; ++j) y[j] = __fmul2_rn(os2, actf2<ACT>(__ffma2_rn(y[j], a[j], b[j])));
            *(uint4 *)(out + r * ld_out + rm.c0) = pk8p(y);
        };
        long r = (long)blockIdx.x * rm.rpb + rm.rsub;
        for (; r + 3 * step < M; r += 4 * step) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ldg_batch(Y + (r + u * step) * C + rm.c0);
#pragma unroll
            for (int u = 0; u < 4; ++u) body(r + u * step, v[u]);
        }
        for (; r < M; r += step) body(r, ldg_batch(Y + r * C + rm.c0));
        return;
    }
    const void *const base[2] = {Y, residual};
    const long pitch[2] = {C * 2L, ld_res * 2};
    row_stream<2>(base, pitch, rm.c0 * 2L, (long)blockIdx.x * rm.rpb + rm.rsub, (long)gridDim.x * rm.rpb, M,
                  [&](long r, const uint4 (&v)[2]) {
        float2 y[4], q[4];
        up8q(v[0], y);
        up8q(v[1], q);
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = __fadd2_rn(__fmul2_rn(os2, actf2<ACT>(__ffma2_rn(y[j], a[j], b[j]))), q[j]);
        *(uint4 *)(out + r * ld_out + rm.c0) = pk8p(y);
    });
}

// out = act(y * scale + shift) for the layer in front of an SE / ECA block, plus the squeeze of that block: per-(image, row
// block) channel sums of the bf16-ROUNDED outputs, pool [B, gridDim.x, C] (written, not accumulated: the consumers fold the
// parts in a fixed order).  grid (row blocks, image); the separate pool_sum pass over `out` disappears.
template <int ACT>
__global__ void __launch_bounds__(384, 2)
bn_apply_pool_kernel(const __nv_bfloat16 *__restrict__ Y, long HW, int C, const float *__restrict__ scale_shift,
                     __nv_bfloat16 *__restrict__ out, float *__restrict__ pool) {
    const RowMap rm(C);
    const long base = (long)blockIdx.y * HW;
    float2 a[4], b[4], s[4];
    ld8p(scale_shift + rm.c0, a);
    ld8p(scale_shift + C + rm.c0, b);
#pragma unroll
    for (int j = 0; j < 4; ++j) s[j] = make_float2(0.f, 0.f);
    const long step = (long)gridDim.x * rm.rpb;
    auto body = [&](long r, const uint4 &vy) {
        float2 y[4], q[4];
        up8q(vy, y);
#pragma unroll
        for (int j = 0; j < 4; ++j) y[j] = actf2<ACT>(__ffma2_rn(y[j], a[j], b[j]));
        const uint4 pk = pk8p(y);
        *(uint4 *)(out + (base + r) * C + rm.c0) = pk;
        up8q(pk, q);
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = __fadd2_rn(s[j], q[j]);
    };
    long r = (long)blockIdx.x * rm.rpb + rm.rsub;
    for (; r + 3 * step < HW; r += 4 * step) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg_batch(Y + (base + r + u * step) * C + rm.c0);
#pragma unroll
        for (int u = 0; u < 4; ++u) body(r + u * step, v[u]);
    }
    for (; r < HW; r += step) body(r, ldg_batch(Y + (base + r) * C + rm.c0));
    float acc[1][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[0][2 * j] = s[j].x; acc[0][2 * j + 1] = s[j].y; }
    rowmap_fold<1, false>(rm, C, acc, pool, ((long)blockIdx.y * gridDim.x + blockIdx.x) * C);
}

// pass 1 of the backward: sums2[c] += dz, sums2[C + c] += dz * xhat, dz = dA * out_scale * act'(z)
template <int ACT>
__global__ void __launch_bounds__(384)
bn_bwd_reduce_kernel(const __nv_bfloat16 *__restrict__ dA, long ld_da, const __nv_bfloat16 *__restrict__ Y, long M, int C,
                     const float *__restrict__ scale_shift, const float *__restrict__ mean_rstd, float out_scale,
                     float *__restrict__ partials) {
    // The second sum is accumulated as sum(dz * y); the coefficient kernel turns it into sum(dz * xhat) =
    // rstd * (sum(dz * y) - mean * sum(dz)).  That drops 16 per-channel constants from the registers of this kernel
    // (80 -> 4 resident CTAs per SM instead of 3) and one FMA per element.
    const RowMap rm(C);
    float2 a[4], b[4];
    ld8p(scale_shift + rm.c0, a);
    ld8p(scale_shift + C + rm.c0, b);
    float2 s1[4], s2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) s1[j] = s2[j] = make_float2(0.f, 0.f);
    const float2 os2 = make_float2(out_scale, out_scale);
    const void *const base[2] = {Y, dA};
    const long pitch[2] = {C * 2L, ld_da * 2};
    row_stream<2>(base, pitch, rm.c0 * 2L, (long)blockIdx.x * rm.rpb + rm.rsub, (long)gridDim.x * rm.rpb, M,
                  [&](long, const uint4 (&v)[2]) {
        float2 y[4], d[4];
        up8q(v[0], y);
        up8q(v[1], d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 dz = __fmul2_rn(__fmul2_rn(d[j], os2), dactf2<ACT>(__ffma2_rn(y[j], a[j], b[j])));
            s1[j] = __fadd2_rn(s1[j], dz);
            s2[j] = __ffma2_rn(dz, y[j], s2[j]);
        }
    });
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        acc[0][2 * j] = s1[j].x; acc[0][2 * j + 1] = s1[j].y;
        acc[1][2 * j] = s2[j].x; acc[1][2 * j + 1] = s2[j].y;
    }
    rowmap_fold<2, false, true>(rm, C, acc, partials, (long)blockIdx.x * 2 * C);
}

// between the passes: fold the per-block partials in a fixed order, accumulate dgamma += s2, dbeta += s1, and emit the two
// per-channel coefficients of pass 2 (dy = a dz - k2 y - k3)
__global__ void __launch_bounds__(256)
bn_bwd_coef_kernel(const float *__restrict__ partials, int parts, float inv_m, const float *__restrict__ scale_shift,
                   const float *__restrict__ mean_rstd, int C, float *__restrict__ coef, float *__restrict__ dgamma,
                   float *__restrict__ dbeta) {
    const int c = blockIdx.x * 8 + (threadIdx.x & 7);
    float s1, s2;
    fold_parts(partials, parts, C, c, s1, s2);
    if (threadIdx.x < 8 && c < C) {
        const float a = scale_shift[c], mean = mean_rstd[c], rstd = mean_rstd[C + c];
        s2 = rstd * (s2 - mean * s1);                      // sum(dz * y) -> sum(dz * xhat)
        dgamma[c] += s2;
        dbeta[c] += s1;
        const float c1 = s1 * inv_m, c2 = s2 * inv_m;
        coef[c] = a * c2 * rstd;
        coef[C + c] = a * (c1 - c2 * mean * rstd);
    }
}

// pass 2: dY = gamma * rstd * (dz - s1 / M - xhat * s2 / M) = a dz - k2 y - k3
template <int ACT>
__global__ void __launch_bounds__(384)
bn_bwd_apply_kernel(const __nv_bfloat16 *__restrict__ dA, long ld_da, const __nv_bfloat16 *__restrict__ Y, long M, int C,
                    const float *__restrict__ scale_shift, const float *__restrict__ coef, float out_scale,
                    __nv_bfloat16 *__restrict__ dY) {
    const RowMap rm(C);
    float2 a[4], b[4], k2[4], k3[4];
    ld8p(scale_shift + rm.c0, a);
    ld8p(scale_shift + C + rm.c0, b);
    ld8p(coef + rm.c0, k2);
    ld8p(coef + C + rm.c0, k3);
    const float2 os2 = make_float2(out_scale, out_scale);
    const void *const base[2] = {Y, dA};
    const long pitch[2] = {C * 2L, ld_da * 2};
    row_stream<2>(base, pitch, rm.c0 * 2L, (long)blockIdx.x * rm.rpb + rm.rsub, (long)gridDim.x * rm.rpb, M,
                  [&](long r, const uint4 (&v)[2]) {
        float2 y[4], d[4];
        up8q(v[0], y);
        up8q(v[1], d);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 dz = __fmul2_rn(__fmul2_rn(d[j], os2), dactf2<ACT>(__ffma2_rn(y[j], a[j], b[j])));
            const float2 t = __ffma2_rn(k2[j], y[j], k3[j]);
            d[j] = __ffma2_rn(a[j], dz, make_float2(-t.x, -t.y));              // a dz - (k2 y + k3)
        }
        *(uint4 *)(dY + r * C + rm.c0) = pk8p(d);
    });
}

// ---- depthwise 3x3 backward ---------------------------------------------------------------------------------
// dX[b, iy, ix, c] = add + sum_{ky,kx} dY[b, oy, ox, c] * w[ky*3+kx][c], oy * s + ky - 1 = iy (likewise x)
__global__ void __launch_bounds__(256)
dwconv_bwd_data_kernel(const __nv_bfloat16 *__restrict__ dY, const float *__restrict__ Wd, int H, int W, int C, int Ho,
                       int Wo, int stride, const __nv_bfloat16 *__restrict__ add, long total8, __nv_bfloat16 *__restrict__ dX) {
    const int C8 = C >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long p = i / C8;
        const int ix = (int)(p % W);
        p /= W;
        const int iy = (int)(p % H);
        const long b = p / H;
        float acc[8];
        if (add != nullptr) up8(__ldg((const uint4 *)add + i), acc);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        }
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int ty = iy + 1 - ky;
            if (ty < 0 || (ty % stride) != 0) continue;
            const int oy = ty / stride;
            if (oy >= Ho) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int tx = ix + 1 - kx;
                if (tx < 0 || (tx % stride) != 0) continue;
                const int ox = tx / stride;
                if (ox >= Wo) continue;
                float d[8];
                up8(__ldg((const uint4 *)(dY + ((b * Ho + oy) * Wo + ox) * C + cg * 8)), d);
                const float *w = Wd + (long)(ky * 3 + kx) * C + cg * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(d[j], __ldg(w + j), acc[j]);
            }
        }
        ((uint4 *)dX)[i] = pk8(acc);
    }
}

// Stride-2 data gradient, one CTA per input row: an input pixel receives 1, 2 or 4 taps depending on the parity of its
// coordinates (even: the centre tap; odd: the two outer taps), so the tap list is decided per row / per column instead of
// testing all nine taps with runtime divisions per element (the generic kernel above ran at ~1.3 TB/s on the two stride-2
// layers of stages 1 and 2: 1.1 ms per step).  A thread walks (pixel, 8-channel group) items of the row: consecutive
// threads write consecutive 16-byte chunks.
__global__ void __launch_bounds__(256)
dwconv_bwd_data_s2_kernel(const __nv_bfloat16 *__restrict__ dY, const float *__restrict__ Wd, int H, int W, int C, int Ho, int Wo,
                          const __nv_bfloat16 *__restrict__ add, __nv_bfloat16 *__restrict__ dX) {
    const int C8 = C >> 3;
    const long row = blockIdx.x;                 // b * H + iy
    const int iy = (int)(row % H);
    const long b = row / H;
    // rows of dY that reach this input row: oy * 2 + ky - 1 = iy
    int ny = 0, kys[2], oys[2];
    if (iy & 1) {
        const int o0 = (iy + 1) >> 1, o1 = (iy - 1) >> 1;
        if (o0 < Ho) { kys[ny] = 0; oys[ny] = o0; ++ny; }
        if (o1 < Ho) { kys[ny] = 2; oys[ny] = o1; ++ny; }
    } else if ((iy >> 1) < Ho) {
        kys[0] = 1; oys[0] = iy >> 1; ny = 1;
    }
    const int items = W * C8;
    __nv_bfloat16 *out = dX + row * (long)W * C;
    const __nv_bfloat16 *addr = add != nullptr ? add + row * (long)W * C : nullptr;
    for (int e = threadIdx.x; e < items; e += blockDim.x) {
        const int ix = e / C8, cg = e - ix * C8;
        float acc[8];
        if (addr != nullptr) up8(__ldg((const uint4 *)addr + e), acc);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        }
        int nx = 0, kxs[2], oxs[2];
        if (ix & 1) {
            const int o0 = (ix + 1) >> 1, o1 = (ix - 1) >> 1;
            if (o0 < Wo) { kxs[nx] = 0; oxs[nx] = o0; ++nx; }
            if (o1 < Wo) { kxs[nx] = 2; oxs[nx] = o1; ++nx; }
        } else if ((ix >> 1) < Wo) {
            kxs[0] = 1; oxs[0] = ix >> 1; nx = 1;
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            if (a >= ny) break;
            const __nv_bfloat16 *drow = dY + ((b * Ho + oys[a]) * Wo) * (long)C + cg * 8;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q >= nx) break;
                float d[8], w[8];
                up8(__ldg((const uint4 *)(drow + (long)oxs[q] * C)), d);
                ld8f(Wd + (long)(kys[a] * 3 + kxs[q]) * C + cg * 8, w);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(d[j], w[j], acc[j]);
            }
        }
        ((uint4 *)out)[e] = pk8(acc);
    }
}

// dW[c * 9 + k] += sum_{b,oy,ox} dY[b,oy,ox,c] * X[b, oy*s+ky-1, ox*s+kx-1, c]   (parameter layout [C,1,3,3])
// Same tiling as the forward kernel: a CTA walks output tiles (TH x TW pixels x 64 channels), stages the input halo
// tile in shared memory once, and every thread keeps 9 x 8 fp32 partials in registers across all its tiles; one fold
// and 576 atomics per CTA at the end.
template <int STRIDE>
struct DwT {
    static constexpr int TH = STRIDE == 1 ? 8 : 4, TW = STRIDE == 1 ? 16 : 8;
    static constexpr int IH = (TH - 1) * STRIDE + 3, IW = (TW - 1) * STRIDE + 3;
};
// the input halo tile is one 4-D TMA box (zero fill outside the image = the padding), as in the forward kernel
template <int STRIDE>
__device__ __forceinline__ void dww_stage(unsigned char *xbuf, const CUtensorMap *map, uint64_t *bar, int b, int tile, int tiles_x,
                                          int c_slab) {
    using T = DwT<STRIDE>;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    mbar_expect_tx(bar, T::IH * T::IW * 128);
    tma_load_4d(xbuf, map, bar, c_slab, tx * T::TW * STRIDE - 1, ty * T::TH * STRIDE - 1, b);
}
// four bf16 channels as two fp32 pairs (operands of the packed FFMA2: two fused multiply-adds per issue slot)
__device__ __forceinline__ void up4(const uint2 &p, float2 (&f)[2]) {
    f[0] = make_float2(__uint_as_float(p.x << 16), __uint_as_float(p.x & 0xffff0000u));
    f[1] = make_float2(__uint_as_float(p.y << 16), __uint_as_float(p.y & 0xffff0000u));
}
// A thread owns FOUR channels of one tile column (16 channel groups x 16 pixel lanes): 36 fp32 partials instead of 72, and
// the output gradients -- each used by exactly one thread -- come straight from global memory into registers while the
// halo tile is in flight, so three CTAs share an SM (was one, at 136 registers and 79 KB of shared memory).
template <int STRIDE>
__global__ void __launch_bounds__(256, 3)
dwconv_bwd_weight_kernel(const __nv_bfloat16 *__restrict__ dY, const __grid_constant__ CUtensorMap mapX, int B, int H, int W, int C,
                         int Ho, int Wo, int tiles_x, int tiles_y, float *__restrict__ dW) {
    using T = DwT<STRIDE>;
    constexpr int kX = T::IH * T::IW * 128;
    constexpr int kPix = T::TH * T::TW, kMine = kPix / 16;      // output pixels per thread
    extern __shared__ __align__(128) unsigned char s_dyn[];     // 2 input halo tiles; the fold reuses them
    __shared__ uint64_t s_bar[2];
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int cgi = threadIdx.x & 15, pl = threadIdx.x >> 4;
    const int c_slab = blockIdx.y * 64;
    const int c0 = c_slab + cgi * 4;
    float2 acc[9][2];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[t][j] = make_float2(0.f, 0.f);
    const int per_img = tiles_x * tiles_y;
    const long n_tiles = (long)B * per_img;
    long item = blockIdx.x;
    if (item < n_tiles && threadIdx.x == 0)
        dww_stage<STRIDE>(s_dyn, &mapX, &s_bar[0], (int)(item / per_img), (int)(item % per_img), tiles_x, c_slab);
    for (int it = 0; item < n_tiles; item += gridDim.x, ++it) {
        const long next = item + gridDim.x;
        if (next < n_tiles && threadIdx.x == 0)
            dww_stage<STRIDE>(s_dyn + ((it + 1) & 1) * kX, &mapX, &s_bar[(it + 1) & 1], (int)(next / per_img), (int)(next % per_img),
                              tiles_x, c_slab);
        // this thread's output gradients (zero outside the image / channel range: they contribute nothing)
        const int b = (int)(item / per_img), tile = (int)(item - (long)b * per_img);
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int oy0 = ty * T::TH, ox0 = tx * T::TW;
        uint2 dyr[kMine];
#pragma unroll
        for (int q = 0; q < kMine; ++q) {
            const int p = STRIDE == 1 ? q * T::TW + pl : q * 16 + pl;     // stride 1: column pl, row q
            const int py = p / T::TW, pxx = p - py * T::TW;
            const int oy = oy0 + py, ox = ox0 + pxx;
            dyr[q] = make_uint2(0u, 0u);
            if (oy < Ho && ox < Wo && c0 < C) dyr[q] = __ldg((const uint2 *)(dY + (((long)b * Ho + oy) * Wo + ox) * C + c0));
        }
        mbar_wait(&s_bar[it & 1], (it >> 1) & 1);
        const unsigned char *s_in = s_dyn + (it & 1) * kX + cgi * 8;
        if (STRIDE == 1) {
            // two vertically adjacent outputs share their 4 x 3 input vectors (12 loads instead of 18)
            const int pxx = pl;
#pragma unroll
            for (int q = 0; q < kMine / 2; ++q) {
                const int r0 = q * 2;
                float2 d0[2], d1[2];
                up4(dyr[r0], d0);
                up4(dyr[r0 + 1], d1);
#pragma unroll
                for (int ir = 0; ir < 4; ++ir)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        float2 x[2];
                        up4(*(const uint2 *)(s_in + ((r0 + ir) * T::IW + pxx + kx) * 128), x);
                        if (ir < 3) {
#pragma unroll
                            for (int j = 0; j < 2; ++j) acc[ir * 3 + kx][j] = __ffma2_rn(d0[j], x[j], acc[ir * 3 + kx][j]);
                        }
                        if (ir > 0) {
#pragma unroll
                            for (int j = 0; j < 2; ++j) acc[(ir - 1) * 3 + kx][j] = __ffma2_rn(d1[j], x[j], acc[(ir - 1) * 3 + kx][j]);
                        }
                    }
            }
        } else {
#pragma unroll
            for (int q = 0; q < kMine; ++q) {
                const int p = q * 16 + pl;
                const int py = p / T::TW, pxx = p - py * T::TW;
                float2 d[2];
                up4(dyr[q], d);
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        float2 x[2];
                        up4(*(const uint2 *)(s_in + ((py * STRIDE + ky) * T::IW + pxx * STRIDE + kx) * 128), x);
#pragma unroll
                        for (int j = 0; j < 2; ++j) acc[ky * 3 + kx][j] = __ffma2_rn(d[j], x[j], acc[ky * 3 + kx][j]);
                    }
            }
        }
        __syncthreads();
    }
    float(*red)[65] = (float(*)[65])s_dyn;
#pragma unroll                        // (a rolled loop would index acc dynamically and push it to local memory)
    for (int t = 0; t < 9; ++t) {
        __syncthreads();
        red[pl][cgi * 4 + 0] = acc[t][0].x; red[pl][cgi * 4 + 1] = acc[t][0].y;
        red[pl][cgi * 4 + 2] = acc[t][1].x; red[pl][cgi * 4 + 3] = acc[t][1].y;
        __syncthreads();
        if (threadIdx.x < 64 && c_slab + threadIdx.x < C) {
            float sacc = 0.f;
#pragma unroll 8
            for (int q = 0; q < 16; ++q) sacc += red[q][threadIdx.x];
            atomicAdd(dW + (long)(c_slab + threadIdx.x) * 9 + t, sacc);
        }
    }
}

// ---- SE / ECA gates -------------------------------------------------------------------------------------------
// out[b, c] += sum_p dOut[b, p, c] * X[b, p, c]   (gradient reaching the gate); grid (row blocks, B)
__global__ void __launch_bounds__(384)
gate_bwd_reduce_kernel(const __nv_bfloat16 *__restrict__ dO, const __nv_bfloat16 *__restrict__ X, long HW, int C,
                       float *__restrict__ out) {
    const RowMap rm(C);
    const long base = (long)blockIdx.y * HW;
    float acc[1][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
    const void *const src[2] = {dO + base * C, X + base * C};
    const long pitch[2] = {C * 2L, C * 2L};
    float2 a2[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) a2[j] = make_float2(0.f, 0.f);
    row_stream<2>(src, pitch, rm.c0 * 2L, (long)blockIdx.x * rm.rpb + rm.rsub, (long)gridDim.x * rm.rpb, HW,
                  [&](long, const uint4 (&v)[2]) {
        float2 d[4], x[4];
        up8q(v[0], d);
        up8q(v[1], x);
#pragma unroll
        for (int j = 0; j < 4; ++j) a2[j] = __ffma2_rn(d[j], x[j], a2[j]);
    });
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[0][2 * j] = a2[j].x; acc[0][2 * j + 1] = a2[j].y; }
    rowmap_fold<1, true, true>(rm, C, acc, out, (long)blockIdx.y * C);
}

// dX[b,p,c] = add + dOut[b,p,c] * gate[b,c] + dmean[b,c] * inv_hw
// grid (row blocks, B); a thread owns one 8-channel group of its image for its whole life (gate and dmean / HW in
// registers) and walks rows -- no per-element index arithmetic (the flat-index version decoded (image, channel group) with
// 64-bit divisions per element and ran at half the bandwidth of its neighbours)
__global__ void __launch_bounds__(384, 2)
gate_bwd_apply_kernel(const __nv_bfloat16 *__restrict__ dO, const float *__restrict__ gate, const __nv_bfloat16 *__restrict__ dmean,
                      float inv_hw, long HW, int C, const __nv_bfloat16 *__restrict__ add, __nv_bfloat16 *__restrict__ dX) {
    const RowMap rm(C);
    const long b = blockIdx.y, base = b * HW;
    float g[8], m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        g[j] = gate != nullptr ? __ldg(gate + b * C + rm.c0 + j) : 1.0f;
        m[j] = 0.f;
    }
    if (dmean != nullptr) {
        up8(__ldg((const uint4 *)(dmean + b * C + rm.c0)), m);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] *= inv_hw;
    }
    const long step = (long)gridDim.x * rm.rpb;
    long r = (long)blockIdx.x * rm.rpb + rm.rsub;
    if (dO == nullptr) {                               // broadcast of dmean / HW (+ add): gradient of a global average
        for (; r < HW; r += step) {
            const long off = (base + r) * C + rm.c0;
            float d[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = m[j];
            if (add != nullptr) {
                float a[8];
                up8(__ldg((const uint4 *)(add + off)), a);
#pragma unroll
                for (int j = 0; j < 8; ++j) d[j] += a[j];
            }
            *(uint4 *)(dX + off) = pk8(d);
        }
        return;
    }
    auto body = [&](long off, const uint4 &vd, const uint4 &va) {
        float d[8];
        up8(vd, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = fmaf(d[j], g[j], m[j]);
        if (add != nullptr) {
            float a[8];
            up8(va, a);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] += a[j];
        }
        *(uint4 *)(dX + off) = pk8(d);
    };
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    for (; r + 3 * step < HW; r += 4 * step) {         // a batch of four (eight with `add`) 16-byte loads in flight per thread
        uint4 vd[4], va[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long off = (base + r + u * step) * C + rm.c0;
            vd[u] = ldg_batch(dO + off);
            va[u] = add != nullptr ? ldg_batch(add + off) : z4;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) body((base + r + u * step) * C + rm.c0, vd[u], va[u]);
    }
    for (; r < HW; r += step) {
        const long off = (base + r) * C + rm.c0;
        body(off, ldg_batch(dO + off), add != nullptr ? ldg_batch(add + off) : z4);
    }
}

// The same with the BatchNorm-backward tail of the layer that produced the gated map (the depthwise ConvBnAct in front of an
// SE / ECA block): dA = dO * gate + dmean / HW never reaches memory; with that layer's saved conv output y the kernel writes
// dz = dA * act'(y * scale + shift) and the per-CTA partial sums of dz and dz * y ([gridDim.y * gridDim.x, 2, C]): the first
// pass of pose_bn_bwd_bf16 (one read of dA and y) disappears.
template <int ACT>
__global__ void __launch_bounds__(384, 2)
gate_bwd_apply_bn_kernel(const __nv_bfloat16 *__restrict__ dO, const float *__restrict__ gate, const __nv_bfloat16 *__restrict__ dmean,
                         float inv_hw, long HW, int C, const __nv_bfloat16 *__restrict__ Ybn, const float *__restrict__ scale_shift,
                         __nv_bfloat16 *__restrict__ dZ, float *__restrict__ partials) {
    const RowMap rm(C);
    const long b = blockIdx.y, base = b * HW;
    float2 g[4], m[4], a[4], sh[4], s1[4], s2[4];
    ld8p(gate + b * C + rm.c0, g);
    up8q(__ldg((const uint4 *)(dmean + b * C + rm.c0)), m);
    ld8p(scale_shift + rm.c0, a);
    ld8p(scale_shift + C + rm.c0, sh);
    const float2 ih2 = make_float2(inv_hw, inv_hw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        m[j] = __fmul2_rn(m[j], ih2);
        s1[j] = s2[j] = make_float2(0.f, 0.f);
    }
    const void *const src[2] = {dO + base * C, Ybn + base * C};
    const long pitch[2] = {C * 2L, C * 2L};
    row_stream<2>(src, pitch, rm.c0 * 2L, (long)blockIdx.x * rm.rpb + rm.rsub, (long)gridDim.x * rm.rpb, HW,
                  [&](long r, const uint4 (&v)[2]) {
        float2 d[4], y[4];
        up8q(v[0], d);
        up8q(v[1], y);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 dz = __fmul2_rn(__ffma2_rn(d[j], g[j], m[j]), dactf2<ACT>(__ffma2_rn(y[j], a[j], sh[j])));
            s1[j] = __fadd2_rn(s1[j], dz);
            s2[j] = __ffma2_rn(dz, y[j], s2[j]);
            d[j] = dz;
        }
        *(uint4 *)(dZ + (base + r) * C + rm.c0) = pk8p(d);
    });
    float acc[2][8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        acc[0][2 * j] = s1[j].x; acc[0][2 * j + 1] = s1[j].y;
        acc[1][2 * j] = s2[j].x; acc[1][2 * j + 1] = s2[j].y;
    }
    rowmap_fold<2, false, true>(rm, C, acc, partials, ((long)blockIdx.y * gridDim.x + blockIdx.x) * 2 * C);
}

// dz[b,c] (bf16) = dgate[b,c] * g (1 - g)     (through the sigmoid of the SE gate; operand of the SE backward GEMMs)
__global__ void __launch_bounds__(256)
sigmoid_bwd_kernel(const float *__restrict__ dgate, const float *__restrict__ gate, long n, __nv_bfloat16 *__restrict__ dz) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float g = gate[i];
        dz[i] = __float2bfloat16_rn(dgate[i] * g * (1.0f - g));
    }
}

// ECA backward, one CTA per image.  mode 0: gate multiplies a map (dgate given);  mode 1: feat = mean * gate (the
// global_features tail): dgate = dfeat * mean and dmean gets the direct term dfeat * gate.
__global__ void __launch_bounds__(256)
eca_bwd_kernel(const float *__restrict__ dgate_in, const __nv_bfloat16 *__restrict__ dfeat, const float *__restrict__ gate,
               const float *__restrict__ pool, int parts, float inv_hw, const float *__restrict__ w, int k, int C, int mode,
               __nv_bfloat16 *__restrict__ dmean_out, float *__restrict__ dw) {
    extern __shared__ float sm[];   // mean[C], dz[C]
    float *mean = sm, *dz = sm + C;
    __shared__ float wred[8];
    const int b = blockIdx.x, half = (k - 1) / 2;
    for (int c = threadIdx.x; c < C; c += 256) {
        float t = 0.f;
        for (int q = 0; q < parts; ++q) t += pool[((long)b * parts + q) * C + c];
        mean[c] = t * inv_hw;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        const float g = gate[(long)b * C + c];
        const float dg = mode == 0 ? dgate_in[(long)b * C + c] : __bfloat162float(dfeat[(long)b * C + c]) * mean[c];
        dz[c] = dg * g * (1.0f - g);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;                            // z[c'] = sum_t w[t] mean[c' + t - half]  ->  dmean[c] = sum_t w[t] dz[c - t + half]
        for (int t = 0; t < k; ++t) {
            const int cc = c - t + half;
            if (cc >= 0 && cc < C) s = fmaf(__ldg(w + t), dz[cc], s);
        }
        if (mode == 1) s += __bfloat162float(dfeat[(long)b * C + c]) * gate[(long)b * C + c];
        dmean_out[(long)b * C + c] = __float2bfloat16_rn(s);
    }
    for (int t = 0; t < k; ++t) {                 // dw[t] += sum_c dz[c] * mean[c + t - half]
        float s = 0.f;
        for (int c = threadIdx.x; c < C; c += 256) {
            const int cc = c + t - half;
            if (cc >= 0 && cc < C) s = fmaf(dz[c], mean[cc], s);
        }
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) wred[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tt = 0.f;
            for (int q = 0; q < 8; ++q) tt += wred[q];
            atomicAdd(dw + t, tt);
        }
        __syncthreads();
    }
}

// ---- CoordAttention backward ------------------------------------------------------------------------------
// dZ [B, H+W, 2C] bf16 = gradient at the INPUT of the sigmoids of G (zero in the halves the forward does not use):
//   rows h < H, cols c      : (sum_w dOut * X * a_w) * a_h (1 - a_h)
//   rows H + w, cols C + c  : (sum_h dOut * X * a_h) * a_w (1 - a_w)
__global__ void __launch_bounds__(256, 2)
coord_bwd_reduce_kernel(const __nv_bfloat16 *__restrict__ dO, const __nv_bfloat16 *__restrict__ X, const __nv_bfloat16 *__restrict__ G,
                        int H, int W, int C, __nv_bfloat16 *__restrict__ dZ) {
    // grid (image, chunk): the (H + W) x C / 8 reductions of an image are spread over gridDim.y CTAs (one CTA per image left
    // 20 SMs idle and every thread walking eight strictly serial reductions: 71 us for 134 MB)
    const int b = blockIdx.x, C8 = C >> 3;
    const int n_rows = H + W;
    for (int i = blockIdx.y * 256 + threadIdx.x; i < n_rows * C8; i += gridDim.y * 256) {
        const int r = i / C8, cg = i - r * C8;
        const bool is_h = r < H;
        const int n = is_h ? W : H;
        float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        auto offs = [&](int q, long &off, long &grow) {
            const int y = is_h ? r : q, x = is_h ? q : r - H;
            off = (((long)b * H + y) * W + x) * C + cg * 8;
            // the OTHER direction's gate at this pixel
            grow = (is_h ? ((long)b * n_rows + H + x) * 2L * C + C : ((long)b * n_rows + y) * 2L * C) + cg * 8;
        };
        auto acc = [&](const uint4 &vd, const uint4 &vx, const uint4 &vg) {
            float d[8], xv[8], go[8];
            up8(vd, d);
            up8(vx, xv);
            up8(vg, go);
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = fmaf(d[j] * xv[j], go[j], s[j]);
        };
        int q = 0;
        for (; q + 3 < n; q += 4) {                       // twelve independent 16-byte loads in flight
            uint4 vd[4], vx[4], vg[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                long off, grow;
                offs(q + u, off, grow);
                vd[u] = ldg_batch(dO + off);
                vx[u] = ldg_batch(X + off);
                vg[u] = ldg_batch(G + grow);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc(vd[u], vx[u], vg[u]);
        }
        for (; q < n; ++q) {
            long off, grow;
            offs(q, off, grow);
            acc(ldg_batch(dO + off), ldg_batch(X + off), ldg_batch(G + grow));
        }
        float gs[8], z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long row = ((long)b * n_rows + r) * 2L * C;
        up8(__ldg((const uint4 *)(G + row + (is_h ? 0 : C) + cg * 8)), gs);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] *= gs[j] * (1.0f - gs[j]);
        *(uint4 *)(dZ + row + (is_h ? 0 : C) + cg * 8) = pk8(s);
        *(uint4 *)(dZ + row + (is_h ? C : 0) + cg * 8) = pk8(z);
    }
}

// dX = dOut * a_h * a_w + dP[b, h, c] / W + dP[b, H + w, c] / H
__global__ void __launch_bounds__(256)
coord_bwd_apply_kernel(const __nv_bfloat16 *__restrict__ dO, const __nv_bfloat16 *__restrict__ G, const __nv_bfloat16 *__restrict__ dP,
                       int H, int W, int C, long total8, __nv_bfloat16 *__restrict__ dX) {
    const int C8 = C >> 3;
    const float iw = 1.0f / (float)W, ih = 1.0f / (float)H;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long p = i / C8;
        const int x = (int)(p % W);
        p /= W;
        const int y = (int)(p % H);
        const long b = p / H;
        float d[8], gh[8], gw[8], ph[8], pw[8];
        up8(__ldg((const uint4 *)dO + i), d);
        up8(__ldg((const uint4 *)(G + ((b * (H + W) + y) * 2L * C) + cg * 8)), gh);
        up8(__ldg((const uint4 *)(G + ((b * (H + W) + H + x) * 2L * C) + C + cg * 8)), gw);
        up8(__ldg((const uint4 *)(dP + (b * (H + W) + y) * (long)C + cg * 8)), ph);
        up8(__ldg((const uint4 *)(dP + (b * (H + W) + H + x) * (long)C + cg * 8)), pw);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = d[j] * gh[j] * gw[j] + ph[j] * iw + pw[j] * ih;
        ((uint4 *)dX)[i] = pk8(d);
    }
}

// ---- WASP branch mix (cnn.py:451-479) -----------------------------------------------------------------------
// out = sum_i w_i * branch_i + w_last * global[b, c],  w = softmax(raw); branches [nb, B*HW, C] contiguous
__global__ void __launch_bounds__(256)
wasp_mix_kernel(const __nv_bfloat16 *__restrict__ branches, int nb, long branch_stride, const __nv_bfloat16 *__restrict__ glob,
                const float *__restrict__ raw, long HW, int C, long total8, __nv_bfloat16 *__restrict__ out) {
    __shared__ float w[8];
    if (threadIdx.x == 0) {
        float m = -1e30f, s = 0.f;
        for (int i = 0; i <= nb; ++i) m = fmaxf(m, raw[i]);
        for (int i = 0; i <= nb; ++i) s += __expf(raw[i] - m);
        for (int i = 0; i <= nb; ++i) w[i] = __expf(raw[i] - m) / s;
    }
    __syncthreads();
    const int C8 = C >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        const long b = (i / C8) / HW;
        float acc[8], f[8];
        up8(__ldg((const uint4 *)(glob + b * C + cg * 8)), acc);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] *= w[nb];
        for (int q = 0; q < nb; ++q) {
            up8(__ldg((const uint4 *)(branches + q * branch_stride) + i), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(w[q], f[j], acc[j]);
        }
        ((uint4 *)out)[i] = pk8(acc);
    }
}

// backward: dbranch_i = w_i * dOut (written), dglob[b,c] = w_last * sum_p dOut (atomic, fp32 [B,C]),
// dots[i] += <dOut, branch_i> (i < nb), dots[nb] += <dOut, glob broadcast>   (fp32 [nb+1], zeroed by the caller)
__global__ void __launch_bounds__(384, 2)
wasp_mix_bwd_kernel(const __nv_bfloat16 *__restrict__ dO, const __nv_bfloat16 *__restrict__ branches, int nb, long branch_stride,
                    const __nv_bfloat16 *__restrict__ glob, const float *__restrict__ raw, long HW, int C,
                    __nv_bfloat16 *__restrict__ dbranches, float *__restrict__ dglob, float *__restrict__ dots) {
    // grid (row blocks, image); a thread owns one 8-channel group of its image and walks rows: the broadcast branch's
    // gradient (a per-image column sum of dO) accumulates in registers and is folded once per CTA -- the flat-index
    // version issued one atomicAdd per ELEMENT (33 M per step) and decoded (image, group) with 64-bit divisions
    __shared__ float w[8];
    __shared__ float sdots[8];
    if (threadIdx.x == 0) {
        float m = -1e30f, s = 0.f;
        for (int i = 0; i <= nb; ++i) m = fmaxf(m, raw[i]);
        for (int i = 0; i <= nb; ++i) s += __expf(raw[i] - m);
        for (int i = 0; i <= nb; ++i) w[i] = __expf(raw[i] - m) / s;
    }
    if (threadIdx.x < 8) sdots[threadIdx.x] = 0.f;
    __syncthreads();
    const RowMap rm(C);
    const long b = blockIdx.y, base = b * HW;
    float g8[8], acc[1][8];
    up8(__ldg((const uint4 *)(glob + b * C + rm.c0)), g8);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
    float dot[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long step = (long)gridDim.x * rm.rpb;
    for (long r = (long)blockIdx.x * rm.rpb + rm.rsub; r < HW; r += step) {
        const long off = (base + r) * C + rm.c0;
        uint4 vb[7];
        const uint4 vd = ldg_batch(dO + off);
#pragma unroll
        for (int q = 0; q < 7; ++q)
            if (q < nb) vb[q] = ldg_batch(branches + q * branch_stride + off);
        float d[8];
        up8(vd, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            dot[7] = fmaf(d[j], g8[j], dot[7]);
            acc[0][j] += d[j];
        }
#pragma unroll
        for (int q = 0; q < 7; ++q)
            if (q < nb) {
                float f[8], o[8];
                up8(vb[q], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    dot[q] = fmaf(d[j], f[j], dot[q]);
                    o[j] = w[q] * d[j];
                }
                *(uint4 *)(dbranches + q * branch_stride + off) = pk8(o);
            }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] *= w[nb];
    rowmap_fold<1>(rm, C, acc, dglob, b * C);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float s = warp_sum(dot[q]);
        if ((threadIdx.x & 31) == 0 && (q < nb || q == 7)) atomicAdd(&sdots[q == 7 ? nb : q], s);
    }
    __syncthreads();
    if (threadIdx.x <= nb) atomicAdd(dots + threadIdx.x, sdots[threadIdx.x]);
}

__global__ void wasp_weights_bwd_kernel(const float *__restrict__ raw, const float *__restrict__ dots, int n, float *__restrict__ draw) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float w[8], m = -1e30f, s = 0.f, t = 0.f;
        for (int i = 0; i < n; ++i) m = fmaxf(m, raw[i]);
        for (int i = 0; i < n; ++i) s += __expf(raw[i] - m);
        for (int i = 0; i < n; ++i) {
            w[i] = __expf(raw[i] - m) / s;
            t += w[i] * dots[i];
        }
        for (int i = 0; i < n; ++i) draw[i] += w[i] * (dots[i] - t);
    }
}

// ---- pooling / striding / elementwise ---------------------------------------------------------------------
// backward of nn.AdaptiveAvgPool2d: dX[b, y, x, c] = sum over the windows (oy, ox) that contain (y, x) of dY / window size
// (gather form: deterministic, no atomics; a pixel lies in at most two windows per axis when OH <= H)
__global__ void __launch_bounds__(256)
adaptive_avgpool_bwd_kernel(const __nv_bfloat16 *__restrict__ dY, int H, int W, int C, int OH, int OW, long total8,
                            __nv_bfloat16 *__restrict__ dX) {
    const int C8 = C >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long p = i / C8;
        const int x = (int)(p % W);
        p /= W;
        const int y = (int)(p % H);
        const long b = p / H;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const int oy_c = (y * OH) / H, ox_c = (x * OW) / W;
        for (int oy = max(0, oy_c - 1); oy <= min(OH - 1, oy_c + 1); ++oy) {
            const int y0 = (oy * H) / OH, y1 = ((oy + 1) * H + OH - 1) / OH;
            if (y < y0 || y >= y1) continue;
            for (int ox = max(0, ox_c - 1); ox <= min(OW - 1, ox_c + 1); ++ox) {
                const int x0 = (ox * W) / OW, x1 = ((ox + 1) * W + OW - 1) / OW;
                if (x < x0 || x >= x1) continue;
                float f[8];
                up8(__ldg((const uint4 *)(dY + ((b * OH + oy) * OW + ox) * C + cg * 8)), f);
                const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], inv, acc[j]);
            }
        }
        ((uint4 *)dX)[i] = pk8(acc);
    }
}

// dX[b, 2y+dy, 2x+dx, c] = 0.25 * dY[b, y, x, c]
__global__ void __launch_bounds__(256)
avgpool2x2_bwd_kernel(const __nv_bfloat16 *__restrict__ dY, int H, int W, int C, long total8, __nv_bfloat16 *__restrict__ dX) {
    const int C8 = C >> 3, Ho = H >> 1, Wo = W >> 1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long p = i / C8;
        const int x = (int)(p % W);
        p /= W;
        const int y = (int)(p % H);
        const long b = p / H;
        float f[8];
        up8(__ldg((const uint4 *)(dY + ((b * Ho + (y >> 1)) * Wo + (x >> 1)) * C + cg * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] *= 0.25f;
        ((uint4 *)dX)[i] = pk8(f);
    }
}

// dX[b, s*i, s*j, c] += dXs[b, i, j, c]   (data gradient of a strided 1x1 convolution, added into the full-size map)
__global__ void __launch_bounds__(256)
scatter_strided_add_kernel(const __nv_bfloat16 *__restrict__ dXs, int Ho, int Wo, int H, int W, int C, int stride, long total8,
                           __nv_bfloat16 *__restrict__ dX) {
    const int C8 = C >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long p = i / C8;
        const int x = (int)(p % Wo);
        p /= Wo;
        const int y = (int)(p % Ho);
        const long b = p / Ho;
        float a[8], d[8];
        uint4 *dst = (uint4 *)(dX + ((b * H + (long)y * stride) * W + (long)x * stride) * C + cg * 8);
        up8(*dst, a);
        up8(__ldg((const uint4 *)dXs + i), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] += d[j];
        *dst = pk8(a);
    }
}

__global__ void __launch_bounds__(256)
add_kernel(const __nv_bfloat16 *__restrict__ a, const __nv_bfloat16 *__restrict__ b, long n8, __nv_bfloat16 *__restrict__ out) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long)gridDim.x * blockDim.x) {
        float x[8], y[8];
        up8(__ldg((const uint4 *)a + i), x);
        up8(__ldg((const uint4 *)b + i), y);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += y[j];
        ((uint4 *)out)[i] = pk8(x);
    }
}

// counter-based dropout mask (common.cuh: drop_keep), 8 elements per thread; the same mask is regenerated in backward
__global__ void __launch_bounds__(256)
dropout_kernel(const __nv_bfloat16 *__restrict__ x, long n, uint32_t thresh, float keep_scale, DropSeed seed,
               __nv_bfloat16 *__restrict__ out) {
    const long n8 = n >> 3;
    const bool vec = (((uintptr_t)x | (uintptr_t)out) & 15) == 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; vec && i < n8; i += (long)gridDim.x * blockDim.x) {
        float v[8];
        up8(__ldg((const uint4 *)x + i), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = drop_keep(seed, (uint64_t)(i * 8 + j), thresh) ? v[j] * keep_scale : 0.f;
        ((uint4 *)out)[i] = pk8(v);
    }
    for (long i = (vec ? n8 * 8 : 0) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        out[i] = __float2bfloat16_rn(drop_keep(seed, (uint64_t)i, thresh) ? __bfloat162float(x[i]) * keep_scale : 0.f);
}

// ---- table-driven parameter re-layout (one launch per step) --------------------------------------------------
// kind 0: depthwise weight [C,1,3,3] fp32 -> fp32 [9, C]                                     (d0 = C)
// kind 1: dense conv weight [Co,Ci,K,K] fp32 -> bf16 KRSC [Co, K, K, Cp]                       (d0 Co, d1 Ci, d2 K, d3 Cp)
// kind 2: dense conv weight -> bf16 [Ci, K, K, Co] spatially flipped (the data-gradient conv)  (same dims, Cp ignored)
// kind 3: matrix [R, Cc] fp32 -> bf16 [.., ld d2] top-left block (zero padding is never touched) (d0 R, d1 Cc, d2 ld)
// kind 4: fp32 vector copy (d0 = n)
// kind 5: (gradient, reverse of 1) fp32 KRSC [Co,K,K,Cp] workspace -> += into fp32 [Co,Ci,K,K]   (src = workspace)
// kind 6: depthwise weight [C,1,3,3] fp32 -> fp32 [9, C] with the taps flipped (data gradient = forward kernel over dY)
struct RepackEntry {
    long src, dst;
    int kind, d0, d1, d2, d3, pad;
};
__global__ void __launch_bounds__(256)
repack_kernel(const RepackEntry *__restrict__ table, const float *__restrict__ src_f32, float *__restrict__ dst_f32,
              __nv_bfloat16 *__restrict__ dst_bf16) {
    const RepackEntry e = table[blockIdx.y];
    const float *s = src_f32 + e.src;
    long n;
    switch (e.kind) {
        case 0: case 6: n = (long)e.d0 * 9; break;
        case 1: case 2: case 5: n = (long)e.d0 * e.d1 * e.d2 * e.d2; break;
        case 3: n = (long)e.d0 * e.d1; break;
        default: n = e.d0; break;
    }
    // all index arithmetic in 32 bits (a parameter tensor has far fewer than 2^31 elements; the host checks): the 64-bit
    // divisions and remainders that decoded every element cost ~150 instructions each and made this the slowest
    // "copy" of the step (182 us for 43 tensors, 24 us per weight-gradient unpack)
    if (e.kind == 1 || e.kind == 2) {
        // destination-major: coalesced bf16 stores, strided fp32 gathers (L2 hits) -- the source-major loop scattered 2-byte
        // stores over 32 sectors per warp
        const unsigned K = e.d2, Ci = e.d1, Co = e.d0;
        const unsigned inner = e.kind == 1 ? e.d3 : Co;          // fastest destination index: padded ci (kind 1) / co (kind 2)
        const unsigned nd = e.kind == 1 ? Co * K * K * e.d3 : Ci * K * K * Co;
        for (unsigned j = blockIdx.x * 256u + threadIdx.x; j < nd; j += gridDim.x * 256u) {
            unsigned t = j;
            const unsigned in = t % inner; t /= inner;
            const unsigned kw = t % K; t /= K;
            const unsigned kh = t % K;
            const unsigned outer = t / K;
            if (e.kind == 1) {
                if (in < Ci) dst_bf16[e.dst + j] = __float2bfloat16_rn(s[((outer * Ci + in) * K + kh) * K + kw]);
            } else {
                dst_bf16[e.dst + j] = __float2bfloat16_rn(s[((in * Ci + outer) * K + (K - 1 - kh)) * K + (K - 1 - kw)]);
            }
        }
        return;
    }
    if (e.kind == 5) {
        // destination-major over the [Co,Ci,K,K] gradient (coalesced read-modify-write), gathers from the KRSC staging
        const unsigned K = e.d2, Ci = e.d1, KK = K * K, Cp = e.d3;
        for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < (unsigned)n; i += gridDim.x * 256u) {
            const unsigned tap = i % KK, t = i / KK;
            const unsigned ci = t % Ci, co = t / Ci;
            dst_f32[e.dst + i] += s[(co * KK + tap) * Cp + ci];
        }
        return;
    }
    for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < (unsigned)n; i += gridDim.x * 256u) {
        if (e.kind == 0 || e.kind == 6) {
            const unsigned c = i / 9, k = i - c * 9;
            dst_f32[e.dst + (e.kind == 0 ? k : 8 - k) * e.d0 + c] = s[i];
        } else if (e.kind == 3) {
            const unsigned r = i / e.d1, c = i - r * e.d1;
            dst_bf16[e.dst + (long)r * e.d2 + c] = __float2bfloat16_rn(s[i]);
        } else {
            dst_f32[e.dst + i] = s[i];
        }
    }
}

int launch_bn_finalize_parts(const float *partials, int parts, long count, const float *gamma, const float *beta, float eps,
                             float momentum, int C, float *mean_rstd, float *scale_shift, float *running_mean,
                             float *running_var, cudaStream_t stream) {
    bn_finalize_kernel<<<(C + 7) / 8, 256, 0, stream>>>(partials, parts, (float)count, gamma, beta, eps, momentum, C, mean_rstd,
                                                         scale_shift, running_mean, running_var);
    return launch_status();
}

}  // namespace pose

using namespace pose;

#define REQ(c, e) do { if (!(c)) return (e); } while (0)

// number of per-block partial slots a row reduction over [M, C] uses, given a scratch capacity in floats: one wave of
// RESIDENT CTAs (per_sm from the occupancy calculator) -- with a fixed 4 per SM the 80-register backward reduction (3
// resident CTAs per SM) ran a second, one-third-full wave
static int bn_parts(long M, int C, long cap_floats, int per_sm) {
    long parts = rowmap_grid(M, C, 8);
    if (parts > (long)kNumSMs * per_sm) parts = (long)kNumSMs * per_sm;
    const long fit = cap_floats / (2L * C);
    if (parts > fit) parts = fit;
    return (int)(parts < 1 ? 1 : parts);
}
template <class K>
static int resident_ctas(K kern, int threads, size_t smem = 0) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem) != cudaSuccess || n < 1) n = 1;
    return n > 8 ? 8 : n;
}
static int one_wave(int grid, int per_sm) { return grid < kNumSMs * per_sm ? grid : kNumSMs * per_sm; }   // grid-stride kernels
static int bn_stats_parts(long M, int C, long cap_floats) {
    static const bool once = (rowpipe_optin(bn_stats_kernel), true);
    (void)once;
    const int thr = rowmap_threads(C);
    return bn_parts(M, C, cap_floats, resident_ctas(bn_stats_kernel, thr, rowpipe_bytes(1, thr)));
}

POSE_API int pose_bn_stats_bf16(const void *Y, long M, int C, long ld, float *partials, long cap_floats, pose_stream_t stream) {
    REQ(Y && partials, POSE_E_NULL);
    REQ(M > 0 && C > 0 && C % 8 == 0 && ld >= C && ld % 8 == 0, POSE_E_SHAPE);
    REQ((uintptr_t)Y % 16 == 0, POSE_E_ALIGN);
    REQ(C <= 3072, POSE_E_UNSUPPORTED);
    REQ(cap_floats >= 2L * C, POSE_E_WORKSPACE);
    const int thr = rowmap_threads(C);
    bn_stats_kernel<<<bn_stats_parts(M, C, cap_floats), thr, rowpipe_bytes(1, thr), (cudaStream_t)stream>>>((const __nv_bfloat16 *)Y, M, C,
                                                                                                      ld, partials);
    return launch_status();
}

POSE_API int pose_bn_finalize(const float *partials, long cap_floats, long count, const float *gamma, const float *beta,
                              float eps, float momentum, int C, float *mean_rstd, float *scale_shift, float *running_mean,
                              float *running_var, pose_stream_t stream) {
    REQ(partials && gamma && beta && mean_rstd && scale_shift, POSE_E_NULL);
    REQ(count > 0 && C > 0, POSE_E_SHAPE);
    bn_finalize_kernel<<<(C + 7) / 8, 256, 0, (cudaStream_t)stream>>>(partials, bn_stats_parts(count, C, cap_floats), (float)count,
                                                                         gamma, beta, eps, momentum, C, mean_rstd, scale_shift,
                                                                         running_mean, running_var);
    return launch_status();
}

/* the same fold over an explicit number of partial rows [parts, 2, C] (statistics emitted by the producing kernel:
   pose_dwconv3x3_bn_stats_bf16) */
POSE_API int pose_bn_finalize_parts(const float *partials, int parts, long count, const float *gamma, const float *beta, float eps,
                                    float momentum, int C, float *mean_rstd, float *scale_shift, float *running_mean,
                                    float *running_var, pose_stream_t stream) {
    REQ(partials && gamma && beta && mean_rstd && scale_shift, POSE_E_NULL);
    REQ(count > 0 && C > 0 && parts > 0, POSE_E_SHAPE);
    bn_finalize_kernel<<<(C + 7) / 8, 256, 0, (cudaStream_t)stream>>>(partials, parts, (float)count, gamma, beta, eps, momentum, C,
                                                                         mean_rstd, scale_shift, running_mean, running_var);
    return launch_status();
}

POSE_API int pose_bn_apply_bf16(const void *Y, long M, int C, const float *scale_shift, int act, float out_scale,
                                const void *residual, long ld_res, void *out, long ld_out, pose_stream_t stream) {
    REQ(Y && scale_shift && out, POSE_E_NULL);
    REQ(M > 0 && C > 0 && C % 8 == 0 && ld_out >= C && ld_out % 8 == 0 && (!residual || (ld_res >= C && ld_res % 8 == 0)),
        POSE_E_SHAPE);
    REQ((uintptr_t)Y % 16 == 0 && (uintptr_t)out % 16 == 0 && (uintptr_t)residual % 16 == 0, POSE_E_ALIGN);
    REQ(C <= 3072 && act >= 0 && act <= 2, POSE_E_UNSUPPORTED);
    const int thr = rowmap_threads(C);
    cudaStream_t s = (cudaStream_t)stream;
#define BN_APPLY(A_, R_)                                                                                               \
    {                                                                                                                  \
        static const bool once = (rowpipe_optin(bn_apply_kernel<A_, R_>), true);                                       \
        (void)once;                                                                                                    \
        const size_t sm = R_ ? rowpipe_bytes(2, thr) : 0;                                                              \
        bn_apply_kernel<A_, R_><<<one_wave(rowmap_grid(M, C, 4), resident_ctas(bn_apply_kernel<A_, R_>, thr, sm)), thr, sm, s>>>(     \
            (const __nv_bfloat16 *)Y, M, C, scale_shift, out_scale, (const __nv_bfloat16 *)residual, ld_res,           \
            (__nv_bfloat16 *)out, ld_out);                                                                             \
    }
#define BN_APPLY_R(A_) if (residual) BN_APPLY(A_, true) else BN_APPLY(A_, false)
    if (act == 0) BN_APPLY_R(0) else if (act == 1) BN_APPLY_R(1) else BN_APPLY_R(2)
#undef BN_APPLY_R
#undef BN_APPLY
    return launch_status();
}

POSE_API int pose_bn_apply_pool_bf16(const void *Y, int B, long HW, int C, const float *scale_shift, int act, void *out,
                                     float *pool, int parts, pose_stream_t stream) {
    REQ(Y && scale_shift && out && pool, POSE_E_NULL);
    REQ(B > 0 && HW > 0 && C > 0 && C % 8 == 0 && parts > 0, POSE_E_SHAPE);
    REQ((uintptr_t)Y % 16 == 0 && (uintptr_t)out % 16 == 0, POSE_E_ALIGN);
    REQ(C <= 3072 && act >= 0 && act <= 2, POSE_E_UNSUPPORTED);
    const int thr = rowmap_threads(C);
    const dim3 grid((unsigned)parts, (unsigned)B);
    cudaStream_t s = (cudaStream_t)stream;
    if (act == 0) bn_apply_pool_kernel<0><<<grid, thr, 0, s>>>((const __nv_bfloat16 *)Y, HW, C, scale_shift, (__nv_bfloat16 *)out, pool);
    else if (act == 1) bn_apply_pool_kernel<1><<<grid, thr, 0, s>>>((const __nv_bfloat16 *)Y, HW, C, scale_shift, (__nv_bfloat16 *)out, pool);
    else bn_apply_pool_kernel<2><<<grid, thr, 0, s>>>((const __nv_bfloat16 *)Y, HW, C, scale_shift, (__nv_bfloat16 *)out, pool);
    return launch_status();
}

POSE_API int pose_bn_bwd_bf16(const void *dA, long ld_da, const void *Y, long M, int C, const float *scale_shift,
                              const float *mean_rstd, int act, float out_scale, float *partials, long cap_floats, float *coef,
                              void *dY, float *dgamma, float *dbeta, pose_stream_t stream) {
    REQ(dA && Y && scale_shift && mean_rstd && partials && coef && dY && dgamma && dbeta, POSE_E_NULL);
    REQ(M > 0 && C > 0 && C % 8 == 0 && ld_da >= C && ld_da % 8 == 0, POSE_E_SHAPE);
    REQ((uintptr_t)dA % 16 == 0 && (uintptr_t)Y % 16 == 0 && (uintptr_t)dY % 16 == 0, POSE_E_ALIGN);
    REQ(C <= 3072 && act >= 0 && act <= 2, POSE_E_UNSUPPORTED);
    REQ(cap_floats >= 2L * C, POSE_E_WORKSPACE);
    cudaStream_t s = (cudaStream_t)stream;
    const int thr = rowmap_threads(C);
#define BN_BWD(A_)                                                                                                     \
    static const bool once = (rowpipe_optin(bn_bwd_reduce_kernel<A_>), rowpipe_optin(bn_bwd_apply_kernel<A_>), true);  \
    (void)once;                                                                                                        \
    const size_t sm = rowpipe_bytes(2, thr);                                                                           \
    const int parts = bn_parts(M, C, cap_floats, resident_ctas(bn_bwd_reduce_kernel<A_>, thr, sm));                    \
    bn_bwd_reduce_kernel<A_><<<parts, thr, sm, s>>>((const __nv_bfloat16 *)dA, ld_da, (const __nv_bfloat16 *)Y, M, C,    \
                                                   scale_shift, mean_rstd, out_scale, partials);                       \
    bn_bwd_coef_kernel<<<(C + 7) / 8, 256, 0, s>>>(partials, parts, 1.0f / (float)M, scale_shift, mean_rstd, C, coef,     \
                                                    dgamma, dbeta);                                                    \
    bn_bwd_apply_kernel<A_><<<one_wave(rowmap_grid(M, C, 4), resident_ctas(bn_bwd_apply_kernel<A_>, thr, sm)), thr, sm, s>>>(      \
        (const __nv_bfloat16 *)dA, ld_da, (const __nv_bfloat16 *)Y, M, C, scale_shift, coef, out_scale, (__nv_bfloat16 *)dY)
    if (act == 0) { BN_BWD(0); } else if (act == 1) { BN_BWD(1); } else { BN_BWD(2); }
#undef BN_BWD
    return launch_status();
}

POSE_API int pose_dwconv3x3_bwd_bf16(const void *dY, const void *X, const float *Wd, int B, int H, int W, int C, int stride,
                                     const void *add, void *dX, float *dW, pose_stream_t stream) {
    REQ(dY && Wd && (dX || dW) && (!dW || X), POSE_E_NULL);
    REQ(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2), POSE_E_SHAPE);
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    cudaStream_t s = (cudaStream_t)stream;
    if (dX) {
        const long total8 = (long)B * H * W * (C / 8);
        if (stride == 2 && (long)B * H < 2147483647L && (uintptr_t)Wd % 16 == 0)
            dwconv_bwd_data_s2_kernel<<<(unsigned)((long)B * H), 256, 0, s>>>((const __nv_bfloat16 *)dY, Wd, H, W, C, Ho, Wo,
                                                                            (const __nv_bfloat16 *)add, (__nv_bfloat16 *)dX);
        else
            dwconv_bwd_data_kernel<<<grid_for(total8), 256, 0, s>>>((const __nv_bfloat16 *)dY, Wd, H, W, C, Ho, Wo, stride,
                                                                   (const __nv_bfloat16 *)add, total8, (__nv_bfloat16 *)dX);
    }
    if (dW) {
        const int th = stride == 1 ? DwT<1>::TH : DwT<2>::TH, tw = stride == 1 ? DwT<1>::TW : DwT<2>::TW;
        const int tiles_y = (Ho + th - 1) / th, tiles_x = (Wo + tw - 1) / tw;
        const int gy = (C + 63) / 64;
        long gx = (long)B * tiles_x * tiles_y;
        const long cap = (kNumSMs * 3 + gy - 1) / gy;
        if (gx > cap) gx = cap;
        const int smem1 = 2 * DwT<1>::IH * DwT<1>::IW * 128;
        const int smem2 = 2 * DwT<2>::IH * DwT<2>::IW * 128;
        CUtensorMap mapX;
        {
            const int e = make_map_dw_halo(&mapX, X, B, H, W, C, stride == 1 ? DwT<1>::IW : DwT<2>::IW,
                                           stride == 1 ? DwT<1>::IH : DwT<2>::IH);
            if (e) return e;
        }
        static bool cfg = false;
        if (!cfg) {
            cudaError_t ce = cudaFuncSetAttribute(dwconv_bwd_weight_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
            if (ce == cudaSuccess)
                ce = cudaFuncSetAttribute(dwconv_bwd_weight_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
            if (ce != cudaSuccess) return (int)ce;
            cfg = true;
        }
        if (stride == 1)
            dwconv_bwd_weight_kernel<1><<<dim3((unsigned)gx, gy), 256, smem1, s>>>((const __nv_bfloat16 *)dY, mapX,
                                                                                  B, H, W, C, Ho, Wo, tiles_x, tiles_y, dW);
        else
            dwconv_bwd_weight_kernel<2><<<dim3((unsigned)gx, gy), 256, smem2, s>>>((const __nv_bfloat16 *)dY, mapX,
                                                                                  B, H, W, C, Ho, Wo, tiles_x, tiles_y, dW);
    }
    return launch_status();
}

POSE_API int pose_gate_bwd_reduce_bf16(const void *dOut, const void *X, int B, long HW, int C, float *dgate, pose_stream_t stream) {
    REQ(dOut && X && dgate, POSE_E_NULL);
    REQ(B > 0 && HW > 0 && C > 0 && C % 8 == 0, POSE_E_SHAPE);
    REQ(C <= 3072, POSE_E_UNSUPPORTED);
    const int rpb = rowmap_threads(C) / (C / 8);
    long gx = (HW + (long)rpb * 8 - 1) / ((long)rpb * 8);
    const long cap = (kNumSMs * 8 + B - 1) / B;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    static const bool once = (rowpipe_optin(gate_bwd_reduce_kernel), true);
    (void)once;
    const int thr = rowmap_threads(C);
    gate_bwd_reduce_kernel<<<dim3((unsigned)gx, B), thr, rowpipe_bytes(2, thr), (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)dOut, (const __nv_bfloat16 *)X, HW, C, dgate);
    return launch_status();
}

POSE_API int pose_gate_bwd_apply_bf16(const void *dOut, const float *gate, const void *dmean, float inv_hw, int B, long HW, int C,
                                      const void *add, void *dX, pose_stream_t stream) {
    REQ((dOut || dmean) && dX, POSE_E_NULL);      /* dOut NULL: dX = dmean / HW broadcast (gradient of a global average) */
    REQ(B > 0 && HW > 0 && C > 0 && C % 8 == 0, POSE_E_SHAPE);
    REQ(C <= 3072, POSE_E_UNSUPPORTED);
    // enough row blocks per image for ~8 CTAs per SM over the batch, at least 4 rows per thread
    const int rpb = rowmap_threads(C) / (C / 8);
    long gx = (HW + (long)rpb * 4 - 1) / ((long)rpb * 4);
    const long cap = ((long)kNumSMs * 8 + B - 1) / B;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    gate_bwd_apply_kernel<<<dim3((unsigned)gx, B), rowmap_threads(C), 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16 *)dOut, gate, (const __nv_bfloat16 *)dmean, inv_hw, HW, C, (const __nv_bfloat16 *)add,
        (__nv_bfloat16 *)dX);
    return launch_status();
}

POSE_API int pose_gate_bwd_apply_bn_bf16(const void *dOut, const float *gate, const void *dmean, float inv_hw, int B, long HW, int C,
                                         const void *Yprev, const float *scale_shift_prev, int act_prev, void *dZ, float *partials,
                                         long cap_floats, int *parts_out, pose_stream_t stream) {
    REQ(dOut && gate && dmean && Yprev && scale_shift_prev && dZ && partials && parts_out, POSE_E_NULL);
    REQ(B > 0 && HW > 0 && C > 0 && C % 8 == 0, POSE_E_SHAPE);
    REQ(C <= 3072 && (act_prev == 1 || act_prev == 2), POSE_E_UNSUPPORTED);
    REQ((uintptr_t)dOut % 16 == 0 && (uintptr_t)Yprev % 16 == 0 && (uintptr_t)dZ % 16 == 0 && (uintptr_t)dmean % 16 == 0, POSE_E_ALIGN);
    const int thr = rowmap_threads(C), rpb = thr / (C / 8);
    const size_t sm = rowpipe_bytes(2, thr);
    static const bool once = (rowpipe_optin(gate_bwd_apply_bn_kernel<1>), rowpipe_optin(gate_bwd_apply_bn_kernel<2>), true);
    (void)once;
    const int per_sm = act_prev == 1 ? resident_ctas(gate_bwd_apply_bn_kernel<1>, thr, sm) : resident_ctas(gate_bwd_apply_bn_kernel<2>, thr, sm);
    long gx = (HW + (long)rpb * 4 - 1) / ((long)rpb * 4);
    const long cap = ((long)kNumSMs * per_sm + B - 1) / B;       // one wave of resident CTAs over the batch
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    REQ(gx * B * 2L * C <= cap_floats, POSE_E_WORKSPACE);
    *parts_out = (int)(gx * B);
    cudaStream_t s = (cudaStream_t)stream;
    if (act_prev == 1)
        gate_bwd_apply_bn_kernel<1><<<dim3((unsigned)gx, B), thr, sm, s>>>((const __nv_bfloat16 *)dOut, gate, (const __nv_bfloat16 *)dmean, inv_hw,
                                                                         HW, C, (const __nv_bfloat16 *)Yprev, scale_shift_prev,
                                                                         (__nv_bfloat16 *)dZ, partials);
    else
        gate_bwd_apply_bn_kernel<2><<<dim3((unsigned)gx, B), thr, sm, s>>>((const __nv_bfloat16 *)dOut, gate, (const __nv_bfloat16 *)dmean, inv_hw,
                                                                         HW, C, (const __nv_bfloat16 *)Yprev, scale_shift_prev,
                                                                         (__nv_bfloat16 *)dZ, partials);
    return launch_status();
}

/* second half of the BatchNorm backward when the producer of the incoming gradient already emitted dz = dA * act'(z) and the
   [parts, 2, C] partial sums of dz and dz * y (pose_gate_bwd_apply_bn_bf16, pose_dwconv3x3_bnbwd_bf16): fold -> dgamma, dbeta,
   coefficients -> dY = a dz - k2 y - k3 */
POSE_API int pose_bn_bwd_from_dz_bf16(const void *dZ, long ld_dz, const void *Y, long M, int C, const float *scale_shift,
                                      const float *mean_rstd, const float *partials, int parts, float *coef, void *dY,
                                      float *dgamma, float *dbeta, pose_stream_t stream) {
    REQ(dZ && Y && scale_shift && mean_rstd && partials && coef && dY && dgamma && dbeta, POSE_E_NULL);
    REQ(M > 0 && C > 0 && C % 8 == 0 && ld_dz >= C && ld_dz % 8 == 0 && parts > 0, POSE_E_SHAPE);
    REQ((uintptr_t)dZ % 16 == 0 && (uintptr_t)Y % 16 == 0 && (uintptr_t)dY % 16 == 0, POSE_E_ALIGN);
    REQ(C <= 3072, POSE_E_UNSUPPORTED);
    cudaStream_t s = (cudaStream_t)stream;
    const int thr = rowmap_threads(C);
    static const bool once = (rowpipe_optin(bn_bwd_apply_kernel<0>), true);
    (void)once;
    const size_t sm = rowpipe_bytes(2, thr);
    bn_bwd_coef_kernel<<<(C + 7) / 8, 256, 0, s>>>(partials, parts, 1.0f / (float)M, scale_shift, mean_rstd, C, coef, dgamma, dbeta);
    bn_bwd_apply_kernel<0><<<one_wave(rowmap_grid(M, C, 4), resident_ctas(bn_bwd_apply_kernel<0>, thr, sm)), thr, sm, s>>>(
        (const __nv_bfloat16 *)dZ, ld_dz, (const __nv_bfloat16 *)Y, M, C, scale_shift, coef, 1.0f, (__nv_bfloat16 *)dY);
    return launch_status();
}

POSE_API int pose_sigmoid_bwd(const float *dgate, const float *gate, long n, void *dz, pose_stream_t stream) {
    REQ(dgate && gate && dz, POSE_E_NULL);
    REQ(n > 0, POSE_E_SHAPE);
    sigmoid_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(dgate, gate, n, (__nv_bfloat16 *)dz);
    return launch_status();
}

POSE_API int pose_eca_bwd(const float *dgate, const void *dfeat, const float *gate, const float *pool_sum, int parts, float inv_hw,
                          const float *w, int k, int B, int C, int mode, void *dmean, float *dw, pose_stream_t stream) {
    REQ(gate && pool_sum && w && dmean && dw && (mode == 0 ? (const void *)dgate : dfeat), POSE_E_NULL);
    REQ(B > 0 && C > 0 && k > 0 && (k & 1) && parts >= 1, POSE_E_SHAPE);
    eca_bwd_kernel<<<B, 256, (size_t)2 * C * sizeof(float), (cudaStream_t)stream>>>(dgate, (const __nv_bfloat16 *)dfeat, gate, pool_sum,
                                                                                   parts, inv_hw, w, k, C, mode,
                                                                                   (__nv_bfloat16 *)dmean, dw);
    return launch_status();
}

POSE_API int pose_coord_bwd_reduce_bf16(const void *dOut, const void *X, const void *G, int B, int H, int W, int C, void *dZ,
                                        pose_stream_t stream) {
    REQ(dOut && X && G && dZ, POSE_E_NULL);
    REQ(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, POSE_E_SHAPE);
    const int items = (H + W) * (C / 8);
    int chunks = (items + 255) / 256;
    if (chunks > 16) chunks = 16;
    coord_bwd_reduce_kernel<<<dim3(B, chunks), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)dOut, (const __nv_bfloat16 *)X,
                                                                (const __nv_bfloat16 *)G, H, W, C, (__nv_bfloat16 *)dZ);
    return launch_status();
}

POSE_API int pose_coord_bwd_apply_bf16(const void *dOut, const void *G, const void *dP, int B, int H, int W, int C, void *dX,
                                       pose_stream_t stream) {
    REQ(dOut && G && dP && dX, POSE_E_NULL);
    REQ(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, POSE_E_SHAPE);
    const long total8 = (long)B * H * W * (C / 8);
    coord_bwd_apply_kernel<<<grid_for(total8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)dOut, (const __nv_bfloat16 *)G,
                                                                              (const __nv_bfloat16 *)dP, H, W, C, total8,
                                                                              (__nv_bfloat16 *)dX);
    return launch_status();
}

POSE_API int pose_wasp_mix_bf16(const void *branches, int nb, const void *glob, const float *raw_weights, int B, long HW, int C,
                                void *out, pose_stream_t stream) {
    REQ(branches && glob && raw_weights && out, POSE_E_NULL);
    REQ(nb > 0 && nb < 8 && B > 0 && HW > 0 && C > 0 && C % 8 == 0, POSE_E_SHAPE);
    const long total8 = (long)B * HW * (C / 8);
    wasp_mix_kernel<<<grid_for(total8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)branches, nb, (long)B * HW * C,
                                                                       (const __nv_bfloat16 *)glob, raw_weights, HW, C, total8,
                                                                       (__nv_bfloat16 *)out);
    return launch_status();
}

POSE_API int pose_wasp_mix_bwd_bf16(const void *dOut, const void *branches, int nb, const void *glob, const float *raw_weights,
                                    int B, long HW, int C, void *dbranches, float *dglob, float *dots, float *draw,
                                    pose_stream_t stream) {
    REQ(dOut && branches && glob && raw_weights && dbranches && dglob && dots && draw, POSE_E_NULL);
    REQ(nb > 0 && nb < 8 && B > 0 && HW > 0 && C > 0 && C % 8 == 0, POSE_E_SHAPE);
    REQ(C <= 3072, POSE_E_UNSUPPORTED);
    cudaStream_t s = (cudaStream_t)stream;
    const int thr = rowmap_threads(C), rpb = thr / (C / 8);
    long gx = (HW + (long)rpb * 4 - 1) / ((long)rpb * 4);
    const long cap = ((long)kNumSMs * 4 + B - 1) / B;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    wasp_mix_bwd_kernel<<<dim3((unsigned)gx, B), thr, 0, s>>>((const __nv_bfloat16 *)dOut, (const __nv_bfloat16 *)branches, nb,
                                                             (long)B * HW * C, (const __nv_bfloat16 *)glob, raw_weights, HW, C,
                                                             (__nv_bfloat16 *)dbranches, dglob, dots);
    wasp_weights_bwd_kernel<<<1, 32, 0, s>>>(raw_weights, dots, nb + 1, draw);
    return launch_status();
}

POSE_API int pose_adaptive_avgpool_bwd_bf16(const void *dY, int B, int H, int W, int C, int OH, int OW, void *dX,
                                            pose_stream_t stream) {
    REQ(dY && dX, POSE_E_NULL);
    REQ(B > 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && OH <= H && OW <= W && C > 0 && C % 8 == 0, POSE_E_SHAPE);
    REQ((uintptr_t)dY % 16 == 0 && (uintptr_t)dX % 16 == 0, POSE_E_ALIGN);
    const long total8 = (long)B * H * W * (C / 8);
    adaptive_avgpool_bwd_kernel<<<grid_for(total8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)dY, H, W, C, OH, OW, total8,
                                                                                   (__nv_bfloat16 *)dX);
    return launch_status();
}

POSE_API int pose_avgpool2x2_bwd_bf16(const void *dY, int B, int H, int W, int C, void *dX, pose_stream_t stream) {
    REQ(dY && dX, POSE_E_NULL);
    REQ(B > 0 && H > 0 && W > 0 && !(H & 1) && !(W & 1) && C > 0 && C % 8 == 0, POSE_E_SHAPE);
    const long total8 = (long)B * H * W * (C / 8);
    avgpool2x2_bwd_kernel<<<grid_for(total8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)dY, H, W, C, total8,
                                                                             (__nv_bfloat16 *)dX);
    return launch_status();
}

POSE_API int pose_scatter_strided_add_bf16(const void *dXs, int B, int Ho, int Wo, int H, int W, int C, int stride, void *dX,
                                           pose_stream_t stream) {
    REQ(dXs && dX, POSE_E_NULL);
    REQ(B > 0 && Ho > 0 && Wo > 0 && C > 0 && C % 8 == 0 && stride >= 1 && (Ho - 1) * stride < H && (Wo - 1) * stride < W,
        POSE_E_SHAPE);
    const long total8 = (long)B * Ho * Wo * (C / 8);
    scatter_strided_add_kernel<<<grid_for(total8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)dXs, Ho, Wo, H, W, C, stride,
                                                                                  total8, (__nv_bfloat16 *)dX);
    return launch_status();
}

POSE_API int pose_add_bf16(const void *a, const void *b, long n, void *out, pose_stream_t stream) {
    REQ(a && b && out, POSE_E_NULL);
    REQ(n > 0 && n % 8 == 0, POSE_E_SHAPE);
    add_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)a, (const __nv_bfloat16 *)b, n / 8,
                                                                 (__nv_bfloat16 *)out);
    return launch_status();
}

POSE_API int pose_dropout_bf16(const void *x, long n, float p, uint64_t seed, void *out, pose_stream_t stream) {
    REQ(x && out, POSE_E_NULL);
    REQ(n > 0 && p >= 0.f && p < 1.f, POSE_E_SHAPE);
    const uint32_t thresh = drop_threshold(p);
    dropout_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)x, n, thresh, 1.0f / (1.0f - p), make_drop_seed(seed),
                                                                 (__nv_bfloat16 *)out);
    return launch_status();
}

POSE_API int pose_param_repack(const void *table, int n_entries, const float *src_f32, float *dst_f32, void *dst_bf16,
                               pose_stream_t stream) {
    REQ(table && src_f32, POSE_E_NULL);
    REQ(n_entries > 0, POSE_E_SHAPE);
    repack_kernel<<<dim3(n_entries >= 8 ? 64 : 296, n_entries), 256, 0, (cudaStream_t)stream>>>((const RepackEntry *)table, src_f32, dst_f32,
                                                                        (__nv_bfloat16 *)dst_bf16);
    return launch_status();
}
