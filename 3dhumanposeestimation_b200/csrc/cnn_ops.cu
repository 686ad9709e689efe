// cnn_ops.cu -- the bandwidth-bound pieces of CNNPoseEstimation.forward (src/models/cnn.py) in channels-last
// bf16: fused input assembly (heat-map render + concat + cast), depthwise 3x3 + BN + SiLU (+ squeeze sums),
// SE / ECA / CoordAttention gates, pooling.  The dense convolutions and Linear layers run on the tcgen05
// GEMM (gemm_tcgen05.cu); everything here moves each activation through HBM exactly once per op.
#include "tc_common.cuh"

namespace pose {

__device__ __forceinline__ float act_f(float v, int act) {   // run-time activation (small kernels only)
    switch (act) {
        case 1: return v > 0.f ? v : 0.f;
        case 2: return v / (1.0f + __expf(-v));
        case 3: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
        case 4: return 1.0f / (1.0f + __expf(-v));
        default: return v;
    }
}

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// compile-time activation for the bandwidth kernels: silu / sigmoid through one MUFU.TANH
template <int ACT>
__device__ __forceinline__ float act_t(float v) {
    if (ACT == 1) return fmaxf(v, 0.f);
    if (ACT == 2) return 0.5f * v * (1.0f + tanh_approx(0.5f * v));
    if (ACT == 3) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    if (ACT == 4) return 0.5f * (1.0f + tanh_approx(0.5f * v));
    return v;
}

__device__ __forceinline__ void unpack8(const uint4 &p, float (&f)[8]) {
    const __nv_bfloat162 *h = (const __nv_bfloat162 *)&p;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 t = __bfloat1622float2(h[q]);
        f[2 * q] = t.x;
        f[2 * q + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    __nv_bfloat162 h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
    return *(uint4 *)h;
}

// ---------------------------------------------------------------------------------------------------------
// Input assembly: torch.cat([image, depth, heatmaps], 1) (cnn.py:644-648) written directly as the conv1
// operand: [B, S, S, 32] bf16 = {R, G, B, depth, 17 heat-maps, 11 zero pad}.  The 17 fp32 planes of the
// reference (4.46 MB/sample at S=256) never exist.  One thread per pixel, 64 B written per pixel.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cnn_input_pack_kernel(const float *__restrict__ image, const float *__restrict__ depth, const float *__restrict__ kp,
                      int B, int S, int J, float scale, float denom, float rcp, int c_stride, __nv_bfloat16 *__restrict__ out) {
    __shared__ float s_mu[32][2];
    __shared__ float s_valid[32];
    const long npx = (long)S * S;
    const int blocks_per_img = (int)((npx + 255) / 256);
    const int b = blockIdx.x / blocks_per_img;
    const long px = (long)(blockIdx.x % blocks_per_img) * 256 + threadIdx.x;
    if (threadIdx.x < J) {
        const float kx = kp[((long)b * J + threadIdx.x) * 2], ky = kp[((long)b * J + threadIdx.x) * 2 + 1];
        s_mu[threadIdx.x][0] = __fmul_rn(kx, scale);
        s_mu[threadIdx.x][1] = __fmul_rn(ky, scale);
        s_valid[threadIdx.x] = (kx > 0.0f && ky > 0.0f) ? 1.0f : 0.0f;
    }
    __syncthreads();
    if (px >= npx) return;
    const int y = (int)(px / S), x = (int)(px - (long)y * S);
    float v[32];
    v[0] = __ldg(image + ((long)b * 3 + 0) * npx + px);
    v[1] = __ldg(image + ((long)b * 3 + 1) * npx + px);
    v[2] = __ldg(image + ((long)b * 3 + 2) * npx + px);
    v[3] = __ldg(depth + (long)b * npx + px);
#pragma unroll
    for (int j = 0; j < 28; ++j) {
        float r = 0.0f;
        if (j < J) {
            const float dx = __fsub_rn((float)x, s_mu[j][0]), dy = __fsub_rn((float)y, s_mu[j][1]);
            const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
            const float q0 = __fmul_rn(d2, rcp);
            const float q1 = __fmaf_rn(__fmaf_rn(-q0, denom, d2), rcp, q0);  // RN(d2 / denom)
            r = (q1 > 104.0f ? 0.0f : expf(-q1)) * s_valid[j];
        }
        v[4 + j] = r;
    }
    uint4 *dst = (uint4 *)(out + ((long)b * npx + px) * c_stride);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float f[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = v[q * 8 + k];
        dst[q] = pack8(f);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Depthwise 3x3 conv (pad 1, stride 1 or 2) + folded BatchNorm + activation, NHWC bf16.
// A CTA owns an output tile of TH x TW pixels and a slab of 64 channels: the input halo tile is staged in shared
// memory once with coalesced 16-byte loads (each input element crosses HBM/L2 once per tile instead of nine
// times), then every thread computes 8 channels of several output pixels from shared memory with its 72 filter
// taps in registers.  Optionally emits per-(image, tile, channel) partial sums of the OUTPUT (the squeeze of
// SE / ECA, cnn.py:22-23,40-41): pool [B, tiles_per_image, C], written -- not accumulated -- and reduced by the
// consumers in a fixed order, so the result is deterministic.
// ---------------------------------------------------------------------------------------------------------
constexpr int kDwSlab = 64;  // channels per CTA

template <int STRIDE>
struct DwTile {
    static constexpr int TH = STRIDE == 1 ? 8 : 4, TW = STRIDE == 1 ? 16 : 8;     // output tile
    static constexpr int IH = (TH - 1) * STRIDE + 3, IW = (TW - 1) * STRIDE + 3;  // input halo tile
    static constexpr int kSmem = IH * IW * kDwSlab * 2;
};

// The halo tile of (image b, tile) x 64-channel slab is ONE 4-D TMA box {64 channels, IW, IH, 1} over the NHWC tensor:
// the box starts one pixel above / left of the tile, and everything outside the image or the channel range is zero
// filled by the copy engine (= the convolution padding).  The per-thread cp.async version spent 27 % of the kernel's
// instructions on halo index arithmetic.  Called by one thread; the bytes land on `bar`.
template <int STRIDE>
__device__ __forceinline__ void dw_stage_tile(unsigned char *buf, const CUtensorMap *map, uint64_t *bar, int b, int tile,
                                              int tiles_x, int c_slab) {
    using T = DwTile<STRIDE>;
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    mbar_expect_tx(bar, T::kSmem);
    tma_load_4d(buf, map, bar, c_slab, tx * T::TW * STRIDE - 1, ty * T::TH * STRIDE - 1, b);
}

// PERSISTENT: a CTA walks (image, tile) pairs of its 64-channel slab; the next tile's halo is in flight (cp.async, second
// buffer) while the current one is computed -- the one-tile-per-CTA version left every global-load latency exposed at
// two CTAs per SM.
// four bf16 channels as two fp32 pairs (the operands of the packed FFMA2 of sm_100: two fused multiply-adds per issue slot)
__device__ __forceinline__ void unpack4(const uint2 &p, float2 (&f)[2]) {
    f[0] = make_float2(__uint_as_float(p.x << 16), __uint_as_float(p.x & 0xffff0000u));
    f[1] = make_float2(__uint_as_float(p.y << 16), __uint_as_float(p.y & 0xffff0000u));
}
__device__ __forceinline__ uint2 pack4(const float (&f)[4]) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
    return make_uint2(*(uint32_t *)&p0, *(uint32_t *)&p1);
}

// activation, store (bf16) and the per-thread partial sums of one 4-channel output (see `pool` / `stats` of the kernel)
template <int ACT>
__device__ __forceinline__ void dw_emit(float (&a)[4], int stats, float (&psum)[4], float (&psq)[4], __nv_bfloat16 *dst) {
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = act_t<ACT>(a[k]);
    const uint2 pk = pack4(a);
    *(uint2 *)dst = pk;
    if (stats) {          // BatchNorm statistics are those of the tensor the next kernel will read: the rounded values
        float2 r[2];
        unpack4(pk, r);
        psum[0] += r[0].x; psum[1] += r[0].y; psum[2] += r[1].x; psum[3] += r[1].y;
        psq[0] = fmaf(r[0].x, r[0].x, psq[0]); psq[1] = fmaf(r[0].y, r[0].y, psq[1]);
        psq[2] = fmaf(r[1].x, r[1].x, psq[2]); psq[3] = fmaf(r[1].y, r[1].y, psq[3]);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) psum[k] += a[k];
    }
}

// BatchNorm-backward tail (ACT 11 / 12, the data gradient of a stride-1 depthwise layer whose input came out of a
// BatchNorm + ReLU / SiLU): the value just computed is dA of that layer; with its saved conv output y (4 channels, `yv`) the
// kernel emits dz = dA * act'(y * scale + shift) instead, and the per-thread partial sums of dz and dz * y -- the first
// pass of that BatchNorm's backward (pose_bn_bwd_bf16's reduction) never runs.
template <int ACT>
__device__ __forceinline__ void dw_emit_bnb(const float (&a)[4], const uint2 &yv, const float2 (&sc)[2], const float2 (&sh)[2],
                                            float (&psum)[4], float (&psq)[4], __nv_bfloat16 *dst) {
    float2 y[2];
    unpack4(yv, y);
    float dz[4];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const float2 z = __ffma2_rn(y[k], sc[k], sh[k]);
        float2 g;
        if (ACT == 12) {
            const float tx = tanh_approx(0.5f * z.x), ty = tanh_approx(0.5f * z.y);
            const float sx = 0.5f + 0.5f * tx, sy = 0.5f + 0.5f * ty;
            g = make_float2(sx * (1.0f + z.x * (1.0f - sx)), sy * (1.0f + z.y * (1.0f - sy)));
        } else {
            g = make_float2(z.x > 0.f ? 1.f : 0.f, z.y > 0.f ? 1.f : 0.f);
        }
        dz[2 * k] = a[2 * k] * g.x;
        dz[2 * k + 1] = a[2 * k + 1] * g.y;
        psum[2 * k] += dz[2 * k];
        psum[2 * k + 1] += dz[2 * k + 1];
        psq[2 * k] = fmaf(dz[2 * k], y[k].x, psq[2 * k]);
        psq[2 * k + 1] = fmaf(dz[2 * k + 1], y[k].y, psq[2 * k + 1]);
    }
    *(uint2 *)dst = pack4(dz);
}

// A thread owns FOUR channels (16 channel groups x 16 pixel lanes): 36 filter taps in registers instead of 72 keeps the
// kernel under 85 registers, so three CTAs (24 warps, three halo tiles in flight) share an SM; with eight channels per
// thread it ran one CTA per SM at 20 % of the HBM roofline.
template <int STRIDE, int ACT>
__global__ void __launch_bounds__(256, ACT >= 10 ? 2 : 3)
dwconv3x3_kernel(const __grid_constant__ CUtensorMap mapX, const float *__restrict__ Wd, const float *__restrict__ bias,
                 int B, int H, int W, int C, __nv_bfloat16 *__restrict__ Y, float *__restrict__ pool, int stats, int Ho, int Wo,
                 int tiles_x, int tiles_y, const __nv_bfloat16 *__restrict__ Ybn, const float *__restrict__ bn_ss) {
    constexpr bool BNB = ACT >= 10;        // BatchNorm-backward tail (stride 1 only): see dw_emit_bnb
    // pool (optional): stats == 0: [B * tiles, C] sums of the activated outputs (SE / ECA squeeze);
    //                  stats == 1: [B * tiles, 2, C] sums of y and y^2 of the bf16-ROUNDED outputs = the first stage of
    //                  the BatchNorm batch statistics (the separate pass over the conv output disappears)
    using T = DwTile<STRIDE>;
    extern __shared__ __align__(128) unsigned char s_dyn[];     // 2 halo buffers, then the pooled-sum scratch
    float(*s_part)[kDwSlab] = (float(*)[kDwSlab])(s_dyn + 2 * T::kSmem);
    float(*s_part2)[kDwSlab] = s_part + 16;
    __shared__ uint64_t s_bar[2];                                // one per halo buffer (TMA completion)
    if (threadIdx.x == 0) {
        mbar_init(&s_bar[0], 1);
        mbar_init(&s_bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int per_img = tiles_x * tiles_y;
    const long n_items = (long)B * per_img;
    const int c_slab = blockIdx.y * kDwSlab;
    const int cg = threadIdx.x & 15, pl = threadIdx.x >> 4;      // 16 channel groups of 4 x 16 pixel lanes
    const int c0 = c_slab + cg * 4;
    const bool c_ok = c0 < C;
    float2 w[9][2], bs[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        bs[k] = (c_ok && bias != nullptr) ? __ldg((const float2 *)(bias + c0) + k) : make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t][k] = c_ok ? __ldg((const float2 *)(Wd + (long)t * C + c0) + k) : make_float2(0.f, 0.f);
    }
    float2 bsc[2], bsh[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        bsc[k] = (BNB && c_ok) ? __ldg((const float2 *)(bn_ss + c0) + k) : make_float2(0.f, 0.f);
        bsh[k] = (BNB && c_ok) ? __ldg((const float2 *)(bn_ss + C + c0) + k) : make_float2(0.f, 0.f);
    }
    long item = blockIdx.x;
    if (item < n_items && threadIdx.x == 0)
        dw_stage_tile<STRIDE>(s_dyn, &mapX, &s_bar[0], (int)(item / per_img), (int)(item % per_img), tiles_x, c_slab);
    for (int it = 0; item < n_items; item += gridDim.x, ++it) {
        const long next = item + gridDim.x;
        if (next < n_items && threadIdx.x == 0)     // (the buffer was released by the barrier that ended the previous iteration)
            dw_stage_tile<STRIDE>(s_dyn + ((it + 1) & 1) * T::kSmem, &mapX, &s_bar[(it + 1) & 1], (int)(next / per_img),
                                  (int)(next % per_img), tiles_x, c_slab);
        const unsigned char *s_in = s_dyn + (it & 1) * T::kSmem + cg * 8;      // this thread's channels of pixel 0
        const int b = (int)(item / per_img), tile = (int)(item - (long)b * per_img);
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int oy0 = ty * T::TH, ox0 = tx * T::TW;
        // BNB: this thread's column of the consumer layer's saved conv outputs (one 8-byte load per tile row), requested
        // before the wait on the halo tile so that all of them are in flight together
        uint2 yv[BNB ? T::TH : 1];
        if (BNB) {
#pragma unroll
            for (int q = 0; q < (BNB ? T::TH : 1); ++q) {
                yv[q] = make_uint2(0u, 0u);
                if (ox0 + pl < Wo && oy0 + q < Ho && c_ok)
                    yv[q] = __ldg((const uint2 *)(Ybn + (((long)b * Ho + oy0 + q) * Wo + ox0 + pl) * C + c0));
            }
        }
        mbar_wait(&s_bar[it & 1], (it >> 1) & 1);
        float psum[4], psq[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) psum[k] = psq[k] = 0.f;
        constexpr int kPix = T::TH * T::TW;
        if (STRIDE == 1) {
            // a thread owns one column of the 8 x 16 tile, two output rows at a time: the 4 x 3 input vectors of a pair are
            // read once and feed both outputs (24 shared-memory loads per 4 outputs instead of 36)
            static_assert(STRIDE != 1 || T::TW == 16, "one pixel lane per tile column");
            const int pxx = pl;
            const int ox = ox0 + pxx;
#pragma unroll(BNB ? T::TH / 2 : 1)
            for (int r0 = 0; r0 < T::TH; r0 += 2) {
                float2 v0[2] = {bs[0], bs[1]}, v1[2] = {bs[0], bs[1]};
#pragma unroll
                for (int ir = 0; ir < 4; ++ir)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        float2 f[2];
                        unpack4(*(const uint2 *)(s_in + ((r0 + ir) * T::IW + pxx + kx) * 128), f);
                        if (ir < 3) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) v0[k] = __ffma2_rn(f[k], w[ir * 3 + kx][k], v0[k]);
                        }
                        if (ir > 0) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) v1[k] = __ffma2_rn(f[k], w[(ir - 1) * 3 + kx][k], v1[k]);
                        }
                    }
                float a0[4] = {v0[0].x, v0[0].y, v0[1].x, v0[1].y}, a1[4] = {v1[0].x, v1[0].y, v1[1].x, v1[1].y};
                if (ox < Wo && c_ok) {
                    __nv_bfloat16 *d0 = Y + (((long)b * Ho + oy0 + r0) * Wo + ox) * C + c0, *d1 = d0 + (long)Wo * C;
                    if (BNB) {
                        if (oy0 + r0 < Ho) dw_emit_bnb<ACT>(a0, yv[BNB ? r0 : 0], bsc, bsh, psum, psq, d0);
                        if (oy0 + r0 + 1 < Ho) dw_emit_bnb<ACT>(a1, yv[BNB ? r0 + 1 : 0], bsc, bsh, psum, psq, d1);
                    } else {
                        if (oy0 + r0 < Ho) dw_emit<ACT>(a0, stats, psum, psq, d0);
                        if (oy0 + r0 + 1 < Ho) dw_emit<ACT>(a1, stats, psum, psq, d1);
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int q = 0; q < kPix / 16; ++q) {
                const int p = q * 16 + pl;
                const int py = p / T::TW, pxx = p - py * T::TW;
                const int oy = oy0 + py, ox = ox0 + pxx;
                float2 v[2] = {bs[0], bs[1]};
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        float2 f[2];
                        unpack4(*(const uint2 *)(s_in + ((py * STRIDE + ky) * T::IW + pxx * STRIDE + kx) * 128), f);
#pragma unroll
                        for (int k = 0; k < 2; ++k) v[k] = __ffma2_rn(f[k], w[ky * 3 + kx][k], v[k]);
                    }
                float acc[4] = {v[0].x, v[0].y, v[1].x, v[1].y};
                if (!BNB && oy < Ho && ox < Wo && c_ok) dw_emit<ACT>(acc, stats, psum, psq, Y + (((long)b * Ho + oy) * Wo + ox) * C + c0);
            }
        }
        if (pool != nullptr) {  // fixed-order reduction over the 16 pixel lanes of the CTA, one write per channel
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s_part[pl][cg * 4 + k] = psum[k];
                s_part2[pl][cg * 4 + k] = psq[k];
            }
            __syncthreads();
            const int which = threadIdx.x >> 6, ch = threadIdx.x & 63;      // threads 0..63: sums, 64..127: sums of squares
            if (which <= stats && c_slab + ch < C) {
                const float(*src)[kDwSlab] = which ? s_part2 : s_part;
                float t = 0.f;
#pragma unroll 8
                for (int q = 0; q < 16; ++q) t += src[q][ch];
                const long part = (long)b * per_img + tile;
                pool[(stats ? part * 2 + which : part) * C + c_slab + ch] = t;
            }
        }
        __syncthreads();     // the buffer just read is the target of the copy issued in the next iteration
    }
}

// ---------------------------------------------------------------------------------------------------------
// Channel sums over the pixels of each image: X [B, HW, C] bf16 -> partial sums S [B, chunks, C] fp32 (written)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pool_sum_kernel(const __nv_bfloat16 *__restrict__ X, int HW, int C, float *__restrict__ S, int chunks) {
    // all 256 threads work whatever C is: G = min(C / 8, 256) channel groups side by side, 256 / G pixel lanes each (with a
    // thread per channel group only, a 128-channel map kept 16 threads of the CTA busy: 1 TB/s), four independent 16-byte
    // loads in flight per thread; the lanes are folded through shared memory in a fixed order (deterministic)
    __shared__ float red[256 * 8];
    const int b = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
    const int C8 = C >> 3;
    const int per = (HW + chunks - 1) / chunks;
    const int p0 = chunk * per, p1 = min(HW, p0 + per);
    const int G = C8 < 256 ? C8 : 256, lanes = 256 / G;
    const int g = threadIdx.x % G, ln = threadIdx.x / G;
    const __nv_bfloat16 *base = X + (long)b * HW * C;
    for (int cg0 = 0; cg0 < C8; cg0 += G) {
        const int cg = cg0 + g;
        float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (ln < lanes && cg < C8) {
            int p = p0 + ln;
            for (; p + 3 * lanes < p1; p += 4 * lanes) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg((const uint4 *)(base + (long)(p + u * lanes) * C + cg * 8));
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float f[8];
                    unpack8(v[u], f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) s[k] += f[k];
                }
            }
            for (; p < p1; p += lanes) {
                float f[8];
                unpack8(__ldg((const uint4 *)(base + (long)p * C + cg * 8)), f);
#pragma unroll
                for (int k = 0; k < 8; ++k) s[k] += f[k];
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) red[k * 256 + threadIdx.x] = s[k];
        __syncthreads();
        if (ln == 0 && cg < C8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                float t = 0.f;
                for (int q = 0; q < lanes; ++q) t += red[k * 256 + q * G + g];
                S[((long)b * chunks + chunk) * C + cg * 8 + k] = t;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// SE gate (cnn.py:9-26): gate = sigmoid(W2 . act(W1 . mean)); one CTA per image.
// ECA gate (cnn.py:29-45): gate = sigmoid(conv1d_k(mean)) across channels (zero padded).
// Both read channel SUMS and scale by inv_hw.  `feat_out` (optional, bf16 [B, C]) receives mean * gate: the
// global_features tail (ECABlock followed by AdaptiveAvgPool2d(1), cnn.py:612-613) needs nothing else.
// ---------------------------------------------------------------------------------------------------------
constexpr int kSeSamples = 2;  // images per CTA: every weight row fetched once serves both

__global__ void __launch_bounds__(256)
se_gate_kernel(const float *__restrict__ pool, int parts, float inv_hw, const float *__restrict__ W1,
               const float *__restrict__ W2, int B, int C, int Cr, int act, float *__restrict__ gate) {
    extern __shared__ float sm[];  // mean[kSeSamples][C] + hidden[kSeSamples][Cr]
    float *mean = sm, *hid = sm + kSeSamples * C;
    const int b0 = blockIdx.x * kSeSamples, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kSeSamples * C; i += 256) {
        const int sb = i / C, c = i - sb * C;
        float t = 0.f;
        if (b0 + sb < B)
            for (int q = 0; q < parts; ++q) t += pool[((long)(b0 + sb) * parts + q) * C + c];
        mean[i] = t * inv_hw;
    }
    __syncthreads();
    for (int r = warp; r < Cr; r += 8) {          // a warp per hidden unit: W1 row read once, coalesced
        float s[kSeSamples];
#pragma unroll
        for (int q = 0; q < kSeSamples; ++q) s[q] = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float w = __ldg(W1 + (long)r * C + c);
#pragma unroll
            for (int q = 0; q < kSeSamples; ++q) s[q] = fmaf(w, mean[q * C + c], s[q]);
        }
#pragma unroll
        for (int q = 0; q < kSeSamples; ++q) {
            s[q] = warp_sum(s[q]);
            if (lane == 0) hid[q * Cr + r] = act_f(s[q], act);
        }
    }
    __syncthreads();
    for (int c = warp; c < C; c += 8) {            // a warp per output channel: W2 row (Cr floats) coalesced
        float s[kSeSamples];
#pragma unroll
        for (int q = 0; q < kSeSamples; ++q) s[q] = 0.f;
        for (int r = lane; r < Cr; r += 32) {
            const float w = __ldg(W2 + (long)c * Cr + r);
#pragma unroll
            for (int q = 0; q < kSeSamples; ++q) s[q] = fmaf(w, hid[q * Cr + r], s[q]);
        }
#pragma unroll
        for (int q = 0; q < kSeSamples; ++q) {
            s[q] = warp_sum(s[q]);
            if (lane == 0 && b0 + q < B) gate[(long)(b0 + q) * C + c] = 1.0f / (1.0f + __expf(-s[q]));
        }
    }
}

__global__ void __launch_bounds__(256)
eca_gate_kernel(const float *__restrict__ pool, int parts, float inv_hw, const float *__restrict__ w, int k, int C,
                float *__restrict__ gate, __nv_bfloat16 *__restrict__ feat_out) {
    extern __shared__ float mean[];  // [C]
    const int b = blockIdx.x;
    const int half = (k - 1) / 2;
    for (int c = threadIdx.x; c < C; c += 256) {
        float t = 0.f;
        for (int q = 0; q < parts; ++q) t += pool[((long)b * parts + q) * C + c];
        mean[c] = t * inv_hw;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
        for (int t = 0; t < k; ++t) {
            const int cc = c + t - half;
            if (cc >= 0 && cc < C) s = fmaf(__ldg(w + t), mean[cc], s);
        }
        const float g = 1.0f / (1.0f + __expf(-s));
        if (gate != nullptr) gate[(long)b * C + c] = g;
        if (feat_out != nullptr) feat_out[(long)b * C + c] = __float2bfloat16_rn(mean[c] * g);
    }
}

// X[b, p, c] = X[b, p, c] * mul[b, c] + add[b, c]   (either vector may be null); in place or to Y
__global__ void __launch_bounds__(256)
channel_affine_kernel(const __nv_bfloat16 *__restrict__ X, const float *__restrict__ mul, const __nv_bfloat16 *__restrict__ add,
                      long HW, int C, long total8, __nv_bfloat16 *__restrict__ Y) {
    const int C8 = C >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        const long b = (i / C8) / HW;
        float f[8];
        unpack8(__ldg((const uint4 *)X + i), f);
        if (mul != nullptr) {
            const float4 m0 = __ldg((const float4 *)(mul + b * C + cg * 8)), m1 = __ldg((const float4 *)(mul + b * C + cg * 8) + 1);
            f[0] *= m0.x; f[1] *= m0.y; f[2] *= m0.z; f[3] *= m0.w;
            f[4] *= m1.x; f[5] *= m1.y; f[6] *= m1.z; f[7] *= m1.w;
        }
        if (add != nullptr) {
            float a[8];
            unpack8(__ldg((const uint4 *)(add + b * C + cg * 8)), a);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += a[k];
        }
        ((uint4 *)Y)[i] = pack8(f);
    }
}

// ---------------------------------------------------------------------------------------------------------
// CoordAttention (cnn.py:48-98): directional means, then out = x * a_h * a_w.
//   coord_pool:  X [B,H,W,C] -> P [B, H+W, C] bf16: rows 0..H-1 = mean over w, rows H.. = mean over h
//   coord_apply: G [B, H+W, 2C] bf16 sigmoid gates (cols 0..C-1 from conv_h, C..2C-1 from conv_w)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
coord_pool_kernel(const __nv_bfloat16 *__restrict__ X, int H, int W, int C, __nv_bfloat16 *__restrict__ P) {
    const int b = blockIdx.x, C8 = C >> 3;
    const int n_rows = H + W;
    for (int i = threadIdx.x; i < n_rows * C8; i += 256) {
        const int r = i / C8, cg = i - r * C8;
        float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const bool is_h = r < H;
        const int n = is_h ? W : H;
        for (int q = 0; q < n; ++q) {
            const int y = is_h ? r : q, x = is_h ? q : r - H;
            float f[8];
            unpack8(__ldg((const uint4 *)(X + (((long)b * H + y) * W + x) * C + cg * 8)), f);
#pragma unroll
            for (int k = 0; k < 8; ++k) s[k] += f[k];
        }
        const float inv = 1.0f / (float)n;
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] *= inv;
        *(uint4 *)(P + ((long)b * n_rows + r) * C + cg * 8) = pack8(s);
    }
}

__global__ void __launch_bounds__(256)
coord_apply_kernel(const __nv_bfloat16 *__restrict__ X, const __nv_bfloat16 *__restrict__ G, int H, int W, int C,
                   long total8, __nv_bfloat16 *__restrict__ Y) {
    const int C8 = C >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long p = i / C8;
        const int x = (int)(p % W);
        p /= W;
        const int y = (int)(p % H);
        const long b = p / H;
        float f[8], gh[8], gw[8];
        unpack8(__ldg((const uint4 *)X + i), f);
        unpack8(__ldg((const uint4 *)(G + ((b * (H + W) + y) * 2L * C) + cg * 8)), gh);
        unpack8(__ldg((const uint4 *)(G + ((b * (H + W) + H + x) * 2L * C) + C + cg * 8)), gw);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = f[k] * gh[k] * gw[k];
        ((uint4 *)Y)[i] = pack8(f);
    }
}

// nn.AdaptiveAvgPool2d((OH, OW)) (cnn.py:602), NHWC bf16: window i covers [floor(i H / OH), ceil((i + 1) H / OH))
// (windows overlap when H is not a multiple of OH), fp32 sum / window size
__global__ void __launch_bounds__(256)
adaptive_avgpool_kernel(const __nv_bfloat16 *__restrict__ X, int H, int W, int C, int OH, int OW, long total8,
                        __nv_bfloat16 *__restrict__ Y) {
    const int C8 = C >> 3;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long p = i / C8;
        const int ox = (int)(p % OW);
        p /= OW;
        const int oy = (int)(p % OH);
        const long b = p / OH;
        const int y0 = (oy * H) / OH, y1 = ((oy + 1) * H + OH - 1) / OH;
        const int x0 = (ox * W) / OW, x1 = ((ox + 1) * W + OW - 1) / OW;
        float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) {
                float f[8];
                unpack8(__ldg((const uint4 *)(X + (((b * H + y) * W) + x) * C + cg * 8)), f);
#pragma unroll
                for (int k = 0; k < 8; ++k) s[k] += f[k];
            }
        const float inv = 1.0f / (float)((y1 - y0) * (x1 - x0));
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] *= inv;
        ((uint4 *)Y)[i] = pack8(s);
    }
}

// 2x2 average pooling (AdaptiveAvgPool2d(8) on a 16x16 map, cnn.py:602), NHWC bf16
__global__ void __launch_bounds__(256)
avgpool2x2_kernel(const __nv_bfloat16 *__restrict__ X, int H, int W, int C, long total8, __nv_bfloat16 *__restrict__ Y) {
    const int C8 = C >> 3, Ho = H >> 1, Wo = W >> 1;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(i % C8);
        long p = i / C8;
        const int x = (int)(p % Wo);
        p /= Wo;
        const int y = (int)(p % Ho);
        const long b = p / Ho;
        float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                float f[8];
                unpack8(__ldg((const uint4 *)(X + (((b * H + 2 * y + dy) * W) + 2 * x + dx) * C + cg * 8)), f);
#pragma unroll
                for (int k = 0; k < 8; ++k) s[k] += f[k];
            }
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] *= 0.25f;
        ((uint4 *)Y)[i] = pack8(s);
    }
}

// S [B, parts, C] fp32 partial sums -> bf16 means [B, C] (operand of the WASP global-branch GEMM)
__global__ void __launch_bounds__(256)
sums_to_bf16_kernel(const float *__restrict__ S, int parts, int C, float scale, long n, __nv_bfloat16 *__restrict__ out) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const long b = i / C;
        const int c = (int)(i - b * C);
        float t = 0.f;
        for (int q = 0; q < parts; ++q) t += S[(b * parts + q) * C + c];
        out[i] = __float2bfloat16_rn(t * scale);
    }
}

static int grid_for(long items, int per_block = 256, int max_waves = 16) {
    long blocks = (items + per_block - 1) / per_block;
    long cap = (long)kNumSMs * max_waves;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

}  // namespace pose

using namespace pose;

POSE_API int pose_cnn_input_pack(const float *image, const float *depth, const float *kp, int B, int S, int J,
                                 float sigma, void *out, pose_stream_t stream) {
    return pose_cnn_input_pack_ex(image, depth, kp, B, S, J, sigma, 32, out, stream);
}

POSE_API int pose_cnn_input_pack_ex(const float *image, const float *depth, const float *kp, int B, int S, int J,
                                    float sigma, int c_stride, void *out, pose_stream_t stream) {
    if (!image || !depth || !kp || !out) return POSE_E_NULL;
    if (B <= 0 || S <= 0 || J <= 0 || J > 28 || !(sigma > 0.f) || c_stride < 32 || c_stride % 8) return POSE_E_SHAPE;
    if ((uintptr_t)out % 16) return POSE_E_ALIGN;
    const float denom = (float)(2.0 * (double)sigma * (double)sigma);
    const long npx = (long)S * S;
    const int blocks_per_img = (int)((npx + 255) / 256);
    cnn_input_pack_kernel<<<B * blocks_per_img, 256, 0, (cudaStream_t)stream>>>(
        image, depth, kp, B, S, J, (float)(S - 1), denom, (float)(1.0 / (double)denom), c_stride, (__nv_bfloat16 *)out);
    return launch_status();
}

POSE_API int pose_dwconv3x3_pool_parts(int H, int W, int stride) {
    if (H <= 0 || W <= 0 || (stride != 1 && stride != 2)) return POSE_E_SHAPE;
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    const int th = stride == 1 ? DwTile<1>::TH : DwTile<2>::TH, tw = stride == 1 ? DwTile<1>::TW : DwTile<2>::TW;
    return ((Ho + th - 1) / th) * ((Wo + tw - 1) / tw);
}

static int dwconv3x3_launch(const void *X, int B, int H, int W, int C, const float *Wd, const float *bias, int stride, int act,
                            void *Y, float *pool_sum, int pool_parts, int stats, pose_stream_t stream,
                            const void *Ybn = nullptr, const float *bn_ss = nullptr);

POSE_API int pose_dwconv3x3_bf16(const void *X, int B, int H, int W, int C, const float *Wd, const float *bias, int stride,
                                 int act, void *Y, float *pool_sum, int pool_parts, pose_stream_t stream) {
    return dwconv3x3_launch(X, B, H, W, C, Wd, bias, stride, act, Y, pool_sum, pool_parts, 0, stream);
}

POSE_API int pose_dwconv3x3_bn_stats_bf16(const void *X, int B, int H, int W, int C, const float *Wd, int stride, void *Y,
                                          float *partials, long cap_floats, pose_stream_t stream) {
    if (!partials) return POSE_E_NULL;
    if (H <= 0 || W <= 0 || (stride != 1 && stride != 2)) return POSE_E_SHAPE;
    const int parts = pose_dwconv3x3_pool_parts(H, W, stride);
    if ((long)B * parts * 2 * C > cap_floats) return POSE_E_WORKSPACE;
    return dwconv3x3_launch(X, B, H, W, C, Wd, nullptr, stride, 0, Y, partials, parts, 1, stream);
}

POSE_API int pose_dwconv3x3_bnbwd_bf16(const void *dY, int B, int H, int W, int C, const float *Wflip, const void *Yprev,
                                       const float *scale_shift_prev, int act_prev, void *dZ, float *partials, long cap_floats,
                                       pose_stream_t stream) {
    if (!partials || !Yprev || !scale_shift_prev) return POSE_E_NULL;
    if (H <= 0 || W <= 0) return POSE_E_SHAPE;
    if (act_prev != 1 && act_prev != 2) return POSE_E_UNSUPPORTED;
    if ((uintptr_t)Yprev % 8) return POSE_E_ALIGN;
    const int parts = pose_dwconv3x3_pool_parts(H, W, 1);
    if ((long)B * parts * 2 * C > cap_floats) return POSE_E_WORKSPACE;
    return dwconv3x3_launch(dY, B, H, W, C, Wflip, nullptr, 1, 10 + act_prev, dZ, partials, parts, 1, stream, Yprev,
                            scale_shift_prev);
}

static int dwconv3x3_launch(const void *X, int B, int H, int W, int C, const float *Wd, const float *bias, int stride, int act,
                            void *Y, float *pool_sum, int pool_parts, int stats, pose_stream_t stream, const void *Ybn,
                            const float *bn_ss) {
    if (!X || !Wd || (!bias && !stats) || !Y) return POSE_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8 || (stride != 1 && stride != 2)) return POSE_E_SHAPE;
    if ((uintptr_t)X % 16 || (uintptr_t)Y % 16) return POSE_E_ALIGN;
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    const int parts = pose_dwconv3x3_pool_parts(H, W, stride);
    if (pool_sum != nullptr && pool_parts != parts) return POSE_E_SHAPE;  // pool_sum is [B, parts, C]
    const int th = stride == 1 ? DwTile<1>::TH : DwTile<2>::TH, tw = stride == 1 ? DwTile<1>::TW : DwTile<2>::TW;
    const int tiles_y = (Ho + th - 1) / th, tiles_x = (Wo + tw - 1) / tw;
    const int gy = (C + kDwSlab - 1) / kDwSlab;
    long gx = (long)B * tiles_x * tiles_y;
    const long cap = ((long)kNumSMs * (act >= 10 ? 2 : 3) + gy - 1) / gy;   // resident CTAs per SM (3; 2 with the BatchNorm tail) over all channel slabs
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, gy);
    cudaStream_t s = (cudaStream_t)stream;
    if (act < 0 || (act > 4 && act != 11 && act != 12) || (act >= 10 && (stride != 1 || !Ybn || !bn_ss))) return POSE_E_UNSUPPORTED;
    const int smem = 2 * (stride == 1 ? DwTile<1>::kSmem : DwTile<2>::kSmem) + 2 * 16 * kDwSlab * 4;
    CUtensorMap mapX;
    {
        const int e = make_map_dw_halo(&mapX, X, B, H, W, C, stride == 1 ? DwTile<1>::IW : DwTile<2>::IW,
                                       stride == 1 ? DwTile<1>::IH : DwTile<2>::IH);
        if (e) return e;
    }
#define DW_LAUNCH(S_, A_)                                                                                             \
    {                                                                                                                 \
        static bool cfg = false;                                                                                      \
        if (!cfg) {                                                                                                   \
            cudaError_t ce = cudaFuncSetAttribute(dwconv3x3_kernel<S_, A_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); \
            if (ce != cudaSuccess) return (int)ce;                                                                    \
            cfg = true;                                                                                               \
        }                                                                                                             \
        dwconv3x3_kernel<S_, A_><<<grid, 256, smem, s>>>(mapX, Wd, bias, B, H, W, C, (__nv_bfloat16 *)Y, \
                                                         pool_sum, stats, Ho, Wo, tiles_x, tiles_y,                   \
                                                         (const __nv_bfloat16 *)Ybn, bn_ss);                          \
    }
#define DW_ACT(S_)                                                                                                    \
    switch (act) {                                                                                                    \
        case 0: DW_LAUNCH(S_, 0); break;                                                                              \
        case 1: DW_LAUNCH(S_, 1); break;                                                                              \
        case 2: DW_LAUNCH(S_, 2); break;                                                                              \
        case 3: DW_LAUNCH(S_, 3); break;                                                                              \
        default: DW_LAUNCH(S_, 4); break;                                                                             \
    }
    if (act == 11) { DW_LAUNCH(1, 11) } else if (act == 12) { DW_LAUNCH(1, 12) } else
    if (stride == 1) { DW_ACT(1) } else { DW_ACT(2) }
#undef DW_ACT
#undef DW_LAUNCH
    return launch_status();
}

POSE_API int pose_pool_sum_bf16(const void *X, int B, int HW, int C, float *sums, int parts, pose_stream_t stream) {
    if (!X || !sums) return POSE_E_NULL;
    if (B <= 0 || HW <= 0 || C <= 0 || C % 8 || parts < 1) return POSE_E_SHAPE;
    const int chunks = parts;
    pool_sum_kernel<<<B * chunks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)X, HW, C, sums, chunks);
    return launch_status();
}

POSE_API int pose_se_gate(const float *pool_sum, int parts, float inv_hw, const float *W1, const float *W2, int B, int C,
                          int Cr, int act, float *gate, pose_stream_t stream) {
    if (!pool_sum || !W1 || !W2 || !gate) return POSE_E_NULL;
    if (B <= 0 || C <= 0 || Cr <= 0 || parts < 1) return POSE_E_SHAPE;
    const size_t smem = (size_t)kSeSamples * (C + Cr) * sizeof(float);
    if (smem > 48 * 1024) return POSE_E_UNSUPPORTED;
    se_gate_kernel<<<(B + kSeSamples - 1) / kSeSamples, 256, smem, (cudaStream_t)stream>>>(pool_sum, parts, inv_hw, W1, W2, B,
                                                                                          C, Cr, act, gate);
    return launch_status();
}

POSE_API int pose_eca_gate(const float *pool_sum, int parts, float inv_hw, const float *w, int k, int B, int C, float *gate,
                           void *feat_out, pose_stream_t stream) {
    if (!pool_sum || !w || (!gate && !feat_out)) return POSE_E_NULL;
    if (B <= 0 || C <= 0 || k <= 0 || !(k & 1) || parts < 1) return POSE_E_SHAPE;
    eca_gate_kernel<<<B, 256, (size_t)C * sizeof(float), (cudaStream_t)stream>>>(pool_sum, parts, inv_hw, w, k, C, gate,
                                                                                (__nv_bfloat16 *)feat_out);
    return launch_status();
}

POSE_API int pose_channel_affine_bf16(const void *X, const float *mul, const void *add, int B, long HW, int C, void *Y,
                                      pose_stream_t stream) {
    if (!X || !Y) return POSE_E_NULL;
    if (B <= 0 || HW <= 0 || C <= 0 || C % 8) return POSE_E_SHAPE;
    if ((uintptr_t)X % 16 || (uintptr_t)Y % 16 || (mul && (uintptr_t)mul % 16) || (add && (uintptr_t)add % 16)) return POSE_E_ALIGN;
    const long total8 = (long)B * HW * (C / 8);
    channel_affine_kernel<<<grid_for(total8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)X, mul, (const __nv_bfloat16 *)add,
                                                                             HW, C, total8, (__nv_bfloat16 *)Y);
    return launch_status();
}

POSE_API int pose_coord_pool_bf16(const void *X, int B, int H, int W, int C, void *P, pose_stream_t stream) {
    if (!X || !P) return POSE_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8) return POSE_E_SHAPE;
    coord_pool_kernel<<<B, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)X, H, W, C, (__nv_bfloat16 *)P);
    return launch_status();
}

POSE_API int pose_coord_apply_bf16(const void *X, const void *G, int B, int H, int W, int C, void *Y, pose_stream_t stream) {
    if (!X || !G || !Y) return POSE_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || C % 8) return POSE_E_SHAPE;
    const long total8 = (long)B * H * W * (C / 8);
    coord_apply_kernel<<<grid_for(total8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)X, (const __nv_bfloat16 *)G, H, W,
                                                                          C, total8, (__nv_bfloat16 *)Y);
    return launch_status();
}

POSE_API int pose_adaptive_avgpool_bf16(const void *X, int B, int H, int W, int C, int OH, int OW, void *Y, pose_stream_t stream) {
    if (!X || !Y) return POSE_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0 || OH > H || OW > W || C <= 0 || C % 8) return POSE_E_SHAPE;
    if ((uintptr_t)X % 16 || (uintptr_t)Y % 16) return POSE_E_ALIGN;
    const long total8 = (long)B * OH * OW * (C / 8);
    adaptive_avgpool_kernel<<<grid_for(total8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)X, H, W, C, OH, OW, total8,
                                                                               (__nv_bfloat16 *)Y);
    return launch_status();
}

POSE_API int pose_avgpool2x2_bf16(const void *X, int B, int H, int W, int C, void *Y, pose_stream_t stream) {
    if (!X || !Y) return POSE_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1) || C <= 0 || C % 8) return POSE_E_SHAPE;
    const long total8 = (long)B * (H / 2) * (W / 2) * (C / 8);
    avgpool2x2_kernel<<<grid_for(total8), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16 *)X, H, W, C, total8,
                                                                         (__nv_bfloat16 *)Y);
    return launch_status();
}

POSE_API int pose_sums_to_bf16(const float *sums, int parts, int B, int C, float scale, void *out, pose_stream_t stream) {
    if (!sums || !out) return POSE_E_NULL;
    if (B <= 0 || C <= 0 || parts < 1) return POSE_E_SHAPE;
    const long n = (long)B * C;
    sums_to_bf16_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(sums, parts, C, scale, n, (__nv_bfloat16 *)out);
    return launch_status();
}
