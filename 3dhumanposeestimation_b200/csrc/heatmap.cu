// heatmap.cu -- Gaussian key-point heat-map rendering.
// Reference semantics: src/models/common.py:23-51 (GaussianHeatmapGenerator.forward):
//   mu = kp * (hs - 1);  d2 = (x - mu_x)^2 + (y - mu_y)^2;  hm = exp(-d2 / (2 sigma^2)) * valid
// with valid = (kp > 0).all(-1) (strict; a NaN key-point yields a NaN plane as NaN * 0 does).
//
// Bandwidth-bound: the planes are written once (J * hs^2 * elt bytes per sample, 136 B read) with
// streaming 128-bit stores.  d2 is evaluated with separate, unfused multiplies and adds in the
// reference's order so the arg-max location of every plane is bit-identical; the division by the
// constant 2 sigma^2 is a correctly rounded Markstein sequence (mul, fma, fma) instead of the
// ~10-instruction IEEE division, keeping the ALU work under the store time.
#include "common.cuh"

namespace pose {

struct HeatmapConst {
    float scale;      // hs - 1
    float denom;      // fp32(2 sigma^2), as torch wraps the Python scalar
    float rcp_denom;  // RN(1 / denom)
};

// RN(-(d2 / denom)): q0 = RN(d2 * y), r = d2 - q0 * denom (exact in one fma), q1 = RN(q0 + r * y)
__device__ __forceinline__ float neg_div_const(float d2, float denom, float rcp) {
    float q0 = __fmul_rn(d2, rcp);
    float r = __fmaf_rn(-q0, denom, d2);
    float q1 = __fmaf_rn(r, rcp, q0);
    return -q1;
}

__device__ __forceinline__ float gauss(float dx2, float dy2, float denom, float rcp, float valid) {
    float d2 = __fadd_rn(dx2, dy2);
    float a = neg_div_const(d2, denom, rcp);
    // exp underflows to exactly 0 below -103.98 (smallest fp32 denormal is 2^-149)
    float v = (a < -104.0f) ? 0.0f : expf(a);
    return v * valid;
}

// layout 0: [B, J, hs, hs]; a thread owns 4 consecutive x of one plane and walks the rows, so the
// four dx^2 stay in registers and a warp writes 512 contiguous bytes per row.
template <typename OutT>
__global__ void __launch_bounds__(256)
heatmap_planes_kernel(const float *__restrict__ kp, long n_planes, int hs, HeatmapConst c, OutT *__restrict__ out,
                      int rows_per_cta) {
    const int X4 = hs >> 2;                        // float4 groups per row (hs % 4 == 0 on this path)
    const int RG = X4 >= 256 ? 1 : 256 / X4;       // rows covered by one sweep of the CTA
    const int r0 = threadIdx.x / X4;               // 0 when a row is wider than the CTA
    if (r0 >= RG) return;                          // tail threads when X4 does not divide 256
    const int bands = (hs + rows_per_cta - 1) / rows_per_cta;
    for (long work = blockIdx.x; work < n_planes * bands; work += gridDim.x) {
        const long plane = work / bands;
        const int band = (int)(work % bands);
        const float kx = __ldg(kp + plane * 2), ky = __ldg(kp + plane * 2 + 1);
        const float mux = __fmul_rn(kx, c.scale), muy = __fmul_rn(ky, c.scale);
        const float valid = (kx > 0.0f && ky > 0.0f) ? 1.0f : 0.0f;
        OutT *o = out + plane * (long)hs * hs;
        const int y_begin = band * rows_per_cta, y_end = min(hs, y_begin + rows_per_cta);
        for (int x4 = threadIdx.x % X4; x4 < X4; x4 += 256) {  // one trip unless hs > 1024
            const int x0 = x4 * 4;
            float dx2[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float dx = __fsub_rn((float)(x0 + k), mux);
                dx2[k] = __fmul_rn(dx, dx);
            }
            for (int y = y_begin + r0; y < y_end; y += RG) {
                float dy = __fsub_rn((float)y, muy);
                float dy2 = __fmul_rn(dy, dy);
                float v0 = gauss(dx2[0], dy2, c.denom, c.rcp_denom, valid);
                float v1 = gauss(dx2[1], dy2, c.denom, c.rcp_denom, valid);
                float v2 = gauss(dx2[2], dy2, c.denom, c.rcp_denom, valid);
                float v3 = gauss(dx2[3], dy2, c.denom, c.rcp_denom, valid);
                OutT *dst = o + (long)y * hs + x0;
                if constexpr (sizeof(OutT) == 4) {
                    st_stream_f4((float *)dst, make_float4(v0, v1, v2, v3));
                } else {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
                    uint2 pk = make_uint2(*(unsigned *)&lo, *(unsigned *)&hi);
                    *(uint2 *)dst = pk;
                }
            }
        }
    }
}

// generic scalar path: any hs, either layout (channels-last writes plane j at channel c_offset + j)
template <typename OutT>
__global__ void __launch_bounds__(256)
heatmap_generic_kernel(const float *__restrict__ kp, int B, int J, int hs, HeatmapConst c, OutT *__restrict__ out,
                       int layout, int c_stride, int c_offset) {
    const long total = (long)B * J * hs * hs;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        int j, x, y;
        long b;
        if (layout == 0) {
            x = (int)(i % hs);
            y = (int)((i / hs) % hs);
            long pj = i / ((long)hs * hs);
            j = (int)(pj % J);
            b = pj / J;
        } else {  // consecutive threads -> consecutive channels of one pixel
            j = (int)(i % J);
            long px = i / J;
            x = (int)(px % hs);
            y = (int)((px / hs) % hs);
            b = px / ((long)hs * hs);
        }
        const float kx = __ldg(kp + (b * J + j) * 2), ky = __ldg(kp + (b * J + j) * 2 + 1);
        const float mux = __fmul_rn(kx, c.scale), muy = __fmul_rn(ky, c.scale);
        const float valid = (kx > 0.0f && ky > 0.0f) ? 1.0f : 0.0f;
        float dx = __fsub_rn((float)x, mux), dy = __fsub_rn((float)y, muy);
        float v = gauss(__fmul_rn(dx, dx), __fmul_rn(dy, dy), c.denom, c.rcp_denom, valid);
        long off = (layout == 0) ? i : (((b * hs + y) * (long)hs + x) * c_stride + c_offset + j);
        if constexpr (sizeof(OutT) == 4) out[off] = v;
        else out[off] = __float2bfloat16_rn(v);
    }
}

// Heat-maps rendered straight into the operand of the ViT's heat-map patch embedding (transformers.py:41-46, :348-350):
// out[(b * PH + py) * PW + px][(j * P + ky) * P + kx] = bf16(gauss_j(y = py P + ky, x = px P + kx)) -- the [B, J, hs, hs] fp32
// planes of GaussianHeatmapGenerator never exist.  Same arithmetic as heatmap_planes_kernel (the bf16 value is the
// rounding of the reference's fp32 value).  A thread writes 8 consecutive kx (16 bytes); P % 8 == 0.
__global__ void __launch_bounds__(256)
heatmap_patchify_kernel(const float *__restrict__ kp, int J, int hs, int P, HeatmapConst c, long total8,
                        __nv_bfloat16 *__restrict__ out) {
    const int PW = hs / P, P8 = P >> 3;
    const int per_patch8 = J * P * P8;            // 16-byte groups of one patch row of the GEMM operand
    const int patches = PW * PW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long)gridDim.x * blockDim.x) {
        const int k8 = (int)(i % per_patch8);
        const long pidx = i / per_patch8;
        const int pp = (int)(pidx % patches);
        const long b = pidx / patches;
        const int py = pp / PW, px = pp - py * PW;
        const int kx8 = k8 % P8, t = k8 / P8;
        const int ky = t % P, j = t / P;
        const float kx = __ldg(kp + (b * J + j) * 2), kyf = __ldg(kp + (b * J + j) * 2 + 1);
        const float mux = __fmul_rn(kx, c.scale), muy = __fmul_rn(kyf, c.scale);
        const float valid = (kx > 0.0f && kyf > 0.0f) ? 1.0f : 0.0f;
        const float dy = __fsub_rn((float)(py * P + ky), muy), dy2 = __fmul_rn(dy, dy);
        const int x0 = px * P + kx8 * 8;
        __nv_bfloat162 pk[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float dx0 = __fsub_rn((float)(x0 + 2 * q), mux), dx1 = __fsub_rn((float)(x0 + 2 * q + 1), mux);
            pk[q] = __floats2bfloat162_rn(gauss(__fmul_rn(dx0, dx0), dy2, c.denom, c.rcp_denom, valid),
                                          gauss(__fmul_rn(dx1, dx1), dy2, c.denom, c.rcp_denom, valid));
        }
        ((uint4 *)out)[i] = make_uint4(*(unsigned *)&pk[0], *(unsigned *)&pk[1], *(unsigned *)&pk[2], *(unsigned *)&pk[3]);
    }
}

}  // namespace pose

POSE_API int pose_heatmap_patchify_bf16(const float *kp, int B, int J, int hs, float sigma, int P, void *out,
                                        pose_stream_t stream) {
    using namespace pose;
    if (!kp || !out) return POSE_E_NULL;
    if (B <= 0 || J <= 0 || hs <= 0 || P <= 0 || hs % P || P % 8 || !(sigma > 0.0f)) return POSE_E_SHAPE;
    if ((uintptr_t)out % 16) return POSE_E_ALIGN;
    HeatmapConst c;
    c.scale = (float)(hs - 1);
    c.denom = (float)(2.0 * (double)sigma * (double)sigma);
    c.rcp_denom = (float)(1.0 / (double)c.denom);
    const long total8 = (long)B * (hs / P) * (hs / P) * J * P * (P / 8);
    long blocks = (total8 + 255) / 256;
    const int grid = (int)(blocks < (long)kNumSMs * 16 ? blocks : (long)kNumSMs * 16);
    heatmap_patchify_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kp, J, hs, P, c, total8, (__nv_bfloat16 *)out);
    return launch_status();
}

POSE_API int pose_heatmap_render(const float *kp, int B, int J, int hs, float sigma, void *out, int out_dtype,
                                 int out_layout, int c_stride, int c_offset, pose_stream_t stream) {
    using namespace pose;
    if (!kp || !out) return POSE_E_NULL;
    if (B <= 0 || J <= 0 || hs <= 0 || !(sigma > 0.0f)) return POSE_E_SHAPE;
    if (out_dtype != 0 && out_dtype != 1) return POSE_E_UNSUPPORTED;
    if (out_layout != 0 && out_layout != 1) return POSE_E_UNSUPPORTED;
    if (out_layout == 1 && (c_offset < 0 || c_offset + J > c_stride)) return POSE_E_SHAPE;
    HeatmapConst c;
    c.scale = (float)(hs - 1);
    c.denom = (float)(2.0 * (double)sigma * (double)sigma);
    c.rcp_denom = (float)(1.0 / (double)c.denom);
    cudaStream_t s = (cudaStream_t)stream;
    const bool fast = out_layout == 0 && hs % 4 == 0 && hs <= 4096 &&
                      (uintptr_t)out % 16 == 0;
    if (fast) {
        const long n_planes = (long)B * J;
        // split a plane into row bands so small batches still fill 148 SMs x 8 CTAs
        int rows_per_cta = hs;
        const long target = (long)kNumSMs * 8;
        while (n_planes * ((hs + rows_per_cta - 1) / rows_per_cta) < target && rows_per_cta > 16) rows_per_cta >>= 1;
        long work = n_planes * ((hs + rows_per_cta - 1) / rows_per_cta);
        int grid = (int)(work < target * 4 ? work : target * 4);
        if (out_dtype == 0)
            heatmap_planes_kernel<float><<<grid, 256, 0, s>>>(kp, n_planes, hs, c, (float *)out, rows_per_cta);
        else
            heatmap_planes_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(kp, n_planes, hs, c, (__nv_bfloat16 *)out,
                                                                     rows_per_cta);
    } else {
        const long total = (long)B * J * hs * hs;
        long blocks = (total + 255) / 256;
        int grid = (int)(blocks < (long)kNumSMs * 32 ? blocks : (long)kNumSMs * 32);
        if (out_dtype == 0)
            heatmap_generic_kernel<float><<<grid, 256, 0, s>>>(kp, B, J, hs, c, (float *)out, out_layout, c_stride, c_offset);
        else
            heatmap_generic_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(kp, B, J, hs, c, (__nv_bfloat16 *)out, out_layout,
                                                                      c_stride, c_offset);
    }
    return launch_status();
}
