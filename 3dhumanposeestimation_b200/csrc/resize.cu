// resize.cu -- transforms.Resize(image_size) of decoded frames on the device (reference: main.py:171-173
// applied per frame in src/dataset/chunked_dataset.py:100-129, followed by the depth rescale of :159-164).
//
// transforms.Resize on a float tensor is torch.nn.functional.interpolate(mode="bilinear", align_corners=False,
// antialias=True): ATen's separable anti-aliased resampling (UpSampleKernel.cpp): width pass into an fp32 temporary, then
// height pass; triangle filter of support max(scale, 1); per output index i
//     center = scale * (i + 0.5), xmin = max(int(center - support + 0.5), 0), xsize = min(int(center + support + 0.5), in) - xmin
//     w_j = tri((j + xmin - center + 0.5) * invscale), normalised by DIVISION by their fp32 sum
// with the C++ source's float / double promotions (center is an fp32 rounding of a double product, ...).  The accumulation
// follows the shipped x86 build tap by tap (probed against torch 2.11; frozen in tests/golden/resize.npz):
// t = s0 * w0, the following taps in blocks of four with separate multiply / add, the last (n - 1) % 4 taps fused -- so the
// result equals the reference's CPU tensor bit for bit.  uint8 frames are converted as the reference does
// (`.float() / 255.0`, IEEE division) on the fly; the optional per-sample affine is the depth rescale
// `depth * (max - min) + min` (separate multiply and add).
//
// One CTA per (frame, channel, band of output rows): the width pass of the input rows the band needs goes to shared memory,
// the height pass reads it from there -- the fp32 temporary never reaches HBM.
#include "common.cuh"

namespace pose {

struct AaDim {
    int in, out, max_interp;
    float scale, support, invscale;
};

__host__ __device__ inline AaDim aa_dim(int in, int out) {
    AaDim d;
    d.in = in;
    d.out = out;
    d.scale = (float)in / (float)out;                       // area_pixel_compute_scale<float>, align_corners = false
    d.support = d.scale >= 1.0f ? 1.0f * d.scale : 1.0f;     // (interp_size * 0.5) * scale, interp_size = 2
    d.invscale = d.scale >= 1.0f ? (float)(1.0 / (double)d.scale) : 1.0f;
    d.max_interp = (int)ceilf(d.support) * 2 + 1;
    return d;
}

// table entry of one output index: [0] = xmin, [1] = xsize, [2 ..] = weights (as float bits)
__global__ void __launch_bounds__(128)
aa_weights_kernel(AaDim d, int *__restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.out) return;
    int *t = table + (long)i * (2 + d.max_interp);
    const float center = (float)((double)d.scale * ((double)i + 0.5));
    long xmin = (long)((double)center - (double)d.support + 0.5);
    if (xmin < 0) xmin = 0;
    long xmax = (long)((double)center + (double)d.support + 0.5);
    if (xmax > d.in) xmax = d.in;
    long n = xmax - xmin;
    n = n < 0 ? 0 : (n > d.max_interp ? d.max_interp : n);
    float total = 0.0f;
    float *w = (float *)(t + 2);
    for (long j = 0; j < n; ++j) {
        float x = (float)(((double)__fsub_rn((float)(j + xmin), center) + 0.5) * (double)d.invscale);
        x = fabsf(x);
        const float wt = x < 1.0f ? __fsub_rn(1.0f, x) : 0.0f;
        w[j] = wt;
        total = __fadd_rn(total, wt);
    }
    if (total != 0.0f)
        for (long j = 0; j < n; ++j) w[j] = __fdiv_rn(w[j], total);
    for (long j = n; j < d.max_interp; ++j) w[j] = 0.0f;
    t[0] = (int)xmin;
    t[1] = (int)n;
}

template <typename InT>
__device__ __forceinline__ float load_px(const InT *p);
template <>
__device__ __forceinline__ float load_px<float>(const float *p) { return __ldg(p); }
template <>
__device__ __forceinline__ float load_px<unsigned char>(const unsigned char *p) { return __fdiv_rn((float)__ldg(p), 255.0f); }

// ATen's tap order: first tap a plain product, then blocks of four (multiply, add), then the fused tail
#define AA_ACCUMULATE(T_, N_, SRC_, W_)                                  \
    do {                                                                 \
        T_ = (N_) > 0 ? __fmul_rn(SRC_(0), (W_)[0]) : 0.0f;              \
        const int blk_ = 1 + (((N_) - 1) / 4) * 4;                       \
        for (int j_ = 1; j_ < (N_); ++j_)                                \
            T_ = j_ >= blk_ ? __fmaf_rn(SRC_(j_), (W_)[j_], T_) : __fadd_rn(T_, __fmul_rn(SRC_(j_), (W_)[j_])); \
    } while (0)

template <typename InT>
__global__ void __launch_bounds__(256)
resize_aa_kernel(const InT *__restrict__ src, int C, AaDim dw, AaDim dh, const int *__restrict__ tab_w, const int *__restrict__ tab_h,
                 int band, int rows_cap, const float *__restrict__ mul, const float *__restrict__ add, float *__restrict__ dst) {
    extern __shared__ float tmp[];                 // [rows_cap][OW]: width-pass results of the input rows of this band
    const int bands = (dh.out + band - 1) / band;
    const int plane = blockIdx.x / bands, bi = blockIdx.x - plane * bands;      // plane = frame * C + channel
    const int oy0 = bi * band, oy1 = min(dh.out, oy0 + band);
    const int sw = 2 + dw.max_interp, sh = 2 + dh.max_interp;
    // input rows needed by the band (xmin is monotone in the output index)
    const int r0 = __ldg(tab_h + (long)oy0 * sh);
    const int r1 = __ldg(tab_h + (long)(oy1 - 1) * sh) + __ldg(tab_h + (long)(oy1 - 1) * sh + 1);
    const int rows = min(r1 - r0, rows_cap);
    const InT *in = src + (long)plane * dh.in * dw.in;
    const int OW = dw.out;
    // ---- width pass: tmp[r][ox] ----
    for (int e = threadIdx.x; e < rows * OW; e += blockDim.x) {
        const int r = e / OW, ox = e - r * OW;
        const int *t = tab_w + (long)ox * sw;
        const int xmin = __ldg(t), n = __ldg(t + 1);
        const float *w = (const float *)(t + 2);
        const InT *row = in + (long)(r0 + r) * dw.in + xmin;
        float acc;
#define SRC_H(j) load_px<InT>(row + (j))
        AA_ACCUMULATE(acc, n, SRC_H, w);
#undef SRC_H
        tmp[e] = acc;
    }
    __syncthreads();
    // ---- height pass ----
    const int b = plane / C;
    const float m = mul != nullptr ? __ldg(mul + b) : 1.0f, a = add != nullptr ? __ldg(add + b) : 0.0f;
    float *out = dst + (long)plane * dh.out * OW;
    for (int e = threadIdx.x; e < (oy1 - oy0) * OW; e += blockDim.x) {
        const int oyl = e / OW, ox = e - oyl * OW, oy = oy0 + oyl;
        const int *t = tab_h + (long)oy * sh;
        const int ymin = __ldg(t) - r0, n = __ldg(t + 1);
        const float *w = (const float *)(t + 2);
        const float *col = tmp + (long)ymin * OW + ox;
        float acc;
#define SRC_V(j) col[(long)(j) * OW]
        AA_ACCUMULATE(acc, n, SRC_V, w);
#undef SRC_V
        if (mul != nullptr || add != nullptr) acc = __fadd_rn(__fmul_rn(acc, m), a);   // depth * (max - min) + min
        out[(long)oy * OW + ox] = acc;
    }
}

}  // namespace pose

using namespace pose;

static int resize_plan(int H, int W, int OH, int OW, int *band, int *rows_cap, size_t *smem) {
    const AaDim dh = aa_dim(H, OH);
    // a band of output rows needs about band * scale + 2 * support + 2 input rows; the width-pass tile must fit in 200 KB
    int b = 16;
    for (;;) {
        const int rows = (int)ceilf((float)b * dh.scale + 2.0f * dh.support) + 3;
        const size_t bytes = (size_t)rows * OW * sizeof(float);
        if (bytes <= 200 * 1024 || b == 1) {
            *band = b;
            *rows_cap = rows;
            *smem = bytes;
            return bytes <= 227 * 1024 ? POSE_OK : POSE_E_UNSUPPORTED;
        }
        b >>= 1;
    }
}

POSE_API size_t pose_resize_workspace_bytes(int H, int W, int OH, int OW) {
    if (H <= 0 || W <= 0 || OH <= 0 || OW <= 0) return 0;
    const AaDim dw = aa_dim(W, OW), dh = aa_dim(H, OH);
    return ((size_t)OW * (2 + dw.max_interp) + (size_t)OH * (2 + dh.max_interp)) * sizeof(int) + 256;
}

POSE_API int pose_resize_bilinear_aa(const void *src, int in_dtype, int B, int C, int H, int W, int OH, int OW, const float *mul,
                                     const float *add, void *workspace, size_t workspace_bytes, float *dst, pose_stream_t stream) {
    if (!src || !dst || !workspace) return POSE_E_NULL;
    if (B <= 0 || C <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0) return POSE_E_SHAPE;
    if (in_dtype != 0 && in_dtype != 1) return POSE_E_UNSUPPORTED;          // 0: fp32 in [0, 1], 1: uint8
    if (workspace_bytes < pose_resize_workspace_bytes(H, W, OH, OW)) return POSE_E_WORKSPACE;
    if ((uintptr_t)workspace % 16) return POSE_E_ALIGN;
    const AaDim dw = aa_dim(W, OW), dh = aa_dim(H, OH);
    int band, rows_cap;
    size_t smem;
    int e = resize_plan(H, W, OH, OW, &band, &rows_cap, &smem);
    if (e) return e;
    int *tab_w = (int *)workspace, *tab_h = tab_w + (size_t)OW * (2 + dw.max_interp);
    cudaStream_t s = (cudaStream_t)stream;
    aa_weights_kernel<<<(OW + 127) / 128, 128, 0, s>>>(dw, tab_w);
    aa_weights_kernel<<<(OH + 127) / 128, 128, 0, s>>>(dh, tab_h);
    const int bands = (OH + band - 1) / band;
    const long grid = (long)B * C * bands;
    if (grid > 2147483647L) return POSE_E_SHAPE;
    cudaError_t ce;
    if (in_dtype == 0) {
        ce = cudaFuncSetAttribute(resize_aa_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return (int)ce;
        resize_aa_kernel<float><<<(unsigned)grid, 256, smem, s>>>((const float *)src, C, dw, dh, tab_w, tab_h, band, rows_cap, mul, add, dst);
    } else {
        ce = cudaFuncSetAttribute(resize_aa_kernel<unsigned char>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return (int)ce;
        resize_aa_kernel<unsigned char><<<(unsigned)grid, 256, smem, s>>>((const unsigned char *)src, C, dw, dh, tab_w, tab_h, band,
                                                                        rows_cap, mul, add, dst);
    }
    return launch_status();
}
