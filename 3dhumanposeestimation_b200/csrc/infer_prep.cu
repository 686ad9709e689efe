// infer_prep.cu -- model-input preparation of the inference path (reference: infer.py:319-380).
//   depth : F.interpolate(depth_map [B,1,h,w], size = model input, mode="bilinear", align_corners=False)  (infer.py:362-367)
//   kpts  : pixel key-points -> (x / img_w, y / img_h)                                                   (infer.py:217-221)
// The interpolation restates ATen's upsample_bilinear2d (area_pixel_compute_source_index with align_corners = false):
//   scale = in / out (float); src = max(scale * (dst + 0.5) - 0.5, 0); i0 = (int)src; i1 = i0 + (i0 < in - 1);
//   l1 = src - i0; l0 = 1 - l1;  out = h0 * (w0 * v00 + w1 * v01) + h1 * (w0 * v10 + w1 * v11)
// evaluated with separate multiplies / adds (no contraction) so that it agrees with the CPU path to the last bit.
#include "common.cuh"

namespace pose {

__global__ void __launch_bounds__(256)
depth_resize_kernel(const float *__restrict__ src, int B, int h, int w, int H, int W, float sh, float sw,
                    float *__restrict__ dst) {
    const long total = (long)B * H * W;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % W);
        const long q = i / W;
        const int y = (int)(q % H);
        const long b = q / H;
        const float fy = fmaxf(__fsub_rn(__fmul_rn(sh, (float)y + 0.5f), 0.5f), 0.f);
        const float fx = fmaxf(__fsub_rn(__fmul_rn(sw, (float)x + 0.5f), 0.5f), 0.f);
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
        const float h1 = __fsub_rn(fy, (float)y0), h0 = __fsub_rn(1.0f, h1);
        const float w1 = __fsub_rn(fx, (float)x0), w0 = __fsub_rn(1.0f, w1);
        const float *p = src + b * h * w;
        const float v00 = __ldg(p + (long)y0 * w + x0), v01 = __ldg(p + (long)y0 * w + x1);
        const float v10 = __ldg(p + (long)y1 * w + x0), v11 = __ldg(p + (long)y1 * w + x1);
        const float top = __fadd_rn(__fmul_rn(w0, v00), __fmul_rn(w1, v01));
        const float bot = __fadd_rn(__fmul_rn(w0, v10), __fmul_rn(w1, v11));
        dst[i] = __fadd_rn(__fmul_rn(h0, top), __fmul_rn(h1, bot));
    }
}

__global__ void __launch_bounds__(256)
kpts_normalise_kernel(const float *__restrict__ px_conf, long n, float img_w, float img_h, float *__restrict__ kp2,
                      float *__restrict__ kp3) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __fdiv_rn(px_conf[i * 3], img_w), y = __fdiv_rn(px_conf[i * 3 + 1], img_h), c = px_conf[i * 3 + 2];
    kp2[i * 2] = x;
    kp2[i * 2 + 1] = y;
    if (kp3 != nullptr) {
        kp3[i * 3] = x;
        kp3[i * 3 + 1] = y;
        kp3[i * 3 + 2] = c;
    }
}

}  // namespace pose

using namespace pose;

POSE_API int pose_infer_prep(const float *depth, int B, int h, int w, int H, int W, float *depth_out, const float *kpts_px_conf,
                             int K, float img_w, float img_h, float *kp_norm, float *kp_norm_conf, pose_stream_t stream) {
    if (!depth || !depth_out) return POSE_E_NULL;
    if (B <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return POSE_E_SHAPE;
    cudaStream_t s = (cudaStream_t)stream;
    const long total = (long)B * H * W;
    long grid = (total + 255) / 256;
    if (grid > (long)kNumSMs * 16) grid = (long)kNumSMs * 16;
    depth_resize_kernel<<<(unsigned)grid, 256, 0, s>>>(depth, B, h, w, H, W, (float)h / (float)H, (float)w / (float)W, depth_out);
    if (kpts_px_conf != nullptr) {
        if (!kp_norm) return POSE_E_NULL;
        if (K <= 0 || !(img_w > 0.f) || !(img_h > 0.f)) return POSE_E_SHAPE;
        const long n = (long)B * K;
        kpts_normalise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(kpts_px_conf, n, img_w, img_h, kp_norm, kp_norm_conf);
    }
    return launch_status();
}

// ---------------------------------------------------------------------------------------------------------
// Collate (reference: src/dataset/collator.py:10-61 Human36MCollator): variable-sized samples
// image [3, h_b, w_b] / depth [1, h_b, w_b] fp32 -> zero-padded batch [B, 3, Hm, Wm] / [B, 1, Hm, Wm] (F.pad on the right /
// bottom, then torch.stack), one launch for the whole batch through a device table of sample pointers.  Optionally the
// depth is rescaled on the way, depth * (max - min) + min (src/dataset/chunked_dataset.py:159-164).
// ---------------------------------------------------------------------------------------------------------
namespace pose {
struct CollateEntry {
    const float *image, *depth;
    int h, w;
    float dscale, dshift;
};
static_assert(sizeof(CollateEntry) == 32, "CollateEntry must stay 32 bytes (host mirror in dataset/collator.py)");

__global__ void __launch_bounds__(256)
collate_pad_kernel(const CollateEntry *__restrict__ table, int Hm, int Wm, float *__restrict__ image, float *__restrict__ depth) {
    const CollateEntry e = table[blockIdx.y];
    const long plane = (long)Hm * Wm;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < 4 * plane; i += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(i / plane);
        const long p = i - ch * plane;
        const int y = (int)(p / Wm), x = (int)(p - (long)y * Wm);
        const bool in = y < e.h && x < e.w;
        if (ch < 3) {
            image[((long)blockIdx.y * 3 + ch) * plane + p] = in ? __ldg(e.image + ((long)ch * e.h + y) * e.w + x) : 0.f;
        } else {
            // the padding is applied AFTER the rescale in the reference (the dataset rescales, the collator pads): zeros stay zeros
            depth[(long)blockIdx.y * plane + p] = in ? __fadd_rn(__fmul_rn(__ldg(e.depth + (long)y * e.w + x), e.dscale), e.dshift) : 0.f;
        }
    }
}
}  // namespace pose

POSE_API int pose_collate_pad(const void *table, int B, int Hm, int Wm, float *image, float *depth, pose_stream_t stream) {
    if (!table || !image || !depth) return POSE_E_NULL;
    if (B <= 0 || Hm <= 0 || Wm <= 0) return POSE_E_SHAPE;
    long gx = (4L * Hm * Wm + 255) / 256;
    const long cap = ((long)kNumSMs * 16 + B - 1) / B;
    if (gx > cap) gx = cap;
    collate_pad_kernel<<<dim3((unsigned)gx, (unsigned)B), 256, 0, (cudaStream_t)stream>>>((const CollateEntry *)table, Hm, Wm, image, depth);
    return launch_status();
}
