// augment.cu -- batched, fused PoseAugmentor for B200.
// Reference semantics: src/dataset/augmentation.py:182-351 (PoseAugmentor.__call__), whose pixel
// arithmetic is torchvision.transforms.functional (PIL backend) + Pillow libImaging:
//   flip (:222-238) -> rotate BILINEAR / depth NEAREST (:241-263) -> antialiased 2-pass BILINEAR
//   resize / depth NEAREST (:266-296) -> NEAREST integer-shift translate (:299-325) -> brightness,
//   contrast (:328-339) -> u8 / 255 (:342-349), with the 2-D key-points re-projected from the
//   Y-rotated 3-D joints in fp64.
// Every stage re-quantises to uint8, so the chain is restated stage by stage and kept ON CHIP:
//
//   pack kernel    fp32 (or u8) planar RGB + depth  ->  uchar4 RGBD [B,H,W]     (4 B/px)
//   table kernel   per sample: resize coefficient/bounds tables and the fp64-accumulated nearest
//                  index tables (sequential by definition in libImaging's ImagingScaleAffine)
//   fused kernel   one thread-block CLUSTER per sample, one row band per CTA:
//                    A  rotated band   (bilinear, fp64 coordinates)        -> smem u8
//                    B  horizontal resize pass (22-bit fixed point)          -> smem u8
//                    C  vertical pass + translate gather + brightness        -> smem u8, grey sum
//                    D  grey mean across the cluster through distributed shared memory
//                    E  contrast, /255, zero padding -> fp32 planar output, 128-bit streaming stores
//                  plus the key-point / joint / camera update in fp64 on one warp.
//
// Compiled with -fmad=false: libImaging evaluates a*b+c with separate roundings.
// HBM traffic per sample (fp32 in, S x S out): 16 H W read + 16 S^2 written + 4 H W packed
// round trip (L2 resident for the benchmark batch).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace pose {

constexpr int kAugThreads = 512;
constexpr int kAugCluster = 8;
constexpr int kAugMaxKsize = 9;       // antialias taps: scale >= 0.25
constexpr int kPrecisionBits = 22;    // Resample.c PRECISION_BITS = 32 - 8 - 2

struct alignas(8) AugPlan {
    double rot[6];          // PIL rotate matrix (output -> input), fp64
    double cos_y, sin_y;    // R_y for the 3-D joints
    double scale;           // scale factor
    double tx, ty;          // translation in pixels
    double tr_a2, tr_a5;    // TF.affine inverse matrix offsets
    double rs_ax, rs_ay;    // nearest resize steps (in / out)
    int32_t fix[6];         // 16.16 fixed-point matrix for the nearest rotate
    int32_t rot_mode;       // 0 affine, 1 copy, 2 ROTATE_90, 3 ROTATE_180, 4 ROTATE_270
    int32_t flip;
    int32_t oH, oW;         // output size (what TF.resize produces)
    int32_t nW, nH;         // the reference's new_size tuple (key-point normalisation)
    int32_t need_h, need_v; // resize passes Pillow actually runs
    float bright, contrast; // Image.blend alpha (C float)
    int32_t bright_ex, contrast_ex;
    int32_t pad_[16];
};
static_assert(sizeof(AugPlan) == POSE_AUG_PLAN_BYTES, "AugPlan must match POSE_AUG_PLAN_BYTES");

// ---------------------------------------------------------------------------------------------
// host: plan
// ---------------------------------------------------------------------------------------------
static double py_round15(double x) {  // Python round(x, 15)
    char buf[64];
    snprintf(buf, sizeof buf, "%.15f", x);
    return strtod(buf, nullptr);
}
static double py_mod(double v, double w) {
    double m = fmod(v, w);
    if (m != 0.0 && ((w < 0.0) != (m < 0.0))) m += w;
    return m;
}
static int pil_floor_h(double v) { return v >= 0.0 ? (int)v : (int)floor(v); }

static void make_plan(const double *p, int H, int W, int flags, AugPlan &pl) {
    memset(&pl, 0, sizeof pl);
    pl.flip = (flags & POSE_AUG_FLIP) && p[0] != 0.0;
    pl.rot_mode = 1;
    pl.cos_y = 1.0;
    if (flags & POSE_AUG_ROTATE) {
        const double ang_in = p[1];
        const double rad = ang_in * (M_PI / 180.0);
        pl.cos_y = cos(rad);
        pl.sin_y = sin(rad);
        const double ang = py_mod(ang_in, 360.0);
        if (ang == 0.0) pl.rot_mode = 1;
        else if (ang == 180.0) pl.rot_mode = 3;
        else if ((ang == 90.0 || ang == 270.0) && W == H) pl.rot_mode = ang == 90.0 ? 2 : 4;
        else {
            pl.rot_mode = 0;
            const double cx = W / 2.0, cy = H / 2.0, r = -(ang * (M_PI / 180.0));
            double a[6] = {py_round15(cos(r)), py_round15(sin(r)), 0.0, py_round15(-sin(r)), py_round15(cos(r)), 0.0};
            const double x = -cx - 0.0, y = -cy - 0.0;
            const double m2 = a[0] * x + a[1] * y + a[2], m5 = a[3] * x + a[4] * y + a[5];
            a[2] = m2 + cx;
            a[5] = m5 + cy;
            for (int i = 0; i < 6; ++i) pl.rot[i] = a[i];
#define FIX(v) pil_floor_h((v) * 65536.0 + 0.5)
            pl.fix[0] = FIX(a[0]);
            pl.fix[1] = FIX(a[1]);
            pl.fix[3] = FIX(a[3]);
            pl.fix[4] = FIX(a[4]);
            pl.fix[2] = FIX(a[2] + a[0] * 0.5 + a[1] * 0.5);
            pl.fix[5] = FIX(a[5] + a[3] * 0.5 + a[4] * 0.5);
#undef FIX
        }
    }
    pl.oH = H;
    pl.oW = W;
    pl.nW = W;
    pl.nH = H;
    pl.scale = 1.0;
    pl.rs_ax = pl.rs_ay = 1.0;
    if (flags & POSE_AUG_SCALE) {
        pl.scale = p[2];
        pl.nW = (int)((double)W * p[2]);
        pl.nH = (int)((double)H * p[2]);
        pl.oH = pl.nW;  // TF.resize reads the (w, h) tuple as (h, w): augmentation.py:270-279
        pl.oW = pl.nH;
        pl.need_h = pl.oW != W;
        pl.need_v = pl.oH != H;
        if (pl.need_h || pl.need_v) {
            pl.rs_ax = (double)((float)W - 0.0f) / pl.oW;
            pl.rs_ay = (double)((float)H - 0.0f) / pl.oH;
        }
    }
    if (flags & POSE_AUG_TRANSLATE) {
        pl.tx = p[3] * (double)pl.oW;
        pl.ty = p[4] * (double)pl.oH;
        const double ccx = pl.oW * 0.5, ccy = pl.oH * 0.5;
        double a2 = 0.0, a5 = 0.0;
        a2 += 1.0 * (-ccx - pl.tx) + 0.0 * (-ccy - pl.ty);
        a5 += -0.0 * (-ccx - pl.tx) + 1.0 * (-ccy - pl.ty);
        a2 += ccx;
        a5 += ccy;
        pl.tr_a2 = a2;
        pl.tr_a5 = a5;
    }
    pl.bright = 1.0f;
    pl.contrast = 1.0f;
    if (flags & POSE_AUG_COLOR) {
        pl.bright = (float)p[5];
        pl.contrast = (float)p[6];
        pl.bright_ex = !(pl.bright >= 0 && pl.bright <= 1.0);
        pl.contrast_ex = !(pl.contrast >= 0 && pl.contrast <= 1.0);
    }
}

struct AugSmemLayout {
    int tab_ints;   // bh + kh + trx + nnx
    int h_px;       // s_h pixels
    int rot_px;     // s_rot / s_out pixels (aliased)
    size_t bytes;
};
static AugSmemLayout smem_layout(int W, const pose_aug_launch &l) {
    AugSmemLayout s;
    s.tab_ints = l.max_out_w * (2 + l.max_ksize + 2) + l.max_band_rows * (4 + l.max_ksize);
    s.tab_ints = (s.tab_ints + 3) & ~3;
    s.h_px = (l.max_rot_rows * l.max_out_w + 3) & ~3;
    int a = l.max_rot_rows * W, b = l.max_band_rows * ((l.max_out_w + 3) & ~3);
    s.rot_px = a > b ? a : b;
    s.bytes = (size_t)s.tab_ints * 4 + (size_t)s.h_px * 4 + (size_t)s.rot_px * 4;
    return s;
}

// ---------------------------------------------------------------------------------------------
// device: pack
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned q8(float v) { return (unsigned)__float2int_rz(v * 255.0f) & 0xffu; }
// uint8 input = the decoded pixel BEFORE the reference's u8 / 255 (chunked_dataset.py) and * 255 -> byte
// (augmentation.py:197) round trip, which is not the identity in fp32 (e.g. 3 -> 2): reproduce it exactly
__device__ __forceinline__ unsigned q8_u8(unsigned v) { return q8(__fdiv_rn((float)v, 255.0f)); }

template <typename InT>
__global__ void __launch_bounds__(256)
aug_pack_kernel(const InT *__restrict__ image, const InT *__restrict__ depth, long n_px_per, int B,
                uchar4 *__restrict__ packed) {
    // 4 pixels per thread: one 128-bit (fp32) or 32-bit (u8) load per plane, one 128-bit store
    const long n4 = n_px_per >> 2, total4 = n4 * B;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long)gridDim.x * blockDim.x) {
        const long b = i / n4, q = i - b * n4;
        uchar4 o[4];
        if constexpr (sizeof(InT) == 4) {
            const float4 r = __ldcs((const float4 *)(image + (b * 3 + 0) * n_px_per) + q);
            const float4 g = __ldcs((const float4 *)(image + (b * 3 + 1) * n_px_per) + q);
            const float4 bl = __ldcs((const float4 *)(image + (b * 3 + 2) * n_px_per) + q);
            const float4 d = __ldcs((const float4 *)(depth + b * n_px_per) + q);
            o[0] = make_uchar4(q8(r.x), q8(g.x), q8(bl.x), q8(d.x));
            o[1] = make_uchar4(q8(r.y), q8(g.y), q8(bl.y), q8(d.y));
            o[2] = make_uchar4(q8(r.z), q8(g.z), q8(bl.z), q8(d.z));
            o[3] = make_uchar4(q8(r.w), q8(g.w), q8(bl.w), q8(d.w));
        } else {
            const uchar4 r = ((const uchar4 *)(image + (b * 3 + 0) * n_px_per))[q];
            const uchar4 g = ((const uchar4 *)(image + (b * 3 + 1) * n_px_per))[q];
            const uchar4 bl = ((const uchar4 *)(image + (b * 3 + 2) * n_px_per))[q];
            const uchar4 d = ((const uchar4 *)(depth + b * n_px_per))[q];
            o[0] = make_uchar4(q8_u8(r.x), q8_u8(g.x), q8_u8(bl.x), q8_u8(d.x));
            o[1] = make_uchar4(q8_u8(r.y), q8_u8(g.y), q8_u8(bl.y), q8_u8(d.y));
            o[2] = make_uchar4(q8_u8(r.z), q8_u8(g.z), q8_u8(bl.z), q8_u8(d.z));
            o[3] = make_uchar4(q8_u8(r.w), q8_u8(g.w), q8_u8(bl.w), q8_u8(d.w));
        }
        ((uint4 *)packed)[i] = *(uint4 *)o;
    }
}

// ---------------------------------------------------------------------------------------------
// device: per-sample tables
// ---------------------------------------------------------------------------------------------
struct AugTables {  // offsets (ints) into one sample's table block
    int bh, kh, bv, kv, nnx, nny, trx, try_, stride;
};
__host__ __device__ inline AugTables table_offsets(int maxW, int maxH, int KS) {
    AugTables t;
    t.bh = 0;
    t.kh = t.bh + maxW * 2;
    t.nnx = t.kh + maxW * KS;
    t.trx = t.nnx + maxW;
    t.bv = t.trx + maxW;
    t.kv = t.bv + maxH * 2;
    t.nny = t.kv + maxH * KS;
    t.try_ = t.nny + maxH;
    t.stride = t.try_ + maxH;
    t.stride = (t.stride + 3) & ~3;
    return t;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for one output index (BILINEAR, box = full)
__device__ void resample_coeff(int xx, int inSize, int outSize, int KS, int *bounds, int *kk) {
    const double scale = (double)inSize / (double)outSize;
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 1.0 * filterscale;
    const double center = 0.0 + ((double)xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > inSize) xmax = inSize;
    xmax -= xmin;
    double k[kAugMaxKsize];
    double ww = 0.0;
#pragma unroll
    for (int x = 0; x < kAugMaxKsize; ++x) {
        double w = 0.0;
        if (x < xmax) {
            double t = ((double)(x + xmin) - center + 0.5) * ss;
            if (t < 0.0) t = -t;
            w = t < 1.0 ? 1.0 - t : 0.0;
            ww += w;
        }
        k[x] = w;
    }
#pragma unroll
    for (int x = 0; x < kAugMaxKsize; ++x) {
        double v = k[x];
        if (x < xmax && ww != 0.0) v = v / ww;
        if (x < KS) kk[x] = v < 0 ? (int)(-0.5 + v * (double)(1 << kPrecisionBits)) : (int)(0.5 + v * (double)(1 << kPrecisionBits));
    }
    bounds[0] = xmin;
    bounds[1] = xmax;
}

__global__ void __launch_bounds__(128)
aug_tables_kernel(const AugPlan *__restrict__ plans, int H, int W, int maxW, int maxH, int KS, int flags,
                  int *__restrict__ tables) {
    const AugPlan &pl = plans[blockIdx.x];
    const AugTables t = table_offsets(maxW, maxH, KS);
    int *T = tables + (size_t)blockIdx.x * t.stride;
    const int oW = pl.oW, oH = pl.oH;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // sequential fp64 accumulations (ImagingScaleAffine): one lane each, on separate warps
    if (lane == 0) {
        if (warp == 0) {  // nearest resize, x
            double xo = 0.0 + pl.rs_ax * 0.5;
            for (int x = 0; x < oW; ++x) {
                int xin = xo < 0.0 ? -1 : (int)xo;
                T[t.nnx + x] = (xin >= 0 && xin < W) ? xin : -1;
                xo += pl.rs_ax;
            }
        } else if (warp == 1) {  // nearest resize, y
            double yo = 0.0 + pl.rs_ay * 0.5;
            for (int y = 0; y < oH; ++y) {
                int yin = yo < 0.0 ? -1 : (int)yo;
                T[t.nny + y] = (yin >= 0 && yin < H) ? yin : -1;
                yo += pl.rs_ay;
            }
        } else if (warp == 2) {  // translate, x
            if (flags & POSE_AUG_TRANSLATE) {
                double xo = pl.tr_a2 + 1.0 * 0.5;
                for (int x = 0; x < oW; ++x) {
                    int xin = xo < 0.0 ? -1 : (int)xo;
                    T[t.trx + x] = (xin >= 0 && xin < oW) ? xin : -1;
                    xo += 1.0;
                }
            } else {
                for (int x = 0; x < oW; ++x) T[t.trx + x] = x;
            }
        } else {  // translate, y
            if (flags & POSE_AUG_TRANSLATE) {
                double yo = pl.tr_a5 + 1.0 * 0.5;
                for (int y = 0; y < oH; ++y) {
                    int yin = yo < 0.0 ? -1 : (int)yo;
                    T[t.try_ + y] = (yin >= 0 && yin < oH) ? yin : -1;
                    yo += 1.0;
                }
            } else {
                for (int y = 0; y < oH; ++y) T[t.try_ + y] = y;
            }
        }
    }
    // antialias coefficient tables: independent per output index
    for (int i = threadIdx.x; i < oW + oH; i += blockDim.x) {
        const bool horiz = i < oW;
        const int xx = horiz ? i : i - oW;
        int *bounds = T + (horiz ? t.bh : t.bv) + xx * 2;
        int *kk = T + (horiz ? t.kh : t.kv) + xx * KS;
        const bool need = horiz ? pl.need_h : pl.need_v;
        if (need) {
            resample_coeff(xx, horiz ? W : H, horiz ? oW : oH, KS, bounds, kk);
        } else {  // identity pass: ((v << 22) + (1 << 21)) >> 22 == v
            bounds[0] = xx;
            bounds[1] = 1;
            kk[0] = 1 << kPrecisionBits;
            for (int k = 1; k < KS; ++k) kk[k] = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// device: fused kernel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ unsigned clip8(int v) {
    v >>= kPrecisionBits;
    return (unsigned)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
// Blend.c ImagingBlend, one channel (used to build the 256-entry brightness / contrast tables)
__device__ __forceinline__ unsigned blend8(int p1, int p2, float alpha, int extrapolate) {
    float tmp = (float)p1 + alpha * (float)(p2 - p1);
    if (!extrapolate) return (unsigned)__float2int_rz(tmp) & 0xffu;
    if (tmp <= 0.0f) return 0u;
    if (tmp >= 255.0f) return 255u;
    return (unsigned)__float2int_rz(tmp);
}

// byte k of a packed pixel as an exact float without the conversion pipe: 0x4B0000bb = 2^23 + bb
__device__ __forceinline__ float byte_f(unsigned w, unsigned sel) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - 8388608.0f;
}

// bilinear_filter32RGB for one channel: fp32 evaluation guarded by an error bound, exact fp64
// re-evaluation only when the truncation could differ (|fp32 - fp64| < 2.5e-4 by construction)
__device__ __forceinline__ unsigned bilerp8(float p0, float p1, float q0, float q1, float dxf, float dyf, double dx,
                                            double dy) {
    const float v1 = fmaf(p1 - p0, dxf, p0);
    const float v2 = fmaf(q1 - q0, dxf, q0);
    const float v = fmaf(v2 - v1, dyf, v1);
    const float t = __fadd_rz(v, 8388608.0f);  // v in [0, 256): truncation == floor, via the 2^23 trick
    const float fl = t - 8388608.0f, fr = v - fl;
    if (fr > 2.5e-4f && fr < 1.0f - 2.5e-4f) return __float_as_uint(t) & 0xffu;
    const double a = (double)p0 + ((double)p1 - (double)p0) * dx;
    const double b = (double)q0 + ((double)q1 - (double)q0) * dx;
    const double r = a + (b - a) * dy;
    return (unsigned)(int)r & 0xffu;
}

__device__ __forceinline__ int src_col(int x, int W, int flip) { return flip ? W - 1 - x : x; }

// source pixel of the rotated image for the exact-multiple-of-90 shortcuts (Geometry.c ImagingRotate*)
__device__ __forceinline__ void transpose_src(int mode, int r, int c, int H, int W, int &sy, int &sx) {
    if (mode == 1) { sy = r; sx = c; }
    else if (mode == 2) { sy = c; sx = W - 1 - r; }
    else if (mode == 3) { sy = H - 1 - r; sx = W - 1 - c; }
    else { sy = H - 1 - c; sx = r; }
}

// work items are (column, group of consecutive rows): a thread keeps its column's constants in registers
// for the rows of the group, and items are dealt round-robin so every phase is balanced for any width.
// Exact i / d for i * d < 2^32 with one multiply-high.
struct FastDiv {
    unsigned m;
    int d;
    __device__ explicit FastDiv(int d_) : m(d_ > 1 ? 0xFFFFFFFFu / (unsigned)d_ + 1u : 0u), d(d_) {}
    __device__ __forceinline__ void divmod(int i, int &q, int &r) const {
        q = d > 1 ? (int)__umulhi((unsigned)i, m) : i;
        r = i - q * d;
    }
};

constexpr double kFloorMagic = 6755399441055744.0;  // 2^52 + 2^51

template <int KS>
__global__ void __launch_bounds__(kAugThreads, 2)
aug_fused_kernel(const uchar4 *__restrict__ packed, const AugPlan *__restrict__ plans, const int *__restrict__ tables,
                 const float *__restrict__ kp_in, const float *__restrict__ joints_in, const double *__restrict__ cam_in,
                 float *__restrict__ image_out, float *__restrict__ depth_out, float *__restrict__ kp_out,
                 float *__restrict__ joints_out, double *__restrict__ cam_out, int32_t *__restrict__ out_hw,
                 int H, int W, int J, int PH, int PW, int maxW, int maxH, int max_rot_rows, int max_band_rows,
                 int tab_ints, int h_px, int flags, int *__restrict__ err_flag) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / CL;
    const int tid = threadIdx.x;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *s_bh = (int *)smem_raw;              // [maxW][2]  horizontal bounds
    int *s_kh = s_bh + maxW * 2;              // [maxW][KS] horizontal coefficients
    int *s_nnx = s_kh + maxW * KS;            // [maxW]     nearest-resize source column
    int *s_trx = s_nnx + maxW;                // [maxW]     translate source column
    int *s_row = s_trx + maxW;                // [max_band_rows][4 + KS]: sy, ry, ymin - rr0, cnt, kv[KS]
    uchar4 *s_h = (uchar4 *)(smem_raw + (size_t)tab_ints * 4);
    uchar4 *s_rot = s_h + h_px;   // phases A-B
    uchar4 *s_out = s_rot;        // phases C-E (the rotated band is dead by then)

    __shared__ AugPlan pl;
    __shared__ int s_rng[4];          // R0, R1 (resized rows), rr0, rr1 (rotated rows)
    __shared__ unsigned s_grey;       // this CTA's grey-level partial sum
    __shared__ unsigned s_total;
    __shared__ unsigned char s_lutb[256];  // brightness: u8 -> u8
    __shared__ float s_lutc[256];          // contrast then /255: u8 -> fp32
    __shared__ float s_lutd[256];          // /255: u8 -> fp32

    if (tid < (int)(sizeof(AugPlan) / 4)) ((int *)&pl)[tid] = ((const int *)(plans + b))[tid];
    if (tid == 0) s_grey = 0u;
    __syncthreads();

    const AugTables t = table_offsets(maxW, maxH, KS);
    const int *T = tables + (size_t)b * t.stride;
    const int oW = pl.oW, oH = pl.oH, flip = pl.flip, mode = pl.rot_mode;
    const int opitch = (oW + 3) & ~3;
    const int rows_per = (oH + CL - 1) / CL;
    const int y0 = min(oH, rank * rows_per), y1 = min(oH, y0 + rows_per);
    const int nband = y1 - y0;
    const uchar4 *src = packed + (size_t)b * H * W;
    const bool color = flags & POSE_AUG_COLOR;

    // tables shared by every row of the band -> smem (one contiguous block: bh, kh, nnx, trx)
    for (int i = tid; i < maxW * (2 + KS + 2); i += kAugThreads) s_bh[i] = __ldg(T + i);
    if (tid < 256) {
        s_lutb[tid] = (unsigned char)(color ? blend8(0, tid, pl.bright, pl.bright_ex) : (unsigned)tid);
        s_lutd[tid] = __fdiv_rn((float)tid, 255.0f);
    }
    // resized-row range needed by this band (translate is a monotone shift), then rotated-row range
    if (tid < 32) {
        int lo = 0x7fffffff, hi = -1;
        for (int y = y0 + tid; y < y1; y += 32) {
            int sy = __ldg(T + t.try_ + y);
            if (sy >= 0) { lo = min(lo, sy); hi = max(hi, sy); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (tid == 0) {
            int rr0 = 0, rr1 = 0;
            if (hi >= 0) {
                rr0 = __ldg(T + t.bv + lo * 2);
                rr1 = __ldg(T + t.bv + hi * 2) + __ldg(T + t.bv + hi * 2 + 1);
            }
            if (rr1 - rr0 > max_rot_rows || nband > max_band_rows) {  // plan/launch mismatch: never write out of bounds
                atomicExch(err_flag, 1);
                rr1 = rr0;
                hi = -1;
            }
            s_rng[0] = lo; s_rng[1] = hi + 1; s_rng[2] = rr0; s_rng[3] = rr1;
        }
    }
    __syncthreads();
    const int R1 = s_rng[1], rr0 = s_rng[2], rr1 = s_rng[3];
    const int n_rot = rr1 - rr0;
    const bool have_rows = R1 > 0;
    // per-row table of the band for phase C
    for (int i = tid; i < nband; i += kAugThreads) {
        int *rt = s_row + i * (4 + KS);
        const int sy = have_rows ? __ldg(T + t.try_ + y0 + i) : -1;
        rt[0] = sy;
        if (sy >= 0) {
            rt[1] = __ldg(T + t.nny + sy);
            rt[2] = __ldg(T + t.bv + sy * 2) - rr0;
            rt[3] = __ldg(T + t.bv + sy * 2 + 1);
#pragma unroll
            for (int q = 0; q < KS; ++q) rt[4 + q] = __ldg(T + t.kv + sy * KS + q);
        }
    }

    // ---- phase A: rotated band (rows rr0..rr1 of the rotated image, all W columns) ----
    if (mode == 0) {
        const double a1 = pl.rot[1], a2 = pl.rot[2], a4 = pl.rot[4], a5 = pl.rot[5];
        constexpr int RG = 2;
        const FastDiv dv(W);
        const int n_items = W * ((n_rot + RG - 1) / RG);
        for (int it = tid; it < n_items; it += kAugThreads) {
            int rq, xo;
            dv.divmod(it, rq, xo);
            const double xs = (double)xo + 0.5;
            const double a0xs = pl.rot[0] * xs, a3xs = pl.rot[3] * xs;
#pragma unroll
            for (int u = 0; u < RG; ++u) {
                const int ry = rq * RG + u;
                if (ry >= n_rot) break;
                const double ys = (double)(rr0 + ry) + 0.5;
                double xin = (a0xs + a1 * ys) + a2;
                double yin = (a3xs + a4 * ys) + a5;
                unsigned o = 0;
                if (!(xin < 0.0 || xin >= (double)W || yin < 0.0 || yin >= (double)H)) {
                    xin -= 0.5;
                    yin -= 0.5;
                    // floor without the conversion pipe: round-down add of 2^52 + 2^51 leaves floor() in the low word
                    const double tx_ = __dadd_rd(xin, kFloorMagic), ty_ = __dadd_rd(yin, kFloorMagic);
                    const int x = __double2loint(tx_), y = __double2loint(ty_);
                    const double dx = xin - (tx_ - kFloorMagic), dy = yin - (ty_ - kFloorMagic);
                    const int x0 = src_col(clampi(x, 0, W - 1), W, flip), x1 = src_col(clampi(x + 1, 0, W - 1), W, flip);
                    const int yc = clampi(y, 0, H - 1);
                    const int yn = (y + 1 >= 0 && y + 1 < H) ? y + 1 : yc;  // v2 = v1 when the lower row is outside
                    const uchar4 *r0p = src + (size_t)yc * W, *r1p = src + (size_t)yn * W;
                    const unsigned p0 = __ldg((const unsigned *)(r0p + x0)), p1 = __ldg((const unsigned *)(r0p + x1));
                    const unsigned q0 = __ldg((const unsigned *)(r1p + x0)), q1 = __ldg((const unsigned *)(r1p + x1));
                    if (((p0 ^ p1) | (p0 ^ q0) | (p0 ^ q1)) << 8 == 0) {
                        o = p0 & 0x00ffffffu;  // flat RGB neighbourhood: exact in any arithmetic
                    } else {
                        const float dxf = (float)dx, dyf = (float)dy;
                        const unsigned r = bilerp8(byte_f(p0, 0x7540), byte_f(p1, 0x7540), byte_f(q0, 0x7540), byte_f(q1, 0x7540), dxf, dyf, dx, dy);
                        const unsigned g = bilerp8(byte_f(p0, 0x7541), byte_f(p1, 0x7541), byte_f(q0, 0x7541), byte_f(q1, 0x7541), dxf, dyf, dx, dy);
                        const unsigned bl = bilerp8(byte_f(p0, 0x7542), byte_f(p1, 0x7542), byte_f(q0, 0x7542), byte_f(q1, 0x7542), dxf, dyf, dx, dy);
                        o = r | (g << 8) | (bl << 16);
                    }
                }
                ((unsigned *)s_rot)[ry * W + xo] = o;
            }
        }
    } else {
        const FastDiv dv(W);
        for (int it = tid; it < n_rot * W; it += kAugThreads) {
            int ry, xo, sy, sx;
            dv.divmod(it, ry, xo);
            transpose_src(mode, rr0 + ry, xo, H, W, sy, sx);
            s_rot[it] = __ldg(src + (size_t)sy * W + src_col(sx, W, flip));
        }
    }
    __syncthreads();

    // ---- phase B: horizontal antialias pass over the band (coefficients of the column in registers) ----
    {
        constexpr int RG = 4;
        const FastDiv dv(oW);
        const int n_items = oW * ((n_rot + RG - 1) / RG);
        for (int it = tid; it < n_items; it += kAugThreads) {
            int rq, xx;
            dv.divmod(it, rq, xx);
            const int xmin = s_bh[xx * 2], cnt = s_bh[xx * 2 + 1];
            int k[KS];
#pragma unroll
            for (int q = 0; q < KS; ++q) k[q] = q < cnt ? s_kh[xx * KS + q] : 0;
#pragma unroll
            for (int u = 0; u < RG; ++u) {
                const int row = rq * RG + u;
                if (row >= n_rot) break;
                const unsigned *in = (const unsigned *)s_rot + row * W + xmin;
                int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
#pragma unroll
                for (int q = 0; q < KS; ++q) {
                    if (q < cnt) {
                        const unsigned p = in[q];
                        s0 += (int)(p & 0xffu) * k[q];
                        s1 += (int)((p >> 8) & 0xffu) * k[q];
                        s2 += (int)((p >> 16) & 0xffu) * k[q];
                    }
                }
                ((unsigned *)s_h)[row * oW + xx] = clip8(s0) | (clip8(s1) << 8) | (clip8(s2) << 16);
            }
        }
    }
    __syncthreads();  // s_rot is dead from here on; s_out aliases it

    // ---- phase C: vertical pass at the translated source position, depth gather, brightness ----
    unsigned grey = 0;
    {
        constexpr int RG = 4;
        const FastDiv dv(opitch);
        const int n_items = opitch * ((nband + RG - 1) / RG);
        for (int it = tid; it < n_items; it += kAugThreads) {
            int rq, x;
            dv.divmod(it, rq, x);
            const int sx = x < oW ? s_trx[x] : -1;
            const int rx = sx >= 0 ? s_nnx[sx] : -1;
            const int cxf = rx * pl.fix[0], cyf = rx * pl.fix[3];
#pragma unroll
            for (int u = 0; u < RG; ++u) {
                const int yy = rq * RG + u;
                if (yy >= nband) break;
                const int *rt = s_row + yy * (4 + KS);
                const int sy = rt[0];
                unsigned o = 0;
                if (sy >= 0 && sx >= 0) {
                    const int ry = rt[1], ymin = rt[2], cnt = rt[3];
                    const unsigned *in = (const unsigned *)s_h + ymin * oW + sx;
                    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
#pragma unroll
                    for (int q = 0; q < KS; ++q) {
                        if (q < cnt) {
                            const unsigned p = in[q * oW];
                            const int kq = rt[4 + q];
                            s0 += (int)(p & 0xffu) * kq;
                            s1 += (int)((p >> 8) & 0xffu) * kq;
                            s2 += (int)((p >> 16) & 0xffu) * kq;
                        }
                    }
                    const unsigned r = s_lutb[clip8(s0)], g = s_lutb[clip8(s1)], bl = s_lutb[clip8(s2)];
                    grey += (19595u * r + 38470u * g + 7471u * bl + 0x8000u) >> 16;
                    // depth: nearest translate -> nearest resize -> nearest rotate (16.16) -> flip
                    unsigned d = 0;
                    if (ry >= 0 && rx >= 0) {
                        int iy, ix;
                        bool ok = true;
                        if (mode == 0) {
                            ix = (pl.fix[2] + ry * pl.fix[1] + cxf) >> 16;
                            iy = (pl.fix[5] + ry * pl.fix[4] + cyf) >> 16;
                            ok = ix >= 0 && ix < W && iy >= 0 && iy < H;
                        } else {
                            transpose_src(mode, ry, rx, H, W, iy, ix);
                        }
                        if (ok) d = __ldg((const unsigned char *)(src + (size_t)iy * W + src_col(ix, W, flip)) + 3);
                    }
                    o = r | (g << 8) | (bl << 16) | (d << 24);
                }
                ((unsigned *)s_out)[yy * opitch + x] = o;
            }
        }
    }
    grey = warp_sum_u32(grey);
    if ((tid & 31) == 0 && grey) atomicAdd(&s_grey, grey);

    // ---- key-points / joints / camera: fp64 on one warp of the cluster's first CTA ----
    if (rank == 0 && tid < 32) {
        const double fx = cam_in[b * 4], fy = cam_in[b * 4 + 1], cx = cam_in[b * 4 + 2], cy = cam_in[b * 4 + 3];
        const double sf = pl.scale;
        const bool do_scale = flags & POSE_AUG_SCALE;
        for (int i = tid; i < J; i += 32) {
            int s = i;  // source joint after the left/right swap
            if (flip && J >= 17) {
                if (i >= 1 && i <= 3) s = i + 3;
                else if (i >= 4 && i <= 6) s = i - 3;
                else if (i >= 11 && i <= 13) s = i + 3;
                else if (i >= 14 && i <= 16) s = i - 3;
            }
            float jx = joints_in[((size_t)b * J + s) * 3], jy = joints_in[((size_t)b * J + s) * 3 + 1],
                  jz = joints_in[((size_t)b * J + s) * 3 + 2];
            float kxf = kp_in[((size_t)b * J + s) * 2], kyf = kp_in[((size_t)b * J + s) * 2 + 1];
            if (flip) {
                jx = -jx;
                kxf = 1.0f - kxf;
            }
            double X = jx, Y = jy, Z = jz, kx = 0.0, ky = 0.0;
            bool j64 = false, k64 = false;
            if (flags & POSE_AUG_ROTATE) {
                const double c = pl.cos_y, s_ = pl.sin_y;
                const double x = jx, y = jy, z = jz;
                X = x * c + y * 0.0 + z * s_;
                Y = x * 0.0 + y * 1.0 + z * 0.0;
                Z = x * (-s_) + y * 0.0 + z * c;
                j64 = true;
                double px = -1.0, py = -1.0;
                if (Z > 0) {
                    px = (X * fx / Z) + cx;
                    py = (Y * fy / Z) + cy;
                }
                kx = px / (double)W;
                ky = py / (double)H;
                k64 = true;
            }
            if (do_scale) {
                const double sfx = fx * sf, sfy = fy * sf, scx = cx * sf, scy = cy * sf;
                double px = -1.0, py = -1.0;
                if (j64) {
                    if (Z > 0) {
                        px = (X * sfx / Z) + scx;
                        py = (Y * sfy / Z) + scy;
                    }
                } else if (jz > 0) {  // fp32 joints: numpy >= 2 keeps the scalar maths in fp32
                    px = (double)((jx * (float)sfx / jz) + (float)scx);
                    py = (double)((jy * (float)sfy / jz) + (float)scy);
                }
                kx = px / (double)pl.nW;
                ky = py / (double)pl.nH;
                k64 = true;
            }
            if (flags & POSE_AUG_TRANSLATE) {
                if (k64) {
                    double ux = kx * (double)oW, uy = ky * (double)oH;
                    ux += pl.tx;
                    uy += pl.ty;
                    kx = ux / (double)oW;
                    ky = uy / (double)oH;
                } else {
                    float ux = kxf * (float)oW, uy = kyf * (float)oH;
                    ux = ux + (float)pl.tx;
                    uy = uy + (float)pl.ty;
                    kxf = ux / (float)oW;
                    kyf = uy / (float)oH;
                }
            }
            joints_out[((size_t)b * J + i) * 3] = j64 ? (float)X : jx;
            joints_out[((size_t)b * J + i) * 3 + 1] = j64 ? (float)Y : jy;
            joints_out[((size_t)b * J + i) * 3 + 2] = j64 ? (float)Z : jz;
            kp_out[((size_t)b * J + i) * 2] = k64 ? (float)kx : kxf;
            kp_out[((size_t)b * J + i) * 2 + 1] = k64 ? (float)ky : kyf;
        }
        if (tid == 0) {
            cam_out[b * 4] = do_scale ? fx * sf : fx;
            cam_out[b * 4 + 1] = do_scale ? fy * sf : fy;
            cam_out[b * 4 + 2] = do_scale ? cx * sf : cx;
            cam_out[b * 4 + 3] = do_scale ? cy * sf : cy;
            out_hw[b * 2] = oH;
            out_hw[b * 2 + 1] = oW;
        }
    }

    // ---- phase D: grey mean over the whole image through distributed shared memory ----
    __syncthreads();
    cluster.sync();
    if (tid == 0) {
        unsigned total = 0;
        for (int r = 0; r < CL; ++r) total += *cluster.map_shared_rank(&s_grey, r);
        s_total = total;
    }
    __syncthreads();
    cluster.sync();  // no CTA may exit while its s_grey can still be read remotely
    if (tid < 256) {
        unsigned v = tid;
        if (color) {
            const double m = (double)s_total / (double)((long)oW * oH);
            v = blend8((int)(m + 0.5), tid, pl.contrast, pl.contrast_ex);
        }
        s_lutc[tid] = __fdiv_rn((float)v, 255.0f);
    }
    __syncthreads();

    // ---- phase E: contrast + /255 by table, write fp32 planes (zero padded to PH x PW) ----
    const size_t plane = (size_t)PH * PW;
    float *oimg = image_out + (size_t)b * 3 * plane;
    float *odep = depth_out + (size_t)b * plane;
    const int PW4 = PW >> 2;
    {
        const FastDiv dv(PW4);
        for (int it = tid; it < nband * PW4; it += kAugThreads) {
            int yy, x4;
            dv.divmod(it, yy, x4);
            const int xb = x4 * 4;
            float4 R = make_float4(0.f, 0.f, 0.f, 0.f), G = R, Bc = R, D = R;
            if (xb < oW) {
                const uint4 p = *(const uint4 *)((const unsigned *)s_out + yy * opitch + xb);
                R.x = s_lutc[p.x & 0xff]; G.x = s_lutc[(p.x >> 8) & 0xff]; Bc.x = s_lutc[(p.x >> 16) & 0xff]; D.x = s_lutd[p.x >> 24];
                if (xb + 1 < oW) { R.y = s_lutc[p.y & 0xff]; G.y = s_lutc[(p.y >> 8) & 0xff]; Bc.y = s_lutc[(p.y >> 16) & 0xff]; D.y = s_lutd[p.y >> 24]; }
                if (xb + 2 < oW) { R.z = s_lutc[p.z & 0xff]; G.z = s_lutc[(p.z >> 8) & 0xff]; Bc.z = s_lutc[(p.z >> 16) & 0xff]; D.z = s_lutd[p.z >> 24]; }
                if (xb + 3 < oW) { R.w = s_lutc[p.w & 0xff]; G.w = s_lutc[(p.w >> 8) & 0xff]; Bc.w = s_lutc[(p.w >> 16) & 0xff]; D.w = s_lutd[p.w >> 24]; }
            }
            const size_t off = (size_t)(y0 + yy) * PW + xb;
            st_stream_f4(oimg + off, R);
            st_stream_f4(oimg + plane + off, G);
            st_stream_f4(oimg + 2 * plane + off, Bc);
            st_stream_f4(odep + off, D);
        }
    }
    // padding rows oH..PH, dealt round-robin to the CTAs of the cluster
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int y = oH + rank; y < PH; y += CL)
        for (int x4 = tid; x4 < PW4; x4 += kAugThreads) {
            const size_t off = (size_t)y * PW + x4 * 4;
            st_stream_f4(oimg + off, z);
            st_stream_f4(oimg + plane + off, z);
            st_stream_f4(oimg + 2 * plane + off, z);
            st_stream_f4(odep + off, z);
        }
}

}  // namespace pose

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
POSE_API int pose_augment_plan(const double *params, int B, int H, int W, int flags, void *plan_out,
                               pose_aug_launch *launch_out) {
    using namespace pose;
    if (!params || !plan_out || !launch_out) return POSE_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0) return POSE_E_SHAPE;
    AugPlan *plans = (AugPlan *)plan_out;
    pose_aug_launch l;
    memset(&l, 0, sizeof l);
    l.cluster = kAugCluster;
    l.max_ksize = 3;
    for (int i = 0; i < B; ++i) {
        const double *p = params + (size_t)i * 8;
        if (flags & POSE_AUG_SCALE) {
            if (!(p[2] > 0.0) || (int)((double)W * p[2]) < 1 || (int)((double)H * p[2]) < 1) return POSE_E_SHAPE;
        }
        make_plan(p, H, W, flags, plans[i]);
        const AugPlan &pl = plans[i];
        if (pl.oH > l.max_out_h) l.max_out_h = pl.oH;
        if (pl.oW > l.max_out_w) l.max_out_w = pl.oW;
        const int rows_per = (pl.oH + kAugCluster - 1) / kAugCluster;
        if (rows_per > l.max_band_rows) l.max_band_rows = rows_per;
        // rotated rows feeding one band: (rows - 1) * scale + 2 * support + 2, rounded up
        double sc_v = pl.need_v ? (double)H / pl.oH : 1.0, sup_v = sc_v < 1.0 ? 1.0 : sc_v;
        int rot_rows = pl.need_v ? (int)ceil((rows_per - 1) * sc_v + 2.0 * sup_v + 2.0) + 1 : rows_per;
        if (rot_rows > H) rot_rows = H;
        if (rot_rows > l.max_rot_rows) l.max_rot_rows = rot_rows;
        double sc_h = pl.need_h ? (double)W / pl.oW : 1.0;
        for (int pass = 0; pass < 2; ++pass) {
            double sc = pass ? sc_v : sc_h, sup = sc < 1.0 ? 1.0 : sc;
            int ks = (int)ceil(sup) * 2 + 1;
            if (ks > l.max_ksize) l.max_ksize = ks;
        }
    }
    if (l.max_ksize > kAugMaxKsize) return POSE_E_UNSUPPORTED;  // scale < 0.25
    AugSmemLayout s = smem_layout(W, l);
    l.smem_bytes = (int32_t)s.bytes;
    if (s.bytes > 227 * 1024 - 2048) return POSE_E_UNSUPPORTED;  // band does not fit one SM's shared memory
    *launch_out = l;
    return POSE_OK;
}

POSE_API size_t pose_augment_workspace_bytes(int B, int H, int W, const pose_aug_launch *launch) {
    using namespace pose;
    if (!launch) return 0;
    AugTables t = table_offsets(launch->max_out_w, launch->max_out_h, launch->max_ksize);
    size_t packed = (size_t)B * H * W * 4;
    packed = (packed + 255) & ~(size_t)255;
    return 256 /* error flag */ + packed + (size_t)B * t.stride * 4;
}

POSE_API int pose_augment_batch(const void *image, const void *depth, int in_dtype, const float *kp,
                                const float *joints, const double *cam, const void *plan,
                                const pose_aug_launch *launch, int B, int H, int W, int J, int flags,
                                float *image_out, float *depth_out, int PH, int PW, float *kp_out, float *joints_out,
                                double *cam_out, int32_t *out_hw, void *workspace, size_t workspace_bytes,
                                pose_stream_t stream) {
    using namespace pose;
    if (!image || !depth || !kp || !joints || !cam || !plan || !launch || !image_out || !depth_out || !kp_out ||
        !joints_out || !cam_out || !out_hw || !workspace)
        return POSE_E_NULL;
    if (B <= 0 || H <= 0 || W <= 0 || J <= 0) return POSE_E_SHAPE;
    if (in_dtype != 0 && in_dtype != 1) return POSE_E_UNSUPPORTED;
    if (((long)H * W) % 4) return POSE_E_UNSUPPORTED;  // pack kernel moves 4 pixels per thread
    if (PW % 4 || PW < launch->max_out_w || PH < launch->max_out_h) return POSE_E_SHAPE;
    if (workspace_bytes < pose_augment_workspace_bytes(B, H, W, launch)) return POSE_E_WORKSPACE;
    if ((uintptr_t)workspace % 256 || (uintptr_t)image % 16 || (uintptr_t)depth % 16 || (uintptr_t)image_out % 16 ||
        (uintptr_t)depth_out % 16)
        return POSE_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    int *err_flag = (int *)workspace;
    uchar4 *packed = (uchar4 *)((char *)workspace + 256);
    size_t packed_bytes = ((size_t)B * H * W * 4 + 255) & ~(size_t)255;
    int *tables = (int *)((char *)packed + packed_bytes);

    const long n_px = (long)H * W, total4 = n_px / 4 * B;
    int grid = (int)((total4 + 255) / 256 < (long)kNumSMs * 16 ? (total4 + 255) / 256 : (long)kNumSMs * 16);
    if (in_dtype == 0)
        aug_pack_kernel<float><<<grid, 256, 0, s>>>((const float *)image, (const float *)depth, n_px, B, packed);
    else
        aug_pack_kernel<unsigned char><<<grid, 256, 0, s>>>((const unsigned char *)image, (const unsigned char *)depth,
                                                            n_px, B, packed);
    aug_tables_kernel<<<B, 128, 0, s>>>((const AugPlan *)plan, H, W, launch->max_out_w, launch->max_out_h,
                                        launch->max_ksize, flags, tables);

    AugSmemLayout sl = smem_layout(W, *launch);
    auto kern = launch->max_ksize == 3 ? aug_fused_kernel<3> : launch->max_ksize == 5 ? aug_fused_kernel<5>
              : launch->max_ksize == 7 ? aug_fused_kernel<7> : aug_fused_kernel<9>;
    if (launch->max_ksize != 3 && launch->max_ksize != 5 && launch->max_ksize != 7 && launch->max_ksize != 9)
        return POSE_E_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.bytes);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)B * launch->cluster);
    cfg.blockDim = dim3(kAugThreads);
    cfg.dynamicSmemBytes = sl.bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = launch->cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const uchar4 *cpacked = packed;
    const AugPlan *cplan = (const AugPlan *)plan;
    const int *ctables = tables;
    int maxW = launch->max_out_w, maxH = launch->max_out_h, KS = launch->max_ksize, mrr = launch->max_rot_rows,
        mbr = launch->max_band_rows;
    (void)KS;
    e = cudaLaunchKernelEx(&cfg, kern, cpacked, cplan, ctables, kp, joints, cam, image_out, depth_out,
                           kp_out, joints_out, cam_out, out_hw, H, W, J, PH, PW, maxW, maxH, mrr, mbr,
                           sl.tab_ints, sl.h_px, flags, err_flag);
    if (e != cudaSuccess) return (int)e;
    return launch_status();
}
