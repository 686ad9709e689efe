/*
 * pose_oracle.c -- CPU restatement (TEST INFRASTRUCTURE ONLY) of the arithmetic on the
 * reference's per-sample hot path.  Nothing in the product package may link, import or
 * execute this file; it exists so tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg have an independent checker for the CUDA kernels.
 *
 * Reference: AliEmreSenel/3DHumanPoseEstimation
 *   A. PoseAugmentor.__call__            src/dataset/augmentation.py:182-351
 *      (pixel arithmetic lives in third-party code that is NOT under /root/reference:
 *       torchvision 0.19.0 transforms.functional + Pillow 11.1.0 libImaging
 *       Geometry.c / Resample.c / Blend.c -- restated here from their published algorithm
 *       and pinned bit-exactly against torchvision 0.26.0 / Pillow 12.2.0 running the
 *       reference's own PoseAugmentor in this container: oracle/gen_golden.py,
 *       tests/test_oracle_golden.py)
 *   B. GaussianHeatmapGenerator.forward  src/models/common.py:23-51
 *   F. ComprehensivePoseLoss.forward     src/loss.py:57-85 (+ :29-55)
 *   metrics compute_mpjpe                src/utils.py:55-69
 *
 * Parity status: PINNED by goldens generated from the live reference (tests/golden/).
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off -- no FMA contraction: Pillow's x86-64
 * wheels evaluate these expressions with separate mul/add).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* small helpers                                                                        */
/* ------------------------------------------------------------------------------------ */

/* libImaging FLOOR(): truncation for v >= 0, floor() below zero */
static inline int pil_floor(double v) { return v >= 0.0 ? (int)v : (int)floor(v); }
/* libImaging COORD(): negative -> -1, else truncation */
static inline int pil_coord(double v) { return v < 0.0 ? -1 : (int)v; }

/* Python round(x, 15): correctly rounded decimal rounding, half-even on the exact binary
 * value.  glibc printf/strtod are both exact, so this reproduces it. */
static double py_round15(double x) {
    char buf[64];
    snprintf(buf, sizeof buf, "%.15f", x);
    return strtod(buf, NULL);
}

/* Python float %: result takes the sign of the divisor */
static double py_mod(double v, double w) {
    double m = fmod(v, w);
    if (m != 0.0 && ((w < 0.0) != (m < 0.0))) m += w;
    return m;
}

/* ------------------------------------------------------------------------------------ */
/* A1. fp32 [0,1] -> uint8 (augmentation.py:196-204; to_pil_image == mul(255).byte())    */
/* ------------------------------------------------------------------------------------ */
ORACLE_API void oracle_quantize_u8(const float *in, long n, uint8_t *out) {
    for (long i = 0; i < n; ++i) {
        float v = in[i] * 255.0f; /* fp32 multiply, then truncation toward zero */
        out[i] = (uint8_t)(int)v;
    }
}

/* ------------------------------------------------------------------------------------ */
/* A2. horizontal flip of an interleaved u8 image (augmentation.py:160-161)             */
/* ------------------------------------------------------------------------------------ */
ORACLE_API void oracle_hflip_u8(const uint8_t *in, int H, int W, int C, uint8_t *out) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < C; ++c) out[((long)y * W + x) * C + c] = in[((long)y * W + (W - 1 - x)) * C + c];
}

/* ------------------------------------------------------------------------------------ */
/* A3. PIL Image.rotate matrix (Image.py rotate(): angle % 360, -radians, round(.,15),   */
/*     centre compensation about (w/2, h/2)).  Returns the transpose mode:              */
/*     0 = general affine in a[6]; 1 = copy; 2 = ROTATE_90; 3 = ROTATE_180; 4 = ROTATE_270 */
/* ------------------------------------------------------------------------------------ */
ORACLE_API int oracle_rotate_matrix(double angle_deg, int W, int H, double a[6]) {
    double ang = py_mod(angle_deg, 360.0);
    if (ang == 0.0) return 1;
    if (ang == 180.0) return 3;
    if ((ang == 90.0 || ang == 270.0) && W == H) return ang == 90.0 ? 2 : 4;
    double cx = W / 2.0, cy = H / 2.0;
    double r = -(ang * (M_PI / 180.0)); /* math.radians(x) == x * (pi/180) */
    a[0] = py_round15(cos(r));
    a[1] = py_round15(sin(r));
    a[2] = 0.0;
    a[3] = py_round15(-sin(r));
    a[4] = py_round15(cos(r));
    a[5] = 0.0;
    /* matrix[2], matrix[5] = transform(-cx, -cy, matrix); then += centre */
    double x = -cx - 0.0, y = -cy - 0.0;
    double m2 = a[0] * x + a[1] * y + a[2];
    double m5 = a[3] * x + a[4] * y + a[5];
    a[2] = m2 + cx;
    a[5] = m5 + cy;
    return 0;
}

/* PIL transposes used for the exact-multiple-of-90 shortcuts (Geometry.c ImagingRotate*) */
static void transpose_u8(const uint8_t *in, int H, int W, int C, int mode, uint8_t *out) {
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            int sy, sx;
            if (mode == 1) { sy = r; sx = c; }
            else if (mode == 2) { sy = c; sx = W - 1 - r; }      /* ROTATE_90  (square only) */
            else if (mode == 3) { sy = H - 1 - r; sx = W - 1 - c; } /* ROTATE_180 */
            else { sy = H - 1 - c; sx = r; }                       /* ROTATE_270 (square only) */
            for (int k = 0; k < C; ++k) out[((long)r * W + c) * C + k] = in[((long)sy * W + sx) * C + k];
        }
}

/* Geometry.c ImagingGenericTransform + affine_transform + bilinear_filter32RGB/8:
 * all coordinate and interpolation arithmetic in fp64, result truncated to u8, fill = 0. */
ORACLE_API void oracle_affine_bilinear_u8(const uint8_t *in, int H, int W, int C, const double a[6],
                                          uint8_t *out) {
    for (int yo = 0; yo < H; ++yo)
        for (int xo = 0; xo < W; ++xo) {
            uint8_t *o = out + ((long)yo * W + xo) * C;
            double xs = xo + 0.5, ys = yo + 0.5;
            double xin = a[0] * xs + a[1] * ys + a[2];
            double yin = a[3] * xs + a[4] * ys + a[5];
            if (xin < 0.0 || xin >= (double)W || yin < 0.0 || yin >= (double)H) {
                for (int k = 0; k < C; ++k) o[k] = 0;
                continue;
            }
            xin -= 0.5;
            yin -= 0.5;
            int x = pil_floor(xin), y = pil_floor(yin);
            double dx = xin - x, dy = yin - y;
            int x0 = x < 0 ? 0 : (x < W ? x : W - 1);
            int x1 = (x + 1) < 0 ? 0 : ((x + 1) < W ? x + 1 : W - 1);
            int yc = y < 0 ? 0 : (y < H ? y : H - 1);
            for (int k = 0; k < C; ++k) {
                const uint8_t *row = in + (long)yc * W * C;
                double p0 = row[x0 * C + k], p1 = row[x1 * C + k];
                double v1 = p0 + (p1 - p0) * dx, v2;
                if (y + 1 >= 0 && y + 1 < H) {
                    const uint8_t *row2 = in + (long)(y + 1) * W * C;
                    double q0 = row2[x0 * C + k], q1 = row2[x1 * C + k];
                    v2 = q0 + (q1 - q0) * dx;
                } else {
                    v2 = v1;
                }
                double v = v1 + (v2 - v1) * dy;
                o[k] = (uint8_t)v;
            }
        }
}

/* Geometry.c affine_fixed: nearest neighbour in 16.16 fixed point, fill = 0 */
ORACLE_API void oracle_affine_nearest_fixed_u8(const uint8_t *in, int H, int W, int C, const double a[6],
                                               uint8_t *out) {
#define FIX(v) pil_floor((v) * 65536.0 + 0.5)
    int a0 = FIX(a[0]), a1 = FIX(a[1]), a3 = FIX(a[3]), a4 = FIX(a[4]);
    int a2 = FIX(a[2] + a[0] * 0.5 + a[1] * 0.5);
    int a5 = FIX(a[5] + a[3] * 0.5 + a[4] * 0.5);
#undef FIX
    memset(out, 0, (size_t)H * W * C);
    for (int y = 0; y < H; ++y) {
        int xx = a2, yy = a5;
        for (int x = 0; x < W; ++x) {
            int xin = xx >> 16;
            if (xin >= 0 && xin < W) {
                int yin = yy >> 16;
                if (yin >= 0 && yin < H)
                    for (int k = 0; k < C; ++k) out[((long)y * W + x) * C + k] = in[((long)yin * W + xin) * C + k];
            }
            xx += a0;
            yy += a3;
        }
        a2 += a1;
        a5 += a4;
    }
}

/* Geometry.c ImagingScaleAffine: nearest neighbour for axis-aligned matrices (a1 == a3 == 0),
 * source coordinate ACCUMULATED in fp64.  Used by resize(NEAREST) and by TF.affine translate. */
ORACLE_API void oracle_scale_affine_nearest_u8(const uint8_t *in, int inH, int inW, int C, int outH, int outW,
                                               const double a[6], uint8_t *out) {
    int *xt = (int *)malloc(sizeof(int) * (size_t)(outW > 0 ? outW : 1));
    double xo = a[2] + a[0] * 0.5, yo = a[5] + a[4] * 0.5;
    for (int x = 0; x < outW; ++x) {
        int xin = pil_coord(xo);
        xt[x] = (xin >= 0 && xin < inW) ? xin : -1;
        xo += a[0];
    }
    memset(out, 0, (size_t)outH * outW * C);
    for (int y = 0; y < outH; ++y) {
        int yi = pil_coord(yo);
        if (yi >= 0 && yi < inH)
            for (int x = 0; x < outW; ++x)
                if (xt[x] >= 0)
                    for (int k = 0; k < C; ++k) out[((long)y * outW + x) * C + k] = in[((long)yi * inW + xt[x]) * C + k];
        yo += a[4];
    }
    free(xt);
}

/* ------------------------------------------------------------------------------------ */
/* A4. Resample.c: antialiased 2-pass BILINEAR resize, 8 bits per channel               */
/* ------------------------------------------------------------------------------------ */
#define PRECISION_BITS (32 - 8 - 2)

static inline double tri_filter(double x) {
    if (x < 0.0) x = -x;
    if (x < 1.0) return 1.0 - x;
    return 0.0;
}

/* precompute_coeffs + normalize_coeffs_8bpc; returns ksize; bounds[2*out], kk[out*ksize] malloc'd */
static int resample_coeffs(int inSize, int outSize, int **boundsp, int32_t **kkp) {
    double scale = (double)((float)inSize - 0.0f) / outSize, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    double support = 1.0 * filterscale;
    int ksize = (int)ceil(support) * 2 + 1;
    int *bounds = (int *)malloc(sizeof(int) * 2 * (size_t)outSize);
    int32_t *kk = (int32_t *)malloc(sizeof(int32_t) * (size_t)outSize * ksize);
    double *k = (double *)malloc(sizeof(double) * (size_t)ksize);
    for (int xx = 0; xx < outSize; ++xx) {
        double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0, ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > inSize) xmax = inSize;
        xmax -= xmin;
        int x;
        for (x = 0; x < xmax; ++x) {
            double w = tri_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        for (; x < ksize; ++x) k[x] = 0.0;
        for (x = 0; x < ksize; ++x)
            kk[(long)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRECISION_BITS))
                                                : (int)(0.5 + k[x] * (1 << PRECISION_BITS));
        bounds[xx * 2] = xmin;
        bounds[xx * 2 + 1] = xmax;
    }
    free(k);
    *boundsp = bounds;
    *kkp = kk;
    return ksize;
}

static inline uint8_t clip8(int v) {
    v >>= PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

ORACLE_API void oracle_resize_bilinear_aa_u8(const uint8_t *in, int inH, int inW, int C, int outH, int outW,
                                             uint8_t *out) {
    int need_h = outW != inW, need_v = outH != inH;
    if (!need_h && !need_v) {
        memcpy(out, in, (size_t)inH * inW * C);
        return;
    }
    int *bh, *bv;
    int32_t *kh, *kv;
    int ksh = resample_coeffs(inW, outW, &bh, &kh);
    int ksv = resample_coeffs(inH, outH, &bv, &kv);
    const uint8_t *src = in;
    uint8_t *tmp = NULL;
    int srcH = inH;
    if (need_h) {
        /* only the rows the vertical pass will touch are produced (ybox_first..ybox_last) */
        int first = bv[0], last = bv[outH * 2 - 2] + bv[outH * 2 - 1];
        for (int i = 0; i < outH; ++i) bv[i * 2] -= first;
        srcH = last - first;
        tmp = (uint8_t *)malloc((size_t)(srcH > 0 ? srcH : 1) * outW * C);
        for (int yy = 0; yy < srcH; ++yy)
            for (int xx = 0; xx < outW; ++xx) {
                int xmin = bh[xx * 2], xmax = bh[xx * 2 + 1];
                const int32_t *k = kh + (long)xx * ksh;
                for (int c = 0; c < C; ++c) {
                    int ss = 1 << (PRECISION_BITS - 1);
                    for (int x = 0; x < xmax; ++x) ss += in[((long)(yy + first) * inW + (x + xmin)) * C + c] * k[x];
                    tmp[((long)yy * outW + xx) * C + c] = clip8(ss);
                }
            }
        src = tmp;
    }
    if (need_v) {
        for (int yy = 0; yy < outH; ++yy) {
            int ymin = bv[yy * 2], ymax = bv[yy * 2 + 1];
            const int32_t *k = kv + (long)yy * ksv;
            for (int xx = 0; xx < outW; ++xx)
                for (int c = 0; c < C; ++c) {
                    int ss = 1 << (PRECISION_BITS - 1);
                    for (int y = 0; y < ymax; ++y) ss += src[((long)(y + ymin) * outW + xx) * C + c] * k[y];
                    out[((long)yy * outW + xx) * C + c] = clip8(ss);
                }
        }
    } else {
        memcpy(out, src, (size_t)srcH * outW * C);
    }
    free(bh); free(bv); free(kh); free(kv); free(tmp);
}

/* ------------------------------------------------------------------------------------ */
/* A6. ImageEnhance.Brightness / Contrast == Image.blend (Blend.c), alpha is a C float  */
/* ------------------------------------------------------------------------------------ */
static inline uint8_t blend_px(int p1, int p2, float alpha, int extrapolate) {
    float t = (float)p1 + alpha * (float)(p2 - p1);
    if (!extrapolate) return (uint8_t)t;
    if (t <= 0.0f) return 0;
    if (t >= 255.0f) return 255;
    return (uint8_t)t;
}

ORACLE_API void oracle_brightness_u8(uint8_t *img, long n, double factor) {
    float alpha = (float)factor;
    int ex = !(alpha >= 0 && alpha <= 1.0);
    for (long i = 0; i < n; ++i) img[i] = blend_px(0, img[i], alpha, ex);
}

/* mean of convert("L"): L = (19595 R + 38470 G + 7471 B + 0x8000) >> 16; m = int(mean + 0.5) */
ORACLE_API int oracle_grey_mean_u8(const uint8_t *rgb, long npix) {
    unsigned long long s = 0;
    for (long i = 0; i < npix; ++i)
        s += (unsigned)((19595u * rgb[i * 3] + 38470u * rgb[i * 3 + 1] + 7471u * rgb[i * 3 + 2] + 0x8000u) >> 16);
    double mean = (double)s / (double)npix;
    return (int)(mean + 0.5);
}

ORACLE_API void oracle_contrast_u8(uint8_t *rgb, long npix, double factor) {
    int m = oracle_grey_mean_u8(rgb, npix);
    float alpha = (float)factor;
    int ex = !(alpha >= 0 && alpha <= 1.0);
    for (long i = 0; i < npix * 3; ++i) rgb[i] = blend_px(m, rgb[i], alpha, ex);
}

/* ------------------------------------------------------------------------------------ */
/* A. the whole PoseAugmentor.__call__ chain with EXPLICIT random draws                 */
/* ------------------------------------------------------------------------------------ */
enum { AUG_FLIP = 1, AUG_ROTATE = 2, AUG_SCALE = 4, AUG_TRANSLATE = 8, AUG_COLOR = 16 };
/* p[0]=flip(0/1, already compared with flip_prob)  p[1]=angle_deg  p[2]=scale
 * p[3]=tx fraction  p[4]=ty fraction  p[5]=brightness  p[6]=contrast  p[7]=unused */

ORACLE_API void oracle_augment_out_size(int H, int W, const double p[8], int flags, int *outH, int *outW) {
    if (flags & AUG_SCALE) {
        /* augmentation.py:270-279 hands (int(w*s), int(h*s)) to TF.resize, which reads it as
         * (h, w): for non-square inputs the output is transposed in shape.  Mirrored as is. */
        *outH = (int)((double)W * p[2]);
        *outW = (int)((double)H * p[2]);
    } else {
        *outW = W;
        *outH = H;
    }
}

static const int SYM_PAIRS[6][2] = {{1, 4}, {2, 5}, {3, 6}, {11, 14}, {12, 15}, {13, 16}};

/* augmentation.py:102-117 + :132-136 in fp64 */
static void project_normalise(const double *j3, int J, double fx, double fy, double cx, double cy, double w,
                              double h, double *kp) {
    for (int i = 0; i < J; ++i) {
        double x = j3[i * 3], y = j3[i * 3 + 1], z = j3[i * 3 + 2], px, py;
        if (z > 0) {
            px = (x * fx / z) + cx;
            py = (y * fy / z) + cy;
        } else {
            px = -1.0;
            py = -1.0;
        }
        kp[i * 2] = px / w;
        kp[i * 2 + 1] = py / h;
    }
}

/*
 * image  fp32 [3,H,W], depth fp32 [H,W], kp fp32 [J,2], joints fp32 [J,3], cam = {fx,fy,cx,cy}.
 * Outputs: image_out fp32 [3,oH,oW], depth_out fp32 [oH,oW] (caller sizes them with
 * oracle_augment_out_size), kp_out fp32 [J,2], joints_out fp32 [J,3], cam_out[4].
 * J must be 17 when AUG_FLIP is set and p[0] != 0 (the symmetric-pair table is H36M's).
 */
ORACLE_API int oracle_augment_sample(const float *image, const float *depth, const float *kp, const float *joints,
                                     const double cam[4], int H, int W, int J, const double p[8], int flags,
                                     float *image_out, float *depth_out, float *kp_out, float *joints_out,
                                     double cam_out[4]) {
    long np_ = (long)H * W;
    uint8_t *rgb = (uint8_t *)malloc((size_t)np_ * 3), *rgb2 = (uint8_t *)malloc((size_t)np_ * 3);
    uint8_t *dep = (uint8_t *)malloc((size_t)np_), *dep2 = (uint8_t *)malloc((size_t)np_);
    /* A1: CHW fp32 -> HWC u8 */
    for (int c = 0; c < 3; ++c)
        for (long i = 0; i < np_; ++i) {
            float v = image[c * np_ + i] * 255.0f;
            rgb[i * 3 + c] = (uint8_t)(int)v;
        }
    for (long i = 0; i < np_; ++i) {
        float v = depth[i] * 255.0f;
        dep[i] = (uint8_t)(int)v;
    }
    /* keypoint state: fp32 until a stage promotes it to fp64 (numpy dtype rules) */
    double *j3 = (double *)malloc(sizeof(double) * J * 3), *k2 = (double *)malloc(sizeof(double) * J * 2);
    float *j3f = (float *)malloc(sizeof(float) * J * 3), *k2f = (float *)malloc(sizeof(float) * J * 2);
    memcpy(j3f, joints, sizeof(float) * J * 3);
    memcpy(k2f, kp, sizeof(float) * J * 2);
    int j_is64 = 0, k_is64 = 0;
    double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
    cam_out[0] = fx; cam_out[1] = fy; cam_out[2] = cx; cam_out[3] = cy;

    /* A2 flip (:222-238, :138-180) -- fp32 arithmetic on fp32 arrays */
    if ((flags & AUG_FLIP) && p[0] != 0.0) {
        oracle_hflip_u8(rgb, H, W, 3, rgb2); { uint8_t *t = rgb; rgb = rgb2; rgb2 = t; }
        oracle_hflip_u8(dep, H, W, 1, dep2); { uint8_t *t = dep; dep = dep2; dep2 = t; }
        for (int i = 0; i < J; ++i) {
            j3f[i * 3] = -j3f[i * 3];
            k2f[i * 2] = 1.0f - k2f[i * 2];
        }
        if (J >= 17)
            for (int s = 0; s < 6; ++s) {
                int l = SYM_PAIRS[s][0], r = SYM_PAIRS[s][1];
                for (int d = 0; d < 3; ++d) { float t = j3f[l * 3 + d]; j3f[l * 3 + d] = j3f[r * 3 + d]; j3f[r * 3 + d] = t; }
                for (int d = 0; d < 2; ++d) { float t = k2f[l * 2 + d]; k2f[l * 2 + d] = k2f[r * 2 + d]; k2f[r * 2 + d] = t; }
            }
    }
    /* A3 rotate (:241-263) */
    if (flags & AUG_ROTATE) {
        double ang = p[1], rad = ang * (M_PI / 180.0), c = cos(rad), s = sin(rad);
        /* joints @ R_y.T with R_y = [[c,0,s],[0,1,0],[-s,0,c]]  (fp64) */
        for (int i = 0; i < J; ++i) {
            double x = j3f[i * 3], y = j3f[i * 3 + 1], z = j3f[i * 3 + 2];
            j3[i * 3] = x * c + y * 0.0 + z * s;
            j3[i * 3 + 1] = x * 0.0 + y * 1.0 + z * 0.0;
            j3[i * 3 + 2] = x * (-s) + y * 0.0 + z * c;
        }
        j_is64 = 1;
        double a[6];
        int mode = oracle_rotate_matrix(ang, W, H, a);
        if (mode == 0) {
            oracle_affine_bilinear_u8(rgb, H, W, 3, a, rgb2);
            oracle_affine_nearest_fixed_u8(dep, H, W, 1, a, dep2);
        } else {
            transpose_u8(rgb, H, W, 3, mode, rgb2);
            transpose_u8(dep, H, W, 1, mode, dep2);
        }
        { uint8_t *t = rgb; rgb = rgb2; rgb2 = t; t = dep; dep = dep2; dep2 = t; }
        project_normalise(j3, J, fx, fy, cx, cy, (double)W, (double)H, k2);
        k_is64 = 1;
    }
    int oW = W, oH = H;
    /* A4 scale (:266-296) */
    if (flags & AUG_SCALE) {
        double sf = p[2];
        int nW = (int)((double)W * sf), nH = (int)((double)H * sf); /* the reference's new_size */
        oH = nW; oW = nH;                                             /* what TF.resize produces */
        uint8_t *r3 = (uint8_t *)malloc((size_t)oW * oH * 3 + 1), *d3 = (uint8_t *)malloc((size_t)oW * oH + 1);
        oracle_resize_bilinear_aa_u8(rgb, H, W, 3, oH, oW, r3);
        if (oW == W && oH == H) {
            memcpy(d3, dep, (size_t)np_); /* Image.resize returns a copy when the size is unchanged */
        } else {
            double a[6] = {(double)((float)W - 0.0f) / oW, 0, 0, 0, (double)((float)H - 0.0f) / oH, 0};
            oracle_scale_affine_nearest_u8(dep, H, W, 1, oH, oW, a, d3);
        }
        free(rgb); free(dep); free(rgb2); free(dep2);
        rgb = r3; dep = d3;
        rgb2 = (uint8_t *)malloc((size_t)oW * oH * 3 + 1);
        dep2 = (uint8_t *)malloc((size_t)oW * oH + 1);
        double sfx = fx * sf, sfy = fy * sf, scx = cx * sf, scy = cy * sf;
        if (!j_is64) { /* fp32 joints promoted element-wise by the python loop (numpy>=2: fp32 scalar maths) */
            for (int i = 0; i < J; ++i) {
                float x = j3f[i * 3], y = j3f[i * 3 + 1], z = j3f[i * 3 + 2];
                double px, py;
                if (z > 0) {
                    px = (double)((x * (float)sfx / z) + (float)scx);
                    py = (double)((y * (float)sfy / z) + (float)scy);
                } else { px = -1.0; py = -1.0; }
                k2[i * 2] = px / (double)nW;
                k2[i * 2 + 1] = py / (double)nH;
            }
        } else {
            project_normalise(j3, J, sfx, sfy, scx, scy, (double)nW, (double)nH, k2);
        }
        k_is64 = 1;
        cam_out[0] = sfx; cam_out[1] = sfy; cam_out[2] = scx; cam_out[3] = scy;
    }
    /* A5 translate (:299-325): TF.affine(angle=0, translate) -> NEAREST, ImagingScaleAffine */
    if (flags & AUG_TRANSLATE) {
        double tx = p[3] * (double)oW, ty = p[4] * (double)oH;
        double ccx = oW * 0.5, ccy = oH * 0.5;
        double a[6] = {1.0, 0.0, 0.0, -0.0, 1.0, 0.0};
        a[2] += a[0] * (-ccx - tx) + a[1] * (-ccy - ty);
        a[5] += a[3] * (-ccx - tx) + a[4] * (-ccy - ty);
        a[2] += ccx;
        a[5] += ccy;
        oracle_scale_affine_nearest_u8(rgb, oH, oW, 3, oH, oW, a, rgb2);
        oracle_scale_affine_nearest_u8(dep, oH, oW, 1, oH, oW, a, dep2);
        { uint8_t *t = rgb; rgb = rgb2; rgb2 = t; t = dep; dep = dep2; dep2 = t; }
        if (k_is64) {
            for (int i = 0; i < J; ++i) {
                double ux = k2[i * 2] * (double)oW, uy = k2[i * 2 + 1] * (double)oH;
                ux += tx; uy += ty;
                k2[i * 2] = ux / (double)oW;
                k2[i * 2 + 1] = uy / (double)oH;
            }
        } else { /* fp32 array; tx/ty are Python floats (weak scalars) so every op stays fp32 */
            for (int i = 0; i < J; ++i) {
                float ux = k2f[i * 2] * (float)oW, uy = k2f[i * 2 + 1] * (float)oH;
                ux = ux + (float)tx; uy = uy + (float)ty;
                k2f[i * 2] = ux / (float)oW;
                k2f[i * 2 + 1] = uy / (float)oH;
            }
        }
    }
    /* A6 colour (:328-339), RGB only */
    if (flags & AUG_COLOR) {
        oracle_brightness_u8(rgb, (long)oW * oH * 3, p[5]);
        oracle_contrast_u8(rgb, (long)oW * oH, p[6]);
    }
    /* A7 back to tensors (:342-349) */
    long onp = (long)oW * oH;
    for (int c = 0; c < 3; ++c)
        for (long i = 0; i < onp; ++i) image_out[c * onp + i] = (float)rgb[i * 3 + c] / 255.0f;
    for (long i = 0; i < onp; ++i) depth_out[i] = (float)dep[i] / 255.0f;
    for (int i = 0; i < J * 3; ++i) joints_out[i] = j_is64 ? (float)j3[i] : j3f[i];
    for (int i = 0; i < J * 2; ++i) kp_out[i] = k_is64 ? (float)k2[i] : k2f[i];
    free(rgb); free(rgb2); free(dep); free(dep2); free(j3); free(k2); free(j3f); free(k2f);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* B. GaussianHeatmapGenerator.forward (common.py:23-51), fp32 in the reference's op order */
/* ------------------------------------------------------------------------------------ */
ORACLE_API void oracle_heatmap(const float *kp, int B, int J, int hs, float sigma, float *out) {
    float scale = (float)(hs - 1);
    float denom = (float)(2.0 * (double)sigma * (double)sigma);
    for (long bj = 0; bj < (long)B * J; ++bj) {
        float kx = kp[bj * 2], ky = kp[bj * 2 + 1];
        float mux = kx * scale, muy = ky * scale;
        float valid = (kx > 0.0f && ky > 0.0f) ? 1.0f : 0.0f; /* strict >, NaN -> invalid */
        float *o = out + bj * (long)hs * hs;
        for (int y = 0; y < hs; ++y) {
            float dy = (float)y - muy;
            float dy2 = dy * dy;
            for (int x = 0; x < hs; ++x) {
                float dx = (float)x - mux;
                float d2 = dx * dx + dy2;
                float v = expf(-d2 / denom);
                o[(long)y * hs + x] = v * valid; /* NaN * 0 = NaN, as in the reference */
            }
        }
    }
}

/* argmax of the squared distance (first index on ties == torch.argmax of the rendered plane
 * wherever exp() is injective); -1 for an invalid keypoint (all-zero plane) */
ORACLE_API void oracle_heatmap_peak(const float *kp, int B, int J, int hs, int *peak) {
    float scale = (float)(hs - 1);
    for (long bj = 0; bj < (long)B * J; ++bj) {
        float kx = kp[bj * 2], ky = kp[bj * 2 + 1];
        if (!(kx > 0.0f && ky > 0.0f)) { peak[bj] = -1; continue; }
        float mux = kx * scale, muy = ky * scale;
        float best = INFINITY; int bi = 0;
        for (int y = 0; y < hs; ++y) {
            float dy = (float)y - muy, dy2 = dy * dy;
            for (int x = 0; x < hs; ++x) {
                float dx = (float)x - mux;
                float d2 = dx * dx + dy2;
                if (d2 < best) { best = d2; bi = y * hs + x; }
            }
        }
        peak[bj] = bi;
    }
}

/* ------------------------------------------------------------------------------------ */
/* F. ComprehensivePoseLoss.forward (loss.py:57-85) + analytic gradient of `total`      */
/*    w = {mse_weight, l1_weight, inter_joint_weight, abs_root_weight}                  */
/*    out5 = {mse, l1, inter_joint, abs_root, total}; grad (may be NULL) = d total/d pred */
/* ------------------------------------------------------------------------------------ */
ORACLE_API void oracle_pose_loss(const float *pred, const float *gt, int B, int J, const float w[4], float out5[5],
                                 float *grad) {
    double s_mse = 0, s_l1 = 0, s_ij = 0, s_root = 0;
    long n = (long)B * J * 3, npairs = (long)J * (J - 1) / 2;
    double c_mse = (double)w[0] * 2.0 / (double)n, c_l1 = (double)w[1] / (double)n;
    double c_ij = npairs ? (double)w[2] / ((double)B * npairs) : 0.0, c_root = (double)w[3] / ((double)B * 3);
    for (int b = 0; b < B; ++b) {
        const float *p = pred + (long)b * J * 3, *g = gt + (long)b * J * 3;
        float *gr = grad ? grad + (long)b * J * 3 : NULL;
        for (int i = 0; i < J * 3; ++i) {
            float d = p[i] - g[i];
            s_mse += (double)(d * d);
            s_l1 += fabsf(d);
            if (i < 3) s_root += fabsf(d);
            if (gr) {
                double sg = (d > 0) - (d < 0);
                gr[i] = (float)(c_mse * d + c_l1 * sg + (i < 3 ? c_root * sg : 0.0));
            }
        }
        for (int i = 0; i < J; ++i)
            for (int j = i + 1; j < J; ++j) {
                float px = p[i * 3] - p[j * 3], py = p[i * 3 + 1] - p[j * 3 + 1], pz = p[i * 3 + 2] - p[j * 3 + 2];
                float gx = g[i * 3] - g[j * 3], gy = g[i * 3 + 1] - g[j * 3 + 1], gz = g[i * 3 + 2] - g[j * 3 + 2];
                float dp = sqrtf(px * px + py * py + pz * pz), dg = sqrtf(gx * gx + gy * gy + gz * gz);
                float e = dp - dg;
                s_ij += fabsf(e);
                if (gr && dp > 0.0f) {
                    double sg = ((e > 0) - (e < 0)) * c_ij / dp;
                    gr[i * 3] += (float)(sg * px); gr[i * 3 + 1] += (float)(sg * py); gr[i * 3 + 2] += (float)(sg * pz);
                    gr[j * 3] -= (float)(sg * px); gr[j * 3 + 1] -= (float)(sg * py); gr[j * 3 + 2] -= (float)(sg * pz);
                }
            }
    }
    float mse = (float)(s_mse / (double)n), l1 = (float)(s_l1 / (double)n);
    float ij = npairs ? (float)(s_ij / ((double)B * npairs)) : NAN; /* mean of an empty tensor is NaN in torch */
    float root = (float)(s_root / ((double)B * 3));
    out5[0] = mse; out5[1] = l1; out5[2] = ij; out5[3] = root;
    out5[4] = w[0] * mse + w[1] * l1 + w[2] * ij + w[3] * root;
}

/* utils.py:55-69 compute_mpjpe */
ORACLE_API float oracle_mpjpe(const float *pred, const float *gt, int B, int J) {
    double s = 0;
    for (long i = 0; i < (long)B * J; ++i) {
        float dx = pred[i * 3] - gt[i * 3], dy = pred[i * 3 + 1] - gt[i * 3 + 1], dz = pred[i * 3 + 2] - gt[i * 3 + 2];
        s += sqrtf(dx * dx + dy * dy + dz * dz);
    }
    return (float)(s / ((double)B * J));
}

/* utils.py:72-165 compute_pa_mpjpe: per sample centre both point sets, M = Pc^T Gc, SVD M = U S V^T, R = V U^T (the
 * reference applies it as Pc @ R, i.e. the transpose of the optimal Procrustes rotation: mirrored, not fixed), reflection
 * fix on the last right-singular vector, scale = sum(S') / |Pc|^2 (1 if |Pc|^2 <= 1e-9), error = mean |s Pc R + mu_g - G|.
 * The 3x3 SVD is a cyclic Jacobi eigen-decomposition of M^T M in double precision (torch.linalg.svd = LAPACK gesdd). */
static void jacobi_eig3(double a[3][3], double v[3][3]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v[i][j] = i == j;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(a[p][q]) < 1e-300) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - sn * akq;
                    a[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - sn * aqk;
                    a[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - sn * vkq;
                    v[k][q] = sn * vkp + c * vkq;
                }
            }
    }
}

static void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}

static double det3(double m[3][3]) {
    return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
           m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

ORACLE_API float oracle_pa_mpjpe(const float *pred, const float *gt, int B, int J, float *per_sample) {
    double total = 0;
    for (int b = 0; b < B; ++b) {
        const float *P = pred + (long)b * J * 3, *G = gt + (long)b * J * 3;
        double mp[3] = {0, 0, 0}, mg[3] = {0, 0, 0};
        for (int j = 0; j < J; ++j) for (int d = 0; d < 3; ++d) { mp[d] += P[j * 3 + d]; mg[d] += G[j * 3 + d]; }
        for (int d = 0; d < 3; ++d) { mp[d] = (float)(mp[d] / J); mg[d] = (float)(mg[d] / J); }
        double M[3][3] = {{0}}, var = 0;
        for (int j = 0; j < J; ++j)
            for (int r = 0; r < 3; ++r) {
                double pc = (float)(P[j * 3 + r] - mp[r]);
                var += pc * pc;
                for (int c = 0; c < 3; ++c) M[r][c] += pc * (double)(float)(G[j * 3 + c] - mg[c]);
            }
        double A[3][3], V[3][3];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { A[r][c] = 0; for (int k = 0; k < 3; ++k) A[r][c] += M[k][r] * M[k][c]; }
        jacobi_eig3(A, V);
        int idx[3] = {0, 1, 2};                                  /* sort eigenvalues descending */
        for (int i = 0; i < 2; ++i) for (int k = i + 1; k < 3; ++k) if (A[idx[k]][idx[k]] > A[idx[i]][idx[i]]) { int t = idx[i]; idx[i] = idx[k]; idx[k] = t; }
        double S[3], Vs[3][3], U[3][3];
        for (int i = 0; i < 3; ++i) {
            S[i] = sqrt(fmax(A[idx[i]][idx[i]], 0.0));
            for (int r = 0; r < 3; ++r) Vs[r][i] = V[r][idx[i]];
        }
        double vc[3]; { double c0[3] = {Vs[0][0], Vs[1][0], Vs[2][0]}, c1[3] = {Vs[0][1], Vs[1][1], Vs[2][1]}; cross3(c0, c1, vc); }
        if (vc[0] * Vs[0][2] + vc[1] * Vs[1][2] + vc[2] * Vs[2][2] < 0) for (int r = 0; r < 3; ++r) Vs[r][2] = -Vs[r][2];   /* det V = +1 */
        const double tiny = 1e-12 * (S[0] > 0 ? S[0] : 1.0);
        int rank = 0;
        for (int i = 0; i < 3; ++i) {
            if (S[i] > tiny) {
                for (int r = 0; r < 3; ++r) { U[r][i] = 0; for (int k = 0; k < 3; ++k) U[r][i] += M[r][k] * Vs[k][i]; U[r][i] /= S[i]; }
                rank = i + 1;
            }
        }
        if (rank == 0) { for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) U[r][c] = Vs[r][c]; }   /* M = 0: any U; R = I */
        else if (rank == 1) {                                   /* complete an orthonormal basis */
            double u0[3] = {U[0][0], U[1][0], U[2][0]}, e[3] = {0, 0, 0}, u1[3], u2[3];
            int m = fabs(u0[0]) < fabs(u0[1]) ? (fabs(u0[0]) < fabs(u0[2]) ? 0 : 2) : (fabs(u0[1]) < fabs(u0[2]) ? 1 : 2);
            e[m] = 1; cross3(u0, e, u1);
            double n = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
            for (int r = 0; r < 3; ++r) u1[r] /= n;
            cross3(u0, u1, u2);
            for (int r = 0; r < 3; ++r) { U[r][1] = u1[r]; U[r][2] = u2[r]; }
        } else if (rank == 2) {
            double u0[3] = {U[0][0], U[1][0], U[2][0]}, u1[3] = {U[0][1], U[1][1], U[2][1]}, u2[3];
            cross3(u0, u1, u2);
            for (int r = 0; r < 3; ++r) U[r][2] = u2[r];
        }
        double R[3][3];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) { R[r][c] = 0; for (int k = 0; k < 3; ++k) R[r][c] += Vs[r][k] * U[c][k]; }
        double s_sum = S[0] + S[1] + S[2];
        if (det3(R) < 0) {
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) R[r][c] -= 2.0 * Vs[r][2] * U[c][2];
            s_sum = S[0] + S[1] - S[2];
        }
        const double sc = var > 1e-9 ? s_sum / var : 1.0;
        double err = 0;
        for (int j = 0; j < J; ++j) {
            double d2 = 0;
            for (int c = 0; c < 3; ++c) {
                double a = 0;
                for (int k = 0; k < 3; ++k) a += (double)(float)(P[j * 3 + k] - mp[k]) * R[k][c];
                double diff = sc * a + mg[c] - G[j * 3 + c];
                d2 += diff * diff;
            }
            err += sqrt(d2);
        }
        err /= J;
        if (per_sample) per_sample[b] = (float)err;
        total += err;
    }
    return (float)(total / B);
}

/* ---------------------------------------------------------------------------------------------
 * transforms.Resize(image_size) on a float32 [C, H, W] tensor (reference: main.py:171-173 applied in
 * src/dataset/chunked_dataset.py:100-129) = torch.nn.functional.interpolate(mode="bilinear", align_corners=False,
 * antialias=True): ATen's separable anti-aliased resampling on the CPU (UpSampleKernel.cpp,
 * _compute_indices_min_size_weights_aa + basic_loop_aa_horizontal / _vertical; torch is a third-party dependency, not
 * under /root/reference: restated from its published algorithm and pinned by tests/golden/resize.npz from the live call).
 * Width pass first into an fp32 temporary [C, H, OW], then the height pass.  Triangle filter of support max(scale, 1),
 * weights normalised by their sum, fp32 arithmetic with the mixed float / double promotions of the C++ source.
 * Accumulation order of the shipped x86 build (probed tap by tap against torch 2.11, see oracle/gen_golden.py::gen_resize):
 * t = s0 * w0, then the remaining taps in order, in blocks of four with separate multiply and add, and the last (n - 1) % 4
 * taps with a fused multiply-add (the compiler's unrolled body vs. its contracted scalar tail).  `fma` = 0 / 1 force
 * plain / fused accumulation everywhere (kept for the probe), 2 = the shipped pattern.
 * ------------------------------------------------------------------------------------------- */
static int aa_weights(int i, int in_size, float scale, float support, int max_interp, float *w, int *xmin_out) {
    const float center = (float)((double)scale * ((double)i + 0.5));
    const float invscale = scale >= 1.0f ? (float)(1.0 / (double)scale) : 1.0f;
    long xmin = (long)((double)center - (double)support + 0.5);
    if (xmin < 0) xmin = 0;
    long xmax = (long)((double)center + (double)support + 0.5);
    if (xmax > in_size) xmax = in_size;
    long xsize = xmax - xmin;
    if (xsize < 0) xsize = 0;
    if (xsize > max_interp) xsize = max_interp;
    float total = 0.0f;
    for (long j = 0; j < xsize; ++j) {
        /* (j + xmin - center + 0.5) * invscale: int64 - float -> float, + 0.5 -> double, * float -> double, -> float */
        float x = (float)(((double)((float)(j + xmin) - center) + 0.5) * (double)invscale);
        if (x < 0.0f) x = -x;
        const float wt = x < 1.0f ? 1.0f - x : 0.0f;
        w[j] = wt;
        total += wt;
    }
    if (total != 0.0f)
        for (long j = 0; j < xsize; ++j) w[j] /= total;
    *xmin_out = (int)xmin;
    return (int)xsize;
}

ORACLE_API void oracle_resize_bilinear_aa(const float *src, int C, int H, int W, int OH, int OW, int fma, float *dst) {
    const float sw = (float)W / (float)OW, sh = (float)H / (float)OH;
    const float supw = sw >= 1.0f ? 1.0f * sw : 1.0f, suph = sh >= 1.0f ? 1.0f * sh : 1.0f;
    const int miw = (int)ceilf(supw) * 2 + 1, mih = (int)ceilf(suph) * 2 + 1;
    float *tmp = (float *)malloc((size_t)C * H * OW * sizeof(float));
    float *w = (float *)malloc((size_t)(miw > mih ? miw : mih) * sizeof(float));
    for (int ox = 0; ox < OW; ++ox) {
        int xmin;
        const int n = aa_weights(ox, W, sw, supw, miw, w, &xmin);
        for (int c = 0; c < C; ++c)
            for (int y = 0; y < H; ++y) {
                const float *s = src + ((size_t)c * H + y) * W + xmin;
                float t = n > 0 ? s[0] * w[0] : 0.0f;
                const int blk = fma == 2 ? 1 + ((n - 1) / 4) * 4 : (fma ? 1 : n);      /* taps [1, blk): mul + add */
                for (int j = 1; j < n; ++j) t = j >= blk ? fmaf(s[j], w[j], t) : t + s[j] * w[j];
                tmp[((size_t)c * H + y) * OW + ox] = t;
            }
    }
    for (int oy = 0; oy < OH; ++oy) {
        int ymin;
        const int n = aa_weights(oy, H, sh, suph, mih, w, &ymin);
        for (int c = 0; c < C; ++c)
            for (int x = 0; x < OW; ++x) {
                const float *s = tmp + ((size_t)c * H + ymin) * OW + x;
                float t = n > 0 ? s[0] * w[0] : 0.0f;
                const int blk = fma == 2 ? 1 + ((n - 1) / 4) * 4 : (fma ? 1 : n);
                for (int j = 1; j < n; ++j) t = j >= blk ? fmaf(s[(size_t)j * OW], w[j], t) : t + s[(size_t)j * OW] * w[j];
                dst[((size_t)c * OH + oy) * OW + x] = t;
            }
    }
    free(tmp);
    free(w);
}
