/*
 * oracle/mvs_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Plain-C, CPU, fp32 restatement of the plane-sweep hot path of fdarmon/wild_deep_mvs.
 * Every function cites the reference lines it restates.  Tensor layouts are the
 * reference's own (PyTorch NCHW / NCDHW, one batch element per call) so that the
 * Python wrappers in oracle/__init__.py can hand numpy views of reference tensors
 * straight in.  Nothing here is tuned; loops are written in the order the maths
 * is stated.  OpenMP is used only over independent output elements.
 *
 * Pinning: tests/test_oracle_golden.py checks every function below against tensors
 * produced by importing the reference itself (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* bilinear sampling, F.grid_sample(mode='bilinear', padding_mode='zeros',    */
/* align_corners=True) as called at models/MVSNet/module.py:165-166,          */
/* models/VisMVSNet/homography.py:101-102, CVP modules.py:120-121,284-286.    */
/* gx, gy are the NORMALISED coordinates the reference hands to grid_sample.  */
/* ------------------------------------------------------------------------- */
typedef struct {
    int x0, y0;      /* north-west tap */
    float w[4];      /* nw, ne, sw, se */
    int ok[4];       /* tap inside the source map? (zeros padding is per tap) */
} taps_t;

static inline taps_t make_taps(float gx, float gy, int Hs, int Ws)
{
    taps_t t;
    /* align_corners=True un-normalisation: ((g + 1) / 2) * (size - 1) */
    float ix = ((gx + 1.0f) / 2.0f) * (float)(Ws - 1);
    float iy = ((gy + 1.0f) / 2.0f) * (float)(Hs - 1);
    float fx = floorf(ix), fy = floorf(iy);
    t.x0 = (int)fx;
    t.y0 = (int)fy;
    float x1 = fx + 1.0f, y1 = fy + 1.0f;
    t.w[0] = (x1 - ix) * (y1 - iy);
    t.w[1] = (ix - fx) * (y1 - iy);
    t.w[2] = (x1 - ix) * (iy - fy);
    t.w[3] = (ix - fx) * (iy - fy);
    int xin0 = t.x0 >= 0 && t.x0 < Ws, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < Ws;
    int yin0 = t.y0 >= 0 && t.y0 < Hs, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < Hs;
    t.ok[0] = xin0 && yin0;
    t.ok[1] = xin1 && yin0;
    t.ok[2] = xin0 && yin1;
    t.ok[3] = xin1 && yin1;
    return t;
}

static inline float sample_taps(const float *plane, int Ws, const taps_t *t)
{
    float v = 0.0f;
    if (t->ok[0]) v += plane[(size_t)t->y0 * Ws + t->x0] * t->w[0];
    if (t->ok[1]) v += plane[(size_t)t->y0 * Ws + t->x0 + 1] * t->w[1];
    if (t->ok[2]) v += plane[(size_t)(t->y0 + 1) * Ws + t->x0] * t->w[2];
    if (t->ok[3]) v += plane[(size_t)(t->y0 + 1) * Ws + t->x0 + 1] * t->w[3];
    return v;
}

static inline float clampf(float v, float lo, float hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}

/* ------------------------------------------------------------------------- */
/* 4x4 inverse (fp32, Gauss-Jordan with partial pivoting) and                 */
/* proj = src_proj @ inverse(ref_proj)   -- models/MVSNet/module.py:128,      */
/* CVP modules.py:95, 250.  Output: rot[9] (row-major), trans[3].             */
/* ------------------------------------------------------------------------- */
static int inv4f(const float *a, float *out)
{
    float m[4][8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            m[i][j] = a[i * 4 + j];
            m[i][4 + j] = (i == j) ? 1.0f : 0.0f;
        }
    for (int c = 0; c < 4; c++) {
        int p = c;
        for (int r = c + 1; r < 4; r++)
            if (fabsf(m[r][c]) > fabsf(m[p][c])) p = r;
        if (m[p][c] == 0.0f) return -1;
        if (p != c)
            for (int j = 0; j < 8; j++) { float tmp = m[c][j]; m[c][j] = m[p][j]; m[p][j] = tmp; }
        float piv = m[c][c];
        for (int j = 0; j < 8; j++) m[c][j] /= piv;
        for (int r = 0; r < 4; r++) {
            if (r == c) continue;
            float f = m[r][c];
            for (int j = 0; j < 8; j++) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) out[i * 4 + j] = m[i][4 + j];
    return 0;
}

ORC_API int orc_relative_proj(const float *src_proj, const float *ref_proj, float *rot, float *trans)
{
    float inv[16], p[16];
    if (inv4f(ref_proj, inv)) return -1;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += src_proj[i * 4 + k] * inv[k * 4 + j];
            p[i * 4 + j] = s;
        }
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) rot[i * 3 + j] = p[i * 4 + j];
        trans[i] = p[i * 4 + 3];
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* a1 / a15 / a16: homo_warping, models/MVSNet/module.py:111-169              */
/* (identical maths at CVP modules.py:74-128 and inside proj_cost :229-286).  */
/*   src   [C,Hs,Ws]   rot[9] trans[3]                                         */
/*   depth [D] (per_pixel=0) or [D,H,W] (per_pixel=1)                          */
/*   out   [C,D,H,W]                                                           */
/* ------------------------------------------------------------------------- */
static inline void mvs_grid(const float *rot, const float *trans, int x, int y, float d,
                            int Hs, int Ws, float *gx, float *gy)
{
    /* module.py:138 rot_xyz = rot @ (x, y, 1) */
    float fx = (float)x, fy = (float)y;
    float rx = rot[0] * fx + rot[1] * fy + rot[2];
    float ry = rot[3] * fx + rot[4] * fy + rot[5];
    float rz = rot[6] * fx + rot[7] * fy + rot[8];
    /* :142-144 */
    float qx = rx * d + trans[0];
    float qy = ry * d + trans[1];
    float qz = rz * d + trans[2];
    /* :146-150 */
    float px = qx / qz, py = qy / qz;
    if (qz <= 0.0f) { px = -10.0f; py = -10.0f; }
    /* :151-155 */
    float nx = px / ((float)(Ws - 1) / 2.0f) - 1.0f;
    float ny = py / ((float)(Hs - 1) / 2.0f) - 1.0f;
    *gx = clampf(nx, -10.0f, 10.0f);
    *gy = clampf(ny, -10.0f, 10.0f);
}

ORC_API void orc_homo_warp_mvs(const float *src, int C, int Hs, int Ws, const float *rot,
                               const float *trans, const float *depth, int per_pixel, int D, int H,
                               int W, float *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int d = 0; d < D; d++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                float dv = per_pixel ? depth[((size_t)d * H + y) * W + x] : depth[d];
                float gx, gy;
                mvs_grid(rot, trans, x, y, dv, Hs, Ws, &gx, &gy);
                taps_t t = make_taps(gx, gy, Hs, Ws);
                for (int c = 0; c < C; c++)
                    out[(((size_t)c * D + d) * H + y) * W + x] =
                        sample_taps(src + (size_t)c * Hs * Ws, Ws, &t);
            }
}

/* ------------------------------------------------------------------------- */
/* a2: variance aggregation, models/MVSNet/model.py:113-139                    */
/*     cost = M2/V - M1^2/V^2                         (order 0, :134)          */
/* a15/a16: CVP net.py:152, modules.py:289   cost = M2/V - (M1/V)^2 (order 1)  */
/*   ref [C,H,W]; warped [S][C,D,H,W]; out [C,D,H,W]                           */
/* ------------------------------------------------------------------------- */
ORC_API void orc_variance(const float *ref, const float *const *warped, int S, int C, int D, int H,
                          int W, int order, float *out)
{
    const float V = (float)(S + 1);
    size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int c = 0; c < C; c++)
        for (int d = 0; d < D; d++)
            for (size_t p = 0; p < HW; p++) {
                float r = ref[(size_t)c * HW + p];
                float m1 = r, m2 = r * r;
                size_t o = ((size_t)c * D + d) * HW + p;
                for (int s = 0; s < S; s++) {
                    float w = warped[s][o];
                    m1 = m1 + w;
                    m2 = m2 + w * w;
                }
                if (order == 0)
                    out[o] = m2 / V - (m1 * m1) / (V * V);
                else {
                    float mean = m1 / V;
                    out[o] = m2 / V - mean * mean;
                }
            }
}

/* ------------------------------------------------------------------------- */
/* a3: softmin aggregation (MVSNet-s), models/MVSNet/model.py:141-173          */
/*   diff = (warp - ref)^2 ; e = exp(-temp * sum_c diff)                       */
/*   cost = sum_v e*diff / (sum_v e + 1e-6)                                    */
/* ------------------------------------------------------------------------- */
ORC_API void orc_softmin(const float *ref, const float *const *warped, int S, int C, int D, int H,
                         int W, float temp, float *out)
{
    size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (int d = 0; d < D; d++)
        for (size_t p = 0; p < HW; p++) {
            float sum_exp = 0.0f;
            for (int c = 0; c < C; c++) out[((size_t)c * D + d) * HW + p] = 0.0f;
            for (int s = 0; s < S; s++) {
                float ssd = 0.0f;
                for (int c = 0; c < C; c++) {
                    float df = warped[s][((size_t)c * D + d) * HW + p] - ref[(size_t)c * HW + p];
                    ssd += df * df;
                }
                float e = expf(-temp * ssd);
                sum_exp += e;
                for (int c = 0; c < C; c++) {
                    float df = warped[s][((size_t)c * D + d) * HW + p] - ref[(size_t)c * HW + p];
                    out[((size_t)c * D + d) * HW + p] += (df * df) * e;
                }
            }
            for (int c = 0; c < C; c++) out[((size_t)c * D + d) * HW + p] /= (sum_exp + 1e-6f);
        }
}

/* ------------------------------------------------------------------------- */
/* a8 + a9: Vis-MVSNet homography + warp                                       */
/*   get_homographies      models/VisMVSNet/homography.py:23-74                */
/*   homography_warping    :107-121, interpolate :85-104, pixel grid :77-82    */
/* The reference materialises H_d = K_r R_r (I - c_rel n / (depth+1e-9)) R_l^T */
/* K_l^-1 per voxel; here the same 3x3 product chain is evaluated per voxel    */
/* in the same association order (fp32).                                       */
/*   cams: ref/src [2,4,4] ([0]=R|t, [1]=K) already scaled by scale_camera     */
/*   depth_start: [1] or [H,W]; depth_interval: scalar                         */
/*   src [C,Hs,Ws] -> out [C,D,H,W]                                            */
/* ------------------------------------------------------------------------- */
static void mat3mul(const float *a, const float *b, float *o)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            float s = 0.0f;
            for (int k = 0; k < 3; k++) s += a[i * 3 + k] * b[k * 3 + j];
            o[i * 3 + j] = s;
        }
}

static int inv3f(const float *a, float *o)
{
    float det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
                a[2] * (a[3] * a[7] - a[4] * a[6]);
    if (det == 0.0f) return -1;
    float id = 1.0f / det;
    o[0] = (a[4] * a[8] - a[5] * a[7]) * id;
    o[1] = (a[2] * a[7] - a[1] * a[8]) * id;
    o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = (a[5] * a[6] - a[3] * a[8]) * id;
    o[4] = (a[0] * a[8] - a[2] * a[6]) * id;
    o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = (a[3] * a[7] - a[4] * a[6]) * id;
    o[7] = (a[1] * a[6] - a[0] * a[7]) * id;
    o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
    return 0;
}

ORC_API int orc_vis_warp(const float *src, int C, int Hs, int Ws, const float *ref_cam,
                         const float *src_cam, const float *depth_start, int start_per_pixel,
                         float depth_interval, int D, int H, int W, float *out)
{
    float Rl[9], Rr[9], tl[3], tr[3], Kl[9], Kr[9];
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) {
            Rl[i * 3 + j] = ref_cam[i * 4 + j];
            Rr[i * 3 + j] = src_cam[i * 4 + j];
            Kl[i * 3 + j] = ref_cam[16 + i * 4 + j];
            Kr[i * 3 + j] = src_cam[16 + i * 4 + j];
        }
        tl[i] = ref_cam[i * 4 + 3];
        tr[i] = src_cam[i * 4 + 3];
    }
    float Kli[9];
    if (inv3f(Kl, Kli)) return -1;
    float RlT[9], RrT[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { RlT[i * 3 + j] = Rl[j * 3 + i]; RrT[i * 3 + j] = Rr[j * 3 + i]; }
    /* c = -R^T t ; c_rel = c_right - c_left  (homography.py:53-55) */
    float cl[3], cr[3], crel[3];
    for (int i = 0; i < 3; i++) {
        cl[i] = -(RlT[i * 3] * tl[0] + RlT[i * 3 + 1] * tl[1] + RlT[i * 3 + 2] * tl[2]);
        cr[i] = -(RrT[i * 3] * tr[0] + RrT[i * 3 + 1] * tr[1] + RrT[i * 3 + 2] * tr[2]);
        crel[i] = cr[i] - cl[i];
    }
    /* temp_vec = c_rel @ fronto_direction (outer product, :58) */
    float tv[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) tv[i * 3 + j] = crel[i] * Rl[6 + j];
    float mm1[9], KrRr[9];
    mat3mul(RlT, Kli, mm1);  /* :62 */
    mat3mul(Kr, Rr, KrRr);   /* :65, left-to-right association */
    int nan_seen = 0;
#pragma omp parallel for collapse(2) schedule(static) reduction(| : nan_seen)
    for (int d = 0; d < D; d++)
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                float ds = start_per_pixel ? depth_start[(size_t)y * W + x] : depth_start[0];
                float depth = ds + depth_interval * (float)d; /* :39 */
                float mm0[9], mm2[9], Hm[9];
                for (int k = 0; k < 9; k++)
                    mm0[k] = ((k % 4 == 0) ? 1.0f : 0.0f) - tv[k] / (depth + 1e-9f); /* :60-61 */
                mat3mul(mm0, mm1, mm2);
                mat3mul(KrRr, mm2, Hm);
                for (int k = 0; k < 9; k++)
                    if (isnan(Hm[k])) nan_seen |= 1; /* :71-72 raises */
                float px = (float)x + 0.5f, py = (float)y + 0.5f; /* :78-79 */
                float qx = Hm[0] * px + Hm[1] * py + Hm[2];
                float qy = Hm[3] * px + Hm[4] * py + Hm[5];
                float qz = Hm[6] * px + Hm[7] * py + Hm[8];
                float zc = qz < 1e-9f ? 1e-9f : qz; /* :116 clamp(min=1e-9) */
                float u = qx / zc, v = qy / zc;
                if (!(qz > 0.0f)) { u = -10.0f; v = -10.0f; } /* :115,117 */
                /* interpolate(): normalise by SIZE (not size-1), clamp +-1.1 (:93-96) */
                float gx = clampf((u / (float)Ws) * 2.0f - 1.0f, -1.1f, 1.1f);
                float gy = clampf((v / (float)Hs) * 2.0f - 1.0f, -1.1f, 1.1f);
                taps_t t = make_taps(gx, gy, Hs, Ws);
                for (int c = 0; c < C; c++)
                    out[(((size_t)c * D + d) * H + y) * W + x] =
                        sample_taps(src + (size_t)c * Hs * Ws, Ws, &t);
            }
    return nan_seen ? -2 : 0;
}

/* a10: groupwise_correlation, models/VisMVSNet/nn_utils.py:473-490 (SUM over the group) */
ORC_API void orc_groupcorr(const float *ref, const float *warped, int C, int G, int D, int H, int W,
                           float *out)
{
    size_t HW = (size_t)H * W;
    int cpg = C / G;
#pragma omp parallel for collapse(2) schedule(static)
    for (int g = 0; g < G; g++)
        for (int d = 0; d < D; d++)
            for (size_t p = 0; p < HW; p++) {
                float s = 0.0f;
                for (int k = 0; k < cpg; k++) {
                    int c = g * cpg + k;
                    s += ref[(size_t)c * HW + p] * warped[((size_t)c * D + d) * HW + p];
                }
                out[((size_t)g * D + d) * HW + p] = s;
            }
}

/* ------------------------------------------------------------------------- */
/* a4 / a11 / a18: nn.Conv3d and nn.ConvTranspose3d as used by                 */
/*   ConvBnReLU3D (models/MVSNet/module.py:41-48), CostRegNet (model.py:43-84) */
/*   UNet / BasicBlock (VisMVSNet/nn_utils.py:123-278), CVP net.py:50-85.      */
/* in [Cin,D,H,W]; w: conv [Cout,Cin,kd,kh,kw], deconv [Cin,Cout,kd,kh,kw];    */
/* bias may be NULL.  out [Cout,Do,Ho,Wo] with PyTorch's size rules.           */
/* ------------------------------------------------------------------------- */
ORC_API void orc_conv3d(const float *in, int Cin, int D, int H, int W, const float *w,
                        const float *bias, int Cout, int kd, int kh, int kw, int stride, int pd,
                        int ph, int pw, float *out)
{
    int Do = (D + 2 * pd - kd) / stride + 1;
    int Ho = (H + 2 * ph - kh) / stride + 1;
    int Wo = (W + 2 * pw - kw) / stride + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int o = 0; o < Cout; o++)
        for (int z = 0; z < Do; z++)
            for (int y = 0; y < Ho; y++)
                for (int x = 0; x < Wo; x++) {
                    float acc = 0.0f;
                    for (int c = 0; c < Cin; c++)
                        for (int a = 0; a < kd; a++) {
                            int zi = z * stride + a - pd;
                            if (zi < 0 || zi >= D) continue;
                            for (int b = 0; b < kh; b++) {
                                int yi = y * stride + b - ph;
                                if (yi < 0 || yi >= H) continue;
                                for (int e = 0; e < kw; e++) {
                                    int xi = x * stride + e - pw;
                                    if (xi < 0 || xi >= W) continue;
                                    acc += in[(((size_t)c * D + zi) * H + yi) * W + xi] *
                                           w[((((size_t)o * Cin + c) * kd + a) * kh + b) * kw + e];
                                }
                            }
                        }
                    if (bias) acc += bias[o];
                    out[(((size_t)o * Do + z) * Ho + y) * Wo + x] = acc;
                }
}

ORC_API void orc_deconv3d(const float *in, int Cin, int D, int H, int W, const float *w,
                          const float *bias, int Cout, int k, int stride, int pad, int outpad,
                          float *out)
{
    int Do = (D - 1) * stride - 2 * pad + k + outpad;
    int Ho = (H - 1) * stride - 2 * pad + k + outpad;
    int Wo = (W - 1) * stride - 2 * pad + k + outpad;
#pragma omp parallel for collapse(2) schedule(static)
    for (int o = 0; o < Cout; o++)
        for (int z = 0; z < Do; z++)
            for (int y = 0; y < Ho; y++)
                for (int x = 0; x < Wo; x++) {
                    float acc = 0.0f;
                    for (int c = 0; c < Cin; c++)
                        for (int a = 0; a < k; a++) {
                            int zn = z + pad - a;
                            if (zn < 0 || zn % stride) continue;
                            int zi = zn / stride;
                            if (zi >= D) continue;
                            for (int b = 0; b < k; b++) {
                                int yn = y + pad - b;
                                if (yn < 0 || yn % stride) continue;
                                int yi = yn / stride;
                                if (yi >= H) continue;
                                for (int e = 0; e < k; e++) {
                                    int xn = x + pad - e;
                                    if (xn < 0 || xn % stride) continue;
                                    int xi = xn / stride;
                                    if (xi >= W) continue;
                                    acc += in[(((size_t)c * D + zi) * H + yi) * W + xi] *
                                           w[((((size_t)c * Cout + o) * k + a) * k + b) * k + e];
                                }
                            }
                        }
                    if (bias) acc += bias[o];
                    out[(((size_t)o * Do + z) * Ho + y) * Wo + x] = acc;
                }
}

/* eval-mode nn.BatchNorm3d (+ optional ReLU), in place: y = (x-mean)/sqrt(var+eps)*gamma+beta */
ORC_API void orc_bn_relu(float *x, int C, size_t n_per_c, const float *gamma, const float *beta,
                         const float *mean, const float *var, float eps, int relu)
{
#pragma omp parallel for schedule(static)
    for (int c = 0; c < C; c++) {
        float inv = 1.0f / sqrtf(var[c] + eps);
        for (size_t i = 0; i < n_per_c; i++) {
            float v = (x[(size_t)c * n_per_c + i] - mean[c]) * inv * gamma[c] + beta[c];
            if (relu && v < 0.0f) v = 0.0f;
            x[(size_t)c * n_per_c + i] = v;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* a5 / a6 / a12 / a19: softmax over D + regression heads.                     */
/*   score [D,H,W]                                                             */
/*   depth mode 0: values [D]           (MVSNet module.py:174-178, CVP :356)   */
/*              1: values [D,H,W]       (CVP depth_regression_refine :362-365) */
/*              2: start([1]|[H,W]) + class*interval (Vis model_cas.py:348,405)*/
/*   conf mode  0: none                                                        */
/*              1: MVSNet/CVP 4-bin: idx=(long)sum p*d; sum p[idx-1..idx+2]    */
/*                 (models/MVSNet/model.py:211-215, CVP net.py:213-219)        */
/*              2: Vis window: sum p_d where |d - cls| <= 2 (nn_utils.py:464)  */
/*   entropy (optional): sum -p*log(clamp(p,1e-9,1)) (nn_utils.py:469-470)     */
/*   prob (optional) [D,H,W]                                                   */
/* ------------------------------------------------------------------------- */
ORC_API void orc_softmax_regress(const float *score, int D, int H, int W, int depth_mode,
                                 const float *dvals, int start_per_pixel, float interval,
                                 int conf_mode, float *depth, float *conf, float *entropy,
                                 float *prob)
{
    size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < HW; p++) {
        float mx = -INFINITY;
        for (int d = 0; d < D; d++) mx = fmaxf(mx, score[(size_t)d * HW + p]);
        float sum = 0.0f;
        for (int d = 0; d < D; d++) sum += expf(score[(size_t)d * HW + p] - mx);
        float e_idx = 0.0f, e_dep = 0.0f, ent = 0.0f;
        for (int d = 0; d < D; d++) {
            float pr = expf(score[(size_t)d * HW + p] - mx) / sum;
            if (prob) prob[(size_t)d * HW + p] = pr;
            e_idx += pr * (float)d;
            if (depth_mode == 0) e_dep += pr * dvals[d];
            else if (depth_mode == 1) e_dep += pr * dvals[(size_t)d * HW + p];
            float pc = pr < 1e-9f ? 1e-9f : (pr > 1.0f ? 1.0f : pr);
            ent += -pr * logf(pc);
        }
        if (depth_mode == 2) {
            float st = start_per_pixel ? dvals[p] : dvals[0];
            e_dep = e_idx * interval + st;
        }
        depth[p] = e_dep;
        if (entropy) entropy[p] = ent;
        if (conf && conf_mode == 1) {
            long idx = (long)e_idx; /* .long() truncation */
            float c = 0.0f;
            for (long j = idx - 1; j <= idx + 2; j++)
                if (j >= 0 && j < D) c += expf(score[(size_t)j * HW + p] - mx) / sum;
            conf[p] = c;
        } else if (conf && conf_mode == 2) {
            float c = 0.0f;
            for (int d = 0; d < D; d++)
                if (fabsf((float)d - e_idx) <= 2.0f) c += expf(score[(size_t)d * HW + p] - mx) / sum;
            conf[p] = c;
        }
    }
}

/* a12: visibility-weighted fusion, models/VisMVSNet/model_cas.py:354-357,385-386 */
/*   interm [S][G,D,H,W]; uncert [S][H,W] -> fused [G,D,H,W]                       */
ORC_API void orc_vis_fuse(const float *const *interm, const float *const *uncert, int S, int G,
                          int D, int H, int W, float *fused)
{
    size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int g = 0; g < G; g++)
        for (int d = 0; d < D; d++)
            for (size_t p = 0; p < HW; p++) {
                float acc = 0.0f, wsum = 0.0f;
                size_t o = ((size_t)g * D + d) * HW + p;
                for (int s = 0; s < S; s++) {
                    float wgt = expf(-uncert[s][p]);
                    wsum = wsum + wgt;
                    acc = acc + interm[s][o] * wgt;
                }
                fused[o] = acc / wsum;
            }
}
