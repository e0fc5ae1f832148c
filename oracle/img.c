/*
 * oracle/img.c -- CPU ORACLE (test infrastructure only; see rvio_oracle.h).
 *
 * Restatement of the OpenCV arithmetic that R-VIO's Tracker::track invokes
 * (src/rvio/Tracker.cc:198-202 CLAHE, :237-244 calcOpticalFlowPyrLK, :100-132 undistortPoints).
 * OpenCV is a non-vendored dependency of the reference (CMakeLists.txt:44-51, any >= 2.4.3); the
 * algorithm restated here is OpenCV 4.x's published one and is pinned bit-exactly against the cv2
 * 4.13.0 wheel by tests/test_oracle_cv2.py.
 *
 * All float32 expressions are evaluated operation by operation (no FMA contraction: compile with
 * -ffp-contract=off), in the accumulation order of OpenCV's SSE2 (SIMD128) code path.
 */
#include "rvio_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

static inline int round_half_even_f(float v) { return (int)lrintf(v); }   /* cvRound: default FE_TONEAREST */
static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/* ------------------------------------------------------------------ CLAHE */
/* OpenCV CLAHE_Impl::apply with clipLimit=3.0, tiles 5x5 (Tracker.cc:200). */
void orc_clahe(const uint8_t* src, int w, int h, int src_stride, uint8_t* dst, int dst_stride)
{
    const int TX = 5, TY = 5, HS = 256;
    int ew = w, eh = h;
    if (w % TX != 0 || h % TY != 0) {          /* both dims padded when either is ragged */
        ew = w + (TX - w % TX);
        eh = h + (TY - h % TY);
    }
    uint8_t* ext = (uint8_t*)malloc((size_t)ew * eh);
    for (int y = 0; y < eh; ++y) {
        int sy = reflect101(y, h);
        for (int x = 0; x < ew; ++x) ext[(size_t)y * ew + x] = src[(size_t)sy * src_stride + reflect101(x, w)];
    }
    const int tw = ew / TX, th = eh / TY, area = tw * th;
    const float lutScale = (float)(HS - 1) / area;
    int clip = (int)(3.0 * area / HS);
    if (clip < 1) clip = 1;

    uint8_t lut[25][256];
    for (int ty = 0; ty < TY; ++ty)
        for (int tx = 0; tx < TX; ++tx) {
            int hist[256];
            memset(hist, 0, sizeof hist);
            for (int y = 0; y < th; ++y)
                for (int x = 0; x < tw; ++x) hist[ext[(size_t)(ty * th + y) * ew + tx * tw + x]]++;
            int clipped = 0;
            for (int i = 0; i < HS; ++i)
                if (hist[i] > clip) { clipped += hist[i] - clip; hist[i] = clip; }
            int batch = clipped / HS, residual = clipped - batch * HS;
            for (int i = 0; i < HS; ++i) hist[i] += batch;
            if (residual != 0) {
                int step = HS / residual;
                if (step < 1) step = 1;
                for (int i = 0; i < HS && residual > 0; i += step, residual--) hist[i]++;
            }
            int sum = 0;
            for (int i = 0; i < HS; ++i) {
                sum += hist[i];
                lut[ty * TX + tx][i] = sat_u8(round_half_even_f((float)sum * lutScale));
            }
        }
    free(ext);

    const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
    for (int y = 0; y < h; ++y) {
        float tyf = (float)y * inv_th - 0.5f;
        int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
        float ya = tyf - (float)ty1, ya1 = 1.0f - ya;
        if (ty1 < 0) ty1 = 0;
        if (ty2 > TY - 1) ty2 = TY - 1;
        for (int x = 0; x < w; ++x) {
            float txf = (float)x * inv_tw - 0.5f;
            int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
            float xa = txf - (float)tx1, xa1 = 1.0f - xa;
            if (tx1 < 0) tx1 = 0;
            if (tx2 > TX - 1) tx2 = TX - 1;
            int v = src[(size_t)y * src_stride + x];
            float a = (float)lut[ty1 * TX + tx1][v] * xa1;
            float b = (float)lut[ty1 * TX + tx2][v] * xa;
            float c = (float)lut[ty2 * TX + tx1][v] * xa1;
            float d = (float)lut[ty2 * TX + tx2][v] * xa;
            float top = a + b, bot = c + d;
            float t0 = top * ya1, t1 = bot * ya;
            float res = t0 + t1;
            dst[(size_t)y * dst_stride + x] = sat_u8(round_half_even_f(res));
        }
    }
}

/* ------------------------------------------------------------------ pyrDown */
void orc_pyr_down(const uint8_t* src, int w, int h, int src_stride, uint8_t* dst, int dst_stride)
{
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int* rows = (int*)malloc(sizeof(int) * (size_t)dw * h);   /* horizontal pass for every source row */
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * src_stride;
        for (int x = 0; x < dw; ++x) {
            int c = 2 * x;
            int m2 = s[reflect101(c - 2, w)], m1 = s[reflect101(c - 1, w)], p0 = s[reflect101(c, w)];
            int p1 = s[reflect101(c + 1, w)], p2 = s[reflect101(c + 2, w)];
            rows[(size_t)y * dw + x] = p0 * 6 + (m1 + p1) * 4 + m2 + p2;
        }
    }
    for (int y = 0; y < dh; ++y) {
        int c = 2 * y;
        const int* r0 = rows + (size_t)reflect101(c - 2, h) * dw;
        const int* r1 = rows + (size_t)reflect101(c - 1, h) * dw;
        const int* r2 = rows + (size_t)reflect101(c, h) * dw;
        const int* r3 = rows + (size_t)reflect101(c + 1, h) * dw;
        const int* r4 = rows + (size_t)reflect101(c + 2, h) * dw;
        for (int x = 0; x < dw; ++x)
            dst[(size_t)y * dst_stride + x] = (uint8_t)((r2[x] * 6 + (r1[x] + r3[x]) * 4 + r0[x] + r4[x] + 128) >> 8);
    }
    free(rows);
}

/* ------------------------------------------------------------------ Scharr */
void orc_scharr(const uint8_t* src, int w, int h, int src_stride, int16_t* dst)
{
    for (int y = 0; y < h; ++y) {
        const uint8_t* s0 = src + (size_t)reflect101(y - 1, h) * src_stride;
        const uint8_t* s1 = src + (size_t)y * src_stride;
        const uint8_t* s2 = src + (size_t)reflect101(y + 1, h) * src_stride;
        for (int x = 0; x < w; ++x) {
            int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            /* vertical smooth (3,10,3) and vertical difference at columns xm, x, xp */
            int sm_m = (s0[xm] + s2[xm]) * 3 + s1[xm] * 10;
            int sm_p = (s0[xp] + s2[xp]) * 3 + s1[xp] * 10;
            int df_m = s2[xm] - s0[xm], df_c = s2[x] - s0[x], df_p = s2[xp] - s0[xp];
            dst[((size_t)y * w + x) * 2 + 0] = (int16_t)(sm_p - sm_m);
            dst[((size_t)y * w + x) * 2 + 1] = (int16_t)((df_p + df_m) * 3 + df_c * 10);
        }
    }
}

/* ------------------------------------------------------------------ pyramidal LK */
typedef struct {
    int w, h;
    int bw;            /* padded width  (w + 2*B) */
    uint8_t* img;      /* padded by B with REFLECT_101 */
    int16_t* der;      /* padded by B with constant 0, interleaved dx,dy (prev only) */
} lk_level_t;

static void make_border_u8(const uint8_t* src, int w, int h, int stride, int B, uint8_t* dst)
{
    const int bw = w + 2 * B;
    for (int y = -B; y < h + B; ++y) {
        int sy = reflect101(y, h);
        for (int x = -B; x < w + B; ++x)
            dst[(size_t)(y + B) * bw + (x + B)] = src[(size_t)sy * stride + reflect101(x, w)];
    }
}

#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))

int orc_lk(const uint8_t* prev, const uint8_t* next, int w, int h, int stride,
           const float* prev_pts, int n, float* next_pts, uint8_t* status,
           int win, int max_level, int max_iter, double eps, double min_eig_thr)
{
    /* wrapper clamps (SparsePyrLKOpticalFlowImpl::calc): maxCount in [0,100], eps in [0,10], eps squared */
    if (max_iter < 0) max_iter = 0;
    if (max_iter > 100) max_iter = 100;
    if (eps < 0) eps = 0;
    if (eps > 10) eps = 10;
    eps *= eps;

    const int B = win;
    lk_level_t P[16], N[16];
    uint8_t *pl[16], *nl[16];
    int lw[16], lh[16];
    int levels = 0;
    /* buildOpticalFlowPyramid: stop before a level that is <= the window in either dimension */
    {
        int cw = w, ch = h;
        for (int l = 0; l <= max_level && l < 16; ++l) {
            if (l > 0) { cw = (cw + 1) / 2; ch = (ch + 1) / 2; }
            if (l > 0 && (cw <= win || ch <= win)) break;
            lw[l] = cw; lh[l] = ch;
            pl[l] = (uint8_t*)malloc((size_t)cw * ch);
            nl[l] = (uint8_t*)malloc((size_t)cw * ch);
            if (l == 0) {
                for (int y = 0; y < ch; ++y) {
                    memcpy(pl[0] + (size_t)y * cw, prev + (size_t)y * stride, (size_t)cw);
                    memcpy(nl[0] + (size_t)y * cw, next + (size_t)y * stride, (size_t)cw);
                }
            } else {
                orc_pyr_down(pl[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], pl[l], cw);
                orc_pyr_down(nl[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], nl[l], cw);
            }
            levels = l + 1;
        }
    }
    for (int l = 0; l < levels; ++l) {
        const int cw = lw[l], ch = lh[l], bw = cw + 2 * B, bh = ch + 2 * B;
        P[l].w = N[l].w = cw; P[l].h = N[l].h = ch; P[l].bw = N[l].bw = bw;
        P[l].img = (uint8_t*)malloc((size_t)bw * bh);
        N[l].img = (uint8_t*)malloc((size_t)bw * bh);
        make_border_u8(pl[l], cw, ch, cw, B, P[l].img);
        make_border_u8(nl[l], cw, ch, cw, B, N[l].img);
        int16_t* d = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)cw * ch);
        orc_scharr(pl[l], cw, ch, cw, d);
        P[l].der = (int16_t*)calloc((size_t)bw * bh * 2, sizeof(int16_t));
        for (int y = 0; y < ch; ++y)
            memcpy(P[l].der + ((size_t)(y + B) * bw + B) * 2, d + (size_t)y * cw * 2, sizeof(int16_t) * 2 * (size_t)cw);
        free(d);
        N[l].der = NULL;
    }

    for (int i = 0; i < n; ++i) status[i] = 1;

    const float half = (float)(win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    int16_t* Iw = (int16_t*)malloc(sizeof(int16_t) * (size_t)win * win);
    int16_t* dIw = (int16_t*)malloc(sizeof(int16_t) * 2 * (size_t)win * win);
    const int simd_px = (win >= 8) ? 8 * (win / 8) : 0;  /* columns handled by the 8-wide SIMD loop */

    for (int level = levels - 1; level >= 0; --level) {
        const lk_level_t* I = &P[level];
        const lk_level_t* J = &N[level];
        const int bw = I->bw;
        const float lscale = (float)(1. / (1 << level));
        for (int p = 0; p < n; ++p) {
            float px = prev_pts[2 * p] * lscale, py = prev_pts[2 * p + 1] * lscale;
            float nx, ny;
            if (level == levels - 1) { nx = px; ny = py; }
            else { nx = next_pts[2 * p] * 2.f; ny = next_pts[2 * p + 1] * 2.f; }
            next_pts[2 * p] = nx; next_pts[2 * p + 1] = ny;

            px -= half; py -= half;
            int ipx = (int)floorf(px), ipy = (int)floorf(py);
            if (ipx < -win || ipx >= I->w || ipy < -win || ipy >= I->h) {
                if (level == 0) status[p] = 0;
                continue;
            }
            float a = px - (float)ipx, b = py - (float)ipy;
            int iw00 = round_half_even_f((1.f - a) * (1.f - b) * (float)(1 << 14));
            int iw01 = round_half_even_f(a * (1.f - b) * (float)(1 << 14));
            int iw10 = round_half_even_f((1.f - a) * b * (float)(1 << 14));
            int iw11 = (1 << 14) - iw00 - iw01 - iw10;

            float qA11[4] = {0, 0, 0, 0}, qA12[4] = {0, 0, 0, 0}, qA22[4] = {0, 0, 0, 0};
            float sA11 = 0, sA12 = 0, sA22 = 0;
            for (int y = 0; y < win; ++y) {
                const uint8_t* src = I->img + (size_t)(y + ipy + B) * bw + (ipx + B);
                const int16_t* dsrc = I->der + ((size_t)(y + ipy + B) * bw + (ipx + B)) * 2;
                for (int x = 0; x < win; ++x) {
                    int ival = DESCALE(src[x] * iw00 + src[x + 1] * iw01 + src[x + bw] * iw10 + src[x + bw + 1] * iw11, 14 - 5);
                    int ixval = DESCALE(dsrc[2 * x] * iw00 + dsrc[2 * x + 2] * iw01 + dsrc[2 * x + 2 * bw] * iw10 + dsrc[2 * x + 2 * bw + 2] * iw11, 14);
                    int iyval = DESCALE(dsrc[2 * x + 1] * iw00 + dsrc[2 * x + 3] * iw01 + dsrc[2 * x + 2 * bw + 1] * iw10 + dsrc[2 * x + 2 * bw + 3] * iw11, 14);
                    Iw[y * win + x] = (int16_t)ival;
                    dIw[(y * win + x) * 2] = (int16_t)ixval;
                    dIw[(y * win + x) * 2 + 1] = (int16_t)iyval;
                    if (x < simd_px) {
                        int k = x & 3;
                        float fx = (float)ixval, fy = (float)iyval;
                        float t;
                        t = fy * fy; qA22[k] = t + qA22[k];
                        t = fx * fy; qA12[k] = t + qA12[k];
                        t = fx * fx; qA11[k] = t + qA11[k];
                    } else {
                        sA11 += (float)(ixval * ixval);
                        sA12 += (float)(ixval * iyval);
                        sA22 += (float)(iyval * iyval);
                    }
                }
            }
            if (simd_px) {
                sA11 += (qA11[0] + qA11[2]) + (qA11[1] + qA11[3]);
                sA12 += (qA12[0] + qA12[2]) + (qA12[1] + qA12[3]);
                sA22 += (qA22[0] + qA22[2]) + (qA22[1] + qA22[3]);
            }
            float A11 = sA11 * FLT_SCALE, A12 = sA12 * FLT_SCALE, A22 = sA22 * FLT_SCALE;
            float D;
            {
                float m0 = A11 * A22, m1 = A12 * A12;
                D = m0 - m1;
            }
            float minEig;
            {
                float df = A11 - A22;
                float t0 = df * df;
                float t1 = 4.f * A12;
                float t2 = t1 * A12;
                float t3 = t0 + t2;
                float sq = sqrtf(t3);
                float sm = A22 + A11;
                minEig = (sm - sq) / (float)(2 * win * win);
            }
            if (minEig < (float)min_eig_thr || D < FLT_EPSILON) {
                /* comparison happens in double in OpenCV (float promoted vs double threshold) */
                if (level == 0) status[p] = 0;
                continue;
            }
            D = 1.f / D;
            nx -= half; ny -= half;
            float pdx = 0, pdy = 0;
            for (int j = 0; j < max_iter; ++j) {
                int inx = (int)floorf(nx), iny = (int)floorf(ny);
                if (inx < -win || inx >= J->w || iny < -win || iny >= J->h) {
                    if (level == 0) status[p] = 0;
                    break;
                }
                a = nx - (float)inx; b = ny - (float)iny;
                iw00 = round_half_even_f((1.f - a) * (1.f - b) * (float)(1 << 14));
                iw01 = round_half_even_f(a * (1.f - b) * (float)(1 << 14));
                iw10 = round_half_even_f((1.f - a) * b * (float)(1 << 14));
                iw11 = (1 << 14) - iw00 - iw01 - iw10;
                float qb0[4] = {0, 0, 0, 0}, qb1[4] = {0, 0, 0, 0};
                float sb1 = 0, sb2 = 0;
                for (int y = 0; y < win; ++y) {
                    const uint8_t* Jp = J->img + (size_t)(y + iny + B) * bw + (inx + B);
                    const int16_t* Ip = Iw + y * win;
                    const int16_t* dIp = dIw + y * win * 2;
                    int diff[64];
                    for (int x = 0; x < win; ++x)
                        diff[x] = DESCALE(Jp[x] * iw00 + Jp[x + 1] * iw01 + Jp[x + bw] * iw10 + Jp[x + bw + 1] * iw11, 14 - 5) - Ip[x];
                    int x = 0;
                    for (; x + 8 <= win; x += 8) {
                        const int* d = diff + x;
                        const int16_t* g = dIp + 2 * x;
                        qb0[0] += (float)(d[0] * g[0] + d[4] * g[8]);
                        qb0[1] += (float)(d[0] * g[1] + d[4] * g[9]);
                        qb0[2] += (float)(d[1] * g[2] + d[5] * g[10]);
                        qb0[3] += (float)(d[1] * g[3] + d[5] * g[11]);
                        qb1[0] += (float)(d[2] * g[4] + d[6] * g[12]);
                        qb1[1] += (float)(d[2] * g[5] + d[6] * g[13]);
                        qb1[2] += (float)(d[3] * g[6] + d[7] * g[14]);
                        qb1[3] += (float)(d[3] * g[7] + d[7] * g[15]);
                    }
                    for (; x < win; ++x) {
                        sb1 += (float)(diff[x] * dIp[2 * x]);
                        sb2 += (float)(diff[x] * dIp[2 * x + 1]);
                    }
                }
                if (win >= 8) {
                    float q0 = qb0[0] + qb1[0], q1 = qb0[1] + qb1[1], q2 = qb0[2] + qb1[2], q3 = qb0[3] + qb1[3];
                    /* v_reduce_sum of [q0 q2 0 0] and [q1 q3 0 0]: (q0+0)+(q2+0) */
                    sb1 += (q0 + 0.f) + (q2 + 0.f);
                    sb2 += (q1 + 0.f) + (q3 + 0.f);
                }
                float b1 = sb1 * FLT_SCALE, b2 = sb2 * FLT_SCALE;
                float dx, dy;
                {
                    float m0 = A12 * b2, m1 = A22 * b1;
                    dx = (m0 - m1) * D;
                    float m2 = A12 * b1, m3 = A11 * b2;
                    dy = (m2 - m3) * D;
                }
                nx += dx; ny += dy;
                next_pts[2 * p] = nx + half; next_pts[2 * p + 1] = ny + half;
                if ((double)dx * (double)dx + (double)dy * (double)dy <= eps) break;
                if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                    next_pts[2 * p] -= dx * 0.5f;
                    next_pts[2 * p + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (status[p] && level == 0) {
                /* err branch of OpenCV (err array is always requested by Tracker.cc:244) */
                float fx = next_pts[2 * p] - half, fy = next_pts[2 * p + 1] - half;
                int ix = (int)floorf(fx), iy = (int)floorf(fy);
                if (ix < -win || ix >= J->w || iy < -win || iy >= J->h) status[p] = 0;
            }
        }
    }
    free(Iw); free(dIw);
    for (int l = 0; l < levels; ++l) {
        free(P[l].img); free(N[l].img); free(P[l].der); free(pl[l]); free(nl[l]);
    }
    return levels - 1;
}

/* ------------------------------------------------------------------ undistortPoints */
/* cv::undistortPoints with K (float32-rounded values), D (k1,k2,p1,p2[,k3]); 5 fixed-point iterations. */
void orc_undistort(const float* px, int n, const float* K4, const float* D5, float* out)
{
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    const double k1 = D5[0], k2 = D5[1], p1 = D5[2], p2 = D5[3], k3 = D5[4];
    for (int i = 0; i < n; ++i) {
        double u = px[2 * i], v = px[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            double r2 = x * x + y * y;
            double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            double deltaX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
            double deltaY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        out[2 * i] = (float)x;
        out[2 * i + 1] = (float)y;
    }
}

/* cv::fisheye::undistortPoints(src, dst, K, D) with no R/P (Tracker.cc:119), OpenCV 4.x: Newton iterations on
 * theta (at most 10, stop when |fix| < 1e-8), scale = tan(theta) / theta_d; a point that did not converge or whose theta
 * changed sign becomes (-1e6, -1e6).  D4 = the four coefficients the reference passes (Camera.k1, k2, p1, p2 read as k1..k4). */
void orc_undistort_fisheye(const float* px, int n, const float* K4, const float* D4, float* out)
{
    const double fx = K4[0], fy = K4[1], cx = K4[2], cy = K4[3];
    const double k0 = D4[0], k1 = D4[1], k2 = D4[2], k3 = D4[3];
    const double eps = 1e-8, half_pi = 3.1415926535897932384626433832795 / 2.;
    for (int i = 0; i < n; ++i) {
        const double u = px[2 * i], v = px[2 * i + 1];
        const double pwx = (u - cx) / fx, pwy = (v - cy) / fy;
        double theta_d = sqrt(pwx * pwx + pwy * pwy);
        theta_d = fmin(fmax(-half_pi, theta_d), half_pi);
        int converged = 0;
        double theta = theta_d, scale = 0.0;
        if (fabs(theta_d) > eps) {
            for (int j = 0; j < 10; ++j) {
                const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
                const double a = k0 * t2, b = k1 * t4, c = k2 * t6, d = k3 * t8;
                const double fix = (theta * (1 + a + b + c + d) - theta_d) / (1 + 3 * a + 5 * b + 7 * c + 9 * d);
                theta = theta - fix;
                if (fabs(fix) < eps) { converged = 1; break; }
            }
            scale = tan(theta) / theta_d;
        } else {
            converged = 1;
        }
        const int flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
        if (converged && !flipped) { out[2 * i] = (float)(pwx * scale); out[2 * i + 1] = (float)(pwy * scale); }
        else { out[2 * i] = -1000000.0f; out[2 * i + 1] = -1000000.0f; }
    }
}

