/* oracle/detector.c -- TEST INFRASTRUCTURE ONLY (see rvio_oracle.h).
 * CPU restatement of FeatureDetector::DetectWithSubPix (reference src/rvio/FeatureDetector.cc:55-75):
 *   cv::goodFeaturesToTrack(im, corners, nCorners, nQualLvl, s*nMinDist)      [blockSize 3, Sobel 3, min-eigenvalue]
 *   cv::cornerSubPix(im, corners, (floor(.5 nMinDist),)*2, (-1,-1), COUNT+EPS 30 / 1e-2)
 * OpenCV itself is not under /root/reference; the algorithm below follows OpenCV 4.x imgproc (featureselect.cpp,
 * corner.cpp, cornersubpix.cpp, samplers.cpp).  Arithmetic contract (what the CUDA path reproduces bit for bit):
 *   Sobel (scale 1/(255*4*3) folded into the smoothing taps k1 = s, k0 = 2s, float32):
 *        Dx = fma(r[y-1] + r[y+1], k1, r[y]*k0),  r = p[x+1] - p[x-1]            (== cv2 4.13 on AVX2/FMA hosts)
 *        Dy = q[y+1] - q[y-1],  q = fma(p[x+1], k1, fma(p[x], k0, p[x-1]*k1))    (== cv2 except its scalar tail columns)
 *   covariance products in float32, 3x3 box sums in double rounded to float32 (== cv2), min eigenvalue in float32
 *        (a*.5 + c*.5) - sqrt((a*.5 - c*.5)^2 + b*b)  without fma (== cv2)
 *   threshold (float)(max * quality), 3x3 local maximum on rows/cols 1..n-2, order by (value desc, index desc), greedy
 *   minimum-distance selection on a cell grid, at most nCorners  (== cv2)
 *   cornerSubPix: float32 bilinear 17x17 patch (a11*p00 + a12*p01 + a21*p10 + a22*p11, left to right, no fma), float32
 *   gradients, double accumulation: 32 strided partial sums (window index mod 32) combined by an xor butterfly -- the
 *   order a warp produces; OpenCV adds row-major, the difference is O(1e-16) relative.
 * Pinned against cv2 4.13 in tests/test_oracle_detector.py (corner sets and sub-pixel positions, stated tolerances). */
#include "rvio_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int refl101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * n - 2 - i; }
    return i;
}

void orc_min_eig_map(const uint8_t* img, int w, int h, int stride, float* eig)
{
    const float s = (float)(1.0 / (255.0 * 4 * 3));
    const float k1 = s, k0 = 2.f * s;
    float* dxx = (float*)malloc(sizeof(float) * (size_t)w * h);
    float* dxy = (float*)malloc(sizeof(float) * (size_t)w * h);
    float* dyy = (float*)malloc(sizeof(float) * (size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* rows[3] = {img + (size_t)refl101(y - 1, h) * stride, img + (size_t)y * stride, img + (size_t)refl101(y + 1, h) * stride};
        for (int x = 0; x < w; ++x) {
            const int xm = refl101(x - 1, w), xp = refl101(x + 1, w);
            float r[3], q[3];
            for (int k = 0; k < 3; ++k) {
                const float pm = rows[k][xm], pc = rows[k][x], pp = rows[k][xp];
                r[k] = pp - pm;
                q[k] = fmaf(pp, k1, fmaf(pc, k0, pm * k1));
            }
            const float dx = fmaf(r[0] + r[2], k1, r[1] * k0);
            const float dy = q[2] - q[0];
            const size_t o = (size_t)y * w + x;
            dxx[o] = dx * dx; dxy[o] = dx * dy; dyy[o] = dy * dy;
        }
    }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            double a = 0, b = 0, c = 0;
            for (int j = -1; j <= 1; ++j) {
                const size_t ro = (size_t)refl101(y + j, h) * w;
                for (int i = -1; i <= 1; ++i) {
                    const size_t o = ro + refl101(x + i, w);
                    a += dxx[o]; b += dxy[o]; c += dyy[o];
                }
            }
            const float fa = (float)a * 0.5f, fb = (float)b, fc = (float)c * 0.5f;
            const float t = fa - fc;
            eig[(size_t)y * w + x] = (fa + fc) - sqrtf(t * t + fb * fb);
        }
    free(dxx); free(dxy); free(dyy);
}

typedef struct { float v; int idx; } cand_t;
static int cand_cmp(const void* A, const void* B)
{
    const cand_t* a = (const cand_t*)A; const cand_t* b = (const cand_t*)B;
    if (a->v > b->v) return -1;
    if (a->v < b->v) return 1;
    return (a->idx > b->idx) ? -1 : (a->idx < b->idx);      /* greaterThanPtr: equal values -> higher address first */
}

/* cv::goodFeaturesToTrack(img, corners, max_corners, quality, min_dist): integer pixel corners, strongest first. */
int orc_good_features(const uint8_t* img, int w, int h, int stride, int max_corners, double quality, double min_dist, float* out_xy)
{
    float* eig = (float*)malloc(sizeof(float) * (size_t)w * h);
    orc_min_eig_map(img, w, h, stride, eig);
    float mx = 0.f;
    for (size_t i = 0; i < (size_t)w * h; ++i) if (eig[i] > mx) mx = eig[i];
    const float thr = (float)((double)mx * quality);
    cand_t* cand = (cand_t*)malloc(sizeof(cand_t) * (size_t)w * h / 4 + 64);
    int nc = 0;
    for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x) {
            const float v = eig[(size_t)y * w + x];
            if (!(v > thr)) continue;                       /* THRESH_TOZERO then val != 0 */
            float m = v;
            for (int j = -1; j <= 1; ++j)
                for (int i = -1; i <= 1; ++i) {
                    const float u = eig[(size_t)(y + j) * w + x + i];
                    const float uu = (u > thr) ? u : 0.f;   /* the dilation sees the thresholded map */
                    if (uu > m) m = uu;
                }
            if (v == m) { cand[nc].v = v; cand[nc].idx = y * w + x; ++nc; }
        }
    qsort(cand, nc, sizeof(cand_t), cand_cmp);
    int n_out = 0;
    if (min_dist >= 1) {
        const int cell = (int)lrint(min_dist);
        const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
        int* head = (int*)malloc(sizeof(int) * (size_t)gw * gh);
        int* next = (int*)malloc(sizeof(int) * (size_t)(max_corners > 0 ? max_corners : nc) + 4);
        for (int i = 0; i < gw * gh; ++i) head[i] = -1;
        const double md2 = min_dist * min_dist;
        for (int k = 0; k < nc; ++k) {
            const int y = cand[k].idx / w, x = cand[k].idx - y * w;
            const int xc = x / cell, yc = y / cell;
            int x1 = xc - 1, y1 = yc - 1, x2 = xc + 1, y2 = yc + 1;
            if (x1 < 0) x1 = 0;
            if (y1 < 0) y1 = 0;
            if (x2 > gw - 1) x2 = gw - 1;
            if (y2 > gh - 1) y2 = gh - 1;
            int good = 1;
            for (int yy = y1; yy <= y2 && good; ++yy)
                for (int xx = x1; xx <= x2 && good; ++xx)
                    for (int e = head[yy * gw + xx]; e >= 0; e = next[e]) {
                        const float dx = (float)x - out_xy[2 * e], dy = (float)y - out_xy[2 * e + 1];
                        if ((double)(dx * dx + dy * dy) < md2) { good = 0; break; }
                    }
            if (!good) continue;
            out_xy[2 * n_out] = (float)x; out_xy[2 * n_out + 1] = (float)y;
            next[n_out] = head[yc * gw + xc]; head[yc * gw + xc] = n_out;
            ++n_out;
            if (max_corners > 0 && n_out == max_corners) break;
        }
        free(head); free(next);
    } else {
        for (int k = 0; k < nc && (max_corners <= 0 || n_out < max_corners); ++k) {
            const int y = cand[k].idx / w, x = cand[k].idx - y * w;
            out_xy[2 * n_out] = (float)x; out_xy[2 * n_out + 1] = (float)y; ++n_out;
        }
    }
    free(cand); free(eig);
    return n_out;
}

/* cv::getRectSubPix(8U -> 32F) of a pw x ph patch centred at (cx, cy): bilinear, replicated border. */
static void rect_subpix(const uint8_t* img, int w, int h, int stride, float cx, float cy, int pw, int ph, float* dst)
{
    cx -= (pw - 1) * 0.5f; cy -= (ph - 1) * 0.5f;
    const int ix = (int)floorf(cx), iy = (int)floorf(cy);
    const float a = cx - ix, b = cy - iy;
    const float a11 = (1.f - a) * (1.f - b), a12 = a * (1.f - b), a21 = (1.f - a) * b, a22 = a * b;
    for (int i = 0; i < ph; ++i) {
        int y0 = iy + i, y1 = y0 + 1;
        y0 = y0 < 0 ? 0 : (y0 > h - 1 ? h - 1 : y0); y1 = y1 < 0 ? 0 : (y1 > h - 1 ? h - 1 : y1);
        const uint8_t* r0 = img + (size_t)y0 * stride; const uint8_t* r1 = img + (size_t)y1 * stride;
        for (int j = 0; j < pw; ++j) {
            int x0 = ix + j, x1 = x0 + 1;
            x0 = x0 < 0 ? 0 : (x0 > w - 1 ? w - 1 : x0); x1 = x1 < 0 ? 0 : (x1 > w - 1 ? w - 1 : x1);
            dst[i * pw + j] = r0[x0] * a11 + r0[x1] * a12 + r1[x0] * a21 + r1[x1] * a22;
        }
    }
}

/* cv::cornerSubPix(img, corners, (hw, hw), (-1,-1), TermCriteria(COUNT+EPS, max_iter, eps)), in place. */
void orc_corner_subpix(const uint8_t* img, int w, int h, int stride, float* xy, int n, int hw, int max_iter, double eps)
{
    const int ww = 2 * hw + 1;
    if (max_iter < 1) max_iter = 1;
    if (max_iter > 100) max_iter = 100;
    eps = eps < 0 ? 0 : eps; eps *= eps;
    float* mask = (float*)malloc(sizeof(float) * ww * ww);
    float* mx = (float*)malloc(sizeof(float) * ww);
    float* patch = (float*)malloc(sizeof(float) * (ww + 2) * (ww + 2));
    for (int j = 0; j < ww; ++j) { const float x = (float)(j - hw) / hw; mx[j] = (float)exp(-x * x); }
    for (int i = 0; i < ww; ++i) {
        const float y = (float)(i - hw) / hw;
        const float vy = (float)exp(-y * y);
        for (int j = 0; j < ww; ++j) mask[i * ww + j] = (float)(vy * mx[j]);
    }
    const int pw = ww + 2;
    for (int k = 0; k < n; ++k) {
        const float cTx = xy[2 * k], cTy = xy[2 * k + 1];
        float cIx = cTx, cIy = cTy;
        int iter = 0; double err = 0;
        do {
            rect_subpix(img, w, h, stride, cIx, cIy, pw, pw, patch);
            /* accumulation order = the CUDA kernel's: window position k = i*ww + j goes to partial sum k % 32 (a lane),
             * the 32 partials are combined by an xor butterfly (16, 8, 4, 2, 1) */
            double part[5][32];
            memset(part, 0, sizeof part);
            for (int k = 0; k < ww * ww; ++k) {
                const int i = k / ww, j = k - i * ww, l = k & 31;
                const double m = mask[k];
                const double tgx = patch[(i + 1) * pw + j + 2] - patch[(i + 1) * pw + j];
                const double tgy = patch[(i + 2) * pw + j + 1] - patch[i * pw + j + 1];
                const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                const double px = j - hw, py = i - hw;
                part[0][l] += gxx; part[1][l] += gxy; part[2][l] += gyy;
                part[3][l] += gxx * px + gxy * py;
                part[4][l] += gxy * px + gyy * py;
            }
            for (int q = 0; q < 5; ++q)
                for (int msk = 16; msk >= 1; msk >>= 1) {
                    double nw[32];
                    for (int l = 0; l < 32; ++l) nw[l] = part[q][l] + part[q][l ^ msk];
                    memcpy(part[q], nw, sizeof nw);
                }
            const double a = part[0][0], b = part[1][0], c = part[2][0], bb1 = part[3][0], bb2 = part[4][0];
            const double det = a * c - b * b;
            if (fabs(det) <= DBL_EPSILON * DBL_EPSILON) break;
            const double scale = 1.0 / det;
            const float nx = (float)(cIx + c * scale * bb1 - b * scale * bb2);
            const float ny = (float)(cIy - b * scale * bb1 + a * scale * bb2);
            err = (double)((nx - cIx) * (nx - cIx) + (ny - cIy) * (ny - cIy));
            cIx = nx; cIy = ny;
            if (cIx < 0 || cIx >= w || cIy < 0 || cIy >= h) break;
        } while (++iter < max_iter && err > eps);
        if (fabsf(cIx - cTx) > hw || fabsf(cIy - cTy) > hw) { cIx = cTx; cIy = cTy; }
        xy[2 * k] = cIx; xy[2 * k + 1] = cIy;
    }
    free(mask); free(mx); free(patch);
}

/* FeatureDetector::DetectWithSubPix (FeatureDetector.cc:55-75). */
int orc_detect_with_subpix(const uint8_t* img, int w, int h, int stride, int n_corners, int s, double quality, double min_dist, float* out_xy)
{
    const int n = orc_good_features(img, w, h, stride, n_corners, (double)(float)quality, (double)(s * (float)min_dist), out_xy);
    if (n > 0) orc_corner_subpix(img, w, h, stride, out_xy, n, (int)floor(.5 * (float)min_dist), 30, 1e-2);
    return n;
}
