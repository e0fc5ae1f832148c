/*
 * oracle/tracker.c -- CPU ORACLE (test infrastructure only; see rvio_oracle.h).
 *
 * Restates src/rvio/Tracker.cc:37-396 (per-frame front end and track bookkeeping) and
 * src/rvio/FeatureDetector.cc:78-150 (grid filter used by the refill).  The corner detector itself
 * (goodFeaturesToTrack + cornerSubPix, FeatureDetector.cc:55-75) stays outside: the caller runs it on
 * orc_tracker_image() and hands the corners in, so oracle and product are fed identical seeds.
 * Display / rviz code (Tracker.cc:135-176,266-269,353-356) is out of scope.
 */
#include "rvio_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

struct orc_tracker {
    orc_tracker_cfg_t cfg;
    int F, Fu, Lmax, Lmin;
    int first;
    int w, h;
    uint8_t *last, *cur;
    /* mvlTrackingHistory: F lists, capacity Lmax+1 */
    float* hist; int* hlen;
    /* mlFreeIndices FIFO (ring of capacity F+1) */
    int* freeq; int fq_head, fq_n;
    /* mvInlierIndices, mvFeatsToTrack, mPoints1ForRansac */
    int* slots; float* feats; double* pts1; int n_track;
    /* work state between track() and refill() */
    int* slots_new; float* feats_new; double* pts1_new; int n_new;
    /* outputs */
    uint8_t* up_types; int32_t* up_off; float* up_xy; int n_up;
    /* debug */
    int last_n; uint8_t *last_status, *last_flags; float *last_lk, *last_un;
    orc_ransac_t ransac;
};

static float* H(orc_tracker_t* t, int slot) { return t->hist + (size_t)slot * (t->Lmax + 1) * 2; }

static void hist_push(orc_tracker_t* t, int slot, float x, float y)
{
    float* h = H(t, slot);
    int n = t->hlen[slot];
    h[2 * n] = x; h[2 * n + 1] = y;
    t->hlen[slot] = n + 1;
}
static void hist_pop_front(orc_tracker_t* t, int slot)
{
    float* h = H(t, slot);
    int n = t->hlen[slot];
    memmove(h, h + 2, sizeof(float) * 2 * (size_t)(n - 1));
    t->hlen[slot] = n - 1;
}
static void fq_push(orc_tracker_t* t, int v) { t->freeq[(t->fq_head + t->fq_n) % (t->F + 1)] = v; t->fq_n++; }
static int fq_pop(orc_tracker_t* t) { int v = t->freeq[t->fq_head]; t->fq_head = (t->fq_head + 1) % (t->F + 1); t->fq_n--; return v; }

orc_tracker_t* orc_tracker_create(const orc_tracker_cfg_t* cfg)
{
    orc_tracker_t* t = (orc_tracker_t*)calloc(1, sizeof *t);
    t->cfg = *cfg;
    t->F = cfg->n_features;
    t->Fu = (int)ceil(.5 * t->F);                     /* Tracker.cc:74 */
    t->Lmax = cfg->max_track_len; t->Lmin = cfg->min_track_len;
    t->w = cfg->img_w; t->h = cfg->img_h;
    t->first = 1;
    size_t npx = (size_t)t->w * t->h;
    t->last = (uint8_t*)calloc(npx, 1); t->cur = (uint8_t*)calloc(npx, 1);
    t->hist = (float*)calloc((size_t)t->F * (t->Lmax + 1) * 2, sizeof(float));
    t->hlen = (int*)calloc((size_t)t->F, sizeof(int));
    t->freeq = (int*)calloc((size_t)t->F + 1, sizeof(int));
    t->slots = (int*)calloc((size_t)t->F, sizeof(int)); t->slots_new = (int*)calloc((size_t)t->F, sizeof(int));
    t->feats = (float*)calloc((size_t)t->F * 2, sizeof(float)); t->feats_new = (float*)calloc((size_t)t->F * 2, sizeof(float));
    t->pts1 = (double*)calloc((size_t)t->F * 3, sizeof(double)); t->pts1_new = (double*)calloc((size_t)t->F * 3, sizeof(double));
    t->up_types = (uint8_t*)calloc((size_t)t->Fu + 1, 1);
    t->up_off = (int32_t*)calloc((size_t)t->Fu + 2, sizeof(int32_t));
    t->up_xy = (float*)calloc((size_t)(t->Fu + 1) * (t->Lmax + 1) * 2, sizeof(float));
    t->last_status = (uint8_t*)calloc((size_t)t->F, 1); t->last_flags = (uint8_t*)calloc((size_t)t->F, 1);
    t->last_lk = (float*)calloc((size_t)t->F * 2, sizeof(float)); t->last_un = (float*)calloc((size_t)t->F * 2, sizeof(float));
    orc_ransac_init(&t->ransac, cfg->use_sampson, cfg->inlier_thr, cfg->small_angle, cfg->T_BC0);
    return t;
}

void orc_tracker_destroy(orc_tracker_t* t)
{
    if (!t) return;
    free(t->last); free(t->cur); free(t->hist); free(t->hlen); free(t->freeq);
    free(t->slots); free(t->slots_new); free(t->feats); free(t->feats_new); free(t->pts1); free(t->pts1_new);
    free(t->up_types); free(t->up_off); free(t->up_xy);
    free(t->last_status); free(t->last_flags); free(t->last_lk); free(t->last_un);
    free(t);
}

static void undist(const orc_tracker_t* t, const float* px, int n, float* out)
{
    float K[4] = {t->cfg.fx, t->cfg.fy, t->cfg.cx, t->cfg.cy};
    float D[5] = {t->cfg.k1, t->cfg.k2, t->cfg.p1, t->cfg.p2, t->cfg.k3};
    if (t->cfg.is_fisheye) orc_undistort_fisheye(px, n, K, D, out);
    else orc_undistort(px, n, K, D, out);
}

static void emit(orc_tracker_t* t, uint8_t type, int slot)
{
    int n = t->n_up;
    t->up_types[n] = type;
    int off = t->up_off[n], len = t->hlen[slot];
    memcpy(t->up_xy + 2 * (size_t)off, H(t, slot), sizeof(float) * 2 * (size_t)len);
    t->up_off[n + 1] = off + len;
    t->n_up = n + 1;
}

static int track_tail(orc_tracker_t* t, const double* imu, int n_imu);

int orc_tracker_track(orc_tracker_t* t, const uint8_t* img, int stride, const double* imu, int n_imu)
{
    const int w = t->w, h = t->h;
    /* Tracker.cc:183-202: mono input only here; CLAHE when enabled */
    if (t->cfg.enable_equalizer) orc_clahe(img, w, h, stride, t->cur, w);
    else for (int y = 0; y < h; ++y) memcpy(t->cur + (size_t)y * w, img + (size_t)y * stride, (size_t)w);

    if (t->first) return 1;                                      /* caller seeds: Tracker.cc:204-234 */

    const int n = t->n_track;
    t->last_n = n;
    if (n == 0) return 2;                                        /* Tracker.cc:246-250: early return, no commit */

    /* Tracker.cc:237-244 */
    orc_lk(t->last, t->cur, w, h, w, t->feats, n, t->last_lk, t->last_status, 15, 3, 30, 1e-2, 1e-3);
    return track_tail(t, imu, n_imu);
}

/* Same frame step with the OpenCV stages done by the caller through the real OpenCV (cv2): `eq` is the equalised
 * image (cv2 CLAHE), lk_px/lk_status the output of cv2.calcOpticalFlowPyrLK on (last image, eq, features).  Used by
 * bench.py's reference arm so that the CPU baseline runs OpenCV's own SIMD/threaded code, as the reference does. */
int orc_tracker_track_ext(orc_tracker_t* t, const uint8_t* eq, const float* lk_px, const uint8_t* lk_status,
                          const double* imu, int n_imu)
{
    memcpy(t->cur, eq, (size_t)t->w * t->h);
    if (t->first) return 1;
    const int n = t->n_track;
    t->last_n = n;
    if (n == 0) return 2;
    memcpy(t->last_lk, lk_px, sizeof(float) * 2 * (size_t)n);
    memcpy(t->last_status, lk_status, (size_t)n);
    return track_tail(t, imu, n_imu);
}
const float* orc_tracker_feats(const orc_tracker_t* t) { return t->feats; }     /* mvFeatsToTrack (committed) */
int orc_tracker_n_feats(const orc_tracker_t* t) { return t->n_track; }
const uint8_t* orc_tracker_last_image(const orc_tracker_t* t) { return t->last; }

static int track_tail(orc_tracker_t* t, const double* imu, int n_imu)
{
    const int n = t->n_track;
    /* Tracker.cc:252-261 */
    undist(t, t->last_lk, n, t->last_un);
    double* pts2 = (double*)malloc(sizeof(double) * 3 * (size_t)n);
    for (int i = 0; i < n; ++i) { pts2[3 * i] = t->last_un[2 * i]; pts2[3 * i + 1] = t->last_un[2 * i + 1]; pts2[3 * i + 2] = 1; }
    memcpy(t->last_flags, t->last_status, (size_t)n);
    /* Tracker.cc:264 */
    orc_ransac_find_inliers(&t->ransac, t->pts1, pts2, n, imu, n_imu, t->last_flags);
    free(pts2);

    /* Tracker.cc:271-342 */
    t->n_up = 0; t->up_off[0] = 0;
    int n_in = 0;
    for (int i = 0; i < n; ++i) {
        if (!t->last_flags[i]) {
            int idx = t->slots[i];
            fq_push(t, idx);
            if (t->hlen[idx] >= t->Lmin && t->n_up < t->Fu) emit(t, '1', idx);
            t->hlen[idx] = 0;
        }
    }
    for (int i = 0; i < n; ++i) {
        if (t->last_flags[i]) {
            int idx = t->slots[i];
            t->slots_new[n_in] = idx;
            t->feats_new[2 * n_in] = t->last_lk[2 * i]; t->feats_new[2 * n_in + 1] = t->last_lk[2 * i + 1];
            float ux = t->last_un[2 * i], uy = t->last_un[2 * i + 1];
            if (t->hlen[idx] == t->Lmax) {
                if (t->n_up < t->Fu) {
                    emit(t, '2', idx);
                    /* Tracker.cc:327-328: keep the newer part */
                    const double keep = t->Lmax - (ceil(.5 * t->Lmax) - 1);
                    while ((double)t->hlen[idx] > keep) hist_pop_front(t, idx);
                } else
                    hist_pop_front(t, idx);
            }
            hist_push(t, idx, ux, uy);
            t->pts1_new[3 * n_in] = ux; t->pts1_new[3 * n_in + 1] = uy; t->pts1_new[3 * n_in + 2] = 1;
            n_in++;
        }
    }
    t->n_new = n_in;
    return 0;
}

const uint8_t* orc_tracker_image(const orc_tracker_t* t) { return t->cur; }
int orc_tracker_n_free(const orc_tracker_t* t) { return t->fq_n; }
int orc_tracker_n_tracked(const orc_tracker_t* t) { return t->n_new; }
const float* orc_tracker_tracked_px(const orc_tracker_t* t) { return t->feats_new; }

void orc_tracker_seed(orc_tracker_t* t, const float* px, int n)   /* Tracker.cc:204-234 */
{
    if (n > t->F) n = t->F;
    if (n == 0) return;                                           /* Tracker.cc:209-213 */
    float* un = (float*)malloc(sizeof(float) * 2 * (size_t)n);
    undist(t, px, n, un);
    for (int i = 0; i < n; ++i) {
        t->feats_new[2 * i] = px[2 * i]; t->feats_new[2 * i + 1] = px[2 * i + 1];
        hist_push(t, i, un[2 * i], un[2 * i + 1]);
        t->pts1_new[3 * i] = un[2 * i]; t->pts1_new[3 * i + 1] = un[2 * i + 1]; t->pts1_new[3 * i + 2] = 1;
        t->slots_new[i] = i;
    }
    for (int i = n; i < t->F; ++i) fq_push(t, i);
    t->n_new = n;
    t->first = 0;
    free(un);
}

int orc_tracker_refill(orc_tracker_t* t, const float* px, int n)   /* Tracker.cc:358-386 */
{
    if (n == 0 || t->fq_n == 0) return 0;
    float* un = (float*)malloc(sizeof(float) * 2 * (size_t)n);
    undist(t, px, n, un);
    int k = 0;
    for (;;) {
        int idx = fq_pop(t);
        int m = t->n_new;
        t->slots_new[m] = idx;
        t->feats_new[2 * m] = px[2 * k]; t->feats_new[2 * m + 1] = px[2 * k + 1];
        hist_push(t, idx, un[2 * k], un[2 * k + 1]);
        t->pts1_new[3 * m] = un[2 * k]; t->pts1_new[3 * m + 1] = un[2 * k + 1]; t->pts1_new[3 * m + 2] = 1;
        t->n_new = m + 1;
        k++;
        if (t->fq_n == 0 || k == n || t->n_new == t->F) break;
    }
    free(un);
    return k;
}

void orc_tracker_commit(orc_tracker_t* t)   /* Tracker.cc:389-395 */
{
    t->n_track = t->n_new;
    memcpy(t->slots, t->slots_new, sizeof(int) * (size_t)t->n_new);
    memcpy(t->feats, t->feats_new, sizeof(float) * 2 * (size_t)t->n_new);
    memcpy(t->pts1, t->pts1_new, sizeof(double) * 3 * (size_t)t->n_new);
    uint8_t* tmp = t->last; t->last = t->cur; t->cur = tmp;
}

/* FeatureDetector.cc:29-48,78-150 */
int orc_find_newer(const orc_tracker_cfg_t* cfg, const float* corners, int n_corners,
                   const float* ref, int n_ref, float* out)
{
    const float min_dist = (float)cfg->min_dist;                   /* float members in FeatureDetector.h */
    const float bx = (float)cfg->block_x, by = (float)cfg->block_y;
    const int W = cfg->img_w, Hh = cfg->img_h;
    const int gc = (int)floor((double)W / cfg->block_x), gr = (int)floor((double)Hh / cfg->block_y);
    const int nb = gc * gr;
    const int offx = (int)(.5 * (W - gc * bx)), offy = (int)(.5 * (Hh - gr * by));   /* int members, FeatureDetector.h:66-67 */
    const int max_per_block = (int)((float)cfg->n_features / nb);                     /* int member, FeatureDetector.h:74 */
    int cap = n_ref + n_corners + 1;
    float* gpts = (float*)malloc(sizeof(float) * 2 * (size_t)cap * (size_t)nb);
    int* gcnt = (int*)calloc((size_t)nb, sizeof(int));
    for (int i = 0; i < n_ref; ++i) {
        float x = ref[2 * i], y = ref[2 * i + 1];
        if (x <= offx || y <= offy || x >= (W - offx) || y >= (Hh - offy)) continue;
        int col = (int)floorf((x - offx) / bx), row = (int)floorf((y - offy) / by);
        int b = row * gc + col;
        float* g = gpts + 2 * ((size_t)b * cap + gcnt[b]++);
        g[0] = x; g[1] = y;
    }
    int nout = 0;
    for (int i = 0; i < n_corners; ++i) {
        float x = corners[2 * i], y = corners[2 * i + 1];
        if (x <= offx || y <= offy || x >= (W - offx) || y >= (Hh - offy)) continue;
        int col = (int)floorf((x - offx) / bx), row = (int)floorf((y - offy) / by);
        float xl = col * bx + offx, xr = xl + bx, yt = row * by + offy, yb = yt + by;
        if (fabs(x - xl) < min_dist || fabs(x - xr) < min_dist || fabs(y - yt) < min_dist || fabs(y - yb) < min_dist) continue;
        int b = row * gc + col;
        if ((float)gcnt[b] < .75 * max_per_block) {
            int cnt = 0;
            for (int k = 0; k < gcnt[b]; ++k) {
                const float* g = gpts + 2 * ((size_t)b * cap + k);
                double dx = (double)(x - g[0]), dy = (double)(y - g[1]);
                if (sqrt(dx * dx + dy * dy) > 1 * min_dist) cnt++;
                else break;
            }
            if (cnt == gcnt[b]) {
                out[2 * nout] = x; out[2 * nout + 1] = y; nout++;
                float* g = gpts + 2 * ((size_t)b * cap + gcnt[b]++);
                g[0] = x; g[1] = y;
            }
        }
    }
    free(gpts); free(gcnt);
    return nout;
}

int orc_tracker_n_update(const orc_tracker_t* t) { return t->n_up; }
const uint8_t* orc_tracker_update_types(const orc_tracker_t* t) { return t->up_types; }
const int32_t* orc_tracker_update_offsets(const orc_tracker_t* t) { return t->up_off; }
const float* orc_tracker_update_xy(const orc_tracker_t* t) { return t->up_xy; }
int orc_tracker_last_n(const orc_tracker_t* t) { return t->last_n; }
const uint8_t* orc_tracker_last_status(const orc_tracker_t* t) { return t->last_status; }
const uint8_t* orc_tracker_last_flags(const orc_tracker_t* t) { return t->last_flags; }
const float* orc_tracker_last_lk(const orc_tracker_t* t) { return t->last_lk; }
const float* orc_tracker_last_un(const orc_tracker_t* t) { return t->last_un; }
const int32_t* orc_tracker_slots(const orc_tracker_t* t) { return t->slots; }
orc_tracker_t* orc_tracker_self(orc_tracker_t* t) { return t; }
orc_ransac_t* orc_tracker_ransac(orc_tracker_t* t) { return &t->ransac; }
