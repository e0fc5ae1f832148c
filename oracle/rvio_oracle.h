/*
 * rvio_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never a fallback).
 *
 * Plain-C restatement of the R-VIO per-frame hot path (Tracker::track + Updater::update) and
 * of the small host stages that feed it (PreIntegrator::propagate, clone augmentation,
 * composition).  Every function cites the reference file:line it follows
 * (paths relative to the reference checkout, e.g. src/rvio/Tracker.cc:179-396).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * link or call this library.  The product (r-vio_b200/csrc) never does.
 *
 * PARITY PINNING
 *   - OpenCV stages (CLAHE, pyrDown, Scharr, pyramidal LK, undistortPoints): the reference
 *     calls OpenCV, which is not vendored.  These restatements are pinned BIT-EXACTLY against
 *     the real OpenCV code in the cv2 4.13.0 wheel (tests/test_oracle_cv2.py, fixtures in
 *     tests/golden/ made by oracle/make_golden.py).
 *   - glibc rand(): pinned against libc.so.6 rand() (tests/test_oracle_ransac.py).
 *   - Eigen stages (RANSAC algebra, Updater, PreIntegrator, System): Eigen is absent in this
 *     image, the reference cannot be compiled (needs ROS + Eigen + OpenCV C++), and the
 *     reference ships no tests or golden vectors.  PARITY UNPINNED by the reference for these
 *     stages; they are cross-checked against an independent NumPy/LAPACK float64 restatement
 *     (oracle/np_updater.py) to ~1e-9.
 */
#ifndef RVIO_ORACLE_H
#define RVIO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------- image stages (oracle/img.c) ---------- */

/* CLAHE(clipLimit 3.0, tiles 5x5), Tracker.cc:198-202 -> OpenCV CLAHE_Impl::apply. */
void orc_clahe(const uint8_t* src, int w, int h, int src_stride, uint8_t* dst, int dst_stride);

/* cv::pyrDown (5x5 binomial, BORDER_REFLECT_101), used inside calcOpticalFlowPyrLK, Tracker.cc:244. */
void orc_pyr_down(const uint8_t* src, int w, int h, int src_stride, uint8_t* dst, int dst_stride);

/* Scharr derivative, int16 interleaved (dx,dy), as OpenCV calcScharrDeriv (lkpyramid.cpp). */
void orc_scharr(const uint8_t* src, int w, int h, int src_stride, int16_t* dst /* 2*w*h */);

/* cv::calcOpticalFlowPyrLK(prev, next, pts, ..., win x win, max_level, (COUNT+EPS,max_iter,eps), 0, min_eig)
 * Tracker.cc:237-244.  pts are float2 pixels.  Returns the number of pyramid levels actually used - 1. */
int orc_lk(const uint8_t* prev, const uint8_t* next, int w, int h, int stride,
           const float* prev_pts, int n, float* next_pts, uint8_t* status,
           int win, int max_level, int max_iter, double eps, double min_eig_thr);

/* cv::undistortPoints(src, dst, K, D) with no R/P: pixel -> undistorted normalized, Tracker.cc:100-132.
 * K = {fx,fy,cx,cy} (float32-rounded), D = {k1,k2,p1,p2,k3}. */
void orc_undistort(const float* px, int n, const float* K4, const float* D5, float* out);
/* cv::fisheye::undistortPoints(src, dst, K, D) (Tracker.cc:119), D4 = (k1..k4). */
void orc_undistort_fisheye(const float* px, int n, const float* K4, const float* D4, float* out);

/* ---------- glibc rand() restatement (oracle/ransac.c) ---------- */
typedef struct { int32_t r[34]; int f, b; } orc_rand_t;
void orc_rand_seed(orc_rand_t* st, unsigned seed);   /* srand(seed); never-seeded == seed 1 */
int  orc_rand_next(orc_rand_t* st);                   /* rand() */

/* ---------- RANSAC (oracle/ransac.c), Ransac.cc:50-266 ---------- */
typedef struct {
    int use_sampson;          /* Tracker.UseSampson */
    double inlier_thr;        /* Tracker.nInlierThrd */
    double small_angle;       /* IMU.nSmallAngle */
    double Ric[9];            /* row-major 3x3, from Camera.T_BC0 */
    orc_rand_t rng;
    /* last-call diagnostics */
    int two_points[32];
    int n_inliers[16];
    int winner;
    double hyp[16*9];         /* row-major E per hypothesis */
    double R[9];
} orc_ransac_t;

void orc_ransac_init(orc_ransac_t* rs, int use_sampson, double thr, double small_angle, const double* T_BC0_rowmajor16);
/* Points are 3xF homogeneous doubles stored column-major (x,y,1 per feature) as Tracker.cc:218-226,255-261.
 * imu: n_imu x 8 doubles (w[3], a[3], t, dt).  flags in/out.  Returns inlier count, or -1 when 17..31
 * candidates would hang the reference (Ransac.cc:57-82) -- flags untouched then. */
int orc_ransac_find_inliers(orc_ransac_t* rs, const double* pts1, const double* pts2, int F,
                            const double* imu, int n_imu, uint8_t* flags);

/* ---------- Tracker (oracle/tracker.c), Tracker.cc:37-396 ---------- */
typedef struct orc_tracker orc_tracker_t;

typedef struct {
    float fx, fy, cx, cy;             /* Camera.* read as float, Tracker.cc:39-42 */
    float k1, k2, p1, p2, k3;         /* Tracker.cc:51-61 */
    int   n_features;                 /* Tracker.nFeatures */
    int   max_track_len, min_track_len;
    int   enable_equalizer;
    int   use_sampson;
    double inlier_thr, small_angle;
    double T_BC0[16];                 /* row-major */
    /* FeatureDetector.cc:29-48 */
    int   img_w, img_h;
    double min_dist; int block_x, block_y;
    int   is_fisheye;                 /* Camera.Fisheye, Tracker.cc:119 */
} orc_tracker_cfg_t;

orc_tracker_t* orc_tracker_create(const orc_tracker_cfg_t* cfg);
void orc_tracker_destroy(orc_tracker_t* t);

/* One Tracker::track call minus the detector (which stays host/cv2):
 *   step 1: orc_tracker_track() does gray(identity for mono)/CLAHE, and either reports that the first
 *           image needs seeding (returns 1) or runs LK+undistort+RANSAC+bookkeeping (returns 0), or
 *           2 when there was nothing to track (Tracker.cc:246-250 early return).
 *   step 2: the caller detects corners on orc_tracker_image() and calls orc_tracker_seed() (first image,
 *           Tracker.cc:204-234) or orc_tracker_refill() (Tracker.cc:344-387, corners already passed through
 *           FindNewer) -- then orc_tracker_commit() stores the last image (Tracker.cc:395).
 */
int  orc_tracker_track(orc_tracker_t* t, const uint8_t* img, int stride, const double* imu, int n_imu);
int  orc_tracker_track_ext(orc_tracker_t* t, const uint8_t* eq, const float* lk_px, const uint8_t* lk_status,
                           const double* imu, int n_imu);   /* OpenCV stages supplied by the caller (cv2) */
const float* orc_tracker_feats(const orc_tracker_t* t);
int  orc_tracker_n_feats(const orc_tracker_t* t);
const uint8_t* orc_tracker_last_image(const orc_tracker_t* t);
const uint8_t* orc_tracker_image(const orc_tracker_t* t);      /* equalised current image, w*h */
int  orc_tracker_n_free(const orc_tracker_t* t);
int  orc_tracker_n_tracked(const orc_tracker_t* t);             /* mvFeatsToTrack.size() after bookkeeping */
const float* orc_tracker_tracked_px(const orc_tracker_t* t);    /* mvFeatsToTrack */
void orc_tracker_seed(orc_tracker_t* t, const float* px, int n);
int  orc_tracker_refill(orc_tracker_t* t, const float* px, int n);
void orc_tracker_commit(orc_tracker_t* t);

/* ---------- detector (oracle/detector.c), FeatureDetector.cc:55-75 via OpenCV imgproc (restated; see the file header) */
void orc_min_eig_map(const uint8_t* img, int w, int h, int stride, float* eig);     /* cv::cornerMinEigenVal(img, 3, 3) */
int  orc_good_features(const uint8_t* img, int w, int h, int stride, int max_corners, double quality, double min_dist, float* out_xy);
void orc_corner_subpix(const uint8_t* img, int w, int h, int stride, float* xy, int n, int half_win, int max_iter, double eps);
int  orc_detect_with_subpix(const uint8_t* img, int w, int h, int stride, int n_corners, int s, double quality, double min_dist, float* out_xy);

/* FeatureDetector::FindNewer grid filter (FeatureDetector.cc:78-150). Returns number kept, writes px. */
int  orc_find_newer(const orc_tracker_cfg_t* cfg, const float* corners, int n_corners,
                    const float* ref, int n_ref, float* out);

/* outputs == Tracker.h:70,74 in CSR form */
int  orc_tracker_n_update(const orc_tracker_t* t);
const uint8_t* orc_tracker_update_types(const orc_tracker_t* t);
const int32_t* orc_tracker_update_offsets(const orc_tracker_t* t);
const float*   orc_tracker_update_xy(const orc_tracker_t* t);
/* debug observables */
int  orc_tracker_last_n(const orc_tracker_t* t);                /* features fed to LK this frame */
const uint8_t* orc_tracker_last_status(const orc_tracker_t* t); /* LK status */
const uint8_t* orc_tracker_last_flags(const orc_tracker_t* t);  /* after RANSAC */
const float*   orc_tracker_last_lk(const orc_tracker_t* t);     /* vFeatsTracked px */
const float*   orc_tracker_last_un(const orc_tracker_t* t);     /* undistorted normalized */
const int32_t* orc_tracker_slots(const orc_tracker_t* t);       /* mvInlierIndices (current) */
orc_ransac_t*  orc_tracker_ransac(orc_tracker_t* t);

/* ---------- Updater (oracle/updater.c), Updater.cc:38-628 ---------- */
typedef struct {
    double sigma;            /* max((float)sigma_px,(float)sigma_py) widened, Updater.cc:42-44 */
    double Ric[9], tic[3];   /* row-major */
} orc_updater_cfg_t;

typedef struct {
    int n_feat, n_good, rows_stacked, rank, compressed, updated;
    int n_reject_init, n_reject_lm, n_reject_gate;
    int rank_full;           /* rows with norm >= 1e-4 anywhere (== rank unless the reference's first-small-row cut dropped rows) */
} orc_update_info_t;
void orc_updater_set_rank_rule(int full_info);

void orc_updater_cfg_init(orc_updater_cfg_t* c, float sigma_px, float sigma_py, const double* T_BC0_rowmajor16);

/* x: 26+7N, P: d x d column-major (d=24+6N). types/offsets/xy: CSR lists of normalized float2 measurements.
 * Optional debug outputs may be NULL:
 *   feat_status[n_feat]: 0 accepted, 1 rejected at init, 2 rejected after LM, 3 rejected by gate
 *   feat_pfinv[3*n_feat], feat_gamma[n_feat]
 *   Hstack (rows_stacked x 6N, row-major), rstack: the stacked system BEFORE compression. */
void orc_updater_update(const orc_updater_cfg_t* c, const double* x, int xdim, const double* P,
                        const uint8_t* types, const int32_t* offsets, const float* xy, int n_feat,
                        double* x_out, double* P_out, orc_update_info_t* info,
                        uint8_t* feat_status, double* feat_pfinv, double* feat_gamma,
                        double* Hstack, double* rstack);

/* ---------- host stages (oracle/filter.c) ---------- */
typedef struct {
    double gravity, small_angle, sigma_g, sigma_wg, sigma_a, sigma_wa;
} orc_imu_cfg_t;

/* PreIntegrator::propagate, PreIntegrator.cc:51-194.  x,P in -> x_out (same dim), P_out. */
void orc_propagate(const orc_imu_cfg_t* c, const double* x, int xdim, const double* P,
                   const double* imu, int n_imu, double* x_out, double* P_out);

/* System.cc:280-365: clone augmentation (or window slide) then composition.  x/P are resized in place:
 * buffers must hold 26+7*window and (24+6*window)^2 doubles.  n_clones in/out.  pose_out = [pGk(3), qkG(4)]. */
void orc_augment_compose(double* x, double* P, int* n_clones, int window, int do_augment, double* pose_out);

/* System::initialize, System.cc:115-170 */
void orc_initialize(const orc_imu_cfg_t* c, double imu_rate, const double* w, const double* a, int n_imu_data,
                    int enable_alignment, double* x26, double* P24);

/* Numerics.h helpers, exported for unit tests */
void orc_quat_mul(const double* q1, const double* q2, double* out);   /* Numerics.h:30-63 */
void orc_quat_to_rot(const double* q, double* R_rowmajor);            /* Numerics.h:111-120 */
void orc_rot_to_quat(const double* R_rowmajor, double* q);            /* Numerics.h:126-167 */
double orc_chi2_95(int dof);                                          /* Numerics.h:173-224 table */

#ifdef __cplusplus
}
#endif
#endif
