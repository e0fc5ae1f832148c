// tracker_kernels.cuh -- device side of the tracker hot path (sm_100a).
//
// Replaces, on the device, what Tracker::track does through OpenCV/Eigen on the host
// (reference src/rvio/Tracker.cc:179-342, src/rvio/Ransac.cc:50-266):
//   k_gray            cvtColor *2GRAY                                   Tracker.cc:183-196
//   k_clahe_lut/apply CLAHE(3.0, 5x5)                                   Tracker.cc:198-202
//   k_pyr_down        pyramid of cv::calcOpticalFlowPyrLK               Tracker.cc:244
//   k_lk              pyramidal LK, one warp per feature, all levels    Tracker.cc:237-244
//                     + undistortPoints in the epilogue                 Tracker.cc:100-132,252-261
//   k_ransac          2-point RANSAC incl. glibc rand() stream          Tracker.cc:264, Ransac.cc:180-247
//   k_bookkeep        track-history bookkeeping / update-list emission  Tracker.cc:271-342
//   k_seed / k_refill first-image seeding and refill                    Tracker.cc:215-233,358-386
//
// Arithmetic contract: integer / float32 results are bit-identical to OpenCV 4.x's SSE2 code path
// (accumulation order documented at k_lk); this translation unit is compiled with -fmad=false and uses
// IEEE-rounded sqrt/div so no operation is contracted or approximated.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvio {

constexpr int kBorder = 16;          // pyramid border (>= LK window 15 + 1 tap)
constexpr int kMaxLevels = 4;        // maxLevel = 3 (Tracker.cc:244)
constexpr int kWin = 15;             // LK window (Tracker.cc:237)
constexpr int kRansacIters = 16;     // RansacModel::nIterations (Ransac.h:52)

struct PyrLevel {
    uint8_t* base;                   // address of pixel (0,0); valid x in [-16, w+16), y in [-16, h+16)
    int pitch, w, h;
};
struct Pyramid {
    PyrLevel lv[kMaxLevels];
    int levels;
};

struct CamParams {                   // float-rounded intrinsics widened to double (Tracker.cc:39-61)
    double fx, fy, cx, cy, ifx, ify, k1, k2, p1, p2, k3;
    int fisheye;                     // Camera.Fisheye: (k1, k2, p1, p2) are the equidistant model's k1..k4 (Tracker.cc:119)
};

// Tracker::UndistortAndNormalize for one point (Tracker.cc:100-132): pixel -> undistorted normalized coordinates.
//   radtan   cv::undistortPoints, 5 fixed-point iterations in double
//   fisheye  cv::fisheye::undistortPoints (OpenCV 4.x): Newton on theta, at most 10 iterations / |fix| < 1e-8,
//            scale = tan(theta) / theta_d; (-1e6, -1e6) when not converged or theta changed sign
// Both translation units that use it are compiled with -fmad=false (the host libraries evaluate these in plain double).
__device__ __forceinline__ void cam_undistort(const CamParams& c, float u_, float v_, float* ox, float* oy)
{
    const double u = (double)u_, v = (double)v_;
    if (c.fisheye) {
        const double pwx = (u - c.cx) / c.fx, pwy = (v - c.cy) / c.fy;
        const double half_pi = 3.1415926535897932384626433832795 / 2.;
        double theta_d = sqrt(pwx * pwx + pwy * pwy);
        theta_d = fmin(fmax(-half_pi, theta_d), half_pi);
        bool converged = false;
        double theta = theta_d, scale = 0.0;
        if (fabs(theta_d) > 1e-8) {
            for (int j = 0; j < 10; ++j) {
                const double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
                const double a = c.k1 * t2, b = c.k2 * t4, cc = c.p1 * t6, d = c.p2 * t8;
                const double fix = (theta * (1 + a + b + cc + d) - theta_d) / (1 + 3 * a + 5 * b + 7 * cc + 9 * d);
                theta = theta - fix;
                if (fabs(fix) < 1e-8) { converged = true; break; }
            }
            scale = tan(theta) / theta_d;
        } else {
            converged = true;
        }
        const bool flipped = (theta_d < 0 && theta > 0) || (theta_d > 0 && theta < 0);
        if (converged && !flipped) { *ox = (float)(pwx * scale); *oy = (float)(pwy * scale); }
        else { *ox = -1000000.0f; *oy = -1000000.0f; }
        return;
    }
    double x = (u - c.cx) * c.ifx, y = (v - c.cy) * c.ify;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; ++j) {
        const double r2 = x * x + y * y;
        const double icdist = 1. / (1 + ((c.k3 * r2 + c.k2) * r2 + c.k1) * r2);
        if (icdist < 0) { x = (u - c.cx) * c.ifx; y = (v - c.cy) * c.ify; break; }
        const double dX = 2 * c.p1 * x * y + c.p2 * (r2 + 2 * x * x);
        const double dY = c.p1 * (r2 + 2 * y * y) + 2 * c.p2 * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    *ox = (float)x;
    *oy = (float)y;
}

// Device-resident scalar state of one tracker (mirrors a handful of Tracker members).
struct TrackerScalars {
    int n_track;        // mnFeatsToTrack
    int n_new;          // size of the set being assembled for the next frame
    int fq_head, fq_n;  // mlFreeIndices ring
    int n_up, n_meas;   // emitted update lists
    int n_cand;         // RANSAC candidates of the last call
    int winner;
    int ransac_ran;     // 1 if hypotheses were evaluated
    int rng_f, rng_b;   // glibc rand() feedback pointers
    int pad;
    int rng_r[34];      // glibc rand() state words
};

struct TrackerBuffers {
    int F, Fu, Lmax, Lmin, hist_cap;
    float2* feats;      // mvFeatsToTrack (pixels)
    int* slots;         // mvInlierIndices
    float2* pts1;       // mPoints1ForRansac (x,y; z == 1)
    float2* feats_new; int* slots_new; float2* pts1_new;
    float2* lk;         // vFeatsTracked
    float2* un;         // vFeatsUndistNorm
    uint8_t* status;    // LK status
    uint8_t* flags;     // after RANSAC
    float2* hist;       // mvlTrackingHistory: F rings of hist_cap
    int* hist_head; int* hist_len;
    int* freeq;         // ring of capacity F+1
    uint8_t* up_types; int* up_off; float2* up_xy;
    int* cand;          // candidate index list (scratch, F)
    int* two_points;    // 32
    int* n_inliers;     // 16
    double* hyp;        // 16*9
    TrackerScalars* sc;
};

}  // namespace rvio
