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
};

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
