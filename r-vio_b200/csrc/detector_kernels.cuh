// detector_kernels.cuh -- parameter blocks of the device detector (see detector.cu).
#pragma once
#include <cuda_runtime.h>
#include "tracker_kernels.cuh"

namespace rvio {

constexpr int kDetKeyCap = 65536;       // local maxima above the quality threshold that are kept (more: overflow flag)
constexpr int kDetCellSlots = 2;        // accepted corners per grid cell: integer pixels in a cell of side ~minDistance that are >= minDistance apart
                                        // can only be the two ends of a diagonal (any axis-aligned pair is at most side - 1 apart)

struct DetCtrl { unsigned max_bits; int n_cand; int n_out; int overflow; };

struct DetParams {
    PyrLevel img;                       // equalised frame (pyramid level 0, reflect-101 border)
    float* eig;                         // W * H
    DetCtrl* ctrl;
    unsigned long long* keys; int key_cap;
    float2* out; int max_corners;
    double quality, min_dist;
    int cell, gw, gh;                   // (unused since the shadow bitmap replaced the minimum-distance grid)
    unsigned* shadow;                   // W x H bit map of the pixels closer than min_dist to a kept corner (global copy: images too large for shared memory)
    int shadow_in_smem, wpr;            // words per bitmap row
    int hw, subpix_iters; double subpix_eps;
    const float* mask;
};

struct Detector {
    int W, H, max_corners;
    float* eig; unsigned long long* keys; DetCtrl* ctrl; float2* out; float* mask; float* h_mask; int mask_hw;
    unsigned* shadow;
};

int detector_create(Detector* D, int W, int H, int max_corners);
void detector_destroy(Detector* D);
int detector_enqueue(Detector* D, cudaStream_t st, const PyrLevel& level0, int s, float min_dist, float quality);

}  // namespace rvio
