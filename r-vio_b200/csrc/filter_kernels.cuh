// filter_kernels.cuh -- parameter blocks of the device-resident filter stages (see filter.cu).
#pragma once
#include <cuda_runtime.h>
#include "tracker_kernels.cuh"

namespace rvio {

struct ImuConsts { double gravity, small_angle, sigma_g, sigma_wg, sigma_a, sigma_wa; };

struct PropagateParams {
    const double* x_in; const double* P_in; int xdim, d;
    const double* imu; int n_imu;         // device, n_imu x 8
    const int* hdr;                        // optional device frame header {n_imu, n_cand}: overrides n_imu (frame graphs)
    double* x_out; double* P_out;
    ImuConsts c;
};

struct AugmentParams {
    double* x;                             // in place (capacity 26+7*window)
    const double* P_in; double* P_out;     // d x d  ->  d' x d'
    int d, N, window, do_augment;
    double* pose_out;                      // [pGk(3), qkG(4)]
};

constexpr int kFindNewerCellCap = 128;

struct FindNewerParams {
    TrackerBuffers B;
    const float2* cand; int n_cand;        // detector output (device); n_cand is the capacity when n_cand_dev is set
    const int* n_cand_dev;                 // optional: the count lives on the device (frame header / device detector)
    int raw;                               // 1: candidates are already FindNewer-filtered
    int W, H, gc, gr, offx, offy, max_per_block;
    float bx, by, min_dist;
    CamParams cam;
};

int launch_propagate(cudaStream_t s, const PropagateParams& p);
int launch_augment_compose(cudaStream_t s, const AugmentParams& p);
int launch_find_newer_refill(cudaStream_t s, const FindNewerParams& p);

}  // namespace rvio
