// updater_kernels.cuh -- device side of the MSCKF measurement update (sm_100a), float64.
//
// Replaces Updater::update (reference src/rvio/Updater.cc:72-628):
//   k_feature       one CTA per feature: relative-pose chain (:118-141), inverse-depth LM triangulation (:146-269),
//                   Jacobian blocks (:271-368), left-nullspace projection (:370-402), chi^2 gate (:404-455)
//   k_dmma_hp + k_gate (n >= 96)   the gate's H~ Pcc of all features as one tall FP64 tensor-core GEMM, then the per-feature decision
//   k_gram / k_gram_dmma           the stacked system's normal terms G = H^T H, z = H^T r and the per-class information
//   compress.cu                    which rows the reference's compression keeps (Updater.cc:474-536) and the EKF step
//                                  (Updater.cc:540-619) on those rows:  S = R Pcc R^T + s^2 I,  K = P R^T S^-1, Joseph form
//                                  == dx = Y^T y~,  P+ = P - Y^T Y  with  Y = L^-1 [R P[c,:] | y],  S = L L^T
// See DESIGN.md for the kernels, the rank rule and the remaining deliberate deviations.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvio {

struct UpdaterConsts {
    double sigma, sig2;
    double Ric[9], tic[3], Rci[9], tci[3];
};

// shared-memory layout of k_feature (offsets in doubles), computed on the host from the capacities
struct FeatLayout {
    int Lc, Pc, Mc, Wc, Dc;          // capacities: track length, phases, rows, clone columns, dof
    int o_relI, o_RI, o_RC, o_tC, o_HRR, o_SUB, o_Hf, o_r, o_v, o_Hx, o_S, o_T, o_meas;
    int total_bytes;
};

struct FeatureParams {
    const double* x; int xdim; int N;           // state, clone count
    const double* P; int d;                     // covariance, column-major
    const uint8_t* types; const int32_t* offsets; const float2* xy; int n_feat;
    const int* n_feat_dev;                      // optional device-resident count (fused path); grid covers capacity
    int rank, world;                            // feature sharding (f % world == rank)
    const double* chi2;                         // 500-entry table
    // outputs
    uint8_t* f_status; double* f_pfinv; double* f_gamma; int32_t* f_dof; int32_t* f_c0; int32_t* f_wc; double* f_fro2;
    double* Hblk; double* rblk;                 // per-feature projected blocks: [f][Mc][n], [f][Mc]
    int gate_mode;                              // 0: chi^2 gate inside k_feature; 1: publish every triangulated block, k_dmma_hp + k_gate decide
    int32_t* f_pend;                            // gate_mode 1: dof of the block waiting for its gate
    int blk_rows;                               // Mc
    UpdaterConsts c;
    FeatLayout lay;
};

}  // namespace rvio
