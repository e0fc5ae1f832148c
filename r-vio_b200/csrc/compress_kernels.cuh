// compress_kernels.cuh -- parameters of the reference compression rule stage (compress.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rvio {

struct RankRuleParams {
    double* red;              // [G (n*n) | z (n) | counters (8)], rewritten in place when the cut discards rows
    int n, world;
    const int* rule_dev;      // 0 = the reference's rule (Updater.cc:515-524), 1 = full information
    double* L;                // n x (n+1) scratch: column j of the lower Cholesky factor of G (= row j of R) and y_j at index N'
    int emit_R;               // also write the kept rows of R (zero padded to n x n) and y: the large-window EKF step is R-form
    double* Rc; double* yc;
    int32_t* rr;              // record: [0] mode (0 off, 1 unchanged, 2 rebuilt, 3 sweep needed, 4 undecided), [1] rows kept,
                              //         [2] N' (columns after the trailing-zero drop), [3] sweep requested, [4] rows >= 1e-4 after the sweep, [5] kept-column limit
};

struct GivensRefParams {
    const double* Hblk; const double* rblk; const int32_t* f_dof;
    int n_feat; const int* n_feat_dev; int n, blk_rows;
    double* red; int32_t* rr;
    double* win;              // window storage in global memory when it does not fit in shared memory
    int emit_R; double* Rc; double* yc;       // as in RankRuleParams
};

// single-CTA R-form EKF step of a small window (k_solve_small_R)
struct SolveSmallRParams {
    const double* Rc; const double* yc;      // kept rows of R (n x n, zero rows where dropped), y
    const double* x; const double* P; int xdim, N, d;
    double sig2;
    const double* gate;                       // counters: gate[0] = accepted features
    double* x_out; double* P_out; int* bad;
};

int compress_configure(int nmax);
int enqueue_solve_small_R(cudaStream_t s, const SolveSmallRParams& q);
int enqueue_update_small(cudaStream_t s, const RankRuleParams& rq, const GivensRefParams& gq, const SolveSmallRParams& q);
size_t givens_window_doubles(int n);
int enqueue_rank_rule(cudaStream_t s, const RankRuleParams& rq, const GivensRefParams& gq, int n);
int enqueue_chol_trsm(cudaStream_t s, const double* S, int n, double* L, double* B, int ldb, int nb, int* bad, const double* gate);

}  // namespace rvio
