// updater.cu -- CUDA kernels + C ABI of the MSCKF update half of the hot path (see updater_kernels.cuh).
#include <stdlib.h>
#include "common.cuh"
#include "tracker_kernels.cuh"
#include "updater_kernels.cuh"
#include "compress_kernels.cuh"
#include "tile_cholesky.cuh"
#include "chi2_table.h"

#include <math.h>
#include <new>
#include <vector>

namespace rvio {

// tracker.cu accessors (same shared library) for the fused path
const TrackerBuffers* tracker_buffers(const rvio_tracker* t);
int tracker_update_counts(const rvio_tracker* t, int* n_meas);
int tracker_device(const rvio_tracker* t);
cudaStream_t tracker_stream(const rvio_tracker* t);

// ================================================================================================
// small fp64 device helpers (Numerics.h:30-167 restated)
// ================================================================================================
__device__ __forceinline__ void d_m3mul(const double* A, const double* B, double* C)
{
    double T[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
    for (int i = 0; i < 9; ++i) C[i] = T[i];
}
__device__ __forceinline__ void d_m3v(const double* A, const double* v, double* o)
{
    const double t0 = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
    const double t1 = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
    const double t2 = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
    o[0] = t0; o[1] = t1; o[2] = t2;
}
__device__ __forceinline__ void d_quat_to_rot(const double* q, double* R)   // I - 2w[q x] + 2[q x]^2
{
    const double qx[9] = {0, -q[2], q[1], q[2], 0, -q[0], -q[1], q[0], 0};
    double qx2[9];
    d_m3mul(qx, qx, qx2);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        R[i] = I - 2 * q[3] * qx[i] + 2 * qx2[i];
    }
}
__device__ __forceinline__ void d_quat_mul(const double* q1, const double* q2, double* out)
{
    double q[4];
    q[0] = q1[3] * q2[0] + q1[2] * q2[1] - q1[1] * q2[2] + q1[0] * q2[3];
    q[1] = -q1[2] * q2[0] + q1[3] * q2[1] + q1[0] * q2[2] + q1[1] * q2[3];
    q[2] = q1[1] * q2[0] - q1[0] * q2[1] + q1[3] * q2[2] + q1[2] * q2[3];
    q[3] = -q1[0] * q2[0] - q1[1] * q2[1] - q1[2] * q2[2] + q1[3] * q2[3];
    const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double sg = (q[3] / nrm < 0) ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = sg * (q[i] / nrm);
}
__device__ __forceinline__ void d_rot_to_quat(const double* R, double* q)   // Breckenridge, Numerics.h:126-167
{
    const double T = R[0] + R[4] + R[8];
    if (R[0] > T && R[0] > R[4] && R[0] > R[8]) {
        q[0] = sqrt((1 + 2 * R[0] - T) / 4);
        const double s = 1 / (4 * q[0]);
        q[1] = s * (R[1] + R[3]); q[2] = s * (R[2] + R[6]); q[3] = s * (R[5] - R[7]);
    } else if (R[4] > T && R[4] > R[0] && R[4] > R[8]) {
        q[1] = sqrt((1 + 2 * R[4] - T) / 4);
        const double s = 1 / (4 * q[1]);
        q[0] = s * (R[1] + R[3]); q[2] = s * (R[5] + R[7]); q[3] = s * (R[6] - R[2]);
    } else if (R[8] > T && R[8] > R[0] && R[8] > R[4]) {
        q[2] = sqrt((1 + 2 * R[8] - T) / 4);
        const double s = 1 / (4 * q[2]);
        q[0] = s * (R[2] + R[6]); q[1] = s * (R[5] + R[7]); q[3] = s * (R[1] - R[3]);
    } else {
        q[3] = sqrt((1 + T) / 4);
        const double s = 1 / (4 * q[3]);
        q[0] = s * (R[5] - R[7]); q[1] = s * (R[6] - R[2]); q[2] = s * (R[1] - R[3]);
    }
    const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double sg = (q[3] / nrm < 0) ? -1.0 : 1.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = sg * (q[i] / nrm);
}
__device__ __forceinline__ void d_set_dir(double phi, double psi, double* e, double* J /* 3x2 */)
{
    double sp, cp, ss, cs;
    sincos(phi, &sp, &cp);
    sincos(psi, &ss, &cs);
    e[0] = cp * ss; e[1] = sp; e[2] = cp * cs;
    J[0] = -sp * ss; J[1] = cp * cs;
    J[2] = cp; J[3] = 0;
    J[4] = -sp * cs; J[5] = -cp * ss;
}
__device__ __forceinline__ void d_hproj(const double* h, double* Hp /* 2x3 */)
{
    const double iz = 1 / h[2], iz2 = 1 / (h[2] * h[2]);
    Hp[0] = iz; Hp[1] = 0; Hp[2] = -h[0] * iz2;
    Hp[3] = 0; Hp[4] = iz; Hp[5] = -h[1] * iz2;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return __shfl_sync(0xffffffffu, v, 0);       // identical bits in every lane
}

// 3x3 solve by Gaussian elimination with partial pivoting (LM normal equations, Updater.cc:239)
__device__ __forceinline__ void d_solve3(const double* A_, const double* b_, double* x)
{
    double A[9], b[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) A[i] = A_[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) b[i] = b_[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int p = k;
        double mx = fabs(A[3 * k + k]);
        for (int i = k + 1; i < 3; ++i)
            if (fabs(A[3 * i + k]) > mx) { mx = fabs(A[3 * i + k]); p = i; }
        if (p != k) {
            for (int j = 0; j < 3; ++j) { const double t = A[3 * k + j]; A[3 * k + j] = A[3 * p + j]; A[3 * p + j] = t; }
            const double t = b[k]; b[k] = b[p]; b[p] = t;
        }
        const double piv = A[3 * k + k];
        for (int i = k + 1; i < 3; ++i) {
            const double f = A[3 * i + k] / piv;
            for (int j = k; j < 3; ++j) A[3 * i + j] -= f * A[3 * k + j];
            b[i] -= f * b[k];
        }
    }
    x[2] = b[2] / A[8];
    x[1] = (b[1] - A[5] * x[2]) / A[4];
    x[0] = (b[0] - A[1] * x[1] - A[2] * x[2]) / A[0];
}

// ================================================================================================
// k_feature: one CTA (128 threads) per feature
// ================================================================================================
#ifdef RVIO_B200_PHASE_CLOCKS
// (profiling build only) phase clocks of the k_feature CTA with the longest track of the launch
__device__ long long g_feat_clk[16];
#define FEAT_CLK(k) do { if (threadIdx.x == 0) fclk[k] = clock64(); } while (0)
#else
#define FEAT_CLK(k) do { } while (0)
#endif
constexpr int kFeatThreads = 256;
constexpr int kSolveSmallMaxClones = 14;     // n = 84 (the EuRoC default window): the single-CTA step (k_update_small) fits in 227 KB of shared memory

__global__ void __launch_bounds__(kFeatThreads) k_feature(FeatureParams P)
{
    rvio::pdl_wait(); rvio::pdl_trigger();      // programmatic dependent launch (common.cuh): nothing above touches memory
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_pf[4];          // phi, psi, rho, (unused)
    __shared__ int s_flag[4];           // [0] reject code, [1] Nc
    __shared__ double s_hh[4];          // householder: beta, alpha
    __shared__ double s_fro[kFeatThreads / 32];
    __shared__ double s_lm[10][33];     // LM normal-equation terms per measurement lane (+ the totals in column 32)

    const int f = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef RVIO_B200_PHASE_CLOCKS
    long long fclk[12];
    for (int k = 0; k < 12; ++k) fclk[k] = 0;
#endif
    FEAT_CLK(0);
    const int n_feat = P.n_feat_dev ? *P.n_feat_dev : P.n_feat;
    if (f >= n_feat) return;
    if (f % P.world != P.rank) return;   // feature sharding: this rank owns f % world == rank
    const FeatLayout& Y = P.lay;
    const UpdaterConsts& C = P.c;
    double* relI = sm + Y.o_relI; double* RI = sm + Y.o_RI; double* RC = sm + Y.o_RC; double* tC = sm + Y.o_tC;
    double* HRR = sm + Y.o_HRR; double* SUB = sm + Y.o_SUB; double* Hf = sm + Y.o_Hf; double* rr = sm + Y.o_r;
    double* vv = sm + Y.o_v; double* Hx = sm + Y.o_Hx; double* S = sm + Y.o_S; double* Tc = sm + Y.o_T;
    float2* meas = reinterpret_cast<float2*>(sm + Y.o_meas);

    const int N = P.N, n = 6 * N;
    const int off0 = P.offsets[f];
    int L = P.offsets[f + 1] - off0;
    const int type = P.types[f];
    const int phases_full = L - 1;

    if (tid == 0) {
        P.f_status[f] = 0; P.f_dof[f] = 0; P.f_c0[f] = 0; P.f_wc[f] = 0;
        if (P.f_pend) P.f_pend[f] = 0;
        P.f_gamma[f] = nan("");
        P.f_pfinv[3 * f] = P.f_pfinv[3 * f + 1] = P.f_pfinv[3 * f + 2] = nan("");
        s_flag[0] = 0;
    }
    // capacity / consistency guards (the reference would read out of bounds here)
    if (L < 2 || L > Y.Lc || phases_full > N) {
        if (tid == 0) P.f_status[f] = 1;
        return;
    }
    for (int i = tid; i < L; i += kFeatThreads) meas[i] = P.xy[off0 + i];

    // ---- relative-pose chain, Updater.cc:118-132 (serial in i)
    const double* rel = (type == '1') ? (P.x + P.xdim - 7 * phases_full) : (P.x + 26);
    if (warp == 0) {
        // relI_i = A_i o A_(i-1) o ... o A_0 with A_i = (q_i, -R(q_i) p_i) and (q, t) o (q', t') = (q q', R(q) t' + t): rigid
        // transforms compose associatively, so the chain is an inclusive scan over the lanes (log2 steps instead of L - 1)
        const bool act = lane < phases_full;
        double q[4] = {0, 0, 0, 1}, t[3] = {0, 0, 0};
        if (act) {
            double R0[9], tt[3];
            for (int k = 0; k < 4; ++k) q[k] = rel[7 * lane + k];
            d_quat_to_rot(q, R0);
            d_m3v(R0, rel + 7 * lane + 4, tt);
            t[0] = -tt[0]; t[1] = -tt[1]; t[2] = -tt[2];
        }
        for (int sft = 1; sft < phases_full; sft <<= 1) {
            double q2[4], t2[3];
#pragma unroll
            for (int k = 0; k < 4; ++k) q2[k] = __shfl_up_sync(0xffffffffu, q[k], sft);
#pragma unroll
            for (int k = 0; k < 3; ++k) t2[k] = __shfl_up_sync(0xffffffffu, t[k], sft);
            if (act && lane >= sft) {
                double R[9], r[3], qn[4];
                d_quat_to_rot(q, R);
                d_m3v(R, t2, r);
                t[0] += r[0]; t[1] += r[1]; t[2] += r[2];
                d_quat_mul(q, q2, qn);
                for (int k = 0; k < 4; ++k) q[k] = qn[k];
            }
        }
        if (act) {
            for (int k = 0; k < 4; ++k) relI[7 * lane + k] = q[k];
            for (int k = 0; k < 3; ++k) relI[7 * lane + 4 + k] = t[k];
        }
    }
    __syncthreads();
    FEAT_CLK(1);
    // ---- camera poses, Updater.cc:134-141 (parallel in i)
    for (int i = tid; i < phases_full; i += kFeatThreads) {
        double R[9], T[9], M[9], qC[4], a[3], b[3];
        d_quat_to_rot(relI + 7 * i, R);
        for (int k = 0; k < 9; ++k) RI[9 * i + k] = R[k];
        d_m3mul(C.Rci, R, T);
        d_m3mul(T, C.Ric, M);
        d_rot_to_quat(M, qC);
        d_quat_to_rot(qC, M);
        for (int k = 0; k < 9; ++k) RC[9 * i + k] = M[k];
        d_m3v(T, C.tic, a);
        d_m3v(C.Rci, relI + 7 * i + 4, b);
        for (int k = 0; k < 3; ++k) tC[3 * i + k] = a[k] + b[k] + C.tci[k];
    }
    __syncthreads();

    FEAT_CLK(2);
    // ---- inverse-depth initialisation + LM (warp 0), Updater.cc:143-269
    if (warp == 0) {
        const float2 m0 = meas[0];
        double phi = atan2((double)m0.y, sqrt((double)m0.x * (double)m0.x + 1));
        double psi = atan2((double)m0.x, 1.0);
        double rho = 0.;
        int reject = 0;
        if (fabs(phi) > .5 * 3.14 || fabs(psi) > .5 * 3.14) reject = 1;
        if (!reject) {
            double e[3], J[6];
            d_set_dir(phi, psi, e, J);
            const double rinv = 1. / (C.sigma * C.sigma);
            double lambda = 0.01, lastCost = INFINITY;
            for (int it = 0; it < 10; ++it) {
                double a00 = 0, a01 = 0, a02 = 0, a11 = 0, a12 = 0, a22 = 0, g0 = 0, g1 = 0, g2 = 0, cost = 0;
                for (int i = lane; i < L; i += 32) {
                    double h[3], Hp[6], Hm[6];
                    if (i == 0) {
                        h[0] = e[0]; h[1] = e[1]; h[2] = e[2];
                        d_hproj(h, Hp);
                        for (int a = 0; a < 2; ++a) {
                            Hm[3 * a] = Hp[3 * a] * J[0] + Hp[3 * a + 1] * J[2] + Hp[3 * a + 2] * J[4];
                            Hm[3 * a + 1] = Hp[3 * a] * J[1] + Hp[3 * a + 1] * J[3] + Hp[3 * a + 2] * J[5];
                            Hm[3 * a + 2] = 0;
                        }
                    } else {
                        const double* Rc = RC + 9 * (i - 1);
                        const double* tc = tC + 3 * (i - 1);
                        d_m3v(Rc, e, h);
                        h[0] += rho * tc[0]; h[1] += rho * tc[1]; h[2] += rho * tc[2];
                        d_hproj(h, Hp);
                        double HR[6];
                        for (int a = 0; a < 2; ++a)
                            for (int b = 0; b < 3; ++b) HR[3 * a + b] = Hp[3 * a] * Rc[b] + Hp[3 * a + 1] * Rc[3 + b] + Hp[3 * a + 2] * Rc[6 + b];
                        for (int a = 0; a < 2; ++a) {
                            Hm[3 * a] = HR[3 * a] * J[0] + HR[3 * a + 1] * J[2] + HR[3 * a + 2] * J[4];
                            Hm[3 * a + 1] = HR[3 * a] * J[1] + HR[3 * a + 1] * J[3] + HR[3 * a + 2] * J[5];
                            Hm[3 * a + 2] = Hp[3 * a] * tc[0] + Hp[3 * a + 1] * tc[1] + Hp[3 * a + 2] * tc[2];
                        }
                    }
                    const float ptx = (float)(h[0] / h[2]), pty = (float)(h[1] / h[2]);
                    const float2 mi = meas[i];
                    const double e0 = (double)(mi.x - ptx), e1 = (double)(mi.y - pty);    // float subtraction (cv::Point2f)
                    cost += (e0 * rinv) * e0 + (e1 * rinv) * e1;
                    const double h00 = Hm[0] * rinv, h01 = Hm[1] * rinv, h02 = Hm[2] * rinv;
                    const double h10 = Hm[3] * rinv, h11 = Hm[4] * rinv, h12 = Hm[5] * rinv;
                    a00 += h00 * Hm[0] + h10 * Hm[3]; a01 += h00 * Hm[1] + h10 * Hm[4]; a02 += h00 * Hm[2] + h10 * Hm[5];
                    a11 += h01 * Hm[1] + h11 * Hm[4]; a12 += h01 * Hm[2] + h11 * Hm[5]; a22 += h02 * Hm[2] + h12 * Hm[5];
                    g0 += h00 * e0 + h10 * e1; g1 += h01 * e0 + h11 * e1; g2 += h02 * e0 + h12 * e1;
                }
                // ten sums over the measurements: through shared memory, lane v adds up quantity v over the (few) measurement
                // lanes in index order -- one short dependent chain instead of ten 5-stage shuffle reductions
                {
                    const int nl = min(L, 32);
                    __syncwarp();
                    if (lane < nl) {
                        s_lm[0][lane] = a00; s_lm[1][lane] = a01; s_lm[2][lane] = a02; s_lm[3][lane] = a11; s_lm[4][lane] = a12;
                        s_lm[5][lane] = a22; s_lm[6][lane] = g0; s_lm[7][lane] = g1; s_lm[8][lane] = g2; s_lm[9][lane] = cost;
                    }
                    __syncwarp();
                    if (lane < 10) {
                        double acc = 0;
                        for (int i = 0; i < nl; ++i) acc += s_lm[lane][i];
                        s_lm[lane][32] = acc;
                    }
                    __syncwarp();
                    a00 = s_lm[0][32]; a01 = s_lm[1][32]; a02 = s_lm[2][32]; a11 = s_lm[3][32]; a12 = s_lm[4][32];
                    a22 = s_lm[5][32]; g0 = s_lm[6][32]; g1 = s_lm[7][32]; g2 = s_lm[8][32]; cost = s_lm[9][32];
                }
                if (cost <= lastCost) {
                    double A[9] = {a00, a01, a02, a01, a11, a12, a02, a12, a22};
                    const double g[3] = {g0, g1, g2};
                    A[0] += lambda * A[0]; A[4] += lambda * A[4]; A[8] += lambda * A[8];
                    double dp[3];
                    d_solve3(A, g, dp);
                    phi += dp[0]; psi += dp[1]; rho += dp[2];
                    d_set_dir(phi, psi, e, J);
                    if (fabs(lastCost - cost) < 1e-6 && dp[2] < 1e-6) break;
                    lambda *= .1;
                    lastCost = cost;
                } else {
                    lambda *= 10;
                    lastCost = cost;
                }
            }
            if (fabs(phi) > .5 * 3.14 || fabs(psi) > .5 * 3.14 || isinf(rho) || rho < 0) reject = 2;
        }
        if (lane == 0) {
            s_pf[0] = phi; s_pf[1] = psi; s_pf[2] = rho;
            s_flag[0] = reject;
            if (!(reject == 1)) { P.f_pfinv[3 * f] = phi; P.f_pfinv[3 * f + 1] = psi; P.f_pfinv[3 * f + 2] = rho; }
            if (reject) P.f_status[f] = (uint8_t)reject;
        }
    }
    __syncthreads();
    if (s_flag[0]) return;

    FEAT_CLK(3);
    const double phi = s_pf[0], psi = s_pf[1], rho = s_pf[2];
    double e[3], J[6];
    d_set_dir(phi, psi, e, J);
    if (type == '2') L = (L + 1) / 2;                 // Updater.cc:271-275
    const int Mr = 2 * L;                             // rows of this feature's block
    const int wc = 6 * (L - 1);                       // non-zero clone columns of the block
    const int c0 = (type == '1') ? 6 * (N - phases_full) : 0;   // Updater.cc:288-293
    const int ld = Y.Wc;

    // ---- Jacobians, Updater.cc:281-368
    for (int i = tid; i < Mr * ld; i += kFeatThreads) Hx[i] = 0.0;
    double Rice[3];
    d_m3v(C.Ric, e, Rice);
    // per-row quantities (thread i <-> measurement i)
    for (int i = tid; i < L; i += kFeatThreads) {
        double h[3], Hp[6];
        if (i == 0) {
            h[0] = e[0]; h[1] = e[1]; h[2] = e[2];
            d_hproj(h, Hp);
            for (int a = 0; a < 2; ++a) {
                Hf[3 * a] = Hp[3 * a] * J[0] + Hp[3 * a + 1] * J[2] + Hp[3 * a + 2] * J[4];
                Hf[3 * a + 1] = Hp[3 * a] * J[1] + Hp[3 * a + 1] * J[3] + Hp[3 * a + 2] * J[5];
                Hf[3 * a + 2] = 0;
            }
        } else {
            const double* Rc = RC + 9 * (i - 1);
            const double* tc = tC + 3 * (i - 1);
            d_m3v(Rc, e, h);
            h[0] += rho * tc[0]; h[1] += rho * tc[1]; h[2] += rho * tc[2];
            d_hproj(h, Hp);
            double HR[6], HRci[6];
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 3; ++b) {
                    HR[3 * a + b] = Hp[3 * a] * Rc[b] + Hp[3 * a + 1] * Rc[3 + b] + Hp[3 * a + 2] * Rc[6 + b];
                    HRci[3 * a + b] = Hp[3 * a] * C.Rci[b] + Hp[3 * a + 1] * C.Rci[3 + b] + Hp[3 * a + 2] * C.Rci[6 + b];
                }
            const double* R = RI + 9 * (i - 1);
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 3; ++b)
                    HRR[6 * i + 3 * a + b] = HRci[3 * a] * R[b] + HRci[3 * a + 1] * R[3 + b] + HRci[3 * a + 2] * R[6 + b];
            for (int a = 0; a < 2; ++a) {
                Hf[3 * (2 * i + a)] = HR[3 * a] * J[0] + HR[3 * a + 1] * J[2] + HR[3 * a + 2] * J[4];
                Hf[3 * (2 * i + a) + 1] = HR[3 * a] * J[1] + HR[3 * a + 1] * J[3] + HR[3 * a + 2] * J[5];
                Hf[3 * (2 * i + a) + 2] = Hp[3 * a] * tc[0] + Hp[3 * a + 1] * tc[1] + Hp[3 * a + 2] * tc[2];
            }
        }
        const float ptx = (float)(h[0] / h[2]), pty = (float)(h[1] / h[2]);
        const float2 mi = meas[i];
        rr[2 * i] = (double)(mi.x - ptx);
        rr[2 * i + 1] = (double)(mi.y - pty);
    }
    // per-clone 3x6 factors: SUB_j = [ skew(Ric e + rho tic + rho Rj^T tj) Rj^T , -rho R_{j-1}^T ]  (j=0: -rho I)
    for (int j = tid; j < L - 1; j += kFeatThreads) {
        const double* Rj = RI + 9 * j;
        const double* tj = relI + 7 * j + 4;
        double RjT[9], tmp[3], v[3];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) RjT[3 * a + b] = Rj[3 * b + a];
        d_m3v(RjT, tj, tmp);
        for (int k = 0; k < 3; ++k) v[k] = Rice[k] + rho * C.tic[k] + rho * tmp[k];
        const double sk[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
        double blkL[9];
        d_m3mul(sk, RjT, blkL);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                SUB[18 * j + 6 * a + b] = blkL[3 * a + b];
                SUB[18 * j + 6 * a + 3 + b] = (j == 0) ? ((a == b) ? -rho : 0.0) : -rho * RI[9 * (j - 1) + 3 * b + a];
            }
    }
    __syncthreads();
    // 2x6 blocks (i, j<i): Hx[2i..2i+1, 6j..6j+5] = HRR_i * SUB_j
    {
        const int npairs = L * (L - 1) / 2;
        for (int p = tid; p < npairs; p += kFeatThreads) {
            // p -> (i, j), i in 1..L-1, j in 0..i-1 ; p = i(i-1)/2 + j
            int i = (int)((1.0 + sqrt(1.0 + 8.0 * (double)p)) * 0.5);
            while (i * (i - 1) / 2 > p) --i;
            while ((i + 1) * i / 2 <= p) ++i;
            const int j = p - i * (i - 1) / 2;
            const double* A = HRR + 6 * i;
            const double* B = SUB + 18 * j;
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 6; ++b)
                    Hx[(2 * i + a) * ld + 6 * j + b] = A[3 * a] * B[b] + A[3 * a + 1] * B[6 + b] + A[3 * a + 2] * B[12 + b];
        }
    }
    __syncthreads();

    FEAT_CLK(4);
    // ---- left-nullspace projection (Updater.cc:370-402) by Householder reflections on H_f
    if (tid == 0) {
        double s = 0;
        for (int i = 0; i < Mr; ++i) s += Hf[3 * i + 2] * Hf[3 * i + 2];
        s_flag[1] = (sqrt(s) < 1e-4) ? 2 : 3;         // rank-deficient H_f: use two columns
    }
    __syncthreads();
    const int Nc = s_flag[1];
    for (int k = 0; k < Nc; ++k) {
        if (warp == 0) {
            double s = 0;
            for (int i = k + lane; i < Mr; i += 32) s += Hf[3 * i + k] * Hf[3 * i + k];
            s = warp_sum(s);
            const double x0 = Hf[3 * k + k];
            const double nrm = sqrt(s);
            const double alpha = (x0 >= 0) ? -nrm : nrm;
            for (int i = k + lane; i < Mr; i += 32) vv[i] = (i == k) ? (x0 - alpha) : Hf[3 * i + k];
            const double vtv = s - x0 * x0 + (x0 - alpha) * (x0 - alpha);
            if (lane == 0) s_hh[0] = (vtv > 0) ? 2.0 / vtv : 0.0;
        }
        __syncthreads();
        const double beta = s_hh[0];
        // columns: Hx (wc), then remaining Hf columns, then r
        const int ncols = wc + (Nc - k) + 1;
        for (int c = tid; c < ncols; c += kFeatThreads) {
            double* col; int stride;
            if (c < wc) { col = Hx + c; stride = ld; }
            else if (c < wc + (Nc - k)) { col = Hf + k + (c - wc); stride = 3; }
            else { col = rr; stride = 1; }
            double w = 0;
            for (int i = k; i < Mr; ++i) w += vv[i] * col[i * stride];
            w *= beta;
            for (int i = k; i < Mr; ++i) col[i * stride] -= w * vv[i];
        }
        __syncthreads();
    }
    FEAT_CLK(5);
    const int dof = Mr - Nc;
    const double* Hn = Hx + Nc * ld;       // projected block, dof x wc (row stride ld)
    const double* rn = rr + Nc;

    if (P.gate_mode == 1) {
        // large windows: the gate's tall product H_stack * Pcc runs on the tensor cores afterwards (k_dmma_hp, FP64 DMMA) and
        // k_gate decides; here the projected block of EVERY triangulated feature is published, its dof kept aside
        double* Hout = P.Hblk + (size_t)f * P.blk_rows * n;
        double* rout = P.rblk + (size_t)f * P.blk_rows;
        double sq = 0;
        for (int o = tid; o < dof * n; o += kFeatThreads) {
            const int a = o / n, c = o - a * n;
            const int k = c - c0;
            const double hv = (k >= 0 && k < wc) ? Hn[a * ld + k] : 0.0;
            Hout[o] = hv;
            sq += hv * hv;
        }
        for (int a = tid; a < dof; a += kFeatThreads) rout[a] = rn[a];
        sq = warp_sum(sq);
        if (lane == 0) s_fro[warp] = sq;
        __syncthreads();
        if (tid == 0) {
            double t = 0;
            for (int w = 0; w < kFeatThreads / 32; ++w) t += s_fro[w];
            P.f_fro2[f] = t;
            P.f_pend[f] = dof; P.f_c0[f] = c0; P.f_wc[f] = wc;
        }
        return;
    }

    // ---- Mahalanobis gate, Updater.cc:404-422:  S = Hn Pcc Hn^T + s^2 I ; gamma = |r^T S^-1 r|
    for (int i = tid; i < dof * dof; i += kFeatThreads) S[i] = 0.0;
    __syncthreads();
    for (int j0 = 0; j0 < wc; j0 += 32) {
        const int jw = min(32, wc - j0);
        // T[a][jj] = sum_k Hn[a][k] * Pcc[c0+k][c0+j0+jj]   (P symmetric: read the row-contiguous element)
        for (int o = tid; o < dof * 32; o += kFeatThreads) {
            const int a = o >> 5, jj = o & 31;
            double acc = 0;
            if (jj < jw) {
                const double* prow = P.P + (size_t)(24 + c0) * P.d + (24 + c0 + j0 + jj);
                const double* hrow = Hn + a * ld;
                for (int k = 0; k < wc; ++k) acc += hrow[k] * prow[(size_t)k * P.d];
            }
            Tc[o] = acc;
        }
        __syncthreads();
        for (int o = tid; o < dof * dof; o += kFeatThreads) {
            const int a = o / dof, b = o - a * dof;
            const double* trow = Tc + a * 32;
            const double* hrow = Hn + b * ld + j0;
            double acc = 0;
            for (int jj = 0; jj < jw; ++jj) acc += trow[jj] * hrow[jj];
            S[o] += acc;
        }
        __syncthreads();
    }
    FEAT_CLK(6);
    // symmetrise + noise
    for (int o = tid; o < dof * dof; o += kFeatThreads) {
        const int a = o / dof, b = o - a * dof;
        if (a <= b) {
            double v = .5 * (S[a * dof + b] + S[b * dof + a]);
            if (a == b) v = S[a * dof + a] + C.sig2;
            S[a * dof + b] = v;
        }
    }
    __syncthreads();
    for (int o = tid; o < dof * dof; o += kFeatThreads) {
        const int a = o / dof, b = o - a * dof;
        if (a > b) S[a * dof + b] = S[b * dof + a];
    }
    __syncthreads();
    FEAT_CLK(7);
    // Cholesky S = L L^T (lower), all threads; the forward substitution y = L^-1 r rides along (column j of L is applied to
    // the right-hand side in the same step), gamma = y^T y
    for (int i = tid; i < dof; i += kFeatThreads) vv[i] = rn[i];
    double gamma = 0;                                  // (thread 0's copy is the one that counts)
    __syncthreads();
    if (tid == 0) {
        const double dj = S[0];
        const double sd = (dj > 0) ? sqrt(dj) : nan("");
        const double yj = vv[0] / sd;
        s_hh[1] = sd; s_hh[2] = yj;
        gamma += yj * yj;
    }
    __syncthreads();
    for (int j = 0; j < dof; ++j) {
        const double dj = s_hh[1], yj = s_hh[2];
        for (int i = j + tid; i < dof; i += kFeatThreads) {
            const double lij = (i == j) ? dj : S[i * dof + j] / dj;
            S[i * dof + j] = lij;
            if (i > j) vv[i] -= lij * yj;
        }
        __syncthreads();
        const int rem = dof - j - 1;
        for (int o = tid; o < rem * rem; o += kFeatThreads) {
            const int a = j + 1 + o / rem, b = j + 1 + o % rem;
            if (b <= a) S[a * dof + b] -= S[a * dof + j] * S[b * dof + j];
        }
        if (tid == 0 && rem > 0) {                     // thread 0 has just finished the next pivot: prepare the next column here
            const double dn = S[(j + 1) * dof + j + 1];
            const double sd = (dn > 0) ? sqrt(dn) : nan("");
            const double yn = vv[j + 1] / sd;
            s_hh[1] = sd; s_hh[2] = yn;
            gamma += yn * yn;
        }
        __syncthreads();
    }
    if (warp == 0) {
        if (lane == 0) {
            gamma = fabs(gamma);
            P.f_gamma[f] = gamma;
            const bool ok = gamma < P.chi2[dof - 1];
            s_flag[2] = ok ? 1 : 0;
            if (!ok) P.f_status[f] = 3;
            P.f_dof[f] = ok ? dof : 0;
            P.f_c0[f] = c0; P.f_wc[f] = wc;
        }
    }
    __syncthreads();
    FEAT_CLK(8);
    if (!s_flag[2]) return;
    // ---- accepted: publish the projected block (full width n, zero outside [c0, c0+wc))
    double* Hout = P.Hblk + (size_t)f * P.blk_rows * n;
    double* rout = P.rblk + (size_t)f * P.blk_rows;
    double sq = 0;
    for (int o = tid; o < dof * n; o += kFeatThreads) {
        const int a = o / n, c = o - a * n;
        const int k = c - c0;
        const double hv = (k >= 0 && k < wc) ? Hn[a * ld + k] : 0.0;
        Hout[o] = hv;
        sq += hv * hv;
    }
    for (int a = tid; a < dof; a += kFeatThreads) rout[a] = rn[a];
    // ||block||_F^2 (fixed reduction order): the rank rule needs the information each column-support class carries
    sq = warp_sum(sq);
    if (lane == 0) s_fro[warp] = sq;
    __syncthreads();
    if (tid == 0) {
        double t = 0;
        for (int w = 0; w < kFeatThreads / 32; ++w) t += s_fro[w];
        P.f_fro2[f] = t;
    }
#ifdef RVIO_B200_PHASE_CLOCKS
    FEAT_CLK(9);
    if (tid == 0 && (P.offsets[f + 1] - P.offsets[f]) >= 11 && type == '2') {      // a full-length track: the slow kind
        for (int k = 0; k < 10; ++k) g_feat_clk[k] = fclk[k];
    }
#endif
}

// ================================================================================================
// Large windows: the gate's dense product on the tensor cores.
//   k_dmma_hp   T = H_stack * Pcc    ((F * Mc) x n) * (n x n), FP64 tensor-core MMA (mma.sync m8n8k4 .f64 -> DMMA): the
//               one genuinely dense GEMM of the update (5.2 GFLOP at configs[4]); FP64 because the accept / reject
//               decision and the filter state must match the reference to 1e-9 -- tcgen05 has no FP64 kind.
//               CTA tile 128 x 64, K chunks of 32 through a cp.async double buffer, 8 warps x (32 x 32) register tiles.
//   k_gate      per feature: S = T_f H_f^T + s^2 I (only the feature's own column range), Cholesky, Mahalanobis distance,
//               chi^2 decision (Updater.cc:404-455).
// ================================================================================================
constexpr int kGM = 128, kGN = 64, kGK = 32, kGLdA = kGK + 4, kGLdB = kGN + 8;

__device__ __forceinline__ void cp_async16_zfill(void* dst_smem, const void* src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void dmma_8x8x4(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

struct DmmaParams {
    const double* A; int lda;     // M x K row-major
    const double* B; int ldb;     // K x N row-major (row k at B + k * ldb)
    double* C; int ldc;           // M x N row-major
    int M, N, K;
    const int* n_rows_dev; int rows_per;      // optional: M = *n_rows_dev * rows_per (device-side feature count)
};

__global__ void __launch_bounds__(256) k_dmma_hp(DmmaParams Q)
{
    extern __shared__ __align__(16) double gsm[];
    double* As = gsm;                                   // [2][kGM][kGLdA]
    double* Bs = gsm + 2 * kGM * kGLdA;                 // [2][kGK][kGLdB]
    const int M = Q.n_rows_dev ? min(Q.M, *Q.n_rows_dev * Q.rows_per) : Q.M;
    const int m0 = blockIdx.x * kGM, n0 = blockIdx.y * kGN;
    if (m0 >= M) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, t = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0; acc[i][j][1] = 0; }
    const int nk = (Q.K + kGK - 1) / kGK;
    auto load_chunk = [&](int kc, int buf) {
        const int k0 = kc * kGK;
        double* a = As + buf * kGM * kGLdA;
        double* b = Bs + buf * kGK * kGLdB;
        for (int p = tid; p < kGM * (kGK / 2); p += 256) {          // 16-byte pieces of the A tile
            const int r = p / (kGK / 2), q = p - r * (kGK / 2);
            const int row = m0 + r, k = k0 + 2 * q;
            const bool ok = row < M && k < Q.K;
            cp_async16_zfill(a + r * kGLdA + 2 * q, Q.A + (size_t)(ok ? row : 0) * Q.lda + (ok ? k : 0), ok);
        }
        for (int p = tid; p < kGK * (kGN / 2); p += 256) {          // B tile
            const int r = p / (kGN / 2), q = p - r * (kGN / 2);
            const int k = k0 + r, col = n0 + 2 * q;
            const bool ok = k < Q.K && col < Q.N;
            cp_async16_zfill(b + r * kGLdB + 2 * q, Q.B + (size_t)(ok ? k : 0) * Q.ldb + (ok ? col : 0), ok);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_chunk(0, 0);
    for (int kc = 0; kc < nk; ++kc) {
        const int buf = kc & 1;
        if (kc + 1 < nk) { load_chunk(kc + 1, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const double* a = As + buf * kGM * kGLdA + (wm * 32 + g) * kGLdA + t;
        const double* b = Bs + buf * kGK * kGLdB + t * kGLdB + wn * 32 + g;
#pragma unroll
        for (int kk = 0; kk < kGK; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = a[i * 8 * kGLdA + kk];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = b[kk * kGLdB + j * 8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j], af[i], bf[j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = m0 + wm * 32 + i * 8 + g;
        if (row >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = n0 + wn * 32 + j * 8 + 2 * t;
            if (col + 1 < Q.N) *reinterpret_cast<double2*>(Q.C + (size_t)row * Q.ldc + col) = make_double2(acc[i][j][0], acc[i][j][1]);
            else if (col < Q.N) Q.C[(size_t)row * Q.ldc + col] = acc[i][j][0];
        }
    }
}

struct GateParams {
    const double* Hblk; const double* rblk; const double* T;      // [f][Mc][n], [f][Mc], [f][Mc][n]
    const int32_t* f_pend; const int32_t* f_c0; const int32_t* f_wc;
    int n_feat; const int* n_feat_dev; int n, blk_rows, rank, world;
    double sig2; const double* chi2;
    uint8_t* f_status; double* f_gamma; int32_t* f_dof;
};

constexpr int kGateThreads = kBCThreads;      // the blocked Cholesky's CTA shape (tile_cholesky.cuh)
__global__ void __launch_bounds__(kGateThreads, 1) k_gate(GateParams P)
{
    // dynamic shared memory: [tile-packed S (+ one extra tile row for r): 44 tiles][chunk T[64][34]][chunk H[64][34]]; after the
    // product the two chunk buffers hold the full 64 x 64 S (both triangles: the symmetrisation needs S(a,b) and S(b,a))
    extern __shared__ __align__(16) double sg[];
    __shared__ double s_pv[72];
    __shared__ PanelPub s_pub;
    __shared__ int s_bad;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_feat = P.n_feat_dev ? *P.n_feat_dev : P.n_feat;
    if (f >= n_feat || f % P.world != P.rank) return;
    const int dof = P.f_pend[f];
    if (dof <= 0) return;
    const int n = P.n, c0 = P.f_c0[f], wc = P.f_wc[f];
    constexpr int kCLd = 34;                                       // chunk row stride (even: 16-byte fragment loads)
    constexpr int kTileDoubles = 44 * 64;                          // 8 tile rows of the lower triangle + the row of r
    double* Tt = sg;
    double* cT = sg + kTileDoubles;                                // 64 x kCLd: rows of T_f, 32 columns at a time
    double* cH = cT + 64 * kCLd;                                   // 64 x kCLd: rows of H_f
    const double* Hf = P.Hblk + (size_t)f * P.blk_rows * n;
    const double* Tf = P.T + (size_t)f * P.blk_rows * n;
    // S = T_f H_f^T over the feature's column range on the FP64 tensor pipe: 8 x 8 output tiles, warp w owns tiles w, w + 16, ...;
    // 32-column chunks of T_f and H_f staged in shared memory, both fragments of a tile product are one 16-byte load
    // (columns 2t, 2t+1 of row g: the k index is permuted identically)
    const int TR = (dof + 7) >> 3;                                 // <= 8 (dof <= 64)
    constexpr int kW = kGateThreads / 32;
    const int gq = lane >> 2, t4 = lane & 3;
    double acc[4][2];
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc[u][0] = 0; acc[u][1] = 0; }
    for (int o = tid; o < 2 * 64 * kCLd; o += kGateThreads) cT[o] = 0.0;                    // rows >= dof stay zero
    if (tid == 0) s_bad = 0;
    __syncthreads();
    for (int j0 = 0; j0 < wc; j0 += 32) {
        const int jw = min(32, wc - j0);
        {
            // all of a thread's loads of the chunk are requested before the first store (dof <= 64: at most 4 elements per thread)
            double vt[4], vh[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int o = tid + q * kGateThreads;
                const int a = o >> 5, jj = o & 31;
                const bool ok = o < dof * 32 && jj < jw;
                vt[q] = ok ? Tf[(size_t)a * n + c0 + j0 + jj] : 0.0;
                vh[q] = ok ? Hf[(size_t)a * n + c0 + j0 + jj] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int o = tid + q * kGateThreads;
                if (o < dof * 32) { cT[(o >> 5) * kCLd + (o & 31)] = vt[q]; cH[(o >> 5) * kCLd + (o & 31)] = vh[q]; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int tl = warp + kW * u;
            if (tl < TR * TR) {
                const int I = tl / TR, J = tl - I * TR;
                const double* ap = cT + (8 * I + gq) * kCLd + 2 * t4;
                const double* bp = cH + (8 * J + gq) * kCLd + 2 * t4;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const double2 av = *reinterpret_cast<const double2*>(ap + 8 * ks), bv = *reinterpret_cast<const double2*>(bp + 8 * ks);
                    dmma_8x8x4(acc[u], av.x, bv.x);
                    dmma_8x8x4(acc[u], av.y, bv.y);
                }
            }
        }
        __syncthreads();
    }
    // the full S into the (now dead) chunk buffers, row stride 64
    double* Sf = cT;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int tl = warp + kW * u;
        if (tl < TR * TR) {
            const int I = tl / TR, J = tl - I * TR;
            *reinterpret_cast<double2*>(Sf + (8 * I + gq) * 64 + 8 * J + 2 * t4) = make_double2(acc[u][0], acc[u][1]);
        }
    }
    __syncthreads();
    // symmetrise + noise (Updater.cc:417-418) into the tile-packed lower triangle; identity padding; r as the extra row
    TileTri T;
    T.t = Tt; T.tc = TR; T.tr = TR + 1;
    const int ncp = 8 * TR;
    for (int o = tid; o < tile_tri_count(TR, TR + 1) * 64; o += kGateThreads) Tt[o] = 0.0;
    __syncthreads();
    for (int o = tid; o < ncp * ncp; o += kGateThreads) {
        const int a = o / ncp, b = o - a * ncp;
        if (b > a) continue;
        double v;
        if (a < dof) v = (a == b) ? Sf[a * 64 + a] + P.sig2 : .5 * (Sf[b * 64 + a] + Sf[a * 64 + b]);
        else v = (a == b) ? 1.0 : 0.0;
        *T.at(a, b) = v;
    }
    {
        const double* rn = P.rblk + (size_t)f * P.blk_rows;
        for (int j = tid; j < dof; j += kGateThreads) *T.at(ncp, j) = rn[j];
    }
    // Cholesky S = L L^T with r riding along as an extra row (-> y = L^-1 r): blocked, look-ahead, DMMA trailing updates
    SpdPivot piv; piv.bad = &s_bad;
    tile_cholesky<false>(T, 0, piv, nullptr, s_pv, &s_pub, nullptr, nullptr);
    if (warp == 0) {                                               // gamma = |L^-1 r|^2
        double g2 = 0;
        for (int j = lane; j < dof; j += 32) { const double y = *T.at(ncp, j); g2 += y * y; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) g2 += __shfl_xor_sync(0xffffffffu, g2, o);
        if (lane == 0) {
            double gamma = s_bad ? nan("") : fabs(g2);             // a non-positive pivot: not a valid innovation covariance -> rejected
            P.f_gamma[f] = gamma;
            const bool ok = gamma < P.chi2[dof - 1];
            if (!ok) P.f_status[f] = 3;
            P.f_dof[f] = ok ? dof : 0;
        }
    }
}

// ================================================================================================
// normal terms  G = sum_f Hn_f^T Hn_f ,  z = sum_f Hn_f^T rn_f   (deterministic two-stage reduction)
// ================================================================================================
struct GramParams {
    const double* Hblk; const double* rblk; const int32_t* f_dof; const int32_t* f_c0; const int32_t* f_wc; const double* f_fro2;
    int n_feat, n, blk_rows, groups, nt;
    const int* n_feat_dev;   // optional device-resident feature count (fused path)
    double* Gpart;     // [groups][n][n]
    double* zpart;     // [groups][n]
};

// One launch: grid (tiles, groups).  Each CTA accumulates its 32x32 tile of G (and, for diagonal tiles, its 32 entries
// of z) over the features of its group into a partial buffer; the LAST CTA of a tile to finish (atomic ticket) sums the
// partials in fixed group order (deterministic) into the reduce buffer; the last CTA of tile 0 also writes the counters.
// reduce buffer layout: [G (n*n) | z (n) | counters (8) | cls (n+1)]; counters: n_good, rows, rej_init, rej_lm, rej_gate, n_local
#ifdef RVIO_B200_PHASE_CLOCKS
// (profiling build only) wall-clock (globaltimer, ns) span of k_gram over all CTAs and the phases of tile 0's reducing CTA
__device__ unsigned long long g_gram_ns[16];
__device__ __forceinline__ unsigned long long gram_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define GRAM_T(k) do { if (threadIdx.x == 0) g_gram_ns[k] = gram_now(); } while (0)
#else
#define GRAM_T(k) do { } while (0)
#endif
__global__ void __launch_bounds__(256) k_gram(GramParams P, const uint8_t* f_status, int rank, int world, double* red, int* tickets)
{
    rvio::pdl_wait(); rvio::pdl_trigger();      // programmatic dependent launch (common.cuh): nothing above touches memory
#ifdef RVIO_B200_PHASE_CLOCKS
    const unsigned long long t_in = gram_now();
    if (threadIdx.x == 0) { atomicMin(&g_gram_ns[0], t_in); }
#endif
    __shared__ double sA[2][8][33], sB[2][8][33], s_r[2][8];      // double-buffered 8-row chunks of the two column ranges (+ residuals)
    __shared__ int s_list[256], s_ldof[256];                      // features of this group that touch the tile, in index order
    __shared__ int s_wsum[8], s_last, s_cnt[6];
    __shared__ double s_cls[192];
    __shared__ int s_fd[512], s_fc[512];                          // (tail) dof / first column of every feature
    __shared__ double s_ff[512];
    const int ti = blockIdx.x / P.nt, tj = blockIdx.x % P.nt, g = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int n = P.n;
    const int i0 = ti * 32, j0 = tj * 32;
    const int n_feat = P.n_feat_dev ? *P.n_feat_dev : P.n_feat;
    // ---- which of this group's features (f = g, g + groups, ...) touch the tile: one feature per thread, order-preserving
    //      compaction (the accumulation order over the features is part of the deterministic result)
    int n_act = 0;
    for (int base = 0; base * P.groups + g < n_feat; base += 256) {
        const int f = g + (base + tid) * P.groups;
        int dof = 0;
        bool act = false;
        if (f < n_feat) {
            dof = P.f_dof[f];
            if (dof > 0) {
                const int c0 = P.f_c0[f], c1 = c0 + P.f_wc[f];
                act = !(i0 >= c1 || i0 + 32 <= c0 || j0 >= c1 || j0 + 32 <= c0);     // block is zero on this tile otherwise
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, act);
        if (lane == 0) s_wsum[warp] = __popc(m);
        __syncthreads();
        int off = n_act;
        for (int w = 0; w < warp; ++w) off += s_wsum[w];
        int tot = 0;
        for (int w = 0; w < 8; ++w) tot += s_wsum[w];
        if (act) { const int pos = off + __popc(m & ((1u << lane) - 1u)); if (pos < 256) { s_list[pos] = f; s_ldof[pos] = dof; } }
        n_act += tot;
        __syncthreads();
    }
    if (n_act > 256) n_act = 256;                                  // (cannot happen: <= 4096 features over 16 groups)
    double acc[2][2] = {{0, 0}, {0, 0}};
    double zacc = 0;
    {
        // chunk stream (feature li, rows a0 .. a0 + 7): the loads of chunk t + 1 are in flight while chunk t is multiplied
        const int r = tid >> 5, c = tid & 31;                      // 8 rows x 32 columns
        int li = 0, a0 = 0;
        double ra = 0, rb = 0, rr = 0;
        auto fetch = [&](int li_, int a0_) {
            const int f = s_list[li_], dof = s_ldof[li_];
            const double* H = P.Hblk + (size_t)f * P.blk_rows * n;
            const int a = a0_ + r;
            ra = (a < dof && i0 + c < n) ? H[(size_t)a * n + i0 + c] : 0.0;
            rb = (a < dof && j0 + c < n) ? H[(size_t)a * n + j0 + c] : 0.0;
            rr = (c == 0 && a < dof && ti == tj) ? P.rblk[(size_t)f * P.blk_rows + a] : 0.0;
        };
        if (n_act > 0) fetch(0, 0);
        int buf = 0;
        while (li < n_act) {
            sA[buf][r][c] = ra; sB[buf][r][c] = rb;
            if (c == 0) s_r[buf][r] = rr;
            __syncthreads();
            int nli = li, na0 = a0 + 8;
            if (na0 >= s_ldof[li]) { nli = li + 1; na0 = 0; }
            if (nli < n_act) fetch(nli, na0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const double a0v = sA[buf][q][2 * ty], a1v = sA[buf][q][2 * ty + 1];
                const double b0v = sB[buf][q][2 * tx], b1v = sB[buf][q][2 * tx + 1];
                acc[0][0] += a0v * b0v; acc[0][1] += a0v * b1v; acc[1][0] += a1v * b0v; acc[1][1] += a1v * b1v;
            }
            if (ti == tj && tid < 32) {
                // z rows of this tile (a diagonal tile sees every feature that touches these columns); rows past dof are zero
#pragma unroll
                for (int q = 0; q < 8; ++q) zacc += sA[buf][q][tid] * s_r[buf][q];
            }
            li = nli; a0 = na0; buf ^= 1;
        }
    }
#ifdef RVIO_B200_PHASE_CLOCKS
    if (threadIdx.x == 0) { atomicMax(&g_gram_ns[1], gram_now()); }       // products done (latest CTA)
#endif
    double* Gp = P.Gpart + (size_t)g * n * n;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const int i = i0 + 2 * ty + a, j = j0 + 2 * tx + b;
            if (i < n && j < n) Gp[(size_t)i * n + j] = acc[a][b];
        }
    if (ti == tj && tid < 32 && i0 + tid < n) P.zpart[(size_t)g * n + i0 + tid] = zacc;
    __threadfence();
    __syncthreads();
#ifdef RVIO_B200_PHASE_CLOCKS
    if (threadIdx.x == 0) { atomicMax(&g_gram_ns[2], gram_now()); }       // partials written + fence (latest CTA)
#endif
    if (tid == 0) {
        const int t = atomicAdd(&tickets[blockIdx.x], 1);
        s_last = (t == P.groups - 1) ? 1 : 0;
        if (s_last) tickets[blockIdx.x] = 0;           // self-reset for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        // every partial of this thread's four elements (and its z entry) is requested before the first sum: one round trip to
        // L2 instead of one per element (the stores into `red` would otherwise fence the next element's loads)
        const int ng = P.groups;                                   // <= 16 on this path
        const double* __restrict__ Gpart = P.Gpart;
        const double* __restrict__ zpart = P.zpart;
        double v[4][16], vz[16];
        const bool zrow = ti == tj && tid < 32 && i0 + tid < n;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + 2 * ty + (e >> 1), j = j0 + 2 * tx + (e & 1);
            const bool ok = i < n && j < n;
#pragma unroll
            for (int gg = 0; gg < 16; ++gg) v[e][gg] = (ok && gg < ng) ? Gpart[(size_t)gg * n * n + (size_t)i * n + j] : 0.0;
        }
#pragma unroll
        for (int gg = 0; gg < 16; ++gg) vz[gg] = (zrow && gg < ng) ? zpart[(size_t)gg * n + i0 + tid] : 0.0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = i0 + 2 * ty + (e >> 1), j = j0 + 2 * tx + (e & 1);
            if (i < n && j < n) {
                double sum = 0;
#pragma unroll
                for (int gg = 0; gg < 16; ++gg) if (gg < ng) sum += v[e][gg];
                red[(size_t)i * n + j] = sum;
            }
        }
        if (zrow) {
            double sum = 0;
#pragma unroll
            for (int gg = 0; gg < 16; ++gg) if (gg < ng) sum += vz[gg];
            red[(size_t)n * n + i0 + tid] = sum;
        }
    }
    if (blockIdx.x != 0) return;
    GRAM_T(3);
    // ---- (tile 0 only) counters and the information per column-support class
    if (tid < 6) s_cnt[tid] = 0;
    __syncthreads();
    {
        int good = 0, rows = 0, r1 = 0, r2 = 0, r3 = 0, loc = 0;
        for (int f = rank + tid * world; f < n_feat; f += 256 * world) {
            loc++;
            const int st = f_status[f];
            if (st == 0) { good++; rows += P.f_dof[f]; }
            else if (st == 1) r1++;
            else if (st == 2) r2++;
            else r3++;
        }
        if (good) atomicAdd(&s_cnt[0], good);
        if (rows) atomicAdd(&s_cnt[1], rows);
        if (r1) atomicAdd(&s_cnt[2], r1);
        if (r2) atomicAdd(&s_cnt[3], r2);
        if (r3) atomicAdd(&s_cnt[4], r3);
        if (loc) atomicAdd(&s_cnt[5], loc);
    }
    const bool staged = n_feat <= 512;
    if (staged)
        for (int f = tid; f < n_feat; f += 256) { s_fd[f] = P.f_dof[f]; s_fc[f] = P.f_c0[f]; s_ff[f] = P.f_fro2[f]; }
    __syncthreads();
    if (tid == 0) {
        double* c = red + (size_t)n * n + n;
        c[0] = s_cnt[0]; c[1] = s_cnt[1]; c[2] = s_cnt[2]; c[3] = s_cnt[3]; c[4] = s_cnt[4]; c[5] = s_cnt[5]; c[6] = 0; c[7] = 0;
    }
    // cls[c] = sum of ||H_f||_F^2 over the accepted features whose first non-zero column is c (c = 0 for '2' features,
    // 6 (N - (L-1)) for '1' features: multiples of 6).  One warp per class, lanes stride over the features, fixed butterfly:
    // deterministic
    for (int c = tid; c <= n; c += 256) s_cls[c] = 0.0;
    __syncthreads();
    for (int c = 6 * warp; c <= n; c += 6 * 8) {
        double a2 = 0;
        if (staged) {
            for (int f = rank + lane * world; f < n_feat; f += 32 * world)
                if (s_fd[f] > 0 && s_fc[f] == c) a2 += s_ff[f];
        } else {
            for (int f = rank + lane * world; f < n_feat; f += 32 * world)
                if (P.f_dof[f] > 0 && P.f_c0[f] == c) a2 += P.f_fro2[f];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        if (lane == 0) s_cls[c] = a2;
    }
    __syncthreads();
    double* cls = red + (size_t)n * n + n + 8;
    for (int c = tid; c <= n; c += 256) cls[c] = s_cls[c];
    GRAM_T(4);
}

// Normal terms on the tensor cores (large windows): the same two-stage deterministic reduction as k_gram, 64 x 64 tiles of
// the UPPER triangle of G (mirrored by the reducing CTA), the products through FP64 DMMA (A(i,k) = H(k,i), B(k,j) = H(k,j):
// both fragments come from the same k-major row chunk in shared memory).  128 threads = 4 warps x (32 x 32).
constexpr int kGramT = 64, kGramLd = kGramT + 8, kGramRows = 32;
constexpr size_t kGramDmmaSmem = sizeof(double) * (2 * 2 * kGramRows * kGramLd + 2 * kGramRows);      // two [sI | sJ] buffers + residuals
// 8-byte asynchronous copy (zero when !valid)
__device__ __forceinline__ void gram_cp8(double* dst_smem, const double* src, bool valid)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(valid ? 8 : 0) : "memory");
}
// 16-byte asynchronous copy of which only the first n_valid (0, 1, 2) doubles are read; the rest is zero filled
__device__ __forceinline__ void gram_cp16(double* dst_smem, const double* src, int n_valid, const double* safe)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const double* sp = n_valid > 0 ? src : safe;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(sp), "r"(8 * n_valid) : "memory");
}
__global__ void __launch_bounds__(128) k_gram_dmma(GramParams P, const uint8_t* f_status, int rank, int world, double* red, int* tickets)
{
#ifdef RVIO_B200_PHASE_CLOCKS
    if (threadIdx.x == 0) atomicMin(&g_gram_ns[8], gram_now());
#endif
    extern __shared__ __align__(16) double gsm[];                  // two buffers of [sI | sJ] (kGramRows x kGramLd each) + residuals
    __shared__ int s_last;
    __shared__ double s_cls[192];
    __shared__ int s_list[256], s_ldof[256], s_wsum[4];
    const int n = P.n, nt = P.nt, g = blockIdx.y;
    // upper-triangle tile pair from the linear index
    int ti = 0, rem = blockIdx.x;
    while (rem >= nt - ti) { rem -= nt - ti; ++ti; }
    const int tj = ti + rem;
    const int i0 = ti * kGramT, j0 = tj * kGramT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 1, wn = warp & 1, gq = lane >> 2, t = lane & 3;
    constexpr int kBuf = 2 * kGramRows * kGramLd;                  // doubles per buffer
    double* s_r = gsm + 2 * kBuf;                                  // [2][kGramRows]
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[a][b][0] = 0; acc[a][b][1] = 0; }
    double zacc = 0;
    const int n_feat = P.n_feat_dev ? *P.n_feat_dev : P.n_feat;
    // ---- this group's features that touch the tile, in index order (the accumulation order is part of the deterministic result)
    int n_act = 0;
    for (int base = 0; base * P.groups + g < n_feat; base += 128) {
        const int f = g + (base + tid) * P.groups;
        int dof = 0;
        bool act = false;
        if (f < n_feat) {
            dof = P.f_dof[f];
            if (dof > 0) {
                const int c0 = P.f_c0[f], c1 = c0 + P.f_wc[f];
                act = !(i0 >= c1 || i0 + kGramT <= c0 || j0 >= c1 || j0 + kGramT <= c0);
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, act);
        if (lane == 0) s_wsum[warp] = __popc(m);
        __syncthreads();
        int off = n_act, tot = 0;
        for (int w = 0; w < 4; ++w) { if (w < warp) off += s_wsum[w]; tot += s_wsum[w]; }
        if (act) { const int pos = off + __popc(m & ((1u << lane) - 1u)); if (pos < 256) { s_list[pos] = f; s_ldof[pos] = dof; } }
        n_act += tot;
        __syncthreads();
    }
#ifdef RVIO_B200_PHASE_CLOCKS
    if (threadIdx.x == 0) atomicMax(&g_gram_ns[9], gram_now());
#endif
    if (n_act > 256) n_act = 256;                                  // (cannot happen: <= 4096 features over >= 16 groups)
    // ---- chunk stream (feature li, rows a0 .. a0 + 31): chunk t + 1 is on its way (cp.async, 16-byte pieces, zero filled past
    //      the block's rows / the matrix' columns) while chunk t is multiplied
    auto issue = [&](int li_, int a0_, int buf) {
        const int f = s_list[li_], dof = s_ldof[li_];
        const double* H = P.Hblk + (size_t)f * P.blk_rows * n;
        double* bI = gsm + buf * kBuf; double* bJ = bI + kGramRows * kGramLd;
        for (int o = tid; o < kGramRows * (kGramT / 2); o += 128) {
            const int r = o / (kGramT / 2), c = 2 * (o - r * (kGramT / 2));
            const int a = a0_ + r;
            const bool rok = a < dof;
            const double* src = H + (size_t)(rok ? a : 0) * n;
            gram_cp16(bI + r * kGramLd + c, src + i0 + c, rok ? max(0, min(2, n - (i0 + c))) : 0, H);
            if (ti != tj) gram_cp16(bJ + r * kGramLd + c, src + j0 + c, rok ? max(0, min(2, n - (j0 + c))) : 0, H);
        }
        if (ti == tj && tid < kGramRows) {
            const int a = a0_ + tid;
            gram_cp8(s_r + buf * kGramRows + tid, P.rblk + (size_t)f * P.blk_rows + (a < dof ? a : 0), a < dof);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int li = 0, a0 = 0, buf = 0;
    if (n_act > 0) issue(0, 0, 0);
    while (li < n_act) {
        int nli = li, na0 = a0 + kGramRows;
        if (na0 >= s_ldof[li]) { nli = li + 1; na0 = 0; }
        if (nli < n_act) { issue(nli, na0, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const double* bI = gsm + buf * kBuf; const double* bJ = (ti == tj) ? bI : bI + kGramRows * kGramLd;
#pragma unroll
        for (int kk = 0; kk < kGramRows; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) af[a] = bI[(kk + t) * kGramLd + wm * 32 + a * 8 + gq];
#pragma unroll
            for (int b = 0; b < 4; ++b) bf[b] = bJ[(kk + t) * kGramLd + wn * 32 + b * 8 + gq];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) dmma_8x8x4(acc[a][b], af[a], bf[b]);
        }
        if (ti == tj && tid < kGramT) {
            const double* rr = s_r + buf * kGramRows;
#pragma unroll 8
            for (int r = 0; r < kGramRows; ++r) zacc = fma(bI[r * kGramLd + tid], rr[r], zacc);
        }
        __syncthreads();                                           // everyone is done with this buffer before the chunk after next lands in it
        li = nli; a0 = na0; buf ^= 1;
    }
#ifdef RVIO_B200_PHASE_CLOCKS
    if (threadIdx.x == 0) atomicMax(&g_gram_ns[10], gram_now());
#endif
    double* Gp = P.Gpart + (size_t)g * n * n;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = i0 + wm * 32 + a * 8 + gq, j = j0 + wn * 32 + b * 8 + 2 * t;
            if (i < n && j < n) Gp[(size_t)i * n + j] = acc[a][b][0];
            if (i < n && j + 1 < n) Gp[(size_t)i * n + j + 1] = acc[a][b][1];
        }
    if (ti == tj && tid < kGramT && i0 + tid < n) P.zpart[(size_t)g * n + i0 + tid] = zacc;
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int tk = atomicAdd(&tickets[blockIdx.x], 1);
        s_last = (tk == P.groups - 1) ? 1 : 0;
        if (s_last) tickets[blockIdx.x] = 0;           // self-reset for the next launch
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    {
        // all partials of two elements are requested before the first sum (one L2 round trip per pair of elements; the
        // stores into `red` would otherwise fence the next element's loads); sums in group order: deterministic
        const int ng = P.groups;                                   // <= 48
        const double* __restrict__ Gpart = P.Gpart;
        for (int e0 = tid; e0 < kGramT * kGramT; e0 += 2 * 128) {
            double v[2][48];
            int ii[2], jj[2];
            bool ok[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int e = e0 + q * 128;
                ii[q] = i0 + e / kGramT; jj[q] = j0 + e % kGramT;
                ok[q] = e < kGramT * kGramT && ii[q] < n && jj[q] < n && (ti != tj || jj[q] >= ii[q]);      // diagonal tiles: the upper part, mirrored
#pragma unroll
                for (int gg = 0; gg < 48; ++gg) v[q][gg] = (ok[q] && gg < ng) ? Gpart[(size_t)gg * n * n + (size_t)ii[q] * n + jj[q]] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (ok[q]) {
                    double sum = 0;
#pragma unroll
                    for (int gg = 0; gg < 48; ++gg) if (gg < ng) sum += v[q][gg];
                    red[(size_t)ii[q] * n + jj[q]] = sum;
                    red[(size_t)jj[q] * n + ii[q]] = sum;        // mirror: G bitwise symmetric
                }
            }
        }
    }
    if (ti == tj && tid < kGramT && i0 + tid < n) {
        double vz[48];
#pragma unroll
        for (int gg = 0; gg < 48; ++gg) vz[gg] = (gg < P.groups) ? P.zpart[(size_t)gg * n + i0 + tid] : 0.0;
        double sum = 0;
#pragma unroll
        for (int gg = 0; gg < 48; ++gg) if (gg < P.groups) sum += vz[gg];
        red[(size_t)n * n + i0 + tid] = sum;
    }
    if (blockIdx.x != 0) return;
    // ---- (tile 0 only) counters and the information per column-support class
    __shared__ int s_cnt[6];
    if (tid < 6) s_cnt[tid] = 0;
    __syncthreads();
    {
        int good = 0, rows = 0, r1 = 0, r2 = 0, r3 = 0, loc = 0;
        for (int f = rank + tid * world; f < n_feat; f += 128 * world) {
            loc++;
            const int st = f_status[f];
            if (st == 0) { good++; rows += P.f_dof[f]; }
            else if (st == 1) r1++;
            else if (st == 2) r2++;
            else r3++;
        }
        if (good) atomicAdd(&s_cnt[0], good);
        if (rows) atomicAdd(&s_cnt[1], rows);
        if (r1) atomicAdd(&s_cnt[2], r1);
        if (r2) atomicAdd(&s_cnt[3], r2);
        if (r3) atomicAdd(&s_cnt[4], r3);
        if (loc) atomicAdd(&s_cnt[5], loc);
    }
    // feature tables staged in the (now dead) chunk buffers: dof, first column, ||H_f||_F^2
    const bool staged = n_feat <= 4096;                            // 4096 x 16 B = 64 KB <= the two chunk buffers
    int* s_fd = reinterpret_cast<int*>(gsm);
    int* s_fc = s_fd + 4096;
    double* s_ff = gsm + 4096;                                     // (after the two int tables)
    if (staged)
        for (int f = tid; f < n_feat; f += 128) { s_fd[f] = P.f_dof[f]; s_fc[f] = P.f_c0[f]; s_ff[f] = P.f_fro2[f]; }
    for (int c = tid; c <= n; c += 128) s_cls[c] = 0.0;
    __syncthreads();
    if (tid == 0) {
        double* c = red + (size_t)n * n + n;
        c[0] = s_cnt[0]; c[1] = s_cnt[1]; c[2] = s_cnt[2]; c[3] = s_cnt[3]; c[4] = s_cnt[4]; c[5] = s_cnt[5]; c[6] = 0; c[7] = 0;
    }
    // cls[c] = sum of ||H_f||_F^2 over the accepted features whose first non-zero column is c (multiples of 6): one warp per
    // class, lanes stride over the features, fixed butterfly -> deterministic
    for (int c = 6 * warp; c <= n; c += 6 * 4) {
        double a2 = 0;
        if (staged) {
            for (int f = rank + lane * world; f < n_feat; f += 32 * world)
                if (s_fd[f] > 0 && s_fc[f] == c) a2 += s_ff[f];
        } else {
            for (int f = rank + lane * world; f < n_feat; f += 32 * world)
                if (P.f_dof[f] > 0 && P.f_c0[f] == c) a2 += P.f_fro2[f];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        if (lane == 0) s_cls[c] = a2;
    }
    __syncthreads();
    double* cls = red + (size_t)n * n + n + 8;
    for (int c = tid; c <= n; c += 128) cls[c] = s_cls[c];
}

// ================================================================================================
// generic fp64 GEMM (strided operands), Gauss-Jordan solve, finalisation
// ================================================================================================
struct GemmParams {
    int M, N, K;
    const double* A; long long ars, acs;      // A(i,k) = A[i*ars + k*acs]
    const double* B; long long brs, bcs;      // B(k,j)
    const double* C0; long long c0rs, c0cs;   // optional addend
    double* C; long long crs, ccs;
    double alpha, beta, diag_add;
    const double* gate;       // optional: skip the launch when gate[0] <= 2 (fewer than 3 accepted features)
};

__global__ void __launch_bounds__(256) k_dgemm(GemmParams P)
{
    __shared__ double sA[16][33], sB[16][33];
    if (P.gate && !(P.gate[0] > 2.0)) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
    double acc[2][2] = {{0, 0}, {0, 0}};
    for (int k0 = 0; k0 < P.K; k0 += 16) {
        for (int o = threadIdx.x; o < 16 * 32; o += 256) {
            const int kk = o >> 5, c = o & 31;
            const int k = k0 + kk;
            sA[kk][c] = (k < P.K && i0 + c < P.M) ? P.A[(long long)(i0 + c) * P.ars + (long long)k * P.acs] : 0.0;
            sB[kk][c] = (k < P.K && j0 + c < P.N) ? P.B[(long long)k * P.brs + (long long)(j0 + c) * P.bcs] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const double a0 = sA[kk][2 * ty], a1 = sA[kk][2 * ty + 1];
            const double b0 = sB[kk][2 * tx], b1 = sB[kk][2 * tx + 1];
            acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
        }
        __syncthreads();
    }
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const int i = i0 + 2 * ty + a, j = j0 + 2 * tx + b;
            if (i < P.M && j < P.N) {
                double v = P.alpha * acc[a][b];
                if (P.C0) v += P.beta * P.C0[(long long)i * P.c0rs + (long long)j * P.c0cs];
                if (i == j) v += P.diag_add;
                P.C[(long long)i * P.crs + (long long)j * P.ccs] = v;
            }
        }
}

// Solves M Y = R in place (Y overwrites R) by Gauss-Jordan elimination with partial pivoting.
// M: n x n row-major, R: n x m row-major.  Single CTA of 1024 threads; operands stay in L2.
__global__ void __launch_bounds__(1024) k_gauss_jordan(double* M, double* R, int n, int m, int* singular, const double* gate)
{
    if (gate && !(gate[0] > 2.0)) return;
    __shared__ double s_val[32];
    __shared__ int s_idx[32];
    __shared__ int s_piv;
    __shared__ double s_pinv;
    extern __shared__ double s_fac[];          // n multipliers
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = 0; k < n; ++k) {
        // pivot search over rows k..n-1 of column k
        double best = -1.0; int bi = k;
        for (int i = k + tid; i < n; i += 1024) {
            const double v = fabs(M[(size_t)i * n + k]);
            if (v > best) { best = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_down_sync(0xffffffffu, best, o);
            const int oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { s_val[warp] = best; s_idx[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = s_val[lane]; bi = s_idx[lane];
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_down_sync(0xffffffffu, best, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (lane == 0) {
                s_piv = bi;
                if (!(best > 0)) { *singular = 1; s_pinv = 0.0; }
                else s_pinv = 1.0 / M[(size_t)bi * n + k];
            }
        }
        __syncthreads();
        const int p = s_piv;
        const double pinv = s_pinv;
        // swap rows k,p and scale the pivot row (columns >= k of M, all of R)
        for (int j = k + tid; j < n + m; j += 1024) {
            double* rowk = (j < n) ? &M[(size_t)k * n + j] : &R[(size_t)k * m + (j - n)];
            double* rowp = (j < n) ? &M[(size_t)p * n + j] : &R[(size_t)p * m + (j - n)];
            const double vk = *rowk, vp = *rowp;
            *rowp = vk;
            *rowk = vp * pinv;
        }
        __syncthreads();
        for (int i = tid; i < n; i += 1024) s_fac[i] = (i == k) ? 0.0 : M[(size_t)i * n + k];
        __syncthreads();
        // eliminate column k from every other row
        const int wj = n - (k + 1) + m;          // columns k+1..n-1 of M, then all of R
        for (long long o = tid; o < (long long)n * wj; o += 1024) {
            const int i = (int)(o / wj), jj = (int)(o - (long long)i * wj);
            const double fct = s_fac[i];
            if (fct == 0.0) continue;
            if (jj < n - (k + 1)) {
                const int j = k + 1 + jj;
                M[(size_t)i * n + j] -= fct * M[(size_t)k * n + j];
            } else {
                const int j = jj - (n - (k + 1));
                R[(size_t)i * m + j] -= fct * R[(size_t)k * m + j];
            }
        }
        __syncthreads();
    }
}

struct FinalizeParams {
    const double* x; int xdim, N, d;
    const double* dx;
    const double* Pnew;      // column-major d x d (unsymmetrised)
    const double* P;         // prior (pass-through source)
    const double* gate;      // counters: gate[0] = accepted features
    double* x_out; double* P_out;
};

__device__ __forceinline__ void d_apply_dq(const double* dth, const double* q, double* out)
{
    double dq[4] = {.5 * dth[0], .5 * dth[1], .5 * dth[2], 0};
    const double vn = sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2]);
    if (vn < 1) dq[3] = sqrt(1 - vn * vn);
    else {
        const double sc = 1 / sqrt(1 + vn * vn);
        dq[0] *= sc; dq[1] *= sc; dq[2] *= sc; dq[3] = sc;
    }
    d_quat_mul(dq, q, out);
}

// State correction (Updater.cc:546-613) + covariance symmetrisation (Updater.cc:619).
__global__ void __launch_bounds__(256) k_finalize(FinalizeParams P)
{
    const int d = P.d;
    if (!(P.gate[0] > 2.0)) {
        // Updater.cc:621-627: too few measurements, posterior = prior
        for (long long o = (long long)blockIdx.x * 256 + threadIdx.x; o < (long long)d * d; o += (long long)gridDim.x * 256) P.P_out[o] = P.P[o];
        if (blockIdx.x == 0) for (int i = threadIdx.x; i < P.xdim; i += 256) P.x_out[i] = P.x[i];
        return;
    }
    for (long long o = (long long)blockIdx.x * 256 + threadIdx.x; o < (long long)d * d; o += (long long)gridDim.x * 256) {
        const int i = (int)(o % d), j = (int)(o / d);
        P.P_out[o] = .5 * (P.Pnew[(size_t)j * d + i] + P.Pnew[(size_t)i * d + j]);
    }
    if (blockIdx.x == 0) {
        const double* x = P.x; const double* dx = P.dx; double* xo = P.x_out;
        for (int b = threadIdx.x; b < 2 + P.N; b += 256) {
            int xq, eq;
            if (b == 0) { xq = 0; eq = 0; }
            else if (b == 1) { xq = 10; eq = 9; }
            else { xq = 26 + 7 * (b - 2); eq = 24 + 6 * (b - 2); }
            d_apply_dq(dx + eq, x + xq, xo + xq);
            if (b >= 2) for (int k = 0; k < 3; ++k) xo[xq + 4 + k] = dx[eq + 3 + k] + x[xq + 4 + k];
        }
        if (threadIdx.x == 0) {
            double g[3];
            for (int k = 0; k < 3; ++k) xo[4 + k] = dx[3 + k] + x[4 + k];
            for (int k = 0; k < 3; ++k) g[k] = dx[6 + k] + x[7 + k];
            const double nn = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
            for (int k = 0; k < 3; ++k) xo[7 + k] = g[k] / nn;
            for (int k = 0; k < 12; ++k) xo[14 + k] = dx[12 + k] + x[14 + k];
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Small-window EKF stage (n = 6N <= 78), Updater.cc:540-619, in three launches:
//   k_wgemm          W = G * P[c,:]  ->  R = [z | W],  M = W[:,24:] + s^2 I            (multi-CTA, 32x32 tiles)
//   k_gj_small<T>    Gauss-Jordan on [M | R], register resident, ONE CTA               (the serial part)
//   k_pout_finalize  dx = P[:,c] y, P+ = sym(P - P[:,c] Y), state correction          (multi-CTA, 32x32 tiles)
// P is symmetric by construction (PreIntegrator.cc:192, System.cc:300,361): P(a,b) is read as P[a d + b] with the
// fastest-varying index on consecutive threads.
// ------------------------------------------------------------------------------------------------
struct SolveSmallParams {
    const double* red;        // [G | z | counters]
    const double* x; const double* P; int xdim, N, d;
    double sig2;
    double* T;                // [M | z | W] TRANSPOSED: column c of the n x (n + d + 1) system at T[c * n .. c * n + n)
    double* Yt;               // solution, transposed: Yt[c * n + k] = Y(k, c), c = 0 (dx part) .. d
    double* dx;
    double* x_out; double* P_out; int* singular;
};

__global__ void __launch_bounds__(256) k_wgemm(SolveSmallParams Q)
{
    __shared__ double sA[16][33], sB[16][33];
    const int N = Q.N, n = 6 * N, d = Q.d;
    const double* gate = Q.red + (size_t)n * n + n;
    if (!(gate[0] > 2.0)) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;         // tx walks the rows of G (fastest in the transposed output)
    const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;           // rows of G, columns of P[c,:]
    const double* G = Q.red;
    const double* Pc = Q.P + (size_t)24 * d;                          // row k of P[c,:] = P[(24+k) d + col]
    double acc[2][2] = {{0, 0}, {0, 0}};
    for (int k0 = 0; k0 < n; k0 += 16) {
        for (int o = threadIdx.x; o < 16 * 32; o += 256) {
            const int kk = o >> 5, c = o & 31;
            const int k = k0 + kk;
            sA[kk][c] = (k < n && i0 + c < n) ? G[(size_t)k * n + i0 + c] : 0.0;       // G(i,k) == G(k,i) bit for bit (mirror tiles of k_gram)
            sB[kk][c] = (k < n && j0 + c < d) ? Pc[(size_t)k * d + j0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            const double a0 = sA[kk][2 * tx], a1 = sA[kk][2 * tx + 1];
            const double b0 = sB[kk][2 * ty], b1 = sB[kk][2 * ty + 1];
            acc[0][0] += a0 * b0; acc[0][1] += a0 * b1; acc[1][0] += a1 * b0; acc[1][1] += a1 * b1;
        }
        __syncthreads();
    }
    double* Tz = Q.T + (size_t)n * n;                                 // column n: z, columns n+1..n+d: W
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const int i = i0 + 2 * tx + a, col = j0 + 2 * ty + b;
            if (i < n && col < d) {
                Tz[(size_t)(1 + col) * n + i] = acc[a][b];
                if (col >= 24) Q.T[(size_t)(col - 24) * n + i] = acc[a][b] + ((i == col - 24) ? Q.sig2 : 0.0);
            }
        }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        const int i = i0 + threadIdx.x;
        if (i < n) Tz[i] = Q.red[(size_t)n * n + i];
    }
}

// Blocked Gauss-Jordan on the augmented system [M | z | W] (n x (n + d + 1)) with implicit row pivoting, ONE CTA, warp
// specialised.  The 6N serial pivot steps are grouped into panels of 8 (= one column block):
//   warps 0..2 (panel group, one row per lane) hold the 8 panel columns in registers and eliminate inside the panel, one
//                 named barrier per step: a REDUX max over packed keys (high word of |v| | 1023 - row) gives each warp's
//                 best unused row, whose lane publishes its panel entries, coefficients and pivot reciprocal
//                 speculatively; after the barrier every row takes the winner's record and updates its remaining panel
//                 entries and the coefficients C(row, t) of  new_row = keep * row + sum_t C(row, t) * old_pivot_row_t.
//                 The step loop is NOT unrolled (the current column is always register 0, entries shift down by one per
//                 step): these kernels are instruction-fetch bound when a single warp walks through unrolled code.
//                 The group then applies the panel to the NEXT panel's columns itself (their pre-update values were staged
//                 in shared memory by the owning workers one panel earlier) and continues with the next panel at once:
//                 the serial chain never waits for the block update.
//   warps 3.. (workers)   thread (g, cb) owns the 4 x 8 register tile rows 4g..4g+3, columns 8cb..8cb+7 and applies the 8
//                 steps of a panel at once: 8 DFMAs per element, pivot rows and C read from shared memory as 128-bit words;
//                 they run up to one panel behind the panel group (C and the pivot list are double buffered).
// After the last panel row p_k holds Y(k, :) in the right-hand-side columns.
constexpr int kGJPanel = 8, kGJMaxRows = 96, kGJMaxCols = 192, kGJPanelThreads = 96;

__device__ __forceinline__ void gj_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void gj_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// 1 / v to double precision from the approximate float32 reciprocal and two Newton steps (|v| is inside the float32
// range for any usable pivot).
__device__ __forceinline__ double gj_rcp(double v)
{
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"((float)v));
    double r = (double)r0;
    double e = fma(-v, r, 1.0);
    r = fma(r, e, r);
    e = fma(-v, r, 1.0);
    return fma(r, e, r);
}

struct GJShared {
    double old[2][kGJPanel][kGJMaxRows];       // a column block before the previous panel is applied, [block parity][col][row]
    double C[2][kGJPanel][kGJMaxRows];         // coefficients of the block update, [panel parity][t][row]
    double rows[kGJPanel][kGJMaxCols];         // old pivot rows, [t][column]
    double cand[2][3][18];                     // per step parity, per panel warp: a[8], C[8], 1/pivot
    unsigned key[2][4];
    short piv[2][kGJPanel];
    short prow[kGJMaxRows], var[kGJMaxRows];
    int sing;
};

template <int THREADS>
__global__ void __launch_bounds__(THREADS) k_gj_block(SolveSmallParams Q)
{
    __shared__ __align__(16) GJShared S;
    constexpr int NW = THREADS - kGJPanelThreads;                     // worker threads
    const int tid = threadIdx.x, lane = tid & 31;
    const int N = Q.N, n = 6 * N, m = Q.d + 1, ncols = n + m;
    const double* gate = Q.red + (size_t)n * n + n;
    if (!(gate[0] > 2.0)) return;
    const int ng = (n + 3) >> 2, ncb = (ncols + 7) >> 3;
    const int npan = (n + kGJPanel - 1) / kGJPanel;
    if (tid == 0) S.sing = 0;
    for (int o = tid; o < kGJPanel * kGJMaxCols; o += THREADS) (&S.rows[0][0])[o] = 0.0;

    if (tid < kGJPanelThreads) {
        // ================================================================== panel group: row = tid
        const int r = tid, warp = tid >> 5;
        const bool valid = r < n;
        bool free_r = valid;
        __syncthreads();                                                         // blocks 0 and 1 staged by their owners
        double a[kGJPanel], C[kGJPanel];                                         // a[0]: current column; C[0]: newest step
#pragma unroll
        for (int t = 0; t < kGJPanel; ++t) a[t] = valid ? S.old[0][t][r] : 0.0;
        bool sing = false;
#pragma unroll 1
        for (int pi = 0; pi < npan; ++pi) {
            const int k0 = pi * kGJPanel, buf = pi & 1;
            const int pw = (n - k0 < kGJPanel) ? n - k0 : kGJPanel;
#pragma unroll
            for (int t = 0; t < kGJPanel; ++t) C[t] = 0.0;
            bool mine = false;                                                   // row r became a pivot row in this panel
#pragma unroll 1
            for (int t = 0; t < pw; ++t) {
                const int par = t & 1;
                const unsigned key = free_r ? (((unsigned)__double2hiint(fabs(a[0])) & ~1023u) | (unsigned)(1023 - r)) : 0u;
                double my_rcp = gj_rcp(free_r ? a[0] : 1.0);                     // started before the maximum is known
                asm volatile("" : "+d"(my_rcp));
                const unsigned wkey = __reduce_max_sync(0xffffffffu, key);
                if (key == wkey && (wkey >> 10) != 0u) {                         // this warp's candidate (keys are unique)
                    double2* dst = reinterpret_cast<double2*>(&S.cand[par][warp][0]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) { dst[u] = make_double2(a[2 * u], a[2 * u + 1]); dst[4 + u] = make_double2(C[2 * u], C[2 * u + 1]); }
                    S.cand[par][warp][16] = my_rcp;
                }
                if (lane == 0) S.key[par][warp] = wkey;
                gj_bar_sync(4, kGJPanelThreads);
                const unsigned best = max(max(S.key[par][0], S.key[par][1]), S.key[par][2]);
                if ((best >> 10) == 0u) { sing = true; break; }                  // uniform over the panel group
                const int p = 1023 - (int)(best & 1023u);
                const double2* src = reinterpret_cast<const double2*>(&S.cand[par][p >> 5][0]);
                double pa[kGJPanel], pc[kGJPanel];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const double2 va = src[u], vc = src[4 + u];
                    pa[2 * u] = va.x; pa[2 * u + 1] = va.y; pc[2 * u] = vc.x; pc[2 * u + 1] = vc.y;
                }
                const double rcp = S.cand[par][p >> 5][16];
                const bool is_p = r == p;
                const double f = is_p ? 0.0 : a[0] * rcp;                        // eliminate the current column from row r
#pragma unroll
                for (int u = 0; u < kGJPanel - 1; ++u) a[u] = fma(-f, pa[u + 1], a[u + 1]);
                a[kGJPanel - 1] = 0.0;
#pragma unroll
                for (int u = kGJPanel - 1; u > 0; --u) C[u] = fma(-f, pc[u - 1], C[u - 1]);
                C[0] = -f;
                if (is_p) {                                                      // the pivot row itself: scaled old values
#pragma unroll
                    for (int u = 0; u < kGJPanel - 1; ++u) a[u] = pa[u + 1] * rcp;
#pragma unroll
                    for (int u = kGJPanel - 1; u > 0; --u) C[u] = pc[u - 1] * rcp;
                    C[0] = rcp;
                    S.piv[buf][t] = (short)p; S.prow[k0 + t] = (short)p;
                    free_r = false; mine = true;
                }
            }
#pragma unroll
            for (int i = 0; i < kGJPanel; ++i) {                                 // C[i] belongs to step pw - 1 - i
                const int tgt = (i < pw) ? pw - 1 - i : i;
                S.C[buf][tgt][r] = (i < pw) ? C[i] : 0.0;
            }
            if (sing && tid == 0) S.sing = 1;
            __threadfence_block();
            gj_bar_arrive(1, THREADS);                                           // panel pi factored
            if (sing) break;
            if (pi + 1 < npan) {
                // look-ahead: apply this panel to the next panel's columns (staged, pre-update) and keep them in registers
                if (pi > 0) gj_bar_sync(3, THREADS);                             // block pi + 1 staged by its owners
                gj_bar_sync(4, kGJPanelThreads);                                 // S.piv of this panel visible to the whole group
                const int ob = (pi + 1) & 1;
#pragma unroll
                for (int j = 0; j < kGJPanel; ++j) a[j] = (valid && !mine) ? S.old[ob][j][r] : 0.0;
#pragma unroll 1
                for (int i = 0; i < pw; ++i) {                                   // C[i] multiplies the old pivot row of step pw - 1 - i
                    const int pr = S.piv[buf][pw - 1 - i];
                    double ci = 0.0;
#pragma unroll
                    for (int u = 0; u < kGJPanel; ++u) if (u == i) ci = C[u];
#pragma unroll
                    for (int j = 0; j < kGJPanel; ++j) a[j] = fma(ci, S.old[ob][j][pr], a[j]);
                }
            }
        }
        __syncthreads();
        if (r < n) S.var[S.prow[r]] = (short)r;
        __syncthreads();
        if (S.sing && tid == 0) *Q.singular = 1;
        return;
    }

    // ====================================================================== workers
    const int w = tid - kGJPanelThreads;
    const int g = w % ng, cb = w / ng;
    const bool active = cb < ncb;
    const int r0 = 4 * g, c0 = 8 * cb;
    double reg[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
#pragma unroll
        for (int r = 0; r < 4; ++r) reg[r][j] = (active && c < ncols && r0 + r < n) ? Q.T[(size_t)c * n + r0 + r] : 0.0;
    }
    if (active && cb < 2) {
#pragma unroll
        for (int t = 0; t < kGJPanel; ++t) {
            *reinterpret_cast<double2*>(&S.old[cb][t][r0]) = make_double2(reg[0][t], reg[1][t]);
            *reinterpret_cast<double2*>(&S.old[cb][t][r0 + 2]) = make_double2(reg[2][t], reg[3][t]);
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int pi = 0; pi < npan; ++pi) {
        const int k0 = pi * kGJPanel, buf = pi & 1;
        const int pw = (n - k0 < kGJPanel) ? n - k0 : kGJPanel;
        gj_bar_sync(1, THREADS);                                                 // C and the pivot list of panel pi are ready
        if (S.sing) break;                                                       // uniform
        int myt[4] = {-1, -1, -1, -1};
#pragma unroll
        for (int t = 0; t < kGJPanel; ++t) {
            const int pr = (t < pw) ? S.piv[buf][t] - r0 : -1;
#pragma unroll
            for (int r = 0; r < 4; ++r) if (pr == r) myt[r] = t;
        }
        // left of the panel everything is final (a short last panel shares its block with right-hand-side columns)
        const bool live = active && (cb > (k0 >> 3) || (cb == (k0 >> 3) && pw < kGJPanel));
        if (live) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (myt[r] >= 0) {
                    double2* dst = reinterpret_cast<double2*>(&S.rows[myt[r]][c0]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_double2(reg[r][2 * j], reg[r][2 * j + 1]);
                }
        }
        gj_bar_sync(2, NW);                                                      // pivot rows published (workers only)
        if (live) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (myt[r] >= 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) reg[r][j] = 0.0;
                }
#pragma unroll 1
            for (int t = 0; t < kGJPanel; ++t) {
                double row[8], cc[4];
                const double2* rs = reinterpret_cast<const double2*>(&S.rows[t][c0]);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const double2 v2 = rs[j]; row[2 * j] = v2.x; row[2 * j + 1] = v2.y; }
                const double2 c01 = *reinterpret_cast<const double2*>(&S.C[buf][t][r0]);
                const double2 c23 = *reinterpret_cast<const double2*>(&S.C[buf][t][r0 + 2]);
                cc[0] = c01.x; cc[1] = c01.y; cc[2] = c23.x; cc[3] = c23.y;
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int j = 0; j < 8; ++j) reg[r][j] += cc[r] * row[j];
            }
        }
        if (pi + 2 < npan) {                                                     // block pi + 2 (panels 0..pi applied) for the look-ahead
            if (active && cb == pi + 2) {
#pragma unroll
                for (int t = 0; t < kGJPanel; ++t) {
                    *reinterpret_cast<double2*>(&S.old[buf][t][r0]) = make_double2(reg[0][t], reg[1][t]);
                    *reinterpret_cast<double2*>(&S.old[buf][t][r0 + 2]) = make_double2(reg[2][t], reg[3][t]);
                }
            }
            __threadfence_block();
            gj_bar_arrive(3, THREADS);
        }
    }
    __syncthreads();
    __syncthreads();                                                             // S.var written by the panel group
    if (S.sing) return;
    if (active) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (r0 + r >= n) continue;
            const int row = S.var[r0 + r];                                       // this register row is row `row` of the solution
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = c0 + j;
                if (c >= n && c < ncols) Q.Yt[(size_t)(c - n) * n + row] = reg[r][j];
            }
        }
    }
}

// dx = P[:,c] y ; P_out = sym(P - P[:,c] Y) on 32x32 tiles (each tile also forms its transposed product so that both
// mirror tiles write bit-identical values); block (0,0) then applies the state correction (Updater.cc:546-613).
__global__ void __launch_bounds__(256) k_pout_finalize(SolveSmallParams Q)
{
    __shared__ double sPi[16][33], sPj[16][33], sYi[16][33], sYj[16][33];
    __shared__ double s_dx[24 + 6 * 13 + 8];
    const int N = Q.N, n = 6 * N, d = Q.d;
    const double* gate = Q.red + (size_t)n * n + n;
    const int tid = threadIdx.x;
    const int nb = (d + 31) / 32;
    if (!(gate[0] > 2.0)) {                       // Updater.cc:621-627: posterior = prior
        for (int o = blockIdx.x * 256 + tid; o < d * d; o += gridDim.x * 256) Q.P_out[o] = Q.P[o];
        if (blockIdx.x == 0) for (int o = tid; o < Q.xdim; o += 256) Q.x_out[o] = Q.x[o];
        return;
    }
    const int ti = blockIdx.x / nb, tj = blockIdx.x % nb;
    const int tx = tid & 15, ty = tid >> 4;
    const int i0 = ti * 32, j0 = tj * 32;
    const double* Pc = Q.P + (size_t)24 * d;                          // Pc[k][i] = P(i, 24+k)
    const double* Y = Q.Yt + n;                                       // Y(k, j) at Yt[(1 + j) * n + k]
    double a[2][2] = {{0, 0}, {0, 0}}, b[2][2] = {{0, 0}, {0, 0}};
    for (int k0 = 0; k0 < n; k0 += 16) {
        for (int o = tid; o < 16 * 32; o += 256) {
            const int kk = o >> 5, c = o & 31;
            const int k = k0 + kk;
            const bool kv = k < n;
            sPi[kk][c] = (kv && i0 + c < d) ? Pc[(size_t)k * d + i0 + c] : 0.0;
            sPj[kk][c] = (kv && j0 + c < d) ? Pc[(size_t)k * d + j0 + c] : 0.0;
        }
        for (int o = tid; o < 16 * 32; o += 256) {                        // Y is stored transposed: k fastest
            const int kk = o & 15, c = o >> 4;
            const int k = k0 + kk;
            const bool kv = k < n;
            sYi[kk][c] = (kv && i0 + c < d) ? Y[(size_t)(i0 + c) * n + k] : 0.0;
            sYj[kk][c] = (kv && j0 + c < d) ? Y[(size_t)(j0 + c) * n + k] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    a[u][v] += sPi[kk][2 * ty + u] * sYj[kk][2 * tx + v];          // (P[:,c] Y)(i, j)
                    b[u][v] += sPj[kk][2 * tx + v] * sYi[kk][2 * ty + u];          // (P[:,c] Y)(j, i)
                }
        }
        __syncthreads();
    }
    for (int u = 0; u < 2; ++u)
        for (int v = 0; v < 2; ++v) {
            const int i = i0 + 2 * ty + u, j = j0 + 2 * tx + v;
            if (i < d && j < d) {
                const double pij = Q.P[(size_t)j * d + i] - a[u][v], pji = Q.P[(size_t)i * d + j] - b[u][v];
                Q.P_out[(size_t)j * d + i] = .5 * (pij + pji);
            }
        }
    if (blockIdx.x != 0) return;
    // dx and the state correction (block 0)
    for (int i = tid; i < d; i += 256) {
        double acc = 0;
        for (int k = 0; k < n; ++k) acc += Pc[(size_t)k * d + i] * Q.Yt[k];
        s_dx[i] = acc;
    }
    __syncthreads();
    const double* x = Q.x; double* xo = Q.x_out; const double* dx = s_dx;
    for (int bq = tid; bq < 2 + N; bq += 256) {
        int xq, eq;
        if (bq == 0) { xq = 0; eq = 0; }
        else if (bq == 1) { xq = 10; eq = 9; }
        else { xq = 26 + 7 * (bq - 2); eq = 24 + 6 * (bq - 2); }
        d_apply_dq(dx + eq, x + xq, xo + xq);
        if (bq >= 2) for (int k = 0; k < 3; ++k) xo[xq + 4 + k] = dx[eq + 3 + k] + x[xq + 4 + k];
    }
    if (tid == 64) {
        double g[3];
        for (int k = 0; k < 3; ++k) xo[4 + k] = dx[3 + k] + x[4 + k];
        for (int k = 0; k < 3; ++k) g[k] = dx[6 + k] + x[7 + k];
        const double nn = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
        for (int k = 0; k < 3; ++k) xo[7 + k] = g[k] / nn;
        for (int k = 0; k < 12; ++k) xo[14 + k] = dx[12 + k] + x[14 + k];
    }
}

}  // namespace rvio

// ================================================================================================
// host side
// ================================================================================================
using namespace rvio;

struct rvio_updater {
    rvio_updater_cfg cfg;
    int device;
    cudaStream_t stream;
    int Nmax, Fmax, Lmax, nmax, dmax, xmax;
    UpdaterConsts consts;
    FeatLayout lay;
    // device
    double *d_x, *d_P, *d_xout, *d_Pout, *d_Pnew, *d_dx;
    uint8_t* d_types; int32_t* d_off; float2* d_xy;
    uint8_t* d_fstatus; double *d_fpfinv, *d_fgamma, *d_ffro2, *d_Tg; int32_t *d_fdof, *d_fc0, *d_fwc, *d_fpend;
    int gate_tensor_min_n;      // windows with n >= this run the gate's H_stack * Pcc product on the tensor cores (FP64 DMMA)
    double *d_Hblk, *d_rblk, *d_Gpart, *d_zpart, *d_red, *d_M, *d_R, *d_chi2, *d_T, *d_Yt;
    int* d_sing; int* d_tickets;
    int* d_rule; int32_t* d_rr; double *d_L, *d_gwin, *d_Rc, *d_yc, *d_S, *d_LS, *d_W;      // reference compression rule + R-form EKF step (compress.cu)
    int rank_rule;
    bool split_small_solve;     // RVIO_B200_SPLIT_SOLVE=1: rank rule / sweep / EKF step as three launches instead of k_update_small, for A/B timing
    bool legacy_small_solve;    // RVIO_B200_LEGACY_SOLVE=1: the round-1 G-form solve (k_wgemm + k_gj_block + k_pout_finalize), for A/B timing
    int groups_cap;
    // pinned
    double *h_x, *h_P, *h_red; uint8_t* h_types; int32_t* h_off; float* h_xy; int* h_sing; int* h_rule;
    // state of an open begin/finish pair
    int cur_N, cur_xdim, cur_d, cur_nfeat, cur_rank, cur_world; bool open; const int* cur_nfeat_dev;
    bool pred_is_kernel = false;     // the next updater_enqueue_normal_terms directly follows a kernel on its stream (PDL allowed)
    const double* cur_x_dev; const double* cur_P_dev;
    std::vector<void*> allocs, hallocs;
};

namespace {

template <typename T>
int ualloc(rvio_updater* u, T** p, size_t count)
{
    void* q = nullptr;
    RVIO_CUDA_TRY(cudaMalloc(&q, count * sizeof(T) + 16));
    RVIO_CUDA_TRY(cudaMemsetAsync(q, 0, count * sizeof(T) + 16, u->stream));
    u->allocs.push_back(q);
    *p = (T*)q;
    return RVIO_OK;
}
template <typename T>
int uhalloc(rvio_updater* u, T** p, size_t count)
{
    void* q = nullptr;
    RVIO_CUDA_TRY(cudaMallocHost(&q, count * sizeof(T) + 16));
    u->hallocs.push_back(q);
    *p = (T*)q;
    return RVIO_OK;
}

FeatLayout make_layout(int Lc)
{
    FeatLayout y;
    y.Lc = Lc; y.Pc = Lc - 1; y.Mc = 2 * Lc; y.Wc = 6 * (Lc - 1); y.Dc = 2 * Lc - 2;
    int o = 0;
    auto take = [&](int cnt) { int r = o; o += (cnt + 1) & ~1; return r; };
    y.o_relI = take(7 * y.Pc); y.o_RI = take(9 * y.Pc); y.o_RC = take(9 * y.Pc); y.o_tC = take(3 * y.Pc);
    y.o_HRR = take(6 * y.Lc); y.o_SUB = take(18 * y.Pc); y.o_Hf = take(3 * y.Mc); y.o_r = take(y.Mc);
    y.o_v = take(y.Mc); y.o_Hx = take(y.Mc * y.Wc); y.o_S = take(y.Dc * y.Dc); y.o_T = take(y.Dc * 32);
    y.o_meas = take(Lc);      // float2 == one double each
    y.total_bytes = o * 8;
    return y;
}

void launch_gemm(cudaStream_t s, const GemmParams& g)
{
    dim3 grid(div_up(g.N, 32), div_up(g.M, 32));
    RVIO_LAUNCH(k_dgemm, grid, 256, 0, s, g);
}

}  // namespace

extern "C" int rvio_updater_create(const rvio_updater_cfg* cfg, int device, rvio_updater** out)
{
    RVIO_ARG_CHECK(cfg && out);
    RVIO_ARG_CHECK(cfg->max_clones >= 1 && cfg->max_features >= 1 && cfg->max_track_len >= 2);
    if (cfg->max_clones > 31) { rvio::set_error("rvio_updater_create", "more than 31 clones (n = 186 clone columns) is not supported"); return RVIO_ERR_CAPACITY; }
    int rc = require_b200(device);
    if (rc != RVIO_OK) return rc;
    rvio_updater* u = new (std::nothrow) rvio_updater();
    if (!u) return RVIO_ERR_CUDA;
    u->cfg = *cfg; u->device = device; u->open = false;
    u->Nmax = cfg->max_clones; u->Fmax = cfg->max_features; u->Lmax = cfg->max_track_len;
    u->nmax = 6 * u->Nmax; u->dmax = 24 + u->nmax; u->xmax = 26 + 7 * u->Nmax;
    RVIO_CUDA_TRY(cudaSetDevice(device));
    RVIO_CUDA_TRY(cudaStreamCreateWithFlags(&u->stream, cudaStreamNonBlocking));
    UpdaterConsts& c = u->consts;
    const float sx = cfg->sigma_px, sy = cfg->sigma_py;
    c.sigma = (double)(sx > sy ? sx : sy);               // Updater.cc:42-44
    c.sig2 = c.sigma * c.sigma;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { c.Ric[3 * i + j] = cfg->T_BC0[4 * i + j]; c.Rci[3 * j + i] = cfg->T_BC0[4 * i + j]; }
        c.tic[i] = cfg->T_BC0[4 * i + 3];
    }
    for (int i = 0; i < 3; ++i)
        c.tci[i] = -(c.Rci[3 * i] * c.tic[0] + c.Rci[3 * i + 1] * c.tic[1] + c.Rci[3 * i + 2] * c.tic[2]);   // Updater.cc:53
    u->lay = make_layout(u->Lmax);
    if (u->lay.total_bytes > 200 * 1024) { set_error("rvio_updater_create", "max_track_len too large for shared memory"); return RVIO_ERR_CAPACITY; }
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_feature, cudaFuncAttributeMaxDynamicSharedMemorySize, u->lay.total_bytes));
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_gauss_jordan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (u->nmax + 2))));
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_dmma_hp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (2 * kGM * kGLdA + 2 * kGK * kGLdB))));
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_gram_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGramDmmaSmem));
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_gate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * (44 * 64 + 2 * 64 * 34 + 16))));
    const size_t F = u->Fmax, n = u->nmax, d = u->dmax, Mc = u->lay.Mc;
    u->groups_cap = cfg->max_clones * 6 >= 96 ? 48 : 16;      // partial normal terms per tile (the tensor-core variant spreads wider)
#define A(p, cnt) if ((rc = ualloc(u, &(p), (cnt))) != RVIO_OK) return rc
    A(u->d_x, u->xmax); A(u->d_xout, u->xmax); A(u->d_P, d * d); A(u->d_Pout, d * d); A(u->d_Pnew, d * d); A(u->d_dx, d);
    A(u->d_types, F + 1); A(u->d_off, F + 2); A(u->d_xy, F * u->Lmax + 1);
    A(u->d_fstatus, F + 1); A(u->d_fpfinv, 3 * F + 3); A(u->d_fgamma, F + 1); A(u->d_fdof, F + 1); A(u->d_fc0, F + 1); A(u->d_fwc, F + 1); A(u->d_ffro2, F + 1); A(u->d_fpend, F + 1);
    u->gate_tensor_min_n = 96;
    { const char* e = getenv("RVIO_B200_GATE_TENSOR_MIN_N"); if (e) u->gate_tensor_min_n = atoi(e); }
    if ((int)n >= u->gate_tensor_min_n) { A(u->d_Tg, F * Mc * n); } else u->d_Tg = nullptr;
    A(u->d_Hblk, F * Mc * n); A(u->d_rblk, F * Mc);
    A(u->d_Gpart, (size_t)u->groups_cap * n * n); A(u->d_zpart, (size_t)u->groups_cap * n);
    A(u->d_red, n * n + n + 8 + n + 1); A(u->d_M, n * n); A(u->d_R, n * (d + 1)); A(u->d_T, n * (n + d + 1)); A(u->d_Yt, n * (d + 1)); A(u->d_chi2, 500); A(u->d_sing, 1);
    A(u->d_tickets, (size_t)div_up((int)n, 32) * div_up((int)n, 32) + 64);
    A(u->d_rule, 4); A(u->d_rr, 8); A(u->d_L, n * (n + 1) + 8); A(u->d_gwin, givens_window_doubles((int)n));
    A(u->d_Rc, n * n); A(u->d_yc, n); A(u->d_S, n * n); A(u->d_LS, n * n + 64 * n + 64);      // d_LS: tile-packed factor (8 x 8 tiles of the padded lower triangle)
    A(u->d_W, n * (d + 1));
#undef A
#define HA(p, cnt) if ((rc = uhalloc(u, &(p), (cnt))) != RVIO_OK) return rc
    HA(u->h_x, u->xmax); HA(u->h_P, d * d); HA(u->h_red, 16); HA(u->h_types, F + 1); HA(u->h_off, F + 2); HA(u->h_xy, 2 * (F * u->Lmax + 1)); HA(u->h_sing, 4); HA(u->h_rule, 4);
#undef HA
    RVIO_CUDA_TRY(cudaMemcpyAsync(u->d_chi2, RVIO_CHI2_95_HOST, sizeof(double) * 500, cudaMemcpyHostToDevice, u->stream));
    u->rank_rule = 0;                                    // the reference's rule (d_rule is zero-initialised)
    { const char* e = getenv("RVIO_B200_LEGACY_SOLVE"); u->legacy_small_solve = e && e[0] == '1'; }
    { const char* e = getenv("RVIO_B200_SPLIT_SOLVE"); u->split_small_solve = e && e[0] == '1'; }
    if ((rc = compress_configure(u->nmax)) != RVIO_OK) return rc;
    RVIO_CUDA_TRY(cudaStreamSynchronize(u->stream));
    *out = u;
    return RVIO_OK;
}

extern "C" void rvio_updater_destroy(rvio_updater* u)
{
    if (!u) return;
    cudaSetDevice(u->device);
    cudaStreamSynchronize(u->stream);
    for (void* p : u->allocs) cudaFree(p);
    for (void* p : u->hallocs) cudaFreeHost(p);
    cudaStreamDestroy(u->stream);
    delete u;
}

// ------------------------------------------------------------------------------------------------
// Device-resident core.  Stage A: per-feature kernel + normal terms into the reduce buffer.
// Stage B: EKF solve + finalisation from the (possibly all-reduced) buffer.  Neither stage synchronises.
// ------------------------------------------------------------------------------------------------
namespace rvio {

void updater_hint_kernel_predecessor(rvio_updater* u, bool yes) { u->pred_is_kernel = yes; }

int updater_enqueue_normal_terms(rvio_updater* u, cudaStream_t s, const double* x_dev, int xdim, const double* P_dev, int d,
                                 const uint8_t* types_dev, const int32_t* off_dev, const float2* xy_dev,
                                 int n_feat_cap, const int* n_feat_dev, int rank, int world)
{
    const int N = (xdim - 26) / 7, n = 6 * N;
    u->cur_N = N; u->cur_xdim = xdim; u->cur_d = d; u->cur_nfeat = n_feat_cap; u->cur_rank = rank; u->cur_world = world;
    u->cur_x_dev = x_dev; u->cur_P_dev = P_dev; u->cur_nfeat_dev = n_feat_dev;
    u->open = true;
    if (n_feat_cap > 0 && N > 0) {
        FeatureParams fp;
        fp.x = x_dev; fp.xdim = xdim; fp.N = N; fp.P = P_dev; fp.d = d;
        fp.types = types_dev; fp.offsets = off_dev; fp.xy = xy_dev; fp.n_feat = n_feat_cap; fp.n_feat_dev = n_feat_dev;
        fp.rank = rank; fp.world = world; fp.chi2 = u->d_chi2;
        fp.f_status = u->d_fstatus; fp.f_pfinv = u->d_fpfinv; fp.f_gamma = u->d_fgamma; fp.f_dof = u->d_fdof;
        fp.f_c0 = u->d_fc0; fp.f_wc = u->d_fwc; fp.f_fro2 = u->d_ffro2; fp.Hblk = u->d_Hblk; fp.rblk = u->d_rblk; fp.blk_rows = u->lay.Mc;
        fp.c = u->consts; fp.lay = u->lay;
        if (world > 1) {
            // features owned by other ranks must not leave stale status/dof behind
            RVIO_ENQ(cudaMemsetAsync(u->d_fdof, 0, sizeof(int32_t) * n_feat_cap, s));
            RVIO_ENQ(cudaMemsetAsync(u->d_fstatus, 0xff, n_feat_cap, s));
        }
        const bool tensor_gate = u->d_Tg != nullptr && n >= u->gate_tensor_min_n;
        fp.gate_mode = tensor_gate ? 1 : 0; fp.f_pend = u->d_fpend;
        // (behind k_ransac_bookkeep in the fused frame: vio.cu says so through updater_hint_kernel_predecessor)
        RVIO_LAUNCH_PDL(world == 1 && u->pred_is_kernel, k_feature, n_feat_cap, kFeatThreads, u->lay.total_bytes, s, fp);
        u->pred_is_kernel = false;
        if (tensor_gate) {
            // Updater.cc:416: the per-feature products H~ Pcc as ONE tall GEMM on the tensor cores, then the per-feature gate
            DmmaParams dq;
            dq.A = u->d_Hblk; dq.lda = n; dq.B = P_dev + (size_t)24 * d + 24; dq.ldb = d; dq.C = u->d_Tg; dq.ldc = n;
            dq.M = n_feat_cap * u->lay.Mc; dq.N = n; dq.K = n; dq.n_rows_dev = n_feat_dev; dq.rows_per = u->lay.Mc;
            RVIO_LAUNCH(k_dmma_hp, dim3(div_up(dq.M, kGM), div_up(n, kGN)), 256, sizeof(double) * (2 * kGM * kGLdA + 2 * kGK * kGLdB), s, dq);
            GateParams gq;
            gq.Hblk = u->d_Hblk; gq.rblk = u->d_rblk; gq.T = u->d_Tg; gq.f_pend = u->d_fpend; gq.f_c0 = u->d_fc0; gq.f_wc = u->d_fwc;
            gq.n_feat = n_feat_cap; gq.n_feat_dev = n_feat_dev; gq.n = n; gq.blk_rows = u->lay.Mc; gq.rank = rank; gq.world = world;
            gq.sig2 = u->consts.sig2; gq.chi2 = u->d_chi2; gq.f_status = u->d_fstatus; gq.f_gamma = u->d_fgamma; gq.f_dof = u->d_fdof;
            RVIO_LAUNCH(k_gate, n_feat_cap, kGateThreads, sizeof(double) * (44 * 64 + 2 * 64 * 34 + 16), s, gq);
        }
        GramParams gp;
        gp.Hblk = u->d_Hblk; gp.rblk = u->d_rblk; gp.f_dof = u->d_fdof; gp.f_c0 = u->d_fc0; gp.f_wc = u->d_fwc; gp.f_fro2 = u->d_ffro2;
        gp.n_feat = n_feat_cap; gp.n_feat_dev = n_feat_dev; gp.n = n; gp.blk_rows = u->lay.Mc;
        gp.groups = n_feat_cap < 16 ? n_feat_cap : 16;
        gp.nt = div_up(n, 32); gp.Gpart = u->d_Gpart; gp.zpart = u->d_zpart;
        if (tensor_gate) {
            gp.groups = n_feat_cap < u->groups_cap ? n_feat_cap : u->groups_cap;
            gp.nt = div_up(n, kGramT);
            RVIO_LAUNCH(k_gram_dmma, dim3(gp.nt * (gp.nt + 1) / 2, gp.groups), 128, kGramDmmaSmem, s, gp, u->d_fstatus, rank, world, u->d_red, u->d_tickets);
        } else {
            RVIO_LAUNCH_PDL(true, k_gram, dim3(gp.nt * gp.nt, gp.groups), 256, 0, s, gp, u->d_fstatus, rank, world, u->d_red, u->d_tickets);
        }
    } else {
        RVIO_ENQ(cudaMemsetAsync(u->d_red, 0, sizeof(double) * ((size_t)n * n + n + 8 + n + 1), s));
    }
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

int updater_enqueue_solve(rvio_updater* u, cudaStream_t s, double* x_out_dev, double* P_out_dev)
{
    const int N = u->cur_N, n = 6 * N, d = u->cur_d, xdim = u->cur_xdim;
    const double* x_dev = u->cur_x_dev; const double* P_dev = u->cur_P_dev;
    u->open = false;
    if (n == 0) {
        RVIO_ENQ(cudaMemcpyAsync(x_out_dev, x_dev, sizeof(double) * xdim, cudaMemcpyDeviceToDevice, s));
        RVIO_ENQ(cudaMemcpyAsync(P_out_dev, P_dev, sizeof(double) * (size_t)d * d, cudaMemcpyDeviceToDevice, s));
        return RVIO_OK;
    }
    const bool small = N <= kSolveSmallMaxClones;
    RankRuleParams rq;
    GivensRefParams gq;
    {
        // which rows the reference's compression keeps (Updater.cc:474-536): may rewrite [G | z] in place; it also hands over
        // the kept rows R (G = R^T R) and y (R^T y = z) to the EKF step
        rq.red = u->d_red; rq.n = n; rq.world = u->cur_world; rq.rule_dev = u->d_rule; rq.L = u->d_L; rq.rr = u->d_rr;
        rq.emit_R = 1; rq.Rc = u->d_Rc; rq.yc = u->d_yc;
        gq.Hblk = u->d_Hblk; gq.rblk = u->d_rblk; gq.f_dof = u->d_fdof; gq.n_feat = u->cur_nfeat; gq.n_feat_dev = u->cur_nfeat_dev;
        gq.n = n; gq.blk_rows = u->lay.Mc; gq.red = u->d_red; gq.rr = u->d_rr; gq.win = u->d_gwin;
        gq.emit_R = rq.emit_R; gq.Rc = u->d_Rc; gq.yc = u->d_yc;
    }
    if (small && !u->legacy_small_solve && !u->split_small_solve) {
        // rank rule, (rare) sweep and the whole EKF step in ONE single-CTA launch (compress.cu: k_update_small)
        SolveSmallRParams q;
        q.Rc = u->d_Rc; q.yc = u->d_yc; q.x = x_dev; q.P = P_dev; q.xdim = xdim; q.N = N; q.d = d; q.sig2 = u->consts.sig2;
        q.gate = u->d_red + (size_t)n * n + n; q.x_out = x_out_dev; q.P_out = P_out_dev; q.bad = u->d_sing;
        return enqueue_update_small(s, rq, gq, q);
    }
    {
        const int r2 = enqueue_rank_rule(s, rq, gq, n);
        if (r2 != RVIO_OK) return r2;
    }
    if (small && !u->legacy_small_solve) {
        // (RVIO_B200_SPLIT_SOLVE=1: the three launches of the un-fused form, kept for A/B timing)
        SolveSmallRParams q;
        q.Rc = u->d_Rc; q.yc = u->d_yc; q.x = x_dev; q.P = P_dev; q.xdim = xdim; q.N = N; q.d = d; q.sig2 = u->consts.sig2;
        q.gate = u->d_red + (size_t)n * n + n; q.x_out = x_out_dev; q.P_out = P_out_dev; q.bad = u->d_sing;
        RVIO_ENQ(cudaMemsetAsync(u->d_sing, 0, sizeof(int), s));
        return enqueue_solve_small_R(s, q);
    }
    if (small) {
        SolveSmallParams sp;
        sp.red = u->d_red; sp.x = x_dev; sp.P = P_dev; sp.xdim = xdim; sp.N = N; sp.d = d; sp.sig2 = u->consts.sig2;
        sp.T = u->d_T; sp.Yt = u->d_Yt; sp.dx = u->d_dx;
        sp.x_out = x_out_dev; sp.P_out = P_out_dev; sp.singular = u->d_sing;
        RVIO_ENQ(cudaMemsetAsync(u->d_sing, 0, sizeof(int), s));
        RVIO_LAUNCH(k_wgemm, dim3(div_up(d, 32), div_up(n, 32)), 256, 0, s, sp);
        if (kGJPanelThreads + div_up(n, 4) * div_up(n + d + 1, 8) <= 448) RVIO_LAUNCH(k_gj_block<448>, 1, 448, 0, s, sp);   // 3 panel warps + one thread per 4 x 8 tile
        else RVIO_LAUNCH(k_gj_block<576>, 1, 576, 0, s, sp);
        const int nb = div_up(d, 32);
        RVIO_LAUNCH(k_pout_finalize, nb * nb, 256, 0, s, sp);
        RVIO_ENQ(cudaGetLastError());
        return RVIO_OK;
    }
    // ---- large windows (N > 14: configs[2] / [4]): the EKF step in the reference's own
    //      form on the kept rows R (n x n, zero padded; Updater.cc:540-619 with Hn = R):
    //          W = R P[c,:]        S = W[:,c] R^T + s^2 I  (SPD)        S = L L^T        Y = L^-1 [W | y]
    //          dx = Y^T y~         P+ = P - Y^T Y                       (y~ = L^-1 y = last column of Y)
    //      GEMMs on all SMs, the Cholesky in one register-resident CTA, the triangular solves one CTA per 8 columns.
    const double* gate = u->d_red + (size_t)n * n + n;      // counters[0] = accepted features (Updater.cc:460)
    const int m = d + 1;
    {
        GemmParams g;                                        // W = Rc * P[c,:]   -> d_W (n x m row-major, column d reserved for y)
        g.M = n; g.N = d; g.K = n;
        g.A = u->d_Rc; g.ars = n; g.acs = 1;
        g.B = P_dev + 24; g.brs = 1; g.bcs = d;              // P[c,:](k, j) = P(24 + k, j) = P_dev[j * d + 24 + k]
        g.C0 = nullptr; g.c0rs = g.c0cs = 0; g.C = u->d_W; g.crs = m; g.ccs = 1;
        g.alpha = 1; g.beta = 0; g.diag_add = 0; g.gate = gate;
        launch_gemm(s, g);
        RVIO_ENQ(cudaMemcpy2DAsync(u->d_W + d, sizeof(double) * m, u->d_yc, sizeof(double), sizeof(double), n, cudaMemcpyDeviceToDevice, s));
    }
    {
        GemmParams g;                                        // S = W[:, 24:] * Rc^T + s^2 I
        g.M = n; g.N = n; g.K = n;
        g.A = u->d_W + 24; g.ars = m; g.acs = 1;
        g.B = u->d_Rc; g.brs = 1; g.bcs = n;                 // Rc^T(k, j) = Rc[j * n + k]
        g.C0 = nullptr; g.c0rs = g.c0cs = 0; g.C = u->d_S; g.crs = n; g.ccs = 1;
        g.alpha = 1; g.beta = 0; g.diag_add = u->consts.sig2; g.gate = gate;
        launch_gemm(s, g);
    }
    RVIO_ENQ(cudaMemsetAsync(u->d_sing, 0, sizeof(int), s));
    {
        const int r2 = enqueue_chol_trsm(s, u->d_S, n, u->d_LS, u->d_W, m, m, u->d_sing, gate);
        if (r2 != RVIO_OK) return r2;
    }
    {
        GemmParams g;                                        // dx = Y^T y~
        g.M = d; g.N = 1; g.K = n;
        g.A = u->d_W; g.ars = 1; g.acs = m;
        g.B = u->d_W + d; g.brs = m; g.bcs = 1;
        g.C0 = nullptr; g.c0rs = g.c0cs = 0; g.C = u->d_dx; g.crs = 1; g.ccs = 1;
        g.alpha = 1; g.beta = 0; g.diag_add = 0; g.gate = gate;
        launch_gemm(s, g);
    }
    {
        GemmParams g;                                        // Pnew = P - Y^T Y   (column-major d x d)
        g.M = d; g.N = d; g.K = n;
        g.A = u->d_W; g.ars = 1; g.acs = m;
        g.B = u->d_W; g.brs = m; g.bcs = 1;
        g.C0 = P_dev; g.c0rs = 1; g.c0cs = d; g.C = u->d_Pnew; g.crs = 1; g.ccs = d;
        g.alpha = -1; g.beta = 1; g.diag_add = 0; g.gate = gate;
        launch_gemm(s, g);
    }
    {
        FinalizeParams fp;
        fp.x = x_dev; fp.xdim = xdim; fp.N = N; fp.d = d; fp.dx = u->d_dx; fp.Pnew = u->d_Pnew; fp.P = P_dev; fp.gate = gate;
        fp.x_out = x_out_dev; fp.P_out = P_out_dev;
        RVIO_LAUNCH(k_finalize, div_up(d * d, 256), 256, 0, s, fp);
    }
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

// Solve against a state other than the one the normal terms were formed on (the propagated x, P of the fused path).
int updater_enqueue_solve_on(rvio_updater* u, cudaStream_t s, const double* x_dev, const double* P_dev, double* x_out_dev, double* P_out_dev)
{
    u->cur_x_dev = x_dev; u->cur_P_dev = P_dev;
    return updater_enqueue_solve(u, s, x_out_dev, P_out_dev);
}

const double* updater_counters_dev(const rvio_updater* u)
{
    const int n = 6 * u->cur_N;
    return u->d_red + (size_t)n * n + n;
}
int updater_device(const rvio_updater* u) { return u->device; }
double* updater_reduce_dev(rvio_updater* u, int* count)
{
    const int n = 6 * u->cur_N;
    *count = n * n + n + 8 + n + 1;
    return u->d_red;
}

}  // namespace rvio

static int check_shapes(rvio_updater* u, const double* x, int xdim, const double* P, int d, int n_feat, int rank, int world)
{
    RVIO_ARG_CHECK(x && P);
    RVIO_ARG_CHECK(xdim >= 26 && (xdim - 26) % 7 == 0);
    const int N = (xdim - 26) / 7;
    RVIO_ARG_CHECK(d == 24 + 6 * N);
    RVIO_ARG_CHECK(world >= 1 && rank >= 0 && rank < world);
    if (N > u->Nmax || n_feat > u->Fmax) { set_error("rvio_updater_update", "exceeds capacity given at create"); return RVIO_ERR_CAPACITY; }
    return RVIO_OK;
}

static int upload_state(rvio_updater* u, const double* x, int xdim, const double* P, int d)
{
    cudaStream_t s = u->stream;
    memcpy(u->h_x, x, sizeof(double) * xdim);
    memcpy(u->h_P, P, sizeof(double) * (size_t)d * d);
    RVIO_ENQ(cudaMemcpyAsync(u->d_x, u->h_x, sizeof(double) * xdim, cudaMemcpyHostToDevice, s));
    RVIO_ENQ(cudaMemcpyAsync(u->d_P, u->h_P, sizeof(double) * (size_t)d * d, cudaMemcpyHostToDevice, s));
    return RVIO_OK;
}

static int upload_lists(rvio_updater* u, const uint8_t* types, const int32_t* offsets, const float* xy, int n_feat)
{
    RVIO_ARG_CHECK(n_feat >= 0 && (n_feat == 0 || (types && offsets && xy)));
    if (n_feat > u->Fmax) { set_error("rvio_updater_update", "n_feat exceeds capacity"); return RVIO_ERR_CAPACITY; }
    if (n_feat == 0) return RVIO_OK;
    const int n_meas = offsets[n_feat] - offsets[0];
    RVIO_ARG_CHECK(offsets[0] == 0 && n_meas >= 0);
    if ((size_t)n_meas > (size_t)u->Fmax * u->Lmax) { set_error("rvio_updater_update", "too many measurements"); return RVIO_ERR_CAPACITY; }
    cudaStream_t s = u->stream;
    memcpy(u->h_types, types, n_feat);
    memcpy(u->h_off, offsets, sizeof(int32_t) * (n_feat + 1));
    memcpy(u->h_xy, xy, sizeof(float) * 2 * n_meas);
    RVIO_ENQ(cudaMemcpyAsync(u->d_types, u->h_types, n_feat, cudaMemcpyHostToDevice, s));
    RVIO_ENQ(cudaMemcpyAsync(u->d_off, u->h_off, sizeof(int32_t) * (n_feat + 1), cudaMemcpyHostToDevice, s));
    RVIO_ENQ(cudaMemcpyAsync(u->d_xy, u->h_xy, sizeof(float) * 2 * n_meas, cudaMemcpyHostToDevice, s));
    return RVIO_OK;
}

extern "C" int rvio_updater_update_begin(rvio_updater* u, const double* x, int xdim, const double* P, int d,
                                         const uint8_t* types, const int32_t* offsets, const float* xy, int n_feat,
                                         int rank, int world)
{
    RVIO_ARG_CHECK(u);
    RVIO_CUDA_TRY(cudaSetDevice(u->device));
    int rc = check_shapes(u, x, xdim, P, d, n_feat, rank, world);
    if (rc != RVIO_OK) return rc;
    if ((rc = upload_lists(u, types, offsets, xy, n_feat)) != RVIO_OK) return rc;
    if ((rc = upload_state(u, x, xdim, P, d)) != RVIO_OK) return rc;
    return updater_enqueue_normal_terms(u, u->stream, u->d_x, xdim, u->d_P, d, u->d_types, u->d_off, u->d_xy, n_feat, nullptr, rank, world);
}

extern "C" int rvio_updater_reduce_buffer(rvio_updater* u, double** buf_dev, int* count)
{
    RVIO_ARG_CHECK(u && buf_dev && count);
    if (!u->open) { set_error("rvio_updater_reduce_buffer", "no update in flight"); return RVIO_ERR_STATE; }
    const int n = 6 * u->cur_N;
    *buf_dev = u->d_red;
    *count = n * n + n + 8 + n + 1;
    return RVIO_OK;
}

extern "C" int rvio_updater_update_finish(rvio_updater* u, double* x_out, double* P_out, rvio_update_info* info)
{
    RVIO_ARG_CHECK(u && x_out && P_out);
    if (!u->open) { set_error("rvio_updater_update_finish", "no update in flight"); return RVIO_ERR_STATE; }
    RVIO_CUDA_TRY(cudaSetDevice(u->device));
    cudaStream_t s = u->stream;
    const int n = 6 * u->cur_N, d = u->cur_d, xdim = u->cur_xdim;
    int rc = updater_enqueue_solve(u, s, u->d_xout, u->d_Pout);
    if (rc != RVIO_OK) return rc;
    RVIO_CUDA_TRY(cudaMemcpyAsync(u->h_red, u->d_red + (size_t)n * n + n, sizeof(double) * 8, cudaMemcpyDeviceToHost, s));
    RVIO_CUDA_TRY(cudaMemcpyAsync(u->h_x, u->d_xout, sizeof(double) * xdim, cudaMemcpyDeviceToHost, s));
    RVIO_CUDA_TRY(cudaMemcpyAsync(u->h_P, u->d_Pout, sizeof(double) * (size_t)d * d, cudaMemcpyDeviceToHost, s));
    RVIO_CUDA_TRY(cudaMemcpyAsync(u->h_sing, u->d_sing, sizeof(int), cudaMemcpyDeviceToHost, s));
    RVIO_CUDA_TRY(cudaStreamSynchronize(s));
    rvio_update_info inf;
    inf.n_feat = u->cur_nfeat; inf.n_good = (int)u->h_red[0]; inf.rows_stacked = (int)u->h_red[1];
    inf.n_reject_init = (int)u->h_red[2]; inf.n_reject_lm = (int)u->h_red[3]; inf.n_reject_gate = (int)u->h_red[4];
    inf.updated = inf.n_good > 2 ? 1 : 0;
    inf.rank = (int)u->h_red[6]; inf.rank_flags = (int)u->h_red[7];
    if (info) *info = inf;
    memcpy(x_out, u->h_x, sizeof(double) * xdim);
    memcpy(P_out, u->h_P, sizeof(double) * (size_t)d * d);
    if (*u->h_sing) { set_error("rvio_updater_update_finish", "singular innovation system"); return RVIO_ERR_STATE; }
    return RVIO_OK;
}

extern "C" int rvio_updater_update(rvio_updater* u, const double* x, int xdim, const double* P, int d,
                                   const uint8_t* types, const int32_t* offsets, const float* xy, int n_feat,
                                   double* x_out, double* P_out, rvio_update_info* info)
{
    int rc = rvio_updater_update_begin(u, x, xdim, P, d, types, offsets, xy, n_feat, 0, 1);
    if (rc != RVIO_OK) return rc;
    return rvio_updater_update_finish(u, x_out, P_out, info);
}

extern "C" int rvio_updater_update_from_tracker(rvio_updater* u, rvio_tracker* trk, const double* x, int xdim,
                                                const double* P, int d, double* x_out, double* P_out, rvio_update_info* info)
{
    RVIO_ARG_CHECK(u && trk);
    if (tracker_device(trk) != u->device) { set_error("rvio_updater_update_from_tracker", "handles on different devices"); return RVIO_ERR_ARG; }
    RVIO_CUDA_TRY(cudaSetDevice(u->device));
    const TrackerBuffers* B = tracker_buffers(trk);
    int n_meas = 0;
    const int n_feat = tracker_update_counts(trk, &n_meas);
    int rc = check_shapes(u, x, xdim, P, d, n_feat, 0, 1);
    if (rc != RVIO_OK) return rc;
    if ((rc = upload_state(u, x, xdim, P, d)) != RVIO_OK) return rc;
    // the tracker finished its stream work before returning (it synchronises to publish its counters)
    rc = updater_enqueue_normal_terms(u, u->stream, u->d_x, xdim, u->d_P, d, B->up_types, B->up_off, B->up_xy, n_feat, nullptr, 0, 1);
    if (rc != RVIO_OK) return rc;
    return rvio_updater_update_finish(u, x_out, P_out, info);
}

extern "C" int rvio_updater_get_debug(rvio_updater* u, int n_feat, uint8_t* status, double* pfinv, double* gamma, int32_t* dof)
{
    RVIO_ARG_CHECK(u && n_feat >= 0 && n_feat <= u->Fmax);
    if (n_feat == 0) return RVIO_OK;
    RVIO_CUDA_TRY(cudaSetDevice(u->device));
    cudaStream_t s = u->stream;
    if (status) RVIO_CUDA_TRY(cudaMemcpyAsync(status, u->d_fstatus, n_feat, cudaMemcpyDeviceToHost, s));
    if (pfinv) RVIO_CUDA_TRY(cudaMemcpyAsync(pfinv, u->d_fpfinv, sizeof(double) * 3 * n_feat, cudaMemcpyDeviceToHost, s));
    if (gamma) RVIO_CUDA_TRY(cudaMemcpyAsync(gamma, u->d_fgamma, sizeof(double) * n_feat, cudaMemcpyDeviceToHost, s));
    if (dof) RVIO_CUDA_TRY(cudaMemcpyAsync(dof, u->d_fdof, sizeof(int32_t) * n_feat, cudaMemcpyDeviceToHost, s));
    RVIO_CUDA_TRY(cudaStreamSynchronize(s));
    return RVIO_OK;
}

extern "C" int rvio_updater_get_normal_terms(rvio_updater* u, double* G, double* z, int n)
{
    RVIO_ARG_CHECK(u && n == 6 * u->cur_N && n > 0);
    RVIO_CUDA_TRY(cudaSetDevice(u->device));
    cudaStream_t s = u->stream;
    if (G) RVIO_CUDA_TRY(cudaMemcpyAsync(G, u->d_red, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, s));
    if (z) RVIO_CUDA_TRY(cudaMemcpyAsync(z, u->d_red + (size_t)n * n, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    RVIO_CUDA_TRY(cudaStreamSynchronize(s));
    return RVIO_OK;
}

extern "C" void* rvio_updater_stream(rvio_updater* u) { return u ? (void*)u->stream : nullptr; }

extern "C" int rvio_updater_set_rank_rule(rvio_updater* u, int mode)
{
    RVIO_ARG_CHECK(u && (mode == RVIO_RANK_RULE_REFERENCE || mode == RVIO_RANK_RULE_FULL_INFORMATION));
    RVIO_CUDA_TRY(cudaSetDevice(u->device));
    u->rank_rule = mode;
    *u->h_rule = mode;
    RVIO_CUDA_TRY(cudaMemcpyAsync(u->d_rule, u->h_rule, sizeof(int), cudaMemcpyHostToDevice, u->stream));
    RVIO_CUDA_TRY(cudaStreamSynchronize(u->stream));
    return RVIO_OK;
}

#ifdef RVIO_B200_PHASE_CLOCKS
extern "C" int rvio_b200_gram_ns(unsigned long long* out, int reset)
{
    cudaDeviceSynchronize();
    int rc = (int)cudaMemcpyFromSymbol(out, rvio::g_gram_ns, sizeof(unsigned long long) * 16);
    if (reset) { unsigned long long z[16]; for (int i = 0; i < 16; ++i) z[i] = 0; z[0] = ~0ull; cudaMemcpyToSymbol(rvio::g_gram_ns, z, sizeof z); }
    return rc;
}
#endif

#ifdef RVIO_B200_PHASE_CLOCKS
extern "C" int rvio_b200_feat_clocks(long long* out)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, rvio::g_feat_clk, sizeof(long long) * 16);
}
#endif
