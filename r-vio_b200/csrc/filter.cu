// filter.cu -- device-resident versions of the small stages either side of the hot path (SURVEY 8f "next" rows 2,3
// and the refill half of row 1), so that x / P never leave the GPU between track() and update():
//   k_propagate        PreIntegrator::propagate            reference src/rvio/PreIntegrator.cc:51-194
//   k_augment_compose  clone augmentation + composition    reference src/rvio/System.cc:280-365
//   k_find_newer_refill FeatureDetector::FindNewer + refill reference src/rvio/FeatureDetector.cc:78-150, Tracker.cc:358-386
// float64 throughout (single CTA each: 24x24 dense algebra, d <= 204).  Compiled with -fmad=false so that the float32
// cell arithmetic of FindNewer matches the host code exactly.
#include "common.cuh"
#include "tracker_kernels.cuh"
#include "filter_kernels.cuh"
#include "device_utils.cuh"

#include <mutex>

namespace rvio {

#define PP(P, d, i, j) (P)[(size_t)(j) * (d) + (i)]

__device__ __forceinline__ void f_m3mul(const double* A, const double* B, double* C)
{
    double T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    for (int i = 0; i < 9; ++i) C[i] = T[i];
}
__device__ __forceinline__ void f_m3v(const double* A, const double* v, double* o)
{
    const double t0 = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
    const double t1 = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
    const double t2 = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
    o[0] = t0; o[1] = t1; o[2] = t2;
}
__device__ __forceinline__ void f_skew(const double* w, double* M)
{
    M[0] = 0; M[1] = -w[2]; M[2] = w[1]; M[3] = w[2]; M[4] = 0; M[5] = -w[0]; M[6] = -w[1]; M[7] = w[0]; M[8] = 0;
}
__device__ __forceinline__ void f_quat_to_rot(const double* q, double* R)
{
    double qx[9], qx2[9];
    f_skew(q, qx);
    f_m3mul(qx, qx, qx2);
    for (int i = 0; i < 9; ++i) R[i] = ((i == 0 || i == 4 || i == 8) ? 1.0 : 0.0) - 2 * q[3] * qx[i] + 2 * qx2[i];
}
__device__ __forceinline__ void f_quat_mul(const double* q1, const double* q2, double* out)
{
    double q[4];
    q[0] = q1[3] * q2[0] + q1[2] * q2[1] - q1[1] * q2[2] + q1[0] * q2[3];
    q[1] = -q1[2] * q2[0] + q1[3] * q2[1] + q1[0] * q2[2] + q1[1] * q2[3];
    q[2] = q1[1] * q2[0] - q1[0] * q2[1] + q1[3] * q2[2] + q1[2] * q2[3];
    q[3] = -q1[0] * q2[0] - q1[1] * q2[1] - q1[2] * q2[2] + q1[3] * q2[3];
    const double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= nrm;
    const double sg = q[3] < 0 ? -1.0 : 1.0;
    for (int i = 0; i < 4; ++i) out[i] = sg * q[i];
}
__device__ __forceinline__ void f_rot_to_quat(const double* R, double* q)
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
    for (int i = 0; i < 4; ++i) q[i] /= nrm;
    if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
}

// ------------------------------------------------------------------------------------------------
// k_propagate: 576 threads, thread (i,j) of the 24x24 blocks.  P_in -> P_out (full d x d), x_in -> x_out.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void set_blk3(double (*M)[24], int r, int c, const double* B, double scale)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[r + i][c + j] = scale * B[3 * i + j];
}

__global__ void __launch_bounds__(576) k_propagate(PropagateParams Q)
{
    __shared__ double F[24][24], Phi[24][24], Psi[24][24], Tm[24][24], Pb[24][24];
    __shared__ double G[24][12];
    __shared__ double s_dt;
    __shared__ double s_st[40];     // Rk(9) RkT(9) vk(3) gk(3) pk(3) dp(3) dv(3) ...
    const int tid = threadIdx.x, i = tid / 24, j = tid % 24;
    const int d = Q.d, n = d - 24;
    const double* x = Q.x_in;
    F[i][j] = 0; Psi[i][j] = (i == j) ? 1.0 : 0.0;
    if (j < 12) G[i][j] = 0;
    Pb[i][j] = PP(Q.P_in, d, i, j);
    double* Rk = s_st; double* RkT = s_st + 9; double* vk = s_st + 18; double* gk = s_st + 21; double* pk = s_st + 24;
    double* dp = s_st + 27; double* dv = s_st + 30;
    __shared__ double s_fix[16];    // gR(3) vR(3) bg(3) ba(3) Dt
    if (tid == 0) {
        for (int k = 0; k < 3; ++k) {
            gk[k] = x[7 + k]; pk[k] = x[14 + k]; vk[k] = x[17 + k];
            s_fix[k] = x[7 + k]; s_fix[3 + k] = x[17 + k]; s_fix[6 + k] = x[20 + k]; s_fix[9 + k] = x[23 + k];
            dp[k] = 0; dv[k] = 0;
        }
        s_fix[12] = 0;
        f_quat_to_rot(x + 10, Rk);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) RkT[3 * a + b] = Rk[3 * b + a];
    }
    __syncthreads();
    const double grav = Q.c.gravity;
    const double nz[12] = {Q.c.sigma_g * Q.c.sigma_g, Q.c.sigma_g * Q.c.sigma_g, Q.c.sigma_g * Q.c.sigma_g,
                           Q.c.sigma_wg * Q.c.sigma_wg, Q.c.sigma_wg * Q.c.sigma_wg, Q.c.sigma_wg * Q.c.sigma_wg,
                           Q.c.sigma_a * Q.c.sigma_a, Q.c.sigma_a * Q.c.sigma_a, Q.c.sigma_a * Q.c.sigma_a,
                           Q.c.sigma_wa * Q.c.sigma_wa, Q.c.sigma_wa * Q.c.sigma_wa, Q.c.sigma_wa * Q.c.sigma_wa};
    __shared__ double s_w[3], s_a[3];
    const int n_imu = Q.hdr ? Q.hdr[0] : Q.n_imu;
    for (int s = 0; s < n_imu; ++s) {
        if (tid == 0) {
            const double* wm = Q.imu + 8 * s;
            const double* am = wm + 3;
            const double dt = wm[7];
            s_dt = dt;
            const double* bg = s_fix + 6; const double* ba = s_fix + 9;
            double w[3] = {wm[0] - bg[0], wm[1] - bg[1], wm[2] - bg[2]};
            for (int k = 0; k < 3; ++k) { s_w[k] = w[k]; s_a[k] = am[k] - ba[k]; }
            double wx[9], vx[9], gx[9], I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, M[9];
            f_skew(w, wx); f_skew(vk, vx); f_skew(gk, gx);
            // PreIntegrator.cc:119-131
            set_blk3(F, 9, 9, wx, -1); set_blk3(F, 9, 18, I3, -1);
            f_m3mul(RkT, vx, M);
            set_blk3(F, 12, 9, M, -1); set_blk3(F, 12, 15, RkT, 1);
            set_blk3(F, 15, 6, Rk, -grav); set_blk3(F, 15, 9, gx, -grav);
            set_blk3(F, 15, 15, wx, -1); set_blk3(F, 15, 18, vx, -1); set_blk3(F, 15, 21, I3, -1);
            // PreIntegrator.cc:133-137
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) {
                    G[9 + a][b] = -I3[3 * a + b]; G[15 + a][b] = -vx[3 * a + b]; G[15 + a][6 + b] = -I3[3 * a + b];
                    G[18 + a][3 + b] = I3[3 * a + b]; G[21 + a][9 + b] = I3[3 * a + b];
                }
        }
        __syncthreads();
        const double dt = s_dt;
        Phi[i][j] = ((i == j) ? 1.0 : 0.0) + dt * F[i][j];
        __syncthreads();
        {
            double acc = 0;
            for (int k = 0; k < 24; ++k) acc += Phi[i][k] * Psi[k][j];
            double acc2 = 0;
            for (int k = 0; k < 24; ++k) acc2 += Phi[i][k] * Pb[k][j];
            __syncthreads();
            Psi[i][j] = acc;
            Tm[i][j] = acc2;
        }
        __syncthreads();
        {
            double acc = 0;
            for (int k = 0; k < 24; ++k) acc += Tm[i][k] * Phi[j][k];
            double q = 0;
            for (int k = 0; k < 12; ++k) q += (dt * G[i][k]) * nz[k] * G[j][k];
            __syncthreads();
            Pb[i][j] = acc + q;
        }
        // state, PreIntegrator.cc:142-178
        if (tid == 0) {
            const double* w = s_w; const double* a = s_a;
            const double* gR = s_fix; const double* vR = s_fix + 3;
            const double w1 = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
            const bool small = w1 < Q.c.small_angle;
            const double wdt = w1 * dt, wdt2 = wdt * wdt, cw = cos(wdt), sw = sin(wdt);
            double wx[9], wx2[9], dR[9], f1, f2, f3, f4;
            const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
            f_skew(w, wx);
            f_m3mul(wx, wx, wx2);
            if (small) {
                const double c2 = (dt * dt) / 2;
                for (int k = 0; k < 9; ++k) dR[k] = I3[k] - dt * wx[k] + c2 * wx2[k];
                f1 = -(dt * dt * dt) / 3; f2 = (dt * dt * dt * dt) / 8; f3 = -(dt * dt) / 2; f4 = (dt * dt * dt) / 6;
            } else {
                const double c1 = sw / w1, c2 = (1 - cw) / (w1 * w1);
                for (int k = 0; k < 9; ++k) dR[k] = I3[k] - c1 * wx[k] + c2 * wx2[k];
                f1 = (wdt * cw - sw) / (w1 * w1 * w1);
                f2 = .5 * (wdt2 - 2 * cw - 2 * wdt * sw + 2) / (w1 * w1 * w1 * w1);
                f3 = (cw - 1) / (w1 * w1);
                f4 = (wdt - sw) / (w1 * w1 * w1);
            }
            f_m3mul(dR, Rk, Rk);
            for (int p = 0; p < 3; ++p)
                for (int q = 0; q < 3; ++q) RkT[3 * p + q] = Rk[3 * q + p];
            double A1[9], A2[9], B[9], t1[3], t2[3];
            const double hdt2 = .5 * (dt * dt);
            for (int k = 0; k < 9; ++k) {
                A1[k] = hdt2 * I3[k] + f1 * wx[k] + f2 * wx2[k];
                A2[k] = dt * I3[k] + f3 * wx[k] + f4 * wx2[k];
            }
            for (int k = 0; k < 3; ++k) dp[k] += dv[k] * dt;
            f_m3mul(RkT, A1, B); f_m3v(B, a, t1);
            for (int k = 0; k < 3; ++k) dp[k] += t1[k];
            f_m3mul(RkT, A2, B); f_m3v(B, a, t2);
            for (int k = 0; k < 3; ++k) dv[k] += t2[k];
            s_fix[12] += dt;
            const double Dt = s_fix[12];
            double u[3];
            for (int k = 0; k < 3; ++k) {
                pk[k] = vR[k] * Dt - .5 * grav * gR[k] * (Dt * Dt) + dp[k];
                u[k] = vR[k] - grav * gR[k] * Dt + dv[k];
            }
            f_m3v(Rk, u, vk);
            f_m3v(Rk, gR, gk);
            const double gn = sqrt(gk[0] * gk[0] + gk[1] * gk[1] + gk[2] * gk[2]);
            for (int k = 0; k < 3; ++k) gk[k] /= gn;
        }
        __syncthreads();
    }
    // state out
    for (int k = tid; k < Q.xdim; k += 576) Q.x_out[k] = x[k];
    __syncthreads();
    if (tid == 0) {
        double q[4];
        f_rot_to_quat(Rk, q);
        for (int k = 0; k < 4; ++k) Q.x_out[10 + k] = q[k];
        for (int k = 0; k < 3; ++k) { Q.x_out[14 + k] = pk[k]; Q.x_out[17 + k] = vk[k]; }
    }
    // covariance out: [ Pb  Psi*P0c ; (.)^T  Pcc ], then .5 (P + P^T)   (PreIntegrator.cc:183-193)
    double* Po = Q.P_out;
    for (int o = tid; o < 24 * n; o += 576) {
        const int r = o / n, c = o % n;
        double acc = 0;
        for (int k = 0; k < 24; ++k) acc += Psi[r][k] * PP(Q.P_in, d, k, 24 + c);
        PP(Po, d, r, 24 + c) = acc;
        PP(Po, d, 24 + c, r) = acc;
    }
    {
        const double v = .5 * (Pb[i][j] + Pb[j][i]);
        PP(Po, d, i, j) = v;
    }
    for (int o = tid; o < n * n; o += 576) {
        const int r = o / n, c = o % n;
        PP(Po, d, 24 + r, 24 + c) = .5 * (PP(Q.P_in, d, 24 + r, 24 + c) + PP(Q.P_in, d, 24 + c, 24 + r));
    }
}

// ------------------------------------------------------------------------------------------------
// k_augment_compose: P_in (d x d) -> P_out (d' x d'), x in place.  do_augment / slide decided on the host
// (deterministic counters, System.cc:280-323), composition System.cc:326-365.
// ------------------------------------------------------------------------------------------------
#ifdef RVIO_B200_PHASE_CLOCKS
__device__ long long g_aug_clk[16];
#define AUG_CLK(k) do { if (threadIdx.x == 0) g_aug_clk[k] = clock64(); } while (0)
#else
#define AUG_CLK(k) do { } while (0)
#endif
constexpr int kCrossCols = 96, kCrossLd = kCrossCols + 1;          // clone columns staged per pass (+1: conflict-free column walks)
__global__ void __launch_bounds__(576) k_augment_compose(AugmentParams Q)
{
    __shared__ double V[24][24], T1[24][24], P00[24][24];
    __shared__ double s_q[16];
    __shared__ double s_cross[24 * kCrossLd];
    const int tid = threadIdx.x, i = tid / 24, j = tid % 24;
    const int d = Q.d, N = Q.N, W = Q.window;
    double* x = Q.x;
    int dn = d, Nn = N;                      // new dimension / clone count
    AUG_CLK(0);
    // ---- augmentation
    if (Q.do_augment) {
        const int d1 = d + 6;
        const bool slide = !(N < W);
        dn = slide ? d : d1;
        Nn = slide ? N : N + 1;
        // element (a,b) of J P J^T, symmetrised, with the oldest clone removed when the window is full; four elements per
        // thread and pass (eight independent loads in flight: the loop is bound by the latency of the transposed reads)
        for (int o0 = tid; o0 < dn * dn; o0 += 4 * 576) {
            double u[4], w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int o = o0 + q * 576;
                u[q] = 0; w[q] = 0;
                if (o < dn * dn) {
                    const int a = o % dn, b = o / dn;
                    int sa = a, sb = b;
                    if (slide) { if (a >= 24) sa = a + 6; if (b >= 24) sb = b + 6; }
                    // index in the (d+6) system -> source row/col of P
                    const int ra = sa < d ? sa : (sa - d < 3 ? 9 + (sa - d) : 12 + (sa - d - 3));
                    const int rb = sb < d ? sb : (sb - d < 3 ? 9 + (sb - d) : 12 + (sb - d - 3));
                    u[q] = PP(Q.P_in, d, ra, rb); w[q] = PP(Q.P_in, d, rb, ra);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int o = o0 + q * 576;
                if (o < dn * dn) Q.P_out[o] = .5 * (u[q] + w[q]);
            }
        }
        // state: the new clone is the current IMU pose (x[10..16]); a full window drops its oldest clone first.  One element
        // per thread, loads before the barrier, stores after it
        AUG_CLK(1);
        const int nshift = slide ? 7 * (W - 1) : 0;
        double keep = 0, cl = 0;
        if (tid < nshift) keep = x[33 + tid];
        if (tid >= 576 - 7) cl = x[10 + tid - (576 - 7)];
        __syncthreads();
        if (tid < nshift) x[26 + tid] = keep;
        if (tid >= 576 - 7) x[(slide ? 26 + 7 * (W - 1) : 26 + 7 * N) + tid - (576 - 7)] = cl;
    } else {
        for (int o = tid; o < d * d; o += 576) Q.P_out[o] = Q.P_in[o];
    }
    __syncthreads();
    double* P = Q.P_out;
    const int n = 6 * Nn;
    AUG_CLK(2);
    // ---- composition
    V[i][j] = 0.0;
    __syncthreads();
    if (tid == 0) {
        double qG[4], pG[3], gk[3], qk[4], pk[3], RG[9], Rk[9], t[3];
        for (int k = 0; k < 4; ++k) { qG[k] = x[k]; qk[k] = x[10 + k]; }
        for (int k = 0; k < 3; ++k) { pG[k] = x[4 + k]; gk[k] = x[7 + k]; pk[k] = x[14 + k]; }
        f_quat_to_rot(qG, RG);
        f_quat_to_rot(qk, Rk);
        f_m3v(Rk, gk, t);
        const double gn = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
        for (int k = 0; k < 3; ++k) gk[k] = t[k] / gn;
        double qkG[4], pkG[3], pGk[3], dv[3];
        f_quat_mul(qk, qG, qkG);
        for (int k = 0; k < 3; ++k) dv[k] = pG[k] - pk[k];
        f_m3v(Rk, dv, pkG);
        for (int k = 0; k < 3; ++k) dv[k] = pk[k] - pG[k];
        for (int k = 0; k < 3; ++k) pGk[k] = RG[k] * dv[0] + RG[3 + k] * dv[1] + RG[6 + k] * dv[2];   // RG^T dv
        double S1[9], S2[9];
        f_skew(pkG, S1); f_skew(gk, S2);
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                V[a][b] = Rk[3 * a + b]; V[a][9 + b] = (a == b) ? 1.0 : 0.0;
                V[3 + a][3 + b] = Rk[3 * a + b]; V[3 + a][9 + b] = S1[3 * a + b]; V[3 + a][12 + b] = -Rk[3 * a + b];
                V[6 + a][6 + b] = Rk[3 * a + b]; V[6 + a][9 + b] = S2[3 * a + b];
            }
        for (int k = 15; k < 24; ++k) V[k][k] = 1;
        for (int k = 0; k < 4; ++k) s_q[k] = qkG[k];
        for (int k = 0; k < 3; ++k) { s_q[4 + k] = pkG[k]; s_q[7 + k] = gk[k]; s_q[10 + k] = pGk[k]; }
    }
    P00[i][j] = PP(P, dn, i, j);
    __syncthreads();
    AUG_CLK(3);
    {
        double acc = 0;
        for (int k = 0; k < 24; ++k) acc += V[i][k] * P00[k][j];
        T1[i][j] = acc;
    }
    __syncthreads();
    {
        double acc = 0;
        for (int k = 0; k < 24; ++k) acc += T1[i][k] * V[j][k];
        P00[i][j] = acc;
    }
    __syncthreads();
    PP(P, dn, i, j) = .5 * (P00[i][j] + P00[j][i]);
    AUG_CLK(4);
    // cross terms V P0c, kCrossCols columns per pass: the block P(0:24, 24 + c0 ..) is staged in shared memory (rows are
    // contiguous in memory: coalesced), then one output element per thread (same summation order as before: k ascending)
    for (int c0 = 0; c0 < n; c0 += kCrossCols) {
        const int nc = min(kCrossCols, n - c0);
        for (int o = tid; o < 24 * nc; o += 576) { const int k = o / nc, c = o - k * nc; s_cross[k * kCrossLd + c] = PP(P, dn, k, 24 + c0 + c); }
        __syncthreads();
        for (int o = tid; o < 24 * nc; o += 576) {
            const int r = o / nc, c = o - r * nc;
            double acc = 0;
#pragma unroll
            for (int k = 0; k < 24; ++k) acc += V[r][k] * s_cross[k * kCrossLd + c];
            PP(P, dn, r, 24 + c0 + c) = acc;
            PP(P, dn, 24 + c0 + c, r) = acc;
        }
        __syncthreads();
    }
    AUG_CLK(5);
    // clone-clone block is already symmetric (augmentation symmetrised it; propagate/update symmetrise their outputs)
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) x[k] = s_q[k];
        for (int k = 0; k < 3; ++k) { x[4 + k] = s_q[4 + k]; x[7 + k] = s_q[7 + k]; }
        x[10] = x[11] = x[12] = 0; x[13] = 1;
        x[14] = x[15] = x[16] = 0;
        for (int k = 0; k < 3; ++k) Q.pose_out[k] = s_q[10 + k];
        for (int k = 0; k < 4; ++k) Q.pose_out[3 + k] = s_q[k];
    }
    AUG_CLK(6);
}

// ------------------------------------------------------------------------------------------------
// k_find_newer_refill: FeatureDetector::FindNewer (one warp per grid cell) + Tracker refill, single CTA.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int fn_cell(const FindNewerParams& Q, float2 p)
{
    // FeatureDetector.cc:83-91 / :101-105 ; returns -1 outside the grid area
    if (p.x <= (float)Q.offx || p.y <= (float)Q.offy || p.x >= (float)(Q.W - Q.offx) || p.y >= (float)(Q.H - Q.offy)) return -1;
    const int col = (int)floorf(__fdiv_rn(__fsub_rn(p.x, (float)Q.offx), Q.bx));
    const int row = (int)floorf(__fdiv_rn(__fsub_rn(p.y, (float)Q.offy), Q.by));
    return row * Q.gc + col;
}

__global__ void __launch_bounds__(1024) k_find_newer_refill(FindNewerParams Q)
{
    extern __shared__ __align__(8) unsigned char s_dyn[];
    __shared__ int sh[34];
    __shared__ float2 s_list[32][kFindNewerCellCap];  // accepted points per warp's current cell
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    TrackerBuffers& B = Q.B;
    TrackerScalars* sc = B.sc;
    const int n_ref = sc->n_new;
    const int nc = Q.n_cand_dev ? min(*Q.n_cand_dev, Q.n_cand) : Q.n_cand;
    // dynamic shared memory: candidates, their grid cells, the cells of the features already there, accepted flags
    float2* s_cand = reinterpret_cast<float2*>(s_dyn);                              // Q.n_cand
    short* s_cell = reinterpret_cast<short*>(s_cand + Q.n_cand);                    // Q.n_cand
    short* s_rcell = s_cell + Q.n_cand;                                             // B.F
    unsigned char* s_acc = reinterpret_cast<unsigned char*>(s_rcell + B.F);         // Q.n_cand
    for (int k = tid; k < nc; k += 1024) { const float2 p = Q.cand[k]; s_cand[k] = p; s_cell[k] = (short)fn_cell(Q, p); s_acc[k] = 0; }
    for (int r = tid; r < n_ref; r += 1024) s_rcell[r] = (short)fn_cell(Q, B.feats_new[r]);
    __syncthreads();
    const int n_cells = Q.gc * Q.gr;
    const double lim = .75 * (double)Q.max_per_block;
    if (sc->fq_n > 0 && !Q.raw) {
        for (int cell = warp; cell < n_cells; cell += 32) {
            // refs already in this cell
            int cnt = 0;
            for (int r = lane; r < n_ref; r += 32) cnt += (s_rcell[r] == cell) ? 1 : 0;
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            const int col = cell % Q.gc, row = cell / Q.gc;
            const float xl = __fadd_rn(__fmul_rn((float)col, Q.bx), (float)Q.offx), xr = __fadd_rn(xl, Q.bx);
            const float yt = __fadd_rn(__fmul_rn((float)row, Q.by), (float)Q.offy), yb = __fadd_rn(yt, Q.by);
            int n_acc = 0;
            for (int k0 = 0; k0 < nc; k0 += 32) {
                // the candidates of this cell among k0 .. k0+31, visited in candidate order (FeatureDetector.cc:99)
                unsigned in_cell = __ballot_sync(0xffffffffu, k0 + lane < nc && s_cell[k0 + lane] == cell);
                while (in_cell) {
                    const int k = k0 + __ffs(in_cell) - 1;
                    in_cell &= in_cell - 1;
                    const float2 p = s_cand[k];
                    if (fabs((double)__fsub_rn(p.x, xl)) < (double)Q.min_dist || fabs((double)__fsub_rn(p.x, xr)) < (double)Q.min_dist ||
                        fabs((double)__fsub_rn(p.y, yt)) < (double)Q.min_dist || fabs((double)__fsub_rn(p.y, yb)) < (double)Q.min_dist) continue;
                    if (!((double)(float)(cnt + n_acc) < lim)) continue;
                    // every point of the cell must be farther than min_dist (FeatureDetector.cc:125-134)
                    int close = 0;
                    for (int r = lane; r < n_ref; r += 32) {
                        if (s_rcell[r] != cell) continue;
                        const float2 q = B.feats_new[r];
                        const double dx = (double)__fsub_rn(p.x, q.x), dy = (double)__fsub_rn(p.y, q.y);
                        if (!(sqrt(dx * dx + dy * dy) > (double)Q.min_dist)) close = 1;
                    }
                    for (int a = lane; a < n_acc; a += 32) {
                        const float2 q = s_list[warp][a];
                        const double dx = (double)__fsub_rn(p.x, q.x), dy = (double)__fsub_rn(p.y, q.y);
                        if (!(sqrt(dx * dx + dy * dy) > (double)Q.min_dist)) close = 1;
                    }
                    close = __any_sync(0xffffffffu, close);
                    if (!close) {
                        if (lane == 0) { s_acc[k] = 1; if (n_acc < kFindNewerCellCap) s_list[warp][n_acc] = p; }
                        n_acc++;
                        __syncwarp();
                    }
                }
            }
        }
    } else if (Q.raw) {
        for (int k = tid; k < nc; k += 1024) s_acc[k] = 1;       // candidates already filtered by the host
    }
    __syncthreads();
    // ---- refill in candidate order (Tracker.cc:358-386)
    const int head = sc->fq_head, fq_n = sc->fq_n, n_new0 = sc->n_new;
    int base = 0;
    for (int start = 0; start < nc; start += 1024) {
        const int k = start + tid;
        const int f = (k < nc) ? s_acc[k] : 0;
        int tot;
        const int ex = block_exscan_1024(f, sh, &tot);
        const int r = base + ex;
        if (f && r < fq_n) {
            const int slot = B.freeq[(head + r) % (B.F + 1)];
            const float2 p = s_cand[k];
            float ux, uy;
            cam_undistort(Q.cam, p.x, p.y, &ux, &uy);
            B.slots_new[n_new0 + r] = slot;
            B.feats_new[n_new0 + r] = p;
            B.pts1_new[n_new0 + r] = make_float2(ux, uy);
            const int len = B.hist_len[slot];
            B.hist[(size_t)slot * B.hist_cap + (B.hist_head[slot] + len) % B.hist_cap] = make_float2(ux, uy);
            B.hist_len[slot] = len + 1;
        }
        base += tot;
    }
    __syncthreads();
    if (tid == 0) {
        const int use = base < fq_n ? base : fq_n;
        sc->fq_head = (head + use) % (B.F + 1);
        sc->fq_n = fq_n - use;
        sc->n_new = n_new0 + use;
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
int launch_propagate(cudaStream_t s, const PropagateParams& p)
{
    RVIO_LAUNCH(k_propagate, 1, 576, 0, s, p);
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}
int launch_augment_compose(cudaStream_t s, const AugmentParams& p)
{
    RVIO_LAUNCH(k_augment_compose, 1, 576, 0, s, p);
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}
int launch_find_newer_refill(cudaStream_t s, const FindNewerParams& p)
{
    static std::once_flag once;
    std::call_once(once, [] { cudaFuncSetAttribute(k_find_newer_refill, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); });
    RVIO_LAUNCH(k_find_newer_refill, 1, 1024, (size_t)p.n_cand * (8 + 2 + 1) + (size_t)p.B.F * 2 + 64, s, p);
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

}  // namespace rvio

#ifdef RVIO_B200_PHASE_CLOCKS
extern "C" int rvio_b200_aug_clocks(long long* out)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, rvio::g_aug_clk, sizeof(long long) * 16);
}
#endif
