/*
 * oracle/filter.c -- CPU ORACLE (test infrastructure only; see rvio_oracle.h).
 *
 * Restates the host stages around the hot path so that a per-frame pose can be produced and compared:
 *   PreIntegrator::propagate        src/rvio/PreIntegrator.cc:30-48,51-194
 *   System::initialize              src/rvio/System.cc:115-170
 *   clone augmentation+composition  src/rvio/System.cc:280-365
 * float64, Eigen semantics restated (PARITY UNPINNED by the reference; see rvio_oracle.h).
 * P is column-major with leading dimension d = 24+6N, as Eigen::MatrixXd.
 */
#include "rvio_oracle.h"
#include "linalg.h"
#include <stdlib.h>

#define PP(P, d, i, j) (P)[(size_t)(j) * (d) + (i)]

static void set_blk3(double* M, int ld, int r, int c, const double* B /* row-major 3x3 */, double scale)
{
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[(r + i) * ld + (c + j)] = scale * B[3 * i + j];
}

void orc_propagate(const orc_imu_cfg_t* cfg, const double* x, int xdim, const double* P_in,
                   const double* imu, int n_imu, double* x_out, double* P)
{
    const int N = (xdim - 26) / 7, d = 24 + 6 * N;
    memcpy(P, P_in, sizeof(double) * (size_t)d * d);
    double gk[3], vk[3], bg[3], ba[3], pk[3], gR[3], vR[3];
    memcpy(gk, x + 7, sizeof gk);
    memcpy(pk, x + 14, sizeof pk);
    memcpy(vk, x + 17, sizeof vk);
    memcpy(bg, x + 20, sizeof bg);
    memcpy(ba, x + 23, sizeof ba);
    memcpy(gR, gk, sizeof gR);
    memcpy(vR, vk, sizeof vR);
    double Rk[9], RkT[9];
    orc_quat_to_rot(x + 10, Rk);
    m3_T(Rk, RkT);
    double dp[3] = {0, 0, 0}, dv[3] = {0, 0, 0};
    double F[24 * 24], Phi[24 * 24], Psi[24 * 24], T[24 * 24], G[24 * 12], Q[24 * 24];   /* row-major */
    memset(F, 0, sizeof F);
    memset(G, 0, sizeof G);
    memset(Psi, 0, sizeof Psi);
    for (int i = 0; i < 24; ++i) Psi[i * 24 + i] = 1;
    const double nz[12] = {cfg->sigma_g * cfg->sigma_g, cfg->sigma_g * cfg->sigma_g, cfg->sigma_g * cfg->sigma_g,
                           cfg->sigma_wg * cfg->sigma_wg, cfg->sigma_wg * cfg->sigma_wg, cfg->sigma_wg * cfg->sigma_wg,
                           cfg->sigma_a * cfg->sigma_a, cfg->sigma_a * cfg->sigma_a, cfg->sigma_a * cfg->sigma_a,
                           cfg->sigma_wa * cfg->sigma_wa, cfg->sigma_wa * cfg->sigma_wa, cfg->sigma_wa * cfg->sigma_wa};
    double I3[9];
    m3_eye(I3);
    const double grav = cfg->gravity;
    double Dt = 0;

    for (int s = 0; s < n_imu; ++s) {
        const double* wm = imu + 8 * s;
        const double* am = imu + 8 * s + 3;
        const double dt = imu[8 * s + 7];
        Dt += dt;
        double w[3] = {wm[0] - bg[0], wm[1] - bg[1], wm[2] - bg[2]};
        double a[3] = {am[0] - ba[0], am[1] - ba[1], am[2] - ba[2]};
        const double w1 = v3_norm(w);
        const int small = w1 < cfg->small_angle;
        const double wdt = w1 * dt, wdt2 = wdt * wdt, cw = cos(wdt), sw = sin(wdt);
        double wx[9], wx2[9], vx[9], gx[9], M[9];
        skew(w, wx);
        m3_mul(wx, wx, wx2);
        skew(vk, vx);
        skew(gk, gx);
        /* PreIntegrator.cc:119-131 */
        set_blk3(F, 24, 9, 9, wx, -1);
        set_blk3(F, 24, 9, 18, I3, -1);
        m3_mul(RkT, vx, M);
        set_blk3(F, 24, 12, 9, M, -1);
        set_blk3(F, 24, 12, 15, RkT, 1);
        set_blk3(F, 24, 15, 6, Rk, -grav);
        set_blk3(F, 24, 15, 9, gx, -grav);
        set_blk3(F, 24, 15, 15, wx, -1);
        set_blk3(F, 24, 15, 18, vx, -1);
        set_blk3(F, 24, 15, 21, I3, -1);
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 24; ++j) Phi[i * 24 + j] = ((i == j) ? 1.0 : 0.0) + dt * F[i * 24 + j];
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 24; ++j) {
                double acc = 0;
                for (int k = 0; k < 24; ++k) acc += Phi[i * 24 + k] * Psi[k * 24 + j];
                T[i * 24 + j] = acc;
            }
        memcpy(Psi, T, sizeof T);
        /* PreIntegrator.cc:133-138 */
        set_blk3(G, 12, 9, 0, I3, -1);
        set_blk3(G, 12, 15, 0, vx, -1);
        set_blk3(G, 12, 15, 6, I3, -1);
        set_blk3(G, 12, 18, 3, I3, 1);
        set_blk3(G, 12, 21, 9, I3, 1);
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 24; ++j) {
                double acc = 0;
                for (int k = 0; k < 12; ++k) acc += (dt * G[i * 12 + k]) * nz[k] * G[j * 12 + k];
                Q[i * 24 + j] = acc;
            }
        /* PreIntegrator.cc:140 */
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 24; ++j) {
                double acc = 0;
                for (int k = 0; k < 24; ++k) acc += Phi[i * 24 + k] * PP(P, d, k, j);
                T[i * 24 + j] = acc;
            }
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 24; ++j) {
                double acc = 0;
                for (int k = 0; k < 24; ++k) acc += T[i * 24 + k] * Phi[j * 24 + k];
                Q[i * 24 + j] = acc + Q[i * 24 + j];
            }
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < 24; ++j) PP(P, d, i, j) = Q[i * 24 + j];

        /* State, PreIntegrator.cc:142-178 */
        double dR[9], f1, f2, f3, f4;
        if (small) {
            const double c2 = (dt * dt) / 2;
            for (int i = 0; i < 9; ++i) dR[i] = I3[i] - dt * wx[i] + c2 * wx2[i];
            f1 = -(dt * dt * dt) / 3;
            f2 = (dt * dt * dt * dt) / 8;
            f3 = -(dt * dt) / 2;
            f4 = (dt * dt * dt) / 6;
        } else {
            const double c1 = sw / w1, c2 = (1 - cw) / (w1 * w1);
            for (int i = 0; i < 9; ++i) dR[i] = I3[i] - c1 * wx[i] + c2 * wx2[i];
            f1 = (wdt * cw - sw) / (w1 * w1 * w1);
            f2 = .5 * (wdt2 - 2 * cw - 2 * wdt * sw + 2) / (w1 * w1 * w1 * w1);
            f3 = (cw - 1) / (w1 * w1);
            f4 = (wdt - sw) / (w1 * w1 * w1);
        }
        m3_mul(dR, Rk, Rk);
        m3_T(Rk, RkT);
        double A1[9], A2[9], t1[3], t2[3], B[9];
        const double hdt2 = .5 * (dt * dt);
        for (int i = 0; i < 9; ++i) {
            A1[i] = hdt2 * I3[i] + f1 * wx[i] + f2 * wx2[i];
            A2[i] = dt * I3[i] + f3 * wx[i] + f4 * wx2[i];
        }
        for (int k = 0; k < 3; ++k) dp[k] += dv[k] * dt;
        m3_mul(RkT, A1, B);
        m3_v(B, a, t1);
        for (int k = 0; k < 3; ++k) dp[k] += t1[k];
        m3_mul(RkT, A2, B);
        m3_v(B, a, t2);
        for (int k = 0; k < 3; ++k) dv[k] += t2[k];
        double u[3];
        for (int k = 0; k < 3; ++k) {
            pk[k] = vR[k] * Dt - .5 * grav * gR[k] * (Dt * Dt) + dp[k];
            u[k] = vR[k] - grav * gR[k] * Dt + dv[k];
        }
        m3_v(Rk, u, vk);
        m3_v(Rk, gR, gk);
        double gn = v3_norm(gk);
        for (int k = 0; k < 3; ++k) gk[k] /= gn;
    }
    memcpy(x_out, x, sizeof(double) * (size_t)xdim);
    orc_rot_to_quat(Rk, x_out + 10);
    memcpy(x_out + 14, pk, sizeof pk);
    memcpy(x_out + 17, vk, sizeof vk);
    if (N > 0) {   /* PreIntegrator.cc:186-191 */
        const int n = 6 * N;
        double* C = (double*)malloc(sizeof(double) * 24 * (size_t)n);
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0;
                for (int k = 0; k < 24; ++k) acc += Psi[i * 24 + k] * PP(P, d, k, 24 + j);
                C[(size_t)i * n + j] = acc;
            }
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < n; ++j) { PP(P, d, i, 24 + j) = C[(size_t)i * n + j]; PP(P, d, 24 + j, i) = C[(size_t)i * n + j]; }
        free(C);
    }
    for (int i = 0; i < d; ++i)
        for (int j = i + 1; j < d; ++j) { double v = .5 * (PP(P, d, i, j) + PP(P, d, j, i)); PP(P, d, i, j) = v; PP(P, d, j, i) = v; }
    for (int i = 0; i < d; ++i) PP(P, d, i, i) = .5 * (PP(P, d, i, i) + PP(P, d, i, i));
}

void orc_initialize(const orc_imu_cfg_t* cfg, double imu_rate, const double* w, const double* a, int n_imu_data,
                    int enable_alignment, double* x, double* P)
{
    double g[3] = {a[0], a[1], a[2]};
    double gn = v3_norm(g);
    for (int k = 0; k < 3; ++k) g[k] /= gn;
    double R[9];
    m3_eye(R);
    if (enable_alignment) {   /* System.cc:122-141 */
        const double* zv = g;
        double ex[3] = {1, 0, 0};
        double zz[9], t[3], xv[3], yv[3], zx[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) zz[3 * i + j] = zv[i] * zv[j];
        m3_v(zz, ex, t);
        for (int k = 0; k < 3; ++k) xv[k] = ex[k] - t[k];
        double n1 = v3_norm(xv);
        for (int k = 0; k < 3; ++k) xv[k] /= n1;
        skew(zv, zx);
        m3_v(zx, xv, yv);
        double n2 = v3_norm(yv);
        for (int k = 0; k < 3; ++k) yv[k] /= n2;
        for (int i = 0; i < 3; ++i) { R[3 * i] = xv[i]; R[3 * i + 1] = yv[i]; R[3 * i + 2] = zv[i]; }
    }
    memset(x, 0, 26 * sizeof(double));
    orc_rot_to_quat(R, x);
    memcpy(x + 7, g, sizeof g);
    if (n_imu_data > 1) {
        for (int k = 0; k < 3; ++k) { x[20 + k] = w[k]; x[23 + k] = a[k] - cfg->gravity * g[k]; }
    }
    const double dt = 1. / imu_rate;
    memset(P, 0, 24 * 24 * sizeof(double));
    for (int k = 0; k < 6; ++k) P[k * 24 + k] = 1e-3 * 1e-3;
    for (int k = 6; k < 9; ++k) P[k * 24 + k] = n_imu_data * dt * (cfg->sigma_a * cfg->sigma_a);
    for (int k = 18; k < 21; ++k) P[k * 24 + k] = n_imu_data * dt * (cfg->sigma_wg * cfg->sigma_wg);
    for (int k = 21; k < 24; ++k) P[k * 24 + k] = n_imu_data * dt * (cfg->sigma_wa * cfg->sigma_wa);
}

void orc_augment_compose(double* x, double* P, int* n_clones, int window, int do_augment, double* pose_out)
{
    int N = *n_clones;
    int d = 24 + 6 * N;
    if (do_augment) {   /* System.cc:280-323 */
        const int d1 = d + 6;
        double* T = (double*)calloc((size_t)d1 * d1, sizeof(double));   /* J P J^T, column-major ld d1 */
        /* row map: new row i <- old row src(i) */
        int* src = (int*)malloc(sizeof(int) * (size_t)d1);
        for (int i = 0; i < d; ++i) src[i] = i;
        for (int k = 0; k < 3; ++k) { src[d + k] = 9 + k; src[d + 3 + k] = 12 + k; }
        for (int j = 0; j < d1; ++j)
            for (int i = 0; i < d1; ++i) T[(size_t)j * d1 + i] = PP(P, d, src[i], src[j]);
        for (int i = 0; i < d1; ++i)
            for (int j = i + 1; j < d1; ++j) { double v = .5 * (T[(size_t)j * d1 + i] + T[(size_t)i * d1 + j]); T[(size_t)j * d1 + i] = v; T[(size_t)i * d1 + j] = v; }
        double clone[7];
        memcpy(clone, x + 10, sizeof clone);
        if (N < window) {
            memcpy(x + 26 + 7 * N, clone, sizeof clone);
            memcpy(P, T, sizeof(double) * (size_t)d1 * d1);
            N++;
            d = d1;
        } else {
            memmove(x + 26, x + 33, sizeof(double) * 7 * (size_t)(window - 1));
            memcpy(x + 26 + 7 * (window - 1), clone, sizeof clone);
            /* drop rows/cols 24..29 of T */
            for (int j = 0; j < d; ++j)
                for (int i = 0; i < d; ++i) {
                    int si = i < 24 ? i : i + 6, sj = j < 24 ? j : j + 6;
                    PP(P, d, i, j) = T[(size_t)sj * d1 + si];
                }
        }
        free(T); free(src);
    }
    *n_clones = N;
    /* Composition, System.cc:326-365 */
    double qG[4], pG[3], gk[3], qk[4], pk[3];
    memcpy(qG, x, sizeof qG); memcpy(pG, x + 4, sizeof pG); memcpy(gk, x + 7, sizeof gk);
    memcpy(qk, x + 10, sizeof qk); memcpy(pk, x + 14, sizeof pk);
    double RG[9], Rk[9];
    orc_quat_to_rot(qG, RG);
    orc_quat_to_rot(qk, Rk);
    double t[3];
    m3_v(Rk, gk, t);
    double gn = v3_norm(t);
    for (int k = 0; k < 3; ++k) gk[k] = t[k] / gn;
    double qkG[4], pkG[3], pGk[3], dv[3];
    orc_quat_mul(qk, qG, qkG);
    for (int k = 0; k < 3; ++k) dv[k] = pG[k] - pk[k];
    m3_v(Rk, dv, pkG);
    for (int k = 0; k < 3; ++k) dv[k] = pk[k] - pG[k];
    m3T_v(RG, dv, pGk);

    double V[24 * 24];
    memset(V, 0, sizeof V);
    double I3[9], S1[9], S2[9];
    m3_eye(I3);
    skew(pkG, S1);
    skew(gk, S2);
    set_blk3(V, 24, 0, 0, Rk, 1);
    set_blk3(V, 24, 0, 9, I3, 1);
    set_blk3(V, 24, 3, 3, Rk, 1);
    set_blk3(V, 24, 3, 9, S1, 1);
    set_blk3(V, 24, 3, 12, Rk, -1);
    set_blk3(V, 24, 6, 6, Rk, 1);
    set_blk3(V, 24, 6, 9, S2, 1);
    for (int k = 15; k < 24; ++k) V[k * 24 + k] = 1;

    const int n = 6 * N;
    double T1[24 * 24], T2[24 * 24];
    for (int i = 0; i < 24; ++i)
        for (int j = 0; j < 24; ++j) {
            double acc = 0;
            for (int k = 0; k < 24; ++k) acc += V[i * 24 + k] * PP(P, d, k, j);
            T1[i * 24 + j] = acc;
        }
    for (int i = 0; i < 24; ++i)
        for (int j = 0; j < 24; ++j) {
            double acc = 0;
            for (int k = 0; k < 24; ++k) acc += T1[i * 24 + k] * V[j * 24 + k];
            T2[i * 24 + j] = acc;
        }
    for (int i = 0; i < 24; ++i)
        for (int j = 0; j < 24; ++j) PP(P, d, i, j) = T2[i * 24 + j];
    if (n > 0) {
        double* C = (double*)malloc(sizeof(double) * 24 * (size_t)n);
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < n; ++j) {
                double acc = 0;
                for (int k = 0; k < 24; ++k) acc += V[i * 24 + k] * PP(P, d, k, 24 + j);
                C[(size_t)i * n + j] = acc;
            }
        for (int i = 0; i < 24; ++i)
            for (int j = 0; j < n; ++j) { PP(P, d, i, 24 + j) = C[(size_t)i * n + j]; PP(P, d, 24 + j, i) = C[(size_t)i * n + j]; }
        free(C);
    }
    for (int i = 0; i < d; ++i)
        for (int j = i + 1; j < d; ++j) { double v = .5 * (PP(P, d, i, j) + PP(P, d, j, i)); PP(P, d, i, j) = v; PP(P, d, j, i) = v; }

    memcpy(x, qkG, sizeof qkG);
    memcpy(x + 4, pkG, sizeof pkG);
    memcpy(x + 7, gk, sizeof gk);
    x[10] = x[11] = x[12] = 0; x[13] = 1;
    x[14] = x[15] = x[16] = 0;
    if (pose_out) { memcpy(pose_out, pGk, sizeof pGk); memcpy(pose_out + 3, qkG, sizeof qkG); }
}
