/*
 * oracle/ransac.c -- CPU ORACLE (test infrastructure only; see rvio_oracle.h).
 *
 * Restates src/rvio/Ransac.cc:32-266 (gyro-aided 2-point RANSAC), src/util/Numerics.h quaternion helpers,
 * and glibc's rand() (TYPE_3 additive-feedback generator of random_r.c, the generator behind the
 * never-seeded rand() calls at Ransac.cc:63,69; pinned against libc.so.6 in tests/test_oracle_ransac.py).
 * Eigen stages: parity unpinned by the reference (no tests, cannot be built here) -- see rvio_oracle.h.
 */
#include "rvio_oracle.h"
#include "linalg.h"
#include <stdlib.h>

/* ------------------------------------------------------------------ glibc rand() */
void orc_rand_seed(orc_rand_t* st, unsigned seed)
{
    if (seed == 0) seed = 1;
    int32_t* r = st->r;
    r[0] = (int32_t)seed;
    for (int i = 1; i < 31; ++i) {
        long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        r[i] = (int32_t)word;
    }
    st->f = 3; st->b = 0;
    for (int i = 0; i < 310; ++i) (void)orc_rand_next(st);
}

int orc_rand_next(orc_rand_t* st)
{
    uint32_t v = (uint32_t)st->r[st->f] + (uint32_t)st->r[st->b];
    st->r[st->f] = (int32_t)v;
    int res = (int)(v >> 1);
    if (++st->f >= 31) { st->f = 0; ++st->b; }
    else if (++st->b >= 31) st->b = 0;
    return res;
}

/* ------------------------------------------------------------------ Numerics.h */
void orc_quat_mul(const double* q1, const double* q2, double* out)   /* Numerics.h:30-63 */
{
    double q[4];
    q[0] =  q1[3] * q2[0] + q1[2] * q2[1] - q1[1] * q2[2] + q1[0] * q2[3];
    q[1] = -q1[2] * q2[0] + q1[3] * q2[1] + q1[0] * q2[2] + q1[1] * q2[3];
    q[2] =  q1[1] * q2[0] - q1[0] * q2[1] + q1[3] * q2[2] + q1[2] * q2[3];
    q[3] = -q1[0] * q2[0] - q1[1] * q2[1] - q1[2] * q2[2] + q1[3] * q2[3];
    double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= nrm;
    if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
    memcpy(out, q, sizeof q);
}

void orc_quat_to_rot(const double* q, double* R)   /* Numerics.h:111-120: I - 2w[q x] + 2[q x]^2 */
{
    double qx[9], qx2[9];
    skew(q, qx);
    m3_mul(qx, qx, qx2);
    for (int i = 0; i < 9; ++i) {
        double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        R[i] = I - 2 * q[3] * qx[i] + 2 * qx2[i];
    }
}

void orc_rot_to_quat(const double* R, double* q)   /* Numerics.h:126-167 (Breckenridge) */
{
    const double T = R[0] + R[4] + R[8];
    if (R[0] > T && R[0] > R[4] && R[0] > R[8]) {
        q[0] = sqrt((1 + 2 * R[0] - T) / 4);
        q[1] = (1 / (4 * q[0])) * (R[1] + R[3]);
        q[2] = (1 / (4 * q[0])) * (R[2] + R[6]);
        q[3] = (1 / (4 * q[0])) * (R[5] - R[7]);
    } else if (R[4] > T && R[4] > R[0] && R[4] > R[8]) {
        q[1] = sqrt((1 + 2 * R[4] - T) / 4);
        q[0] = (1 / (4 * q[1])) * (R[1] + R[3]);
        q[2] = (1 / (4 * q[1])) * (R[5] + R[7]);
        q[3] = (1 / (4 * q[1])) * (R[6] - R[2]);
    } else if (R[8] > T && R[8] > R[0] && R[8] > R[4]) {
        q[2] = sqrt((1 + 2 * R[8] - T) / 4);
        q[0] = (1 / (4 * q[2])) * (R[2] + R[6]);
        q[1] = (1 / (4 * q[2])) * (R[5] + R[7]);
        q[3] = (1 / (4 * q[2])) * (R[1] - R[3]);
    } else {
        q[3] = sqrt((1 + T) / 4);
        q[0] = (1 / (4 * q[3])) * (R[5] - R[7]);
        q[1] = (1 / (4 * q[3])) * (R[6] - R[2]);
        q[2] = (1 / (4 * q[3])) * (R[1] - R[3]);
    }
    double nrm = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= nrm;
    if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
}

/* ------------------------------------------------------------------ Ransac */
void orc_ransac_init(orc_ransac_t* rs, int use_sampson, double thr, double small_angle, const double* T)
{
    memset(rs, 0, sizeof *rs);
    rs->use_sampson = use_sampson;
    rs->inlier_thr = thr;
    rs->small_angle = small_angle;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) rs->Ric[3 * i + j] = T[4 * i + j];   /* Ransac.cc:41-46 */
    orc_rand_seed(&rs->rng, 1);                                          /* rand() never seeded */
}

/* Ransac.cc:120-155 */
static void get_rotation(const orc_ransac_t* rs, const double* imu, int n_imu, double* R)
{
    double tempR[9], I[9];
    m3_eye(tempR); m3_eye(I);
    for (int k = 0; k < n_imu; ++k) {
        const double* wm = imu + 8 * k;
        const double dt = imu[8 * k + 7];
        const double w1 = v3_norm(wm);
        const int small = w1 < rs->small_angle;
        const double wdt = w1 * dt;
        double wx[9], wx2[9], dR[9];
        skew(wm, wx);
        m3_mul(wx, wx, wx2);
        if (small) {
            const double c2 = .5 * (dt * dt);
            for (int i = 0; i < 9; ++i) dR[i] = I[i] - dt * wx[i] + c2 * wx2[i];
        } else {
            const double c1 = sin(wdt) / w1, c2 = (1 - cos(wdt)) / (w1 * w1);
            for (int i = 0; i < 9; ++i) dR[i] = I[i] - c1 * wx[i] + c2 * wx2[i];
        }
        m3_mul(dR, tempR, tempR);
    }
    double Rci[9], T[9];
    m3_T(rs->Ric, Rci);
    m3_mul(Rci, tempR, T);
    m3_mul(T, rs->Ric, R);
}

/* Ransac.cc:250-266 */
static double sampson(const double* p1, const double* p2, const double* E)
{
    double Fx1[3], Fx2[3], p2E[3];
    m3_v(E, p1, Fx1);
    m3T_v(E, p2, Fx2);
    m3T_v(E, p2, p2E);      /* pt2^T * E as a row vector */
    double num = p2E[0] * p1[0] + p2E[1] * p1[1] + p2E[2] * p1[2];
    return (num * num) / (Fx1[0] * Fx1[0] + Fx1[1] * Fx1[1] + Fx2[0] * Fx2[0] + Fx2[1] * Fx2[1]);
}
static double algebraic(const double* p1, const double* p2, const double* E)
{
    double p2E[3];
    m3T_v(E, p2, p2E);
    return fabs(p2E[0] * p1[0] + p2E[1] * p1[1] + p2E[2] * p1[2]);
}

int orc_ransac_find_inliers(orc_ransac_t* rs, const double* P1, const double* P2, int F,
                            const double* imu, int n_imu, uint8_t* flags)
{
    const int NIT = 16;
    memset(rs->hyp, 0, sizeof rs->hyp);
    memset(rs->n_inliers, 0, sizeof rs->n_inliers);
    memset(rs->two_points, 0, sizeof rs->two_points);
    rs->winner = 0;

    int* cand = (int*)malloc(sizeof(int) * (size_t)(F > 0 ? F : 1));
    int nc = 0;
    for (int i = 0; i < F; ++i)
        if (flags[i]) cand[nc++] = i;
    if (nc <= NIT) { free(cand); return 0; }            /* Ransac.cc:201-205: flags untouched */
    if (nc < 2 * NIT) { free(cand); return -1; }        /* 17..31: reference never returns; defined as "untouched" */

    /* SetPointPair, Ransac.cc:50-83 */
    {
        int* v = (int*)malloc(sizeof(int) * (size_t)nc);
        for (int i = 0; i < nc; ++i) v[i] = i;
        for (int it = 0; it < NIT; ++it) {
            int a, b;
            do { a = orc_rand_next(&rs->rng) % nc; } while (v[a] == -1);
            do { b = orc_rand_next(&rs->rng) % nc; } while (v[b] == -1 || a == b);
            rs->two_points[2 * it] = cand[v[a]];
            rs->two_points[2 * it + 1] = cand[v[b]];
            v[a] = -1; v[b] = -1;
        }
        free(v);
    }
    get_rotation(rs, imu, n_imu, rs->R);

    int best = 0, best_idx = 0;
    for (int it = 0; it < NIT; ++it) {
        /* SetRansacModel, Ransac.cc:86-117 */
        const double* A1 = P1 + 3 * rs->two_points[2 * it];
        const double* A2 = P2 + 3 * rs->two_points[2 * it];
        const double* B1 = P1 + 3 * rs->two_points[2 * it + 1];
        const double* B2 = P2 + 3 * rs->two_points[2 * it + 1];
        double A0[3], B0[3];
        m3_v(rs->R, A1, A0);
        m3_v(rs->R, B1, B0);
        double c1 = A2[0] * A0[1] - A0[0] * A2[1];
        double c2 = A0[1] * A2[2] - A2[1] * A0[2];
        double c3 = A2[0] * A0[2] - A0[0] * A2[2];
        double c4 = B2[0] * B0[1] - B0[0] * B2[1];
        double c5 = B0[1] * B2[2] - B2[1] * B0[2];
        double c6 = B2[0] * B0[2] - B0[0] * B2[2];
        double alpha = atan2(c3 * c5 - c2 * c6, c1 * c6 - c3 * c4);
        double beta = atan2(-c3, c1 * sin(alpha) + c2 * cos(alpha));
        double t[3] = {sin(beta) * cos(alpha), cos(beta), -sin(beta) * sin(alpha)};
        double tx[9];
        skew(t, tx);
        double* E = rs->hyp + 9 * it;
        m3_mul(tx, rs->R, E);
        /* CountInliers, Ransac.cc:158-177 */
        for (int k = 0; k < nc; ++k) {
            int idx = cand[k];
            double d = rs->use_sampson ? sampson(P1 + 3 * idx, P2 + 3 * idx, E) : algebraic(P1 + 3 * idx, P2 + 3 * idx, E);
            if (d < rs->inlier_thr) rs->n_inliers[it] += 1;
        }
        if (rs->n_inliers[it] > best) { best = rs->n_inliers[it]; best_idx = it; }
    }
    rs->winner = best_idx;
    const double* W = rs->hyp + 9 * best_idx;
    int new_out = 0;
    for (int k = 0; k < nc; ++k) {
        int idx = cand[k];
        double d = rs->use_sampson ? sampson(P1 + 3 * idx, P2 + 3 * idx, W) : algebraic(P1 + 3 * idx, P2 + 3 * idx, W);
        if (d > rs->inlier_thr || isnan(d)) { flags[idx] = 0; new_out++; }
    }
    free(cand);
    return nc - new_out;
}
