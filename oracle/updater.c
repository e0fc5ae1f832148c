/*
 * oracle/updater.c -- CPU ORACLE (test infrastructure only; see rvio_oracle.h).
 *
 * Restates src/rvio/Updater.cc:38-628 (MSCKF-style robocentric measurement update) in float64, in the
 * reference's operation order: per-feature relative-pose chain (:118-141), inverse-depth LM triangulation
 * (:146-269), Jacobian blocks (:271-368), Givens left-nullspace projection (:370-402), chi^2 gate (:404-455),
 * stacking, Givens QR compression (:460-536), EKF gain / state correction / Joseph covariance (:540-619).
 * Eigen primitives (makeGivens, colPivHouseholderQr().solve, inverse()) are restated from Eigen's public
 * semantics (Eigen is not vendored and absent here): PARITY UNPINNED by the reference for this file; it is
 * cross-checked against oracle/np_updater.py (NumPy/LAPACK).
 * rviz landmark publishing (:78-87,430-448,458) is out of scope.
 */
#include "rvio_oracle.h"
#include "linalg.h"
#include "chi2_table.h"
#include <stdlib.h>
#include <float.h>

static int g_full_info = 0;
/* 0 (default): the reference's rule (rows up to the first row with norm < 1e-4, Updater.cc:515-524);
 * 1: keep every row with norm >= 1e-4 (no information discarded) -- used to quantify the rule's effect. */
void orc_updater_set_rank_rule(int full_info) { g_full_info = full_info; }

double orc_chi2_95(int dof) { return (dof >= 1 && dof <= 500) ? ORC_CHI2_95[dof - 1] : NAN; }

void orc_updater_cfg_init(orc_updater_cfg_t* c, float sx, float sy, const double* T)
{
    c->sigma = (double)(sx > sy ? sx : sy);            /* Updater.cc:42-44 */
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) c->Ric[3 * i + j] = T[4 * i + j];
        c->tic[i] = T[4 * i + 3];
    }
}

/* Eigen JacobiRotation::makeGivens (real), G^* [p q]^T = [r 0]^T */
static void make_givens(double p, double q, double* c, double* s)
{
    if (q == 0) { *c = p < 0 ? -1 : 1; *s = 0; }
    else if (p == 0) { *c = 0; *s = q < 0 ? 1 : -1; }
    else if (fabs(p) > fabs(q)) {
        double t = q / p, u = sqrt(1 + t * t);
        if (p < 0) u = -u;
        *c = 1 / u; *s = -t * (*c);
    } else {
        double t = p / q, u = sqrt(1 + t * t);
        if (q < 0) u = -u;
        *s = -1 / u; *c = -t * (*s);
    }
}
/* rows x,y of length n (stride 1): x' = c x - s y ; y' = s x + c y */
static void rot_rows(double* x, double* y, int n, double c, double s)
{
    if (c == 1 && s == 0) return;
    for (int i = 0; i < n; ++i) {
        double xi = x[i], yi = y[i];
        x[i] = c * xi - s * yi;
        y[i] = s * xi + c * yi;
    }
}

/* Solve A x = b (n x n, row-major, A destroyed) by column-pivoted Householder QR, the semantics of
 * Eigen's colPivHouseholderQr().solve for a full-rank system (Updater.cc:239,420). */
static void colpiv_qr_solve(double* A, int n, const double* b, double* x)
{
    int* perm = (int*)malloc(sizeof(int) * (size_t)n);
    double* c = (double*)malloc(sizeof(double) * (size_t)n);
    double* v = (double*)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; ++i) { perm[i] = i; c[i] = b[i]; }
    int rank = n;
    double maxnorm0 = 0;
    for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += A[i * n + j] * A[i * n + j];
        if (s > maxnorm0) maxnorm0 = s;
    }
    const double tiny = maxnorm0 * (DBL_EPSILON / n) * (DBL_EPSILON / n);
    for (int k = 0; k < n; ++k) {
        int best = k; double bn = -1;
        for (int j = k; j < n; ++j) {
            double s = 0;
            for (int i = k; i < n; ++i) s += A[i * n + j] * A[i * n + j];
            if (s > bn) { bn = s; best = j; }
        }
        if (rank == n && bn < tiny * (double)(n - k)) rank = k;
        if (best != k) {
            for (int i = 0; i < n; ++i) { double t = A[i * n + k]; A[i * n + k] = A[i * n + best]; A[i * n + best] = t; }
            int t = perm[k]; perm[k] = perm[best]; perm[best] = t;
        }
        /* Householder on column k, rows k..n-1 */
        double tail = 0;
        for (int i = k + 1; i < n; ++i) tail += A[i * n + k] * A[i * n + k];
        double a0 = A[k * n + k];
        if (tail <= DBL_MIN) continue;
        double beta = sqrt(a0 * a0 + tail);
        if (a0 >= 0) beta = -beta;
        double tau = (beta - a0) / beta;
        v[k] = 1;
        for (int i = k + 1; i < n; ++i) v[i] = A[i * n + k] / (a0 - beta);
        A[k * n + k] = beta;
        for (int i = k + 1; i < n; ++i) A[i * n + k] = 0;
        for (int j = k + 1; j < n; ++j) {
            double d = 0;
            for (int i = k; i < n; ++i) d += v[i] * A[i * n + j];
            d *= tau;
            for (int i = k; i < n; ++i) A[i * n + j] -= d * v[i];
        }
        double d = 0;
        for (int i = k; i < n; ++i) d += v[i] * c[i];
        d *= tau;
        for (int i = k; i < n; ++i) c[i] -= d * v[i];
    }
    for (int i = 0; i < n; ++i) x[i] = 0;
    for (int i = rank - 1; i >= 0; --i) {
        double s = c[i];
        for (int j = i + 1; j < rank; ++j) s -= A[i * n + j] * c[j];
        c[i] = s / A[i * n + i];
    }
    for (int i = 0; i < rank; ++i) x[perm[i]] = c[i];
    free(perm); free(c); free(v);
}

/* In-place inverse by partial-pivot LU (Eigen MatrixXd::inverse(), Updater.cc:543). A n x n row-major. */
static void lu_inverse(double* A, int n)
{
    int* piv = (int*)malloc(sizeof(int) * (size_t)n);
    double* B = (double*)calloc((size_t)n * n, sizeof(double));
    for (int k = 0; k < n; ++k) {
        int p = k; double mx = fabs(A[k * n + k]);
        for (int i = k + 1; i < n; ++i) if (fabs(A[i * n + k]) > mx) { mx = fabs(A[i * n + k]); p = i; }
        piv[k] = p;
        if (p != k) for (int j = 0; j < n; ++j) { double t = A[k * n + j]; A[k * n + j] = A[p * n + j]; A[p * n + j] = t; }
        double d = A[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            double f = A[i * n + k] / d;
            A[i * n + k] = f;
            for (int j = k + 1; j < n; ++j) A[i * n + j] -= f * A[k * n + j];
        }
    }
    /* B = P (permuted identity), then solve L U X = B */
    for (int i = 0; i < n; ++i) B[i * n + i] = 1;
    for (int k = 0; k < n; ++k)
        if (piv[k] != k) for (int j = 0; j < n; ++j) { double t = B[k * n + j]; B[k * n + j] = B[piv[k] * n + j]; B[piv[k] * n + j] = t; }
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < i; ++k) { double f = A[i * n + k]; if (f != 0) for (int j = 0; j < n; ++j) B[i * n + j] -= f * B[k * n + j]; }
    for (int i = n - 1; i >= 0; --i) {
        for (int k = i + 1; k < n; ++k) { double f = A[i * n + k]; if (f != 0) for (int j = 0; j < n; ++j) B[i * n + j] -= f * B[k * n + j]; }
        double d = A[i * n + i];
        for (int j = 0; j < n; ++j) B[i * n + j] /= d;
    }
    memcpy(A, B, sizeof(double) * (size_t)n * n);
    free(piv); free(B);
}

static void hproj(const double* h, double* Hp /* 2x3 */)
{
    Hp[0] = 1 / h[2]; Hp[1] = 0; Hp[2] = -h[0] / (h[2] * h[2]);
    Hp[3] = 0; Hp[4] = 1 / h[2]; Hp[5] = -h[1] / (h[2] * h[2]);
}
static void m23_m33(const double* A, const double* B, double* C)
{
    double T[6];
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(C, T, sizeof T);
}
static void set_dir(double phi, double psi, double* e, double* Jang /* 3x2 row-major */)
{
    e[0] = cos(phi) * sin(psi); e[1] = sin(phi); e[2] = cos(phi) * cos(psi);
    Jang[0] = -sin(phi) * sin(psi); Jang[1] = cos(phi) * cos(psi);
    Jang[2] = cos(phi); Jang[3] = 0;
    Jang[4] = -sin(phi) * cos(psi); Jang[5] = -cos(phi) * sin(psi);
}

void orc_updater_update(const orc_updater_cfg_t* cfg, const double* x, int xdim, const double* P,
                        const uint8_t* types, const int32_t* offsets, const float* xy, int n_feat,
                        double* x_out, double* P_out, orc_update_info_t* info,
                        uint8_t* feat_status, double* feat_pfinv, double* feat_gamma,
                        double* Hstack_out, double* rstack_out)
{
    const int N = (xdim - 26) / 7;               /* Updater.cc:98 */
    const int n = 6 * N, d = 24 + n;
    const double sigma = cfg->sigma, sig2 = sigma * sigma;
    double Rci[9], tci[3];
    m3_T(cfg->Ric, Rci);
    m3_v(Rci, cfg->tic, tci);
    for (int i = 0; i < 3; ++i) tci[i] = -tci[i];  /* Updater.cc:53 */

    int nRows = 0;
    for (int f = 0; f < n_feat; ++f) nRows += 2 * (offsets[f + 1] - offsets[f]);
    double* r = (double*)calloc((size_t)(nRows > 0 ? nRows : 1), sizeof(double));
    double* Hx = (double*)calloc((size_t)(nRows > 0 ? nRows : 1) * (size_t)(n > 0 ? n : 1), sizeof(double));   /* clone columns only */
    int nRowCount = 0, nGood = 0;
    orc_update_info_t inf;
    memset(&inf, 0, sizeof inf);
    inf.n_feat = n_feat;

    /* P_cc in row-major (P is symmetric in practice but keep the exact element mapping) */
    double* Pcc = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Pcc[i * n + j] = P[(size_t)(24 + j) * d + (24 + i)];

    const int Lcap = N + 2;
    double* relI = (double*)malloc(sizeof(double) * 7 * (size_t)Lcap);
    double* relC = (double*)malloc(sizeof(double) * 7 * (size_t)Lcap);
    double* RI = (double*)malloc(sizeof(double) * 9 * (size_t)Lcap);    /* QuatToRot(relI.q) */
    double* RC = (double*)malloc(sizeof(double) * 9 * (size_t)Lcap);    /* QuatToRot(relC.q) */

    for (int f = 0; f < n_feat; ++f) {
        if (feat_status) feat_status[f] = 0;
        if (feat_gamma) feat_gamma[f] = NAN;
        if (feat_pfinv) feat_pfinv[3 * f] = feat_pfinv[3 * f + 1] = feat_pfinv[3 * f + 2] = NAN;
        const char type = (char)types[f];
        const float* meas = xy + 2 * (size_t)offsets[f];
        int L = offsets[f + 1] - offsets[f];
        int phases = L - 1;
        const double* rel = (type == '1') ? (x + xdim - 7 * phases) : (x + 26);   /* Updater.cc:118-122 */

        /* [qIi1,tIi1], Updater.cc:124-132 */
        {
            double R0[9], t[3];
            orc_quat_to_rot(rel, R0);
            m3_v(R0, rel + 4, t);
            memcpy(relI, rel, 4 * sizeof(double));
            relI[4] = -t[0]; relI[5] = -t[1]; relI[6] = -t[2];
            for (int i = 1; i < phases; ++i) {
                double Ri[9], dv[3];
                orc_quat_mul(rel + 7 * i, relI + 7 * (i - 1), relI + 7 * i);
                orc_quat_to_rot(rel + 7 * i, Ri);
                for (int k = 0; k < 3; ++k) dv[k] = relI[7 * (i - 1) + 4 + k] - rel[7 * i + 4 + k];
                m3_v(Ri, dv, relI + 7 * i + 4);
            }
        }
        /* [qCi1,tCi1], Updater.cc:134-141 */
        for (int i = 0; i < phases; ++i) {
            double T[9], M[9], a[3], b[3];
            orc_quat_to_rot(relI + 7 * i, RI + 9 * i);
            m3_mul(Rci, RI + 9 * i, T);
            m3_mul(T, cfg->Ric, M);
            orc_rot_to_quat(M, relC + 7 * i);
            m3_v(T, cfg->tic, a);
            m3_v(Rci, relI + 7 * i + 4, b);
            for (int k = 0; k < 3; ++k) relC[7 * i + 4 + k] = a[k] + b[k] + tci[k];
            orc_quat_to_rot(relC + 7 * i, RC + 9 * i);
        }

        /* Feature initialization, Updater.cc:143-158 */
        const float fx0 = meas[0], fy0 = meas[1];
        double phi = atan2((double)fy0, sqrt((double)fx0 * (double)fx0 + 1));
        double psi = atan2((double)fx0, 1);
        double rho = 0.;
        if (fabs(phi) > .5 * 3.14 || fabs(psi) > .5 * 3.14) {
            if (feat_status) feat_status[f] = 1;
            inf.n_reject_init++;
            continue;
        }
        double e[3], Jang[6];
        set_dir(phi, psi, e, Jang);
        const double rinv = 1. / (sigma * sigma);

        /* LM refinement, Updater.cc:173-263 */
        double lambda = 0.01, lastCost = INFINITY;
        for (int it = 0; it < 10; ++it) {
            double A[9] = {0}, g[3] = {0}, cost = 0;
            for (int i = 0; i <= phases; ++i) {
                double h[3], Hp[6], Hm[6];        /* Hm: 2x3 measurement Jacobian wrt [phi psi rho] */
                if (i == 0) {
                    memcpy(h, e, sizeof h);
                    hproj(h, Hp);
                    for (int a = 0; a < 2; ++a) {
                        for (int b = 0; b < 2; ++b) Hm[3 * a + b] = Hp[3 * a] * Jang[b] + Hp[3 * a + 1] * Jang[2 + b] + Hp[3 * a + 2] * Jang[4 + b];
                        Hm[3 * a + 2] = 0;
                    }
                } else {
                    const double* Rc = RC + 9 * (i - 1);
                    const double* tc = relC + 7 * (i - 1) + 4;
                    m3_v(Rc, e, h);
                    for (int k = 0; k < 3; ++k) h[k] = h[k] + rho * tc[k];
                    hproj(h, Hp);
                    double HR[6];
                    m23_m33(Hp, Rc, HR);
                    for (int a = 0; a < 2; ++a) {
                        for (int b = 0; b < 2; ++b) Hm[3 * a + b] = HR[3 * a] * Jang[b] + HR[3 * a + 1] * Jang[2 + b] + HR[3 * a + 2] * Jang[4 + b];
                        Hm[3 * a + 2] = Hp[3 * a] * tc[0] + Hp[3 * a + 1] * tc[1] + Hp[3 * a + 2] * tc[2];
                    }
                }
                float ptx = (float)(h[0] / h[2]), pty = (float)(h[1] / h[2]);
                double e0 = (double)(meas[2 * i] - ptx), e1 = (double)(meas[2 * i + 1] - pty);   /* float subtraction */
                cost += (e0 * rinv) * e0 + (e1 * rinv) * e1;
                for (int a = 0; a < 3; ++a) {
                    double ha0 = Hm[a] * rinv, ha1 = Hm[3 + a] * rinv;
                    for (int b = 0; b < 3; ++b) A[3 * a + b] += ha0 * Hm[b] + ha1 * Hm[3 + b];
                    g[a] += ha0 * e0 + ha1 * e1;
                }
            }
            if (cost <= lastCost) {
                double dp[3], At[9];
                for (int k = 0; k < 3; ++k) A[4 * k] += lambda * A[4 * k];
                memcpy(At, A, sizeof At);
                colpiv_qr_solve(At, 3, g, dp);
                phi += dp[0]; psi += dp[1]; rho += dp[2];
                set_dir(phi, psi, e, Jang);
                if (fabs(lastCost - cost) < 1e-6 && dp[2] < 1e-6) break;
                lambda *= .1;
                lastCost = cost;
            } else {
                lambda *= 10;
                lastCost = cost;
            }
        }
        if (feat_pfinv) { feat_pfinv[3 * f] = phi; feat_pfinv[3 * f + 1] = psi; feat_pfinv[3 * f + 2] = rho; }
        if (fabs(phi) > .5 * 3.14 || fabs(psi) > .5 * 3.14 || isinf(rho) || rho < 0) {
            if (feat_status) feat_status[f] = 2;
            inf.n_reject_lm++;
            continue;
        }
        const int phases_full = phases;
        if (type == '2') { L = (int)ceil(.5 * L); phases = L - 1; }    /* Updater.cc:271-275 */

        /* Jacobians, Updater.cc:281-368 */
        const int M = 2 * L;
        double* tr = (double*)calloc((size_t)M, sizeof(double));
        double* tHx = (double*)calloc((size_t)M * (size_t)(n > 0 ? n : 1), sizeof(double));
        double* tHf = (double*)calloc((size_t)M * 3, sizeof(double));
        const int startCol = (type == '1') ? 6 * (N - phases_full) : 0;
        {
            double Hp[6];
            hproj(e, Hp);
            float ptx = (float)(e[0] / e[2]), pty = (float)(e[1] / e[2]);
            tr[0] = (double)(meas[0] - ptx); tr[1] = (double)(meas[1] - pty);
            for (int a = 0; a < 2; ++a) {
                for (int b = 0; b < 2; ++b) tHf[3 * a + b] = Hp[3 * a] * Jang[b] + Hp[3 * a + 1] * Jang[2 + b] + Hp[3 * a + 2] * Jang[4 + b];
                tHf[3 * a + 2] = 0;
            }
        }
        double Rice[3];
        m3_v(cfg->Ric, e, Rice);
        for (int i = 1; i < L; ++i) {
            const double* R = RI + 9 * (i - 1);
            const double* Rc = RC + 9 * (i - 1);
            const double* tc = relC + 7 * (i - 1) + 4;
            double h[3], Hp[6];
            m3_v(Rc, e, h);
            for (int k = 0; k < 3; ++k) h[k] = h[k] + rho * tc[k];
            float ptx = (float)(h[0] / h[2]), pty = (float)(h[1] / h[2]);
            hproj(h, Hp);
            tr[2 * i] = (double)(meas[2 * i] - ptx); tr[2 * i + 1] = (double)(meas[2 * i + 1] - pty);
            double HRci[6], HRR[6];
            m23_m33(Hp, Rci, HRci);
            m23_m33(HRci, R, HRR);                 /* Hproj*mRci*R */
            for (int j = 0; j < i; ++j) {
                double RjT[9], sub[18], v[3], tmp[3], dpx[9], blkL[9];
                m3_T(RI + 9 * j, RjT);
                m3_v(RjT, relI + 7 * j + 4, tmp);
                for (int k = 0; k < 3; ++k) v[k] = Rice[k] + rho * cfg->tic[k] + rho * tmp[k];
                skew(v, dpx);
                m3_mul(dpx, RjT, blkL);
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b) {
                        sub[6 * a + b] = blkL[3 * a + b];
                        if (j == 0) sub[6 * a + 3 + b] = -rho * ((a == b) ? 1.0 : 0.0);
                        else sub[6 * a + 3 + b] = -rho * RI[9 * (j - 1) + 3 * b + a];   /* -rho*R_{j-1}^T */
                    }
                for (int a = 0; a < 2; ++a)
                    for (int b = 0; b < 6; ++b)
                        tHx[(size_t)(2 * i + a) * n + startCol + 6 * j + b] =
                            HRR[3 * a] * sub[b] + HRR[3 * a + 1] * sub[6 + b] + HRR[3 * a + 2] * sub[12 + b];
            }
            double HR[6];
            m23_m33(Hp, Rc, HR);
            for (int a = 0; a < 2; ++a) {
                for (int b = 0; b < 2; ++b) tHf[3 * (2 * i + a) + b] = HR[3 * a] * Jang[b] + HR[3 * a + 1] * Jang[2 + b] + HR[3 * a + 2] * Jang[4 + b];
                tHf[3 * (2 * i + a) + 2] = Hp[3 * a] * tc[0] + Hp[3 * a + 1] * tc[1] + Hp[3 * a + 2] * tc[2];
            }
        }

        /* Nullspace projection, Updater.cc:370-402 */
        int Nc = 3;
        {
            double s = 0;
            for (int i = 0; i < M; ++i) s += tHf[3 * i + 2] * tHf[3 * i + 2];
            if (sqrt(s) < 1e-4) Nc--;
        }
        for (int c0 = 0; c0 < Nc; ++c0)
            for (int m = M - 1; m > c0; --m) {
                double c, s;
                make_givens(tHf[3 * (m - 1) + c0], tHf[3 * m + c0], &c, &s);
                rot_rows(tHf + 3 * (m - 1) + c0, tHf + 3 * m + c0, Nc - c0, c, s);
                rot_rows(tHx + (size_t)(m - 1) * n, tHx + (size_t)m * n, n, c, s);
                rot_rows(tr + (m - 1), tr + m, 1, c, s);
            }

        /* Mahalanobis gate, Updater.cc:404-455 */
        const int dof = M - Nc;
        const double* r_ = tr + Nc;
        const double* H_ = tHx + (size_t)Nc * n;
        double* T = (double*)malloc(sizeof(double) * (size_t)dof * (size_t)(n > 0 ? n : 1));
        double* S = (double*)malloc(sizeof(double) * (size_t)dof * dof);
        for (int i = 0; i < dof; ++i)
            for (int j = 0; j < n; ++j) {
                double a = 0;
                for (int k = 0; k < n; ++k) a += H_[(size_t)i * n + k] * Pcc[k * n + j];
                T[(size_t)i * n + j] = a;
            }
        for (int i = 0; i < dof; ++i)
            for (int j = 0; j < dof; ++j) {
                double a = 0;
                for (int k = 0; k < n; ++k) a += T[(size_t)i * n + k] * H_[(size_t)j * n + k];
                S[i * dof + j] = a;
            }
        for (int i = 0; i < dof; ++i) S[i * dof + i] += sig2;
        for (int i = 0; i < dof; ++i)
            for (int j = i + 1; j < dof; ++j) { double a = .5 * (S[i * dof + j] + S[j * dof + i]); S[i * dof + j] = S[j * dof + i] = a; }
        double* sol = (double*)malloc(sizeof(double) * (size_t)dof);
        colpiv_qr_solve(S, dof, r_, sol);
        double gamma = 0;
        for (int i = 0; i < dof; ++i) gamma += r_[i] * sol[i];
        gamma = fabs(gamma);
        free(sol); free(S); free(T);
        if (feat_gamma) feat_gamma[f] = gamma;

        if (gamma < ORC_CHI2_95[dof - 1]) {
            memcpy(r + nRowCount, r_, sizeof(double) * (size_t)dof);
            memcpy(Hx + (size_t)nRowCount * n, H_, sizeof(double) * (size_t)dof * n);
            nRowCount += dof;
            nGood++;
        } else {
            if (feat_status) feat_status[f] = 3;
            inf.n_reject_gate++;
        }
        free(tr); free(tHx); free(tHf);
    }
    free(relI); free(relC); free(RI); free(RC); free(Pcc);

    inf.n_good = nGood;
    inf.rows_stacked = nRowCount;
    if (Hstack_out) memcpy(Hstack_out, Hx, sizeof(double) * (size_t)nRowCount * n);
    if (rstack_out) memcpy(rstack_out, r, sizeof(double) * (size_t)nRowCount);

    if (nGood > 2) {
        inf.updated = 1;
        int rk = nRowCount;
        /* Model compression, Updater.cc:474-536 */
        if (nRowCount > n) {
            inf.compressed = 1;
            const int M = nRowCount;
            int Nc = n;
            for (int i = n; i > 0; --i) {
                double s = 0;
                for (int m = 0; m < M; ++m) s += Hx[(size_t)m * n + (i - 1)] * Hx[(size_t)m * n + (i - 1)];
                if (sqrt(s) == 0) Nc--;
                else break;
            }
            for (int c0 = 0; c0 < Nc; ++c0)
                for (int m = M - 1; m > c0; --m) {
                    double c, s;
                    make_givens(Hx[(size_t)(m - 1) * n + c0], Hx[(size_t)m * n + c0], &c, &s);
                    rot_rows(Hx + (size_t)(m - 1) * n + c0, Hx + (size_t)m * n + c0, Nc - c0, c, s);
                    rot_rows(r + (m - 1), r + m, 1, c, s);
                }
            rk = 0;
            int rk_full = 0, stopped = 0;
            for (int i = 0; i < M; ++i) {
                double s = 0;
                for (int j = 0; j < n; ++j) s += Hx[(size_t)i * n + j] * Hx[(size_t)i * n + j];
                if (sqrt(s) < 1e-4) { stopped = 1; continue; }
                if (!stopped) rk++;
                if (g_full_info && rk_full != i) {       /* compact informative rows to the top */
                    memcpy(Hx + (size_t)rk_full * n, Hx + (size_t)i * n, sizeof(double) * (size_t)n);
                    r[rk_full] = r[i];
                }
                rk_full++;
            }
            inf.rank_full = rk_full;
            if (g_full_info) rk = rk_full;
        } else
            inf.rank_full = nRowCount;
        inf.rank = rk;
        /* EKF update, Updater.cc:540-544.  Hn = [0 (rk x 24) | Hx(0:rk,:)] */
        double* PHt = (double*)malloc(sizeof(double) * (size_t)d * (size_t)(rk > 0 ? rk : 1));     /* d x rk */
        double* S = (double*)malloc(sizeof(double) * (size_t)(rk > 0 ? rk : 1) * (size_t)(rk > 0 ? rk : 1));
        double* HP = (double*)malloc(sizeof(double) * (size_t)(rk > 0 ? rk : 1) * (size_t)d);      /* rk x d  = Hn*P */
        for (int i = 0; i < rk; ++i)
            for (int j = 0; j < d; ++j) {
                double a = 0;
                for (int k = 0; k < n; ++k) a += Hx[(size_t)i * n + k] * P[(size_t)j * d + (24 + k)];   /* P(24+k, j) col-major */
                HP[(size_t)i * d + j] = a;
            }
        for (int i = 0; i < rk; ++i)
            for (int j = 0; j < rk; ++j) {
                double a = 0;
                for (int k = 0; k < n; ++k) a += HP[(size_t)i * d + 24 + k] * Hx[(size_t)j * n + k];
                S[i * rk + j] = a;
            }
        for (int i = 0; i < rk; ++i) S[i * rk + i] += sig2;
        for (int i = 0; i < rk; ++i)
            for (int j = i + 1; j < rk; ++j) { double a = .5 * (S[i * rk + j] + S[j * rk + i]); S[i * rk + j] = S[j * rk + i] = a; }
        lu_inverse(S, rk);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < rk; ++j) {
                double a = 0;
                for (int k = 0; k < n; ++k) a += P[(size_t)(24 + k) * d + i] * Hx[(size_t)j * n + k];   /* P(i,24+k) */
                PHt[(size_t)i * rk + j] = a;
            }
        double* K = (double*)malloc(sizeof(double) * (size_t)d * (size_t)(rk > 0 ? rk : 1));
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < rk; ++j) {
                double a = 0;
                for (int k = 0; k < rk; ++k) a += PHt[(size_t)i * rk + k] * S[k * rk + j];
                K[(size_t)i * rk + j] = a;
            }
        double* dx = (double*)malloc(sizeof(double) * (size_t)d);
        for (int i = 0; i < d; ++i) {
            double a = 0;
            for (int k = 0; k < rk; ++k) a += K[(size_t)i * rk + k] * r[k];
            dx[i] = a;
        }
        /* State correction, Updater.cc:546-613 */
        memcpy(x_out, x, sizeof(double) * (size_t)xdim);
        for (int blk = 0; blk < 2 + N; ++blk) {
            int xo, eo;
            if (blk == 0) { xo = 0; eo = 0; }
            else if (blk == 1) { xo = 10; eo = 9; }
            else { xo = 26 + 7 * (blk - 2); eo = 24 + 6 * (blk - 2); }
            double dq[4] = {.5 * dx[eo], .5 * dx[eo + 1], .5 * dx[eo + 2], 0};
            double vn = sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2]);
            if (vn < 1) dq[3] = sqrt(1 - vn * vn);
            else {
                double sc = 1 / sqrt(1 + vn * vn);
                dq[0] *= sc; dq[1] *= sc; dq[2] *= sc; dq[3] = sc;
            }
            orc_quat_mul(dq, x + xo, x_out + xo);
        }
        for (int k = 0; k < 6; ++k) x_out[4 + k] = dx[3 + k] + x[4 + k];
        {
            double* g = x_out + 7;
            double nn = v3_norm(g);
            g[0] /= nn; g[1] /= nn; g[2] /= nn;
        }
        for (int k = 0; k < 12; ++k) x_out[14 + k] = dx[12 + k] + x[14 + k];
        for (int c = 0; c < N; ++c)
            for (int k = 0; k < 3; ++k) x_out[26 + 7 * c + 4 + k] = dx[24 + 6 * c + 3 + k] + x[26 + 7 * c + 4 + k];

        /* Joseph form, Updater.cc:615-619 */
        double* A = (double*)malloc(sizeof(double) * (size_t)d * d);     /* I - K*Hn, row-major */
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                double a = 0;
                if (j >= 24) for (int k = 0; k < rk; ++k) a += K[(size_t)i * rk + k] * Hx[(size_t)k * n + (j - 24)];
                A[(size_t)i * d + j] = ((i == j) ? 1.0 : 0.0) - a;
            }
        double* AP = (double*)malloc(sizeof(double) * (size_t)d * d);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                double a = 0;
                for (int k = 0; k < d; ++k) a += A[(size_t)i * d + k] * P[(size_t)j * d + k];   /* P(k,j) */
                AP[(size_t)i * d + j] = a;
            }
        double* Pn = (double*)malloc(sizeof(double) * (size_t)d * d);
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) {
                double a = 0, b = 0;
                for (int k = 0; k < d; ++k) a += AP[(size_t)i * d + k] * A[(size_t)j * d + k];
                for (int k = 0; k < rk; ++k) b += K[(size_t)i * rk + k] * K[(size_t)j * rk + k];
                Pn[(size_t)i * d + j] = a + sig2 * b;
            }
        for (int i = 0; i < d; ++i)
            for (int j = 0; j < d; ++j) P_out[(size_t)j * d + i] = .5 * (Pn[(size_t)i * d + j] + Pn[(size_t)j * d + i]);
        free(A); free(AP); free(Pn); free(dx); free(K); free(PHt); free(S); free(HP);
    } else {
        /* Updater.cc:621-627 */
        memcpy(x_out, x, sizeof(double) * (size_t)xdim);
        memcpy(P_out, P, sizeof(double) * (size_t)d * d);
    }
    if (info) *info = inf;
    free(r); free(Hx);
}
