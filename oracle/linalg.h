/*
 * oracle/linalg.h -- CPU ORACLE (test infrastructure only).
 * Small dense helpers in float64 used by the Eigen-stage restatements.  Row-major 3x3 unless noted.
 * Numerics.h citations refer to src/util/Numerics.h of the reference.
 */
#ifndef ORC_LINALG_H
#define ORC_LINALG_H
#include <math.h>
#include <string.h>

static inline void m3_mul(const double* A, const double* B, double* C)
{
    double T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(C, T, sizeof T);
}
static inline void m3_T(const double* A, double* B)
{
    double T[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
    memcpy(B, T, sizeof T);
}
static inline void m3_v(const double* A, const double* v, double* o)
{
    double t0 = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
    double t1 = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
    double t2 = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
    o[0] = t0; o[1] = t1; o[2] = t2;
}
static inline void m3T_v(const double* A, const double* v, double* o)
{
    double t0 = A[0] * v[0] + A[3] * v[1] + A[6] * v[2];
    double t1 = A[1] * v[0] + A[4] * v[1] + A[7] * v[2];
    double t2 = A[2] * v[0] + A[5] * v[1] + A[8] * v[2];
    o[0] = t0; o[1] = t1; o[2] = t2;
}
static inline double v3_norm(const double* v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
/* Numerics.h:97-105 */
static inline void skew(const double* w, double* M)
{
    M[0] = 0; M[1] = -w[2]; M[2] = w[1];
    M[3] = w[2]; M[4] = 0; M[5] = -w[0];
    M[6] = -w[1]; M[7] = w[0]; M[8] = 0;
}
static inline void m3_eye(double* M) { memset(M, 0, 9 * sizeof(double)); M[0] = M[4] = M[8] = 1; }

#endif
