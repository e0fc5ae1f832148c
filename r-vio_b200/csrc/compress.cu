// compress.cu -- the reference's model-compression rule (Updater.cc:474-536) on the device.
//
// The reference compresses the stacked system with a bottom-up sweep of Givens rotations on adjacent rows, column by
// column (Updater.cc:494-512), after dropping trailing all-zero columns (:480-489), and then keeps the rows of the
// resulting trapezoid UP TO THE FIRST ROW WHOSE NORM IS BELOW 1e-4 (:515-524).  With noise s^2 I the EKF step depends on
// the kept rows Hn, rn only through  G = Hn^T Hn,  z = Hn^T rn, so the product path keeps its normal-term form
// (k_gram -> solve) and this file only decides WHICH rows the reference keeps and rewrites [G | z] accordingly:
//
//   k_rank_rule   (every updating frame, one CTA)   pivot-free Cholesky of G = R^T R (one thread per row, shared memory),
//                 which yields the rows of the reference's trapezoid as long as the leading columns are independent (R is
//                 unique up to row signs for ANY orthogonal triangularisation) -- the reference's own test (norm < 1e-4)
//                 applies to them.  At a dependent column the reference's row is a unit combination of the rows still
//                 "active" there (features whose column support starts at or before it) with rounding-residue weights:
//                 below 1e-4 for certain when the active rows are exhausted (trace of the active Schur complement < 1e-8:
//                 the cut is there, every class of features that would only start later is discarded and [G | z] are
//                 rebuilt from the kept rows), orders of magnitude above it otherwise.  Everything in between is left to
//                 k_givens_ref.  Also hands the kept rows R and y (R^T y = z) to the EKF step (R-form).
//   k_givens_ref  (only then, one CTA)   the reference's Givens sweep itself, rotation for rotation, scheduled as a
//                 wavefront: rotation (column n, rows m-1,m) runs at step t = (M-1-m) + 2n; all rotations of a step touch
//                 disjoint row pairs, M + N' - 2 dependent steps instead of ~M N'.  The rows live in a circular
//                 shared-memory window of 2N'+8 rows that slides up the stacked matrix (each row of H is read from HBM
//                 exactly once, by cp.async, a few steps ahead).  Special cases of Eigen's makeGivens (q == 0 -> identity,
//                 p == 0 -> row swap) are kept exactly: they are what moves the exact zeros around and make the reference's
//                 outcome deterministic.  Then the first-small-row cut, and [G | z] := kept rows.
//   k_solve_small_R / k_chol_S + k_trsm   the EKF step itself on the kept rows (Updater.cc:540-619), see below.
//
// Both kernels read the mode from device memory (0 = reference rule, 1 = full information: keep [G | z] of all rows), so
// captured frame graphs stay valid when the mode is switched.
#include "common.cuh"
#include "compress_kernels.cuh"

namespace rvio {

namespace {

constexpr int kGVThreads = 512;
constexpr int kGVPrefetch = 6;        // rows in flight ahead of the wavefront
constexpr int kGVMaxFeat = 1024;      // accepted-feature list kept in shared memory

__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Single-CTA factorisations of this file (Cholesky of G for the rank rule, of S for the EKF step) all use ONE THREAD PER ROW
// with the matrix in shared memory: per elimination step a thread walks only its own remaining row, and warps whose rows
// are finished just meet the barrier.  A register-tiled, block-wide variant (4 x 8 tiles, 576 threads, one published column
// per step) was built and measured first: 0.9 - 1.1 us per step regardless of n (60 us for N' = 30, 194 us for n = 180) --
// every one of its 18 warps walks the whole ~700-instruction step body, so it is instruction-issue bound, not latency
// bound.  The per-row form issues only the instructions of live rows.
// ------------------------------------------------------------------------------------------------
#ifdef RVIO_B200_PHASE_CLOCKS
// (profiling build only: make PHASES=1) per-phase SM clock stamps of the single-CTA kernels, read back by tools/prof_update.py
__device__ long long g_phase_clk[64];
#define PHASE_CLK(k) do { if (threadIdx.x == 0) g_phase_clk[k] = clock64(); } while (0)
#else
#define PHASE_CLK(k) do { } while (0)
#endif
constexpr int kSFThreads = 576;           // rank-rule CTA (post-pass loops are sized for it)
constexpr int kSFMaxRows = 188;           // n + 1 <= 188 (31 clones)

// ------------------------------------------------------------------------------------------------
// k_rank_rule
// counters (red + n*n + n): [0] accepted features, [1] stacked rows, ..., [6] rows kept (rank), [7] flags (RVIO_RANK_*):
//   1 = the cut discarded information, 2 = decided by the Givens sweep, 4 = undecided (feature-sharded call: the stacked
//   rows are distributed; full information is used), 8 = [G | z] rewritten from the kept rows, 16 = dependent columns inside
//   the active set (every row is kept; the reference's nRank additionally counts its linearly dependent rows)
// cls (red + n*n + n + 8): information ||H_f||_F^2 summed per column-support class (first non-zero column of the feature).
//
// What decides the reference's outcome.  Its sweep leaves in row j the normalised combination of all rows that are
// "active" at column j (rows of features whose support starts at or before j).  While the leading columns are independent
// that row is unique (= row j of the pivot-free Cholesky factor of G) and the reference's own test applies to it.  At a
// dependent column the combination is taken with weights proportional to rounding residue: it is below 1e-4 for certain
// when the active rows carry nothing any more (trace of the active Schur complement < 1e-8) -- the cut is there, and every
// feature that would only start later is discarded -- and above it (by orders of magnitude: a random unit combination of
// rows holding >= 1e-3 of information) otherwise.  The factorisation therefore runs over the TOTAL G once, skipping
// dependent columns, and at every column where a new class of features starts it compares the information left in the
// active part with what is still to come:
//     active part exhausted (< 1e-8), something dependent behind us  ->  cut here, kept = good pivot rows so far
//     exactly one dependent column behind us, active part alive (>= 1e-3)  ->  the reference is still going: continue
//     anything in between  ->  k_givens_ref replays the reference's sweep
// With Q.emit_R the kept rows (scaled, zero padded to n x n) and y are also written out: the large-window EKF step works on
// them (R-form: S = R Pcc R^T + s^2 I).
// ------------------------------------------------------------------------------------------------
// PACKED = false: rows of the factor in a rectangular shared-memory array (n <= 96), which the post-pass reads in place;
// PACKED = true: rows packed (row i holds columns i .. N' and the right-hand side), written out to Q.L afterwards.
template <bool PACKED>
__global__ void __launch_bounds__(kSFThreads, 1) k_rank_rule(RankRuleParams Q)
{
    extern __shared__ __align__(16) double rsm[];            // the factor itself, one thread per row
    __shared__ double s_pv[kSFMaxRows], s_gd[kSFMaxRows], s_late[kSFMaxRows + 2], s_nr2[kSFMaxRows];
    __shared__ int s_np, s_k, s_smin, s_ncls;
    __shared__ double s_tau;
    const int n = Q.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* cnt = Q.red + (size_t)n * n + n;
    const double* cls = cnt + 8;
    const double* G = Q.red;
    const int rows = (int)cnt[1];
    const bool updating = cnt[0] > 2.0;                                            // Updater.cc:460
    const bool rule = updating && (*Q.rule_dev == 0) && rows > n;                  // Updater.cc:474
    if (tid == 0) {
        Q.rr[0] = 0; Q.rr[1] = rows; Q.rr[2] = n; Q.rr[3] = 0;
        s_np = 0; s_smin = n + 1; s_ncls = 0; s_k = n + 1;
        cnt[6] = updating ? (double)rows : 0.0; cnt[7] = 0.0;
    }
    if (!updating || (!rule && !Q.emit_R)) return;
    __syncthreads();
    // trailing all-zero columns (Updater.cc:480-489): column norm == 0  <=>  G(j,j) == 0
    for (int j = tid; j < n; j += kSFThreads) {
        const double g = G[(size_t)j * n + j];
        s_gd[j] = g;
        if (g != 0.0) atomicMax(&s_np, j + 1);
    }
    for (int c = tid; c <= n; c += kSFThreads)
        if (cls[c] > 0.0) { atomicMin(&s_smin, c); atomicAdd(&s_ncls, 1); }
    __syncthreads();
    const int Np = s_np;
    PHASE_CLK(16);
    for (int c = tid; c <= n; c += kSFThreads) s_nr2[c < kSFMaxRows ? c : kSFMaxRows - 1] = cls[c];      // staged: the suffix sum below must not walk global memory
    __syncthreads();
    if (tid == 0) {                                               // late[j] = information of the classes starting at column >= j
        double acc = 0;
        for (int c = n; c > Np; --c) acc += s_nr2[c];
        for (int j = Np; j >= 0; --j) { acc += s_nr2[j]; s_late[j] = acc; }
        s_late[Np + 1] = 0.0;
    }
    __syncthreads();
    // no feature starts at column 0 (no '2' feature, no full-length '1'): the first rows of the reference's trapezoid are
    // raw rows in list order -- with several classes only the sweep knows what a later cut would keep
    const bool raw_top = rule && s_smin > 0 && s_ncls > 1;
    const bool boundaries = rule && s_smin == 0 && !raw_top;

    const int ldl = PACKED ? (n + 1) : ((Np + 2) & ~1);      // even: thread i walks row i from column i, stride ldl + 1 doubles (odd) between lanes
    double* L = PACKED ? Q.L : rsm;                               // rows of R (= columns of the lower factor), y at index Np
    int q = 0, first_dep = Np;                                    // (thread-uniform copies)
    int mode = 1, kcut = 0;
    bool undecided = false;
    auto after_pivot = [&](int j, bool dep) { if (!dep) q++; else if (first_dep == Np) first_dep = j; };
    // class-boundary test (see above); tau = trace of the current Schur complement - information that has not started yet
    auto boundary_decision = [&](int j, double tau) -> bool {
        const int dd = j - q;
        if (dd >= 1) {
            if (tau < 1e-8) { mode = 2; kcut = j; return false; }                // exhausted: the reference cuts, later classes are discarded
            if (tau < 1e-3 || dd >= 2) {                                        // only the reference's own sweep can tell
                if (Q.world == 1) { mode = 3; return false; }
                undecided = true;                                               // feature-sharded: keep everything, say so
            }
        }
        return true;
    };
    int jstop = Np;
    {
        // ---- ONE THREAD PER ROW of the upper factor in shared memory.  Per step a thread walks its own row
        //      (U(i,k) -= U(j,i) / p * U(j,k), k = i .. N', plus the right-hand side): the cost of a step is the length of the
        //      remaining rows, and warps whose rows are finished only meet the barrier.  (A register-tiled block-wide scheme
        //      was measured 5x slower here: every warp walks the whole step body, and instruction issue, not latency, bounds it.)
        double* U = rsm;
        auto row = [&](int i) -> double* { return PACKED ? (U + (size_t)i * (Np + 1) - (size_t)i * (i - 1) / 2 - i) : (U + (size_t)i * ldl); };   // row(i)[k], k >= i
        for (int o = tid; o < Np * (Np + 1); o += kSFThreads) {
            const int i = o / (Np + 1), k = o - i * (Np + 1);
            if (k >= i) row(i)[k] = (k < Np) ? G[(size_t)i * n + k] : Q.red[(size_t)n * n + i];
        }
        __syncthreads();
        PHASE_CLK(17);
        for (int j = 0; j < Np; ++j) {
            if (boundaries && j > 0 && s_late[j] > s_late[j + 1]) {
                if (warp == 0) {
                    double t = 0;
                    for (int k = j + lane; k < Np; k += 32) t += row(k)[k];
                    t = warp_sum_d(t);
                    if (lane == 0) s_tau = t - s_late[j];
                }
                __syncthreads();
                const double tau = s_tau;
                __syncthreads();
                if (!boundary_decision(j, tau)) { jstop = j; break; }
            }
            const double* rj = row(j);
            const double pj = rj[j];
            const bool dep = !(pj >= fmax(1e-12, 1e-12 * s_gd[j]));
            if (tid == 0) s_pv[j] = dep ? -1.0 : pj;
            after_pivot(j, dep);
            if (dep) continue;                                     // nothing changes: no barrier needed
            const int i = j + 1 + tid;
            if (i < Np) {
                const double f = -rj[i] / pj;
                double* __restrict__ ri = row(i);
                const double* __restrict__ rjj = rj;
                int k = i;
                for (; k + 8 <= Np + 1; k += 8) {                  // 8 loads in flight, then 8 stores (the rows do not alias)
                    double a[8], b[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { a[u] = rjj[k + u]; b[u] = ri[k + u]; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) ri[k + u] = fma(f, a[u], b[u]);
                }
                for (; k <= Np; ++k) ri[k] = fma(f, rjj[k], ri[k]);
            }
            __syncthreads();
        }
        __syncthreads();
        PHASE_CLK(18);
        // rows of R: U(j, j..) / sqrt(p_j); the right-hand side becomes y
        for (int j = warp; j < jstop; j += kSFThreads / 32) {
            if (s_pv[j] < 0.0) continue;
            const double rs = rsqrt(s_pv[j]);
            double* rj = row(j);
            for (int k = j + lane; k <= Np; k += 32) {
                const double v = rj[k] * rs;
                if (PACKED) L[(size_t)j * ldl + k] = v; else rj[k] = v;
            }
        }
    }
    __syncthreads();
    PHASE_CLK(19);
    if (raw_top) { if (Q.world == 1) mode = 3; else undecided = true; }
    const int jend = (mode == 2) ? kcut : jstop;                  // columns whose rows may be kept
    // the reference's own test on the rows that are unique (before the first dependent column): norm < 1e-4 (Updater.cc:519)
    const int ulim = rule ? min(first_dep, jend) : 0;
    for (int j = warp; j < jend; j += kSFThreads / 32) {
        if (s_pv[j] < 0.0) { if (lane == 0) s_nr2[j] = 0.0; continue; }
        const double* cj = L + (size_t)j * ldl;
        double sq = 0;
        for (int i = j + lane; i < Np; i += 32) sq += cj[i] * cj[i];
        sq = warp_sum_d(sq);
        if (lane == 0) { s_nr2[j] = sq; if (j < ulim && sq < 1e-8) atomicMin(&s_k, j); }
    }
    __syncthreads();
    PHASE_CLK(20);
    int kept_rows = q;
    bool rebuild = false, discards = false, generic = false;
    int klim = jend;             // columns with index < klim (and a good pivot) are kept
    if (rule && s_k <= n) {      // an early small row: the reference stops there, whatever comes later
        klim = s_k; kept_rows = s_k; rebuild = true; mode = 2;
        if (warp == 0) {         // information dropped: everything except the kept rows
            double tr = 0, kept = 0;
            for (int j = lane; j < Np; j += 32) tr += s_gd[j];
            for (int j = lane; j < klim; j += 32) kept += s_nr2[j];
            tr = warp_sum_d(tr); kept = warp_sum_d(kept);
            if (lane == 0) s_tau = tr - kept;
        }
        __syncthreads();
        discards = s_tau >= 1e-8;
    } else if (mode == 2) {      // cut at a class boundary: the later classes are discarded
        rebuild = true; discards = s_late[kcut] >= 1e-8;
    } else if (mode == 1) {      // ran to the end: every row is kept
        generic = rule && first_dep < Np && q > first_dep;        // dependent columns in the middle of the active set
    }
    if (undecided && mode == 1) { mode = 4; generic = false; }
    if (!rule) { mode = 0; rebuild = false; }
    if (tid == 0) {
        Q.rr[0] = mode; Q.rr[1] = kept_rows; Q.rr[2] = Np; Q.rr[3] = (mode == 3) ? 1 : 0; Q.rr[5] = klim;
        cnt[6] = rule ? (double)kept_rows : (double)rows;
        cnt[7] = (rebuild ? 8.0 : 0.0) + (discards ? 1.0 : 0.0) + (mode == 4 ? 4.0 : 0.0) + (generic ? 16.0 : 0.0);
    }
    if (Q.emit_R && mode != 3) {
        // kept rows of R (= columns of L with a good pivot below klim), zero padded to n x n, and y
        const int kl = klim;
        for (int o = tid; o < n * n; o += kSFThreads) {
            const int j = o / n, i = o - j * n;
            double v = 0;
            if (j < kl && j < Np && s_pv[j] >= 0.0 && i >= j && i < Np) v = L[(size_t)j * ldl + i];
            Q.Rc[o] = v;
        }
        for (int j = tid; j < n; j += kSFThreads) Q.yc[j] = (j < kl && j < Np && s_pv[j] >= 0.0) ? L[(size_t)j * ldl + Np] : 0.0;
    }
    PHASE_CLK(21);
    if (!rebuild) return;
    // G' = sum_{kept j} L_j L_j^T ,  z' = sum_{kept j} L_j y_j   (bitwise symmetric: products commute)
    double* Gw = Q.red; double* zw = Q.red + (size_t)n * n;
    for (int o = tid; o < Np * Np; o += kSFThreads) {
        const int r = o / Np, c = o - r * Np;
        const int lim = min(klim, min(r, c) + 1);
        double acc = 0;
        for (int j = 0; j < lim; ++j)
            if (s_pv[j] >= 0.0) acc = fma(L[(size_t)j * ldl + r], L[(size_t)j * ldl + c], acc);
        Gw[(size_t)r * n + c] = acc;
    }
    for (int r = tid; r < Np; r += kSFThreads) {
        const int lim = min(klim, r + 1);
        double acc = 0;
        for (int j = 0; j < lim; ++j)
            if (s_pv[j] >= 0.0) acc = fma(L[(size_t)j * ldl + r], L[(size_t)j * ldl + Np], acc);
        zw[r] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Large-window EKF step, serial part:  S = L L^T (k_chol_S, one CTA, one thread per row, packed triangle in shared memory), then  Y = L^-1 [W | y]
// (k_trsm, one CTA per 8 right-hand-side columns, L in shared memory).
// ------------------------------------------------------------------------------------------------
constexpr int kCholThreads = 192;
__global__ void __launch_bounds__(kCholThreads, 1) k_chol_S(const double* S, int m, double* Lp, double* invd, int* bad, const double* gate)
{
    extern __shared__ __align__(16) double csm[];                // lower triangle, row i at i (i + 1) / 2
    __shared__ double s_rs[kSFMaxRows + 8];
    if (gate && !(gate[0] > 2.0)) return;
    const int tid = threadIdx.x;
    double* A = csm;
    for (int o = tid; o < m * m; o += kCholThreads) {
        const int r = o / m, c = o - r * m;
        if (c <= r) A[r * (r + 1) / 2 + c] = S[(size_t)r * m + c];
    }
    __syncthreads();
    // right-looking Cholesky, one thread per row: A(i,k) -= A(i,j) / p * A(k,j), k = j+1 .. i; columns unscaled until the end
    const int i = tid;
    double* __restrict__ ri = A + i * (i + 1) / 2;
    for (int j = 0; j < m; ++j) {
        const double p = A[j * (j + 1) / 2 + j];
        if (!(p > 0.0)) { if (tid == 0) *bad = 1; }
        if (tid == 0) s_rs[j] = (p > 0.0) ? rsqrt(p) : 0.0;
        if (i > j && i < m && p > 0.0) {
            const double f = -ri[j] / p;
            int offk = (j + 1) * (j + 2) / 2 + j;                  // A(k, j), k = j + 1
            int k = j + 1;
            for (; k + 8 <= i + 1; k += 8) {
                double a[8], b[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) { a[u] = A[offk]; offk += k + u + 1; b[u] = ri[k + u]; }
#pragma unroll
                for (int u = 0; u < 8; ++u) ri[k + u] = fma(f, a[u], b[u]);
            }
            for (; k <= i; ++k) { ri[k] = fma(f, A[offk], ri[k]); offk += k + 1; }
        }
        __syncthreads();
    }
    // the factor, column packed (what k_trsm reads): column j at j m - j (j-1) / 2, element i at + (i - j)
    for (int o = tid; o < m * m; o += kCholThreads) {
        const int r = o / m, c = o - r * m;
        if (c <= r) Lp[((size_t)c * m - (size_t)c * (c - 1) / 2) + (r - c)] = A[r * (r + 1) / 2 + c] * s_rs[c];
    }
    for (int j = tid; j < m; j += kCholThreads) invd[j] = s_rs[j];
}

// B: m x nb row-major right-hand sides (leading dimension ldb); solves L Y = B in place.  Lp: packed lower factor (column j
// at j m - j (j-1) / 2), invd[j] = 1 / L(j,j).  One CTA per kTrsmCols columns, one thread per row, the factor in shared memory.
constexpr int kTrsmCols = 8;
__global__ void __launch_bounds__(192) k_trsm(const double* Lp, const double* invd, int m, double* B, int ldb, int nb, const double* gate)
{
    extern __shared__ __align__(16) double sL[];                 // packed lower triangle, then invd[m]
    __shared__ double s_y[2][kTrsmCols];
    if (gate && !(gate[0] > 2.0)) return;
    const int tid = threadIdx.x;
    const int c0 = blockIdx.x * kTrsmCols;
    const int np = m * (m + 1) / 2;
    double* sI = sL + np;
    for (int e = tid; e < np; e += 192) sL[e] = Lp[e];           // linear copy: many loads in flight
    for (int e = tid; e < m; e += 192) sI[e] = invd[e];
    double b[kTrsmCols];
#pragma unroll
    for (int c = 0; c < kTrsmCols; ++c) b[c] = (tid < m && c0 + c < nb) ? B[(size_t)tid * ldb + c0 + c] : 0.0;
    __syncthreads();
    int buf = 0;
    for (int j = 0; j < m; ++j) {
        const int off = j * m - j * (j - 1) / 2 - j;             // column j, element i at off + i
        if (tid == j) {
            const double inv = sI[j];
#pragma unroll
            for (int c = 0; c < kTrsmCols; ++c) { b[c] *= inv; s_y[buf][c] = b[c]; }
        }
        __syncthreads();
        if (tid > j && tid < m) {
            const double l = sL[off + tid];
#pragma unroll
            for (int c = 0; c < kTrsmCols; ++c) b[c] = fma(-l, s_y[buf][c], b[c]);
        }
        buf ^= 1;
    }
#pragma unroll
    for (int c = 0; c < kTrsmCols; ++c) if (tid < m && c0 + c < nb) B[(size_t)tid * ldb + c0 + c] = b[c];
}

// ------------------------------------------------------------------------------------------------
// k_solve_small_R -- the whole EKF step of a small window (n + d + 1 <= 188, i.e. N <= 13) in ONE CTA, R-form
// (Updater.cc:540-619 with Hn = R, the kept rows handed over by the rank rule, zero rows where a column was dropped):
//     W = R P[c,:]            (n x d, shared memory, 2 x 4 register tiles)
//     S = W[:,c] R^T + s^2 I  (n x n, SPD)
//     [ S ; W^T ; y^T ]  ->  register-resident Cholesky with the d + 1 right-hand sides as extra rows: the extra rows of the
//                            factor are Y^T = (L^-1 [W | y])^T, i.e. factorisation and triangular solves in the same sweep
//     dx = Y^T y~ ,  P+ = sym(P) - Y^T Y ,  state correction (quaternions multiplicative, Updater.cc:546-613)
// Replaces k_wgemm + k_gj_block + k_pout_finalize (three launches, a pivoted Gauss-Jordan on the non-symmetric G Pcc + s^2 I).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sr_quat_mul(const double* q1, const double* q2, double* out)      // Numerics.h:30-63
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
__device__ __forceinline__ void sr_apply_dq(const double* dth, const double* q, double* out)      // Updater.cc:549-566
{
    double dq[4] = {.5 * dth[0], .5 * dth[1], .5 * dth[2], 0};
    const double vn = sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2]);
    if (vn < 1) dq[3] = sqrt(1 - vn * vn);
    else {
        const double sc = 1 / sqrt(1 + vn * vn);
        dq[0] *= sc; dq[1] *= sc; dq[2] *= sc; dq[3] = sc;
    }
    sr_quat_mul(dq, q, out);
}

constexpr int kSRThreads = 256;
__global__ void __launch_bounds__(kSRThreads, 1) k_solve_small_R(SolveSmallRParams Q)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ double s_rs[kSFMaxRows];
    __shared__ double s_dx[kSFMaxRows];
    const int tid = threadIdx.x;
    const int N = Q.N, n = 6 * N, d = Q.d;
    if (!(Q.gate[0] > 2.0)) {                                      // Updater.cc:621-627: too few features, posterior = prior
        for (int o = tid; o < d * d; o += kSRThreads) Q.P_out[o] = Q.P[o];
        for (int o = tid; o < Q.xdim; o += kSRThreads) Q.x_out[o] = Q.x[o];
        return;
    }
    // A: (n + d + 1) x lda.  Rows 0..n-1: lower triangle of S, then of its Cholesky factor.  Rows n + c: column c of [W | y],
    // after the factorisation column c of Y = L^-1 [W | y] (the right-hand sides ride along as extra rows).
    const int rows_total = n + d + 1;
    const int lda = n + 1, ldp = d + 1;                            // both odd: rows of consecutive threads fall into different banks
    double* A = sm;
    double* sR = A + (size_t)rows_total * lda;                     // n x lda
    double* sP = sR + (size_t)n * lda;                             // n x ldp: P[c,:]
    PHASE_CLK(0);
    for (int o = tid; o < n * n; o += kSRThreads) { const int r = o / n, c = o - r * n; sR[r * lda + c] = Q.Rc[o]; }
    for (int o = tid; o < n * d; o += kSRThreads) { const int j = o / n, k = o - j * n; sP[k * ldp + j] = Q.P[(size_t)j * d + 24 + k]; }   // P(24+k, j)
    __syncthreads();
    PHASE_CLK(1);
    // ---- W = R P[c,:]  (R upper triangular), written transposed into the extra rows of A
    {
        const int tr_n = (n + 1) / 2, tc_n = (d + 3) / 4;
        for (int t = tid; t < tr_n * tc_n; t += kSRThreads) {
            const int r0 = 2 * (t / tc_n), j0 = 4 * (t % tc_n);
            double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
            const bool r1ok = r0 + 1 < n;
            for (int k = r0; k < n; ++k) {
                const double a0 = sR[r0 * lda + k], a1 = r1ok ? sR[(r0 + 1) * lda + k] : 0.0;
                const double* pk = sP + k * ldp + j0;
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const double b = (j0 + y < d) ? pk[y] : 0.0;
                    acc[0][y] = fma(a0, b, acc[0][y]);
                    acc[1][y] = fma(a1, b, acc[1][y]);
                }
            }
#pragma unroll
            for (int y = 0; y < 4; ++y)
                if (j0 + y < d) { A[(n + j0 + y) * lda + r0] = acc[0][y]; if (r1ok) A[(n + j0 + y) * lda + r0 + 1] = acc[1][y]; }
        }
        for (int r = tid; r < n; r += kSRThreads) A[(n + d) * lda + r] = Q.yc[r];
    }
    __syncthreads();
    PHASE_CLK(2);
    // ---- S = W[:, 24:] R^T + s^2 I  (lower triangle; R[c][k] = 0 for k < c) into rows 0..n-1 of A
    {
        const int tn = (n + 1) / 2, tcn = (n + 3) / 4;
        for (int t = tid; t < tn * tcn; t += kSRThreads) {
            const int r0 = 2 * (t / tcn), c0 = 4 * (t % tcn);
            if (c0 > r0 + 1) continue;                              // tile entirely above the diagonal
            double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
            const bool r1ok = r0 + 1 < n;
            for (int k = c0; k < n; ++k) {
                const double* wk = A + (n + 24 + k) * lda;          // W(:, 24 + k)
                const double w0 = wk[r0], w1 = r1ok ? wk[r0 + 1] : 0.0;
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const double b = (c0 + y < n) ? sR[(c0 + y) * lda + k] : 0.0;
                    acc[0][y] = fma(w0, b, acc[0][y]);
                    acc[1][y] = fma(w1, b, acc[1][y]);
                }
            }
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int c = c0 + y;
                if (c < n) {
                    if (c <= r0) A[r0 * lda + c] = acc[0][y] + (c == r0 ? Q.sig2 : 0.0);
                    if (r1ok && c <= r0 + 1) A[(r0 + 1) * lda + c] = acc[1][y] + (c == r0 + 1 ? Q.sig2 : 0.0);
                }
            }
        }
    }
    __syncthreads();
    PHASE_CLK(3);
    // ---- right-looking Cholesky, ONE THREAD PER ROW (rows of S and the extra rows alike): at step j row i does
    //      A(i,k) -= A(i,j) / p * A(k,j) for k = j+1 .. min(i, n-1).  Columns stay unscaled until the end (every thread reads
    //      column j while its owners would be rescaling it).  One barrier per step; the work of a step is the remaining row.
    {
        const int i = tid;
        for (int j = 0; j < n; ++j) {
            const double p = A[j * lda + j];
            if (!(p > 0.0)) { if (tid == 0) *Q.bad = 1; }
            if (tid == 0) s_rs[j] = (p > 0.0) ? rsqrt(p) : 0.0;
            if (i > j && i < rows_total && p > 0.0) {
                double* __restrict__ ri = A + i * lda;
                const double f = -ri[j] / p;
                const int kmax = min(i, n - 1);
                const double* __restrict__ cj = A + j;             // column j: A(k, j) = cj[k * lda] (never written during step j)
                int k = j + 1;
                for (; k + 8 <= kmax + 1; k += 8) {
                    double a[8], b[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { a[u] = cj[(k + u) * lda]; b[u] = ri[k + u]; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) ri[k + u] = fma(f, a[u], b[u]);
                }
                for (; k <= kmax; ++k) ri[k] = fma(f, cj[k * lda], ri[k]);
            }
            __syncthreads();
        }
        PHASE_CLK(4);
        for (int o = tid; o < rows_total * n; o += kSRThreads) {   // scale the columns: L(i,j) = A(i,j) / sqrt(p_j)
            const int r = o / n, c = o - r * n;
            if (c <= r) A[r * lda + c] *= s_rs[c];
        }
    }
    __syncthreads();
    PHASE_CLK(5);
    // ---- dx = Y^T y~ ;  Y(j, c) = A(n + c, j)
    for (int c = tid; c < d; c += kSRThreads) {
        const double* yc_ = A + (n + c) * lda; const double* yd = A + (n + d) * lda;
        double acc = 0;
        for (int k = 0; k < n; ++k) acc = fma(yc_[k], yd[k], acc);
        s_dx[c] = acc;
    }
    // ---- P+ = sym(P) - Y^T Y
    {
        const int tn = (d + 1) / 2, tcn = (d + 3) / 4;
        for (int t = tid; t < tn * tcn; t += kSRThreads) {
            const int i0 = 2 * (t / tcn), j0 = 4 * (t % tcn);
            double acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
            const bool i1ok = i0 + 1 < d;
            const double* u0p = A + (n + i0) * lda; const double* u1p = i1ok ? u0p + lda : u0p;
            const double* vp[4];
#pragma unroll
            for (int y = 0; y < 4; ++y) vp[y] = A + (n + min(j0 + y, d - 1)) * lda;
            for (int k = 0; k < n; ++k) {
                const double u0 = u0p[k], u1 = u1p[k];
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const double b = vp[y][k];
                    acc[0][y] = fma(u0, b, acc[0][y]);
                    acc[1][y] = fma(u1, b, acc[1][y]);
                }
            }
#pragma unroll
            for (int x = 0; x < 2; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int i = i0 + x, j = j0 + y;
                    if (i < d && j < d) Q.P_out[(size_t)j * d + i] = .5 * (Q.P[(size_t)j * d + i] + Q.P[(size_t)i * d + j]) - acc[x][y];
                }
        }
    }
    __syncthreads();
    PHASE_CLK(6);
    // ---- state correction (Updater.cc:546-613)
    const double* x = Q.x; double* xo = Q.x_out; const double* dx = s_dx;
    for (int bq = tid; bq < 2 + N; bq += kSRThreads) {
        int xq, eq;
        if (bq == 0) { xq = 0; eq = 0; }
        else if (bq == 1) { xq = 10; eq = 9; }
        else { xq = 26 + 7 * (bq - 2); eq = 24 + 6 * (bq - 2); }
        sr_apply_dq(dx + eq, x + xq, xo + xq);
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

// ------------------------------------------------------------------------------------------------
// k_givens_ref
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool SMEM>
__global__ void __launch_bounds__(kGVThreads) k_givens_ref(GivensRefParams Q)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ short s_accf[kGVMaxFeat], s_accd[kGVMaxFeat];
    __shared__ int s_nacc, s_k, s_full;
    __shared__ double s_tr[2];
    const int n = Q.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (Q.rr[3] == 0) return;                                     // k_rank_rule settled it
    double* cnt = Q.red + (size_t)n * n + n;
    const int n_feat = Q.n_feat_dev ? *Q.n_feat_dev : Q.n_feat;
    const int Np = Q.rr[2];
    const int LD = Np + 1;                                        // column Np carries r
    const int WIN = 2 * Np + kGVPrefetch + 2;
    double* W = SMEM ? sm : Q.win;

    // accepted features in list order (= the reference's stacking order, Updater.cc:424-428)
    if (warp == 0) {
        int base = 0;
        for (int f0 = 0; f0 < n_feat; f0 += 32) {
            const int f = f0 + lane;
            const int dof = (f < n_feat) ? Q.f_dof[f] : 0;
            const unsigned m = __ballot_sync(0xffffffffu, dof > 0);
            const int pos = base + __popc(m & ((1u << lane) - 1u));
            if (dof > 0 && pos < kGVMaxFeat) { s_accf[pos] = (short)f; s_accd[pos] = (short)dof; }
            base += __popc(m);
        }
        if (lane == 0) s_nacc = base;
    }
    __syncthreads();
    const int nacc = s_nacc;
    const int M = (int)cnt[1];
    if (nacc > kGVMaxFeat || M <= Np) {                            // cannot happen with the capacities checked at create
        if (tid == 0) cnt[7] = 4.0;
        return;
    }

    // row loader (last warp): cursor over the stacked rows from the bottom
    int cur_k = nacc - 1, cur_a = (nacc > 0) ? s_accd[nacc - 1] - 1 : 0, cur_row = M - 1;
    auto issue_row = [&]() {
        if (cur_row >= 0) {
            const int f = s_accf[cur_k];
            const double* src = Q.Hblk + ((size_t)f * Q.blk_rows + cur_a) * n;
            double* dst = W + (size_t)(cur_row % WIN) * LD;
            if (SMEM) {
                for (int j = lane; j < Np; j += 32) cp_async8(dst + j, src + j);
                if (lane == 0) cp_async8(dst + Np, Q.rblk + (size_t)f * Q.blk_rows + cur_a);
            } else {
                for (int j = lane; j < Np; j += 32) dst[j] = src[j];
                if (lane == 0) dst[Np] = Q.rblk[(size_t)f * Q.blk_rows + cur_a];
            }
            cur_row--;
            if (--cur_a < 0) { cur_k--; cur_a = (cur_k >= 0) ? s_accd[cur_k] - 1 : 0; }
        }
        if (SMEM) cp_async_commit();
    };
    const bool loader = warp == kGVThreads / 32 - 1;
    if (loader) {
        for (int i = 0; i <= kGVPrefetch; ++i) issue_row();        // rows M-1 .. M-1-PF
        if (SMEM) cp_async_wait<0>();
    }
    __syncthreads();

    const int slot = tid >> 3, l = tid & 7;                          // 8 lanes per rotation, 64 rotations per pass
    constexpr int NSLOT = kGVThreads / 8;
    const unsigned hmask = 0xFFu << (8 * (slot & 3));
    const int T = M + Np - 2;
    for (int t = 0; t < T; ++t) {
        if (loader) issue_row();                                   // row M-2-(t+PF), needed at step t+PF
        const int n_lo = max(0, t - M + 2), n_hi = min(Np - 1, t >> 1);
        for (int nn = n_lo + slot; nn <= n_hi; nn += NSLOT) {
            const int m = M - 1 - t + 2 * nn;                      // rotation on rows (m-1, m), columns nn..Np
            const int sa = (m - 1) % WIN;
            const int sb = (sa + 1 == WIN) ? 0 : sa + 1;
            double* ra = W + (size_t)sa * LD;
            double* rb = W + (size_t)sb * LD;
            const double p = ra[nn], q = rb[nn];
            double c, s;
            // Eigen JacobiRotation::makeGivens(p, q): both general branches reduce to c = p / h, s = -q / h, h = hypot
            if (q == 0.0) { c = (p < 0.0) ? -1.0 : 1.0; s = 0.0; }
            else if (p == 0.0) { c = 0.0; s = (q < 0.0) ? 1.0 : -1.0; }
            else {
                double pp = p, qq = q, h2 = p * p + q * q;
                if (h2 < 1e-200) { pp *= 0x1p300; qq *= 0x1p300; h2 = pp * pp + qq * qq; }      // no underflow in the squares
                const double rh = rsqrt(h2);
                c = pp * rh; s = -qq * rh;
            }
            __syncwarp(hmask);                                      // p, q read by all 8 lanes before column nn is rewritten
            if (!(c == 1.0 && s == 0.0)) {
                for (int j = nn + l; j <= Np; j += 8) {
                    const double x = ra[j], y = rb[j];
                    ra[j] = c * x - s * y;                          // applyOnTheLeft(0, 1, G.adjoint())
                    rb[j] = s * x + c * y;
                }
            }
        }
        if (SMEM && loader) cp_async_wait<kGVPrefetch - 1>();
        __syncthreads();
    }

    // first-small-row cut (Updater.cc:515-524) on rows 0..Np-1 (rows >= Np are eliminated: norm ~ 1e-16)
    double* nrm = SMEM ? (sm + (size_t)WIN * LD) : sm;             // Np doubles
    if (tid == 0) { s_k = Np; s_full = 0; }
    for (int i = warp; i < Np; i += kGVThreads / 32) {
        const double* ri = W + (size_t)(i % WIN) * LD;
        double sq = 0;
        for (int j = i + lane; j < Np; j += 32) sq += ri[j] * ri[j];
        sq = warp_sum_d(sq);
        if (lane == 0) nrm[i] = sq;
    }
    __syncthreads();
    for (int i = tid; i < Np; i += kGVThreads)
        if (sqrt(nrm[i]) < 1e-4) atomicMin(&s_k, i); else atomicAdd(&s_full, 1);
    __syncthreads();
    const int k = s_k;
    if (warp == 0) {
        double tr = 0, kept = 0;
        for (int j = lane; j < Np; j += 32) tr += Q.red[(size_t)j * n + j];
        for (int j = lane; j < k; j += 32) kept += nrm[j];
        tr = warp_sum_d(tr); kept = warp_sum_d(kept);
        if (lane == 0) { s_tr[0] = tr; s_tr[1] = kept; }
    }
    __syncthreads();
    // [G | z] := kept rows
    double* Gw = Q.red; double* zw = Q.red + (size_t)n * n;
    for (int o = tid; o < Np * Np; o += kGVThreads) {
        const int a = o / Np, b = o - a * Np;
        const int lim = min(k, min(a, b) + 1);
        double acc = 0;
        for (int i = 0; i < lim; ++i) {
            const double* ri = W + (size_t)(i % WIN) * LD;
            acc = fma(ri[a], ri[b], acc);
        }
        Gw[(size_t)a * n + b] = acc;
    }
    for (int a = tid; a < Np; a += kGVThreads) {
        const int lim = min(k, a + 1);
        double acc = 0;
        for (int i = 0; i < lim; ++i) {
            const double* ri = W + (size_t)(i % WIN) * LD;
            acc = fma(ri[a], ri[Np], acc);
        }
        zw[a] = acc;
    }
    if (Q.emit_R) {                                                // R-form consumers: the kept rows themselves
        for (int o = tid; o < n * n; o += kGVThreads) {
            const int i = o / n, c = o - i * n;
            Q.Rc[o] = (i < k && c >= i && c < Np) ? W[(size_t)(i % WIN) * LD + c] : 0.0;
        }
        for (int i = tid; i < n; i += kGVThreads) Q.yc[i] = (i < k) ? W[(size_t)(i % WIN) * LD + Np] : 0.0;
    }
    if (tid == 0) {
        Q.rr[1] = k; Q.rr[4] = s_full;
        cnt[6] = (double)k;
        cnt[7] = 2.0 + 8.0 + ((s_tr[0] - s_tr[1]) >= 1e-8 ? 1.0 : 0.0);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t givens_window_doubles(int n) { return (size_t)(2 * n + kGVPrefetch + 2) * (size_t)(n + 1); }

size_t givens_smem_bytes(int n, bool* smem_window)
{
    const size_t full = sizeof(double) * (givens_window_doubles(n) + (size_t)n + 8);
    if (full <= 200 * 1024) { *smem_window = true; return full; }
    *smem_window = false;
    return sizeof(double) * ((size_t)n + 8);
}

size_t solve_small_smem_bytes(int n, int d)
{
    return sizeof(double) * ((size_t)(n + d + 1) * (n + 1) + (size_t)n * (n + 1) + (size_t)n * (d + 1) + 16);
}

constexpr int kRankSmallMaxN = 96;        // windows up to 16 clones: one thread per row of the factor in shared memory
constexpr int kSolveSmallRMaxN = 72;      // windows up to 12 clones: the whole EKF step in one CTA (197 KB of shared memory)

int compress_configure(int nmax)
{
    {
        const int ns = nmax < kRankSmallMaxN ? nmax : kRankSmallMaxN;
        RVIO_CUDA_TRY(cudaFuncSetAttribute(k_rank_rule<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * ((size_t)ns * ((ns + 2) | 1) + 8))));
        if (nmax > kRankSmallMaxN)
            RVIO_CUDA_TRY(cudaFuncSetAttribute(k_rank_rule<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * ((size_t)(nmax + 1) * (nmax + 2) / 2 + 8))));
    }
    // a handle sized for nmax also serves smaller windows (the filter's warm-up, other configurations): every variant is
    // given the largest dynamic shared memory it can be launched with
    bool w;
    size_t gv = givens_smem_bytes(nmax, &w);
    if (!w) {
        RVIO_CUDA_TRY(cudaFuncSetAttribute(k_givens_ref<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gv));
        gv = 200 * 1024;
    }
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_givens_ref<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gv));
    {
        const int ns = nmax < kSolveSmallRMaxN ? nmax : kSolveSmallRMaxN;
        RVIO_CUDA_TRY(cudaFuncSetAttribute(k_solve_small_R, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_small_smem_bytes(ns, 24 + ns)));
    }
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_chol_S, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * ((size_t)nmax * (nmax + 1) / 2 + 8))));
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * ((size_t)nmax * (nmax + 1) / 2 + nmax + 8))));
    return RVIO_OK;
}

// Enqueues the rank rule on [G | z | counters | classes] (after the all-reduce in the feature-sharded form).
int enqueue_rank_rule(cudaStream_t s, const RankRuleParams& rq, const GivensRefParams& gq_in, int n)
{
    if (n + 1 > kSFMaxRows) { set_error("enqueue_rank_rule", "window too large"); return RVIO_ERR_CAPACITY; }
    if (n <= kRankSmallMaxN) RVIO_LAUNCH(k_rank_rule<false>, 1, kSFThreads, sizeof(double) * ((size_t)n * ((n + 2) | 1) + 8), s, rq);
    else RVIO_LAUNCH(k_rank_rule<true>, 1, kSFThreads, sizeof(double) * ((size_t)(n + 1) * (n + 2) / 2 + 8), s, rq);
    if (rq.world == 1) {
        bool w;
        const size_t gv = givens_smem_bytes(n, &w);
        if (w) RVIO_LAUNCH(k_givens_ref<true>, 1, kGVThreads, gv, s, gq_in);
        else RVIO_LAUNCH(k_givens_ref<false>, 1, kGVThreads, gv, s, gq_in);
    }
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

// Whole small-window EKF step in one CTA (n + d + 1 <= kSFMaxRows).
int enqueue_solve_small_R(cudaStream_t s, const SolveSmallRParams& q)
{
    const int n = 6 * q.N;
    if (n > kSolveSmallRMaxN || n + q.d + 1 > kSRThreads) { set_error("enqueue_solve_small_R", "window too large for the single-CTA solve"); return RVIO_ERR_CAPACITY; }
    RVIO_LAUNCH(k_solve_small_R, 1, kSRThreads, solve_small_smem_bytes(n, q.d), s, q);
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

// Serial part of the large-window EKF step: S (n x n) -> L ; B (n x nb, leading dimension ldb) <- L^-1 B.
int enqueue_chol_trsm(cudaStream_t s, const double* S, int n, double* L, double* B, int ldb, int nb, int* bad, const double* gate)
{
    if (n > kCholThreads) { set_error("enqueue_chol_trsm", "window too large"); return RVIO_ERR_CAPACITY; }
    double* invd = L + (size_t)n * (n + 1) / 2;                 // the scratch is n x n: room for the packed factor and 1 / diag
    RVIO_LAUNCH(k_chol_S, 1, kCholThreads, sizeof(double) * ((size_t)n * (n + 1) / 2 + 8), s, S, n, L, invd, bad, gate);
    RVIO_LAUNCH(k_trsm, div_up(nb, kTrsmCols), 192, sizeof(double) * ((size_t)n * (n + 1) / 2 + n + 8), s, L, invd, n, B, ldb, nb, gate);
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

}  // namespace rvio

#ifdef RVIO_B200_PHASE_CLOCKS
extern "C" int rvio_b200_phase_clocks(long long* out, int n)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, rvio::g_phase_clk, sizeof(long long) * (n < 64 ? n : 64));
}
#endif
