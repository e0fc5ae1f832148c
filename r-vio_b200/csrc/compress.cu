// compress.cu -- the reference's model-compression rule (Updater.cc:474-536) on the device.
//
// The reference compresses the stacked system with a bottom-up sweep of Givens rotations on adjacent rows, column by
// column (Updater.cc:494-512), after dropping trailing all-zero columns (:480-489), and then keeps the rows of the
// resulting trapezoid UP TO THE FIRST ROW WHOSE NORM IS BELOW 1e-4 (:515-524).  With noise s^2 I the EKF step depends on
// the kept rows Hn, rn only through  G = Hn^T Hn,  z = Hn^T rn, so the product path keeps its normal-term form
// (k_gram -> solve) and this file only decides WHICH rows the reference keeps and rewrites [G | z] accordingly:
//
//   k_rank_rule   (every updating frame, one CTA)   pivot-free blocked Cholesky of G = R^T R (tile_cholesky.cuh),
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
//   k_update_small (<= 14 clones: rule + sweep + EKF step in one launch) / k_chol_S + k_trsm   the EKF step itself on the kept
//                 rows (Updater.cc:540-619), see below.
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

#ifdef RVIO_B200_PHASE_CLOCKS
// (profiling build only: make PHASES=1) per-phase SM clock stamps of the single-CTA kernels, read back by tools/prof_update.py
__device__ long long g_phase_clk[64];
#define PHASE_CLK(k) do { if (threadIdx.x == 0) g_phase_clk[k] = clock64(); } while (0)
#else
#define PHASE_CLK(k) do { } while (0)
#endif

}  // namespace rvio
#include "tile_cholesky.cuh"
namespace rvio {

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
// rsm: the tile-packed factor (lower: column j of L = row j of R), y in the extra row.  Rt_smem (optional, fused small-window
// step): the kept rows go straight into the EKF step's shared-memory tiles of R (upper triangle by tile, rt_tc tile columns)
// and yc_smem instead of global memory.
__device__ __forceinline__ void rank_rule_body(const RankRuleParams& Q, double* rsm, double* Rt_smem, int rt_tc, double* yc_smem)
{
    __shared__ double s_pv[kSFMaxRows], s_gd[kSFMaxRows], s_late[kSFMaxRows + 2], s_nr2[kSFMaxRows];
    __shared__ double s_part[9 * kBCWarps];
    __shared__ PanelPub s_pub;
    __shared__ int s_np, s_k, s_smin, s_ncls, s_verdict[6];
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
    for (int j = tid; j < n; j += kBCThreads) {
        const double gd = G[(size_t)j * n + j];
        s_gd[j] = gd;
        if (gd != 0.0) atomicMax(&s_np, j + 1);
    }
    for (int c = tid; c <= n; c += kBCThreads)
        if (cls[c] > 0.0) { atomicMin(&s_smin, c); atomicAdd(&s_ncls, 1); }
    __syncthreads();
    const int Np = s_np;
    PHASE_CLK(16);
    for (int c = tid; c <= n; c += kBCThreads) s_nr2[c < kSFMaxRows ? c : kSFMaxRows - 1] = cls[c];      // staged: the suffix sum below must not walk global memory
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

    TileTri T;
    T.t = rsm; T.tc = (Np + 7) >> 3; T.tr = T.tc + 1;
    const int ncp = 8 * T.tc;                                     // padded column count; row ncp carries z, later y
    int q = 0, first_dep = Np;                                    // (thread-uniform copies)
    int mode = 1, kcut = 0;
    bool undecided = false;
    int jstop = Np;
    {
        // [G | z] -> tiles, asynchronously (every 16-byte piece of every tile is written: zero where there is no data)
        const int nsq = T.tc * (T.tc + 1) / 2;
        for (int o = tid; o < (nsq + T.tc) * 32; o += kBCThreads) {
            const int tl = o >> 5, r = (o >> 2) & 7, q2 = 2 * (o & 3);
            int I, J;
            if (tl < nsq) {
                I = (int)((sqrtf(8.f * (float)tl + 1.f) - 1.f) * .5f);
                while (I * (I + 1) / 2 > tl) --I;
                while ((I + 1) * (I + 2) / 2 <= tl) ++I;
                J = tl - I * (I + 1) / 2;
            } else { I = T.tc; J = tl - nsq; }
            const int i = 8 * I + r, j = 8 * J + q2;
            const double* src; int nv = 0;
            if (I < T.tc) { src = G + (size_t)i * n + j; if (i < Np) nv = max(0, min(2, Np - j)); }
            else { src = Q.red + (size_t)n * n + j; if (r == 0) nv = max(0, min(2, Np - j)); }
            cp_async16_zfill(rsm + ((size_t)tl << 6) + r * 8 + q2, src, nv, G);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        for (int j = Np + tid; j < ncp; j += kBCThreads) *T.at(j, j) = 1.0;      // identity padding up to a multiple of 8 columns
        PHASE_CLK(17);
        // class-boundary test (see above), walked by warp 0 along the panels; the verdict comes back through shared memory
        RankPivot piv; piv.gd = s_gd; piv.np = Np;
        RankWalk rw;
        rw.q = 0; rw.first_dep = Np; rw.mode = 1; rw.kcut = 0; rw.jstop = Np; rw.undecided = 0;
        rw.np = Np; rw.world = Q.world; rw.boundaries = boundaries ? 1 : 0; rw.late = s_late; rw.nr2 = s_nr2;
        if (Np > 0) tile_cholesky<true>(T, Np, piv, &rw, s_pv, &s_pub, s_part, nullptr);
        if (tid == 0) { s_verdict[0] = rw.q; s_verdict[1] = rw.first_dep; s_verdict[2] = rw.mode; s_verdict[3] = rw.kcut; s_verdict[4] = rw.jstop; s_verdict[5] = rw.undecided; }
        __syncthreads();
        q = s_verdict[0]; first_dep = s_verdict[1]; mode = s_verdict[2]; kcut = s_verdict[3]; jstop = s_verdict[4]; undecided = s_verdict[5] != 0;
    }
    PHASE_CLK(18);
    __syncthreads();
    if (raw_top) { if (Q.world == 1) mode = 3; else undecided = true; }
    const int jend = (mode == 2) ? kcut : jstop;                  // columns whose rows may be kept
    // the reference's own test on the rows that are unique (before the first dependent column): norm < 1e-4 (Updater.cc:519)
    const int ulim = rule ? min(first_dep, jend) : 0;
    for (int j = tid; j < ulim; j += kBCThreads)
        if (s_pv[j] >= 0.0 && s_nr2[j] < 1e-8) atomicMin(&s_k, j);
    __syncthreads();
    PHASE_CLK(19);
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
    PHASE_CLK(20);
    if (Q.emit_R && mode != 3 && Rt_smem) {
        const int kl = klim, ntl = rt_tc * (rt_tc + 1) / 2;
        for (int o = tid; o < ntl * 64; o += kBCThreads) {       // tile (Rj, Kb >= Rj) at Rj rt_tc - Rj (Rj - 1) / 2 + Kb - Rj
            const int tl = o >> 6, e = o & 63;
            int Rj = 0, rem = tl;
            while (rem >= rt_tc - Rj) { rem -= rt_tc - Rj; ++Rj; }
            const int r = 8 * Rj + (e >> 3), k = 8 * (Rj + rem) + (e & 7);
            double v = 0;
            if (r < kl && r < Np && s_pv[r] >= 0.0 && k >= r && k < Np) v = *T.at(k, r);     // R(r, k) = L(k, r)
            Rt_smem[o] = v;
        }
        for (int j = tid; j < 8 * rt_tc; j += kBCThreads) yc_smem[j] = (j < kl && j < Np && s_pv[j] >= 0.0) ? *T.at(ncp, j) : 0.0;
    } else if (Q.emit_R && mode != 3) {
        // kept rows of R (= columns of L with a good pivot below klim), zero padded to n x n, and y.  Lane (a, b) of a warp
        // takes R(j0 + b, i0 + a .. ): 8 columns j of one tile row (contiguous in shared memory), 4 consecutive i per row of R
        const int kl = klim;
        const int ib = (n + 3) >> 2, jb = (n + 7) >> 3;
        for (int blk = warp; blk < ib * jb; blk += kBCWarps) {
            const int j = 8 * (blk / ib) + (lane & 7), i = 4 * (blk % ib) + (lane >> 3);
            if (i < n && j < n) {
                double v = 0;
                if (j < kl && j < Np && s_pv[j] >= 0.0 && i >= j && i < Np) v = *T.at(i, j);
                Q.Rc[(size_t)j * n + i] = v;
            }
        }
        for (int j = tid; j < n; j += kBCThreads) Q.yc[j] = (j < kl && j < Np && s_pv[j] >= 0.0) ? *T.at(ncp, j) : 0.0;
    }
    PHASE_CLK(21);
    if (!rebuild) return;
    // G' = sum_{kept j} L_j L_j^T ,  z' = sum_{kept j} L_j y_j   (bitwise symmetric: products commute)
    double* Gw = Q.red; double* zw = Q.red + (size_t)n * n;
    for (int o = tid; o < Np * Np; o += kBCThreads) {
        const int r = o / Np, c = o - r * Np;
        const int lim = min(klim, min(r, c) + 1);
        double acc = 0;
        for (int j = 0; j < lim; ++j)
            if (s_pv[j] >= 0.0) acc = fma(*T.at(r, j), *T.at(c, j), acc);
        Gw[(size_t)r * n + c] = acc;
    }
    for (int r = tid; r < Np; r += kBCThreads) {
        const int lim = min(klim, r + 1);
        double acc = 0;
        for (int j = 0; j < lim; ++j)
            if (s_pv[j] >= 0.0) acc = fma(*T.at(r, j), *T.at(ncp, j), acc);
        zw[r] = acc;
    }
}

__global__ void __launch_bounds__(kBCThreads, 1) k_rank_rule(RankRuleParams Q)
{
    extern __shared__ __align__(16) double rsm_dyn[];
    rank_rule_body(Q, rsm_dyn, nullptr, 0, nullptr);
}

// ------------------------------------------------------------------------------------------------
// Large-window EKF step, serial part:  S = L L^T (k_chol_S, one CTA, blocked / DMMA as above), then  Y = L^-1 [W | y]
// (k_trsm: one warp per 8 right-hand-side columns, tile row by tile row on the tensor pipe; the diagonal blocks are applied
// through their inverses, which k_chol_S leaves in the diagonal tiles of the factor it writes out).
// Lt: tile-packed lower factor in global memory, tile (I, J) at (I (I + 1) / 2 + J) * 64; diagonal tiles = inv(L_II).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBCThreads, 1) k_chol_S(const double* S, int m, double* Lt, int* bad, const double* gate)
{
    extern __shared__ __align__(16) double csm[];                // tile-packed lower triangle, then tc tiles for the inverses
    __shared__ double s_pv[kSFMaxRows];
    __shared__ PanelPub s_pub;
    if (gate && !(gate[0] > 2.0)) return;
    const int tid = threadIdx.x;
    TileTri T;
    T.t = csm; T.tc = (m + 7) >> 3; T.tr = T.tc;
    const int ntile = tile_tri_count(T.tc, T.tc);
    double* X = csm + (size_t)ntile * 64;
    for (int o = tid; o < ntile * 32; o += kBCThreads) {           // S -> tiles, asynchronously, zero where there is no data
        const int tl = o >> 5, r = (o >> 2) & 7, q2 = 2 * (o & 3);
        int I = (int)((sqrtf(8.f * (float)tl + 1.f) - 1.f) * .5f);
        while (I * (I + 1) / 2 > tl) --I;
        while ((I + 1) * (I + 2) / 2 <= tl) ++I;
        const int J = tl - I * (I + 1) / 2;
        const int i = 8 * I + r, j = 8 * J + q2;
        cp_async16_zfill(csm + ((size_t)tl << 6) + r * 8 + q2, S + (size_t)i * m + j, (i < m) ? max(0, min(2, m - j)) : 0, S);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    for (int j = m + tid; j < 8 * T.tc; j += kBCThreads) *T.at(j, j) = 1.0;
    PHASE_CLK(32);
    SpdPivot piv; piv.bad = bad;
    tile_cholesky<false>(T, 0, piv, nullptr, s_pv, &s_pub, nullptr, X);      // X: inv(L_JJ) of every panel
    PHASE_CLK(33);
    for (int o = tid; o < ntile * 64; o += kBCThreads) Lt[o] = csm[o];
    __syncthreads();
    for (int o = tid; o < T.tc * 64; o += kBCThreads) { const int J = o >> 6; Lt[((size_t)(J * (J + 1) / 2 + J) << 6) + (o & 63)] = X[o]; }
    PHASE_CLK(34);
}

// B: m x nb row-major right-hand sides (leading dimension ldb); solves L Y = B in place.  One warp per 8 columns:
//   for every tile row I:  C = B_I - sum_{K<I} L(I,K) Y_K   (DMMA, Y_K kept transposed in shared memory so that its fragment
//   is one LDS.128),  Y_I = inv(L_II) C  (one more tile product through a transposed scratch tile).
constexpr int kTrsmWarps = 4;
__global__ void __launch_bounds__(kTrsmWarps * 32) k_trsm(const double* Lt, int m, double* B, int ldb, int nb, const double* gate)
{
    extern __shared__ __align__(16) double sL[];                 // the factor's tiles, then per warp: tc transposed Y tiles + 1 scratch tile
    if (gate && !(gate[0] > 2.0)) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int tc = (m + 7) >> 3;
    const int ntile = tile_tri_count(tc, tc);
    {
        const double2* src = reinterpret_cast<const double2*>(Lt);
        double2* dst = reinterpret_cast<double2*>(sL);
        for (int e = tid; e < ntile * 32; e += kTrsmWarps * 32) dst[e] = src[e];      // linear copy: many loads in flight
    }
    double* YT = sL + (size_t)ntile * 64 + (size_t)warp * (tc + 1) * 64;
    double* CT = YT + (size_t)tc * 64;
    __syncthreads();
    const int col0 = (blockIdx.x * kTrsmWarps + warp) * 8;
    if (col0 >= nb) return;
    const int cj = col0 + 2 * t4;
    for (int I = 0; I < tc; ++I) {
        const int row = 8 * I + g;
        double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);
        if (row < m) {
            if (cj < nb) c0.x = B[(size_t)row * ldb + cj];
            if (cj + 1 < nb) c0.y = B[(size_t)row * ldb + cj + 1];
        }
        const double* lrow = sL + ((size_t)(I * (I + 1) / 2) << 6) + g * 8 + 2 * t4;
        int K = 0;
        for (; K + 1 < I; K += 2) {                              // two accumulators: the chain of dependent DMMAs is halved
            double2 a0 = *reinterpret_cast<const double2*>(lrow + K * 64), a1 = *reinterpret_cast<const double2*>(lrow + K * 64 + 64);
            const double2 y0 = *reinterpret_cast<const double2*>(YT + K * 64 + g * 8 + 2 * t4);
            const double2 y1 = *reinterpret_cast<const double2*>(YT + K * 64 + 64 + g * 8 + 2 * t4);
            a0.x = -a0.x; a0.y = -a0.y; a1.x = -a1.x; a1.y = -a1.y;
            dmma884(c0.x, c0.y, a0.x, y0.x);
            dmma884(c1.x, c1.y, a1.x, y1.x);
            dmma884(c0.x, c0.y, a0.y, y0.y);
            dmma884(c1.x, c1.y, a1.y, y1.y);
        }
        if (K < I) {
            double2 a0 = *reinterpret_cast<const double2*>(lrow + K * 64);
            const double2 y0 = *reinterpret_cast<const double2*>(YT + K * 64 + g * 8 + 2 * t4);
            a0.x = -a0.x; a0.y = -a0.y;
            tile_mma(c0, a0, y0);
        }
        c0.x += c1.x; c0.y += c1.y;
        // C (rows g, columns 2t, 2t+1) -> scratch, transposed: CT[column][row]
        CT[(2 * t4) * 8 + g] = c0.x;
        CT[(2 * t4 + 1) * 8 + g] = c0.y;
        __syncwarp();
        const double2 xi = *reinterpret_cast<const double2*>(lrow + I * 64);          // inv(L_II)[g][2t..]
        const double2 ct = *reinterpret_cast<const double2*>(CT + g * 8 + 2 * t4);     // C[2t..][g]
        double2 y = make_double2(0.0, 0.0);
        tile_mma(y, xi, ct);
        YT[I * 64 + (2 * t4) * 8 + g] = y.x;
        YT[I * 64 + (2 * t4 + 1) * 8 + g] = y.y;
        if (row < m) {
            if (cj < nb) B[(size_t)row * ldb + cj] = y.x;
            if (cj + 1 < nb) B[(size_t)row * ldb + cj + 1] = y.y;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// k_solve_small_R -- the whole EKF step of a small window (N <= 14 clones) in ONE CTA, R-form
// (Updater.cc:540-619 with Hn = R, the kept rows handed over by the rank rule, zero rows where a column was dropped):
//     W = R P[c,:]            (n x d: FP64 DMMA tile products, operands as 8 x 8 tiles in shared memory)
//     S = W[:,c] R^T + s^2 I  (n x n, SPD; DMMA)
//     [ S ; W^T ; y^T ]  ->  blocked Cholesky (tile_cholesky.cuh) with the d + 1 right-hand sides as extra tile rows: the extra
//                            rows of the factor are Y^T = (L^-1 [W | y])^T, i.e. factorisation and triangular solves in one sweep
//     dx = Y^T y~ ,  P+ = sym(P) - Y^T Y (DMMA; P copied into shared memory behind the factorisation),
//     state correction (quaternions multiplicative, Updater.cc:546-613)
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

constexpr int kSRThreads = kBCThreads;
// shared-memory tiles of k_solve_small_R: T (S + extra rows), R (upper triangle by tile), Pc (extra tile rows x tile columns)
// (the R / Pc region is reused for a copy of P during the factorisation: it is at least d x d doubles)
// Windows up to kPtStagedMaxN columns also stage P[c,:] as tiles next to R (all operands of W = R P[c,:] in shared memory);
// above that (13, 14 clones: the EuRoC default) the fragments of P come straight from global memory so that the step still
// fits in one CTA.
constexpr int kPtStagedMaxN = 72;
__host__ __device__ inline int solve_small_stage_doubles(int n, int d)
{
    const int tc = (n + 7) >> 3, tre = (d + 1 + 7) >> 3;
    const int a = (tc * (tc + 1) / 2 + (n <= kPtStagedMaxN ? tre * tc : 0)) * 64, b = d * d;
    return a > b ? a : b;
}
__host__ __device__ inline size_t solve_small_doubles(int n, int d)
{
    const int tc = (n + 7) >> 3, tre = (d + 1 + 7) >> 3;
    return (size_t)tile_tri_count(tc, tc + tre) * 64 + solve_small_stage_doubles(n, d);
}
// r_staged: the tiles of R are already in place (written by rank_rule_body in the same CTA) and y is in yc_smem.
__device__ __forceinline__ void solve_small_body(const SolveSmallRParams& Q, double* sm, bool r_staged, const double* yc_smem)
{
    __shared__ double s_pv[kSFMaxRows];
    __shared__ double s_dx[kSFMaxRows];
    __shared__ PanelPub s_pub;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int N = Q.N, n = 6 * N, d = Q.d;
    if (!(Q.gate[0] > 2.0)) {                                      // Updater.cc:621-627: too few features, posterior = prior
        for (int o = tid; o < d * d; o += kSRThreads) Q.P_out[o] = Q.P[o];
        for (int o = tid; o < Q.xdim; o += kSRThreads) Q.x_out[o] = Q.x[o];
        return;
    }
    // T: rows 0 .. 8 tc - 1: lower triangle of S, then of its Cholesky factor (columns n .. 8 tc - 1: s^2 I padding).  Row
    // 8 tc + c: column c of [W | y], after the factorisation column c of Y = L^-1 [W | y] (the right-hand sides ride along).
    TileTri T;
    T.t = sm; T.tc = (n + 7) >> 3;
    const int tc = T.tc, tre = (d + 1 + 7) >> 3;
    T.tr = tc + tre;
    double* Rt = sm + (size_t)tile_tri_count(tc, T.tr) * 64;       // tile (Rj, Kb), Kb >= Rj, at (Rj tc - Rj (Rj - 1) / 2 + Kb - Rj) * 64: R(8 Rj + r, 8 Kb + k) at [r][k]
    double* Pt = Rt + (size_t)(tc * (tc + 1) / 2) * 64;             // tile (Ci, Kb) at (Ci tc + Kb) * 64: Pc(8 Kb + k, 8 Ci + c) at [c][k]
    auto rtile = [&](int Rj, int Kb) -> double* { return Rt + ((size_t)(Rj * tc - (Rj * (Rj - 1)) / 2 + Kb - Rj) << 6); };
    PHASE_CLK(0);
    if (!r_staged)
    for (int o = tid; o < (tc * (tc + 1) / 2) * 32; o += kSRThreads) {          // R -> tiles (tiles of row Rj are consecutive)
        const int tl = o >> 5, rr = (o >> 2) & 7, q2 = 2 * (o & 3);
        int Rj = 0, rem = tl;
        while (rem >= tc - Rj) { rem -= tc - Rj; ++Rj; }
        const int r = 8 * Rj + rr, k = 8 * (Rj + rem) + q2;
        cp_async16_zfill(Rt + ((size_t)tl << 6) + rr * 8 + q2, Q.Rc + (size_t)r * n + k, (r < n) ? max(0, min(2, n - k)) : 0, Q.Rc);
    }
    const bool pt_staged = n <= kPtStagedMaxN;
    if (pt_staged)
    for (int o = tid; o < tre * tc * 32; o += kSRThreads) {                       // P(c, 24 + k) = Pc(k, c) -> tiles
        const int tl = o >> 5, cc = (o >> 2) & 7, q2 = 2 * (o & 3);
        const int Ci = tl / tc, Kb = tl - Ci * tc;
        const int c = 8 * Ci + cc, k = 8 * Kb + q2;
        cp_async16_zfill(Pt + ((size_t)tl << 6) + cc * 8 + q2, Q.P + (size_t)c * d + 24 + k, (c < d) ? max(0, min(2, n - k)) : 0, Q.P);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    PHASE_CLK(1);
    // ---- W^T tiles: Wt(Ci, Rj) = sum_{Kb >= Rj} Pt(Ci, Kb) Rt(Rj, Kb)^T   (R upper triangular)
    for (int w = warp; w < tre * tc; w += kBCWarps) {
        const int Ci = w / tc, Rj = w - Ci * tc;
        double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);
        const double* bp = rtile(Rj, Rj) + g * 8 + 2 * t4;                           // tiles (Rj, Kb), Kb = Rj .. : consecutive
        int nk = tc - Rj;
        if (pt_staged) {
            const double* ap = Pt + ((size_t)(Ci * tc + Rj) << 6) + g * 8 + 2 * t4;  // tiles (Ci, Kb), Kb = Rj .. : consecutive
            for (; nk >= 2; nk -= 2, ap += 128, bp += 128) {
                const double2 a0 = *reinterpret_cast<const double2*>(ap), a1 = *reinterpret_cast<const double2*>(ap + 64);
                const double2 b0 = *reinterpret_cast<const double2*>(bp), b1 = *reinterpret_cast<const double2*>(bp + 64);
                dmma884(c0.x, c0.y, a0.x, b0.x);
                dmma884(c1.x, c1.y, a1.x, b1.x);
                dmma884(c0.x, c0.y, a0.y, b0.y);
                dmma884(c1.x, c1.y, a1.y, b1.y);
            }
            if (nk) {
                const double2 a0 = *reinterpret_cast<const double2*>(ap);
                const double2 b0 = *reinterpret_cast<const double2*>(bp);
                tile_mma(c0, a0, b0);
            }
        } else {
            // the fragments of P (row c of P, columns 24 + k: P is symmetric) straight from global memory, all of a tile's
            // requests issued before its first product
            constexpr int kMaxKb = 11;                                               // ceil(84 / 8)
            const int cP = 8 * Ci + g;
            const double* pg = Q.P + (size_t)cP * d + 24 + 8 * Rj + 2 * t4;
            double2 av[kMaxKb];
#pragma unroll
            for (int q = 0; q < kMaxKb; ++q) {
                const int k = 8 * (Rj + q) + 2 * t4;
                av[q] = (q < nk && cP < d && k < n) ? *reinterpret_cast<const double2*>(pg + 8 * q) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int q = 0; q < kMaxKb; ++q) {
                if (q < nk) {
                    const double2 b0 = *reinterpret_cast<const double2*>(bp + 64 * q);
                    if (q & 1) tile_mma(c1, av[q], b0); else tile_mma(c0, av[q], b0);
                }
            }
        }
        c0.x += c1.x; c0.y += c1.y;
        // the row c = d of the extra rows is y, not a column of W
        const int c = 8 * Ci + g, r = 8 * Rj + 2 * t4;
        if (c == d) { const double* ycv = r_staged ? yc_smem : Q.yc; c0.x = (r < n) ? ycv[r] : 0.0; c0.y = (r + 1 < n) ? ycv[r + 1] : 0.0; }
        *reinterpret_cast<double2*>(T.tile(tc + Ci, Rj) + g * 8 + 2 * t4) = c0;
    }
    __syncthreads();
    PHASE_CLK(2);
    // ---- S tiles (lower): S(Ri, Cj) = sum_{Kb >= Cj} Wc(Ri, Kb) Rt(Cj, Kb)^T + s^2 I ;  Wc(r, k) = W(r, 24 + k) = extra row 24 + k, column r
    for (int w = warp; w < tc * (tc + 1) / 2; w += kBCWarps) {
        int Ri = 0, rem = w;
        while (rem > Ri) { rem -= Ri + 1; ++Ri; }
        const int Cj = rem;
        double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);
        const double* wt = T.tile(tc + 3 + Cj, Ri) + (2 * t4) * 8 + g;                  // extra tile rows: tc tiles each -> stride tc * 64
        const double* bp = rtile(Cj, Cj) + g * 8 + 2 * t4;
        for (int nk = tc - Cj, par = 0; nk > 0; --nk, par ^= 1, wt += (size_t)tc << 6, bp += 64) {
            const double a0 = wt[0], a1 = wt[8];                                         // W(r, 24 + k) at [k][r]
            const double2 b0 = *reinterpret_cast<const double2*>(bp);
            if (par) { dmma884(c1.x, c1.y, a0, b0.x); dmma884(c1.x, c1.y, a1, b0.y); }
            else { dmma884(c0.x, c0.y, a0, b0.x); dmma884(c0.x, c0.y, a1, b0.y); }
        }
        c0.x += c1.x; c0.y += c1.y;
        if (Ri == Cj) { if (g == 2 * t4) c0.x += Q.sig2; if (g == 2 * t4 + 1) c0.y += Q.sig2; }
        *reinterpret_cast<double2*>(T.tile(Ri, Cj) + g * 8 + 2 * t4) = c0;
    }
    __syncthreads();
    PHASE_CLK(3);
    // R and Pc are dead: the region takes a copy of P (the epilogue of P+ reads P(i,j) and P(j,i)); the copy runs behind the
    // factorisation
    double* sPm = Rt;
    for (int o = tid; o < (d * d) / 2; o += kSRThreads) cp_async16_zfill(sPm + 2 * o, Q.P + 2 * o, 2, Q.P);
    cp_async_commit();
    // ---- blocked Cholesky; the extra rows become Y^T
    {
        SpdPivot piv; piv.bad = Q.bad;
        tile_cholesky<false>(T, 0, piv, nullptr, s_pv, &s_pub, nullptr, nullptr);
    }
    cp_async_wait<0>();
    __syncthreads();
    PHASE_CLK(4);
    // ---- P+ = sym(P) - Y^T Y on the upper tile triangle (mirrored); the column j = d of Y^T Y is dx = Y^T y~
    for (int w = warp; w < tre * (tre + 1) / 2; w += kBCWarps) {
        int Jj = 0, rem = w;
        while (rem > Jj) { rem -= Jj + 1; ++Jj; }
        const int Ii = rem;                                                              // Ii <= Jj
        const int i = 8 * Ii + g, j = 8 * Jj + 2 * t4;
        double p00 = 0, p01 = 0, q0 = 0, q1 = 0;                                         // P(i, j), P(i, j+1), P(j, i), P(j+1, i) from the shared copy
        if (i < d && j < d) { p00 = sPm[i * d + j]; q0 = sPm[j * d + i]; }
        if (i < d && j + 1 < d) { p01 = sPm[i * d + j + 1]; q1 = sPm[(j + 1) * d + i]; }
        double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);
        const double* ai = T.tile(tc + Ii, 0) + g * 8 + 2 * t4;
        const double* bj = T.tile(tc + Jj, 0) + g * 8 + 2 * t4;
        int Kb = 0;
        for (; Kb + 1 < tc; Kb += 2) {
            const double2 a0 = *reinterpret_cast<const double2*>(ai + Kb * 64), a1 = *reinterpret_cast<const double2*>(ai + Kb * 64 + 64);
            const double2 b0 = *reinterpret_cast<const double2*>(bj + Kb * 64), b1 = *reinterpret_cast<const double2*>(bj + Kb * 64 + 64);
            dmma884(c0.x, c0.y, a0.x, b0.x);
            dmma884(c1.x, c1.y, a1.x, b1.x);
            dmma884(c0.x, c0.y, a0.y, b0.y);
            dmma884(c1.x, c1.y, a1.y, b1.y);
        }
        if (Kb < tc) {
            const double2 a0 = *reinterpret_cast<const double2*>(ai + Kb * 64);
            const double2 b0 = *reinterpret_cast<const double2*>(bj + Kb * 64);
            tile_mma(c0, a0, b0);
        }
        c0.x += c1.x; c0.y += c1.y;
        if (i < d) {
            if (j < d) {
                const double v = .5 * (p00 + q0) - c0.x;
                Q.P_out[(size_t)i * d + j] = v;
                if (Ii != Jj) Q.P_out[(size_t)j * d + i] = v;
            } else if (j == d) s_dx[i] = c0.x;
            if (j + 1 < d) {
                const double v = .5 * (p01 + q1) - c0.y;
                Q.P_out[(size_t)i * d + j + 1] = v;
                if (Ii != Jj) Q.P_out[(size_t)(j + 1) * d + i] = v;
            } else if (j + 1 == d) s_dx[i] = c0.y;
        }
    }
    __syncthreads();
    PHASE_CLK(5);
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
        double g3[3];
        for (int k = 0; k < 3; ++k) xo[4 + k] = dx[3 + k] + x[4 + k];
        for (int k = 0; k < 3; ++k) g3[k] = dx[6 + k] + x[7 + k];
        const double nn = sqrt(g3[0] * g3[0] + g3[1] * g3[1] + g3[2] * g3[2]);
        for (int k = 0; k < 3; ++k) xo[7 + k] = g3[k] / nn;
        for (int k = 0; k < 3; ++k) { }
        for (int k = 0; k < 12; ++k) xo[14 + k] = dx[12 + k] + x[14 + k];
    }
    PHASE_CLK(6);
}

__global__ void __launch_bounds__(kSRThreads, 1) k_solve_small_R(SolveSmallRParams Q)
{
    extern __shared__ __align__(16) double sm_dyn[];
    solve_small_body(Q, sm_dyn, false, nullptr);
}

// ------------------------------------------------------------------------------------------------
// k_givens_ref
// ------------------------------------------------------------------------------------------------

template <bool SMEM>
__device__ __forceinline__ void givens_ref_body(const GivensRefParams& Q, double* sm)
{
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

template <bool SMEM>
__global__ void __launch_bounds__(kGVThreads) k_givens_ref(GivensRefParams Q)
{
    extern __shared__ __align__(16) double gsm_dyn[];
    givens_ref_body<SMEM>(Q, gsm_dyn);
}

// ------------------------------------------------------------------------------------------------
// k_update_small -- small windows (N <= 14 clones): rank rule, (rarely) the reference's sweep, and the whole EKF step in ONE launch.
// The kept rows of R never leave shared memory (the rule writes them straight into the EKF step's tiles), the sweep's
// no-op launch between the two and a memset node disappear, and so do two kernel boundaries of the frame's critical path.
// Dynamic shared memory: [EKF step: T | R / Pc (later P)] [rank rule's factor]; the sweep (when it runs, the rule's factor is
// dead and the step has not started) takes the front of the same buffer.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBCThreads, 1) k_update_small(RankRuleParams rq, GivensRefParams gq, SolveSmallRParams sq)
{
    extern __shared__ __align__(16) double usm[];
    __shared__ double s_yc[kSFMaxRows];
    static_assert(kGVThreads == kBCThreads, "the fused step runs the sweep with the CTA it has");
    const int n = rq.n, d = sq.d;
    const int tc = (n + 7) >> 3, tre = (d + 1 + 7) >> 3;
    if (threadIdx.x == 0) *sq.bad = 0;
    double* Rt = usm + (size_t)tile_tri_count(tc, tc + tre) * 64;
    double* rank_tiles = usm;                                       // (inside the step's T region, which is written only after the rule has emitted R)
    rank_rule_body(rq, rank_tiles, Rt, tc, s_yc);
    __syncthreads();
    const bool updating = sq.gate[0] > 2.0;
    const bool sweep = updating && rq.rr[3] != 0;                  // (written by thread 0 of this CTA before the barrier)
    if (sweep) {
        givens_ref_body<true>(gq, usm);
        __threadfence_block();
        __syncthreads();
    }
    solve_small_body(sq, usm, updating && !sweep, s_yc);
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

size_t solve_small_smem_bytes(int n, int d) { return sizeof(double) * solve_small_doubles(n, d) + 64; }
static size_t rank_rule_smem_bytes(int n) { const int tc = (n + 7) / 8; return sizeof(double) * 64 * (size_t)tile_tri_count(tc, tc + 1) + 64; }
static size_t chol_smem_bytes(int n) { const int tc = (n + 7) / 8; return sizeof(double) * 64 * (size_t)(tile_tri_count(tc, tc) + tc) + 64; }
static size_t trsm_smem_bytes(int n) { const int tc = (n + 7) / 8; return sizeof(double) * 64 * (size_t)(tile_tri_count(tc, tc) + kTrsmWarps * (tc + 1)) + 64; }

static size_t update_small_smem_bytes(int n, int d)
{
    bool w;
    size_t a = solve_small_smem_bytes(n, d);
    const size_t r = rank_rule_smem_bytes(n), b = givens_smem_bytes(n, &w);
    if (r > a) a = r;
    return a > b ? a : b;
}
constexpr int kSolveSmallRMaxN = 84;      // windows up to 14 clones (the EuRoC default): the whole EKF step in one CTA (206 KB of shared memory)

int compress_configure(int nmax)
{
    // a handle sized for nmax also serves smaller windows (the filter's warm-up, other configurations): every kernel is
    // given the largest dynamic shared memory it can be launched with
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_rank_rule, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rank_rule_smem_bytes(nmax)));
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
    {
        const int ns = nmax < kSolveSmallRMaxN ? nmax : kSolveSmallRMaxN;
        RVIO_CUDA_TRY(cudaFuncSetAttribute(k_update_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_small_smem_bytes(ns, 24 + ns)));
    }
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_chol_S, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chol_smem_bytes(nmax)));
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_trsm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsm_smem_bytes(nmax)));
    return RVIO_OK;
}

// Enqueues the rank rule on [G | z | counters | classes] (after the all-reduce in the feature-sharded form).
int enqueue_rank_rule(cudaStream_t s, const RankRuleParams& rq, const GivensRefParams& gq_in, int n)
{
    if (n + 9 > kSFMaxRows) { set_error("enqueue_rank_rule", "window too large"); return RVIO_ERR_CAPACITY; }
    RVIO_LAUNCH(k_rank_rule, 1, kBCThreads, rank_rule_smem_bytes(n), s, rq);
    if (rq.world == 1) {
        bool w;
        const size_t gv = givens_smem_bytes(n, &w);
        if (w) RVIO_LAUNCH(k_givens_ref<true>, 1, kGVThreads, gv, s, gq_in);
        else RVIO_LAUNCH(k_givens_ref<false>, 1, kGVThreads, gv, s, gq_in);
    }
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

// Whole small-window EKF step in one CTA.
int enqueue_solve_small_R(cudaStream_t s, const SolveSmallRParams& q)
{
    const int n = 6 * q.N;
    if (n > kSolveSmallRMaxN) { set_error("enqueue_solve_small_R", "window too large for the single-CTA solve"); return RVIO_ERR_CAPACITY; }
    RVIO_LAUNCH(k_solve_small_R, 1, kSRThreads, solve_small_smem_bytes(n, q.d), s, q);
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

// Small windows: rank rule + (rare) sweep + EKF step in one launch.
int enqueue_update_small(cudaStream_t s, const RankRuleParams& rq, const GivensRefParams& gq, const SolveSmallRParams& q)
{
    const int n = 6 * q.N;
    bool w;
    givens_smem_bytes(n, &w);
    if (n > kSolveSmallRMaxN || n + 9 > kSFMaxRows || !w) { set_error("enqueue_update_small", "window too large for the single-CTA step"); return RVIO_ERR_CAPACITY; }
    RVIO_LAUNCH(k_update_small, 1, kBCThreads, update_small_smem_bytes(n, q.d), s, rq, gq, q);
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

// Serial part of the large-window EKF step: S (n x n) -> L (tile packed, needs tile_tri_count * 64 doubles of scratch: <= n x n + 64 n);
// B (n x nb, leading dimension ldb) <- L^-1 B.
int enqueue_chol_trsm(cudaStream_t s, const double* S, int n, double* L, double* B, int ldb, int nb, int* bad, const double* gate)
{
    if (n + 8 > kSFMaxRows) { set_error("enqueue_chol_trsm", "window too large"); return RVIO_ERR_CAPACITY; }
    RVIO_LAUNCH(k_chol_S, 1, kBCThreads, chol_smem_bytes(n), s, S, n, L, bad, gate);
    RVIO_LAUNCH(k_trsm, div_up(div_up(nb, 8), kTrsmWarps), kTrsmWarps * 32, trsm_smem_bytes(n), s, L, n, B, ldb, nb, gate);
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
