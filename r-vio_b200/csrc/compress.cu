// compress.cu -- the reference's model-compression rule (Updater.cc:474-536) on the device.
//
// The reference compresses the stacked system with a bottom-up sweep of Givens rotations on adjacent rows, column by
// column (Updater.cc:494-512), after dropping trailing all-zero columns (:480-489), and then keeps the rows of the
// resulting trapezoid UP TO THE FIRST ROW WHOSE NORM IS BELOW 1e-4 (:515-524).  With noise s^2 I the EKF step depends on
// the kept rows Hn, rn only through  G = Hn^T Hn,  z = Hn^T rn, so the product path keeps its normal-term form
// (k_gram -> solve) and this file only decides WHICH rows the reference keeps and rewrites [G | z] accordingly:
//
//   k_rank_rule   (every compressed frame, one CTA)   pivot-free Cholesky of G = R^T R, which yields the rows of the
//                 reference's trapezoid as long as the leading columns are independent (R is unique up to row signs for
//                 ANY orthogonal triangularisation).  Let j* be the first dependent column.  If the information left after
//                 j* columns (the trace of the Schur complement = ||B||_F^2 of the remaining block B) is below (1e-4)^2, the
//                 reference's row j* (a unit combination of the rows of B) is below 1e-4: the cut is at j* and discards
//                 nothing -> [G | z] stay as they are.  If an earlier row is already below 1e-4 the cut is there and
//                 [G | z] are rebuilt from the kept rows.  Otherwise (a dependent column in the MIDDLE with information
//                 after it -- e.g. '2' features covering the first clones, '1' features the last ones, disjoint supports)
//                 the outcome depends on the reference's own rotation order and exact-zero pattern: k_givens_ref decides.
//   k_givens_ref  (only then, one CTA)   the reference's Givens sweep itself, rotation for rotation, scheduled as a
//                 wavefront: rotation (column n, rows m-1,m) runs at step t = (M-1-m) + 2n; all rotations of a step touch
//                 disjoint row pairs, M + N' - 2 dependent steps instead of ~M N'.  The rows live in a circular
//                 shared-memory window of 2N'+8 rows that slides up the stacked matrix (each row of H is read from HBM
//                 exactly once, by cp.async, a few steps ahead).  Special cases of Eigen's makeGivens (q == 0 -> identity,
//                 p == 0 -> row swap) are kept exactly: they are what moves the exact zeros around and make the reference's
//                 outcome deterministic.  Then the first-small-row cut, and [G | z] := kept rows.
//
// Both kernels read the mode from device memory (0 = reference rule, 1 = full information: keep [G | z] of all rows), so
// captured frame graphs stay valid when the mode is switched.
#include "common.cuh"
#include "compress_kernels.cuh"

namespace rvio {

namespace {

constexpr int kRRThreads = 512;
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
// feature that would only start later is discarded -- and above it (by orders of magnitude, it is a random unit
// combination of rows holding >= 1e-3 of information) otherwise.  The factorisation therefore runs over the TOTAL G once,
// skipping dependent columns, and at every column where a new class of features starts it compares the information left
// in the active part with what is still to come:
//     active part exhausted (< 1e-8), something dependent behind us  ->  cut here, kept = good pivot rows so far
//     exactly one dependent column behind us, active part alive (>= 1e-3)  ->  the reference is still going: continue
//     anything in between  ->  k_givens_ref replays the reference's sweep
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRRThreads) k_rank_rule(RankRuleParams Q)
{
    extern __shared__ __align__(16) double sm[];
    __shared__ int s_np, s_k, s_smin, s_ncls;
    __shared__ double s_tau;
    const int n = Q.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* cnt = Q.red + (size_t)n * n + n;
    const double* cls = cnt + 8;
    const double* G = Q.red;
    const int rows = (int)cnt[1];
    if (tid == 0) { Q.rr[0] = 0; Q.rr[1] = rows; Q.rr[2] = n; Q.rr[3] = 0; s_np = 0; s_smin = n + 1; s_ncls = 0; }
    const bool active = (cnt[0] > 2.0) && (*Q.rule_dev == 0) && rows > n;     // Updater.cc:460,474
    __syncthreads();
    if (!active) {
        if (tid == 0) { cnt[6] = (cnt[0] > 2.0) ? (double)rows : 0.0; cnt[7] = 0.0; }
        return;
    }
    // trailing all-zero columns (Updater.cc:480-489): column norm == 0  <=>  G(j,j) == 0
    for (int j = tid; j < n; j += kRRThreads)
        if (G[(size_t)j * n + j] != 0.0) atomicMax(&s_np, j + 1);
    for (int c = tid; c <= n; c += kRRThreads)
        if (cls[c] > 0.0) { atomicMin(&s_smin, c); atomicAdd(&s_ncls, 1); }
    __syncthreads();
    const int Np = s_np;
    const int ld = Np | 1;
    double* U = Q.use_glob ? Q.U_glob : sm;                       // Np x ld, upper triangle used
    double* aux = Q.use_glob ? sm : sm + (size_t)Np * ld;        // zt[Np], gd[Np], nr2[Np], late[Np + 1]
    double* zt = aux; double* gd = aux + Np; double* nr2 = aux + 2 * Np; double* late = aux + 3 * Np;
    for (int o = tid; o < Np * Np; o += kRRThreads) {
        const int i = o / Np, k = o - i * Np;
        if (k >= i) U[(size_t)i * ld + k] = G[(size_t)i * n + k];
    }
    for (int j = tid; j < Np; j += kRRThreads) { zt[j] = Q.red[(size_t)n * n + j]; gd[j] = G[(size_t)j * n + j]; }
    if (tid == 0) {                                               // late[j] = information of the classes starting at column >= j
        double acc = 0;
        for (int c = n; c > Np; --c) acc += cls[c];
        for (int j = Np; j >= 0; --j) { acc += cls[j]; late[j] = acc; }
    }
    __syncthreads();
    // no feature starts at column 0 (no '2' feature, no full-length '1'): the first rows of the reference's trapezoid are
    // raw rows in list order -- with several classes only the sweep knows what a later cut would keep
    if (s_smin > 0 && s_ncls > 1) {
        if (tid == 0) {
            const int mode = (Q.world == 1) ? 3 : 4;
            Q.rr[0] = mode; Q.rr[2] = Np; Q.rr[3] = (mode == 3) ? 1 : 0;
            cnt[6] = (double)rows; cnt[7] = (mode == 4) ? 4.0 : 0.0;
        }
        return;
    }

    // pivot-free right-looking Cholesky of the total G, dependent columns skipped (rows are left unscaled: row j of the
    // factor = U[j][j..] / sqrt(U[j][j]))
    const int ty = tid >> 4, tx = tid & 15;
    int q = 0, first_dep = Np, mode = 1, kcut = 0;
    for (int j = 0; j < Np; ++j) {
        if (j > 0 && late[j] > late[j + 1] && s_smin == 0) {      // a class of features starts here (cls[j] > 0)
            if (warp == 0) {
                const double* dg = U;
                double t = 0;
                for (int k = j + lane; k < Np; k += 32) t += dg[(size_t)k * ld + k];
                t = warp_sum_d(t);
                if (lane == 0) s_tau = t - late[j];               // information left in the active rows
            }
            __syncthreads();
            const double tau = s_tau;
            const int dd = j - q;
            if (dd >= 1) {
                if (tau < 1e-8) { mode = 2; kcut = j; break; }                 // exhausted: the reference cuts, later classes are discarded
                if (tau < 1e-3 || dd >= 2) { mode = 3; break; }               // only the reference's own sweep can tell
            }
            __syncthreads();                                       // s_tau is rewritten at the next boundary
        }
        const double pj = U[(size_t)j * ld + j];
        const double thr = fmax(1e-12, 1e-12 * gd[j]);
        if (!(pj >= thr)) {                                        // dependent (or empty) column: skipped, nothing changes
            if (first_dep == Np) first_dep = j;
            if (tid == 0) nr2[j] = -1.0;
            continue;
        }
        q++;
        const double rp = 1.0 / pj;
        const double* rowj = U + (size_t)j * ld;
        for (int i = j + 1 + ty; i < Np; i += kRRThreads / 16) {
            const double f = rowj[i] * rp;
            double* rowi = U + (size_t)i * ld;
            for (int k = i + tx; k < Np; k += 16) rowi[k] -= f * rowj[k];
            if (tx == 0) zt[i] -= f * zt[j];
        }
        __syncthreads();
    }
    const int jend = (mode == 2) ? kcut : Np;                      // columns whose rows may be kept
    // the reference's own test on the rows that are unique (before the first dependent column): norm < 1e-4 (Updater.cc:519)
    if (tid == 0) s_k = Np + 1;
    __syncthreads();
    const int ulim = min(first_dep, jend);
    for (int j = warp; j < ulim; j += kRRThreads / 32) {
        const double* rowj = U + (size_t)j * ld;
        double s = 0;
        for (int k = j + lane; k < Np; k += 32) s += rowj[k] * rowj[k];
        s = warp_sum_d(s);
        if (lane == 0) { nr2[j] = s / rowj[j]; if (s / rowj[j] < 1e-8) atomicMin(&s_k, j); }
    }
    __syncthreads();
    int kept_rows;               // what is reported as the rank
    bool rebuild = false, discards = false, generic = false;
    int klim = jend;             // rows with index < klim (and a good pivot) are kept
    if (s_k <= Np) {             // an early small row: the reference stops there, whatever comes later
        klim = s_k; kept_rows = s_k; rebuild = true; mode = 2;
        // information dropped: everything except the kept rows
        if (warp == 0) {
            double tr = 0, kept = 0;
            for (int j = lane; j < Np; j += 32) tr += gd[j];
            for (int j = lane; j < klim; j += 32) kept += nr2[j];
            tr = warp_sum_d(tr); kept = warp_sum_d(kept);
            if (lane == 0) s_tau = tr - kept;
        }
        __syncthreads();
        discards = s_tau >= 1e-8;
    } else if (mode == 2) {      // cut at a class boundary: the later classes are discarded
        kept_rows = q; rebuild = true; discards = late[kcut] >= 1e-8;
    } else if (mode == 1) {      // ran to the end: every row is kept
        kept_rows = q; generic = first_dep < Np && q > first_dep;   // dependent columns in the middle of the active set
    } else {
        kept_rows = q;
    }
    if (mode == 3 && Q.world != 1) mode = 4;
    if (tid == 0) {
        Q.rr[0] = mode; Q.rr[1] = kept_rows; Q.rr[2] = Np; Q.rr[3] = (mode == 3) ? 1 : 0;
        cnt[6] = (double)kept_rows;
        cnt[7] = (rebuild ? 8.0 : 0.0) + (discards ? 1.0 : 0.0) + (mode == 4 ? 4.0 : 0.0) + (generic ? 16.0 : 0.0);
    }
    if (!rebuild) return;
    // G' = sum_{kept i} row_i row_i^T / p_i ,  z' = sum_{kept i} row_i zt_i / p_i   (bitwise symmetric: products commute)
    double* Gw = Q.red; double* zw = Q.red + (size_t)n * n;
    for (int o = tid; o < Np * Np; o += kRRThreads) {
        const int a = o / Np, b = o - a * Np;
        const int lim = min(klim, min(a, b) + 1);
        double acc = 0;
        for (int i = 0; i < lim; ++i) {
            const double* rowi = U + (size_t)i * ld;
            const double pi = rowi[i];
            if (pi >= fmax(1e-12, 1e-12 * gd[i])) acc = fma(rowi[a] * rowi[b], 1.0 / pi, acc);
        }
        Gw[(size_t)a * n + b] = acc;
    }
    for (int a = tid; a < Np; a += kRRThreads) {
        const int lim = min(klim, a + 1);
        double acc = 0;
        for (int i = 0; i < lim; ++i) {
            const double* rowi = U + (size_t)i * ld;
            const double pi = rowi[i];
            if (pi >= fmax(1e-12, 1e-12 * gd[i])) acc = fma(rowi[a] * zt[i], 1.0 / pi, acc);
        }
        zw[a] = acc;
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
    if (tid == 0) {
        Q.rr[1] = k; Q.rr[4] = s_full;
        cnt[6] = (double)k;
        cnt[7] = 2.0 + 8.0 + ((s_tr[0] - s_tr[1]) >= 1e-8 ? 1.0 : 0.0);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
size_t rank_rule_smem_bytes(int n, bool* use_glob)
{
    const size_t ld = (size_t)(n | 1);
    const size_t full = sizeof(double) * ((size_t)n * ld + 4 * (size_t)n + 16);
    if (full <= 200 * 1024) { *use_glob = false; return full; }
    *use_glob = true;
    return sizeof(double) * (4 * (size_t)n + 16);
}

size_t givens_window_doubles(int n) { return (size_t)(2 * n + kGVPrefetch + 2) * (size_t)(n + 1); }

size_t givens_smem_bytes(int n, bool* smem_window)
{
    const size_t full = sizeof(double) * (givens_window_doubles(n) + (size_t)n + 8);
    if (full <= 200 * 1024) { *smem_window = true; return full; }
    *smem_window = false;
    return sizeof(double) * ((size_t)n + 8);
}

int compress_configure(int nmax)
{
    // a handle sized for nmax also serves smaller windows (the filter's warm-up, other configurations): every variant is
    // given the largest dynamic shared memory it can be launched with
    bool g;
    size_t rr = rank_rule_smem_bytes(nmax, &g);
    if (g) rr = 200 * 1024;
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_rank_rule, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rr));
    bool w;
    size_t gv = givens_smem_bytes(nmax, &w);
    if (!w) {
        RVIO_CUDA_TRY(cudaFuncSetAttribute(k_givens_ref<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gv));
        gv = 200 * 1024;
    }
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_givens_ref<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gv));
    return RVIO_OK;
}

// Enqueues the rank rule on [G | z | counters] (after the all-reduce in the feature-sharded form).
int enqueue_rank_rule(cudaStream_t s, const RankRuleParams& rq_in, const GivensRefParams& gq_in, int nmax)
{
    RankRuleParams rq = rq_in;
    bool g;
    const size_t rr = rank_rule_smem_bytes(nmax, &g);
    rq.use_glob = g ? 1 : 0;
    RVIO_LAUNCH(k_rank_rule, 1, kRRThreads, rr, s, rq);
    if (rq.world == 1) {
        bool w;
        const size_t gv = givens_smem_bytes(nmax, &w);
        if (w) RVIO_LAUNCH(k_givens_ref<true>, 1, kGVThreads, gv, s, gq_in);
        else RVIO_LAUNCH(k_givens_ref<false>, 1, kGVThreads, gv, s, gq_in);
    }
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

}  // namespace rvio
