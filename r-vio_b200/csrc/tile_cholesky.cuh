// tile_cholesky.cuh -- blocked single-CTA Cholesky on a tile-packed lower trapezoid in shared memory, FP64 DMMA trailing updates.
// Used by compress.cu (rank rule, S factorisation, small-window EKF step) and updater.cu (the gate of large windows).
#pragma once
#include "common.cuh"

#ifndef PHASE_CLK
#define PHASE_CLK(k) do { } while (0)
#endif

namespace rvio {

// ------------------------------------------------------------------------------------------------
// Single-CTA factorisations (Cholesky of G for the rank rule, of S for the EKF step, of the per-feature gate matrix): BLOCKED,
// 8 columns per panel, on a TILE-PACKED lower trapezoid in shared memory (8 x 8 tiles of 64 doubles; tile (I, J) of the lower
// triangle at I (I + 1) / 2 + J, full tile rows below the square part for right-hand sides that ride along as extra rows):
//   per panel   (A) every warp turns its tile rows into L(I,J) = A(I,J) inv(L_JJ)^T                        -> barrier
//               (B) trailing tiles C(I,K) -= L(I,J) L(K,J)^T, tile rows handed out dynamically; warp 0 first updates and
//                   factors the NEXT diagonal tile (look-ahead) and publishes its inverse                   -> barrier
// All tile products are FP64 DMMA (mma.sync m8n8k4): both fragments are one conflict-free LDS.128 each (lane (g, t) takes
// columns 2t, 2t+1 of row g -- the k index of the product is permuted the same way on both operands), the C tile one
// LDS.128 / STS.128.  See tile_cholesky() below and DESIGN.md section 4 for the measured history of this routine.
// ------------------------------------------------------------------------------------------------
constexpr int kBCThreads = 512;           // threads of the blocked-Cholesky CTAs: 16 warps (a warp issues a dependent instruction every ~8 cycles:
                                          // the tile products need 4 warps per scheduler to keep the FP64 tensor pipe busy)
constexpr int kBCWarps = kBCThreads / 32;
constexpr int kSFMaxRows = 200;           // padded columns + right-hand side row (31 clones: 186 -> 192 + 8)

__device__ __forceinline__ void cp_async8(double* dst_smem, const double* src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// 16-byte copy of which only the first n_valid (0, 1, 2) doubles are read; the rest is zero filled.  `safe`: any valid
// 16-byte aligned global address (used when nothing is read).
__device__ __forceinline__ void cp_async16_zfill(double* dst_smem, const double* src, int n_valid, const double* safe)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    const double* sp = n_valid > 0 ? src : safe;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(sp), "r"(8 * n_valid) : "memory");
}

struct TileTri {
    double* t; int tc, tr;                // tc tile columns (square lower triangle), tr >= tc tile rows
    __device__ __forceinline__ int toff(int I, int J) const { return (I < tc ? (I * (I + 1)) / 2 : (tc * (tc + 1)) / 2 + (I - tc) * tc) + J; }
    __device__ __forceinline__ double* tile(int I, int J) const { return t + ((size_t)toff(I, J) << 6); }
    __device__ __forceinline__ double* at(int i, int j) const { return tile(i >> 3, j >> 3) + ((i & 7) << 3) + (j & 7); }
};
__host__ __device__ inline int tile_tri_count(int tc, int tr) { return tc * (tc + 1) / 2 + (tr - tc) * tc; }

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// C (8 x 8, lane (g, t) holds C[g][2t], C[g][2t+1]) += A B^T for two 8 x 8 row-major tiles given as their (g, 2t..2t+1) pairs
__device__ __forceinline__ void tile_mma(double2& c, const double2& a, const double2& b)
{
    dmma884(c.x, c.y, a.x, b.x);
    dmma884(c.x, c.y, a.y, b.y);
}

// What the factoring warp publishes about the diagonal tile of a panel.
struct __align__(16) PanelPub {
    double X[64];             // inv(L_JJ), row-major; rows / columns of skipped pivots are zero
    double cin[8];            // RANK: the tile's own share of |L(:, c)|^2 (real rows)
    double pv[8];             // pivots (as found, before the square root)
    double below[9];          // RANK scratch of the walk
    unsigned okmask;
    int next_row;             // work counter of the trailing update (tile rows are handed out from the bottom up)
    int stop;                 // RANK: the walk ended the factorisation in this panel
};

// Pivot rules (thread-uniform).  SPD: the matrix is S = W R^T + s^2 I; a non-positive pivot is reported.
struct SpdPivot {
    int* bad;
    __device__ __forceinline__ bool ok(int, double p) const { const bool k = p > 0.0; if (!k) *bad = 1; return k; }
};
// Rank rule: a column whose pivot drowned in rounding (relative to its original diagonal) is dependent and skipped;
// columns >= np are identity padding.
struct RankPivot {
    const double* gd; int np;
    __device__ __forceinline__ bool ok(int j, double p) const { return j >= np || p >= fmax(1e-12, 1e-12 * gd[j]); }
};

// State of the reference's rule along the columns; lives in warp 0 (every lane holds the same values).
struct RankWalk {
    int q, first_dep, mode, kcut, jstop, undecided;
    int np, world, boundaries;
    const double* late;       // late[j] = information of the classes starting at column >= j
    double* nr2;              // out: |row j of R|^2
    __device__ __forceinline__ bool is_boundary(int j) const { return boundaries && j > 0 && j < np && late[j] > late[j + 1]; }
};

// One warp factors the diagonal tile (J, J): lane (g, t) owns L[g][2t], L[g][2t+1] (the fragment layout); per column one
// broadcast of the pivot, one rsqrt, one broadcast of the scaled column, two FMAs -- ~30 instructions per column and lane
// instead of the ~100 of a redundant per-lane factorisation (a single warp issues one dependent instruction every ~8 cycles,
// so the instruction count of this chain IS the critical path of a small factorisation).  Then X = inv(L_JJ) (lanes 0..7,
// one column each), the pivots and the column norms.  Returns the trace of the tile's real rows before the factorisation
// when want_trace (RANK: only panels with a class boundary need it).
template <bool RANK, class Pivot>
__device__ __forceinline__ double factor_diag_tile(const TileTri& T, int J, int nact, const Pivot& piv, double* s_pv, PanelPub* pub,
                                                   double* Xkeep, bool want_trace, unsigned& okmask_out)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, g = lane >> 2, t4 = lane & 3;
    double* dt = T.tile(J, J);
    double2 e = *reinterpret_cast<const double2*>(dt + g * 8 + 2 * t4);
    double tr0 = 0;
    if (RANK && want_trace) {
        double v = (lane < 8 && 8 * J + lane < nact) ? dt[lane * 9] : 0.0;
        v += __shfl_xor_sync(FULL, v, 1); v += __shfl_xor_sync(FULL, v, 2); v += __shfl_xor_sync(FULL, v, 4);
        tr0 = __shfl_sync(FULL, v, 0);
    }
    if (J == 2) PHASE_CLK(RANK ? 56 : 60);
    unsigned okmask = 0;
    double pvc[8], rsv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const double colv = (c & 1) ? e.y : e.x;                    // this lane's element of the column pair c belongs to (meaningful when t4 == c / 2)
        const double p = __shfl_sync(FULL, colv, 4 * c + (c >> 1));
        pvc[c] = p;
        const bool ok = piv.ok(8 * J + c, p);
        okmask |= (ok ? 1u : 0u) << c;
        const double r_ = ok ? rsqrt(p) : 0.0;
        rsv[c] = r_;
        const double lc = (g >= c) ? colv * r_ : 0.0;               // L[g][c] in the lanes with t4 == c / 2 (rows above the diagonal: 0)
        if (t4 == (c >> 1)) { if (c & 1) e.y = lc; else e.x = lc; }
        const double lg = __shfl_sync(FULL, lc, 4 * g + (c >> 1));
        const double la = __shfl_sync(FULL, lc, 4 * (2 * t4) + (c >> 1));
        const double lb = __shfl_sync(FULL, lc, 4 * (2 * t4 + 1) + (c >> 1));
        if (2 * t4 > c) e.x = fma(-lg, la, e.x);
        if (2 * t4 + 1 > c) e.y = fma(-lg, lb, e.y);
    }
    if (J == 2) PHASE_CLK(RANK ? 57 : 61);
    if (2 * t4 > g) e.x = 0.0;                                      // strictly upper part
    if (2 * t4 + 1 > g) e.y = 0.0;
    *reinterpret_cast<double2*>(dt + g * 8 + 2 * t4) = e;
    if (RANK) {                                                     // column norms over the real rows: reduce over g
        const bool real = 8 * J + g < nact;
        double v0 = real ? e.x * e.x : 0.0, v1 = real ? e.y * e.y : 0.0;
#pragma unroll
        for (int o = 4; o <= 16; o <<= 1) { v0 += __shfl_xor_sync(FULL, v0, o); v1 += __shfl_xor_sync(FULL, v1, o); }
        if (lane < 4) { pub->cin[2 * lane] = v0; pub->cin[2 * lane + 1] = v1; }
    }
    if (lane == 0) {                                                // (one lane, straight-line stores: a per-lane selection of pvc[lane] compiles to a jump table)
#pragma unroll
        for (int c = 0; c < 8; ++c) { s_pv[8 * J + c] = ((okmask >> c) & 1u) ? pvc[c] : -1.0; pub->pv[c] = pvc[c]; }
        pub->okmask = okmask;
    }
    okmask_out = okmask;
    __syncwarp();
    if (J == 2) PHASE_CLK(RANK ? 58 : 62);
    // X = inv(L_JJ): lane k < 8 runs the forward substitution on e_k against the rows of L_JJ just written (broadcast reads);
    // a skipped pivot has a zero diagonal: its row of X is zero
    {
        const int k = lane & 7;
        double x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double v = (k == c) ? 1.0 : 0.0;
#pragma unroll
            for (int c1 = 0; c1 < c; ++c1) v = fma(-dt[c * 8 + c1], x[c1], v);
            x[c] = v * rsv[c];                                      // 1 / L[c][c] = rsqrt(p); 0 for a skipped pivot
        }
        if (lane < 8) {
#pragma unroll
            for (int c = 0; c < 8; ++c) { pub->X[c * 8 + lane] = x[c]; if (Xkeep) Xkeep[c * 8 + lane] = x[c]; }
        }
    }
    if (J == 2) PHASE_CLK(RANK ? 59 : 63);
    return tr0;
}

// The rule's bookkeeping for panel J (warp 0, after the panel's rows have been turned into L): column norms -> nr2, count of
// independent columns, first dependent column, and the class-boundary tests (see k_rank_rule).  Returns true to stop.
__device__ __forceinline__ bool rank_walk_panel(RankWalk& rw, int J, unsigned okmask, double tr0, PanelPub* pub, const double* s_part)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int j0 = 8 * J;
    unsigned colmask = (j0 + 8 <= rw.np) ? 0xffu : ((1u << (rw.np - j0)) - 1u);       // real columns of this panel
    double below = 0;                                                                  // lane c < 8: |L(rows below the tile, c)|^2 ; lane 8: their trace
    if (lane < 9) {
#pragma unroll
        for (int w = 0; w < kBCWarps; ++w) below += s_part[w * 9 + lane];
    }
    if (lane < 8 && ((colmask >> lane) & 1u)) rw.nr2[j0 + lane] = ((okmask >> lane) & 1u) ? pub->cin[lane] + below : 0.0;
    const unsigned bmask = __ballot_sync(FULL, lane < 8 && rw.is_boundary(j0 + lane)) & colmask;
    bool stop = false;
    if (bmask) {                                                                       // rare: a class of features starts inside this panel
        if (lane < 9) pub->below[lane] = below;
        __syncwarp();
        unsigned bm = bmask;
        while (bm) {
            const int cb = __ffs(bm) - 1;
            bm &= bm - 1;
            // trace of the Schur complement at column j over the real rows: the tile's rows (recursively: every good pivot
            // takes its column norm, a skipped one its pivot), the rows below (their trace minus what the columns took)
            double tau = tr0 + pub->below[8];
            for (int c = 0; c < cb; ++c) tau -= (((okmask >> c) & 1u) ? pub->cin[c] : pub->pv[c]) + pub->below[c];
            const int j = j0 + cb;
            const double tt = tau - rw.late[j];
            const int dd = j - (rw.q + __popc(okmask & colmask & ((1u << cb) - 1u)));
            if (dd >= 1) {
                if (tt < 1e-8) { rw.mode = 2; rw.kcut = j; stop = true; }              // exhausted: the reference cuts, later classes are discarded
                else if (tt < 1e-3 || dd >= 2) {                                       // only the reference's own sweep can tell
                    if (rw.world == 1) { rw.mode = 3; stop = true; }
                    else rw.undecided = 1;                                             // feature-sharded: keep everything, say so
                }
            }
            if (stop) { rw.jstop = j; colmask &= (1u << cb) - 1u; break; }
        }
    }
    rw.q += __popc(okmask & colmask);
    const unsigned deps = ~okmask & colmask;
    if (deps && rw.first_dep == rw.np) rw.first_dep = j0 + __ffs(deps) - 1;
    return stop;
}

// Blocked Cholesky of the trapezoid T (T.tr tile rows, T.tc tile columns), in place: L below and on the diagonal.
//   per panel J:  (A) every warp turns its tile rows of the panel into L(I,J) = A(I,J) inv(L_JJ)^T (two DMMAs per tile)
//                 (B) trailing update C(I,K) -= L(I,J) L(K,J)^T on the tensor pipe, tile rows handed out dynamically;
//                     warp 0 takes tile (J+1, J+1) first and factors it while the others are still updating (look-ahead:
//                     the dependent chain of the 8 x 8 factorisation is off the critical path when there is enough
//                     trailing work)
//   RANK: rw carries the reference's rule (walked by warp 0 only, which publishes the verdict; the other warps never wait
//   for it -- a trailing update past the cut is harmless, nothing behind the cut is read afterwards); nact = real columns.
//   s_part: kBCWarps x 9 doubles.  Xall (optional): tc tiles, receives inv(L_JJ) of every panel.
template <bool RANK, class Pivot>
__device__ __forceinline__ void tile_cholesky(const TileTri& T, int nact, const Pivot& piv, RankWalk* rw, double* s_pv, PanelPub* pub,
                                              double* s_part, double* Xall)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int fo = g * 8 + 2 * t4;                                  // this lane's pair inside a tile
    double tr0 = 0;
    unsigned okmask = 0;
    if (tid == 0) pub->stop = 0;
    __syncthreads();
    if (warp == 0) {
        bool wt = false;
        if (RANK) { for (int c = 0; c < 8; ++c) wt = wt || rw->is_boundary(c); }
        tr0 = factor_diag_tile<RANK>(T, 0, nact, piv, s_pv, pub, Xall, wt, okmask);
    }
    __syncthreads();
    for (int J = 0; J < T.tc; ++J) {
        if (J == 1) PHASE_CLK(RANK ? 48 : 40);
        // ---- (A) the panel below the diagonal tile
        double cn0 = 0, cn1 = 0, trp = 0;
        {
            const double2 xf = *reinterpret_cast<const double2*>(pub->X + fo);
            for (int I = J + 1 + warp; I < T.tr; I += 2 * kBCWarps) {
                const int I2 = I + kBCWarps;
                const bool two = I2 < T.tr;
                double* p0 = T.tile(I, J) + fo;
                double* p1 = two ? T.tile(I2, J) + fo : p0;
                const double2 a0 = *reinterpret_cast<const double2*>(p0), a1 = *reinterpret_cast<const double2*>(p1);
                double2 c0 = make_double2(0.0, 0.0), c1 = make_double2(0.0, 0.0);
                dmma884(c0.x, c0.y, a0.x, xf.x);
                dmma884(c1.x, c1.y, a1.x, xf.x);
                dmma884(c0.x, c0.y, a0.y, xf.y);
                dmma884(c1.x, c1.y, a1.y, xf.y);
                *reinterpret_cast<double2*>(p0) = c0;
                if (two) *reinterpret_cast<double2*>(p1) = c1;
                if (RANK) {
                    if (8 * I + g < nact) { cn0 = fma(c0.x, c0.x, cn0); cn1 = fma(c0.y, c0.y, cn1); }
                    if (two && 8 * I2 + g < nact) { cn0 = fma(c1.x, c1.x, cn0); cn1 = fma(c1.y, c1.y, cn1); }
                    if (lane < 8) {
                        if (I < T.tc && 8 * I + lane < nact) trp += T.tile(I, I)[lane * 9];
                        if (two && I2 < T.tc && 8 * I2 + lane < nact) trp += T.tile(I2, I2)[lane * 9];
                    }
                }
            }
        }
        if (RANK) {
#pragma unroll
            for (int o = 4; o <= 16; o <<= 1) { cn0 += __shfl_xor_sync(0xffffffffu, cn0, o); cn1 += __shfl_xor_sync(0xffffffffu, cn1, o); }
#pragma unroll
            for (int o = 1; o <= 4; o <<= 1) trp += __shfl_xor_sync(0xffffffffu, trp, o);
            if (lane < 4) { s_part[warp * 9 + 2 * lane] = cn0; s_part[warp * 9 + 2 * lane + 1] = cn1; }
            if (lane == 0) s_part[warp * 9 + 8] = trp;
        }
        if (J == 1) PHASE_CLK(RANK ? 49 : 41);
        if (tid == 0) pub->next_row = T.tr - 1;
        __syncthreads();
        if (J == 1) PHASE_CLK(RANK ? 50 : 42);
        const bool last = J + 1 >= T.tc;                            // the last panel has nothing behind it
        if (warp == 0) {
            bool stop = false;
            if (RANK) {
                stop = rank_walk_panel(*rw, J, okmask, tr0, pub, s_part);
                if (stop && lane == 0) pub->stop = 1;
            }
            if (J == 1) PHASE_CLK(RANK ? 51 : 43);
            if (!stop && !last) {                                   // look-ahead: the next diagonal tile first, then its factorisation
                double2 a = *reinterpret_cast<const double2*>(T.tile(J + 1, J) + fo);
                const double2 b = a;
                a.x = -a.x; a.y = -a.y;
                double* cp = T.tile(J + 1, J + 1) + fo;
                double2 c = *reinterpret_cast<double2*>(cp);
                tile_mma(c, a, b);
                *reinterpret_cast<double2*>(cp) = c;
                __syncwarp();
                if (J == 1) PHASE_CLK(RANK ? 52 : 44);
                bool wt = false;
                if (RANK) { for (int c = 0; c < 8; ++c) wt = wt || rw->is_boundary(8 * (J + 1) + c); }
                tr0 = factor_diag_tile<RANK>(T, J + 1, nact, piv, s_pv, pub, Xall ? Xall + (size_t)(J + 1) * 64 : nullptr, wt, okmask);
            }
        }
        if (last) break;
        // ---- (B) trailing update
        if (J == 1) PHASE_CLK(RANK ? 53 : 45);
        while (true) {
            int I = 0;
            if (lane == 0) I = atomicSub(&pub->next_row, 1);
            I = __shfl_sync(0xffffffffu, I, 0);
            if (I < J + 2) break;
            double2 a = *reinterpret_cast<const double2*>(T.tile(I, J) + fo);
            a.x = -a.x; a.y = -a.y;
            const int Kmax = min(I, T.tc - 1);
            const double* brow = T.tile(J + 1, J) + fo;             // tiles (K, J): K (K + 1) / 2 + J -> stride grows by K + 1 tiles
            double* crow = T.tile(I, J + 1) + fo;                   // tiles (I, J+1), (I, J+2), ... are consecutive
            int K = J + 1;
            for (; K + 1 <= Kmax; K += 2, crow += 128) {
                const double2 b0 = *reinterpret_cast<const double2*>(brow);
                brow += (size_t)(K + 1) << 6;
                const double2 b1 = *reinterpret_cast<const double2*>(brow);
                brow += (size_t)(K + 2) << 6;
                double2 c0 = *reinterpret_cast<double2*>(crow), c1 = *reinterpret_cast<double2*>(crow + 64);
                dmma884(c0.x, c0.y, a.x, b0.x);
                dmma884(c1.x, c1.y, a.x, b1.x);
                dmma884(c0.x, c0.y, a.y, b0.y);
                dmma884(c1.x, c1.y, a.y, b1.y);
                *reinterpret_cast<double2*>(crow) = c0;
                *reinterpret_cast<double2*>(crow + 64) = c1;
            }
            if (K <= Kmax) {
                const double2 b0 = *reinterpret_cast<const double2*>(brow);
                double2 c0 = *reinterpret_cast<double2*>(crow);
                tile_mma(c0, a, b0);
                *reinterpret_cast<double2*>(crow) = c0;
            }
        }
        if (J == 1) PHASE_CLK(RANK ? 54 : 46);
        __syncthreads();
        if (J == 1) PHASE_CLK(RANK ? 55 : 47);
        if (RANK && pub->stop) break;
    }
    __syncthreads();
}


}  // namespace rvio
