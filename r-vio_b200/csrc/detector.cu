// detector.cu -- FeatureDetector::DetectWithSubPix on the device (reference src/rvio/FeatureDetector.cc:55-75:
// cv::goodFeaturesToTrack + cv::cornerSubPix on the equalised frame, called from Tracker.cc:207,350).
// Compiled for sm_100a with -fmad=false; the arithmetic contract is the one written down in oracle/detector.c (the CPU
// restatement this path is checked against bit for bit; that file is pinned against cv2).
//   k_det_eig     min-eigenvalue map (Sobel 3 + 3x3 box, float32 / double sums) + global maximum          W*H/1024 CTAs
//   k_det_nms     threshold (float)(max * quality), 3x3 local maxima -> packed keys (value bits | y | x)    W*H/1024 CTAs
//   k_det_select  strongest-first blocks: histogram, gather, bitonic sort, greedy minimum-distance selection      1 CTA
//   k_det_subpix  cornerSubPix, one warp per corner                                                         n/4 CTAs
#include "common.cuh"
#include "tracker_kernels.cuh"
#include "detector_kernels.cuh"

#include <math.h>

namespace rvio {

__device__ __forceinline__ int d_refl101(int i, int n)
{
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// ------------------------------------------------------------------------------------------------ min-eigenvalue map
__global__ void __launch_bounds__(1024) k_det_eig(DetParams P)
{
    __shared__ float s_xx[34][35], s_xy[34][35], s_yy[34][35];
    __shared__ unsigned s_max;
    const int W = P.img.w, H = P.img.h;
    const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const float k1 = (float)(1.0 / (255.0 * 4 * 3)), k0 = __fmul_rn(2.f, k1);
    if (tid == 0) s_max = 0u;
    // covariance products on the 34 x 34 halo tile; a position outside the image takes the products OF its reflected
    // position (cv::boxFilter reflects the product images, not the pixels)
    for (int o = tid; o < 34 * 34; o += 1024) {
        const int ty = o / 34, tx = o - ty * 34;
        const int y = d_refl101(y0 + ty - 1, H), x = d_refl101(x0 + tx - 1, W);
        float xx = 0.f, xy = 0.f, yy = 0.f;
        if (y < H && x < W && y >= 0 && x >= 0) {
            const uint8_t* c = P.img.base + (ptrdiff_t)y * P.img.pitch + x;        // the level keeps a reflect-101 border
            float r[3], q[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint8_t* rp = c + (ptrdiff_t)(k - 1) * P.img.pitch;
                const float pm = (float)rp[-1], pc = (float)rp[0], pp = (float)rp[1];
                r[k] = __fsub_rn(pp, pm);
                q[k] = __fmaf_rn(pp, k1, __fmaf_rn(pc, k0, __fmul_rn(pm, k1)));
            }
            const float dx = __fmaf_rn(__fadd_rn(r[0], r[2]), k1, __fmul_rn(r[1], k0));
            const float dy = __fsub_rn(q[2], q[0]);
            xx = __fmul_rn(dx, dx); xy = __fmul_rn(dx, dy); yy = __fmul_rn(dy, dy);
        }
        s_xx[ty][tx] = xx; s_xy[ty][tx] = xy; s_yy[ty][tx] = yy;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    float e = 0.f;
    if (x < W && y < H) {
        double a = 0, b = 0, c = 0;
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                a += (double)s_xx[threadIdx.y + j][threadIdx.x + i];
                b += (double)s_xy[threadIdx.y + j][threadIdx.x + i];
                c += (double)s_yy[threadIdx.y + j][threadIdx.x + i];
            }
        const float fa = __fmul_rn((float)a, 0.5f), fb = (float)b, fc = __fmul_rn((float)c, 0.5f);
        const float t = __fsub_rn(fa, fc);
        e = __fsub_rn(__fadd_rn(fa, fc), sqrtf(__fadd_rn(__fmul_rn(t, t), __fmul_rn(fb, fb))));
        P.eig[(size_t)y * W + x] = e;
    }
    // maximum (the map is >= 0 up to rounding; negative values never win against the 0 start, like cv::minMaxLoc on a
    // map that contains a positive value)
    unsigned bits = (e > 0.f) ? __float_as_uint(e) : 0u;
    bits = __reduce_max_sync(0xffffffffu, bits);
    if ((tid & 31) == 0 && bits) atomicMax(&s_max, bits);
    __syncthreads();
    if (tid == 0 && s_max) atomicMax(&P.ctrl->max_bits, s_max);
}

// ------------------------------------------------------------------------------------------------ threshold + NMS
__global__ void __launch_bounds__(1024) k_det_nms(DetParams P)
{
    const int W = P.img.w, H = P.img.h;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
    if (x < 1 || y < 1 || x > W - 2 || y > H - 2) return;
    const float thr = (float)((double)__uint_as_float(P.ctrl->max_bits) * P.quality);
    const float* e = P.eig + (size_t)y * W + x;
    const float v = e[0];
    if (!(v > thr)) return;
    bool is_max = true;
#pragma unroll
    for (int j = -1; j <= 1; ++j)
#pragma unroll
        for (int i = -1; i <= 1; ++i) is_max = is_max && !(e[(ptrdiff_t)j * W + i] > v);
    if (!is_max) return;
    const int pos = atomicAdd(&P.ctrl->n_cand, 1);
    if (pos < P.key_cap) P.keys[pos] = ((unsigned long long)__float_as_uint(v) << 32) | ((unsigned)y << 16) | (unsigned)x;
    else P.ctrl->overflow = 1;
}

// ------------------------------------------------------------------------------------------------ sort + greedy selection
// cv::goodFeaturesToTrack walks the local maxima in descending (value, address) order and keeps a corner when no kept corner
// lies closer than the minimum distance, until max_corners are kept.  Only a prefix of that order is ever looked at, so
// the candidates are taken block by block from the top of a histogram over the leading float bits: gather <= 4096 keys,
// bitonic sort (registers / shuffles / shared memory by partner distance), then the block is settled 256 candidates at a
// time: all threads build, for every candidate of the batch, the bit mask of the STRONGER batch members within the minimum
// distance; one warp resolves the masks in rounds (rejected once a stronger neighbour is kept, kept once every stronger
// neighbour is decided) -- the greedy order's outcome without its one-candidate-at-a-time chain; the rest of the block is
// then tested against the kept corners' grid by all threads and compacted in rank order.
// (measured on B200, 18.8 k candidates, 200 corners: the previous shared-memory bitonic sort took 52 us and the warp-serial
//  settling 85 us of the kernel's 148 us.)

// Shadow bitmap: bit (x, y) is set when the pixel lies closer than the minimum distance to a corner kept so far -- the test a
// candidate has to pass is then one bit.  A kept corner paints its disc { dx^2 + dy^2 < d^2 } (one warp per corner, one row
// per lane, word-wise atomicOr).
__device__ __forceinline__ bool det_shadowed(const unsigned* sb, int wpr, int x, int y)
{
    return (sb[y * wpr + (x >> 5)] >> (x & 31)) & 1u;
}
__device__ __forceinline__ void det_paint(unsigned* sb, int wpr, int W, int H, int kx, int ky, double md2, int R, int lane)
{
    for (int r = lane; r <= 2 * R; r += 32) {
        const int dy = r - R, y = ky + dy;
        if (y < 0 || y >= H) continue;
        const double rem = md2 - (double)(dy * dy);
        if (!(rem > 0.0)) continue;
        int dxm = (int)sqrt(rem);
        while ((double)((dxm + 1) * (dxm + 1) + dy * dy) < md2) ++dxm;
        while (dxm >= 0 && !((double)(dxm * dxm + dy * dy) < md2)) --dxm;
        if (dxm < 0) continue;
        const int x0 = max(kx - dxm, 0), x1 = min(kx + dxm, W - 1);
        for (int w = x0 >> 5; w <= (x1 >> 5); ++w) {
            const int lo = max(x0 - 32 * w, 0), hi = min(x1 - 32 * w, 31);
            const unsigned m = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
            atomicOr(&sb[y * wpr + w], m);
        }
    }
}

#ifdef RVIO_B200_PHASE_CLOCKS
// (profiling build only) clock stamps / counters of the selection kernel: [0] start, [1] after the first histogram, [2..] per
// block: gathered, sorted, settled; [30] candidates, [31] blocks, [32] settle iterations, [33] corners kept
__device__ long long g_det_clk[64];
#define DET_CLK(k) do { if (threadIdx.x == 0 && (k) < 64) g_det_clk[k] = clock64(); } while (0)
#define DET_VAL(k, v) do { if (threadIdx.x == 0) g_det_clk[k] = (v); } while (0)
#else
#define DET_CLK(k) do { } while (0)
#define DET_VAL(k, v) do { } while (0)
#endif
constexpr int kDetBlock = 4096, kDetBins = 1024;
constexpr int kDetWarpBatch = 8;          // sub-batches of 32 the settling warp takes per iteration

__global__ void __launch_bounds__(1024) k_det_select(DetParams P)
{
    extern __shared__ __align__(16) unsigned char dsm[];
    __shared__ int s_hist[kDetBins];
    __shared__ int s_lo, s_take, s_count, s_nout, s_stop;
    __shared__ int s_warp[32];
    __shared__ unsigned s_M[32 * kDetWarpBatch][kDetWarpBatch + 1], s_A[kDetWarpBatch], s_U[kDetWarpBatch];      // (+1: rows of consecutive lanes in different banks)
    __shared__ unsigned s_new[32 * kDetWarpBatch];         // corners kept in the current batch (y << 16 | x)
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(dsm);
    const int tid = threadIdx.x, lane = tid & 31;
    int nc = P.ctrl->n_cand;
    if (nc > P.key_cap) nc = P.key_cap;
    const unsigned max_bits = P.ctrl->max_bits;
    const unsigned thr_bits = __float_as_uint((float)((double)__uint_as_float(max_bits) * P.quality));
    int shift = 17;
    while ((int)((max_bits >> shift) - (thr_bits >> shift)) + 1 > kDetBins) ++shift;
    const unsigned base = thr_bits >> shift;
    const int nbins = (int)((max_bits >> shift) - base) + 1;
    for (int i = tid; i < kDetBins; i += 1024) s_hist[i] = 0;
    const int IW = P.img.w, IH = P.img.h, wpr = P.wpr;
    unsigned* sb = P.shadow_in_smem ? reinterpret_cast<unsigned*>(dsm + (size_t)kDetBlock * 16) : P.shadow;
    const int paintR = (int)ceil(P.min_dist);
    unsigned long long* keys2 = reinterpret_cast<unsigned long long*>(dsm + (size_t)kDetBlock * 8);     // second buffer of the compaction
    for (int i = tid; i < wpr * IH; i += 1024) sb[i] = 0u;
    if (tid == 0) { s_nout = 0; s_stop = 0; }
    __syncthreads();
    const double md2 = P.min_dist * P.min_dist;
    bool first_round = true;
    int dbg_blocks = 0, dbg_iters = 0;
    DET_CLK(0); DET_VAL(30, nc);
    while (true) {
        // candidates that lie within the minimum distance of a corner kept in an earlier round can never be kept: drop them
        // (all 1024 threads), so that the serial part below only sees candidates that still have a chance
        for (int i = tid; i < kDetBins; i += 1024) s_hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < nc; i += 1024) {
            unsigned long long kk = P.keys[i];
            if (kk == 0ull) continue;
            if (!first_round) {
                const unsigned lo32 = (unsigned)kk;
                const int y = (int)(lo32 >> 16), x = (int)(lo32 & 0xffffu);
                const bool alive = !det_shadowed(sb, wpr, x, y);
                if (!alive) { P.keys[i] = 0ull; continue; }
            }
            atomicAdd(&s_hist[(int)(((unsigned)(kk >> 32) >> shift) - base)], 1);
        }
        first_round = false;
        __syncthreads();
        DET_CLK(1 + 4 * dbg_blocks);
        {
            // next block of bins from the top, at most kDetBlock candidates: suffix sums of the histogram (thread t = bin t; warp
            // shuffles + one pass over the 32 warp totals), lo = first bin whose suffix still fits, instead of a serial walk
            const int h = (tid < nbins) ? s_hist[tid] : 0;
            int suf = h;                                      // inclusive suffix sum inside the warp
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t2 = __shfl_down_sync(0xffffffffu, suf, o); if (lane + o < 32) suf += t2; }
            if (lane == 0) s_warp[tid >> 5] = suf;
            __syncthreads();
            int above = 0;                                    // everything in the warps above mine
            for (int w = (tid >> 5) + 1; w < 32; ++w) above += s_warp[w];
            suf += above;                                     // = number of candidates in bins >= tid
            const unsigned fits = __ballot_sync(0xffffffffu, tid < nbins && suf <= kDetBlock);
            const unsigned nz = __ballot_sync(0xffffffffu, h > 0);
            __syncthreads();
            unsigned* scratch = &s_M[0][0];
            if (lane == 0) { s_warp[tid >> 5] = __popc(fits); scratch[tid >> 5] = nz; }
            __syncthreads();
            if (tid == 0) {
                int nfit = 0, hi = 0;
                for (int w = 0; w < 32; ++w) { nfit += s_warp[w]; if (scratch[w]) hi = 32 * w + (32 - __clz(scratch[w])); }
                const int lo = nbins - nfit;                  // suffix sums are non-increasing in the bin index: the fitting bins are the top nfit ones
                s_lo = lo; s_count = 0;
                s_take = -1;                                  // filled in below by the thread that owns bin lo
                if (hi > 0 && lo >= hi) { P.ctrl->overflow = 1; s_stop = 1; }    // one bin alone exceeds the block
                if (hi == 0) s_stop = 1;                      // nothing left
            }
            __syncthreads();
            if (tid == s_lo && tid < nbins) s_take = suf;
            if (tid == 0 && s_lo >= nbins) s_take = 0;
        }
        __syncthreads();
        if (s_stop) break;
        const int lo = s_lo, take = s_take;
        for (int i0 = 0; i0 < nc; i0 += 4 * 1024) {             // four keys per thread in flight; one shared-memory atomic per warp and key slot
            unsigned long long kk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int i = i0 + q * 1024 + tid; kk[q] = (i < nc) ? P.keys[i] : 0ull; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int i = i0 + q * 1024 + tid;
                bool hit = false;
                if (kk[q] != 0ull) hit = (int)(((unsigned)(kk[q] >> 32) >> shift) - base) >= lo;
                const unsigned hm = __ballot_sync(0xffffffffu, hit);
                int wbase = 0;
                if (lane == 0 && hm) wbase = atomicAdd(&s_count, __popc(hm));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (hit) { keys[wbase + __popc(hm & ((1u << lane) - 1u))] = kk[q]; P.keys[i] = 0ull; }      // consumed by this round
            }
        }
        __syncthreads();
        for (int i = take + tid; i < kDetBlock; i += 1024) keys[i] = 0ull;
        __syncthreads();
        DET_CLK(2 + 4 * dbg_blocks); DET_VAL(34 + dbg_blocks, take);
        // bitonic sort of the 4096 slots, descending.  Thread t holds slots 4t .. 4t+3 in registers: partner distances 1, 2 stay
        // inside the thread, 4 .. 64 are warp shuffles, only 128 .. 2048 (15 of the 78 passes) go through shared memory
        {
            unsigned long long r[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) r[q] = keys[4 * tid + q];
            for (int k = 2; k <= kDetBlock; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    if (j >= 128) {
                        __syncthreads();
#pragma unroll
                        for (int q = 0; q < 4; ++q) keys[4 * tid + q] = r[q];
                        __syncthreads();
                        const int pt = tid ^ (j >> 2);
                        const bool lower = (tid & (j >> 2)) == 0;
                        const bool desc = ((4 * tid) & k) == 0;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const unsigned long long pq = keys[4 * pt + q];
                            const bool take_max = (lower == desc);
                            r[q] = take_max ? (r[q] > pq ? r[q] : pq) : (r[q] < pq ? r[q] : pq);
                        }
                    } else if (j >= 4) {
                        const bool lower = (tid & (j >> 2)) == 0;
                        const bool desc = ((4 * tid) & k) == 0;
                        const bool take_max = (lower == desc);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const unsigned long long pq = __shfl_xor_sync(0xffffffffu, r[q], j >> 2);
                            r[q] = take_max ? (r[q] > pq ? r[q] : pq) : (r[q] < pq ? r[q] : pq);
                        }
                    } else {
                        // (static register indices: pairs (0,2),(1,3) for distance 2, (0,1),(2,3) for distance 1)
                        auto cx = [&](unsigned long long& a, unsigned long long& b, int qlow) {
                            const bool desc = ((4 * tid + qlow) & k) == 0;
                            const bool sw = desc ? (a < b) : (a > b);
                            const unsigned long long t0 = sw ? b : a, t1 = sw ? a : b;
                            a = t0; b = t1;
                        };
                        if (j == 2) { cx(r[0], r[2], 0); cx(r[1], r[3], 1); }
                        else { cx(r[0], r[1], 0); cx(r[2], r[3], 2); }
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < 4; ++q) keys[4 * tid + q] = r[q];
            __syncthreads();
        }
        DET_CLK(3 + 4 * dbg_blocks);
        // ---- greedy minimum-distance selection of the sorted block: resolve / filter / compact iterations.
        //      (1) one warp settles the 32 strongest candidates still alive, in rank order (they are known to be clear of
        //          every corner kept so far); (2) ALL threads test the remaining candidates against the corners kept so far
        //          (3 x 3 cells of the minimum-distance grid) -- a candidate within the distance of a kept, stronger corner can
        //          never be kept; (3) the survivors are compacted, in rank order, to the front (block scan).  Every iteration
        //          retires the 32 strongest plus everything they shadow, so the serial warp only ever sees genuine contenders.
        {
            unsigned long long* cur = keys;
            unsigned long long* alt = keys2;
            int n_alive = take;
            int n_out = s_nout;
            bool block_first = true;
            while (n_alive > 0 && n_out < P.max_corners) {
                ++dbg_iters;
                if (block_first) {
                    // the sorted block has not been tested against the corners kept in EARLIER blocks' later iterations: the filter
                    // pass above did that before the gather, so the front of the block is clear
                    block_first = false;
                }
                // (1) the kDetWarpBatch * 32 strongest candidates still alive (all clear of every corner kept so far): for each
                //     of them the set of STRONGER batch members closer than the minimum distance, as a 256-bit mask (all threads)
                if (dbg_blocks == 0 && dbg_iters == 1) DET_CLK(40);
                const int B = min(n_alive, 32 * kDetWarpBatch);
                for (int item = tid; item < B * kDetWarpBatch; item += 1024) {
                    const int i = item / kDetWarpBatch, w = item - i * kDetWarpBatch;
                    unsigned m = 0;
                    if (32 * w < i) {
                        const unsigned pi = (unsigned)cur[i];
                        const int xi = (int)(pi & 0xffffu), yi = (int)(pi >> 16);
                        const int jmax = min(32, i - 32 * w);
                        for (int b2 = 0; b2 < jmax; ++b2) {
                            const unsigned pj = (unsigned)cur[32 * w + b2];
                            const int dx = xi - (int)(pj & 0xffffu), dy = yi - (int)(pj >> 16);
                            if ((double)(dx * dx + dy * dy) < md2) m |= 1u << b2;
                        }
                    }
                    s_M[i][w] = m;
                }
                if (tid < kDetWarpBatch) {
                    s_A[tid] = 0u;
                    s_U[tid] = (32 * tid + 32 <= B) ? 0xffffffffu : (32 * tid < B ? ((1u << (B - 32 * tid)) - 1u) : 0u);
                }
                __syncthreads();
                if (dbg_blocks == 0 && dbg_iters == 1) DET_CLK(41);
                if (tid < 32) {
                    // (2) a candidate is kept iff no stronger kept candidate lies within the distance.  The batch is in rank order,
                    //     so word k (32 candidates) depends only on the words before it, which are final, and on itself: first the
                    //     candidates shadowed by a kept member of an earlier word drop out, then the word settles in rounds held in
                    //     registers (rejected as soon as a stronger neighbour is kept, kept as soon as every stronger neighbour
                    //     inside the word is decided and none is kept).
                    for (int k = 0; k < kDetWarpBatch && 32 * k < B; ++k) {
                        const int i = 32 * k + lane;
                        bool dead = !((s_U[k] >> lane) & 1u);                 // (beyond the batch)
                        for (int w = 0; w < k; ++w) dead = dead || (s_M[i][w] & s_A[w]);
                        const unsigned mk = s_M[i][k];
                        unsigned U = __ballot_sync(0xffffffffu, !dead), A = 0u;
                        while (U) {
                            const bool und = (U >> lane) & 1u;
                            const bool rej = und && (mk & A);
                            const bool acc = und && !rej && !(mk & U);
                            const unsigned am = __ballot_sync(0xffffffffu, acc), rm = __ballot_sync(0xffffffffu, rej);
                            A |= am; U &= ~(am | rm);
                        }
                        if (lane == 0) { s_A[k] = A; s_U[k] = 0u; }
                        __syncwarp();
                    }
                    // (3) the kept ones in rank order: output, and their coordinates for the painters
                    int no = n_out;
                    for (int k = 0; k < kDetWarpBatch && 32 * k < B; ++k) {
                        const unsigned am = s_A[k];
                        const int pos = no + __popc(am & ((1u << lane) - 1u));
                        if (((am >> lane) & 1u) && pos < P.max_corners) {
                            const unsigned lo32 = (unsigned)cur[32 * k + lane];
                            P.out[pos] = make_float2((float)(lo32 & 0xffffu), (float)(lo32 >> 16));
                            s_new[pos - n_out] = lo32;
                        }
                        no += __popc(am);
                    }
                    if (no > P.max_corners) no = P.max_corners;
                    if (lane == 0) s_nout = no;
                }
                __syncthreads();
                {
                    // (3b) every corner kept in this batch paints its disc into the shadow bitmap: one warp per corner
                    const int n_new = s_nout - n_out;
                    for (int c = (tid >> 5); c < n_new; c += 32) {
                        const unsigned lo32 = s_new[c];
                        det_paint(sb, wpr, IW, IH, (int)(lo32 & 0xffffu), (int)(lo32 >> 16), md2, paintR, lane);
                    }
                }
                __syncthreads();
                if (dbg_blocks == 0 && dbg_iters == 1) DET_CLK(43);
                n_out = s_nout;
                if (n_out >= P.max_corners) break;
                // (2) + (3): candidates 32.. against the grid, survivors compacted in rank order
                const int rest = n_alive - 32 * kDetWarpBatch;
                if (rest <= 0) { n_alive = 0; break; }
                const int per = (rest + 1023) / 1024;
                const int k0 = 32 * kDetWarpBatch + tid * per;
                unsigned long long mine[4];
                int local = 0;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int k = k0 + u;
                    unsigned long long kk = 0ull;
                    if (u < per && k < n_alive) {
                        kk = cur[k];
                        const unsigned lo32 = (unsigned)kk;
                        const int y = (int)(lo32 >> 16), x = (int)(lo32 & 0xffffu);
                        if (det_shadowed(sb, wpr, x, y)) kk = 0ull;
                    }
                    mine[u] = kk;
                    local += kk != 0ull;
                }
                int incl = local;
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                if (lane == 31) s_warp[tid >> 5] = incl;
                __syncthreads();
                if (tid < 32) {
                    int w = s_warp[tid];
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
                    s_warp[tid] = w;
                }
                __syncthreads();
                int pos = incl - local + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0);
#pragma unroll
                for (int u = 0; u < 4; ++u) if (mine[u] != 0ull) alt[pos++] = mine[u];
                n_alive = s_warp[31];
                __syncthreads();
                if (dbg_blocks == 0 && dbg_iters == 1) { DET_CLK(44); DET_VAL(45, n_alive); }
                unsigned long long* t2 = cur; cur = alt; alt = t2;
            }
            if (tid == 0 && s_nout >= P.max_corners) s_stop = 1;
        }
        __syncthreads();
        DET_CLK(4 + 4 * dbg_blocks);
        ++dbg_blocks;
        if (s_stop) break;
    }
    DET_CLK(29); DET_VAL(31, dbg_blocks); DET_VAL(32, dbg_iters); DET_VAL(33, s_nout);
    if (tid == 0) P.ctrl->n_out = s_nout;
}

// ------------------------------------------------------------------------------------------------ cornerSubPix
__global__ void __launch_bounds__(128) k_det_subpix(DetParams P)
{
    __shared__ float s_patch[4][33 * 33];
    __shared__ unsigned char s_px[4][34 * 36];                       // large windows: (pw + 1)^2 source pixels of the current patch; small: a 32 x 32 tile
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int k = blockIdx.x * 4 + wib;
    if (k >= P.ctrl->n_out) return;
    const int W = P.img.w, H = P.img.h;
    const int hw = P.hw, ww = 2 * hw + 1, pw = ww + 2, sw = pw + 1;
    float* patch = s_patch[wib];
    unsigned char* px = s_px[wib];
    const float2 c0 = P.out[k];
    float cx = c0.x, cy = c0.y;
    int iter = 0;
    double err = 0;
    const double eps = P.subpix_eps * P.subpix_eps;
    int tx0 = 0, ty0 = 0;
    bool have_tile = false;
    // per-lane element tables, fixed for the whole refinement (window sizes up to 15 x 15 take the register path)
    constexpr int kPatPer = 10, kWinPer = 8;                          // ceil(17*17/32), ceil(15*15/32)
    const bool small = hw <= 7;
    short pat_o[kPatPer], win_o[kWinPer];
    signed char win_x[kWinPer], win_y[kWinPer];
    float win_m[kWinPer];
    if (small) {
#pragma unroll
        for (int u = 0; u < kPatPer; ++u) { const int o = u * 32 + lane; pat_o[u] = (short)((o / pw) * 32 + (o - (o / pw) * pw)); }      // (row stride of the staged tile)
#pragma unroll
        for (int u = 0; u < kWinPer; ++u) {
            const int q = u * 32 + lane, i = q / ww, j = q - i * ww;
            win_o[u] = (short)(i * pw + j); win_x[u] = (signed char)(j - hw); win_y[u] = (signed char)(i - hw);
            win_m[u] = (q < ww * ww) ? P.mask[q] : 0.f;
        }
    }
    do {
        // getRectSubPix: (ww+2)^2 float patch centred at (cx, cy), bilinear, replicated border
        const float ox = __fsub_rn(cx, __fmul_rn((float)(pw - 1), 0.5f)), oy = __fsub_rn(cy, __fmul_rn((float)(pw - 1), 0.5f));
        const int ix = (int)floorf(ox), iy = (int)floorf(oy);
        const float a = __fsub_rn(ox, (float)ix), b = __fsub_rn(oy, (float)iy);
        const float a11 = __fmul_rn(__fsub_rn(1.f, a), __fsub_rn(1.f, b)), a12 = __fmul_rn(a, __fsub_rn(1.f, b));
        const float a21 = __fmul_rn(__fsub_rn(1.f, a), b), a22 = __fmul_rn(a, b);
        double sa = 0, sb = 0, sc = 0, s1 = 0, s2 = 0;
        if (small) {
            // the (pw+1)^2 source pixels come from a 32 x 32 tile of the image (coordinates clamped to the image) staged in
            // shared memory: the refinement moves the window by fractions of a pixel, so the tile is fetched once and only
            // re-fetched if the window leaves it (one global round trip per corner instead of one per iteration)
            if (!have_tile || ix < tx0 || iy < ty0 || ix + sw > tx0 + 32 || iy + sw > ty0 + 32) {
                tx0 = ix - (32 - sw) / 2; ty0 = iy - (32 - sw) / 2;
                unsigned char v[32];
                const int yy = min(max(ty0 + lane, 0), H - 1);
                const unsigned char* row = P.img.base + (ptrdiff_t)yy * P.img.pitch;
#pragma unroll
                for (int u = 0; u < 32; ++u) v[u] = row[min(max(tx0 + u, 0), W - 1)];
                __syncwarp();
#pragma unroll
                for (int u = 0; u < 32; u += 4)
                    *reinterpret_cast<unsigned*>(px + lane * 32 + u) = (unsigned)v[u] | ((unsigned)v[u + 1] << 8) | ((unsigned)v[u + 2] << 16) | ((unsigned)v[u + 3] << 24);
                have_tile = true;
                __syncwarp();
            }
            const unsigned char* pxo = px + (iy - ty0) * 32 + (ix - tx0);
#pragma unroll
            for (int u = 0; u < kPatPer; ++u) {
                const int o = u * 32 + lane;
                if (o < pw * pw) {
                    const unsigned char* q = pxo + pat_o[u];
                    float t = __fmul_rn((float)q[0], a11);
                    t = __fadd_rn(t, __fmul_rn((float)q[1], a12));
                    t = __fadd_rn(t, __fmul_rn((float)q[32], a21));
                    t = __fadd_rn(t, __fmul_rn((float)q[33], a22));
                    patch[o] = t;
                }
            }
            __syncwarp();
#pragma unroll
            for (int u = 0; u < kWinPer; ++u) {
                if (u * 32 + lane < ww * ww) {
                    const float* pp = patch + win_o[u];
                    const double m = (double)win_m[u];
                    const double tgx = (double)__fsub_rn(pp[pw + 2], pp[pw]);
                    const double tgy = (double)__fsub_rn(pp[2 * pw + 1], pp[1]);
                    const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                    const double px_ = (double)win_x[u], py_ = (double)win_y[u];
                    sa += gxx; sb += gxy; sc += gyy;
                    s1 += gxx * px_ + gxy * py_;
                    s2 += gxy * px_ + gyy * py_;
                }
            }
        } else {
            for (int o = lane; o < sw * sw; o += 32) {
                const int i = o / sw, j = o - i * sw;
                const int yy = min(max(iy + i, 0), H - 1), xx = min(max(ix + j, 0), W - 1);
                px[o] = P.img.base[(ptrdiff_t)yy * P.img.pitch + xx];
            }
            __syncwarp();
            for (int o = lane; o < pw * pw; o += 32) {
                const int i = o / pw, j = o - i * pw;
                const unsigned char* q = px + i * sw + j;
                float t = __fmul_rn((float)q[0], a11);
                t = __fadd_rn(t, __fmul_rn((float)q[1], a12));
                t = __fadd_rn(t, __fmul_rn((float)q[sw], a21));
                t = __fadd_rn(t, __fmul_rn((float)q[sw + 1], a22));
                patch[o] = t;
            }
            __syncwarp();
            for (int q = lane; q < ww * ww; q += 32) {
                const int i = q / ww, j = q - i * ww;
                const double m = (double)P.mask[q];
                const double tgx = (double)__fsub_rn(patch[(i + 1) * pw + j + 2], patch[(i + 1) * pw + j]);
                const double tgy = (double)__fsub_rn(patch[(i + 2) * pw + j + 1], patch[i * pw + j + 1]);
                const double gxx = tgx * tgx * m, gxy = tgx * tgy * m, gyy = tgy * tgy * m;
                const double px_ = (double)(j - hw), py_ = (double)(i - hw);
                sa += gxx; sb += gxy; sc += gyy;
                s1 += gxx * px_ + gxy * py_;
                s2 += gxy * px_ + gyy * py_;
            }
        }
#pragma unroll
        for (int msk = 16; msk >= 1; msk >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, msk); sb += __shfl_xor_sync(0xffffffffu, sb, msk);
            sc += __shfl_xor_sync(0xffffffffu, sc, msk); s1 += __shfl_xor_sync(0xffffffffu, s1, msk);
            s2 += __shfl_xor_sync(0xffffffffu, s2, msk);
        }
        __syncwarp();
        const double det = sa * sc - sb * sb;
        if (fabs(det) <= 2.220446049250313e-16 * 2.220446049250313e-16) break;
        const double scale = 1.0 / det;
        const float nx = (float)((double)cx + sc * scale * s1 - sb * scale * s2);
        const float ny = (float)((double)cy - sb * scale * s1 + sa * scale * s2);
        const float ex = __fsub_rn(nx, cx), ey = __fsub_rn(ny, cy);
        err = (double)__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
        cx = nx; cy = ny;
        if (cx < 0 || cx >= (float)W || cy < 0 || cy >= (float)H) break;
    } while (++iter < P.subpix_iters && err > eps);
    if (fabsf(__fsub_rn(cx, c0.x)) > (float)hw || fabsf(__fsub_rn(cy, c0.y)) > (float)hw) { cx = c0.x; cy = c0.y; }
    if (lane == 0) P.out[k] = make_float2(cx, cy);
}

// ------------------------------------------------------------------------------------------------ host
int detector_create(Detector* D, int W, int H, int max_corners)
{
    memset(D, 0, sizeof *D);
    D->W = W; D->H = H; D->max_corners = max_corners;
    RVIO_CUDA_TRY(cudaMalloc((void**)&D->eig, sizeof(float) * (size_t)W * H));
    RVIO_CUDA_TRY(cudaMalloc((void**)&D->keys, sizeof(unsigned long long) * kDetKeyCap));
    RVIO_CUDA_TRY(cudaMalloc((void**)&D->ctrl, sizeof(DetCtrl)));
    RVIO_CUDA_TRY(cudaMalloc((void**)&D->out, sizeof(float2) * ((size_t)max_corners + 1)));
    RVIO_CUDA_TRY(cudaMalloc((void**)&D->mask, sizeof(float) * 31 * 31));
    RVIO_CUDA_TRY(cudaMalloc((void**)&D->shadow, sizeof(unsigned) * (size_t)((W + 31) / 32) * H));
    RVIO_CUDA_TRY(cudaMemset(D->ctrl, 0, sizeof(DetCtrl)));
    RVIO_CUDA_TRY(cudaMallocHost((void**)&D->h_mask, sizeof(float) * 31 * 31));
    RVIO_CUDA_TRY(cudaFuncSetAttribute(k_det_select, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    D->mask_hw = -1;
    return RVIO_OK;
}

void detector_destroy(Detector* D)
{
    cudaFree(D->eig); cudaFree(D->keys); cudaFree(D->ctrl); cudaFree(D->out); cudaFree(D->mask); cudaFree(D->shadow);
    cudaFreeHost(D->h_mask);
}

// Enqueues DetectWithSubPix(level0, max_corners, s) on `st`; the corners land in D->out, their number in D->ctrl->n_out.
int detector_enqueue(Detector* D, cudaStream_t st, const PyrLevel& level0, int s, float min_dist_f, float quality_f)
{
    DetParams P;
    P.img = level0; P.eig = D->eig; P.ctrl = D->ctrl; P.keys = D->keys; P.key_cap = kDetKeyCap;
    P.out = D->out; P.max_corners = D->max_corners;
    P.quality = (double)quality_f;
    P.min_dist = (double)((float)s * min_dist_f);
    P.hw = (int)floor(.5 * min_dist_f);                               // FeatureDetector.cc:68
    if (P.hw < 1 || P.hw > 15 || P.min_dist < 1) { set_error("detector_enqueue", "Tracker.nMinDist outside the supported range [2, 31]"); return RVIO_ERR_ARG; }
    P.subpix_iters = 30; P.subpix_eps = 1e-2;                        // FeatureDetector.cc:70
    P.cell = P.gw = P.gh = 0;
    P.wpr = (D->W + 31) / 32;
    const size_t bitmap = sizeof(unsigned) * (size_t)P.wpr * D->H;
    P.shadow_in_smem = ((size_t)kDetBlock * 16 + bitmap + 64 <= 200 * 1024) ? 1 : 0;      // 752x480: 45 KB, 1280x720: 115 KB; larger: the global copy
    P.shadow = D->shadow;
    const size_t smem = (size_t)kDetBlock * 16 + (P.shadow_in_smem ? bitmap : 0) + 64;
    if (smem > 200 * 1024) { set_error("detector_enqueue", "minimum-distance grid does not fit in shared memory (Tracker.nMinDist too small for this image size)"); return RVIO_ERR_CAPACITY; }
    if (D->mask_hw != P.hw) {                                        // cornerSubPix window weights (host libm, like OpenCV)
        const int ww = 2 * P.hw + 1;
        float mx[31];
        for (int j = 0; j < ww; ++j) { const float x = (float)(j - P.hw) / P.hw; mx[j] = (float)exp(-x * x); }
        for (int i = 0; i < ww; ++i) {
            const float y = (float)(i - P.hw) / P.hw;
            const float vy = (float)exp(-y * y);
            for (int j = 0; j < ww; ++j) D->h_mask[i * ww + j] = (float)(vy * mx[j]);
        }
        RVIO_CUDA_TRY(cudaMemcpy(D->mask, D->h_mask, sizeof(float) * ww * ww, cudaMemcpyHostToDevice));   // once per window size
        D->mask_hw = P.hw;
    }
    P.mask = D->mask;
    RVIO_ENQ(cudaMemsetAsync(D->ctrl, 0, sizeof(DetCtrl), st));
    const dim3 grid(div_up(D->W, 32), div_up(D->H, 32)), blk(32, 32);
    RVIO_LAUNCH(k_det_eig, grid, blk, 0, st, P);
    RVIO_LAUNCH(k_det_nms, grid, blk, 0, st, P);
    RVIO_LAUNCH(k_det_select, 1, 1024, smem, st, P);
    RVIO_LAUNCH(k_det_subpix, div_up(D->max_corners, 4), 128, 0, st, P);
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

}  // namespace rvio

#ifdef RVIO_B200_PHASE_CLOCKS
extern "C" int rvio_b200_det_clocks(long long* out, int n)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, rvio::g_det_clk, sizeof(long long) * (n < 64 ? n : 64));
}
#endif
