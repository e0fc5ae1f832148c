// tracker.cu -- CUDA kernels + C ABI of the tracker half of the hot path (see tracker_kernels.cuh).
// Compiled for sm_100a with -fmad=false (bit-exact float32/float64 contract).
#include <cuda.h>
#include "common.cuh"
#include "tracker_kernels.cuh"
#include "detector_kernels.cuh"
#include "device_utils.cuh"

#include <math.h>
#include <new>
#include <vector>

namespace rvio {

// ================================================================================================
// small device helpers
// ================================================================================================
__device__ __forceinline__ int reflect101(int i, int n)
{
    // single reflection is enough for |overshoot| < n (callers guarantee n > kBorder)
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// Writes pixel (x,y) and every border position that BORDER_REFLECT_101 maps onto it.
__device__ __forceinline__ void store_with_border(const PyrLevel& L, int x, int y, uint8_t v)
{
    int xs[3], ys[3], nx = 0, ny = 0;
    xs[nx++] = x;
    if (x >= 1 && x <= kBorder) xs[nx++] = -x;
    if (x <= L.w - 2 && x >= L.w - 1 - kBorder) xs[nx++] = 2 * (L.w - 1) - x;
    ys[ny++] = y;
    if (y >= 1 && y <= kBorder) ys[ny++] = -y;
    if (y <= L.h - 2 && y >= L.h - 1 - kBorder) ys[ny++] = 2 * (L.h - 1) - y;
    for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) L.base[(ptrdiff_t)ys[a] * L.pitch + xs[b]] = v;
}

// ================================================================================================
// gray / CLAHE / pyramid
// ================================================================================================
// cvtColor(*2GRAY) for 8-bit input (OpenCV 4.x 15-bit fixed point), Tracker.cc:183-196.
__global__ void k_gray(const uint8_t* __restrict__ src, int src_pitch, int channels, int is_rgb,
                       uint8_t* __restrict__ dst, int dst_pitch, int w, int h)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    const uint8_t* p = src + (size_t)y * src_pitch + (size_t)x * channels;
    int c0 = p[0], c1 = p[1], c2 = p[2];
    int b = is_rgb ? c2 : c0, r = is_rgb ? c0 : c2;
    dst[(size_t)y * dst_pitch + x] = (uint8_t)((b * 3735 + c1 * 19235 + r * 9798 + (1 << 14)) >> 15);
}

// One CTA (256 threads) per CLAHE tile: histogram of the (reflect-padded) tile, clip, redistribute, LUT.
__global__ void __launch_bounds__(1024) k_clahe_lut(const uint8_t* __restrict__ src, int pitch, int w, int h,
                                                    int tw, int th, int clip, float lut_scale,
                                                    uint8_t* __restrict__ lut /* 25 x 256 */)
{
    rvio::pdl_wait(); rvio::pdl_trigger();      // programmatic dependent launch (common.cuh): nothing above touches memory
    __shared__ int hist[256];
    __shared__ int sh[34];
    const int tid = threadIdx.x;
    const int tx = blockIdx.x % 5, ty = blockIdx.x / 5;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const int area = tw * th;
    for (int i = tid; i < area; i += 1024) {
        const int ex = tx * tw + i % tw, ey = ty * th + i / tw;
        const int sx = reflect101(ex, w), sy = reflect101(ey, h);
        atomicAdd(&hist[src[(size_t)sy * pitch + sx]], 1);
    }
    __syncthreads();
    // the rest is per-bin work: bins live in threads 0..255 (threads >= 256 carry zeros through the block scans)
    int v = tid < 256 ? hist[tid] : 0;
    int excess = v > clip ? v - clip : 0;
    if (v > clip) v = clip;
    const int clipped = block_reduce_sum(excess, sh);
    const int batch = clipped / 256;
    const int residual = clipped - batch * 256;
    if (tid < 256) {
        v += batch;
        if (residual != 0) {
            int step = 256 / residual;
            if (step < 1) step = 1;
            if (tid % step == 0 && tid / step < residual) v++;
        }
    }
    int total;
    const int ex = block_exscan(v, sh, &total);
    if (tid < 256) {
        const int sum = ex + v;
        int q = __float2int_rn(__fmul_rn((float)sum, lut_scale));
        q = q < 0 ? 0 : (q > 255 ? 255 : q);
        lut[blockIdx.x * 256 + tid] = (uint8_t)q;
    }
}

// Bilinear LUT interpolation; writes level 0 of the current pyramid including its reflect border.
__global__ void __launch_bounds__(256) k_clahe_apply(const uint8_t* __restrict__ src, int pitch,
                                                     const uint8_t* __restrict__ lut, float inv_tw, float inv_th,
                                                     PyrLevel dst)
{
    rvio::pdl_wait(); rvio::pdl_trigger();      // programmatic dependent launch (common.cuh): nothing above touches memory
    __shared__ uint8_t slut[25 * 256];
    for (int i = threadIdx.x; i < 25 * 256 / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(slut)[i] = reinterpret_cast<const uint32_t*>(lut)[i];
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dst.w) return;
    const float tyf = __fsub_rn(__fmul_rn((float)y, inv_th), 0.5f);
    int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
    const float ya = __fsub_rn(tyf, (float)ty1), ya1 = __fsub_rn(1.0f, ya);
    ty1 = max(ty1, 0); ty2 = min(ty2, 4);
    const float txf = __fsub_rn(__fmul_rn((float)x, inv_tw), 0.5f);
    int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
    const float xa = __fsub_rn(txf, (float)tx1), xa1 = __fsub_rn(1.0f, xa);
    tx1 = max(tx1, 0); tx2 = min(tx2, 4);
    const int v = src[(size_t)y * pitch + x];
    const float a = __fmul_rn((float)slut[(ty1 * 5 + tx1) * 256 + v], xa1);
    const float b = __fmul_rn((float)slut[(ty1 * 5 + tx2) * 256 + v], xa);
    const float c = __fmul_rn((float)slut[(ty2 * 5 + tx1) * 256 + v], xa1);
    const float d = __fmul_rn((float)slut[(ty2 * 5 + tx2) * 256 + v], xa);
    const float res = __fadd_rn(__fmul_rn(__fadd_rn(a, b), ya1), __fmul_rn(__fadd_rn(c, d), ya));
    int q = __float2int_rn(res);
    q = q < 0 ? 0 : (q > 255 ? 255 : q);
    store_with_border(dst, x, y, (uint8_t)q);
}

// Equalizer off: plain copy into level 0 (with border).
__global__ void k_copy_level0(const uint8_t* __restrict__ src, int pitch, PyrLevel dst)
{
    rvio::pdl_wait(); rvio::pdl_trigger();      // programmatic dependent launch (common.cuh): nothing above touches memory
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dst.w) return;
    store_with_border(dst, x, y, src[(size_t)y * pitch + x]);
}

// cv::pyrDown: separable [1 4 6 4 1], BORDER_REFLECT_101 (read from the source border), (sum+128)>>8.
__global__ void __launch_bounds__(256) k_pyr_down(PyrLevel src, PyrLevel dst)
{
    rvio::pdl_wait(); rvio::pdl_trigger();      // programmatic dependent launch (common.cuh): nothing above touches memory
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dst.w) return;
    const uint8_t* s = src.base + (ptrdiff_t)(2 * y - 2) * src.pitch + (2 * x - 2);
    int acc = 0;
    const int wgt[5] = {1, 4, 6, 4, 1};
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const uint8_t* p = s + (ptrdiff_t)r * src.pitch;
        const int row = p[2] * 6 + (p[1] + p[3]) * 4 + p[0] + p[4];
        acc += row * wgt[r];
    }
    store_with_border(dst, x, y, (uint8_t)((acc + 128) >> 8));
}


// Three pyrDown steps in ONE launch (levels 1, 2, 3 from level 0): each CTA owns an 8 x 8 tile of level 3 and everything
// above it (16 x 16 of level 2, 32 x 32 of level 1), and recomputes the halo it needs (19 x 19 of level 2, 41 x 41 of level
// 1, from 85 x 85 of level 0) in shared memory -- three dependent launches of a few microseconds each become one.
// BORDER_REFLECT_101 is applied with the level's own size on every read; results are identical to k_pyr_down's.
__device__ __forceinline__ int pyr_reflect(int c, int n) { return c < 0 ? -c : (c >= n ? 2 * (n - 1) - c : c); }
template <int RS, int PS, int RD, int PD>
__device__ __forceinline__ void pyr_down_region(const uint8_t* src, int sox, int soy, int sw, int sh,      // source region (origin, level size)
                                                uint8_t* dst, int dox, int doy, const PyrLevel& L,          // destination region (origin), destination level
                                                int own_x0, int own_y0, int own_n)                          // owned square of the destination level
{
    for (int o = threadIdx.x; o < RD * RD; o += 256) {
        const int ry = o / RD, rx = o - ry * RD;
        const int x = dox + rx, y = doy + ry;
        if (x < 0 || y < 0 || x >= L.w || y >= L.h) continue;
        int xi[5], acc = 0;
#pragma unroll
        for (int k = 0; k < 5; ++k) xi[k] = pyr_reflect(2 * x - 2 + k, sw) - sox;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const uint8_t* p = src + (pyr_reflect(2 * y - 2 + r, sh) - soy) * PS;
            const int row = p[xi[2]] * 6 + (p[xi[1]] + p[xi[3]]) * 4 + p[xi[0]] + p[xi[4]];
            acc += row * ((r == 0 || r == 4) ? 1 : ((r == 2) ? 6 : 4));
        }
        const uint8_t v = (uint8_t)((acc + 128) >> 8);
        dst[ry * PD + rx] = v;
        if (x >= own_x0 && x < own_x0 + own_n && y >= own_y0 && y < own_y0 + own_n) store_with_border(L, x, y, v);
    }
}
__global__ void __launch_bounds__(256) k_pyr_down3(PyrLevel l0, PyrLevel l1, PyrLevel l2, PyrLevel l3)
{
    rvio::pdl_wait(); rvio::pdl_trigger();      // programmatic dependent launch (common.cuh): nothing above touches memory
    __shared__ uint8_t s0[85 * 88], s1[41 * 44], s2[19 * 20], s3[8 * 8];
    const int X3 = 8 * blockIdx.x, Y3 = 8 * blockIdx.y;
    const int o2x = 2 * X3 - 2, o2y = 2 * Y3 - 2, o1x = 2 * o2x - 2, o1y = 2 * o2y - 2, o0x = 2 * o1x - 2, o0y = 2 * o1y - 2;
    for (int o = threadIdx.x; o < 85 * 85; o += 256) {
        const int ry = o / 85, rx = o - ry * 85;
        const int x = o0x + rx, y = o0y + ry;
        uint8_t v = 0;
        if (x >= 0 && y >= 0 && x < l0.w && y < l0.h) v = l0.base[(ptrdiff_t)y * l0.pitch + x];
        s0[ry * 88 + rx] = v;
    }
    __syncthreads();
    pyr_down_region<85, 88, 41, 44>(s0, o0x, o0y, l0.w, l0.h, s1, o1x, o1y, l1, 4 * X3, 4 * Y3, 32);
    __syncthreads();
    pyr_down_region<41, 44, 19, 20>(s1, o1x, o1y, l1.w, l1.h, s2, o2x, o2y, l2, 2 * X3, 2 * Y3, 16);
    __syncthreads();
    pyr_down_region<19, 20, 8, 8>(s2, o2x, o2y, l2.w, l2.h, s3, X3, Y3, l3, X3, Y3, 8);
}

// ================================================================================================
// pyramidal Lucas-Kanade, one warp per feature
// ================================================================================================
struct LKParams {
    Pyramid prev, cur;
    const float2* feats;
    int n;                // number of features to track (upper bound when n_dev is set)
    const int* n_dev;     // optional: the count lives on the device (fused pipeline / frame graphs)
    int first, last;      // feature index range of this launch (feature sharding across GPUs; [0, INT_MAX) otherwise)
    float2* out;
    uint8_t* status;
    float2* un;
    CamParams cam;
    int max_iter;
    float eps_sq_f;       // unused placeholder (epsilon compared in double)
    double eps_sq;
    float min_eig_thr;
};

// Accumulation order reproduced from OpenCV's SIMD128 path (lkpyramid.cpp, LKTrackerInvoker):
//   A-matrix: for every window row, columns 0..7 feed four float lanes l[k] (+= v(y,k); += v(y,4+k)),
//             columns 8..14 feed one scalar float accumulator in (y,x) order; total = scalar + ((l0+l2)+(l1+l3)).
//   b-vector: per row eight int32 pair-sums (d_k*g_k + d_{k+4}*g_{k+4}) are converted to float and added to eight
//             lanes; columns 8..14 feed two scalar chains; b1 = s1 + ((q0+0)+(q2+0)), b2 = s2 + ((q1+0)+(q3+0)).
// Work split: lane 2y owns window pixels (y, 0..7), lane 2y+1 owns (y, 8..14) (lanes 30,31 idle); the interpolated
// previous-image patch stays in registers for the whole level, the next-image window is read from a 32x48 shared-memory
// tile staged once per level with 16-byte loads (re-staged only if the window leaves it), and the float32 sums are
// formed in OpenCV's order by walking the owning lanes with shuffles.  The two 105-term scalar chains take an exact
// shortcut whenever every partial sum (in OpenCV's order) and every term is an integer below 2^24: float32 addition
// is then exact, so the sum equals the integer total and is order independent; otherwise they are added one by one.
constexpr int kTileW = 48, kTileH = 32;

// TMA descriptors of the CURRENT pyramid's levels (one 2-D u8 tensor per level: the whole padded level buffer), passed
// to k_lk as a __grid_constant__ parameter; the next-image tile is fetched with cp.async.bulk.tensor.2d into shared memory
// and its arrival is awaited on a per-warp mbarrier (north_star: "image pyramid staged through TMA into shared memory").
struct LKTmaps { CUtensorMap lv[kMaxLevels]; };

struct __align__(128) LKWarpSmem {
    __align__(128) unsigned char tile[kTileH * kTileW];      // TMA destination (128-byte aligned)
    __align__(8) unsigned long long mbar;                    // transaction barrier of the tile loads
    short dgrid[16 * 16 * 2];            // Scharr (dx,dy) on the 16x16 tap grid
    unsigned char Ireg[18 * 20];
    __align__(16) float ch[3 * 232];     // per matrix / chain set: [y*8 + slot] (120 floats) then [120 + y*7 + t] (105 floats)
};

constexpr int kLKWarps = 2;

__device__ __forceinline__ void lk_weights(float a, float b, int* w00, int* w01, int* w10, int* w11)
{
    const float s = 16384.f;
    const float a1 = __fsub_rn(1.f, a), b1 = __fsub_rn(1.f, b);
    *w00 = __float2int_rn(__fmul_rn(__fmul_rn(a1, b1), s));
    *w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, b1), s));
    *w10 = __float2int_rn(__fmul_rn(__fmul_rn(a1, b), s));
    *w11 = 16384 - *w00 - *w01 - *w10;
}

// ---- TMA / mbarrier primitives (sm_100a PTX)
__device__ __forceinline__ unsigned lk_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lk_mbar_init(unsigned long long* bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(lk_smem_addr(bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// One lane: order the warp's earlier generic-proxy reads of the tile before the async-proxy write, arm the barrier with
// the tile's byte count and start the 2-D tensor copy of the kTileW x kTileH box whose top-left element is (x, y).
__device__ __forceinline__ void lk_tma_load_tile(const CUtensorMap* map, unsigned char* tile, unsigned long long* bar, int x, int y)
{
    const unsigned dst = lk_smem_addr(tile), mb = lk_smem_addr(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"((unsigned)(kTileW * kTileH)) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(mb) : "memory");
}
__device__ __forceinline__ void lk_mbar_wait(unsigned long long* bar, unsigned parity)
{
    const unsigned mb = lk_smem_addr(bar);
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(mb), "r"(parity) : "memory");
    } while (!ok);
}

// (pre-TMA staging, kept for the sharded / reference comparison build: -DRVIO_LK_NO_TMA)
// Stage the kTileH x kTileW byte tile whose top-left pixel is (tx0, ty0) (tx0 multiple of 16) from level J.
__device__ __forceinline__ void lk_stage_tile(const PyrLevel& J, int tx0, int ty0, unsigned char* tile, int lane)
{
    const long long total = (long long)J.pitch * (J.h + 2 * kBorder);
    const unsigned char* origin = J.base - (ptrdiff_t)kBorder * J.pitch - kBorder;      // first byte of the level buffer
    for (int c = lane; c < kTileH * (kTileW / 16); c += 32) {
        const int r = c / (kTileW / 16), q = c - r * (kTileW / 16);
        const long long flat = (long long)(ty0 + r + kBorder) * J.pitch + (tx0 + 16 * q + kBorder);
        uint4 v = make_uint4(0, 0, 0, 0);
        if (flat >= 0 && flat + 16 <= total) v = __ldg(reinterpret_cast<const uint4*>(origin + flat));
        *reinterpret_cast<uint4*>(tile + r * kTileW + 16 * q) = v;
    }
}

__global__ void __launch_bounds__(kLKWarps * 32) k_lk(LKParams P, const __grid_constant__ LKTmaps M)
{
    rvio::pdl_wait(); rvio::pdl_trigger();      // programmatic dependent launch (common.cuh): nothing above touches memory
    __shared__ LKWarpSmem smem_all[kLKWarps];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int pt = P.first + blockIdx.x * kLKWarps + wib;
    if (pt >= (P.n_dev ? *P.n_dev : P.n) || pt >= P.last) return;   // whole warp exits together
    LKWarpSmem& S = smem_all[wib];
    unsigned tma_phase = 0;
    if (lane == 0) lk_mbar_init(&S.mbar);
    __syncwarp();
    const unsigned FULL = 0xffffffffu;
    // pixel ownership
    const int oy = lane >> 1, half_ = lane & 1;
    const bool owner = lane < 30;
    const int ox0 = half_ * 8, onx = half_ ? 7 : 8;

    const float2 p0 = P.feats[pt];
    float nx = 0.f, ny = 0.f;                   // nextPts[pt] (kept by all lanes identically)
    float outx = 0.f, outy = 0.f;
    int status = 1;
    const float half = 7.0f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const int levels = P.cur.levels;

    for (int level = levels - 1; level >= 0; --level) {
        const PyrLevel I = P.prev.lv[level];
        const PyrLevel J = P.cur.lv[level];
        const float lscale = (float)(1. / (1 << level));
        float px = __fmul_rn(p0.x, lscale), py = __fmul_rn(p0.y, lscale);
        if (level == levels - 1) { nx = px; ny = py; }
        else { nx = __fmul_rn(outx, 2.f); ny = __fmul_rn(outy, 2.f); }
        outx = nx; outy = ny;

        px = __fsub_rn(px, half); py = __fsub_rn(py, half);
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -kWin || ipx >= I.w || ipy < -kWin || ipy >= I.h) {
            if (level == 0) status = 0;
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), &w00, &w01, &w10, &w11);

        // ---- stage the 18x18 neighbourhood of the previous image (origin ipx-1, ipy-1)
        __syncwarp();
        for (int i = lane; i < 18 * 18; i += 32) {
            const int r = i / 18, c = i - r * 18;
            S.Ireg[r * 20 + c] = I.base[(ptrdiff_t)(ipy - 1 + r) * I.pitch + (ipx - 1 + c)];
        }
        __syncwarp();
        // ---- Scharr on the 16x16 tap grid; zero outside the image (derivative border is constant 0)
        for (int i = lane; i < 256; i += 32) {
            const int r = i >> 4, c = i & 15;
            const int gx = ipx + c, gy = ipy + r;
            int dx = 0, dy = 0;
            if (gx >= 0 && gx < I.w && gy >= 0 && gy < I.h) {
                const unsigned char* q = &S.Ireg[r * 20 + c];      // top-left of the 3x3 block
                const int a00 = q[0], a01 = q[1], a02 = q[2];
                const int a10 = q[20], a12 = q[22];
                const int a20 = q[40], a21 = q[41], a22 = q[42];
                dx = ((a02 + a22) * 3 + a12 * 10) - ((a00 + a20) * 3 + a10 * 10);
                dy = ((a22 - a02) + (a20 - a00)) * 3 + (a21 - a01) * 10;
            }
            S.dgrid[2 * i] = (short)dx;
            S.dgrid[2 * i + 1] = (short)dy;
        }
        __syncwarp();
        // ---- interpolated patch I, Ix, Iy of the pixels this lane owns (registers) + A-matrix products
        int Iw[8], Ixw[8], Iyw[8];
        float pa11[8], pa12[8], pa22[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            Iw[j] = 0; Ixw[j] = 0; Iyw[j] = 0;
            if (owner && j < onx) {
                const int x = ox0 + j;
                const unsigned char* s = &S.Ireg[(oy + 1) * 20 + (x + 1)];
                Iw[j] = (s[0] * w00 + s[1] * w01 + s[20] * w10 + s[21] * w11 + (1 << 8)) >> 9;
                const short* g = &S.dgrid[2 * (oy * 16 + x)];
                Ixw[j] = (g[0] * w00 + g[2] * w01 + g[32] * w10 + g[34] * w11 + (1 << 13)) >> 14;
                Iyw[j] = (g[1] * w00 + g[3] * w01 + g[33] * w10 + g[35] * w11 + (1 << 13)) >> 14;
            }
            pa11[j] = (float)(Ixw[j] * Ixw[j]);
            pa12[j] = (float)(Ixw[j] * Iyw[j]);
            pa22[j] = (float)(Iyw[j] * Iyw[j]);
        }
        float A11, A12, A22;
        {
            // stage the products: lane 2y -> slots [y*8 .. y*8+7], lane 2y+1 -> tail [120 + y*7 .. +6]
            __syncwarp();
            if (owner) {
                if (half_ == 0) {
                    float4* d0 = reinterpret_cast<float4*>(&S.ch[oy * 8]);
                    float4* d1 = reinterpret_cast<float4*>(&S.ch[232 + oy * 8]);
                    float4* d2 = reinterpret_cast<float4*>(&S.ch[464 + oy * 8]);
                    d0[0] = make_float4(pa11[0], pa11[1], pa11[2], pa11[3]); d0[1] = make_float4(pa11[4], pa11[5], pa11[6], pa11[7]);
                    d1[0] = make_float4(pa12[0], pa12[1], pa12[2], pa12[3]); d1[1] = make_float4(pa12[4], pa12[5], pa12[6], pa12[7]);
                    d2[0] = make_float4(pa22[0], pa22[1], pa22[2], pa22[3]); d2[1] = make_float4(pa22[4], pa22[5], pa22[6], pa22[7]);
                } else {
#pragma unroll
                    for (int t = 0; t < 7; ++t) {
                        S.ch[120 + oy * 7 + t] = pa11[t];
                        S.ch[232 + 120 + oy * 7 + t] = pa12[t];
                        S.ch[464 + 120 + oy * 7 + t] = pa22[t];
                    }
                }
            }
            __syncwarp();
            // 15 chain lanes: lane = 5*m + j ; j<4: SIMD lane j of matrix m (30 terms: (y,j),(y,4+j)) ; j==4: scalar chain (105 terms)
            float acc = 0.f;
            {
                const int m = lane / 5, jn = lane - 5 * m;
                const float* pr = S.ch + 232 * (m < 3 ? m : 0);
                const bool is_long = jn == 4;
                const int cnt = lane < 15 ? (is_long ? 105 : 30) : 0;
#pragma unroll 5
                for (int i = 0; i < 105; ++i) {
                    const int off = is_long ? (120 + i) : ((i >> 1) * 8 + (i & 1) * 4 + jn);
                    if (i < cnt) acc = __fadd_rn(acc, pr[off]);
                }
            }
            float t[3];
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const float l0 = __shfl_sync(FULL, acc, m * 5 + 0), l1 = __shfl_sync(FULL, acc, m * 5 + 1);
                const float l2 = __shfl_sync(FULL, acc, m * 5 + 2), l3 = __shfl_sync(FULL, acc, m * 5 + 3);
                const float sc = __shfl_sync(FULL, acc, m * 5 + 4);
                t[m] = __fadd_rn(sc, __fadd_rn(__fadd_rn(l0, l2), __fadd_rn(l1, l3)));
            }
            A11 = __fmul_rn(t[0], FLT_SCALE); A12 = __fmul_rn(t[1], FLT_SCALE); A22 = __fmul_rn(t[2], FLT_SCALE);
        }
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        float minEig;
        {
            const float df = __fsub_rn(A11, A22);
            const float t3 = __fadd_rn(__fmul_rn(df, df), __fmul_rn(__fmul_rn(4.f, A12), A12));
            minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(t3)), 450.f);
        }
        if (minEig < P.min_eig_thr || D < 1.1920928955078125e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
        float pdx = 0.f, pdy = 0.f;
        int tx0 = 0, ty0 = 0;
        bool have_tile = false;

        for (int j = 0; j < P.max_iter; ++j) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -kWin || inx >= J.w || iny < -kWin || iny >= J.h) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), &w00, &w01, &w10, &w11);
            // ---- make sure the 16x16 window lies inside the staged tile
            if (!have_tile || inx < tx0 || inx + 16 > tx0 + kTileW || iny < ty0 || iny + 16 > ty0 + kTileH) {
                tx0 = ((inx - 8 + 1024) & ~15) - 1024;           // floor to a multiple of 16 (also for negatives)
                ty0 = iny - 8;
                __syncwarp();
#ifdef RVIO_LK_NO_TMA
                lk_stage_tile(J, tx0, ty0, S.tile, lane);
#else
                // TMA: one lane issues the box copy (coordinates relative to the padded level buffer; out-of-buffer parts are
                // zero filled by the hardware), the warp waits on the tile's transaction barrier
                if (lane == 0) lk_tma_load_tile(&M.lv[level], S.tile, &S.mbar, tx0 + kBorder, ty0 + kBorder);
                lk_mbar_wait(&S.mbar, tma_phase);
                tma_phase ^= 1u;
#endif
                __syncwarp();
                have_tile = true;
            }
            // ---- diff for the owned pixels, SIMD-slot pair sums / tail products
            float v[8];              // half 0: eight SIMD slot values ; half 1: v[0..6] = d*Ix
            float u[8];              // half 1: u[0..6] = d*Iy
            int dterm_ok = 1;        // every tail term representable (|term| < 2^24)
            int ti1[7], ti2[7];
            {
                int dd[8];
                const unsigned char* r0 = S.tile + (iny - ty0 + oy) * kTileW + (inx - tx0 + ox0);
                const unsigned char* r1 = r0 + kTileW;
                int a = owner ? r0[0] : 0, c = owner ? r1[0] : 0;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    dd[q] = 0;
                    if (owner && q < onx) {
                        const int b = r0[q + 1], e = r1[q + 1];
                        dd[q] = ((a * w00 + b * w01 + c * w10 + e * w11 + (1 << 8)) >> 9) - Iw[q];
                        a = b; c = e;
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) { v[q] = 0.f; u[q] = 0.f; }
                if (half_ == 0) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        v[2 * k] = (float)(dd[k] * Ixw[k] + dd[k + 4] * Ixw[k + 4]);
                        v[2 * k + 1] = (float)(dd[k] * Iyw[k] + dd[k + 4] * Iyw[k + 4]);
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < 7; ++t) {
                        ti1[t] = dd[t] * Ixw[t]; ti2[t] = dd[t] * Iyw[t];
                        v[t] = (float)ti1[t]; u[t] = (float)ti2[t];
                        if (abs(ti1[t]) >= (1 << 24) || abs(ti2[t]) >= (1 << 24)) dterm_ok = 0;
                    }
                }
                if (half_ == 0) {
#pragma unroll
                    for (int t = 0; t < 7; ++t) { ti1[t] = 0; ti2[t] = 0; }
                }
            }
            // ---- stage: lane 2y -> eight SIMD slot values [y*8 + c]; lane 2y+1 -> tails [120 + y*7 + t] (Ix in set 0, Iy in set 1)
            __syncwarp();
            if (owner) {
                if (half_ == 0) {
                    float4* d0 = reinterpret_cast<float4*>(&S.ch[oy * 8]);
                    d0[0] = make_float4(v[0], v[1], v[2], v[3]); d0[1] = make_float4(v[4], v[5], v[6], v[7]);
                } else {
#pragma unroll
                    for (int t = 0; t < 7; ++t) { S.ch[120 + oy * 7 + t] = v[t]; S.ch[232 + 120 + oy * 7 + t] = u[t]; }
                }
            }
            // ---- exact shortcut for the two scalar chains: every term and every partial sum (in OpenCV's order) below 2^24
            bool exact;
            int tot1 = 0, tot2 = 0;
            {
                int loc1 = 0, loc2 = 0;
#pragma unroll
                for (int t = 0; t < 7; ++t) { loc1 += ti1[t]; loc2 += ti2[t]; }
                int inc1 = loc1, inc2 = loc2;               // inclusive scan over lanes (even lanes contribute 0)
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int a1 = __shfl_up_sync(FULL, inc1, o), a2 = __shfl_up_sync(FULL, inc2, o);
                    if (lane >= o) { inc1 += a1; inc2 += a2; }
                }
                int ok = dterm_ok;
                int p1 = inc1 - loc1, p2 = inc2 - loc2;      // exclusive prefix of this lane
#pragma unroll
                for (int t = 0; t < 7; ++t) {
                    p1 += ti1[t]; p2 += ti2[t];
                    if (abs(p1) >= (1 << 24) || abs(p2) >= (1 << 24)) ok = 0;
                }
                // with every |term| < 2^24 no prefix can leave int32 (105 * 2^24 < 2^31); otherwise ok is already 0
                exact = __all_sync(FULL, ok) != 0;
                tot1 = __shfl_sync(FULL, inc1, 31); tot2 = __shfl_sync(FULL, inc2, 31);
            }
            __syncwarp();
            // ---- chain lanes: 0..7 SIMD lanes (15 terms each), 8 / 9 the scalar chains (105 terms, only when not exact)
            float acc = 0.f;
            {
                const bool is_long = lane >= 8;
                const float* pr = S.ch + (lane == 9 ? 232 : 0);
                const int cnt = lane < 8 ? 15 : ((lane < 10 && !exact) ? 105 : 0);
                const int n_it = exact ? 15 : 105;
#pragma unroll 5
                for (int i = 0; i < n_it; ++i) {
                    const int off = is_long ? (120 + i) : (i * 8 + lane);
                    if (i < cnt) acc = __fadd_rn(acc, pr[off]);
                }
            }
            const float q0 = __fadd_rn(__shfl_sync(FULL, acc, 0), __shfl_sync(FULL, acc, 4));
            const float q1 = __fadd_rn(__shfl_sync(FULL, acc, 1), __shfl_sync(FULL, acc, 5));
            const float q2 = __fadd_rn(__shfl_sync(FULL, acc, 2), __shfl_sync(FULL, acc, 6));
            const float q3 = __fadd_rn(__shfl_sync(FULL, acc, 3), __shfl_sync(FULL, acc, 7));
            float s1 = __shfl_sync(FULL, acc, 8), s2 = __shfl_sync(FULL, acc, 9);
            if (exact) { s1 = (float)tot1; s2 = (float)tot2; }
            const float sb1 = __fadd_rn(s1, __fadd_rn(__fadd_rn(q0, 0.f), __fadd_rn(q2, 0.f)));
            const float sb2 = __fadd_rn(s2, __fadd_rn(__fadd_rn(q1, 0.f), __fadd_rn(q3, 0.f)));
            const float b1 = __fmul_rn(sb1, FLT_SCALE), b2 = __fmul_rn(sb2, FLT_SCALE);
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            outx = __fadd_rn(nx, half); outy = __fadd_rn(ny, half);
            if ((double)dx * (double)dx + (double)dy * (double)dy <= P.eps_sq) break;
            if (j > 0 && fabs((double)__fadd_rn(dx, pdx)) < 0.01 && fabs((double)__fadd_rn(dy, pdy)) < 0.01) {
                outx = __fsub_rn(outx, __fmul_rn(dx, 0.5f));
                outy = __fsub_rn(outy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {
            const float fx = __fsub_rn(outx, half), fy = __fsub_rn(outy, half);
            const int ix = (int)floorf(fx), iy = (int)floorf(fy);
            if (ix < -kWin || ix >= J.w || iy < -kWin || iy >= J.h) status = 0;
        }
    }
    if (lane == 0) {
        P.out[pt] = make_float2(outx, outy);
        P.status[pt] = (uint8_t)status;
        float ux, uy;
        cam_undistort(P.cam, outx, outy, &ux, &uy);      // all points, also status 0 (Tracker.cc:253)
        P.un[pt] = make_float2(ux, uy);
    }
}

// ================================================================================================
// RANSAC (single CTA)
// ================================================================================================
// glibc rand() state, staged in shared memory for the draw loop (the state lives in the tracker's device scalars: every
// draw would otherwise be a dependent load-modify-store round trip to L2)
struct RngLocal { int r[34]; int f, b; };
__device__ __forceinline__ int glibc_rand_next(RngLocal* sc)
{
    // glibc random_r.c TYPE_3: r[f] += r[b]; result = r[f] >> 1
    unsigned v = (unsigned)sc->r[sc->f] + (unsigned)sc->r[sc->b];
    sc->r[sc->f] = (int)v;
    const int res = (int)(v >> 1);
    if (++sc->f >= 31) { sc->f = 0; ++sc->b; }
    else if (++sc->b >= 31) sc->b = 0;
    return res;
}

struct RansacParams {
    TrackerBuffers B;
    int n;                  // features fed to LK (ignored when n_dev is set)
    const int* n_dev;
    int use_sampson;
    double thr, small_angle;
    const double* R;        // device, 9 doubles: Rci * prod(dR_k) * Ric, Ransac.cc:120-155 (evaluated on the host with libm)
};

__device__ __forceinline__ void m3mul(const double* A, const double* B, double* C)
{
    double T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    for (int i = 0; i < 9; ++i) C[i] = T[i];
}

__device__ __forceinline__ double epi_dist(const double* E, double x1, double y1, double x2, double y2, int sampson)
{
    // points are (x,y,1); Ransac.cc:250-266
    const double Fx10 = E[0] * x1 + E[1] * y1 + E[2], Fx11 = E[3] * x1 + E[4] * y1 + E[5], Fx12 = E[6] * x1 + E[7] * y1 + E[8];
    const double Fx20 = E[0] * x2 + E[3] * y2 + E[6], Fx21 = E[1] * x2 + E[4] * y2 + E[7], Fx22 = E[2] * x2 + E[5] * y2 + E[8];
    const double num = Fx20 * x1 + Fx21 * y1 + Fx22;      // (p2^T E) p1
    (void)Fx12;
    if (!sampson) return fabs(num);
    return (num * num) / (Fx10 * Fx10 + Fx11 * Fx11 + Fx20 * Fx20 + Fx21 * Fx21);
}

#ifdef RVIO_B200_PHASE_CLOCKS
__device__ long long g_trk_clk[32];
#define TRK_CLK(k) do { if (threadIdx.x == 0) g_trk_clk[k] = clock64(); } while (0)
#else
#define TRK_CLK(k) do { } while (0)
#endif
__device__ __forceinline__ void ransac_body(const RansacParams& P, const int n)
{
    __shared__ int sh[34];
    __shared__ double sR[9];
    __shared__ double sE[kRansacIters * 9];
    __shared__ int sCnt[kRansacIters];
    __shared__ int sWinner;
    __shared__ unsigned s_used[128];
    __shared__ RngLocal s_rng;
    __shared__ int s_pick[2 * kRansacIters];
    const int tid = threadIdx.x;
    const TrackerBuffers& B = P.B;
    TrackerScalars* sc = B.sc;
    if (tid < 34) s_rng.r[tid] = sc->rng_r[tid];
    if (tid == 34) { s_rng.f = sc->rng_f; s_rng.b = sc->rng_b; }
    if (tid >= 64 && tid < 73) sR[tid - 64] = P.R[tid - 64];   // GetRotation: evaluated on the host, uploaded per frame (see tracker_enqueue)

    TRK_CLK(0);
    // flags start as the LK status (Tracker.cc:264 passes vInlierFlag in/out)
    for (int i = tid; i < n; i += 256) B.flags[i] = B.status[i];
    // candidate compaction in index order (Ransac.cc:190-199)
    int base = 0;
    for (int start = 0; start < n; start += 256) {
        const int i = start + tid;
        const int f = (i < n && B.status[i]) ? 1 : 0;
        int tot;
        const int ex = block_exscan(f, sh, &tot);
        if (f) B.cand[base + ex] = i;
        base += tot;
    }
    const int nc = base;
    // (diagnostics are zeroed before the barrier: lanes of warp 0 may not reconverge between divergent ifs)
    if (tid < kRansacIters) { sCnt[tid] = 0; B.n_inliers[tid] = 0; }
    if (tid < 32) B.two_points[tid] = 0;
    __syncthreads();
    if (tid == 0) { sc->n_cand = nc; sc->winner = 0; sc->ransac_ran = 0; }
    // <=16 candidates: flags untouched (Ransac.cc:201-205).  17..31 candidates make the reference spin forever in
    // SetPointPair (Ransac.cc:57-82); defined here as "flags untouched", no rand() consumed.
    if (nc < 2 * kRansacIters) return;

    if (tid < 128) s_used[tid] = 0u;
    __syncthreads();
    TRK_CLK(1);
    if (tid == 0) {
        // SetPointPair (Ransac.cc:50-83); "used" is a shared-memory bitmask (nc <= 4096); the draws run on the staged state
        for (int it = 0; it < kRansacIters; ++it) {
            int a, b;
            do { a = glibc_rand_next(&s_rng) % nc; } while ((s_used[a >> 5] >> (a & 31)) & 1u);
            do { b = glibc_rand_next(&s_rng) % nc; } while (((s_used[b >> 5] >> (b & 31)) & 1u) || a == b);
            s_pick[2 * it] = a; s_pick[2 * it + 1] = b;
            s_used[a >> 5] |= 1u << (a & 31); s_used[b >> 5] |= 1u << (b & 31);
        }
        sc->ransac_ran = 1;
    }
    __syncthreads();
    TRK_CLK(2);
    if (tid < 34) sc->rng_r[tid] = s_rng.r[tid];
    if (tid == 34) { sc->rng_f = s_rng.f; sc->rng_b = s_rng.b; }
    if (tid >= 64 && tid < 64 + 2 * kRansacIters) { const int pi = B.cand[s_pick[tid - 64]]; s_pick[tid - 64] = pi; B.two_points[tid - 64] = pi; }
    __syncthreads();
    if (tid < kRansacIters) {
        // SetRansacModel (Ransac.cc:86-117)
        const int ia = s_pick[2 * tid], ib = s_pick[2 * tid + 1];
        const double A1[3] = {(double)B.pts1[ia].x, (double)B.pts1[ia].y, 1.0};
        const double A2[3] = {(double)B.un[ia].x, (double)B.un[ia].y, 1.0};
        const double B1[3] = {(double)B.pts1[ib].x, (double)B.pts1[ib].y, 1.0};
        const double B2[3] = {(double)B.un[ib].x, (double)B.un[ib].y, 1.0};
        double A0[3], B0[3];
        for (int i = 0; i < 3; ++i) {
            A0[i] = sR[3 * i] * A1[0] + sR[3 * i + 1] * A1[1] + sR[3 * i + 2] * A1[2];
            B0[i] = sR[3 * i] * B1[0] + sR[3 * i + 1] * B1[1] + sR[3 * i + 2] * B1[2];
        }
        const double c1 = A2[0] * A0[1] - A0[0] * A2[1];
        const double c2 = A0[1] * A2[2] - A2[1] * A0[2];
        const double c3 = A2[0] * A0[2] - A0[0] * A2[2];
        const double c4 = B2[0] * B0[1] - B0[0] * B2[1];
        const double c5 = B0[1] * B2[2] - B2[1] * B0[2];
        const double c6 = B2[0] * B0[2] - B0[0] * B2[2];
        const double alpha = atan2(c3 * c5 - c2 * c6, c1 * c6 - c3 * c4);
        const double beta = atan2(-c3, c1 * sin(alpha) + c2 * cos(alpha));
        const double t[3] = {sin(beta) * cos(alpha), cos(beta), -sin(beta) * sin(alpha)};
        const double tx[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
        double E[9];
        m3mul(tx, sR, E);
        for (int i = 0; i < 9; ++i) { sE[9 * tid + i] = E[i]; B.hyp[9 * tid + i] = E[i]; }
    }
    __syncthreads();
    TRK_CLK(3);
    // CountInliers (Ransac.cc:158-177): 16 x nc Sampson tests
    {
        // point-major: every candidate is loaded once and tested against all hypotheses; the counts are formed with warp votes
        // (one ballot per hypothesis) instead of sixteen shuffle reductions
        int wcnt[kRansacIters];
#pragma unroll
        for (int it = 0; it < kRansacIters; ++it) wcnt[it] = 0;
        for (int k0 = 0; k0 < nc; k0 += 256) {
            const int k = k0 + tid;
            unsigned msk = 0;
            if (k < nc) {
                const int idx = B.cand[k];
                const float2 p1 = B.pts1[idx], p2 = B.un[idx];
#pragma unroll
                for (int it = 0; it < kRansacIters; ++it)
                    if (epi_dist(&sE[9 * it], p1.x, p1.y, p2.x, p2.y, P.use_sampson) < P.thr) msk |= 1u << it;
            }
#pragma unroll
            for (int it = 0; it < kRansacIters; ++it) wcnt[it] += __popc(__ballot_sync(0xffffffffu, (msk >> it) & 1u));
        }
        if ((tid & 31) == 0) {
#pragma unroll
            for (int it = 0; it < kRansacIters; ++it) if (wcnt[it]) atomicAdd(&sCnt[it], wcnt[it]);
        }
    }
    __syncthreads();
    TRK_CLK(4);
    if (tid == 0) {
        int best = 0, bi = 0;
        for (int it = 0; it < kRansacIters; ++it) {
            B.n_inliers[it] = sCnt[it];
            if (sCnt[it] > best) { best = sCnt[it]; bi = it; }      // strict '>' (Ransac.cc:218)
        }
        sWinner = bi;
        sc->winner = bi;
    }
    __syncthreads();
    const double* W = &sE[9 * sWinner];
    for (int k = tid; k < nc; k += 256) {
        const int idx = B.cand[k];
        const double d = epi_dist(W, B.pts1[idx].x, B.pts1[idx].y, B.un[idx].x, B.un[idx].y, P.use_sampson);
        if (d > P.thr || isnan(d)) B.flags[idx] = 0;                 // Ransac.cc:238
    }
}

// ================================================================================================
// bookkeeping (single CTA, 256 threads, features processed in chunks of 256 in index order)
// ================================================================================================
__device__ __forceinline__ float2 hist_get(const TrackerBuffers& B, int slot, int k)
{
    return B.hist[(size_t)slot * B.hist_cap + (B.hist_head[slot] + k) % B.hist_cap];
}
__device__ __forceinline__ void hist_push(const TrackerBuffers& B, int slot, float2 v)
{
    const int len = B.hist_len[slot];
    B.hist[(size_t)slot * B.hist_cap + (B.hist_head[slot] + len) % B.hist_cap] = v;
    B.hist_len[slot] = len + 1;
}

__device__ __forceinline__ void bookkeep_body(const TrackerBuffers& B, int n)
{
    __shared__ int sh[34];
    const int tid = threadIdx.x;
    TrackerScalars* sc = B.sc;
    const int fq_head = sc->fq_head, fq_n0 = sc->fq_n;
    const int cap = B.F + 1;
    // keep = Lmax - (ceil(.5*Lmax) - 1)   (Tracker.cc:327)
    const int keep = B.Lmax - ((B.Lmax + 1) / 2 - 1);
    __syncthreads();

    int n_lost = 0, n_up = 0, n_meas = 0;
    // ---- pass 1: lost features, in LK order (Tracker.cc:283-303)
    for (int start = 0; start < n; start += 256) {
        const int i = start + tid;
        const bool lost = (i < n) && !B.flags[i];
        const int slot = lost ? B.slots[i] : 0;
        const int len = lost ? B.hist_len[slot] : 0;
        const int want = (lost && len >= B.Lmin) ? 1 : 0;
        int tot_l, tot_w, tot_m;
        const int ex_l = block_exscan(lost ? 1 : 0, sh, &tot_l);
        const int ex_w = block_exscan(want, sh, &tot_w);
        const bool emit = want && (n_up + ex_w) < B.Fu;
        const int ex_m = block_exscan(emit ? len : 0, sh, &tot_m);
        if (lost) {
            B.freeq[(fq_head + fq_n0 + n_lost + ex_l) % cap] = slot;
            if (emit) {
                const int u = n_up + ex_w, off = n_meas + ex_m;
                B.up_types[u] = '1';
                B.up_off[u] = off;
                for (int k = 0; k < len; ++k) B.up_xy[off + k] = hist_get(B, slot, k);
            }
            B.hist_len[slot] = 0;
            B.hist_head[slot] = 0;
        }
        n_lost += tot_l;
        const int emitted = min(tot_w, max(B.Fu - n_up, 0));
        n_up += emitted;
        n_meas += tot_m;
    }
    TRK_CLK(6);
    // ---- pass 2: tracked features (Tracker.cc:305-342)
    int n_in = 0;
    for (int start = 0; start < n; start += 256) {
        const int i = start + tid;
        const bool trk = (i < n) && B.flags[i];
        const int slot = trk ? B.slots[i] : 0;
        const int len = trk ? B.hist_len[slot] : 0;
        const int full = (trk && len == B.Lmax) ? 1 : 0;
        int tot_t, tot_f, tot_m;
        const int ex_t = block_exscan(trk ? 1 : 0, sh, &tot_t);
        const int ex_f = block_exscan(full, sh, &tot_f);
        const bool emit = full && (n_up + ex_f) < B.Fu;
        const int ex_m = block_exscan(emit ? len : 0, sh, &tot_m);
        if (trk) {
            const int r = n_in + ex_t;
            B.slots_new[r] = slot;
            B.feats_new[r] = B.lk[i];
            const float2 u2 = B.un[i];
            if (full) {
                if (emit) {
                    const int u = n_up + ex_f, off = n_meas + ex_m;
                    B.up_types[u] = '2';
                    B.up_off[u] = off;
                    for (int k = 0; k < len; ++k) B.up_xy[off + k] = hist_get(B, slot, k);
                    const int drop = len - keep;                   // pop_front until size <= keep
                    if (drop > 0) {
                        B.hist_head[slot] = (B.hist_head[slot] + drop) % B.hist_cap;
                        B.hist_len[slot] = len - drop;
                    }
                } else {
                    B.hist_head[slot] = (B.hist_head[slot] + 1) % B.hist_cap;
                    B.hist_len[slot] = len - 1;
                }
            }
            hist_push(B, slot, u2);
            B.pts1_new[r] = u2;
        }
        n_in += tot_t;
        const int emitted = min(tot_f, max(B.Fu - n_up, 0));
        n_up += emitted;
        n_meas += tot_m;
    }
    TRK_CLK(7);
    if (tid == 0) {
        B.up_off[n_up] = n_meas;
        sc->fq_n = fq_n0 + n_lost;
        sc->n_new = n_in;
        sc->n_up = n_up;
        sc->n_meas = n_meas;
    }
}

// RANSAC followed by the bookkeeping in ONE single-CTA launch (Tracker.cc:264-342).
__global__ void __launch_bounds__(256) k_ransac_bookkeep(RansacParams P)
{
    rvio::pdl_wait(); rvio::pdl_trigger();      // programmatic dependent launch (common.cuh): nothing above touches memory
    const int n = P.n_dev ? *P.n_dev : P.n;       // read before anything below rewrites the scalars
    ransac_body(P, n);
    __syncthreads();
    TRK_CLK(5);
    __threadfence_block();
    bookkeep_body(P.B, n);
    TRK_CLK(8);
}

// First image (Tracker.cc:215-233): slot i <- corner i, free list = n..F-1.
__global__ void k_seed(TrackerBuffers B, const float2* __restrict__ px, int n, const int* __restrict__ n_dev, CamParams cam)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n_dev) n = min(*n_dev, B.F);                 // corner count produced on the device (device detector)
    if (i < n) {
        const float2 p = px[i];
        float ux, uy;
        cam_undistort(cam, p.x, p.y, &ux, &uy);
        B.feats_new[i] = p;
        B.slots_new[i] = i;
        B.pts1_new[i] = make_float2(ux, uy);
        B.hist_head[i] = 0;
        B.hist_len[i] = 0;
        hist_push(B, i, make_float2(ux, uy));
    } else if (i < B.F) {
        B.freeq[i - n] = i;
        B.hist_head[i] = 0;
        B.hist_len[i] = 0;
    }
    if (i == 0) {
        B.sc->fq_head = 0;
        B.sc->fq_n = B.F - n;
        B.sc->n_new = n;
    }
}

// Refill (Tracker.cc:358-386): corner k takes the k-th free slot.
__global__ void k_refill(TrackerBuffers B, const float2* __restrict__ px, int n_use, CamParams cam)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    TrackerScalars* sc = B.sc;
    const int head = sc->fq_head, n_new = sc->n_new;
    if (k < n_use) {
        const int slot = B.freeq[(head + k) % (B.F + 1)];
        const float2 p = px[k];
        float ux, uy;
        cam_undistort(cam, p.x, p.y, &ux, &uy);
        B.slots_new[n_new + k] = slot;
        B.feats_new[n_new + k] = p;
        B.pts1_new[n_new + k] = make_float2(ux, uy);
        hist_push(B, slot, make_float2(ux, uy));
    }
}
__global__ void k_refill_commit(TrackerBuffers B, int n_use)
{
    TrackerScalars* sc = B.sc;
    sc->fq_head = (sc->fq_head + n_use) % (B.F + 1);
    sc->fq_n -= n_use;
    sc->n_new += n_use;
}

}  // namespace rvio

// ================================================================================================
// host side: handle + C ABI
// ================================================================================================
using namespace rvio;

struct rvio_tracker {
    rvio_tracker_cfg cfg;
    int device;
    cudaStream_t stream;
    int W, H, F, Fu, Lmax, Lmin;
    bool first;
    int n_track;                // host mirror of mnFeatsToTrack
    int last_n;                 // features fed to LK in the last track()
    bool frame_open;            // track() ran, commit() pending
    bool shard_open;            // track_begin() ran, track_finish() pending
    Detector det;               // FeatureDetector::DetectWithSubPix on the device (detector.cu)
    cudaEvent_t ev_level0;      // recorded when the equalised level 0 of the current frame is complete
    CamParams cam;
    double Ric[9];
    // device memory
    uint8_t* d_in; size_t in_pitch; int in_channels_cap;
    uint8_t* d_gray; size_t gray_pitch;
    uint8_t* d_lut;
    uint8_t* d_pyr_mem[2];
    Pyramid pyr[2];             // [cur_idx] = current, [1-cur_idx] = previous
    LKTmaps tmaps[2];           // TMA descriptors of the levels of pyr[0] / pyr[1]
    int cur_idx;
    TrackerBuffers B;
    double* d_R; int imu_cap;   // RANSAC rotation of the frame (9 doubles)
    float2* d_px_in;            // seed/refill staging
    // pinned host staging
    uint8_t* h_img; size_t h_img_bytes;
    double* h_R;
    float* h_px;
    TrackerScalars* h_sc;
    // CLAHE constants
    int tw, th, clip;
    float lut_scale, inv_tw, inv_th;
    std::vector<void*> allocs;
};

namespace {

template <typename T>
int dalloc(rvio_tracker* t, T** p, size_t count)
{
    void* q = nullptr;
    RVIO_CUDA_TRY(cudaMalloc(&q, count * sizeof(T) + 16));
    RVIO_CUDA_TRY(cudaMemsetAsync(q, 0, count * sizeof(T) + 16, t->stream));
    t->allocs.push_back(q);
    *p = (T*)q;
    return RVIO_OK;
}

int build_pyramid_layout(rvio_tracker* t, int which)
{
    Pyramid& P = t->pyr[which];
    size_t total = 0;
    int w = t->W, h = t->H;
    size_t offs[kMaxLevels];
    int pitches[kMaxLevels], ws[kMaxLevels], hs[kMaxLevels];
    int levels = 0;
    for (int l = 0; l < kMaxLevels; ++l) {
        if (l > 0) { w = (w + 1) / 2; h = (h + 1) / 2; }
        if (l > 0 && (w <= kWin || h <= kWin)) break;      // buildOpticalFlowPyramid stops here
        if (w <= kBorder + 1 || h <= kBorder + 1) break;   // single-reflection border needs > 17 px
        const int pitch = ((w + 2 * kBorder + 127) / 128) * 128;
        offs[l] = total; pitches[l] = pitch; ws[l] = w; hs[l] = h;
        total += (size_t)pitch * (h + 2 * kBorder);
        levels = l + 1;
    }
    if (levels == 0) { set_error("tracker", "image too small"); return RVIO_ERR_ARG; }
    uint8_t* mem = nullptr;
    RVIO_CUDA_TRY(cudaMalloc((void**)&mem, total + 256));
    RVIO_CUDA_TRY(cudaMemsetAsync(mem, 0, total + 256, t->stream));
    t->allocs.push_back(mem);
    t->d_pyr_mem[which] = mem;
    P.levels = levels;
    for (int l = 0; l < levels; ++l) {
        P.lv[l].base = mem + offs[l] + (size_t)kBorder * pitches[l] + kBorder;
        P.lv[l].pitch = pitches[l]; P.lv[l].w = ws[l]; P.lv[l].h = hs[l];
    }
    return RVIO_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int build_tensor_maps(rvio_tracker* t, int which)
{
    static PFN_tmapEncodeTiled enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        RVIO_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess) { set_error("cudaGetDriverEntryPoint", "cuTensorMapEncodeTiled not available"); return RVIO_ERR_CUDA; }
        enc = (PFN_tmapEncodeTiled)fn;
    }
    const Pyramid& P = t->pyr[which];
    memset(&t->tmaps[which], 0, sizeof(LKTmaps));
    for (int l = 0; l < P.levels; ++l) {
        const PyrLevel& L = P.lv[l];
        void* origin = (void*)(L.base - (ptrdiff_t)kBorder * L.pitch - kBorder);      // first byte of the padded level buffer
        const cuuint64_t dims[2] = {(cuuint64_t)L.pitch, (cuuint64_t)(L.h + 2 * kBorder)};
        const cuuint64_t strides[1] = {(cuuint64_t)L.pitch};
        const cuuint32_t box[2] = {(cuuint32_t)kTileW, (cuuint32_t)kTileH};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&t->tmaps[which].lv[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, origin, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled", "failed for a pyramid level"); return RVIO_ERR_CUDA; }
    }
    for (int l = P.levels; l < kMaxLevels; ++l) t->tmaps[which].lv[l] = t->tmaps[which].lv[0];
    return RVIO_OK;
}

int sync_scalars(rvio_tracker* t)
{
    RVIO_ENQ(cudaMemcpyAsync(t->h_sc, t->B.sc, sizeof(TrackerScalars), cudaMemcpyDeviceToHost, t->stream));
    RVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    return RVIO_OK;
}

}  // namespace

extern "C" int rvio_tracker_create(const rvio_tracker_cfg* cfg, int device, rvio_tracker** out)
{
    RVIO_ARG_CHECK(cfg && out);
    RVIO_ARG_CHECK(cfg->width > 0 && cfg->height > 0 && cfg->n_features > 0);
    RVIO_ARG_CHECK(cfg->max_track_len >= 2 && cfg->min_track_len >= 1);
    if (cfg->is_fisheye && cfg->k3 != 0.f) {    // the reference would hand cv::fisheye five coefficients: OpenCV asserts D.total() == 4
        set_error("rvio_tracker_create", "Camera.Fisheye with Camera.k3 != 0: cv::fisheye::undistortPoints takes exactly 4 coefficients (Tracker.cc:56-61,119)");
        return RVIO_ERR_ARG;
    }
    if (cfg->n_features > 4096) {               // the RANSAC pair sampler keeps its 'used' set as a 4096-bit mask in shared memory
        set_error("rvio_tracker_create", "Tracker.nFeatures > 4096 is not supported");
        return RVIO_ERR_CAPACITY;
    }
    int rc = require_b200(device);
    if (rc != RVIO_OK) return rc;
    rvio_tracker* t = new (std::nothrow) rvio_tracker();
    if (!t) return RVIO_ERR_CUDA;
    t->cfg = *cfg; t->device = device;
    t->W = cfg->width; t->H = cfg->height; t->F = cfg->n_features;
    t->Fu = (cfg->n_features + 1) / 2;                     // ceil(.5*nFeatures), Tracker.cc:74
    t->Lmax = cfg->max_track_len; t->Lmin = cfg->min_track_len;
    t->first = true; t->n_track = 0; t->last_n = 0; t->frame_open = false; t->shard_open = false; t->cur_idx = 0;
    RVIO_CUDA_TRY(cudaSetDevice(device));
    RVIO_CUDA_TRY(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
    RVIO_CUDA_TRY(cudaEventCreateWithFlags(&t->ev_level0, cudaEventDisableTiming));
    if ((rc = detector_create(&t->det, t->W, t->H, t->F)) != RVIO_OK) return rc;
    // camera (float-rounded values widened to double: Tracker.cc:39-61)
    CamParams& c = t->cam;
    c.fx = cfg->fx; c.fy = cfg->fy; c.cx = cfg->cx; c.cy = cfg->cy;
    c.ifx = 1. / c.fx; c.ify = 1. / c.fy;
    c.k1 = cfg->k1; c.k2 = cfg->k2; c.p1 = cfg->p1; c.p2 = cfg->p2; c.k3 = cfg->k3; c.fisheye = cfg->is_fisheye ? 1 : 0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) t->Ric[3 * i + j] = cfg->T_BC0[4 * i + j];   // Ransac.cc:41-46
    // CLAHE geometry (OpenCV CLAHE_Impl::apply)
    int ew = t->W, eh = t->H;
    if (t->W % 5 != 0 || t->H % 5 != 0) { ew = t->W + (5 - t->W % 5); eh = t->H + (5 - t->H % 5); }
    t->tw = ew / 5; t->th = eh / 5;
    const int area = t->tw * t->th;
    t->clip = (int)(3.0 * area / 256);
    if (t->clip < 1) t->clip = 1;
    t->lut_scale = (float)255 / area;
    t->inv_tw = 1.0f / t->tw; t->inv_th = 1.0f / t->th;

    // device buffers
    t->in_pitch = ((size_t)t->W * 4 + 255) / 256 * 256;
    t->gray_pitch = ((size_t)t->W + 255) / 256 * 256;
    if ((rc = dalloc(t, &t->d_in, t->in_pitch * t->H)) != RVIO_OK) return rc;
    if ((rc = dalloc(t, &t->d_gray, t->gray_pitch * t->H)) != RVIO_OK) return rc;
    if ((rc = dalloc(t, &t->d_lut, 25 * 256)) != RVIO_OK) return rc;
    if ((rc = build_pyramid_layout(t, 0)) != RVIO_OK) return rc;
    if ((rc = build_pyramid_layout(t, 1)) != RVIO_OK) return rc;
    if ((rc = build_tensor_maps(t, 0)) != RVIO_OK) return rc;
    if ((rc = build_tensor_maps(t, 1)) != RVIO_OK) return rc;
    TrackerBuffers& B = t->B;
    B.F = t->F; B.Fu = t->Fu; B.Lmax = t->Lmax; B.Lmin = t->Lmin; B.hist_cap = t->Lmax + 1;
    const size_t F = (size_t)t->F;
#define A(p, n) if ((rc = dalloc(t, &(p), (n))) != RVIO_OK) return rc
    A(B.feats, F); A(B.slots, F); A(B.pts1, F); A(B.feats_new, F); A(B.slots_new, F); A(B.pts1_new, F);
    A(B.lk, F + 64); A(B.un, F + 64); A(B.status, F + 64); A(B.flags, F);     // + 64: equal-sized shards of an all-gather may overhang F
    A(B.hist, F * B.hist_cap); A(B.hist_head, F); A(B.hist_len, F); A(B.freeq, F + 1);
    A(B.up_types, (size_t)t->Fu + 1); A(B.up_off, (size_t)t->Fu + 2); A(B.up_xy, ((size_t)t->Fu + 1) * t->Lmax);
    A(B.cand, 2 * F); A(B.two_points, 32); A(B.n_inliers, 16); A(B.hyp, 16 * 9); A(B.sc, 1);
    t->imu_cap = 512;
    A(t->d_R, 16);
    A(t->d_px_in, F);
#undef A
    t->h_img_bytes = (size_t)t->W * t->H * 4;
    RVIO_CUDA_TRY(cudaMallocHost((void**)&t->h_img, t->h_img_bytes));
    RVIO_CUDA_TRY(cudaMallocHost((void**)&t->h_R, sizeof(double) * 16));
    RVIO_CUDA_TRY(cudaMallocHost((void**)&t->h_px, sizeof(float) * 2 * F));
    RVIO_CUDA_TRY(cudaMallocHost((void**)&t->h_sc, sizeof(TrackerScalars)));
    // glibc rand() never seeded == srand(1): build the state on the host (random_r.c) and upload
    {
        TrackerScalars s;
        memset(&s, 0, sizeof s);
        int32_t* r = s.rng_r;
        r[0] = 1;
        for (int i = 1; i < 31; ++i) {
            long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
            long word = 16807 * lo - 2836 * hi;
            if (word < 0) word += 2147483647;
            r[i] = (int32_t)word;
        }
        int f = 3, b = 0;
        for (int i = 0; i < 310; ++i) {
            r[f] = (int32_t)((uint32_t)r[f] + (uint32_t)r[b]);
            if (++f >= 31) { f = 0; ++b; } else if (++b >= 31) b = 0;
        }
        s.rng_f = f; s.rng_b = b;
        *t->h_sc = s;
        RVIO_CUDA_TRY(cudaMemcpyAsync(B.sc, t->h_sc, sizeof s, cudaMemcpyHostToDevice, t->stream));
    }
    RVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    *out = t;
    return RVIO_OK;
}

extern "C" void rvio_tracker_destroy(rvio_tracker* t)
{
    if (!t) return;
    cudaSetDevice(t->device);
    cudaStreamSynchronize(t->stream);
    for (void* p : t->allocs) cudaFree(p);
    detector_destroy(&t->det); cudaEventDestroy(t->ev_level0);
    cudaFreeHost(t->h_img); cudaFreeHost(t->h_R); cudaFreeHost(t->h_px); cudaFreeHost(t->h_sc);
    cudaStreamDestroy(t->stream);
    delete t;
}

// Ransac::GetRotation (Ransac.cc:120-155): gyro integration over the frame's IMU samples, raw gyro (no bias removal),
// R = Rci * (prod_k dR_k) * Ric.  Depends on host data only, so it is evaluated here with the host libm.
static void host_gyro_rotation(const double* Ric, const double* imu, int n_imu, double small_angle, double* R)
{
    auto mul = [](const double* A, const double* B, double* C) {
        double T[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
        for (int i = 0; i < 9; ++i) C[i] = T[i];
    };
    double acc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int k = 0; k < n_imu; ++k) {
        const double* wm = imu + 8 * k;
        const double dt = imu[8 * k + 7];
        const double w1 = sqrt(wm[0] * wm[0] + wm[1] * wm[1] + wm[2] * wm[2]);
        const double wdt = w1 * dt;
        const double wx[9] = {0, -wm[2], wm[1], wm[2], 0, -wm[0], -wm[1], wm[0], 0};
        double wx2[9], dR[9];
        mul(wx, wx, wx2);
        double c1, c2;
        if (w1 < small_angle) { c1 = dt; c2 = .5 * (dt * dt); }
        else { c1 = sin(wdt) / w1; c2 = (1 - cos(wdt)) / (w1 * w1); }
        for (int i = 0; i < 9; ++i) dR[i] = ((i == 0 || i == 4 || i == 8) ? 1.0 : 0.0) - c1 * wx[i] + c2 * wx2[i];
        mul(dR, acc, acc);
    }
    double Rci[9], T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rci[3 * i + j] = Ric[3 * j + i];
    mul(Rci, acc, T);
    mul(T, Ric, R);
}

// Enqueues one frame (no synchronisation).  dev_count: the number of features to track is read from the device scalars
// (sc->n_new of the previous frame) instead of the host mirror, which makes the launch parameters frame-invariant (the
// fused pipeline replays this sequence as a CUDA graph).
static int tracker_enqueue(rvio_tracker* t, const uint8_t* gray_dev, int gray_pitch, const double* imu, int n_imu, bool dev_count = false,
                           int shard_rank = 0, int shard_world = 1, bool finish = true)
{
    RVIO_ARG_CHECK(n_imu >= 0 && n_imu <= t->imu_cap);
    cudaStream_t s = t->stream;
    const Pyramid& cur = t->pyr[t->cur_idx];
    const Pyramid& prev = t->pyr[1 - t->cur_idx];
    const dim3 blk(256), grd(div_up(t->W, 256), t->H);
    const bool will_track = !t->first && t->n_track > 0;
    if (will_track) {
        // Ransac::GetRotation on the host (libm) into a pinned block that k_ransac_bookkeep reads directly (nine doubles over
        // PCIe at the start of the kernel, behind its candidate compaction): no copy node on the frame's critical path
        host_gyro_rotation(t->Ric, imu, n_imu, t->cfg.small_angle, t->h_R);
    }
    // Tracker.cc:198-202
    if (t->cfg.enable_equalizer) {
        RVIO_LAUNCH(k_clahe_lut, 25, 1024, 0, s, gray_dev, gray_pitch, t->W, t->H, t->tw, t->th, t->clip, t->lut_scale, t->d_lut);
        RVIO_LAUNCH_PDL(true, k_clahe_apply, grd, blk, 0, s, gray_dev, gray_pitch, t->d_lut, t->inv_tw, t->inv_th, cur.lv[0]);
    } else {
        RVIO_LAUNCH(k_copy_level0, grd, blk, 0, s, gray_dev, gray_pitch, cur.lv[0]);
    }
    RVIO_ENQ(cudaEventRecord(t->ev_level0, s));            // what the detector needs (Tracker.cc:207,350 pass the equalised image)
    if (cur.levels == 4) {
        const dim3 g(div_up(cur.lv[3].w, 8), div_up(cur.lv[3].h, 8));
        RVIO_LAUNCH_PDL(true, k_pyr_down3, g, blk, 0, s, cur.lv[0], cur.lv[1], cur.lv[2], cur.lv[3]);
    } else {
        for (int l = 1; l < cur.levels; ++l) {
            const dim3 g(div_up(cur.lv[l].w, 256), cur.lv[l].h);
            RVIO_LAUNCH_PDL(true, k_pyr_down, g, blk, 0, s, cur.lv[l - 1], cur.lv[l]);
        }
    }
    t->frame_open = true;
    if (t->first) { RVIO_ENQ(cudaGetLastError()); return RVIO_FIRST_IMAGE; }
    const int n = t->n_track;
    t->last_n = n;
    if (n == 0) { t->frame_open = false; RVIO_ENQ(cudaGetLastError()); return RVIO_NO_FEATURES; }

    LKParams lp;
    lp.prev = prev; lp.cur = cur; lp.feats = t->B.feats; lp.out = t->B.lk; lp.status = t->B.status;
    lp.n = dev_count ? t->F : n; lp.n_dev = dev_count ? &t->B.sc->n_new : nullptr;
    lp.un = t->B.un; lp.cam = t->cam; lp.max_iter = 30; lp.eps_sq_f = 0.f; lp.eps_sq = 1e-2 * 1e-2;
    lp.min_eig_thr = 1e-3f;
    lp.first = 0; lp.last = 0x7fffffff;
    int n_lk = lp.n;
    if (shard_world > 1) {                                   // equal shards of ceil(F / world) feature indices
        const int S = div_up(t->F, shard_world);
        lp.first = shard_rank * S; lp.last = lp.first + S;
        n_lk = S;
    }
    RVIO_LAUNCH_PDL(true, k_lk, div_up(n_lk, kLKWarps), kLKWarps * 32, 0, s, lp, t->tmaps[t->cur_idx]);
    if (!finish) { RVIO_ENQ(cudaGetLastError()); return RVIO_OK; }
    RansacParams rp;
    rp.B = t->B; rp.n = n; rp.n_dev = lp.n_dev; rp.use_sampson = t->cfg.use_sampson;
    rp.thr = t->cfg.inlier_thr; rp.small_angle = t->cfg.small_angle; rp.R = t->h_R;
    RVIO_LAUNCH_PDL(true, k_ransac_bookkeep, 1, 256, 0, s, rp);
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}

static int tracker_run(rvio_tracker* t, const uint8_t* gray_dev, int gray_pitch, const double* imu, int n_imu)
{
    const int rc = tracker_enqueue(t, gray_dev, gray_pitch, imu, n_imu);
    if (rc != RVIO_OK) return rc;
    return sync_scalars(t);
}

static int upload_image(rvio_tracker* t, const uint8_t* img, int width, int height, int stride_bytes, int channels)
{
    const size_t row = (size_t)width * channels;
    for (int y = 0; y < height; ++y) memcpy(t->h_img + (size_t)y * row, img + (size_t)y * stride_bytes, row);
    cudaStream_t s = t->stream;
    if (channels == 1) {
        RVIO_ENQ(cudaMemcpy2DAsync(t->d_gray, t->gray_pitch, t->h_img, row, row, height, cudaMemcpyHostToDevice, s));
    } else {
        RVIO_ENQ(cudaMemcpy2DAsync(t->d_in, t->in_pitch, t->h_img, row, row, height, cudaMemcpyHostToDevice, s));
        RVIO_LAUNCH(k_gray, dim3(div_up(width, 256), height), 256, 0, s, t->d_in, (int)t->in_pitch, channels,
                    t->cfg.is_rgb, t->d_gray, (int)t->gray_pitch, width, height);
    }
    return RVIO_OK;
}

extern "C" int rvio_tracker_track(rvio_tracker* t, const uint8_t* img, int width, int height, int stride_bytes,
                                  int channels, const double* imu, int n_imu)
{
    RVIO_ARG_CHECK(t && img);
    RVIO_ARG_CHECK(width == t->W && height == t->H);
    RVIO_ARG_CHECK(channels == 1 || channels == 3 || channels == 4);
    RVIO_ARG_CHECK(stride_bytes >= width * channels);
    RVIO_ARG_CHECK(n_imu == 0 || imu);
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    const int rcu = upload_image(t, img, width, height, stride_bytes, channels);
    if (rcu != RVIO_OK) return rcu;
    return tracker_run(t, t->d_gray, (int)t->gray_pitch, imu, n_imu);
}

extern "C" int rvio_tracker_track_dev(rvio_tracker* t, const uint8_t* img_dev, int pitch_bytes, const double* imu, int n_imu)
{
    RVIO_ARG_CHECK(t && img_dev && pitch_bytes >= t->W);
    RVIO_ARG_CHECK(n_imu == 0 || imu);
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    return tracker_run(t, img_dev, pitch_bytes, imu, n_imu);
}

// ---- feature-sharded tracking: LK on this rank's share of the feature indices, the caller all-gathers the
//      per-feature LK results between _begin and _finish, RANSAC + bookkeeping then run replicated on the complete arrays.
extern "C" int rvio_tracker_track_begin(rvio_tracker* t, const uint8_t* img, int width, int height, int stride_bytes,
                                        int channels, const double* imu, int n_imu, int rank, int world)
{
    RVIO_ARG_CHECK(t && img);
    RVIO_ARG_CHECK(width == t->W && height == t->H);
    RVIO_ARG_CHECK(channels == 1 || channels == 3 || channels == 4);
    RVIO_ARG_CHECK(stride_bytes >= width * channels);
    RVIO_ARG_CHECK(n_imu == 0 || imu);
    RVIO_ARG_CHECK(world >= 1 && world <= 64 && rank >= 0 && rank < world);
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    const int rcu = upload_image(t, img, width, height, stride_bytes, channels);
    if (rcu != RVIO_OK) return rcu;
    const int rc = tracker_enqueue(t, t->d_gray, (int)t->gray_pitch, imu, n_imu, false, rank, world, false);
    if (rc < 0) return rc;
    t->shard_open = rc == RVIO_OK;
    RVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    return rc;
}

extern "C" int rvio_tracker_lk_results(rvio_tracker* t, int world, void** lk_px_dev, void** undist_dev, void** status_dev, int* shard)
{
    RVIO_ARG_CHECK(t && world >= 1 && world <= 64);
    if (lk_px_dev) *lk_px_dev = t->B.lk;
    if (undist_dev) *undist_dev = t->B.un;
    if (status_dev) *status_dev = t->B.status;
    if (shard) *shard = div_up(t->F, world);
    return RVIO_OK;
}

extern "C" int rvio_tracker_track_finish(rvio_tracker* t)
{
    RVIO_ARG_CHECK(t);
    if (!t->shard_open) { set_error("rvio_tracker_track_finish", "no sharded frame open"); return RVIO_ERR_STATE; }
    t->shard_open = false;
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    RansacParams rp;
    rp.B = t->B; rp.n = t->last_n; rp.n_dev = nullptr; rp.use_sampson = t->cfg.use_sampson;
    rp.thr = t->cfg.inlier_thr; rp.small_angle = t->cfg.small_angle; rp.R = t->h_R;
    RVIO_LAUNCH(k_ransac_bookkeep, 1, 256, 0, t->stream, rp);
    RVIO_CUDA_TRY(cudaGetLastError());
    return sync_scalars(t);
}

// FeatureDetector::DetectWithSubPix(equalised current image, nFeatures, s) on the device (FeatureDetector.cc:55-75).
extern "C" int rvio_tracker_detect(rvio_tracker* t, int s_factor, float min_dist, float quality, float* xy_out, int* n_out)
{
    RVIO_ARG_CHECK(t && xy_out && n_out && (s_factor == 1 || s_factor == 2));
    if (!t->frame_open) { set_error("rvio_tracker_detect", "no open frame (call rvio_tracker_track first)"); return RVIO_ERR_STATE; }
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    const int rc = detector_enqueue(&t->det, t->stream, t->pyr[t->cur_idx].lv[0], s_factor, min_dist, quality);
    if (rc != RVIO_OK) return rc;
    DetCtrl ctrl;
    RVIO_CUDA_TRY(cudaMemcpyAsync(&ctrl, t->det.ctrl, sizeof ctrl, cudaMemcpyDeviceToHost, t->stream));
    RVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    if (ctrl.overflow) { set_error("rvio_tracker_detect", "more local maxima than the detector keeps (or a grid cell overflowed)"); return RVIO_ERR_CAPACITY; }
    *n_out = ctrl.n_out;
    if (ctrl.n_out > 0) RVIO_CUDA_TRY(cudaMemcpy(xy_out, t->det.out, sizeof(float2) * ctrl.n_out, cudaMemcpyDeviceToHost));
    return RVIO_OK;
}

extern "C" int rvio_tracker_get_image(rvio_tracker* t, uint8_t* out, int out_stride)
{
    RVIO_ARG_CHECK(t && out && out_stride >= t->W);
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    const PyrLevel& L = t->pyr[t->cur_idx].lv[0];
    RVIO_CUDA_TRY(cudaMemcpy2DAsync(t->h_img, t->W, L.base, L.pitch, t->W, t->H, cudaMemcpyDeviceToHost, t->stream));
    RVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    for (int y = 0; y < t->H; ++y) memcpy(out + (size_t)y * out_stride, t->h_img + (size_t)y * t->W, t->W);
    return RVIO_OK;
}

extern "C" int rvio_tracker_n_free(rvio_tracker* t, int* n_free)
{
    RVIO_ARG_CHECK(t && n_free);
    *n_free = t->first ? t->F : t->h_sc->fq_n;
    return RVIO_OK;
}

extern "C" int rvio_tracker_get_tracked_px(rvio_tracker* t, float* xy, int* n)
{
    RVIO_ARG_CHECK(t && n);
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    const int m = t->first ? 0 : t->h_sc->n_new;
    *n = m;
    if (m > 0 && xy) {
        RVIO_CUDA_TRY(cudaMemcpyAsync(t->h_px, t->B.feats_new, sizeof(float2) * m, cudaMemcpyDeviceToHost, t->stream));
        RVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
        memcpy(xy, t->h_px, sizeof(float) * 2 * m);
    }
    return RVIO_OK;
}

extern "C" int rvio_tracker_seed(rvio_tracker* t, const float* px, int n)
{
    RVIO_ARG_CHECK(t && (n == 0 || px) && n >= 0);
    if (!t->first || !t->frame_open) { set_error("rvio_tracker_seed", "not expecting a seed"); return RVIO_ERR_STATE; }
    if (n == 0) return RVIO_OK;                       // Tracker.cc:209-213: stays "first image"
    if (n > t->F) n = t->F;
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    memcpy(t->h_px, px, sizeof(float) * 2 * n);
    RVIO_CUDA_TRY(cudaMemcpyAsync(t->d_px_in, t->h_px, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, t->stream));
    RVIO_LAUNCH(k_seed, div_up(t->F, 256), 256, 0, t->stream, t->B, t->d_px_in, n, (const int*)nullptr, t->cam);
    RVIO_CUDA_TRY(cudaGetLastError());
    t->first = false;
    return sync_scalars(t);
}

extern "C" int rvio_tracker_refill(rvio_tracker* t, const float* px, int n, int* n_used)
{
    RVIO_ARG_CHECK(t && (n == 0 || px) && n >= 0);
    if (t->first || !t->frame_open) { set_error("rvio_tracker_refill", "no open frame"); return RVIO_ERR_STATE; }
    int use = n < t->h_sc->fq_n ? n : t->h_sc->fq_n;
    if (n_used) *n_used = use;
    if (use <= 0) return RVIO_OK;
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    memcpy(t->h_px, px, sizeof(float) * 2 * use);
    RVIO_CUDA_TRY(cudaMemcpyAsync(t->d_px_in, t->h_px, sizeof(float) * 2 * use, cudaMemcpyHostToDevice, t->stream));
    RVIO_LAUNCH(k_refill, div_up(use, 128), 128, 0, t->stream, t->B, t->d_px_in, use, t->cam);
    RVIO_LAUNCH(k_refill_commit, 1, 1, 0, t->stream, t->B, use);
    RVIO_CUDA_TRY(cudaGetLastError());
    return sync_scalars(t);
}

extern "C" int rvio_tracker_commit(rvio_tracker* t)
{
    RVIO_ARG_CHECK(t);
    if (!t->frame_open) { set_error("rvio_tracker_commit", "no open frame"); return RVIO_ERR_STATE; }
    t->frame_open = false;
    if (t->first) return RVIO_OK;                      // nothing seeded: image is not kept (Tracker.cc:209-213)
    // Tracker.cc:389-395
    TrackerBuffers& B = t->B;
    std::swap(B.feats, B.feats_new); std::swap(B.slots, B.slots_new); std::swap(B.pts1, B.pts1_new);
    t->n_track = t->h_sc->n_new;
    t->cur_idx = 1 - t->cur_idx;                       // current pyramid becomes mLastImage's pyramid
    return RVIO_OK;
}

extern "C" int rvio_tracker_get_update_count(rvio_tracker* t, int* n_feat, int* n_meas)
{
    RVIO_ARG_CHECK(t && n_feat);
    *n_feat = t->h_sc->n_up;
    if (n_meas) *n_meas = t->h_sc->n_meas;
    return RVIO_OK;
}

extern "C" int rvio_tracker_get_update_lists(rvio_tracker* t, uint8_t* types, int32_t* offsets, float* xy)
{
    RVIO_ARG_CHECK(t && types && offsets && xy);
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    const int nu = t->h_sc->n_up, nm = t->h_sc->n_meas;
    offsets[0] = 0;
    if (nu == 0) return RVIO_OK;
    RVIO_CUDA_TRY(cudaMemcpyAsync(types, t->B.up_types, nu, cudaMemcpyDeviceToHost, t->stream));
    RVIO_CUDA_TRY(cudaMemcpyAsync(offsets, t->B.up_off, sizeof(int32_t) * (nu + 1), cudaMemcpyDeviceToHost, t->stream));
    RVIO_CUDA_TRY(cudaMemcpyAsync(xy, t->B.up_xy, sizeof(float2) * nm, cudaMemcpyDeviceToHost, t->stream));
    RVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    return RVIO_OK;
}

extern "C" int rvio_tracker_get_debug(rvio_tracker* t, int* n, uint8_t* lk_status, uint8_t* inlier_flags,
                                      float* lk_px, float* undist, int32_t* slots)
{
    RVIO_ARG_CHECK(t && n);
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    const int m = t->last_n;
    *n = m;
    if (m == 0) return RVIO_OK;
    cudaStream_t s = t->stream;
    // note: after commit() the pre-frame slot list lives in slots_new (buffers were swapped)
    const int* slots_src = t->frame_open ? t->B.slots : t->B.slots_new;
    if (lk_status) RVIO_CUDA_TRY(cudaMemcpyAsync(lk_status, t->B.status, m, cudaMemcpyDeviceToHost, s));
    if (inlier_flags) RVIO_CUDA_TRY(cudaMemcpyAsync(inlier_flags, t->B.flags, m, cudaMemcpyDeviceToHost, s));
    if (lk_px) RVIO_CUDA_TRY(cudaMemcpyAsync(lk_px, t->B.lk, sizeof(float2) * m, cudaMemcpyDeviceToHost, s));
    if (undist) RVIO_CUDA_TRY(cudaMemcpyAsync(undist, t->B.un, sizeof(float2) * m, cudaMemcpyDeviceToHost, s));
    if (slots) RVIO_CUDA_TRY(cudaMemcpyAsync(slots, slots_src, sizeof(int) * m, cudaMemcpyDeviceToHost, s));
    RVIO_CUDA_TRY(cudaStreamSynchronize(s));
    return RVIO_OK;
}

extern "C" int rvio_tracker_get_ransac_debug(rvio_tracker* t, int32_t* two_points, int32_t* n_inliers,
                                             int32_t* winner, int32_t* n_candidates, double* hypotheses)
{
    RVIO_ARG_CHECK(t);
    RVIO_CUDA_TRY(cudaSetDevice(t->device));
    cudaStream_t s = t->stream;
    if (two_points) RVIO_CUDA_TRY(cudaMemcpyAsync(two_points, t->B.two_points, sizeof(int) * 32, cudaMemcpyDeviceToHost, s));
    if (n_inliers) RVIO_CUDA_TRY(cudaMemcpyAsync(n_inliers, t->B.n_inliers, sizeof(int) * 16, cudaMemcpyDeviceToHost, s));
    if (hypotheses) RVIO_CUDA_TRY(cudaMemcpyAsync(hypotheses, t->B.hyp, sizeof(double) * 144, cudaMemcpyDeviceToHost, s));
    RVIO_CUDA_TRY(cudaStreamSynchronize(s));
    if (winner) *winner = t->h_sc->winner;
    if (n_candidates) *n_candidates = t->h_sc->n_cand;
    return RVIO_OK;
}

extern "C" int rvio_tracker_get_pyramid(rvio_tracker* t, int which, int level, uint8_t* out, int* lw, int* lh)
{
    RVIO_ARG_CHECK(t && (which == 0 || which == 1));
    // frame_open: cur_idx is the image being processed; after commit the roles are swapped
    const int idx = which == 0 ? t->cur_idx : 1 - t->cur_idx;
    const Pyramid& P = t->pyr[idx];
    RVIO_ARG_CHECK(level >= 0 && level < P.levels);
    const PyrLevel& L = P.lv[level];
    if (lw) *lw = L.w;
    if (lh) *lh = L.h;
    if (out) {
        RVIO_CUDA_TRY(cudaSetDevice(t->device));
        RVIO_CUDA_TRY(cudaMemcpy2DAsync(out, L.w, L.base, L.pitch, L.w, L.h, cudaMemcpyDeviceToHost, t->stream));
        RVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    }
    return RVIO_OK;
}

extern "C" void* rvio_tracker_stream(rvio_tracker* t) { return t ? (void*)t->stream : nullptr; }

// accessors for the updater's fused path and for vio.cu (same shared library)
namespace rvio {
int tracker_enqueue_frame_host(rvio_tracker* t, const uint8_t* img, int w, int h, int stride, int ch, const double* imu, int n_imu)
{
    RVIO_ARG_CHECK(t && img && w == t->W && h == t->H && (ch == 1 || ch == 3 || ch == 4) && stride >= w * ch);
    const int rc = upload_image(t, img, w, h, stride, ch);
    if (rc != RVIO_OK) return rc;
    return tracker_enqueue(t, t->d_gray, (int)t->gray_pitch, imu, n_imu, true);
}
int tracker_enqueue_frame_dev(rvio_tracker* t, const uint8_t* img_dev, int pitch, const double* imu, int n_imu)
{
    return tracker_enqueue(t, img_dev, pitch, imu, n_imu, true);
}
// Frame already staged in the tracker's own gray buffer (see tracker_gray): frame-invariant launch parameters.
int tracker_enqueue_frame_staged(rvio_tracker* t, const double* imu, int n_imu)
{
    return tracker_enqueue(t, t->d_gray, (int)t->gray_pitch, imu, n_imu, true);
}
// Feature-sharded frame of the fused path: image pipeline + LK for this rank's share of the feature indices (device-side
// feature count); the caller all-gathers the LK arrays on the same stream and then calls tracker_enqueue_ransac.
int tracker_enqueue_frame_sharded(rvio_tracker* t, const uint8_t* img_host, int w, int h, int stride, int ch, const uint8_t* img_dev, int pitch,
                                  bool staged, const double* imu, int n_imu, int rank, int world)
{
    if (!staged && !img_dev) {
        RVIO_ARG_CHECK(img_host && w == t->W && h == t->H && (ch == 1 || ch == 3 || ch == 4) && stride >= w * ch);
        const int rc = upload_image(t, img_host, w, h, stride, ch);
        if (rc != RVIO_OK) return rc;
    }
    const uint8_t* g = (staged || !img_dev) ? t->d_gray : img_dev;
    const int gp = (staged || !img_dev) ? (int)t->gray_pitch : pitch;
    return tracker_enqueue(t, g, gp, imu, n_imu, true, rank, world, false);
}
int tracker_enqueue_ransac(rvio_tracker* t)
{
    RansacParams rp;
    rp.B = t->B; rp.n = t->last_n; rp.n_dev = &t->B.sc->n_new; rp.use_sampson = t->cfg.use_sampson;
    rp.thr = t->cfg.inlier_thr; rp.small_angle = t->cfg.small_angle; rp.R = t->h_R;
    RVIO_LAUNCH(k_ransac_bookkeep, 1, 256, 0, t->stream, rp);
    RVIO_ENQ(cudaGetLastError());
    return RVIO_OK;
}
int tracker_shard_size(const rvio_tracker* t, int world) { return div_up(t->F, world); }
uint8_t* tracker_gray(rvio_tracker* t, size_t* pitch) { *pitch = t->gray_pitch; return t->d_gray; }
int tracker_enqueue_seed_dev(rvio_tracker* t, const float2* px_dev, int n, const int* n_dev)
{
    if (n > t->F) n = t->F;
    RVIO_LAUNCH(k_seed, div_up(t->F, 256), 256, 0, t->stream, t->B, px_dev, n, n_dev, t->cam);
    RVIO_ENQ(cudaGetLastError());
    t->first = false;
    return RVIO_OK;
}
int tracker_sync(rvio_tracker* t) { return sync_scalars(t); }
// The two halves of tracker_sync: the device->host copy of the scalars is part of the frame's stream work (replayable),
// the wait is not.
int tracker_enqueue_scalars(rvio_tracker* t)
{
    RVIO_ENQ(cudaMemcpyAsync(t->h_sc, t->B.sc, sizeof(TrackerScalars), cudaMemcpyDeviceToHost, t->stream));
    return RVIO_OK;
}
int tracker_wait(rvio_tracker* t)
{
    RVIO_CUDA_TRY(cudaStreamSynchronize(t->stream));
    return RVIO_OK;
}
int tracker_parity(const rvio_tracker* t) { return t->cur_idx; }
Detector* tracker_detector(rvio_tracker* t) { return &t->det; }
const PyrLevel* tracker_level0(const rvio_tracker* t) { return &t->pyr[t->cur_idx].lv[0]; }
cudaEvent_t tracker_level0_event(const rvio_tracker* t) { return t->ev_level0; }
void tracker_set_first(rvio_tracker* t, bool first) { t->first = first; }
bool tracker_is_first(const rvio_tracker* t) { return t->first; }
const CamParams* tracker_cam(const rvio_tracker* t) { return &t->cam; }
int tracker_n_track(const rvio_tracker* t) { return t->n_track; }
const TrackerScalars* tracker_host_scalars(const rvio_tracker* t) { return t->h_sc; }
const TrackerBuffers* tracker_buffers(const rvio_tracker* t) { return &t->B; }
int tracker_update_counts(const rvio_tracker* t, int* n_meas) { if (n_meas) *n_meas = t->h_sc->n_meas; return t->h_sc->n_up; }
int tracker_device(const rvio_tracker* t) { return t->device; }
cudaStream_t tracker_stream(const rvio_tracker* t) { return t->stream; }
}

#ifdef RVIO_B200_PHASE_CLOCKS
extern "C" int rvio_b200_trk_clocks(long long* out, int n)
{
    cudaDeviceSynchronize();
    return (int)cudaMemcpyFromSymbol(out, rvio::g_trk_clk, sizeof(long long) * (n < 32 ? n : 32));
}
#endif
