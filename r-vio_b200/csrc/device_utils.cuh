// device_utils.cuh -- block-level reduce / scan helpers shared by the tracker and filter kernels.
#pragma once
#include <cuda_runtime.h>

namespace rvio {

__device__ __forceinline__ int block_reduce_sum(int v, int* sh /* >= 32 ints */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    int r = 0;
    if (wid == 0) {
        r = lane < nw ? sh[lane] : 0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
        if (lane == 0) sh[0] = r;
    }
    __syncthreads();
    r = sh[0];
    __syncthreads();
    return r;
}

// Exclusive scan of one int per thread over the block; returns the exclusive prefix, *total = block sum.
__device__ __forceinline__ int block_exscan(int v, int* sh /* >= 33 ints */, int* total)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sh[wid] = inc;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    if (wid == 0) {
        int w = lane < nw ? sh[lane] : 0;
        int winc = w;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        sh[lane] = winc - w;          // exclusive warp offsets
        if (lane == 31) sh[32] = winc;
    }
    __syncthreads();
    int res = sh[wid] + inc - v;
    *total = sh[32];
    __syncthreads();
    return res;
}


// same scan, name kept for call sites that run 1024-thread blocks
__device__ __forceinline__ int block_exscan_1024(int v, int* sh, int* total) { return block_exscan(v, sh, total); }

}  // namespace rvio
