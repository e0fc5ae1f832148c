// common.cuh -- shared host/device helpers for librvio_b200.so (sm_100a only, no CPU fallback).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <atomic>

#include "../../include/rvio_b200.h"

namespace rvio {

extern thread_local char g_last_error[512];
extern std::atomic<uint64_t> g_kernel_launches;

inline void set_error(const char* where, const char* what)
{
    snprintf(g_last_error, sizeof g_last_error, "%s: %s", where, what);
}

#define RVIO_CUDA_TRY(expr)                                                             \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            rvio::set_error(#expr, cudaGetErrorString(_e));                             \
            return RVIO_ERR_CUDA;                                                       \
        }                                                                               \
    } while (0)

#define RVIO_ARG_CHECK(cond)                                                            \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            rvio::set_error("argument check failed", #cond);                            \
            return RVIO_ERR_ARG;                                                        \
        }                                                                               \
    } while (0)

extern std::atomic<int> g_profile_on;
void profile_begin(const char* name, cudaStream_t s);
void profile_end(cudaStream_t s);

// Frame-graph replay (vio.cu): the per-frame enqueue code runs once under stream capture to build a CUDA graph; on later
// frames with the same signature it runs again with t_replay set, which turns every stream operation (RVIO_LAUNCH,
// RVIO_ENQ) into a no-op so that only the host-side state advances, and the instantiated graph is launched instead.
extern thread_local bool t_replay;

#define RVIO_ENQ(expr)                                                                  \
    do {                                                                                \
        if (!rvio::t_replay) RVIO_CUDA_TRY(expr);                                       \
    } while (0)

// Counts every kernel this library launches (bench.py reports it as gpu_launches); with profiling enabled
// (rvio_b200_profile) each launch is bracketed by CUDA events on its own stream.
#define RVIO_LAUNCH(kernel, grid, block, smem, stream, ...)                             \
    do {                                                                                \
        rvio::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);                \
        if (rvio::t_replay) break;                                                      \
        const bool _prof = rvio::g_profile_on.load(std::memory_order_relaxed) != 0;     \
        if (_prof) rvio::profile_begin(#kernel, (stream));                              \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                     \
        if (_prof) rvio::profile_end((stream));                                         \
    } while (0)

// Programmatic dependent launch (sm_90+): a kernel launched with RVIO_LAUNCH_PDL directly behind another kernel of the same
// stream may become resident while its predecessor is still running (the predecessor calls pdl_trigger() first thing) and
// blocks in pdl_wait() -- its first statement -- until the predecessor's grid has completed and its writes are visible.  The
// stream order is therefore unchanged; what disappears is the launch gap between two short dependent kernels.  Only used where
// the predecessor in the stream is known to be a kernel.  rvio_b200_pdl() / RVIO_B200_PDL select it process-wide (frame graphs
// are keyed by it).
extern std::atomic<int> g_pdl_on;

template <typename... KArgs, typename... Args>
static inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<Args&&>(args)...);      // errors surface through the cudaGetLastError() that ends every enqueue
}

#define RVIO_LAUNCH_PDL(allowed, kernel, grid, block, smem, stream, ...)                \
    do {                                                                                \
        rvio::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);                \
        if (rvio::t_replay) break;                                                      \
        const bool _prof = rvio::g_profile_on.load(std::memory_order_relaxed) != 0;     \
        if (_prof) rvio::profile_begin(#kernel, (stream));                              \
        if ((allowed) && !_prof && rvio::g_pdl_on.load(std::memory_order_relaxed) != 0) \
            rvio::launch_pdl(kernel, dim3(grid), dim3(block), (smem), (stream), __VA_ARGS__); \
        else                                                                            \
            kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                 \
        if (_prof) rvio::profile_end((stream));                                         \
    } while (0)

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// Verifies that `device` is a usable sm_100 part; the product has no other path.
int require_b200(int device);

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

}  // namespace rvio
