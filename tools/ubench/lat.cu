// Micro-benchmarks for the latency-bound single-CTA kernels (dependent-issue latencies on sm_100a).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat lat.cu ; run: ./lat
#include <cstdio>
#include <cuda_runtime.h>

#define N_IT 512

template <int MODE>
__global__ void k_chain(double* out, long long* cyc, double seed, int iseed)
{
    __shared__ int s_ptr[256];
    __shared__ double s_d[64];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_ptr[i] = (i * 33 + 7) & 255;
    for (int i = threadIdx.x; i < 64; i += blockDim.x) s_d[i] = seed + i;
    __syncthreads();
    double x = seed + threadIdx.x, y = seed * 0.5;
    float xf = (float)seed + threadIdx.x;
    int xi = iseed + threadIdx.x;
    unsigned xu = iseed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N_IT; ++i) {
        if (MODE == 0) x = fma(x, y, 1.0);                       // DFMA
        if (MODE == 1) x = x + y;                                // DADD
        if (MODE == 2) x = x * y;                                // DMUL
        if (MODE == 3) xf = fmaf(xf, 1.0001f, 1.0f);             // FFMA
        if (MODE == 4) xi = xi * 3 + 1;                          // IMAD
        if (MODE == 5) xi = s_ptr[xi & 255];                     // LDS chase
        if (MODE == 6) xu = __reduce_max_sync(0xffffffffu, xu + 1);   // REDUX
        if (MODE == 7) xi = __shfl_sync(0xffffffffu, xi + 1, (xi & 31));   // SHFL
        if (MODE == 8) { xf = (float)x; x = (double)xf + 1.0; }  // F2F both ways + DADD
        if (MODE == 9) { asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(xf) : "f"(xf)); }
        if (MODE == 10) __syncthreads();                         // CTA barrier
        if (MODE == 11) { s_d[threadIdx.x & 63] = x; __syncwarp(); x = s_d[(threadIdx.x + 1) & 63] + 1.0; __syncwarp(); }   // STS->LDS round trip + DADD
        if (MODE == 12) x = 1.0 / x + 2.0;                       // DDIV + DADD
        if (MODE == 13) xu = max(xu + 1, (unsigned)xi) ^ 5u;     // IADD+MAX+LOP chain (3 dependent ALU)
        if (MODE == 14) { asm volatile("bar.sync 1, 96;" ::: "memory"); }
        if (MODE == 15) x = rsqrt(x) + 2.0;                      // rsqrt(double) + DADD
        if (MODE == 16) { double c1 = y; asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(x), "+d"(c1) : "d"(y), "d"(y)); }   // dependent DMMA
        if (MODE == 17) { double v = x; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); x = v * 1e-3; }   // warp sum of doubles
        if (MODE == 18) x = sqrt(x) + 2.0;                       // sqrt(double) + DADD
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = x + xf + xi + xu;
}

// throughput: each thread has K independent DFMA chains
template <int K>
__global__ void k_tp(double* out, long long* cyc, double seed)
{
    double x[K];
    for (int k = 0; k < K; ++k) x[k] = seed + k + threadIdx.x;
    const double y = seed * 0.5;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N_IT; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k) x[k] = fma(x[k], y, 1.0);
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
    for (int k = 0; k < K; ++k) s += x[k];
    out[threadIdx.x] = s;
}

template <int K>
__global__ void k_tp_dmma(double* out, long long* cyc, double seed)
{
    double x[K][2];
    for (int k = 0; k < K; ++k) { x[k][0] = seed + k; x[k][1] = seed - k; }
    const double y = seed * 0.5;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N_IT; ++i)
#pragma unroll
        for (int k = 0; k < K; ++k)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(x[k][0]), "+d"(x[k][1]) : "d"(y), "d"(y));
    long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
    for (int k = 0; k < K; ++k) s += x[k][0] + x[k][1];
    out[threadIdx.x] = s;
}

int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
    const char* names[] = {"DFMA", "DADD", "DMUL", "FFMA", "IMAD", "LDS chase", "REDUX", "SHFL", "F2F x2 + DADD", "MUFU.RCP", "bar.sync 0 (CTA)",
                           "STS+LDS+DADD+2 syncwarp", "DDIV+DADD", "3 ALU", "bar.sync 1,96",
                           "rsqrt(double)+DADD", "DMMA m8n8k4 dependent", "warp_sum(double)", "sqrt(double)+DADD"};
    long long h;
#define RUN(MODE, THREADS)                                                                  \
    for (int rep = 0; rep < 2; ++rep) k_chain<MODE><<<1, THREADS>>>(out, cyc, 1.000001, 3); \
    cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                \
    printf("%-28s threads %4d : %7.1f cycles / iteration\n", names[MODE], THREADS, (double)h / N_IT);
    RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(5, 32) RUN(6, 32) RUN(7, 32) RUN(8, 32) RUN(9, 32)
    RUN(10, 32) RUN(10, 96) RUN(10, 384) RUN(10, 1024) RUN(11, 32) RUN(12, 32) RUN(13, 32) RUN(14, 96)
    RUN(0, 128) RUN(0, 384) RUN(0, 256) RUN(15, 32) RUN(15, 256) RUN(16, 32) RUN(16, 256) RUN(17, 32) RUN(17, 256) RUN(18, 32) RUN(10, 256)
#define TP(K, THREADS)                                                                      \
    for (int rep = 0; rep < 2; ++rep) k_tp<K><<<1, THREADS>>>(out, cyc, 1.000001);           \
    cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                \
    printf("DFMA throughput K=%2d threads %4d : %7.2f cycles / DFMA-per-thread, %6.1f DFMA/clk/SM\n", K, THREADS, (double)h / N_IT / K, (double)THREADS * K * N_IT / h);
    TP(1, 32) TP(8, 32) TP(16, 32) TP(8, 128) TP(8, 384) TP(8, 1024) TP(16, 384)
#define TPM(K, THREADS)                                                                     \
    for (int rep = 0; rep < 2; ++rep) k_tp_dmma<K><<<1, THREADS>>>(out, cyc, 1.000001);      \
    cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                \
    printf("DMMA throughput K=%2d threads %4d : %7.2f cycles / DMMA-per-warp, %6.1f FMA/clk/SM\n", K, THREADS, (double)h / N_IT / K, (double)(THREADS / 32) * K * N_IT * 256 / h);
    TPM(1, 32) TPM(4, 32) TPM(8, 32) TPM(4, 128) TPM(4, 256) TPM(8, 256) TPM(4, 1024)
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
