// api_misc.cu -- library-wide state: error text, launch counter, device check, version.
#include "common.cuh"

namespace rvio {

thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_kernel_launches{0};

int require_b200(int device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        set_error("require_b200", e != cudaSuccess ? cudaGetErrorString(e) : "no CUDA device (this library has no CPU path)");
        return RVIO_ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("require_b200", "bad device ordinal"); return RVIO_ERR_ARG; }
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) { set_error("cudaGetDeviceProperties", cudaGetErrorString(e)); return RVIO_ERR_CUDA; }
    if (p.major != 10) {
        set_error("require_b200", "device is not sm_100-class; librvio_b200 is built for sm_100a only");
        return RVIO_ERR_CUDA;
    }
    return RVIO_OK;
}

}  // namespace rvio

extern "C" const char* rvio_b200_version(void) { return "rvio_b200 0.1 (sm_100a)"; }
extern "C" const char* rvio_b200_last_error(void) { return rvio::g_last_error; }
extern "C" uint64_t rvio_b200_kernel_launches(void) { return rvio::g_kernel_launches.load(); }
