// api_misc.cu -- library-wide state: error text, launch counter, device check, version.
#include "common.cuh"
#include <stdlib.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace rvio {

thread_local char g_last_error[512] = "";
std::atomic<uint64_t> g_kernel_launches{0};

std::atomic<int> g_profile_on{0};
thread_local bool t_replay = false;
static int pdl_default() { const char* e = getenv("RVIO_B200_PDL"); return e ? (atoi(e) != 0) : 1; }
std::atomic<int> g_pdl_on{pdl_default()};
namespace {
struct ProfRec { const char* name; cudaEvent_t e0, e1; };
std::mutex g_prof_mu;
std::vector<ProfRec> g_prof;
}
void profile_begin(const char* name, cudaStream_t s)
{
    ProfRec r;
    r.name = name;
    cudaEventCreate(&r.e0);
    cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, s);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
}
void profile_end(cudaStream_t s)
{
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof.empty()) cudaEventRecord(g_prof.back().e1, s);
}

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

extern "C" int rvio_b200_pdl(int enable)
{
    if (enable < 0) return rvio::g_pdl_on.load();
    return rvio::g_pdl_on.exchange(enable ? 1 : 0);
}

extern "C" void rvio_b200_profile(int enable)
{
    rvio::g_profile_on.store(enable ? 1 : 0);
}

// Writes "kernel count total_ms\n" lines (CUDA-event time of every launch since the last report) and clears the log.
extern "C" int rvio_b200_profile_report(char* buf, int cap)
{
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(rvio::g_prof_mu);
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& r : rvio::g_prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { auto& a = agg[r.name]; a.first++; a.second += ms; }
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    rvio::g_prof.clear();
    int off = 0;
    for (auto& kv : agg) {
        const int n = snprintf(buf + off, cap > off ? cap - off : 0, "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        if (n < 0 || off + n >= cap) break;
        off += n;
    }
    return off;
}
