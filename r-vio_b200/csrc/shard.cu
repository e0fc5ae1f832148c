// shard.cu -- NCCL plumbing of the feature-sharded single-stream form (BASELINE configs[4]).
//
// One process per GPU; every rank is fed the same frames and holds the same tracker / filter state.  Per frame the path
// has exactly two exchange steps, both enqueued ON THE PIPELINE'S OWN STREAM (so they are part of the frame's stream
// order and of the captured frame graph):
//   * ncclAllGather of the per-feature LK results (pixels, normalised points, status: <= 17 B / feature) after the
//     sharded k_lk -- RANSAC + bookkeeping then run replicated and bit-identical on every rank;
//   * ncclAllReduce(sum) of the reduce buffer [G | z | counters | per-class information] after the sharded
//     k_feature / k_gram -- the rank rule and the EKF solve then run replicated.
// libnccl is loaded at run time (dlopen, re-using a copy already in the process such as PyTorch's): a host that never
// shards has no NCCL dependency.
#include "common.cuh"
#include "shard.cuh"

#include <dlfcn.h>
#include <nccl.h>

namespace rvio {

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl()
{
    if (g_nccl.lib) return RVIO_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // a copy already mapped (e.g. PyTorch's bundled one)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("dlopen(libnccl.so.2)", dlerror()); return RVIO_ERR_STATE; }
#define SYM(field, name)                                                            \
    *(void**)(&g_nccl.field) = dlsym(h, name);                                      \
    if (!g_nccl.field) { set_error("dlsym", name); return RVIO_ERR_STATE; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.lib = h;
    return RVIO_OK;
}
}  // namespace

#define RVIO_NCCL_TRY(expr)                                                         \
    do {                                                                            \
        ncclResult_t _r = (expr);                                                   \
        if (_r != ncclSuccess) { set_error(#expr, g_nccl.GetErrorString(_r)); return RVIO_ERR_CUDA; } \
    } while (0)

int shard_unique_id(void* id128)
{
    int rc = load_nccl();
    if (rc != RVIO_OK) return rc;
    ncclUniqueId id;
    RVIO_NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return RVIO_OK;
}

int shard_comm_create(ShardComm* sc, int rank, int world, const void* id128, int device)
{
    int rc = load_nccl();
    if (rc != RVIO_OK) return rc;
    RVIO_CUDA_TRY(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t c = nullptr;
    RVIO_NCCL_TRY(g_nccl.CommInitRank(&c, world, id, rank));
    sc->comm = c; sc->rank = rank; sc->world = world;
    return RVIO_OK;
}

void shard_comm_destroy(ShardComm* sc)
{
    if (sc->comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)sc->comm);
    sc->comm = nullptr; sc->world = 1; sc->rank = 0;
}

// In-place all-gather of the three per-feature LK arrays (equal shards of S feature indices, this rank's shard at r * S).
int shard_allgather_lk(const ShardComm* sc, cudaStream_t s, float2* lk, float2* un, uint8_t* status, int S)
{
    if (t_replay) return RVIO_OK;                              // graph replay: the collectives are nodes of the captured graph
    ncclComm_t c = (ncclComm_t)sc->comm;
    const int r = sc->rank;
    RVIO_NCCL_TRY(g_nccl.GroupStart());
    RVIO_NCCL_TRY(g_nccl.AllGather(lk + (size_t)r * S, lk, (size_t)2 * S, ncclFloat, c, s));
    RVIO_NCCL_TRY(g_nccl.AllGather(un + (size_t)r * S, un, (size_t)2 * S, ncclFloat, c, s));
    RVIO_NCCL_TRY(g_nccl.AllGather(status + (size_t)r * S, status, (size_t)S, ncclUint8, c, s));
    RVIO_NCCL_TRY(g_nccl.GroupEnd());
    return RVIO_OK;
}

// The single reduce of the sharded update: sum of [G | z | counters | per-class information] over the ranks, in place.
int shard_allreduce_terms(const ShardComm* sc, cudaStream_t s, double* red, int count)
{
    if (t_replay) return RVIO_OK;
    RVIO_NCCL_TRY(g_nccl.AllReduce(red, red, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)sc->comm, s));
    return RVIO_OK;
}

}  // namespace rvio

extern "C" int rvio_b200_nccl_unique_id(void* id128)
{
    RVIO_ARG_CHECK(id128);
    return rvio::shard_unique_id(id128);
}
