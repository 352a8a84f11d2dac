// The ONE collective of the path: an all-gather of the per-view depth maps over NVLink / NVSwitch, exported
// through the C ABI so that a caller without torch.distributed can issue it (include/mvsb200.h, "multi-GPU").
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy the process already holds when PyTorch is loaded, otherwise
// the system's): libmvsb200.so itself has no link-time dependency on it and loads on a box without NCCL; only these entry
// points then fail, loudly.  The gather is IN PLACE: every rank owns the slice [rank * count, (rank + 1) * count) of the
// [world * count] buffer -- the slice its regression kernel (K3) wrote its depth maps into -- so there is no staging copy
// and nothing is allocated; the call is stream-ordered and legal inside a CUDA-graph capture.
#include <dlfcn.h>

#include <cstring>

#include <mutex>

#include "common.cuh"

namespace mvsb200 {

struct NcclUniqueId {
    char internal[128];
};
typedef void *NcclComm;
typedef int (*FnGetUniqueId)(NcclUniqueId *);
typedef int (*FnCommInitRank)(NcclComm *, int, NcclUniqueId, int);
typedef int (*FnAllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t);
typedef int (*FnCommDestroy)(NcclComm);
typedef const char *(*FnGetErrorString)(int);

static struct {
    void *handle;
    FnGetUniqueId get_unique_id;
    FnCommInitRank comm_init_rank;
    FnAllGather all_gather;
    FnCommDestroy comm_destroy;
    FnGetErrorString error_string;
} g_nccl;
static std::once_flag g_nccl_once;

static bool nccl_load()
{
    std::call_once(g_nccl_once, [] {
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy PyTorch (or the caller) already loaded
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        g_nccl.get_unique_id = reinterpret_cast<FnGetUniqueId>(dlsym(h, "ncclGetUniqueId"));
        g_nccl.comm_init_rank = reinterpret_cast<FnCommInitRank>(dlsym(h, "ncclCommInitRank"));
        g_nccl.all_gather = reinterpret_cast<FnAllGather>(dlsym(h, "ncclAllGather"));
        g_nccl.comm_destroy = reinterpret_cast<FnCommDestroy>(dlsym(h, "ncclCommDestroy"));
        g_nccl.error_string = reinterpret_cast<FnGetErrorString>(dlsym(h, "ncclGetErrorString"));
        if (g_nccl.get_unique_id && g_nccl.comm_init_rank && g_nccl.all_gather && g_nccl.comm_destroy) g_nccl.handle = h;
    });
    if (!g_nccl.handle) set_error("gather: libnccl.so.2 could not be loaded (no NCCL in this process or on this box)");
    return g_nccl.handle != nullptr;
}

static int nccl_check(int rc, const char *what)
{
    if (rc == 0) return MVSB200_OK;
    set_error("%s: NCCL error %d (%s)", what, rc, g_nccl.error_string ? g_nccl.error_string(rc) : "?");
    return MVSB200_E_CUDA;
}

}  // namespace mvsb200

using namespace mvsb200;

extern "C" int mvsb200_gather_unique_id(void *id128)
{
    MVSB200_REQUIRE(id128, "gather_unique_id: null pointer");
    if (!nccl_load()) return MVSB200_E_CUDA;
    NcclUniqueId id;
    if (int rc = nccl_check(g_nccl.get_unique_id(&id), "gather_unique_id")) return rc;
    memcpy(id128, &id, sizeof(id));
    return MVSB200_OK;
}

extern "C" int mvsb200_gather_init(const void *id128, int world, int rank, void **comm)
{
    MVSB200_REQUIRE(id128 && comm, "gather_init: null pointer");
    MVSB200_REQUIRE(world >= 1 && rank >= 0 && rank < world, "gather_init: rank %d of %d", rank, world);
    if (!nccl_load()) return MVSB200_E_CUDA;
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NcclComm c = nullptr;
    if (int rc = nccl_check(g_nccl.comm_init_rank(&c, world, id, rank), "gather_init")) return rc;
    *comm = c;
    return MVSB200_OK;
}

extern "C" int mvsb200_allgather_depth(void *comm, float *all_maps, long long count_per_rank, int rank, mvsb200_stream_t stream)
{
    MVSB200_REQUIRE(comm && all_maps && count_per_rank > 0 && rank >= 0, "allgather_depth: bad argument");
    if (!nccl_load()) return MVSB200_E_CUDA;
    // in place: the send buffer is this rank's slice of the receive buffer (ncclAllGather's in-place convention)
    return nccl_check(g_nccl.all_gather(all_maps + (size_t)rank * count_per_rank, all_maps, (size_t)count_per_rank, /* ncclFloat32 */ 7, comm,
                                        (cudaStream_t)stream),
                      "allgather_depth");
}

extern "C" int mvsb200_gather_destroy(void *comm)
{
    if (!comm) return MVSB200_OK;
    if (!nccl_load()) return MVSB200_E_CUDA;
    return nccl_check(g_nccl.comm_destroy(comm), "gather_destroy");
}
