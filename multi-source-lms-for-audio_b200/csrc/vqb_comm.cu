// The one exchange on the path: sum-all-reduce of the per-rank statistics buffer over NCCL (NVLink 5 / NVSwitch).
// libnccl.so.2 is resolved at run time with dlopen so that the library the host process already uses (torch's bundled
// NCCL) is shared instead of a second copy being linked in.
#include "vqb_internal.h"

#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

namespace vqb {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char* (*GetErrorString)(ncclResult_t);
    bool ok;
};

static NcclApi* nccl() {
    static NcclApi api{};
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void* h = nullptr;
    if (const char* p = getenv("VQB_NCCL_PATH")) h = dlopen(p, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot dlopen libnccl.so.2: %s (set VQB_NCCL_PATH)", dlerror()); return nullptr; }
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(dlsym(h, "ncclAllReduce"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllReduce && api.CommDestroy && api.GetErrorString;
    if (!api.ok) { set_error("libnccl.so.2 lacks a required symbol"); return nullptr; }
    return &api;
}

static int nccl_fail(NcclApi* a, ncclResult_t r, const char* what) {
    set_error("%s: %s", what, a->GetErrorString(r));
    return 1000 + (int)r;
}

}  // namespace vqb

using namespace vqb;

extern "C" {

int vqb_comm_unique_id(void* id_out_host) {
    static_assert(sizeof(ncclUniqueId) == VQB_UNIQUE_ID_BYTES, "ncclUniqueId size changed");
    if (!id_out_host) { set_error("vqb_comm_unique_id: NULL output"); return VQB_E_NULL; }
    NcclApi* a = nccl();
    if (!a) return VQB_E_NCCL;
    ncclUniqueId id;
    ncclResult_t r = a->GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(a, r, "ncclGetUniqueId");
    memcpy(id_out_host, &id, sizeof(id));
    return 0;
}

int vqb_comm_init(const void* id_host, int rank, int world, void** comm_out) {
    if (!id_host || !comm_out) { set_error("vqb_comm_init: NULL pointer argument"); return VQB_E_NULL; }
    if (world < 1 || rank < 0 || rank >= world) { set_error("vqb_comm_init: bad rank %d / world %d", rank, world); return VQB_E_SHAPE; }
    NcclApi* a = nccl();
    if (!a) return VQB_E_NCCL;
    ncclUniqueId id;
    memcpy(&id, id_host, sizeof(id));
    ncclComm_t comm = nullptr;
    ncclResult_t r = a->CommInitRank(&comm, world, id, rank);
    if (r != ncclSuccess) return nccl_fail(a, r, "ncclCommInitRank");
    *comm_out = comm;
    return 0;
}

int vqb_allreduce_stats(void* comm, float* stats, size_t n_floats, void* stream) {
    if (!comm || !stats) { set_error("vqb_allreduce_stats: NULL pointer argument"); return VQB_E_NULL; }
    NcclApi* a = nccl();
    if (!a) return VQB_E_NCCL;
    ncclResult_t r = a->AllReduce(stats, stats, n_floats, ncclFloat32, ncclSum, static_cast<ncclComm_t>(comm),
                                  static_cast<cudaStream_t>(stream));
    if (r != ncclSuccess) return nccl_fail(a, r, "ncclAllReduce");
    return 0;
}

int vqb_comm_destroy(void* comm) {
    if (!comm) return 0;
    NcclApi* a = nccl();
    if (!a) return VQB_E_NCCL;
    ncclResult_t r = a->CommDestroy(static_cast<ncclComm_t>(comm));
    if (r != ncclSuccess) return nccl_fail(a, r, "ncclCommDestroy");
    return 0;
}

}  // extern "C"
