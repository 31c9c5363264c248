// Multi-GPU exchange of the path: ONE collective, an all-reduce-min of the best-known objective, issued
// in-stream on the device-resident bound so a portfolio epoch needs no host round trip:
//     sls kernel -> best_reduce (bounds[0] = min(bounds[0], local best)) -> ncclAllReduce(bounds, min) -> next epoch
// One process per GPU; the host (Rust in the reference's world, torch.distributed / a file in the tests) only carries
// the 128-byte ncclUniqueId from rank 0 to the others.  NCCL is bound at run time with dlopen("libnccl.so.2"): libtss has
// no link-time NCCL dependency and, inside a process that already loaded NCCL (PyTorch), shares that copy.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "engine.hpp"

namespace tss {

struct NcclUniqueId { char internal[128]; };  // NCCL_UNIQUE_ID_BYTES
typedef void* NcclComm;
constexpr int kNcclInt32 = 2, kNcclMin = 3;   // ncclDataType_t / ncclRedOp_t values (nccl.h)

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
        api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
        api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
        api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
        if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) api.error = "libnccl lacks a required symbol";
    });
    return &api;
}

struct Comm {
    NcclComm comm = nullptr;
    int rank = 0, world = 1;
};

void comm_destroy(Comm* c) {
    if (!c) return;
    if (c->comm && nccl_api()->CommDestroy) nccl_api()->CommDestroy(c->comm);
    delete c;
}

// in-stream all-reduce-min of n int32 values in device memory
int comm_allreduce_min(tss_engine* e, Comm* c, int* dev, int n) {
    if (!c || c->world <= 1) return TSS_OK;
    NcclApi* api = nccl_api();
    int r = api->AllReduce(dev, dev, (size_t)n, kNcclInt32, kNcclMin, c->comm, e->stream);
    if (r != 0) return e->fail(TSS_E_CUDA, "ncclAllReduce failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    e->stats.kernel_launches++;
    return TSS_OK;
}

// in-stream all-reduce-min of n uint32 words: src -> dst (the window-decomposed portfolio ships the winner's layout this
// way: losers contribute all-ones, so the min is the winner's bitboard — no broadcast root has to be known on the host)
int comm_allreduce_min_u32(tss_engine* e, Comm* c, const uint32_t* src, uint32_t* dst, int n) {
    if (!c || c->world <= 1) return TSS_OK;
    NcclApi* api = nccl_api();
    int r = api->AllReduce(src, dst, (size_t)n, 3 /* ncclUint32 */, kNcclMin, c->comm, e->stream);
    if (r != 0) return e->fail(TSS_E_CUDA, "ncclAllReduce failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    e->stats.kernel_launches++;
    return TSS_OK;
}
// in-stream all-reduce-SUM of n uint32 words: src -> dst.  The window-decomposed portfolio assembles its next layout this way:
// every bit of the result is contributed by exactly one rank (the winner of that window; rank 0 for the frozen supports), so
// the sum is a bitwise OR without carries.
int comm_allreduce_sum_u32(tss_engine* e, Comm* c, const uint32_t* src, uint32_t* dst, int n) {
    if (!c || c->world <= 1) return TSS_OK;
    NcclApi* api = nccl_api();
    int r = api->AllReduce(src, dst, (size_t)n, 3 /* ncclUint32 */, 0 /* ncclSum */, c->comm, e->stream);
    if (r != 0) return e->fail(TSS_E_CUDA, "ncclAllReduce failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    e->stats.kernel_launches++;
    return TSS_OK;
}
int comm_rank(const Comm* c) { return c ? c->rank : 0; }
int comm_world(const Comm* c) { return c ? c->world : 1; }

}  // namespace tss

using namespace tss;

extern "C" {

int tss_comm_unique_id(tss_engine* e, uint8_t* out_id128) {
    if (!e) return TSS_E_INVALID;
    if (!out_id128) return e->fail(TSS_E_INVALID, "tss_comm_unique_id: null output");
    NcclApi* api = nccl_api();
    if (!api->error.empty()) return e->fail(TSS_E_UNSUPPORTED, "%s", api->error.c_str());
    NcclUniqueId id;
    int r = api->GetUniqueId(&id);
    if (r != 0) return e->fail(TSS_E_CUDA, "ncclGetUniqueId failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?");
    std::memcpy(out_id128, id.internal, sizeof id.internal);
    return TSS_OK;
}

int tss_comm_init(tss_engine* e, const uint8_t* id128, int32_t rank, int32_t world) {
    if (!e) return TSS_E_INVALID;
    if (!id128 || world < 1 || rank < 0 || rank >= world) return e->fail(TSS_E_INVALID, "tss_comm_init: bad arguments");
    NcclApi* api = nccl_api();
    if (!api->error.empty()) return e->fail(TSS_E_UNSUPPORTED, "%s", api->error.c_str());
    TSS_CUDA(e, cudaSetDevice(e->device));
    if (e->comm) { comm_destroy(e->comm); e->comm = nullptr; }
    Comm* c = new Comm();
    c->rank = rank;
    c->world = world;
    NcclUniqueId id;
    std::memcpy(id.internal, id128, sizeof id.internal);
    int r = api->CommInitRank(&c->comm, world, id, rank);
    if (r != 0) { delete c; return e->fail(TSS_E_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString ? api->GetErrorString(r) : "?"); }
    e->comm = c;
    return TSS_OK;
}

int tss_comm_world(const tss_engine* e) { return (e && e->comm) ? e->comm->world : 1; }

}  // extern "C"
