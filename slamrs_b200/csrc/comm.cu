#include "comm.h"

#include <dlfcn.h>
#include <stdlib.h>
#include <nccl.h>
#include <mutex>

namespace slamrs {

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string load_error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // SLAMRS_NCCL_LIB (tests): load this library instead, e.g. a path that does not exist
        const char* forced = getenv("SLAMRS_NCCL_LIB");
        const char* names[] = {forced ? forced : "libnccl.so.2", forced ? forced : "libnccl.so"};
        std::string first_error;
        for (const char* n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
            const char* e = dlerror();   // (a second dlerror() call returns NULL: read it once)
            if (first_error.empty()) first_error = e ? e : "?";
        }
        if (!api.lib) {
            api.load_error = std::string("dlopen(") + names[0] + ") failed: " + first_error;
            return;
        }
#define SLAMRS_SYM(field, name)                                                   \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name));      \
    if (!api.field) { api.load_error = std::string("missing NCCL symbol ") + name; return; }
        SLAMRS_SYM(GetUniqueId, "ncclGetUniqueId")
        SLAMRS_SYM(CommInitRank, "ncclCommInitRank")
        SLAMRS_SYM(CommDestroy, "ncclCommDestroy")
        SLAMRS_SYM(AllGather, "ncclAllGather")
        SLAMRS_SYM(Broadcast, "ncclBroadcast")
        SLAMRS_SYM(AllReduce, "ncclAllReduce")
        SLAMRS_SYM(GetErrorString, "ncclGetErrorString")
#undef SLAMRS_SYM
    });
    return &api;
}

bool nccl_ok(NcclApi* a, ncclResult_t r, const char* what, std::string* err) {
    if (r == ncclSuccess) return true;
    if (err) *err = std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "nccl error");
    return false;
}
}  // namespace

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
};

int comm_unique_id(uint8_t out[128], std::string* err) {
    NcclApi* a = nccl_api();
    if (!a->load_error.empty()) { if (err) *err = a->load_error; return -1; }
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (!nccl_ok(a, a->GetUniqueId(&id), "ncclGetUniqueId", err)) return -1;
    memcpy(out, &id, 128);
    return 0;
}

Comm* comm_create(const uint8_t id_bytes[128], int rank, int world, std::string* err) {
    NcclApi* a = nccl_api();
    if (!a->load_error.empty()) { if (err) *err = a->load_error; return nullptr; }
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    Comm* c = new Comm();
    c->rank = rank;
    c->world = world;
    if (!nccl_ok(a, a->CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank", err)) {
        delete c;
        return nullptr;
    }
    return c;
}

void comm_destroy(Comm* c) {
    if (!c) return;
    NcclApi* a = nccl_api();
    if (c->comm && a->CommDestroy) a->CommDestroy(c->comm);
    delete c;
}

int comm_all_gather(Comm* c, const void* send, void* recv, size_t bytes_per_rank, cudaStream_t s, std::string* err) {
    NcclApi* a = nccl_api();
    return nccl_ok(a, a->AllGather(send, recv, bytes_per_rank, ncclUint8, c->comm, s), "ncclAllGather", err) ? 0 : -1;
}

int comm_broadcast(Comm* c, void* buf, size_t bytes, int root, cudaStream_t s, std::string* err) {
    NcclApi* a = nccl_api();
    return nccl_ok(a, a->Broadcast(buf, buf, bytes, ncclUint8, root, c->comm, s), "ncclBroadcast", err) ? 0 : -1;
}

int comm_barrier(Comm* c, int* scratch, cudaStream_t s, std::string* err) {
    NcclApi* a = nccl_api();
    return nccl_ok(a, a->AllReduce(scratch, scratch, 1, ncclInt32, ncclSum, c->comm, s), "ncclAllReduce(barrier)", err)
               ? 0
               : -1;
}

}  // namespace slamrs
