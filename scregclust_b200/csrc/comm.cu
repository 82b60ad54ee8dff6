#include "comm.cuh"

#include <dlfcn.h>
#include <mutex>

#include "../../include/scregclust_b200.h"

namespace scrc {

namespace {
// The slice of nccl.h this library uses (NCCL 2.x ABI; ncclDataType_t / ncclRedOp_t values are stable).
struct UniqueId { char internal[128]; };
enum { kNcclInt8 = 0, kNcclUint8 = 1, kNcclInt32 = 2, kNcclFloat64 = 8 };
enum { kNcclSum = 0 };
struct Api {
    int (*GetUniqueId)(UniqueId*) = nullptr;
    int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};
Api g_api;
std::once_flag g_once;

void load_api() {
    std::call_once(g_once, [] {
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { g_api.why = std::string("libnccl.so.2 could not be loaded: ") + dlerror(); return; }
        auto sym = [&](const char* n) { return dlsym(h, n); };
        g_api.GetUniqueId = (decltype(g_api.GetUniqueId))sym("ncclGetUniqueId");
        g_api.CommInitRank = (decltype(g_api.CommInitRank))sym("ncclCommInitRank");
        g_api.CommDestroy = (decltype(g_api.CommDestroy))sym("ncclCommDestroy");
        g_api.AllReduce = (decltype(g_api.AllReduce))sym("ncclAllReduce");
        g_api.Broadcast = (decltype(g_api.Broadcast))sym("ncclBroadcast");
        g_api.GroupStart = (decltype(g_api.GroupStart))sym("ncclGroupStart");
        g_api.GroupEnd = (decltype(g_api.GroupEnd))sym("ncclGroupEnd");
        g_api.GetErrorString = (decltype(g_api.GetErrorString))sym("ncclGetErrorString");
        g_api.ok = g_api.GetUniqueId && g_api.CommInitRank && g_api.CommDestroy && g_api.AllReduce && g_api.Broadcast &&
                   g_api.GroupStart && g_api.GroupEnd && g_api.GetErrorString;
        if (!g_api.ok) g_api.why = "libnccl.so.2 lacks a required symbol";
    });
    if (!g_api.ok) throw Error(SCRC_ERR_COMM, "multi-GPU exchange unavailable: " + g_api.why);
}

void check(int rc, const char* what) {
    if (rc != 0) throw Error(SCRC_ERR_COMM, std::string("NCCL failure in ") + what + ": " + g_api.GetErrorString(rc));
}
}  // namespace

void comm_unique_id(char* out_128) {
    load_api();
    UniqueId id;
    check(g_api.GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out_128, id.internal, 128);
}

void Comm::init(const char* unique_id_128, int world_, int rank_, int device_) {
    if (world_ < 1 || rank_ < 0 || rank_ >= world_) throw Error(SCRC_ERR_ARG, "scrc_comm_create: rank / world out of range");
    load_api();
    world = world_; rank = rank_; device = device_;
    SCRC_CUDA(cudaSetDevice(device));
    UniqueId id;
    std::memcpy(id.internal, unique_id_128, 128);
    check(g_api.CommInitRank(&nccl, world, id, rank), "ncclCommInitRank");
}
Comm::~Comm() {
    if (nccl) { cudaSetDevice(device); g_api.CommDestroy(nccl); nccl = nullptr; }
}
void Comm::allreduce_sum_f64(double* p, size_t n, cudaStream_t st) {
    if (world > 1 && n) check(g_api.AllReduce(p, p, n, kNcclFloat64, kNcclSum, nccl, st), "ncclAllReduce(f64)");
}
void Comm::allreduce_sum_i32(int* p, size_t n, cudaStream_t st) {
    if (world > 1 && n) check(g_api.AllReduce(p, p, n, kNcclInt32, kNcclSum, nccl, st), "ncclAllReduce(i32)");
}
void Comm::allreduce_sum_u8(unsigned char* p, size_t n, cudaStream_t st) {
    if (world > 1 && n) check(g_api.AllReduce(p, p, n, kNcclUint8, kNcclSum, nccl, st), "ncclAllReduce(u8)");
}
void Comm::broadcast_bytes(void* p, size_t bytes, int root, cudaStream_t st) {
    if (world > 1 && bytes) check(g_api.Broadcast(p, p, bytes, kNcclUint8, root, nccl, st), "ncclBroadcast");
}
void Comm::group_begin() { if (world > 1) check(g_api.GroupStart(), "ncclGroupStart"); }
void Comm::group_end() { if (world > 1) check(g_api.GroupEnd(), "ncclGroupEnd"); }

}  // namespace scrc
