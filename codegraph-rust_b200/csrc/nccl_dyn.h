// nccl_dyn.h — NCCL bound at run time with dlopen, so libcgvec_b200.so loads on machines (and in
// processes) without NCCL and, inside a PyTorch process, shares the libnccl.so.2 torch already loaded.
// Only the handful of entry points the top-k exchange needs.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdlib.h>

namespace cgv {

struct NcclComm;
typedef NcclComm* nccl_comm_t;
struct NcclUniqueId { char internal[128]; };
enum { kNcclSuccess = 0, kNcclUint64 = 5 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    const char* load_error = nullptr;

    bool load() {
        if (handle) return true;
        const char* env = getenv("CGVEC_NCCL_LIB");
        const char* names[] = {env, "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n) continue;
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { load_error = "libnccl.so.2 not found (set CGVEC_NCCL_LIB)"; return false; }
#define CGV_SYM(field, name)                                                     \
    field = reinterpret_cast<decltype(field)>(dlsym(handle, name));              \
    if (!field) { load_error = "missing NCCL symbol " name; handle = nullptr; return false; }
        CGV_SYM(GetUniqueId, "ncclGetUniqueId")
        CGV_SYM(CommInitRank, "ncclCommInitRank")
        CGV_SYM(CommInitAll, "ncclCommInitAll")
        CGV_SYM(CommDestroy, "ncclCommDestroy")
        CGV_SYM(AllGather, "ncclAllGather")
        CGV_SYM(GroupStart, "ncclGroupStart")
        CGV_SYM(GroupEnd, "ncclGroupEnd")
        CGV_SYM(GetErrorString, "ncclGetErrorString")
#undef CGV_SYM
        return true;
    }
};

inline NcclApi& nccl_api() {
    static NcclApi api;
    return api;
}

}  // namespace cgv
