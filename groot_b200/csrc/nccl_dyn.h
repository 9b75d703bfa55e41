// NCCL, bound at run time: libgrootgpu.so does not link libnccl. The single-GPU path needs no NCCL at all, and a host
// process may already carry its own copy (a Python host: torch's bundled libnccl.so.2) — the symbols are taken from that
// copy when it is loaded, otherwise from the system library. <nccl.h> is used for the types only.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>

namespace groot {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string source;
};

inline const NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    static std::string error;
    std::call_once(once, [] {
        void* h = nullptr;
        const char* env = getenv("GROOTGPU_NCCL_LIB");
        if (env && *env) { h = dlopen(env, RTLD_NOW | RTLD_GLOBAL); api.source = env; }
        if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); api.source = "libnccl.so.2 (already loaded by the host process)"; }
        if (!h) { h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL); api.source = "libnccl.so.2"; }
        if (!h) { h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL); api.source = "libnccl.so"; }
        if (!h) { error = "NCCL not found (libnccl.so.2; set GROOTGPU_NCCL_LIB): multi-GPU entry points need it"; return; }
        auto sym = [&](const char* name) {
            void* p = dlsym(h, name);
            if (!p && error.empty()) error = std::string("NCCL symbol missing: ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.CommAbort = reinterpret_cast<decltype(api.CommAbort)>(sym("ncclCommAbort"));
        api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
        api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
        api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    if (!error.empty()) throw std::runtime_error(error);
    return api;
}

}  // namespace groot
