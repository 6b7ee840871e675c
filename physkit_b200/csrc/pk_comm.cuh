// pk_comm.cuh — the one exchange step of a world that spans GPUs: NCCL all-gathers over NVLink on the
// context's stream.  NCCL is resolved at run time (dlopen of libnccl.so.2: the copy already in the process when the
// host has loaded one — e.g. the one bundled with torch — else the system library), so the collision library itself
// has no link-time dependency on it and loads on a box without NCCL; pk_comm_init then reports why it cannot work.
#pragma once

#include <dlfcn.h>
#include <nccl.h>

#include <mutex>
#include <string>

namespace pk
{

struct NcclApi
{
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

inline NcclApi &nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once,
                   []()
                   {
                       void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD); // the host's copy, if it has one
                       if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
                       if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
                       if (!h)
                       {
                           api.error = std::string("NCCL not found: ") + dlerror();
                           return;
                       }
                       api.handle = h;
#define PK_NCCL_SYM(name)                                                          \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(h, "nccl" #name));       \
    if (!api.name) api.error = "NCCL symbol nccl" #name " not found"
                       PK_NCCL_SYM(GetUniqueId);
                       PK_NCCL_SYM(CommInitRank);
                       PK_NCCL_SYM(CommDestroy);
                       PK_NCCL_SYM(AllGather);
                       PK_NCCL_SYM(Broadcast);
                       PK_NCCL_SYM(GroupStart);
                       PK_NCCL_SYM(GroupEnd);
                       PK_NCCL_SYM(GetErrorString);
#undef PK_NCCL_SYM
                   });
    return api;
}

} // namespace pk
