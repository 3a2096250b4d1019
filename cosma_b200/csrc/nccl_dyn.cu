#include "nccl_dyn.h"

#include <cstdlib>
#include <dlfcn.h>
#include <mutex>
#include <string>

namespace cosma_b200 {
void set_last_error(const std::string& msg);

const NcclApi* nccl() {
    static NcclApi api;
    static bool ok = false;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        // COSMA_B200_NCCL_LIB: explicit path; otherwise whatever libnccl the process already loaded (torch's bundled one
        // under Python) or the dynamic loader finds (the system NCCL for plain C++ / Fortran hosts)
        if (const char* path = std::getenv("COSMA_B200_NCCL_LIB"))
            if (*path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            if (h) break;
            h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (h) break;
        }
        if (!h) {
            set_last_error(std::string("cannot load NCCL: ") + dlerror());
            return;
        }
        bool all = true;
#define COSMA_B200_SYM(field, sym)                                              \
    api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, sym));           \
    if (!api.field) { all = false; set_last_error(std::string("NCCL symbol missing: ") + sym); }
        COSMA_B200_SYM(GetUniqueId, "ncclGetUniqueId")
        COSMA_B200_SYM(CommInitRank, "ncclCommInitRank")
        COSMA_B200_SYM(CommSplit, "ncclCommSplit")
        COSMA_B200_SYM(CommDestroy, "ncclCommDestroy")
        COSMA_B200_SYM(CommCount, "ncclCommCount")
        COSMA_B200_SYM(CommUserRank, "ncclCommUserRank")
        COSMA_B200_SYM(AllGather, "ncclAllGather")
        COSMA_B200_SYM(ReduceScatter, "ncclReduceScatter")
        COSMA_B200_SYM(Reduce, "ncclReduce")
        COSMA_B200_SYM(AllReduce, "ncclAllReduce")
        COSMA_B200_SYM(Send, "ncclSend")
        COSMA_B200_SYM(Recv, "ncclRecv")
        COSMA_B200_SYM(GroupStart, "ncclGroupStart")
        COSMA_B200_SYM(GroupEnd, "ncclGroupEnd")
        COSMA_B200_SYM(GetErrorString, "ncclGetErrorString")
        COSMA_B200_SYM(GetVersion, "ncclGetVersion")
#undef COSMA_B200_SYM
        ok = all;
    });
    return ok ? &api : nullptr;
}

}  // namespace cosma_b200
