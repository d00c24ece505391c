// nccl_dyn.h -- NCCL is bound at run time (dlopen "libnccl.so.2"), not at link time.
//
// Why: inside a Python process torch has usually loaded ITS bundled NCCL (2.28.x) already; a DT_NEEDED on the system NCCL
// (2.27.x) would either load a second copy or -- if this library is loaded first -- shadow torch's copy by SONAME and break
// `import torch`.  dlopen by SONAME returns whatever copy the process already has, and single-GPU users never touch NCCL.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;   // optional
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    const char *error = nullptr;

    bool load() {
        if (handle) return true;
#ifdef CDBG_EMU_ACTIVE   // CPU emulation build of the tests: ranks are threads of one process, see tests/emu/nccl.h
        handle = this;
        GetUniqueId = emu_ncclGetUniqueId; CommInitRank = emu_ncclCommInitRank; CommDestroy = emu_ncclCommDestroy; CommAbort = emu_ncclCommDestroy;
        AllGather = emu_ncclAllGather; Send = emu_ncclSend; Recv = emu_ncclRecv; GroupStart = emu_ncclGroupStart; GroupEnd = emu_ncclGroupEnd;
        GetErrorString = emu_ncclGetErrorString;
        return true;
#endif
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (handle) break; }
        if (!handle) { error = "cannot dlopen libnccl.so.2"; return false; }
#define NCCL_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(handle, sym)); if (!field) { error = "missing NCCL symbol " sym; handle = nullptr; return false; }
        NCCL_SYM(GetUniqueId, "ncclGetUniqueId") NCCL_SYM(CommInitRank, "ncclCommInitRank") NCCL_SYM(CommDestroy, "ncclCommDestroy")
        NCCL_SYM(AllGather, "ncclAllGather") NCCL_SYM(Send, "ncclSend") NCCL_SYM(Recv, "ncclRecv") NCCL_SYM(GroupStart, "ncclGroupStart")
        NCCL_SYM(GroupEnd, "ncclGroupEnd") NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
        CommAbort = reinterpret_cast<decltype(CommAbort)>(dlsym(handle, "ncclCommAbort"));
        return true;
    }
};

static NcclApi g_nccl;
// route the plain NCCL names used in the host code through the table
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclAllGather g_nccl.AllGather
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclGetErrorString g_nccl.GetErrorString
