// tests/emu/nccl.h -- TEST INFRASTRUCTURE ONLY: an in-process stand-in for NCCL used by the CPU emulation build
// (tests/emu/cuda_runtime.h).  Every "rank" is a host thread of ONE process driving its own cdbg context; the few collectives the
// sharded path uses (all-gather, grouped send/recv, communicator setup) are memcpy between the ranks' buffers behind a barrier.
// This lets the no-GPU suite run the real sharded orchestration (shard_host.inl) with its real exchange layouts.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <map>
#include <mutex>
#include <random>
#include <string>
#include <vector>

struct ncclComm;
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0, ncclUnhandledCudaError = 1, ncclSystemError = 2, ncclInternalError = 3, ncclInvalidArgument = 4 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclUint8 = 1, ncclInt32 = 2, ncclInt = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5 } ncclDataType_t;

namespace emu_nccl {

inline size_t dtype_size(ncclDataType_t t) { return t == ncclInt8 || t == ncclUint8 ? 1 : (t == ncclInt32 || t == ncclUint32 ? 4 : 8); }

struct Op { int peer; const void *src; void *dst; size_t bytes; };

struct World {
    int size = 0, joined = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0; uint64_t gen = 0;
    std::vector<const void *> posted;                 // all-gather: every rank's send buffer
    std::vector<std::vector<Op>> sends;               // grouped send/recv: sends[rank] in program order
    void barrier() {
        std::unique_lock<std::mutex> l(m);
        uint64_t g = gen;
        if (++arrived == size) { arrived = 0; gen++; cv.notify_all(); return; }
        cv.wait(l, [&] { return gen != g; });
    }
};

inline std::mutex &registry_mutex() { static std::mutex m; return m; }
inline std::map<std::string, World *> &registry() { static std::map<std::string, World *> r; return r; }

}  // namespace emu_nccl

struct ncclComm {
    emu_nccl::World *w; int rank;
    bool grouping = false;
    std::vector<emu_nccl::Op> sends, recvs;
};

// the communicator of the calling thread (one rank = one host thread): lets a group without any operation still rendezvous
inline ncclComm_t &emu_nccl_thread_comm() { static thread_local ncclComm_t c = nullptr; return c; }

inline ncclResult_t emu_ncclGetUniqueId(ncclUniqueId *id) {
    static std::mt19937_64 rng(std::random_device{}());
    std::lock_guard<std::mutex> l(emu_nccl::registry_mutex());
    for (int i = 0; i < 16; i++) { uint64_t v = rng(); memcpy(id->internal + 8 * i, &v, 8); }
    return ncclSuccess;
}
inline ncclResult_t emu_ncclCommInitRank(ncclComm_t *comm, int nranks, ncclUniqueId id, int rank) {
    emu_nccl::World *w;
    {
        std::lock_guard<std::mutex> l(emu_nccl::registry_mutex());
        auto &r = emu_nccl::registry();
        std::string key(id.internal, 128);
        auto it = r.find(key);
        if (it == r.end()) { w = new emu_nccl::World(); w->size = nranks; w->posted.resize(nranks); w->sends.resize(nranks); r[key] = w; }
        else w = it->second;
    }
    if (w->size != nranks || rank < 0 || rank >= nranks) return ncclInvalidArgument;
    *comm = new ncclComm{w, rank};
    emu_nccl_thread_comm() = *comm;
    w->barrier();                                     // every rank has joined (like the real call, this is a rendezvous)
    return ncclSuccess;
}
inline ncclResult_t emu_ncclCommDestroy(ncclComm_t c) { delete c; return ncclSuccess; }
inline const char *emu_ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "emulated NCCL error"; }

inline ncclResult_t emu_ncclAllGather(const void *send, void *recv, size_t count, ncclDataType_t t, ncclComm_t c, void *) {
    emu_nccl::World *w = c->w;
    emu_nccl_thread_comm() = c;
    const size_t bytes = count * emu_nccl::dtype_size(t);
    w->posted[c->rank] = send;
    w->barrier();
    for (int r = 0; r < w->size; r++) memmove((uint8_t *)recv + (size_t)r * bytes, w->posted[r], bytes);
    w->barrier();                                     // nobody reuses its send buffer before everyone has copied
    return ncclSuccess;
}

inline ncclComm_t &emu_nccl_group_comm() { static thread_local ncclComm_t c = nullptr; return c; }
inline bool &emu_nccl_in_group() { static thread_local bool g = false; return g; }

inline ncclResult_t emu_ncclGroupStart() { emu_nccl_in_group() = true; emu_nccl_group_comm() = nullptr; return ncclSuccess; }
inline ncclResult_t emu_ncclSend(const void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, void *) {
    if (!emu_nccl_in_group()) return ncclInvalidArgument;   // (the library only uses grouped point-to-point)
    emu_nccl_group_comm() = c;
    c->sends.push_back({peer, buf, nullptr, count * emu_nccl::dtype_size(t)});
    return ncclSuccess;
}
inline ncclResult_t emu_ncclRecv(void *buf, size_t count, ncclDataType_t t, int peer, ncclComm_t c, void *) {
    if (!emu_nccl_in_group()) return ncclInvalidArgument;
    emu_nccl_group_comm() = c;
    c->recvs.push_back({peer, nullptr, buf, count * emu_nccl::dtype_size(t)});
    return ncclSuccess;
}
// Every rank closes every group (the library's exchanges are collective), also a rank that had nothing to send or receive: it
// still has to rendezvous, with the communicator its thread initialised.
inline ncclResult_t emu_ncclGroupEnd() {
    ncclComm_t c = emu_nccl_group_comm() ? emu_nccl_group_comm() : emu_nccl_thread_comm();
    emu_nccl_in_group() = false;
    if (!c) return ncclInvalidArgument;
    emu_nccl_thread_comm() = c;
    emu_nccl::World *w = c->w;
    w->sends[c->rank] = c->sends;
    w->barrier();
    ncclResult_t res = ncclSuccess;
    std::vector<size_t> next(w->size, 0);             // k-th receive from peer p pairs with p's k-th send to this rank
    for (auto &rv : c->recvs) {
        auto &ps = w->sends[rv.peer];
        size_t &i = next[rv.peer];
        while (i < ps.size() && ps[i].peer != c->rank) i++;
        if (i == ps.size() || ps[i].bytes != rv.bytes) { res = ncclInternalError; break; }
        memmove(rv.dst, ps[i].src, rv.bytes);
        i++;
    }
    w->barrier();
    c->sends.clear(); c->recvs.clear();
    return res;
}
