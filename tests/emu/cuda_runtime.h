// tests/emu/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY: a CPU stand-in for <cuda_runtime.h>.
//
// tests/emu/build_emu.py compiles the UNCHANGED library sources (bcalm_b200/csrc/*.cu/.cuh/.inl) with g++ and this header
// first on the include path, giving tests/_emu_build/libcdbg_emu.so: the same host orchestration and the same kernel
// bodies, with every CUDA thread run by an OS thread.  It exists so that kernel and host LOGIC (indexing, reservations,
// scans, overflow paths, stage hand-offs) can be checked against the oracle in a container without a GPU.  It is not a
// product path: nothing under bcalm_b200/ loads it, it says nothing about performance, and it cannot show memory-model or
// occupancy problems.  The GPU parity tests (-m gpu) remain the gate.
//
// Model: the blocks of a launch run one after another (so `__shared__` variables can be plain statics and dynamic shared memory
// one global buffer); inside a block every CUDA thread is a fiber (default) or an OS thread (CDBG_EMU_ENGINE=threads), see
// "execution engine" below.  The "device" is host memory and every runtime call is synchronous.
#pragma once
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <condition_variable>
#include <functional>
#include <algorithm>
#include <map>
#include <mutex>
#include <thread>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#define CDBG_EMU_ACTIVE 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ------------------------------------------------------------------------------------------------ vector types
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }

// ------------------------------------------------------------------------------------------------ execution engine
//
// Two engines behind one interface.  "fibers" (default): the CUDA threads of a block are user-level contexts that one OS thread
// switches between (a 20-instruction x86-64 context switch, ucontext elsewhere); a thread runs until it blocks in __syncthreads
// or a warp collective.  Fast -- a barrier is a handful of switches instead of hundreds of futex wake-ups -- and deterministic,
// and a kernel that can never make progress is reported as a deadlock instead of hanging.  "threads" (CDBG_EMU_ENGINE=threads):
// one OS thread per CUDA thread, truly concurrent, so shared-memory and global atomics really race; slow.
#if !defined(__x86_64__)
#include <ucontext.h>
#endif
namespace emu {

constexpr unsigned MAX_THREADS = 1024, MAX_WARPS = MAX_THREADS / 32;
constexpr size_t DYN_SMEM_BYTES = 256 * 1024, FIBER_STACK = 96 * 1024;

struct Tls { unsigned lane = 0, warp = 0; };
inline thread_local Tls tls;
inline void set_builtin(unsigned tid, unsigned bid);
inline void set_dims(unsigned block, unsigned grid);

struct Barrier {   // OS-thread generation barrier whose party count can shrink (threads that returned from the kernel)
    std::mutex m;
    std::condition_variable cv;
    unsigned expected = 0, arrived = 0;
    uint64_t gen = 0;
    void reset(unsigned n) { expected = n; arrived = 0; }
    void wait() {
        std::unique_lock<std::mutex> l(m);
        uint64_t g = gen;
        if (++arrived >= expected) { arrived = 0; gen++; cv.notify_all(); return; }
        cv.wait(l, [&] { return gen != g; });
    }
    void leave() {
        std::unique_lock<std::mutex> l(m);
        if (expected) expected--;
        if (expected && arrived >= expected) { arrived = 0; gen++; cv.notify_all(); }
    }
};

struct FBarrier { unsigned expected = 0, arrived = 0; uint64_t gen = 0; };   // the same for fibers (single OS thread: no locking)

#if defined(__x86_64__)
extern "C" void emu_ctx_switch(void **save_sp, void *load_sp);
__asm__(".text\n.p2align 4\n.globl emu_ctx_switch\n.type emu_ctx_switch,@function\nemu_ctx_switch:\n"
        "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
        "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
        "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
        ".size emu_ctx_switch,.-emu_ctx_switch\n");
#endif

struct Fiber {
#if defined(__x86_64__)
    void *sp = nullptr;
#else
    ucontext_t uc;
#endif
    bool done = false;
    const uint64_t *wait_gen = nullptr;   // blocked until *wait_gen != wait_val
    uint64_t wait_val = 0;
};

struct Engine {
    // ---- common
    unsigned grid = 0, block = 0;
    const std::function<void()> *body = nullptr;
    uint64_t warp_slot[MAX_WARPS][32];
    unsigned char *dyn = nullptr;
    std::mutex launch_mutex;   // one launch at a time (contexts on different host threads serialise here)
    bool use_threads = false;

    // ---- fibers
    std::vector<Fiber> fib;
    unsigned char *stacks = nullptr;
    unsigned cur = 0, alive = 0;
    FBarrier fblock, fwarp[MAX_WARPS];
#if defined(__x86_64__)
    void *sched_sp = nullptr;
#else
    ucontext_t sched_uc;
#endif
    static void fiber_main();
    void to_fiber(unsigned i) {
        cur = i;
        tls.lane = i & 31; tls.warp = i >> 5;
#if defined(__x86_64__)
        emu_ctx_switch(&sched_sp, fib[i].sp);
#else
        swapcontext(&sched_uc, &fib[i].uc);
#endif
    }
    void to_scheduler() {
#if defined(__x86_64__)
        emu_ctx_switch(&fib[cur].sp, sched_sp);
#else
        swapcontext(&fib[cur].uc, &sched_uc);
#endif
    }
    void fwait(FBarrier &b) {
        if (++b.arrived >= b.expected) { b.arrived = 0; b.gen++; return; }
        Fiber &f = fib[cur];
        f.wait_gen = &b.gen; f.wait_val = b.gen;
        to_scheduler();
    }
    static void fleave(FBarrier &b) {
        if (b.expected) b.expected--;
        if (b.expected && b.arrived >= b.expected) { b.arrived = 0; b.gen++; }
    }
    void run_block_fibers(unsigned bidx) {
        if (!stacks) { if (posix_memalign((void **)&stacks, 4096, (size_t)MAX_THREADS * FIBER_STACK)) abort(); }
        fib.assign(block, Fiber());
        fblock = FBarrier(); fblock.expected = block;
        for (unsigned w = 0; w * 32 < block; w++) { fwarp[w] = FBarrier(); fwarp[w].expected = block - w * 32 < 32 ? block - w * 32 : 32; }
        for (unsigned i = 0; i < block; i++) {
            unsigned char *top = stacks + (size_t)(i + 1) * FIBER_STACK;
#if defined(__x86_64__)
            void **sp = (void **)top;
            *(--sp) = nullptr;                       // keeps the entry aligned like after a call
            *(--sp) = (void *)&Engine::fiber_main;   // `ret` of the first switch lands here
            for (int r = 0; r < 6; r++) *(--sp) = nullptr;
            fib[i].sp = sp;
#else
            getcontext(&fib[i].uc);
            fib[i].uc.uc_stack.ss_sp = top - FIBER_STACK; fib[i].uc.uc_stack.ss_size = FIBER_STACK; fib[i].uc.uc_link = nullptr;
            makecontext(&fib[i].uc, (void (*)()) & Engine::fiber_main, 0);
#endif
        }
        alive = block;
        while (alive) {
            bool progressed = false;
            for (unsigned i = 0; i < block; i++) {
                Fiber &f = fib[i];
                if (f.done || (f.wait_gen && *f.wait_gen == f.wait_val)) continue;
                f.wait_gen = nullptr;
                set_builtin(i, bidx);
                to_fiber(i);
                progressed = true;
            }
            if (!progressed) {
                fprintf(stderr, "[emu] deadlock in block %u: %u of %u threads are blocked in a barrier or warp collective that the others never reach\n", bidx, alive, block);
                abort();
            }
        }
    }

    // ---- OS threads
    std::mutex m;
    std::condition_variable cv_start, cv_done;
    std::vector<std::thread> workers;
    uint64_t gen = 0;
    unsigned done = 0;
    Barrier pre, block_bar, warp_bar[MAX_WARPS];
    void run(unsigned id) {
        for (unsigned b = 0; b < grid; b++) {
            pre.wait();                      // every thread has left the previous block
            if (id == 0) {
                block_bar.reset(block);
                for (unsigned w = 0; w * 32 < block; w++) warp_bar[w].reset(block - w * 32 < 32 ? block - w * 32 : 32);
            }
            pre.wait();
            tls.lane = id & 31; tls.warp = id >> 5;
            set_builtin(id, b);
            (*body)();
            block_bar.leave();
            warp_bar[id >> 5].leave();
        }
    }
    void worker(unsigned id, uint64_t seen) {
        for (;;) {
            {
                std::unique_lock<std::mutex> l(m);
                cv_start.wait(l, [&] { return gen != seen; });
                seen = gen;
                if (id >= block) continue;
            }
            run(id);
            {
                std::unique_lock<std::mutex> l(m);
                if (++done == block) cv_done.notify_one();
            }
        }
    }

    void launch(unsigned g, unsigned b, size_t smem, const std::function<void()> &fn) {
        std::lock_guard<std::mutex> one(launch_mutex);
        if (b == 0 || b > MAX_THREADS || smem > DYN_SMEM_BYTES) { fprintf(stderr, "[emu] bad launch geometry %u x %u, %zu B smem\n", g, b, smem); abort(); }
        if (g == 0) { fprintf(stderr, "[emu] launch with an empty grid (block %u): cudaErrorInvalidConfiguration on a GPU\n", b); abort(); }
        if (!dyn) {
            if (posix_memalign((void **)&dyn, 1024, DYN_SMEM_BYTES)) abort();
            const char *e = getenv("CDBG_EMU_ENGINE");
            use_threads = e && strcmp(e, "threads") == 0;
        }
        grid = g; block = b; body = &fn;
        set_dims(b, g);
        if (!use_threads) {
            for (unsigned bidx = 0; bidx < g; bidx++) run_block_fibers(bidx);
            return;
        }
        std::unique_lock<std::mutex> l(m);
        while (workers.size() < b) {   // new workers start from the current generation: they must not replay an old launch
            unsigned id = (unsigned)workers.size(); uint64_t g0 = gen;
            workers.emplace_back([this, id, g0] { worker(id, g0); });
            workers.back().detach();
        }
        done = 0;
        pre.reset(b);
        gen++;
        cv_start.notify_all();
        cv_done.wait(l, [&] { return done == block; });
    }

    // ---- what the intrinsics call
    void sync_block() { if (use_threads) block_bar.wait(); else fwait(fblock); }
    void sync_warp() { if (use_threads) warp_bar[tls.warp].wait(); else fwait(fwarp[tls.warp]); }
    void yield_now() {   // a spin-wait on memory another CUDA thread writes must let that thread run
        if (use_threads) sched_yield(); else { fib[cur].wait_gen = nullptr; to_scheduler(); }
    }
};

inline Engine &engine() { static Engine *e = new Engine(); return *e; }   // leaked on purpose: worker threads are detached
inline void Engine::fiber_main() {
    Engine &e = engine();
    (*e.body)();
    Fiber &f = e.fib[e.cur];
    f.done = true;
    e.alive--;
    fleave(e.fblock);
    fleave(e.fwarp[e.cur >> 5]);
    e.to_scheduler();
    abort();   // a finished fiber is never resumed
}
inline unsigned char *dyn_smem() { return engine().dyn; }
inline void launch(unsigned g, unsigned b, size_t smem, const std::function<void()> &fn) { engine().launch(g, b, smem, fn); }
// dynamic shared memory above 48 KB needs cudaFuncAttributeMaxDynamicSharedMemorySize on THAT kernel (at most 227 KB on sm_100)
inline std::map<const void *, size_t> &smem_optin() { static std::map<const void *, size_t> m; return m; }
inline std::mutex &smem_optin_mutex() { static std::mutex m; return m; }
inline void launch_kernel(const void *kern, const char *name, unsigned g, unsigned b, size_t smem, const std::function<void()> &fn) {
    if (smem > 48 * 1024) {
        std::lock_guard<std::mutex> l(smem_optin_mutex());
        auto it = smem_optin().find(kern);
        if (it == smem_optin().end() || it->second < smem) {
            fprintf(stderr, "[emu] %s launched with %zu B of dynamic shared memory without the opt-in attribute: cudaErrorInvalidValue on a GPU\n", name, smem);
            abort();
        }
    }
    engine().launch(g, b, smem, fn);
}
inline void yield_now() { engine().yield_now(); }

template <class T>
inline T warp_exchange(T v, unsigned src_lane) {
    static_assert(sizeof(T) <= 8 && std::is_integral<T>::value, "emu shuffles carry integers up to 64 bits");
    Engine &e = engine();
    const unsigned w = tls.warp, lane = tls.lane;
    e.warp_slot[w][lane] = (uint64_t)v;
    e.sync_warp();
    T r = (T)e.warp_slot[w][src_lane & 31];
    e.sync_warp();
    return r;
}
inline void need_full_mask(unsigned mask) {
    if (mask != 0xFFFFFFFFu) { fprintf(stderr, "[emu] warp collective with a partial mask %08x is not modelled\n", mask); abort(); }
}

}  // namespace emu

// the built-in variables are real objects (not macros: cudaLaunchConfig_t has members called gridDim / blockDim)
inline thread_local dim3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

inline void emu::set_builtin(unsigned tid, unsigned bid) { threadIdx = dim3(tid, 0, 0); blockIdx = dim3(bid, 0, 0); }
inline void emu::set_dims(unsigned block, unsigned grid) { blockDim = dim3(block, 1, 1); gridDim = dim3(grid, 1, 1); }

static inline void __syncthreads() { emu::engine().sync_block(); }
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { emu::need_full_mask(mask); emu::engine().sync_warp(); }
static inline void __threadfence() { __sync_synchronize(); }
static inline unsigned __activemask() { return 0xFFFFFFFFu; }

template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) { emu::need_full_mask(mask); return emu::warp_exchange(v, (unsigned)src); }
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int d) { emu::need_full_mask(mask); return emu::warp_exchange(v, emu::tls.lane ^ (unsigned)d); }
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d) {
    emu::need_full_mask(mask);
    return emu::warp_exchange(v, emu::tls.lane >= d ? emu::tls.lane - d : emu::tls.lane);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d) {
    emu::need_full_mask(mask);
    return emu::warp_exchange(v, emu::tls.lane + d < 32 ? emu::tls.lane + d : emu::tls.lane);
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    emu::need_full_mask(mask);
    emu::Engine &e = emu::engine();
    const unsigned w = emu::tls.warp;
    e.warp_slot[w][emu::tls.lane] = pred ? 1 : 0;
    e.sync_warp();
    unsigned r = 0;
    const unsigned lanes = e.block - w * 32 < 32 ? e.block - w * 32 : 32;
    for (unsigned i = 0; i < lanes; i++) r |= (unsigned)e.warp_slot[w][i] << i;
    e.sync_warp();
    return r;
}

static inline unsigned __match_any_sync(unsigned mask, unsigned v) {
    emu::need_full_mask(mask);
    emu::Engine &e = emu::engine();
    const unsigned w = emu::tls.warp;
    e.warp_slot[w][emu::tls.lane] = v;
    e.sync_warp();
    unsigned r = 0;
    const unsigned lanes = e.block - w * 32 < 32 ? e.block - w * 32 : 32;
    for (unsigned i = 0; i < lanes; i++) if ((unsigned)e.warp_slot[w][i] == v) r |= 1u << i;
    e.sync_warp();
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }
static inline unsigned __reduce_add_sync(unsigned mask, unsigned v) {
    emu::need_full_mask(mask);
    emu::Engine &e = emu::engine();
    const unsigned w = emu::tls.warp;
    e.warp_slot[w][emu::tls.lane] = v;
    e.sync_warp();
    unsigned r = 0;
    const unsigned lanes = e.block - w * 32 < 32 ? e.block - w * 32 : 32;
    for (unsigned i = 0; i < lanes; i++) r += (unsigned)e.warp_slot[w][i];
    e.sync_warp();
    return r;
}

static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {
    emu::need_full_mask(mask);
    emu::Engine &e = emu::engine();
    const unsigned w = emu::tls.warp;
    e.warp_slot[w][emu::tls.lane] = v;
    e.sync_warp();
    unsigned r = 0;
    const unsigned lanes = e.block - w * 32 < 32 ? e.block - w * 32 : 32;
    for (unsigned i = 0; i < lanes; i++) r = (unsigned)e.warp_slot[w][i] > r ? (unsigned)e.warp_slot[w][i] : r;
    e.sync_warp();
    return r;
}

// ------------------------------------------------------------------------------------------------ intrinsics
static inline unsigned __brev(unsigned x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(x);
}
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __vcmpeq4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xFF) == ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i);
    return r;
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }

template <class A, class B> static inline typename std::common_type<A, B>::type min(A a, B b) { typedef typename std::common_type<A, B>::type T; return (T)a < (T)b ? (T)a : (T)b; }
template <class A, class B> static inline typename std::common_type<A, B>::type max(A a, B b) { typedef typename std::common_type<A, B>::type T; return (T)a > (T)b ? (T)a : (T)b; }

template <class T, class V> static inline T atomicAdd(T *p, V v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class V> static inline T atomicSub(T *p, V v) { return __atomic_fetch_sub(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class V> static inline T atomicOr(T *p, V v) { return __atomic_fetch_or(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class V> static inline T atomicAnd(T *p, V v) { return __atomic_fetch_and(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class V> static inline T atomicExch(T *p, V v) { return __atomic_exchange_n(p, (T)v, __ATOMIC_SEQ_CST); }
template <class T, class V> static inline T atomicMin(T *p, V v) {
    T cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while ((T)v < cur && !__atomic_compare_exchange_n(p, &cur, (T)v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return cur;
}
template <class T, class V> static inline T atomicMax(T *p, V v) {
    T cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while ((T)v > cur && !__atomic_compare_exchange_n(p, &cur, (T)v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return cur;
}
template <class T, class A, class B> static inline T atomicCAS(T *p, A cmp, B val) {
    T expected = (T)cmp;
    __atomic_compare_exchange_n(p, &expected, (T)val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;
}

// ------------------------------------------------------------------------------------------------ runtime API
typedef enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600, cudaErrorUnknown = 999 } cudaError_t;
typedef void *cudaStream_t;
struct emuEvent { timespec t; bool pending = false; };
typedef emuEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaHostAllocDefault = 0, cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttributeValue { int programmaticStreamSerializationAllowed; };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; cudaLaunchAttributeValue val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes = 0; cudaStream_t stream = nullptr; cudaLaunchAttribute *attrs = nullptr; unsigned numAttrs = 0; };
struct cudaDeviceProp { int multiProcessorCount; char name[64]; size_t totalGlobalMem; };

static inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 16; return cudaSuccess; }   // (all ordinals are the same host memory)
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    const char *s = getenv("CDBG_EMU_SMS");
    p->multiProcessorCount = s ? atoi(s) : 2;
    snprintf(p->name, sizeof p->name, "CPU emulation (tests/emu)");
    p->totalGlobalMem = (size_t)180 << 30;
    return cudaSuccess;
}
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) {
    void *q = nullptr;
    if (posix_memalign(&q, 512, (n + 511) & ~size_t(511))) { *p = nullptr; return cudaErrorMemoryAllocation; }
    memset(q, 0xCD, (n + 511) & ~size_t(511));   // poison: device memory is not zero-initialised
    *p = (T *)q;
    return cudaSuccess;
}
// Kernels run synchronously here, so the one thing a missing stream synchronisation can break on a GPU -- the host reading the
// destination of an asynchronous device -> host copy (or the time of an event) too early -- is modelled explicitly: such a copy
// snapshots its source in stream order, POISONS the destination, and only delivers at the next synchronising call of the thread.
namespace emu {
struct PendingCopy { void *dst; std::vector<unsigned char> data; };
inline thread_local std::vector<PendingCopy> pending_d2h;
inline thread_local std::vector<emuEvent *> pending_events;
inline void stream_sync() {
    for (auto &c : pending_d2h) memcpy(c.dst, c.data.data(), c.data.size());
    pending_d2h.clear();
    for (auto *e : pending_events) e->pending = false;
    pending_events.clear();
}
}  // namespace emu
template <class T> static inline cudaError_t cudaFree(T *p) { emu::stream_sync(); free((void *)p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaHostAlloc(T **p, size_t n, unsigned) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaFreeHost(T *p) { emu::stream_sync(); free((void *)p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind kind, cudaStream_t) {
    if (kind == cudaMemcpyDeviceToHost && n) {
        emu::PendingCopy c; c.dst = dst; c.data.assign((const unsigned char *)src, (const unsigned char *)src + n);
        memset(dst, 0xEE, n);
        emu::pending_d2h.push_back(std::move(c));
        return cudaSuccess;
    }
    emu::stream_sync();          // (a later copy in the same stream sees the earlier ones done)
    memmove(dst, src, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { emu::stream_sync(); memmove(dst, src, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *dst, int v, size_t n, cudaStream_t) { memset(dst, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { emu::stream_sync(); return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { emu::stream_sync(); return cudaSuccess; }
static inline cudaError_t cudaStreamQuery(cudaStream_t) { emu::stream_sync(); return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }   // (the emulated streams run in issue order)
static inline cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b) { *free_b = *total_b = (size_t)180 << 30; return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emuEvent(); return cudaSuccess; }
#define cudaEventDisableTiming 2u
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new emuEvent(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    auto &v = emu::pending_events; v.erase(std::remove(v.begin(), v.end(), e), v.end());
    delete e; return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    clock_gettime(CLOCK_MONOTONIC, &e->t);
    if (!e->pending) { e->pending = true; emu::pending_events.push_back(e); }
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { emu::stream_sync(); return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    if (a->pending || b->pending) { *ms = 0; return cudaErrorNotReady; }      // a GPU answers the same before the events have completed
    *ms = (float)((b->t.tv_sec - a->t.tv_sec) * 1e3 + (b->t.tv_nsec - a->t.tv_nsec) * 1e-6);
    return cudaSuccess;
}
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
// CDBG_EMU_PINNED=1 makes every host pointer look pinned + mapped, so that the zero-copy path of cdbg_build can be exercised
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
    const char *s = getenv("CDBG_EMU_PINNED");
    const bool pinned = s && atoi(s) != 0;
    a->type = pinned ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
    a->device = 0; a->devicePointer = pinned ? const_cast<void *>(p) : nullptr; a->hostPointer = const_cast<void *>(p);
    return cudaSuccess;
}
template <class F> static inline cudaError_t cudaFuncSetAttribute(F f, cudaFuncAttribute, int v) {
    if (v < 0 || v > 227 * 1024) { fprintf(stderr, "[emu] cudaFuncSetAttribute: %d B of dynamic shared memory is above the 227 KB limit\n", v); abort(); }
    std::lock_guard<std::mutex> l(emu::smem_optin_mutex());
    emu::smem_optin()[(const void *)f] = (size_t)v;
    return cudaSuccess;
}

template <class... Params, class... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t *cfg, void (*kern)(Params...), Args &&...args) {
    std::tuple<typename std::decay<Params>::type...> a{static_cast<typename std::decay<Params>::type>(args)...};
    emu::launch_kernel((const void *)kern, "cudaLaunchKernelEx", cfg->gridDim.x, cfg->blockDim.x, cfg->dynamicSmemBytes, [&] { std::apply(kern, a); });
    return cudaSuccess;
}
