// cdbg.cu -- host side of libcdbg_b200.so: context, device memory, stage orchestration and the C ABI of include/cdbg.h.
//
// The library is the product path: every stage runs as CUDA kernels from kernels.cuh on one B200; there is no CPU
// implementation behind these entry points.  Stage boundaries that are disk files in the reference
// (.h5 partitions -> .glue.N -> .unitigs.fa -> .links.N) are device arrays here.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/cdbg.h"
#include "kernels.cuh"
#include "shard.cuh"
#include "k1_tma.cuh"
#include "k2_smem.cuh"
#include "k5_smem.cuh"
#include "sk.cuh"
#include "parse.cuh"
#include "wide.cuh"

namespace {

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
    bool used = false;
};

enum Stage { ST_EMPTY = 0, ST_PUSHED, ST_COUNTED, ST_COMPACTED, ST_LINKED };

enum { T_PUSH = 0, T_COUNT, T_COMPACT, T_LINK, T_FETCH, T_K1H, T_K1S, T_K2, T_K3, T_K5, T_K6R, T_K6S, T_K2H, T_K2S, T_K2C, T_X1, T_X2, T_X3, T_GLUE, T_XLINK, T_N };

}  // namespace

static std::atomic<int> g_shard_abort{0};   // a sharded context of this process failed: its sibling ranks stop waiting for it (stream_sync)

struct cdbg_ctx {
    cdbg_params prm{};
    int k = 0, m = 0;
    int W = 1;                     // 64-bit words per k-mer: 1 for k <= 31, ceil(2k/64) rounded like the reference spans above
    uint32_t amin = 1, amax = 0x7FFFFFFFu;
    int n_sm = 148;
    uint64_t total_mem = 0;       // device memory (cudaDeviceProp::totalGlobalMem), for the wave planning of stage 1
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;                 // sharded glue: the all-gather of the piece strings runs beside the piece-level ranking
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    std::string err;
    Stage stage = ST_EMPTY;
    cdbg_stats st{};
    Timer tm[T_N];
    uint32_t launches = 0;
    bool use_pdl = true;           // CDBG_PDL=0 disables programmatic dependent launch (A/B)
    bool trace = false;            // CDBG_TRACE=1: print a host-side timeline (stream-synchronised, debugging only)
    double trace_t0 = 0;

    // device block cache behind dalloc/dfree (see dalloc_bytes)
    struct Block { void *p; size_t cap; bool used; };
    std::vector<Block> blocks;
    uint64_t last_n_bytes = 0;

    // reads
    uint8_t *d_bytes = nullptr;
    uint64_t n_bytes = 0, cap_bytes = 0;
    bool bytes_borrowed = false;
    bool bytes_hostmapped = false;        // d_bytes is pinned HOST memory (cdbg_build zero-copy): only k1t_single may read it
    const uint8_t *h_bytes_src = nullptr; // the caller's host pointer behind it
    bool shard_failed = false;    // a collective was given up: the communicator is aborted, not destroyed
    uint64_t sk_solid_cap = 0;    // capacity of d_solid_kmers / d_solid_counts while the waves of stage 1 append to them
    uint8_t *d_synth = nullptr;   // owned by the context across resets (bench input)
    uint64_t synth_cap = 0;

    // stage 1 results
    uint64_t *d_solid_kmers = nullptr;
    uint32_t *d_solid_counts = nullptr;
    uint64_t S = 0;

    // stage 2/3 results
    uint32_t *d_link = nullptr;
    uint64_t *d_edges = nullptr;
    uint64_t n_edges = 0, edge_cap = 0;
    uint64_t *d_info = nullptr;
    uint32_t *d_unitig_of_kmer = nullptr;
    char *d_seq = nullptr;
    uint64_t *d_seq_off = nullptr;
    unsigned long long *d_kc = nullptr;
    uint32_t *d_abund = nullptr;
    uint32_t *d_term_begin = nullptr, *d_term_end = nullptr;   // sharded mode: k-mer ends at which each local piece begins / ends
    uint64_t U = 0, total_nt = 0;

    // stage 4
    LinkRec *d_links = nullptr;
    uint64_t *d_link_off = nullptr;   // [U+1] CSR offsets of d_links by `from`

    // sharded mode (world > 1)
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    uint64_t *d_link64 = nullptr;            // [2S] partner gid of every local k-mer end
    ulonglong2 *d_sedges = nullptr;          // unitig-level edges (from gid, to gid) this GPU is responsible for
    uint64_t n_sedges = 0, sedge_cap = 0;
    uint64_t *d_pinfo = nullptr;             // [G] replicated piece-level ranking result
    uint32_t *d_unitig_of_piece = nullptr;   // [G]
    GidTable gid_table{};                    // replicated k-mer-end gid -> piece terminus lookup (glue + links)
    uint64_t piece_base_me = 0, n_pieces = 0, U_total = 0, u_first = 0;

    // small device scratch for counters (zeroed per use)
    unsigned long long *d_ctr = nullptr;   // [64]
    int *d_flag = nullptr;                 // [4]
    unsigned long long *h_ctr = nullptr;   // pinned [64]

    // host result buffers (pinned, grown on demand)
    char *h_seq = nullptr; size_t h_seq_cap = 0;
    uint64_t *h_seq_off = nullptr; size_t h_seq_off_cap = 0;
    uint64_t *h_kc = nullptr; size_t h_kc_cap = 0;
    uint32_t *h_abund = nullptr; size_t h_abund_cap = 0;
    cdbg_link_t *h_links = nullptr; size_t h_links_cap = 0;
    uint64_t *h_link_off = nullptr; size_t h_link_off_cap = 0;

    int fail(int code, const char *fmt, ...) {
        char buf[1024];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        err = buf;
        if (prm.flags & CDBG_FLAG_VERBOSE) fprintf(stderr, "[cdbg] error: %s\n", buf);
        if (world > 1) { shard_failed = true; g_shard_abort.store(1, std::memory_order_relaxed); }   // sibling ranks in this process stop waiting (stream_sync)
        return code;
    }
};

namespace {

#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess)                                                                                     \
            return ctx->fail(e_ == cudaErrorMemoryAllocation ? CDBG_ERR_OOM : CDBG_ERR_CUDA, "%s:%d %s -> %s", __FILE__, \
                             __LINE__, #call, cudaGetErrorString(e_));                                             \
    } while (0)

inline double host_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }   // milliseconds

// Waiting for the stream.  A sharded context waits by polling: a rank that failed earlier never enters the next collective, and the
// NCCL kernels of the other ranks would spin for ever inside cudaStreamSynchronize.  Polling lets them give up -- at once when the
// failed rank lives in this process (g_shard_abort, set by cdbg_ctx::fail), after CDBG_SHARD_TIMEOUT_S (default 120) otherwise --
// so that every rank returns an error instead of hanging (the context is then only good for cdbg_destroy, which aborts the communicator).
int stream_sync(cdbg_ctx *ctx) {
    if (ctx->world <= 1) { CK(cudaStreamSynchronize(ctx->stream)); return CDBG_OK; }
    static const double limit_s = getenv("CDBG_SHARD_TIMEOUT_S") ? atof(getenv("CDBG_SHARD_TIMEOUT_S")) : 120.0;
    const double t0 = host_now();
    for (uint64_t spins = 0;; spins++) {
        const cudaError_t e = cudaStreamQuery(ctx->stream);
        if (e == cudaSuccess) return CDBG_OK;
        if (e != cudaErrorNotReady) return ctx->fail(CDBG_ERR_CUDA, "cudaStreamQuery -> %s", cudaGetErrorString(e));
        if ((spins & 1023) == 1023) {
            if (g_shard_abort.load(std::memory_order_relaxed)) { ctx->shard_failed = true; return ctx->fail(CDBG_ERR_NCCL, "another rank of this process failed: giving up the collective"); }
            if (host_now() - t0 > limit_s * 1e3) { ctx->shard_failed = true; return ctx->fail(CDBG_ERR_NCCL, "a peer GPU did not reach the collective within %.0f s (it failed earlier?): giving up", limit_s); }
            if (spins > (1u << 16)) { struct timespec ts = {0, 20000}; nanosleep(&ts, nullptr); }
        }
    }
}
#define SYNC_STREAM(ctx)                          \
    do {                                          \
        int s_ = stream_sync(ctx);                \
        if (s_) return s_;                        \
    } while (0)

// CDBG_EMU_ACTIVE is only ever defined by tests/emu/cuda_runtime.h (the CPU emulation build used by the no-GPU logic tests)
#ifdef CDBG_EMU_ACTIVE
#define KLAUNCH(kern, grid, block, smem, ...) emu::launch_kernel((const void *)(kern), #kern, (unsigned)(grid), (unsigned)(block), (size_t)(smem), [&] { kern(__VA_ARGS__); })
#else
#define KLAUNCH(kern, grid, block, smem, ...) kern<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__)
#endif
#define LAUNCH(kern, grid, block, smem, ...)                                   \
    do {                                                                       \
        KLAUNCH(kern, grid, block, smem, __VA_ARGS__);                         \
        ctx->launches++;                                                       \
        CK(cudaGetLastError());                                                \
    } while (0)

inline void trace_point(cdbg_ctx *ctx, const char *label) {
    if (!ctx->trace) return;
    cudaStreamSynchronize(ctx->stream);
    double t = host_now();
    fprintf(stderr, "[cdbg trace r%d] %-28s +%8.3f ms\n", ctx->rank, label, t - ctx->trace_t0);
    ctx->trace_t0 = t;
}
#define TRACE(label) trace_point(ctx, label)

// Device memory: a per-context block cache.  Every stage allocates through dalloc/dfree; freed blocks are kept and handed
// out again (best fit, at most 25% + 1 MB larger than asked), so a repeated build touches the driver allocator only while
// warming up.  All work of a context is on ONE stream, so reuse right after dfree is safe in stream order -- the semantics of
// cudaMallocAsync, which this replaced because its pool re-mapped ~GB blocks on every build once NCCL was active (2.4 ms).
int dalloc_bytes(cdbg_ctx *ctx, void **out, size_t bytes) {
    bytes = (std::max<size_t>(bytes, 256) + 511) & ~size_t(511);
    int best = -1;
    for (size_t i = 0; i < ctx->blocks.size(); i++) {
        const auto &b = ctx->blocks[i];
        if (b.used || b.cap < bytes || b.cap > bytes + bytes / 4 + (1u << 20)) continue;
        if (best < 0 || b.cap < ctx->blocks[best].cap) best = (int)i;
    }
    if (best >= 0) { ctx->blocks[best].used = true; *out = ctx->blocks[best].p; return CDBG_OK; }
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) {   // give cached-but-unused blocks back and retry once
        cudaGetLastError();
        cudaStreamSynchronize(ctx->stream);
        for (size_t i = 0; i < ctx->blocks.size();) {
            if (!ctx->blocks[i].used) { cudaFree(ctx->blocks[i].p); ctx->blocks.erase(ctx->blocks.begin() + i); } else i++;
        }
        CK(cudaMalloc(&q, bytes));
    }
    ctx->blocks.push_back({q, bytes, true});
    *out = q;
    return CDBG_OK;
}

// launch with programmatic stream serialization (the kernel must start with pdl_wait(), see kernels.cuh)
#define LAUNCH_PDL(kern, grid, block, smem, ...)                                                             \
    do {                                                                                                     \
        cudaLaunchConfig_t cfg_ = {};                                                                        \
        cfg_.gridDim = dim3((unsigned)(grid)); cfg_.blockDim = dim3((unsigned)(block));                      \
        cfg_.dynamicSmemBytes = (smem); cfg_.stream = ctx->stream;                                           \
        cudaLaunchAttribute at_[1];                                                                          \
        at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                      \
        at_[0].val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;                            \
        cfg_.attrs = at_; cfg_.numAttrs = 1;                                                                 \
        ctx->launches++;                                                                                     \
        CK(cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__));                                                    \
    } while (0)

template <typename T>
int dalloc(cdbg_ctx *ctx, T **p, size_t n_elems) {
    void *q = nullptr;
    int rc = dalloc_bytes(ctx, &q, n_elems * sizeof(T));
    *p = reinterpret_cast<T *>(q);
    return rc;
}

template <typename T>
int dfree(cdbg_ctx *ctx, T *&p) {
    if (!p) return CDBG_OK;
    for (auto &b : ctx->blocks) if (b.p == (void *)p) { b.used = false; break; }
    p = nullptr;
    return CDBG_OK;
}

template <typename T>
int hgrow(cdbg_ctx *ctx, T **p, size_t *cap, size_t n_elems) {
    if (n_elems <= *cap && *p) return CDBG_OK;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    size_t want = std::max<size_t>(n_elems + n_elems / 8, 1024);
    void *q = nullptr;
    CK(cudaHostAlloc(&q, want * sizeof(T), cudaHostAllocDefault));
    *p = reinterpret_cast<T *>(q); *cap = want;
    return CDBG_OK;
}

int tstart(cdbg_ctx *ctx, int i) {
    Timer &t = ctx->tm[i];
    if (!t.a) { CK(cudaEventCreate(&t.a)); CK(cudaEventCreate(&t.b)); }
    CK(cudaEventRecord(t.a, ctx->stream));
    return CDBG_OK;
}
int tstop(cdbg_ctx *ctx, int i) {
    Timer &t = ctx->tm[i];
    CK(cudaEventRecord(t.b, ctx->stream));
    t.used = true;
    return CDBG_OK;
}
float tread(cdbg_ctx *ctx, int i) {
    Timer &t = ctx->tm[i];
    if (!t.used) return 0.f;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, t.a, t.b) != cudaSuccess) { cudaGetLastError(); return 0.f; }
    return ms;
}

inline int grid_for(cdbg_ctx *ctx, uint64_t n, int block, int per_sm) {
    uint64_t need = (n + block - 1) / block;
    uint64_t cap = (uint64_t)ctx->n_sm * per_sm;
    return (int)std::max<uint64_t>(1, std::min(need, cap));
}


// zero the counter scratch
int ctr_zero(cdbg_ctx *ctx) {
    CK(cudaMemsetAsync(ctx->d_ctr, 0, 64 * sizeof(unsigned long long), ctx->stream));
    CK(cudaMemsetAsync(ctx->d_flag, 0, 4 * sizeof(int), ctx->stream));
    return CDBG_OK;
}
int ctr_read(cdbg_ctx *ctx) {
    CK(cudaMemcpyAsync(ctx->h_ctr, ctx->d_ctr, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_ctr + 64, ctx->d_flag, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SYNC_STREAM(ctx);
    return CDBG_OK;
}

// exclusive scan of a u32 array into u64, total to *h_total (synchronises)
int exclusive_scan(cdbg_ctx *ctx, const uint32_t *d_in, uint64_t n, uint64_t *d_out, uint64_t *h_total) {
    uint64_t n_tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (n_tiles == 0) { *h_total = 0; return CDBG_OK; }
    uint64_t *d_sums = nullptr;
    int rc = dalloc(ctx, &d_sums, n_tiles + 1);
    if (rc) return rc;
    LAUNCH(scan_tile_sums, (unsigned)n_tiles, SCAN_THREADS, 0, d_in, n, d_sums);
    LAUNCH(scan_sums_inplace, 1, SCAN_THREADS, 0, d_sums, n_tiles, d_sums + n_tiles);
    LAUNCH(scan_apply, (unsigned)n_tiles, SCAN_THREADS, 0, d_in, n, d_sums, d_out);
    CK(cudaMemcpyAsync(ctx->h_ctr + 32, d_sums + n_tiles, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SYNC_STREAM(ctx);
    *h_total = ctx->h_ctr[32];
    return dfree(ctx, d_sums);
}

int free_results(cdbg_ctx *ctx) {
    for (auto &b : ctx->blocks) b.used = false;
    ctx->d_bytes = nullptr; ctx->n_bytes = 0; ctx->cap_bytes = 0; ctx->bytes_borrowed = false; ctx->bytes_hostmapped = false; ctx->h_bytes_src = nullptr;
    ctx->d_solid_kmers = nullptr; ctx->d_solid_counts = nullptr; ctx->S = 0;
    ctx->d_link = nullptr; ctx->d_edges = nullptr; ctx->n_edges = 0; ctx->edge_cap = 0;
    ctx->d_info = nullptr; ctx->d_unitig_of_kmer = nullptr; ctx->d_seq = nullptr; ctx->d_seq_off = nullptr;
    ctx->d_kc = nullptr; ctx->d_abund = nullptr; ctx->U = 0; ctx->total_nt = 0; ctx->d_links = nullptr; ctx->d_link_off = nullptr; ctx->d_term_begin = nullptr; ctx->d_term_end = nullptr;
    ctx->d_link64 = nullptr; ctx->d_sedges = nullptr; ctx->n_sedges = 0; ctx->sedge_cap = 0; ctx->d_pinfo = nullptr; ctx->d_unitig_of_piece = nullptr; ctx->gid_table = GidTable{};
    ctx->piece_base_me = 0; ctx->n_pieces = 0; ctx->U_total = 0; ctx->u_first = 0;
    ctx->stage = ST_EMPTY;
    memset(&ctx->st, 0, sizeof ctx->st);
    for (auto &t : ctx->tm) t.used = false;
    ctx->launches = 0;
    return CDBG_OK;
}

// ------------------------------------------------------------------------------------------------ stage 1

#ifndef K1_512_DEFAULT
#define K1_512_DEFAULT 0   // k1t_single with 512 threads x 16 positions (twice the resident warps): opt-in (CDBG_K1_THREADS=512) until timed on a GPU
#endif
#ifndef K2P_REGIONS_DEFAULT
#define K2P_REGIONS_DEFAULT 0   // fine partition of K2 into fixed-capacity regions (no histogram pass): opt-in (CDBG_K2P=regions) until timed on a GPU
#endif
#ifndef K2C_HASH_MUL32_DEFAULT
#define K2C_HASH_MUL32_DEFAULT 1   // cheaper slot hash of the counting kernels (k2c_hash_mul32; C2: table pass 1.246 -> 1.229 ms, k-mer path 0.846 -> 0.814 ms);
                                   // CDBG_K2C_HASH=mix64 selects the murmur finaliser
#endif
#ifndef K2C_FLAT_DEFAULT
#define K2C_FLAT_DEFAULT 0   // flat (one probe step per iteration) insert loop of k2c_count: opt-in (CDBG_K2C=flat) until timed on a GPU
#endif
#ifndef RANK_MATCH_DEFAULT
#define RANK_MATCH_DEFAULT 0   // tile-local ranks by match.any instead of shared-memory atomics: opt-in (CDBG_RANK=match) until timed on a GPU
#endif
bool rank_by_match() {
    const char *r = getenv("CDBG_RANK");
    return RANK_MATCH_DEFAULT ? !(r && strcmp(r, "atomic") == 0) : (r && strcmp(r, "match") == 0);
}

// Sizing of the shared-memory K2 variant.  slots = table size per CTA, fill = k-mer INSTANCES per slot a fine bin is sized for:
// 1.5 suits sequencing coverage (30x data has ~0.3 distinct k-mers per instance, i.e. a load factor below 0.5); when an input
// has more distinct k-mers than that, k2c_count splits the affected bins on the fly, so the value is a tuning knob only.
struct K2SmemPlan { bool enabled; uint32_t slots; double fill; uint32_t P1; };

K2SmemPlan k2_smem_plan(uint64_t n_bytes) {
    K2SmemPlan pl{};
    const char *mode = getenv("CDBG_K2");
    pl.enabled = !(mode && strcmp(mode, "l2") == 0);  // default: shared-memory tables (C2 on a B200: K2 2.15 -> 1.52 ms); CDBG_K2=l2 for A/B
    pl.slots = getenv("CDBG_K2S_SLOTS") ? (uint32_t)atoi(getenv("CDBG_K2S_SLOTS")) : 8192u;
    if (pl.slots != 4096 && pl.slots != 8192 && pl.slots != 16384) pl.slots = 8192;
    pl.fill = getenv("CDBG_K2S_FILL") ? atof(getenv("CDBG_K2S_FILL")) : 1.5;
    if (!(pl.fill >= 0.05 && pl.fill <= 8.0)) pl.fill = 1.5;
    const double per_fine = pl.fill * pl.slots;
    const double n_fine = std::max(1.0, (double)n_bytes * 0.8 / per_fine);
    uint64_t p1 = (uint64_t)ceil(sqrt(n_fine));
    if (getenv("CDBG_K2S_P1")) p1 = strtoull(getenv("CDBG_K2S_P1"), nullptr, 10);
    pl.P1 = (uint32_t)std::min<uint64_t>(K1_MAX_BINS, std::max<uint64_t>(1, p1));
    return pl;
}

// K2 with shared-memory tables: fine histogram -> scan -> fine scatter -> one CTA per fine bin (k2_smem.cuh).  d_kmers holds the
// I instances in P coarse bins (h_cnt, d_bin_off).  Leaves ctx->d_solid_kmers / d_solid_counts / S like the L2 path.
// `h_group` (optional): group of every coarse bin (see K2PParams::group), `n_groups` of them; fine bins are then per group.
int stage_count_k2_smem(cdbg_ctx *ctx, const K2SmemPlan &pl, const uint64_t *d_kmers, uint64_t I, uint32_t P, const std::vector<uint64_t> &h_cnt,
                        const std::vector<uint64_t> &h_off, const std::vector<uint32_t> *h_group = nullptr, uint32_t n_groups = 0, bool allow_regions = true) {
    int rc;
    // CDBG_K2P=regions (opt-in, not timed yet): no fine histogram pass, fixed-capacity regions per fine bin, exact redo on overflow
    const char *k2p = getenv("CDBG_K2P");
    const bool regions = allow_regions && !h_group && (K2P_REGIONS_DEFAULT ? !(k2p && strcmp(k2p, "exact") == 0) : (k2p && strcmp(k2p, "regions") == 0));
    const double per_fine = pl.fill * pl.slots;
    const uint32_t G = h_group ? n_groups : P;
    const uint32_t F = (uint32_t)std::min<uint64_t>(K2P_MAX_F, std::max<uint64_t>(1, (uint64_t)ceil((double)I / (per_fine * G))));
    const uint64_t NF = (uint64_t)G * F;
    std::vector<uint32_t> h_tile_first(P + 1);
    uint64_t tiles = 0;
    for (uint32_t b = 0; b < P; b++) { h_tile_first[b] = (uint32_t)tiles; tiles += (h_cnt[b] + K2P_TILE - 1) / K2P_TILE; }
    h_tile_first[P] = (uint32_t)tiles;
    if (tiles >= (1ULL << 32)) return ctx->fail(CDBG_ERR_PARAM, "too many k-mer tiles for one device");

    uint32_t *d_tile_first = nullptr, *d_fine_cnt = nullptr, *d_fine_cursor = nullptr;
    uint64_t *d_fine_off = nullptr, *d_kmers2 = nullptr, *d_bin_range = nullptr;
    std::vector<uint64_t> h_range(2 * (size_t)P);   // [begin of bin 0..P-1 | end of bin 0..P-1] inside d_kmers
    for (uint32_t b = 0; b < P; b++) { h_range[b] = h_off[b]; h_range[P + b] = h_off[b] + h_cnt[b]; }
    if ((rc = dalloc(ctx, &d_bin_range, 2 * (size_t)P))) return rc;
    CK(cudaMemcpyAsync(d_bin_range, h_range.data(), 2 * (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = dalloc(ctx, &d_tile_first, P + 1))) return rc;
    uint32_t *d_group = nullptr;
    if (h_group) {
        if ((rc = dalloc(ctx, &d_group, P))) return rc;
        CK(cudaMemcpyAsync(d_group, h_group->data(), (size_t)P * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = dalloc(ctx, &d_fine_cnt, NF + 1))) return rc;
    if ((rc = dalloc(ctx, &d_fine_cursor, NF + 1))) return rc;
    if ((rc = dalloc(ctx, &d_fine_off, 2 * NF + 2))) return rc;     // exact: prefix offsets [NF + 1]; regions: begins [NF] then ends [NF]
    uint64_t *d_region_base = nullptr; uint32_t *d_region_cap = nullptr;
    uint64_t out_cap = I;
    if (regions) {
        std::vector<uint64_t> h_base(P); std::vector<uint32_t> h_cap(P);
        out_cap = 0;
        for (uint32_t b = 0; b < P; b++) {
            const uint64_t mean = (h_cnt[b] + F - 1) / F;
            h_cap[b] = (uint32_t)std::min<uint64_t>(mean + mean / 16 + 256, 0xFFFFFFFFu);
            h_base[b] = out_cap; out_cap += (uint64_t)F * h_cap[b];
        }
        if ((rc = dalloc(ctx, &d_region_base, P))) return rc;
        if ((rc = dalloc(ctx, &d_region_cap, P))) return rc;
        CK(cudaMemcpyAsync(d_region_base, h_base.data(), (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_region_cap, h_cap.data(), (size_t)P * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = dalloc(ctx, &d_kmers2, out_cap + 1))) return rc;
    CK(cudaMemsetAsync(d_fine_cnt, 0, (NF + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(d_fine_cursor, 0, (NF + 1) * 4, ctx->stream));
    CK(cudaMemcpyAsync(d_tile_first, h_tile_first.data(), (P + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));

    tstart(ctx, T_K2);
    K2PParams pp{};
    pp.kmers = d_kmers; pp.bin_beg = d_bin_range; pp.bin_end = d_bin_range + P; pp.tile_first = d_tile_first; pp.group = d_group; pp.P1 = P; pp.F = F; pp.fine_shift = h_group ? 13u : 32u; pp.n_tiles = (uint32_t)tiles;
    pp.fine_cnt = d_fine_cnt; pp.fine_off = d_fine_off; pp.fine_cursor = d_fine_cursor; pp.out = d_kmers2;
    pp.region_base = d_region_base; pp.region_cap = d_region_cap; pp.overflow = ctx->d_flag + 2;
    const uint64_t *d_fine_beg = d_fine_off, *d_fine_end = d_fine_off + 1;
    if (regions) {
        if ((rc = ctr_zero(ctx))) return rc;       // the overflow flag must survive until the counters are read after k2c_count
        d_fine_end = d_fine_off + NF;
    }
    if (tiles) {
        if (!regions) {
            tstart(ctx, T_K2H);
            LAUNCH(k2p_hist, (int)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * 4), K2P_THREADS, 0, pp);
            tstop(ctx, T_K2H);
            uint64_t total = 0;
            if ((rc = exclusive_scan(ctx, d_fine_cnt, NF + 1, d_fine_off, &total))) return rc;   // entry NF = I
            if (total != I) return ctx->fail(CDBG_ERR_INTERNAL, "fine histogram lost k-mers (%llu != %llu)", (unsigned long long)total, (unsigned long long)I);
        }
        const bool match = rank_by_match() && F <= K2P_MATCH_MAX_F;
        const size_t sm = k2p_scatter_smem_bytes(F, match);
        tstart(ctx, T_K2S);
        if (match) {
            CK(cudaFuncSetAttribute(k2p_scatter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            LAUNCH(k2p_scatter<true>, (int)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * 2), K2P_THREADS, sm, pp);
        } else {
            CK(cudaFuncSetAttribute(k2p_scatter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            LAUNCH(k2p_scatter<false>, (int)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * 2), K2P_THREADS, sm, pp);
        }
        tstop(ctx, T_K2S);
        if (regions) LAUNCH(k2p_region_ranges, grid_for(ctx, NF, 256, 8), 256, 0, d_region_base, d_region_cap, d_fine_cursor, F, NF, d_fine_off, d_fine_off + NF);
    } else {
        CK(cudaMemsetAsync(d_fine_off, 0, (2 * NF + 2) * 8, ctx->stream));
    }
    TRACE("K2 fine partition");

    uint64_t solid_cap = std::max<uint64_t>(I <= (1ULL << 28) ? I : std::max<uint64_t>(I / 4, 1ULL << 28), 1);
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = dalloc(ctx, &ctx->d_solid_kmers, solid_cap))) return rc;
        if ((rc = dalloc(ctx, &ctx->d_solid_counts, solid_cap))) return rc;
        if (!(regions && attempt == 0)) { if ((rc = ctr_zero(ctx))) return rc; }
        K2CParams cp{};
        cp.kmers = d_kmers2; cp.fine_beg = d_fine_beg; cp.fine_end = d_fine_end; cp.n_fine = (uint32_t)NF; cp.amin = ctx->amin; cp.amax = ctx->amax;
        cp.solid_kmers = ctx->d_solid_kmers; cp.solid_counts = ctx->d_solid_counts; cp.solid_cap = solid_cap;
        cp.n_solid = ctx->d_ctr + 0; cp.n_distinct = ctx->d_ctr + 1; cp.n_splits = ctx->d_ctr + 2; cp.failed = ctx->d_flag + 1;
        tstart(ctx, T_K2C);
        if (I) {
#define K2C_LAUNCH1(SL, FL, MU, PER_SM)                                                                                                              \
    do {                                                                                                                                            \
        CK(cudaFuncSetAttribute(k2c_count<SL, FL, MU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2c_smem_bytes<SL>()));                     \
        LAUNCH((k2c_count<SL, FL, MU>), (int)std::min<uint64_t>(NF, (uint64_t)ctx->n_sm * (PER_SM)), k2c_threads<SL>(), k2c_smem_bytes<SL>(), cp); \
    } while (0)
#define K2C_LAUNCH(SL, PER_SM)                                                                                  \
    do {                                                                                                        \
        if (flat) { if (mul32) K2C_LAUNCH1(SL, true, true, PER_SM); else K2C_LAUNCH1(SL, true, false, PER_SM); }  \
        else { if (mul32) K2C_LAUNCH1(SL, false, true, PER_SM); else K2C_LAUNCH1(SL, false, false, PER_SM); }     \
    } while (0)
            const char *k2c = getenv("CDBG_K2C"), *hm = getenv("CDBG_K2C_HASH");
            const bool flat = K2C_FLAT_DEFAULT ? !(k2c && strcmp(k2c, "nested") == 0) : (k2c && strcmp(k2c, "flat") == 0);
            const bool mul32 = K2C_HASH_MUL32_DEFAULT ? !(hm && strcmp(hm, "mix64") == 0) : (hm && strcmp(hm, "mul32") == 0);
            if (pl.slots == 4096) K2C_LAUNCH(4096, 4);
            else if (pl.slots == 8192) K2C_LAUNCH(8192, 2);
            else K2C_LAUNCH(16384, 1);
#undef K2C_LAUNCH1
#undef K2C_LAUNCH
        }
        tstop(ctx, T_K2C);
        if ((rc = ctr_read(ctx))) return rc;
        if (regions && attempt == 0 && ((int *)(ctx->h_ctr + 64))[2]) {
            // a fine bin outgrew its region (repeated k-mers): what was counted is incomplete -- redo with exact offsets
            dfree(ctx, ctx->d_solid_kmers); dfree(ctx, ctx->d_solid_counts);
            dfree(ctx, d_bin_range); dfree(ctx, d_tile_first); dfree(ctx, d_group); dfree(ctx, d_fine_cnt); dfree(ctx, d_fine_cursor);
            dfree(ctx, d_fine_off); dfree(ctx, d_kmers2); dfree(ctx, d_region_base); dfree(ctx, d_region_cap);
            ctx->st.k2p_fallback = 1;
            return stage_count_k2_smem(ctx, pl, d_kmers, I, P, h_cnt, h_off, h_group, n_groups, false);
        }
        if (((int *)(ctx->h_ctr + 64))[1]) return ctx->fail(CDBG_ERR_INTERNAL, "K2: a fine bin does not fit a shared-memory table after %d splits", K2C_MAX_LEVEL);
        ctx->st.n_k2_splits = ctx->h_ctr[2];
        ctx->S = ctx->h_ctr[0];
        ctx->st.n_distinct = ctx->h_ctr[1];
        if (ctx->S <= solid_cap) break;
        if (attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "solid buffer overflow");
        dfree(ctx, ctx->d_solid_kmers); dfree(ctx, ctx->d_solid_counts);
        solid_cap = ctx->S;
    }
    tstop(ctx, T_K2);
    ctx->st.n_solid = ctx->S;
    if (ctx->S >= (1ULL << 31)) return ctx->fail(CDBG_ERR_PARAM, "more than 2^31 solid k-mers on one device (%llu): shard over more GPUs", (unsigned long long)ctx->S);
    dfree(ctx, d_bin_range); dfree(ctx, d_tile_first); dfree(ctx, d_group); dfree(ctx, d_fine_cnt); dfree(ctx, d_fine_cursor); dfree(ctx, d_fine_off); dfree(ctx, d_kmers2);
    dfree(ctx, d_region_base); dfree(ctx, d_region_cap);
    return CDBG_OK;
}

int stage_count(cdbg_ctx *ctx) {
    const int k = ctx->k;
    const uint64_t n = ctx->n_bytes;
    int rc;
    tstart(ctx, T_COUNT);

    // ---- K1a: histogram over bins
    uint64_t n_tiles = (n + K1_TILE - 1) / K1_TILE;
    const uint64_t bin_target = getenv("CDBG_BIN_TARGET") ? strtoull(getenv("CDBG_BIN_TARGET"), nullptr, 10) : 5000000ULL;   // instances per bin (C2 sweep: 25 bins beat 49 and 13; fewer launches, table still L2 resident)
    uint32_t P = (uint32_t)std::min<uint64_t>(K1_MAX_BINS, std::max<uint64_t>(1, (n * 4 / 5 + bin_target - 1) / bin_target));
    // K2 variant: "l2" = one L2-resident table per coarse bin (k2_insert/k2_emit); "smem" = fine partition + shared-memory
    // tables (k2_smem.cuh).  With shared-memory tables the coarse fan-out P and the fine fan-out F split the total number of
    // fine bins about evenly, so that both partition passes write runs of similar length.
    K2SmemPlan k2s = k2_smem_plan(n);
    if (k2s.enabled) P = k2s.P1;
    ctx->st.n_count_bins = P;
    unsigned long long *d_bin_count = nullptr, *d_bin_cursor = nullptr;
    uint64_t *d_bin_off = nullptr;
    if ((rc = dalloc(ctx, &d_bin_count, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_cursor, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_off, P + 1))) return rc;
    CK(cudaMemsetAsync(d_bin_count, 0, P * 8, ctx->stream));
    CK(cudaMemsetAsync(d_bin_cursor, 0, P * 8, ctx->stream));

    // K1 variants (CDBG_K1): default "single" = one pass into fixed-capacity bin regions (k1t_single), falling back to the exact
    // two-pass TMA kernels ("twopass": histogram, then scatter) when a bin outgrows its region; "classic" = the earlier LDG-based
    // two-pass kernel, kept for A/B runs.
    const char *k1mode = getenv("CDBG_K1");
    const bool use_classic = k1mode && strcmp(k1mode, "classic") == 0;
    bool use_single = !use_classic && !(k1mode && strcmp(k1mode, "twopass") == 0);
    const bool use_tma = !use_classic;

    std::vector<uint64_t> h_cnt(P), h_off(P + 1);   // per coarse bin: number of k-mers, start inside d_kmers (regions may have gaps)
    uint64_t I = 0, max_bin = 0;
    uint64_t *d_kmers = nullptr;

    if (use_single && n_tiles) {
        // every bin gets the mean number of POSITIONS per bin plus 3 % and a constant: instances <= positions, so a bin only
        // overflows when repeated k-mers pile up in it
        const uint64_t slack = getenv("CDBG_K1_SLACK") ? strtoull(getenv("CDBG_K1_SLACK"), nullptr, 10) : 16384;   // (tests shrink it to force the fallback)
        const uint64_t bin_cap = ((n / P + n / P / 32 + slack) + 31) & ~31ULL;
        if ((rc = dalloc(ctx, &d_kmers, (uint64_t)P * bin_cap + 1))) return rc;
        if ((rc = ctr_zero(ctx))) return rc;
        K1XParams kx{};
        kx.bytes = ctx->d_bytes; kx.n_bytes = n; kx.n_tiles = n_tiles; kx.k = k; kx.P = P;
        kx.bin_cap = bin_cap; kx.bin_cursor = d_bin_cursor; kx.out_kmers = d_kmers; kx.overflow = ctx->d_flag + 2;
        // the ragged last tiles are read with plain loads: keep them in device memory when the reads are mapped host memory
        kx.tail_start = n > 2 * (uint64_t)K1T_RAW ? ((n - 2 * (uint64_t)K1T_RAW) & ~15ULL) : 0;
        uint8_t *d_tail = nullptr;
        if (ctx->bytes_hostmapped) {
            if ((rc = dalloc(ctx, &d_tail, n - kx.tail_start + 16))) return rc;
            CK(cudaMemcpyAsync(d_tail, ctx->h_bytes_src + kx.tail_start, n - kx.tail_start, cudaMemcpyHostToDevice, ctx->stream));
            kx.tail = d_tail;
        } else kx.tail = ctx->d_bytes + kx.tail_start;
        const bool match = rank_by_match() && P <= K1X_MATCH_MAX_P;
        const size_t sm = k1x_smem_bytes(P, match);
        const char *k1t = getenv("CDBG_K1_THREADS");
        const bool wide_cta = K1_512_DEFAULT ? !(k1t && atoi(k1t) == 256) : (k1t && atoi(k1t) == 512);
        tstart(ctx, T_K1S);
        if (wide_cta && !match) {
            CK(cudaFuncSetAttribute(k1t_single512, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            LAUNCH(k1t_single512, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 2), K1H_THREADS, sm, kx);
        } else if (match) {
            CK(cudaFuncSetAttribute(k1t_single<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            LAUNCH(k1t_single<true>, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 2), K1_THREADS, sm, kx);
        } else {
            CK(cudaFuncSetAttribute(k1t_single<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            LAUNCH(k1t_single<false>, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 2), K1_THREADS, sm, kx);
        }
        tstop(ctx, T_K1S);
        CK(cudaMemcpyAsync(h_cnt.data(), d_bin_cursor, P * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = ctr_read(ctx))) return rc;
        dfree(ctx, d_tail);
        if (((int *)(ctx->h_ctr + 64))[2]) {   // a bin outgrew its region: redo with exact offsets
            dfree(ctx, d_kmers);
            CK(cudaMemsetAsync(d_bin_cursor, 0, P * 8, ctx->stream));
            use_single = false;
            ctx->st.k1_fallback = 1;
        } else {
            for (uint32_t b = 0; b < P; b++) { h_off[b] = (uint64_t)b * bin_cap; I += h_cnt[b]; max_bin = std::max(max_bin, h_cnt[b]); }
            h_off[P] = (uint64_t)P * bin_cap;
        }
    }
    if (ctx->bytes_hostmapped && !(use_single && n_tiles)) {
        // the two-pass kernels read the input twice: bring it into HBM first
        uint8_t *nb = nullptr;
        if ((rc = dalloc(ctx, &nb, n + 64))) return rc;
        CK(cudaMemcpyAsync(nb, ctx->h_bytes_src, n, cudaMemcpyHostToDevice, ctx->stream));
        ctx->d_bytes = nb; ctx->cap_bytes = n; ctx->bytes_borrowed = false; ctx->bytes_hostmapped = false;
    }

    if (!(use_single && n_tiles)) {
        K1Params kp{};
        kp.bytes = ctx->d_bytes; kp.n_bytes = n; kp.n_tiles = n_tiles; kp.k = k; kp.P = P;
        kp.bin_count = d_bin_count; kp.bin_off = d_bin_off; kp.bin_cursor = d_bin_cursor; kp.out_kmers = nullptr;
        K1TParams kt{};
        kt.bytes = ctx->d_bytes; kt.n_bytes = n; kt.n_tiles = n_tiles; kt.k = k; kt.P = P;
        kt.bin_count = d_bin_count; kt.bin_off = d_bin_off; kt.bin_cursor = d_bin_cursor; kt.out_kmers = nullptr;
        uint16_t *d_pos_bin = nullptr;
        if (use_tma) { if ((rc = dalloc(ctx, &d_pos_bin, n_tiles * K1_TILE + 8))) return rc; }
        kt.pos_bin = d_pos_bin;
        if (n_tiles) {
            tstart(ctx, T_K1H);
            if (use_tma) {
                size_t sm0 = k1t_smem_bytes(0, P);
                CK(cudaFuncSetAttribute(k1t_extract<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm0));
                CK(cudaFuncSetAttribute(k1t_extract<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1t_smem_bytes(1, P)));
                LAUNCH(k1t_extract<0>, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 6), K1_THREADS, sm0, kt);
            } else {
                size_t sm0 = k1_smem_bytes(0, P), sm1 = k1_smem_bytes(1, P);
                CK(cudaFuncSetAttribute(k1_extract<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm0));
                CK(cudaFuncSetAttribute(k1_extract<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
                LAUNCH(k1_extract<0>, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 6), K1_THREADS, sm0, kp);
            }
            tstop(ctx, T_K1H);
            CK(cudaMemcpyAsync(h_cnt.data(), d_bin_count, P * 8, cudaMemcpyDeviceToHost, ctx->stream));
            SYNC_STREAM(ctx);
        } else std::fill(h_cnt.begin(), h_cnt.end(), 0);
        I = 0; max_bin = 0;
        for (uint32_t b = 0; b < P; b++) { h_off[b] = I; I += h_cnt[b]; max_bin = std::max(max_bin, h_cnt[b]); }
        h_off[P] = I;
        if ((rc = dalloc(ctx, &d_kmers, I + 1))) return rc;
        if (I) {
            CK(cudaMemcpyAsync(d_bin_off, h_off.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
            kp.out_kmers = d_kmers; kt.out_kmers = d_kmers;
            tstart(ctx, T_K1S);
            if (use_tma) LAUNCH(k1t_extract<1>, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 2), K1_THREADS, k1t_smem_bytes(1, P), kt);
            else LAUNCH(k1_extract<1>, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 2), K1_THREADS, k1_smem_bytes(1, P), kp);
            tstop(ctx, T_K1S);
        }
        dfree(ctx, d_pos_bin);
    }
    ctx->st.n_instances = I;

    // ---- K2, shared-memory variant (falls through to the L2 path when a coarse bin is too large for 32-bit fine counters)
    if (k2s.enabled && max_bin < (1ULL << 32)) {
        if ((rc = stage_count_k2_smem(ctx, k2s, d_kmers, I, P, h_cnt, h_off))) return rc;
        dfree(ctx, d_kmers);
        dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
        tstop(ctx, T_COUNT);
        ctx->stage = ST_COUNTED;
        return CDBG_OK;
    }

    // ---- K2: per-bin counting in one L2-resident table
    auto table_cap = [](uint64_t n) { return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1024, (n * 3 / 2 + 63) & ~63ULL), 0xFFFFFF00u); };
    uint32_t cap = table_cap(max_bin);
    uint64_t *d_keys = nullptr; uint32_t *d_counts = nullptr;
    if ((rc = dalloc(ctx, &d_keys, cap))) return rc;
    if ((rc = dalloc(ctx, &d_counts, cap))) return rc;
    CK(cudaMemsetAsync(d_keys, 0xFF, (size_t)cap * 8, ctx->stream));
    CK(cudaMemsetAsync(d_counts, 0, (size_t)cap * 4, ctx->stream));

    uint64_t solid_cap = std::max<uint64_t>(I <= (1ULL << 28) ? I : std::max<uint64_t>(I / 4, 1ULL << 28), 1);
    // Table sizing: the first bin is sized for the worst case (every instance distinct, 1.5 slots each).  Its measured
    // distinct/instance ratio r then sizes the other bins at ~2.5 slots per expected distinct k-mer (hash bins are statistically
    // alike), which shrinks the table scan of k2_emit and its L2 footprint.  If a bin overflows anyway, everything is redone
    // with worst-case tables (attempt 1).
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = dalloc(ctx, &ctx->d_solid_kmers, solid_cap))) return rc;
        if ((rc = dalloc(ctx, &ctx->d_solid_counts, solid_cap))) return rc;
        if ((rc = ctr_zero(ctx))) return rc;
        tstart(ctx, T_K2);
        double slots_per_instance = 1.5;
        bool first = true;
        for (uint32_t b = 0; b < P; b++) {
            uint64_t nb = h_cnt[b];
            if (!nb) continue;
            uint32_t cap_b = std::min(cap, std::max<uint32_t>(1024, (uint32_t)std::min<double>((double)nb * slots_per_instance + 64.0, 4.0e9) & ~63u));
            LAUNCH_PDL(k2_insert, grid_for(ctx, nb, 256, 8), 256, 0, d_kmers + h_off[b], nb, d_keys, d_counts, cap_b, ctx->d_flag);
            LAUNCH_PDL(k2_emit, grid_for(ctx, cap_b, 256, 8), 256, 0, d_keys, d_counts, cap_b, ctx->amin, ctx->amax, ctx->d_solid_kmers,
                   ctx->d_solid_counts, solid_cap, ctx->d_ctr + 0, ctx->d_ctr + 1);
            if (first && attempt == 0 && P > 2) {
                first = false;
                if ((rc = ctr_read(ctx))) return rc;
                double r = (double)ctx->h_ctr[1] / (double)nb;
                slots_per_instance = std::min(1.5, std::max(0.02, 2.5 * r + 0.02));
            }
        }
        tstop(ctx, T_K2);
        if ((rc = ctr_read(ctx))) return rc;
        const bool overflow = ((int *)(ctx->h_ctr + 64))[0] != 0;
        if (overflow && attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "K2 hash table overflow");
        ctx->S = ctx->h_ctr[0];
        ctx->st.n_distinct = ctx->h_ctr[1];
        if (!overflow && ctx->S <= solid_cap) break;
        if (attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "solid buffer overflow");
        dfree(ctx, ctx->d_solid_kmers); dfree(ctx, ctx->d_solid_counts);
        if (!overflow) solid_cap = ctx->S;
        else {   // a partially filled table may hold stale entries: clear both arrays before the conservative rerun
            CK(cudaMemsetAsync(d_keys, 0xFF, (size_t)cap * 8, ctx->stream));
            CK(cudaMemsetAsync(d_counts, 0, (size_t)cap * 4, ctx->stream));
        }
    }
    ctx->st.n_solid = ctx->S;
    if (ctx->S >= (1ULL << 31)) return ctx->fail(CDBG_ERR_PARAM, "more than 2^31 solid k-mers on one device (%llu): shard over more GPUs", (unsigned long long)ctx->S);
    dfree(ctx, d_kmers); dfree(ctx, d_keys); dfree(ctx, d_counts);
    dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
    tstop(ctx, T_COUNT);
    ctx->stage = ST_COUNTED;
    return CDBG_OK;
}

// ------------------------------------------------------------------------------------------------ stages 2+3

#ifndef K6_LIST_DEFAULT
#define K6_LIST_DEFAULT 0   // pointer jumping over shrinking lists of moving elements: opt-in (CDBG_K6=list) until timed on a GPU
#endif

int rank_rounds(cdbg_ctx *ctx, uint64_t *d_pair, uint64_t n2, uint64_t *remaining, bool prefer_list = false) {
    // pointer jumping until no element moves (paths done) or the number of moving elements stops shrinking (cycles)
    uint64_t prev = ~0ULL;
    *remaining = 0;
    const char *mode = getenv("CDBG_K6");
    // prefer_list: the piece-level ranking of the sharded glue -- chains of a handful of pieces, so after the first batch of rounds only a
    // small fraction of the elements still moves (8 GPUs, C3: 69 M elements, the full-array rounds were most of the 29 ms glue)
    const bool use_list = ((K6_LIST_DEFAULT || prefer_list) ? !(mode && strcmp(mode, "full") == 0) : (mode && strcmp(mode, "list") == 0)) && n2 < (1ULL << 32);
    int rc;
    if (use_list) {
        // rounds come in batches of 4; the last round of a batch writes the list of elements that moved, the next batch walks it
        uint32_t *d_list[2] = {nullptr, nullptr};
        if ((rc = dalloc(ctx, &d_list[0], n2 + 1))) return rc;
        if ((rc = dalloc(ctx, &d_list[1], n2 + 1))) return rc;
        const uint32_t *in_list = nullptr;
        uint64_t n_cur = n2;
        int result = -1;
        for (int round = 0, batch = 0; round < 80 && result < 0; round += 4, batch++) {
            if ((rc = ctr_zero(ctx))) return rc;
            uint32_t *out_list = d_list[batch & 1];
            for (int j = 0; j < 4; j++)
                LAUNCH_PDL(k6_jump_list, grid_for(ctx, n_cur, 256, 8), 256, 0, d_pair, in_list, n_cur, j == 3 ? out_list : (uint32_t *)nullptr, ctx->d_ctr + 4,
                           ctx->d_ctr + j);
            ctx->st.rank_rounds += 4;
            if ((rc = ctr_read(ctx))) return rc;
            for (int j = 0; j < 4 && result < 0; j++) {
                uint64_t c = ctx->h_ctr[j];
                if (c == 0) { *remaining = 0; result = 0; }
                else if (c == prev) { *remaining = c; result = 0; }
                prev = c;
            }
            in_list = out_list; n_cur = ctx->h_ctr[4];
        }
        dfree(ctx, d_list[0]); dfree(ctx, d_list[1]);
        if (result == 0) return CDBG_OK;
        return ctx->fail(CDBG_ERR_INTERNAL, "list ranking did not converge");
    }
    for (int round = 0; round < 80; round += 4) {
        rc = ctr_zero(ctx);
        if (rc) return rc;
        for (int j = 0; j < 4; j++) LAUNCH_PDL(k6_jump, grid_for(ctx, n2, 256, 8), 256, 0, d_pair, n2, ctx->d_ctr + j);
        ctx->st.rank_rounds += 4;
        if ((rc = ctr_read(ctx))) return rc;
        for (int j = 0; j < 4; j++) {
            uint64_t c = ctx->h_ctr[j];
            if (c == 0) { *remaining = 0; return CDBG_OK; }
            if (c == prev) { *remaining = c; return CDBG_OK; }
            prev = c;
        }
    }
    return ctx->fail(CDBG_ERR_INTERNAL, "list ranking did not converge");
}

#ifndef K6_TWO_LEVEL_DEFAULT
#define K6_TWO_LEVEL_DEFAULT 1   // window-local ranking in shared memory + global rounds over representatives (C2 on a B200: K6 0.618 -> 0.407 ms,
                                 // profiles/r02_sk_ab3_c2.txt); CDBG_K6=full / list select the one-level variants
#endif

// Two-level ranking of an unweighted list (kernels.cuh: k6_local / k6_jump_list / k6_resolve).  *remaining as rank_rounds.
int rank_two_level(cdbg_ctx *ctx, const uint32_t *d_link, uint64_t *d_pair, uint64_t n2, uint64_t *remaining) {
    int rc;
    *remaining = 0;
    uint32_t *d_list[2] = {nullptr, nullptr};
    if ((rc = dalloc(ctx, &d_list[0], n2 + 1))) return rc;
    if ((rc = dalloc(ctx, &d_list[1], n2 + 1))) return rc;
    if ((rc = ctr_zero(ctx))) return rc;
    const uint64_t n_win = (n2 + 2 * K6L_NV - 1) / (2 * K6L_NV);
    LAUNCH(k6_local, (int)std::min<uint64_t>(n_win, (uint64_t)ctx->n_sm * 4), K6L_THREADS, 0, d_link, d_pair, n2, d_list[0], ctx->d_ctr + 4);
    if ((rc = ctr_read(ctx))) return rc;
    uint64_t n_cur = ctx->h_ctr[4];
    const uint32_t *in_list = d_list[0];
    int result = n_cur ? -1 : 0;
    // Termination.  A round in which nothing moves means every path is ranked.  What still moves after log2(#representatives) + a few
    // rounds is on a cycle: a representative's pointer at least doubles its reach per round once it points at another representative
    // (one extra round to step over a parked element).  NOT "the number of moving elements stopped shrinking", the rule of the
    // one-level ranking: there every chain has an element at every distance from its tail (its closest unfinished element finishes in
    // the next round, so the count strictly shrinks while a path is unfinished), here the representatives are sparse and
    // two consecutive rounds can move the same number of them on plain paths (seen on a B200 with golden case wide_9_k127_a1, where
    // the false alarm made the cycle cutter split a unitig).
    int max_rounds = 8;                                          // 2 * ceil(log2(#representatives)) + 8: with asynchronous updates a pointer can
    for (uint64_t r = 1; r < n_cur; r <<= 1) max_rounds += 2;    // land on a parked element every other round, which costs a round without doubling
    int rounds_done = 0;
    for (int batch = 1; result < 0; batch++) {
        if ((rc = ctr_zero(ctx))) return rc;
        uint32_t *out_list = d_list[batch & 1];
        for (int j = 0; j < 4; j++)
            LAUNCH_PDL(k6_jump_list, grid_for(ctx, n_cur, 256, 8), 256, 0, d_pair, in_list, n_cur, j == 3 ? out_list : (uint32_t *)nullptr, ctx->d_ctr + 4,
                       ctx->d_ctr + j);
        ctx->st.rank_rounds += 4;
        rounds_done += 4;
        if ((rc = ctr_read(ctx))) return rc;
        for (int j = 0; j < 4 && result < 0; j++)
            if (ctx->h_ctr[j] == 0) { *remaining = 0; result = 0; }
        in_list = out_list; n_cur = ctx->h_ctr[4];
        if (result < 0 && n_cur == 0) { *remaining = 0; result = 0; }
        if (result < 0 && rounds_done >= max_rounds) {          // cycles: every moving representative stands for up to a window of elements
            *remaining = std::min<uint64_t>(n2, std::max<uint64_t>(ctx->h_ctr[3], 1) * 2 * K6L_NV);
            result = 0;
        }
    }
    dfree(ctx, d_list[0]); dfree(ctx, d_list[1]);
    if (result != 0) return ctx->fail(CDBG_ERR_INTERNAL, "list ranking did not converge");
    LAUNCH(k6_resolve, grid_for(ctx, n2, 256, 8), 256, 0, d_pair, n2);
    return CDBG_OK;
}

// smallest edge-buffer size (CDBG_EDGE_MIN: tests shrink it to exercise the growth paths)
uint64_t edge_min() { static const uint64_t v = getenv("CDBG_EDGE_MIN") ? strtoull(getenv("CDBG_EDGE_MIN"), nullptr, 10) : (1ULL << 16); return std::max<uint64_t>(v, 16); }

// List ranking over `n_nodes` nodes with two ends each (link[2*n_nodes], NOLINK = chain end).  weight == nullptr: every node
// counts 1 (k-mers); otherwise hop counts are weighted (pieces of several k-mers).  Cycles are cut at their smallest node;
// the two edges of each cut junction are appended to edges[] (packed end ids) and *n_edges advanced.
int run_ranking(cdbg_ctx *ctx, uint32_t *d_link, uint64_t n_nodes, const uint32_t *d_weight, uint64_t **d_edges_io, uint64_t *edge_cap_io,
                uint64_t *n_edges, uint64_t **d_pair_out) {
    uint64_t *d_edges = *d_edges_io;
    uint64_t edge_cap = *edge_cap_io;
    int rc;
    const uint64_t n2 = 2 * n_nodes;
    uint64_t *d_pair = nullptr;
    if ((rc = dalloc(ctx, &d_pair, n2 + 1))) return rc;
    *d_pair_out = d_pair;
    if (!n_nodes) return CDBG_OK;
    uint64_t remaining = 0;
    const char *k6m = getenv("CDBG_K6");
    const bool two_level = !d_weight && n2 < (1ULL << 32) && (K6_TWO_LEVEL_DEFAULT ? !(k6m && (strcmp(k6m, "full") == 0 || strcmp(k6m, "list") == 0)) : (k6m && strcmp(k6m, "local") == 0));
    if (two_level) {
        if ((rc = rank_two_level(ctx, d_link, d_pair, n2, &remaining))) return rc;
    } else {
        if (d_weight) LAUNCH(k6w_init, grid_for(ctx, n2, 256, 8), 256, 0, d_link, d_weight, d_pair, n2);
        else LAUNCH(k6_init, grid_for(ctx, n2, 256, 8), 256, 0, d_link, d_pair, n2);
        if ((rc = rank_rounds(ctx, d_pair, n2, &remaining, d_weight != nullptr))) return rc;
    }
    for (int guard = 0; remaining && guard < 4; guard++) {
        uint8_t *d_cyc = nullptr; uint64_t *d_lab = nullptr;
        if ((rc = dalloc(ctx, &d_cyc, n2 + 1))) return rc;
        if ((rc = dalloc(ctx, &d_lab, n2 + 1))) return rc;
        LAUNCH(k6_cycle_mark, grid_for(ctx, n2, 256, 8), 256, 0, d_pair, d_link, n2, d_cyc, d_lab);
        int rounds = 2;
        for (uint64_t r = 1; r < remaining; r <<= 1) rounds++;
        for (int r = 0; r < rounds; r++) LAUNCH(k6_cycle_min, grid_for(ctx, n2, 256, 8), 256, 0, d_cyc, d_lab, n2);
        // room for the two edges of every cut: count the cycles first and grow the edge buffer if need be (inputs made of
        // many small circular components used to end in "edge buffer overflow")
        if ((rc = ctr_zero(ctx))) return rc;
        LAUNCH(k6_cycle_count, grid_for(ctx, n2, 256, 8), 256, 0, d_cyc, d_lab, n2, ctx->d_ctr + 1);
        if ((rc = ctr_read(ctx))) return rc;
        if (*n_edges + 2 * ctx->h_ctr[1] > edge_cap) {
            const uint64_t new_cap = *n_edges + 2 * ctx->h_ctr[1] + 1024;
            uint64_t *ne = nullptr;
            if ((rc = dalloc(ctx, &ne, new_cap))) return rc;
            if (*n_edges) CK(cudaMemcpyAsync(ne, d_edges, *n_edges * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            dfree(ctx, d_edges);
            d_edges = ne; edge_cap = new_cap;
            *d_edges_io = ne; *edge_cap_io = new_cap;
        }
        if ((rc = ctr_zero(ctx))) return rc;
        CK(cudaMemcpyAsync(ctx->d_ctr + 0, n_edges, 8, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(k6_cycle_cut, grid_for(ctx, n2, 256, 8), 256, 0, d_cyc, d_lab, n2, d_link, d_edges, edge_cap, ctx->d_ctr + 0, ctx->d_ctr + 1);
        if (d_weight) LAUNCH(k6w_reinit_flagged, grid_for(ctx, n2, 256, 8), 256, 0, d_cyc, d_link, d_weight, d_pair, n2);
        else LAUNCH(k6_reinit_flagged, grid_for(ctx, n2, 256, 8), 256, 0, d_cyc, d_link, d_pair, n2);
        if ((rc = ctr_read(ctx))) return rc;
        if (ctx->h_ctr[0] > edge_cap) return ctx->fail(CDBG_ERR_INTERNAL, "edge buffer overflow while cutting cycles");
        *n_edges = ctx->h_ctr[0];
        ctx->st.n_cycles += ctx->h_ctr[1];
        dfree(ctx, d_cyc); dfree(ctx, d_lab);
        if ((rc = rank_rounds(ctx, d_pair, n2, &remaining))) return rc;
    }
    if (remaining) return ctx->fail(CDBG_ERR_INTERNAL, "cycles remain after cutting");
    return CDBG_OK;
}

// Unit (unitig / piece) numbering from the ranked pairs: info per node, exclusive scans of head flags and head lengths.
int number_units(cdbg_ctx *ctx, const uint64_t *d_pair, uint64_t n_nodes, const uint32_t *d_weight, uint64_t **d_info, uint64_t **d_uid_scan,
                 uint64_t **d_nt_scan, uint64_t *U, uint64_t *total_nt) {
    int rc;
    uint32_t *d_head_flag = nullptr, *d_head_nt = nullptr;
    if ((rc = dalloc(ctx, d_info, n_nodes + 1))) return rc;
    if ((rc = dalloc(ctx, &d_head_flag, n_nodes + 1))) return rc;
    if ((rc = dalloc(ctx, &d_head_nt, n_nodes + 1))) return rc;
    if ((rc = dalloc(ctx, d_uid_scan, n_nodes + 1))) return rc;
    if ((rc = dalloc(ctx, d_nt_scan, n_nodes + 1))) return rc;
    *U = 0; *total_nt = 0;
    if (n_nodes) {
        if (d_weight) LAUNCH(k6w_finalize, grid_for(ctx, n_nodes, 256, 8), 256, 0, d_pair, d_weight, n_nodes, ctx->k, *d_info, d_head_flag, d_head_nt);
        else LAUNCH(k6_finalize, grid_for(ctx, n_nodes, 256, 8), 256, 0, d_pair, n_nodes, ctx->k, *d_info, d_head_flag, d_head_nt);
        if ((rc = exclusive_scan(ctx, d_head_flag, n_nodes, *d_uid_scan, U))) return rc;
        if ((rc = exclusive_scan(ctx, d_head_nt, n_nodes, *d_nt_scan, total_nt))) return rc;
    }
    dfree(ctx, d_head_flag); dfree(ctx, d_head_nt);
    return CDBG_OK;
}

// Spell the U units (unitigs, or per-GPU pieces in sharded mode) numbered by number_units into ctx->d_seq / d_seq_off / d_kc.
int spell_units(cdbg_ctx *ctx, const uint64_t *d_uid_scan, const uint64_t *d_nt_scan, bool want_terms) {
    int rc;
    const uint64_t S = ctx->S, U = ctx->U, total_nt = ctx->total_nt;
    if ((rc = dalloc(ctx, &ctx->d_seq, total_nt + 1))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_seq_off, U + 1))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_kc, U + 1))) return rc;
    CK(cudaMemsetAsync(ctx->d_kc, 0, (U + 1) * 8, ctx->stream));
    if (ctx->prm.flags & CDBG_FLAG_ALL_ABUNDANCE_COUNTS) { if ((rc = dalloc(ctx, &ctx->d_abund, S + 1))) return rc; }
    if (want_terms) {
        if ((rc = dalloc(ctx, &ctx->d_term_begin, U + 1))) return rc;
        if ((rc = dalloc(ctx, &ctx->d_term_end, U + 1))) return rc;
    }
    if (S) {
        SpellParams sp{};
        sp.solid_kmers = ctx->d_solid_kmers; sp.solid_counts = ctx->d_solid_counts; sp.S = S; sp.k = ctx->k;
        sp.info = ctx->d_info; sp.uid_scan = d_uid_scan; sp.nt_scan = d_nt_scan;
        sp.seq = ctx->d_seq; sp.seq_off = ctx->d_seq_off; sp.kc = ctx->d_kc; sp.abund = ctx->d_abund;
        sp.unitig_of_kmer = ctx->d_unitig_of_kmer;
        sp.link = ctx->d_link; sp.term_begin = want_terms ? ctx->d_term_begin : nullptr; sp.term_end = want_terms ? ctx->d_term_end : nullptr;
        switch (ctx->W) {      // (multi-word k-mers reach this function only on the sharded path: its pieces are spelled here)
            case 1: LAUNCH(k8_spell, grid_for(ctx, S, 256, 8), 256, 0, sp); break;
            case 2: LAUNCH(kw8_spell<2>, grid_for(ctx, S, 256, 8), 256, 0, sp); break;
            case 3: LAUNCH(kw8_spell<3>, grid_for(ctx, S, 256, 8), 256, 0, sp); break;
            default: LAUNCH(kw8_spell<4>, grid_for(ctx, S, 256, 8), 256, 0, sp); break;
        }
    }
    CK(cudaMemcpyAsync(ctx->d_seq_off + U, &ctx->total_nt, 8, cudaMemcpyHostToDevice, ctx->stream));
    return CDBG_OK;
}

#ifndef K5_SMEM_DEFAULT
#define K5_SMEM_DEFAULT 1   // shared-memory join tables by default (C2 on a B200: K5 0.91 -> 0.455 ms); CDBG_K5=l2 selects the L2 tables
#endif

// K5 with shared-memory tables: fine histogram -> scan -> fine scatter -> one CTA per fine bin (k5_smem.cuh).  The end records are
// in P coarse bins (h_cnt, h_off).  Leaves ctx->d_link (pre-allocated) / d_edges / n_edges like the L2 path.
int stage_compact_k5_smem(cdbg_ctx *ctx, uint32_t slots, double per_fine, const uint64_t *d_rec_key, const uint64_t *d_rec_val, uint32_t P,
                          const std::vector<uint64_t> &h_cnt, const std::vector<uint64_t> &h_off) {
    int rc;
    const uint64_t S = ctx->S, n2 = 2 * S;
    const uint32_t F = (uint32_t)std::min<uint64_t>(K5P_MAX_F, std::max<uint64_t>(1, (uint64_t)ceil((double)n2 / (per_fine * P))));
    const uint64_t NF = (uint64_t)P * F;
    std::vector<uint32_t> h_tile_first(P + 1);
    std::vector<uint64_t> h_range(2 * (size_t)P);
    uint64_t tiles = 0;
    for (uint32_t b = 0; b < P; b++) {
        h_tile_first[b] = (uint32_t)tiles; tiles += (h_cnt[b] + K5P_TILE - 1) / K5P_TILE;
        h_range[b] = h_off[b]; h_range[P + b] = h_off[b] + h_cnt[b];
    }
    h_tile_first[P] = (uint32_t)tiles;
    if (tiles >= (1ULL << 32) || NF >= (1ULL << 32)) return ctx->fail(CDBG_ERR_PARAM, "too many end-record tiles for one device");

    uint32_t *d_tile_first = nullptr, *d_fine_cnt = nullptr, *d_fine_cursor = nullptr;
    uint64_t *d_fine_off = nullptr, *d_key2 = nullptr, *d_val2 = nullptr, *d_bin_range = nullptr;
    if ((rc = dalloc(ctx, &d_bin_range, 2 * (size_t)P))) return rc;
    if ((rc = dalloc(ctx, &d_tile_first, P + 1))) return rc;
    if ((rc = dalloc(ctx, &d_fine_cnt, NF + 1))) return rc;
    if ((rc = dalloc(ctx, &d_fine_cursor, NF + 1))) return rc;
    if ((rc = dalloc(ctx, &d_fine_off, NF + 2))) return rc;
    if ((rc = dalloc(ctx, &d_key2, n2 + 1))) return rc;
    if ((rc = dalloc(ctx, &d_val2, n2 + 1))) return rc;
    CK(cudaMemsetAsync(d_fine_cnt, 0, (NF + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(d_fine_cursor, 0, (NF + 1) * 4, ctx->stream));
    CK(cudaMemcpyAsync(d_bin_range, h_range.data(), 2 * (size_t)P * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_tile_first, h_tile_first.data(), (P + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));

    tstart(ctx, T_K5);
    K5PParams pp{};
    pp.key = d_rec_key; pp.val = d_rec_val; pp.bin_beg = d_bin_range; pp.bin_end = d_bin_range + P; pp.tile_first = d_tile_first;
    pp.P1 = P; pp.F = F; pp.n_tiles = (uint32_t)tiles; pp.group = nullptr; pp.G = P;
    pp.fine_cnt = d_fine_cnt; pp.fine_off = d_fine_off; pp.fine_cursor = d_fine_cursor; pp.out_key = d_key2; pp.out_val = d_val2;
    if (tiles) {
        LAUNCH(k5p_hist, (int)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * 8), K5P_THREADS, 0, pp);
        uint64_t total = 0;
        if ((rc = exclusive_scan(ctx, d_fine_cnt, NF + 1, d_fine_off, &total))) return rc;
        if (total != n2) return ctx->fail(CDBG_ERR_INTERNAL, "fine histogram lost end records (%llu != %llu)", (unsigned long long)total, (unsigned long long)n2);
        CK(cudaFuncSetAttribute(k5p_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5p_scatter_smem_bytes()));
        LAUNCH(k5p_scatter, (int)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * 2), K5P_THREADS, k5p_scatter_smem_bytes(), pp);
    } else {
        CK(cudaMemsetAsync(d_fine_off, 0, (NF + 2) * 8, ctx->stream));
    }

    ctx->edge_cap = std::max<uint64_t>(S / 2, edge_min());
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = dalloc(ctx, &ctx->d_edges, ctx->edge_cap))) return rc;
        CK(cudaMemsetAsync(ctx->d_link, 0xFF, (n2 + 1) * 4, ctx->stream));
        if ((rc = ctr_zero(ctx))) return rc;
        K5CParams cp{};
        cp.key = d_key2; cp.val = d_val2; cp.fine_off = d_fine_off; cp.n_fine = (uint32_t)NF;
        cp.link = ctx->d_link; cp.edges = ctx->d_edges; cp.edge_cap = ctx->edge_cap;
        cp.n_edges = ctx->d_ctr + 0; cp.n_compact = ctx->d_ctr + 1; cp.n_splits = ctx->d_ctr + 2; cp.failed = ctx->d_flag + 1;
        if (n2) {
            if (slots == 4096) {
                CK(cudaFuncSetAttribute(k5c_join<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5c_smem_bytes<4096>()));
                LAUNCH(k5c_join<4096>, (int)std::min<uint64_t>(NF, (uint64_t)ctx->n_sm), K5C_THREADS, k5c_smem_bytes<4096>(), cp);
            } else {
                CK(cudaFuncSetAttribute(k5c_join<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5c_smem_bytes<2048>()));
                LAUNCH(k5c_join<2048>, (int)std::min<uint64_t>(NF, (uint64_t)ctx->n_sm * 2), K5C_THREADS, k5c_smem_bytes<2048>(), cp);
            }
        }
        if (attempt == 0) tstop(ctx, T_K5);
        if ((rc = ctr_read(ctx))) return rc;
        if (((int *)(ctx->h_ctr + 64))[1]) return ctx->fail(CDBG_ERR_INTERNAL, "K5: a fine bin does not fit a shared-memory table after %d splits", K5C_MAX_LEVEL);
        ctx->n_edges = ctx->h_ctr[0];
        ctx->st.n_compactions = ctx->h_ctr[1];
        ctx->st.n_k5_splits = ctx->h_ctr[2];
        if (ctx->n_edges <= ctx->edge_cap) break;          // (the edges of cut cycles make their own room: run_ranking grows the buffer)
        if (attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "edge buffer overflow");
        dfree(ctx, ctx->d_edges);
        ctx->edge_cap = ctx->n_edges + std::max<uint64_t>(ctx->n_edges / 8, 1 << 16);
    }
    dfree(ctx, d_bin_range); dfree(ctx, d_tile_first); dfree(ctx, d_fine_cnt); dfree(ctx, d_fine_cursor); dfree(ctx, d_fine_off);
    dfree(ctx, d_key2); dfree(ctx, d_val2);
    return CDBG_OK;
}

int stage_compact(cdbg_ctx *ctx) {
    const int k = ctx->k;
    const uint64_t S = ctx->S, n2 = 2 * S;
    int rc;
    tstart(ctx, T_COMPACT);

    // ---- K3: end records, partitioned by hash of the canonical (k-1)-mer
    const uint64_t join_target = getenv("CDBG_JOIN_TARGET") ? strtoull(getenv("CDBG_JOIN_TARGET"), nullptr, 10) : 500000ULL;   // records per bin
    uint32_t P = (uint32_t)std::min<uint64_t>(4096, std::max<uint64_t>(1, (n2 + join_target - 1) / join_target));
    // K5 variant (CDBG_K5): "l2" = one L2-resident join table per coarse bin; "smem" = fine partition + shared-memory tables
    // (k5_smem.cuh): the coarse and the fine fan-out split the number of fine bins evenly, as for K2.
    const char *k5mode = getenv("CDBG_K5");
    const bool k5_smem = K5_SMEM_DEFAULT ? !(k5mode && strcmp(k5mode, "l2") == 0) : (k5mode && strcmp(k5mode, "smem") == 0);
    const uint32_t k5_slots = getenv("CDBG_K5S_SLOTS") && atoi(getenv("CDBG_K5S_SLOTS")) == 4096 ? 4096u : 2048u;
    double k5_fill = getenv("CDBG_K5S_FILL") ? atof(getenv("CDBG_K5S_FILL")) : 1.0;   // end records per table slot (2 records per key are typical)
    if (!(k5_fill >= 0.05 && k5_fill <= 8.0)) k5_fill = 1.0;
    const double k5_per_fine = k5_fill * k5_slots;
    if (k5_smem) {
        uint64_t p1 = (uint64_t)ceil(sqrt(std::max(1.0, (double)n2 / k5_per_fine)));
        if (getenv("CDBG_K5S_P1")) p1 = strtoull(getenv("CDBG_K5S_P1"), nullptr, 10);
        P = (uint32_t)std::min<uint64_t>(4096, std::max<uint64_t>(1, p1));
    }
    ctx->st.n_join_bins = P;
    ctx->st.n_junction_records = n2;
    unsigned long long *d_bin_count = nullptr, *d_bin_cursor = nullptr;
    uint64_t *d_bin_off = nullptr, *d_rec_key = nullptr, *d_rec_val = nullptr;
    if ((rc = dalloc(ctx, &d_bin_count, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_cursor, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_off, P + 1))) return rc;
    if ((rc = dalloc(ctx, &d_rec_key, n2 + 1))) return rc;
    if ((rc = dalloc(ctx, &d_rec_val, n2 + 1))) return rc;
    CK(cudaMemsetAsync(d_bin_count, 0, P * 8, ctx->stream));
    CK(cudaMemsetAsync(d_bin_cursor, 0, P * 8, ctx->stream));
    K3Params p3{};
    p3.solid_kmers = ctx->d_solid_kmers; p3.n_solid = S; p3.k = k; p3.P = P;
    p3.bin_count = d_bin_count; p3.bin_off = d_bin_off; p3.bin_cursor = d_bin_cursor; p3.rec_key = d_rec_key; p3.rec_val = d_rec_val;
    std::vector<uint64_t> h_cnt(P, 0), h_off(P + 1, 0);
    uint64_t max_bin = 0;
    tstart(ctx, T_K3);
    if (S) {
        CK(cudaFuncSetAttribute(k3_ends<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3_smem_bytes(1, P)));
        const int g3 = (int)std::min<uint64_t>((S + K3_TILE - 1) / K3_TILE, (uint64_t)ctx->n_sm * 4);
        LAUNCH(k3_ends<0>, g3, K3_THREADS, k3_smem_bytes(0, P), p3);
        CK(cudaMemcpyAsync(h_cnt.data(), d_bin_count, P * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SYNC_STREAM(ctx);
        uint64_t acc = 0;
        for (uint32_t b = 0; b < P; b++) { h_off[b] = acc; acc += h_cnt[b]; max_bin = std::max(max_bin, h_cnt[b]); }
        h_off[P] = acc;
        CK(cudaMemcpyAsync(d_bin_off, h_off.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(k3_ends<1>, g3, K3_THREADS, k3_smem_bytes(1, P), p3);
    }
    tstop(ctx, T_K3);

    // ---- K5: per-bin join
    if ((rc = dalloc(ctx, &ctx->d_link, n2 + 1))) return rc;
    if (k5_smem && max_bin < (1ULL << 32)) {
        if ((rc = stage_compact_k5_smem(ctx, k5_slots, k5_per_fine, d_rec_key, d_rec_val, P, h_cnt, h_off))) return rc;
        dfree(ctx, d_rec_key); dfree(ctx, d_rec_val);
        dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
    } else {
    auto table_cap = [](uint64_t n) { return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1024, (n * 3 / 2 + 63) & ~63ULL), 0xFFFFFF00u); };
    uint32_t cap = table_cap(max_bin);
    JoinTable T{};
    if ((rc = dalloc(ctx, &T.keys, cap))) return rc;
    if ((rc = dalloc(ctx, &T.cnt, cap))) return rc;
    if ((rc = dalloc(ctx, &T.ids, (size_t)cap * 8))) return rc;
    CK(cudaMemsetAsync(T.keys, 0xFF, (size_t)cap * 8, ctx->stream));
    CK(cudaMemsetAsync(T.cnt, 0, (size_t)cap * 4, ctx->stream));
    ctx->edge_cap = std::max<uint64_t>(S / 2, 1 << 16);
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = dalloc(ctx, &ctx->d_edges, ctx->edge_cap + 2 * 0))) return rc;
        CK(cudaMemsetAsync(ctx->d_link, 0xFF, (n2 + 1) * 4, ctx->stream));
        if ((rc = ctr_zero(ctx))) return rc;
        tstart(ctx, T_K5);
        for (uint32_t b = 0; b < P; b++) {
            uint64_t nb = h_cnt[b];
            if (!nb) continue;
            uint32_t cap_b = table_cap(nb);
            T.cap = cap_b;
            LAUNCH_PDL(k5_insert, grid_for(ctx, nb, 256, 8), 256, 0, d_rec_key + h_off[b], d_rec_val + h_off[b], nb, T, ctx->d_flag);
            LAUNCH_PDL(k5_resolve, grid_for(ctx, cap_b, 256, 8), 256, 0, T, cap_b, ctx->d_link, ctx->d_edges, ctx->edge_cap, ctx->d_ctr + 0,
                   ctx->d_ctr + 1);
        }
        tstop(ctx, T_K5);
        if ((rc = ctr_read(ctx))) return rc;
        if (((int *)(ctx->h_ctr + 64))[0]) return ctx->fail(CDBG_ERR_INTERNAL, "K5 hash table overflow");
        ctx->n_edges = ctx->h_ctr[0];
        ctx->st.n_compactions = ctx->h_ctr[1];
        // leave room for the two edges of every cut cycle (at most S/2 cycles, in practice a handful)
        if (ctx->n_edges + 2048 <= ctx->edge_cap) break;
        if (attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "edge buffer overflow");
        dfree(ctx, ctx->d_edges);
        ctx->edge_cap = ctx->n_edges + std::max<uint64_t>(ctx->n_edges / 8, 1 << 16);
    }
    dfree(ctx, d_rec_key); dfree(ctx, d_rec_val); dfree(ctx, T.keys); dfree(ctx, T.cnt); dfree(ctx, T.ids);
    dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
    }

    // ---- K6: list ranking + unitig numbering
    uint64_t *d_pair = nullptr, *d_uid_scan = nullptr, *d_nt_scan = nullptr;
    tstart(ctx, T_K6R);
    if ((rc = run_ranking(ctx, ctx->d_link, S, nullptr, &ctx->d_edges, &ctx->edge_cap, &ctx->n_edges, &d_pair))) return rc;
    tstop(ctx, T_K6R);
    tstart(ctx, T_K6S);
    uint64_t U = 0, total_nt = 0;
    if ((rc = number_units(ctx, d_pair, S, nullptr, &ctx->d_info, &d_uid_scan, &d_nt_scan, &U, &total_nt))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_unitig_of_kmer, S + 1))) return rc;
    ctx->U = U; ctx->total_nt = total_nt;
    dfree(ctx, d_pair);
    if ((rc = spell_units(ctx, d_uid_scan, d_nt_scan, false))) return rc;
    dfree(ctx, d_uid_scan); dfree(ctx, d_nt_scan);
    tstop(ctx, T_K6S);
    ctx->st.n_unitigs = U; ctx->st.total_nt = total_nt;
    tstop(ctx, T_COMPACT);
    ctx->stage = ST_COMPACTED;
    return CDBG_OK;
}

// ------------------------------------------------------------------------------------------------ stage 4

int stage_link(cdbg_ctx *ctx) {
    int rc;
    const uint64_t n = ctx->n_edges, U = ctx->U;
    tstart(ctx, T_LINK);
    LinkRec *d_tmp = nullptr; uint32_t *d_per = nullptr, *d_cur = nullptr;
    if ((rc = dalloc(ctx, &ctx->d_links, n + 1))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_link_off, U + 1))) return rc;
    if ((rc = dalloc(ctx, &d_tmp, n + 1))) return rc;
    if ((rc = dalloc(ctx, &d_per, U + 1))) return rc;
    if ((rc = dalloc(ctx, &d_cur, U + 1))) return rc;
    CK(cudaMemsetAsync(d_per, 0, (U + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(d_cur, 0, (U + 1) * 4, ctx->stream));
    if (n) LAUNCH(k9_translate, grid_for(ctx, n, 256, 8), 256, 0, ctx->d_edges, n, ctx->d_info, ctx->d_unitig_of_kmer, d_tmp, d_per);
    uint64_t total = 0;
    if ((rc = exclusive_scan(ctx, d_per, U + 1, ctx->d_link_off, &total))) return rc;   // entry U = n
    if (total != n) return ctx->fail(CDBG_ERR_INTERNAL, "link grouping lost records (%llu != %llu)", (unsigned long long)total, (unsigned long long)n);
    if (n) LAUNCH(k9_group, grid_for(ctx, n, 256, 8), 256, 0, d_tmp, n, ctx->d_link_off, d_cur, ctx->d_links);
    dfree(ctx, d_tmp); dfree(ctx, d_per); dfree(ctx, d_cur);
    tstop(ctx, T_LINK);
    ctx->st.n_links = n;
    ctx->stage = ST_LINKED;
    return CDBG_OK;
}

#include "sk_host.inl"
#include "shard_host.inl"
#include "wide_host.inl"
#include "shard_wide_host.inl"

void fill_stats(cdbg_ctx *ctx, cdbg_stats *out) {
    cudaStreamSynchronize(ctx->stream);
    cdbg_stats &s = ctx->st;
    s.n_bytes = ctx->n_bytes;
    s.kernel_launches = ctx->launches;
    s.ms_push = tread(ctx, T_PUSH); s.ms_count = tread(ctx, T_COUNT); s.ms_compact = tread(ctx, T_COMPACT);
    s.ms_link = tread(ctx, T_LINK); s.ms_fetch = tread(ctx, T_FETCH);
    s.ms_k1_hist = tread(ctx, T_K1H); s.ms_k1_scatter = tread(ctx, T_K1S); s.ms_k2 = tread(ctx, T_K2);
    s.ms_k3 = tread(ctx, T_K3); s.ms_k5 = tread(ctx, T_K5); s.ms_k6_rank = tread(ctx, T_K6R); s.ms_k6_spell = tread(ctx, T_K6S);
    s.ms_k2_hist = tread(ctx, T_K2H); s.ms_k2_scatter = tread(ctx, T_K2S); s.ms_k2_table = tread(ctx, T_K2C);
    s.ms_x1 = tread(ctx, T_X1); s.ms_x2 = tread(ctx, T_X2); s.ms_x3 = tread(ctx, T_X3); s.ms_glue = tread(ctx, T_GLUE); s.ms_xlink = tread(ctx, T_XLINK);
    if (out) *out = s;
}

}  // namespace

// ================================================================================================ C ABI

extern "C" {

const char *cdbg_version(void) { return "cdbg_b200 0.2 (sm_100a, k<=127)"; }

int cdbg_create(cdbg_ctx **out, const cdbg_params *params) {
    if (!out || !params) return CDBG_ERR_PARAM;
    *out = nullptr;
    cdbg_ctx *ctx = new cdbg_ctx();
    ctx->prm = *params;
    auto bail = [&](int code) { *out = ctx; return code; };   // caller can still read cdbg_last_error, then destroy
    // span dispatch of the reference (Integer.hpp:77-85: spans 32/64/96/128): one 64-bit word for k <= 31, 2..4 words up to k = 127
    if (params->kmer_size < 4 || params->kmer_size > 127)
        return bail(ctx->fail(CDBG_ERR_PARAM, "unhandled kmer size %d (this build supports 4 <= k <= 127)", params->kmer_size));
    ctx->W = params->kmer_size <= 31 ? 1 : std::max(2, (2 * params->kmer_size + 63) / 64);   // k = 32 also takes two words (reference: k < span)
    if (params->abundance_min < 0) return bail(ctx->fail(CDBG_ERR_PARAM, "abundance-min must be >= 0"));
    if (params->world_size > MAX_WORLD || params->world_size < 0 || (params->world_size > 1 && (params->rank < 0 || params->rank >= params->world_size)))
        return bail(ctx->fail(CDBG_ERR_PARAM, "bad rank/world_size %d/%d (at most %d shards)", params->rank, params->world_size, MAX_WORLD));
    ctx->world = params->world_size > 1 ? params->world_size : 1;
    ctx->rank = ctx->world > 1 ? params->rank : 0;
    ctx->k = params->kmer_size;
    ctx->m = params->minimizer_size > 0 ? params->minimizer_size : 10;
    if (ctx->m > ctx->k - 1) ctx->m = ctx->k - 1;   // ConfigurationAlgorithm.cpp:251
    ctx->trace = getenv("CDBG_TRACE") != nullptr;
    ctx->use_pdl = !(getenv("CDBG_PDL") && atoi(getenv("CDBG_PDL")) == 0);
    ctx->amin = (uint32_t)std::max(params->abundance_min, 1);
    ctx->amax = params->abundance_max > 0 ? (uint32_t)params->abundance_max : 0x7FFFFFFFu;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return bail(ctx->fail(CDBG_ERR_CUDA, "no CUDA device available (%s): this library has no CPU fallback",
                              e != cudaSuccess ? cudaGetErrorString(e) : "device count 0"));
    }
    if (params->device < 0 || params->device >= ndev) return bail(ctx->fail(CDBG_ERR_PARAM, "bad device ordinal %d", params->device));
    *out = ctx;
    CK(cudaSetDevice(params->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, params->device));
    ctx->n_sm = prop.multiProcessorCount;
    ctx->total_mem = (uint64_t)prop.totalGlobalMem;
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    if (ctx->world > 1) {
        CK(cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    CK(cudaMalloc(&ctx->d_ctr, 64 * sizeof(unsigned long long)));
    CK(cudaMalloc(&ctx->d_flag, 4 * sizeof(int)));
    CK(cudaHostAlloc(&ctx->h_ctr, 80 * sizeof(unsigned long long), cudaHostAllocDefault));
    return CDBG_OK;
}

void cdbg_destroy(cdbg_ctx *ctx) {
    if (!ctx) return;
    if (ctx->stream) {
        cudaSetDevice(ctx->prm.device);
        free_results(ctx);
        cudaStreamSynchronize(ctx->stream);
        if (ctx->comm) { if (ctx->shard_failed && g_nccl.CommAbort) g_nccl.CommAbort(ctx->comm); else ncclCommDestroy(ctx->comm); }
        for (auto &b : ctx->blocks) cudaFree(b.p);
        ctx->blocks.clear();
        if (ctx->d_synth) cudaFree(ctx->d_synth);
        if (ctx->d_ctr) cudaFree(ctx->d_ctr);
        if (ctx->d_flag) cudaFree(ctx->d_flag);
        if (ctx->h_ctr) cudaFreeHost(ctx->h_ctr);
        if (ctx->h_seq) cudaFreeHost(ctx->h_seq);
        if (ctx->h_seq_off) cudaFreeHost(ctx->h_seq_off);
        if (ctx->h_kc) cudaFreeHost(ctx->h_kc);
        if (ctx->h_abund) cudaFreeHost(ctx->h_abund);
        if (ctx->h_links) cudaFreeHost(ctx->h_links);
        if (ctx->h_link_off) cudaFreeHost(ctx->h_link_off);
        for (auto &t : ctx->tm) { if (t.a) cudaEventDestroy(t.a); if (t.b) cudaEventDestroy(t.b); }
        if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
        if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
        if (ctx->stream2) cudaStreamDestroy(ctx->stream2);
        cudaStreamDestroy(ctx->stream);
    }
    delete ctx;
}

const char *cdbg_last_error(const cdbg_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int cdbg_reset(cdbg_ctx *ctx) {
    if (!ctx || !ctx->stream) return CDBG_ERR_PARAM;
    CK(cudaSetDevice(ctx->prm.device));
    return free_results(ctx);
}

int cdbg_push_reads(cdbg_ctx *ctx, const uint8_t *bytes, uint64_t n_bytes, int location) {
    if (!ctx || !ctx->stream) return CDBG_ERR_PARAM;
    if (!bytes && n_bytes) return ctx->fail(CDBG_ERR_PARAM, "null reads pointer");
    if (ctx->stage > ST_PUSHED) return ctx->fail(CDBG_ERR_STATE, "push_reads after count: call cdbg_reset first");
    CK(cudaSetDevice(ctx->prm.device));
    int rc;
    tstart(ctx, T_PUSH);
    if (location == CDBG_MEM_DEVICE && ctx->stage == ST_EMPTY && ctx->cap_bytes == 0 && (reinterpret_cast<uintptr_t>(bytes) & 15) == 0) {
        ctx->d_bytes = const_cast<uint8_t *>(bytes);   // used in place
        ctx->n_bytes = n_bytes; ctx->bytes_borrowed = true;
    } else {
        // append (a break byte separates pushes)
        uint64_t need = ctx->n_bytes + n_bytes + 1 + (ctx->bytes_borrowed ? 1 : 0);
        if (ctx->bytes_borrowed || need > ctx->cap_bytes) {
            uint64_t cap = std::max<uint64_t>(need, ctx->cap_bytes * 2);
            uint8_t *nb = nullptr;
            if ((rc = dalloc(ctx, &nb, cap + 64))) return rc;
            if (ctx->n_bytes) CK(cudaMemcpyAsync(nb, ctx->d_bytes, ctx->n_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
            if (ctx->bytes_borrowed) {
                // a buffer used in place carries no break byte after its last read: add it now that something follows
                CK(cudaMemsetAsync(nb + ctx->n_bytes, '\n', 1, ctx->stream));
                ctx->n_bytes += 1;
            } else dfree(ctx, ctx->d_bytes);
            ctx->d_bytes = nb; ctx->cap_bytes = cap; ctx->bytes_borrowed = false;
        }
        CK(cudaMemcpyAsync(ctx->d_bytes + ctx->n_bytes, bytes, n_bytes,
                           location == CDBG_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->d_bytes + ctx->n_bytes + n_bytes, '\n', 1, ctx->stream));
        ctx->n_bytes += n_bytes + 1;
        // the caller may refill or free `bytes` as soon as this returns (include/cdbg.h: "copied inside the call"): a copy from
        // pinned or device memory is otherwise still in flight
        SYNC_STREAM(ctx);
    }
    tstop(ctx, T_PUSH);
    ctx->stage = ST_PUSHED;
    return CDBG_OK;
}

int cdbg_reserve_reads(cdbg_ctx *ctx, uint64_t n_bytes) {
    if (!ctx || !ctx->stream) return CDBG_ERR_PARAM;
    if (ctx->stage > ST_PUSHED) return ctx->fail(CDBG_ERR_STATE, "reserve after count: call cdbg_reset first");
    CK(cudaSetDevice(ctx->prm.device));
    if (n_bytes + 64 <= ctx->cap_bytes && !ctx->bytes_borrowed) return CDBG_OK;
    int rc;
    const uint64_t cap = std::max<uint64_t>(n_bytes, ctx->n_bytes + 2) + 64;
    uint8_t *nb = nullptr;
    if ((rc = dalloc(ctx, &nb, cap + 64))) return rc;
    if (ctx->n_bytes) CK(cudaMemcpyAsync(nb, ctx->d_bytes, ctx->n_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->bytes_borrowed) {          // (a buffer used in place carries no break byte after its last read)
        if (ctx->n_bytes) { CK(cudaMemsetAsync(nb + ctx->n_bytes, '\n', 1, ctx->stream)); ctx->n_bytes += 1; }
    } else if (ctx->d_bytes) dfree(ctx, ctx->d_bytes);
    ctx->d_bytes = nb; ctx->cap_bytes = cap; ctx->bytes_borrowed = false; ctx->bytes_hostmapped = false;
    SYNC_STREAM(ctx);
    return CDBG_OK;
}

int cdbg_push_sequence_text(cdbg_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int location, uint64_t *n_records) {
    if (!ctx || !ctx->stream) return CDBG_ERR_PARAM;
    if (n_records) *n_records = 0;
    if (!text && n_bytes) return ctx->fail(CDBG_ERR_PARAM, "null text pointer");
    if (ctx->stage > ST_PUSHED) return ctx->fail(CDBG_ERR_STATE, "push after count: call cdbg_reset first");
    if (n_bytes == 0) return CDBG_OK;
    CK(cudaSetDevice(ctx->prm.device));
    const uint64_t n_tiles = (n_bytes + PARSE_TILE - 1) / PARSE_TILE;
    if (n_tiles >= (1ULL << 32)) return ctx->fail(CDBG_ERR_PARAM, "text too large for one push (%llu bytes): push it in pieces cut at record starts", (unsigned long long)n_bytes);
    int rc;
    uint8_t first = 0;
    if (location == CDBG_MEM_DEVICE) {
        CK(cudaMemcpyAsync(&first, text, 1, cudaMemcpyDeviceToHost, ctx->stream));
        SYNC_STREAM(ctx);
    } else first = text[0];
    if (first != '>' && first != '@') return ctx->fail(CDBG_ERR_PARAM, "text does not start with a FASTA ('>') or FASTQ ('@') record");
    const int fmt = first == '>' ? PARSE_FASTA : PARSE_FASTQ;

    uint8_t *d_text = nullptr, *d_fn = nullptr, *d_state = nullptr, *d_out = nullptr;
    uint32_t *d_cnt = nullptr; uint64_t *d_off = nullptr;
    const uint8_t *src = text;
    if (location != CDBG_MEM_DEVICE) {
        if ((rc = dalloc(ctx, &d_text, n_bytes + 64))) return rc;
        CK(cudaMemcpyAsync(d_text, text, n_bytes, cudaMemcpyHostToDevice, ctx->stream));
        src = d_text;
    }
    if ((rc = dalloc(ctx, &d_fn, n_tiles + 1))) return rc;
    if ((rc = dalloc(ctx, &d_state, n_tiles + 2))) return rc;
    if ((rc = dalloc(ctx, &d_cnt, n_tiles + 1))) return rc;
    if ((rc = dalloc(ctx, &d_off, n_tiles + 2))) return rc;
    if ((rc = dalloc(ctx, &d_out, n_bytes + 64))) return rc;
    if ((rc = ctr_zero(ctx))) return rc;
    ParseParams pp{};
    pp.text = src; pp.n = n_bytes; pp.n_tiles = (uint32_t)n_tiles; pp.tile_fn = d_fn; pp.tile_state = d_state; pp.tile_cnt = d_cnt;
    pp.tile_off = d_off; pp.out = d_out; pp.n_records = ctx->d_ctr + 0; pp.bad = ctx->d_flag + 3;
    const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 8);
    uint64_t total = 0;
    if (fmt == PARSE_FASTA) {
        LAUNCH(parse_tile_fn<PARSE_FASTA>, grid, PARSE_THREADS, 0, pp);
        LAUNCH(parse_tile_states, 1, PARSE_THREADS, 0, pp, fmt);
        LAUNCH((parse_emit<PARSE_FASTA, false>), grid, PARSE_THREADS, 0, pp);
        if ((rc = exclusive_scan(ctx, d_cnt, n_tiles, d_off, &total))) return rc;
        LAUNCH((parse_emit<PARSE_FASTA, true>), grid, PARSE_THREADS, 0, pp);
    } else {
        LAUNCH(parse_tile_fn<PARSE_FASTQ>, grid, PARSE_THREADS, 0, pp);
        LAUNCH(parse_tile_states, 1, PARSE_THREADS, 0, pp, fmt);
        LAUNCH((parse_emit<PARSE_FASTQ, false>), grid, PARSE_THREADS, 0, pp);
        if ((rc = exclusive_scan(ctx, d_cnt, n_tiles, d_off, &total))) return rc;
        LAUNCH((parse_emit<PARSE_FASTQ, true>), grid, PARSE_THREADS, 0, pp);
    }
    uint8_t final_state = 0;
    CK(cudaMemcpyAsync(&final_state, d_state + n_tiles, 1, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = ctr_read(ctx))) return rc;
    bool bad = ((int *)(ctx->h_ctr + 64))[3] != 0;
    if (fmt == PARSE_FASTQ && final_state != 3) bad = true;                      // the last record is incomplete
    // the record that is still open at the end of the text gets its '\n'
    if (!bad && fmt == PARSE_FASTA && final_state != 0) {
        CK(cudaMemsetAsync(d_out + total, '\n', 1, ctx->stream));
        total++;
    }
    dfree(ctx, d_text); dfree(ctx, d_fn); dfree(ctx, d_state); dfree(ctx, d_cnt); dfree(ctx, d_off);
    if (bad) { dfree(ctx, d_out); return ctx->fail(CDBG_ERR_PARAM, "FASTQ text is not in strict 4-line records (multi-line FASTQ or stray lines): parse it on the host"); }
    if (n_records) *n_records = ctx->h_ctr[0];
    // hand the nucleotide stream to the normal path; as the first push it is used in place (the block stays allocated until reset)
    const bool in_place = ctx->stage == ST_EMPTY && ctx->cap_bytes == 0;
    rc = cdbg_push_reads(ctx, d_out, total, CDBG_MEM_DEVICE);
    if (!in_place || rc != CDBG_OK) dfree(ctx, d_out);
    return rc;
}

int cdbg_get_pushed_reads(cdbg_ctx *ctx, uint8_t *dst, uint64_t cap, uint64_t *n) {
    if (!ctx || !ctx->stream || !n) return CDBG_ERR_PARAM;
    if (ctx->bytes_hostmapped) return ctx->fail(CDBG_ERR_STATE, "the reads of a zero-copy build stay in the caller's buffer");
    CK(cudaSetDevice(ctx->prm.device));
    *n = ctx->n_bytes;
    const uint64_t c = std::min<uint64_t>(cap, ctx->n_bytes);
    if (c && dst) CK(cudaMemcpyAsync(dst, ctx->d_bytes, c, cudaMemcpyDeviceToHost, ctx->stream));
    SYNC_STREAM(ctx);
    return CDBG_OK;
}

int cdbg_count(cdbg_ctx *ctx, cdbg_stats *stats) {
    if (!ctx || !ctx->stream) return CDBG_ERR_PARAM;
    if (ctx->stage != ST_PUSHED) return ctx->fail(CDBG_ERR_STATE, "cdbg_count needs pushed reads");
    if (ctx->world > 1 && !ctx->comm) return ctx->fail(CDBG_ERR_STATE, "sharded context: call cdbg_shard_init first");
    if (ctx->last_n_bytes && ctx->last_n_bytes != ctx->n_bytes) {   // a different workload: drop cached blocks of the old sizes
        cudaStreamSynchronize(ctx->stream);
        for (size_t i = 0; i < ctx->blocks.size();) {
            if (!ctx->blocks[i].used) { cudaFree(ctx->blocks[i].p); ctx->blocks.erase(ctx->blocks.begin() + i); } else i++;
        }
    }
    ctx->last_n_bytes = ctx->n_bytes;
    CK(cudaSetDevice(ctx->prm.device));
    int rc = ctx->world > 1 ? (ctx->W > 1 ? shard_stage_count_wide_dispatch(ctx) : sk_enabled(ctx) ? shard_stage_count_sk(ctx) : shard_stage_count(ctx))
                            : (ctx->W > 1 ? stage_count_wide_dispatch(ctx) : sk_enabled(ctx) ? stage_count_sk(ctx) : stage_count(ctx));
    if (rc == CDBG_OK && stats) fill_stats(ctx, stats);
    return rc;
}

int cdbg_compact(cdbg_ctx *ctx, cdbg_stats *stats) {
    if (!ctx || !ctx->stream) return CDBG_ERR_PARAM;
    if (ctx->stage != ST_COUNTED) return ctx->fail(CDBG_ERR_STATE, "cdbg_compact needs cdbg_count first");
    CK(cudaSetDevice(ctx->prm.device));
    int rc = ctx->world > 1 ? shard_stage_compact(ctx) : ctx->W > 1 ? stage_compact_wide_dispatch(ctx) : stage_compact(ctx);
    if (rc == CDBG_OK && stats) fill_stats(ctx, stats);
    return rc;
}

int cdbg_link(cdbg_ctx *ctx, cdbg_stats *stats) {
    if (!ctx || !ctx->stream) return CDBG_ERR_PARAM;
    if (ctx->stage != ST_COMPACTED) return ctx->fail(CDBG_ERR_STATE, "cdbg_link needs cdbg_compact first");
    CK(cudaSetDevice(ctx->prm.device));
    int rc = ctx->world > 1 ? shard_stage_link(ctx) : stage_link(ctx);
    if (rc == CDBG_OK && stats) fill_stats(ctx, stats);
    return rc;
}

int cdbg_get_stats(cdbg_ctx *ctx, cdbg_stats *stats) {
    if (!ctx || !stats) return CDBG_ERR_PARAM;
    fill_stats(ctx, stats);
    return CDBG_OK;
}

int cdbg_get_unitigs(cdbg_ctx *ctx, cdbg_unitig_view *view) {
    if (!ctx || !ctx->stream || !view) return CDBG_ERR_PARAM;
    if (ctx->stage < ST_COMPACTED) return ctx->fail(CDBG_ERR_STATE, "cdbg_get_unitigs needs cdbg_compact first");
    CK(cudaSetDevice(ctx->prm.device));
    int rc;
    const uint64_t U = ctx->U, nt = ctx->total_nt;
    const uint64_t S = ctx->world > 1 ? nt - U * (uint64_t)(ctx->k - 1) : ctx->S;     // k-mers of the unitigs in this view
    const uint64_t nl = ctx->stage >= ST_LINKED ? ctx->n_edges : 0;
    if ((rc = hgrow(ctx, &ctx->h_seq, &ctx->h_seq_cap, nt + 1))) return rc;
    if ((rc = hgrow(ctx, &ctx->h_seq_off, &ctx->h_seq_off_cap, U + 1))) return rc;
    if ((rc = hgrow(ctx, &ctx->h_kc, &ctx->h_kc_cap, U + 1))) return rc;
    if ((rc = hgrow(ctx, &ctx->h_links, &ctx->h_links_cap, nl + 1))) return rc;
    if ((rc = hgrow(ctx, &ctx->h_link_off, &ctx->h_link_off_cap, U + 2))) return rc;
    const bool ab = ctx->d_abund != nullptr;
    if (ab) { if ((rc = hgrow(ctx, &ctx->h_abund, &ctx->h_abund_cap, S + 1))) return rc; }
    tstart(ctx, T_FETCH);
    if (nt) CK(cudaMemcpyAsync(ctx->h_seq, ctx->d_seq, nt, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_seq_off, ctx->d_seq_off, (U + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (U) CK(cudaMemcpyAsync(ctx->h_kc, ctx->d_kc, U * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (nl) CK(cudaMemcpyAsync(ctx->h_links, ctx->d_links, nl * sizeof(LinkRec), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->stage >= ST_LINKED) CK(cudaMemcpyAsync(ctx->h_link_off, ctx->d_link_off, (U + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    else memset(ctx->h_link_off, 0, (U + 1) * 8);
    if (ab && S) CK(cudaMemcpyAsync(ctx->h_abund, ctx->d_abund, S * 4, cudaMemcpyDeviceToHost, ctx->stream));
    tstop(ctx, T_FETCH);
    SYNC_STREAM(ctx);
    view->kmer_size = ctx->k;
    view->n_unitigs = U; view->total_nt = nt;
    view->seq = ctx->h_seq; view->seq_off = ctx->h_seq_off; view->kc = ctx->h_kc;
    view->abund = ab ? ctx->h_abund : nullptr;
    view->n_links = nl; view->links = ctx->h_links; view->link_off = ctx->h_link_off;
    view->first_unitig = ctx->world > 1 ? ctx->u_first : 0;
    view->n_unitigs_total = ctx->world > 1 ? ctx->U_total : U;
    return CDBG_OK;
}

// cdbg_build with reads in PINNED host memory: instead of an H2D copy followed by K1, the single-pass K1 reads the host buffer
// in place (its TMA bulk loads go over PCIe), so the transfer and the extraction are one step.  Only inside cdbg_build: the call
// is blocking, so the caller's buffer is guaranteed to stay valid while the kernel reads it.
static bool try_zero_copy(cdbg_ctx *ctx, const uint8_t *bytes, uint64_t n_bytes) {
    const char *z = getenv("CDBG_ZEROCOPY"), *k1 = getenv("CDBG_K1"), *zmin = getenv("CDBG_ZEROCOPY_MIN");
    const uint64_t min_bytes = zmin ? strtoull(zmin, nullptr, 10) : (1u << 20);   // small inputs: a plain copy is as good
    if (ctx->W != 1 || n_bytes < min_bytes || n_bytes == 0 || (reinterpret_cast<uintptr_t>(bytes) & 15)) return false;
    if (ctx->world != 1 && !sk_enabled(ctx)) return false;     // sharded: only the super-k-mer K1 reads its input once, through TMA
    if ((z && atoi(z) == 0) || (k1 && strcmp(k1, "single") != 0)) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, bytes) != cudaSuccess) { cudaGetLastError(); return false; }
    if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
    ctx->d_bytes = static_cast<uint8_t *>(at.devicePointer);
    ctx->h_bytes_src = bytes;
    ctx->n_bytes = n_bytes; ctx->bytes_borrowed = true; ctx->bytes_hostmapped = true;
    ctx->stage = ST_PUSHED;
    return true;
}

int cdbg_build(cdbg_ctx *ctx, const uint8_t *bytes, uint64_t n_bytes, int location, cdbg_unitig_view *view, cdbg_stats *stats) {
    int rc;
    if ((rc = cdbg_reset(ctx))) return rc;
    if (location == CDBG_MEM_HOST && bytes && try_zero_copy(ctx, bytes, n_bytes)) ctx->st.zero_copy = 1;
    else if ((rc = cdbg_push_reads(ctx, bytes, n_bytes, location))) return rc;
    if ((rc = cdbg_count(ctx, nullptr))) return rc;
    if ((rc = cdbg_compact(ctx, nullptr))) return rc;
    if ((rc = cdbg_link(ctx, nullptr))) return rc;
    if (view) { if ((rc = cdbg_get_unitigs(ctx, view))) return rc; }
    if (stats) fill_stats(ctx, stats);
    return CDBG_OK;
}

int cdbg_get_solid_kmers(cdbg_ctx *ctx, uint64_t *kmers, uint32_t *counts, uint64_t cap, uint64_t *n) {
    if (!ctx || !ctx->stream || !n) return CDBG_ERR_PARAM;
    if (ctx->stage < ST_COUNTED) return ctx->fail(CDBG_ERR_STATE, "cdbg_get_solid_kmers needs cdbg_count first");
    CK(cudaSetDevice(ctx->prm.device));
    *n = ctx->S;
    uint64_t c = std::min<uint64_t>(cap, ctx->S);
    if (c && kmers) CK(cudaMemcpyAsync(kmers, ctx->d_solid_kmers, c * 8 * ctx->W, cudaMemcpyDeviceToHost, ctx->stream));
    if (c && counts) CK(cudaMemcpyAsync(counts, ctx->d_solid_counts, c * 4, cudaMemcpyDeviceToHost, ctx->stream));
    SYNC_STREAM(ctx);
    return CDBG_OK;
}

int cdbg_set_solid_kmers(cdbg_ctx *ctx, const uint64_t *kmers, const uint32_t *counts, uint64_t n, int location) {
    if (!ctx || !ctx->stream) return CDBG_ERR_PARAM;
    if ((!kmers || !counts) && n) return ctx->fail(CDBG_ERR_PARAM, "null k-mer / count pointer");
    if (ctx->world != 1) return ctx->fail(CDBG_ERR_PARAM, "cdbg_set_solid_kmers needs a single-GPU context");
    if (ctx->stage != ST_EMPTY) return ctx->fail(CDBG_ERR_STATE, "cdbg_set_solid_kmers on a context that already holds reads or results: call cdbg_reset first");
    if (n >= (1ULL << 31)) return ctx->fail(CDBG_ERR_PARAM, "more than 2^31 solid k-mers on one device (%llu)", (unsigned long long)n);
    CK(cudaSetDevice(ctx->prm.device));
    int rc;
    tstart(ctx, T_COUNT);
    if ((rc = dalloc(ctx, &ctx->d_solid_kmers, n * ctx->W + 1))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_solid_counts, n + 1))) return rc;
    const cudaMemcpyKind kind = location == CDBG_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    if (n) {
        CK(cudaMemcpyAsync(ctx->d_solid_kmers, kmers, n * 8 * ctx->W, kind, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_solid_counts, counts, n * 4, kind, ctx->stream));
        switch (ctx->W) {
            case 1: LAUNCH(k2_canonicalise, grid_for(ctx, n, 256, 8), 256, 0, ctx->d_solid_kmers, n, ctx->k); break;
            case 2: LAUNCH(kw2_canonicalise<2>, grid_for(ctx, n, 256, 8), 256, 0, ctx->d_solid_kmers, n, ctx->k); break;
            case 3: LAUNCH(kw2_canonicalise<3>, grid_for(ctx, n, 256, 8), 256, 0, ctx->d_solid_kmers, n, ctx->k); break;
            default: LAUNCH(kw2_canonicalise<4>, grid_for(ctx, n, 256, 8), 256, 0, ctx->d_solid_kmers, n, ctx->k); break;
        }
    }
    SYNC_STREAM(ctx);   // the caller may free its arrays when this returns
    tstop(ctx, T_COUNT);
    ctx->S = n;
    ctx->st.n_solid = n; ctx->st.n_distinct = n;
    ctx->stage = ST_COUNTED;
    return CDBG_OK;
}

int cdbg_synth_reads(cdbg_ctx *ctx, uint64_t genome_len, uint32_t read_len, uint64_t n_reads, uint32_t err_ppm, uint64_t seed,
                     uint8_t **d_bytes, uint64_t *n_bytes) {
    return cdbg_synth_reads_range(ctx, genome_len, read_len, 0, n_reads, err_ppm, seed, d_bytes, n_bytes);
}

int cdbg_synth_reads_range(cdbg_ctx *ctx, uint64_t genome_len, uint32_t read_len, uint64_t first_read, uint64_t n_reads, uint32_t err_ppm,
                           uint64_t seed, uint8_t **d_bytes, uint64_t *n_bytes) {
    if (!ctx || !ctx->stream || !d_bytes || !n_bytes) return CDBG_ERR_PARAM;
    if (read_len == 0 || genome_len < read_len) return ctx->fail(CDBG_ERR_PARAM, "bad synthetic read geometry");
    CK(cudaSetDevice(ctx->prm.device));
    uint64_t total = n_reads * ((uint64_t)read_len + 1);
    if (total + 64 > ctx->synth_cap) {
        if (ctx->d_synth) CK(cudaFree(ctx->d_synth));
        ctx->d_synth = nullptr; ctx->synth_cap = 0;
        CK(cudaMalloc(&ctx->d_synth, total + 64));
        ctx->synth_cap = total + 64;
    }
    if (total) LAUNCH(synth_reads_kernel, grid_for(ctx, total, 256, 16), 256, 0, ctx->d_synth, genome_len, read_len, first_read, n_reads, err_ppm, seed);
    SYNC_STREAM(ctx);
    *d_bytes = ctx->d_synth; *n_bytes = total;
    return CDBG_OK;
}

int cdbg_nccl_unique_id(uint8_t *out128) {
    if (!out128) return CDBG_ERR_PARAM;
    if (!g_nccl.load()) return CDBG_ERR_NCCL;
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return CDBG_ERR_NCCL;
    static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return CDBG_OK;
}

int cdbg_shard_init(cdbg_ctx *ctx, const uint8_t *id128) {
    if (!ctx || !ctx->stream || !id128) return CDBG_ERR_PARAM;
    if (ctx->world <= 1) return ctx->fail(CDBG_ERR_STATE, "cdbg_shard_init on a context created with world_size <= 1");
    if (ctx->comm) return CDBG_OK;
    g_shard_abort.store(0, std::memory_order_relaxed);          // a new generation of ranks
    ctx->shard_failed = false;
    if (!g_nccl.load()) return ctx->fail(CDBG_ERR_NCCL, "%s", g_nccl.error);
    CK(cudaSetDevice(ctx->prm.device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NK(ncclCommInitRank(&ctx->comm, ctx->world, id, ctx->rank));
    return CDBG_OK;
}

uint32_t cdbg_owner_of(uint64_t x, int32_t n_nt, int32_t m, uint32_t world) {
    if (m > n_nt) m = n_nt;
    return owner_of_rank(min_rank_of(x, n_nt, m), world ? world : 1);
}

int cdbg_plan_shard_exchange(uint32_t world, uint32_t rank, uint32_t p_local, const uint64_t *all_cnt, uint64_t *send_cnt, uint64_t *recv_cnt,
                             uint64_t *run_off, uint64_t *run_len) {
    if (!world || world > MAX_WORLD || rank >= world || !p_local || !all_cnt || !send_cnt || !recv_cnt || !run_off || !run_len) return CDBG_ERR_PARAM;
    plan_shard_exchange(world, rank, p_local, all_cnt, send_cnt, recv_cnt, run_off, run_len);
    return CDBG_OK;
}

void *cdbg_host_alloc(uint64_t n_bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, n_bytes ? n_bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void cdbg_host_free(void *p) { if (p) cudaFreeHost(p); }

int cdbg_copy_to_host(cdbg_ctx *ctx, void *dst, const void *d_src, uint64_t n_bytes) {
    if (!ctx || !ctx->stream) return CDBG_ERR_PARAM;
    CK(cudaSetDevice(ctx->prm.device));
    CK(cudaMemcpyAsync(dst, d_src, n_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SYNC_STREAM(ctx);
    return CDBG_OK;
}

}  // extern "C"
