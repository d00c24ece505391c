// k1_tma.cuh -- K1 (reads -> canonical k-mers, partitioned) as a persistent kernel fed by the TMA bulk-copy engine.
//
// Each CTA walks tiles of 8192 read positions.  One elected thread keeps the NEXT tile's raw bytes in flight with
// cp.async.bulk (global -> shared, completion signalled on an mbarrier with expect_tx) while all threads work on the current
// one, so the loads never sit on the critical path and no LSU instructions are spent on the byte stream.
//   pass 0 (HIST)   : pack ASCII -> 2 bit, roll fwd/rc k-mers, canonical select, 32-bit bin hash; CTA-private histogram in
//                     shared memory, and the bin of every position is written out (uint16, coalesced 64 B per thread)
//   pass 1 (SCATTER): reads the bins back (no hashing, no validity logic), counts the tile per bin, reserves one run per
//                     (tile, bin) with a single global atomic, rolls the k-mers again and counting-sorts them in shared memory so
//                     that every run leaves the SM as contiguous, coalesced stores.
#pragma once
#include "kernels.cuh"

constexpr int K1T_RAW = K1_TILE + 32;       // bytes per tile incl. halo (multiple of 16)
#define K1_NOBIN 0xFFFFu

#ifndef CDBG_EMU_ACTIVE
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared through the TMA engine (SASS: UBLKCP); size and both addresses multiples of 16
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
#else
// CPU emulation build of the tests (tests/emu): an mbarrier (arrival count 1) is a count of completed phases in the low half and
// the bytes still expected by the current phase in the high half; a bulk copy is a memcpy that must deliver exactly the bytes
// announced on its barrier (on a GPU a mismatch is a phase that never completes: a hang) and then completes the phase.
// wait(parity) passes once the phase of that parity has completed.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    if (count != 1) { fprintf(stderr, "[emu] mbarrier arrival counts other than 1 are not modelled\n"); abort(); }
    __atomic_store_n(bar, 0, __ATOMIC_SEQ_CST);
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    const uint64_t cur = __atomic_load_n(bar, __ATOMIC_SEQ_CST);
    if (cur >> 32) { fprintf(stderr, "[emu] expect_tx on an mbarrier whose phase still waits for %u bytes\n", (uint32_t)(cur >> 32)); abort(); }
    if (bytes == 0) __atomic_fetch_add(bar, 1, __ATOMIC_SEQ_CST);
    else __atomic_fetch_add(bar, (uint64_t)bytes << 32, __ATOMIC_SEQ_CST);
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (;;) {
        const uint32_t done = (uint32_t)__atomic_load_n(bar, __ATOMIC_SEQ_CST);
        if (done > 0 && ((done - 1) & 1) == parity) return;
        emu::yield_now();
    }
}
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    // cp.async.bulk: source, destination and size must be multiples of 16 bytes (a GPU raises a misaligned-address fault)
    if ((reinterpret_cast<uintptr_t>(dst_smem) | reinterpret_cast<uintptr_t>(src_gmem) | bytes) & 15) {
        fprintf(stderr, "[emu] bulk copy not 16-byte aligned: dst %p src %p bytes %u\n", dst_smem, src_gmem, bytes);
        abort();
    }
    const uint64_t cur = __atomic_load_n(bar, __ATOMIC_SEQ_CST);
    if ((uint32_t)(cur >> 32) != bytes) { fprintf(stderr, "[emu] bulk copy of %u bytes on an mbarrier expecting %u\n", bytes, (uint32_t)(cur >> 32)); abort(); }
    memcpy(dst_smem, src_gmem, bytes);
    __atomic_fetch_add(bar, 1 - ((uint64_t)bytes << 32), __ATOMIC_SEQ_CST);      // bytes delivered, phase complete
}
__device__ __forceinline__ void mbar_fence_init() {}
#endif

struct K1TParams {
    const uint8_t *bytes; uint64_t n_bytes; uint64_t n_tiles;
    int k; uint32_t P;
    unsigned long long *bin_count; const uint64_t *bin_off; unsigned long long *bin_cursor;
    uint64_t *out_kmers;
    uint16_t *pos_bin;         // [n_tiles * K1_TILE]
};

template <int MODE>
__global__ void __launch_bounds__(K1_THREADS) k1t_extract(K1TParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint8_t *s_raw = smem_raw;                                                              // [2][K1T_RAW]  TMA destinations
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + 2 * K1T_RAW);                 // [2] mbarriers
    uint64_t *s_pk = s_bar + 2;                                                             // [THREADS + 2]
    uint64_t *s_gbase = s_pk + K1_THREADS + 2;                                              // [P]        MODE 1
    uint64_t *s_stage = s_gbase + (MODE == 1 ? p.P : 0);                                    // [K1_TILE]  MODE 1
    uint32_t *s_vm = reinterpret_cast<uint32_t *>(s_stage + (MODE == 1 ? K1_TILE : 0));     // [THREADS + 2]
    uint32_t *s_cnt = s_vm + K1_THREADS + 2;                                                // [P]
    uint32_t *s_off = s_cnt + p.P;                                                          // [P + 1]    MODE 1
    uint32_t *s_cur = s_off + (MODE == 1 ? p.P + 1 : 0);                                    // [P]        MODE 1
    uint16_t *s_sb = reinterpret_cast<uint16_t *>(s_cur + (MODE == 1 ? p.P : 0));           // [K1_TILE]  MODE 1: bin of each staged k-mer
    __shared__ uint32_t s_warp[K1_THREADS / 32];

    const int t = threadIdx.x, k = p.k;
    const uint64_t first = blockIdx.x, stride = gridDim.x;

    auto issue = [&](uint64_t tile, int stage) {      // thread 0 only
        const uint64_t start = tile * (uint64_t)K1_TILE;
        if (start + K1T_RAW <= p.n_bytes) {
            mbar_expect_tx(&s_bar[stage], K1T_RAW);
            tma_load_1d(s_raw + stage * K1T_RAW, p.bytes + start, K1T_RAW, &s_bar[stage]);
        } else {                                      // ragged last tile: plain loads, positions past the end read as breaks
            for (int i = 0; i < K1T_RAW; i++) { uint64_t q = start + i; s_raw[stage * K1T_RAW + i] = q < p.n_bytes ? p.bytes[q] : (uint8_t)'\n'; }
            mbar_expect_tx(&s_bar[stage], 0);
        }
    };

    if (t == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    for (uint32_t i = t; i < p.P; i += K1_THREADS) s_cnt[i] = 0;
    __syncthreads();
    if (t == 0 && first < p.n_tiles) issue(first, 0);

    uint32_t it = 0;
    for (uint64_t tile = first; tile < p.n_tiles; tile += stride, it++) {
        const int stage = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        const uint64_t tile_start = tile * (uint64_t)K1_TILE;
        if (t == 0 && tile + stride < p.n_tiles) issue(tile + stride, stage ^ 1);   // prefetch the next tile
        mbar_wait(&s_bar[stage], parity);
        // ---- phase A: ASCII -> 2 bit from the staged bytes
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(s_raw + stage * K1T_RAW + t * 32);
            uint64_t pk; uint32_t vm;
            pack32(src[0], src[1], pk, vm);
            s_pk[t] = pk; s_vm[t] = vm;
            if (t == 0) {
                const uint4 *h = reinterpret_cast<const uint4 *>(s_raw + stage * K1T_RAW + K1_TILE);
                pack32(h[0], h[1], pk, vm);
                s_pk[K1_THREADS] = pk; s_vm[K1_THREADS] = vm;
            }
        }
        __syncthreads();
        uint64_t hi = s_pk[t], lo = s_pk[t + 1];
        uint64_t fwd = hi >> (64 - 2 * k);
        uint64_t rc = revcomp64(fwd, k);

        if (MODE == 0) {
            const uint64_t V = ((uint64_t)s_vm[t] << 32) | s_vm[t + 1];
            const uint32_t valid32 = (uint32_t)(kmer_valid_mask(V, k) >> 32);
            uint32_t packed[K1_PER_THREAD / 2];
#pragma unroll
            for (int i = 0; i < K1_PER_THREAD; i++) {
                if (i > 0) {
                    hi = (hi << 2) | (lo >> 62); lo <<= 2;
                    fwd = hi >> (64 - 2 * k);
                    rc = (rc >> 2) | (((fwd & 3) ^ 2) << (2 * (k - 1)));
                }
                const uint64_t c = fwd < rc ? fwd : rc;
                uint32_t bin = hash32_to_bin(bin_hash32(c), p.P);
                const bool ok = (valid32 >> (31 - i)) & 1;
                if (ok) atomicAdd(&s_cnt[bin], 1u); else bin = K1_NOBIN;
                if (i & 1) packed[i >> 1] |= bin << 16; else packed[i >> 1] = bin;
            }
            uint4 *dst = reinterpret_cast<uint4 *>(p.pos_bin + tile_start + (uint64_t)t * K1_PER_THREAD);
#pragma unroll
            for (int v = 0; v < K1_PER_THREAD / 8; v++) dst[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
            __syncthreads();     // s_pk / the raw stage are reused by the next iterations
        } else {
            uint32_t packed[K1_PER_THREAD / 2];
            const uint4 *srcb = reinterpret_cast<const uint4 *>(p.pos_bin + tile_start + (uint64_t)t * K1_PER_THREAD);
#pragma unroll
            for (int v = 0; v < K1_PER_THREAD / 8; v++) { uint4 q = __ldg(srcb + v); packed[4 * v] = q.x; packed[4 * v + 1] = q.y; packed[4 * v + 2] = q.z; packed[4 * v + 3] = q.w; }
            // tile counts per bin (the bins come from pass 0: no hashing here)
#pragma unroll
            for (int i = 0; i < K1_PER_THREAD; i++) {
                uint32_t bin = (packed[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
                if (bin != K1_NOBIN) atomicAdd(&s_cnt[bin], 1u);
            }
            __syncthreads();
            {   // exclusive scan of s_cnt + one global reservation per non-empty bin
                const uint32_t per = (p.P + K1_THREADS - 1) / K1_THREADS;
                uint32_t b0 = t * per, sum = 0;
                for (uint32_t j = 0; j < per; j++) if (b0 + j < p.P) sum += s_cnt[b0 + j];
                uint32_t incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane_id() >= d) incl += n; }
                if (lane_id() == 31) s_warp[t >> 5] = incl;
                __syncthreads();
                if (t < 32) {
                    uint32_t wv = t < K1_THREADS / 32 ? s_warp[t] : 0, wi = wv;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, d); if ((int)lane_id() >= d) wi += n; }
                    if (t < K1_THREADS / 32) s_warp[t] = wi - wv;
                }
                __syncthreads();
                uint32_t run = s_warp[t >> 5] + incl - sum;
                for (uint32_t j = 0; j < per; j++) if (b0 + j < p.P) {
                    uint32_t cn = s_cnt[b0 + j];
                    s_off[b0 + j] = run; s_cur[b0 + j] = run; s_cnt[b0 + j] = 0;
                    if (cn) s_gbase[b0 + j] = p.bin_off[b0 + j] + atomicAdd(&p.bin_cursor[b0 + j], (unsigned long long)cn);
                    run += cn;
                }
                if (t == K1_THREADS - 1) s_off[p.P] = run;
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < K1_PER_THREAD; i++) {
                if (i > 0) {
                    hi = (hi << 2) | (lo >> 62); lo <<= 2;
                    fwd = hi >> (64 - 2 * k);
                    rc = (rc >> 2) | (((fwd & 3) ^ 2) << (2 * (k - 1)));
                }
                uint32_t bin = (packed[i >> 1] >> ((i & 1) * 16)) & 0xFFFFu;
                if (bin != K1_NOBIN) {
                    uint32_t pos = atomicAdd(&s_cur[bin], 1u);
                    s_stage[pos] = fwd < rc ? fwd : rc; s_sb[pos] = (uint16_t)bin;
                }
            }
            __syncthreads();
            const uint32_t total = s_off[p.P];
            for (uint32_t j = t; j < total; j += K1_THREADS) {
                uint32_t b = s_sb[j];
                p.out_kmers[s_gbase[b] + (j - s_off[b])] = s_stage[j];
            }
            __syncthreads();
        }
    }
    if (MODE == 0) {
        for (uint32_t i = t; i < p.P; i += K1_THREADS) { uint32_t c = s_cnt[i]; if (c) atomicAdd(&p.bin_count[i], (unsigned long long)c); }
    }
}

static inline size_t k1t_smem_bytes(int mode, uint32_t P) {
    size_t s = 2 * (size_t)K1T_RAW + 16 + (size_t)(K1_THREADS + 2) * 8 + (size_t)(K1_THREADS + 2) * 4 + (size_t)P * 4;
    if (mode == 1) s += (size_t)P * 8 + (size_t)K1_TILE * 8 + (size_t)(P + 1) * 4 + (size_t)P * 4 + (size_t)K1_TILE * 2;
    return s + 128;
}

// ------------------------------------------------------------------------------------------------ single pass
//
// k1t_single: the same extraction in ONE pass over the reads.  Hash bins are statistically alike, so instead of counting
// first (pass 0 above) every bin gets a fixed region of `bin_cap` k-mers (mean + slack, the host sizes it from the byte count)
// and tiles reserve their runs from per-bin cursors directly.  A bin that outgrows its region (possible only with heavily
// repeated k-mers) raises *overflow and the host redoes stage 1 with the exact two-pass kernels -- the result never depends
// on the estimate.  Per tile: pack, roll + hash once (the rank inside the tile's bin comes back from the counting atomic and
// stays in a register with the bin), scan + one global reservation per (tile, bin), roll again into the staging area at
// offset + rank, coalesced run stores.  Compared with the two passes this reads the input once (which is what allows `bytes`
// to be pinned HOST memory read over PCIe by the TMA engine: the H2D copy and K1 become one step), and it neither writes nor
// re-reads the per-position bins.
struct K1XParams {
    const uint8_t *bytes;            // device memory, or pinned host memory mapped into the device address space
    const uint8_t *tail;             // device copy of bytes[tail_start, n_bytes): the ragged last tiles are read from here
    uint64_t tail_start;
    uint64_t n_bytes, n_tiles;
    int k; uint32_t P;
    uint64_t bin_cap;                // bin b owns out_kmers[b * bin_cap, (b + 1) * bin_cap)
    unsigned long long *bin_cursor;  // [P] zero-initialised; ends as the bin counts
    uint64_t *out_kmers;
    int *overflow;
};

#define K1X_NONE 0xFFFFFFFFu
#define K1X_SKIP 0xFFFFFFFFFFFFFFFFULL

// MATCH = true ranks the positions of a tile inside their bin without atomics: lanes of a warp that hit the same bin find each
// other with match.any, the first of them bumps a counter that is private to (bin, warp) with a plain read-modify-write, and the
// per-warp counts are prefix-summed per bin before the second sweep.  (A shared-memory atomic that returns a value costs more than
// one that does not; this trades it for match + shuffle.)  Needs P * 8 warp counters in shared memory: P <= K1X_MATCH_MAX_P.
constexpr uint32_t K1X_MATCH_MAX_P = 512;

template <bool MATCH>
__global__ void __launch_bounds__(K1_THREADS) k1t_single(K1XParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint8_t *s_raw = smem_raw;                                                              // [2][K1T_RAW]  TMA destinations
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + 2 * K1T_RAW);                 // [2] mbarriers
    uint64_t *s_pk = s_bar + 2;                                                             // [THREADS + 2]
    uint64_t *s_gbase = s_pk + K1_THREADS + 2;                                              // [P]
    uint64_t *s_stage = s_gbase + p.P;                                                      // [K1_TILE]
    uint32_t *s_vm = reinterpret_cast<uint32_t *>(s_stage + K1_TILE);                       // [THREADS + 2]
    uint32_t *s_cnt = s_vm + K1_THREADS + 2;                                                // [P]
    uint32_t *s_off = s_cnt + p.P;                                                          // [P + 1]
    uint32_t *s_wc = s_off + p.P + 1;                                                       // [P * 8] MATCH: per (bin, warp) counters
    uint16_t *s_sb = reinterpret_cast<uint16_t *>(s_wc + (MATCH ? p.P * (K1_THREADS / 32) : 0));   // [K1_TILE] bin of each staged k-mer
    __shared__ uint32_t s_warp[K1_THREADS / 32];
    constexpr uint32_t NW = K1_THREADS / 32;

    const int t = threadIdx.x, k = p.k;
    const uint32_t lane = lane_id(), warp = (uint32_t)t >> 5;
    const uint64_t first = blockIdx.x, stride = gridDim.x;

    auto issue = [&](uint64_t tile, int stage) {      // thread 0 only
        const uint64_t start = tile * (uint64_t)K1_TILE;
        if (start + K1T_RAW <= p.n_bytes) {
            mbar_expect_tx(&s_bar[stage], K1T_RAW);
            tma_load_1d(s_raw + stage * K1T_RAW, p.bytes + start, K1T_RAW, &s_bar[stage]);
        } else {                                      // ragged last tiles: plain loads from the device-resident tail
            for (int i = 0; i < K1T_RAW; i++) { uint64_t q = start + i; s_raw[stage * K1T_RAW + i] = q < p.n_bytes ? p.tail[q - p.tail_start] : (uint8_t)'\n'; }
            mbar_expect_tx(&s_bar[stage], 0);
        }
    };

    if (t == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    for (uint32_t i = t; i < p.P; i += K1_THREADS) s_cnt[i] = 0;
    if (MATCH) for (uint32_t i = t; i < p.P * NW; i += K1_THREADS) s_wc[i] = 0;
    __syncthreads();
    if (t == 0 && first < p.n_tiles) issue(first, 0);

    uint32_t it = 0;
    for (uint64_t tile = first; tile < p.n_tiles; tile += stride, it++) {
        const int stage = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        if (t == 0 && tile + stride < p.n_tiles) issue(tile + stride, stage ^ 1);   // prefetch the next tile
        mbar_wait(&s_bar[stage], parity);
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(s_raw + stage * K1T_RAW + t * 32);
            uint64_t pk; uint32_t vm;
            pack32(src[0], src[1], pk, vm);
            s_pk[t] = pk; s_vm[t] = vm;
            if (t == 0) {
                const uint4 *h = reinterpret_cast<const uint4 *>(s_raw + stage * K1T_RAW + K1_TILE);
                pack32(h[0], h[1], pk, vm);
                s_pk[K1_THREADS] = pk; s_vm[K1_THREADS] = vm;
            }
        }
        __syncthreads();
        const uint64_t hi0 = s_pk[t], lo0 = s_pk[t + 1];
        const uint64_t V = ((uint64_t)s_vm[t] << 32) | s_vm[t + 1];
        const uint32_t valid32 = (uint32_t)(kmer_valid_mask(V, k) >> 32);

        // ---- sweep 1: bin of every valid position + its rank inside the tile's bin (one shared-memory atomic)
        uint32_t br[K1_PER_THREAD];                   // rank << 11 | bin   (P <= 2048, rank < 8192)
        {
            uint64_t hi = hi0, lo = lo0;
            uint64_t fwd = hi >> (64 - 2 * k);
            uint64_t rc = revcomp64(fwd, k);
#pragma unroll
            for (int i = 0; i < K1_PER_THREAD; i++) {
                if (i > 0) {
                    hi = (hi << 2) | (lo >> 62); lo <<= 2;
                    fwd = hi >> (64 - 2 * k);
                    rc = (rc >> 2) | (((fwd & 3) ^ 2) << (2 * (k - 1)));
                }
                const uint64_t c = fwd < rc ? fwd : rc;
                const uint32_t bin = hash32_to_bin(bin_hash32(c), p.P);
                const bool ok = (valid32 >> (31 - i)) & 1;
                if (!MATCH) {
                    br[i] = ok ? ((atomicAdd(&s_cnt[bin], 1u) << 11) | bin) : K1X_NONE;
                } else {
                    const unsigned m = __match_any_sync(0xFFFFFFFFu, ok ? bin : K1X_NONE);
                    const int leader = __ffs(m) - 1;
                    uint32_t base = 0;
                    if (ok && (int)lane == leader) { base = s_wc[bin * NW + warp]; s_wc[bin * NW + warp] = base + __popc(m); }
                    base = __shfl_sync(0xFFFFFFFFu, base, leader);
                    br[i] = ok ? (((base + __popc(m & ((1u << lane) - 1))) << 11) | bin) : K1X_NONE;
                    __syncwarp();
                }
            }
        }
        __syncthreads();
        if (MATCH) {   // per bin: prefix over the warps' counts (becomes each warp's base), total into s_cnt
            for (uint32_t b = t; b < p.P; b += K1_THREADS) {
                uint32_t run = 0;
#pragma unroll
                for (uint32_t w = 0; w < NW; w++) { const uint32_t c = s_wc[b * NW + w]; s_wc[b * NW + w] = run; run += c; }
                s_cnt[b] = run;
            }
            __syncthreads();
        }
        {   // exclusive scan of s_cnt + one global reservation per non-empty bin, checked against the bin's region
            const uint32_t per = (p.P + K1_THREADS - 1) / K1_THREADS;
            uint32_t b0 = t * per, sum = 0;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.P) sum += s_cnt[b0 + j];
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane_id() >= d) incl += n; }
            if (lane_id() == 31) s_warp[t >> 5] = incl;
            __syncthreads();
            if (t < 32) {
                uint32_t wv = t < K1_THREADS / 32 ? s_warp[t] : 0, wi = wv;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, d); if ((int)lane_id() >= d) wi += n; }
                if (t < K1_THREADS / 32) s_warp[t] = wi - wv;
            }
            __syncthreads();
            uint32_t run = s_warp[t >> 5] + incl - sum;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.P) {
                const uint32_t cn = s_cnt[b0 + j];
                s_off[b0 + j] = run; s_cnt[b0 + j] = 0;
                if (cn) {
                    const unsigned long long old = atomicAdd(&p.bin_cursor[b0 + j], (unsigned long long)cn);
                    if (old + cn > p.bin_cap) { *p.overflow = 1; s_gbase[b0 + j] = K1X_SKIP; }
                    else s_gbase[b0 + j] = (uint64_t)(b0 + j) * p.bin_cap + old;
                }
                run += cn;
            }
            if (t == K1_THREADS - 1) s_off[p.P] = run;
        }
        __syncthreads();
        // ---- sweep 2: roll again, counting-sort into the staging area
        {
            uint64_t hi = hi0, lo = lo0;
            uint64_t fwd = hi >> (64 - 2 * k);
            uint64_t rc = revcomp64(fwd, k);
#pragma unroll
            for (int i = 0; i < K1_PER_THREAD; i++) {
                if (i > 0) {
                    hi = (hi << 2) | (lo >> 62); lo <<= 2;
                    fwd = hi >> (64 - 2 * k);
                    rc = (rc >> 2) | (((fwd & 3) ^ 2) << (2 * (k - 1)));
                }
                if (br[i] != K1X_NONE) {
                    const uint32_t bin = br[i] & 2047u, pos = s_off[bin] + (br[i] >> 11) + (MATCH ? s_wc[bin * NW + warp] : 0u);
                    s_stage[pos] = fwd < rc ? fwd : rc; s_sb[pos] = (uint16_t)bin;
                }
            }
        }
        __syncthreads();
        const uint32_t total = s_off[p.P];
        for (uint32_t j = t; j < total; j += K1_THREADS) {
            const uint32_t b = s_sb[j];
            const uint64_t g = s_gbase[b];
            if (g != K1X_SKIP) p.out_kmers[g + (j - s_off[b])] = s_stage[j];
        }
        if (MATCH) for (uint32_t i = t; i < p.P * NW; i += K1_THREADS) s_wc[i] = 0;   // (sweep 2 is behind the barrier above)
        __syncthreads();
    }
}

static inline size_t k1x_smem_bytes(uint32_t P, bool match) {
    return (match ? (size_t)P * (K1_THREADS / 32) * 4 : 0) + 2 * (size_t)K1T_RAW + 16 + (size_t)(K1_THREADS + 2) * 8 + (size_t)P * 8 + (size_t)K1_TILE * 8 + (size_t)(K1_THREADS + 2) * 4 +
           (size_t)P * 4 + (size_t)(P + 1) * 4 + (size_t)K1_TILE * 2 + 128;
}

// ------------------------------------------------------------------------------------------------ single pass, 512 threads
//
// k1t_single512: the same kernel with 512 threads per 8192-position tile, 16 positions per thread, capped at 64 registers so that
// two CTAs (32 warps) are resident per SM instead of 16 warps.  ncu on k1t_single: IPC 1.85, 23 % occupancy, the top stall is
// `wait` (dependent ALU chains of the rolling k-mer and the hash) -- more resident warps is the direct remedy.  Thread t owns
// positions [16t, 16t + 16): its 46 bases come from the packed words t/2 and t/2 + 1, shifted by half a word for odd t.
constexpr int K1H_THREADS = 512;
constexpr int K1H_PER_THREAD = K1_TILE / K1H_THREADS;     // 16

__global__ void __launch_bounds__(K1H_THREADS, 2) k1t_single512(K1XParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint8_t *s_raw = smem_raw;                                                              // [2][K1T_RAW]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + 2 * K1T_RAW);                 // [2]
    uint64_t *s_pk = s_bar + 2;                                                             // [256 + 2] packed bases, 32 per word
    uint64_t *s_gbase = s_pk + K1_THREADS + 2;                                              // [P]
    uint64_t *s_stage = s_gbase + p.P;                                                      // [K1_TILE]
    uint32_t *s_vm = reinterpret_cast<uint32_t *>(s_stage + K1_TILE);                       // [256 + 2]
    uint32_t *s_cnt = s_vm + K1_THREADS + 2;                                                // [P]
    uint32_t *s_off = s_cnt + p.P;                                                          // [P + 1]
    uint16_t *s_sb = reinterpret_cast<uint16_t *>(s_off + p.P + 1);                         // [K1_TILE]
    __shared__ uint32_t s_warp[K1H_THREADS / 32];

    const int t = threadIdx.x, k = p.k;
    const uint64_t first = blockIdx.x, stride = gridDim.x;

    auto issue = [&](uint64_t tile, int stage) {      // thread 0 only
        const uint64_t start = tile * (uint64_t)K1_TILE;
        if (start + K1T_RAW <= p.n_bytes) {
            mbar_expect_tx(&s_bar[stage], K1T_RAW);
            tma_load_1d(s_raw + stage * K1T_RAW, p.bytes + start, K1T_RAW, &s_bar[stage]);
        } else {
            for (int i = 0; i < K1T_RAW; i++) { uint64_t q = start + i; s_raw[stage * K1T_RAW + i] = q < p.n_bytes ? p.tail[q - p.tail_start] : (uint8_t)'\n'; }
            mbar_expect_tx(&s_bar[stage], 0);
        }
    };

    if (t == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    for (uint32_t i = t; i < p.P; i += K1H_THREADS) s_cnt[i] = 0;
    __syncthreads();
    if (t == 0 && first < p.n_tiles) issue(first, 0);

    uint32_t it = 0;
    for (uint64_t tile = first; tile < p.n_tiles; tile += stride, it++) {
        const int stage = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        if (t == 0 && tile + stride < p.n_tiles) issue(tile + stride, stage ^ 1);
        mbar_wait(&s_bar[stage], parity);
        if (t <= K1_THREADS) {                        // 256 words of the tile + the halo word
            const uint4 *src = reinterpret_cast<const uint4 *>(s_raw + stage * K1T_RAW + t * 32);
            uint64_t pk; uint32_t vm;
            pack32(src[0], src[1], pk, vm);
            s_pk[t] = pk; s_vm[t] = vm;
        }
        __syncthreads();
        const int w = t >> 1;
        uint64_t hi0, lo0, V;
        if (t & 1) {
            hi0 = (s_pk[w] << 32) | (s_pk[w + 1] >> 32); lo0 = s_pk[w + 1] << 32;
            V = ((uint64_t)s_vm[w] << 48) | ((uint64_t)s_vm[w + 1] << 16);
        } else {
            hi0 = s_pk[w]; lo0 = s_pk[w + 1];
            V = ((uint64_t)s_vm[w] << 32) | s_vm[w + 1];
        }
        const uint32_t valid16 = (uint32_t)(kmer_valid_mask(V, k) >> 48);   // bit (15-i): k-mer at position 16t + i valid

        uint32_t br[K1H_PER_THREAD];                  // rank << 11 | bin
        {
            uint64_t hi = hi0, lo = lo0;
            uint64_t fwd = hi >> (64 - 2 * k);
            uint64_t rc = revcomp64(fwd, k);
#pragma unroll
            for (int i = 0; i < K1H_PER_THREAD; i++) {
                if (i > 0) {
                    hi = (hi << 2) | (lo >> 62); lo <<= 2;
                    fwd = hi >> (64 - 2 * k);
                    rc = (rc >> 2) | (((fwd & 3) ^ 2) << (2 * (k - 1)));
                }
                const uint64_t c = fwd < rc ? fwd : rc;
                const uint32_t bin = hash32_to_bin(bin_hash32(c), p.P);
                br[i] = ((valid16 >> (15 - i)) & 1) ? ((atomicAdd(&s_cnt[bin], 1u) << 11) | bin) : K1X_NONE;
            }
        }
        __syncthreads();
        {
            const uint32_t per = (p.P + K1H_THREADS - 1) / K1H_THREADS;
            uint32_t b0 = t * per, sum = 0;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.P) sum += s_cnt[b0 + j];
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane_id() >= d) incl += n; }
            if (lane_id() == 31) s_warp[t >> 5] = incl;
            __syncthreads();
            if (t < 32) {
                uint32_t wv = t < K1H_THREADS / 32 ? s_warp[t] : 0, wi = wv;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, d); if ((int)lane_id() >= d) wi += n; }
                if (t < K1H_THREADS / 32) s_warp[t] = wi - wv;
            }
            __syncthreads();
            uint32_t run = s_warp[t >> 5] + incl - sum;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.P) {
                const uint32_t cn = s_cnt[b0 + j];
                s_off[b0 + j] = run; s_cnt[b0 + j] = 0;
                if (cn) {
                    const unsigned long long old = atomicAdd(&p.bin_cursor[b0 + j], (unsigned long long)cn);
                    if (old + cn > p.bin_cap) { *p.overflow = 1; s_gbase[b0 + j] = K1X_SKIP; }
                    else s_gbase[b0 + j] = (uint64_t)(b0 + j) * p.bin_cap + old;
                }
                run += cn;
            }
            if (t == K1H_THREADS - 1) s_off[p.P] = run;
        }
        __syncthreads();
        {
            uint64_t hi = hi0, lo = lo0;
            uint64_t fwd = hi >> (64 - 2 * k);
            uint64_t rc = revcomp64(fwd, k);
#pragma unroll
            for (int i = 0; i < K1H_PER_THREAD; i++) {
                if (i > 0) {
                    hi = (hi << 2) | (lo >> 62); lo <<= 2;
                    fwd = hi >> (64 - 2 * k);
                    rc = (rc >> 2) | (((fwd & 3) ^ 2) << (2 * (k - 1)));
                }
                if (br[i] != K1X_NONE) {
                    const uint32_t bin = br[i] & 2047u, pos = s_off[bin] + (br[i] >> 11);
                    s_stage[pos] = fwd < rc ? fwd : rc; s_sb[pos] = (uint16_t)bin;
                }
            }
        }
        __syncthreads();
        const uint32_t total = s_off[p.P];
        for (uint32_t j = t; j < total; j += K1H_THREADS) {
            const uint32_t b = s_sb[j];
            const uint64_t g = s_gbase[b];
            if (g != K1X_SKIP) p.out_kmers[g + (j - s_off[b])] = s_stage[j];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ sharded histogram pass, rolling
//
// k1t_shard_hist: pass 0 of the SHARDED K1 in the style of k1t_extract<0> (TMA-fed tiles of 8192 positions, 32 consecutive
// positions per thread, rolling k-mers) instead of k1s_extract<0>'s lane-consecutive, non-rolling 2048-position tiles.  The bin of
// a k-mer is destination-major: owner(minimizer_k) * P_local + local hash bin, like k1s_extract, so that pass 1 is simply
// k1t_extract<1> over P = world * P_local bins.  Minimizer = smallest mmer_rank over the w = k-m+1 canonical m-mers inside the
// k-mer (the same function as k1s_extract / min_rank_of): the ranks of all m-mers of the tile (+ halo) are computed once, rolling,
// into shared memory, a van Herk / Gil-Werman pass turns them into block prefix / suffix minima, and every k-mer then needs two
// shared-memory reads.  Rank arrays are padded by one word per 32 (a thread owns 32 consecutive positions).
// Opt-in (CDBG_SHARD_K1=roll): written after the round's last GPU run, verified under the CPU emulator only.
constexpr int K1TS_RAW = K1_TILE + 64;                  // m-mers starting in the 32-position halo need 32 more bases
constexpr int K1TS_RANKS = K1_TILE + 32;                // ranks incl. halo (w - 1 <= 30)
__device__ __forceinline__ uint32_t k1ts_pad(uint32_t q) { return q + (q >> 5); }
constexpr int K1TS_RANK_WORDS = K1TS_RANKS + K1TS_RANKS / 32 + 2;

struct K1TSParams {
    const uint8_t *bytes; uint64_t n_bytes; uint64_t n_tiles;
    int k, m; uint32_t world, P_local;
    unsigned long long *bin_count;
    uint16_t *pos_bin;         // [n_tiles * K1_TILE]
};

__global__ void __launch_bounds__(K1_THREADS) k1t_shard_hist(K1TSParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t P = p.world * p.P_local;
    uint8_t *s_raw = smem_raw;                                                              // [2][K1TS_RAW]
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + 2 * K1TS_RAW);                // [2]
    uint64_t *s_pk = s_bar + 2;                                                             // [THREADS + 2] (2 halo words)
    uint32_t *s_vm = reinterpret_cast<uint32_t *>(s_pk + K1_THREADS + 2);                   // [THREADS + 2]
    uint32_t *s_R = s_vm + K1_THREADS + 2;                                                  // [K1TS_RANK_WORDS] ranks -> block prefix minima
    uint32_t *s_S = s_R + K1TS_RANK_WORDS;                                                  // [K1TS_RANK_WORDS] block suffix minima
    uint32_t *s_cnt = s_S + K1TS_RANK_WORDS;                                                // [P]

    const int t = threadIdx.x, k = p.k, m = p.m, w = k - m + 1;
    const uint64_t first = blockIdx.x, stride = gridDim.x;

    auto issue = [&](uint64_t tile, int stage) {      // thread 0 only
        const uint64_t start = tile * (uint64_t)K1_TILE;
        if (start + K1TS_RAW <= p.n_bytes) {
            mbar_expect_tx(&s_bar[stage], K1TS_RAW);
            tma_load_1d(s_raw + stage * K1TS_RAW, p.bytes + start, K1TS_RAW, &s_bar[stage]);
        } else {
            for (int i = 0; i < K1TS_RAW; i++) { uint64_t q = start + i; s_raw[stage * K1TS_RAW + i] = q < p.n_bytes ? p.bytes[q] : (uint8_t)'\n'; }
            mbar_expect_tx(&s_bar[stage], 0);
        }
    };

    if (t == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        mbar_fence_init();
    }
    for (uint32_t i = t; i < P; i += K1_THREADS) s_cnt[i] = 0;
    __syncthreads();
    if (t == 0 && first < p.n_tiles) issue(first, 0);

    uint32_t it = 0;
    for (uint64_t tile = first; tile < p.n_tiles; tile += stride, it++) {
        const int stage = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        const uint64_t tile_start = tile * (uint64_t)K1_TILE;
        if (t == 0 && tile + stride < p.n_tiles) issue(tile + stride, stage ^ 1);
        mbar_wait(&s_bar[stage], parity);
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(s_raw + stage * K1TS_RAW + t * 32);
            uint64_t pk; uint32_t vm;
            pack32(src[0], src[1], pk, vm);
            s_pk[t] = pk; s_vm[t] = vm;
            if (t < 2) {                              // the two halo words
                const uint4 *h = reinterpret_cast<const uint4 *>(s_raw + stage * K1TS_RAW + K1_TILE + t * 32);
                pack32(h[0], h[1], pk, vm);
                s_pk[K1_THREADS + t] = pk; s_vm[K1_THREADS + t] = vm;
            }
        }
        __syncthreads();
        // ---- ranks of the canonical m-mers starting at every position (rolling; thread 0 also does the 32 halo positions)
        for (int seg = t; seg < K1_THREADS + 1; seg += K1_THREADS) {
            uint64_t hi = s_pk[seg], lo = s_pk[seg + 1];
            uint64_t fm = hi >> (64 - 2 * m);
            uint64_t rm = revcomp64(fm, m);
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
                if (i > 0) {
                    hi = (hi << 2) | (lo >> 62); lo <<= 2;
                    fm = hi >> (64 - 2 * m);
                    rm = (rm >> 2) | (((fm & 3) ^ 2) << (2 * (m - 1)));
                }
                s_R[k1ts_pad((uint32_t)(seg * 32 + i))] = mmer_rank((uint32_t)(fm < rm ? fm : rm));
            }
        }
        __syncthreads();
        // ---- van Herk / Gil-Werman: per block of w ranks, suffix minima into s_S, prefix minima in place
        for (int b = t; b * w < K1TS_RANKS; b += K1_THREADS) {
            const int lo_q = b * w, hi_q = min(lo_q + w, K1TS_RANKS);
            uint32_t run = 0xFFFFFFFFu;
            for (int q = hi_q - 1; q >= lo_q; q--) { run = min(run, s_R[k1ts_pad((uint32_t)q)]); s_S[k1ts_pad((uint32_t)q)] = run; }
            run = 0xFFFFFFFFu;
            for (int q = lo_q; q < hi_q; q++) { run = min(run, s_R[k1ts_pad((uint32_t)q)]); s_R[k1ts_pad((uint32_t)q)] = run; }
        }
        __syncthreads();
        // ---- rolling k-mers, destination-major bin, CTA-private histogram, per-position bins written out
        {
            uint64_t hi = s_pk[t], lo = s_pk[t + 1];
            uint64_t fwd = hi >> (64 - 2 * k);
            uint64_t rc = revcomp64(fwd, k);
            const uint64_t V = ((uint64_t)s_vm[t] << 32) | s_vm[t + 1];
            const uint32_t valid32 = (uint32_t)(kmer_valid_mask(V, k) >> 32);
            uint32_t packed[K1_PER_THREAD / 2];
#pragma unroll
            for (int i = 0; i < K1_PER_THREAD; i++) {
                if (i > 0) {
                    hi = (hi << 2) | (lo >> 62); lo <<= 2;
                    fwd = hi >> (64 - 2 * k);
                    rc = (rc >> 2) | (((fwd & 3) ^ 2) << (2 * (k - 1)));
                }
                const uint64_t c = fwd < rc ? fwd : rc;
                const uint32_t q = (uint32_t)(t * 32 + i);
                const uint32_t mr = min(s_S[k1ts_pad(q)], s_R[k1ts_pad(q + (uint32_t)w - 1)]);
                uint32_t bin = owner_of_rank(mr, p.world) * p.P_local + hash32_to_bin(bin_hash32(c), p.P_local);
                const bool ok = (valid32 >> (31 - i)) & 1;
                if (ok) atomicAdd(&s_cnt[bin], 1u); else bin = K1_NOBIN;
                if (i & 1) packed[i >> 1] |= bin << 16; else packed[i >> 1] = bin;
            }
            uint4 *dst = reinterpret_cast<uint4 *>(p.pos_bin + tile_start + (uint64_t)t * K1_PER_THREAD);
#pragma unroll
            for (int v = 0; v < K1_PER_THREAD / 8; v++) dst[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
        }
        __syncthreads();
    }
    for (uint32_t i = t; i < P; i += K1_THREADS) { uint32_t c = s_cnt[i]; if (c) atomicAdd(&p.bin_count[i], (unsigned long long)c); }
}

static inline size_t k1ts_smem_bytes(uint32_t P) {
    return 2 * (size_t)K1TS_RAW + 16 + (size_t)(K1_THREADS + 2) * 8 + (size_t)(K1_THREADS + 2) * 4 + 2 * (size_t)K1TS_RANK_WORDS * 4 + (size_t)P * 4 + 128;
}
