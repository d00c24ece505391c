// k2_smem.cuh -- K2 (k-mer counting + solidity filter, SURVEY a6/a7: PartitionsCommand.cpp:1600-1805,
// CountProcessorSolidity.hpp:186-189) as "partition finer, then count in shared memory".
//
// The L2 path (k2_insert/k2_emit in kernels.cuh) keeps one hash table per coarse bin resident in L2 and pays ~2.3 L2
// operations per k-mer instance; it is bound by L2 atomic throughput (6 % of the HBM roofline on C2).  Here the k-mer stream
// that K1 wrote into P1 coarse bins is split once more, into F fine bins per coarse bin, sized so that the distinct k-mers of a
// fine bin fit a hash table in SHARED memory; one CTA then owns one fine bin at a time and all probing, CAS and counting are
// shared-memory operations.  HBM sees three streaming passes instead of table traffic:
//   k2p_hist     read the stream, histogram of fine bins per coarse bin            I*8 read
//   k2p_scatter  read it again (L2-warm per coarse bin), tile-local counting sort
//                in shared memory, one coalesced run per (tile, fine bin)          I*8 read + I*8 write
//   k2c_count    one CTA per fine bin: stream its k-mers through the table,
//                then scan the table and emit the solid (k-mer, count) pairs       I*8 read + S*12 write
// A fine bin whose distinct k-mers do not fit its table is split further inside k2c_count (see there), so the result never
// depends on the sizing estimate.
#pragma once
#include "kernels.cuh"

constexpr int K2P_THREADS = 512;
constexpr int K2P_PER_THREAD = 16;
constexpr int K2P_TILE = K2P_THREADS * K2P_PER_THREAD;    // 8192 k-mers per tile
constexpr int K2P_MAX_F = 1024;                            // fine bins per coarse bin (10 bits in the packed rank word)
#define K2P_NOKEY 0xFFFFFFFFu

// fine bin inside a coarse bin: 32 bits of mix64 starting at `shift`.  Single GPU: shift 32, the HIGH half (K1's coarse bin comes
// from bin_hash32, the table slot from the LOW half).  Sharded: the local bin already comes from the high half (its top bits), so
// the fine bin uses bits 13..44 -- its leading bits 44..35 are clear of the local bin's (63..) and of the slot's (12..0).
__device__ __forceinline__ uint32_t k2p_fine(uint64_t h, uint32_t F, uint32_t shift) {
    return (uint32_t)((((h >> shift) & 0xFFFFFFFFULL) * (uint64_t)F) >> 32);
}

struct K2PParams {
    const uint64_t *kmers;          // K1 output, coarse bin b = [bin_beg[b], bin_end[b]) (regions may have gaps)
    const uint64_t *bin_beg, *bin_end;   // [P1] each
    const uint32_t *tile_first;     // [P1 + 1] first tile of each coarse bin (tiles never span coarse bins)
    const uint32_t *group;          // [P1] or null: fine bins are shared by all coarse "bins" of one group (sharded mode: a bin received
                                    //      from several GPUs is one run per source; the runs of a bin form a group).  null: group = bin
    uint32_t P1, F;
    uint32_t fine_shift;            // see k2p_fine
    uint32_t n_tiles;
    uint32_t *fine_cnt;             // [P1 * F]     hist: out
    const uint64_t *fine_off;       // [P1 * F + 1] scatter: exclusive offsets into `out`
    uint32_t *fine_cursor;          // [P1 * F]     scatter: running cursors, zero-initialised
    uint64_t *out;                  // [I] scatter: the stream re-ordered by (coarse, fine) bin
    // "regions" mode (no histogram pass): fine bin f of coarse bin b owns out[region_base[b] + f * region_cap[b] ...) with room for
    // region_cap[b] k-mers (mean + slack: fine bins are hash-uniform inside a coarse bin); a bin that outgrows its region raises
    // *overflow and the host redoes the partition with exact offsets.  fine_cursor ends as the fine counts.
    const uint64_t *region_base;    // [P1] or null (exact mode: fine_off)
    const uint32_t *region_cap;     // [P1]
    int *overflow;
};
#define K2P_SKIP 0xFFFFFFFFFFFFFFFFULL

// coarse bin that owns `tile`: largest b with tile_first[b] <= tile (empty coarse bins have tile_first[b] == tile_first[b+1])
__device__ __forceinline__ uint32_t k2p_bin_of_tile(const uint32_t *__restrict__ tile_first, uint32_t P1, uint32_t tile) {
    uint32_t lo = 0, hi = P1;       // invariant: tile_first[lo] <= tile < tile_first[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (tile_first[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(K2P_THREADS) k2p_hist(K2PParams p) {
    __shared__ uint32_t s_cnt[K2P_MAX_F];
    const uint32_t t = threadIdx.x;
    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t b = k2p_bin_of_tile(p.tile_first, p.P1, tile);
        const uint64_t beg = p.bin_beg[b] + (uint64_t)(tile - p.tile_first[b]) * K2P_TILE;
        const uint64_t end = min(beg + (uint64_t)K2P_TILE, p.bin_end[b]);
        for (uint32_t i = t; i < p.F; i += K2P_THREADS) s_cnt[i] = 0;
        __syncthreads();
        uint64_t key[K2P_PER_THREAD];
#pragma unroll
        for (int j = 0; j < K2P_PER_THREAD; j++) {
            const uint64_t i = beg + (uint64_t)j * K2P_THREADS + t;
            key[j] = i < end ? p.kmers[i] : EMPTY_KEY;
        }
#pragma unroll
        for (int j = 0; j < K2P_PER_THREAD; j++)
            if (key[j] != EMPTY_KEY) atomicAdd(&s_cnt[k2p_fine(mix64(key[j]), p.F, p.fine_shift)], 1u);
        __syncthreads();
        // the thread that flushes counter i is the one that zeroes it for the next tile: no barrier needed in between
        for (uint32_t i = t; i < p.F; i += K2P_THREADS) {
            const uint32_t c = s_cnt[i];
            if (c) atomicAdd(&p.fine_cnt[(size_t)(p.group ? p.group[b] : b) * p.F + i], c);
        }
    }
}

// MATCH = true: ranks without atomics (match.any inside the warp + counters private to (fine bin, warp)), see k1t_single.
constexpr uint32_t K2P_MATCH_MAX_F = 256;

template <bool MATCH>
__global__ void __launch_bounds__(K2P_THREADS, 2) k2p_scatter(K2PParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *s_stage = reinterpret_cast<uint64_t *>(smem_raw);                  // [K2P_TILE] the tile sorted by fine bin
    uint16_t *s_sb = reinterpret_cast<uint16_t *>(s_stage + K2P_TILE);           // [K2P_TILE] fine bin of each staged k-mer
    uint32_t *s_wc = reinterpret_cast<uint32_t *>(s_sb + K2P_TILE);              // [F * 16] MATCH: per (fine bin, warp) counters
    constexpr uint32_t NW = K2P_THREADS / 32;
    __shared__ uint64_t s_gbase[K2P_MAX_F];                                      // global position of each fine bin's run
    __shared__ uint32_t s_cnt[K2P_MAX_F];
    __shared__ uint32_t s_off[K2P_MAX_F + 1];
    __shared__ uint32_t s_warp[K2P_THREADS / 32];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (p.F + K2P_THREADS - 1) / K2P_THREADS;                  // fine bins scanned per thread (1 or 2)
    const uint32_t lane = lane_id(), warp = t >> 5;
    if (MATCH) { for (uint32_t i = t; i < p.F * NW; i += K2P_THREADS) s_wc[i] = 0; }

    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t b = k2p_bin_of_tile(p.tile_first, p.P1, tile);
        const uint64_t beg = p.bin_beg[b] + (uint64_t)(tile - p.tile_first[b]) * K2P_TILE;
        const uint64_t end = min(beg + (uint64_t)K2P_TILE, p.bin_end[b]);
        for (uint32_t i = t; i < p.F; i += K2P_THREADS) s_cnt[i] = 0;
        __syncthreads();

        // rank of every k-mer inside its fine bin of this tile: one shared-memory atomic per k-mer
        uint64_t key[K2P_PER_THREAD];
        uint32_t fr[K2P_PER_THREAD];                                             // rank << 10 | fine bin
#pragma unroll
        for (int j = 0; j < K2P_PER_THREAD; j++) {
            const uint64_t i = beg + (uint64_t)j * K2P_THREADS + t;
            key[j] = i < end ? p.kmers[i] : EMPTY_KEY;
        }
#pragma unroll
        for (int j = 0; j < K2P_PER_THREAD; j++) {
            if (!MATCH) {
                if (key[j] == EMPTY_KEY) { fr[j] = K2P_NOKEY; continue; }
                const uint32_t f = k2p_fine(mix64(key[j]), p.F, p.fine_shift);
                fr[j] = (atomicAdd(&s_cnt[f], 1u) << 10) | f;
            } else {   // (every lane takes part: the tile's ragged end is warp-divergent otherwise)
                const bool ok = key[j] != EMPTY_KEY;
                const uint32_t f = ok ? k2p_fine(mix64(key[j]), p.F, p.fine_shift) : K2P_NOKEY;
                const unsigned m = __match_any_sync(0xFFFFFFFFu, f);
                const int leader = __ffs(m) - 1;
                uint32_t base = 0;
                if (ok && (int)lane == leader) { base = s_wc[f * NW + warp]; s_wc[f * NW + warp] = base + __popc(m); }
                base = __shfl_sync(0xFFFFFFFFu, base, leader);
                fr[j] = ok ? (((base + __popc(m & ((1u << lane) - 1))) << 10) | f) : K2P_NOKEY;
                __syncwarp();
            }
        }
        __syncthreads();
        if (MATCH) {   // per fine bin: prefix over the warps' counts (each warp's base), total into s_cnt
            for (uint32_t f = t; f < p.F; f += K2P_THREADS) {
                uint32_t run = 0;
#pragma unroll
                for (uint32_t w = 0; w < NW; w++) { const uint32_t c = s_wc[f * NW + w]; s_wc[f * NW + w] = run; run += c; }
                s_cnt[f] = run;
            }
            __syncthreads();
        }

        // exclusive scan of the tile's fine counts + ONE global reservation per non-empty (tile, fine bin)
        {
            const uint32_t b0 = t * per;
            uint32_t sum = 0;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.F) sum += s_cnt[b0 + j];
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane_id() >= d) incl += n; }
            if (lane_id() == 31) s_warp[t >> 5] = incl;
            __syncthreads();
            if (t < 32) {
                uint32_t wv = t < K2P_THREADS / 32 ? s_warp[t] : 0, wi = wv;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, d); if ((int)lane_id() >= d) wi += n; }
                if (t < K2P_THREADS / 32) s_warp[t] = wi - wv;
            }
            __syncthreads();
            uint32_t run = s_warp[t >> 5] + incl - sum;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.F) {
                const uint32_t cn = s_cnt[b0 + j];
                s_off[b0 + j] = run;
                if (cn) {
                    const uint32_t gb = p.group ? p.group[b] : b;
                    const size_t g = (size_t)gb * p.F + b0 + j;
                    const uint32_t old = atomicAdd(&p.fine_cursor[g], cn);
                    if (!p.region_base) s_gbase[b0 + j] = p.fine_off[g] + old;
                    else if ((uint64_t)old + cn > p.region_cap[gb]) { *p.overflow = 1; s_gbase[b0 + j] = K2P_SKIP; }
                    else s_gbase[b0 + j] = p.region_base[gb] + (uint64_t)(b0 + j) * p.region_cap[gb] + old;
                }
                run += cn;
            }
        }
        __syncthreads();

        // counting sort into the staging area, then every run leaves as contiguous stores
#pragma unroll
        for (int j = 0; j < K2P_PER_THREAD; j++) {
            if (fr[j] == K2P_NOKEY) continue;
            const uint32_t f = fr[j] & 1023u, pos = s_off[f] + (fr[j] >> 10) + (MATCH ? s_wc[f * NW + warp] : 0u);
            s_stage[pos] = key[j]; s_sb[pos] = (uint16_t)f;
        }
        __syncthreads();
        const uint32_t total = (uint32_t)(end - beg);
        for (uint32_t j = t; j < total; j += K2P_THREADS) {
            const uint32_t f = s_sb[j];
            const uint64_t gbase = s_gbase[f];
            if (gbase != K2P_SKIP) p.out[gbase + (j - s_off[f])] = s_stage[j];
        }
        if (MATCH) { for (uint32_t i = t; i < p.F * NW; i += K2P_THREADS) s_wc[i] = 0; }   // (the staging loop is behind the barrier above)
        __syncthreads();     // the staging area and the counters are reused by the next tile
    }
}

// regions mode: [begin, end) of every fine bin from the region layout and the final cursors
__global__ void __launch_bounds__(256) k2p_region_ranges(const uint64_t *__restrict__ region_base, const uint32_t *__restrict__ region_cap,
                                                         const uint32_t *__restrict__ fine_cursor, uint32_t F, uint64_t NF, uint64_t *fine_beg,
                                                         uint64_t *fine_end) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < NF; g += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(g / F), f = (uint32_t)(g % F);
        const uint64_t beg = region_base[b] + (uint64_t)f * region_cap[b];
        fine_beg[g] = beg; fine_end[g] = beg + min(fine_cursor[g], region_cap[b]);
    }
}

static inline size_t k2p_scatter_smem_bytes(uint32_t F, bool match) {
    return (size_t)K2P_TILE * 8 + (size_t)K2P_TILE * 2 + (match ? (size_t)F * (K2P_THREADS / 32) * 4 : 0) + 128;
}

// ------------------------------------------------------------------------------------------------ counting in shared memory

struct K2CParams {
    const uint64_t *kmers;          // the stream ordered by fine bin
    const uint64_t *fine_beg, *fine_end;   // [n_fine] each (exact offsets: fine_end = fine_beg + 1 of one prefix array)
    uint32_t n_fine;
    uint32_t amin, amax;
    uint64_t *solid_kmers; uint32_t *solid_counts; uint64_t solid_cap;
    unsigned long long *n_solid, *n_distinct;
    unsigned long long *n_splits;   // statistics: how many times a (part of a) fine bin had to be split
    int *failed;                    // set when a fine bin still does not fit after K2C_MAX_LEVEL splits (host reports an error)
};

// Slot / split-bit hash of the table kernel.  The fine bin was chosen from the HIGH half of mix64(key), so anything that is not a
// function of those bits will do inside a bin.  mix64 costs two 64-bit multiplies (~20 instructions of an issue-bound kernel);
// mul32 is two 32-bit multiplies and a shift-xor.  Placement only: the (k-mer, count) multiset that comes out is the same.
__device__ __forceinline__ uint32_t k2c_hash_mul32(uint64_t key) {
    uint32_t h = (uint32_t)key * 0x9E3779B1u ^ (uint32_t)(key >> 32) * 0x85EBCA77u;
    return h ^ (h >> 15);
}
// (a template parameter of the kernel, not a run-time switch: a run-time one changed the register allocation of the default kernel)
template <bool MUL32> __device__ __forceinline__ uint32_t k2c_hash(uint64_t key) { return MUL32 ? k2c_hash_mul32(key) : (uint32_t)mix64(key); }

constexpr int K2C_MAX_LEVEL = 10;   // a fine bin is split into at most 2^10 parts by further hash bits

template <int SLOTS> constexpr int k2c_threads() { return SLOTS / 16; }
template <int SLOTS> static inline size_t k2c_smem_bytes() { return (size_t)SLOTS * 12 + 128; }

// One CTA per fine bin (grid-strided).  Table = SLOTS keys (u64, EMPTY = all ones: never a k-mer for k <= 31) + SLOTS counts.
// Every warp owns SLOTS/warps consecutive slots for clearing and emitting, so those phases need no barrier between them.
//
// Sizing is only an estimate (the number of DISTINCT k-mers of a bin is not known in advance), so the kernel adapts: new keys
// are tallied while inserting, and when the table passes 3/4 full the pass is abandoned BEFORE anything was emitted and the
// (part of the) bin is redone as two halves selected by one more hash bit, each half a full pass over the bin's k-mers (they
// are L2-warm).  Low-coverage input (every instance distinct) therefore costs one extra read per bin instead of failing.
//
// FLAT = true is a second form of the insert loop (same table, same result).  In the nested form every lane probes for ITS key
// until it is placed, so a warp waits for its slowest lane at every key: ncu shows 16.6 of 32 lanes active on average and
// 38 % of the warp-instructions issued for fewer than 4 lanes (the tail of the probe-length distribution).  The flat form does
// ONE probe step per iteration and a lane that has placed its key takes its next one immediately (keys are prefetched two
// ahead into registers), so lanes drift apart across keys instead of idling: the warp runs for max-over-lanes of the SUM of
// probe lengths instead of the sum of the per-key maxima.
template <int SLOTS, bool FLAT, bool MUL32 = false>
__global__ void __launch_bounds__(SLOTS / 16) k2c_count(K2CParams p) {
    constexpr int T = SLOTS / 16, NW = T / 32, PER_WARP = SLOTS / NW;   // PER_WARP = 512 slots
    constexpr uint32_t MASK = SLOTS - 1, FILL_LIMIT = SLOTS / 4 * 3;
    static_assert((SLOTS & (SLOTS - 1)) == 0 && SLOTS >= 1024 && T <= 1024, "SLOTS must be a power of two in [1024, 16384]");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_keys + SLOTS);
    __shared__ uint32_t s_wsolid[32];
    __shared__ unsigned long long s_base;
    __shared__ int s_over;
    __shared__ uint32_t s_fill;
    const uint32_t t = threadIdx.x, lane = lane_id(), warp = t >> 5;
    const uint32_t wb = warp * PER_WARP;
    unsigned long long my_distinct = 0, my_splits = 0;

    for (uint32_t bin = blockIdx.x; bin < p.n_fine; bin += gridDim.x) {
        const uint64_t beg = p.fine_beg[bin], end = p.fine_end[bin];
        if (beg == end) continue;                                    // (uniform over the block)
        // work stack of (level << 16 | part): the k-mers of this bin whose hash bits 16.. (level of them) equal `part`.
        // Every thread keeps the same copy: all decisions below are uniform over the block.
        uint32_t stk[K2C_MAX_LEVEL + 2];
        int sp = 0;
        stk[sp++] = 0;
        while (sp) {
            const uint32_t item = stk[--sp], level = item >> 16, part = item & 0xFFFFu, sel_mask = (1u << level) - 1u;
#pragma unroll 4
            for (uint32_t i = wb + lane; i < wb + PER_WARP; i += 32) { s_keys[i] = EMPTY_KEY; s_cnt[i] = 0; }
            if (t == 0) { s_over = 0; s_fill = 0; }
            __syncthreads();

            // ---- insert: load on the slot, CAS only when it looks empty, count with a shared-memory add
            if (!FLAT)
            for (uint64_t base = beg + warp * 32; base < end; base += (uint64_t)T * 4) {     // warp-uniform trip count
                uint64_t key[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { const uint64_t i = base + lane + (uint64_t)u * T; key[u] = i < end ? p.kmers[i] : EMPTY_KEY; }
                if (__any_sync(0xFFFFFFFFu, *(volatile int *)&s_over)) break;     // table (nearly) full: this pass is abandoned
                uint32_t fresh = 0;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (key[u] == EMPTY_KEY) continue;
                    const uint32_t h = k2c_hash<MUL32>(key[u]);
                    if (((h >> 16) & sel_mask) != part) continue;
                    uint32_t slot = h & MASK;
                    for (uint32_t probes = 0;;) {
                        uint64_t cur = ld_volatile_u64(&s_keys[slot]);
                        if (cur == EMPTY_KEY) {
                            cur = atomicCAS((unsigned long long *)&s_keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key[u]);
                            if (cur == EMPTY_KEY) { cur = key[u]; fresh++; }
                        }
                        if (cur == key[u]) { atomicAdd(&s_cnt[slot], 1u); break; }
                        slot = (slot + 1) & MASK;
                        if (++probes >= (uint32_t)SLOTS) { s_over = 1; break; }
                    }
                }
                fresh = __reduce_add_sync(0xFFFFFFFFu, fresh);
                if (lane == 0 && fresh && atomicAdd(&s_fill, fresh) + fresh > FILL_LIMIT) s_over = 1;
            }
            if (FLAT) {
                uint64_t nxt = beg + t;                                   // index of the next key to load; this thread owns beg + t + j*T
                uint64_t q0 = nxt < end ? p.kmers[nxt] : EMPTY_KEY; nxt += T;
                uint64_t q1 = nxt < end ? p.kmers[nxt] : EMPTY_KEY; nxt += T;
                uint64_t cur = EMPTY_KEY;
                uint32_t slot = 0, probes = 0, fresh = 0;
                bool have = false, drained = false;
                for (uint32_t iter = 0;; iter++) {
                    if (!have && !drained) {                              // take the next key of this thread that belongs to this pass
                        cur = q0; q0 = q1; q1 = nxt < end ? p.kmers[nxt] : EMPTY_KEY; nxt += T;
                        if (cur == EMPTY_KEY) drained = true;
                        else {
                            const uint32_t h = k2c_hash<MUL32>(cur);
                            if (((h >> 16) & sel_mask) == part) { slot = h & MASK; probes = 0; have = true; }
                        }
                    }
                    if (have) {                                           // one probe step
                        uint64_t c = ld_volatile_u64(&s_keys[slot]);
                        if (c == EMPTY_KEY) {
                            c = atomicCAS((unsigned long long *)&s_keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)cur);
                            if (c == EMPTY_KEY) { c = cur; fresh++; }
                        }
                        if (c == cur) { atomicAdd(&s_cnt[slot], 1u); have = false; }
                        else {
                            slot = (slot + 1) & MASK;
                            if (++probes >= (uint32_t)SLOTS) { s_over = 1; have = false; drained = true; }
                        }
                    }
                    if ((iter & 7u) == 7u) {                              // tally the new keys, look at the overflow flag (warp-uniform)
                        const uint32_t f = __reduce_add_sync(0xFFFFFFFFu, fresh);
                        fresh = 0;
                        if (lane == 0 && f && atomicAdd(&s_fill, f) + f > FILL_LIMIT) s_over = 1;
                        if (__any_sync(0xFFFFFFFFu, *(volatile int *)&s_over)) break;
                    }
                    if (!__any_sync(0xFFFFFFFFu, have || !drained)) break;
                }
                const uint32_t f = __reduce_add_sync(0xFFFFFFFFu, fresh);
                if (lane == 0 && f && atomicAdd(&s_fill, f) + f > FILL_LIMIT) s_over = 1;
            }
            __syncthreads();
            const int over = s_over;
            __syncthreads();                                         // everyone has read s_over before thread 0 may reset it
            if (over) {
                my_splits += t == 0;
                if (level >= (uint32_t)K2C_MAX_LEVEL) { if (t == 0) *p.failed = 1; break; }
                stk[sp++] = ((level + 1) << 16) | part;
                stk[sp++] = ((level + 1) << 16) | part | (1u << level);
                continue;
            }

            // ---- emit: count the solid entries, reserve the output range with ONE global atomic, write ballot-ranked
            uint32_t my_solid = 0;
#pragma unroll 4
            for (uint32_t i = wb + lane; i < wb + PER_WARP; i += 32) {
                const bool occ = s_keys[i] != EMPTY_KEY;
                const uint32_t c = s_cnt[i];
                my_distinct += occ;
                my_solid += occ && c >= p.amin && c <= p.amax;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) my_solid += __shfl_xor_sync(0xFFFFFFFFu, my_solid, d);
            if (lane == 0) s_wsolid[warp] = my_solid;
            __syncthreads();
            if (t == 0) {
                uint32_t tot = 0;
                for (int w = 0; w < NW; w++) { uint32_t v = s_wsolid[w]; s_wsolid[w] = tot; tot += v; }
                s_base = tot ? atomicAdd(p.n_solid, (unsigned long long)tot) : 0ULL;
            }
            __syncthreads();
            uint64_t out = s_base + s_wsolid[warp];
            if (my_solid) {                                          // (uniform over the warp after the reduction)
                for (uint32_t i = wb + lane; i < wb + PER_WARP; i += 32) {
                    const uint64_t key = s_keys[i];
                    const uint32_t c = s_cnt[i];
                    const bool solid = key != EMPTY_KEY && c >= p.amin && c <= p.amax;
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, solid);
                    if (solid) {
                        const uint64_t pos = out + __popc(m & ((1u << lane) - 1));
                        if (pos < p.solid_cap) { p.solid_kmers[pos] = key; p.solid_counts[pos] = c; }
                    }
                    out += __popc(m);
                }
            }
            // no barrier here: the next pass's clear touches only this warp's own slots, and s_wsolid / s_base / s_over / s_fill
            // are next written after at least two more barriers
        }
    }
    block_add(p.n_distinct, my_distinct);
    block_add(p.n_splits, my_splits);
}
