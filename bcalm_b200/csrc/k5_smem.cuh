// k5_smem.cuh -- K5 (the junction join on canonical (k-1)-mers, SURVEY a11: ograph.cpp:228-318, LinkTigs.cpp:268-457) with
// the join tables in SHARED memory, the same way k2_smem.cuh does the counting.
//
// The L2 path (k5_insert / k5_resolve in kernels.cuh) runs one L2-resident table per coarse bin: ~4 L2 operations per end
// record and two short launches per bin.  Here the end records K3 wrote into P1 coarse bins are split once more into F fine
// bins per coarse bin (k5p_hist, scan, k5p_scatter: a tile-local counting sort like k2p_scatter, records are 16 bytes), and
// k5c_join gives every fine bin to one CTA: insert into a shared-memory table {key, counters, <=4 IN ids, <=4 OUT ids},
// then walk the table and decide every junction exactly as k5_resolve does.  A fine bin with more distinct keys than the
// table takes is split on further hash bits inside the kernel (as in k2c_count).
#pragma once
#include "kernels.cuh"

constexpr int K5P_THREADS = 256;
constexpr int K5P_PER_THREAD = 16;
constexpr int K5P_TILE = K5P_THREADS * K5P_PER_THREAD;    // 4096 records per tile
constexpr int K5P_MAX_F = 1024;
#define K5P_NOKEY 0xFFFFFFFFu

// Fine bin of a record inside its coarse bin b.  g = floor(h_hi * P1*F / 2^32) is a bin index over ALL fine bins, and
// floor(g / F) == floor(h_hi * P1 / 2^32) == hash_to_bin(h, P1), the coarse bin K3 used -- so g - b*F is in [0, F).
__device__ __forceinline__ uint32_t k5p_fine(uint64_t h, uint32_t NF, uint32_t F, uint32_t b) {
    return (uint32_t)(((h >> 32) * (uint64_t)NF) >> 32) - b * F;
}

struct K5PParams {
    const uint64_t *key, *val;          // K3 output (SoA), coarse bin b = [bin_beg[b], bin_end[b])
    const uint64_t *bin_beg, *bin_end;  // [P1] each
    const uint32_t *tile_first;         // [P1 + 1]
    const uint32_t *group;              // [P1] or null.  Sharded mode: the coarse "bins" are the received runs (one per local bin and source
                                        // GPU); group[b] = the run's local bin = hash_to_bin(h, G), whose fine bins its records share
    uint32_t P1, F, n_tiles, G;         // G = number of groups (= P1 without `group`)
    uint32_t *fine_cnt;                 // [G * F]
    const uint64_t *fine_off;           // [P1 * F + 1]
    uint32_t *fine_cursor;              // [P1 * F]
    uint64_t *out_key, *out_val;        // records re-ordered by (coarse, fine) bin
};

__device__ __forceinline__ uint32_t k5p_bin_of_tile(const uint32_t *__restrict__ tile_first, uint32_t P1, uint32_t tile) {
    uint32_t lo = 0, hi = P1;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (tile_first[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(K5P_THREADS) k5p_hist(K5PParams p) {
    __shared__ uint32_t s_cnt[K5P_MAX_F];
    const uint32_t t = threadIdx.x, NF = p.G * p.F;
    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t b = k5p_bin_of_tile(p.tile_first, p.P1, tile);
        const uint32_t gb = p.group ? p.group[b] : b;
        const uint64_t beg = p.bin_beg[b] + (uint64_t)(tile - p.tile_first[b]) * K5P_TILE;
        const uint64_t end = min(beg + (uint64_t)K5P_TILE, p.bin_end[b]);
        for (uint32_t i = t; i < p.F; i += K5P_THREADS) s_cnt[i] = 0;
        __syncthreads();
        uint64_t key[K5P_PER_THREAD];
#pragma unroll
        for (int j = 0; j < K5P_PER_THREAD; j++) {
            const uint64_t i = beg + (uint64_t)j * K5P_THREADS + t;
            key[j] = i < end ? p.key[i] : EMPTY_KEY;
        }
#pragma unroll
        for (int j = 0; j < K5P_PER_THREAD; j++)
            if (key[j] != EMPTY_KEY) atomicAdd(&s_cnt[k5p_fine(mix64(key[j]), NF, p.F, gb)], 1u);
        __syncthreads();
        for (uint32_t i = t; i < p.F; i += K5P_THREADS) {     // (same thread zeroes counter i for the next tile)
            const uint32_t c = s_cnt[i];
            if (c) atomicAdd(&p.fine_cnt[(size_t)gb * p.F + i], c);
        }
    }
}

__global__ void __launch_bounds__(K5P_THREADS, 2) k5p_scatter(K5PParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *s_key = reinterpret_cast<uint64_t *>(smem_raw);                    // [K5P_TILE]
    uint64_t *s_val = s_key + K5P_TILE;                                          // [K5P_TILE]
    uint16_t *s_sb = reinterpret_cast<uint16_t *>(s_val + K5P_TILE);             // [K5P_TILE]
    __shared__ uint64_t s_gbase[K5P_MAX_F];
    __shared__ uint32_t s_cnt[K5P_MAX_F];
    __shared__ uint32_t s_off[K5P_MAX_F + 1];
    __shared__ uint32_t s_warp[K5P_THREADS / 32];
    const uint32_t t = threadIdx.x, NF = p.G * p.F;
    const uint32_t per = (p.F + K5P_THREADS - 1) / K5P_THREADS;                  // fine bins scanned per thread (1..4)

    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t b = k5p_bin_of_tile(p.tile_first, p.P1, tile);
        const uint32_t gb = p.group ? p.group[b] : b;
        const uint64_t beg = p.bin_beg[b] + (uint64_t)(tile - p.tile_first[b]) * K5P_TILE;
        const uint64_t end = min(beg + (uint64_t)K5P_TILE, p.bin_end[b]);
        for (uint32_t i = t; i < p.F; i += K5P_THREADS) s_cnt[i] = 0;
        __syncthreads();

        uint64_t key[K5P_PER_THREAD];
        uint32_t fr[K5P_PER_THREAD];                                             // rank << 10 | fine bin
#pragma unroll
        for (int j = 0; j < K5P_PER_THREAD; j++) {
            const uint64_t i = beg + (uint64_t)j * K5P_THREADS + t;
            key[j] = i < end ? p.key[i] : EMPTY_KEY;
        }
#pragma unroll
        for (int j = 0; j < K5P_PER_THREAD; j++) {
            if (key[j] == EMPTY_KEY) { fr[j] = K5P_NOKEY; continue; }
            const uint32_t f = k5p_fine(mix64(key[j]), NF, p.F, gb);
            fr[j] = (atomicAdd(&s_cnt[f], 1u) << 10) | f;
        }
        __syncthreads();
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
                uint32_t wv = t < K5P_THREADS / 32 ? s_warp[t] : 0, wi = wv;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, d); if ((int)lane_id() >= d) wi += n; }
                if (t < K5P_THREADS / 32) s_warp[t] = wi - wv;
            }
            __syncthreads();
            uint32_t run = s_warp[t >> 5] + incl - sum;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.F) {
                const uint32_t cn = s_cnt[b0 + j];
                s_off[b0 + j] = run;
                if (cn) {
                    const size_t g = (size_t)gb * p.F + b0 + j;
                    s_gbase[b0 + j] = p.fine_off[g] + atomicAdd(&p.fine_cursor[g], cn);
                }
                run += cn;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < K5P_PER_THREAD; j++) {
            if (fr[j] == K5P_NOKEY) continue;
            const uint64_t i = beg + (uint64_t)j * K5P_THREADS + t;
            const uint32_t f = fr[j] & 1023u, pos = s_off[f] + (fr[j] >> 10);
            s_key[pos] = key[j]; s_val[pos] = p.val[i]; s_sb[pos] = (uint16_t)f;
        }
        __syncthreads();
        const uint32_t total = (uint32_t)(end - beg);
        for (uint32_t j = t; j < total; j += K5P_THREADS) {
            const uint32_t f = s_sb[j];
            const uint64_t g = s_gbase[f] + (j - s_off[f]);
            p.out_key[g] = s_key[j]; p.out_val[g] = s_val[j];
        }
        __syncthreads();
    }
}

static inline size_t k5p_scatter_smem_bytes() { return (size_t)K5P_TILE * 18 + 128; }

// ------------------------------------------------------------------------------------------------ the join in shared memory

struct K5CParams {
    const uint64_t *key, *val;          // records ordered by fine bin
    const uint64_t *fine_off;           // [n_fine + 1]
    uint32_t n_fine;
    uint32_t *link;                     // [2S] compaction links between k-mer ends (pre-set to NOLINK)
    uint64_t *edges; uint64_t edge_cap; // unitig-level edges (from end << 32 | to end)
    unsigned long long *n_edges, *n_compact, *n_splits;
    int *failed;
};

constexpr int K5C_THREADS = 256;
constexpr int K5C_MAX_LEVEL = 10;
template <int SLOTS> static inline size_t k5c_smem_bytes() { return (size_t)SLOTS * (8 + 4 + 32) + 128; }

template <int SLOTS>
__global__ void __launch_bounds__(K5C_THREADS) k5c_join(K5CParams p) {
    constexpr int T = K5C_THREADS, NW = T / 32, PER_WARP = SLOTS / NW;
    constexpr uint32_t MASK = SLOTS - 1, FILL_LIMIT = SLOTS / 4 * 3;
    static_assert((SLOTS & (SLOTS - 1)) == 0 && SLOTS >= 1024, "SLOTS must be a power of two");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(smem_raw);                   // [SLOTS]
    uint32_t *s_ids = reinterpret_cast<uint32_t *>(s_keys + SLOTS);              // [SLOTS * 8]  0..3 IN, 4..7 OUT
    uint32_t *s_cnt = s_ids + (size_t)SLOTS * 8;                                 // [SLOTS] IN count | OUT count << 8 | palindrome << 31
    __shared__ int s_over;
    __shared__ uint32_t s_fill;
    const uint32_t t = threadIdx.x, lane = lane_id(), warp = t >> 5;
    const uint32_t wb = warp * PER_WARP;
    unsigned long long compact = 0, my_splits = 0;

    for (uint32_t bin = blockIdx.x; bin < p.n_fine; bin += gridDim.x) {
        const uint64_t beg = p.fine_off[bin], end = p.fine_off[bin + 1];
        if (beg == end) continue;
        uint32_t stk[K5C_MAX_LEVEL + 2];      // (level << 16 | part), identical in every thread: see k2c_count
        int sp = 0;
        stk[sp++] = 0;
        while (sp) {
            const uint32_t item = stk[--sp], level = item >> 16, part = item & 0xFFFFu, sel_mask = (1u << level) - 1u;
            for (uint32_t i = wb + lane; i < wb + PER_WARP; i += 32) { s_keys[i] = EMPTY_KEY; s_cnt[i] = 0; }
            if (t == 0) { s_over = 0; s_fill = 0; }
            __syncthreads();

            for (uint64_t base = beg + warp * 32; base < end; base += (uint64_t)T * 2) {     // warp-uniform trip count
                uint64_t key[2], val[2];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const uint64_t i = base + lane + (uint64_t)u * T;
                    key[u] = i < end ? p.key[i] : EMPTY_KEY;
                    val[u] = i < end ? p.val[i] : 0;
                }
                if (__any_sync(0xFFFFFFFFu, *(volatile int *)&s_over)) break;
                uint32_t fresh = 0;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    if (key[u] == EMPTY_KEY || (val[u] & REC_SKIP)) continue;      // (REC_SKIP: second end of a self-complementary k-mer)
                    const uint64_t h = mix64(key[u]);
                    if ((((uint32_t)h >> 16) & sel_mask) != part) continue;
                    uint32_t slot = (uint32_t)h & MASK;
                    for (uint32_t probes = 0;;) {
                        uint64_t cur = ld_volatile_u64(&s_keys[slot]);
                        if (cur == EMPTY_KEY) {
                            cur = atomicCAS((unsigned long long *)&s_keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key[u]);
                            if (cur == EMPTY_KEY) { cur = key[u]; fresh++; }
                        }
                        if (cur == key[u]) {
                            const bool out = (val[u] & REC_SIDE_OUT) != 0, pal = (val[u] & REC_PALIN) != 0;
                            const bool use_out = out && !pal;          // palindromic ends all go to the IN list ("all pairs")
                            const uint32_t old = atomicAdd(&s_cnt[slot], use_out ? 0x100u : 0x1u);
                            if (pal) atomicOr(&s_cnt[slot], 0x80000000u);
                            const uint32_t pos = use_out ? ((old >> 8) & 0xFF) : (old & 0xFF);
                            if (pos < 4) {
                                s_ids[(size_t)slot * 8 + (use_out ? 4 : 0) + pos] = (uint32_t)val[u];
                                if (val[u] & REC_BOTH) atomicOr(&s_cnt[slot], 1u << ((use_out ? JOIN_BOTH_SHIFT_OUT : JOIN_BOTH_SHIFT_IN) + pos));
                            }
                            break;
                        }
                        slot = (slot + 1) & MASK;
                        if (++probes >= (uint32_t)SLOTS) { s_over = 1; break; }
                    }
                }
                fresh = __reduce_add_sync(0xFFFFFFFFu, fresh);
                if (lane == 0 && fresh && atomicAdd(&s_fill, fresh) + fresh > FILL_LIMIT) s_over = 1;
            }
            __syncthreads();
            const int over = s_over;
            __syncthreads();
            if (over) {
                my_splits += t == 0;
                if (level >= (uint32_t)K5C_MAX_LEVEL) { if (t == 0) *p.failed = 1; break; }
                stk[sp++] = ((level + 1) << 16) | part;
                stk[sp++] = ((level + 1) << 16) | part | (1u << level);
                continue;
            }

            // ---- resolve (the decision rules of k5_resolve); every warp walks its own slots
            for (uint32_t i = wb + lane; i < wb + PER_WARP; i += 32) {
                if (s_keys[i] == EMPTY_KEY) continue;
                const uint32_t c = s_cnt[i];
                uint32_t nin = c & 0xFF, nout = (c >> 8) & 0xFF; const bool pal = (c >> 31) != 0;
                if (c == 0x101u) {                      // the common junction: one IN end, one OUT end, no flags -- no id arrays needed
                    const uint32_t a = s_ids[(size_t)i * 8], b = s_ids[(size_t)i * 8 + 4];
                    if ((a >> 1) != (b >> 1)) { p.link[a] = b; p.link[b] = a; compact++; continue; }
                }
                if (nin > 4) nin = 4;
                if (nout > 4) nout = 4;
                uint32_t in[5], out[5];
                {
                    const uint4 vin = *reinterpret_cast<const uint4 *>(&s_ids[(size_t)i * 8]);
                    const uint4 vout = *reinterpret_cast<const uint4 *>(&s_ids[(size_t)i * 8 + 4]);
                    const uint32_t sin[4] = {vin.x, vin.y, vin.z, vin.w}, sout[4] = {vout.x, vout.y, vout.z, vout.w};
                    nin = join_expand_side(sin, nin, (c >> JOIN_BOTH_SHIFT_IN) & 0xFu, in);
                    nout = join_expand_side(sout, nout, (c >> JOIN_BOTH_SHIFT_OUT) & 0xFu, out);
                }
                if (!pal && nin == 1 && nout == 1 && (in[0] >> 1) != (out[0] >> 1)) {
                    p.link[in[0]] = out[0]; p.link[out[0]] = in[0];
                    compact++;
                    continue;
                }
                const uint32_t ne = pal ? nin * nin : 2 * nin * nout;
                if (ne == 0) continue;
                unsigned long long base = atomicAdd(p.n_edges, (unsigned long long)ne);
                if (base + ne > p.edge_cap) continue;   // counted; the host retries with a larger buffer
                if (pal) {
                    for (uint32_t a = 0; a < nin; a++) for (uint32_t b2 = 0; b2 < nin; b2++) p.edges[base++] = ((uint64_t)in[a] << 32) | in[b2];
                } else {
                    for (uint32_t a = 0; a < nin; a++) for (uint32_t b2 = 0; b2 < nout; b2++) {
                        p.edges[base++] = ((uint64_t)in[a] << 32) | out[b2];
                        p.edges[base++] = ((uint64_t)out[b2] << 32) | in[a];
                    }
                }
            }
            // (no barrier: the next pass clears only this warp's own slots; inserts into them start after the next barrier)
        }
    }
    block_add(p.n_compact, compact);
    block_add(p.n_splits, my_splits);
}
