// kernels.cuh -- the hand-written sm_100a kernels of the reads -> unitigs path (k <= 31).
//
// Stage map (DESIGN.md gives the byte counts and rooflines):
//   K1  k1_extract<HIST|SCATTER>  reads -> canonical k-mers, radix-partitioned by hash           (a1,a2,a5; Model.hpp:857-884)
//   K2  k2_insert / k2_emit       per-bin counting in an L2-resident table + solidity filter     (a6,a7; PartitionsCommand.cpp:1600-1805)
//   K3  k3_ends<HIST|SCATTER>     solid k-mer -> two (k-1)-mer end records, partitioned          (a8,a9,a10; bcalm_algo.cpp:182-223, ograph.cpp:384-434)
//   K5  k5_insert / k5_resolve    per-bin join on the canonical (k-1)-mer: compaction links
//                                 or unitig-level edges                                          (a11; ograph.cpp:228-318, LinkTigs.cpp:268-457)
//   K6  k6_* / k8_*               list ranking over the links, cycle cutting, unitig numbering,
//                                 spelling + abundance sums                                      (a12,a13,a15,a16; ograph.cpp:126-213, bglue_algo.cpp:262-479)
//   K9  k9_translate              k-mer-end edges -> L: records                                  (a17; LinkTigs.cpp:365-369,428-432)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kmer.cuh"

#define NOLINK 0xFFFFFFFFu
#define EMPTY_KEY 0xFFFFFFFFFFFFFFFFULL

// ------------------------------------------------------------------------------------------------ small helpers

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// warp-aggregated append: every active lane with pred gets a distinct slot from *cursor (one atomic per warp)
__device__ __forceinline__ uint64_t warp_append(unsigned long long *cursor, bool pred) {
    unsigned m = __ballot_sync(__activemask(), pred);
    if (!pred) return 0;
    int leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if ((int)lane_id() == leader) base = atomicAdd(cursor, (unsigned long long)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & ((1u << lane_id()) - 1));
}

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p) { return *(const volatile uint64_t *)p; }

// Programmatic dependent launch: kernels of the short per-bin loops (K2, K5, K6) are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so their CTAs are scheduled while the previous kernel drains; they must
// wait here before touching anything the predecessor wrote.  A no-op for a normal launch.
__device__ __forceinline__ void pdl_wait() {
#ifndef CDBG_EMU_ACTIVE   // (the CPU emulation build of the tests runs launches back to back)
    asm volatile("griddepcontrol.wait;" ::: "memory");               // predecessor grid complete and its writes visible
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let the successor's CTAs take SM slots as ours retire
#endif
}

// Sum `v` over the block and add it to *dst with ONE atomic per block: a same-address global atomic costs ~4 ns, so
// per-warp accumulation (thousands of warps per launch) would dominate these short kernels.  All threads must call.
__device__ __forceinline__ void block_add(unsigned long long *dst, unsigned long long v) {
    __shared__ unsigned long long s_part[32];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    if (lane_id() == 0) s_part[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long w = threadIdx.x < (blockDim.x + 31) / 32 ? s_part[threadIdx.x] : 0;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(0xFFFFFFFFu, w, d);
        if (threadIdx.x == 0 && w) atomicAdd(dst, w);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ K1: extract

constexpr int K1_THREADS = 256;
constexpr int K1_PER_THREAD = 32;                       // positions per thread
constexpr int K1_TILE = K1_THREADS * K1_PER_THREAD;     // 8192 positions per tile
constexpr int K1_MAX_BINS = 2048;

// pack 4 ASCII bytes (little-endian word, first char = lowest byte) into 8 bits, first char most significant,
// plus a 4-bit validity mask (bit 3 = first char)
__device__ __forceinline__ void pack4(uint32_t w, uint32_t &codes, uint32_t &valid) {
    uint32_t c = (w >> 1) & 0x03030303u;
    codes = (c * 0x40100401u) >> 24;
    uint32_t u = w & 0xDFDFDFDFu;                       // fold lower case
    uint32_t v = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
    valid = ((v & 0x01010101u) * 0x08040201u) >> 24 & 0xFu;
}

// 32 bytes (two uint4) -> 64-bit packed codes (first base in bits 63:62) + 32-bit validity (first base in bit 31)
__device__ __forceinline__ void pack32(const uint4 a, const uint4 b, uint64_t &pk, uint32_t &vm) {
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    pk = 0; vm = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t c, v;
        pack4(w[i], c, v);
        pk = (pk << 8) | c;
        vm = (vm << 4) | v;
    }
}

// 32 bytes at `off` (a multiple of 32; `bytes` is 16-byte aligned); positions >= n read as a break
__device__ __forceinline__ void load_pack32(const uint8_t *bytes, uint64_t n, uint64_t off, uint64_t &pk, uint32_t &vm) {
    uint4 a, b;
    if (off + 32 <= n) {
        const uint4 *src = reinterpret_cast<const uint4 *>(bytes + off);
        a = __ldg(src); b = __ldg(src + 1);
    } else {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t x = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint64_t q = off + 4 * i + j;
                x |= (uint32_t)(q < n ? bytes[q] : (uint8_t)'\n') << (8 * j);
            }
            w[i] = x;
        }
        a = make_uint4(w[0], w[1], w[2], w[3]); b = make_uint4(w[4], w[5], w[6], w[7]);
    }
    pack32(a, b, pk, vm);
}

// bit (63-i) of the result is set iff bases i .. i+k-1 are all valid (V has base j at bit 63-j)
__device__ __forceinline__ uint64_t kmer_valid_mask(uint64_t V, int k) {
    uint64_t W1 = V, W2 = W1 & (W1 << 1), W4 = W2 & (W2 << 2), W8 = W4 & (W4 << 4), W16 = W8 & (W8 << 8);
    uint64_t r = ~0ULL;
    int off = 0;
    if (k & 16) { r &= W16 << off; off += 16; }
    if (k & 8)  { r &= W8 << off;  off += 8; }
    if (k & 4)  { r &= W4 << off;  off += 4; }
    if (k & 2)  { r &= W2 << off;  off += 2; }
    if (k & 1)  { r &= W1 << off; }
    return r;
}

struct K1Params {
    const uint8_t *bytes;      // 16-byte aligned; reads past n_bytes are treated as breaks
    uint64_t n_bytes;
    uint64_t n_tiles;
    int k;
    uint32_t P;                // number of bins
    unsigned long long *bin_count;   // [P]  (HIST: out)
    const uint64_t *bin_off;         // [P]  (SCATTER: in, exclusive offsets)
    unsigned long long *bin_cursor;  // [P]  (SCATTER: running cursors, zero-initialised)
    uint64_t *out_kmers;             // (SCATTER) partitioned canonical k-mers
};

// MODE 0: histogram of bins.  MODE 1: scatter with a tile-local counting sort so that each bin's k-mers of one tile
// leave the SM as one contiguous run.
template <int MODE>
__global__ void __launch_bounds__(K1_THREADS) k1_extract(K1Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // 8-byte arrays first, then the 4-byte ones
    uint64_t *s_pk = reinterpret_cast<uint64_t *>(smem_raw);                 // [THREADS + 2] packed bases, 32 per word
    uint64_t *s_gbase = s_pk + K1_THREADS + 2;                               // [P]       MODE 1: global base of each bin's run
    uint64_t *s_stage = s_gbase + (MODE == 1 ? p.P : 0);                     // [K1_TILE] MODE 1: tile sorted by bin
    uint32_t *s_vm = reinterpret_cast<uint32_t *>(s_stage + (MODE == 1 ? K1_TILE : 0));  // [THREADS + 2] validity bits
    uint32_t *s_cnt = s_vm + K1_THREADS + 2;                                 // [P]   tile counts (MODE 1) / CTA histogram (MODE 0)
    uint32_t *s_off = s_cnt + p.P;                                           // [P+1] MODE 1: exclusive offsets
    uint32_t *s_cur = s_off + p.P + 1;                                       // [P]   MODE 1: running cursors
    __shared__ uint32_t s_warp[K1_THREADS / 32];

    const int t = threadIdx.x;
    const int k = p.k;
    const uint64_t kmask = kmer_mask(k);

    if (MODE == 0) {
        for (uint32_t i = t; i < p.P; i += K1_THREADS) s_cnt[i] = 0;
        __syncthreads();
    }

    for (uint64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint64_t tile_start = tile * (uint64_t)K1_TILE;
        // ---- phase A: ASCII -> 2-bit, coalesced 32 B per thread
        {
            uint64_t pk; uint32_t vm;
            load_pack32(p.bytes, p.n_bytes, tile_start + (uint64_t)t * 32, pk, vm);
            s_pk[t] = pk; s_vm[t] = vm;
            if (t == 0) {  // halo: the next 32 bases
                load_pack32(p.bytes, p.n_bytes, tile_start + K1_TILE, pk, vm);
                s_pk[K1_THREADS] = pk; s_vm[K1_THREADS] = vm;
            }
        }
        if (MODE == 1) {
            for (uint32_t i = t; i < p.P; i += K1_THREADS) s_cnt[i] = 0;
        }
        __syncthreads();

        // ---- phase B: rolling canonical k-mers of 32 consecutive positions
        const uint64_t hi0 = s_pk[t], lo0 = s_pk[t + 1];
        const uint64_t V = ((uint64_t)s_vm[t] << 32) | s_vm[t + 1];
        const uint64_t KV = kmer_valid_mask(V, k);
        const uint32_t valid32 = (uint32_t)(KV >> 32);   // bit (31-i): k-mer at position i valid

        auto sweep = [&](auto &&emit) {
            uint64_t hi = hi0, lo = lo0;
            uint64_t fwd = hi >> (64 - 2 * k);
            uint64_t rc = revcomp64(fwd, k);
#pragma unroll 4
            for (int i = 0; i < K1_PER_THREAD; i++) {
                if (i > 0) {
                    hi = (hi << 2) | (lo >> 62); lo <<= 2;
                    fwd = hi >> (64 - 2 * k);
                    rc = (rc >> 2) | (((fwd & 3) ^ 2) << (2 * (k - 1)));
                }
                if ((valid32 >> (31 - i)) & 1) {
                    uint64_t c = fwd < rc ? fwd : rc;
                    emit(c);
                }
            }
        };
        (void)kmask;

        if (MODE == 0) {
            sweep([&](uint64_t c) { atomicAdd(&s_cnt[hash_to_bin(mix64(c), p.P)], 1u); });
            __syncthreads();   // s_pk reuse in the next tile
        } else {
            sweep([&](uint64_t c) { atomicAdd(&s_cnt[hash_to_bin(mix64(c), p.P)], 1u); });
            __syncthreads();
            // exclusive scan of s_cnt (P <= 2048: 256 threads x up to 8 bins each) + global reservation per bin
            {
                const uint32_t per = (p.P + K1_THREADS - 1) / K1_THREADS;
                uint32_t b0 = t * per, sum = 0;
                for (uint32_t j = 0; j < per; j++) if (b0 + j < p.P) sum += s_cnt[b0 + j];
                // block scan of `sum` via warp shuffles
                uint32_t incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane_id() >= d) incl += n; }
                if (lane_id() == 31) s_warp[t >> 5] = incl;
                __syncthreads();
                if (t < 32) {
                    uint32_t w = t < K1_THREADS / 32 ? s_warp[t] : 0, wi = w;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, d); if ((int)lane_id() >= d) wi += n; }
                    if (t < K1_THREADS / 32) s_warp[t] = wi - w;
                }
                __syncthreads();
                uint32_t run = s_warp[t >> 5] + incl - sum;
                for (uint32_t j = 0; j < per; j++) if (b0 + j < p.P) {
                    uint32_t c = s_cnt[b0 + j];
                    s_off[b0 + j] = run; s_cur[b0 + j] = run;
                    if (c) s_gbase[b0 + j] = p.bin_off[b0 + j] + atomicAdd(&p.bin_cursor[b0 + j], (unsigned long long)c);
                    run += c;
                }
                if (t == K1_THREADS - 1) s_off[p.P] = run;
            }
            __syncthreads();
            sweep([&](uint64_t c) {
                uint32_t pos = atomicAdd(&s_cur[hash_to_bin(mix64(c), p.P)], 1u);
                s_stage[pos] = c;
            });
            __syncthreads();
            const uint32_t total = s_off[p.P];
            for (uint32_t j = t; j < total; j += K1_THREADS) {
                uint64_t c = s_stage[j];
                uint32_t b = hash_to_bin(mix64(c), p.P);
                p.out_kmers[s_gbase[b] + (j - s_off[b])] = c;
            }
            __syncthreads();
        }
    }

    if (MODE == 0) {
        __syncthreads();
        for (uint32_t i = t; i < p.P; i += K1_THREADS) {
            uint32_t c = s_cnt[i];
            if (c) atomicAdd(&p.bin_count[i], (unsigned long long)c);
        }
    }
}

static inline size_t k1_smem_bytes(int mode, uint32_t P) {
    size_t s = (size_t)(K1_THREADS + 2) * 8 + (size_t)(K1_THREADS + 2) * 4 + (size_t)P * 4;
    if (mode == 1) s += (size_t)P * 8 + (size_t)K1_TILE * 8 + (size_t)(P + 1) * 4 + (size_t)P * 4;
    return s + 16;
}

// ------------------------------------------------------------------------------------------------ K2: count

// One bin at a time: the table (keys + counts) is sized to the bin (1.5 slots per instance, any capacity: slot = mulhi(hash, cap))
// and stays resident in the 126 MB L2.
__device__ __forceinline__ uint32_t table_slot(uint64_t key, uint32_t cap) { return (uint32_t)(((mix64(key) & 0xFFFFFFFFULL) * cap) >> 32); }

// (Tried: 4 k-mers per thread with their first probes issued together -- 2.16 -> 2.75 ms on C2.  The kernel is bound by L2
// operation throughput, not by per-thread latency, so more requests in flight only add contention.)
__global__ void __launch_bounds__(256) k2_insert(const uint64_t *__restrict__ kmers, uint64_t n, uint64_t *keys, uint32_t *counts,
                                                 uint32_t cap, int *overflow) {
    pdl_wait();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = kmers[i];
        uint32_t slot = table_slot(key, cap);
        uint32_t probes = 0;
        for (;;) {
            // load first, CAS only on an empty slot: probing with the CAS itself (2.0 instead of 2.3 L2 operations per instance)
            // measured slower (2.15 -> 2.23 ms): an L2 atomic costs more than the load it saves
            uint64_t cur = ld_volatile_u64(&keys[slot]);
            if (cur == EMPTY_KEY) {
                cur = atomicCAS((unsigned long long *)&keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
                if (cur == EMPTY_KEY) cur = key;
            }
            if (cur == key) { atomicAdd(&counts[slot], 1u); break; }
            slot = slot + 1 == cap ? 0 : slot + 1;
            if (++probes > min(cap, 8192u)) { *overflow = 1; break; }
        }
    }
}

// scan the table: emit solid (k-mer, count), count distinct, and reset every touched slot for the next bin.
// Each warp owns a contiguous run of slots.  Pass 1 counts the solid entries of the run (loads batched K2_EMIT_ILP deep),
// the block reserves its output range with ONE global atomic, pass 2 re-reads the run (now in L2), writes the solid pairs
// at warp-private offsets computed with ballots, and resets the slots.  Two barriers per block, no per-entry atomics.
constexpr int K2_EMIT_ILP = 4;
constexpr int K2_EMIT_THREADS = 256;
__global__ void __launch_bounds__(K2_EMIT_THREADS) k2_emit(uint64_t *keys, uint32_t *counts, uint32_t cap, uint32_t amin, uint32_t amax,
                                                           uint64_t *solid_kmers, uint32_t *solid_counts, uint64_t solid_cap,
                                                           unsigned long long *n_solid, unsigned long long *n_distinct) {
    pdl_wait();
    __shared__ uint32_t s_wsolid[K2_EMIT_THREADS / 32];
    __shared__ unsigned long long s_base;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t n_warps = gridDim.x * (K2_EMIT_THREADS / 32);
    const uint32_t per_warp = ((cap + n_warps - 1) / n_warps + 31u) & ~31u;
    const uint32_t gw = blockIdx.x * (K2_EMIT_THREADS / 32) + warp;
    const uint32_t wb = (uint32_t)min((uint64_t)cap, (uint64_t)gw * per_warp), we = (uint32_t)min((uint64_t)cap, (uint64_t)wb + per_warp);

    uint32_t my_solid = 0, my_distinct = 0;
    for (uint32_t base = wb; base < we; base += 32 * K2_EMIT_ILP) {
        uint64_t key[K2_EMIT_ILP]; uint32_t c[K2_EMIT_ILP];
#pragma unroll
        for (int j = 0; j < K2_EMIT_ILP; j++) { uint32_t i = base + j * 32 + lane; key[j] = i < we ? keys[i] : EMPTY_KEY; }
#pragma unroll
        for (int j = 0; j < K2_EMIT_ILP; j++) { uint32_t i = base + j * 32 + lane; c[j] = key[j] != EMPTY_KEY ? counts[i] : 0; }
#pragma unroll
        for (int j = 0; j < K2_EMIT_ILP; j++) {
            my_distinct += key[j] != EMPTY_KEY;
            my_solid += (key[j] != EMPTY_KEY) && c[j] >= amin && c[j] <= amax;
        }
    }
    uint32_t wsolid = my_solid;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) wsolid += __shfl_xor_sync(0xFFFFFFFFu, wsolid, d);
    if (lane == 0) s_wsolid[warp] = wsolid;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int w = 0; w < K2_EMIT_THREADS / 32; w++) { uint32_t v = s_wsolid[w]; s_wsolid[w] = tot; tot += v; }
        s_base = tot ? atomicAdd(n_solid, (unsigned long long)tot) : 0ULL;
    }
    __syncthreads();
    uint64_t out = s_base + s_wsolid[warp];

    for (uint32_t base = wb; base < we; base += 32 * K2_EMIT_ILP) {
        uint64_t key[K2_EMIT_ILP]; uint32_t c[K2_EMIT_ILP];
#pragma unroll
        for (int j = 0; j < K2_EMIT_ILP; j++) { uint32_t i = base + j * 32 + lane; key[j] = i < we ? keys[i] : EMPTY_KEY; }
#pragma unroll
        for (int j = 0; j < K2_EMIT_ILP; j++) { uint32_t i = base + j * 32 + lane; c[j] = key[j] != EMPTY_KEY ? counts[i] : 0; }
#pragma unroll
        for (int j = 0; j < K2_EMIT_ILP; j++) {
            uint32_t i = base + j * 32 + lane;
            bool occ = key[j] != EMPTY_KEY;
            bool solid = occ && c[j] >= amin && c[j] <= amax;
            unsigned m = __ballot_sync(0xFFFFFFFFu, solid);
            if (solid) {
                uint64_t pos = out + __popc(m & ((1u << lane) - 1));
                if (pos < solid_cap) { solid_kmers[pos] = key[j]; solid_counts[pos] = c[j]; }
            }
            out += __popc(m);
            if (occ) { keys[i] = EMPTY_KEY; counts[i] = 0; }
        }
    }
    block_add(n_distinct, (unsigned long long)my_distinct);
}

// counted k-mers handed in from outside (cdbg_set_solid_kmers): bring every k-mer to the canonical orientation of this encoding
__global__ void __launch_bounds__(256) k2_canonicalise(uint64_t *kmers, uint64_t n, int k) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        kmers[i] = canonical64(kmers[i] & kmer_mask(k), k);
}

// ------------------------------------------------------------------------------------------------ K3: junction end records

// A solid canonical k-mer x has two ends.  Each end is keyed by the canonical form c of its (k-1)-mer and lies on one SIDE
// of c: IN  (the k-mer, suitably oriented, ends with c)  or OUT (it begins with c).  ograph.cpp:384-434 calls these the
// `right` and `left` arrays; a palindromic c is on both sides (never compacts).
//   record value: bits 31..0 = end id (2*kmer_index + {0: left end, 1: right end}), bit 32 = side (0 IN, 1 OUT), bit 33 = palindrome
#define REC_SIDE_OUT (1ULL << 32)
#define REC_PALIN (1ULL << 33)
// A k-mer that is its own reverse complement (even k only) has BOTH its ends on the same side of the same junction (its suffix
// is the reverse complement of its prefix).  With such a k-mer a side can hold 5 ends -- 4 extensions of which one is the
// palindrome, counted twice -- one more than the 4 id slots per side.  So the two ends travel as ONE record: the left end's
// record carries REC_BOTH and stands for ends 2v and 2v+1, the right end's record carries REC_SKIP and is ignored by the join;
// the table remembers which stored ids are such pairs in bits 16..19 (IN list) / 20..23 (OUT list) of its counter word.
#define REC_BOTH (1ULL << 34)
#define REC_SKIP (1ULL << 35)
#define JOIN_BOTH_SHIFT_IN 16
#define JOIN_BOTH_SHIFT_OUT 20

// ids of one side of a junction with the (2v, 2v+1) pairs written out; at most 4 stored ids, at most one of them a pair
template <class ID>
__device__ __forceinline__ uint32_t join_expand_side(const ID *ids, uint32_t n, uint32_t both_mask, ID *out) {
    uint32_t m = 0;
    for (uint32_t a = 0; a < n; a++) {
        out[m++] = ids[a];
        if ((both_mask >> a) & 1u) out[m++] = ids[a] ^ (ID)1;
    }
    return m;
}

__device__ __forceinline__ void kmer_end_record(uint64_t x, int k, int end, uint32_t idx, uint64_t &key, uint64_t &val) {
    const int k1 = k - 1;
    uint64_t e = end == 0 ? (x >> 2) : (x & kmer_mask(k1));
    uint64_t r = revcomp64(e, k1);
    key = e < r ? e : r;
    bool canon_fwd = e <= r;
    // left end:  (k-1)-mer read forward is the prefix -> the k-mer begins with c (OUT) when canonical, else rc(k-mer) ends with c (IN)
    // right end: suffix -> the k-mer ends with c (IN) when canonical, else rc(k-mer) begins with c (OUT)
    bool out = (end == 0) ? canon_fwd : !canon_fwd;
    val = (uint64_t)(2u * idx + (uint32_t)end) | (out ? REC_SIDE_OUT : 0) | (e == r ? REC_PALIN : 0);
    if (!(k & 1) && x == revcomp64(x, k)) val |= end == 0 ? REC_BOTH : REC_SKIP;
}

struct K3Params {
    const uint64_t *solid_kmers;
    uint64_t n_solid;
    int k;
    uint32_t P;
    unsigned long long *bin_count;   // [P]
    const uint64_t *bin_off;         // [P]
    unsigned long long *bin_cursor;  // [P]
    uint64_t *rec_key;               // [2S]
    uint64_t *rec_val;               // [2S]
};

constexpr int K3_THREADS = 256;
constexpr int K3_PER_THREAD = 4;                          // k-mers per thread per tile
constexpr int K3_TILE = K3_THREADS * K3_PER_THREAD;       // 1024 k-mers = 2048 records per tile

// MODE 0: histogram (CTA-private in shared memory, flushed once).  MODE 1: scatter; each tile reserves one run per bin with a
// single global atomic, so the hot cursors are hit once per (tile, bin) instead of once per record.
template <int MODE>
__global__ void __launch_bounds__(K3_THREADS) k3_ends(K3Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *s_base = reinterpret_cast<uint64_t *>(smem_raw);                  // [P] MODE 1
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_base + (MODE == 1 ? p.P : 0));   // [P]
    const int t = threadIdx.x;
    const uint64_t n_tiles = (p.n_solid + K3_TILE - 1) / K3_TILE;
    for (uint32_t i = t; i < p.P; i += K3_THREADS) s_cnt[i] = 0;
    __syncthreads();
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t key[2 * K3_PER_THREAD], val[2 * K3_PER_THREAD];
        uint32_t bin[2 * K3_PER_THREAD], rank[2 * K3_PER_THREAD];
#pragma unroll
        for (int j = 0; j < K3_PER_THREAD; j++) {
            uint64_t i = tile * K3_TILE + (uint64_t)j * K3_THREADS + t;
            bool ok = i < p.n_solid;
            uint64_t x = ok ? p.solid_kmers[i] : 0;
#pragma unroll
            for (int end = 0; end < 2; end++) {
                const int r = 2 * j + end;
                kmer_end_record(x, p.k, end, (uint32_t)i, key[r], val[r]);
                bin[r] = ok ? hash_to_bin(mix64(key[r]), p.P) : 0xFFFFFFFFu;
                rank[r] = ok ? atomicAdd(&s_cnt[bin[r]], 1u) : 0;
            }
        }
        if (MODE == 1) {
            __syncthreads();
            for (uint32_t b = t; b < p.P; b += K3_THREADS) {
                uint32_t c = s_cnt[b];
                if (c) s_base[b] = p.bin_off[b] + atomicAdd(&p.bin_cursor[b], (unsigned long long)c);
                s_cnt[b] = 0;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 2 * K3_PER_THREAD; r++) {
                if (bin[r] == 0xFFFFFFFFu) continue;
                uint64_t pos = s_base[bin[r]] + rank[r];
                p.rec_key[pos] = key[r]; p.rec_val[pos] = val[r];
            }
            __syncthreads();
        }
    }
    if (MODE == 0) {
        __syncthreads();
        for (uint32_t b = t; b < p.P; b += K3_THREADS) {
            uint32_t c = s_cnt[b];
            if (c) atomicAdd(&p.bin_count[b], (unsigned long long)c);
        }
    }
}

static inline size_t k3_smem_bytes(int mode, uint32_t P) { return (size_t)P * 4 + (mode == 1 ? (size_t)P * 8 : 0) + 16; }

// ------------------------------------------------------------------------------------------------ K5: junction join

// table entry: key, packed counters (IN count in bits 7..0, OUT count in bits 15..8, REC_BOTH masks in bits 16..23, palindrome
// flag bit 31) and up to 4 end ids per side.  A (k-1)-mer has at most 4 k-mers on each side (one per extension base); the one
// case of 5 ENDS on a side (even k, one of the 4 k-mers is its own reverse complement) is stored as 4 ids + a pair flag.
struct JoinTable {
    uint64_t *keys;     // [cap]
    uint32_t *cnt;      // [cap]
    uint32_t *ids;      // [cap * 8]  (0..3 IN, 4..7 OUT)
    uint32_t cap;
};

__global__ void __launch_bounds__(256) k5_insert(const uint64_t *__restrict__ rec_key, const uint64_t *__restrict__ rec_val, uint64_t n,
                                                 JoinTable T, int *overflow) {
    pdl_wait();
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = rec_key[i], val = rec_val[i];
        if (val & REC_SKIP) continue;
        uint32_t slot = table_slot(key, T.cap);
        uint32_t probes = 0;
        for (;;) {
            uint64_t cur = ld_volatile_u64(&T.keys[slot]);
            if (cur == EMPTY_KEY) {
                cur = atomicCAS((unsigned long long *)&T.keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
                if (cur == EMPTY_KEY) cur = key;
            }
            if (cur == key) {
                const bool out = (val & REC_SIDE_OUT) != 0, pal = (val & REC_PALIN) != 0;
                // palindromic ends all go to the IN list (the join treats them as "all pairs")
                const bool use_out = out && !pal;
                uint32_t old = atomicAdd(&T.cnt[slot], use_out ? 0x100u : 0x1u);
                if (pal) atomicOr(&T.cnt[slot], 0x80000000u);
                uint32_t pos = use_out ? ((old >> 8) & 0xFF) : (old & 0xFF);
                if (pos < 4) {
                    T.ids[(size_t)slot * 8 + (use_out ? 4 : 0) + pos] = (uint32_t)val;
                    if (val & REC_BOTH) atomicOr(&T.cnt[slot], 1u << ((use_out ? JOIN_BOTH_SHIFT_OUT : JOIN_BOTH_SHIFT_IN) + pos));
                }
                break;
            }
            slot = slot + 1 == T.cap ? 0 : slot + 1;
            if (++probes > min(T.cap, 8192u)) { *overflow = 1; break; }
        }
    }
}

// Decide every junction of the bin (ograph.cpp:228-318): exactly one IN end and one OUT end belonging to two different
// k-mers and a non-palindromic key -> compaction link; otherwise every (IN, OUT) pair (every ordered pair for a
// palindrome) is an edge between unitig extremities (LinkTigs.cpp:365-369,428-432).  Slots are reset for the next bin.
__global__ void __launch_bounds__(256) k5_resolve(JoinTable T, uint32_t cap, uint32_t *link, uint64_t *edges, uint64_t edge_cap,
                                                  unsigned long long *n_edges, unsigned long long *n_compact) {
    pdl_wait();
    unsigned long long compact = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        if (T.keys[i] == EMPTY_KEY) continue;
        uint32_t c = T.cnt[i];
        T.keys[i] = EMPTY_KEY; T.cnt[i] = 0;
        uint32_t nin = c & 0xFF, nout = (c >> 8) & 0xFF; bool pal = (c >> 31) != 0;
        if (nin > 4) nin = 4; if (nout > 4) nout = 4;
        uint32_t in[5], out[5];
        {
            const uint4 vin = *reinterpret_cast<const uint4 *>(&T.ids[(size_t)i * 8]);
            const uint4 vout = *reinterpret_cast<const uint4 *>(&T.ids[(size_t)i * 8 + 4]);
            const uint32_t sin[4] = {vin.x, vin.y, vin.z, vin.w}, sout[4] = {vout.x, vout.y, vout.z, vout.w};
            nin = join_expand_side(sin, nin, (c >> JOIN_BOTH_SHIFT_IN) & 0xFu, in);
            nout = join_expand_side(sout, nout, (c >> JOIN_BOTH_SHIFT_OUT) & 0xFu, out);
        }
        if (!pal && nin == 1 && nout == 1 && (in[0] >> 1) != (out[0] >> 1)) {
            link[in[0]] = out[0]; link[out[0]] = in[0];
            compact++;
            continue;
        }
        uint32_t ne = pal ? nin * nin : 2 * nin * nout;
        if (ne == 0) continue;
        unsigned long long base = atomicAdd(n_edges, (unsigned long long)ne);
        if (base + ne > edge_cap) continue;   // counted; host retries with a larger buffer
        if (pal) {
            for (uint32_t a = 0; a < nin; a++) for (uint32_t b = 0; b < nin; b++) edges[base++] = ((uint64_t)in[a] << 32) | in[b];
        } else {
            for (uint32_t a = 0; a < nin; a++) for (uint32_t b = 0; b < nout; b++) {
                edges[base++] = ((uint64_t)in[a] << 32) | out[b];
                edges[base++] = ((uint64_t)out[b] << 32) | in[a];
            }
        }
    }
    block_add(n_compact, compact);
}

// ------------------------------------------------------------------------------------------------ K6: list ranking
#define K6_DONE 0x80000000u                                   // bit 31 of the hop count: nothing left to do for this element
#define K6_NIL 0xFFFFFFFFu                                    // next of a tail
#define K6_TAIL (((uint64_t)K6_NIL << 32) | K6_DONE)          // a tail: next = NIL, 0 hops, done
#define K6_PARKED 0x40000000u                                 // bit 30 (two-level ranking): waits on the representative of its window-local chain
#define K6_HOPS 0x3FFFFFFFu
//
// Directed element x (0 <= x < 2S) = "traverse k-mer x>>1 entering through end x&1 (0 = left, 1 = right)"; it leaves through
// end x^1, so succ(x) = link[x^1].  pair[x] = (next << 32 | hops to next); a traversal's last element (the tail) has next = NIL.
// (A self-pointer must NOT mean "tail": on a cycle whose length is a power of two, doubling lands an element on itself.)
// Updates are in place and asynchronous: each pair is read/written as one 64-bit word and the invariant
// "hops = distance from x to next along the list" holds at all times, so any interleaving converges.

__global__ void __launch_bounds__(256) k6_init(const uint32_t *__restrict__ link, uint64_t *pair, uint64_t n2) {
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t s = link[x ^ 1];
        pair[x] = s == NOLINK ? K6_TAIL : (((uint64_t)s << 32) | 1u);
    }
}


__global__ void __launch_bounds__(256) k6_jump(uint64_t *pair, uint64_t n2, unsigned long long *n_updated) {
    pdl_wait();
    unsigned long long upd = 0;
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p = ld_volatile_u64(&pair[x]);
        if ((uint32_t)p & K6_DONE) continue;
        uint32_t n = (uint32_t)(p >> 32);
        uint64_t q = ld_volatile_u64(&pair[n]);
        uint32_t n2_ = (uint32_t)(q >> 32);
        if (n2_ == K6_NIL) { *(volatile uint64_t *)&pair[x] = p | K6_DONE; continue; }        // next is the tail: finished
        *(volatile uint64_t *)&pair[x] = ((uint64_t)n2_ << 32) | (uint32_t)(((uint32_t)p & K6_HOPS) + ((uint32_t)q & K6_HOPS));
        upd++;
    }
    block_add(n_updated, upd);
}

// ---- two-level ranking (unweighted lists).  Solid k-mers leave the counting kernel bucket by bucket, so most compaction links join
// k-mers whose indices are close: consecutive k-mers of a unitig share their minimizer -- and with it their bucket -- for ~(k-m+2)/2
// positions.  k6_local ranks every window of K6L_NV consecutive k-mers in SHARED memory first: an element whose successor lies
// outside its window becomes the REPRESENTATIVE of its window-local chain, the chain's other elements PARK on it (pair = (rep,
// hops to rep) | K6_PARKED) or finish (chains that end inside the window).  Only the representatives (about one element in
// (k-m+2)/2) take part in the global, HBM-latency-bound rounds (k6_jump_list over a shrinking list); k6_resolve then adds the
// representative's result to every parked element.  A parked element is a valid jumping target: "hops = distance to next" holds.
constexpr int K6L_NV = 2048;                  // k-mers per window: 4096 directed elements, 32 KB of pairs
constexpr int K6L_THREADS = 512;
constexpr int K6L_MAX_ROUNDS = 14;

__global__ void __launch_bounds__(K6L_THREADS) k6_local(const uint32_t *__restrict__ link, uint64_t *pair, uint64_t n2, uint32_t *rep_list,
                                                       unsigned long long *n_reps) {
    __shared__ uint64_t s_pair[2 * K6L_NV];
    __shared__ int s_changed;
    __shared__ uint32_t s_wcnt[K6L_THREADS / 32];
    __shared__ unsigned long long s_obase;
    const uint32_t t = threadIdx.x, lane = lane_id(), warp = t >> 5;
    const uint64_t n_win = (n2 + 2 * K6L_NV - 1) / (2 * K6L_NV);
    for (uint64_t win = blockIdx.x; win < n_win; win += gridDim.x) {
        const uint64_t base = win * (2ULL * K6L_NV);
        const uint32_t cnt = (uint32_t)min((uint64_t)(2 * K6L_NV), n2 - base);      // even: n2 and the window size are
        for (uint32_t i = t; i < cnt; i += K6L_THREADS) {
            const uint32_t s_ = link[base + (i ^ 1u)];
            s_pair[i] = s_ == NOLINK ? K6_TAIL : (((uint64_t)s_ << 32) | 1u);
        }
        __syncthreads();
        for (int round = 0; round < K6L_MAX_ROUNDS; round++) {
            if (t == 0) s_changed = 0;
            __syncthreads();
            bool ch = false;
            for (uint32_t i = t; i < cnt; i += K6L_THREADS) {
                const uint64_t p = *(volatile uint64_t *)&s_pair[i];
                if ((uint32_t)p & (K6_DONE | K6_PARKED)) continue;
                const uint64_t n = (uint32_t)(p >> 32);
                if (n < base || n >= base + cnt) continue;                            // a representative (or a tail's predecessor outside: same thing)
                const uint64_t q = *(volatile uint64_t *)&s_pair[n - base];
                const uint32_t qn = (uint32_t)(q >> 32);
                if (qn == K6_NIL) { *(volatile uint64_t *)&s_pair[i] = p | K6_DONE; continue; }                 // next is the tail
                const uint32_t hops = ((uint32_t)p & K6_HOPS) + ((uint32_t)q & K6_HOPS);
                if ((uint32_t)q & K6_PARKED) { *(volatile uint64_t *)&s_pair[i] = ((uint64_t)qn << 32) | hops | K6_PARKED; continue; }   // park on the same representative
                if (!((uint32_t)q & K6_DONE) && ((uint64_t)qn < base || (uint64_t)qn >= base + cnt)) {
                    *(volatile uint64_t *)&s_pair[i] = p | K6_PARKED;                 // next IS a representative: park on it
                    continue;
                }
                *(volatile uint64_t *)&s_pair[i] = ((uint64_t)qn << 32) | hops;
                ch = true;
            }
            if (ch) s_changed = 1;
            __syncthreads();
            if (!s_changed) break;
            __syncthreads();
        }
        // write back; the representatives (unfinished, not parked) go on the list
        for (uint32_t i0 = 0; i0 < cnt; i0 += K6L_THREADS) {                          // block-uniform
            const uint32_t i = i0 + t;
            bool rep = false;
            if (i < cnt) {
                const uint64_t p = s_pair[i];
                pair[base + i] = p;
                rep = !((uint32_t)p & (K6_DONE | K6_PARKED));
            }
            const unsigned m = __ballot_sync(0xFFFFFFFFu, rep);
            if (lane == 0) s_wcnt[warp] = __popc(m);
            __syncthreads();
            if (t == 0) {
                uint32_t tot = 0;
                for (int w = 0; w < K6L_THREADS / 32; w++) { uint32_t v = s_wcnt[w]; s_wcnt[w] = tot; tot += v; }
                s_obase = tot ? atomicAdd(n_reps, (unsigned long long)tot) : 0ULL;
            }
            __syncthreads();
            if (rep) rep_list[s_obase + s_wcnt[warp] + __popc(m & ((1u << lane) - 1))] = (uint32_t)(base + i);
            __syncthreads();
        }
    }
}

// every parked element takes over its representative's result (or, on a cycle, becomes a plain moving element again)
__global__ void __launch_bounds__(256) k6_resolve(uint64_t *pair, uint64_t n2) {
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t p = pair[x];
        if (!((uint32_t)p & K6_PARKED)) continue;
        const uint32_t e = (uint32_t)(p >> 32);
        const uint64_t q = pair[e];                                                   // representatives are never parked: not written by this kernel
        if ((uint32_t)q & K6_DONE) pair[x] = (q & 0xFFFFFFFF00000000ULL) | (((uint32_t)p & K6_HOPS) + ((uint32_t)q & K6_HOPS)) | K6_DONE;
        else pair[x] = p & ~(uint64_t)K6_PARKED;
    }
}

// Variant that walks a LIST of still-moving elements instead of the whole array (in_list == nullptr: all x < n), and can emit
// the elements that moved -- exactly the ones that may move again -- as the next list (out_list != nullptr; one reservation
// atomic per block).  After a few rounds only the elements far from a tail are left, so later rounds touch a small fraction.
// (An element whose next became the tail is simply not listed again; it stays without its DONE bit, which nothing relies on.)
__global__ void __launch_bounds__(256) k6_jump_list(uint64_t *pair, const uint32_t *__restrict__ in_list, uint64_t n, uint32_t *out_list,
                                                    unsigned long long *n_out, unsigned long long *n_updated) {
    pdl_wait();
    __shared__ uint32_t s_wcnt[8];
    __shared__ unsigned long long s_obase;
    unsigned long long upd = 0;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (uint64_t base = blockIdx.x * (uint64_t)blockDim.x; base < n; base += (uint64_t)gridDim.x * blockDim.x) {   // block-uniform
        const uint64_t i = base + threadIdx.x;
        bool moved = false;
        uint32_t x = 0;
        if (i < n) {
            x = in_list ? in_list[i] : (uint32_t)i;
            uint64_t p = ld_volatile_u64(&pair[x]);
            if (!((uint32_t)p & K6_DONE)) {
                uint32_t nx = (uint32_t)(p >> 32);
                uint64_t q = ld_volatile_u64(&pair[nx]);
                uint32_t n2_ = (uint32_t)(q >> 32);
                if (n2_ == K6_NIL) *(volatile uint64_t *)&pair[x] = p | K6_DONE;
                else {
                    *(volatile uint64_t *)&pair[x] = ((uint64_t)n2_ << 32) | (uint32_t)(((uint32_t)p & K6_HOPS) + ((uint32_t)q & K6_HOPS));
                    moved = true; upd++;
                }
            }
        }
        if (out_list) {
            const unsigned m = __ballot_sync(0xFFFFFFFFu, moved);
            if (lane == 0) s_wcnt[warp] = __popc(m);
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t tot = 0;
                for (int w = 0; w < 8; w++) { uint32_t v = s_wcnt[w]; s_wcnt[w] = tot; tot += v; }
                s_obase = tot ? atomicAdd(n_out, (unsigned long long)tot) : 0ULL;
            }
            __syncthreads();
            if (moved) out_list[s_obase + s_wcnt[warp] + __popc(m & ((1u << lane) - 1))] = x;
            __syncthreads();
        }
    }
    block_add(n_updated, upd);
}

// ---- cycles: elements that never reach a tail.  Find the smallest k-mer index on each cycle by pointer doubling with a
// running minimum, cut the link at that k-mer's left end (the reference cuts wherever its scan order happens to close
// the cycle: bglue_algo.cpp:400-416), record the two edges of the cut junction, and re-rank the affected elements.

__device__ __forceinline__ bool k6_unfinished(const uint64_t *pair, uint64_t x) {
    uint64_t p = pair[x];
    if ((uint32_t)p & K6_DONE) return false;
    uint32_t n = (uint32_t)(p >> 32);
    return (uint32_t)(pair[n] >> 32) != K6_NIL;
}

__global__ void __launch_bounds__(256) k6_cycle_mark(const uint64_t *pair, const uint32_t *link, uint64_t n2, uint8_t *flag,
                                                     uint64_t *lab) {
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x) {
        bool u = k6_unfinished(pair, x);
        flag[x] = u;
        if (u) lab[x] = ((uint64_t)link[x ^ 1] << 32) | (uint32_t)(x >> 1);   // (succ, min k-mer index so far)
    }
}

__global__ void __launch_bounds__(256) k6_cycle_min(const uint8_t *flag, uint64_t *lab, uint64_t n2) {
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x) {
        if (!flag[x]) continue;
        uint64_t a = ld_volatile_u64(&lab[x]);
        uint64_t b = ld_volatile_u64(&lab[(uint32_t)(a >> 32)]);
        uint32_t m = min((uint32_t)a, (uint32_t)b);
        *(volatile uint64_t *)&lab[x] = (b & 0xFFFFFFFF00000000ULL) | m;
    }
}

// number of cycles about to be cut (= flagged minima): lets the host make room for their 2 edges each before k6_cycle_cut runs
__global__ void __launch_bounds__(256) k6_cycle_count(const uint8_t *flag, const uint64_t *lab, uint64_t n2, unsigned long long *n_cycles) {
    unsigned long long c = 0;
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x)
        c += flag[x] && !(x & 1) && (uint32_t)lab[x] == (uint32_t)(x >> 1);
    block_add(n_cycles, c);
}

__global__ void __launch_bounds__(256) k6_cycle_cut(const uint8_t *flag, const uint64_t *lab, uint64_t n2, uint32_t *link,
                                                    uint64_t *edges, uint64_t edge_cap, unsigned long long *n_edges,
                                                    unsigned long long *n_cycles) {
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x) {
        if (!flag[x] || (x & 1)) continue;
        if ((uint32_t)lab[x] != (uint32_t)(x >> 1)) continue;    // not the minimum k-mer of its cycle
        uint32_t e = (uint32_t)x;            // left end of the minimum k-mer
        uint32_t f = link[e];
        link[e] = NOLINK; link[f] = NOLINK;
        unsigned long long base = atomicAdd(n_edges, 2ULL);
        if (base + 2 <= edge_cap) {
            edges[base] = ((uint64_t)e << 32) | f;
            edges[base + 1] = ((uint64_t)f << 32) | e;
        }
        atomicAdd(n_cycles, 1ULL);
    }
}

__global__ void __launch_bounds__(256) k6_reinit_flagged(const uint8_t *flag, const uint32_t *link, uint64_t *pair, uint64_t n2) {
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x) {
        if (!flag[x]) continue;
        uint32_t s = link[x ^ 1];
        pair[x] = s == NOLINK ? K6_TAIL : (((uint64_t)s << 32) | 1u);
    }
}

// ---- unitig numbering.  For k-mer v: T0 = terminus reached going forward, T1 = terminus reached going backward; the unitig
// starts at B = min(T0, T1).  info[v] = (B << 32) | pos << 1 | reversed.  Heads (pos == 0) also publish the unitig length.
__global__ void __launch_bounds__(256) k6_finalize(const uint64_t *__restrict__ pair, uint64_t S, int k, uint64_t *info, uint32_t *head_flag,
                                                   uint32_t *head_nt) {
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < S; v += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p0 = pair[2 * v], p1 = pair[2 * v + 1];
        uint32_t n0 = (uint32_t)(p0 >> 32), n1 = (uint32_t)(p1 >> 32);
        uint32_t T0 = (n0 == K6_NIL ? (uint32_t)(2 * v) : n0) ^ 1u, T1 = (n1 == K6_NIL ? (uint32_t)(2 * v + 1) : n1) ^ 1u;
        uint32_t d0 = (uint32_t)p0 & ~K6_DONE, d1 = (uint32_t)p1 & ~K6_DONE;
        uint32_t B, pos, rev;
        if (T1 < T0) { B = T1; pos = d1; rev = 0; } else { B = T0; pos = d0; rev = 1; }
        info[v] = ((uint64_t)B << 32) | ((uint64_t)pos << 1) | rev;
        bool head = pos == 0;
        head_flag[v] = head;
        head_nt[v] = head ? (d0 + d1 + 1 + (uint32_t)(k - 1)) : 0;
    }
}

// ------------------------------------------------------------------------------------------------ exclusive scan (u32 -> u64)

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint64_t block_exclusive_scan_u64(uint64_t v, uint64_t &total) {
    __shared__ uint64_t s_w[SCAN_THREADS / 32];
    uint64_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint64_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane_id() >= d) incl += n; }
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint64_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) { uint64_t s = s_w[w]; if (w < (int)(threadIdx.x >> 5)) woff += s; tot += s; }
    total = tot;
    __syncthreads();
    return woff + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const uint32_t *__restrict__ in, uint64_t n, uint64_t *tile_sums) {
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    uint64_t s = 0;
    for (int j = 0; j < SCAN_ITEMS; j++) { uint64_t i = base + (uint64_t)j * SCAN_THREADS + threadIdx.x; if (i < n) s += in[i]; }
    uint64_t tot; block_exclusive_scan_u64(s, tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_sums_inplace(uint64_t *tile_sums, uint64_t n_tiles, uint64_t *total_out) {
    uint64_t carry = 0;
    for (uint64_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        uint64_t i = base + threadIdx.x;
        uint64_t v = i < n_tiles ? tile_sums[i] : 0, tot;
        uint64_t ex = block_exclusive_scan_u64(v, tot);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const uint32_t *__restrict__ in, uint64_t n, const uint64_t *tile_sums, uint64_t *out) {
    // thread-contiguous items for a simple in-register scan
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS]; uint64_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) { v[j] = (base + j < n) ? in[base + j] : 0; s += v[j]; }
    uint64_t tot; uint64_t ex = block_exclusive_scan_u64(s, tot) + tile_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) { if (base + j < n) out[base + j] = ex; ex += v[j]; }
}

// ------------------------------------------------------------------------------------------------ K6/K8: spell

struct SpellParams {
    const uint64_t *solid_kmers; const uint32_t *solid_counts; uint64_t S; int k;
    const uint64_t *info;          // per k-mer
    const uint64_t *uid_scan;      // exclusive scan of head flags  (index by head k-mer)
    const uint64_t *nt_scan;       // exclusive scan of head_nt     (index by head k-mer)
    char *seq;                     // [total_nt]
    uint64_t *seq_off;             // [U+1]
    unsigned long long *kc;        // [U]
    uint32_t *abund;               // [S] or null
    uint32_t *unitig_of_kmer;      // [S] (for K9)
    const uint32_t *link;          // [2S] local links (only read when the terminus arrays are requested)
    uint32_t *term_begin;          // [U] or null: k-mer end at which each unit begins
    uint32_t *term_end;            // [U] or null: k-mer end at which it ends
};

__global__ void __launch_bounds__(256) k8_spell(SpellParams p) {
    const char NT[4] = {'A', 'C', 'T', 'G'};
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < p.S; v += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t inf = p.info[v];
        uint32_t B = (uint32_t)(inf >> 32), pos = (uint32_t)inf >> 1, rev = (uint32_t)inf & 1u;
        uint32_t h = B >> 1;
        uint64_t u = p.uid_scan[h], off = p.nt_scan[h];
        uint64_t x = p.solid_kmers[v];
        if (rev) x = revcomp64(x, p.k);
        if (p.term_begin) {
            if (pos == 0) p.term_begin[u] = B;
            uint32_t exit_end = 2u * (uint32_t)v + (rev ? 0u : 1u);
            if (p.link[exit_end] == NOLINK) p.term_end[u] = exit_end;
        }
        if (pos == 0) {
            p.seq_off[u] = off;
            for (int j = 0; j < p.k; j++) p.seq[off + j] = NT[(x >> (2 * (p.k - 1 - j))) & 3];
        } else {
            p.seq[off + (p.k - 1) + pos] = NT[x & 3];
        }
        uint32_t c = p.solid_counts[v];
        atomicAdd(&p.kc[u], (unsigned long long)c);
        if (p.abund) p.abund[off - u * (uint64_t)(p.k - 1) + pos] = c;
        p.unitig_of_kmer[v] = (uint32_t)u;
    }
}

// ------------------------------------------------------------------------------------------------ K9: links

struct LinkRec { uint32_t from, to; uint8_t from_sign, to_sign, pad0, pad1; };

// edge (e -> f) between k-mer ends that are unitig extremities:  L:<'-' if e is the BEGIN of its unitig else '+'>:<unitig(f)>:
// <'+' if f is a BEGIN else '-'>   (LinkTigs.cpp:377,438: rc = (pos == UNITIG_END))
__global__ void __launch_bounds__(256) k9_translate(const uint64_t *__restrict__ edges, uint64_t n, const uint64_t *__restrict__ info,
                                                    const uint32_t *__restrict__ unitig_of_kmer, LinkRec *out, uint32_t *per_unitig) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t ed = edges[i];
        uint32_t e = (uint32_t)(ed >> 32), f = (uint32_t)ed;
        bool e_begin = (uint32_t)(info[e >> 1] >> 32) == e;
        bool f_begin = (uint32_t)(info[f >> 1] >> 32) == f;
        LinkRec r;
        r.from = unitig_of_kmer[e >> 1]; r.to = unitig_of_kmer[f >> 1];
        r.from_sign = e_begin ? '-' : '+'; r.to_sign = f_begin ? '+' : '-';
        r.pad0 = r.pad1 = 0;
        out[i] = r;
        atomicAdd(&per_unitig[r.from], 1u);
    }
}

// group the links by `from` unitig (CSR): link_off = exclusive scan of per_unitig, cursor zero-initialised
__global__ void __launch_bounds__(256) k9_group(const LinkRec *__restrict__ in, uint64_t n, const uint64_t *__restrict__ link_off,
                                                uint32_t *cursor, LinkRec *out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        LinkRec r = in[i];
        out[link_off[r.from] + atomicAdd(&cursor[r.from], 1u)] = r;
    }
}

// ------------------------------------------------------------------------------------------------ misc


// synthetic reads (bench): genome base g(i) = hash(seed, i) & 3 generated on the fly, so no genome array is needed
__device__ __forceinline__ uint32_t synth_base(uint64_t seed, uint64_t i) { return (uint32_t)(mix64(seed ^ (i * 0x9E3779B97F4A7C15ULL)) >> 61) & 3u; }

__global__ void __launch_bounds__(256) synth_reads_kernel(uint8_t *out, uint64_t genome_len, uint32_t read_len, uint64_t first_read, uint64_t n_reads,
                                                          uint32_t err_ppm, uint64_t seed) {
    const char NT[4] = {'A', 'C', 'T', 'G'};
    const uint64_t stride = (uint64_t)read_len + 1;
    const uint64_t total = n_reads * stride;
    for (uint64_t idx = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; idx < total; idx += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t r = idx / stride; uint32_t j = (uint32_t)(idx - r * stride);
        r += first_read;
        if (j == read_len) { out[idx] = '\n'; continue; }
        uint64_t hr = mix64(seed * 0xD1342543DE82EF95ULL + r + 1);
        uint64_t start = hr % (genome_len - read_len + 1);
        bool rev = (hr >> 63) != 0;
        uint64_t gpos = rev ? start + (read_len - 1 - j) : start + j;
        uint32_t b = synth_base(seed, gpos);
        if (rev) b ^= 2u;
        uint64_t he = mix64((seed + 0x5851F42D4C957F2DULL) ^ ((r * stride + j) * 0xA24BAED4963EE407ULL));
        if ((uint32_t)(he % 1000000ULL) < err_ppm) b = (b + 1 + (uint32_t)((he >> 40) % 3)) & 3u;
        out[idx] = NT[b];
    }
}
