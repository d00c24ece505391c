// sk.cuh -- stage 1 on SUPER-K-MERS (k <= 31): reads -> minimizer-bucketed super-k-mers -> counted solid k-mers.
//
// Reference rows: a3 minimizer of a k-mer (Model.hpp:1040-1064, 1254-1287), a4 minimizer -> partition
// (PartiInfo.cpp:130-185 Repartitor), a5 super-k-mers (Sequence2SuperKmer.hpp:81-134, Model.hpp:1386-1471 SuperKmer::save,
// SortingCountAlgorithm.cpp:1081-1151 FillPartitions), a6 counting from super-k-mers (PartitionsCommand.cpp:944-1128
// ReadSuperKCommand) and a7 the solidity filter.  Like the reference, a k-mer instance is never materialised in global memory:
// the read stream is cut into maximal runs of consecutive valid k-mers that share their minimizer, each run is stored as its
// k+n-1 nucleotides (2 bits each) in a fixed 16-byte record, records are radix-partitioned by a hash of the minimizer, and the
// counting kernel expands a bucket's records into canonical k-mers in SHARED memory only.  HBM sees ~1.6 bytes per k-mer instance
// instead of the 8 bytes (written once, read three times, rewritten once) of the k-mer-instance path in k1_tma.cuh / k2_smem.cuh,
// and the same records are what crosses NVLink in the sharded path (exchange X1).
//
// Minimizer: the canonical m-mer of smallest mmer_rank (kmer.cuh) among the k-m+1 m-mers of the k-mer -- a function of the
// canonical k-mer, so both strands of a k-mer land in the same bucket.  The order is ours; m is chosen by the host
// from the input size so that no single minimizer dominates a bucket.
//
// Record (ulonglong2):  x = nucleotides 0..31 (first at bits 63:62);  y = nucleotides 32..51 in bits 63..24,
//                       bits 23..6 = `sub` (18 hash bits below the coarse bin: the fine bin is (sub * F) >> 18),
//                       bits 5..0 = n - 1  (n k-mers, n <= 53 - k, i.e. at most 52 nucleotides).
#pragma once
#include "kernels.cuh"
#include "k1_tma.cuh"
#include "k2_smem.cuh"

constexpr int SK_MAX_NT = 52;
constexpr int SK_SUB_BITS = 18;
__host__ __device__ __forceinline__ int sk_nmax(int k) { return SK_MAX_NT + 1 - k; }

// hash of the minimizer's rank: its top bits give the coarse bin (destination-major in sharded mode: owner = bin / P_local, which
// equals owner_of_rank(minrank, world) of shard.cuh), the bits below give `sub`
static CDBG_HD uint32_t sk_hash(uint32_t minrank) {
    uint32_t h = minrank ^ 0x68E31DA4u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
// With V > 1 waves (stage 1 repeated over the input, one slice of the buckets per pass: the analogue of the reference's nb_passes,
// ConfigurationAlgorithm.cpp:350) the hash is cut into Pc * V virtual bins: virtual bin v belongs to wave v % V and is coarse bin
// v / V of that wave -- so the owner GPU (coarse / P_local) of a minimizer does not depend on V.  Returns the wave.
static CDBG_HD uint32_t sk_bins(uint32_t minrank, uint32_t Pc, uint32_t V, uint32_t &coarse, uint32_t &sub) {
    const uint64_t x = (uint64_t)sk_hash(minrank) * ((uint64_t)Pc * V);      // Pc * V < 2^32 (host)
    const uint32_t v = (uint32_t)(x >> 32);
    sub = (uint32_t)x >> (32 - SK_SUB_BITS);
    if (V == 1) { coarse = v; return 0; }
    coarse = v / V;
    return v - coarse * V;
}
__device__ __forceinline__ uint32_t sk_rec_n(const ulonglong2 &r) { return ((uint32_t)r.y & 63u) + 1u; }
__device__ __forceinline__ uint32_t sk_rec_sub(const ulonglong2 &r) { return ((uint32_t)r.y >> 6) & ((1u << SK_SUB_BITS) - 1u); }
__device__ __forceinline__ uint32_t sk_fine(uint32_t sub, uint32_t F) { return (sub * F) >> SK_SUB_BITS; }   // F <= 1024: no overflow

// the record read in the other direction: reverse complement of its nt = n + k - 1 nucleotides, left-aligned again
__device__ __forceinline__ ulonglong2 sk_canon_rec(const ulonglong2 r, int k) {
    const uint32_t nt = sk_rec_n(r) + (uint32_t)k - 1u;
    const uint64_t x = r.x, y = r.y & ~0xFFFFFFULL;
    const uint64_t rx = revcomp64(y, 32), ry = revcomp64(x, 32);      // (rx:ry) = the reversed string right-aligned in 128 bits
    const uint32_t s = 128u - 2u * nt;                                // nt <= 52: s >= 24
    uint64_t cx, cy;
    if (s >= 64) { cx = ry << (s - 64); cy = 0; }
    else { cx = (rx << s) | (ry >> (64 - s)); cy = ry << s; }
    const bool flip = cx < x || (cx == x && cy < y);
    return flip ? make_ulonglong2(cx, cy | (r.y & 0xFFFFFFULL)) : r;
}


// ------------------------------------------------------------------------------------------------ K1: reads -> super-k-mers

constexpr int SK1_LIST_CAP = 2048;        // run descriptors per tile handled through shared memory (more: written directly)
constexpr int SK1_STAGE_CAP = 2048;       // records per tile staged in shared memory for coalesced stores (more: written directly)
constexpr int SK1_R = SK1_LIST_CAP / K1_THREADS;
constexpr int SK1_RANKS = K1_TILE + 32;
constexpr int SK1_RANK_WORDS = SK1_RANKS + SK1_RANKS / 32 + 2;
#define SK1_SKIP 0xFFFFFFFFFFFFFFFFULL
constexpr size_t SK1_X_LIST = (size_t)SK1_LIST_CAP * 8 + (size_t)SK1_STAGE_CAP * 16 + (size_t)SK1_STAGE_CAP * 2;
constexpr size_t SK1_X_BYTES = ((SK1_X_LIST > (size_t)SK1_RANK_WORDS * 4 ? SK1_X_LIST : (size_t)SK1_RANK_WORDS * 4) + 15) & ~size_t(15);

struct SK1Params {
    const uint8_t *bytes;            // device memory, or pinned host memory mapped into the device address space (single pass only)
    const uint8_t *tail;             // device copy of bytes[tail_start, n_bytes): the ragged last tiles are read from here
    uint64_t tail_start;
    uint64_t n_bytes, n_tiles;
    int k, m;
    uint32_t Pc;                     // coarse bins (<= K1_MAX_BINS); sharded: world * P_local, destination-major
    uint32_t V, wave;                // waves: only the runs whose virtual bin belongs to `wave` are kept (V = 1: all)
    const uint64_t *bin_base;        // [Pc] first record of each coarse bin's region in `out`
    const uint64_t *bin_cap;         // [Pc] capacity of the region (records)
    unsigned long long *bin_cursor;  // [Pc] zero-initialised; ends as the number of records per bin (MODE 0: the histogram)
    ulonglong2 *out;
    int *overflow;                   // set when a bin outgrows its region (host: histogram pass, exact regions, redo)
    unsigned long long *n_kmers;     // += valid k-mer instances
    unsigned long long *n_runs;      // += super-k-mer records
};

// one record: n k-mers starting at tile position q (nucleotides q .. q+n+k-2) from the packed tile
__device__ __forceinline__ ulonglong2 sk_make_record(const uint64_t *s_pk, uint32_t q, uint32_t n, int k, uint32_t sub) {
    const uint32_t w = q >> 5, o = (q & 31u) * 2u;
    const uint64_t a = s_pk[w], b = s_pk[w + 1], c = s_pk[w + 2];
    uint64_t hi = o ? (a << o) | (b >> (64 - o)) : a;
    uint64_t lo = o ? (b << o) | (c >> (64 - o)) : b;
    const uint32_t nt = n + (uint32_t)k - 1u;
    if (nt <= 32) { hi &= ~0ULL << (64 - 2 * nt); lo = 0; }
    else lo &= ~0ULL << (64 - 2 * (nt - 32));
    return make_ulonglong2(hi, lo | ((uint64_t)sub << 6) | (uint64_t)(n - 1u));
}

// a run that could not go through the tile's shared-memory list (more than SK1_LIST_CAP runs in one tile): reserve and write alone
#ifdef CDBG_EMU_ACTIVE
#define SK_NOINLINE __attribute__((noinline))
#else
#define SK_NOINLINE __noinline__
#endif
template <int MODE>
__device__ SK_NOINLINE void sk1_direct(const SK1Params &p, const uint64_t *s_pk, uint32_t minrank, uint32_t start, uint32_t len) {
    uint32_t c, sub;
    if (sk_bins(minrank, p.Pc, p.V, c, sub) != p.wave) return;
    const uint32_t nmax = (uint32_t)sk_nmax(p.k), pieces = (len + nmax - 1) / nmax;
    const unsigned long long old = atomicAdd(&p.bin_cursor[c], (unsigned long long)pieces);
    atomicAdd(p.n_runs, (unsigned long long)pieces);
    if (MODE == 0) return;
    if (old + pieces > p.bin_cap[c]) { *p.overflow = 1; return; }
    for (uint32_t j = 0; j < pieces; j++) {
        const uint32_t n = min(nmax, len - j * nmax);
        p.out[p.bin_base[c] + old + j] = sk_make_record(s_pk, start + j * nmax, n, p.k, sub);
    }
}

// MODE 0: histogram only (records per coarse bin).  MODE 1: partitioned write into the bins' regions.
template <int MODE>
__global__ void __launch_bounds__(K1_THREADS, 2) sk1_extract(SK1Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint8_t *s_raw = smem_raw;                                                              // [2][K1T_RAW]  TMA destinations
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + 2 * K1T_RAW);                 // [2] mbarriers
    uint64_t *s_pk = s_bar + 2;                                                             // [THREADS + 4] packed bases, 32 per word
    uint64_t *s_gbase = s_pk + K1_THREADS + 4;                                              // [Pc rounded up to even: region X stays 16-byte aligned]
    uint64_t *s_x = s_gbase + ((p.Pc + 1u) & ~1u);                                          // region X, two lives:
    uint32_t *s_R = reinterpret_cast<uint32_t *>(s_x);                                      //   [SK1_RANK_WORDS] m-mer ranks (phases B, C)
    uint64_t *s_list = s_x;                                                                 //   [SK1_LIST_CAP] run descriptors (phase E on)
    ulonglong2 *s_stage = reinterpret_cast<ulonglong2 *>(s_x + SK1_LIST_CAP);               //   [SK1_STAGE_CAP] records sorted by bin
    uint16_t *s_sb = reinterpret_cast<uint16_t *>(s_stage + SK1_STAGE_CAP);                 //   [SK1_STAGE_CAP] bin of each staged record
    uint32_t *s_vm = reinterpret_cast<uint32_t *>(reinterpret_cast<unsigned char *>(s_x) + SK1_X_BYTES);   // [THREADS + 2]
    uint32_t *s_rs = s_vm + K1_THREADS + 2;                                                 // [THREADS] reset bits 31..1 | "run continues into the next thread" bit 0
    uint32_t *s_cnt = s_rs + K1_THREADS;                                                    // [Pc]
    uint32_t *s_off = s_cnt + p.Pc;                                                         // [Pc + 1]
    __shared__ uint32_t s_warp[K1_THREADS / 32];
    __shared__ uint32_t s_nrec;

    const int t = threadIdx.x, k = p.k, m = p.m, w = k - m + 1;
    const uint32_t lane = lane_id();
    const uint32_t nmax = (uint32_t)sk_nmax(k);
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    int pw = 1;                                       // largest power of two <= w
    while (2 * pw <= w) pw *= 2;
    unsigned long long my_kmers = 0, my_runs = 0;

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
    for (uint32_t i = t; i < p.Pc; i += K1_THREADS) s_cnt[i] = 0;
    if (t < 3) s_pk[K1_THREADS + 1 + t] = 0;          // words past the halo: only ever shifted into bits that are masked off
    __syncthreads();
    if (t == 0 && first < p.n_tiles) issue(first, 0);

    uint32_t it = 0;
    for (uint64_t tile = first; tile < p.n_tiles; tile += stride, it++) {
        const int stage = it & 1;
        const uint32_t parity = (it >> 1) & 1;
        if (t == 0 && tile + stride < p.n_tiles) issue(tile + stride, stage ^ 1);   // prefetch the next tile
        mbar_wait(&s_bar[stage], parity);
        // ---- A: ASCII -> 2 bit + validity
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
        // ---- B: ranks of the canonical m-mers starting at my 32 positions (rolling); kept in registers and published for the neighbour
        uint32_t a[64];
        {
            uint64_t hi = s_pk[t], lo = s_pk[t + 1];
            if (m == 16) {                                     // the usual case (inputs of 10 Mbases and more): the m-mer is the high word, 32-bit rolling
                uint32_t fm = (uint32_t)(hi >> 32);
                uint32_t rm = (uint32_t)revcomp64(fm, 16);
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    if (i > 0) {
                        hi = (hi << 2) | (lo >> 62); lo <<= 2;
                        fm = (uint32_t)(hi >> 32);
                        rm = (rm >> 2) | (((fm & 3u) ^ 2u) << 30);
                    }
                    a[i] = mmer_rank(fm < rm ? fm : rm);
                    s_R[k1ts_pad((uint32_t)(t * 32 + i))] = a[i];
                }
            } else {
                uint64_t fm = hi >> (64 - 2 * m);
                uint64_t rm = revcomp64(fm, m);
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    if (i > 0) {
                        hi = (hi << 2) | (lo >> 62); lo <<= 2;
                        fm = hi >> (64 - 2 * m);
                        rm = (rm >> 2) | (((fm & 3) ^ 2) << (2 * (m - 1)));
                    }
                    a[i] = mmer_rank((uint32_t)(fm < rm ? fm : rm));
                    s_R[k1ts_pad((uint32_t)(t * 32 + i))] = a[i];
                }
            }
            // the halo's m-mers (positions K1_TILE + j, j <= w - 2: all inside the halo word), one per thread
            if (t < w - 1) {
                const uint64_t hw = s_pk[K1_THREADS];
                const uint64_t mf = (t ? (hw << (2 * t)) : hw) >> (64 - 2 * m);
                s_R[k1ts_pad((uint32_t)(K1_TILE + t))] = mmer_rank((uint32_t)canonical64(mf, m));
            }
        }
        __syncthreads();
        // ---- C: minimizer rank of the k-mer at each of my positions (and at the next thread's first): sliding-window minimum over w
        //         ranks by doubling in registers; a[i] = min over [i, i + s) after the level of stride s
#pragma unroll
        for (int j = 0; j < 31; j++) a[32 + j] = j < w ? s_R[k1ts_pad((uint32_t)(t * 32 + 32 + j))] : 0xFFFFFFFFu;
        a[63] = 0xFFFFFFFFu;
        if (pw >= 2) {
#pragma unroll
            for (int i = 0; i < 63; i++) a[i] = min(a[i], a[i + 1]);
        }
        if (pw >= 4) {
#pragma unroll
            for (int i = 0; i < 61; i++) a[i] = min(a[i], a[i + 2]);
        }
        if (pw >= 8) {
#pragma unroll
            for (int i = 0; i < 57; i++) a[i] = min(a[i], a[i + 4]);
        }
        if (pw >= 16) {
#pragma unroll
            for (int i = 0; i < 49; i++) a[i] = min(a[i], a[i + 8]);
        }
        uint32_t mr[33];
        switch (w - pw) {
#define SK1_COMBINE(D) case D: _Pragma("unroll") for (int i = 0; i < 33; i++) mr[i] = min(a[i], a[i + D]); break;
            SK1_COMBINE(0) SK1_COMBINE(1) SK1_COMBINE(2) SK1_COMBINE(3) SK1_COMBINE(4) SK1_COMBINE(5) SK1_COMBINE(6) SK1_COMBINE(7)
            SK1_COMBINE(8) SK1_COMBINE(9) SK1_COMBINE(10) SK1_COMBINE(11) SK1_COMBINE(12) SK1_COMBINE(13) SK1_COMBINE(14)
            default: _Pragma("unroll") for (int i = 0; i < 33; i++) mr[i] = min(a[i], a[i + 15]); break;
#undef SK1_COMBINE
        }
        // ---- D: validity, run structure.  same[i]: positions i and i+1 are both valid k-mers with the same minimizer
        const uint64_t V = ((uint64_t)s_vm[t] << 32) | s_vm[t + 1];
        const uint64_t KV = kmer_valid_mask(V, k);                 // bit 63-i: the k-mer at my position i is valid (i <= 64 - k)
        const uint32_t valid = __brev((uint32_t)(KV >> 32));       // bit i: position i valid
        const uint32_t valid32 = t == K1_THREADS - 1 ? 0u : (uint32_t)(KV >> 31) & 1u;   // the next thread's first position (tile end: a forced break)
        uint32_t same = 0;
#pragma unroll
        for (int i = 0; i < 32; i++) same |= (uint32_t)(mr[i] == mr[i + 1]) << i;
        same &= valid & ((valid >> 1) | (valid32 << 31));
        const uint32_t ends = valid & ~same;                       // bit i: a run ends at position i
        const uint32_t resets_hi = ~(same << 1) & ~1u;             // bit i (i >= 1): position i does not continue a run
        my_kmers += __popc(valid);
        __syncthreads();                                           // everyone is done with s_R: region X changes hands
        s_rs[t] = resets_hi | (same >> 31);
        // ---- E: slots of my runs in the tile's descriptor list (block-wide exclusive scan of the run counts)
        uint32_t slot;
        {
            const uint32_t cnt = __popc(ends);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += n; }
            if (lane == 31) s_warp[t >> 5] = incl;
            __syncthreads();
            uint32_t wbase = 0, tot = 0;
#pragma unroll
            for (int x = 0; x < K1_THREADS / 32; x++) { const uint32_t v = s_warp[x]; if (x < (t >> 5)) wbase += v; tot += v; }
            slot = wbase + incl - cnt;
            if (t == 0) s_nrec = tot;
        }
        // start of the run my position 0 continues (the previous thread's bit 0 says whether it does): the nearest reset before it
        uint32_t cur_start = (uint32_t)t * 32u;
        if (t > 0 && (s_rs[t - 1] & 1u)) {
            for (int u = t - 1;; u--) {
                const uint32_t word = (s_rs[u] & ~1u) | (u == 0 ? 1u : (~s_rs[u - 1] & 1u));
                if (word) { cur_start = (uint32_t)u * 32u + (31u - (uint32_t)__clz((int)word)); break; }
            }
        }
#pragma unroll
        for (int i = 0; i < 32; i++) {
            if (i > 0 && ((resets_hi >> i) & 1u)) cur_start = (uint32_t)t * 32u + i;
            if ((ends >> i) & 1u) {
                const uint32_t len = (uint32_t)t * 32u + i - cur_start + 1u;
                if (slot < (uint32_t)SK1_LIST_CAP) s_list[slot] = ((uint64_t)mr[i] << 32) | ((uint64_t)cur_start << 14) | len;
                else sk1_direct<MODE>(p, s_pk, mr[i], cur_start, len);
                slot++;
            }
        }
        __syncthreads();
        // ---- F: every listed run gets its coarse bin and the rank of its first record inside the tile's bin
        const uint32_t nrec = min(s_nrec, (uint32_t)SK1_LIST_CAP);
        uint32_t cr[SK1_R];                                        // rank << 11 | coarse bin
#pragma unroll
        for (int r = 0; r < SK1_R; r++) {
            const uint32_t j = (uint32_t)t + (uint32_t)r * K1_THREADS;
            cr[r] = 0xFFFFFFFFu;
            if (j < nrec) {
                const uint64_t d = s_list[j];
                uint32_t c, sub;
                if (sk_bins((uint32_t)(d >> 32), p.Pc, p.V, c, sub) == p.wave) {
                    const uint32_t len = (uint32_t)d & 0x3FFFu, pieces = (len + nmax - 1) / nmax;
                    cr[r] = (atomicAdd(&s_cnt[c], pieces) << 11) | c;
                    my_runs += pieces;
                }
            }
        }
        __syncthreads();
        if (MODE == 0) {
            for (uint32_t i = t; i < p.Pc; i += K1_THREADS) { const uint32_t c = s_cnt[i]; s_cnt[i] = 0; if (c) atomicAdd(&p.bin_cursor[i], (unsigned long long)c); }
            __syncthreads();
            continue;
        }
        // ---- G: exclusive scan of the tile's bin counts + one global reservation per non-empty bin
        {
            const uint32_t per = (p.Pc + K1_THREADS - 1) / K1_THREADS;
            uint32_t b0 = t * per, sum = 0;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.Pc) sum += s_cnt[b0 + j];
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += n; }
            if (lane == 31) s_warp[t >> 5] = incl;
            __syncthreads();
            uint32_t wbase = 0;
#pragma unroll
            for (int x = 0; x < K1_THREADS / 32; x++) if (x < (t >> 5)) wbase += s_warp[x];
            uint32_t run = wbase + incl - sum;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.Pc) {
                const uint32_t cn = s_cnt[b0 + j];
                s_off[b0 + j] = run; s_cnt[b0 + j] = 0;
                if (cn) {
                    const unsigned long long old = atomicAdd(&p.bin_cursor[b0 + j], (unsigned long long)cn);
                    if (old + cn > p.bin_cap[b0 + j]) { *p.overflow = 1; s_gbase[b0 + j] = SK1_SKIP; }
                    else s_gbase[b0 + j] = p.bin_base[b0 + j] + old;
                }
                run += cn;
            }
            if (t == K1_THREADS - 1) s_off[p.Pc] = run;
        }
        __syncthreads();
        // ---- H: build the records, counting-sorted by bin in the staging area (what does not fit goes out directly)
#pragma unroll
        for (int r = 0; r < SK1_R; r++) {
            if (cr[r] == 0xFFFFFFFFu) continue;
            const uint64_t d = s_list[(uint32_t)t + (uint32_t)r * K1_THREADS];
            const uint32_t c = cr[r] & 2047u, rank = cr[r] >> 11;
            uint32_t c2, sub;
            sk_bins((uint32_t)(d >> 32), p.Pc, p.V, c2, sub);
            const uint32_t start = ((uint32_t)d >> 14) & 0x3FFFu, len = (uint32_t)d & 0x3FFFu;
            const uint64_t g = s_gbase[c];
            uint32_t pos = s_off[c] + rank;
            for (uint32_t o = 0; o < len; o += nmax, pos++) {
                const ulonglong2 rec = sk_make_record(s_pk, start + o, min(nmax, len - o), k, sub);
                if (pos < (uint32_t)SK1_STAGE_CAP) { s_stage[pos] = rec; s_sb[pos] = (uint16_t)c; }
                else if (g != SK1_SKIP) p.out[g + (pos - s_off[c])] = rec;
            }
        }
        __syncthreads();
        {
            const uint32_t total = min(s_off[p.Pc], (uint32_t)SK1_STAGE_CAP);
            for (uint32_t j = t; j < total; j += K1_THREADS) {
                const uint32_t b = s_sb[j];
                const uint64_t g = s_gbase[b];
                if (g != SK1_SKIP) p.out[g + (j - s_off[b])] = s_stage[j];
            }
        }
        __syncthreads();
    }
    block_add(p.n_kmers, my_kmers);
    block_add(p.n_runs, my_runs);
}

static inline size_t sk1_smem_bytes(uint32_t Pc) {
    return 2 * (size_t)K1T_RAW + 16 + (size_t)(K1_THREADS + 4) * 8 + (size_t)((Pc + 1u) & ~1u) * 8 + SK1_X_BYTES + (size_t)(K1_THREADS + 2) * 4 + (size_t)K1_THREADS * 4 +
           (size_t)Pc * 4 + (size_t)(Pc + 1) * 4 + 128;
}

// gather the used parts of the coarse bins' regions into one contiguous array (sharded mode: the per-destination segments of the
// send buffer must be contiguous).  out_off[Pc + 1] = exclusive prefix of the bin counts.
__global__ void __launch_bounds__(256) sk_compact(const ulonglong2 *__restrict__ in, const uint64_t *__restrict__ bin_base, const uint64_t *__restrict__ out_off,
                                                  uint32_t Pc, ulonglong2 *out) {
    const uint64_t n = out_off[Pc];
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = Pc;       // invariant: out_off[lo] <= i < out_off[hi]
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (out_off[mid] <= i) lo = mid; else hi = mid; }
        out[i] = in[bin_base[lo] + (i - out_off[lo])];
    }
}

// ------------------------------------------------------------------------------------------------ K2a: fine partition of the records

constexpr int SK2P_THREADS = 512;
constexpr int SK2P_PER_THREAD = 8;
constexpr int SK2P_TILE = SK2P_THREADS * SK2P_PER_THREAD;   // 4096 records (64 KB) per tile
constexpr int SK2P_MAX_F = 1024;

struct SK2PParams {
    const ulonglong2 *recs;         // K1 output (or the receive buffer of X1): coarse bin / run b = [bin_beg[b], bin_end[b])
    const uint64_t *bin_beg, *bin_end;
    const uint32_t *tile_first;     // [P1 + 1] first tile of each coarse bin (tiles never span coarse bins)
    const uint32_t *group;          // [P1] or null: sharded mode, the runs (one per source GPU) of one local bin share their fine bins
    uint32_t P1, F, n_tiles;
    int canon, k;                   // canon: write every record in its strand-canonical form (sk_canon_rec), what sk2d_count's dedup needs
    uint32_t *fine_cnt;             // [G * F]     hist: out
    const uint64_t *fine_off;       // [G * F + 1] scatter: exclusive offsets into `out`
    uint32_t *fine_cursor;          // [G * F]     scatter: running cursors, zero-initialised
    ulonglong2 *out;
};

__global__ void __launch_bounds__(SK2P_THREADS) sk2p_hist(SK2PParams p) {
    __shared__ uint32_t s_cnt[SK2P_MAX_F];
    const uint32_t t = threadIdx.x;
    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t b = k2p_bin_of_tile(p.tile_first, p.P1, tile);
        const uint64_t beg = p.bin_beg[b] + (uint64_t)(tile - p.tile_first[b]) * SK2P_TILE;
        const uint64_t end = min(beg + (uint64_t)SK2P_TILE, p.bin_end[b]);
        for (uint32_t i = t; i < p.F; i += SK2P_THREADS) s_cnt[i] = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SK2P_PER_THREAD; j++) {
            const uint64_t i = beg + (uint64_t)j * SK2P_THREADS + t;
            if (i < end) atomicAdd(&s_cnt[sk_fine(((uint32_t)p.recs[i].y >> 6) & ((1u << SK_SUB_BITS) - 1u), p.F)], 1u);
        }
        __syncthreads();
        for (uint32_t i = t; i < p.F; i += SK2P_THREADS) {
            const uint32_t c = s_cnt[i];
            if (c) atomicAdd(&p.fine_cnt[(size_t)(p.group ? p.group[b] : b) * p.F + i], c);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SK2P_THREADS) sk2p_scatter(SK2PParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ulonglong2 *s_stage = reinterpret_cast<ulonglong2 *>(smem_raw);              // [SK2P_TILE] the tile sorted by fine bin
    uint16_t *s_sb = reinterpret_cast<uint16_t *>(s_stage + SK2P_TILE);          // [SK2P_TILE]
    __shared__ uint64_t s_gbase[SK2P_MAX_F];
    __shared__ uint32_t s_cnt[SK2P_MAX_F];
    __shared__ uint32_t s_off[SK2P_MAX_F + 1];
    __shared__ uint32_t s_warp[SK2P_THREADS / 32];
    const uint32_t t = threadIdx.x, lane = lane_id();
    const uint32_t per = (p.F + SK2P_THREADS - 1) / SK2P_THREADS;
    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint32_t b = k2p_bin_of_tile(p.tile_first, p.P1, tile);
        const uint64_t beg = p.bin_beg[b] + (uint64_t)(tile - p.tile_first[b]) * SK2P_TILE;
        const uint64_t end = min(beg + (uint64_t)SK2P_TILE, p.bin_end[b]);
        for (uint32_t i = t; i < p.F; i += SK2P_THREADS) s_cnt[i] = 0;
        __syncthreads();
        ulonglong2 rec[SK2P_PER_THREAD];
        uint32_t fr[SK2P_PER_THREAD];                                            // rank << 10 | fine bin
#pragma unroll
        for (int j = 0; j < SK2P_PER_THREAD; j++) {
            const uint64_t i = beg + (uint64_t)j * SK2P_THREADS + t;
            fr[j] = 0xFFFFFFFFu;
            if (i < end) { rec[j] = p.recs[i]; if (p.canon) rec[j] = sk_canon_rec(rec[j], p.k); }
            else rec[j] = make_ulonglong2(0, 0);
        }
#pragma unroll
        for (int j = 0; j < SK2P_PER_THREAD; j++) {
            const uint64_t i = beg + (uint64_t)j * SK2P_THREADS + t;
            if (i < end) {
                const uint32_t f = sk_fine(sk_rec_sub(rec[j]), p.F);
                fr[j] = (atomicAdd(&s_cnt[f], 1u) << 10) | f;
            }
        }
        __syncthreads();
        {
            const uint32_t b0 = t * per;
            uint32_t sum = 0;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.F) sum += s_cnt[b0 + j];
            uint32_t incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += n; }
            if (lane == 31) s_warp[t >> 5] = incl;
            __syncthreads();
            uint32_t wbase = 0;
#pragma unroll
            for (int x = 0; x < SK2P_THREADS / 32; x++) if (x < (int)(t >> 5)) wbase += s_warp[x];
            uint32_t run = wbase + incl - sum;
            for (uint32_t j = 0; j < per; j++) if (b0 + j < p.F) {
                const uint32_t cn = s_cnt[b0 + j];
                s_off[b0 + j] = run;
                if (cn) {
                    const size_t g = (size_t)(p.group ? p.group[b] : b) * p.F + b0 + j;
                    s_gbase[b0 + j] = p.fine_off[g] + atomicAdd(&p.fine_cursor[g], cn);
                }
                run += cn;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SK2P_PER_THREAD; j++) {
            if (fr[j] == 0xFFFFFFFFu) continue;
            const uint32_t f = fr[j] & 1023u, pos = s_off[f] + (fr[j] >> 10);
            s_stage[pos] = rec[j]; s_sb[pos] = (uint16_t)f;
        }
        __syncthreads();
        const uint32_t total = (uint32_t)(end - beg);
        for (uint32_t j = t; j < total; j += SK2P_THREADS) {
            const uint32_t f = s_sb[j];
            p.out[s_gbase[f] + (j - s_off[f])] = s_stage[j];
        }
        __syncthreads();
    }
}

static inline size_t sk2p_scatter_smem_bytes() { return (size_t)SK2P_TILE * 16 + (size_t)SK2P_TILE * 2 + 128; }

// ------------------------------------------------------------------------------------------------ K2b: counting in shared memory

struct SK2CParams {
    const ulonglong2 *recs;         // the records ordered by fine bin
    const uint64_t *fine_off;       // [n_fine + 1]
    uint32_t n_fine;
    int k;
    uint32_t amin, amax;
    uint64_t *solid_kmers; uint32_t *solid_counts; uint64_t solid_cap;
    unsigned long long *n_solid, *n_distinct, *n_splits;
    int *failed;
};

// k-mers a warp expands per round into its private staging area: 256 (2 KB per warp); 128 for the 512-thread CTAs over a 4096-slot table,
// so that three of them (48 warps) still fit one SM's shared memory
// (SMALL: 128 also for 256-thread CTAs -- 56 KB per CTA, four CTAs per SM instead of three; CDBG_SK_WCAP=128)
template <int SLOTS, int T, bool SMALL = false> constexpr int sk2c_wcap() { return (SMALL || (T >= 512 && SLOTS <= 4096)) ? 128 : 256; }
template <int SLOTS, int T, bool SMALL = false> static inline size_t sk2c_smem_bytes() { return (size_t)SLOTS * 12 + (size_t)(T / 32) * sk2c_wcap<SLOTS, T, SMALL>() * 8 + 128; }

// One CTA per fine bin (grid-strided), the table of k2c_count (keys u64 + counts u32 in shared memory, linear probing, abandon at
// 3/4 full and redo the bin as two halves selected by one more hash bit).  The difference is the input: every warp takes 32 records
// at a time, expands as many of them as fit its private staging area into canonical k-mers (one record per lane, rolling), and then
// inserts the staged k-mers with all lanes busy -- the k-mers of a bucket exist in shared memory only.
// KC: the k-mer size as a compile-time constant (0 = take it from the parameters): the rolling k-mers are all shifts by k.
// T threads per CTA.  The warps of a CTA take batches of 32 records from a shared cursor, so a warp that drew expensive batches
// (long probe sequences, many k-mers per record) is not waited for by warps that would otherwise hold a fixed share.
// TWO: two-phase insert.  In the plain loop every lane probes until ITS key is placed, so a warp runs as many iterations as its
// unluckiest lane (ncu: ~11 of 32 lanes active in the probe loop).  With TWO every staged k-mer gets exactly ONE probe with the whole
// warp converged; the few keys that met a different key there (~the load factor) are written back to the front of the staging
// area -- already consumed -- and finish in a second, short pass.
template <int SLOTS, int T, bool MUL32, int KC, bool TWO = false, bool SMALL = false>
__global__ void __launch_bounds__(T) sk2c_count(SK2CParams p) {
    constexpr int NW = T / 32, PER_WARP = SLOTS / NW, SK2C_WCAP = (SMALL || (T >= 512 && SLOTS <= 4096)) ? 128 : 256;   // (= sk2c_wcap<SLOTS, T, SMALL>())
    constexpr uint32_t MASK = SLOTS - 1, FILL_LIMIT = SLOTS / 4 * 3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(smem_raw);
    uint64_t *s_stage = s_keys + SLOTS;                                          // [NW][SK2C_WCAP]
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_stage + NW * SK2C_WCAP);
    __shared__ uint32_t s_wsolid[32];
    __shared__ unsigned long long s_base;
    __shared__ int s_over;
    __shared__ uint32_t s_fill;
    __shared__ unsigned long long s_next;                              // next unclaimed record of the current pass
    const uint32_t t = threadIdx.x, lane = lane_id(), warp = t >> 5;
    const uint32_t wb = warp * PER_WARP;
    uint64_t *my_stage = s_stage + warp * SK2C_WCAP;
    const int k = KC ? KC : p.k;
    unsigned long long my_distinct = 0, my_splits = 0;

    for (uint32_t bin = blockIdx.x; bin < p.n_fine; bin += gridDim.x) {
        const uint64_t beg = p.fine_off[bin], end = p.fine_off[bin + 1];
        if (beg == end) continue;                                    // (uniform over the block)
        uint32_t stk[K2C_MAX_LEVEL + 2];
        int sp = 0;
        stk[sp++] = 0;
        while (sp) {
            const uint32_t item = stk[--sp], level = item >> 16, part = item & 0xFFFFu, sel_mask = (1u << level) - 1u;
#pragma unroll 4
            for (uint32_t i = wb + lane; i < wb + PER_WARP; i += 32) { s_keys[i] = EMPTY_KEY; s_cnt[i] = 0; }
            if (t == 0) { s_over = 0; s_fill = 0; s_next = beg + (unsigned long long)NW * 32; }
            __syncthreads();

            for (uint64_t base = beg + (uint64_t)warp * 32; base < end;) {   // first batch by warp index, then from the shared cursor
                if (__any_sync(0xFFFFFFFFu, *(volatile int *)&s_over)) break;   // table (nearly) full: this pass is abandoned
                const uint64_t i = base + lane;
                ulonglong2 rec = make_ulonglong2(0, 0);
                uint32_t n = 0;
                if (i < end) { rec = p.recs[i]; n = sk_rec_n(rec); }
                uint32_t incl = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += v; }
                uint32_t done = 0;                                   // k-mers of this warp's 32 records already inserted
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                uint32_t fresh = 0;
                while (done < total) {
                    // the lanes whose records fit the staging area this round: a contiguous range (incl is monotone)
                    const bool mine = n && incl > done && incl - done <= (uint32_t)SK2C_WCAP;
                    const uint32_t round_end = __reduce_max_sync(0xFFFFFFFFu, mine ? incl : done);
                    if (mine) {
                        uint64_t hi = rec.x, lo = rec.y & ~0xFFFFFFULL;
                        uint64_t fwd = hi >> (64 - 2 * k);
                        uint64_t rc = revcomp64(fwd, k);
                        uint32_t o = incl - n - done;
                        my_stage[o] = fwd < rc ? fwd : rc;
                        for (uint32_t j = 1; j < n; j++) {
                            hi = (hi << 2) | (lo >> 62); lo <<= 2;
                            fwd = hi >> (64 - 2 * k);
                            rc = (rc >> 2) | (((fwd & 3) ^ 2) << (2 * (k - 1)));
                            my_stage[o + j] = fwd < rc ? fwd : rc;
                        }
                    }
                    __syncwarp();
                    const uint32_t cnt = round_end - done;
                    if (TWO) {
                        uint32_t n_def = 0;                          // deferred keys so far (warp-uniform), stored at my_stage[0 .. n_def)
                        for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {  // warp-uniform trip count
                            const uint32_t j = j0 + lane;
                            bool defer = false;
                            uint64_t key = 0;
                            if (j < cnt) {
                                key = my_stage[j];
                                const uint32_t h = k2c_hash<MUL32>(key);
                                if (((h >> 16) & sel_mask) == part) {
                                    const uint32_t slot = h & MASK;
                                    uint64_t cur = ld_volatile_u64(&s_keys[slot]);
                                    if (cur == EMPTY_KEY) {
                                        cur = atomicCAS((unsigned long long *)&s_keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
                                        if (cur == EMPTY_KEY) { cur = key; fresh++; }
                                    }
                                    if (cur == key) atomicAdd(&s_cnt[slot], 1u);
                                    else defer = true;
                                }
                            }
                            const unsigned dm = __ballot_sync(0xFFFFFFFFu, defer);   // (also orders this round's reads before its writes)
                            if (defer) my_stage[n_def + __popc(dm & ((1u << lane) - 1u))] = key;   // n_def + 32 <= j0 + 32: never past this round's keys
                            n_def += __popc(dm);
                            __syncwarp();
                        }
                        for (uint32_t j = lane; j < n_def; j += 32) {
                            const uint64_t key = my_stage[j];
                            uint32_t slot = (k2c_hash<MUL32>(key) + 1u) & MASK;
                            for (uint32_t probes = 1;;) {
                                uint64_t cur = ld_volatile_u64(&s_keys[slot]);
                                if (cur == EMPTY_KEY) {
                                    cur = atomicCAS((unsigned long long *)&s_keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
                                    if (cur == EMPTY_KEY) { cur = key; fresh++; }
                                }
                                if (cur == key) { atomicAdd(&s_cnt[slot], 1u); break; }
                                slot = (slot + 1) & MASK;
                                if (++probes >= (uint32_t)SLOTS) { s_over = 1; break; }
                            }
                        }
                    } else {
                    for (uint32_t j = lane; j < cnt; j += 32) {
                        const uint64_t key = my_stage[j];
                        const uint32_t h = k2c_hash<MUL32>(key);
                        if (((h >> 16) & sel_mask) != part) continue;
                        uint32_t slot = h & MASK;
                        for (uint32_t probes = 0;;) {
                            uint64_t cur = ld_volatile_u64(&s_keys[slot]);
                            if (cur == EMPTY_KEY) {
                                cur = atomicCAS((unsigned long long *)&s_keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
                                if (cur == EMPTY_KEY) { cur = key; fresh++; }
                            }
                            if (cur == key) { atomicAdd(&s_cnt[slot], 1u); break; }
                            slot = (slot + 1) & MASK;
                            if (++probes >= (uint32_t)SLOTS) { s_over = 1; break; }
                        }
                    }
                    }
                    __syncwarp();
                    done = round_end;
                }
                fresh = __reduce_add_sync(0xFFFFFFFFu, fresh);
                if (lane == 0 && fresh && atomicAdd(&s_fill, fresh) + fresh > FILL_LIMIT) s_over = 1;
                unsigned long long nb = 0;
                if (lane == 0) nb = atomicAdd(&s_next, 32ULL);
                base = __shfl_sync(0xFFFFFFFFu, nb, 0);
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
                for (int x = 0; x < NW; x++) { uint32_t v = s_wsolid[x]; s_wsolid[x] = tot; tot += v; }
                s_base = tot ? atomicAdd(p.n_solid, (unsigned long long)tot) : 0ULL;
            }
            __syncthreads();
            uint64_t out = s_base + s_wsolid[warp];
            if (my_solid) {                                          // (uniform over the warp after the reduction)
                for (uint32_t i = wb + lane; i < wb + PER_WARP; i += 32) {
                    const uint64_t key = s_keys[i];
                    const uint32_t c = s_cnt[i];
                    const bool solid = key != EMPTY_KEY && c >= p.amin && c <= p.amax;
                    const unsigned mk = __ballot_sync(0xFFFFFFFFu, solid);
                    if (solid) {
                        const uint64_t pos = out + __popc(mk & ((1u << lane) - 1));
                        if (pos < p.solid_cap) { p.solid_kmers[pos] = key; p.solid_counts[pos] = c; }
                    }
                    out += __popc(mk);
                }
            }
            __syncthreads();     // the next pass clears the table and thread 0 resets s_over / s_fill
        }
    }
    block_add(p.n_distinct, my_distinct);
    block_add(p.n_splits, my_splits);
}

// ------------------------------------------------------------------------------------------------ K2b': counting with record dedup
//
// Sequencing reads cover every genome position many times, so most super-k-mers of a bucket are exact copies of one another (at 30x
// about two thirds of the records: the ones without a sequencing error that are not cut by a read end).  sk2d_count therefore counts
// the distinct RECORDS of a fine bin first and expands each distinct record once, adding its multiplicity to its k-mers' counts:
//   B  records -> record table (RS slots: 64-bit fingerprint of the 16 record bytes, multiplicity, index of the first such record);
//      after a barrier every record compares its 16 bytes with the table's representative -- a record whose fingerprint collides with
//      a different record, or that finds the record table full, is expanded directly with weight 1 (exactness never rests on the hash);
//   C  distinct records -> k-mer table: a warp takes 32 record slots, prefix-sums their k-mer counts and expands ONE K-MER PER LANE
//      (record found by a 5-step search of the prefix sums), so expansion, hashing and the first probe run with full warps;
//   D  emit (as sk2c_count).
// Records must be strand-canonical for copies read from opposite strands to be equal: sk2p_scatter does that (sk_canon_rec) when
// its `canon` flag is set.  Same table overflow handling as sk2c_count (abandon at 3/4 full, redo as two halves by one more hash bit).

__device__ __forceinline__ uint64_t sk_rec_fp(const ulonglong2 r) {
    uint64_t h = mix64(r.x ^ (r.y * 0x9E3779B97F4A7C15ULL)) ^ r.y;
    return h == EMPTY_KEY ? 0 : h;
}

template <int SLOTS, int RS, int T> static inline size_t sk2d_smem_bytes() { return (size_t)SLOTS * 12 + (size_t)RS * 16 + (size_t)(T / 32) * (32 * 16 + 32 * 8) + 128; }

template <int SLOTS, int RS, int T, bool MUL32, int KC>
__global__ void __launch_bounds__(T) sk2d_count(SK2CParams p) {
    constexpr int NW = T / 32, PER_WARP = SLOTS / NW, R_PER = 4;
    constexpr uint32_t MASK = SLOTS - 1, FILL_LIMIT = SLOTS / 4 * 3, RMASK = RS - 1, RFILL_LIMIT = RS / 4 * 3, NONE = 0xFFFFFFFFu;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(smem_raw);                   // [SLOTS] k-mer table
    uint64_t *r_fp = s_keys + SLOTS;                                             // [RS]    record table: fingerprints
    ulonglong2 *w_rec = reinterpret_cast<ulonglong2 *>(r_fp + RS);               // [NW][32] the records a warp is expanding
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(w_rec + NW * 32);             // [SLOTS]
    uint32_t *r_cnt = s_cnt + SLOTS;                                             // [RS]    multiplicities
    uint32_t *r_idx = r_cnt + RS;                                                // [RS]    representative (index from the bin's first record)
    uint32_t *w_incl = r_idx + RS;                                               // [NW][32] inclusive prefix sums of k-mer counts
    uint32_t *w_wt = w_incl + NW * 32;                                           // [NW][32] weights
    __shared__ uint32_t s_wsolid[32];
    __shared__ unsigned long long s_base;
    __shared__ int s_over, s_rfull;
    __shared__ uint32_t s_fill, s_rfresh;
    const uint32_t t = threadIdx.x, lane = lane_id(), warp = t >> 5;
    const uint32_t wb = warp * PER_WARP;
    ulonglong2 *my_rec = w_rec + warp * 32;
    uint32_t *my_incl = w_incl + warp * 32, *my_wt = w_wt + warp * 32;
    const int k = KC ? KC : p.k;
    unsigned long long my_distinct = 0, my_splits = 0;

    for (uint32_t bin = blockIdx.x; bin < p.n_fine; bin += gridDim.x) {
        const uint64_t beg = p.fine_off[bin], end = p.fine_off[bin + 1];
        if (beg == end) continue;                                    // (uniform over the block)
        uint32_t stk[K2C_MAX_LEVEL + 2];
        int sp = 0;
        stk[sp++] = 0;
        while (sp) {
            const uint32_t item = stk[--sp], level = item >> 16, part = item & 0xFFFFu, sel_mask = (1u << level) - 1u;
#pragma unroll 4
            for (uint32_t i = wb + lane; i < wb + PER_WARP; i += 32) { s_keys[i] = EMPTY_KEY; s_cnt[i] = 0; }
            for (uint32_t i = t; i < (uint32_t)RS; i += T) { r_fp[i] = EMPTY_KEY; r_cnt[i] = 0; }
            if (t == 0) { s_over = 0; s_fill = 0; s_rfull = 0; s_rfresh = 0; }
            __syncthreads();
            uint32_t fresh = 0;

            // insert one k-mer with weight w (probe until placed; the caller has applied the part selection)
            auto insert = [&](uint64_t key, uint32_t h, uint32_t w) {
                uint32_t slot = h & MASK;
                for (uint32_t probes = 0;;) {
                    uint64_t cur = ld_volatile_u64(&s_keys[slot]);
                    if (cur == EMPTY_KEY) {
                        cur = atomicCAS((unsigned long long *)&s_keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
                        if (cur == EMPTY_KEY) { cur = key; fresh++; }
                    }
                    if (cur == key) { atomicAdd(&s_cnt[slot], w); break; }
                    slot = (slot + 1) & MASK;
                    if (++probes >= (uint32_t)SLOTS) { s_over = 1; break; }
                }
            };

            // ---- B: records -> record table, T * R_PER records at a time
            for (uint64_t c0 = beg; c0 < end; c0 += (uint64_t)T * R_PER) {
                ulonglong2 rec[R_PER];
                uint32_t mys[R_PER];
#pragma unroll
                for (int r = 0; r < R_PER; r++) {
                    const uint64_t i = c0 + (uint64_t)r * T + t;
                    mys[r] = NONE;
                    rec[r] = make_ulonglong2(0, 0);
                    if (i < end) {
                        rec[r] = p.recs[i];
                        const uint64_t fp = sk_rec_fp(rec[r]);
                        uint32_t slot = (uint32_t)(fp >> 32) & RMASK;
                        for (uint32_t probes = 0; probes < (uint32_t)RS; probes++) {
                            uint64_t cur = ld_volatile_u64(&r_fp[slot]);
                            if (cur == EMPTY_KEY) {
                                if (*(volatile int *)&s_rfull) break;                    // no new entries: this record is expanded directly
                                cur = atomicCAS((unsigned long long *)&r_fp[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)fp);
                                if (cur == EMPTY_KEY) {
                                    cur = fp;
                                    r_idx[slot] = (uint32_t)(i - beg);
                                    if (atomicAdd(&s_rfresh, 1u) + 1u >= RFILL_LIMIT) s_rfull = 1;
                                }
                            }
                            if (cur == fp) { atomicAdd(&r_cnt[slot], 1u); mys[r] = slot; break; }
                            slot = (slot + 1) & RMASK;
                        }
                    }
                }
                __syncthreads();                                     // the representatives' indices are published
#pragma unroll
                for (int r = 0; r < R_PER; r++) {
                    const uint64_t i = c0 + (uint64_t)r * T + t;
                    if (i >= end) continue;
                    if (mys[r] != NONE) {
                        const ulonglong2 rep = p.recs[beg + r_idx[mys[r]]];
                        if (rep.x == rec[r].x && rep.y == rec[r].y) continue;
                        atomicSub(&r_cnt[mys[r]], 1u);               // a different record with the same fingerprint
                    }
                    // direct expansion, weight 1 (rare: fingerprint collision or record table full)
                    if (*(volatile int *)&s_over) continue;
                    const uint32_t n = sk_rec_n(rec[r]);
                    uint64_t hi = rec[r].x, lo = rec[r].y & ~0xFFFFFFULL;
                    uint64_t fwd = hi >> (64 - 2 * k);
                    uint64_t rc = revcomp64(fwd, k);
                    for (uint32_t j = 0; j < n; j++) {
                        if (j) {
                            hi = (hi << 2) | (lo >> 62); lo <<= 2;
                            fwd = hi >> (64 - 2 * k);
                            rc = (rc >> 2) | (((fwd & 3) ^ 2) << (2 * (k - 1)));
                        }
                        const uint64_t key = fwd < rc ? fwd : rc;
                        const uint32_t h = k2c_hash<MUL32>(key);
                        if (((h >> 16) & sel_mask) == part) insert(key, h, 1u);
                    }
                }
            }
            __syncthreads();

            // ---- C: distinct records -> k-mer table, one k-mer per lane
            for (uint32_t s0 = warp * 32; s0 < (uint32_t)RS; s0 += NW * 32) {
                if (__any_sync(0xFFFFFFFFu, *(volatile int *)&s_over)) break;
                const uint32_t s = s0 + lane;
                const uint32_t wgt = r_fp[s] != EMPTY_KEY ? r_cnt[s] : 0u;
                ulonglong2 rec = make_ulonglong2(0, 0);
                uint32_t n = 0;
                if (wgt) { rec = p.recs[beg + r_idx[s]]; n = sk_rec_n(rec); }
                uint32_t incl = n;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += v; }
                const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
                if (!total) continue;
                my_rec[lane] = rec; my_incl[lane] = incl; my_wt[lane] = wgt;
                __syncwarp();
                for (uint32_t g0 = 0; g0 < total; g0 += 32) {
                    const uint32_t g = g0 + lane;
                    if (g < total) {
                        uint32_t r = 0;
                        if (my_incl[r + 15] <= g) r += 16;
                        if (my_incl[r + 7] <= g) r += 8;
                        if (my_incl[r + 3] <= g) r += 4;
                        if (my_incl[r + 1] <= g) r += 2;
                        if (my_incl[r] <= g) r += 1;
                        const ulonglong2 rr = my_rec[r];
                        const uint32_t j = g - (my_incl[r] - sk_rec_n(rr));
                        const uint64_t ylo = rr.y & ~0xFFFFFFULL;
                        const uint64_t hi = j ? (rr.x << (2 * j)) | (ylo >> (64 - 2 * j)) : rr.x;
                        const uint64_t fwd = hi >> (64 - 2 * k);
                        const uint64_t rc = revcomp64(fwd, k);
                        const uint64_t key = fwd < rc ? fwd : rc;
                        const uint32_t h = k2c_hash<MUL32>(key);
                        if (((h >> 16) & sel_mask) == part) insert(key, h, my_wt[r]);
                    }
                }
                __syncwarp();
                fresh = __reduce_add_sync(0xFFFFFFFFu, fresh);
                if (lane == 0 && fresh && atomicAdd(&s_fill, fresh) + fresh > FILL_LIMIT) s_over = 1;
                fresh = 0;
            }
            fresh = __reduce_add_sync(0xFFFFFFFFu, fresh);           // (keys placed by direct expansion in B)
            if (lane == 0 && fresh && atomicAdd(&s_fill, fresh) + fresh > FILL_LIMIT) s_over = 1;
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

            // ---- D: emit: count the solid entries, reserve the output range with ONE global atomic, write ballot-ranked
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
                for (int x = 0; x < NW; x++) { uint32_t v = s_wsolid[x]; s_wsolid[x] = tot; tot += v; }
                s_base = tot ? atomicAdd(p.n_solid, (unsigned long long)tot) : 0ULL;
            }
            __syncthreads();
            uint64_t out = s_base + s_wsolid[warp];
            if (my_solid) {                                          // (uniform over the warp after the reduction)
                for (uint32_t i = wb + lane; i < wb + PER_WARP; i += 32) {
                    const uint64_t key = s_keys[i];
                    const uint32_t c = s_cnt[i];
                    const bool solid = key != EMPTY_KEY && c >= p.amin && c <= p.amax;
                    const unsigned mk = __ballot_sync(0xFFFFFFFFu, solid);
                    if (solid) {
                        const uint64_t pos = out + __popc(mk & ((1u << lane) - 1));
                        if (pos < p.solid_cap) { p.solid_kmers[pos] = key; p.solid_counts[pos] = c; }
                    }
                    out += __popc(mk);
                }
            }
            __syncthreads();     // the next pass clears the tables and thread 0 resets the flags
        }
    }
    block_add(p.n_distinct, my_distinct);
    block_add(p.n_splits, my_splits);
}
