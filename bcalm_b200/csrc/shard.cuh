// shard.cuh -- kernels of the minimizer-sharded multi-GPU path.
//
// Ownership: a canonical k-mer lives on GPU  owner(minimizer_k(x));  a junction (canonical (k-1)-mer c) is decided on GPU
// owner(minimizer_{k-1}(c)).  minimizer = the canonical m-mer of smallest mmer_rank (kmer.cuh; the reference's structure,
// Model.hpp:1040-1064,1254-1287, with a hash order instead of the sampled frequency order -- SURVEY N1).  Consecutive k-mers
// of a read share their minimizer with probability ~1-2/(k-m+2), so ~90% of the graph's adjacencies stay on one GPU: each
// GPU compacts its shard into local unitigs ("pieces", stage 2) and the cross-GPU stitching ("glue", stage 3) only sees the
// piece ends.  Global ids: gid = rank << 32 | local end id.
#pragma once
#include "nccl_dyn.h"
#include "kernels.cuh"

#define GID_RANK(g) ((uint32_t)(((g) >> 32) & 0xFFu))
#define GID_LOCAL(g) ((uint32_t)(g))
#define GID_MASK 0x000000FFFFFFFFFFULL
#define NOGID 0xFFFFFFFFFFFFFFFFULL

static CDBG_HD uint32_t owner_of_rank(uint32_t minrank, uint32_t world) {
    uint32_t h = minrank ^ 0x68E31DA4u;
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return (uint32_t)(((uint64_t)h * world) >> 32);
}

// smallest canonical m-mer rank of the n-mer x (straightforward form: used per junction record, not per read position)
static CDBG_HD uint32_t min_rank_of(uint64_t x, int n, int m) {
    const uint64_t mm = kmer_mask(m);
    const uint64_t rx = revcomp64(x, n);      // the reverse complement of the m-mer at offset j is the m-mer of rx at offset n-m-j
    uint32_t best = 0xFFFFFFFFu;
    for (int j = 0; j + m <= n; j++) {
        uint32_t f = (uint32_t)((x >> (2 * (n - m - j))) & mm), r = (uint32_t)((rx >> (2 * j)) & mm);
        uint32_t rk = mmer_rank(f < r ? f : r);
        best = rk < best ? rk : best;
    }
    return best;
}

// ------------------------------------------------------------------------------------------------ K1 (sharded)

constexpr int K1S_THREADS = 256;
constexpr int K1S_TILE = 2048;                              // positions per tile, lane-consecutive (8 per thread)
constexpr int K1S_WORDS = K1S_TILE / 32 + 2;                // packed words incl. 64 halo positions
constexpr int K1S_RANKS = K1S_TILE + 32;                    // m-mer ranks incl. halo (window <= 32)

struct K1SParams {
    const uint8_t *bytes; uint64_t n_bytes; uint64_t n_tiles;
    int k, m; uint32_t world, P_local;                      // P = world * P_local bins, destination-major
    unsigned long long *bin_count; const uint64_t *bin_off; unsigned long long *bin_cursor;
    uint64_t *out_kmers;
    uint16_t *pos_bin;                                      // [n_tiles * K1S_TILE] bin of the k-mer starting at each position (0xFFFF = none):
                                                            // written by the histogram pass, read by the scatter pass
};

__device__ __forceinline__ uint64_t funnel_left(uint64_t hi, uint64_t lo, int bits) {   // (hi:lo) << bits, top 64 bits; bits in [0, 63]
    return bits ? (hi << bits) | (lo >> (64 - bits)) : hi;
}

template <int MODE>
__global__ void __launch_bounds__(K1S_THREADS) k1s_extract(K1SParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t P = p.world * p.P_local;
    uint64_t *s_pk = reinterpret_cast<uint64_t *>(smem_raw);                   // [K1S_WORDS]
    uint64_t *s_gbase = s_pk + K1S_WORDS;                                      // [P]        MODE 1
    uint64_t *s_c = s_gbase + (MODE == 1 ? P : 0);                             // [K1S_TILE] MODE 1: k-mers by position
    uint64_t *s_stage = s_c + (MODE == 1 ? K1S_TILE : 0);                      // [K1S_TILE] MODE 1: k-mers sorted by bin
    uint32_t *s_vm = reinterpret_cast<uint32_t *>(s_stage + (MODE == 1 ? K1S_TILE : 0));   // [K1S_WORDS]
    uint32_t *s_R = s_vm + K1S_WORDS;                                          // [K1S_RANKS] m-mer ranks, then in-block prefix minima
    uint32_t *s_S = s_R + K1S_RANKS;                                           // [K1S_RANKS] in-block suffix minima
    uint32_t *s_cnt = s_S + K1S_RANKS;                                         // [P]
    uint32_t *s_off = s_cnt + P;                                               // [P+1]      MODE 1
    uint32_t *s_cur = s_off + (MODE == 1 ? P + 1 : 0);                         // [P]        MODE 1
    uint16_t *s_b = reinterpret_cast<uint16_t *>(s_cur + (MODE == 1 ? P : 0)); // [K1S_TILE] MODE 1: bin by position
    uint16_t *s_sb = s_b + (MODE == 1 ? K1S_TILE : 0);                         // [K1S_TILE] MODE 1: bin by sorted slot
    __shared__ uint32_t s_warp[K1S_THREADS / 32];
    const int t = threadIdx.x, k = p.k, m = p.m, w = k - m + 1;

    if (MODE == 0) { for (uint32_t i = t; i < P; i += K1S_THREADS) s_cnt[i] = 0; __syncthreads(); }

    for (uint64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint64_t tile_start = tile * (uint64_t)K1S_TILE;
        if (t < K1S_WORDS) { uint64_t pk; uint32_t vm; load_pack32(p.bytes, p.n_bytes, tile_start + (uint64_t)t * 32, pk, vm); s_pk[t] = pk; s_vm[t] = vm; }
        if (MODE == 1) for (uint32_t i = t; i < P; i += K1S_THREADS) s_cnt[i] = 0;
        __syncthreads();
        if (MODE == 0) {
            // ranks of the canonical m-mers starting at every position of the tile (+ halo)
            for (int q = t; q < K1S_RANKS; q += K1S_THREADS) {
                int wi = q >> 5, o = q & 31;
                uint64_t mf = funnel_left(s_pk[wi], s_pk[wi + 1], 2 * o) >> (64 - 2 * m);
                s_R[q] = mmer_rank((uint32_t)canonical64(mf, m));
            }
            __syncthreads();
            // sliding-window minimum (van Herk / Gil-Werman): blocks of w ranks, suffix minima into s_S, prefix minima in place;
            // min over [q, q+w-1] = min(s_S[q], s_R[q+w-1])
            for (int b = t; b * w < K1S_RANKS; b += K1S_THREADS) {
                const int lo = b * w, hi = min(lo + w, K1S_RANKS);
                uint32_t run = 0xFFFFFFFFu;
                for (int q = hi - 1; q >= lo; q--) { run = min(run, s_R[q]); s_S[q] = run; }
                run = 0xFFFFFFFFu;
                for (int q = lo; q < hi; q++) { run = min(run, s_R[q]); s_R[q] = run; }
            }
            __syncthreads();
        }
#pragma unroll 2
        for (int j = 0; j < K1S_TILE / K1S_THREADS; j++) {
            const int q = t + j * K1S_THREADS;
            const int wi = q >> 5, o = q & 31;
            uint64_t c = EMPTY_KEY; uint32_t bin = 0xFFFFu;
            if (MODE == 0) {
                const uint64_t V = (((uint64_t)s_vm[wi] << 32) | s_vm[wi + 1]) << o;
                if (((~V) >> (64 - k)) == 0) {
                    uint64_t fwd = funnel_left(s_pk[wi], s_pk[wi + 1], 2 * o) >> (64 - 2 * k);
                    uint64_t rc = revcomp64(fwd, k);
                    c = fwd < rc ? fwd : rc;
                    const uint32_t mr = min(s_S[q], s_R[q + w - 1]);
                    bin = owner_of_rank(mr, p.world) * p.P_local + hash_to_bin(mix64(c), p.P_local);
                    atomicAdd(&s_cnt[bin], 1u);
                }
                p.pos_bin[tile_start + q] = (uint16_t)bin;
            } else {
                bin = p.pos_bin[tile_start + q];
                if (bin != 0xFFFFu) {
                    uint64_t fwd = funnel_left(s_pk[wi], s_pk[wi + 1], 2 * o) >> (64 - 2 * k);
                    uint64_t rc = revcomp64(fwd, k);
                    c = fwd < rc ? fwd : rc;
                    atomicAdd(&s_cnt[bin], 1u);
                }
                s_c[q] = c; s_b[q] = (uint16_t)bin;
            }
        }
        __syncthreads();
        if (MODE == 1) {
            // exclusive scan of the tile's bin counts + one global reservation per non-empty bin
            {
                const uint32_t per = (P + K1S_THREADS - 1) / K1S_THREADS;
                uint32_t b0 = t * per, sum = 0;
                for (uint32_t j = 0; j < per; j++) if (b0 + j < P) sum += s_cnt[b0 + j];
                uint32_t incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane_id() >= d) incl += n; }
                if (lane_id() == 31) s_warp[t >> 5] = incl;
                __syncthreads();
                if (t < 32) {
                    uint32_t wv = t < K1S_THREADS / 32 ? s_warp[t] : 0, wi2 = wv;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi2, d); if ((int)lane_id() >= d) wi2 += n; }
                    if (t < K1S_THREADS / 32) s_warp[t] = wi2 - wv;
                }
                __syncthreads();
                uint32_t run = s_warp[t >> 5] + incl - sum;
                for (uint32_t j = 0; j < per; j++) if (b0 + j < P) {
                    uint32_t cn = s_cnt[b0 + j];
                    s_off[b0 + j] = run; s_cur[b0 + j] = run;
                    if (cn) s_gbase[b0 + j] = p.bin_off[b0 + j] + atomicAdd(&p.bin_cursor[b0 + j], (unsigned long long)cn);
                    run += cn;
                }
                if (t == K1S_THREADS - 1) s_off[P] = run;
            }
            __syncthreads();
            for (int j = 0; j < K1S_TILE / K1S_THREADS; j++) {
                const int q = t + j * K1S_THREADS;
                uint64_t c = s_c[q];
                if (c == EMPTY_KEY) continue;
                uint32_t b = s_b[q];
                uint32_t pos = atomicAdd(&s_cur[b], 1u);
                s_stage[pos] = c; s_sb[pos] = (uint16_t)b;
            }
            __syncthreads();
            const uint32_t total = s_off[P];
            for (uint32_t j = t; j < total; j += K1S_THREADS) {
                uint32_t b = s_sb[j];
                p.out_kmers[s_gbase[b] + (j - s_off[b])] = s_stage[j];
            }
            __syncthreads();
        }
    }
    if (MODE == 0) {
        __syncthreads();
        for (uint32_t i = t; i < P; i += K1S_THREADS) { uint32_t cn = s_cnt[i]; if (cn) atomicAdd(&p.bin_count[i], (unsigned long long)cn); }
    }
}

static inline size_t k1s_smem_bytes(int mode, uint32_t P) {
    size_t s = (size_t)K1S_WORDS * 8 + (size_t)K1S_WORDS * 4 + 2 * (size_t)K1S_RANKS * 4 + (size_t)P * 4;
    if (mode == 1) s += (size_t)P * 8 + 2 * (size_t)K1S_TILE * 8 + (size_t)(P + 1) * 4 + (size_t)P * 4 + 2 * (size_t)K1S_TILE * 2;
    return s + 32;
}

// ------------------------------------------------------------------------------------------------ multi-run views

// A bin received from `world` sources is a set of runs inside the receive buffer.
constexpr int MAX_WORLD = 16;
struct Runs {
    uint64_t off[MAX_WORLD];     // start of each run (elements)
    uint64_t pre[MAX_WORLD + 1]; // exclusive prefix of run lengths
    int n;
};
__device__ __forceinline__ uint64_t runs_index(const Runs &r, uint64_t i) {
    int s = 0;
#pragma unroll
    for (int j = 1; j < MAX_WORLD; j++) if (j < r.n && i >= r.pre[j]) s = j;
    return r.off[s] + (i - r.pre[s]);
}

__global__ void __launch_bounds__(256) k2s_insert(const uint64_t *__restrict__ kmers, Runs runs, uint64_t *keys, uint32_t *counts, uint32_t cap,
                                                  int *overflow) {
    const uint64_t n = runs.pre[runs.n];
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = kmers[runs_index(runs, i)];
        uint32_t slot = table_slot(key, cap), probes = 0;
        for (;;) {
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

// ------------------------------------------------------------------------------------------------ K3 (sharded)

// record value: bits 39..0 = gid of the k-mer end, bit 62 = side OUT, bit 63 = palindrome
#define SREC_SIDE_OUT (1ULL << 62)
#define SREC_PALIN (1ULL << 63)
#define SREC_BOTH (1ULL << 61)   // see REC_BOTH / REC_SKIP in kernels.cuh
#define SREC_SKIP (1ULL << 60)

struct K3SParams {
    const uint64_t *solid_kmers; uint64_t n_solid; int k, m; uint32_t rank, world, P_local;
    unsigned long long *bin_count; const uint64_t *bin_off; unsigned long long *bin_cursor;
    uint64_t *rec_key, *rec_val;
    uint16_t *bins;                  // [2S] bin of every end record: written by the histogram pass, read by the scatter pass
};

template <int MODE>
__global__ void __launch_bounds__(K3_THREADS) k3s_ends(K3SParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const uint32_t P = p.world * p.P_local;
    uint64_t *s_base = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_base + (MODE == 1 ? P : 0));
    const int t = threadIdx.x;
    const uint64_t n_tiles = (p.n_solid + K3_TILE - 1) / K3_TILE;
    for (uint32_t i = t; i < P; i += K3_THREADS) s_cnt[i] = 0;
    __syncthreads();
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        uint64_t key[2 * K3_PER_THREAD], val[2 * K3_PER_THREAD];
        uint32_t bin[2 * K3_PER_THREAD], rnk[2 * K3_PER_THREAD];
#pragma unroll
        for (int j = 0; j < K3_PER_THREAD; j++) {
            uint64_t i = tile * K3_TILE + (uint64_t)j * K3_THREADS + t;
            bool ok = i < p.n_solid;
            uint64_t x = ok ? p.solid_kmers[i] : 0;
#pragma unroll
            for (int end = 0; end < 2; end++) {
                const int r = 2 * j + end;
                uint64_t v;
                kmer_end_record(x, p.k, end, (uint32_t)i, key[r], v);
                val[r] = ((uint64_t)p.rank << 32) | (uint32_t)v | ((v & REC_SIDE_OUT) ? SREC_SIDE_OUT : 0) | ((v & REC_PALIN) ? SREC_PALIN : 0) |
                         ((v & REC_BOTH) ? SREC_BOTH : 0) | ((v & REC_SKIP) ? SREC_SKIP : 0);
                bin[r] = 0xFFFFFFFFu; rnk[r] = 0;
                if (ok) {
                    if (MODE == 0) {
                        int mm = p.m < p.k - 1 ? p.m : p.k - 1;
                        bin[r] = owner_of_rank(min_rank_of(key[r], p.k - 1, mm), p.world) * p.P_local + hash_to_bin(mix64(key[r]), p.P_local);
                        p.bins[2 * i + end] = (uint16_t)bin[r];
                    } else bin[r] = p.bins[2 * i + end];
                    rnk[r] = atomicAdd(&s_cnt[bin[r]], 1u);
                }
            }
        }
        if (MODE == 1) {
            __syncthreads();
            for (uint32_t b = t; b < P; b += K3_THREADS) {
                uint32_t c = s_cnt[b];
                if (c) s_base[b] = p.bin_off[b] + atomicAdd(&p.bin_cursor[b], (unsigned long long)c);
                s_cnt[b] = 0;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 2 * K3_PER_THREAD; r++) {
                if (bin[r] == 0xFFFFFFFFu) continue;
                uint64_t pos = s_base[bin[r]] + rnk[r];
                p.rec_key[pos] = key[r]; p.rec_val[pos] = val[r];
            }
            __syncthreads();
        }
    }
    if (MODE == 0) {
        __syncthreads();
        for (uint32_t b = t; b < P; b += K3_THREADS) { uint32_t c = s_cnt[b]; if (c) atomicAdd(&p.bin_count[b], (unsigned long long)c); }
    }
}

// ------------------------------------------------------------------------------------------------ K5 (sharded)

struct JoinTable64 { uint64_t *keys; uint32_t *cnt; uint64_t *ids; uint32_t cap; };   // ids[cap*8]: gids, 0..3 IN, 4..7 OUT

__global__ void __launch_bounds__(256) k5s_insert(const uint64_t *__restrict__ rec_key, const uint64_t *__restrict__ rec_val, Runs runs,
                                                  JoinTable64 T, int *overflow) {
    const uint64_t n = runs.pre[runs.n];
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t src = runs_index(runs, i);
        const uint64_t key = rec_key[src], val = rec_val[src];
        if (val & SREC_SKIP) continue;
        uint32_t slot = table_slot(key, T.cap), probes = 0;
        for (;;) {
            uint64_t cur = ld_volatile_u64(&T.keys[slot]);
            if (cur == EMPTY_KEY) {
                cur = atomicCAS((unsigned long long *)&T.keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key);
                if (cur == EMPTY_KEY) cur = key;
            }
            if (cur == key) {
                const bool pal = (val & SREC_PALIN) != 0, use_out = (val & SREC_SIDE_OUT) != 0 && !pal;
                uint32_t old = atomicAdd(&T.cnt[slot], use_out ? 0x100u : 0x1u);
                if (pal) atomicOr(&T.cnt[slot], 0x80000000u);
                uint32_t pos = use_out ? ((old >> 8) & 0xFF) : (old & 0xFF);
                if (pos < 4) {
                    T.ids[(size_t)slot * 8 + (use_out ? 4 : 0) + pos] = val & GID_MASK;
                    if (val & SREC_BOTH) atomicOr(&T.cnt[slot], 1u << ((use_out ? JOIN_BOTH_SHIFT_OUT : JOIN_BOTH_SHIFT_IN) + pos));
                }
                break;
            }
            slot = slot + 1 == T.cap ? 0 : slot + 1;
            if (++probes > min(T.cap, 8192u)) { *overflow = 1; break; }
        }
    }
}

// Same decision rule as k5_resolve, but the results are MESSAGES because the ends may live on other GPUs:
//   link message (a, b): "end a is compacted with end b" (emitted twice, once per direction, routed to rank(a));
//   edge (e, f): unitig-level edge between extremities.  Two-pass warp-chunk scan like k2_emit: count, reserve once per block, write.
__global__ void __launch_bounds__(256) k5s_resolve(JoinTable64 T, uint32_t cap, uint32_t rank, uint64_t *link64, ulonglong2 *link_msgs, uint64_t link_cap,
                                                   unsigned long long *n_link_msgs, ulonglong2 *edges, uint64_t edge_cap, unsigned long long *n_edges,
                                                   unsigned long long *n_compact) {
    __shared__ uint32_t s_wl[8], s_we[8];
    __shared__ unsigned long long s_bl, s_be;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t n_warps = gridDim.x * 8;
    const uint32_t per_warp = ((cap + n_warps - 1) / n_warps + 31u) & ~31u;
    const uint32_t gw = blockIdx.x * 8 + warp;
    const uint32_t wb = (uint32_t)min((uint64_t)cap, (uint64_t)gw * per_warp), we = (uint32_t)min((uint64_t)cap, (uint64_t)wb + per_warp);
    // counts with the (2v, 2v+1) pairs of self-complementary k-mers written out (REC_BOTH); bin/bout say which stored ids are pairs
    auto classify = [&](uint32_t i, uint32_t &nin, uint32_t &nout, bool &pal, bool &compact, uint32_t &sin, uint32_t &sout, uint32_t &bin_, uint32_t &bout) {
        uint32_t c = T.cnt[i];
        sin = min(c & 0xFFu, 4u); sout = min((c >> 8) & 0xFFu, 4u); pal = (c >> 31) != 0;
        bin_ = (c >> JOIN_BOTH_SHIFT_IN) & ((1u << sin) - 1u); bout = (c >> JOIN_BOTH_SHIFT_OUT) & ((1u << sout) - 1u);
        nin = sin + __popc(bin_); nout = sout + __popc(bout);
        compact = !pal && nin == 1 && nout == 1 && (T.ids[(size_t)i * 8] >> 1) != (T.ids[(size_t)i * 8 + 4] >> 1);
    };
    uint32_t my_l = 0, my_e = 0; unsigned long long my_c = 0;
    for (uint32_t i = wb + lane; i < we; i += 32) {
        if (T.keys[i] == EMPTY_KEY) continue;
        uint32_t nin, nout, sin, sout, bin_, bout; bool pal, compact;
        classify(i, nin, nout, pal, compact, sin, sout, bin_, bout);
        if (compact) {   // only the directions whose destination end lives on another GPU become messages
            my_l += (GID_RANK(T.ids[(size_t)i * 8]) != rank) + (GID_RANK(T.ids[(size_t)i * 8 + 4]) != rank);
            my_c++;
        } else my_e += pal ? nin * nin : 2 * nin * nout;
    }
    uint32_t wl = my_l, wed = my_e;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { wl += __shfl_xor_sync(0xFFFFFFFFu, wl, d); wed += __shfl_xor_sync(0xFFFFFFFFu, wed, d); }
    if (lane == 0) { s_wl[warp] = wl; s_we[warp] = wed; }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tl = 0, te = 0;
        for (int x = 0; x < 8; x++) { uint32_t a = s_wl[x]; s_wl[x] = tl; tl += a; uint32_t b = s_we[x]; s_we[x] = te; te += b; }
        s_bl = tl ? atomicAdd(n_link_msgs, (unsigned long long)tl) : 0ULL;
        s_be = te ? atomicAdd(n_edges, (unsigned long long)te) : 0ULL;
    }
    __syncthreads();
    // lane-exclusive offsets inside the warp
    uint32_t xl = my_l, xe = my_e;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t a = __shfl_up_sync(0xFFFFFFFFu, xl, d), b = __shfl_up_sync(0xFFFFFFFFu, xe, d); if ((int)lane >= d) { xl += a; xe += b; } }
    uint64_t ol = s_bl + s_wl[warp] + xl - my_l, oe = s_be + s_we[warp] + xe - my_e;
    for (uint32_t i = wb + lane; i < we; i += 32) {
        if (T.keys[i] == EMPTY_KEY) continue;
        uint32_t nin, nout, sin, sout, bin_, bout; bool pal, compact;
        classify(i, nin, nout, pal, compact, sin, sout, bin_, bout);
        const uint64_t *ids = &T.ids[(size_t)i * 8];
        uint64_t ein[5], eout[5];
        join_expand_side(ids, sin, bin_, ein);
        join_expand_side(ids + 4, sout, bout, eout);
        if (compact) {
            uint64_t a = ids[0], b = ids[4];
            if (GID_RANK(a) == rank) link64[GID_LOCAL(a)] = b; else { if (ol < link_cap) link_msgs[ol] = make_ulonglong2(a, b); ol++; }
            if (GID_RANK(b) == rank) link64[GID_LOCAL(b)] = a; else { if (ol < link_cap) link_msgs[ol] = make_ulonglong2(b, a); ol++; }
        } else if (pal) {
            for (uint32_t a = 0; a < nin; a++) for (uint32_t b = 0; b < nin; b++) { if (oe < edge_cap) edges[oe] = make_ulonglong2(ein[a], ein[b]); oe++; }
        } else {
            for (uint32_t a = 0; a < nin; a++) for (uint32_t b = 0; b < nout; b++) {
                if (oe + 2 <= edge_cap) { edges[oe] = make_ulonglong2(ein[a], eout[b]); edges[oe + 1] = make_ulonglong2(eout[b], ein[a]); }
                oe += 2;
            }
        }
        T.keys[i] = EMPTY_KEY; T.cnt[i] = 0;
    }
    block_add(n_compact, my_c);
}

// ---- the sharded join with the tables in SHARED memory (the sharded counterpart of k5c_join, k5_smem.cuh): the received end records
// are split into fine bins (k5p_hist / k5p_scatter with `group`), one CTA joins a fine bin.  A table slot remembers the POSITION of
// up to 4 IN and 4 OUT records inside the bin (32 bits each, as in k5c_join) -- the 40-bit global end ids are read back from the
// record array when the junction is decided -- and the decisions leave as in k5s_resolve: links whose end is local are written,
// the others become messages, unitig-level edges are (gid, gid) pairs.  Reservations are warp-aggregated.
struct K5CSParams {
    const uint64_t *key, *val;          // records ordered by fine bin
    const uint64_t *fine_off;           // [n_fine + 1]
    uint32_t n_fine, rank;
    uint64_t *link64;                   // [2S] partner gid of every local k-mer end (pre-set to NOGID)
    ulonglong2 *link_msgs; uint64_t link_cap; unsigned long long *n_link_msgs;
    ulonglong2 *edges; uint64_t edge_cap; unsigned long long *n_edges;
    unsigned long long *n_compact, *n_splits;
    int *failed;
};

constexpr int K5CS_THREADS = 256;
constexpr int K5CS_MAX_LEVEL = 10;
template <int SLOTS> static inline size_t k5cs_smem_bytes() { return (size_t)SLOTS * (8 + 4 + 32) + 128; }

template <int SLOTS>
__global__ void __launch_bounds__(K5CS_THREADS) k5cs_join(K5CSParams p) {
    constexpr int T = K5CS_THREADS, NW = T / 32, PER_WARP = SLOTS / NW;
    constexpr uint32_t MASK = SLOTS - 1, FILL_LIMIT = SLOTS / 4 * 3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *s_keys = reinterpret_cast<uint64_t *>(smem_raw);                   // [SLOTS]
    uint32_t *s_ids = reinterpret_cast<uint32_t *>(s_keys + SLOTS);              // [SLOTS * 8]  record positions: 0..3 IN, 4..7 OUT
    uint32_t *s_cnt = s_ids + (size_t)SLOTS * 8;                                 // [SLOTS] IN count | OUT count << 8 | pair flags | palindrome << 31
    __shared__ int s_over;
    __shared__ uint32_t s_fill;
    const uint32_t t = threadIdx.x, lane = lane_id(), warp = t >> 5;
    const uint32_t wb = warp * PER_WARP;
    unsigned long long compact = 0, my_splits = 0;

    for (uint32_t bin = blockIdx.x; bin < p.n_fine; bin += gridDim.x) {
        const uint64_t beg = p.fine_off[bin], end = p.fine_off[bin + 1];
        if (beg == end) continue;
        uint32_t stk[K5CS_MAX_LEVEL + 2];
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
                    if (key[u] == EMPTY_KEY || (val[u] & SREC_SKIP)) continue;
                    const uint64_t h = mix64(key[u]);
                    if ((((uint32_t)h >> 16) & sel_mask) != part) continue;
                    const uint32_t pos_in_bin = (uint32_t)(base + lane + (uint64_t)u * T - beg);
                    uint32_t slot = (uint32_t)h & MASK;
                    for (uint32_t probes = 0;;) {
                        uint64_t cur = ld_volatile_u64(&s_keys[slot]);
                        if (cur == EMPTY_KEY) {
                            cur = atomicCAS((unsigned long long *)&s_keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)key[u]);
                            if (cur == EMPTY_KEY) { cur = key[u]; fresh++; }
                        }
                        if (cur == key[u]) {
                            const bool pal = (val[u] & SREC_PALIN) != 0, use_out = (val[u] & SREC_SIDE_OUT) != 0 && !pal;
                            const uint32_t old = atomicAdd(&s_cnt[slot], use_out ? 0x100u : 0x1u);
                            if (pal) atomicOr(&s_cnt[slot], 0x80000000u);
                            const uint32_t pos = use_out ? ((old >> 8) & 0xFF) : (old & 0xFF);
                            if (pos < 4) {
                                s_ids[(size_t)slot * 8 + (use_out ? 4 : 0) + pos] = pos_in_bin;
                                if (val[u] & SREC_BOTH) atomicOr(&s_cnt[slot], 1u << ((use_out ? JOIN_BOTH_SHIFT_OUT : JOIN_BOTH_SHIFT_IN) + pos));
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
                if (level >= (uint32_t)K5CS_MAX_LEVEL) { if (t == 0) *p.failed = 1; break; }
                stk[sp++] = ((level + 1) << 16) | part;
                stk[sp++] = ((level + 1) << 16) | part | (1u << level);
                continue;
            }

            // ---- resolve: every warp walks its own slots, 32 at a time (uniform trip count: PER_WARP is a multiple of 32)
            for (uint32_t i = wb + lane; i < wb + PER_WARP; i += 32) {
                uint32_t nin = 0, nout = 0, n_msg = 0, n_edge = 0;
                bool pal = false, comp = false, occ = s_keys[i] != EMPTY_KEY;
                uint64_t ein[5], eout[5];
                if (occ) {
                    const uint32_t c = s_cnt[i];
                    const uint32_t sin = min(c & 0xFFu, 4u), sout = min((c >> 8) & 0xFFu, 4u);
                    pal = (c >> 31) != 0;
                    uint64_t gin[4], gout[4];
                    for (uint32_t a = 0; a < sin; a++) gin[a] = p.val[beg + s_ids[(size_t)i * 8 + a]] & GID_MASK;
                    for (uint32_t a = 0; a < sout; a++) gout[a] = p.val[beg + s_ids[(size_t)i * 8 + 4 + a]] & GID_MASK;
                    nin = join_expand_side(gin, sin, (c >> JOIN_BOTH_SHIFT_IN) & ((1u << sin) - 1u), ein);
                    nout = join_expand_side(gout, sout, (c >> JOIN_BOTH_SHIFT_OUT) & ((1u << sout) - 1u), eout);
                    comp = !pal && nin == 1 && nout == 1 && (ein[0] >> 1) != (eout[0] >> 1);
                    if (comp) n_msg = (GID_RANK(ein[0]) != p.rank) + (GID_RANK(eout[0]) != p.rank);
                    else n_edge = pal ? nin * nin : 2 * nin * nout;
                }
                // warp-aggregated reservations
                uint32_t xm = n_msg, xe = n_edge;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t a = __shfl_up_sync(0xFFFFFFFFu, xm, d), b = __shfl_up_sync(0xFFFFFFFFu, xe, d); if ((int)lane >= d) { xm += a; xe += b; } }
                const uint32_t tm = __shfl_sync(0xFFFFFFFFu, xm, 31), te = __shfl_sync(0xFFFFFFFFu, xe, 31);
                unsigned long long bm = 0, be = 0;
                if (lane == 0) { if (tm) bm = atomicAdd(p.n_link_msgs, (unsigned long long)tm); if (te) be = atomicAdd(p.n_edges, (unsigned long long)te); }
                bm = __shfl_sync(0xFFFFFFFFu, bm, 0); be = __shfl_sync(0xFFFFFFFFu, be, 0);
                uint64_t ol = bm + xm - n_msg, oe = be + xe - n_edge;
                if (!occ) continue;
                if (comp) {
                    const uint64_t a = ein[0], b = eout[0];
                    if (GID_RANK(a) == p.rank) p.link64[GID_LOCAL(a)] = b; else { if (ol < p.link_cap) p.link_msgs[ol] = make_ulonglong2(a, b); ol++; }
                    if (GID_RANK(b) == p.rank) p.link64[GID_LOCAL(b)] = a; else { if (ol < p.link_cap) p.link_msgs[ol] = make_ulonglong2(b, a); ol++; }
                    compact++;
                } else if (pal) {
                    for (uint32_t a = 0; a < nin; a++) for (uint32_t b = 0; b < nin; b++) { if (oe < p.edge_cap) p.edges[oe] = make_ulonglong2(ein[a], ein[b]); oe++; }
                } else {
                    for (uint32_t a = 0; a < nin; a++) for (uint32_t b = 0; b < nout; b++) {
                        if (oe + 2 <= p.edge_cap) { p.edges[oe] = make_ulonglong2(ein[a], eout[b]); p.edges[oe + 1] = make_ulonglong2(eout[b], ein[a]); }
                        oe += 2;
                    }
                }
            }
            // (no barrier: the next pass clears only this warp's own slots; inserts into them start after the next barrier)
        }
    }
    block_add(p.n_compact, compact);
    block_add(p.n_splits, my_splits);
}

// ------------------------------------------------------------------------------------------------ generic routing of 16-byte records

// destination = rank field of rec.x (bits 39..32)
// (with 2..16 destinations a shared-memory atomic per record would serialise whole warps on one or two counters -- 24 ms for the 10 M
//  link records of C3 on 2 GPUs -- so ranks and counts are taken per warp from one ballot per destination)
__global__ void __launch_bounds__(256) route_hist(const ulonglong2 *__restrict__ recs, uint64_t n, uint32_t world, unsigned long long *count) {
    __shared__ uint32_t s_c[MAX_WORLD];
    if (threadIdx.x < MAX_WORLD) s_c[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t lane = lane_id();
    uint32_t acc = 0;                                            // lane w counts destination w for its warp
    const uint64_t per_iter = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t iters = (n + per_iter - 1) / per_iter;
    for (uint64_t it = 0; it < iters; it++) {                    // warp-uniform trip count
        const uint64_t i = it * per_iter + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
        const bool ok = i < n;
        const uint32_t d = ok ? GID_RANK(recs[i].x) : 0xFFFFFFFFu;
        for (uint32_t w = 0; w < world; w++) {
            const unsigned m = __ballot_sync(0xFFFFFFFFu, d == w);
            if (lane == w) acc += __popc(m);
        }
    }
    if (lane < world && acc) atomicAdd(&s_c[lane], acc);
    __syncthreads();
    if (threadIdx.x < world && s_c[threadIdx.x]) atomicAdd(&count[threadIdx.x], (unsigned long long)s_c[threadIdx.x]);
}

__global__ void __launch_bounds__(256) route_scatter(const ulonglong2 *__restrict__ recs, uint64_t n, uint32_t world, const uint64_t *__restrict__ off,
                                                     unsigned long long *cursor, ulonglong2 *out) {
    __shared__ uint32_t s_c[MAX_WORLD];
    __shared__ uint64_t s_base[MAX_WORLD];
    const uint32_t lane = lane_id();
    const uint64_t per_iter = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t iters = (n + per_iter - 1) / per_iter;
    for (uint64_t it = 0; it < iters; it++) {
        uint64_t i = it * per_iter + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
        if (threadIdx.x < MAX_WORLD) s_c[threadIdx.x] = 0;
        __syncthreads();
        ulonglong2 r = make_ulonglong2(0, 0); uint32_t d = 0xFFFFFFFFu, rk = 0; const bool ok = i < n;
        if (ok) { r = recs[i]; d = GID_RANK(r.x); }
        for (uint32_t w = 0; w < world; w++) {                   // rank inside the block's share of destination w: one atomic per warp
            const unsigned m = __ballot_sync(0xFFFFFFFFu, d == w);
            uint32_t b = 0;
            if (lane == 0 && m) b = atomicAdd(&s_c[w], (uint32_t)__popc(m));
            b = __shfl_sync(0xFFFFFFFFu, b, 0);
            if (d == w) rk = b + __popc(m & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (threadIdx.x < world && s_c[threadIdx.x]) s_base[threadIdx.x] = off[threadIdx.x] + atomicAdd(&cursor[threadIdx.x], (unsigned long long)s_c[threadIdx.x]);
        __syncthreads();
        if (ok) out[s_base[d] + rk] = r;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ stage 2 on a shard

// apply link messages (a, b): link64[a.local] = b
__global__ void __launch_bounds__(256) shard_apply_links(const ulonglong2 *__restrict__ msgs, uint64_t n, uint64_t *link64) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        ulonglong2 m = msgs[i];
        link64[GID_LOCAL(m.x)] = m.y;
    }
}

// local view of the links: partners on other GPUs are chain ends for the local ranking
__global__ void __launch_bounds__(256) shard_local_links(const uint64_t *__restrict__ link64, uint64_t n2, uint32_t rank, uint32_t *link32) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t g = link64[i];
        link32[i] = (g != NOGID && GID_RANK(g) == rank) ? GID_LOCAL(g) : NOLINK;
    }
}

// after local cycle cutting link32 lost some links: mirror that in link64 (a locally cut link must not become a glue link)
__global__ void __launch_bounds__(256) shard_sync_cut(const uint32_t *__restrict__ link32, uint64_t n2, uint32_t rank, uint64_t *link64) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n2; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t g = link64[i];
        if (g != NOGID && GID_RANK(g) == rank && link32[i] == NOLINK) link64[i] = NOGID;
    }
}

// ------------------------------------------------------------------------------------------------ stage 3: glue of pieces

// piece descriptor exchanged between GPUs (7 x u64)
struct PieceDesc {
    uint64_t nbr[2];     // neighbour piece terminus: (global piece id << 1 | terminus), NOGID if none   [0] = at BEGIN, [1] = at END
    uint64_t term[2];    // gid of the k-mer end at the piece BEGIN / END
    uint64_t nk;         // k-mers in the piece
    uint64_t kc;         // abundance sum
    uint64_t seq_off;    // offset of the piece string in its GPU's piece buffer
};

// per local piece: termini gids, partner gids (the k-mer ends on other GPUs its termini are glued to), sizes
__global__ void __launch_bounds__(256) glue_describe(uint64_t n_pieces, uint32_t rank, const uint32_t *__restrict__ term_begin,
                                                     const uint32_t *__restrict__ term_end, const uint64_t *__restrict__ link64,
                                                     const uint64_t *__restrict__ seq_off, const unsigned long long *__restrict__ kc, int k,
                                                     PieceDesc *desc) {
    for (uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; q < n_pieces; q += (uint64_t)gridDim.x * blockDim.x) {
        PieceDesc d;
        uint32_t e0 = term_begin[q], e1 = term_end[q];
        d.term[0] = ((uint64_t)rank << 32) | e0; d.term[1] = ((uint64_t)rank << 32) | e1;
        d.nbr[0] = link64[e0]; d.nbr[1] = link64[e1];          // NOGID when the terminus is a real chain end
        d.nk = seq_off[q + 1] - seq_off[q] - (uint64_t)(k - 1);
        d.kc = kc[q]; d.seq_off = seq_off[q];
        desc[q] = d;
    }
}

// Replicated lookup table  k-mer-end gid -> global piece terminus id (2 * piece + {0 begin, 1 end}), built on every GPU from the
// all-gathered descriptors.  It replaces two exchanges: a piece finds its neighbour by looking up its partner gid, and the link
// stage resolves both ends of an edge locally.
struct GidTable { uint64_t *keys; uint32_t *vals; uint32_t cap; };
__device__ __forceinline__ uint32_t gid_lookup(const GidTable &T, uint64_t gid) {
    uint32_t slot = table_slot(gid, T.cap);
    for (uint32_t probes = 0; probes <= T.cap; probes++) {
        uint64_t cur = T.keys[slot];
        if (cur == gid) return T.vals[slot];
        if (cur == EMPTY_KEY) return NOLINK;
        slot = slot + 1 == T.cap ? 0 : slot + 1;
    }
    return NOLINK;
}
__global__ void __launch_bounds__(256) glue_table_build(const PieceDesc *__restrict__ gathered, uint64_t max_p, uint32_t world,
                                                        const uint64_t *__restrict__ piece_base, uint64_t G, GidTable T) {
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < 2 * G; x += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = x >> 1; const int t = (int)(x & 1);
        uint32_t r = 0;
        for (uint32_t j = 1; j < world; j++) if (g >= piece_base[j]) r = j;
        const uint64_t gid = gathered[(uint64_t)r * max_p + (g - piece_base[r])].term[t];
        uint32_t slot = table_slot(gid, T.cap);
        for (;;) {
            uint64_t cur = atomicCAS((unsigned long long *)&T.keys[slot], (unsigned long long)EMPTY_KEY, (unsigned long long)gid);
            if (cur == EMPTY_KEY) { T.vals[slot] = (uint32_t)x; break; }
            if (cur == gid) break;                                    // every terminus gid is unique: cannot happen
            slot = slot + 1 == T.cap ? 0 : slot + 1;
        }
    }
}

// dense global piece tables from the padded all-gather; plink = piece-terminus links for the replicated ranking
__global__ void __launch_bounds__(256) glue_build(const PieceDesc *__restrict__ gathered, uint64_t max_p, uint32_t world, const uint64_t *__restrict__ piece_base,
                                                  uint64_t G, GidTable T, uint32_t *plink, uint32_t *weight) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < G; g += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t r = 0;
        for (uint32_t j = 1; j < world; j++) if (g >= piece_base[j]) r = j;
        const PieceDesc &d = gathered[(uint64_t)r * max_p + (g - piece_base[r])];
        plink[2 * g] = d.nbr[0] == NOGID ? NOLINK : gid_lookup(T, d.nbr[0]);
        plink[2 * g + 1] = d.nbr[1] == NOGID ? NOLINK : gid_lookup(T, d.nbr[1]);
        weight[g] = (uint32_t)d.nk;
    }
}

__device__ __forceinline__ const PieceDesc &piece_desc(const PieceDesc *gathered, uint64_t max_p, uint32_t world, const uint64_t *piece_base, uint64_t g,
                                                       uint32_t &r) {
    r = 0;
    for (uint32_t j = 1; j < world; j++) if (g >= piece_base[j]) r = j;
    return gathered[(uint64_t)r * max_p + (g - piece_base[r])];
}

// weighted variants of the ranking kernels: an element's hop count is the number of k-mers of the pieces AFTER it
__global__ void __launch_bounds__(256) k6w_init(const uint32_t *__restrict__ link, const uint32_t *__restrict__ weight, uint64_t *pair, uint64_t n2) {
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t s = link[x ^ 1];
        pair[x] = s == NOLINK ? K6_TAIL : (((uint64_t)s << 32) | weight[s >> 1]);
    }
}
__global__ void __launch_bounds__(256) k6w_reinit_flagged(const uint8_t *flag, const uint32_t *link, const uint32_t *weight, uint64_t *pair, uint64_t n2) {
    for (uint64_t x = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; x < n2; x += (uint64_t)gridDim.x * blockDim.x) {
        if (!flag[x]) continue;
        uint32_t s = link[x ^ 1];
        pair[x] = s == NOLINK ? K6_TAIL : (((uint64_t)s << 32) | weight[s >> 1]);
    }
}
// info[g] = (start terminus << 32) | k-mer offset << 1 | reversed ; heads publish the unitig length in nt
__global__ void __launch_bounds__(256) k6w_finalize(const uint64_t *__restrict__ pair, const uint32_t *__restrict__ weight, uint64_t G, int k, uint64_t *info,
                                                    uint32_t *head_flag, uint32_t *head_nt) {
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < G; v += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t p0 = pair[2 * v], p1 = pair[2 * v + 1];
        uint32_t n0 = (uint32_t)(p0 >> 32), n1 = (uint32_t)(p1 >> 32);
        uint32_t T0 = (n0 == K6_NIL ? (uint32_t)(2 * v) : n0) ^ 1u, T1 = (n1 == K6_NIL ? (uint32_t)(2 * v + 1) : n1) ^ 1u;
        uint32_t d0 = (uint32_t)p0 & ~K6_DONE, d1 = (uint32_t)p1 & ~K6_DONE;
        uint32_t B, pos, rev;
        if (T1 < T0) { B = T1; pos = d1; rev = 0; } else { B = T0; pos = d0; rev = 1; }
        info[v] = ((uint64_t)B << 32) | ((uint64_t)pos << 1) | rev;
        bool head = pos == 0;
        head_flag[v] = head;
        head_nt[v] = head ? (d0 + d1 + weight[v] + (uint32_t)(k - 1)) : 0;
    }
}

// per piece: unitig id; accumulate unitig KC
__global__ void __launch_bounds__(256) glue_unitig_of_piece(const uint64_t *__restrict__ pinfo, const uint64_t *__restrict__ uid_scan, uint64_t G,
                                                            const PieceDesc *__restrict__ gathered, uint64_t max_p, uint32_t world,
                                                            const uint64_t *__restrict__ piece_base, uint32_t *unitig_of_piece, unsigned long long *kc) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < G; g += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t B = (uint32_t)(pinfo[g] >> 32);
        uint32_t u = (uint32_t)uid_scan[B >> 1];
        unitig_of_piece[g] = u;
        uint32_t r;
        const PieceDesc &d = piece_desc(gathered, max_p, world, piece_base, g, r);
        atomicAdd(&kc[u], (unsigned long long)d.kc);
    }
}

// assemble the unitigs this GPU owns (global ids [u_first, u_last)) from the all-gathered piece strings: one warp per piece
__global__ void __launch_bounds__(256) glue_assemble(uint64_t G, const uint64_t *__restrict__ pinfo, const uint32_t *__restrict__ unitig_of_piece,
                                                     const uint64_t *__restrict__ nt_scan /*by head piece*/, const PieceDesc *__restrict__ gathered,
                                                     uint64_t max_p, uint32_t world, const uint64_t *__restrict__ piece_base,
                                                     const char *__restrict__ all_seq, uint64_t max_nt, uint32_t u_first, uint32_t u_last,
                                                     uint64_t nt_first, int k, char *seq, uint64_t *seq_off) {
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = lane_id();
    for (uint64_t g = warp; g < G; g += n_warps) {
        uint32_t u = unitig_of_piece[g];
        if (u < u_first || u >= u_last) continue;
        uint64_t inf = pinfo[g];
        uint32_t B = (uint32_t)(inf >> 32), pos = (uint32_t)inf >> 1, rev = (uint32_t)inf & 1u;
        uint32_t r;
        const PieceDesc &d = piece_desc(gathered, max_p, world, piece_base, g, r);
        const char *src = all_seq + (uint64_t)r * max_nt + d.seq_off;
        const uint64_t len = d.nk + (uint64_t)(k - 1);
        const uint64_t uoff = nt_scan[B >> 1] - nt_first;
        char *dst = seq + uoff + pos;
        if (pos == 0 && lane == 0) seq_off[u - u_first] = uoff;
        for (uint64_t j = lane; j < len; j += 32) {
            char c = rev ? src[len - 1 - j] : src[j];
            if (rev) c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A';
            dst[j] = c;
        }
    }
}

// -all-abundance-counts on shards: the per-k-mer abundances of the unitigs this GPU owns, from the all-gathered per-piece lists
// (piece q of rank r starts at k-mer  seq_off - q * (k-1)  of that rank's list; a reversed piece lists its k-mers backwards).
// One warp per piece; unitig u (local) starts at k-mer  seq_off[u] - u * (k-1)  of `abund`, as in the single-GPU view.
__global__ void __launch_bounds__(256) glue_assemble_abund(uint64_t G, const uint64_t *__restrict__ pinfo, const uint32_t *__restrict__ unitig_of_piece,
                                                           const uint64_t *__restrict__ nt_scan /*by head piece*/, const PieceDesc *__restrict__ gathered,
                                                           uint64_t max_p, uint32_t world, const uint64_t *__restrict__ piece_base,
                                                           const uint32_t *__restrict__ all_abund, uint64_t max_s, uint32_t u_first, uint32_t u_last,
                                                           uint64_t nt_first, int k, uint32_t *abund) {
    const uint64_t warp = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = lane_id();
    for (uint64_t g = warp; g < G; g += n_warps) {
        const uint32_t u = unitig_of_piece[g];
        if (u < u_first || u >= u_last) continue;
        const uint64_t inf = pinfo[g];
        const uint32_t B = (uint32_t)(inf >> 32), pos = (uint32_t)inf >> 1, rev = (uint32_t)inf & 1u;
        uint32_t r;
        const PieceDesc &d = piece_desc(gathered, max_p, world, piece_base, g, r);
        const uint32_t *src = all_abund + (uint64_t)r * max_s + (d.seq_off - (g - piece_base[r]) * (uint64_t)(k - 1));
        uint32_t *dst = abund + (nt_scan[B >> 1] - nt_first) - (uint64_t)(u - u_first) * (uint64_t)(k - 1) + pos;
        for (uint64_t j = lane; j < d.nk; j += 32) dst[j] = rev ? src[d.nk - 1 - j] : src[j];
    }
}

// ------------------------------------------------------------------------------------------------ stage 4 on shards

// An edge (e, f) between k-mer ends that are unitig extremities: both ends are piece termini, found in the replicated table; the
// piece-level ranking (replicated too) says which unitig each terminus belongs to and whether it is the unitig's BEGIN.
__device__ __forceinline__ bool terminus_to_unitig(uint32_t term, const uint64_t *pinfo, const uint32_t *unitig_of_piece, uint32_t &u, bool &is_begin) {
    if (term == NOLINK) return false;
    const uint32_t g = term >> 1;
    const uint32_t rev = (uint32_t)pinfo[g] & 1u;
    u = unitig_of_piece[g];
    is_begin = ((term & 1u) == 0) != (rev != 0);               // a reversed piece swaps its termini
    return true;
}

// edge (e gid, f gid) -> (owner(ue) << 32 | ue,  uf | from_sign << 32 | to_sign << 40), ready to be routed to the owner of unitig ue
__global__ void __launch_bounds__(256) links_resolve(ulonglong2 *edges, uint64_t n, GidTable T, const uint64_t *__restrict__ pinfo,
                                                     const uint32_t *__restrict__ unitig_of_piece, uint64_t U_total, uint32_t world, int *bad) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        ulonglong2 ed = edges[i];
        uint32_t ue = 0, uf = 0; bool eb = false, fb = false;
        bool ok = terminus_to_unitig(gid_lookup(T, ed.x), pinfo, unitig_of_piece, ue, eb) &&
                  terminus_to_unitig(gid_lookup(T, ed.y), pinfo, unitig_of_piece, uf, fb);
        if (!ok) { *bad = 1; ue = 0; uf = 0; }
        uint64_t owner = ((uint64_t)ue * world) / U_total;
        uint64_t fs = eb ? '-' : '+', ts = fb ? '+' : '-';
        edges[i] = make_ulonglong2((owner << 32) | ue, (uint64_t)uf | (fs << 32) | (ts << 40));
    }
}

__global__ void __launch_bounds__(256) links_finalize(const ulonglong2 *__restrict__ recs, uint64_t n, uint32_t u_first, LinkRec *out, uint32_t *per_unitig) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        ulonglong2 r = recs[i];
        LinkRec l;
        l.from = (uint32_t)r.x; l.to = (uint32_t)r.y; l.from_sign = (uint8_t)(r.y >> 32); l.to_sign = (uint8_t)(r.y >> 40); l.pad0 = l.pad1 = 0;
        out[i] = l;
        atomicAdd(&per_unitig[l.from - u_first], 1u);
    }
}

__global__ void __launch_bounds__(256) links_group(const LinkRec *__restrict__ in, uint64_t n, uint32_t u_first, const uint64_t *__restrict__ link_off,
                                                   uint32_t *cursor, LinkRec *out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        LinkRec r = in[i];
        uint32_t lu = r.from - u_first;
        out[link_off[lu] + atomicAdd(&cursor[lu], 1u)] = r;
    }
}

// nucleotide offset of every unitig (indexed by unitig id) from the scans indexed by head piece
__global__ void __launch_bounds__(256) glue_unitig_offsets(const uint64_t *__restrict__ pinfo, const uint64_t *__restrict__ uid_scan,
                                                           const uint64_t *__restrict__ nt_scan, uint64_t G, uint64_t *unt_off) {
    for (uint64_t g = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < G; g += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t inf = pinfo[g];
        if (((uint32_t)inf >> 1) == 0) unt_off[uid_scan[g]] = nt_scan[g];    // head piece (offset 0)
    }
}

// edges of locally cut cycles are packed local end ids: make them (gid, gid)
__global__ void __launch_bounds__(256) shard_local_edges_to_gid(const uint64_t *__restrict__ edges, uint64_t n, uint32_t rank, ulonglong2 *out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t e = edges[i];
        out[i] = make_ulonglong2(((uint64_t)rank << 32) | (e >> 32), ((uint64_t)rank << 32) | (uint32_t)e);
    }
}

// piece-level cycle cuts produce edges between piece termini: translate to k-mer end gids
__global__ void __launch_bounds__(256) glue_cut_edges(const uint64_t *__restrict__ pedges, uint64_t n, const PieceDesc *__restrict__ gathered, uint64_t max_p,
                                                      uint32_t world, const uint64_t *__restrict__ piece_base, uint32_t rank, ulonglong2 *edges,
                                                      uint64_t edge_cap, unsigned long long *n_edges) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t pe = pedges[i];
        uint32_t a = (uint32_t)(pe >> 32), b = (uint32_t)pe, ra, rb;
        uint64_t ga = piece_desc(gathered, max_p, world, piece_base, a >> 1, ra).term[a & 1];
        uint64_t gb = piece_desc(gathered, max_p, world, piece_base, b >> 1, rb).term[b & 1];
        if (ra != rank) continue;                              // the GPU owning the source end records the edge
        uint64_t pos = atomicAdd(n_edges, 1ULL);
        if (pos < edge_cap) edges[pos] = make_ulonglong2(ga, gb);
    }
}
