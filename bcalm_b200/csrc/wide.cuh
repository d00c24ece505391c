// wide.cuh -- the reads -> unitigs kernels for multi-word k-mers (32 <= k <= 127, W = 2..4 64-bit words).
//
// Same stages and the same decision rules as kernels.cuh; what changes is the integer type (the reference's LargeInt<N>,
// G/tools/math/LargeInt.hpp: little-endian words, comparison most-significant word first :502-509) and how a hash-table slot is
// claimed: a W-word key cannot be installed by one CAS, so every slot has a 32-bit TAG
//     0 = empty -> (CAS) 1 = locked, the winner writes the key words -> fingerprint | 0x80000000 = ready
// and probers compare the full key only when the fingerprint matches.  This path favours clarity over speed (plain loads, no TMA,
// no tile-local sort); the k <= 31 path in kernels.cuh / k1_tma.cuh is the tuned one.
#pragma once
#include "kernels.cuh"
#include "shard.cuh"

template <int W> struct KM { uint64_t w[W]; };

template <int W> __device__ __forceinline__ bool km_less(const KM<W> &a, const KM<W> &b) {
#pragma unroll
    for (int i = W - 1; i >= 0; i--) { if (a.w[i] != b.w[i]) return a.w[i] < b.w[i]; }
    return false;
}
template <int W> __device__ __forceinline__ bool km_eq(const KM<W> &a, const KM<W> &b) {
    bool e = true;
#pragma unroll
    for (int i = 0; i < W; i++) e = e && a.w[i] == b.w[i];
    return e;
}
template <int W> __device__ __forceinline__ void km_mask(KM<W> &x, int n) {   // keep the low 2n bits
    const int bits = 2 * n;
#pragma unroll
    for (int i = 0; i < W; i++) {
        const int lo = 64 * i;
        if (bits <= lo) x.w[i] = 0;
        else if (bits < lo + 64) x.w[i] &= ((1ULL << (bits - lo)) - 1ULL);
    }
}
template <int W> __device__ __forceinline__ void km_shl2(KM<W> &x) {
#pragma unroll
    for (int i = W - 1; i > 0; i--) x.w[i] = (x.w[i] << 2) | (x.w[i - 1] >> 62);
    x.w[0] <<= 2;
}
template <int W> __device__ __forceinline__ void km_shr2(KM<W> &x) {
#pragma unroll
    for (int i = 0; i < W - 1; i++) x.w[i] = (x.w[i] >> 2) | (x.w[i + 1] << 62);
    x.w[W - 1] >>= 2;
}
template <int W> __device__ __forceinline__ void km_or_at(KM<W> &x, uint64_t code, int nt_pos) {   // nt_pos from the least significant nt
    const int bit = 2 * nt_pos, wi = bit >> 6;
#pragma unroll
    for (int i = 0; i < W; i++) if (i == wi) x.w[i] |= code << (bit & 63);
}
template <int W> __device__ __forceinline__ uint32_t km_nt(const KM<W> &x, int nt_pos) {
    const int bit = 2 * nt_pos, wi = bit >> 6;
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < W; i++) if (i == wi) v = x.w[i];
    return (uint32_t)(v >> (bit & 63)) & 3u;
}
// reverse complement of the n-mer in the low 2n bits: reverse the 2-bit groups of the whole 64W-bit value, complement, shift down
template <int W> __device__ __forceinline__ KM<W> km_revcomp(const KM<W> &x, int n) {
    KM<W> r;
#pragma unroll
    for (int i = 0; i < W; i++) r.w[i] = revcomp64(x.w[W - 1 - i], 32);   // full-word reverse complement
    const int sh = 64 * W - 2 * n;                                           // drop the (complemented) padding
    const int ws = sh >> 6, bs = sh & 63;
    KM<W> o;
#pragma unroll
    for (int i = 0; i < W; i++) {
        uint64_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < W; j++) { if (j == i + ws) lo = r.w[j]; if (j == i + ws + 1) hi = r.w[j]; }
        o.w[i] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
    }
    return o;
}
template <int W> __device__ __forceinline__ uint64_t km_hash(const KM<W> &x) {
    uint64_t h = 0;
#pragma unroll
    for (int i = 0; i < W; i++) h = mix64(h ^ x.w[i]);
    return h;
}
template <int W> __device__ __forceinline__ KM<W> km_load(const uint64_t *p) { KM<W> x;
#pragma unroll
    for (int i = 0; i < W; i++) x.w[i] = p[i];
    return x; }
template <int W> __device__ __forceinline__ void km_store(uint64_t *p, const KM<W> &x) {
#pragma unroll
    for (int i = 0; i < W; i++) p[i] = x.w[i];
}

// ------------------------------------------------------------------------------------------------ hash-table slot protocol

#define TAG_EMPTY 0u
#define TAG_LOCKED 1u
__device__ __forceinline__ uint32_t tag_of(uint64_t h) { return (uint32_t)(h >> 33) | 0x80000000u; }

// find or claim the slot of `key`; returns the slot (or 0xFFFFFFFF on overflow)
template <int W>
__device__ __forceinline__ uint32_t wide_find_or_insert(const KM<W> &key, uint64_t h, uint32_t *tags, uint64_t *keys, uint32_t cap) {
    const uint32_t tag = tag_of(h);
    uint32_t slot = (uint32_t)(((h & 0xFFFFFFFFULL) * cap) >> 32), probes = 0;
    for (;;) {
        uint32_t cur = *(volatile uint32_t *)&tags[slot];
        if (cur == TAG_EMPTY) {
            cur = atomicCAS(&tags[slot], TAG_EMPTY, TAG_LOCKED);
            if (cur == TAG_EMPTY) {
                km_store<W>(keys + (size_t)slot * W, key);
                __threadfence();
                *(volatile uint32_t *)&tags[slot] = tag;
                return slot;
            }
        }
        while (cur == TAG_LOCKED) cur = *(volatile uint32_t *)&tags[slot];       // the owner is writing the key words
        if (cur == tag) {
            __threadfence();
            KM<W> k2;
#pragma unroll
            for (int i = 0; i < W; i++) k2.w[i] = *(volatile uint64_t *)&keys[(size_t)slot * W + i];
            if (km_eq<W>(k2, key)) return slot;
        }
        slot = slot + 1 == cap ? 0 : slot + 1;
        if (++probes > min(cap, 8192u)) return 0xFFFFFFFFu;
    }
}

// ------------------------------------------------------------------------------------------------ K1 (wide)

constexpr int KW1_THREADS = 256;
constexpr int KW1_PER = 64;                                // positions per thread
constexpr int KW1_TILE = KW1_THREADS * KW1_PER;            // 16384
constexpr int KW1_HALO = 128;                              // >= k - 1

struct KW1Params {
    const uint8_t *bytes; uint64_t n_bytes, n_tiles; int k; uint32_t P;
    unsigned long long *bin_count; const uint64_t *bin_off; unsigned long long *bin_cursor;
    uint64_t *out_kmers;       // [I * W]
};

__device__ __forceinline__ bool nt_valid_dev(uint8_t c) { uint8_t u = c & 0xDF; return u == 'A' || u == 'C' || u == 'G' || u == 'T'; }

// MODE 0 histogram, MODE 1 scatter (two rolling sweeps per tile: count, then place into the run reserved per (tile, bin))
template <int W, int MODE>
__global__ void __launch_bounds__(KW1_THREADS) kw1_extract(KW1Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint8_t *s_raw = smem_raw;                                                             // [TILE + HALO]
    uint64_t *s_base = reinterpret_cast<uint64_t *>(smem_raw + KW1_TILE + KW1_HALO);       // [P]  MODE 1
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_base + (MODE == 1 ? p.P : 0));        // [P]
    uint32_t *s_cur = s_cnt + p.P;                                                         // [P]  MODE 1
    const int t = threadIdx.x, k = p.k;
    for (uint32_t i = t; i < p.P; i += KW1_THREADS) s_cnt[i] = 0;
    __syncthreads();

    for (uint64_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint64_t start = tile * (uint64_t)KW1_TILE;
        for (int i = t * 16; i < KW1_TILE + KW1_HALO; i += KW1_THREADS * 16) {
            uint64_t q = start + i;
            if (q + 16 <= p.n_bytes) *reinterpret_cast<uint4 *>(s_raw + i) = __ldg(reinterpret_cast<const uint4 *>(p.bytes + q));
            else for (int j = 0; j < 16; j++) s_raw[i + j] = (q + j < p.n_bytes) ? p.bytes[q + j] : (uint8_t)'\n';
        }
        __syncthreads();

        auto sweep = [&](auto &&emit) {
            KM<W> fwd, rc;
#pragma unroll
            for (int i = 0; i < W; i++) { fwd.w[i] = 0; rc.w[i] = 0; }
            int run = 0;
            const int base = t * KW1_PER;
            for (int i = 0; i < KW1_PER + k - 1; i++) {          // k-1 warm-up bases, then one k-mer per step
                const uint8_t c = s_raw[base + i];
                if (!nt_valid_dev(c)) { run = 0; continue; }
                const uint64_t code = (c >> 1) & 3;
                km_shl2<W>(fwd); fwd.w[0] |= code; km_mask<W>(fwd, k);
                km_shr2<W>(rc); km_or_at<W>(rc, code ^ 2, k - 1);
                if (++run >= k && i >= k - 1) emit(km_less<W>(fwd, rc) ? fwd : rc);
            }
        };
        if (MODE == 0) {
            sweep([&](const KM<W> &c) { atomicAdd(&s_cnt[hash_to_bin(km_hash<W>(c), p.P)], 1u); });
            __syncthreads();
        } else {
            sweep([&](const KM<W> &c) { atomicAdd(&s_cnt[hash_to_bin(km_hash<W>(c), p.P)], 1u); });
            __syncthreads();
            for (uint32_t b = t; b < p.P; b += KW1_THREADS) {
                uint32_t cn = s_cnt[b];
                if (cn) s_base[b] = p.bin_off[b] + atomicAdd(&p.bin_cursor[b], (unsigned long long)cn);
                s_cnt[b] = 0; s_cur[b] = 0;
            }
            __syncthreads();
            sweep([&](const KM<W> &c) {
                uint32_t b = hash_to_bin(km_hash<W>(c), p.P);
                uint64_t pos = s_base[b] + atomicAdd(&s_cur[b], 1u);
                km_store<W>(p.out_kmers + pos * W, c);
            });
            __syncthreads();
        }
    }
    if (MODE == 0) {
        for (uint32_t i = t; i < p.P; i += KW1_THREADS) { uint32_t cn = s_cnt[i]; if (cn) atomicAdd(&p.bin_count[i], (unsigned long long)cn); }
    }
}
static inline size_t kw1_smem_bytes(int mode, uint32_t P) { return (size_t)KW1_TILE + KW1_HALO + (size_t)P * 4 + (mode == 1 ? (size_t)P * 12 : 0) + 32; }

// ------------------------------------------------------------------------------------------------ K2 (wide)

template <int W>
__global__ void __launch_bounds__(256) kw2_insert(const uint64_t *__restrict__ kmers, uint64_t n, uint32_t *tags, uint64_t *keys, uint32_t *counts,
                                                  uint32_t cap, int *overflow) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        KM<W> key = km_load<W>(kmers + i * W);
        uint32_t slot = wide_find_or_insert<W>(key, km_hash<W>(key) * 0x9E3779B97F4A7C15ULL + 0x7F4A7C15ULL, tags, keys, cap);
        if (slot == 0xFFFFFFFFu) { *overflow = 1; continue; }
        atomicAdd(&counts[slot], 1u);
    }
}

// two-pass warp-chunk scan like k2_emit
template <int W>
__global__ void __launch_bounds__(256) kw2_emit(uint32_t *tags, uint64_t *keys, uint32_t *counts, uint32_t cap, uint32_t amin, uint32_t amax,
                                                uint64_t *solid_kmers, uint32_t *solid_counts, uint64_t solid_cap, unsigned long long *n_solid,
                                                unsigned long long *n_distinct) {
    __shared__ uint32_t s_w[8];
    __shared__ unsigned long long s_base;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t n_warps = gridDim.x * 8;
    const uint32_t per_warp = ((cap + n_warps - 1) / n_warps + 31u) & ~31u;
    const uint32_t gw = blockIdx.x * 8 + warp;
    const uint32_t wb = (uint32_t)min((uint64_t)cap, (uint64_t)gw * per_warp), we = (uint32_t)min((uint64_t)cap, (uint64_t)wb + per_warp);
    uint32_t my_solid = 0, my_distinct = 0;
    for (uint32_t i = wb + lane; i < we; i += 32) {
        if (tags[i] < 2) continue;
        uint32_t c = counts[i];
        my_distinct++;
        my_solid += c >= amin && c <= amax;
    }
    uint32_t ws = my_solid;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ws += __shfl_xor_sync(0xFFFFFFFFu, ws, d);
    if (lane == 0) s_w[warp] = ws;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tot = 0;
        for (int x = 0; x < 8; x++) { uint32_t v = s_w[x]; s_w[x] = tot; tot += v; }
        s_base = tot ? atomicAdd(n_solid, (unsigned long long)tot) : 0ULL;
    }
    __syncthreads();
    uint64_t out = s_base + s_w[warp];
    for (uint32_t base = wb; base < we; base += 32) {
        uint32_t i = base + lane;
        bool occ = i < we && tags[i] >= 2;
        uint32_t c = occ ? counts[i] : 0;
        bool solid = occ && c >= amin && c <= amax;
        unsigned m = __ballot_sync(0xFFFFFFFFu, solid);
        if (solid) {
            uint64_t pos = out + __popc(m & ((1u << lane) - 1));
            if (pos < solid_cap) {
#pragma unroll
                for (int j = 0; j < W; j++) solid_kmers[pos * W + j] = keys[(size_t)i * W + j];
                solid_counts[pos] = c;
            }
        }
        out += __popc(m);
        if (occ) { tags[i] = TAG_EMPTY; counts[i] = 0; }
    }
    block_add(n_distinct, (unsigned long long)my_distinct);
}

// counted k-mers handed in from outside (cdbg_set_solid_kmers, k >= 32): mask to 2k bits and bring to the canonical orientation
template <int W>
__global__ void __launch_bounds__(256) kw2_canonicalise(uint64_t *kmers, uint64_t n, int k) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        KM<W> x = km_load<W>(kmers + i * W);
        km_mask<W>(x, k);
        const KM<W> r = km_revcomp<W>(x, k);
        km_store<W>(kmers + i * W, km_less<W>(r, x) ? r : x);
    }
}

// ------------------------------------------------------------------------------------------------ K3 (wide)

template <int W>
__device__ __forceinline__ void kmer_end_record_w(const KM<W> &x, int k, int end, uint32_t idx, KM<W> &key, uint64_t &val) {
    const int k1 = k - 1;
    KM<W> e = x;
    if (end == 0) km_shr2<W>(e); else km_mask<W>(e, k1);
    KM<W> r = km_revcomp<W>(e, k1);
    const bool lt = km_less<W>(e, r), eq = km_eq<W>(e, r);
    key = lt ? e : r;
    const bool canon_fwd = lt || eq;
    const bool out = (end == 0) ? canon_fwd : !canon_fwd;
    val = (uint64_t)(2u * idx + (uint32_t)end) | (out ? REC_SIDE_OUT : 0) | (eq ? REC_PALIN : 0);
    if (!(k & 1) && km_eq<W>(x, km_revcomp<W>(x, k))) val |= end == 0 ? REC_BOTH : REC_SKIP;   // see REC_BOTH in kernels.cuh
}

struct KW3Params {
    const uint64_t *solid_kmers; uint64_t n_solid; int k; uint32_t P;
    unsigned long long *bin_count; const uint64_t *bin_off; unsigned long long *bin_cursor;
    uint64_t *rec_key;   // [2S * W]
    uint64_t *rec_val;   // [2S]
    int sharded; uint32_t rank;   // sharded path: values carry the gid (rank << 32 | end) and the SREC_* flags of shard.cuh
};

// per-tile shared-memory aggregation like k3_ends: the hot per-bin cursors are hit once per (tile, bin)
template <int W, int MODE>
__global__ void __launch_bounds__(K3_THREADS) kw3_ends(KW3Params p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *s_base = reinterpret_cast<uint64_t *>(smem_raw);
    uint32_t *s_cnt = reinterpret_cast<uint32_t *>(s_base + (MODE == 1 ? p.P : 0));
    const int t = threadIdx.x;
    const uint64_t n_tiles = (p.n_solid + K3_THREADS - 1) / K3_THREADS;
    for (uint32_t i = t; i < p.P; i += K3_THREADS) s_cnt[i] = 0;
    __syncthreads();
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t i = tile * K3_THREADS + t;
        const bool ok = i < p.n_solid;
        KM<W> key[2]; uint64_t val[2]; uint32_t bin[2] = {0xFFFFFFFFu, 0xFFFFFFFFu}, rnk[2] = {0, 0};
        if (ok) {
            KM<W> x = km_load<W>(p.solid_kmers + i * W);
#pragma unroll
            for (int end = 0; end < 2; end++) {
                kmer_end_record_w<W>(x, p.k, end, (uint32_t)i, key[end], val[end]);
                if (p.sharded) {
                    const uint64_t v = val[end];
                    val[end] = ((uint64_t)p.rank << 32) | (uint32_t)v | ((v & REC_SIDE_OUT) ? SREC_SIDE_OUT : 0) | ((v & REC_PALIN) ? SREC_PALIN : 0) |
                               ((v & REC_BOTH) ? SREC_BOTH : 0) | ((v & REC_SKIP) ? SREC_SKIP : 0);
                }
                bin[end] = hash_to_bin(km_hash<W>(key[end]), p.P);
                rnk[end] = atomicAdd(&s_cnt[bin[end]], 1u);
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
            for (int end = 0; end < 2; end++) {
                if (bin[end] == 0xFFFFFFFFu) continue;
                uint64_t pos = s_base[bin[end]] + rnk[end];
                km_store<W>(p.rec_key + pos * W, key[end]); p.rec_val[pos] = val[end];
            }
            __syncthreads();
        }
    }
    if (MODE == 0) {
        __syncthreads();
        for (uint32_t b = t; b < p.P; b += K3_THREADS) { uint32_t c = s_cnt[b]; if (c) atomicAdd(&p.bin_count[b], (unsigned long long)c); }
    }
}

// ------------------------------------------------------------------------------------------------ K5 (wide)

struct JoinTableW { uint32_t *tags; uint64_t *keys; uint32_t *cnt; uint32_t *ids; uint32_t cap; };

template <int W>
__global__ void __launch_bounds__(256) kw5_insert(const uint64_t *__restrict__ rec_key, const uint64_t *__restrict__ rec_val, uint64_t n, JoinTableW T,
                                                  int *overflow) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        KM<W> key = km_load<W>(rec_key + i * W);
        const uint64_t val = rec_val[i];
        if (val & REC_SKIP) continue;
        uint32_t slot = wide_find_or_insert<W>(key, km_hash<W>(key) * 0x9E3779B97F4A7C15ULL + 0x7F4A7C15ULL, T.tags, T.keys, T.cap);
        if (slot == 0xFFFFFFFFu) { *overflow = 1; continue; }
        const bool pal = (val & REC_PALIN) != 0, use_out = (val & REC_SIDE_OUT) != 0 && !pal;
        uint32_t old = atomicAdd(&T.cnt[slot], use_out ? 0x100u : 0x1u);
        if (pal) atomicOr(&T.cnt[slot], 0x80000000u);
        uint32_t pos = use_out ? ((old >> 8) & 0xFF) : (old & 0xFF);
        if (pos < 4) {
            T.ids[(size_t)slot * 8 + (use_out ? 4 : 0) + pos] = (uint32_t)val;
            if (val & REC_BOTH) atomicOr(&T.cnt[slot], 1u << ((use_out ? JOIN_BOTH_SHIFT_OUT : JOIN_BOTH_SHIFT_IN) + pos));
        }
    }
}

// identical decision rule to k5_resolve; occupancy is read from the tags
__global__ void __launch_bounds__(256) kw5_resolve(JoinTableW T, uint32_t cap, uint32_t *link, uint64_t *edges, uint64_t edge_cap,
                                                   unsigned long long *n_edges, unsigned long long *n_compact) {
    unsigned long long compact = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        if (T.tags[i] < 2) continue;
        uint32_t c = T.cnt[i];
        T.tags[i] = TAG_EMPTY; T.cnt[i] = 0;
        uint32_t nin = min(c & 0xFFu, 4u), nout = min((c >> 8) & 0xFFu, 4u); bool pal = (c >> 31) != 0;
        uint32_t in[5], out[5];
        {
            const uint4 vin = *reinterpret_cast<const uint4 *>(&T.ids[(size_t)i * 8]);
            const uint4 vout = *reinterpret_cast<const uint4 *>(&T.ids[(size_t)i * 8 + 4]);
            const uint32_t sin[4] = {vin.x, vin.y, vin.z, vin.w}, sout[4] = {vout.x, vout.y, vout.z, vout.w};
            nin = join_expand_side(sin, nin, (c >> JOIN_BOTH_SHIFT_IN) & 0xFu, in);
            nout = join_expand_side(sout, nout, (c >> JOIN_BOTH_SHIFT_OUT) & 0xFu, out);
        }
        if (!pal && nin == 1 && nout == 1 && (in[0] >> 1) != (out[0] >> 1)) { link[in[0]] = out[0]; link[out[0]] = in[0]; compact++; continue; }
        uint32_t ne = pal ? nin * nin : 2 * nin * nout;
        if (ne == 0) continue;
        unsigned long long base = atomicAdd(n_edges, (unsigned long long)ne);
        if (base + ne > edge_cap) continue;
        if (pal) { for (uint32_t a = 0; a < nin; a++) for (uint32_t b = 0; b < nin; b++) edges[base++] = ((uint64_t)in[a] << 32) | in[b]; }
        else for (uint32_t a = 0; a < nin; a++) for (uint32_t b = 0; b < nout; b++) { edges[base++] = ((uint64_t)in[a] << 32) | out[b]; edges[base++] = ((uint64_t)out[b] << 32) | in[a]; }
    }
    block_add(n_compact, compact);
}

// Sharded join on multi-word keys: the slot is found through the tag protocol; the per-slot state lives in a JoinTable64 (64-bit
// gids, like the k <= 31 sharded join) whose `keys` array only marks occupancy, so that k5s_resolve serves both widths.
template <int W>
__global__ void __launch_bounds__(256) kw5s_insert(const uint64_t *__restrict__ rec_key, const uint64_t *__restrict__ rec_val, uint64_t n, uint32_t *tags,
                                                   uint64_t *wkeys, JoinTable64 T, int *overflow) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const KM<W> key = km_load<W>(rec_key + i * W);
        const uint64_t val = rec_val[i];
        if (val & SREC_SKIP) continue;
        const uint32_t slot = wide_find_or_insert<W>(key, km_hash<W>(key) * 0x9E3779B97F4A7C15ULL + 0x7F4A7C15ULL, tags, wkeys, T.cap);
        if (slot == 0xFFFFFFFFu) { *overflow = 1; continue; }
        T.keys[slot] = 0;                              // occupied (anything but EMPTY_KEY; every writer stores the same value)
        const bool pal = (val & SREC_PALIN) != 0, use_out = (val & SREC_SIDE_OUT) != 0 && !pal;
        const uint32_t old = atomicAdd(&T.cnt[slot], use_out ? 0x100u : 0x1u);
        if (pal) atomicOr(&T.cnt[slot], 0x80000000u);
        const uint32_t pos = use_out ? ((old >> 8) & 0xFF) : (old & 0xFF);
        if (pos < 4) {
            T.ids[(size_t)slot * 8 + (use_out ? 4 : 0) + pos] = val & GID_MASK;
            if (val & SREC_BOTH) atomicOr(&T.cnt[slot], 1u << ((use_out ? JOIN_BOTH_SHIFT_OUT : JOIN_BOTH_SHIFT_IN) + pos));
        }
    }
}

// ------------------------------------------------------------------------------------------------ spell (wide)

template <int W>
__global__ void __launch_bounds__(256) kw8_spell(SpellParams p) {
    const char NT[4] = {'A', 'C', 'T', 'G'};
    for (uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; v < p.S; v += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t inf = p.info[v];
        uint32_t B = (uint32_t)(inf >> 32), pos = (uint32_t)inf >> 1, rev = (uint32_t)inf & 1u;
        uint32_t h = B >> 1;
        uint64_t u = p.uid_scan[h], off = p.nt_scan[h];
        KM<W> x = km_load<W>(p.solid_kmers + v * W);
        if (rev) x = km_revcomp<W>(x, p.k);
        if (p.term_begin) {                          // sharded path: the k-mer ends at which this piece begins and ends (see k8_spell)
            if (pos == 0) p.term_begin[u] = B;
            const uint32_t exit_end = 2u * (uint32_t)v + (rev ? 0u : 1u);
            if (p.link[exit_end] == NOLINK) p.term_end[u] = exit_end;
        }
        if (pos == 0) {
            p.seq_off[u] = off;
            for (int j = 0; j < p.k; j++) p.seq[off + j] = NT[km_nt<W>(x, p.k - 1 - j)];
        } else p.seq[off + (p.k - 1) + pos] = NT[km_nt<W>(x, 0)];
        uint32_t c = p.solid_counts[v];
        atomicAdd(&p.kc[u], (unsigned long long)c);
        if (p.abund) p.abund[off - u * (uint64_t)(p.k - 1) + pos] = c;
        p.unitig_of_kmer[v] = (uint32_t)u;
    }
}
