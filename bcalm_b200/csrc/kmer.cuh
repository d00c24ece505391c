// kmer.cuh -- 2-bit k-mer arithmetic shared by every kernel (k <= 31: one 64-bit word).
//
// Encoding follows the reference so that canonical forms (an integer comparison) are identical:
//   code = (c >> 1) & 3  ->  A=0 C=1 T=2 G=3, complement = code ^ 2      (G/tools/misc/api/Data.hpp:185)
//   first nucleotide in the most significant position                    (G/kmer/impl/Model.hpp:857-884)
//   canonical = min(fwd, revcomp) as integers                            (Model.hpp:294)
#pragma once
#include <stdint.h>

#define CDBG_HD __host__ __device__ __forceinline__

static CDBG_HD uint64_t kmer_mask(int k) { return k >= 32 ? ~0ULL : ((1ULL << (2 * k)) - 1ULL); }

// reverse complement of the k-mer held in the low 2k bits (same function as LargeInt1.pri:137-155 computes with a table)
static CDBG_HD uint64_t revcomp64(uint64_t x, int k) {
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
#ifdef __CUDA_ARCH__
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    lo = __byte_perm(lo, 0, 0x0123);
    hi = __byte_perm(hi, 0, 0x0123);
    x = ((uint64_t)lo << 32) | hi;
#else
    x = __builtin_bswap64(x);
#endif
    x ^= 0xAAAAAAAAAAAAAAAAULL;  // complement every 2-bit group (code ^ 2)
    return x >> (64 - 2 * k);
}

static CDBG_HD uint64_t canonical64(uint64_t x, int k) {
    uint64_t r = revcomp64(x, k);
    return x < r ? x : r;
}

// murmur3 fmix64: the partition hash (high bits -> bin) and the table hash (low bits -> slot)
static CDBG_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

// bin in [0, P) from the high 32 bits of the hash (multiply-shift, P need not be a power of two)
static CDBG_HD uint32_t hash_to_bin(uint64_t h, uint32_t P) {
    return (uint32_t)(((h >> 32) * (uint64_t)P) >> 32);
}

// order of canonical m-mers for minimizers: an invertible 32-bit mix (xor a constant, one odd multiply, one xor-shift: 4 instructions
// on the hottest line of K1 instead of the 9 of murmur3's fmix32), identical to orc_mmer_rank in oracle/cdbg_oracle.c.  Being a
// bijection, min over ranks identifies a unique m-mer; the constant keeps low-complexity m-mers (poly-A = 0) from ranking first.
static CDBG_HD uint32_t mmer_rank(uint32_t cm) {
    uint32_t h = (cm ^ 0x9E3779B9u) * 0x85EBCA6Bu;
    return h ^ (h >> 15);
}

// minimizer of the n-mer x: best canonical m-mer under (rank, value); straightforward form used outside the hot loop
static CDBG_HD uint32_t minimizer_of(uint64_t x, int n, int m) {
    const uint64_t mm = kmer_mask(m);
    uint32_t best = 0, best_r = 0xFFFFFFFFu;
    bool have = false;
    for (int j = 0; j + m <= n; j++) {
        uint64_t mer = (x >> (2 * (n - m - j))) & mm;
        uint32_t c = (uint32_t)canonical64(mer, m);
        uint32_t r = mmer_rank(c);
        if (!have || r < best_r || (r == best_r && c < best)) { best = c; best_r = r; have = true; }
    }
    return best;
}

// cheap 32-bit bin hash for K1 (10 integer ops instead of two 64-bit multiplies); independent of table_slot's mix64
static CDBG_HD uint32_t bin_hash32(uint64_t c) {
    uint32_t h = ((uint32_t)c * 0x9E3779B1u) ^ ((uint32_t)(c >> 32) * 0x85EBCA77u + 0x165667B1u);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    return h;
}
static CDBG_HD uint32_t hash32_to_bin(uint32_t h, uint32_t P) { return (uint32_t)(((uint64_t)h * P) >> 32); }
