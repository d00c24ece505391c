// parse.cuh -- FASTA / FASTQ text -> nucleotide stream on the GPU (the reference's parser is
// G/bank/impl/BankFasta.cpp:390-560, the host-side equivalent here is parse_sequences_raw in bcalm_main.cpp).
//
// Output grammar (what cdbg_push_reads expects): the sequence characters of every record, lines joined, the '\r' of CRLF line
// ends dropped, one '\n' after each record; headers, '+' lines and quality lines are gone.  Any other byte is passed through
// (K1 treats everything that is not ACGTacgt as a break).
//
// A parser is a small state machine over lines, and the state at a byte depends on everything before it -- a scan.  The state
// space is tiny (<= 4 states for both formats), so a chunk of text is summarised by its TRANSITION FUNCTION start state -> end
// state, packed into 8 bits (2 bits per start state); function composition is associative, so chunks combine with an ordinary
// parallel scan:   pass 1  every thread composes the function of its 32 bytes, the block reduces to one function per tile
//                  pass 2  one block scans the tile functions from the initial state -> state at the start of every tile
//                  pass 3  every thread gets its start state from a block scan, walks its 32 bytes, counts what it keeps
//                  (scan of the per-tile counts)
//                  pass 4  the same walk, writing at tile offset + thread offset
// FASTA states: 0 before the first record, 1 inside a header line, 2 inside a sequence line of a record.
// FASTQ states: line number mod 4 (0 '@' header, 1 sequence, 2 '+', 3 quality); strict 4-line records -- a file that does not
// keep to them (multi-line FASTQ, stray lines) raises a flag and the caller parses on the host instead.
#pragma once
#include "kernels.cuh"

constexpr int PARSE_THREADS = 256;
constexpr int PARSE_PER_THREAD = 32;
constexpr int PARSE_TILE = PARSE_THREADS * PARSE_PER_THREAD;   // 8192 bytes

enum { PARSE_FASTA = 0, PARSE_FASTQ = 1 };
#define PARSE_IDENTITY 0xE4u   // the function s -> s: (3 << 6) | (2 << 4) | (1 << 2) | 0

__device__ __forceinline__ uint32_t parse_apply(uint32_t f, uint32_t s) { return (f >> (2 * s)) & 3u; }
// (g after f): first f, then g
__device__ __forceinline__ uint32_t parse_compose(uint32_t f, uint32_t g) {
    return parse_apply(g, parse_apply(f, 0)) | (parse_apply(g, parse_apply(f, 1)) << 2) | (parse_apply(g, parse_apply(f, 2)) << 4) |
           (parse_apply(g, parse_apply(f, 3)) << 6);
}
// state after the line start whose first byte is c
template <int FMT>
__device__ __forceinline__ uint32_t parse_line_start(uint32_t s, uint8_t c) {
    if (FMT == PARSE_FASTA) return c == '>' ? 1u : (s == 0 ? 0u : 2u);
    return (s + 1) & 3u;
}
__device__ __forceinline__ uint32_t parse_initial_state(int fmt) { return fmt == PARSE_FASTA ? 0u : 3u; }   // FASTQ: the first line start gives 0

struct ParseParams {
    const uint8_t *text; uint64_t n;
    uint32_t n_tiles;
    uint8_t *tile_fn;        // [n_tiles]  pass 1: transition function of each tile
    uint8_t *tile_state;     // [n_tiles + 1] pass 2: state at the start of each tile, [n_tiles] = final state
    uint32_t *tile_cnt;      // [n_tiles]  pass 3: bytes kept per tile
    const uint64_t *tile_off;// [n_tiles]  pass 4: exclusive offsets
    uint8_t *out;            // pass 4
    unsigned long long *n_records;
    int *bad;                // FASTQ: a line that should start with '@' / '+' does not
};

// transition function of bytes [beg, end): every byte that begins a line applies parse_line_start
template <int FMT>
__device__ __forceinline__ uint32_t parse_chunk_fn(const uint8_t *text, uint64_t beg, uint64_t end) {
    uint32_t f = PARSE_IDENTITY;
    uint8_t prev = beg ? text[beg - 1] : (uint8_t)'\n';
    for (uint64_t i = beg; i < end; i++) {
        const uint8_t c = text[i];
        if (prev == '\n') {
            const uint32_t g = parse_line_start<FMT>(0, c) | (parse_line_start<FMT>(1, c) << 2) | (parse_line_start<FMT>(2, c) << 4) |
                               (parse_line_start<FMT>(3, c) << 6);
            f = parse_compose(f, g);
        }
        prev = c;
    }
    return f;
}

// inclusive block scan of transition functions; returns the EXCLUSIVE prefix of this thread, *total = function of the whole block
__device__ __forceinline__ uint32_t parse_block_scan(uint32_t f, uint32_t *total) {
    __shared__ uint32_t s_w[PARSE_THREADS / 32];
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    uint32_t incl = f;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl = parse_compose(o, incl);
    }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t wpre = PARSE_IDENTITY, tot = PARSE_IDENTITY;
#pragma unroll
    for (int w = 0; w < PARSE_THREADS / 32; w++) { const uint32_t v = s_w[w]; if (w < (int)warp) wpre = parse_compose(wpre, v); tot = parse_compose(tot, v); }
    *total = tot;
    uint32_t excl = __shfl_up_sync(0xFFFFFFFFu, incl, 1);
    if (lane == 0) excl = PARSE_IDENTITY;
    __syncthreads();
    return parse_compose(wpre, excl);
}

template <int FMT>
__global__ void __launch_bounds__(PARSE_THREADS) parse_tile_fn(ParseParams p) {
    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint64_t beg = (uint64_t)tile * PARSE_TILE + (uint64_t)threadIdx.x * PARSE_PER_THREAD;
        const uint64_t end = min(beg + (uint64_t)PARSE_PER_THREAD, p.n);
        const uint32_t f = beg < p.n ? parse_chunk_fn<FMT>(p.text, beg, end) : PARSE_IDENTITY;
        uint32_t total;
        parse_block_scan(f, &total);
        if (threadIdx.x == 0) p.tile_fn[tile] = (uint8_t)total;
    }
}

// one block: state at the start of every tile.  Every thread owns a contiguous run of tiles.
__global__ void __launch_bounds__(PARSE_THREADS) parse_tile_states(ParseParams p, int fmt) {
    const uint32_t per = (p.n_tiles + PARSE_THREADS - 1) / PARSE_THREADS;
    const uint32_t t0 = min(threadIdx.x * per, p.n_tiles), t1 = min(t0 + per, p.n_tiles);
    uint32_t f = PARSE_IDENTITY;
    for (uint32_t t = t0; t < t1; t++) f = parse_compose(f, p.tile_fn[t]);
    uint32_t total;
    const uint32_t pre = parse_block_scan(f, &total);
    uint32_t s = parse_apply(pre, parse_initial_state(fmt));
    for (uint32_t t = t0; t < t1; t++) { p.tile_state[t] = (uint8_t)s; s = parse_apply(p.tile_fn[t], s); }
    if (threadIdx.x == 0) p.tile_state[p.n_tiles] = (uint8_t)parse_apply(total, parse_initial_state(fmt));
}

// WRITE = false: count the kept bytes of every tile (and the records).  WRITE = true: write them.
template <int FMT, bool WRITE>
__global__ void __launch_bounds__(PARSE_THREADS) parse_emit(ParseParams p) {
    __shared__ uint32_t s_wcnt[PARSE_THREADS / 32];
    __shared__ uint32_t s_tile_total;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    unsigned long long records = 0;
    for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        const uint64_t beg = (uint64_t)tile * PARSE_TILE + (uint64_t)threadIdx.x * PARSE_PER_THREAD;
        const uint64_t end = min(beg + (uint64_t)PARSE_PER_THREAD, p.n);
        const uint32_t f = beg < p.n ? parse_chunk_fn<FMT>(p.text, beg, end) : PARSE_IDENTITY;
        uint32_t total;
        const uint32_t pre = parse_block_scan(f, &total);
        uint32_t s = parse_apply(pre, p.tile_state[tile]);

        // walk the chunk: what is kept, in order
        uint8_t keep[PARSE_PER_THREAD];
        uint32_t nk = 0;
        if (beg < p.n) {
            uint8_t prev = beg ? p.text[beg - 1] : (uint8_t)'\n';
            for (uint64_t i = beg; i < end; i++) {
                const uint8_t c = p.text[i];
                const uint32_t before = s;
                const bool line_start = prev == '\n';
                if (line_start) s = parse_line_start<FMT>(s, c);
                const bool cr_at_eol = c == '\r' && (i + 1 >= p.n || p.text[i + 1] == '\n');
                if (FMT == PARSE_FASTA) {
                    if (line_start && s == 1) { records++; if (before != 0) keep[nk++] = '\n'; }      // a header closes the previous record
                    else if (s == 2 && c != '\n' && !cr_at_eol) keep[nk++] = c;
                } else {
                    if (line_start) {
                        if (s == 1) records++;
                        if ((s == 0 && c != '@') || (s == 2 && c != '+')) *p.bad = 1;
                    }
                    if (s == 1 && !cr_at_eol) keep[nk++] = c;                                           // (the line's own '\n' ends the record)
                }
                prev = c;
            }
        }
        // exclusive scan of nk over the block
        uint32_t incl = nk;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += o; }
        if (lane == 31) s_wcnt[warp] = incl;
        __syncthreads();
        uint32_t wpre = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < PARSE_THREADS / 32; w++) { const uint32_t v = s_wcnt[w]; if (w < (int)warp) wpre += v; tot += v; }
        if (!WRITE) { if (threadIdx.x == 0) p.tile_cnt[tile] = tot; }
        else {
            uint8_t *dst = p.out + p.tile_off[tile] + wpre + incl - nk;
            for (uint32_t j = 0; j < nk; j++) dst[j] = keep[j];
        }
        (void)s_tile_total;
        __syncthreads();
    }
    if (!WRITE) block_add(p.n_records, records);
}
