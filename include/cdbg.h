/*
 * include/cdbg.h -- C ABI of the B200-native compacted de Bruijn graph constructor (libcdbg_b200.so).
 *
 * This is the drop-in boundary for the BCALM 2 hot path  reads -> solid canonical k-mers -> compaction -> glue ->
 * links  (SURVEY.md section 8).  The reference has no C ABI: its boundary is a set of C++ templates that hand data to
 * each other through files.  Each entry point below names the reference interface it replaces
 * ("G/" = gatb-core/gatb-core/src/gatb/ inside the reference tree):
 *
 *   cdbg_create          Tool / GraphUnitigsTemplate<span> construction: option parsing + span dispatch
 *                        (src/bcalm_1.cpp:49-97, G/debruijn/impl/GraphUnitigs.cpp:832-973, G/tools/math/Integer.hpp:77-85)
 *   cdbg_push_reads      Bank iteration feeding SortingCountAlgorithm (G/kmer/impl/SortingCountAlgorithm.cpp:1211-1344,
 *                        G/bank/impl/BankFasta.cpp:390-560 delivers the same nucleotide stream)
 *   cdbg_count           SortingCountAlgorithm<span>::execute + the ICountProcessor chain Solidity -> Dump
 *                        (G/kmer/impl/SortingCountAlgorithm.cpp:636-781, G/kmer/api/ICountProcessor.hpp:105-173,
 *                         G/kmer/impl/CountProcessorSolidity.hpp:186-189, CountProcessorDump.hpp:148-152)
 *   cdbg_set_solid_kmers the "solid k-mers already on disk" entry of the same class (GraphUnitigs.cpp:884-951: a `.h5` with a
 *                        dsk/solid group skips the counting step)
 *   cdbg_compact         bcalm2<SPAN>() + bglue<SPAN>()  (G/bcalm2/bcalm_algo.hpp:29-38, G/bcalm2/bglue_algo.hpp:147-155)
 *   cdbg_link            link_tigs<SPAN>()               (G/debruijn/impl/LinkTigs.hpp:33)
 *   cdbg_get_unitigs     the <prefix>.unitigs.fa content (G/bcalm2/bglue_algo.hpp:73-90 BufferedFasta,
 *                        G/debruijn/impl/LinkTigs.cpp:114-205 write_final_output) as arrays instead of a file
 *   cdbg_build           UnitigsConstructionAlgorithm<span>::execute(), i.e. all of the above in sequence
 *                        (G/debruijn/impl/UnitigsConstructionAlgorithm.cpp:92-122)
 *
 * Conventions: plain pointers and sizes, no C++ or torch types; every call is blocking; a context is not re-entrant but
 * different contexts may be used from different threads; no exceptions and no exit() cross this boundary (the reference
 * throws system::Exception or calls exit(1): bcalm_algo.cpp:216, bglue_algo.cpp:247, LinkTigs.cpp:59) -- every function
 * returns a cdbg_status and cdbg_last_error() gives the message.  There is NO CPU fallback: without a CUDA device
 * cdbg_create fails with CDBG_ERR_CUDA.
 *
 * Input grammar: a byte stream in which A,C,G,T,a,c,g,t are nucleotides (code (c>>1)&3, A=0 C=1 T=2 G=3,
 * G/tools/misc/api/Data.hpp:185) and EVERY other byte (newline, N, IUPAC, '>' ...) is a break that invalidates all
 * k-mers covering it.  One read per line, no headers: what a FASTA/FASTQ parser hands over after stripping headers and
 * quality lines.
 */
#ifndef CDBG_H_
#define CDBG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cdbg_ctx cdbg_ctx;

typedef enum {
    CDBG_OK = 0,
    CDBG_ERR_PARAM = -1,    /* bad k / abundance / pointer ("unhandled kmer size", Integer.hpp:77-85) */
    CDBG_ERR_CUDA = -2,     /* no device, launch failure */
    CDBG_ERR_OOM = -3,      /* device or pinned host memory exhausted */
    CDBG_ERR_STATE = -4,    /* stage called out of order */
    CDBG_ERR_NCCL = -5,
    CDBG_ERR_INTERNAL = -6
} cdbg_status;

enum { CDBG_MEM_HOST = 0, CDBG_MEM_DEVICE = 1 };

enum {
    CDBG_FLAG_ALL_ABUNDANCE_COUNTS = 1, /* -all-abundance-counts: keep the per-k-mer abundance list (ab:Z:) */
    CDBG_FLAG_VERBOSE = 2
};

typedef struct {
    int32_t kmer_size;      /* -kmer-size, 4 <= k <= 127 (k <= 31: one 64-bit word, the tuned path; 32..127: 2..4 words, like the
                               reference's spans 64/96/128, G/tools/math/Integer.hpp:77-85) */
    int32_t abundance_min;  /* -abundance-min (solid <=> min <= count <= max) */
    int32_t abundance_max;  /* -abundance-max, 0 means INT32_MAX (the reference default) */
    int32_t minimizer_size; /* -minimizer-size, 0 means default 10 (clamped to k-1 like ConfigurationAlgorithm.cpp:251) */
    int32_t device;         /* CUDA device ordinal */
    int32_t flags;          /* CDBG_FLAG_* */
    int32_t rank;           /* shard index of this context (0 when world_size == 1) */
    int32_t world_size;     /* number of shards (GPUs); k <= 31: buckets owned by hash(minimizer) % world_size, k >= 32: by hash(k-mer) */
} cdbg_params;

typedef struct {
    uint64_t n_bytes;       /* bytes pushed */
    uint64_t n_instances;   /* valid k-mer occurrences (K1) */
    uint64_t n_distinct;    /* distinct canonical k-mers (K2) */
    uint64_t n_solid;       /* after the abundance filter */
    uint64_t n_junction_records; /* (k-1)-mer end records joined (K3-K5) */
    uint64_t n_compactions; /* junctions compacted (pairs of k-mer ends linked) */
    uint64_t n_cycles;      /* circular chains cut */
    uint64_t n_unitigs;
    uint64_t total_nt;      /* sum of unitig lengths */
    uint64_t n_links;       /* L: tokens */
    uint32_t n_count_bins;  /* partitions used by K1/K2 */
    uint32_t n_join_bins;   /* partitions used by K3-K5 */
    uint32_t rank_rounds;   /* pointer-jumping rounds (K6) */
    uint32_t kernel_launches; /* kernels launched by the library since cdbg_create / the last cdbg_reset */
    float ms_push, ms_count, ms_compact, ms_link, ms_fetch; /* device time per stage, CUDA events */
    float ms_k1_hist, ms_k1_scatter, ms_k2, ms_k3, ms_k5, ms_k6_rank, ms_k6_spell; /* inside the stages */
    float ms_k2_hist, ms_k2_scatter, ms_k2_table; /* inside K2 (shared-memory variant): fine histogram, fine scatter, table pass */
    uint32_t n_k2_splits;   /* K2 (shared-memory variant): table passes abandoned and redone as two halves */
    uint32_t k1_fallback;   /* 1 when the single-pass K1 overflowed a bin region and stage 1 was redone with the exact two-pass kernels */
    uint32_t zero_copy;     /* 1 when cdbg_build read pinned host memory directly from K1 (no separate H2D copy) */
    uint32_t n_k5_splits;   /* K5 (shared-memory variant): join passes abandoned and redone as two halves */
    uint32_t k2p_fallback;  /* 1 when the region-based fine partition of K2 (opt-in) overflowed and was redone with exact offsets */
    uint32_t sk_m;          /* minimizer length used to bucket the super-k-mers (stage 1 on super-k-mers, k <= 31; 0 otherwise) */
    uint64_t n_superkmers;  /* super-k-mer records written by K1 (16 bytes each) */
    float ms_x1, ms_x2, ms_x3, ms_glue, ms_xlink; /* sharded mode, device time: super-k-mer exchange (incl. packing), end-record exchange, link-message
                               routing, glue (piece exchange + piece-level ranking + assembly), link-record routing */
    uint32_t n_pieces;      /* sharded mode: shard-local unitigs before the glue */
    uint32_t sk_waves;      /* passes of stage 1 over the input (1 unless the buckets or their records do not fit in one pass) */
} cdbg_stats;

typedef struct {
    uint32_t from;      /* unitig id */
    uint32_t to;        /* unitig id */
    uint8_t from_sign;  /* '+' leaves the END of `from`, '-' leaves its BEGIN */
    uint8_t to_sign;    /* '+' enters `to` at its BEGIN, '-' at its END (LinkTigs.cpp:377,438) */
    uint8_t pad[2];
} cdbg_link_t;

/* Library-owned host buffers, valid until the next cdbg_reset / cdbg_destroy on the context. */
typedef struct {
    int32_t kmer_size;
    uint64_t n_unitigs;
    uint64_t total_nt;
    const char *seq;          /* concatenated ASCII sequences, unitig u = seq[seq_off[u] .. seq_off[u+1]) */
    const uint64_t *seq_off;  /* [n_unitigs + 1] */
    const uint64_t *kc;       /* [n_unitigs] sum of k-mer abundances (KC:i:) */
    const uint32_t *abund;    /* per-k-mer abundances in unitig order, unitig u starts at seq_off[u] - u*(k-1);
                                 NULL unless CDBG_FLAG_ALL_ABUNDANCE_COUNTS */
    uint64_t n_links;
    const cdbg_link_t *links; /* grouped by `from` (CSR); order inside a group is unspecified, as in the reference */
    const uint64_t *link_off; /* [n_unitigs + 1]: links of unitig u = links[link_off[u] .. link_off[u+1]) */
    /* sharded mode: this context returns the unitigs it owns, global ids first_unitig .. first_unitig + n_unitigs - 1
     * (local index u <-> global id first_unitig + u); `from`/`to` in links are GLOBAL ids. Single GPU: 0 and n_unitigs. */
    uint64_t first_unitig;
    uint64_t n_unitigs_total;
} cdbg_unitig_view;

int cdbg_create(cdbg_ctx **out, const cdbg_params *params);
void cdbg_destroy(cdbg_ctx *ctx);
/* Forget all pushed reads and results; keeps the device memory pool warm. */
int cdbg_reset(cdbg_ctx *ctx);
const char *cdbg_last_error(const cdbg_ctx *ctx);

/* Append reads.  `location` says where `bytes` lives; host buffers are copied H2D inside the call (pinned buffers copy at
 * full PCIe speed), device buffers are used in place when this is the only push (caller keeps them alive until
 * cdbg_count returns). */
int cdbg_push_reads(cdbg_ctx *ctx, const uint8_t *bytes, uint64_t n_bytes, int location);

/* Optional: make room on the device for `n_bytes` of reads before a series of pushes (a host that streams a large file in chunks
 * knows the total; without it the read buffer grows by doubling).  Reads already pushed are kept. */
int cdbg_reserve_reads(cdbg_ctx *ctx, uint64_t n_bytes);

/* Append reads given as raw FASTA or FASTQ TEXT (the bytes of the file, after gunzip): headers, '+' and quality lines are
 * removed and multi-line records joined ON THE GPU (a parallel scan over the parser's state-transition functions, parse.cuh),
 * replacing the host-side Bank parser (G/bank/impl/BankFasta.cpp:390-560) for this path.  The text must start at a record
 * ('>' or '@'); FASTQ must be in strict 4-line records, otherwise CDBG_ERR_PARAM is returned and nothing is appended (parse
 * such input on the host and use cdbg_push_reads).  *n_records (optional) receives the number of records. */
int cdbg_push_sequence_text(cdbg_ctx *ctx, const uint8_t *text, uint64_t n_bytes, int location, uint64_t *n_records);

/* The nucleotide stream pushed so far (after GPU parsing, with the break bytes between pushes), copied to `dst` (capacity `cap`);
 * *n receives its length.  For inspection and tests. */
int cdbg_get_pushed_reads(cdbg_ctx *ctx, uint8_t *dst, uint64_t cap, uint64_t *n);

int cdbg_count(cdbg_ctx *ctx, cdbg_stats *stats);   /* stage 1: K1 extract + K2 count + solidity filter */
int cdbg_compact(cdbg_ctx *ctx, cdbg_stats *stats); /* stages 2+3: K3-K5 junction join, K6/K8 chain + spell */
int cdbg_link(cdbg_ctx *ctx, cdbg_stats *stats);    /* stage 4: K9 links */
int cdbg_get_unitigs(cdbg_ctx *ctx, cdbg_unitig_view *view); /* D2H into library-owned pinned buffers */
int cdbg_get_stats(cdbg_ctx *ctx, cdbg_stats *stats);

/* reads -> unitigs in one call (reset, push, count, compact, link, get_unitigs). */
int cdbg_build(cdbg_ctx *ctx, const uint8_t *bytes, uint64_t n_bytes, int location, cdbg_unitig_view *view,
               cdbg_stats *stats);

/* Stage-1 result for parity checks in isolation (like the reference's `dsk/solid` partitions): copies the solid
 * (k-mer, abundance) pairs to host arrays of capacity `cap` k-mers; returns the number of solid k-mers in *n. Order is
 * unspecified.  A k-mer takes W = 1 (k <= 31) or max(2, ceil(2k/64)) (k >= 32) little-endian 64-bit words in `kmers`. */
int cdbg_get_solid_kmers(cdbg_ctx *ctx, uint64_t *kmers, uint32_t *counts, uint64_t cap, uint64_t *n);

/* The same hand-off in the other direction (start from counted k-mers, like the reference resuming from the solid
 * k-mers of its `.h5`, GraphUnitigs.cpp:884-951): install `n` solid (k-mer, abundance) pairs -- e.g. a DSK / KMC / Jellyfish dump
 * after the caller's abundance filter -- and continue with cdbg_compact / cdbg_link.  K-mers are 2 bits per nucleotide, first
 * nucleotide most significant, A=0 C=1 T=2 G=3, in either orientation (they are canonicalised on the device); every canonical
 * k-mer must occur once (W words per k-mer as for cdbg_get_solid_kmers).  Single GPU, on an empty context (after cdbg_create or
 * cdbg_reset). */
int cdbg_set_solid_kmers(cdbg_ctx *ctx, const uint64_t *kmers, const uint32_t *counts, uint64_t n, int location);

/* Device-side synthetic read generator used by bench.py (inputs resident in HBM when the timed region starts):
 * uniform random genome of `genome_len` nt from `seed`, `n_reads` reads of `read_len` nt, uniform start, random strand,
 * i.i.d. substitution errors at `err_ppm` parts per million, one '\n' after each read.  Writes n_reads*(read_len+1)
 * bytes to device memory owned by the context and returns the device pointer. */
int cdbg_synth_reads(cdbg_ctx *ctx, uint64_t genome_len, uint32_t read_len, uint64_t n_reads, uint32_t err_ppm,
                     uint64_t seed, uint8_t **d_bytes, uint64_t *n_bytes);

/* Sharded (multi-GPU) mode: one context per GPU with params.rank / params.world_size set.  Rank 0 obtains an id with
 * cdbg_nccl_unique_id and distributes its 128 bytes to the other ranks by any means (torch.distributed, MPI, a file);
 * every rank then calls cdbg_shard_init (collective).  From then on cdbg_push_reads takes THIS rank's share of the reads and
 * cdbg_count / cdbg_compact / cdbg_link / cdbg_build are collective calls: k-mers are routed to owner(minimizer) by an NCCL
 * all-to-all, each GPU compacts its shard, and a second round of exchanges glues the shard-local unitigs and their links. */
int cdbg_nccl_unique_id(uint8_t *out128);
int cdbg_shard_init(cdbg_ctx *ctx, const uint8_t *id128);

/* Same generator, reads first_read .. first_read + n_reads - 1 of the same read set: lets every rank of a sharded run
 * generate its own share. */
int cdbg_synth_reads_range(cdbg_ctx *ctx, uint64_t genome_len, uint32_t read_len, uint64_t first_read, uint64_t n_reads,
                           uint32_t err_ppm, uint64_t seed, uint8_t **d_bytes, uint64_t *n_bytes);

/* Host-side pieces of the sharded path, callable without a GPU (used by the 2-process gloo tests):
 *  cdbg_owner_of            GPU that owns the n_nt-mer x (a k-mer for counting, a (k-1)-mer for the junction join):
 *                           hash(min over its canonical m-mers of mmer_rank) mod world -- the same function the kernels use.
 *  cdbg_plan_shard_exchange the layout arithmetic of a destination-major all-to-all: from the all-gathered per-(destination, local
 *                           bin) counts of every rank (all_cnt[s * world*p_local + d * p_local + b]) to this rank's send / receive
 *                           counts and, per local bin b and source s, the run [run_off, run_off + run_len) of the receive buffer. */
uint32_t cdbg_owner_of(uint64_t x, int32_t n_nt, int32_t m, uint32_t world);
int cdbg_plan_shard_exchange(uint32_t world, uint32_t rank, uint32_t p_local, const uint64_t *all_cnt, uint64_t *send_cnt, uint64_t *recv_cnt,
                             uint64_t *run_off, uint64_t *run_len);

/* Pinned (page-locked) host memory for read buffers.  cdbg_push_reads copies from pinned memory at full PCIe speed, and
 * cdbg_build does not copy at all when `bytes` is pinned and 16-byte aligned (memory from here, cudaHostAlloc, or a pinned
 * torch tensor): the extraction kernel reads the buffer in place over PCIe (cdbg_stats.zero_copy says whether that happened).
 * Returns NULL on failure. */
void *cdbg_host_alloc(uint64_t n_bytes);
void cdbg_host_free(void *p);

/* Utility: copy device memory owned by the context to a host buffer (used by tests / bench to fetch synthetic reads). */
int cdbg_copy_to_host(cdbg_ctx *ctx, void *dst, const void *d_src, uint64_t n_bytes);

/* Library version string, e.g. "cdbg_b200 0.1 (sm_100a)". */
const char *cdbg_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CDBG_H_ */
