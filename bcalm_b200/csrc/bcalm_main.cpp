// bcalm_main.cpp -- `bcalm`-compatible command line host on top of the C ABI (include/cdbg.h).
//
// Drop-in for the reference CLI on the reads -> unitigs path:
//   bcalm -in <fasta|fastq[.gz]|comma list|file of files> -kmer-size <k> -abundance-min <a> [-out <prefix>] [-nb-cores n]
//         [-all-abundance-counts] [-edge-km 0|1] [-verbose v] [-minimizer-size m] [-abundance-max x]      (4 <= k <= 127)
// writes <prefix>.unitigs.fa with prefix = -out, else the basename of -in cut at its last dot, in the current directory
// (G/debruijn/impl/GraphUnitigs.cpp:888-904, G/system/impl/FileSystemCommon.cpp:115-137), in the reference's header grammar
//   ><id> LN:i:<len> KC:i:<sum> km:f:<mean, 1 decimal> [L:<+|->:<id>:<+|->]*      (G/bcalm2/bglue_algo.cpp:197-213, LinkTigs.cpp)
// `-version` / `-v` print a version banner (src/main.cpp:30-37); errors print "EXCEPTION: <msg>" and exit non-zero
// (src/main.cpp:44-48).  GATB knobs that only steer the reference's disk/memory plumbing are accepted and ignored.
// Extra (not in the reference): -out-gfa <file> also writes GFA1 like scripts/convertToGFA.py.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <string>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/cdbg.h"

namespace {

struct Options {
    std::string in, out, out_gfa;
    std::string solid_in;   // -solid-kmers-in: start from counted k-mers ("KMER count" lines: dsk2ascii / jellyfish dump -c / kmc dump)
    std::string solid_out;  // -solid-kmers-out: also write the solid k-mers of this run in the same format
    bool kff = false;       // -kff: also write the solid k-mers as <prefix>.kff (KFF 1.0, like the reference's -kff)
    int k = 31, amin = 2, amax = 0, m = 10, verbose = 1, cores = 0, device = 0;
    bool all_abund = false;
    bool parse_only = false; // -parse-only: parse the input, print record/byte counts and a checksum, no GPU (host-side tests)
    int nb_gpus = 1;        // -nb-gpus N: one context per GPU (threads of this process), the graph sharded by minimizer over them
    int gpu_parse = -1;     // -gpu-parse 1: ship the file bytes and strip headers / quality lines on the GPU (cdbg_push_sequence_text); 0: host parser;
                            // default (-1): on the GPU for a single-GPU build from reads, on the host otherwise
    bool edge_km = false;   // -edge-km 1: J:<0|1>:<id>:<0|1> tokens instead of L: (G/debruijn/impl/LinkTigs.cpp:380-384,440-444)
};

[[noreturn]] void die(const std::string &msg) {
    printf("EXCEPTION: %s\n", msg.c_str());
    exit(EXIT_FAILURE);
}

// options of the reference parser that take one argument / no argument and are irrelevant here
const char *IGNORED_1[] = {"-max-memory", "-max-disk", "-out-dir", "-out-tmp", "-out-compress", "-storage-type", "-histo2D", "-histo",
                           "-minimizer-type", "-repartition-type", "-bloom", "-debloom", "-debloom-impl", "-branching-nodes",
                           "-topology-stats", "-integer-precision", "-nb-glue-partitions", "-solidity-custom", "-solidity-kind",
                           "-histo-max"};
const char *IGNORED_0[] = {"-no-mphf", "-config-only", "-redo-bcalm", "-skip-bcalm", "-redo-bglue", "-skip-bglue", "-redo-links",
                           "-skip-links"};

void usage() {
    printf("[bcalm_1 options]\n"
           "       -in                   (1 arg) :    reads file (fasta/fastq, .gz, comma separated list or file of files)\n"
           "       -kmer-size            (1 arg) :    size of a kmer  [default '31']\n"
           "       -abundance-min        (1 arg) :    min abundance threshold for solid kmers  [default '2']\n"
           "       -abundance-max        (1 arg) :    max abundance threshold for solid kmers  [default '2147483647']\n"
           "       -minimizer-size       (1 arg) :    size of a minimizer  [default '10']\n"
           "       -out                  (1 arg) :    output prefix  [default: basename of -in]\n"
           "       -out-gfa              (1 arg) :    also write GFA1 to this file\n"
           "       -solid-kmers-in       (1 arg) :    start from counted k-mers instead of reads: text lines 'KMER count' (dsk2ascii,\n"
           "                                          jellyfish dump -c, kmc dump); -abundance-min/-max are applied; -in is then optional\n"
           "       -solid-kmers-out      (1 arg) :    also write the solid k-mers of this run as 'KMER count' lines\n"
           "       -kff                  (0 arg) :    also write the solid k-mers as <prefix>.kff (KFF 1.0); -solid-kmers-in reads KFF too\n"
           "       -all-abundance-counts (0 arg) :    output all k-mer abundance counts instead of mean\n"
           "       -edge-km              (1 arg) :    edge km representation  [default '0']\n"
           "       -nb-cores             (1 arg) :    number of cores (host parsing threads)  [default '0']\n"
           "       -device               (1 arg) :    CUDA device ordinal  [default '0']\n"
           "       -nb-gpus              (1 arg) :    number of GPUs (devices -device .. -device+N-1), one shard each, NCCL between them  [default '1']\n"
           "       -gpu-parse            (1 arg) :    parse FASTA/FASTQ text on the GPU instead of the host  [default '0']\n"
           "       -verbose              (1 arg) :    verbosity level  [default '1']\n"
           "       -version              (0 arg) :    version\n"
           "       -help                 (0 arg) :    help\n"
           "   GATB storage/memory knobs (-max-memory -max-disk -out-dir -out-tmp -storage-type ...) are accepted and ignored.\n");
}

Options parse(int argc, char **argv) {
    Options o;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto need = [&](const char *name) -> std::string {
            if (i + 1 >= argc) die(std::string("option ") + name + " needs an argument");
            return argv[++i];
        };
        if (a == "-in") o.in = need("-in");
        else if (a == "-out") o.out = need("-out");
        else if (a == "-out-gfa") o.out_gfa = need("-out-gfa");
        else if (a == "-solid-kmers-in") o.solid_in = need("-solid-kmers-in");
        else if (a == "-solid-kmers-out") o.solid_out = need("-solid-kmers-out");
        else if (a == "-kff") o.kff = true;
        else if (a == "-kmer-size") o.k = atoi(need("-kmer-size").c_str());
        else if (a == "-abundance-min") o.amin = atoi(need("-abundance-min").c_str());
        else if (a == "-abundance-max") { long long v = atoll(need("-abundance-max").c_str()); o.amax = v >= 2147483647LL ? 0 : (int)v; }
        else if (a == "-minimizer-size") o.m = atoi(need("-minimizer-size").c_str());
        else if (a == "-nb-cores") o.cores = atoi(need("-nb-cores").c_str());
        else if (a == "-verbose") o.verbose = atoi(need("-verbose").c_str());
        else if (a == "-device") o.device = atoi(need("-device").c_str());
        else if (a == "-nb-gpus") o.nb_gpus = atoi(need("-nb-gpus").c_str());
        else if (a == "-all-abundance-counts") o.all_abund = true;
        else if (a == "-parse-only") o.parse_only = true;
        else if (a == "-gpu-parse") o.gpu_parse = atoi(need("-gpu-parse").c_str()) != 0 ? 1 : 0;
        else if (a == "-edge-km") o.edge_km = atoi(need("-edge-km").c_str()) != 0;
        else if (a == "-help" || a == "-h") { usage(); exit(EXIT_SUCCESS); }
        else {
            bool known = false;
            for (const char *s : IGNORED_1) if (a == s) { need(s); known = true; break; }
            if (!known) for (const char *s : IGNORED_0) if (a == s) { known = true; break; }
            if (!known) { usage(); die("Unknown parameter '" + a + "'"); }
        }
    }
    if (o.in.empty() && o.solid_in.empty()) { usage(); die("Option '-in' is mandatory"); }
    if (o.nb_gpus < 1 || o.nb_gpus > 16) die("-nb-gpus must be between 1 and 16 (the library shards one graph over at most 16 GPUs of one node)");
    if (o.nb_gpus > 1 && (!o.solid_in.empty() || !o.solid_out.empty() || o.gpu_parse == 1 || o.kff))
        die("-nb-gpus > 1 does not support -solid-kmers-in / -solid-kmers-out / -gpu-parse / -kff");
    return o;
}

// FileSystemCommon::getBaseName: posix basename, then cut at the LAST dot
std::string base_name(const std::string &path) {
    size_t s = path.find_last_of('/');
    std::string b = s == std::string::npos ? path : path.substr(s + 1);
    size_t d = b.find_last_of('.');
    if (d != std::string::npos) b = b.substr(0, d);
    return b;
}

std::string dir_name(const std::string &path) {
    size_t s = path.find_last_of('/');
    return s == std::string::npos ? std::string(".") : path.substr(0, s);
}

// A file's bytes without zero-filling or re-growing: plain files are read in one go into an uninitialised buffer sized from
// fstat, gzip files (magic 1f 8b) go through zlib with a doubling buffer.
struct FileBuf {
    char *p = nullptr; size_t n = 0;
    ~FileBuf() { free(p); }
    FileBuf() = default;
    FileBuf(const FileBuf &) = delete;
    FileBuf &operator=(const FileBuf &) = delete;
};

bool read_whole(const std::string &path, FileBuf &buf) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    unsigned char magic[2] = {0, 0};
    size_t got = fread(magic, 1, 2, f);
    const bool gz = got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    if (!gz) {
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        if (sz < 0) { fclose(f); return false; }
        buf.p = (char *)malloc((size_t)sz + 1);
        if (!buf.p) { fclose(f); return false; }
        buf.n = fread(buf.p, 1, (size_t)sz, f);
        fclose(f);
        return true;
    }
    fclose(f);
    gzFile g = gzopen(path.c_str(), "rb");
    if (!g) return false;
    gzbuffer(g, 1 << 20);
    size_t cap = 1 << 24;
    buf.p = (char *)malloc(cap);
    for (;;) {
        if (buf.n == cap) { cap *= 2; buf.p = (char *)realloc(buf.p, cap); }
        if (!buf.p) { gzclose(g); return false; }
        int r = gzread(g, buf.p + buf.n, (unsigned)std::min<size_t>(cap - buf.n, 1u << 30));
        if (r < 0) { gzclose(g); return false; }
        if (r == 0) break;
        buf.n += (size_t)r;
    }
    gzclose(g);
    return true;
}

// growable byte buffer that does not zero-fill (std::vector<uint8_t>::resize would touch every page twice)
struct ByteBuf {
    uint8_t *p = nullptr; size_t n = 0, cap = 0;
    ~ByteBuf() { free(p); }
    ByteBuf() = default;
    ByteBuf(const ByteBuf &) = delete;
    ByteBuf &operator=(const ByteBuf &) = delete;
    void reserve(size_t want) { if (want > cap) { cap = std::max(want, cap * 2); p = (uint8_t *)realloc(p, cap); if (!p) { printf("EXCEPTION: out of host memory\n"); exit(EXIT_FAILURE); } } }
    void append(const uint8_t *src, size_t len) { reserve(n + len); memcpy(p + n, src, len); n += len; }
    void push(uint8_t c) { reserve(n + 1); p[n++] = c; }
    uint8_t *data() { return p; }
    size_t size() const { return n; }
};

// Append the nucleotide stream of one FASTA/FASTQ buffer to `out`: sequence characters only, one '\n' after each record
// (multi-line FASTA records are joined: BankFasta.cpp:390-560).
// Writes the nucleotide stream of [begin, e) to dst (which must have room for e - begin + 1 bytes: the output is never longer
// than the input plus one break) and returns its length.
size_t parse_sequences_raw(const char *begin, const char *e, uint8_t *dst, uint64_t &n_records) {
    const char *p = begin;
    uint8_t *o = dst;
    while (p < e) {
        if (*p == '>') {                                  // FASTA record
            const char *nl = (const char *)memchr(p, '\n', e - p);
            p = nl ? nl + 1 : e;
            while (p < e && *p != '>') {
                const char *l = (const char *)memchr(p, '\n', e - p);
                const char *le = l ? l : e;
                size_t len = le - p;
                if (len && p[len - 1] == '\r') len--;
                memcpy(o, p, len); o += len;
                p = l ? l + 1 : e;
            }
            *o++ = '\n';
            n_records++;
        } else if (*p == '@') {                           // FASTQ record (4 lines; multi-line FASTQ is not produced by sequencers)
            const char *l1 = (const char *)memchr(p, '\n', e - p);
            if (!l1) break;
            const char *s = l1 + 1;
            const char *l2 = (const char *)memchr(s, '\n', e - s);
            const char *se = l2 ? l2 : e;
            size_t len = se - s;
            if (len && s[len - 1] == '\r') len--;
            memcpy(o, s, len); o += len;
            *o++ = '\n';
            n_records++;
            if (!l2) break;
            const char *l3 = (const char *)memchr(l2 + 1, '\n', e - (l2 + 1));   // '+' line
            if (!l3) break;
            const char *l4 = (const char *)memchr(l3 + 1, '\n', e - (l3 + 1));   // quality line
            p = l4 ? l4 + 1 : e;
        } else {
            const char *nl = (const char *)memchr(p, '\n', e - p);   // blank / stray line
            p = nl ? nl + 1 : e;
        }
    }
    return (size_t)(o - dst);
}

void parse_sequences(const char *begin, const char *e, ByteBuf &out, uint64_t &n_records) {
    out.reserve(out.size() + (size_t)(e - begin) + 16);
    out.n += parse_sequences_raw(begin, e, out.p + out.n, n_records);
}

// Parse one file's bytes on `threads` host threads: the buffer is cut at record starts ('>' or, for FASTQ, every 4th line
// would need a global line count, so FASTQ stays single-threaded unless it is cut at "\n@" followed by a plausible 4-line record),
// each piece is parsed independently and the pieces are concatenated in order -- byte-identical to the serial parse.
void parse_sequences_mt(const FileBuf &in, ByteBuf &out, uint64_t &n_records, int threads) {
    const size_t n = in.n;
    const char *b = in.p;
    if (threads <= 1 || n < (8u << 20) || b[0] != '>') { parse_sequences(b, b + n, out, n_records); return; }
    std::vector<size_t> cut{0};
    for (int t = 1; t < threads; t++) {
        size_t p = n / threads * t;
        while (p < n && !(b[p] == '>' && b[p - 1] == '\n')) p++;      // next FASTA record start
        if (p < n && p > cut.back()) cut.push_back(p);
    }
    cut.push_back(n);
    const size_t pieces = cut.size() - 1;
    // every piece writes at its own INPUT offset inside the output reservation (output <= input, so pieces cannot collide),
    // then the pieces are closed up in order: no per-thread buffers, one pass of page faults shared by all threads
    out.reserve(out.size() + n + pieces + 16);
    uint8_t *base = out.p + out.n;
    std::vector<size_t> len(pieces, 0);
    std::vector<uint64_t> recs(pieces, 0);
    std::vector<std::thread> pool;
    for (size_t i = 0; i < pieces; i++) pool.emplace_back([&, i]() { len[i] = parse_sequences_raw(b + cut[i], b + cut[i + 1], base + cut[i] + i, recs[i]); });
    for (auto &th : pool) th.join();
    size_t at = 0;
    for (size_t i = 0; i < pieces; i++) {
        if (at != cut[i] + i) memmove(base + at, base + cut[i] + i, len[i]);
        at += len[i]; n_records += recs[i];
    }
    out.n += at;
}

static int g_parse_threads = 1;
static cdbg_ctx *g_gpu_parse_ctx = nullptr;   // -gpu-parse 1: files are pushed as text and parsed on the device
static std::function<cdbg_ctx *()> g_gpu_parse_wait;   // the context is created on its own thread while the first file is read: join it here

// A large plain-text FASTA file for the GPU parser: the file is mapped, cut at record starts into chunks of ~256 MB
// (CDBG_CLI_CHUNK), and a reader thread copies chunk i + 1 from the page cache into one of two pinned buffers (all host threads
// share the copy) while the GPU receives and parses chunk i -- instead of read() into a pageable buffer followed by one big
// pageable H2D copy.  Returns false when nothing was pushed (the caller then reads the file the ordinary way).
bool push_fasta_chunked(cdbg_ctx *ctx, const std::string &path, uint64_t &n_records) {
    // (below ~0.5 GB the two pinned buffers cost more to allocate than the pipeline saves: C2's 156 MB file took 0.95 s this way, 0.4 s the ordinary way)
    const size_t min_size = getenv("CDBG_CLI_CHUNK_MIN") ? strtoull(getenv("CDBG_CLI_CHUNK_MIN"), nullptr, 10) : (512u << 20);
    const size_t CH = std::max<size_t>(getenv("CDBG_CLI_CHUNK") ? strtoull(getenv("CDBG_CLI_CHUNK"), nullptr, 10) : (256u << 20), 64);
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < min_size) { close(fd); return false; }
    const size_t n = (size_t)sb.st_size;
    const char *map = (const char *)mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return false;
    auto give_up = [&]() { munmap((void *)map, n); return false; };
    if (map[0] != '>') return give_up();                           // gz, FASTQ, file of files, leading blank lines: the ordinary path
    madvise((void *)map, n, MADV_SEQUENTIAL);
    std::vector<size_t> cut{0};                                      // chunk i = [cut[i], cut[i + 1]), every cut at a record start
    while (cut.back() < n) {
        size_t e = std::min(n, cut.back() + CH);
        if (e < n) {
            size_t p = e;
            while (p > cut.back() + 1 && !(map[p] == '>' && map[p - 1] == '\n')) p--;
            if (p <= cut.back() + 1) { p = e; while (p < n && !(map[p] == '>' && map[p - 1] == '\n')) p++; }   // a record longer than a chunk
            e = p;
        }
        cut.push_back(e);
    }
    const size_t n_chunks = cut.size() - 1;
    size_t maxc = 0;
    for (size_t i = 0; i < n_chunks; i++) maxc = std::max(maxc, cut[i + 1] - cut[i]);
    uint8_t *buf[2] = {(uint8_t *)cdbg_host_alloc(maxc), (uint8_t *)cdbg_host_alloc(maxc)};
    if (!buf[0] || !buf[1]) { cdbg_host_free(buf[0]); cdbg_host_free(buf[1]); return give_up(); }
    if (cdbg_reserve_reads(ctx, n + n_chunks + 64) != CDBG_OK) { cdbg_host_free(buf[0]); cdbg_host_free(buf[1]); return give_up(); }
    std::mutex mu; std::condition_variable cv;
    int filled[2] = {-1, -1};                                         // chunk held by each buffer, -1 = free
    bool stop = false;
    const int copy_threads = std::max(1, std::min(g_parse_threads, 16));
    std::thread reader([&]() {
        for (size_t i = 0; i < n_chunks; i++) {
            const int b = (int)(i & 1);
            { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&]() { return filled[b] == -1 || stop; }); if (stop) return; }
            const size_t len = cut[i + 1] - cut[i], per = (len + copy_threads - 1) / copy_threads;
            std::vector<std::thread> pool;
            for (int t = 1; t < copy_threads; t++) {
                const size_t o0 = std::min(len, per * t), o1 = std::min(len, o0 + per);
                if (o1 > o0) pool.emplace_back([=]() { memcpy(buf[b] + o0, map + cut[i] + o0, o1 - o0); });
            }
            memcpy(buf[b], map + cut[i], std::min(len, per));
            for (auto &th : pool) th.join();
            { std::lock_guard<std::mutex> lk(mu); filled[b] = (int)i; }
            cv.notify_all();
        }
    });
    bool ok = true, pushed_any = false;
    std::string err;
    for (size_t i = 0; i < n_chunks && ok; i++) {
        const int b = (int)(i & 1);
        { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&]() { return filled[b] == (int)i; }); }
        uint64_t nrec = 0;
        if (cdbg_push_sequence_text(ctx, buf[b], cut[i + 1] - cut[i], CDBG_MEM_HOST, &nrec) != CDBG_OK) { ok = false; err = cdbg_last_error(ctx); }
        else { n_records += nrec; pushed_any = true; }
        { std::lock_guard<std::mutex> lk(mu); filled[b] = -1; if (!ok) stop = true; }
        cv.notify_all();
    }
    reader.join();
    cdbg_host_free(buf[0]); cdbg_host_free(buf[1]);
    munmap((void *)map, n);
    if (!ok && pushed_any) die("GPU parser failed in the middle of '" + path + "': " + err);
    return ok;
}

void load_input(const std::string &spec, ByteBuf &stream, uint64_t &n_records, int depth = 0) {
    size_t start = 0;
    while (start <= spec.size()) {
        size_t c = spec.find(',', start);
        std::string path = spec.substr(start, c == std::string::npos ? std::string::npos : c - start);
        start = c == std::string::npos ? spec.size() + 1 : c + 1;
        if (path.empty()) continue;
        if (g_gpu_parse_wait || g_gpu_parse_ctx) {                  // large plain FASTA: streamed to the GPU parser in chunks
            if (!g_gpu_parse_ctx) g_gpu_parse_ctx = g_gpu_parse_wait();
            if (push_fasta_chunked(g_gpu_parse_ctx, path, n_records)) continue;
        }
        FileBuf buf;
        if (!read_whole(path, buf)) die("Unable to open file '" + path + "'");
        size_t i = 0;
        while (i < buf.n && (buf.p[i] == '\n' || buf.p[i] == '\r' || buf.p[i] == ' ')) i++;
        if (i == buf.n) continue;
        if (buf.p[i] == '>' || buf.p[i] == '@') {
            bool done = false;
            if (!g_gpu_parse_ctx && g_gpu_parse_wait) g_gpu_parse_ctx = g_gpu_parse_wait();
            if (g_gpu_parse_ctx) {   // the library refuses text it cannot parse exactly (e.g. multi-line FASTQ): then the host parser takes it
                uint64_t nrec = 0;
                done = cdbg_push_sequence_text(g_gpu_parse_ctx, (const uint8_t *)buf.p + i, buf.n - i, CDBG_MEM_HOST, &nrec) == CDBG_OK;
                if (done) n_records += nrec;
            }
            if (!done) parse_sequences_mt(buf, stream, n_records, g_parse_threads);
        } else {
            if (depth > 2) die("file of files nested too deeply: " + path);
            // file of files: one path per line, relative to the list's directory (G/bank/impl/BankAlbum.cpp)
            std::string text(buf.p, buf.p + buf.n), dir = dir_name(path);
            size_t ls = 0;
            while (ls < text.size()) {
                size_t le = text.find('\n', ls);
                std::string line = text.substr(ls, le == std::string::npos ? std::string::npos : le - ls);
                ls = le == std::string::npos ? text.size() : le + 1;
                while (!line.empty() && (line.back() == '\r' || line.back() == ' ')) line.pop_back();
                if (line.empty()) continue;
                std::string sub = line[0] == '/' ? line : dir + "/" + line;
                load_input(sub, stream, n_records, depth + 1);
            }
        }
    }
}

inline void put_u64(std::string &s, uint64_t v) {
    char tmp[24]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) s.push_back(tmp[--n]);
}

// km:f: -- float32 mean printed with one decimal (bglue_algo.cpp:131-145, :75-81)
inline void put_km(std::string &s, uint64_t kc, uint64_t n_kmers, const uint32_t *ab) {
    float sum = 0.f;
    if (ab && kc >= (1ULL << 24)) for (uint64_t i = 0; i < n_kmers; i++) sum += (float)(int)ab[i];
    else sum = (float)kc;
    float mean = sum / (float)n_kmers;
    char tmp[48];
    int n = snprintf(tmp, sizeof tmp, "%.1f", (double)mean);
    s.append(tmp, n);
}

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

// "KMER count" lines -> packed canonical k-mers (A=0 C=1 T=2 G=3, first nucleotide most significant) + abundances.  A dump made
// without canonical counting lists both orientations of a k-mer: their counts are added.  The abundance filter is applied here.
// ---- KFF (k-mer file format) 1.0, the format of the reference's `-kff` dump (G/kmer/impl/CountProcessorDumpKff.hpp through the
// vendored kff-cpp-api).  Layout, all integers big-endian:
//   "KFF" major minor encoding uniqueness canonicity free_size(4) free_block
//   sections:  'v' n(8) { name\0 value(8) }*          global variables: k, max, data_size
//              'r' n_blocks(8) { [n_kmers (ceil(ceil(log2 max)/8) bytes, absent when max = 1)]
//                                sequence of n_kmers+k-1 nucleotides, 2 bits each, left-padded to whole bytes
//                                n_kmers * data_size bytes of data (here: the abundance) }*
//              'i' n(8) { type(1) relative_position(8) }* next_index(8)      index of the sections before it
//   footer:    'v' with first_index and footer_size, then "KFF"
// encoding byte = the 2-bit codes of A, C, G, T from the top; the reference writes 0x1e (A=0 C=1 G=3 T=2), the same as this library.
// 'm' (minimizer) sections are not read.
uint64_t kff_be(const uint8_t *p, int n) { uint64_t v = 0; for (int i = 0; i < n; i++) v = (v << 8) | p[i]; return v; }

// A k-mer of up to 127 nucleotides on the host: 4 little-endian 64-bit words (the library's layout, W = words_of(k) of them are used),
// compared most significant word first like the device code.
struct HostKmer {
    uint64_t w[4];
    bool operator<(const HostKmer &o) const { for (int i = 3; i >= 0; i--) if (w[i] != o.w[i]) return w[i] < o.w[i]; return false; }
    bool operator==(const HostKmer &o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3]; }
};
inline int words_of(int k) { return k <= 31 ? 1 : std::max(2, (2 * k + 63) / 64); }
// roll one nucleotide (our 2-bit code) into the forward k-mer x and its reverse complement r
inline void kmer_roll(HostKmer &x, HostKmer &r, uint64_t code, int k) {
    for (int i = 3; i > 0; i--) x.w[i] = (x.w[i] << 2) | (x.w[i - 1] >> 62);
    x.w[0] = (x.w[0] << 2) | code;
    const int bits = 2 * k;                                  // keep the low 2k bits
    for (int i = 0; i < 4; i++) { const int lo = 64 * i; if (bits <= lo) x.w[i] = 0; else if (bits < lo + 64) x.w[i] &= (1ULL << (bits - lo)) - 1ULL; }
    for (int i = 0; i < 3; i++) r.w[i] = (r.w[i] >> 2) | (r.w[i + 1] << 62);
    r.w[3] >>= 2;
    const int top = 2 * (k - 1);
    r.w[top >> 6] |= (code ^ 2) << (top & 63);
}
inline unsigned kmer_nt(const uint64_t *km, int j) { return (unsigned)(km[(2 * j) >> 6] >> ((2 * j) & 63)) & 3u; }   // nucleotide j from the last

bool load_kff(const Options &o, const FileBuf &in, std::vector<std::pair<HostKmer, uint64_t>> &rec) {
    const uint8_t *b = (const uint8_t *)in.p, *e = b + in.n;
    if (in.n < 12 || memcmp(b, "KFF", 3) != 0) return false;
    if (b[3] != 1) die("unsupported KFF version in '" + o.solid_in + "'");
    const uint8_t enc = b[5];
    uint8_t to_ours[4];                                // file code -> our code (A=0 C=1 T=2 G=3)
    { const uint8_t a = enc >> 6, c = (enc >> 4) & 3, g = (enc >> 2) & 3, t = enc & 3; to_ours[a] = 0; to_ours[c] = 1; to_ours[t] = 2; to_ours[g] = 3; }
    const uint8_t *p = b + 12 + kff_be(b + 8, 4);
    uint64_t kk = 0, maxk = 0, dsz = 0;
    bool have_k = false;
    auto need = [&](size_t n) { if ((size_t)(e - p) < n) die("truncated KFF file '" + o.solid_in + "'"); };
    while (p < e) {
        if (e - p == 3 && memcmp(p, "KFF", 3) == 0) break;           // end signature
        need(9);
        const char type = (char)*p++;
        const uint64_t n = kff_be(p, 8); p += 8;
        if (type == 'v') {
            for (uint64_t i = 0; i < n; i++) {
                const uint8_t *z = (const uint8_t *)memchr(p, 0, (size_t)(e - p));
                if (!z) die("bad variable section in KFF file '" + o.solid_in + "'");
                std::string name((const char *)p, (size_t)(z - p));
                p = z + 1; need(8);
                const uint64_t v = kff_be(p, 8); p += 8;
                if (name == "k") { kk = v; have_k = true; } else if (name == "max") maxk = v; else if (name == "data_size") dsz = v;
            }
        } else if (type == 'i') {
            need(n * 9 + 8); p += n * 9 + 8;
        } else if (type == 'r') {
            if (!have_k || maxk == 0) die("KFF raw section before the k / max variables in '" + o.solid_in + "'");
            if ((int)kk != o.k) die("k of the KFF file (" + std::to_string(kk) + ") differs from -kmer-size");
            if (dsz > 8) die("KFF data_size > 8 is not supported");
            int nbits = 0; while ((1ULL << nbits) < maxk) nbits++;      // ceil(log2(max))
            const int nb_bytes = (nbits + 7) / 8;
            for (uint64_t blk = 0; blk < n; blk++) {
                uint64_t nk = 1;
                if (nb_bytes) { need((size_t)nb_bytes); nk = kff_be(p, nb_bytes); p += nb_bytes; }
                const uint64_t nt = nk + kk - 1, sb = (nt + 3) / 4;
                need(sb + nk * dsz);
                const uint8_t *seq = p, *dat = p + sb;
                const uint64_t pad = sb * 4 - nt;                          // leading padding, in nucleotides
                HostKmer x{}, r{};
                for (uint64_t i = 0; i < nt; i++) {
                    const uint64_t q = pad + i;
                    const uint64_t code = to_ours[(seq[q >> 2] >> (2 * (3 - (q & 3)))) & 3];
                    kmer_roll(x, r, code, o.k);
                    if (i + 1 >= kk) {
                        const uint64_t j = i + 1 - kk;
                        rec.emplace_back(x < r ? x : r, dsz ? kff_be(dat + j * dsz, (int)dsz) : 1);
                    }
                }
                p += sb + nk * dsz;
            }
        } else if (type == 'm') {
            die("KFF minimizer sections are not supported: convert the file to raw sections first");
        } else die("unknown KFF section type in '" + o.solid_in + "'");
    }
    return true;
}

void write_kff(const std::string &path, cdbg_ctx *ctx, int k) {
    uint64_t n = 0;
    if (cdbg_get_solid_kmers(ctx, nullptr, nullptr, 0, &n) != CDBG_OK) die(cdbg_last_error(ctx));
    const int W = words_of(k);
    std::vector<uint64_t> km(n * W + 1); std::vector<uint32_t> ct(n + 1);
    if (cdbg_get_solid_kmers(ctx, km.data(), ct.data(), n, &n) != CDBG_OK) die(cdbg_last_error(ctx));
    std::string f;
    auto be = [&](uint64_t v, int bytes) { for (int i = bytes - 1; i >= 0; i--) f.push_back((char)(v >> (8 * i))); };
    auto var = [&](const char *name, uint64_t v) { f += name; f.push_back('\0'); be(v, 8); };
    const char *note = "b200-cdbg";
    f += "KFF"; f.push_back(1); f.push_back(0); f.push_back((char)0x1e); f.push_back(1); f.push_back(1);
    be(strlen(note), 4); f += note;
    const uint64_t pos_v = f.size();
    f.push_back('v'); be(3, 8); var("data_size", 4); var("k", (uint64_t)k); var("max", 1);
    const uint64_t pos_r = f.size();
    f.push_back('r'); be(n, 8);
    const int sb = (k + 3) / 4;
    f.reserve(f.size() + n * (size_t)(sb + 4) + 128);
    for (uint64_t i = 0; i < n; i++) {
        for (int b = sb - 1; b >= 0; b--) f.push_back((char)(km[i * W + ((8 * b) >> 6)] >> ((8 * b) & 63)));     // 2k bits, big-endian, left-padded
        be(ct[i], 4);
    }
    const uint64_t pos_i = f.size(), end_i = pos_i + 17 + 9 * 2;
    f.push_back('i'); be(2, 8);
    f.push_back('v'); be((uint64_t)((int64_t)pos_v - (int64_t)end_i), 8);
    f.push_back('r'); be((uint64_t)((int64_t)pos_r - (int64_t)end_i), 8);
    be(0, 8);
    f.push_back('v'); be(2, 8); var("first_index", pos_i); var("footer_size", 9 + 2 * (12 + 8));
    f += "KFF";
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) die("Unable to open output file '" + path + "'");
    fwrite(f.data(), 1, f.size(), fp);
    fclose(fp);
}

void load_counted_kmers(const Options &o, std::vector<uint64_t> &kmers, std::vector<uint32_t> &counts) {
    FileBuf in;
    if (!read_whole(o.solid_in, in)) die("Unable to open file '" + o.solid_in + "'");
    const int k = o.k;
    std::vector<std::pair<HostKmer, uint64_t>> rec;
    const int W = words_of(k);
    const char *p = in.p, *e = in.p + in.n;
    uint64_t line = 0;
    if (load_kff(o, in, rec)) p = e;                   // a KFF file: the records are in `rec` already
    while (p < e) {
        line++;
        const char *nl = (const char *)memchr(p, '\n', (size_t)(e - p));
        const char *le = nl ? nl : e;
        const char *q = p;
        while (q < le && (*q == ' ' || *q == '\t')) q++;
        if (q < le && *q != '#' && *q != '\r') {
            HostKmer x{}, r{};
            int len = 0;
            for (; q < le && *q != ' ' && *q != '\t' && *q != '\r'; q++, len++) {
                const char c = *q & 0xDF;
                if (len >= k) die("k-mer longer than -kmer-size in '" + o.solid_in + "' line " + std::to_string(line));
                if (c != 'A' && c != 'C' && c != 'G' && c != 'T') die("bad nucleotide in '" + o.solid_in + "' line " + std::to_string(line));
                kmer_roll(x, r, ((uint64_t)c >> 1) & 3, k);
            }
            if (len != k) die("k-mer length differs from -kmer-size in '" + o.solid_in + "' line " + std::to_string(line));
            while (q < le && (*q == ' ' || *q == '\t')) q++;
            if (q >= le || *q < '0' || *q > '9') die("missing count in '" + o.solid_in + "' line " + std::to_string(line));
            uint64_t c = 0;
            for (; q < le && *q >= '0' && *q <= '9'; q++) c = std::min<uint64_t>(c * 10 + (uint64_t)(*q - '0'), 1ULL << 40);
            rec.emplace_back(x < r ? x : r, c);
        }
        p = nl ? nl + 1 : e;
    }
    std::sort(rec.begin(), rec.end());
    const uint64_t amax = o.amax > 0 ? (uint64_t)o.amax : 2147483647ULL, amin = (uint64_t)std::max(o.amin, 1);
    for (size_t i = 0; i < rec.size();) {
        size_t j = i; uint64_t c = 0;
        for (; j < rec.size() && rec[j].first == rec[i].first; j++) c += rec[j].second;
        if (c >= amin && c <= amax) { for (int w = 0; w < W; w++) kmers.push_back(rec[i].first.w[w]); counts.push_back((uint32_t)c); }
        i = j;
    }
}

void write_counted_kmers(const std::string &path, cdbg_ctx *ctx, int k) {
    uint64_t n = 0;
    if (cdbg_get_solid_kmers(ctx, nullptr, nullptr, 0, &n) != CDBG_OK) die(cdbg_last_error(ctx));
    const int W = words_of(k);
    std::vector<uint64_t> km(n * W + 1); std::vector<uint32_t> ct(n + 1);
    if (cdbg_get_solid_kmers(ctx, km.data(), ct.data(), n, &n) != CDBG_OK) die(cdbg_last_error(ctx));
    std::string text;
    text.reserve(n * (size_t)(k + 12));
    const char NT[4] = {'A', 'C', 'T', 'G'};
    for (uint64_t i = 0; i < n; i++) {
        for (int j = k - 1; j >= 0; j--) text.push_back(NT[kmer_nt(&km[i * W], j)]);
        text.push_back(' '); put_u64(text, ct[i]); text.push_back('\n');
    }
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) die("Unable to open output file '" + path + "'");
    fwrite(text.data(), 1, text.size(), f);
    fclose(f);
}

// <prefix>.unitigs.fa (and GFA1) from one view, or from the views of a sharded run in the order of their global unitig ids
void write_outputs(const Options &o, const std::vector<const cdbg_unitig_view *> &views, const std::string &out_path, cdbg_ctx *ctx_for_cleanup) {
    const int k = o.k;
    uint64_t total_nt = 0, n_unitigs = 0, n_links = 0;
    for (const cdbg_unitig_view *vp : views) { total_nt += vp->total_nt; n_unitigs += vp->n_unitigs; n_links += vp->n_links; }
    // the records of unitigs [u0, u1) of one view, appended to `text`
    auto format_range = [&](const cdbg_unitig_view &v, uint64_t u0, uint64_t u1, std::string &text) {
        for (uint64_t u = u0; u < u1; u++) {
            const uint64_t b = v.seq_off[u], len = v.seq_off[u + 1] - b, nk = len - (uint64_t)(k - 1);
            const uint32_t *ab = v.abund ? v.abund + (b - u * (uint64_t)(k - 1)) : nullptr;
            text.push_back('>'); put_u64(text, v.first_unitig + u);
            text += " LN:i:"; put_u64(text, len);
            if (o.all_abund && ab) {
                text += " ab:Z:";
                for (uint64_t i = 0; i < nk; i++) { put_u64(text, ab[i]); text.push_back(' '); }
            } else {
                text += " KC:i:"; put_u64(text, v.kc[u]);
                text += " km:f:"; put_km(text, v.kc[u], nk, ab);
                text.push_back(' ');
            }
            for (uint64_t i = v.link_off[u]; i < v.link_off[u + 1]; i++) {
                const cdbg_link_t &l = v.links[i];
                if (o.edge_km) {   // J:0 = in-link (leaves the BEGIN), J:1 = out-link; last field 1 when the neighbour is entered at its END
                    text += " J:"; text.push_back(l.from_sign == '-' ? '0' : '1'); text.push_back(':'); put_u64(text, l.to);
                    text.push_back(':'); text.push_back(l.to_sign == '-' ? '1' : '0');
                } else {
                    text += " L:"; text.push_back((char)l.from_sign); text.push_back(':'); put_u64(text, l.to);
                    text.push_back(':'); text.push_back((char)l.to_sign);
                }
            }
            text.push_back('\n');
            text.append(v.seq + b, len);
            text.push_back('\n');
        }
    };
    // formatted by all host threads (every piece a contiguous range of unitigs of about the same number of nucleotides), written in order
    struct Piece { const cdbg_unitig_view *v; uint64_t u0, u1; std::string text; };
    std::vector<Piece> pieces;
    {
        const uint64_t n_thr = (uint64_t)std::max(1, g_parse_threads);
        const uint64_t target = std::max<uint64_t>((total_nt + n_unitigs * 48) / (n_thr * 4) + 1, 1u << 20);
        for (const cdbg_unitig_view *vp : views) {
            uint64_t u0 = 0;
            while (u0 < vp->n_unitigs) {
                uint64_t u1 = u0, acc = 0;
                while (u1 < vp->n_unitigs && acc < target) { acc += vp->seq_off[u1 + 1] - vp->seq_off[u1] + 48; u1++; }
                pieces.push_back(Piece{vp, u0, u1, std::string()});
                u0 = u1;
            }
        }
        std::atomic<size_t> next{0};
        auto work = [&]() {
            for (size_t i; (i = next.fetch_add(1)) < pieces.size();) {
                Piece &pc = pieces[i];
                pc.text.reserve((size_t)(pc.v->seq_off[pc.u1] - pc.v->seq_off[pc.u0]) + (size_t)(pc.u1 - pc.u0) * 72 +
                                (size_t)(pc.v->link_off[pc.u1] - pc.v->link_off[pc.u0]) * 16 + 64);
                format_range(*pc.v, pc.u0, pc.u1, pc.text);
            }
        };
        std::vector<std::thread> pool;
        const size_t nt_ = std::min<size_t>((size_t)n_thr, pieces.size());
        for (size_t t = 1; t < nt_; t++) pool.emplace_back(work);
        work();
        for (auto &th : pool) th.join();
    }
    FILE *f = fopen(out_path.c_str(), "wb");
    if (!f) { if (ctx_for_cleanup) cdbg_destroy(ctx_for_cleanup); die("Unable to open output file '" + out_path + "'"); }
    size_t text_bytes = 0;
    for (const Piece &pc : pieces) { fwrite(pc.text.data(), 1, pc.text.size(), f); text_bytes += pc.text.size(); }
    fclose(f);

    if (!o.out_gfa.empty()) {   // GFA1, same records as scripts/convertToGFA.py:74,97-114
        std::string g;
        g.reserve(text_bytes + n_links * 24);
        g += "H\tVN:Z:1.0\tks:i:"; put_u64(g, (uint64_t)k); g.push_back('\n');
        for (const cdbg_unitig_view *vp : views) {
            const cdbg_unitig_view &v = *vp;
            for (uint64_t u = 0; u < v.n_unitigs; u++) {
                const uint64_t b = v.seq_off[u], len = v.seq_off[u + 1] - b, nk = len - (uint64_t)(k - 1);
                g += "S\t"; put_u64(g, v.first_unitig + u); g.push_back('\t'); g.append(v.seq + b, len);
                g += "\tLN:i:"; put_u64(g, len); g += "\tKC:i:"; put_u64(g, v.kc[u]); g += "\tkm:f:"; put_km(g, v.kc[u], nk, nullptr);
                g.push_back('\n');
            }
        }
        for (const cdbg_unitig_view *vp : views) {
            const cdbg_unitig_view &v = *vp;
            for (uint64_t u = 0; u < v.n_unitigs; u++)
                for (uint64_t i = v.link_off[u]; i < v.link_off[u + 1]; i++) {
                    const cdbg_link_t &l = v.links[i];
                    g += "L\t"; put_u64(g, v.first_unitig + u); g.push_back('\t'); g.push_back((char)l.from_sign); g.push_back('\t'); put_u64(g, l.to);
                    g.push_back('\t'); g.push_back((char)l.to_sign); g.push_back('\t'); put_u64(g, (uint64_t)(k - 1)); g += "M\n";
                }
        }
        FILE *gf = fopen(o.out_gfa.c_str(), "wb");
        if (!gf) { if (ctx_for_cleanup) cdbg_destroy(ctx_for_cleanup); die("Unable to open output file '" + o.out_gfa + "'"); }
        fwrite(g.data(), 1, g.size(), gf);
        fclose(gf);
    }
}

// -nb-gpus N: the reads are dealt round-robin (by record) to N contexts, one per GPU, each driven by its own thread; the library
// shards the graph by minimizer and exchanges k-mers, junction records, glue and link records over NCCL; every context returns
// the unitigs whose global ids fall in its range.
int run_sharded(const Options &o, ByteBuf &stream, uint64_t n_records, double t0) {
    const int N = o.nb_gpus;
    std::vector<ByteBuf> share(N);
    {
        const uint8_t *p = stream.data(), *e = p + stream.size();
        uint64_t r = 0;
        for (ByteBuf &sb : share) sb.reserve(stream.size() / N + 4096);
        while (p < e) {
            const uint8_t *nl = (const uint8_t *)memchr(p, '\n', (size_t)(e - p));
            const uint8_t *le = nl ? nl + 1 : e;
            share[r % N].append(p, (size_t)(le - p));
            r++; p = le;
        }
    }
    const double t1 = now_s();
    std::vector<cdbg_ctx *> ctx(N, nullptr);
    auto destroy_all = [&]() { for (cdbg_ctx *c : ctx) if (c) cdbg_destroy(c); };
    {
        std::vector<int> crc(N, CDBG_OK);
        std::vector<std::thread> mk;
        for (int r = 0; r < N; r++)
            mk.emplace_back([&, r]() {
                cdbg_params prm{};
                prm.kmer_size = o.k; prm.abundance_min = o.amin; prm.abundance_max = o.amax; prm.minimizer_size = o.m; prm.device = o.device + r;
                prm.flags = (o.all_abund ? CDBG_FLAG_ALL_ABUNDANCE_COUNTS : 0) | (o.verbose > 1 ? CDBG_FLAG_VERBOSE : 0);
                prm.rank = r; prm.world_size = N;
                crc[r] = cdbg_create(&ctx[r], &prm);
            });
        for (auto &th : mk) th.join();
        for (int r = 0; r < N; r++)
            if (crc[r] != CDBG_OK) {
                std::string msg = ctx[r] ? cdbg_last_error(ctx[r]) : "cdbg_create failed";
                destroy_all();
                die(msg);
            }
    }
    uint8_t id[128];
    if (cdbg_nccl_unique_id(id) != CDBG_OK) { destroy_all(); die("NCCL is not available (cdbg_nccl_unique_id failed)"); }
    std::vector<cdbg_unitig_view> view(N);
    std::vector<cdbg_stats> st(N);
    std::vector<int> rcs(N, CDBG_OK);
    std::vector<std::thread> pool;
    for (int r = 0; r < N; r++)
        pool.emplace_back([&, r]() {
            int rc = cdbg_shard_init(ctx[r], id);
            if (rc == CDBG_OK) rc = cdbg_build(ctx[r], share[r].data(), share[r].size(), CDBG_MEM_HOST, &view[r], &st[r]);
            rcs[r] = rc;
        });
    for (auto &th : pool) th.join();
    for (int r = 0; r < N; r++)
        if (rcs[r] != CDBG_OK) { std::string msg = cdbg_last_error(ctx[r]); destroy_all(); die(msg); }
    const double t2 = now_s();
    std::vector<const cdbg_unitig_view *> views;
    for (int r = 0; r < N; r++) views.push_back(&view[r]);
    std::sort(views.begin(), views.end(), [](const cdbg_unitig_view *a, const cdbg_unitig_view *b) {
        return a->first_unitig != b->first_unitig ? a->first_unitig < b->first_unitig : a->n_unitigs < b->n_unitigs; });
    const std::string prefix = !o.out.empty() ? o.out : base_name(o.in);
    const std::string out_path = prefix + ".unitigs.fa";
    write_outputs(o, views, out_path, nullptr);
    const double t3 = now_s();
    if (o.verbose) {
        uint64_t inst = 0, solid = 0, links = 0;
        for (int r = 0; r < N; r++) { inst += st[r].n_instances; solid += st[r].n_solid; links += view[r].n_links; }
        printf("sequences: %llu   nucleotide stream: %llu bytes   GPUs: %d\n", (unsigned long long)n_records, (unsigned long long)stream.size(), N);
        printf("kmers: %llu instances, %llu solid\n", (unsigned long long)inst, (unsigned long long)solid);
        printf("unitigs: %llu, links: %llu\n", (unsigned long long)view[0].n_unitigs_total, (unsigned long long)links);
        printf("wall s: parse+deal %.3f  gpu (H2D+stages+exchanges+D2H) %.3f  write %.3f  total %.3f\n", t1 - t0, t2 - t1, t3 - t2, t3 - t0);
        printf("output: %s\n", out_path.c_str());
    }
    destroy_all();
    return EXIT_SUCCESS;
}

int main(int argc, char **argv) {
    if (argc > 1 && (strcmp(argv[1], "-version") == 0 || strcmp(argv[1], "-v") == 0)) {
        printf("BCALM 2 (B200-native cdbg), version v2.2.3-compatible\nUsing %s\n", cdbg_version());
        return EXIT_SUCCESS;
    }
    Options o = parse(argc, argv);
    const double t0 = now_s();
    if (o.verbose) printf("BCALM 2 (B200-native), k=%d abundance-min=%d input=%s\n", o.k, o.amin, o.in.c_str());

    g_parse_threads = o.cores > 0 ? o.cores : (int)std::max(1u, std::thread::hardware_concurrency());
    if (o.parse_only) {
        ByteBuf st0; uint64_t nrec = 0;
        const double tp0 = now_s();
        load_input(o.in, st0, nrec);
        const double tp1 = now_s();
        uint64_t h = 1469598103934665603ULL;                       // FNV-1a of the nucleotide stream
        for (size_t i = 0; i < st0.size(); i++) { h ^= st0.p[i]; h *= 1099511628211ULL; }
        if (o.verbose > 1) printf("parse seconds %.3f\n", tp1 - tp0);
        printf("records %llu bytes %llu fnv1a %016llx threads %d\n", (unsigned long long)nrec, (unsigned long long)st0.size(), (unsigned long long)h, g_parse_threads);
        return EXIT_SUCCESS;
    }
    if (o.nb_gpus > 1) {
        ByteBuf st0; uint64_t nrec = 0;
        load_input(o.in, st0, nrec);
        return run_sharded(o, st0, nrec, t0);
    }
    cdbg_params prm{};
    prm.kmer_size = o.k; prm.abundance_min = o.amin; prm.abundance_max = o.amax; prm.minimizer_size = o.m; prm.device = o.device;
    prm.flags = (o.all_abund ? CDBG_FLAG_ALL_ABUNDANCE_COUNTS : 0) | (o.verbose > 1 ? CDBG_FLAG_VERBOSE : 0);
    prm.rank = 0; prm.world_size = 1;
    // the CUDA context takes a few hundred ms to create: do it on its own thread while the input is read (and parsed, on the host path)
    cdbg_ctx *ctx = nullptr;
    int rc = CDBG_OK;
    double t_ctx = 0;
    std::thread creator([&]() { const double a = now_s(); rc = cdbg_create(&ctx, &prm); t_ctx = now_s() - a; });
    bool joined = false;
    auto wait_ctx = [&]() -> cdbg_ctx * {
        if (!joined) { creator.join(); joined = true; }
        if (rc != CDBG_OK) {
            std::string msg = ctx ? cdbg_last_error(ctx) : "cdbg_create failed";
            if (ctx) cdbg_destroy(ctx);
            die(msg);
        }
        return ctx;
    };

    ByteBuf stream;
    uint64_t n_records = 0;
    cdbg_unitig_view v{};
    cdbg_stats st{};
    double t1;
    if (!o.solid_in.empty()) {   // stages 2-4 only, from counted k-mers
        std::vector<uint64_t> km; std::vector<uint32_t> ct;
        load_counted_kmers(o, km, ct);
        wait_ctx();
        t1 = now_s();
        rc = cdbg_set_solid_kmers(ctx, km.data(), ct.data(), ct.size(), CDBG_MEM_HOST);
        if (rc == CDBG_OK) rc = cdbg_compact(ctx, nullptr);
        if (rc == CDBG_OK) rc = cdbg_link(ctx, nullptr);
        if (rc == CDBG_OK) rc = cdbg_get_unitigs(ctx, &v);
        if (rc == CDBG_OK) rc = cdbg_get_stats(ctx, &st);
    } else if (o.gpu_parse != 0) {
        // (measured on a B200 host: creating the context WHILE another thread faults in the file's pages takes 0.45 s instead of 0.2 s --
        // both fight over the process's address-space lock -- so on this path the file is read after the context exists)
        wait_ctx();
        g_gpu_parse_wait = wait_ctx;
        load_input(o.in, stream, n_records);           // GPU-parsed files are already on the device; `stream` holds the rest
        wait_ctx();
        t1 = now_s();
        rc = stream.size() || n_records == 0 ? cdbg_push_reads(ctx, stream.data(), stream.size(), CDBG_MEM_HOST) : CDBG_OK;
        if (rc == CDBG_OK) rc = cdbg_count(ctx, nullptr);
        if (rc == CDBG_OK) rc = cdbg_compact(ctx, nullptr);
        if (rc == CDBG_OK) rc = cdbg_link(ctx, nullptr);
        if (rc == CDBG_OK) rc = cdbg_get_unitigs(ctx, &v);
        if (rc == CDBG_OK) rc = cdbg_get_stats(ctx, &st);
    } else {
        load_input(o.in, stream, n_records);
        wait_ctx();
        t1 = now_s();
        rc = cdbg_build(ctx, stream.data(), stream.size(), CDBG_MEM_HOST, &v, &st);
    }
    if (rc != CDBG_OK) { std::string msg = cdbg_last_error(ctx); cdbg_destroy(ctx); die(msg); }
    const double t2 = now_s();
    if (!o.solid_out.empty()) write_counted_kmers(o.solid_out, ctx, o.k);
    if (o.kff) {
        write_kff((!o.out.empty() ? o.out : base_name(o.in.empty() ? o.solid_in : o.in)) + ".kff", ctx, o.k);
    }

    const std::string prefix = !o.out.empty() ? o.out : base_name(o.in.empty() ? o.solid_in : o.in);
    const std::string out_path = prefix + ".unitigs.fa";
    std::vector<const cdbg_unitig_view *> views{&v};
    write_outputs(o, views, out_path, ctx);
    const double t3 = now_s();

    if (o.verbose) {
        printf("sequences: %llu   nucleotide stream: %llu bytes\n", (unsigned long long)n_records, (unsigned long long)stream.size());
        printf("kmers: %llu instances, %llu distinct, %llu solid\n", (unsigned long long)st.n_instances, (unsigned long long)st.n_distinct,
               (unsigned long long)st.n_solid);
        printf("unitigs: %llu (%llu nt), links: %llu, compactions: %llu, cycles cut: %llu\n", (unsigned long long)st.n_unitigs,
               (unsigned long long)st.total_nt, (unsigned long long)st.n_links, (unsigned long long)st.n_compactions, (unsigned long long)st.n_cycles);
        printf("device ms: count %.3f (K1 %.3f+%.3f, K2 %.3f)  compact %.3f (K3 %.3f, K5 %.3f, rank %.3f, spell %.3f)  link %.3f  fetch %.3f\n",
               st.ms_count, st.ms_k1_hist, st.ms_k1_scatter, st.ms_k2, st.ms_compact, st.ms_k3, st.ms_k5, st.ms_k6_rank, st.ms_k6_spell, st.ms_link,
               st.ms_fetch);
        printf("wall s: parse %.3f  gpu (H2D+stages+D2H) %.3f  write %.3f  total %.3f  (CUDA context %.3f, created while the input was read%s)\n", t1 - t0, t2 - t1,
               t3 - t2, t3 - t0, t_ctx, o.gpu_parse != 0 ? "; text parsed on the GPU" : "");
        printf("output: %s\n", out_path.c_str());
    }
    // the output files are closed; tearing the CUDA context down costs ~0.1 s that nobody needs
    fflush(nullptr);
    _exit(EXIT_SUCCESS);
}
