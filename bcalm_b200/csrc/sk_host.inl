// sk_host.inl -- host orchestration of stage 1 on super-k-mers (sk.cuh); included by cdbg.cu inside its anonymous namespace.
//
//   sk1_extract   reads -> super-k-mer records in Pc coarse bins (hash of the minimizer); ONE pass into regions sized from the
//                 byte count, exact histogram + redo when a region overflows
//   sk2p_*        fine partition of the records (F fine bins per coarse bin, exact offsets)
//   sk2c_count    one CTA per fine bin: expand in shared memory, count, solidity filter
// The same three steps serve the sharded path (shard_host.inl): there the coarse bins are destination-major, the exchange X1
// carries the records, and the fine partition runs over the received (local bin, source) runs.

#ifndef SK_DEFAULT
#define SK_DEFAULT 1   // stage 1 on super-k-mers by default for k <= 31 (CDBG_S1=kmer selects the k-mer-instance path of k1_tma.cuh / k2_smem.cuh)
#endif

struct SKPlan {
    int m;              // minimizer length used for bucketing
    uint32_t slots;     // table slots per counting CTA
    double per_fine;    // k-mer instances a fine bin is sized for
    uint32_t Pc_local;  // coarse bins per GPU
    uint32_t F;         // fine bins per coarse bin
    double rec_per_byte;  // estimated super-k-mer records per input byte (sizes the single-pass regions)
    uint32_t V;         // waves: stage 1 runs V times over the input, each time keeping 1/V of the buckets (1 unless the fine bins do not
                        // fit K1_MAX_BINS x SK2P_MAX_F or the records of the whole input do not fit the device memory)
};

bool sk_enabled(const cdbg_ctx *ctx) {
    const char *s = getenv("CDBG_S1");
    if (ctx->W != 1) return false;
    return SK_DEFAULT ? !(s && strcmp(s, "kmer") == 0) : (s && strcmp(s, "sk") == 0);
}

// Counting kernel: sk2d_count (distinct records first, CDBG_SK2C=dedup) or sk2c_count (every record expanded, CDBG_SK2C=direct).
#ifndef SK_DEDUP_DEFAULT
#define SK_DEDUP_DEFAULT 0
#endif
#ifndef SK_TWO_PHASE_DEFAULT
#define SK_TWO_PHASE_DEFAULT 1   // sk2c_count's two-phase insert (C2 on a B200: table pass 1.246 -> 1.205 ms, 1.5 Gbases slice 13.65 -> 13.23 ms;
                                 // profiles/r02_sk_ab4_*); CDBG_SK_INSERT=loop selects the plain probe loop
#endif
bool sk_dedup_enabled() {
    const char *s = getenv("CDBG_SK2C");
    return SK_DEDUP_DEFAULT ? !(s && strcmp(s, "direct") == 0) : (s && strcmp(s, "dedup") == 0);
}

// Minimizer length.  A bucket is counted in a shared-memory table, so the k-mers of one minimizer must stay a small part of a
// bucket, or buckets overflow their table and are split (each split re-expands the whole bucket).  Measured on a B200
// (profiles/r02_sk_ab_*): C2 (1e8 instances) m = 11 -> 346 splits, 12 -> 10, 14 -> 0; 1.2e9 instances m = 12 -> 21 835 splits.
// With no splits the table pass still gets faster as m grows (C2: m = 13 1.356 ms, 15 1.336, 16 1.266: buckets are more alike).
// Rule: 4^m >= 40 x instances, capped at 16 (ranks are 32-bit) -- m = 16 from ~10 Mbases of input on.  Lower bound: the window k-m+1 must not exceed the record capacity (53-k k-mers), else every natural run is split.
int sk_choose_m(int k, int user_m, uint64_t total_bytes) {
    if (getenv("CDBG_SK_M")) { int m = atoi(getenv("CDBG_SK_M")); return std::max(1, std::min(m, std::min(16, k - 1))); }
    int m = 7;
    const double need = (double)total_bytes * 0.8 * 40.0;
    while (m < 16 && pow(4.0, m) < need) m++;
    m = std::max(m, 2 * k - SK_MAX_NT);
    if (user_m > m) m = user_m;
    return std::max(1, std::min(m, std::min(16, k - 1)));
}

SKPlan sk_plan(const cdbg_ctx *ctx, uint64_t total_bytes, uint32_t world) {
    SKPlan pl{};
    pl.m = sk_choose_m(ctx->k, ctx->prm.minimizer_size, total_bytes);
    pl.slots = getenv("CDBG_SK_SLOTS") && atoi(getenv("CDBG_SK_SLOTS")) == 8192 ? 8192u : 4096u;
    double fill = getenv("CDBG_SK_FILL") ? atof(getenv("CDBG_SK_FILL")) : 1.5;
    if (!(fill >= 0.05 && fill <= 8.0)) fill = 1.5;
    pl.per_fine = fill * pl.slots;
    const double n_fine = std::max(1.0, (double)total_bytes * 0.8 / pl.per_fine);     // over all GPUs
    uint64_t pc = (uint64_t)ceil(sqrt(n_fine / world));                                // per GPU: coarse and fine fan-out about equal
    if (getenv("CDBG_SK_P1")) pc = strtoull(getenv("CDBG_SK_P1"), nullptr, 10);
    const int w = ctx->k - pl.m + 1;
    pl.rec_per_byte = std::min(0.6, 1.6 / (w + 1) + 0.012) * 1.15;
    if (getenv("CDBG_SK_DENSITY")) pl.rec_per_byte = atof(getenv("CDBG_SK_DENSITY"));   // (tests shrink it to force the exact fallback)
    // waves (the reference's passes, ConfigurationAlgorithm.cpp:350,399-418: nb_passes from the volume and the memory budget):
    //   capacity: one wave addresses K1_MAX_BINS / world coarse x SK2P_MAX_F fine bins per GPU (beyond that bins overflow their table and
    //             are split again and again inside the counting kernel);
    //   memory:   the records of one wave exist three times at the peak (K1 regions, send / receive or fine-partitioned copy).
    const double fine_cap = (double)(K1_MAX_BINS / world) * SK2P_MAX_F * world;
    uint64_t V = (uint64_t)ceil(n_fine / (fine_cap * 1.5));       // (a fine bin may hold 1.5x its nominal share before it splits often)
    if (ctx->total_mem) {          // (the device's total memory, read once at cdbg_create: the same on every rank, no driver call per build)
        const double rec_bytes = (double)total_bytes / world * pl.rec_per_byte * 16.0 * 3.2;
        const double budget = 0.55 * (double)ctx->total_mem;
        V = std::max<uint64_t>(V, (uint64_t)ceil(rec_bytes / budget));
    }
    if (getenv("CDBG_SK_WAVES")) V = strtoull(getenv("CDBG_SK_WAVES"), nullptr, 10);
    pl.V = (uint32_t)std::min<uint64_t>(64, std::max<uint64_t>(1, V));
    const double n_fine_wave = std::max(1.0, n_fine / pl.V);
    pc = (uint64_t)ceil(sqrt(n_fine_wave / world));
    if (getenv("CDBG_SK_P1")) pc = strtoull(getenv("CDBG_SK_P1"), nullptr, 10);
    pl.Pc_local = (uint32_t)std::min<uint64_t>(K1_MAX_BINS / world, std::max<uint64_t>(1, pc));
    pl.F = (uint32_t)std::min<uint64_t>(SK2P_MAX_F, std::max<uint64_t>(1, (uint64_t)ceil(n_fine_wave / ((double)pl.Pc_local * world))));
    return pl;
}

// K1 on super-k-mers over this context's reads.  Pc coarse bins; on return d_recs holds the records, bin b in
// [h_base[b], h_base[b] + h_cnt[b]).  Reads pinned host memory in place when ctx->bytes_hostmapped (single pass only).
int sk_run_k1(cdbg_ctx *ctx, const SKPlan &pl, uint32_t Pc, uint32_t wave, ulonglong2 **d_recs_out, std::vector<uint64_t> &h_cnt, std::vector<uint64_t> &h_base) {
    const uint64_t n = ctx->n_bytes;
    const uint64_t n_tiles = (n + K1_TILE - 1) / K1_TILE;
    int rc;
    h_cnt.assign(Pc, 0); h_base.assign(Pc + 1, 0);
    unsigned long long *d_cursor = nullptr; uint64_t *d_base = nullptr, *d_cap = nullptr;
    if ((rc = dalloc(ctx, &d_cursor, Pc))) return rc;
    if ((rc = dalloc(ctx, &d_base, Pc + 1))) return rc;
    if ((rc = dalloc(ctx, &d_cap, Pc + 1))) return rc;
    ulonglong2 *d_recs = nullptr;
    *d_recs_out = nullptr;
    if (!n_tiles) { if ((rc = dalloc(ctx, &d_recs, 1))) return rc; *d_recs_out = d_recs; dfree(ctx, d_cursor); dfree(ctx, d_base); dfree(ctx, d_cap); return CDBG_OK; }

    SK1Params kp{};
    kp.bytes = ctx->d_bytes; kp.n_bytes = n; kp.n_tiles = n_tiles; kp.k = ctx->k; kp.m = pl.m; kp.Pc = Pc; kp.V = pl.V; kp.wave = wave;
    kp.bin_base = d_base; kp.bin_cap = d_cap; kp.bin_cursor = d_cursor; kp.overflow = ctx->d_flag + 2;
    kp.n_kmers = ctx->d_ctr + 8; kp.n_runs = ctx->d_ctr + 9;
    kp.tail_start = n > 2 * (uint64_t)K1T_RAW ? ((n - 2 * (uint64_t)K1T_RAW) & ~15ULL) : 0;
    uint8_t *d_tail = nullptr;
    if (ctx->bytes_hostmapped && pl.V > 1) {          // every wave reads the input again: bring it into HBM once
        uint8_t *nb = nullptr;
        if ((rc = dalloc(ctx, &nb, n + 64))) return rc;
        CK(cudaMemcpyAsync(nb, ctx->h_bytes_src, n, cudaMemcpyHostToDevice, ctx->stream));
        ctx->d_bytes = nb; ctx->cap_bytes = n; ctx->bytes_borrowed = false; ctx->bytes_hostmapped = false;
        kp.bytes = nb;
    }
    if (ctx->bytes_hostmapped) {
        if ((rc = dalloc(ctx, &d_tail, n - kp.tail_start + 16))) return rc;
        CK(cudaMemcpyAsync(d_tail, ctx->h_bytes_src + kp.tail_start, n - kp.tail_start, cudaMemcpyHostToDevice, ctx->stream));
        kp.tail = d_tail;
    } else kp.tail = ctx->d_bytes + kp.tail_start;
    const size_t sm = sk1_smem_bytes(Pc);
    CK(cudaFuncSetAttribute(sk1_extract<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 2);

    // single pass into regions of the estimated size
    const uint64_t slack = getenv("CDBG_SK_SLACK") ? strtoull(getenv("CDBG_SK_SLACK"), nullptr, 10) : 512;   // (tests use 0 to force the exact fallback)
    const uint64_t cap = (((uint64_t)((double)n * pl.rec_per_byte / Pc / pl.V) + slack) + 7) & ~7ULL;
    std::vector<uint64_t> h_cap(Pc + 1, cap);
    for (uint32_t b = 0; b <= Pc; b++) h_base[b] = (uint64_t)b * cap;
    if ((rc = dalloc(ctx, &d_recs, (uint64_t)Pc * cap + 1))) return rc;
    if ((rc = ctr_zero(ctx))) return rc;
    CK(cudaMemsetAsync(d_cursor, 0, (size_t)Pc * 8, ctx->stream));
    CK(cudaMemcpyAsync(d_base, h_base.data(), (size_t)(Pc + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_cap, h_cap.data(), (size_t)(Pc + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    kp.out = d_recs;
    tstart(ctx, T_K1S);
    LAUNCH(sk1_extract<1>, grid, K1_THREADS, sm, kp);
    tstop(ctx, T_K1S);
    SYNC_STREAM(ctx);   // (a copy to pageable memory blocks inside the call: wait by polling first, see stream_sync)
    CK(cudaMemcpyAsync(h_cnt.data(), d_cursor, (size_t)Pc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if ((rc = ctr_read(ctx))) return rc;
    if (((int *)(ctx->h_ctr + 64))[2]) {
        // a region overflowed (short reads, many breaks, heavy minimizers): the cursors ARE the exact histogram -- every run was
        // counted, whether it fitted or not -- so lay the regions out exactly and run the pass again
        ctx->st.k1_fallback = 1;
        if (ctx->bytes_hostmapped) {      // the second pass reads the input again: bring it into HBM first
            uint8_t *nb = nullptr;
            if ((rc = dalloc(ctx, &nb, n + 64))) return rc;
            CK(cudaMemcpyAsync(nb, ctx->h_bytes_src, n, cudaMemcpyHostToDevice, ctx->stream));
            ctx->d_bytes = nb; ctx->cap_bytes = n; ctx->bytes_borrowed = false; ctx->bytes_hostmapped = false;
            kp.bytes = nb; kp.tail = nb + kp.tail_start;
        }
        dfree(ctx, d_recs);
        uint64_t acc = 0;
        for (uint32_t b = 0; b < Pc; b++) { h_base[b] = acc; h_cap[b] = h_cnt[b]; acc += (h_cnt[b] + 7) & ~7ULL; }
        h_base[Pc] = acc;
        if ((rc = dalloc(ctx, &d_recs, acc + 1))) return rc;
        if ((rc = ctr_zero(ctx))) return rc;
        CK(cudaMemsetAsync(d_cursor, 0, (size_t)Pc * 8, ctx->stream));
        CK(cudaMemcpyAsync(d_base, h_base.data(), (size_t)(Pc + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d_cap, h_cap.data(), (size_t)(Pc + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        kp.out = d_recs;
        tstart(ctx, T_K1H);
        LAUNCH(sk1_extract<1>, grid, K1_THREADS, sm, kp);
        tstop(ctx, T_K1H);
        std::vector<uint64_t> again(Pc);
        SYNC_STREAM(ctx);   // (a copy to pageable memory blocks inside the call: wait by polling first, see stream_sync)
        CK(cudaMemcpyAsync(again.data(), d_cursor, (size_t)Pc * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = ctr_read(ctx))) return rc;
        if (((int *)(ctx->h_ctr + 64))[2] || again != h_cnt) return ctx->fail(CDBG_ERR_INTERNAL, "super-k-mer extraction is not reproducible (exact pass overflowed)");
    }
    if (wave == 0) { ctx->st.n_instances = ctx->h_ctr[8]; ctx->st.n_superkmers = 0; }   // (every wave sees all k-mer positions)
    ctx->st.n_superkmers += ctx->h_ctr[9];
    dfree(ctx, d_tail); dfree(ctx, d_cursor); dfree(ctx, d_base); dfree(ctx, d_cap);
    *d_recs_out = d_recs;
    return CDBG_OK;
}

// Fine partition + counting.  d_recs holds R records in P1 coarse bins / runs (h_cnt, h_off); `h_group` (optional): the runs of one
// local bin (sharded mode: one per source GPU) share their fine bins.  I = k-mer instances (sizes the solid buffers).
// Leaves ctx->d_solid_kmers / d_solid_counts / S.
int sk_partition_count(cdbg_ctx *ctx, const SKPlan &pl, const ulonglong2 *d_recs, uint64_t I, uint32_t P1, const std::vector<uint64_t> &h_cnt,
                       const std::vector<uint64_t> &h_off, const std::vector<uint32_t> *h_group = nullptr, uint32_t n_groups = 0, uint32_t wave = 0) {
    int rc;
    const uint32_t G = h_group ? n_groups : P1, F = pl.F;
    const uint64_t NF = (uint64_t)G * F;
    std::vector<uint32_t> h_tile_first(P1 + 1);
    std::vector<uint64_t> h_range(2 * (size_t)P1);
    uint64_t tiles = 0, R = 0;
    for (uint32_t b = 0; b < P1; b++) {
        h_tile_first[b] = (uint32_t)tiles; tiles += (h_cnt[b] + SK2P_TILE - 1) / SK2P_TILE;
        h_range[b] = h_off[b]; h_range[P1 + b] = h_off[b] + h_cnt[b]; R += h_cnt[b];
    }
    h_tile_first[P1] = (uint32_t)tiles;
    if (tiles >= (1ULL << 32) || NF >= (1ULL << 32)) return ctx->fail(CDBG_ERR_PARAM, "too many super-k-mer tiles for one device");

    uint32_t *d_tile_first = nullptr, *d_fine_cnt = nullptr, *d_fine_cursor = nullptr, *d_group = nullptr;
    uint64_t *d_fine_off = nullptr, *d_bin_range = nullptr;
    ulonglong2 *d_recs2 = nullptr;
    if ((rc = dalloc(ctx, &d_bin_range, 2 * (size_t)P1))) return rc;
    if ((rc = dalloc(ctx, &d_tile_first, P1 + 1))) return rc;
    if (h_group) {
        if ((rc = dalloc(ctx, &d_group, P1))) return rc;
        CK(cudaMemcpyAsync(d_group, h_group->data(), (size_t)P1 * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = dalloc(ctx, &d_fine_cnt, NF + 1))) return rc;
    if ((rc = dalloc(ctx, &d_fine_cursor, NF + 1))) return rc;
    if ((rc = dalloc(ctx, &d_fine_off, NF + 2))) return rc;
    if ((rc = dalloc(ctx, &d_recs2, R + 1))) return rc;
    CK(cudaMemsetAsync(d_fine_cnt, 0, (NF + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(d_fine_cursor, 0, (NF + 1) * 4, ctx->stream));
    CK(cudaMemcpyAsync(d_bin_range, h_range.data(), 2 * (size_t)P1 * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_tile_first, h_tile_first.data(), (size_t)(P1 + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));

    tstart(ctx, T_K2);
    SK2PParams pp{};
    pp.recs = d_recs; pp.bin_beg = d_bin_range; pp.bin_end = d_bin_range + P1; pp.tile_first = d_tile_first; pp.group = d_group;
    pp.P1 = P1; pp.F = F; pp.n_tiles = (uint32_t)tiles;
    const bool dedup = sk_dedup_enabled();
    pp.canon = dedup ? 1 : 0; pp.k = ctx->k;
    pp.fine_cnt = d_fine_cnt; pp.fine_off = d_fine_off; pp.fine_cursor = d_fine_cursor; pp.out = d_recs2;
    if (tiles) {
        tstart(ctx, T_K2H);
        LAUNCH(sk2p_hist, (int)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * 4), SK2P_THREADS, 0, pp);
        tstop(ctx, T_K2H);
        uint64_t total = 0;
        if ((rc = exclusive_scan(ctx, d_fine_cnt, NF + 1, d_fine_off, &total))) return rc;   // entry NF = R
        if (total != R) return ctx->fail(CDBG_ERR_INTERNAL, "fine histogram lost super-k-mers (%llu != %llu)", (unsigned long long)total, (unsigned long long)R);
        tstart(ctx, T_K2S);
        CK(cudaFuncSetAttribute(sk2p_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sk2p_scatter_smem_bytes()));
        LAUNCH(sk2p_scatter, (int)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * 3), SK2P_THREADS, sk2p_scatter_smem_bytes(), pp);
        tstop(ctx, T_K2S);
    } else {
        CK(cudaMemsetAsync(d_fine_off, 0, (NF + 2) * 8, ctx->stream));
    }
    TRACE("SK fine partition");

    // the solid k-mers of all waves are appended to one pair of arrays: wave 0 sizes them for the whole input
    uint64_t solid_cap = wave == 0 ? std::max<uint64_t>(I <= (1ULL << 28) ? I : std::max<uint64_t>(I / 4, 1ULL << 28), 1) : ctx->sk_solid_cap;
    const uint64_t S_prev = wave == 0 ? 0 : ctx->S, D_prev = wave == 0 ? 0 : ctx->st.n_distinct;
    const uint32_t splits_prev = wave == 0 ? 0 : ctx->st.n_k2_splits;
    for (int attempt = 0; attempt < 2; attempt++) {
        if (attempt == 0 && wave == 0) {
            if ((rc = dalloc(ctx, &ctx->d_solid_kmers, solid_cap))) return rc;
            if ((rc = dalloc(ctx, &ctx->d_solid_counts, solid_cap))) return rc;
        }
        ctx->sk_solid_cap = solid_cap;
        if ((rc = ctr_zero(ctx))) return rc;
        if (S_prev || D_prev) {
            CK(cudaMemcpyAsync(ctx->d_ctr + 0, &S_prev, 8, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->d_ctr + 1, &D_prev, 8, cudaMemcpyHostToDevice, ctx->stream));
        }
        SK2CParams cp{};
        cp.recs = d_recs2; cp.fine_off = d_fine_off; cp.n_fine = (uint32_t)NF; cp.k = ctx->k; cp.amin = ctx->amin; cp.amax = ctx->amax;
        cp.solid_kmers = ctx->d_solid_kmers; cp.solid_counts = ctx->d_solid_counts; cp.solid_cap = solid_cap;
        cp.n_solid = ctx->d_ctr + 0; cp.n_distinct = ctx->d_ctr + 1; cp.n_splits = ctx->d_ctr + 2; cp.failed = ctx->d_flag + 1;
        tstart(ctx, T_K2C);
        if (R) {
#define SK2C_LAUNCH2(SL, TH, MU, KC, TW)                                                                                                     \
    do {                                                                                                                                    \
        const size_t sm_ = sk2c_smem_bytes<SL, TH>();                                                                                       \
        CK(cudaFuncSetAttribute(sk2c_count<SL, TH, MU, KC, TW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_));                     \
        const uint64_t per_sm_ = std::max<uint64_t>(1, std::min<uint64_t>(2048 / (TH), (227 * 1024) / (sm_ + 1024)));                       \
        LAUNCH((sk2c_count<SL, TH, MU, KC, TW>), (int)std::min<uint64_t>(NF, (uint64_t)ctx->n_sm * per_sm_), (TH), sm_, cp);                \
    } while (0)
#define SK2C_LAUNCH1(SL, TH, MU, KC)                                                                                  \
    do { if (two_phase) SK2C_LAUNCH2(SL, TH, MU, KC, true); else SK2C_LAUNCH2(SL, TH, MU, KC, false); } while (0)
#define SK2C_LAUNCH(SL, TH)                                                                                           \
    do {                                                                                                              \
        if (k31) { if (mul32) SK2C_LAUNCH1(SL, TH, true, 31); else SK2C_LAUNCH1(SL, TH, false, 31); }                 \
        else { if (mul32) SK2C_LAUNCH1(SL, TH, true, 0); else SK2C_LAUNCH1(SL, TH, false, 0); }                       \
    } while (0)
#define SK2D_LAUNCH1(SL, RS_, TH, MU, KC)                                                                                                    \
    do {                                                                                                                                    \
        const size_t sm_ = sk2d_smem_bytes<SL, RS_, TH>();                                                                                  \
        CK(cudaFuncSetAttribute(sk2d_count<SL, RS_, TH, MU, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_));                    \
        const uint64_t per_sm_ = std::max<uint64_t>(1, std::min<uint64_t>(2048 / (TH), (227 * 1024) / (sm_ + 1024)));                       \
        LAUNCH((sk2d_count<SL, RS_, TH, MU, KC>), (int)std::min<uint64_t>(NF, (uint64_t)ctx->n_sm * per_sm_), (TH), sm_, cp);               \
    } while (0)
#define SK2D_LAUNCH(SL, RS_, TH)                                                                                      \
    do {                                                                                                              \
        if (k31) { if (mul32) SK2D_LAUNCH1(SL, RS_, TH, true, 31); else SK2D_LAUNCH1(SL, RS_, TH, false, 31); }       \
        else { if (mul32) SK2D_LAUNCH1(SL, RS_, TH, true, 0); else SK2D_LAUNCH1(SL, RS_, TH, false, 0); }             \
    } while (0)
            const char *hm = getenv("CDBG_K2C_HASH");
            const bool mul32 = K2C_HASH_MUL32_DEFAULT ? !(hm && strcmp(hm, "mix64") == 0) : (hm && strcmp(hm, "mul32") == 0);
            const bool k31 = ctx->k == 31 && !(getenv("CDBG_SK_KC") && atoi(getenv("CDBG_SK_KC")) == 0);   // k = 31 has its own instantiation
            const int th = getenv("CDBG_SK_THREADS") ? atoi(getenv("CDBG_SK_THREADS")) : 256;
            const char *ins = getenv("CDBG_SK_INSERT");
            const bool two_phase = SK_TWO_PHASE_DEFAULT ? !(ins && strcmp(ins, "loop") == 0) : (ins && strcmp(ins, "two") == 0);
            const bool small_stage = getenv("CDBG_SK_WCAP") && atoi(getenv("CDBG_SK_WCAP")) == 128 && pl.slots == 4096 && th == 256 && two_phase;
            if (small_stage && !dedup) {
#define SK2C_SMALL(MU, KC)                                                                                                                  \
    do {                                                                                                                                    \
        const size_t sm_ = sk2c_smem_bytes<4096, 256, true>();                                                                              \
        CK(cudaFuncSetAttribute(sk2c_count<4096, 256, MU, KC, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_));          \
        const uint64_t per_sm_ = std::max<uint64_t>(1, std::min<uint64_t>(2048 / 256, (227 * 1024) / (sm_ + 1024)));                        \
        LAUNCH((sk2c_count<4096, 256, MU, KC, true, true>), (int)std::min<uint64_t>(NF, (uint64_t)ctx->n_sm * per_sm_), 256, sm_, cp);      \
    } while (0)
                if (k31) { if (mul32) SK2C_SMALL(true, 31); else SK2C_SMALL(false, 31); }
                else { if (mul32) SK2C_SMALL(true, 0); else SK2C_SMALL(false, 0); }
#undef SK2C_SMALL
            } else if (dedup) {
                const int rs = getenv("CDBG_SK_RS") ? atoi(getenv("CDBG_SK_RS")) : 1024;
                if (pl.slots == 8192) { if (rs == 2048) SK2D_LAUNCH(8192, 2048, 256); else SK2D_LAUNCH(8192, 1024, 256); }
                else { if (rs == 2048) SK2D_LAUNCH(4096, 2048, 256); else if (rs == 512) SK2D_LAUNCH(4096, 512, 256); else SK2D_LAUNCH(4096, 1024, 256); }
            } else if (pl.slots == 8192) { if (th == 512) SK2C_LAUNCH(8192, 512); else SK2C_LAUNCH(8192, 256); }
            else { if (th == 128) SK2C_LAUNCH(4096, 128); else if (th == 512) SK2C_LAUNCH(4096, 512); else SK2C_LAUNCH(4096, 256); }
#undef SK2C_LAUNCH
#undef SK2C_LAUNCH1
#undef SK2C_LAUNCH2
#undef SK2D_LAUNCH
#undef SK2D_LAUNCH1
        }
        tstop(ctx, T_K2C);
        if ((rc = ctr_read(ctx))) return rc;
        if (((int *)(ctx->h_ctr + 64))[1]) return ctx->fail(CDBG_ERR_INTERNAL, "K2: a fine bin does not fit a shared-memory table after %d splits", K2C_MAX_LEVEL);
        ctx->st.n_k2_splits = splits_prev + (uint32_t)ctx->h_ctr[2];
        const uint64_t S_now = ctx->h_ctr[0];
        ctx->st.n_distinct = ctx->h_ctr[1];
        if (S_now <= solid_cap) { ctx->S = S_now; break; }
        if (attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "solid buffer overflow");
        // grow: exact for this wave plus the same rate for the waves to come; the solid k-mers of earlier waves move over
        const uint64_t new_cap = S_now + (uint64_t)((double)(S_now - S_prev) * 1.1 * (pl.V - 1 - wave)) + 1024;
        uint64_t *nk = nullptr; uint32_t *nc = nullptr;
        if ((rc = dalloc(ctx, &nk, new_cap))) return rc;
        if ((rc = dalloc(ctx, &nc, new_cap))) return rc;
        if (S_prev) {
            CK(cudaMemcpyAsync(nk, ctx->d_solid_kmers, S_prev * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaMemcpyAsync(nc, ctx->d_solid_counts, S_prev * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        }
        dfree(ctx, ctx->d_solid_kmers); dfree(ctx, ctx->d_solid_counts);
        ctx->d_solid_kmers = nk; ctx->d_solid_counts = nc;
        solid_cap = new_cap;
    }
    tstop(ctx, T_K2);
    ctx->st.n_solid = ctx->S;
    dfree(ctx, d_bin_range); dfree(ctx, d_tile_first); dfree(ctx, d_group); dfree(ctx, d_fine_cnt); dfree(ctx, d_fine_cursor); dfree(ctx, d_fine_off);
    dfree(ctx, d_recs2);
    if (ctx->S >= (1ULL << 31)) return ctx->fail(CDBG_ERR_PARAM, "more than 2^31 solid k-mers on one device (%llu): shard over more GPUs", (unsigned long long)ctx->S);
    return CDBG_OK;
}

// stage 1 of a single-GPU context on super-k-mers
int stage_count_sk(cdbg_ctx *ctx) {
    int rc;
    tstart(ctx, T_COUNT);
    const SKPlan pl = sk_plan(ctx, ctx->n_bytes, 1);
    const uint32_t Pc = pl.Pc_local;
    ctx->st.n_count_bins = Pc;
    ctx->st.sk_m = (uint32_t)pl.m;
    ctx->st.sk_waves = pl.V;
    for (uint32_t wave = 0; wave < pl.V; wave++) {
        std::vector<uint64_t> h_cnt, h_base;
        ulonglong2 *d_recs = nullptr;
        if ((rc = sk_run_k1(ctx, pl, Pc, wave, &d_recs, h_cnt, h_base))) return rc;
        TRACE("SK K1");
        if ((rc = sk_partition_count(ctx, pl, d_recs, ctx->st.n_instances, Pc, h_cnt, h_base, nullptr, 0, wave))) return rc;
        dfree(ctx, d_recs);
    }
    tstop(ctx, T_COUNT);
    ctx->stage = ST_COUNTED;
    return CDBG_OK;
}
