// wide_host.inl -- stage orchestration for multi-word k-mers (k >= 32), included by cdbg.cu inside its anonymous namespace.
// Mirrors stage_count / stage_compact with the kernels of wide.cuh; K6 (ranking), the scans and K9 are shared with the k <= 31 path.

template <int W>
int stage_count_wide(cdbg_ctx *ctx) {
    const int k = ctx->k;
    const uint64_t n = ctx->n_bytes;
    int rc;
    tstart(ctx, T_COUNT);
    const uint64_t n_tiles = (n + KW1_TILE - 1) / KW1_TILE;
    uint32_t P = (uint32_t)std::min<uint64_t>(2048, std::max<uint64_t>(1, (n * 4 / 5 + 1999999) / 2000000));
    ctx->st.n_count_bins = P;
    unsigned long long *d_bin_count = nullptr, *d_bin_cursor = nullptr; uint64_t *d_bin_off = nullptr;
    if ((rc = dalloc(ctx, &d_bin_count, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_cursor, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_off, P + 1))) return rc;
    CK(cudaMemsetAsync(d_bin_count, 0, P * 8, ctx->stream));
    CK(cudaMemsetAsync(d_bin_cursor, 0, P * 8, ctx->stream));
    KW1Params kp{};
    kp.bytes = ctx->d_bytes; kp.n_bytes = n; kp.n_tiles = n_tiles; kp.k = k; kp.P = P;
    kp.bin_count = d_bin_count; kp.bin_off = d_bin_off; kp.bin_cursor = d_bin_cursor; kp.out_kmers = nullptr;
    std::vector<uint64_t> h_cnt(P, 0), h_off(P + 1, 0);
    uint64_t I = 0, max_bin = 0;
    const int g1 = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 4));
    if (n_tiles) {
        CK(cudaFuncSetAttribute(kw1_extract<W, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw1_smem_bytes(0, P)));
        CK(cudaFuncSetAttribute(kw1_extract<W, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw1_smem_bytes(1, P)));
        tstart(ctx, T_K1H);
        LAUNCH((kw1_extract<W, 0>), g1, KW1_THREADS, kw1_smem_bytes(0, P), kp);
        tstop(ctx, T_K1H);
        CK(cudaMemcpyAsync(h_cnt.data(), d_bin_count, P * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SYNC_STREAM(ctx);
    }
    for (uint32_t b = 0; b < P; b++) { h_off[b] = I; I += h_cnt[b]; max_bin = std::max(max_bin, h_cnt[b]); }
    h_off[P] = I;
    ctx->st.n_instances = I;
    uint64_t *d_kmers = nullptr;
    if ((rc = dalloc(ctx, &d_kmers, I * W + 1))) return rc;
    if (I) {
        CK(cudaMemcpyAsync(d_bin_off, h_off.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        kp.out_kmers = d_kmers;
        tstart(ctx, T_K1S);
        LAUNCH((kw1_extract<W, 1>), g1, KW1_THREADS, kw1_smem_bytes(1, P), kp);
        tstop(ctx, T_K1S);
    }
    auto table_cap = [](uint64_t x) { return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1024, (x * 3 / 2 + 63) & ~63ULL), 0xFFFFFF00u); };
    const uint32_t cap = table_cap(max_bin);
    uint32_t *d_tags = nullptr, *d_counts = nullptr; uint64_t *d_keys = nullptr;
    if ((rc = dalloc(ctx, &d_tags, cap))) return rc;
    if ((rc = dalloc(ctx, &d_counts, cap))) return rc;
    if ((rc = dalloc(ctx, &d_keys, (size_t)cap * W))) return rc;
    CK(cudaMemsetAsync(d_tags, 0, (size_t)cap * 4, ctx->stream));
    CK(cudaMemsetAsync(d_counts, 0, (size_t)cap * 4, ctx->stream));
    uint64_t solid_cap = std::max<uint64_t>(I <= (1ULL << 27) ? I : std::max<uint64_t>(I / 4, 1ULL << 27), 1);
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = dalloc(ctx, &ctx->d_solid_kmers, solid_cap * W))) return rc;
        if ((rc = dalloc(ctx, &ctx->d_solid_counts, solid_cap))) return rc;
        if ((rc = ctr_zero(ctx))) return rc;
        tstart(ctx, T_K2);
        for (uint32_t b = 0; b < P; b++) {
            const uint64_t nb = h_cnt[b];
            if (!nb) continue;
            const uint32_t cap_b = table_cap(nb);
            LAUNCH((kw2_insert<W>), grid_for(ctx, nb, 256, 8), 256, 0, d_kmers + h_off[b] * W, nb, d_tags, d_keys, d_counts, cap_b, ctx->d_flag);
            LAUNCH((kw2_emit<W>), grid_for(ctx, cap_b, 256, 8), 256, 0, d_tags, d_keys, d_counts, cap_b, ctx->amin, ctx->amax, ctx->d_solid_kmers,
                   ctx->d_solid_counts, solid_cap, ctx->d_ctr + 0, ctx->d_ctr + 1);
        }
        tstop(ctx, T_K2);
        if ((rc = ctr_read(ctx))) return rc;
        if (((int *)(ctx->h_ctr + 64))[0]) return ctx->fail(CDBG_ERR_INTERNAL, "K2 hash table overflow");
        ctx->S = ctx->h_ctr[0];
        ctx->st.n_distinct = ctx->h_ctr[1];
        if (ctx->S <= solid_cap) break;
        if (attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "solid buffer overflow");
        dfree(ctx, ctx->d_solid_kmers); dfree(ctx, ctx->d_solid_counts);
        solid_cap = ctx->S;
    }
    ctx->st.n_solid = ctx->S;
    if (ctx->S >= (1ULL << 31)) return ctx->fail(CDBG_ERR_PARAM, "more than 2^31 solid k-mers on one device");
    dfree(ctx, d_kmers); dfree(ctx, d_tags); dfree(ctx, d_counts); dfree(ctx, d_keys);
    dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
    tstop(ctx, T_COUNT);
    ctx->stage = ST_COUNTED;
    return CDBG_OK;
}

template <int W>
int stage_compact_wide(cdbg_ctx *ctx) {
    const int k = ctx->k;
    const uint64_t S = ctx->S, n2 = 2 * S;
    int rc;
    tstart(ctx, T_COMPACT);
    uint32_t P = (uint32_t)std::min<uint64_t>(4096, std::max<uint64_t>(1, (n2 + 499999) / 500000));
    ctx->st.n_join_bins = P;
    ctx->st.n_junction_records = n2;
    unsigned long long *d_bin_count = nullptr, *d_bin_cursor = nullptr;
    uint64_t *d_bin_off = nullptr, *d_rec_key = nullptr, *d_rec_val = nullptr;
    if ((rc = dalloc(ctx, &d_bin_count, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_cursor, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_off, P + 1))) return rc;
    if ((rc = dalloc(ctx, &d_rec_key, n2 * W + 1))) return rc;
    if ((rc = dalloc(ctx, &d_rec_val, n2 + 1))) return rc;
    CK(cudaMemsetAsync(d_bin_count, 0, P * 8, ctx->stream));
    CK(cudaMemsetAsync(d_bin_cursor, 0, P * 8, ctx->stream));
    KW3Params p3{};
    p3.solid_kmers = ctx->d_solid_kmers; p3.n_solid = S; p3.k = k; p3.P = P;
    p3.bin_count = d_bin_count; p3.bin_off = d_bin_off; p3.bin_cursor = d_bin_cursor; p3.rec_key = d_rec_key; p3.rec_val = d_rec_val;
    std::vector<uint64_t> h_cnt(P, 0), h_off(P + 1, 0);
    uint64_t max_bin = 0;
    tstart(ctx, T_K3);
    if (S) {
        CK(cudaFuncSetAttribute(kw3_ends<W, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3_smem_bytes(1, P)));
        const int g3 = (int)std::max<uint64_t>(1, std::min<uint64_t>((S + K3_THREADS - 1) / K3_THREADS, (uint64_t)ctx->n_sm * 8));
        LAUNCH((kw3_ends<W, 0>), g3, K3_THREADS, k3_smem_bytes(0, P), p3);
        CK(cudaMemcpyAsync(h_cnt.data(), d_bin_count, P * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SYNC_STREAM(ctx);
        uint64_t acc = 0;
        for (uint32_t b = 0; b < P; b++) { h_off[b] = acc; acc += h_cnt[b]; max_bin = std::max(max_bin, h_cnt[b]); }
        h_off[P] = acc;
        CK(cudaMemcpyAsync(d_bin_off, h_off.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH((kw3_ends<W, 1>), g3, K3_THREADS, k3_smem_bytes(1, P), p3);
    }
    tstop(ctx, T_K3);

    if ((rc = dalloc(ctx, &ctx->d_link, n2 + 1))) return rc;
    auto table_cap = [](uint64_t x) { return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1024, (x * 3 / 2 + 63) & ~63ULL), 0xFFFFFF00u); };
    const uint32_t cap = table_cap(max_bin);
    JoinTableW T{};
    if ((rc = dalloc(ctx, &T.tags, cap))) return rc;
    if ((rc = dalloc(ctx, &T.keys, (size_t)cap * W))) return rc;
    if ((rc = dalloc(ctx, &T.cnt, cap))) return rc;
    if ((rc = dalloc(ctx, &T.ids, (size_t)cap * 8))) return rc;
    CK(cudaMemsetAsync(T.tags, 0, (size_t)cap * 4, ctx->stream));
    CK(cudaMemsetAsync(T.cnt, 0, (size_t)cap * 4, ctx->stream));
    ctx->edge_cap = std::max<uint64_t>(S / 2, 1 << 16);
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = dalloc(ctx, &ctx->d_edges, ctx->edge_cap))) return rc;
        CK(cudaMemsetAsync(ctx->d_link, 0xFF, (n2 + 1) * 4, ctx->stream));
        if ((rc = ctr_zero(ctx))) return rc;
        tstart(ctx, T_K5);
        for (uint32_t b = 0; b < P; b++) {
            const uint64_t nb = h_cnt[b];
            if (!nb) continue;
            T.cap = table_cap(nb);
            LAUNCH((kw5_insert<W>), grid_for(ctx, nb, 256, 8), 256, 0, d_rec_key + h_off[b] * W, d_rec_val + h_off[b], nb, T, ctx->d_flag);
            LAUNCH(kw5_resolve, grid_for(ctx, T.cap, 256, 8), 256, 0, T, T.cap, ctx->d_link, ctx->d_edges, ctx->edge_cap, ctx->d_ctr + 0, ctx->d_ctr + 1);
        }
        tstop(ctx, T_K5);
        if ((rc = ctr_read(ctx))) return rc;
        if (((int *)(ctx->h_ctr + 64))[0]) return ctx->fail(CDBG_ERR_INTERNAL, "K5 hash table overflow");
        ctx->n_edges = ctx->h_ctr[0];
        ctx->st.n_compactions = ctx->h_ctr[1];
        if (ctx->n_edges + 2048 <= ctx->edge_cap) break;
        if (attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "edge buffer overflow");
        dfree(ctx, ctx->d_edges);
        ctx->edge_cap = ctx->n_edges + std::max<uint64_t>(ctx->n_edges / 8, 1 << 16);
    }
    dfree(ctx, d_rec_key); dfree(ctx, d_rec_val); dfree(ctx, T.tags); dfree(ctx, T.keys); dfree(ctx, T.cnt); dfree(ctx, T.ids);
    dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);

    uint64_t *d_pair = nullptr, *d_uid_scan = nullptr, *d_nt_scan = nullptr;
    tstart(ctx, T_K6R);
    if ((rc = run_ranking(ctx, ctx->d_link, S, nullptr, &ctx->d_edges, &ctx->edge_cap, &ctx->n_edges, &d_pair))) return rc;
    tstop(ctx, T_K6R);
    tstart(ctx, T_K6S);
    uint64_t U = 0, total_nt = 0;
    if ((rc = number_units(ctx, d_pair, S, nullptr, &ctx->d_info, &d_uid_scan, &d_nt_scan, &U, &total_nt))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_unitig_of_kmer, S + 1))) return rc;
    ctx->U = U; ctx->total_nt = total_nt;
    dfree(ctx, d_pair);
    if ((rc = dalloc(ctx, &ctx->d_seq, total_nt + 1))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_seq_off, U + 1))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_kc, U + 1))) return rc;
    CK(cudaMemsetAsync(ctx->d_kc, 0, (U + 1) * 8, ctx->stream));
    if (ctx->prm.flags & CDBG_FLAG_ALL_ABUNDANCE_COUNTS) { if ((rc = dalloc(ctx, &ctx->d_abund, S + 1))) return rc; }
    if (S) {
        SpellParams sp{};
        sp.solid_kmers = ctx->d_solid_kmers; sp.solid_counts = ctx->d_solid_counts; sp.S = S; sp.k = k;
        sp.info = ctx->d_info; sp.uid_scan = d_uid_scan; sp.nt_scan = d_nt_scan;
        sp.seq = ctx->d_seq; sp.seq_off = ctx->d_seq_off; sp.kc = ctx->d_kc; sp.abund = ctx->d_abund;
        sp.unitig_of_kmer = ctx->d_unitig_of_kmer; sp.link = ctx->d_link; sp.term_begin = nullptr; sp.term_end = nullptr;
        LAUNCH((kw8_spell<W>), grid_for(ctx, S, 256, 8), 256, 0, sp);
    }
    CK(cudaMemcpyAsync(ctx->d_seq_off + U, &ctx->total_nt, 8, cudaMemcpyHostToDevice, ctx->stream));
    dfree(ctx, d_uid_scan); dfree(ctx, d_nt_scan);
    tstop(ctx, T_K6S);
    ctx->st.n_unitigs = U; ctx->st.total_nt = total_nt;
    tstop(ctx, T_COMPACT);
    ctx->stage = ST_COMPACTED;
    return CDBG_OK;
}

int stage_count_wide_dispatch(cdbg_ctx *ctx) {
    switch (ctx->W) { case 2: return stage_count_wide<2>(ctx); case 3: return stage_count_wide<3>(ctx); case 4: return stage_count_wide<4>(ctx); }
    return ctx->fail(CDBG_ERR_INTERNAL, "bad word count %d", ctx->W);
}
int stage_compact_wide_dispatch(cdbg_ctx *ctx) {
    switch (ctx->W) { case 2: return stage_compact_wide<2>(ctx); case 3: return stage_compact_wide<3>(ctx); case 4: return stage_compact_wide<4>(ctx); }
    return ctx->fail(CDBG_ERR_INTERNAL, "bad word count %d", ctx->W);
}
