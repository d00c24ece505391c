// shard_wide_host.inl -- the sharded (multi-GPU) path for multi-word k-mers (32 <= k <= 127), included by cdbg.cu inside its
// anonymous namespace after shard_host.inl and wide_host.inl.
//
// Only the front of each stage depends on the k-mer width: extraction, the k-mer exchange X1 and counting (stage 1), and end
// records, their exchange X2 and the join (stage 2).  Everything behind the join -- link messages, local ranking, pieces, glue,
// links -- works on ids and is shared with the k <= 31 path (shard_host.inl).
//
// Ownership here is by HASH, not by minimizer: a canonical k-mer lives on GPU  floor(bin / P_local)  of its multiply-shift hash bin
// (the bins of kw1_extract / kw3_ends are destination-major by construction when P = world * P_local), a junction likewise by the
// hash of its canonical (k-1)-mer.  Any function of the canonical value gives the same graph; minimizer ownership (which keeps
// ~90 % of the junction records local in the k <= 31 path) is the next step for this one.  Like wide.cuh it favours clarity over
// speed: one table per local bin, one insert launch per (bin, source) run.

template <int WW>
int shard_stage_count_wide(cdbg_ctx *ctx) {
    const int W = ctx->world, me = ctx->rank, k = ctx->k;
    const uint64_t n = ctx->n_bytes;
    int rc;
    tstart(ctx, T_COUNT);
    ctx->trace_t0 = host_now();
    uint64_t all_n[MAX_WORLD], total_bytes = 0;
    if ((rc = allgather_host_u64(ctx, &n, 1, all_n))) return rc;
    for (int r = 0; r < W; r++) total_bytes += all_n[r];
    const uint32_t P_local = (uint32_t)std::min<uint64_t>(2048 / W, std::max<uint64_t>(1, (total_bytes / W * 4 / 5 + 1999999) / 2000000));
    const uint32_t P = P_local * W;
    ctx->st.n_count_bins = P_local;
    const uint64_t n_tiles = (n + KW1_TILE - 1) / KW1_TILE;
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
    const int g1 = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 4));
    CK(cudaFuncSetAttribute(kw1_extract<WW, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw1_smem_bytes(0, P)));
    CK(cudaFuncSetAttribute(kw1_extract<WW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kw1_smem_bytes(1, P)));
    if (n_tiles) {
        tstart(ctx, T_K1H);
        LAUNCH((kw1_extract<WW, 0>), g1, KW1_THREADS, kw1_smem_bytes(0, P), kp);
        tstop(ctx, T_K1H);
        SYNC_STREAM(ctx);   // (a copy to pageable memory blocks inside the call: wait by polling first, see stream_sync)
        CK(cudaMemcpyAsync(h_cnt.data(), d_bin_count, P * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SYNC_STREAM(ctx);
    }
    uint64_t I_out = 0;
    for (uint32_t b = 0; b < P; b++) { h_off[b] = I_out; I_out += h_cnt[b]; }
    h_off[P] = I_out;
    ctx->st.n_instances = I_out;
    std::vector<uint64_t> all_cnt((size_t)W * P);
    if ((rc = allgather_host_u64(ctx, h_cnt.data(), P, all_cnt.data()))) return rc;
    uint64_t *d_send = nullptr;
    if ((rc = dalloc(ctx, &d_send, I_out * WW + 1))) return rc;
    if (I_out) {
        CK(cudaMemcpyAsync(d_bin_off, h_off.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        kp.out_kmers = d_send;
        tstart(ctx, T_K1S);
        LAUNCH((kw1_extract<WW, 1>), g1, KW1_THREADS, kw1_smem_bytes(1, P), kp);
        tstop(ctx, T_K1S);
    }
    TRACE("k1 (wide) + allgather counts");
    // X1: k-mers (WW words each) to their counting owner
    uint64_t send_cnt[MAX_WORLD] = {0}, recv_cnt[MAX_WORLD] = {0};
    std::vector<uint64_t> run_off((size_t)P_local * W), run_len((size_t)P_local * W);
    plan_shard_exchange(W, me, P_local, all_cnt.data(), send_cnt, recv_cnt, run_off.data(), run_len.data());
    void *recv = nullptr; uint64_t I_in = 0;
    if ((rc = exchange_with_counts(ctx, d_send, send_cnt, recv_cnt, 8 * WW, &recv, &I_in))) return rc;
    uint64_t *d_kmers = (uint64_t *)recv;
    dfree(ctx, d_send);
    TRACE("X1 kmer exchange (wide)");

    // K2: one table per local bin, filled run by run (one run per source rank)
    uint64_t max_bin = 0;
    std::vector<uint64_t> bin_tot(P_local, 0);
    for (uint32_t b = 0; b < P_local; b++) { for (int s = 0; s < W; s++) bin_tot[b] += run_len[(size_t)b * W + s]; max_bin = std::max(max_bin, bin_tot[b]); }
    const uint32_t cap = table_cap_for(max_bin);
    uint32_t *d_tags = nullptr, *d_counts = nullptr; uint64_t *d_keys = nullptr;
    if ((rc = dalloc(ctx, &d_tags, cap))) return rc;
    if ((rc = dalloc(ctx, &d_counts, cap))) return rc;
    if ((rc = dalloc(ctx, &d_keys, (size_t)cap * WW))) return rc;
    CK(cudaMemsetAsync(d_tags, 0, (size_t)cap * 4, ctx->stream));
    CK(cudaMemsetAsync(d_counts, 0, (size_t)cap * 4, ctx->stream));
    uint64_t solid_cap = std::max<uint64_t>(I_in <= (1ULL << 27) ? I_in : std::max<uint64_t>(I_in / 4, 1ULL << 27), 1);
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = dalloc(ctx, &ctx->d_solid_kmers, solid_cap * WW))) return rc;
        if ((rc = dalloc(ctx, &ctx->d_solid_counts, solid_cap))) return rc;
        if ((rc = ctr_zero(ctx))) return rc;
        tstart(ctx, T_K2);
        for (uint32_t b = 0; b < P_local; b++) {
            if (!bin_tot[b]) continue;
            const uint32_t cap_b = table_cap_for(bin_tot[b]);
            for (int s = 0; s < W; s++) {
                const uint64_t nb = run_len[(size_t)b * W + s];
                if (!nb) continue;
                LAUNCH((kw2_insert<WW>), grid_for(ctx, nb, 256, 8), 256, 0, d_kmers + run_off[(size_t)b * W + s] * WW, nb, d_tags, d_keys, d_counts, cap_b, ctx->d_flag);
            }
            LAUNCH((kw2_emit<WW>), grid_for(ctx, cap_b, 256, 8), 256, 0, d_tags, d_keys, d_counts, cap_b, ctx->amin, ctx->amax, ctx->d_solid_kmers,
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
    if (ctx->S >= (1ULL << 31)) return ctx->fail(CDBG_ERR_PARAM, "more than 2^31 solid k-mers on one device: shard over more GPUs");
    dfree(ctx, d_kmers); dfree(ctx, d_tags); dfree(ctx, d_counts); dfree(ctx, d_keys);
    dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
    tstop(ctx, T_COUNT);
    ctx->stage = ST_COUNTED;
    return CDBG_OK;
}

// K3 + X2 + K5 for multi-word k-mers; same contract as shard_join_narrow
template <int WW>
int shard_join_wide(cdbg_ctx *ctx, uint64_t S_total, ulonglong2 **out_msgs, uint64_t *out_n_msgs) {
    const int W = ctx->world, me = ctx->rank, k = ctx->k;
    const uint64_t S = ctx->S, n2 = 2 * S;
    int rc;
    const uint32_t P_local = (uint32_t)std::min<uint64_t>(2048 / W, std::max<uint64_t>(1, (2 * S_total / W + 499999) / 500000));
    const uint32_t P = P_local * W;
    ctx->st.n_join_bins = P_local;
    unsigned long long *d_bin_count = nullptr, *d_bin_cursor = nullptr;
    uint64_t *d_bin_off = nullptr, *d_key_s = nullptr, *d_val_s = nullptr;
    if ((rc = dalloc(ctx, &d_bin_count, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_cursor, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_off, P + 1))) return rc;
    if ((rc = dalloc(ctx, &d_key_s, n2 * WW + 1))) return rc;
    if ((rc = dalloc(ctx, &d_val_s, n2 + 1))) return rc;
    CK(cudaMemsetAsync(d_bin_count, 0, P * 8, ctx->stream));
    CK(cudaMemsetAsync(d_bin_cursor, 0, P * 8, ctx->stream));
    KW3Params p3{};
    p3.solid_kmers = ctx->d_solid_kmers; p3.n_solid = S; p3.k = k; p3.P = P;
    p3.bin_count = d_bin_count; p3.bin_off = d_bin_off; p3.bin_cursor = d_bin_cursor; p3.rec_key = d_key_s; p3.rec_val = d_val_s;
    p3.sharded = 1; p3.rank = (uint32_t)me;
    std::vector<uint64_t> h_cnt(P, 0), h_off(P + 1, 0);
    tstart(ctx, T_K3);
    CK(cudaFuncSetAttribute(kw3_ends<WW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3_smem_bytes(1, P)));
    const int g3 = (int)std::max<uint64_t>(1, std::min<uint64_t>((S + K3_THREADS - 1) / K3_THREADS, (uint64_t)ctx->n_sm * 8));
    if (S) {
        LAUNCH((kw3_ends<WW, 0>), g3, K3_THREADS, k3_smem_bytes(0, P), p3);
        SYNC_STREAM(ctx);   // (a copy to pageable memory blocks inside the call: wait by polling first, see stream_sync)
        CK(cudaMemcpyAsync(h_cnt.data(), d_bin_count, P * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SYNC_STREAM(ctx);
    }
    uint64_t acc = 0;
    for (uint32_t b = 0; b < P; b++) { h_off[b] = acc; acc += h_cnt[b]; }
    h_off[P] = acc;
    std::vector<uint64_t> all_cnt((size_t)W * P);
    if ((rc = allgather_host_u64(ctx, h_cnt.data(), P, all_cnt.data()))) return rc;
    if (S) {
        CK(cudaMemcpyAsync(d_bin_off, h_off.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH((kw3_ends<WW, 1>), g3, K3_THREADS, k3_smem_bytes(1, P), p3);
    }
    tstop(ctx, T_K3);
    // X2: keys (WW words per record) and values to the owner of the junction
    uint64_t send_cnt[MAX_WORLD] = {0}, recv_cnt[MAX_WORLD] = {0};
    std::vector<uint64_t> run_off((size_t)P_local * W), run_len((size_t)P_local * W);
    plan_shard_exchange(W, me, P_local, all_cnt.data(), send_cnt, recv_cnt, run_off.data(), run_len.data());
    void *rk = nullptr, *rv = nullptr; uint64_t T_in = 0, T_in2 = 0;
    if ((rc = exchange_with_counts(ctx, d_key_s, send_cnt, recv_cnt, 8 * WW, &rk, &T_in))) return rc;
    if ((rc = exchange_with_counts(ctx, d_val_s, send_cnt, recv_cnt, 8, &rv, &T_in2))) return rc;
    uint64_t *d_rec_key = (uint64_t *)rk, *d_rec_val = (uint64_t *)rv;
    dfree(ctx, d_key_s); dfree(ctx, d_val_s);
    ctx->st.n_junction_records = T_in;
    TRACE("k3 + X2 record exchange (wide)");

    // K5: per local bin, run by run; the slot state is a JoinTable64, so k5s_resolve (shard.cuh) turns it into messages + edges
    uint64_t max_bin = 0;
    std::vector<uint64_t> bin_tot(P_local, 0);
    for (uint32_t b = 0; b < P_local; b++) { for (int s = 0; s < W; s++) bin_tot[b] += run_len[(size_t)b * W + s]; max_bin = std::max(max_bin, bin_tot[b]); }
    JoinTable64 T{};
    const uint32_t cap = table_cap_for(max_bin);
    uint32_t *d_tags = nullptr; uint64_t *d_wkeys = nullptr;
    if ((rc = dalloc(ctx, &d_tags, cap))) return rc;
    if ((rc = dalloc(ctx, &d_wkeys, (size_t)cap * WW))) return rc;
    if ((rc = dalloc(ctx, &T.keys, cap))) return rc;
    if ((rc = dalloc(ctx, &T.cnt, cap))) return rc;
    if ((rc = dalloc(ctx, &T.ids, (size_t)cap * 8))) return rc;
    CK(cudaMemsetAsync(d_tags, 0, (size_t)cap * 4, ctx->stream));
    CK(cudaMemsetAsync(T.keys, 0xFF, (size_t)cap * 8, ctx->stream));
    CK(cudaMemsetAsync(T.cnt, 0, (size_t)cap * 4, ctx->stream));
    ulonglong2 *d_link_msgs = nullptr;
    const uint64_t link_cap = T_in + 16;
    ctx->sedge_cap = std::max<uint64_t>(T_in / 2, 1 << 16) + 4096;
    if ((rc = dalloc(ctx, &d_link_msgs, link_cap))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_sedges, ctx->sedge_cap))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_link64, n2 + 1))) return rc;
    CK(cudaMemsetAsync(ctx->d_link64, 0xFF, (n2 + 1) * 8, ctx->stream));
    if ((rc = ctr_zero(ctx))) return rc;
    tstart(ctx, T_K5);
    for (uint32_t b = 0; b < P_local; b++) {
        if (!bin_tot[b]) continue;
        T.cap = table_cap_for(bin_tot[b]);
        for (int s = 0; s < W; s++) {
            const uint64_t nb = run_len[(size_t)b * W + s], ro = run_off[(size_t)b * W + s];
            if (!nb) continue;
            LAUNCH((kw5s_insert<WW>), grid_for(ctx, nb, 256, 8), 256, 0, d_rec_key + ro * WW, d_rec_val + ro, nb, d_tags, d_wkeys, T, ctx->d_flag);
        }
        LAUNCH(k5s_resolve, grid_for(ctx, T.cap, 256, 8), 256, 0, T, T.cap, (uint32_t)me, ctx->d_link64, d_link_msgs, link_cap, ctx->d_ctr + 0,
               ctx->d_sedges, ctx->sedge_cap, ctx->d_ctr + 1, ctx->d_ctr + 2);
        CK(cudaMemsetAsync(d_tags, 0, (size_t)T.cap * 4, ctx->stream));      // k5s_resolve clears T.keys / T.cnt; the tags are this path's own
    }
    tstop(ctx, T_K5);
    if ((rc = ctr_read(ctx))) return rc;
    if (((int *)(ctx->h_ctr + 64))[0]) return ctx->fail(CDBG_ERR_INTERNAL, "K5 hash table overflow");
    const uint64_t n_link_msgs = ctx->h_ctr[0];
    ctx->n_sedges = ctx->h_ctr[1];
    if (n_link_msgs > link_cap) return ctx->fail(CDBG_ERR_INTERNAL, "link message buffer overflow");
    if (ctx->n_sedges + 2048 > ctx->sedge_cap) return ctx->fail(CDBG_ERR_INTERNAL, "edge buffer overflow (%llu edges)", (unsigned long long)ctx->n_sedges);
    ctx->st.n_compactions = ctx->h_ctr[2];
    dfree(ctx, d_rec_key); dfree(ctx, d_rec_val); dfree(ctx, d_tags); dfree(ctx, d_wkeys); dfree(ctx, T.keys); dfree(ctx, T.cnt); dfree(ctx, T.ids);
    dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
    *out_msgs = d_link_msgs; *out_n_msgs = n_link_msgs;
    return CDBG_OK;
}

int shard_stage_count_wide_dispatch(cdbg_ctx *ctx) {
    switch (ctx->W) { case 2: return shard_stage_count_wide<2>(ctx); case 3: return shard_stage_count_wide<3>(ctx); case 4: return shard_stage_count_wide<4>(ctx); }
    return ctx->fail(CDBG_ERR_INTERNAL, "bad word count %d", ctx->W);
}
int shard_join_wide_dispatch(cdbg_ctx *ctx, uint64_t S_total, ulonglong2 **out_msgs, uint64_t *out_n_msgs) {
    switch (ctx->W) {
        case 2: return shard_join_wide<2>(ctx, S_total, out_msgs, out_n_msgs);
        case 3: return shard_join_wide<3>(ctx, S_total, out_msgs, out_n_msgs);
        case 4: return shard_join_wide<4>(ctx, S_total, out_msgs, out_n_msgs);
    }
    return ctx->fail(CDBG_ERR_INTERNAL, "bad word count %d", ctx->W);
}
