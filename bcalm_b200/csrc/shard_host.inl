// shard_host.inl -- host orchestration of the minimizer-sharded multi-GPU path (included by cdbg.cu inside its anonymous
// namespace).  One context per GPU / process; every shard_* function is a COLLECTIVE: all ranks call it in the same order.
//
// Exchanges (NCCL grouped ncclSend/ncclRecv over NVLink, counts first via a small ncclAllGather):
//   X1  packed canonical k-mers            -> owner(minimizer_k)            (after K1, before counting)
//   X2  junction end records (key, value)  -> owner(minimizer_{k-1})        (after K3, before the join)
//   X3  compaction link messages           -> owner of each k-mer end       (after K5)
//   AG  piece descriptors + piece strings  -> every GPU                     (glue: replicated gid->terminus table + piece-level ranking)
//   X4  resolved link records              -> owner(from unitig)            (links; both ends resolved locally through the table)

#define NK(call)                                                                                                       \
    do {                                                                                                               \
        ncclResult_t r_ = (call);                                                                                      \
        if (r_ != ncclSuccess)                                                                                         \
            return ctx->fail(CDBG_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r_));      \
    } while (0)

// all-gather `count` u64 host values from every rank: h_out[r * count + i]
int allgather_host_u64(cdbg_ctx *ctx, const uint64_t *h_in, size_t count, uint64_t *h_out) {
    int rc;
    uint64_t *d_in = nullptr, *d_out = nullptr;
    if ((rc = dalloc(ctx, &d_in, count))) return rc;
    if ((rc = dalloc(ctx, &d_out, count * ctx->world))) return rc;
    CK(cudaMemcpyAsync(d_in, h_in, count * 8, cudaMemcpyHostToDevice, ctx->stream));
    NK(ncclAllGather(d_in, d_out, count, ncclUint64, ctx->comm, ctx->stream));
    SYNC_STREAM(ctx);   // (a copy to pageable memory blocks inside the call: wait by polling first, see stream_sync)
    CK(cudaMemcpyAsync(h_out, d_out, count * 8 * ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
    SYNC_STREAM(ctx);
    dfree(ctx, d_in); dfree(ctx, d_out);
    return CDBG_OK;
}

// d_send holds the elements for rank 0, then rank 1, ...; send_cnt / recv_cnt are element counts per peer
int exchange_with_counts(cdbg_ctx *ctx, const void *d_send, const uint64_t *send_cnt, const uint64_t *recv_cnt, size_t esz, void **d_recv,
                         uint64_t *n_recv) {
    int rc;
    uint64_t tot = 0;
    for (int s = 0; s < ctx->world; s++) tot += recv_cnt[s];
    uint8_t *recv = nullptr;
    if ((rc = dalloc(ctx, &recv, tot * esz + 16))) return rc;
    // this rank's own segment (with minimizer ownership: most of the end records) is a device copy, not a trip through NCCL
    {
        uint64_t so = 0, ro = 0;
        for (int peer = 0; peer < ctx->rank; peer++) { so += send_cnt[peer]; ro += recv_cnt[peer]; }
        if (send_cnt[ctx->rank])
            CK(cudaMemcpyAsync(recv + ro * esz, (const uint8_t *)d_send + so * esz, send_cnt[ctx->rank] * esz, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    NK(ncclGroupStart());
    uint64_t so = 0, ro = 0;
    for (int peer = 0; peer < ctx->world; peer++) {
        if (peer != ctx->rank) {
            if (send_cnt[peer]) NK(ncclSend((const uint8_t *)d_send + so * esz, send_cnt[peer] * esz, ncclChar, peer, ctx->comm, ctx->stream));
            if (recv_cnt[peer]) NK(ncclRecv(recv + ro * esz, recv_cnt[peer] * esz, ncclChar, peer, ctx->comm, ctx->stream));
        }
        so += send_cnt[peer]; ro += recv_cnt[peer];
    }
    NK(ncclGroupEnd());
    *d_recv = recv; *n_recv = tot;
    return CDBG_OK;
}

// route 16-byte records to the rank named in bits 39..32 of rec.x; returns the received records
int route16(cdbg_ctx *ctx, const ulonglong2 *d_recs, uint64_t n, ulonglong2 **d_out, uint64_t *n_out) {
    int rc;
    const int W = ctx->world;
    unsigned long long *d_cnt = nullptr, *d_cur = nullptr; uint64_t *d_off = nullptr, *d_all = nullptr; ulonglong2 *d_sorted = nullptr;
    if ((rc = dalloc(ctx, &d_cnt, MAX_WORLD))) return rc;
    if ((rc = dalloc(ctx, &d_cur, MAX_WORLD))) return rc;
    if ((rc = dalloc(ctx, &d_off, MAX_WORLD + 1))) return rc;
    if ((rc = dalloc(ctx, &d_all, (size_t)MAX_WORLD * MAX_WORLD))) return rc;
    if ((rc = dalloc(ctx, &d_sorted, n + 1))) return rc;
    CK(cudaMemsetAsync(d_cnt, 0, MAX_WORLD * 8, ctx->stream));
    CK(cudaMemsetAsync(d_cur, 0, MAX_WORLD * 8, ctx->stream));
    // the per-destination counts never visit the host on their own: histogram -> all-gather of the device counters -> ONE read of
    // the whole world x world matrix (this rank's row included)
    if (n) LAUNCH(route_hist, grid_for(ctx, n, 256, 4), 256, 0, d_recs, n, (uint32_t)W, d_cnt);
    NK(ncclAllGather(d_cnt, d_all, W, ncclUint64, ctx->comm, ctx->stream));
    std::vector<uint64_t> M((size_t)W * W);
    SYNC_STREAM(ctx);   // (a copy to pageable memory blocks inside the call: wait by polling first, see stream_sync)
    CK(cudaMemcpyAsync(M.data(), d_all, (size_t)W * W * 8, cudaMemcpyDeviceToHost, ctx->stream));
    SYNC_STREAM(ctx);
    uint64_t h_cnt[MAX_WORLD] = {0}, h_off[MAX_WORLD + 1] = {0};
    for (int d = 0; d < W; d++) { h_cnt[d] = M[(size_t)ctx->rank * W + d]; h_off[d + 1] = h_off[d] + h_cnt[d]; }
    if (h_off[W] != n) return ctx->fail(CDBG_ERR_INTERNAL, "route16: a record names a rank outside the world");
    if (n) {
        CK(cudaMemcpyAsync(d_off, h_off, (W + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(route_scatter, grid_for(ctx, n, 256, 4), 256, 0, d_recs, n, (uint32_t)W, d_off, d_cur, d_sorted);
    }
    uint64_t recv_cnt[MAX_WORLD];
    for (int s = 0; s < W; s++) recv_cnt[s] = M[(size_t)s * W + ctx->rank];
    void *recv = nullptr;
    if ((rc = exchange_with_counts(ctx, d_sorted, h_cnt, recv_cnt, 16, &recv, n_out))) return rc;
    *d_out = (ulonglong2 *)recv;
    dfree(ctx, d_cnt); dfree(ctx, d_cur); dfree(ctx, d_off); dfree(ctx, d_all); dfree(ctx, d_sorted);
    return CDBG_OK;
}

// Pure planning of a destination-major exchange (no CUDA, no NCCL; exported as cdbg_plan_shard_exchange and exercised by the
// 2-process gloo test on CPU).  all_cnt[s * P + d * P_local + b] = number of elements rank s holds for (destination d, local bin b),
// P = world * P_local.  Outputs for rank `me`: elements to send to / receive from every peer, and for each local bin the runs
// (offset, length) of the receive buffer that make it up, source-major: run_off[b * world + s], run_len[b * world + s].
inline void plan_shard_exchange(uint32_t world, uint32_t me, uint32_t P_local, const uint64_t *all_cnt, uint64_t *send_cnt, uint64_t *recv_cnt,
                                uint64_t *run_off, uint64_t *run_len) {
    const uint64_t P = (uint64_t)world * P_local;
    std::vector<uint64_t> recv_base(world + 1, 0), within(world, 0);
    for (uint32_t d = 0; d < world; d++) { send_cnt[d] = 0; for (uint32_t b = 0; b < P_local; b++) send_cnt[d] += all_cnt[(size_t)me * P + (size_t)d * P_local + b]; }
    for (uint32_t s = 0; s < world; s++) { recv_cnt[s] = 0; for (uint32_t b = 0; b < P_local; b++) recv_cnt[s] += all_cnt[(size_t)s * P + (size_t)me * P_local + b]; }
    for (uint32_t s = 0; s < world; s++) recv_base[s + 1] = recv_base[s] + recv_cnt[s];
    for (uint32_t b = 0; b < P_local; b++)
        for (uint32_t s = 0; s < world; s++) {
            uint64_t len = all_cnt[(size_t)s * P + (size_t)me * P_local + b];
            run_off[(size_t)b * world + s] = recv_base[s] + within[s];
            run_len[(size_t)b * world + s] = len;
            within[s] += len;
        }
}

// make room for `need` unitig-level edges in ctx->d_sedges (the edges of cut cycles are appended after the join has sized the buffer)
int grow_sedges(cdbg_ctx *ctx, uint64_t need) {
    if (need <= ctx->sedge_cap) return CDBG_OK;
    int rc;
    const uint64_t cap = need + need / 8 + 1024;
    ulonglong2 *ne = nullptr;
    if ((rc = dalloc(ctx, &ne, cap))) return rc;
    if (ctx->n_sedges) CK(cudaMemcpyAsync(ne, ctx->d_sedges, ctx->n_sedges * sizeof(ulonglong2), cudaMemcpyDeviceToDevice, ctx->stream));
    dfree(ctx, ctx->d_sedges);
    ctx->d_sedges = ne; ctx->sedge_cap = cap;
    return CDBG_OK;
}

inline uint32_t table_cap_for(uint64_t n) { return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(1024, (n * 3 / 2 + 63) & ~63ULL), 0xFFFFFF00u); }

// ------------------------------------------------------------------------------------------------ stage 1 (sharded)

// Stage 1 on super-k-mers (sk.cuh): K1 cuts this rank's reads into super-k-mer records in destination-major coarse bins
// (owner = bin / P_local = owner_of_rank(minimizer rank)), the records -- not 8-byte k-mer instances -- cross NVLink (X1), and the
// owner runs the single-GPU fine partition + shared-memory counting over the received (local bin, source) runs.
int shard_stage_count_sk(cdbg_ctx *ctx) {
    const int W = ctx->world, me = ctx->rank;
    int rc;
    tstart(ctx, T_COUNT);
    ctx->trace_t0 = host_now();
    uint64_t all_n[MAX_WORLD], total_bytes = 0;
    const uint64_t n = ctx->n_bytes;
    if ((rc = allgather_host_u64(ctx, &n, 1, all_n))) return rc;
    for (int r = 0; r < W; r++) total_bytes += all_n[r];
    const SKPlan pl = sk_plan(ctx, total_bytes, (uint32_t)W);
    const uint32_t P_local = pl.Pc_local, Pc = P_local * (uint32_t)W;
    ctx->st.n_count_bins = P_local;
    ctx->st.sk_m = (uint32_t)pl.m;
    ctx->m = pl.m;      // junctions are owned through the same minimizer order and length: ~1 - 2/(k-m+1) of a k-mer's end records stay local
    ctx->st.sk_waves = pl.V;
    for (uint32_t wave = 0; wave < pl.V; wave++) {
    std::vector<uint64_t> h_cnt, h_base;
    ulonglong2 *d_recs = nullptr;
    if ((rc = sk_run_k1(ctx, pl, Pc, wave, &d_recs, h_cnt, h_base))) return rc;
    TRACE("SK K1");
    // (test hook: tests/test_gpu_sharded.py makes one rank fail here to check that the others do not wait for it for ever)
    if (getenv("CDBG_FAULT_RANK") && atoi(getenv("CDBG_FAULT_RANK")) == me) return ctx->fail(CDBG_ERR_INTERNAL, "injected fault on rank %d (CDBG_FAULT_RANK)", me);
    // counts of every rank (+ its number of k-mer instances, which sizes the solid buffers)
    std::vector<uint64_t> mine(Pc + 1), all((size_t)W * (Pc + 1)), all_cnt((size_t)W * Pc);
    for (uint32_t b = 0; b < Pc; b++) mine[b] = h_cnt[b];
    mine[Pc] = ctx->st.n_instances;
    if ((rc = allgather_host_u64(ctx, mine.data(), Pc + 1, all.data()))) return rc;
    uint64_t I_total = 0;
    for (int r = 0; r < W; r++) {
        for (uint32_t b = 0; b < Pc; b++) all_cnt[(size_t)r * Pc + b] = all[(size_t)r * (Pc + 1) + b];
        I_total += all[(size_t)r * (Pc + 1) + Pc];
    }
    // the bins' regions have gaps: gather them into one contiguous send buffer (destination-major)
    tstart(ctx, T_X1);
    std::vector<uint64_t> h_out(Pc + 1, 0);
    for (uint32_t b = 0; b < Pc; b++) h_out[b + 1] = h_out[b] + h_cnt[b];
    const uint64_t R_out = h_out[Pc];
    ulonglong2 *d_send = nullptr; uint64_t *d_base = nullptr, *d_out_off = nullptr;
    if ((rc = dalloc(ctx, &d_send, R_out + 1))) return rc;
    if ((rc = dalloc(ctx, &d_base, Pc + 1))) return rc;
    if ((rc = dalloc(ctx, &d_out_off, Pc + 1))) return rc;
    CK(cudaMemcpyAsync(d_base, h_base.data(), (size_t)(Pc + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_out_off, h_out.data(), (size_t)(Pc + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (R_out) LAUNCH(sk_compact, grid_for(ctx, R_out, 256, 8), 256, 0, d_recs, d_base, d_out_off, Pc, d_send);
    dfree(ctx, d_recs); dfree(ctx, d_base); dfree(ctx, d_out_off);
    // X1: records to the owner of their minimizer
    uint64_t send_cnt[MAX_WORLD] = {0}, recv_cnt[MAX_WORLD] = {0};
    std::vector<uint64_t> run_off((size_t)P_local * W), run_len((size_t)P_local * W);
    plan_shard_exchange(W, me, P_local, all_cnt.data(), send_cnt, recv_cnt, run_off.data(), run_len.data());
    void *recv = nullptr; uint64_t R_in = 0;
    if ((rc = exchange_with_counts(ctx, d_send, send_cnt, recv_cnt, 16, &recv, &R_in))) return rc;
    dfree(ctx, d_send);
    tstop(ctx, T_X1);
    TRACE("X1 super-k-mer exchange");
    const uint32_t R = P_local * (uint32_t)W;
    std::vector<uint64_t> r_cnt(R), r_off(R + 1, 0);
    std::vector<uint32_t> r_group(R);
    for (uint32_t b = 0; b < P_local; b++)
        for (int s_ = 0; s_ < W; s_++) {
            const uint32_t r = b * (uint32_t)W + (uint32_t)s_;
            r_cnt[r] = run_len[(size_t)b * W + s_]; r_off[r] = run_off[(size_t)b * W + s_]; r_group[r] = b;
        }
    const uint64_t I_est = I_total / W + I_total / (4 * W) + 1024;
    if ((rc = sk_partition_count(ctx, pl, (const ulonglong2 *)recv, I_est, R, r_cnt, r_off, &r_group, P_local, wave))) return rc;
    dfree(ctx, recv);
    }
    tstop(ctx, T_COUNT);
    ctx->stage = ST_COUNTED;
    return CDBG_OK;
}

int shard_stage_count(cdbg_ctx *ctx) {
    const int W = ctx->world, me = ctx->rank, k = ctx->k;
    const uint64_t n = ctx->n_bytes;
    int rc;
    tstart(ctx, T_COUNT);
    ctx->trace_t0 = host_now();
    // global input size -> the same number of local bins on every rank
    uint64_t all_n[MAX_WORLD], total_bytes = 0;
    if ((rc = allgather_host_u64(ctx, &n, 1, all_n))) return rc;
    for (int r = 0; r < W; r++) total_bytes += all_n[r];
    uint32_t P_local = (uint32_t)std::min<uint64_t>(K1_MAX_BINS / W, std::max<uint64_t>(1, (total_bytes / W * 4 / 5 + 4999999) / 5000000));
    const uint32_t P = P_local * W;
    ctx->st.n_count_bins = P_local;
    // K1 variant of the sharded path (CDBG_SHARD_K1): default = k1s_extract (lane-consecutive 2048-position tiles, non-rolling);
    // "roll" = k1t_shard_hist + k1t_extract<1> (TMA-fed 8192-position tiles, rolling k-mers; emulator-verified only so far)
    const bool k1_roll = getenv("CDBG_SHARD_K1") && strcmp(getenv("CDBG_SHARD_K1"), "roll") == 0;
    const uint64_t tile_sz = k1_roll ? (uint64_t)K1_TILE : (uint64_t)K1S_TILE;
    uint64_t n_tiles = (n + tile_sz - 1) / tile_sz;

    unsigned long long *d_bin_count = nullptr, *d_bin_cursor = nullptr; uint64_t *d_bin_off = nullptr;
    if ((rc = dalloc(ctx, &d_bin_count, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_cursor, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_off, P + 1))) return rc;
    CK(cudaMemsetAsync(d_bin_count, 0, P * 8, ctx->stream));
    CK(cudaMemsetAsync(d_bin_cursor, 0, P * 8, ctx->stream));
    uint16_t *d_pos_bin = nullptr;
    if ((rc = dalloc(ctx, &d_pos_bin, n_tiles * tile_sz + 8))) return rc;
    K1SParams kp{};
    kp.pos_bin = d_pos_bin;
    kp.bytes = ctx->d_bytes; kp.n_bytes = n; kp.n_tiles = n_tiles; kp.k = k; kp.m = ctx->m; kp.world = W; kp.P_local = P_local;
    kp.bin_count = d_bin_count; kp.bin_off = d_bin_off; kp.bin_cursor = d_bin_cursor; kp.out_kmers = nullptr;
    K1TSParams ks{};
    ks.bytes = ctx->d_bytes; ks.n_bytes = n; ks.n_tiles = n_tiles; ks.k = k; ks.m = ctx->m; ks.world = W; ks.P_local = P_local;
    ks.bin_count = d_bin_count; ks.pos_bin = d_pos_bin;
    K1TParams kt{};
    kt.bytes = ctx->d_bytes; kt.n_bytes = n; kt.n_tiles = n_tiles; kt.k = k; kt.P = P;
    kt.bin_count = d_bin_count; kt.bin_off = d_bin_off; kt.bin_cursor = d_bin_cursor; kt.out_kmers = nullptr; kt.pos_bin = d_pos_bin;
    std::vector<uint64_t> h_cnt(P, 0), h_off(P + 1, 0);
    if (k1_roll) {
        CK(cudaFuncSetAttribute(k1t_shard_hist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1ts_smem_bytes(P)));
        CK(cudaFuncSetAttribute(k1t_extract<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1t_smem_bytes(1, P)));
    } else {
        CK(cudaFuncSetAttribute(k1s_extract<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1s_smem_bytes(0, P)));
        CK(cudaFuncSetAttribute(k1s_extract<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1s_smem_bytes(1, P)));
    }
    if (n_tiles) {
        tstart(ctx, T_K1H);
        if (k1_roll) LAUNCH(k1t_shard_hist, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 2), K1_THREADS, k1ts_smem_bytes(P), ks);
        else LAUNCH(k1s_extract<0>, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 4), K1S_THREADS, k1s_smem_bytes(0, P), kp);
        tstop(ctx, T_K1H);
        SYNC_STREAM(ctx);   // (a copy to pageable memory blocks inside the call: wait by polling first, see stream_sync)
        CK(cudaMemcpyAsync(h_cnt.data(), d_bin_count, P * 8, cudaMemcpyDeviceToHost, ctx->stream));
        SYNC_STREAM(ctx);
    }
    TRACE("k1 hist");
    uint64_t I_out = 0;
    for (uint32_t b = 0; b < P; b++) { h_off[b] = I_out; I_out += h_cnt[b]; }
    h_off[P] = I_out;
    ctx->st.n_instances = I_out;
    std::vector<uint64_t> all_cnt((size_t)W * P);
    if ((rc = allgather_host_u64(ctx, h_cnt.data(), P, all_cnt.data()))) return rc;

    uint64_t *d_send = nullptr;
    if ((rc = dalloc(ctx, &d_send, I_out + 1))) return rc;
    if (I_out) {
        CK(cudaMemcpyAsync(d_bin_off, h_off.data(), (P + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        kp.out_kmers = d_send; kt.out_kmers = d_send;
        tstart(ctx, T_K1S);
        if (k1_roll) LAUNCH(k1t_extract<1>, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 2), K1_THREADS, k1t_smem_bytes(1, P), kt);
        else LAUNCH(k1s_extract<1>, (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->n_sm * 2), K1S_THREADS, k1s_smem_bytes(1, P), kp);
        tstop(ctx, T_K1S);
    }
    TRACE("k1 scatter + allgather counts");
    dfree(ctx, d_pos_bin);
    // X1: k-mers to their counting owner
    uint64_t send_cnt[MAX_WORLD] = {0}, recv_cnt[MAX_WORLD] = {0};
    std::vector<uint64_t> run_off((size_t)P_local * W), run_len((size_t)P_local * W);
    plan_shard_exchange(W, me, P_local, all_cnt.data(), send_cnt, recv_cnt, run_off.data(), run_len.data());
    void *recv = nullptr; uint64_t I_in = 0;
    if ((rc = exchange_with_counts(ctx, d_send, send_cnt, recv_cnt, 8, &recv, &I_in))) return rc;
    uint64_t *d_kmers = (uint64_t *)recv;
    dfree(ctx, d_send);
    TRACE("  free send buffer");

    TRACE("X1 kmer exchange");
    // K2 with shared-memory tables (k2_smem.cuh), opt-in for the sharded path (CDBG_SHARD_K2=smem; its kernels are the single-GPU
    // default, this wiring has only run under the CPU emulator): every (local bin, source) run is a coarse bin, the runs of one local
    // bin share their fine bins
    if (getenv("CDBG_SHARD_K2") && strcmp(getenv("CDBG_SHARD_K2"), "smem") == 0) {
        K2SmemPlan pl = k2_smem_plan(I_in * 5 / 4);
        pl.enabled = true;
        const uint32_t R = P_local * (uint32_t)W;
        std::vector<uint64_t> r_cnt(R), r_off(R + 1, 0);
        std::vector<uint32_t> r_group(R);
        uint64_t worst = 0;
        for (uint32_t b = 0; b < P_local; b++)
            for (int s = 0; s < W; s++) {
                const uint32_t r = b * (uint32_t)W + (uint32_t)s;
                r_cnt[r] = run_len[(size_t)b * W + s]; r_off[r] = run_off[(size_t)b * W + s]; r_group[r] = b;
                worst = std::max(worst, r_cnt[r]);
            }
        if (worst < (1ULL << 32)) {
            if ((rc = stage_count_k2_smem(ctx, pl, d_kmers, I_in, R, r_cnt, r_off, &r_group, P_local))) return rc;
            dfree(ctx, d_kmers);
            dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
            tstop(ctx, T_COUNT);
            ctx->stage = ST_COUNTED;
            return CDBG_OK;
        }
    }
    // K2 per local bin; a bin is one run per source rank
    uint64_t max_bin = 0;
    std::vector<uint64_t> bin_tot(P_local, 0);
    for (uint32_t b = 0; b < P_local; b++) { for (int s = 0; s < W; s++) bin_tot[b] += all_cnt[(size_t)s * P + (size_t)me * P_local + b]; max_bin = std::max(max_bin, bin_tot[b]); }
    uint32_t cap = table_cap_for(max_bin);
    uint64_t *d_keys = nullptr; uint32_t *d_counts = nullptr;
    if ((rc = dalloc(ctx, &d_keys, cap))) return rc;
    if ((rc = dalloc(ctx, &d_counts, cap))) return rc;
    CK(cudaMemsetAsync(d_keys, 0xFF, (size_t)cap * 8, ctx->stream));
    CK(cudaMemsetAsync(d_counts, 0, (size_t)cap * 4, ctx->stream));
    TRACE("  table alloc + memset");
    uint64_t solid_cap = std::max<uint64_t>(I_in <= (1ULL << 28) ? I_in : std::max<uint64_t>(I_in / 4, 1ULL << 28), 1);
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = dalloc(ctx, &ctx->d_solid_kmers, solid_cap))) return rc;
        if ((rc = dalloc(ctx, &ctx->d_solid_counts, solid_cap))) return rc;
        if ((rc = ctr_zero(ctx))) return rc;
        TRACE("  solid alloc");
        tstart(ctx, T_K2);
        double slots_per_instance = 1.5;   // first bin worst case, then sized from its distinct ratio (see stage_count)
        bool first = true;
        for (uint32_t b = 0; b < P_local; b++) {
            Runs runs{}; runs.n = W; runs.pre[0] = 0;
            for (int s = 0; s < W; s++) { runs.off[s] = run_off[(size_t)b * W + s]; runs.pre[s + 1] = runs.pre[s] + run_len[(size_t)b * W + s]; }
            uint64_t nb = bin_tot[b];
            if (!nb) continue;
            uint32_t cap_b = std::min(cap, std::max<uint32_t>(1024, (uint32_t)std::min<double>((double)nb * slots_per_instance + 64.0, 4.0e9) & ~63u));
            LAUNCH(k2s_insert, grid_for(ctx, nb, 256, 8), 256, 0, d_kmers, runs, d_keys, d_counts, cap_b, ctx->d_flag);
            LAUNCH(k2_emit, grid_for(ctx, cap_b, 256, 8), 256, 0, d_keys, d_counts, cap_b, ctx->amin, ctx->amax, ctx->d_solid_kmers, ctx->d_solid_counts,
                   solid_cap, ctx->d_ctr + 0, ctx->d_ctr + 1);
            if (first && attempt == 0 && P_local > 2) {
                first = false;
                if ((rc = ctr_read(ctx))) return rc;
                slots_per_instance = std::min(1.5, std::max(0.02, 2.5 * (double)ctx->h_ctr[1] / (double)nb + 0.02));
            }
        }
        tstop(ctx, T_K2);
        if ((rc = ctr_read(ctx))) return rc;
        const bool overflow = ((int *)(ctx->h_ctr + 64))[0] != 0;
        if (overflow && attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "K2 hash table overflow");
        ctx->S = ctx->h_ctr[0];
        ctx->st.n_distinct = ctx->h_ctr[1];
        if (!overflow && ctx->S <= solid_cap) break;
        if (attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "solid buffer overflow");
        dfree(ctx, ctx->d_solid_kmers); dfree(ctx, ctx->d_solid_counts);
        if (!overflow) solid_cap = ctx->S;
        else {
            CK(cudaMemsetAsync(d_keys, 0xFF, (size_t)cap * 8, ctx->stream));
            CK(cudaMemsetAsync(d_counts, 0, (size_t)cap * 4, ctx->stream));
        }
    }
    ctx->st.n_solid = ctx->S;
    if (ctx->S >= (1ULL << 31)) return ctx->fail(CDBG_ERR_PARAM, "more than 2^31 solid k-mers on one device: shard over more GPUs");
    dfree(ctx, d_kmers); dfree(ctx, d_keys); dfree(ctx, d_counts);
    dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
    tstop(ctx, T_COUNT);
    ctx->stage = ST_COUNTED;
    return CDBG_OK;
}

// ------------------------------------------------------------------------------------------------ stages 2+3 (sharded)

// The sharded join with shared-memory tables (k5p_hist / k5p_scatter with groups, k5cs_join): the received end records (runs: one
// per local bin and source GPU) are split into fine bins of ~one table each; leaves ctx->d_link64, ctx->d_sedges / n_sedges and the
// link messages for ends on other GPUs.  Buffers that overflow are re-allocated with the exact size and the join is repeated.
int shard_join_smem(cdbg_ctx *ctx, const uint64_t *d_rec_key, const uint64_t *d_rec_val, uint64_t T_in, uint32_t P_local,
                    const std::vector<uint64_t> &run_off, const std::vector<uint64_t> &run_len, ulonglong2 **out_msgs, uint64_t *out_n_msgs) {
    const int W = ctx->world;
    const uint64_t n2 = 2 * ctx->S;
    int rc;
    uint32_t slots = getenv("CDBG_K5S_SLOTS") && atoi(getenv("CDBG_K5S_SLOTS")) == 4096 ? 4096u : 2048u;
    double fill = getenv("CDBG_K5S_FILL") ? atof(getenv("CDBG_K5S_FILL")) : 1.0;
    if (!(fill >= 0.05 && fill <= 8.0)) fill = 1.0;
    // more records than P_local x K5P_MAX_F small tables take: the larger table, before relying on in-kernel splits
    if (slots == 2048 && !getenv("CDBG_K5S_SLOTS") && (double)T_in / (fill * slots * P_local) > (double)K5P_MAX_F) slots = 4096;
    const uint32_t R = P_local * (uint32_t)W;
    const uint32_t F = (uint32_t)std::min<uint64_t>(K5P_MAX_F, std::max<uint64_t>(1, (uint64_t)ceil((double)T_in / (fill * slots * P_local))));
    const uint64_t NF = (uint64_t)P_local * F;
    std::vector<uint32_t> h_tile_first(R + 1), h_group(R);
    std::vector<uint64_t> h_range(2 * (size_t)R);
    uint64_t tiles = 0;
    for (uint32_t b = 0; b < P_local; b++)
        for (int s_ = 0; s_ < W; s_++) {
            const uint32_t r = b * (uint32_t)W + (uint32_t)s_;
            h_tile_first[r] = (uint32_t)tiles; tiles += (run_len[r] + K5P_TILE - 1) / K5P_TILE;
            h_range[r] = run_off[r]; h_range[R + r] = run_off[r] + run_len[r]; h_group[r] = b;
        }
    h_tile_first[R] = (uint32_t)tiles;
    if (tiles >= (1ULL << 32)) return ctx->fail(CDBG_ERR_PARAM, "too many end-record tiles for one device");
    uint32_t *d_tile_first = nullptr, *d_fine_cnt = nullptr, *d_fine_cursor = nullptr, *d_group = nullptr;
    uint64_t *d_fine_off = nullptr, *d_key2 = nullptr, *d_val2 = nullptr, *d_bin_range = nullptr;
    if ((rc = dalloc(ctx, &d_bin_range, 2 * (size_t)R))) return rc;
    if ((rc = dalloc(ctx, &d_tile_first, R + 1))) return rc;
    if ((rc = dalloc(ctx, &d_group, R))) return rc;
    if ((rc = dalloc(ctx, &d_fine_cnt, NF + 1))) return rc;
    if ((rc = dalloc(ctx, &d_fine_cursor, NF + 1))) return rc;
    if ((rc = dalloc(ctx, &d_fine_off, NF + 2))) return rc;
    if ((rc = dalloc(ctx, &d_key2, T_in + 1))) return rc;
    if ((rc = dalloc(ctx, &d_val2, T_in + 1))) return rc;
    CK(cudaMemsetAsync(d_fine_cnt, 0, (NF + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(d_fine_cursor, 0, (NF + 1) * 4, ctx->stream));
    CK(cudaMemcpyAsync(d_bin_range, h_range.data(), 2 * (size_t)R * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_tile_first, h_tile_first.data(), (size_t)(R + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_group, h_group.data(), (size_t)R * 4, cudaMemcpyHostToDevice, ctx->stream));
    tstart(ctx, T_K5);
    K5PParams pp{};
    pp.key = d_rec_key; pp.val = d_rec_val; pp.bin_beg = d_bin_range; pp.bin_end = d_bin_range + R; pp.tile_first = d_tile_first; pp.group = d_group;
    pp.P1 = R; pp.F = F; pp.n_tiles = (uint32_t)tiles; pp.G = P_local;
    pp.fine_cnt = d_fine_cnt; pp.fine_off = d_fine_off; pp.fine_cursor = d_fine_cursor; pp.out_key = d_key2; pp.out_val = d_val2;
    if (tiles) {
        LAUNCH(k5p_hist, (int)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * 8), K5P_THREADS, 0, pp);
        uint64_t total = 0;
        if ((rc = exclusive_scan(ctx, d_fine_cnt, NF + 1, d_fine_off, &total))) return rc;
        if (total != T_in) return ctx->fail(CDBG_ERR_INTERNAL, "fine histogram lost end records (%llu != %llu)", (unsigned long long)total, (unsigned long long)T_in);
        CK(cudaFuncSetAttribute(k5p_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5p_scatter_smem_bytes()));
        LAUNCH(k5p_scatter, (int)std::min<uint64_t>(tiles, (uint64_t)ctx->n_sm * 2), K5P_THREADS, k5p_scatter_smem_bytes(), pp);
    } else {
        CK(cudaMemsetAsync(d_fine_off, 0, (NF + 2) * 8, ctx->stream));
    }
    ulonglong2 *d_link_msgs = nullptr;
    uint64_t link_cap = T_in / 4 + 4096;
    ctx->sedge_cap = std::max<uint64_t>(T_in / 4, edge_min()) + 4096;
    if ((rc = dalloc(ctx, &ctx->d_link64, n2 + 1))) return rc;
    uint64_t n_link_msgs = 0;
    for (int attempt = 0; attempt < 2; attempt++) {
        if ((rc = dalloc(ctx, &d_link_msgs, link_cap))) return rc;
        if ((rc = dalloc(ctx, &ctx->d_sedges, ctx->sedge_cap))) return rc;
        CK(cudaMemsetAsync(ctx->d_link64, 0xFF, (n2 + 1) * 8, ctx->stream));
        if ((rc = ctr_zero(ctx))) return rc;
        K5CSParams cp{};
        cp.key = d_key2; cp.val = d_val2; cp.fine_off = d_fine_off; cp.n_fine = (uint32_t)NF; cp.rank = (uint32_t)ctx->rank;
        cp.link64 = ctx->d_link64; cp.link_msgs = d_link_msgs; cp.link_cap = link_cap; cp.n_link_msgs = ctx->d_ctr + 0;
        cp.edges = ctx->d_sedges; cp.edge_cap = ctx->sedge_cap; cp.n_edges = ctx->d_ctr + 1;
        cp.n_compact = ctx->d_ctr + 2; cp.n_splits = ctx->d_ctr + 3; cp.failed = ctx->d_flag + 1;
        if (T_in) {
            if (slots == 4096) {
                CK(cudaFuncSetAttribute(k5cs_join<4096>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5cs_smem_bytes<4096>()));
                LAUNCH(k5cs_join<4096>, (int)std::min<uint64_t>(NF, (uint64_t)ctx->n_sm), K5CS_THREADS, k5cs_smem_bytes<4096>(), cp);
            } else {
                CK(cudaFuncSetAttribute(k5cs_join<2048>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5cs_smem_bytes<2048>()));
                LAUNCH(k5cs_join<2048>, (int)std::min<uint64_t>(NF, (uint64_t)ctx->n_sm * 2), K5CS_THREADS, k5cs_smem_bytes<2048>(), cp);
            }
        }
        if (attempt == 0) tstop(ctx, T_K5);
        if ((rc = ctr_read(ctx))) return rc;
        if (((int *)(ctx->h_ctr + 64))[1]) return ctx->fail(CDBG_ERR_INTERNAL, "K5: a fine bin does not fit a shared-memory table after %d splits", K5CS_MAX_LEVEL);
        n_link_msgs = ctx->h_ctr[0];
        ctx->n_sedges = ctx->h_ctr[1];
        ctx->st.n_compactions = ctx->h_ctr[2];
        ctx->st.n_k5_splits = (uint32_t)ctx->h_ctr[3];
        // room is kept for the edges of locally cut cycles and of piece-level cycle cuts, which are appended later
        if (n_link_msgs <= link_cap && ctx->n_sedges <= ctx->sedge_cap) break;     // (edges of cut cycles are appended through grow_sedges)
        if (attempt == 1) return ctx->fail(CDBG_ERR_INTERNAL, "join output buffers overflow after resizing");
        dfree(ctx, d_link_msgs); dfree(ctx, ctx->d_sedges);
        link_cap = n_link_msgs + 4096;
        ctx->sedge_cap = ctx->n_sedges + std::max<uint64_t>(ctx->n_sedges / 8, 1 << 17) + 4096;
    }
    dfree(ctx, d_bin_range); dfree(ctx, d_tile_first); dfree(ctx, d_group); dfree(ctx, d_fine_cnt); dfree(ctx, d_fine_cursor); dfree(ctx, d_fine_off);
    dfree(ctx, d_key2); dfree(ctx, d_val2);
    *out_msgs = d_link_msgs; *out_n_msgs = n_link_msgs;
    return CDBG_OK;
}

// K3 + X2 + K5 of the sharded path for one-word k-mers: end records to the owner of their junction, the join there; leaves
// ctx->d_link64 (links whose end is local), the link messages for ends on other GPUs, and the unitig-level edges in ctx->d_sedges.
int shard_join_narrow(cdbg_ctx *ctx, uint64_t S_total, ulonglong2 **out_msgs, uint64_t *out_n_msgs) {
    const int W = ctx->world, me = ctx->rank, k = ctx->k;
    const uint64_t S = ctx->S, n2 = 2 * S;
    int rc;
    // ---- K3: end records, destination-major bins, X2
    uint32_t P_local = (uint32_t)std::min<uint64_t>(2048 / W, std::max<uint64_t>(1, (2 * S_total / W + 499999) / 500000));
    const uint32_t P = P_local * W;
    ctx->st.n_join_bins = P_local;
    unsigned long long *d_bin_count = nullptr, *d_bin_cursor = nullptr;
    uint64_t *d_bin_off = nullptr, *d_key_s = nullptr, *d_val_s = nullptr;
    if ((rc = dalloc(ctx, &d_bin_count, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_cursor, P))) return rc;
    if ((rc = dalloc(ctx, &d_bin_off, P + 1))) return rc;
    if ((rc = dalloc(ctx, &d_key_s, n2 + 1))) return rc;
    if ((rc = dalloc(ctx, &d_val_s, n2 + 1))) return rc;
    uint16_t *d_bins3 = nullptr;
    if ((rc = dalloc(ctx, &d_bins3, n2 + 2))) return rc;
    CK(cudaMemsetAsync(d_bin_count, 0, P * 8, ctx->stream));
    CK(cudaMemsetAsync(d_bin_cursor, 0, P * 8, ctx->stream));
    K3SParams p3{};
    p3.solid_kmers = ctx->d_solid_kmers; p3.n_solid = S; p3.k = k; p3.m = ctx->m; p3.rank = me; p3.world = W; p3.P_local = P_local;
    p3.bin_count = d_bin_count; p3.bin_off = d_bin_off; p3.bin_cursor = d_bin_cursor; p3.rec_key = d_key_s; p3.rec_val = d_val_s; p3.bins = d_bins3;
    std::vector<uint64_t> h_cnt(P, 0), h_off(P + 1, 0);
    tstart(ctx, T_K3);
    CK(cudaFuncSetAttribute(k3s_ends<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3_smem_bytes(1, P)));
    const int g3 = (int)std::max<uint64_t>(1, std::min<uint64_t>((S + K3_TILE - 1) / K3_TILE, (uint64_t)ctx->n_sm * 4));
    if (S) {
        LAUNCH(k3s_ends<0>, g3, K3_THREADS, k3_smem_bytes(0, P), p3);
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
        LAUNCH(k3s_ends<1>, g3, K3_THREADS, k3_smem_bytes(1, P), p3);
    }
    tstop(ctx, T_K3);
    uint64_t send_cnt[MAX_WORLD] = {0}, recv_cnt[MAX_WORLD] = {0};
    std::vector<uint64_t> run_off((size_t)P_local * W), run_len((size_t)P_local * W);
    plan_shard_exchange(W, me, P_local, all_cnt.data(), send_cnt, recv_cnt, run_off.data(), run_len.data());
    // keys and values travel as two segments of ONE exchange: [keys of all bins for peer d | values ...] would need a repack, so
    // instead both arrays are sent in the same NCCL group (one launch, one synchronisation point)
    uint64_t *d_rec_key = nullptr, *d_rec_val = nullptr; uint64_t T_in = 0;
    tstart(ctx, T_X2);
    {
        for (int sidx = 0; sidx < W; sidx++) T_in += recv_cnt[sidx];
        if ((rc = dalloc(ctx, &d_rec_key, T_in + 2))) return rc;
        if ((rc = dalloc(ctx, &d_rec_val, T_in + 2))) return rc;
        {   // this rank's own records (~90 % with minimizer ownership) are a device copy
            uint64_t so = 0, ro = 0;
            for (int peer = 0; peer < me; peer++) { so += send_cnt[peer]; ro += recv_cnt[peer]; }
            if (send_cnt[me]) {
                CK(cudaMemcpyAsync(d_rec_key + ro, d_key_s + so, send_cnt[me] * 8, cudaMemcpyDeviceToDevice, ctx->stream));
                CK(cudaMemcpyAsync(d_rec_val + ro, d_val_s + so, send_cnt[me] * 8, cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
        NK(ncclGroupStart());
        uint64_t so = 0, ro = 0;
        for (int peer = 0; peer < W; peer++) {
            if (peer != me && send_cnt[peer]) {
                NK(ncclSend(d_key_s + so, send_cnt[peer], ncclUint64, peer, ctx->comm, ctx->stream));
                NK(ncclSend(d_val_s + so, send_cnt[peer], ncclUint64, peer, ctx->comm, ctx->stream));
            }
            if (peer != me && recv_cnt[peer]) {
                NK(ncclRecv(d_rec_key + ro, recv_cnt[peer], ncclUint64, peer, ctx->comm, ctx->stream));
                NK(ncclRecv(d_rec_val + ro, recv_cnt[peer], ncclUint64, peer, ctx->comm, ctx->stream));
            }
            so += send_cnt[peer]; ro += recv_cnt[peer];
        }
        NK(ncclGroupEnd());
    }
    tstop(ctx, T_X2);
    dfree(ctx, d_key_s); dfree(ctx, d_val_s); dfree(ctx, d_bins3);
    ctx->st.n_junction_records = T_in;

    TRACE("k3 + X2 record exchange");
    // ---- K5 with shared-memory tables (default; CDBG_SHARD_K5=l2 keeps one L2-resident table per local bin)
    if (!(getenv("CDBG_SHARD_K5") && strcmp(getenv("CDBG_SHARD_K5"), "l2") == 0)) {
        uint64_t worst = 0;
        for (size_t i = 0; i < run_len.size(); i++) worst = std::max(worst, run_len[i]);
        if (worst < (1ULL << 32)) {
            if ((rc = shard_join_smem(ctx, d_rec_key, d_rec_val, T_in, P_local, run_off, run_len, out_msgs, out_n_msgs))) return rc;
            dfree(ctx, d_rec_key); dfree(ctx, d_rec_val);
            dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
            return CDBG_OK;
        }
    }
    // ---- K5: per-bin join -> link messages + edges (gids)
    uint64_t max_bin = 0;
    std::vector<uint64_t> bin_tot(P_local, 0);
    for (uint32_t b = 0; b < P_local; b++) { for (int s = 0; s < W; s++) bin_tot[b] += all_cnt[(size_t)s * P + (size_t)me * P_local + b]; max_bin = std::max(max_bin, bin_tot[b]); }
    JoinTable64 T{};
    uint32_t cap = table_cap_for(max_bin);
    if ((rc = dalloc(ctx, &T.keys, cap))) return rc;
    if ((rc = dalloc(ctx, &T.cnt, cap))) return rc;
    if ((rc = dalloc(ctx, &T.ids, (size_t)cap * 8))) return rc;
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
    {
        for (uint32_t b = 0; b < P_local; b++) {
            Runs runs{}; runs.n = W; runs.pre[0] = 0;
            for (int s = 0; s < W; s++) { runs.off[s] = run_off[(size_t)b * W + s]; runs.pre[s + 1] = runs.pre[s] + run_len[(size_t)b * W + s]; }
            uint64_t nb = bin_tot[b];
            if (!nb) continue;
            T.cap = table_cap_for(nb);
            LAUNCH(k5s_insert, grid_for(ctx, nb, 256, 8), 256, 0, d_rec_key, d_rec_val, runs, T, ctx->d_flag);
            LAUNCH(k5s_resolve, grid_for(ctx, T.cap, 256, 8), 256, 0, T, T.cap, (uint32_t)me, ctx->d_link64, d_link_msgs, link_cap, ctx->d_ctr + 0,
                   ctx->d_sedges, ctx->sedge_cap, ctx->d_ctr + 1, ctx->d_ctr + 2);
        }
    }
    tstop(ctx, T_K5);
    if ((rc = ctr_read(ctx))) return rc;
    if (((int *)(ctx->h_ctr + 64))[0]) return ctx->fail(CDBG_ERR_INTERNAL, "K5 hash table overflow");
    const uint64_t n_link_msgs = ctx->h_ctr[0];
    ctx->n_sedges = ctx->h_ctr[1];
    if (n_link_msgs > link_cap) return ctx->fail(CDBG_ERR_INTERNAL, "link message buffer overflow");
    if (ctx->n_sedges + 2048 > ctx->sedge_cap) return ctx->fail(CDBG_ERR_INTERNAL, "edge buffer overflow (%llu edges)", (unsigned long long)ctx->n_sedges);
    ctx->st.n_compactions = ctx->h_ctr[2];
    dfree(ctx, d_rec_key); dfree(ctx, d_rec_val); dfree(ctx, T.keys); dfree(ctx, T.cnt); dfree(ctx, T.ids);
    dfree(ctx, d_bin_count); dfree(ctx, d_bin_cursor); dfree(ctx, d_bin_off);
    *out_msgs = d_link_msgs; *out_n_msgs = n_link_msgs;
    return CDBG_OK;
}

int shard_join_wide_dispatch(cdbg_ctx *ctx, uint64_t S_total, ulonglong2 **out_msgs, uint64_t *out_n_msgs);   // shard_wide_host.inl

int shard_stage_compact(cdbg_ctx *ctx) {
    const int W = ctx->world, me = ctx->rank, k = ctx->k;
    const uint64_t S = ctx->S, n2 = 2 * S;
    int rc;
    tstart(ctx, T_COMPACT);
    uint64_t all_S[MAX_WORLD], S_total = 0;
    if ((rc = allgather_host_u64(ctx, &S, 1, all_S))) return rc;
    for (int r = 0; r < W; r++) S_total += all_S[r];

    ulonglong2 *d_link_msgs = nullptr; uint64_t n_link_msgs = 0;
    if ((rc = ctx->W > 1 ? shard_join_wide_dispatch(ctx, S_total, &d_link_msgs, &n_link_msgs) : shard_join_narrow(ctx, S_total, &d_link_msgs, &n_link_msgs))) return rc;

    TRACE("k5 join");
    // ---- X3: link messages to the owners of the ends; local view of the links
    ulonglong2 *d_msgs_in = nullptr; uint64_t n_msgs_in = 0;
    tstart(ctx, T_X3);
    if ((rc = route16(ctx, d_link_msgs, n_link_msgs, &d_msgs_in, &n_msgs_in))) return rc;
    tstop(ctx, T_X3);
    dfree(ctx, d_link_msgs);
    if ((rc = dalloc(ctx, &ctx->d_link, n2 + 1))) return rc;
    if (n_msgs_in) LAUNCH(shard_apply_links, grid_for(ctx, n_msgs_in, 256, 8), 256, 0, d_msgs_in, n_msgs_in, ctx->d_link64);
    dfree(ctx, d_msgs_in);
    CK(cudaMemsetAsync(ctx->d_link, 0xFF, (n2 + 1) * 4, ctx->stream));
    if (n2) LAUNCH(shard_local_links, grid_for(ctx, n2, 256, 8), 256, 0, ctx->d_link64, n2, (uint32_t)me, ctx->d_link);

    TRACE("X3 route link msgs + apply");
    // ---- stage 2 proper: local ranking -> pieces (the unitigs of this shard), spelled locally
    uint64_t *d_pair = nullptr, *d_uid_scan = nullptr, *d_nt_scan = nullptr;
    ctx->edge_cap = edge_min();
    if ((rc = dalloc(ctx, &ctx->d_edges, ctx->edge_cap))) return rc;
    ctx->n_edges = 0;
    tstart(ctx, T_K6R);
    if ((rc = run_ranking(ctx, ctx->d_link, S, nullptr, &ctx->d_edges, &ctx->edge_cap, &ctx->n_edges, &d_pair))) return rc;
    if (n2) LAUNCH(shard_sync_cut, grid_for(ctx, n2, 256, 8), 256, 0, ctx->d_link, n2, (uint32_t)me, ctx->d_link64);
    tstop(ctx, T_K6R);
    tstart(ctx, T_K6S);
    uint64_t Pn = 0, piece_nt = 0;
    if ((rc = number_units(ctx, d_pair, S, nullptr, &ctx->d_info, &d_uid_scan, &d_nt_scan, &Pn, &piece_nt))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_unitig_of_kmer, S + 1))) return rc;
    ctx->U = Pn; ctx->total_nt = piece_nt;
    dfree(ctx, d_pair);
    if ((rc = spell_units(ctx, d_uid_scan, d_nt_scan, true))) return rc;
    dfree(ctx, d_uid_scan); dfree(ctx, d_nt_scan);
    // the piece-level arrays move aside; ctx->d_seq etc. will hold the unitigs this GPU owns
    char *d_pseq = ctx->d_seq; uint64_t *d_pseq_off = ctx->d_seq_off; unsigned long long *d_pkc = ctx->d_kc;
    uint32_t *d_pabund = ctx->d_abund;            // -all-abundance-counts: per-k-mer abundances in piece order (else null)
    ctx->d_seq = nullptr; ctx->d_seq_off = nullptr; ctx->d_kc = nullptr; ctx->d_abund = nullptr;

    TRACE("local rank + spell pieces");
    // ---- stage 3: glue.  global piece ids, neighbour messages (X4), all-gather, replicated piece ranking
    tstart(ctx, T_GLUE);
    uint64_t all_P[MAX_WORLD], all_nt[MAX_WORLD], piece_base[MAX_WORLD + 1] = {0}, max_p = 1, max_nt = 16;
    uint64_t mine[3] = {Pn, piece_nt, S}, both[3 * MAX_WORLD], max_s = 4;
    if ((rc = allgather_host_u64(ctx, mine, 3, both))) return rc;
    for (int r = 0; r < W; r++) {
        all_P[r] = both[3 * r]; all_nt[r] = both[3 * r + 1]; piece_base[r + 1] = piece_base[r] + all_P[r];
        max_p = std::max(max_p, all_P[r]); max_nt = std::max(max_nt, all_nt[r]); max_s = std::max(max_s, both[3 * r + 2]);
    }
    max_nt = (max_nt + 15) & ~15ULL;
    const uint64_t G = piece_base[W];
    if (G >= (1ULL << 31)) return ctx->fail(CDBG_ERR_PARAM, "more than 2^31 pieces");
    ctx->piece_base_me = piece_base[me];
    PieceDesc *d_desc = nullptr, *d_desc_all = nullptr;
    if ((rc = dalloc(ctx, &d_desc, max_p))) return rc;
    if ((rc = dalloc(ctx, &d_desc_all, max_p * W))) return rc;
    CK(cudaMemsetAsync(d_desc, 0xFF, max_p * sizeof(PieceDesc), ctx->stream));
    if (Pn) LAUNCH(glue_describe, grid_for(ctx, Pn, 256, 8), 256, 0, Pn, (uint32_t)me, ctx->d_term_begin, ctx->d_term_end, ctx->d_link64, d_pseq_off,
                   d_pkc, k, d_desc);
    NK(ncclAllGather(d_desc, d_desc_all, max_p * sizeof(PieceDesc), ncclChar, ctx->comm, ctx->stream));
    TRACE("glue describe + allgather desc");
    char *d_seq_all = nullptr;
    if ((rc = dalloc(ctx, &d_seq_all, max_nt * W))) return rc;
    char *d_pad_seq = nullptr;
    {   // pad my piece strings to max_nt for the all-gather.  Only glue_assemble needs the strings: the all-gather runs on the second
        // stream, beside the table build and the piece-level ranking (NCCL keeps the issue order of the collectives of a communicator)
        if ((rc = dalloc(ctx, &d_pad_seq, max_nt))) return rc;
        if (piece_nt) CK(cudaMemcpyAsync(d_pad_seq, d_pseq, piece_nt, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
        NK(ncclAllGather(d_pad_seq, d_seq_all, max_nt, ncclChar, ctx->comm, ctx->stream2));
        CK(cudaEventRecord(ctx->ev_join, ctx->stream2));
    }
    uint32_t *d_abund_all = nullptr;
    if (d_pabund) {   // the per-k-mer abundance lists travel like the piece strings
        max_s = (max_s + 3) & ~3ULL;
        uint32_t *d_pad = nullptr;
        if ((rc = dalloc(ctx, &d_abund_all, max_s * W))) return rc;
        if ((rc = dalloc(ctx, &d_pad, max_s))) return rc;
        if (S) CK(cudaMemcpyAsync(d_pad, d_pabund, S * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        NK(ncclAllGather(d_pad, d_abund_all, max_s * 4, ncclChar, ctx->comm, ctx->stream));
        dfree(ctx, d_pad);
    }
    TRACE("allgather piece strings");
    uint64_t *d_piece_base = nullptr; uint32_t *d_plink = nullptr, *d_weight = nullptr;
    if ((rc = dalloc(ctx, &d_piece_base, MAX_WORLD + 1))) return rc;
    if ((rc = dalloc(ctx, &d_plink, 2 * G + 1))) return rc;
    if ((rc = dalloc(ctx, &d_weight, G + 1))) return rc;
    CK(cudaMemcpyAsync(d_piece_base, piece_base, (W + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    // replicated gid -> piece-terminus table (kept for the link stage)
    ctx->gid_table.cap = table_cap_for(2 * G + 16);
    if ((rc = dalloc(ctx, &ctx->gid_table.keys, ctx->gid_table.cap))) return rc;
    if ((rc = dalloc(ctx, &ctx->gid_table.vals, ctx->gid_table.cap))) return rc;
    CK(cudaMemsetAsync(ctx->gid_table.keys, 0xFF, (size_t)ctx->gid_table.cap * 8, ctx->stream));
    if (G) {
        LAUNCH(glue_table_build, grid_for(ctx, 2 * G, 256, 8), 256, 0, d_desc_all, max_p, (uint32_t)W, d_piece_base, G, ctx->gid_table);
        LAUNCH(glue_build, grid_for(ctx, G, 256, 8), 256, 0, d_desc_all, max_p, (uint32_t)W, d_piece_base, G, ctx->gid_table, d_plink, d_weight);
    }
    uint64_t *d_ppair = nullptr, *d_puid = nullptr, *d_pnt = nullptr, *d_pedges = nullptr; uint64_t n_pedges = 0;
    uint64_t pedge_cap = edge_min();
    if ((rc = dalloc(ctx, &d_pedges, pedge_cap))) return rc;
    if ((rc = run_ranking(ctx, d_plink, G, d_weight, &d_pedges, &pedge_cap, &n_pedges, &d_ppair))) return rc;
    uint64_t U_total = 0, nt_total = 0;
    if ((rc = number_units(ctx, d_ppair, G, d_weight, &ctx->d_pinfo, &d_puid, &d_pnt, &U_total, &nt_total))) return rc;
    dfree(ctx, d_ppair); dfree(ctx, d_plink);
    if (U_total >= (1ULL << 32)) return ctx->fail(CDBG_ERR_PARAM, "more than 2^32 unitigs");
    ctx->U_total = U_total;
    unsigned long long *d_kc_all = nullptr;
    if ((rc = dalloc(ctx, &ctx->d_unitig_of_piece, G + 1))) return rc;
    if ((rc = dalloc(ctx, &d_kc_all, U_total + 1))) return rc;
    CK(cudaMemsetAsync(d_kc_all, 0, (U_total + 1) * 8, ctx->stream));
    if (G) LAUNCH(glue_unitig_of_piece, grid_for(ctx, G, 256, 8), 256, 0, ctx->d_pinfo, d_puid, G, d_desc_all, max_p, (uint32_t)W, d_piece_base,
                  ctx->d_unitig_of_piece, d_kc_all);

    TRACE("piece-level ranking");
    // ---- ownership of the final unitigs: contiguous id ranges; assemble mine from the all-gathered piece strings
    const uint64_t u_first = ((uint64_t)me * U_total + W - 1) / W, u_last = ((uint64_t)(me + 1) * U_total + W - 1) / W;
    ctx->u_first = u_first;
    // nucleotide offset of a unitig = nt_scan at its head piece; find the offsets of u_first and u_last
    uint64_t *d_unt_off = nullptr;
    if ((rc = dalloc(ctx, &d_unt_off, U_total + 2))) return rc;
    if (G) LAUNCH(glue_unitig_offsets, grid_for(ctx, G, 256, 8), 256, 0, ctx->d_pinfo, d_puid, d_pnt, G, d_unt_off);
    CK(cudaMemcpyAsync(d_unt_off + U_total, &nt_total, 8, cudaMemcpyHostToDevice, ctx->stream));
    uint64_t nt_range[2] = {0, 0};
    SYNC_STREAM(ctx);   // (a copy to pageable memory blocks inside the call: wait by polling first, see stream_sync)
    CK(cudaMemcpyAsync(&nt_range[0], d_unt_off + u_first, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&nt_range[1], d_unt_off + u_last, 8, cudaMemcpyDeviceToHost, ctx->stream));
    SYNC_STREAM(ctx);
    const uint64_t U_mine = u_last - u_first, nt_mine = nt_range[1] - nt_range[0];
    if ((rc = dalloc(ctx, &ctx->d_seq, nt_mine + 1))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_seq_off, U_mine + 1))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_kc, U_mine + 1))) return rc;
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));      // the piece strings have arrived
    dfree(ctx, d_pad_seq);
    if (G && U_mine) LAUNCH(glue_assemble, grid_for(ctx, G * 32, 256, 8), 256, 0, G, ctx->d_pinfo, ctx->d_unitig_of_piece, d_pnt, d_desc_all, max_p, (uint32_t)W,
                            d_piece_base, d_seq_all, max_nt, (uint32_t)u_first, (uint32_t)u_last, nt_range[0], k, ctx->d_seq, ctx->d_seq_off);
    if (d_abund_all) {
        if ((rc = dalloc(ctx, &ctx->d_abund, nt_mine - U_mine * (uint64_t)(k - 1) + 1))) return rc;
        if (G && U_mine) LAUNCH(glue_assemble_abund, grid_for(ctx, G * 32, 256, 8), 256, 0, G, ctx->d_pinfo, ctx->d_unitig_of_piece, d_pnt, d_desc_all, max_p,
                                (uint32_t)W, d_piece_base, d_abund_all, max_s, (uint32_t)u_first, (uint32_t)u_last, nt_range[0], k, ctx->d_abund);
    }
    CK(cudaMemcpyAsync(ctx->d_seq_off + U_mine, &nt_mine, 8, cudaMemcpyHostToDevice, ctx->stream));
    if (U_mine) CK(cudaMemcpyAsync(ctx->d_kc, d_kc_all + u_first, U_mine * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    SYNC_STREAM(ctx);   // nt_mine is a stack variable

    TRACE("assemble unitigs");
    // ---- collect every edge this GPU is responsible for as (from gid, to gid): join edges are already there;
    //      add the edges of locally cut cycles and of piece-level cut cycles
    if (ctx->n_edges) {
        if ((rc = grow_sedges(ctx, ctx->n_sedges + ctx->n_edges))) return rc;
        LAUNCH(shard_local_edges_to_gid, grid_for(ctx, ctx->n_edges, 256, 8), 256, 0, ctx->d_edges, ctx->n_edges, (uint32_t)me, ctx->d_sedges + ctx->n_sedges);
        ctx->n_sedges += ctx->n_edges;
    }
    if (n_pedges) {
        if ((rc = grow_sedges(ctx, ctx->n_sedges + n_pedges))) return rc;
        if ((rc = ctr_zero(ctx))) return rc;
        CK(cudaMemcpyAsync(ctx->d_ctr + 0, &ctx->n_sedges, 8, cudaMemcpyHostToDevice, ctx->stream));
        LAUNCH(glue_cut_edges, grid_for(ctx, n_pedges, 256, 8), 256, 0, d_pedges, n_pedges, d_desc_all, max_p, (uint32_t)W, d_piece_base, (uint32_t)me,
               ctx->d_sedges, ctx->sedge_cap, ctx->d_ctr + 0);
        if ((rc = ctr_read(ctx))) return rc;
        ctx->n_sedges = ctx->h_ctr[0];
        if (ctx->n_sedges > ctx->sedge_cap) return ctx->fail(CDBG_ERR_INTERNAL, "edge buffer overflow");
    }
    dfree(ctx, d_pedges); dfree(ctx, d_desc); dfree(ctx, d_desc_all); dfree(ctx, d_seq_all); dfree(ctx, d_piece_base); dfree(ctx, d_weight);
    dfree(ctx, d_puid); dfree(ctx, d_pnt); dfree(ctx, d_kc_all); dfree(ctx, d_unt_off);
    dfree(ctx, d_pseq); dfree(ctx, d_pseq_off); dfree(ctx, d_pkc); dfree(ctx, d_pabund); dfree(ctx, d_abund_all);
    tstop(ctx, T_GLUE);
    tstop(ctx, T_K6S);
    TRACE("edges + cleanup");
    ctx->n_pieces = Pn; ctx->st.n_pieces = (uint32_t)Pn;
    ctx->U = U_mine; ctx->total_nt = nt_mine;
    ctx->st.n_unitigs = U_mine; ctx->st.total_nt = nt_mine;
    tstop(ctx, T_COMPACT);
    ctx->stage = ST_COMPACTED;
    return CDBG_OK;
}

// ------------------------------------------------------------------------------------------------ stage 4 (sharded)

int shard_stage_link(cdbg_ctx *ctx) {
    int rc;
    const int W = ctx->world;
    tstart(ctx, T_LINK);
    // both ends of every edge are resolved through the replicated gid table; one exchange to the owner of the FROM unitig
    ulonglong2 *c = nullptr; uint64_t nc = 0;
    if ((rc = ctr_zero(ctx))) return rc;
    if (ctx->n_sedges) LAUNCH(links_resolve, grid_for(ctx, ctx->n_sedges, 256, 8), 256, 0, ctx->d_sedges, ctx->n_sedges, ctx->gid_table, ctx->d_pinfo,
                              ctx->d_unitig_of_piece, std::max<uint64_t>(ctx->U_total, 1), (uint32_t)W, ctx->d_flag);
    tstart(ctx, T_XLINK);
    if ((rc = route16(ctx, ctx->d_sedges, ctx->n_sedges, &c, &nc))) return rc;
    tstop(ctx, T_XLINK);
    {
        int bad = 0;
        SYNC_STREAM(ctx);   // (a copy to pageable memory blocks inside the call: wait by polling first, see stream_sync)
        CK(cudaMemcpyAsync(&bad, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SYNC_STREAM(ctx);
        if (bad) return ctx->fail(CDBG_ERR_INTERNAL, "an edge end is not a piece terminus");
    }
    const uint64_t U = ctx->U;
    LinkRec *d_tmp = nullptr; uint32_t *d_per = nullptr, *d_cur = nullptr;
    if ((rc = dalloc(ctx, &ctx->d_links, nc + 1))) return rc;
    if ((rc = dalloc(ctx, &ctx->d_link_off, U + 1))) return rc;
    if ((rc = dalloc(ctx, &d_tmp, nc + 1))) return rc;
    if ((rc = dalloc(ctx, &d_per, U + 1))) return rc;
    if ((rc = dalloc(ctx, &d_cur, U + 1))) return rc;
    CK(cudaMemsetAsync(d_per, 0, (U + 1) * 4, ctx->stream));
    CK(cudaMemsetAsync(d_cur, 0, (U + 1) * 4, ctx->stream));
    if (nc) LAUNCH(links_finalize, grid_for(ctx, nc, 256, 8), 256, 0, c, nc, (uint32_t)ctx->u_first, d_tmp, d_per);
    uint64_t total = 0;
    if ((rc = exclusive_scan(ctx, d_per, U + 1, ctx->d_link_off, &total))) return rc;
    if (total != nc) return ctx->fail(CDBG_ERR_INTERNAL, "link grouping lost records");
    if (nc) LAUNCH(links_group, grid_for(ctx, nc, 256, 8), 256, 0, d_tmp, nc, (uint32_t)ctx->u_first, ctx->d_link_off, d_cur, ctx->d_links);
    dfree(ctx, c); dfree(ctx, d_tmp); dfree(ctx, d_per); dfree(ctx, d_cur);
    tstop(ctx, T_LINK);
    TRACE("links (resolve + 1 route + group)");
    ctx->n_edges = nc;
    ctx->st.n_links = nc;
    ctx->stage = ST_LINKED;
    return CDBG_OK;
}
