// host_scan.inl — exact-order path orchestration (included by cgvec_api.cu inside its anonymous namespace):
// scan planning (plan_scan), K1 launch, multi-level merge (merge_lists), the NCCL / fused peer-memory exchange and the
// per-batch drivers local_exact / scan_batch.
// ---- scan planning / launch ---------------------------------------------------------------------
// `ld`, `esize`, `dim`, `n` default to the index's own matrix; the int8 scan passes its code matrix (words of 4 codes).
int plan_scan(const Index* ix, uint32_t k, uint32_t nq, ScanGeom* g, uint32_t ld = 0, uint32_t esize = 0, uint32_t dim = 0, uint64_t n = ~0ull,
              uint32_t extra_smem = 0) {
    if (ld == 0) { ld = ix->ld; esize = ix->esize; dim = ix->dim; }
    if (n == ~0ull) n = ix->n;
    g->row_words = scan_row_words(ld, esize);
    const uint32_t try_tiles[4] = {16, 32, 8, 4};
    uint32_t best_bytes = 0;
    for (int t = 0; t < 4; ++t) {
        uint32_t tile = ix->opt_tile_rows ? (uint32_t)ix->opt_tile_rows : try_tiles[t];
        if (tile != 4 && tile != 8 && tile != 16 && tile != 32) return fail(CGVEC_ERR_BAD_ARG, "tile_rows must be 4, 8, 16 or 32");
        const uint32_t gmax = kScanConsumerWarps / (tile / 4);
        uint32_t max_stages = ix->opt_stages ? (uint32_t)ix->opt_stages : 8;
        for (uint32_t s = max_stages; s >= 2; --s) {
            // A stage must always be drained by the same warp group, otherwise a group would wait on a phase of the
            // stage's mbarrier without having observed the previous one (parity aliasing): active groups divide stages.
            uint32_t groups = 1;
            for (uint32_t a = gmax; a >= 1; --a) if (s % a == 0) { groups = a; break; }
            if (ix->opt_stages == 0 && groups < gmax && groups * 2 <= gmax && s > 2) continue;   // prefer well-populated groupings
            uint32_t sync = ix->opt_sync ? (uint32_t)ix->opt_sync : 8;
            sync = ((sync + groups - 1) / groups) * groups;
            uint32_t cand = next_pow2(k + sync * tile);
            if (cand < 64) cand = 64;
            ScanSmemLayout L = scan_smem_layout(g->row_words, tile, s, dim, nq, cand);
            if (L.total + extra_smem > kSmemBudget) continue;   // extra_smem: what the caller appends behind this layout (resident server)
            uint32_t bytes = s * tile * g->row_words * 4;
            if (bytes > best_bytes + best_bytes / 8) {          // keep the first (preferred) tile unless another buffers >12% more
                best_bytes = bytes;
                g->tile_rows = tile; g->stages = s; g->groups = groups; g->sync_interval = sync; g->cand_cap = cand; g->smem = L.total;
            }
            break;
        }
        if (ix->opt_tile_rows) break;
    }
    if (!best_bytes) return fail(CGVEC_ERR_UNSUPPORTED, "dimension %u (k=%u, nq=%u) does not fit the scan kernel's shared memory", ix->dim, k, nq);
    uint64_t tiles = (n + g->tile_rows - 1) / g->tile_rows;
    uint32_t grid = ix->opt_grid ? (uint32_t)ix->opt_grid : (uint32_t)ix->sm_count;
    g->grid = (uint32_t)(tiles < grid ? tiles : grid);
    if (g->grid == 0) g->grid = 1;
    return CGVEC_OK;
}

// cudaFuncSetAttribute is per device: remember which (function, device) pairs have been raised already.
template <typename F>
int ensure_smem_attr(F func, uint32_t bytes) {
    static std::mutex mu;
    static std::vector<std::pair<const void*, int>> done;
    int dev = 0;
    cudaGetDevice(&dev);
    const void* key = reinterpret_cast<const void*>(func);
    std::lock_guard<std::mutex> lk(mu);
    for (auto& d : done) if (d.first == key && d.second == dev) return CGVEC_OK;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return fail(CGVEC_ERR_CUDA, "cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
    done.emplace_back(key, dev);
    return CGVEC_OK;
}

template <typename T, int METRIC, int NQ>
int launch_scan_t(const ScanParams& p, const ScanGeom& g, cudaStream_t st) {
    int arc = ensure_smem_attr(scan_exact_kernel<T, METRIC, NQ>, kSmemBudget);
    if (arc) return arc;
    // Programmatic dependent launch: the kernel ahead of us in the stream is normally the previous query's merge,
    // whose output we do not read (partials are double buffered), so our CTAs may start as soon as it has started.
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g.grid);
    cfg.blockDim = dim3(kScanThreads);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = g.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, scan_exact_kernel<T, METRIC, NQ>, p));
    return CGVEC_OK;
}
template <typename T, int METRIC>
int launch_scan_q(uint32_t nq, const ScanParams& p, const ScanGeom& g, cudaStream_t st) {
    switch (nq) {
        case 1: return launch_scan_t<T, METRIC, 1>(p, g, st);
        case 2: return launch_scan_t<T, METRIC, 2>(p, g, st);
        case 4: return launch_scan_t<T, METRIC, 4>(p, g, st);
    }
    return fail(CGVEC_ERR_BAD_ARG, "internal: scan batch %u", nq);
}
template <typename T>
int launch_scan_m(int metric, uint32_t nq, const ScanParams& p, const ScanGeom& g, cudaStream_t st) {
    switch (metric) {
        case CGVEC_COSINE: return launch_scan_q<T, METRIC_COSINE>(nq, p, g, st);
        case CGVEC_DOT: return launch_scan_q<T, METRIC_DOT>(nq, p, g, st);
        case CGVEC_L2: return launch_scan_q<T, METRIC_L2>(nq, p, g, st);
    }
    return fail(CGVEC_ERR_BAD_ARG, "unknown metric %d", metric);
}

// Merge `lists` key lists per query (element (q, l, i) at in[q*q_stride + l*l_stride + i]) down to one list of k.
// The final level decodes into d_rows/d_scores/d_counts when given, and/or writes keys to `final_keys`.
int merge_lists(Index* ix, SearchCtx* c, const uint64_t* in, uint32_t nq, uint32_t lists, uint32_t k, int ascending,
                uint64_t* final_keys, uint64_t* d_rows, float* d_scores, uint32_t* d_counts, cudaStream_t st,
                size_t q_stride, size_t l_stride, uint32_t list_len = 0, int sorted_in = 1) {
    NvtxRange nvtx_("cgvec.merge");
    if (list_len == 0) list_len = k;
    {
        int arc = ensure_smem_attr(merge_topk_kernel, kMergeMaxKeys * 8);
        if (arc) return arc;
    }
    const uint64_t* cur = in;
    uint32_t cur_lists = lists;
    int pp = 0;
    uint64_t qs = q_stride, ls = l_stride;                       // the kernel reads strided input directly (no re-packing copies)
    while (true) {
        uint32_t per_cta_max = (kMergeMaxKeys - 8 * kTournamentMaxK) / list_len < 2 ? 2 : (kMergeMaxKeys - 8 * kTournamentMaxK) / list_len;
        if (per_cta_max > 256) per_cta_max = 256;
        uint32_t per_cta = cur_lists < per_cta_max ? cur_lists : per_cta_max;
        uint32_t n_out = (cur_lists + per_cta - 1) / per_cta;
        uint32_t sort_n = next_pow2(per_cta * list_len);
        if (sort_n < 2) sort_n = 2;
        const bool last = (n_out == 1);
        uint64_t* out = last ? final_keys : c->d_part[pp];
        dim3 grid(n_out, nq);
        const bool tournament = sorted_in && k <= kTournamentMaxK && per_cta <= 256;
        const size_t smem = tournament ? ((size_t)per_cta * list_len + 8 * k) * 8 : (size_t)sort_n * 8;
        if (tournament) sort_n = per_cta * list_len;            // staging area size (keys) ahead of the level-2 lists
        {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = grid; cfg.blockDim = dim3(kMergeThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = ix->opt_pdl >= 2 ? 1 : 0;
            cfg.attrs = attr; cfg.numAttrs = 1;
            CUDA_TRY(cudaLaunchKernelEx(&cfg, merge_topk_kernel, cur, cur_lists, list_len, k, per_cta, sort_n, out, ascending,
                                        last ? d_rows : (uint64_t*)nullptr, last ? d_scores : (float*)nullptr, last ? d_counts : (uint32_t*)nullptr,
                                        sorted_in, trace_slot(ix, 2), qs, ls));
        }
        ix->launches++;
        CUDA_TRY(cudaGetLastError());
        if (last) break;
        cur = out;
        cur_lists = n_out;
        list_len = k;
        sorted_in = 1;
        qs = (uint64_t)n_out * k; ls = k;                        // intermediate levels are dense [nq][n_out][k]
        pp ^= 1;
    }
    return CGVEC_OK;
}

int ensure_parts(SearchCtx* c, size_t need_part);

// Exchange step of a sharded index: this rank's best-k keys [nq][k] -> one NCCL all-gather -> every rank merges
// the `world` lists and decodes.  The single collective of the path.
int exchange_and_decode(Index* ix, SearchCtx* c, uint64_t* local_keys, uint32_t nq, uint32_t k, int ascending, cudaStream_t st,
                        uint64_t* d_rows, float* d_scores, uint32_t* d_counts) {
    NvtxRange nvtx_("cgvec.exchange.nccl");
    size_t per_rank = (size_t)nq * k;
    {   // the gathered [rank][query][k] layout is re-packed to [query][rank][k] in the merge scratch
        int rc = ensure_parts(c, per_rank * (size_t)ix->world);
        if (rc) return rc;
    }
    {
        std::lock_guard<std::mutex> lk(ix->comm_mu);
        NCCL_TRY(nccl_api().AllGather(local_keys, c->d_gather, per_rank, kNcclUint64, ix->comm, st));
    }
    ix->last_exchange = 2;
    return merge_lists(ix, c, c->d_gather, nq, ix->world, k, ascending, nullptr, d_rows, d_scores, d_counts, st, (size_t)k, per_rank);
}

int ensure_gather(Index* ix, SearchCtx* c, uint32_t nq, uint32_t k, uint64_t** local_keys) {
    size_t per_rank = (size_t)nq * k;
    size_t cap_g = c->gather_cap;
    int rc = ensure(&c->d_gather, &cap_g, per_rank * (ix->world + 1));
    if (rc) return rc;
    c->gather_cap = cap_g;
    *local_keys = c->d_gather + per_rank * ix->world;
    return CGVEC_OK;
}

int ensure_parts(SearchCtx* c, size_t need_part) {
    if (need_part > c->part_cap) {
        size_t cap0 = c->part_cap, cap1 = c->part_cap;
        int rc = ensure(&c->d_part[0], &cap0, need_part); if (rc) return rc;
        rc = ensure(&c->d_part[1], &cap1, need_part); if (rc) return rc;
        c->part_cap = cap0 < cap1 ? cap0 : cap1;
    }
    return CGVEC_OK;
}

// Exact-order scan (K1) of the local shard for `nq` (1, 2 or 4) queries already on the device at `d_q`
// (stride = dim rounded up to 4 floats).  Leaves this shard's best-k keys in `local_keys` when given, else
// decodes straight into d_rows/d_scores/d_counts.
int local_exact(Index* ix, SearchCtx* c, const float* d_q, uint32_t nq, uint32_t k, int metric, cudaStream_t st,
                uint64_t* local_keys, uint64_t* d_rows, float* d_scores, uint32_t* d_counts,
                const uint64_t** partials_out = nullptr, uint32_t* lists_out = nullptr) {
    NvtxRange nvtx_("cgvec.exact_scan");
    ScanGeom g;
    int rc = plan_scan(ix, k, nq, &g);
    if (rc) return rc;
    const int ascending = (metric == CGVEC_L2);
    rc = ensure_parts(c, (size_t)nq * (g.grid > (uint32_t)ix->world ? g.grid : ix->world) * k);
    if (rc) return rc;
    {
        size_t need = (size_t)nq * g.grid * k;
        if (need > c->scan_cap) {
            size_t c0 = c->scan_cap, c1 = c->scan_cap;
            rc = ensure(&c->d_scan[0], &c0, need); if (rc) return rc;
            rc = ensure(&c->d_scan[1], &c1, need); if (rc) return rc;
            c->scan_cap = c0 < c1 ? c0 : c1;
        }
    }
    g.pdl = (uint32_t)ix->opt_pdl;
    uint64_t* partials = c->d_scan[c->scan_flip & 1];
    c->scan_flip++;
    ScanParams p = map_params(ix);
    p.rows = ix->d_rows; p.norms = ix->d_norms; p.queries = d_q; p.partials = partials;
    p.n_rows = ix->n; p.d = ix->dim; p.ld = ix->ld; p.row_words = g.row_words; p.tile_rows = g.tile_rows;
    p.stages = g.stages; p.active_groups = g.groups; p.k = k; p.cand_cap = g.cand_cap; p.sync_interval = g.sync_interval; p.use_l2_hint = ix->opt_l2_hint;
    p.trace = trace_slot(ix, 1);
    p.early_trigger = (ix->opt_pdl >= 2 && g.grid >= (uint32_t)ix->sm_count) ? 1u : 0u;   // only with every SM occupied

    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ix->opt_timing) {
        CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1));
        CUDA_TRY(cudaEventRecord(e0, st));
    }
    rc = (ix->dtype == CGVEC_F32) ? launch_scan_m<float>(metric, nq, p, g, st) : launch_scan_m<__half>(metric, nq, p, g, st);
    if (rc) return rc;
    ix->launches++;
    if (ix->opt_timing) {
        CUDA_TRY(cudaEventRecord(e1, st));
        std::lock_guard<std::mutex> lk(ix->ev_mu);
        ix->timed.push_back({e0, e1, 0});
    }
    ix->last_geom = g;
    if (partials_out) { *partials_out = partials; *lists_out = g.grid; return CGVEC_OK; }   // caller fuses merge + exchange
    return merge_lists(ix, c, partials, nq, g.grid, k, ascending, local_keys, d_rows, d_scores, d_counts, st, (size_t)g.grid * k, k);
}

int scan_batch(Index* ix, SearchCtx* c, const float* d_q, uint32_t nq, uint32_t k, int metric, cudaStream_t st,
               uint64_t* d_rows, float* d_scores, uint32_t* d_counts) {
    if (ix->world == 1) return local_exact(ix, c, d_q, nq, k, metric, st, nullptr, d_rows, d_scores, d_counts);
    if (ix->p2p && ix->opt_p2p && k <= kXchgMaxK && nq <= kXchgMaxQ) {
        // merge + exchange + merge as one kernel over NVLink peer memory (exchange.cuh)
        {
            int arc = ensure_smem_attr(xchg_merge_kernel, 160 * 1024);
            if (arc) return arc;
        }
        const uint64_t* partials = nullptr;
        uint32_t lists = 0;
        int rc = local_exact(ix, c, d_q, nq, k, metric, st, nullptr, nullptr, nullptr, nullptr, &partials, &lists);
        if (rc) return rc;
        if (lists <= 256) {
            NvtxRange nvtx_x("cgvec.exchange.p2p");
            std::lock_guard<std::mutex> lk(ix->comm_mu);          // same step order on every rank
            XchgParams xp{};
            xp.partials = partials; xp.n_lists = lists; xp.k = k; xp.nq = nq; xp.ascending = (metric == CGVEC_L2);
            xp.rank = (uint32_t)ix->rank; xp.world = (uint32_t)ix->world; xp.seq = ++ix->xseq;
            for (int r = 0; r < ix->world; ++r) xp.peer[r] = ix->xpeer[r];
            xp.out_rows = d_rows; xp.out_scores = d_scores; xp.out_counts = d_counts;
            xp.trace = trace_slot(ix, 3);
            xp.err = ix->h_xerr; xp.timeout_ns = (uint64_t)(ix->opt_xchg_timeout_ms > 0 ? ix->opt_xchg_timeout_ms : 5000) * 1000000ull;
            const size_t smem = ((size_t)lists * k + 9 * k) * 8;
            {
                cudaLaunchConfig_t cfg{};
                cfg.gridDim = dim3(nq); cfg.blockDim = dim3(kXchgThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = ix->opt_pdl >= 2 ? 1 : 0;
                cfg.attrs = attr; cfg.numAttrs = 1;
                CUDA_TRY(cudaLaunchKernelEx(&cfg, xchg_merge_kernel, xp));
            }
            ix->launches++;
            ix->last_exchange = 1;
            return CGVEC_OK;
        }
        uint64_t* local_keys = nullptr;
        rc = ensure_gather(ix, c, nq, k, &local_keys);
        if (rc) return rc;
        rc = merge_lists(ix, c, partials, nq, lists, k, metric == CGVEC_L2, local_keys, nullptr, nullptr, nullptr, st, (size_t)lists * k, k);
        if (rc) return rc;
        return exchange_and_decode(ix, c, local_keys, nq, k, metric == CGVEC_L2, st, d_rows, d_scores, d_counts);
    }
    uint64_t* local_keys = nullptr;
    int rc = ensure_gather(ix, c, nq, k, &local_keys);
    if (rc) return rc;
    rc = local_exact(ix, c, d_q, nq, k, metric, st, local_keys, nullptr, nullptr, nullptr);
    if (rc) return rc;
    return exchange_and_decode(ix, c, local_keys, nq, k, metric == CGVEC_L2, st, d_rows, d_scores, d_counts);
}

