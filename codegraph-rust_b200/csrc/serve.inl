// serve.inl — host side of the resident batch-1 server (scan_serve.cuh).  Included by cgvec_api.cu after the search entry
// points (global scope; helpers of the anonymous namespace are visible).
//
// The reference's serving shape is one search_similar call per query (VectorStore::search_similar, traits.rs:11-16;
// SemanticSearch::search_by_embedding, search.rs:91-144).  A session keeps the scan kernel resident between those calls: a
// submit is a 64-byte descriptor + a doorbell word in pinned memory, a completion is a word the host spins on; the kernel is
// (re)launched on demand and leaves by itself when idle.  Results are those of cgvec_search (same arithmetic, same keys).
struct cgvec_server {
    Index* ix = nullptr;
    uint32_t k = 0, qstride = 0, grid = 0;
    int metric = CGVEC_COSINE;
    ServeParams P{};
    ServeHostBlock* h = nullptr;            // pinned + mapped
    ServeCtrl* d_ctrl = nullptr;
    float* d_qbuf = nullptr;
    uint64_t* d_lists = nullptr;
    // pinned staging of host-I/O submissions, one slot per descriptor slot
    float* h_q = nullptr;
    uint64_t* h_rows = nullptr;
    float* h_scores = nullptr;
    uint32_t* h_counts = nullptr;
    float* dv_q = nullptr; uint64_t* dv_rows = nullptr; float* dv_scores = nullptr; uint32_t* dv_counts = nullptr;   // their device addresses
    struct HostOut { uint64_t* rows; uint8_t (*ids)[16]; float* scores; uint32_t* count; bool host_io; };
    HostOut outs[kServeSlots] = {};
    bool uncollected[kServeSlots] = {};     // host-I/O ticket of this slot completed or in flight, results not handed out yet
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // cgvec_serve_timer_*: device-side bracket on the session's launch stream
    uint32_t next_seq = 1;                  // next sequence number to hand out
    bool launched = false;                  // a kernel has been enqueued that has not announced its exit yet
    std::mutex mu;                          // submissions and relaunches
    uint64_t launches = 0;
    int idle_us = 200, life_ms = 2000, abort_ms = 20000, wait_ms = 30000;
    int max_inflight = (int)kServeSlots - 2;   // submissions the host lets run ahead of the oldest incomplete one
};

namespace {

__global__ void serve_reset_kernel(ServeCtrl* c, uint32_t last_done) {
    if (threadIdx.x == 0) {
        c->reserved0 = 0ull; c->go = last_done; c->exit_seq = 0; c->completed = last_done;
        for (uint32_t i = 0; i < kServeSlots; ++i) c->done[i] = 0;
    }
    for (uint32_t i = threadIdx.x; i < kServeMaxGrid; i += blockDim.x) { c->cta_line[i][0] = last_done; c->cta_line[i][1] = 0; }
}

template <typename T>
int serve_launch_t(cgvec_server* s, int metric, uint32_t smem) {
    switch (metric) {
        case CGVEC_COSINE: { int rc = ensure_smem_attr(scan_serve_kernel<T, METRIC_COSINE>, kSmemBudget); if (rc) return rc;
                             scan_serve_kernel<T, METRIC_COSINE><<<s->grid, kServeThreads, smem, s->st>>>(s->P); break; }
        case CGVEC_DOT:    { int rc = ensure_smem_attr(scan_serve_kernel<T, METRIC_DOT>, kSmemBudget); if (rc) return rc;
                             scan_serve_kernel<T, METRIC_DOT><<<s->grid, kServeThreads, smem, s->st>>>(s->P); break; }
        case CGVEC_L2:     { int rc = ensure_smem_attr(scan_serve_kernel<T, METRIC_L2>, kSmemBudget); if (rc) return rc;
                             scan_serve_kernel<T, METRIC_L2><<<s->grid, kServeThreads, smem, s->st>>>(s->P); break; }
        default: return fail(CGVEC_ERR_BAD_ARG, "unknown metric %d", metric);
    }
    CUDA_TRY(cudaGetLastError());
    return CGVEC_OK;
}

// Enqueues a (re)launch that serves sequence numbers from `start` on.  Caller holds s->mu.  Any previous kernel of this session
// has announced its exit (or never existed); launches on s->st serialise behind it.
int serve_launch(cgvec_server* s, uint32_t start) {
    Index* ix = s->ix;
    CUDA_TRY(cudaSetDevice(ix->device));
    s->h->exit_seq = 0;
    s->h->stop = 0;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    serve_reset_kernel<<<1, 32, 0, s->st>>>(s->d_ctrl, start - 1);
    ix->launches++;
    s->P.start_seq = start;
    s->P.sp.rows = ix->d_rows; s->P.sp.norms = ix->d_norms; s->P.sp.n_rows = ix->n;
    int rc = ix->dtype == CGVEC_F32 ? serve_launch_t<float>(s, s->metric, s->P.smem_total) : serve_launch_t<__half>(s, s->metric, s->P.smem_total);
    if (rc) return rc;
    ix->launches++;
    s->launches++;
    s->launched = true;
    return CGVEC_OK;
}

// The kernel announces its exit in h->exit_seq (first sequence number it will not serve).  If that leaves submitted work
// unserved, start the next kernel from there.  Caller holds s->mu.
int serve_keep_alive(cgvec_server* s) {
    const uint32_t ex = s->h->exit_seq;
    if (s->launched && ex == 0) return CGVEC_OK;                 // resident and listening
    if (s->launched && ex != 0) s->launched = false;             // leaving (or gone): everything below ex is or will be served by it
    const uint32_t first_unserved = ex != 0 ? ex : s->h->completed + 1;
    if ((int32_t)(s->next_seq - first_unserved) <= 0) return CGVEC_OK;   // nothing submitted beyond what it serves
    return serve_launch(s, first_unserved);
}

void serve_free(cgvec_server* s) {
    if (!s) return;
    cudaSetDevice(s->ix->device);
    if (s->h) {
        s->h->stop = 1;
        std::atomic_thread_fence(std::memory_order_seq_cst);
    }
    if (s->st) cudaStreamSynchronize(s->st);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    cudaFree(s->d_ctrl); cudaFree(s->d_qbuf); cudaFree(s->d_lists);
    cudaFreeHost(s->h); cudaFreeHost(s->h_q); cudaFreeHost(s->h_rows); cudaFreeHost(s->h_scores); cudaFreeHost(s->h_counts);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
}

int serve_open_impl(Index* ix, uint32_t k, int metric, cgvec_server** out) {
    *out = nullptr;
    if (!ix->parts.empty()) return fail(CGVEC_ERR_UNSUPPORTED, "sessions serve single-device indexes and the ranks of a sharded index");
    if (k == 0 || k > kServeMaxK) return fail(CGVEC_ERR_UNSUPPORTED, "sessions serve 1 <= k <= %u", kServeMaxK);
    if (metric != CGVEC_COSINE && metric != CGVEC_DOT && metric != CGVEC_L2) return fail(CGVEC_ERR_BAD_ARG, "unknown metric %d", metric);
    if (ix->n == 0) return fail(CGVEC_ERR_UNSUPPORTED, "the index is empty");
    if (ix->world > 1 && !(ix->p2p && ix->opt_p2p && k <= kXchgMaxK)) return fail(CGVEC_ERR_UNSUPPORTED, "sessions on a sharded index need the peer-memory exchange");
    CUDA_TRY(cudaSetDevice(ix->device));
    std::unique_ptr<cgvec_server, void (*)(cgvec_server*)> s(new cgvec_server(), serve_free);
    s->ix = ix; s->k = k; s->metric = metric;
    s->qstride = (ix->dim + 3) & ~3u;
    // geometry: K1's plan with room left for the finishing CTA's merge staging
    ScanGeom g;
    const uint32_t sm = (uint32_t)ix->sm_count;
    uint32_t merge_lists = (sm + 31) / 32 * 32;
    int rc = CGVEC_ERR_UNSUPPORTED;
    uint32_t extra = 0;
    for (; merge_lists >= 32; merge_lists -= 32) {
        extra = 16 + kTcMaxStages * 4 + 128 + ((size_t)merge_lists * k + 9 * k) * 8 + 64;
        rc = plan_scan(ix, k, 2, &g, 0, 0, 0, ~0ull, extra);      // two query / candidate buffers (the helper warp works on the other one)
        const bool full_groups = rc == CGVEC_OK && g.groups == kScanConsumerWarps / (g.tile_rows / 4);
        if (full_groups && g.stages >= 3) break;                 // prefer a deep ring over staging every list at once
        if (full_groups && merge_lists == 32) break;
        if (rc == CGVEC_OK && !full_groups) rc = CGVEC_ERR_UNSUPPORTED;
    }
    if (rc) return fail(CGVEC_ERR_UNSUPPORTED, "no session geometry for dimension %u, k = %u", ix->dim, k);
    if (g.grid > kServeMaxGrid) g.grid = kServeMaxGrid;
    s->grid = g.grid;
    ServeParams& P = s->P;
    P.sp = map_params(ix);
    P.sp.d = ix->dim; P.sp.ld = ix->ld; P.sp.row_words = g.row_words; P.sp.tile_rows = g.tile_rows; P.sp.stages = g.stages;
    P.sp.active_groups = g.groups; P.sp.k = k; P.sp.cand_cap = g.cand_cap; P.sp.sync_interval = g.sync_interval; P.sp.use_l2_hint = ix->opt_l2_hint;
    P.epoch_rounds = g.sync_interval / g.groups;
    const ScanSmemLayout L = scan_smem_layout(g.row_words, g.tile_rows, g.stages, ix->dim, 2, g.cand_cap);
    P.off_ctl = (L.total + 15) & ~15u;
    P.off_merge = P.off_ctl + 128;                               // 8 mbarriers + control words + the two query norms
    P.merge_lists = merge_lists;
    P.smem_total = P.off_merge + (uint32_t)(((size_t)merge_lists * k + 9 * k) * 8);
    if (P.smem_total > kSmemBudget) return fail(CGVEC_ERR_UNSUPPORTED, "no shared memory for a session at dimension %u, k = %u", ix->dim, k);
    P.idle_ns = (uint64_t)s->idle_us * 1000ull; P.life_ns = (uint64_t)s->life_ms * 1000000ull; P.abort_ns = (uint64_t)s->abort_ms * 1000000ull;
    P.rank = (uint32_t)ix->rank; P.world = (uint32_t)ix->world;
    for (int r = 0; r < ix->world && r < (int)kXchgMaxWorld; ++r) P.peer[r] = ix->xpeer[r];
    P.xchg_timeout_ns = (uint64_t)(ix->opt_xchg_timeout_ms > 0 ? ix->opt_xchg_timeout_ms : 5000) * 1000000ull;

    CUDA_TRY(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&s->h), sizeof(ServeHostBlock), cudaHostAllocMapped | cudaHostAllocPortable));
    memset(s->h, 0, sizeof(ServeHostBlock));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&s->h_q), (size_t)kServeSlots * s->qstride * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&s->h_rows), (size_t)kServeSlots * k * sizeof(uint64_t), cudaHostAllocMapped | cudaHostAllocPortable));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&s->h_scores), (size_t)kServeSlots * k * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
    CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&s->h_counts), (size_t)kServeSlots * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_ctrl), sizeof(ServeCtrl)));
    CUDA_TRY(cudaMemset(s->d_ctrl, 0, sizeof(ServeCtrl)));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_qbuf), (size_t)kServeSlots * s->qstride * sizeof(float)));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_lists), (size_t)kServeSlots * s->grid * k * sizeof(uint64_t)));
    P.ctrl = s->d_ctrl; P.qbuf = s->d_qbuf; P.lists = s->d_lists;
    void* hdev = nullptr;
    CUDA_TRY(cudaHostGetDevicePointer(&hdev, s->h, 0));
    P.host = static_cast<ServeHostBlock*>(hdev);
    CUDA_TRY(cudaHostGetDevicePointer(&hdev, s->h_q, 0)); s->dv_q = static_cast<float*>(hdev);
    CUDA_TRY(cudaHostGetDevicePointer(&hdev, s->h_rows, 0)); s->dv_rows = static_cast<uint64_t*>(hdev);
    CUDA_TRY(cudaHostGetDevicePointer(&hdev, s->h_scores, 0)); s->dv_scores = static_cast<float*>(hdev);
    CUDA_TRY(cudaHostGetDevicePointer(&hdev, s->h_counts, 0)); s->dv_counts = static_cast<uint32_t*>(hdev);
    ix->last_geom = g;
    *out = s.release();
    return CGVEC_OK;
}

// Spins until sequence number `seq` has completed; relaunches the kernel if it left before serving it.
int serve_wait_impl(cgvec_server* s, uint32_t seq) {
    const auto t0 = std::chrono::steady_clock::now();
    uint32_t spins = 0;
    while ((int32_t)(s->h->completed - seq) < 0) {
        if ((++spins & 63u) == 0) {
            if (s->h->exit_seq != 0) {
                std::lock_guard<std::mutex> lk(s->mu);
                int rc = serve_keep_alive(s);
                if (rc) return rc;
            }
            if ((spins & 0xffffu) == 0) {
                if (cudaStreamQuery(s->st) != cudaErrorNotReady && (int32_t)(s->h->completed - seq) < 0 && s->h->exit_seq == 0) {
                    cudaError_t e = cudaGetLastError();
                    return fail(CGVEC_ERR_CUDA, "the resident scan kernel ended without serving query %u (%s)", seq, cudaGetErrorString(e));
                }
                if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(s->wait_ms))
                    return fail(CGVEC_ERR_CUDA, "query %u did not complete within %d ms", seq, s->wait_ms);
            }
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (s->h->error) return fail(CGVEC_ERR_NCCL, "peer exchange timed out waiting for rank %u", s->h->error - 1);
    return CGVEC_OK;
}

int serve_submit_impl(cgvec_server* s, const float* query, bool device_io, uint64_t* out_rows, uint8_t (*out_ids)[16], float* out_scores,
                      uint32_t* out_count, uint32_t* out_seq) {
    Index* ix = s->ix;
    std::unique_lock<std::mutex> lk(s->mu);
    const uint32_t seq = s->next_seq;
    const uint32_t slot = seq % kServeSlots;
    if ((int32_t)(seq - s->h->completed) > s->max_inflight) {                // enough in flight: wait for the oldest
        lk.unlock();
        int rc = serve_wait_impl(s, seq - (uint32_t)s->max_inflight);
        if (rc) return rc;
        lk.lock();
        if (s->next_seq != seq) return fail(CGVEC_ERR_UNSUPPORTED, "concurrent submitters overran the session's ring; serialise cgvec_serve_submit calls");
    }
    if (s->uncollected[slot]) return fail(CGVEC_ERR_UNSUPPORTED, "ticket %u has not been waited for; at most %u host-I/O tickets may be outstanding", seq - kServeSlots, kServeSlots - 2);
    ServeDesc d{};
    if (device_io) {
        d.q_ptr = reinterpret_cast<uint64_t>(query); d.q_on_host = 0;
        d.out_rows = reinterpret_cast<uint64_t>(out_rows); d.out_scores = reinterpret_cast<uint64_t>(out_scores); d.out_counts = reinterpret_cast<uint64_t>(out_count);
    } else {
        float* hq = s->h_q + (size_t)slot * s->qstride;
        memcpy(hq, query, ix->dim * sizeof(float));
        for (uint32_t i = ix->dim; i < s->qstride; ++i) hq[i] = 0.0f;
        d.q_ptr = reinterpret_cast<uint64_t>(s->dv_q + (size_t)slot * s->qstride); d.q_on_host = 1;
        d.out_rows = reinterpret_cast<uint64_t>(s->dv_rows + (size_t)slot * s->k);
        d.out_scores = reinterpret_cast<uint64_t>(s->dv_scores + (size_t)slot * s->k);
        d.out_counts = reinterpret_cast<uint64_t>(s->dv_counts + slot);
    }
    s->outs[slot] = {out_rows, out_ids, out_scores, out_count, !device_io};
    s->uncollected[slot] = !device_io;
    if (ix->world > 1) { std::lock_guard<std::mutex> clk(ix->comm_mu); d.xseq = ++ix->xseq; ix->last_exchange = 1; }
    memcpy(const_cast<ServeDesc*>(&s->h->desc[slot]), &d, sizeof(d));
    std::atomic_thread_fence(std::memory_order_release);
    s->h->doorbell = seq;
    s->next_seq = seq + 1;
    ix->searches++;
    int rc = serve_keep_alive(s);
    if (rc) return rc;
    if (out_seq) *out_seq = seq;
    return CGVEC_OK;
}

// hands a completed host-I/O submission's results to the caller's buffers
void serve_collect(cgvec_server* s, uint32_t seq) {
    const uint32_t slot = seq % kServeSlots;
    const cgvec_server::HostOut& o = s->outs[slot];
    if (!o.host_io) return;
    s->uncollected[slot] = false;
    const Index* ix = s->ix;
    const uint32_t cnt = s->h_counts[slot];
    if (o.count) *o.count = cnt;
    for (uint32_t i = 0; i < s->k; ++i) {
        const bool valid = i < cnt;
        const uint64_t grow = valid ? s->h_rows[(size_t)slot * s->k + i] : ~0ull;
        if (o.rows) o.rows[i] = grow;
        if (o.scores) o.scores[i] = valid ? s->h_scores[(size_t)slot * s->k + i] : 0.0f;
        if (o.ids) {
            memset(o.ids[i], 0, 16);
            if (valid && grow >= ix->row_offset && grow - ix->row_offset < ix->n && ix->has_id[grow - ix->row_offset])
                memcpy(o.ids[i], &ix->ids[(grow - ix->row_offset) * 16], 16);
        }
    }
}

}  // namespace

CGVEC_EXPORT int cgvec_serve_open(cgvec_index* ix, uint32_t k, cgvec_metric metric, cgvec_server** out) {
    if (!ix || !out) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    RwGuard rw_guard_(ix, false);
    int rc = serve_open_impl(ix, k, (int)metric, out);
    if (rc == CGVEC_OK) ix->open_streams++;                      // the write side refuses to run while a session is open
    return rc;
}
CGVEC_EXPORT int cgvec_serve_submit(cgvec_server* s, const float* query, int device_io, uint64_t* out_rows, float* out_scores,
                                    uint32_t* out_count, uint32_t* out_ticket) {
    if (!s || !query) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    return serve_submit_impl(s, query, device_io != 0, out_rows, nullptr, out_scores, out_count, out_ticket);
}
CGVEC_EXPORT int cgvec_serve_wait(cgvec_server* s, uint32_t ticket) {
    if (!s) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    int rc = serve_wait_impl(s, ticket);
    if (rc) return rc;
    serve_collect(s, ticket);
    return CGVEC_OK;
}
CGVEC_EXPORT int cgvec_serve_search(cgvec_server* s, const float* query, uint64_t* out_rows, uint8_t (*out_ids)[16], float* out_scores,
                                    uint32_t* out_count) {
    if (!s || !query) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    uint32_t seq = 0;
    int rc = serve_submit_impl(s, query, false, out_rows, out_ids, out_scores, out_count, &seq);
    if (rc) return rc;
    rc = serve_wait_impl(s, seq);
    if (rc) return rc;
    serve_collect(s, seq);
    return CGVEC_OK;
}
/* Asks the resident kernel to leave now (it also leaves by itself when idle); the next submit starts it again. */
CGVEC_EXPORT int cgvec_serve_pause(cgvec_server* s) {
    if (!s) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    std::lock_guard<std::mutex> lk(s->mu);
    CUDA_TRY(cudaSetDevice(s->ix->device));
    s->h->stop = 1;
    std::atomic_thread_fence(std::memory_order_seq_cst);
    CUDA_TRY(cudaStreamSynchronize(s->st));
    s->launched = false;
    return CGVEC_OK;
}
/* Device-side timing of a run of submissions (CUDA events on the session's own launch stream).  timer_start pauses the session
 * and records the first event; the kernel is then launched by the next submit, i.e. INSIDE the bracket.  timer_stop waits for
 * everything submitted, makes the kernel leave, records the second event behind it and returns the elapsed milliseconds:
 * launch + all queries + exit, on the device clock. */
CGVEC_EXPORT int cgvec_serve_timer_start(cgvec_server* s) {
    if (!s) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    int rc = cgvec_serve_pause(s);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->ev0) { CUDA_TRY(cudaEventCreate(&s->ev0)); CUDA_TRY(cudaEventCreate(&s->ev1)); }
    CUDA_TRY(cudaEventRecord(s->ev0, s->st));
    return CGVEC_OK;
}
CGVEC_EXPORT int cgvec_serve_timer_stop(cgvec_server* s, float* out_ms) {
    if (!s || !out_ms) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    if (!s->ev0) return fail(CGVEC_ERR_BAD_ARG, "cgvec_serve_timer_start was not called");
    if (s->next_seq > 1) { int rc = serve_wait_impl(s, s->next_seq - 1); if (rc) return rc; }
    int rc = cgvec_serve_pause(s);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(s->mu);
    CUDA_TRY(cudaEventRecord(s->ev1, s->st));
    CUDA_TRY(cudaEventSynchronize(s->ev1));
    CUDA_TRY(cudaEventElapsedTime(out_ms, s->ev0, s->ev1));
    return CGVEC_OK;
}
CGVEC_EXPORT int cgvec_serve_stats(const cgvec_server* s, uint64_t* out_launches, uint64_t* out_served) {
    if (!s) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    if (out_launches) *out_launches = s->launches;
    if (out_served) *out_served = s->h->completed;
    return CGVEC_OK;
}
CGVEC_EXPORT int cgvec_serve_set(cgvec_server* s, const char* key, int64_t value) {
    if (!s || !key) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    std::string k(key);
    if (k == "idle_us") { s->idle_us = (int)value; s->P.idle_ns = (uint64_t)value * 1000ull; }
    else if (k == "life_ms") { s->life_ms = (int)value; s->P.life_ns = (uint64_t)value * 1000000ull; }
    else if (k == "abort_ms") { s->abort_ms = (int)value; s->P.abort_ns = (uint64_t)value * 1000000ull; }
    else if (k == "wait_ms") s->wait_ms = (int)value;
    else if (k == "l2_hint") s->P.sp.use_l2_hint = (uint32_t)value;
    else if (k == "contig") s->P.contig = (uint32_t)value;
    else if (k == "max_inflight") s->max_inflight = value < 1 ? 1 : value > (int64_t)kServeSlots - 2 ? (int)kServeSlots - 2 : (int)value;
    else return fail(CGVEC_ERR_BAD_ARG, "unknown session option '%s'", key);
    return CGVEC_OK;
}
CGVEC_EXPORT int cgvec_serve_close(cgvec_server* s) {
    if (!s) return CGVEC_OK;
    if (s->ix) s->ix->open_streams--;
    serve_free(s);
    return CGVEC_OK;
}
