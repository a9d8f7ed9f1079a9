// stream_search.inl — streaming batches of host queries through an index with double-buffered upload
// (BASELINE config 5: "streaming 1k-query batches").  Included by cgvec_api.cu after the search entry points.
//
// The reference's closest shape is SemanticSearch::multi_vector_search (crates/codegraph-vector/src/search.rs:347-361):
// many query embeddings against one store.  Here a caller feeds batch after batch; cgvec_stream_submit(i+1) first
// stages batch i+1 (pinned copy + asynchronous H2D on a copy stream) and THEN runs the search of batch i on the run
// stream, so the upload of the next batch travels behind the scan of the current one.  Results lag one submit;
// cgvec_stream_flush drains the last batch.
struct cgvec_stream {
    Index* ix = nullptr;
    uint32_t max_batch = 0, k = 0, qstride = 0;
    int metric = CGVEC_COSINE, path = CGVEC_PATH_AUTO;
    cudaStream_t copy_st = nullptr, run_st = nullptr;
    float* h_q[2] = {nullptr, nullptr};
    float* d_q[2] = {nullptr, nullptr};
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
    uint32_t nq_staged[2] = {0, 0};
    int staged = -1, next_slot = 0;
    uint64_t *d_rows = nullptr, *h_rows = nullptr;
    float *d_scores = nullptr, *h_scores = nullptr;
    uint32_t *d_counts = nullptr, *h_counts = nullptr;
    uint64_t batches = 0;
};

static void stream_free(cgvec_stream* s) {
    if (!s) return;
    cudaSetDevice(s->ix->device);
    if (s->run_st) cudaStreamSynchronize(s->run_st);
    if (s->copy_st) cudaStreamSynchronize(s->copy_st);
    for (int i = 0; i < 2; ++i) { cudaFreeHost(s->h_q[i]); cudaFree(s->d_q[i]); if (s->ev_h2d[i]) cudaEventDestroy(s->ev_h2d[i]); }
    cudaFree(s->d_rows); cudaFree(s->d_scores); cudaFree(s->d_counts);
    cudaFreeHost(s->h_rows); cudaFreeHost(s->h_scores); cudaFreeHost(s->h_counts);
    if (s->copy_st) cudaStreamDestroy(s->copy_st);
    if (s->run_st) cudaStreamDestroy(s->run_st);
    delete s;
}

CGVEC_EXPORT int cgvec_stream_open(cgvec_index* ix, uint32_t max_batch, uint32_t k, cgvec_metric metric, cgvec_path path, cgvec_stream** out) {
    if (!ix || !out) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    *out = nullptr;
    if (max_batch == 0 || k == 0) return fail(CGVEC_ERR_BAD_ARG, "max_batch and k must be > 0");
    if (!ix->parts.empty()) return fail(CGVEC_ERR_UNSUPPORTED, "streams serve single-device and one-process-per-GPU indexes");
    if (k > kMaxK) return fail(CGVEC_ERR_UNSUPPORTED, "k = %u exceeds the fused top-k limit of %u", k, kMaxK);
    CUDA_TRY(cudaSetDevice(ix->device));
    std::unique_ptr<cgvec_stream, void (*)(cgvec_stream*)> s(new cgvec_stream(), stream_free);
    s->ix = ix; s->max_batch = max_batch; s->k = k; s->metric = metric; s->path = path;
    s->qstride = (ix->dim + 3) & ~3u;
    if (s->qstride != ix->dim) return fail(CGVEC_ERR_UNSUPPORTED, "streams need dim %% 4 == 0");
    CUDA_TRY(cudaStreamCreateWithFlags(&s->copy_st, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&s->run_st, cudaStreamNonBlocking));
    const size_t qn = (size_t)max_batch * s->qstride, on = (size_t)max_batch * k;
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&s->h_q[i]), qn * sizeof(float)));
        CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_q[i]), qn * sizeof(float)));
        CUDA_TRY(cudaEventCreateWithFlags(&s->ev_h2d[i], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_rows), on * sizeof(uint64_t)));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_scores), on * sizeof(float)));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&s->d_counts), max_batch * sizeof(uint32_t)));
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&s->h_rows), on * sizeof(uint64_t)));
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&s->h_scores), on * sizeof(float)));
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&s->h_counts), max_batch * sizeof(uint32_t)));
    ix->open_streams++;                                        // the write side refuses to run until the session is closed
    *out = s.release();
    return CGVEC_OK;
}

// runs the staged batch (if any) and hands its results out; *out_nq = 0 when nothing was staged
static int stream_run_staged(cgvec_stream* s, uint64_t* out_rows, float* out_scores, uint32_t* out_counts, uint32_t* out_nq) {
    if (out_nq) *out_nq = 0;
    if (s->staged < 0) return CGVEC_OK;
    Index* ix = s->ix;
    const int slot = s->staged;
    const uint32_t nq = s->nq_staged[slot];
    s->staged = -1;
    SearchCtx* c = nullptr;
    int rc = ctx_acquire(ix, &c);
    if (rc) return rc;
    cudaStreamWaitEvent(s->run_st, c->done, 0);
    cudaStreamWaitEvent(s->run_st, s->ev_h2d[slot], 0);
    ix->searches++;
    rc = run_queries(ix, c, s->d_q[slot], s->qstride, nq, s->k, s->metric, s->path, s->run_st, s->d_rows, s->d_scores, s->d_counts);
    cudaError_t e = cudaSuccess;
    if (!rc) {
        const size_t on = (size_t)nq * s->k;
        e = cudaMemcpyAsync(s->h_rows, s->d_rows, on * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->run_st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->h_scores, s->d_scores, on * sizeof(float), cudaMemcpyDeviceToHost, s->run_st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s->h_counts, s->d_counts, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, s->run_st);
    }
    cudaError_t e2 = cudaStreamSynchronize(s->run_st);
    cudaEventRecord(c->done, s->run_st);
    ctx_release(ix, c);
    if (rc) return rc;
    if (e != cudaSuccess || e2 != cudaSuccess) return fail(CGVEC_ERR_CUDA, "stream search failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    const size_t on = (size_t)nq * s->k;
    if (out_rows) memcpy(out_rows, s->h_rows, on * sizeof(uint64_t));
    if (out_scores) memcpy(out_scores, s->h_scores, on * sizeof(float));
    if (out_counts) memcpy(out_counts, s->h_counts, nq * sizeof(uint32_t));
    if (out_nq) *out_nq = nq;
    s->batches++;
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_stream_submit(cgvec_stream* s, const float* queries, uint32_t nq, uint64_t* out_rows, float* out_scores,
                                     uint32_t* out_counts, uint32_t* out_nq) {
    if (!s) return fail(CGVEC_ERR_BAD_ARG, "stream is NULL");
    if (out_nq) *out_nq = 0;
    if (nq > s->max_batch) return fail(CGVEC_ERR_BAD_ARG, "batch of %u queries exceeds the stream's max_batch %u", nq, s->max_batch);
    if (nq && !queries) return fail(CGVEC_ERR_BAD_ARG, "queries is NULL");
    CUDA_TRY(cudaSetDevice(s->ix->device));
    int new_slot = -1;
    if (nq) {                                                  // 1. stage the new batch: its upload overlaps the search below
        new_slot = s->next_slot;
        s->next_slot ^= 1;
        memcpy(s->h_q[new_slot], queries, (size_t)nq * s->qstride * sizeof(float));
        CUDA_TRY(cudaMemcpyAsync(s->d_q[new_slot], s->h_q[new_slot], (size_t)nq * s->qstride * sizeof(float), cudaMemcpyHostToDevice, s->copy_st));
        CUDA_TRY(cudaEventRecord(s->ev_h2d[new_slot], s->copy_st));
        s->nq_staged[new_slot] = nq;
    }
    int rc = stream_run_staged(s, out_rows, out_scores, out_counts, out_nq);   // 2. search the batch staged by the previous submit
    if (new_slot >= 0) s->staged = new_slot;
    return rc;
}

CGVEC_EXPORT int cgvec_stream_flush(cgvec_stream* s, uint64_t* out_rows, float* out_scores, uint32_t* out_counts, uint32_t* out_nq) {
    if (!s) return fail(CGVEC_ERR_BAD_ARG, "stream is NULL");
    CUDA_TRY(cudaSetDevice(s->ix->device));
    return stream_run_staged(s, out_rows, out_scores, out_counts, out_nq);
}

CGVEC_EXPORT int cgvec_stream_close(cgvec_stream* s) {
    if (s && s->ix) s->ix->open_streams--;
    stream_free(s);
    return CGVEC_OK;
}
