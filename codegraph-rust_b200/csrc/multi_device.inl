// multi_device.inl — single-process, multi-GPU index (cgvec_create with n_devices > 1): the deployment a Rust host
// process uses (one process, all GPUs of the box).  Included by cgvec_api.cu inside its anonymous namespace.
//
// The parent Index owns one shard Index per device.  Global rows are dealt to the shards in blocks of
// kMultiBlk rows round-robin (global row g -> block g / kMultiBlk -> shard block % G), so shards stay balanced
// under incremental adds and keys keep carrying global rows (tie rule survives the merge).  A search launches
// the exact-order scan on every device and then exchange.cuh's fused kernel on every device: each shard pushes
// its best-k into all peers' buffers over NVLink (direct peer access, no IPC needed inside one process) and
// every device ends with the global answer; the host reads device 0's copy.
constexpr uint32_t kMultiBlk = 1024;

inline void multi_locate(const Index* mx, uint64_t g, size_t* shard, uint64_t* local) {
    const uint64_t G = mx->parts.size(), b = g / kMultiBlk;
    *shard = (size_t)(b % G);
    *local = (b / G) * kMultiBlk + g % kMultiBlk;
}
// number of rows shard s holds when the index has n global rows
inline uint64_t multi_local_count(const Index* mx, size_t s, uint64_t n) {
    const uint64_t G = mx->parts.size(), full = n / kMultiBlk, rem = n % kMultiBlk;
    uint64_t c = (full / G) * kMultiBlk + ((full % G) > s ? kMultiBlk : 0);
    if (full % G == s) c += rem;
    return c;
}

int multi_create(uint32_t dim, cgvec_dtype storage, const int* device_ids, int n_devices, Index** out) {
    if (n_devices > (int)kXchgMaxWorld) return fail(CGVEC_ERR_UNSUPPORTED, "at most %u devices per index", kXchgMaxWorld);
    std::unique_ptr<cgvec_index> mx(new cgvec_index());
    mx->dim = dim; mx->dtype = storage; mx->esize = storage == CGVEC_F32 ? 4 : 2;
    const uint32_t align_elems = 16 / mx->esize;
    mx->ld = (dim + align_elems - 1) / align_elems * align_elems;
    auto cleanup = [&](int code) {
        for (Index* p : mx->parts) {
            cudaSetDevice(p->device);
            for (auto* c : p->pool) ctx_free(c);
            cudaFree(p->xbuf); cudaFree(p->d_rows); cudaFree(p->d_norms);
            if (p->main_stream) cudaStreamDestroy(p->main_stream);
            delete static_cast<cgvec_index*>(p);
        }
        mx->parts.clear();
        return code;
    };
    for (int i = 0; i < n_devices; ++i) {
        const int dev = device_ids ? device_ids[i] : i;
        for (int j = 0; j < i; ++j)
            if ((device_ids ? device_ids[j] : j) == dev) return cleanup(fail(CGVEC_ERR_BAD_ARG, "device %d listed twice", dev));
        int rc = check_device(dev);
        if (rc) return cleanup(rc);
        cudaError_t e = cudaSetDevice(dev);
        Index* p = new cgvec_index();
        mx->parts.push_back(p);
        p->dim = dim; p->dtype = storage; p->esize = mx->esize; p->ld = mx->ld; p->device = dev;
        p->rank = i; p->world = n_devices; p->blk_rows = kMultiBlk; p->n_shards = (uint32_t)n_devices; p->shard_id = (uint32_t)i;
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->main_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&p->xbuf), kXchgBytes);
        if (e == cudaSuccess) e = cudaMemset(p->xbuf, 0, kXchgBytes);
        if (e != cudaSuccess) return cleanup(fail(CGVEC_ERR_CUDA, "device %d setup failed: %s", dev, cudaGetErrorString(e)));
    }
    for (Index* a : mx->parts) {
        cudaSetDevice(a->device);
        for (Index* b : mx->parts) {
            a->xpeer[b->rank] = b->xbuf;
            if (a == b) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a->device, b->device);
            if (!can) return cleanup(fail(CGVEC_ERR_UNSUPPORTED, "device %d cannot access device %d's memory (no NVLink/P2P)", a->device, b->device));
            cudaError_t e = cudaDeviceEnablePeerAccess(b->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cleanup(fail(CGVEC_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)));
            cudaGetLastError();
        }
        a->p2p = true;
    }
    mx->device = mx->parts[0]->device;
    mx->sm_count = mx->parts[0]->sm_count;
    if (cudaHostAlloc(reinterpret_cast<void**>(&mx->h_xerr), sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess)
        return cleanup(fail(CGVEC_ERR_OOM, "pinned error word"));
    *mx->h_xerr = 0;
    *out = mx.release();
    return CGVEC_OK;
}

void multi_destroy(Index* mx) {
    for (Index* p : mx->parts) {
        cudaSetDevice(p->device);
        cudaDeviceSynchronize();
        drain_timings(p);
        for (auto* c : p->pool) ctx_free(c);
        cudaFree(p->xbuf); cudaFree(p->d_rows); cudaFree(p->d_norms);
        if (p->main_stream) cudaStreamDestroy(p->main_stream);
        delete static_cast<cgvec_index*>(p);
    }
    cudaFreeHost(mx->h_xerr);
    delete static_cast<cgvec_index*>(mx);
}

int multi_reserve(Index* mx, uint64_t n_rows) {
    for (size_t s = 0; s < mx->parts.size(); ++s) {
        Index* p = mx->parts[s];
        CUDA_TRY(cudaSetDevice(p->device));
        int rc = grow(p, multi_local_count(mx, s, n_rows), /*exact=*/true);
        if (rc) return rc;
    }
    return CGVEC_OK;
}

// copies one source row run [i, j) whose targets are the consecutive global rows [g, g + (j-i)) into the shards
int multi_copy_run(Index* mx, const uint8_t* src, size_t src_pitch, uint64_t g, uint64_t count) {
    const size_t dst_pitch = (size_t)mx->ld * mx->esize;
    uint64_t done = 0;
    while (done < count) {
        size_t s; uint64_t local;
        multi_locate(mx, g + done, &s, &local);
        uint64_t seg = kMultiBlk - (g + done) % kMultiBlk;
        if (seg > count - done) seg = count - done;
        Index* p = mx->parts[s];
        CUDA_TRY(cudaSetDevice(p->device));
        int rc = grow(p, local + seg);
        if (rc) return rc;
        uint8_t* dst = static_cast<uint8_t*>(p->d_rows) + local * dst_pitch;
        if (dst_pitch != src_pitch) CUDA_TRY(cudaMemset2DAsync(dst, dst_pitch, 0, dst_pitch, seg, p->main_stream));
        CUDA_TRY(cudaMemcpy2DAsync(dst, dst_pitch, src + done * src_pitch, src_pitch, src_pitch, seg, cudaMemcpyHostToDevice, p->main_stream));
        rc = launch_norms(p, local, seg, p->main_stream);
        if (rc) return rc;
        if (local + seg > p->n) p->n = local + seg;
        done += seg;
    }
    return CGVEC_OK;
}

int multi_sync(Index* mx) {
    for (Index* p : mx->parts) {
        CUDA_TRY(cudaSetDevice(p->device));
        CUDA_TRY(cudaStreamSynchronize(p->main_stream));
    }
    return CGVEC_OK;
}

int multi_add(Index* mx, const uint8_t (*ids)[16], const void* rows, uint64_t n, uint32_t src_esize) {
    if (src_esize != mx->esize)
        return fail(CGVEC_ERR_BAD_ARG, "index stores %s rows; use %s", mx->esize == 4 ? "f32" : "f16", mx->esize == 4 ? "cgvec_add" : "cgvec_add_f16");
    std::vector<uint64_t> target(n);
    std::unordered_map<IdKey, uint64_t, IdHash> staged;            // committed only after every copy succeeded (see add_impl)
    uint64_t next = mx->n;
    for (uint64_t i = 0; i < n; ++i) {
        if (ids) {
            IdKey key = id_key(ids[i]);
            auto it = mx->id2row.find(key);
            if (it != mx->id2row.end()) { target[i] = it->second; continue; }
            auto st = staged.find(key);
            if (st != staged.end()) { target[i] = st->second; continue; }
            staged.emplace(key, next);
        }
        target[i] = next++;
    }
    if (next > 0xfffffffeull) return fail(CGVEC_ERR_UNSUPPORTED, "more than 2^32-2 rows are not supported");
    mx->ids.resize(next * 16, 0);
    mx->has_id.resize(next, 0);
    const size_t src_pitch = (size_t)mx->dim * mx->esize;
    const uint8_t* src = static_cast<const uint8_t*>(rows);
    uint64_t i = 0;
    while (i < n) {
        uint64_t j = i + 1;
        while (j < n && target[j] == target[j - 1] + 1) ++j;
        int rc = multi_copy_run(mx, src + i * src_pitch, src_pitch, target[i], j - i);
        if (rc) return rc;
        if (ids) for (uint64_t r = i; r < j; ++r) { memcpy(&mx->ids[target[r] * 16], ids[r], 16); mx->has_id[target[r]] = 1; }
        i = j;
    }
    int rc = multi_sync(mx);
    if (rc) return rc;
    for (auto& kv : staged) mx->id2row.emplace(kv.first, kv.second);
    mx->n = next;
    return CGVEC_OK;
}

int multi_fill_synthetic(Index* mx, uint64_t n, uint64_t seed, int unit_norm) {
    const uint64_t next = mx->n + n;
    if (next > 0xfffffffeull) return fail(CGVEC_ERR_UNSUPPORTED, "more than 2^32-2 rows are not supported");
    const int threads = 256;
    for (size_t s = 0; s < mx->parts.size(); ++s) {
        Index* p = mx->parts[s];
        CUDA_TRY(cudaSetDevice(p->device));
        const uint64_t lo = multi_local_count(mx, s, mx->n), hi = multi_local_count(mx, s, next);
        if (hi == lo) continue;
        int rc = grow(p, hi);
        if (rc) return rc;
        ScanParams map = map_params(p);
        const uint64_t chunk = 1ull << 22;
        for (uint64_t done = lo; done < hi; done += chunk) {
            uint64_t cnt = hi - done < chunk ? hi - done : chunk;
            uint64_t blocks = (cnt * 8 + threads - 1) / threads;
            if (p->dtype == CGVEC_F32) synth_rows_kernel<float><<<(unsigned)blocks, threads, 0, p->main_stream>>>(static_cast<float*>(p->d_rows), done, cnt, p->dim, p->ld, seed, unit_norm, map);
            else synth_rows_kernel<__half><<<(unsigned)blocks, threads, 0, p->main_stream>>>(static_cast<__half*>(p->d_rows), done, cnt, p->dim, p->ld, seed, unit_norm, map);
            p->launches++;
            CUDA_TRY(cudaGetLastError());
            rc = launch_norms(p, done, cnt, p->main_stream);
            if (rc) return rc;
        }
        p->n = hi;
    }
    int rc = multi_sync(mx);
    if (rc) return rc;
    mx->ids.resize(next * 16, 0);
    mx->has_id.resize(next, 0);
    mx->n = next;
    return CGVEC_OK;
}

int multi_get_rows(Index* mx, uint64_t first, uint64_t n, float* out) {
    if (first + n > mx->n) return fail(CGVEC_ERR_NOT_FOUND, "rows [%llu, %llu) out of range", (unsigned long long)first, (unsigned long long)(first + n));
    uint64_t done = 0;
    while (done < n) {
        size_t s; uint64_t local;
        multi_locate(mx, first + done, &s, &local);
        uint64_t seg = kMultiBlk - (first + done) % kMultiBlk;
        if (seg > n - done) seg = n - done;
        int rc = cgvec_get_rows(static_cast<const cgvec_index*>(mx->parts[s]), local, seg, out + done * mx->dim);
        if (rc) return rc;
        done += seg;
    }
    return CGVEC_OK;
}

int multi_search(Index* mx, const float* queries, uint32_t nq, uint32_t k, const cgvec_search_opts& o, uint64_t* out_rows,
                 uint8_t (*out_ids)[16], float* out_scores, uint32_t* out_counts);
int multi_rescore(Index* mx, const float* query, const uint64_t* rows, uint32_t n, cgvec_metric metric, cgvec_formula formula, float* out);

// Non-SIMD cosine formulas (scalar / sequential / 1 - cos distance) on a multi-device index: the same contract as
// search_formula() on one device — over-fetch under the SIMD order on every device, re-score the merged candidates in the
// formula's own arithmetic on the devices that hold them, re-rank, and prove that no row outside the candidate set can enter
// the top-k (|formula - simd| <= eps for every row); widen and retry otherwise.
int multi_search_formula(Index* mx, const float* queries, uint32_t nq, uint32_t k, const cgvec_search_opts& o, uint64_t* out_rows,
                         uint8_t (*out_ids)[16], float* out_scores, uint32_t* out_counts) {
    if (o.device_io) return fail(CGVEC_ERR_UNSUPPORTED, "device_io supports the SIMD formula only");
    const int ascending = (o.formula == CGVEC_FORMULA_BASELINE);
    const uint64_t n = mx->n;
    const uint32_t want = (uint32_t)(k < n ? k : n);
    if (k > kMaxK) return fail(CGVEC_ERR_UNSUPPORTED, "k = %u exceeds the fused top-k limit of %u", k, kMaxK);
    const float eps = 4.0f * (float)(mx->dim + 8) * 5.9604645e-8f;
    cgvec_search_opts so = o;
    so.formula = CGVEC_FORMULA_SIMD; so.path = CGVEC_PATH_EXACT; so.metric = CGVEC_COSINE;
    std::vector<uint64_t> cand_rows(kMaxK);
    std::vector<float> cand_simd(kMaxK), cand_new(kMaxK);
    std::vector<uint32_t> order;
    for (uint32_t q = 0; q < nq; ++q) {
        const float* qv = queries + (size_t)q * mx->dim;
        if (out_counts) out_counts[q] = want;
        if (want == 0) continue;
        uint32_t kp = want + (want / 4 > 16 ? want / 4 : 16);
        while (true) {
            if (kp > n) kp = (uint32_t)n;
            if (kp > kMaxK) kp = kMaxK;
            uint32_t cnt = 0;
            int rc = multi_search(mx, qv, 1, kp, so, cand_rows.data(), nullptr, cand_simd.data(), &cnt);
            if (rc) return rc;
            if (cnt < kp) kp = cnt;
            rc = multi_rescore(mx, qv, cand_rows.data(), kp, CGVEC_COSINE, (cgvec_formula)o.formula, cand_new.data());
            if (rc) return rc;
            order.resize(kp);
            for (uint32_t i = 0; i < kp; ++i) order[i] = i;
            std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
                const float x = cand_new[a], y = cand_new[b];
                const bool xn = std::isnan(x), yn = std::isnan(y);
                if (xn != yn) return yn;
                if (!xn && x != y) return ascending ? x < y : x > y;
                return cand_rows[a] < cand_rows[b];
            });
            bool proven = (kp >= n);
            if (!proven) {
                const float tau = cand_simd[kp - 1], kth = cand_new[order[want - 1]];
                if (!std::isnan(tau) && !std::isnan(kth)) proven = ascending ? kth < (1.0f - tau) - eps : kth > tau + eps;
            }
            if (proven) break;
            if (kp >= kMaxK) return fail(CGVEC_ERR_UNSUPPORTED, "could not separate the top-%u under formula %d within %u candidates", k, (int)o.formula, kp);
            kp *= 2;
        }
        for (uint32_t i = 0; i < k; ++i) {
            const size_t oi = (size_t)q * k + i;
            const bool valid = i < want;
            const uint64_t grow = valid ? cand_rows[order[i]] : ~0ull;
            if (out_rows) out_rows[oi] = grow;
            if (out_scores) out_scores[oi] = valid ? cand_new[order[i]] : 0.0f;
            if (out_ids) {
                memset(out_ids[oi], 0, 16);
                if (valid && grow < mx->n && mx->has_id[grow]) memcpy(out_ids[oi], &mx->ids[grow * 16], 16);
            }
        }
    }
    return CGVEC_OK;
}

// Search of a single-process multi-device index.  Every batch is scanned on all devices at once (exact-order kernel or the
// tensor path, one host thread per device when the path synchronises internally), each device reduces its shard to its best k
// keys, and the shards meet on the first device:
//   * k <= 128 and <= 4 queries (the batch-1 serving shape): the fused peer-memory exchange kernel (exchange.cuh) on every device;
//   * anything else (k up to 1024, tensor batches): each device's [nq][k] keys are pulled into the first device with
//     cudaMemcpyPeerAsync behind per-device events and merged there (no NCCL in this deployment).
// Host I/O or device I/O (queries / results in the FIRST device's memory, asynchronous on the caller's stream).
int multi_search(Index* mx, const float* queries, uint32_t nq, uint32_t k, const cgvec_search_opts& o, uint64_t* out_rows,
                 uint8_t (*out_ids)[16], float* out_scores, uint32_t* out_counts) {
    NvtxRange nvtx_("cgvec.multi_search");
    if (o.formula != CGVEC_FORMULA_SIMD) return multi_search_formula(mx, queries, nq, k, o, out_rows, out_ids, out_scores, out_counts);
    if (k > kMaxK) return fail(CGVEC_ERR_UNSUPPORTED, "k = %u exceeds the fused top-k limit of %u", k, kMaxK);
    const size_t G = mx->parts.size();
    Index* p0 = mx->parts[0];
    cudaStream_t user_st = (cudaStream_t)o.stream;
    if (mx->n == 0) {
        if (o.device_io) { CUDA_TRY(cudaSetDevice(p0->device)); if (out_counts) CUDA_TRY(cudaMemsetAsync(out_counts, 0, nq * sizeof(uint32_t), user_st)); }
        else if (out_counts) for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
        return CGVEC_OK;
    }
    if (mx->h_xerr && *mx->h_xerr) return fail(CGVEC_ERR_NCCL, "an earlier peer exchange timed out waiting for device %u", *mx->h_xerr - 1);
    const uint32_t qstride = (mx->dim + 3) & ~3u;
    if (o.device_io && qstride != mx->dim) return fail(CGVEC_ERR_UNSUPPORTED, "device_io needs dim %% 4 == 0");
    uint64_t min_n = ~0ull;
    for (Index* p : mx->parts) min_n = std::min<uint64_t>(min_n, p->n);
    std::lock_guard<std::mutex> lk(mx->comm_mu);                 // one exchange sequence at a time
    std::vector<SearchCtx*> ctx(G, nullptr);
    auto finish = [&](int code) {
        for (size_t s = 0; s < G; ++s) {
            if (!ctx[s]) continue;
            cudaSetDevice(mx->parts[s]->device);
            if (!o.device_io || code != CGVEC_OK) cudaStreamSynchronize(ctx[s]->stream);
            cudaEventRecord(ctx[s]->done, ctx[s]->stream);
            ctx_release(mx->parts[s], ctx[s]);
        }
        return code;
    };
    // ---- queries onto every device
    cudaEvent_t q_ready = nullptr;                               // device_io: the caller's stream has produced the queries
    if (o.device_io) {
        CUDA_TRY(cudaSetDevice(p0->device));
        CUDA_TRY(cudaEventCreateWithFlags(&q_ready, cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(q_ready, user_st));
    }
    for (size_t s = 0; s < G; ++s) {
        Index* p = mx->parts[s];
        CUDA_TRY(cudaSetDevice(p->device));
        int rc = ctx_acquire(p, &ctx[s]);
        if (rc) return finish(rc);
        SearchCtx* c = ctx[s];
        cudaStreamWaitEvent(c->stream, c->done, 0);
        rc = ensure(&c->d_q, &c->q_cap, (size_t)nq * qstride); if (rc) return finish(rc);
        cudaError_t e;
        if (o.device_io) {
            cudaStreamWaitEvent(c->stream, q_ready, 0);
            e = cudaMemcpyPeerAsync(c->d_q, p->device, queries, p0->device, (size_t)nq * qstride * sizeof(float), c->stream);
        } else {
            rc = ensure(&c->h_q, &c->hq_cap, (size_t)nq * qstride, true); if (rc) return finish(rc);
            for (uint32_t q = 0; q < nq; ++q) {
                memcpy(c->h_q + (size_t)q * qstride, queries + (size_t)q * mx->dim, mx->dim * sizeof(float));
                for (uint32_t i = mx->dim; i < qstride; ++i) c->h_q[(size_t)q * qstride + i] = 0.0f;
            }
            e = cudaMemcpyAsync(c->d_q, c->h_q, (size_t)nq * qstride * sizeof(float), cudaMemcpyHostToDevice, c->stream);
        }
        if (e != cudaSuccess) return finish(fail(CGVEC_ERR_CUDA, "query upload failed: %s", cudaGetErrorString(e)));
        p->searches++;
    }
    if (q_ready) cudaEventDestroy(q_ready);
    // ---- outputs on the first device
    SearchCtx* c0 = ctx[0];
    uint64_t* d_rows = out_rows; float* d_scores = out_scores; uint32_t* d_counts = out_counts;
    CUDA_TRY(cudaSetDevice(p0->device));
    if (!o.device_io) {
        size_t oc = c0->out_cap, oc2 = c0->out_cap;
        int rc = ensure(&c0->d_rows, &oc, (size_t)nq * k); if (rc) return finish(rc);
        rc = ensure(&c0->d_scores, &oc2, (size_t)nq * k); if (rc) return finish(rc);
        c0->out_cap = oc < oc2 ? oc : oc2;
        rc = ensure(&c0->d_counts, &c0->cnt_cap, nq); if (rc) return finish(rc);
        size_t hc = c0->hout_cap, hc2 = c0->hout_cap;
        rc = ensure(&c0->h_rows, &hc, (size_t)nq * k, true); if (rc) return finish(rc);
        rc = ensure(&c0->h_scores, &hc2, (size_t)nq * k, true); if (rc) return finish(rc);
        c0->hout_cap = hc < hc2 ? hc : hc2;
        rc = ensure(&c0->h_counts, &c0->hcnt_cap, nq, true); if (rc) return finish(rc);
        d_rows = c0->d_rows; d_scores = c0->d_scores; d_counts = c0->d_counts;
    }
    // ---- kernel family (the same on every device: decided on the smallest shard)
    bool tensor = false;
    if (o.path == CGVEC_PATH_TENSOR) {
        if (o.metric != CGVEC_COSINE) return finish(fail(CGVEC_ERR_UNSUPPORTED, "the tensor-core path serves the cosine metric"));
        tensor = true;
    } else if (o.path == CGVEC_PATH_AUTO) {
        tensor = o.metric == CGVEC_COSINE && tensor_auto_rule(p0, min_n, nq, k, tc_batch_limit(p0, nq));
    }
    uint32_t n_max = tensor ? tc_batch_limit(p0, nq) : 0;
    if (tensor && n_max == 0) {
        if (o.path == CGVEC_PATH_TENSOR) return finish(fail(CGVEC_ERR_UNSUPPORTED, "dimension %u leaves no room for a resident query block", mx->dim));
        tensor = false;
    }
    const int ascending = (o.metric == CGVEC_L2);
    std::vector<cudaEvent_t> ev(G, nullptr);
    auto free_events = [&] { for (auto& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; } };
    uint32_t q0 = 0;
    while (q0 < nq) {
        uint32_t b = tensor ? std::min(nq - q0, n_max)
                            : (nq - q0 >= 4 && mx->opt_max_nq >= 4 ? 4u : (nq - q0 >= 2 && mx->opt_max_nq >= 2 ? 2u : 1u));
        const bool fused = !tensor && k <= kXchgMaxK && b <= kXchgMaxQ && mx->opt_p2p;
        if (fused) {
            const uint32_t seq = ++mx->xseq;
            for (size_t s = 0; s < G; ++s) {
                Index* p = mx->parts[s];
                SearchCtx* c = ctx[s];
                CUDA_TRY(cudaSetDevice(p->device));
                int rc = ensure_smem_attr(xchg_merge_kernel, 160 * 1024);
                if (rc) { free_events(); return finish(rc); }
                const uint64_t* partials = nullptr;
                uint32_t lists = 0;
                rc = local_exact(p, c, c->d_q + (size_t)q0 * qstride, b, k, o.metric, c->stream, nullptr, nullptr, nullptr, nullptr, &partials, &lists);
                if (rc) { free_events(); return finish(rc); }
                XchgParams xp{};
                xp.partials = partials; xp.n_lists = lists; xp.k = k; xp.nq = b; xp.ascending = ascending;
                xp.rank = (uint32_t)s; xp.world = (uint32_t)G; xp.seq = seq;
                for (size_t r = 0; r < G; ++r) xp.peer[r] = p->xpeer[r];
                xp.err = mx->h_xerr; xp.timeout_ns = (uint64_t)(mx->opt_xchg_timeout_ms > 0 ? mx->opt_xchg_timeout_ms : 5000) * 1000000ull;
                if (s == 0) { xp.out_rows = d_rows + (size_t)q0 * k; xp.out_scores = d_scores + (size_t)q0 * k; xp.out_counts = d_counts + q0; }
                xchg_merge_kernel<<<b, kXchgThreads, ((size_t)lists * k + 9 * k) * 8, c->stream>>>(xp);
                p->launches++;
                CUDA_TRY(cudaGetLastError());
            }
        } else {
            // each device reduces its shard to [b][k] keys (own gather slot), the first device pulls them in and merges
            const size_t per = (size_t)b * k;
            std::vector<uint64_t*> local_keys(G, nullptr);
            for (size_t s = 0; s < G; ++s) {
                CUDA_TRY(cudaSetDevice(mx->parts[s]->device));
                int rc = ensure_gather(mx->parts[s], ctx[s], b, k, &local_keys[s]);
                if (rc) { free_events(); return finish(rc); }
                if (!ev[s]) CUDA_TRY(cudaEventCreateWithFlags(&ev[s], cudaEventDisableTiming));
            }
            std::vector<int> rcs(G, CGVEC_OK);
            std::vector<std::string> errs(G);
            auto run_part = [&](size_t s) {
                Index* p = mx->parts[s];
                SearchCtx* c = ctx[s];
                if (cudaSetDevice(p->device) != cudaSuccess) { rcs[s] = CGVEC_ERR_CUDA; errs[s] = "cudaSetDevice failed"; return; }
                const float* dq = c->d_q + (size_t)q0 * qstride;
                rcs[s] = tensor ? local_tensor(p, c, dq, qstride, b, k, c->stream, local_keys[s], nullptr, nullptr, nullptr)
                                : local_exact(p, c, dq, b, k, o.metric, c->stream, local_keys[s], nullptr, nullptr, nullptr);
                if (rcs[s]) errs[s] = g_err;                     // the message is thread-local
                else cudaEventRecord(ev[s], c->stream);
            };
            if (tensor) {                                        // local_tensor synchronises (proof flags): one host thread per device
                std::vector<std::thread> th;
                for (size_t s = 1; s < G; ++s) th.emplace_back(run_part, s);
                run_part(0);
                for (auto& t : th) t.join();
            } else {
                for (size_t s = 0; s < G; ++s) run_part(s);
            }
            for (size_t s = 0; s < G; ++s) if (rcs[s]) { free_events(); return finish(fail(rcs[s], "%s", errs[s].c_str())); }
            CUDA_TRY(cudaSetDevice(p0->device));
            for (size_t s = 0; s < G; ++s) {
                cudaStreamWaitEvent(c0->stream, ev[s], 0);
                CUDA_TRY(cudaMemcpyPeerAsync(c0->d_gather + s * per, p0->device, local_keys[s], mx->parts[s]->device, per * sizeof(uint64_t), c0->stream));
            }
            int rc = merge_lists(p0, c0, c0->d_gather, b, (uint32_t)G, k, ascending, nullptr, d_rows + (size_t)q0 * k, d_scores + (size_t)q0 * k,
                                 d_counts + q0, c0->stream, (size_t)k, per);
            if (rc) { free_events(); return finish(rc); }
            // the next batch may overwrite the peers' gather slots only after the first device has pulled them
            if (q0 + b < nq) {
                CUDA_TRY(cudaEventRecord(ev[0], c0->stream));
                for (size_t s = 1; s < G; ++s) { cudaSetDevice(mx->parts[s]->device); cudaStreamWaitEvent(ctx[s]->stream, ev[0], 0); }
            }
        }
        q0 += b;
    }
    CUDA_TRY(cudaSetDevice(p0->device));
    if (o.device_io) {
        // results are on the first device in the caller's buffers: order the caller's stream behind our work and return
        if (!ev[0]) CUDA_TRY(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
        CUDA_TRY(cudaEventRecord(ev[0], c0->stream));
        CUDA_TRY(cudaStreamWaitEvent(user_st, ev[0], 0));
        free_events();
        return finish(CGVEC_OK);
    }
    free_events();
    cudaError_t e = cudaMemcpyAsync(c0->h_rows, c0->d_rows, (size_t)nq * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, c0->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c0->h_scores, c0->d_scores, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, c0->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c0->h_counts, c0->d_counts, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, c0->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c0->stream);
    if (e != cudaSuccess) return finish(fail(CGVEC_ERR_CUDA, "multi-device search failed: %s", cudaGetErrorString(e)));
    if (mx->h_xerr && *mx->h_xerr) return finish(fail(CGVEC_ERR_NCCL, "peer exchange timed out waiting for device %u", *mx->h_xerr - 1));
    for (uint32_t q = 0; q < nq; ++q) {
        const uint32_t cnt = c0->h_counts[q];
        if (out_counts) out_counts[q] = cnt;
        for (uint32_t i = 0; i < k; ++i) {
            const size_t oi = (size_t)q * k + i;
            const bool valid = i < cnt;
            const uint64_t grow = valid ? c0->h_rows[oi] : ~0ull;
            if (out_rows) out_rows[oi] = grow;
            if (out_scores) out_scores[oi] = valid ? c0->h_scores[oi] : 0.0f;
            if (out_ids) {
                memset(out_ids[oi], 0, 16);
                if (valid && grow < mx->n && mx->has_id[grow]) memcpy(out_ids[oi], &mx->ids[grow * 16], 16);
            }
        }
    }
    return finish(CGVEC_OK);
}

int multi_rescore(Index* mx, const float* query, const uint64_t* rows, uint32_t n, cgvec_metric metric, cgvec_formula formula, float* out) {
    std::vector<std::vector<uint64_t>> local(mx->parts.size());
    std::vector<std::vector<uint32_t>> pos(mx->parts.size());
    for (uint32_t i = 0; i < n; ++i) {
        if (rows[i] >= mx->n) return fail(CGVEC_ERR_NOT_FOUND, "row %llu out of range", (unsigned long long)rows[i]);
        size_t s; uint64_t l;
        multi_locate(mx, rows[i], &s, &l);
        local[s].push_back(l); pos[s].push_back(i);
    }
    for (size_t s = 0; s < mx->parts.size(); ++s) {
        if (local[s].empty()) continue;
        std::vector<float> tmp(local[s].size());
        int rc = cgvec_rescore(static_cast<const cgvec_index*>(mx->parts[s]), query, local[s].data(), (uint32_t)local[s].size(), metric, formula, tmp.data());
        if (rc) return rc;
        for (size_t j = 0; j < tmp.size(); ++j) out[pos[s][j]] = tmp[j];
    }
    return CGVEC_OK;
}
