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
                 uint8_t (*out_ids)[16], float* out_scores, uint32_t* out_counts) {
    if (o.device_io) return fail(CGVEC_ERR_UNSUPPORTED, "device_io is not available on a multi-device index");
    if (o.formula != CGVEC_FORMULA_SIMD) return fail(CGVEC_ERR_UNSUPPORTED, "multi-device indexes serve the SIMD formula only");
    if (o.path == CGVEC_PATH_TENSOR) return fail(CGVEC_ERR_UNSUPPORTED, "multi-device indexes use the exact-order kernel");
    if (k > kXchgMaxK) return fail(CGVEC_ERR_UNSUPPORTED, "k = %u exceeds the peer-exchange limit of %u on a multi-device index", k, kXchgMaxK);
    if (mx->n == 0) { if (out_counts) for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0; return CGVEC_OK; }
    const size_t G = mx->parts.size();
    const uint32_t qstride = (mx->dim + 3) & ~3u;
    std::lock_guard<std::mutex> lk(mx->comm_mu);                 // one exchange sequence at a time
    std::vector<SearchCtx*> ctx(G, nullptr);
    auto finish = [&](int code) {
        for (size_t s = 0; s < G; ++s) {
            if (!ctx[s]) continue;
            cudaSetDevice(mx->parts[s]->device);
            cudaStreamSynchronize(ctx[s]->stream);
            ctx_release(mx->parts[s], ctx[s]);
        }
        return code;
    };
    for (size_t s = 0; s < G; ++s) {
        Index* p = mx->parts[s];
        CUDA_TRY(cudaSetDevice(p->device));
        int rc = ctx_acquire(p, &ctx[s]);
        if (rc) return finish(rc);
        SearchCtx* c = ctx[s];
        rc = ensure(&c->h_q, &c->hq_cap, (size_t)nq * qstride, true); if (rc) return finish(rc);
        rc = ensure(&c->d_q, &c->q_cap, (size_t)nq * qstride); if (rc) return finish(rc);
        for (uint32_t q = 0; q < nq; ++q) {
            memcpy(c->h_q + (size_t)q * qstride, queries + (size_t)q * mx->dim, mx->dim * sizeof(float));
            for (uint32_t i = mx->dim; i < qstride; ++i) c->h_q[(size_t)q * qstride + i] = 0.0f;
        }
        cudaError_t e = cudaMemcpyAsync(c->d_q, c->h_q, (size_t)nq * qstride * sizeof(float), cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) return finish(fail(CGVEC_ERR_CUDA, "H2D failed: %s", cudaGetErrorString(e)));
        p->searches++;
    }
    SearchCtx* c0 = ctx[0];
    {
        CUDA_TRY(cudaSetDevice(mx->parts[0]->device));
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
    }
    uint32_t q0 = 0;
    while (q0 < nq) {
        const uint32_t b = nq - q0 >= 4 && mx->opt_max_nq >= 4 ? 4 : (nq - q0 >= 2 && mx->opt_max_nq >= 2 ? 2 : 1);
        const uint32_t seq = ++mx->xseq;
        for (size_t s = 0; s < G; ++s) {
            Index* p = mx->parts[s];
            SearchCtx* c = ctx[s];
            CUDA_TRY(cudaSetDevice(p->device));
            int rc = ensure_smem_attr(xchg_merge_kernel, 160 * 1024);
            if (rc) return finish(rc);
            const uint64_t* partials = nullptr;
            uint32_t lists = 0;
            rc = local_exact(p, c, c->d_q + (size_t)q0 * qstride, b, k, o.metric, c->stream, nullptr, nullptr, nullptr, nullptr, &partials, &lists);
            if (rc) return finish(rc);
            XchgParams xp{};
            xp.partials = partials; xp.n_lists = lists; xp.k = k; xp.nq = b; xp.ascending = (o.metric == CGVEC_L2);
            xp.rank = (uint32_t)s; xp.world = (uint32_t)G; xp.seq = seq;
            for (size_t r = 0; r < G; ++r) xp.peer[r] = p->xpeer[r];
            xp.err = mx->h_xerr; xp.timeout_ns = (uint64_t)(mx->opt_xchg_timeout_ms > 0 ? mx->opt_xchg_timeout_ms : 5000) * 1000000ull;
            if (s == 0) { xp.out_rows = c0->d_rows + (size_t)q0 * k; xp.out_scores = c0->d_scores + (size_t)q0 * k; xp.out_counts = c0->d_counts + q0; }
            xchg_merge_kernel<<<b, kXchgThreads, ((size_t)lists * k + 9 * k) * 8, c->stream>>>(xp);
            p->launches++;
            CUDA_TRY(cudaGetLastError());
        }
        q0 += b;
    }
    CUDA_TRY(cudaSetDevice(mx->parts[0]->device));
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
