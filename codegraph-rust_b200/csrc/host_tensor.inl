// host_tensor.inl — tensor-core path orchestration (included by cgvec_api.cu inside its anonymous namespace):
// TMA descriptors, kernel choice (resident query block vs CTA pairs), the geometric row ranges with threshold refresh,
// exact re-score + proof + exact-kernel fallback (local_tensor), batching (run_queries), scan timing read-out and the
// CUDA-IPC peer mapping of the exchange buffers.
// ---- tensor-core batched path (K2) -----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// Tensor map over a row-major matrix of 128-byte K-blocks (64 halves or 32 floats), SWIZZLE_128B, OOB reads as zero.
//   box4d == 0: 2-D {K, rows}, box = one K-block x box_rows                        (any row pitch that is a multiple of 16 bytes)
//   box4d == 1: 4-D {128 B, 8 rows, K-blocks, row groups}, box = kbs K-blocks x box_rows: lands directly as kbs consecutive
//               K-major swizzle atoms per 8-row group (UMMA SBO = kbs * 1024).  Needs row pitch % 128 == 0 (nkb * 128 <= pitch)
//               and 7 rows of readable slack behind the last row (grow() provides it).
int make_tmap_rows(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes, uint32_t box_rows,
                   int l2promo, bool f32, uint32_t kbs, bool box4d) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return fail(CGVEC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint32_t kbe = (cuuint32_t)(f32 ? kTcKBlock / 2 : kTcKBlock);       // 128 bytes of K either way
    const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const CUtensorMapL2promotion promo = l2promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : l2promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    CUresult r;
    if (box4d) {
        const uint64_t nkb = (inner + kbe - 1) / kbe;
        cuuint64_t gdim[4] = {kbe, 8, nkb, (rows + 7) / 8};
        cuuint64_t gstr[3] = {row_stride_bytes, 128, 8 * row_stride_bytes};
        cuuint32_t box[4] = {kbe, 8, kbs, box_rows / 8};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        r = fn(map, dt, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t gdim[2] = {inner, rows};
        cuuint64_t gstr[1] = {row_stride_bytes};
        cuuint32_t box[2] = {kbe, box_rows};
        cuuint32_t estr[2] = {1, 1};
        r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(CGVEC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return CGVEC_OK;
}

constexpr uint32_t kTcCap = 8192;     // candidate slots per query between selects (one CTA sorts them in smem)

bool tensor_path_applicable(const Index* ix, int metric, uint32_t nq) {
    return metric == CGVEC_COSINE && nq >= 1 && ix->n >= 1;      // f16 rows -> kind::f16, f32 rows -> kind::tf32
}

// Launch plan of the tensor kernels for a batch padded to N columns: kernel (resident query block vs CTA pairs), K-blocks
// per ring stage and ring depth.  Measured (tools/tma_stream_bench.cu, profiles/r02_tma_stream_pipe.txt): one commit per
// 16 KB K-block caps the ring at 5.2 TB/s whatever its depth, two or four K-blocks per stage reach 6.8-7.1 TB/s.
struct TcPlan { bool pairs; uint32_t N, nkb, kbs, groups, stages, box4d, smem; };
bool tc_plan_fill(const Index* ix, bool pairs, uint32_t N, TcPlan* out, uint32_t sel_cap = 2048) {
    const uint32_t kbe = ix->dtype == CGVEC_F32 ? kTcKBlock / 2 : kTcKBlock;
    const uint32_t nkb = (ix->dim + kbe - 1) / kbe;
    const bool can4d = ((size_t)ix->ld * ix->esize) % 128 == 0 && ix->opt_tc_kbs != 1 && nkb > 1;
    const uint32_t try_kbs[3] = {ix->opt_tc_kbs > 0 ? (uint32_t)ix->opt_tc_kbs : 2u, 2u, 1u};
    for (int t = 0; t < 3; ++t) {
        uint32_t kbs = can4d ? try_kbs[t] : 1u;
        if (kbs > nkb) kbs = nkb;
        if (kbs < 1) kbs = 1;
        const uint32_t min_stages = pairs ? 2u : 3u;
        uint32_t max_stages = ix->opt_tc_stages ? (uint32_t)ix->opt_tc_stages : (kbs >= 4 ? 3u : kbs >= 2 ? 6u : 12u);
        if (max_stages > kTcMaxStages) max_stages = kTcMaxStages;
        for (uint32_t s = max_stages; s >= min_stages; --s) {
            const uint32_t total = (pairs ? tc2_smem_layout(N, s, kbs, sel_cap).total : tc_smem_layout(N, nkb, s, kbs, sel_cap).total) + 1024;
            if (total <= kSmemBudget) {
                out->pairs = pairs; out->N = N; out->nkb = nkb; out->kbs = kbs; out->groups = (nkb + kbs - 1) / kbs; out->stages = s;
                out->box4d = kbs > 1 ? 1u : 0u; out->smem = total;
                return true;
            }
        }
        if (!can4d) break;
    }
    return false;
}
// Largest MMA N (multiple of 16) whose resident query block still leaves a ring in shared memory (tc_scan_kernel).
uint32_t tc_max_n(const Index* ix) {
    uint32_t limit = (uint32_t)ix->opt_tc_max_n;
    if (limit > kTcMaxN) limit = kTcMaxN;
    TcPlan pl;
    for (uint32_t N = limit & ~15u; N >= 16; N -= 16)
        if (tc_plan_fill(ix, false, N, &pl)) return N;
    return 0;
}
// which kernel serves a batch of nq queries, and how many queries one pass may take
bool tc_use_pairs(const Index* ix, uint32_t nq) {
    if (ix->opt_tc_kernel == 1) return false;
    if (ix->opt_tc_kernel == 2) return true;
    return nq > tc_max_n(ix);
}
uint32_t tc_batch_limit(const Index* ix, uint32_t nq) {
    if (tc_use_pairs(ix, nq)) { uint32_t m = (uint32_t)ix->opt_tc2_max_n & ~15u; return m >= 16 && m <= kTc2MaxN ? m : kTc2MaxN; }
    return tc_max_n(ix);
}

// Tensor-core scan of the local shard for `nq` <= N_max queries (f32, on the device, stride qstride): approximate
// ordering on tcgen05, exact re-score of the kp survivors per query, proof of exactness; unproven queries are
// re-run on the exact-order kernel.  Produces this shard's exact best-k keys (`local_keys` [nq][k]) or decoded results.
int local_tensor(Index* ix, SearchCtx* c, const float* d_q, uint32_t qstride, uint32_t nq, uint32_t k, cudaStream_t st,
                 uint64_t* local_keys, uint64_t* d_rows, float* d_scores, uint32_t* d_counts, int formula = CGVEC_FORMULA_SIMD) {
    NvtxRange nvtx_("cgvec.tensor_scan");
    {
        int arc = ensure_smem_attr(tc_scan_kernel, kSmemBudget);
        if (!arc) arc = ensure_smem_attr(tc2_scan_kernel, kSmemBudget);
        if (!arc) arc = ensure_smem_attr(tc_select_kernel, kTcCap * 8);
        if (arc) return arc;
    }
    const bool pairs = tc_use_pairs(ix, nq);
    const uint32_t n_max = pairs ? kTc2MaxN : tc_max_n(ix);
    if (n_max == 0 || nq > n_max) return fail(CGVEC_ERR_UNSUPPORTED, "dimension %u leaves no room for a resident query block", ix->dim);
    const uint32_t N = (nq + 15) & ~15u;
    const bool f32 = ix->dtype == CGVEC_F32;
    const uint64_t n = ix->n;
    const uint32_t want = (uint32_t)(k < n ? k : n);
    uint32_t kp = k + (ix->opt_tc_margin > 0 ? (uint32_t)ix->opt_tc_margin : (k / 2 > 32 ? k / 2 : 32));
    if (kp > kTcCap / 8) kp = kTcCap / 8;
    if (kp < k) return fail(CGVEC_ERR_UNSUPPORTED, "k = %u is too large for the tensor path", k);
    uint32_t sel_cap = next_pow2(kp + kTcSelStep);                    // selector sort buffer: best kp + at least one step of new keys
    if (sel_cap < 512) sel_cap = 512;
    TcPlan plan;
    if (!tc_plan_fill(ix, pairs, N, &plan, sel_cap)) return fail(CGVEC_ERR_UNSUPPORTED, "no shared memory for the tensor kernel at N = %u, dimension %u", N, ix->dim);
    const uint32_t kbe = f32 ? kTcKBlock / 2 : kTcKBlock;            // elements per 128-byte K-block
    const uint32_t nkb = plan.nkb, dpad = nkb * kbe, stages = plan.stages;
    const uint32_t tile_rows = pairs ? 2 * kTcTileRows : kTcTileRows;

    int rc;
    rc = ensure(&c->d_B, &c->B_cap, (size_t)N * dpad * (f32 ? 2 : 1)); if (rc) return rc;      // capacity counted in halves
    rc = ensure(&c->d_tc_f, &c->tcf_cap, (size_t)3 * kTc2MaxN); if (rc) return rc;
    rc = ensure(&c->d_tc_u, &c->tcu_cap, (size_t)3 * kTc2MaxN + 4); if (rc) return rc;
    rc = ensure(&c->d_cand, &c->cand_cap, (size_t)kTc2MaxN * kTcCap); if (rc) return rc;
    rc = ensure(&c->d_exact, &c->exact_cap, (size_t)kTc2MaxN * (kTcCap / 8)); if (rc) return rc;
    rc = ensure(&c->h_proven, &c->hprov_cap, (size_t)kTc2MaxN + 4, true); if (rc) return rc;
    float *d_thr = c->d_tc_f, *d_na = c->d_tc_f + kTc2MaxN, *d_rho = c->d_tc_f + 2 * kTc2MaxN;
    uint32_t *d_cnt = c->d_tc_u, *d_proven = c->d_tc_u + kTc2MaxN, *d_consumed = c->d_tc_u + 2 * kTc2MaxN, *d_overflow = c->d_tc_u + 3 * kTc2MaxN;

    if (f32) tc_prep_queries_kernel<float><<<(N * 8 + 127) / 128, 128, 0, st>>>(d_q, qstride, nq, N, ix->dim, dpad, reinterpret_cast<float*>(c->d_B), d_na, d_rho, d_thr, d_cnt, d_overflow);
    else tc_prep_queries_kernel<__half><<<(N * 8 + 127) / 128, 128, 0, st>>>(d_q, qstride, nq, N, ix->dim, dpad, c->d_B, d_na, d_rho, d_thr, d_cnt, d_overflow);
    ix->launches++;
    CUDA_TRY(cudaGetLastError());

    CUtensorMap tmA, tmB;
    rc = make_tmap_rows(&tmA, ix->d_rows, ix->dim, n, (uint64_t)ix->ld * ix->esize, kTcTileRows, ix->opt_tc_l2promo, f32, plan.kbs, plan.box4d != 0); if (rc) return rc;
    if (pairs) rc = make_tmap_rows(&tmB, c->d_B, dpad, N, (uint64_t)dpad * ix->esize, N / 2, 1, f32, plan.kbs, plan.box4d != 0);
    else rc = make_tmap_rows(&tmB, c->d_B, dpad, N, (uint64_t)dpad * ix->esize, N, 1, f32, 1, false);
    if (rc) return rc;

    TcParams p{};
    ScanParams map = map_params(ix);
    p.n_rows = n; p.norms = ix->d_norms; p.thr = d_thr; p.cand = c->d_cand; p.cand_count = d_cnt; p.overflow = d_overflow;
    p.cap = kTcCap; p.nq = nq; p.N = N; p.nkb = nkb; p.stages = stages; p.metric = METRIC_COSINE;
    p.kbs = plan.kbs; p.groups = plan.groups; p.box4d = plan.box4d;
    p.tmem_cols = next_pow2(2 * N) < 32 ? 32 : next_pow2(2 * N);
    p.debug = (uint32_t)ix->opt_tc_debug;
    p.tf32 = f32 ? 1u : 0u;
    p.row_offset = map.row_offset; p.blk_rows = map.blk_rows; p.n_shards = map.n_shards; p.shard_id = map.shard_id;
    const uint32_t smem = plan.smem;

    p.kp = kp; p.sel_cap = sel_cap; p.consumed = d_consumed;
    auto launch_range = [&](uint64_t begin, uint64_t end, uint32_t sel_on, uint32_t boot = 0) -> int {
        p.row_begin = begin; p.row_end = end; p.sel_on = sel_on; p.boot = boot;
        const uint64_t tiles = (end - begin + tile_rows - 1) / tile_rows;
        if (pairs) {
            const uint64_t max_pairs = (uint64_t)ix->sm_count / 2;
            const uint32_t grid = 2 * (uint32_t)(tiles < max_pairs ? tiles : max_pairs);
            tc2_scan_kernel<<<grid, kTcThreads, smem, st>>>(tmA, tmB, p);
        } else {
            const uint32_t grid = (uint32_t)(tiles < (uint64_t)ix->sm_count ? tiles : (uint64_t)ix->sm_count);
            tc_scan_kernel<<<grid, kTcThreads, smem, st>>>(tmA, tmB, p);
        }
        ix->launches++;
        CUDA_TRY(cudaGetLastError());
        return CGVEC_OK;
    };
    auto launch_select = [&](uint32_t mode, uint32_t fixed_cnt = 0) -> int {
        tc_select_kernel<<<nq, 1024, kTcCap * 8, st>>>(c->d_cand, d_cnt, d_consumed, d_thr, kTcCap, kp, kTcCap, mode, fixed_cnt);
        ix->launches++;
        CUDA_TRY(cudaGetLastError());
        return CGVEC_OK;
    };
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ix->opt_tc_flow == 0) {
        // Bootstrap range (every row survives: thresholds are -inf) -> select -> ONE launch over the rest of the shard whose
        // selector warps keep tightening the thresholds in place -> select.  Unwritten list slots must read as 0.
        CUDA_TRY(cudaMemsetAsync(c->d_cand, 0, (size_t)nq * kTcCap * sizeof(uint64_t), st));
        uint64_t S0 = ix->opt_tc_first > 0 ? (uint64_t)ix->opt_tc_first : 4096;
        if (S0 > kTcCap) S0 = kTcCap;
        S0 = (S0 + tile_rows - 1) / tile_rows * tile_rows;
        if (S0 >= n || (n - S0) * 4 < S0) S0 = n;                         // a small shard is one range
        if (S0 > kTcCap) {                                                // (remainder folded in: fall back to fixed ranges of <= cap rows)
            S0 = kTcCap / tile_rows * tile_rows;
        }
        rc = launch_range(0, S0, 0, 1); if (rc) return rc;
        rc = launch_select(S0 >= n ? 2u : 0u, (uint32_t)S0); if (rc) return rc;
        if (S0 < n) {
            if (ix->opt_timing) { CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1)); CUDA_TRY(cudaEventRecord(e0, st)); }
            rc = launch_range(S0, n, 1); if (rc) return rc;
            if (ix->opt_timing) {                                         // the dominant kernel: main range of the tensor scan
                CUDA_TRY(cudaEventRecord(e1, st));
                std::lock_guard<std::mutex> lk(ix->ev_mu);
                ix->timed.push_back({e0, e1, 1});
            }
            rc = launch_select(1); if (rc) return rc;
        }
    } else {
        // legacy flow (option tc_flow = 1, kept for A/B timing): geometrically growing row ranges with a select in between.
        // After T rows the threshold sits at quantile kp/T, so a range of S rows adds about S*kp/T survivors per query.
        uint32_t target = ix->opt_tc_target > 0 ? (uint32_t)ix->opt_tc_target : 1024;
        if (target < 4 * kp) target = 4 * kp;
        if (target > kTcCap / 2) target = kTcCap / 2;
        uint64_t T = 0;
        while (T < n) {
            uint64_t S = (T == 0) ? (ix->opt_tc_first > 0 ? (uint64_t)ix->opt_tc_first : target) : T * (target - kp) / kp;
            if (T == 0 && S > kTcCap) S = kTcCap;
            S = (S + tile_rows - 1) / tile_rows * tile_rows;
            if (S < tile_rows) S = tile_rows;
            if (T + S > n || (n - T - S) * 8 < S) S = n - T;      // fold a small remainder into this range
            rc = launch_range(T, T + S, 0); if (rc) return rc;
            rc = launch_select(2); if (rc) return rc;
            T += S;
        }
    }
    // exact re-score of the survivors, sort, proof
    {
        const uint32_t total = nq * kp;
        if (f32)
            tc_rescore_kernel<float><<<(total * 8 + 127) / 128, 128, 0, st>>>(static_cast<const float*>(ix->d_rows), ix->dim, ix->ld, d_q, qstride, d_na,
                                                                             ix->d_norms, c->d_cand, kTcCap, kp, nq, METRIC_COSINE, formula, ix->row_offset, ix->blk_rows, ix->n_shards, c->d_exact);
        else
            tc_rescore_kernel<__half><<<(total * 8 + 127) / 128, 128, 0, st>>>(static_cast<const __half*>(ix->d_rows), ix->dim, ix->ld, d_q, qstride, d_na,
                                                                              ix->d_norms, c->d_cand, kTcCap, kp, nq, METRIC_COSINE, formula, ix->row_offset, ix->blk_rows, ix->n_shards, c->d_exact);
        ix->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    rc = ensure_parts(c, (size_t)nq * kp * 2); if (rc) return rc;
    // sorted exact keys [nq][kp] (needed by the proof) ...
    uint64_t* sorted = c->d_part[1];
    rc = merge_lists(ix, c, c->d_exact, nq, 1, kp, 0, sorted, nullptr, nullptr, nullptr, st, kp, kp, kp, /*sorted_in=*/0); if (rc) return rc;
    // Error budget of the proof, all relative to |q||r| (cosine units):
    //   (d+8) 2^-23      tensor-core accumulation of d exact products (allowing truncation instead of rounding)
    //   (d/8+12) 2^-24   the oracle's own dot: d/8 fused steps per lane + horizontal sum + tail + division
    //   2 (d/8+16) 2^-24 the two squared norms in oracle order, their square roots, rsqrtf (2 ulp) and the final multiply
    //   rho_q            query rounding to the operand type, measured per query (Cauchy-Schwarz), added in the kernel
    //   2^-10 (+ cross)  TF32 only: the tensor core drops 13 mantissa bits of the ROWS as well
    const float acc_bound = (float)(ix->dim + 8) * 1.1920929e-7f + (float)(ix->dim / 8 + 12) * 5.9604645e-8f +
                            2.0f * (float)(ix->dim / 8 + 16) * 5.9604645e-8f + (f32 ? 1.0e-3f : 0.0f) +
                            (formula != CGVEC_FORMULA_SIMD ? 4.0f * (float)(ix->dim + 8) * 5.9604645e-8f : 0.0f);   // |formula(x) - simd(x)| (search_formula's eps)
    tc_verify_kernel<<<(nq + 127) / 128, 128, 0, st>>>(c->d_cand, kTcCap, d_cnt, kp, sorted, want ? want : 1, d_na, d_rho, acc_bound, METRIC_COSINE, nq,
                                                      d_proven, d_overflow);
    ix->launches++;
    CUDA_TRY(cudaGetLastError());
    // ... and this shard's best-k (keys for the exchange, or decoded results when unsharded)
    rc = merge_lists(ix, c, sorted, nq, 1, k, 0, local_keys, d_rows, d_scores, d_counts, st, kp, kp, kp); if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->h_proven, d_proven, nq * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    ix->tc_batches++;
    for (uint32_t q = 0; q < nq; ++q) {
        if (c->h_proven[q]) continue;
        ix->tc_fallbacks++;                      // could not prove: this query takes the exact-order kernel
        if (formula != CGVEC_FORMULA_SIMD)       // (host-I/O, unsharded only: outputs are the pinned host mirrors)
            rc = search_formula(ix, c, d_q + (size_t)q * qstride, k, formula, st, d_rows + (size_t)q * k, d_scores + (size_t)q * k, d_counts + q);
        else
            rc = local_exact(ix, c, d_q + (size_t)q * qstride, 1, k, CGVEC_COSINE, st, local_keys ? local_keys + (size_t)q * k : nullptr,
                             d_rows ? d_rows + (size_t)q * k : nullptr, d_scores ? d_scores + (size_t)q * k : nullptr, d_counts ? d_counts + q : nullptr);
        if (rc) return rc;
    }
    return CGVEC_OK;
}

int tensor_batch(Index* ix, SearchCtx* c, const float* d_q, uint32_t qstride, uint32_t nq, uint32_t k, cudaStream_t st,
                 uint64_t* d_rows, float* d_scores, uint32_t* d_counts, int formula = CGVEC_FORMULA_SIMD) {
    if (ix->world == 1) return local_tensor(ix, c, d_q, qstride, nq, k, st, nullptr, d_rows, d_scores, d_counts, formula);
    uint64_t* local_keys = nullptr;
    int rc = ensure_gather(ix, c, nq, k, &local_keys);
    if (rc) return rc;
    rc = local_tensor(ix, c, d_q, qstride, nq, k, st, local_keys, nullptr, nullptr, nullptr);
    if (rc) return rc;
    return exchange_and_decode(ix, c, local_keys, nq, k, 0, st, d_rows, d_scores, d_counts);
}

// AUTO's rule for sending a batch to the tensor kernels (rank-invariant on a sharded index: `rows` is the agreed smallest
// shard).  Option tc_min_batch > 0 pins a fixed batch threshold (doubled for f32 storage); the default is a cost model fitted
// to tools/bench_paths.py on a B200 (profiles/r02_exact_vs_tensor_small_batches.txt), in milliseconds:
//   exact-order pass, 1 query   u = 0.03 + rows*dim*4 / 7.0e9   (f16 rows cost as much as f32: the kernel is bound by shared-memory
//                                   wavefronts per element, not by bytes); 2 queries 1.12 u, 4 queries 1.63 u per launch
//   tensor pass (<= n_max queries) 0.15 + rows*dim*esize / 6.6e9 + MMA time (matters beyond ~64 queries only)
// Batch-1 always stays on the exact-order kernel (fully asynchronous, fused peer exchange, resident server).
// the two estimates in milliseconds (pure function of the shape: exported as cgvec_path_cost_model for the CPU test tier)
void tensor_cost_model(bool f32, uint32_t dim, uint64_t rows, uint32_t nq, uint32_t n_max, double* t_exact, double* t_tensor) {
    const double elems = (double)rows * dim;
    const double u = 0.03 + elems * 4.0 / 7.0e9;
    const uint32_t n4 = nq / 4, r = nq % 4;
    *t_exact = n4 * 1.63 * u + (r == 3 ? 2.12 * u : r == 2 ? 1.12 * u : r == 1 ? u : 0.0);
    const uint32_t passes = n_max ? (nq + n_max - 1) / n_max : 1;
    const double mma_rate = f32 ? 6.0e11 : 1.2e12;                               // sustained flop per ms under the power cap
    *t_tensor = passes * (0.15 + elems * (f32 ? 4 : 2) / 6.6e9) + 2.0 * elems * nq / mma_rate;
}
bool tensor_auto_rule(const Index* ix, uint64_t rows, uint32_t nq, uint32_t k, uint32_t n_max) {
    if (rows < 4 * kTcCap || k > kTcCap / 16 || nq < 2 || n_max == 0) return false;
    if (ix->opt_tc_min_nq > 0) return nq >= (ix->dtype == CGVEC_F32 ? (uint32_t)ix->opt_tc_min_nq * 2 : (uint32_t)ix->opt_tc_min_nq);
    double t_exact = 0.0, t_tensor = 0.0;
    tensor_cost_model(ix->dtype == CGVEC_F32, ix->dim, rows, nq, n_max, &t_exact, &t_tensor);
    return t_tensor < t_exact;
}
bool tensor_auto_ok(const Index* ix, int metric, uint32_t nq, uint32_t k) {
    return tensor_path_applicable(ix, metric, nq) && tensor_auto_rule(ix, ix->world > 1 ? ix->agreed_min_n : ix->n, nq, k, tc_batch_limit(ix, nq));
}

// Runs nq queries (device, stride qstride) through whichever kernel family `path` selects, in batches.
int run_queries(Index* ix, SearchCtx* c, const float* d_q, uint32_t qstride, uint32_t nq, uint32_t k, int metric, int path, cudaStream_t st,
                uint64_t* d_rows, float* d_scores, uint32_t* d_counts, int formula = CGVEC_FORMULA_SIMD) {
    bool tensor = false;
    if (path == CGVEC_PATH_TENSOR) {
        if (!tensor_path_applicable(ix, metric, nq)) return fail(CGVEC_ERR_UNSUPPORTED, "the tensor-core path serves the cosine metric");
        tensor = true;
    } else if (path == CGVEC_PATH_AUTO) {
        tensor = tensor_auto_ok(ix, metric, nq, k);
    }
    uint32_t n_max = tensor ? tc_batch_limit(ix, nq) : 0;
    if (tensor && n_max == 0) {
        if (path == CGVEC_PATH_TENSOR) return fail(CGVEC_ERR_UNSUPPORTED, "dimension %u leaves no room for a resident query block", ix->dim);
        tensor = false;
    }
    uint32_t q0 = 0;
    while (q0 < nq) {
        uint32_t b;
        int rc;
        if (tensor) {
            b = nq - q0 < n_max ? nq - q0 : n_max;
            rc = tensor_batch(ix, c, d_q + (size_t)q0 * qstride, qstride, b, k, st, d_rows ? d_rows + (size_t)q0 * k : nullptr,
                              d_scores ? d_scores + (size_t)q0 * k : nullptr, d_counts ? d_counts + q0 : nullptr, formula);
        } else {
            b = nq - q0 >= 4 && ix->opt_max_nq >= 4 ? 4 : (nq - q0 >= 2 && ix->opt_max_nq >= 2 ? 2 : 1);
            rc = scan_batch(ix, c, d_q + (size_t)q0 * qstride, b, k, metric, st, d_rows ? d_rows + (size_t)q0 * k : nullptr,
                            d_scores ? d_scores + (size_t)q0 * k : nullptr, d_counts ? d_counts + q0 : nullptr);
        }
        if (rc) return rc;
        q0 += b;
    }
    return CGVEC_OK;
}

void drain_timings(Index* ix) {
    std::lock_guard<std::mutex> lk(ix->ev_mu);
    for (auto& pr : ix->timed) {
        if (cudaEventSynchronize(pr.e1) == cudaSuccess) {
            float ms = 0.0f;
            if (cudaEventElapsedTime(&ms, pr.e0, pr.e1) == cudaSuccess) {
                if (pr.kind == 0) { ix->scan_ms_total += ms; ix->scan_timed++; ix->last_scan_ms = ms; }
                else { ix->tc_main_ms_total += ms; ix->tc_main_timed++; }
            }
        }
        cudaEventDestroy(pr.e0); cudaEventDestroy(pr.e1);
    }
    ix->timed.clear();
}

// Maps every rank's exchange buffer into this process (CUDA IPC over NVLink peer access).  Collective: all ranks call it.
void setup_peer_exchange(Index* ix) {
    ix->p2p = false;
    if (ix->world > (int)kXchgMaxWorld) return;
    bool ok = cudaMalloc(reinterpret_cast<void**>(&ix->xbuf), kXchgBytes) == cudaSuccess &&
              cudaMemset(ix->xbuf, 0, kXchgBytes) == cudaSuccess;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (ok) ok = cudaIpcGetMemHandle(&mine, ix->xbuf) == cudaSuccess;
    // all-gather (handle, ok) so that every rank takes the same decision
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    std::vector<uint8_t> host(rec * ix->world, 0);
    uint8_t* d_all = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&d_all), rec * ix->world) != cudaSuccess) { cudaGetLastError(); return; }
    std::vector<uint8_t> me(rec, 0);
    memcpy(me.data(), &mine, sizeof(mine));
    me[sizeof(mine)] = ok ? 1 : 0;
    cudaMemcpy(d_all + rec * ix->rank, me.data(), rec, cudaMemcpyHostToDevice);
    bool gathered = nccl_api().AllGather(d_all + rec * ix->rank, d_all, rec, /*ncclUint8*/ 1, ix->comm, ix->main_stream) == kNcclSuccess &&
                    cudaStreamSynchronize(ix->main_stream) == cudaSuccess &&
                    cudaMemcpy(host.data(), d_all, rec * ix->world, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(d_all);
    if (!gathered) { cudaGetLastError(); return; }
    bool all_ok = true;
    for (int r = 0; r < ix->world; ++r) all_ok = all_ok && host[rec * r + sizeof(mine)] == 1;
    if (all_ok) {
        for (int r = 0; r < ix->world && all_ok; ++r) {
            if (r == ix->rank) { ix->xpeer[r] = ix->xbuf; continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, &host[rec * r], sizeof(h));
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); all_ok = false; break; }
            ix->xpeer[r] = static_cast<uint8_t*>(ptr);
        }
    }
    // second agreement round: only use peer memory if EVERY rank mapped every peer
    uint8_t* d_flag = nullptr;
    if (cudaMalloc(reinterpret_cast<void**>(&d_flag), ix->world) != cudaSuccess) { cudaGetLastError(); return; }
    uint8_t f = all_ok ? 1 : 0;
    cudaMemcpy(d_flag + ix->rank, &f, 1, cudaMemcpyHostToDevice);
    std::vector<uint8_t> flags(ix->world, 0);
    bool g2 = nccl_api().AllGather(d_flag + ix->rank, d_flag, 1, 1, ix->comm, ix->main_stream) == kNcclSuccess &&
              cudaStreamSynchronize(ix->main_stream) == cudaSuccess &&
              cudaMemcpy(flags.data(), d_flag, ix->world, cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(d_flag);
    bool every = g2;
    for (int r = 0; r < ix->world; ++r) every = every && flags[r] == 1;
    ix->p2p = every;
    if (!every) cudaGetLastError();
}

