// scan_exact.cuh — K1: the HBM-bound brute-force scan + fused top-k, in the reference's EXACT
// floating-point operation order.
//
// Replaces ParallelVectorOps::parallel_top_k_search (reference crates/codegraph-vector/src/simd_ops.rs:361-383):
// every stored row is scored against the query with adaptive_cosine_similarity (simd_ops.rs:281-295 ->
// cosine_similarity_avx2 :15-78 for d >= 32, cosine_similarity_scalar :257-278 below), or with
// dot_product_avx2 (:149-183) / l2_distance_avx2 (:105-143), and the best k are kept.
//
// How the order is reproduced.  _mm256_fmadd_ps is 8 independent fmaf chains (lane l consumes
// elements l, l+8, l+16, ... in order).  Here 8 THREADS play the 8 AVX lanes of one row: thread L of an
// octet walks elements 8i+L with __fmaf_rn, the octet is reduced with shuffles in exactly
// horizontal_sum_avx2's association ((l0+l4)+(l1+l5))+((l2+l6)+(l3+l7)), octet lane 0 adds the scalar
// tail (un-fused mul+add) and finishes with IEEE sqrt/div.  Scores are therefore bit-identical to the
// CPU reference, so the fused top-k needs no over-fetch / re-rank to return bit-exact indices.
// The row's squared norm (a function of the row only) is computed once at store time in the same
// order (norms.cuh) — the scan issues one FFMA per matrix element.
//
// Data movement.  Persistent CTAs (one per SM).  The last warp is the producer: for every tile of `tile_rows`
// rows it arms an mbarrier and issues one cp.async.bulk (TMA engine, SASS UBLKCP) per row into a
// ring of `stages` shared-memory buffers; rows land with a padded stride (row_words = 8 mod 16 words)
// so the 4 rows a consumer warp reads concurrently fall in disjoint bank octets -> conflict-free
// LDS.32.  8 consumer warps (4 rows each, 1 octet per row) release each stage through an `empty`
// mbarrier.  Per SM up to stages x tile_rows x row bytes (~200 KB) are in flight, far more than
// the ~45 KB Little's law needs for 6.5 TB/s / 148 SMs.
//
// Top-k.  Octet lane 0 turns (score, global row) into a 64-bit key (common.cuh) and appends it to a
// CTA-shared candidate buffer when it beats the CTA's current k-th key.  Every `sync_interval` tiles
// the consumer warps meet on a named barrier and, if the buffer could overflow in the next interval,
// bitonic-sort it and keep the best k (the new threshold).  Each CTA finally writes its sorted best k
// keys; topk.cuh merges the per-CTA lists.
#pragma once
#include "common.cuh"

namespace cgv {

constexpr int kScanConsumerWarps = 8;
constexpr int kScanThreads = 32 * (1 + kScanConsumerWarps);
constexpr int kScanMaxQ = 4;

struct ScanParams {
    const void* rows;          // local shard, row-major, ld elements between rows (16-byte aligned rows)
    const float* norms;        // per-row squared norm in reference order (cosine only); padded to 32 rows
    const float* queries;      // [nq][d] f32 on the device
    uint64_t* partials;        // [nq][grid][k] sorted keys, 0-padded
    uint64_t n_rows;           // local rows
    uint32_t d, ld;            // logical / padded row length in elements
    uint32_t row_words;        // shared-memory row stride in 32-bit words (>= ld*esize/4, = 8 mod 16)
    uint32_t tile_rows;        // rows per stage: 4, 8, 16 or 32
    uint32_t stages;           // multiple of active_groups (a stage is always drained by the same warp group)
    uint32_t active_groups;    // warp groups that consume tiles, <= 8 / (tile_rows/4); the rest only help to sort
    uint32_t k;
    uint32_t cand_cap;         // power of two >= k + sync_interval*tile_rows
    uint32_t sync_interval;    // tiles between threshold syncs (multiple of the consumer group count)
    uint32_t use_l2_hint;
    // global row = ((local / blk_rows) * n_shards + shard_id) * blk_rows + local % blk_rows + row_offset
    uint64_t row_offset;
    uint32_t blk_rows, n_shards, shard_id;
    // METRIC_I8 (int8 quantised scan, scan_i8.cuh): rows are u8 codes viewed as 32-bit words, `norms` holds int32 sum((code-128)^2),
    // `queries` the quantised query as packed s8 words; d / ld count WORDS.
    const float* i8_q_norm;    // [1] sqrt(sum(q^2)) as the reference accumulates it
    const int32_t* i8_q_sum;   // [1] sum(q)
    uint32_t i8_tie_mode;      // 1: keep the k LOWEST rows with score >= i8_tie_vstar (second pass of the reference's tie resolution)
    float i8_tie_vstar;
    uint64_t* trace;           // optional [start, end] slot of this launch (see common.cuh trace_begin)
    uint32_t early_trigger;    // PDL chain mode 2: release dependents at once, order our partial-list WRITE after the predecessor
};

__host__ __device__ inline uint32_t scan_row_words(uint32_t ld, uint32_t esize) {
    uint32_t w = (ld * esize + 3) / 4;
    w = (w + 3) & ~3u;                    // 16-byte rows
    while ((w & 15u) != 8u) w += 4;       // stride = 8 (mod 16) words: 4 concurrent rows -> disjoint bank octets
    return w;
}

struct ScanSmemLayout {
    uint32_t off_rows, off_norms, off_q, off_cand, off_bars, off_misc, total;
};
__host__ __device__ inline ScanSmemLayout scan_smem_layout(uint32_t row_words, uint32_t tile_rows, uint32_t stages,
                                                           uint32_t d, uint32_t nq, uint32_t cand_cap) {
    ScanSmemLayout L;
    uint32_t o = 0;
    L.off_rows = o;   o += stages * tile_rows * row_words * 4;
    L.off_norms = o;  o += stages * 32 * 4;                     // one 128-byte norm slot per stage
    L.off_q = o;      o += nq * ((d * 4 + 15) & ~15u);
    o = (o + 15) & ~15u;
    L.off_cand = o;   o += nq * cand_cap * 8;
    L.off_bars = o;   o += (2 * stages + 1) * 8;
    L.off_misc = o;   o += nq * 16 + 16;                        // per query: thr (u64), count (u32), pad
    L.total = (o + 15) & ~15u;
    return L;
}

__device__ __forceinline__ uint64_t scan_global_row(const ScanParams& p, uint64_t local) {
    uint64_t b = local / p.blk_rows, r = local - b * p.blk_rows;
    return (b * p.n_shards + p.shard_id) * p.blk_rows + r + p.row_offset;
}

// One octet (8 threads = 8 AVX lanes) scores one row for NQ queries.  Valid on octet lane 0.
template <typename T, int METRIC, int NQ>
__device__ __forceinline__ void score_row_octet(const T* __restrict__ row, const float* __restrict__ q, uint32_t qstride,
                                                uint32_t d, int L, const float* na, float nb, float* out) {
    const bool simd = (METRIC != METRIC_COSINE) || d >= 32;     // adaptive_cosine_similarity, simd_ops.rs:284
    float dp[NQ];
    if (simd) {
        float acc[NQ];
#pragma unroll
        for (int j = 0; j < NQ; ++j) acc[j] = 0.0f;
        const uint32_t chunks = d >> 3;                          // simd_ops.rs:30
        const T* rp = row + L;
        const float* qp = q + L;
#pragma unroll 8
        for (uint32_t i = 0; i < chunks; ++i) {                  // simd_ops.rs:31-47 (one lane of the fmadd)
            float vb = ldf(rp + 8 * i);
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                float va = qp[j * qstride + 8 * i];
                if (METRIC == METRIC_L2) {
                    float diff = sub_rn(va, vb);                 // simd_ops.rs:126-129
                    acc[j] = fma_rn(diff, diff, acc[j]);
                } else {
                    acc[j] = fma_rn(va, vb, acc[j]);             // simd_ops.rs:40 / :170
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NQ; ++j) acc[j] = hsum8_ref_order(acc[j]);   // simd_ops.rs:50 / :227-242
        if (L == 0) {
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                float r = 0.0f;                                  // scalar tail, un-fused (simd_ops.rs:55-65)
                for (uint32_t i = chunks * 8; i < d; ++i) {
                    float va = q[j * qstride + i], vb = ldf(row + i);
                    if (METRIC == METRIC_L2) {
                        float diff = sub_rn(va, vb);
                        r = add_rn(r, mul_rn(diff, diff));
                    } else {
                        r = add_rn(r, mul_rn(va, vb));
                    }
                }
                dp[j] = add_rn(acc[j], r);                       // simd_ops.rs:67
            }
        }
    } else if (L == 0) {                                         // cosine_similarity_scalar, simd_ops.rs:262-270
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            float s = 0.0f;
            for (uint32_t i = 0; i < d; ++i) s = add_rn(s, mul_rn(q[j * qstride + i], ldf(row + i)));
            dp[j] = s;
        }
    }
    if (L == 0) {
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            if (METRIC == METRIC_COSINE) {
                float np = sqrt_rn(mul_rn(na[j], nb));           // simd_ops.rs:72 / :272
                out[j] = (np == 0.0f) ? 0.0f : div_rn(dp[j], np);   // :73-77
            } else if (METRIC == METRIC_DOT) {
                out[j] = dp[j];
            } else {
                out[j] = sqrt_rn(dp[j]);                         // simd_ops.rs:142
            }
        }
    }
}

// int8 cosine numerator of one row (search_optimized, optimization.rs:124-130): sum((code - 128) * q) over the row, computed as
// dp4a.u32.s32(codes, q) - 128 * sum(q) by the caller.  Integer arithmetic is associative, so no order has to be reproduced:
// thread L takes words L, L+8, ... and the octet is reduced with shuffles.  Valid on octet lane 0.
__device__ __forceinline__ int i8_dot_octet(const uint32_t* __restrict__ row, const uint32_t* __restrict__ q, uint32_t words, int L) {
    int acc0 = 0, acc1 = 0;
    uint32_t i = L;
    // 8 row words + 8 query words in flight per thread before the first dp4a: at 1 byte per element the scan needs four times
    // the rows per second of the f32 scan from the same 8 consumer warps, so the loads must not wait for each other
    for (; i + 56 < words; i += 64) {
        uint32_t a[8], b[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { a[u] = row[i + 8 * u]; b[u] = q[i + 8 * u]; }
#pragma unroll
        for (int u = 0; u < 8; u += 2) {
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc0) : "r"(a[u]), "r"(b[u]));
            asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc1) : "r"(a[u + 1]), "r"(b[u + 1]));
        }
    }
    for (; i < words; i += 8) asm("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(acc0) : "r"(row[i]), "r"(q[i]));
    int acc = acc0 + acc1;
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    return acc;
}

// Squared norm of a vector held in shared/global memory, in the order the cosine kernels above imply
// (AVX2 lanes + hsum + tail for d >= 32, sequential un-fused below).  Valid on octet lane 0.
template <typename T>
__device__ __forceinline__ float sqnorm_octet(const T* __restrict__ v, uint32_t d, int L) {
    float r = 0.0f;
    if (d >= 32) {
        float acc = 0.0f;
        const uint32_t chunks = d >> 3;
#pragma unroll 8
        for (uint32_t i = 0; i < chunks; ++i) {
            float x = ldf(v + 8 * i + L);
            acc = fma_rn(x, x, acc);                             // simd_ops.rs:43 / :46
        }
        acc = hsum8_ref_order(acc);
        if (L == 0) {
            float t = 0.0f;
            for (uint32_t i = chunks * 8; i < d; ++i) { float x = ldf(v + i); t = add_rn(t, mul_rn(x, x)); }
            r = add_rn(acc, t);
        }
    } else if (L == 0) {
        for (uint32_t i = 0; i < d; ++i) { float x = ldf(v + i); r = add_rn(r, mul_rn(x, x)); }   // simd_ops.rs:268-269
    }
    return r;
}

template <typename T, int METRIC, int NQ>
__global__ void __launch_bounds__(kScanThreads, 1) scan_exact_kernel(const ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem_ex[];
    uint8_t* smem = smem_ex;
    const ScanSmemLayout lay = scan_smem_layout(p.row_words, p.tile_rows, p.stages, p.d, NQ, p.cand_cap);
    uint8_t* s_rows = smem + lay.off_rows;
    float* s_norms = reinterpret_cast<float*>(smem + lay.off_norms);
    float* s_q = reinterpret_cast<float*>(smem + lay.off_q);
    uint64_t* s_cand = reinterpret_cast<uint64_t*>(smem + lay.off_cand);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + lay.off_bars);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* q_bar = empty_bar + p.stages;
    uint64_t* s_thr = reinterpret_cast<uint64_t*>(smem + lay.off_misc);           // [NQ]
    uint32_t* s_count = reinterpret_cast<uint32_t*>(smem + lay.off_misc + NQ * 8);   // [NQ]

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t warps_per_stage = p.tile_rows >> 2;
    const uint32_t ngroups = p.active_groups;
    const uint32_t qstride = ((p.d * 4 + 15) & ~15u) >> 2;
    const uint32_t stage_bytes = p.tile_rows * p.row_words * 4;
    const uint32_t row_bytes = p.ld * sizeof(T);
    const bool ascending = (METRIC == METRIC_L2);

    const uint64_t num_tiles = (p.n_rows + p.tile_rows - 1) / p.tile_rows;
    const uint64_t my_tiles = (num_tiles > blockIdx.x) ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    // PDL chain (mode 2).  Every kernel of the per-query chain scan -> merge/exchange -> scan -> ... releases its
    // dependents immediately and executes griddepcontrol.wait before its first access that could conflict with its
    // predecessor: the merge before READING our partial lists, we before WRITING them (double buffered: the previous
    // reader of this buffer is the merge two kernels back).  By induction along the chain, passing our wait implies every
    // older kernel has completed, so the scanning itself may overlap the previous query's tail and merge.
    if (p.early_trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        trace_begin(p.trace);
        for (uint32_t s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], warps_per_stage);
        }
        mbar_init(q_bar, 1);
        for (int j = 0; j < NQ; ++j) { s_thr[j] = 0; s_count[j] = 0; }
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kScanConsumerWarps) {
        // ===================== producer: TMA-engine bulk copies, one per row =====================
        // (the last warp: the SM's arbiter favours higher warp ids, and this warp must never wait for an issue slot)
        const uint64_t policy = l2_policy_evict_first();
        if (lane == 0) {
            mbar_arrive_expect_tx(q_bar, NQ * qstride * 4);
            for (int j = 0; j < NQ; ++j) bulk_g2s(s_q + j * qstride, p.queries + (size_t)j * qstride, qstride * 4, q_bar);
        }
        for (uint64_t n = 0; n < my_tiles; ++n) {
            const uint32_t s = n % p.stages;
            if (n >= p.stages) mbar_wait(&empty_bar[s], ((n / p.stages) - 1) & 1);
            const uint64_t tile = blockIdx.x + n * gridDim.x;
            const uint64_t row0 = tile * p.tile_rows;
            const uint32_t rows = (uint32_t)min((uint64_t)p.tile_rows, p.n_rows - row0);
            const bool with_norms = (METRIC == METRIC_COSINE || METRIC == METRIC_I8);
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], rows * row_bytes + (with_norms ? p.tile_rows * 4 : 0));
            __syncwarp();
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.rows) + row0 * row_bytes;
            uint8_t* dst = s_rows + (size_t)s * stage_bytes;
            for (uint32_t r = lane; r < rows; r += 32) {
                if (p.use_l2_hint) bulk_g2s_hint(dst + (size_t)r * p.row_words * 4, src + (size_t)r * row_bytes, row_bytes, &full_bar[s], policy);
                else bulk_g2s(dst + (size_t)r * p.row_words * 4, src + (size_t)r * row_bytes, row_bytes, &full_bar[s]);
            }
            if (with_norms && lane == 0) bulk_g2s(s_norms + s * 32, p.norms + row0, p.tile_rows * 4, &full_bar[s]);
        }
    } else {
        // ===================== consumers: exact-order scoring + candidate filter =====================
        const uint32_t cw = warp;
        const uint32_t group = cw / warps_per_stage, sub = cw % warps_per_stage;
        const uint32_t ctid = tid, nct = kScanConsumerWarps * 32;
        const int L = lane & 7;
        const uint32_t lrow = sub * 4 + (lane >> 3);             // row of the tile this octet scores

        mbar_wait(q_bar, 0);
        float i8_qnorm = 0.0f;
        int i8_qsum = 0;
        if constexpr (METRIC == METRIC_I8) { i8_qnorm = *p.i8_q_norm; i8_qsum = *p.i8_q_sum; }
        float na[NQ];
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            na[j] = (METRIC == METRIC_COSINE) ? sqnorm_octet(s_q + j * qstride, p.d, L) : 0.0f;
            na[j] = __shfl_sync(0xffffffffu, na[j], lane & ~7);
        }

        const uint32_t epoch = p.sync_interval;
        const uint32_t flush_limit = p.cand_cap - epoch * p.tile_rows;
        uint64_t next_boundary = epoch;

        auto sync_and_maybe_compact = [&](bool force) {
            named_bar_sync(1, nct);
            uint32_t cnt[NQ];
#pragma unroll
            for (int j = 0; j < NQ; ++j) cnt[j] = s_count[j];
            named_bar_sync(1, nct);
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                if (force || cnt[j] > flush_limit) {
                    uint64_t* c = s_cand + (size_t)j * p.cand_cap;
                    for (uint32_t i = cnt[j] + ctid; i < p.cand_cap; i += nct) c[i] = 0;
                    named_bar_sync(1, nct);
                    bitonic_sort_desc(c, p.cand_cap, ctid, nct, 1);
                    if (ctid == 0) {
                        uint32_t keep = min(cnt[j], p.k);
                        s_count[j] = keep;
                        s_thr[j] = (keep >= p.k) ? c[p.k - 1] : 0ull;
                    }
                    named_bar_sync(1, nct);
                }
            }
        };

        // Idle groups (group >= ngroups) take no tiles but still join every threshold barrier below.
        for (uint64_t n = (group < ngroups ? group : my_tiles); n < my_tiles; n += ngroups) {
            while (next_boundary <= n) { sync_and_maybe_compact(false); next_boundary += epoch; }
            const uint32_t s = n % p.stages;
            mbar_wait(&full_bar[s], (n / p.stages) & 1);
            const uint64_t tile = blockIdx.x + n * gridDim.x;
            const uint64_t row = tile * p.tile_rows + lrow;      // local row index
            const T* rp = reinterpret_cast<const T*>(s_rows + (size_t)s * stage_bytes + (size_t)lrow * p.row_words * 4);
            const float nb = (METRIC == METRIC_COSINE || METRIC == METRIC_I8) ? s_norms[s * 32 + lrow] : 0.0f;
            float sc[NQ];
            bool skip = false;
            if constexpr (METRIC == METRIC_I8) {
                const int s1 = i8_dot_octet(reinterpret_cast<const uint32_t*>(rp), reinterpret_cast<const uint32_t*>(s_q), p.d, L);
                const int nv = __float_as_int(nb);
                skip = nv == 0;                                                     // optimization.rs:132-134
                const int dot = s1 - 128 * i8_qsum;
                sc[0] = skip ? 0.0f : div_rn((float)dot, mul_rn(i8_qnorm, sqrt_rn((float)nv)));   // :136
                if (p.i8_tie_mode) { skip = skip || !(sc[0] >= p.i8_tie_vstar); sc[0] = 1.0f; }
            } else {
                score_row_octet<T, METRIC, NQ>(rp, s_q, qstride, p.d, L, na, nb, sc);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);           // stage may be refilled
            if (L == 0 && row < p.n_rows && !skip) {
                const uint32_t grow = (uint32_t)scan_global_row(p, row);
#pragma unroll
                for (int j = 0; j < NQ; ++j) {
                    const uint64_t key = make_key(sc[j], grow, ascending);
                    if (key > *reinterpret_cast<volatile uint64_t*>(&s_thr[j])) {
                        uint32_t pos = atomicAdd(&s_count[j], 1u);
                        s_cand[(size_t)j * p.cand_cap + pos] = key;
                    }
                }
            }
        }
        while (next_boundary < my_tiles) { sync_and_maybe_compact(false); next_boundary += epoch; }
        sync_and_maybe_compact(true);
        // sorted best-k keys of this CTA (0-padded).  The wait orders this WRITE after the kernel ahead of us in the stream
        // whenever we were launched with the programmatic attribute — also when we did not release our own dependents
        // early (a grid smaller than the SM count): otherwise a small shard's scan could overwrite the double-buffered
        // lists / let the next exchange run while the previous one still spins on a slow peer.  It is a no-op for a
        // plain launch.
        asm volatile("griddepcontrol.wait;" ::: "memory");
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            const uint32_t cnt = s_count[j];
            uint64_t* out = p.partials + ((size_t)j * gridDim.x + blockIdx.x) * p.k;
            for (uint32_t i = ctid; i < p.k; i += nct) out[i] = (i < cnt) ? s_cand[(size_t)j * p.cand_cap + i] : 0ull;
        }
        if (ctid == 0) trace_end(p.trace);
    }
}

}  // namespace cgv
