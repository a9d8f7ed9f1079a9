// aux_kernels.cuh — the small kernels either side of the scan:
//   K3 row_sqnorm_kernel     per-row squared norms in reference order, once at store time
//   K4 rescore_kernel        exact re-scoring of (query, row) pairs in any of the reference's formulas
//   K5 merge_topk_kernel     merges sorted key lists (per-CTA partials, per-shard partials) + decode
//   normalize_rows_kernel    parallel_normalize_vectors (simd_ops.rs:386-419 -> normalize_avx2 :189-222)
//   synth_rows_kernel        deterministic synthetic rows (bench / property tests)
#pragma once
#include "common.cuh"
#include "scan_exact.cuh"

namespace cgv {

// ---- K3 -------------------------------------------------------------------------------------------
// norms[row] = ||row||^2 exactly as cosine_similarity_avx2 (simd_ops.rs:46,52,64,69) / _scalar (:269)
// would accumulate it.  One octet per row.
template <typename T>
__global__ void row_sqnorm_kernel(const T* __restrict__ rows, uint64_t first, uint64_t count, uint32_t d, uint32_t ld,
                                  float* __restrict__ norms) {
    const uint64_t octet = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 3;
    const int L = threadIdx.x & 7;
    const uint64_t r = first + (octet < count ? octet : count - 1);
    float v = sqnorm_octet(rows + r * ld, d, L);
    if (L == 0 && octet < count) norms[r] = v;
}

// sum of squares in dot_product_avx2 order for ANY d (normalize_avx2 calls dot_product_avx2(v, v), simd_ops.rs:195)
template <typename T>
__device__ __forceinline__ float sqnorm_simd_octet(const T* __restrict__ v, uint32_t d, int L) {
    float acc = 0.0f;
    const uint32_t chunks = d >> 3;
    for (uint32_t i = 0; i < chunks; ++i) { float x = ldf(v + 8 * i + L); acc = fma_rn(x, x, acc); }
    acc = hsum8_ref_order(acc);
    float t = 0.0f;
    if (L == 0) for (uint32_t i = chunks * 8; i < d; ++i) { float x = ldf(v + i); t = add_rn(t, mul_rn(x, x)); }
    return add_rn(acc, t);
}

__device__ __forceinline__ void st_elem(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_elem(__half* p, float v) { *p = __float2half_rn(v); }

// normalize_avx2 (simd_ops.rs:189-222): nsq = dot(v,v); if 0 leave; inv = 1/sqrt(nsq); v *= inv.
template <typename T>
__global__ void normalize_rows_kernel(T* __restrict__ rows, uint64_t count, uint32_t d, uint32_t ld) {
    const uint64_t octet = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 3;
    const int L = threadIdx.x & 7;
    const uint64_t r = octet < count ? octet : count - 1;
    T* v = rows + r * ld;
    float nsq = sqnorm_simd_octet(v, d, L);
    nsq = __shfl_sync(0xffffffffu, nsq, threadIdx.x & 31 & ~7);
    if (octet >= count || nsq == 0.0f) return;
    const float inv = div_rn(1.0f, sqrt_rn(nsq));
    for (uint32_t i = L; i < d; i += 8) st_elem(v + i, mul_rn(ldf(v + i), inv));
}

// ---- synthetic rows ---------------------------------------------------------------------------------
// value(seed, global_row, col) = (a+b+c+d - 131070) * 2^-16 with a..d the four 16-bit fields of
// splitmix64(seed ^ row*0x9E3779B97F4A7C15 ^ col*0xC2B2AE3D27D4EB4F): an Irwin-Hall(4) bell, exactly
// representable, reproducible bit-for-bit on the host (tests/synth.py).  unit_norm applies
// normalize_avx2's arithmetic in f32 before the (optional) rounding to f16.
__host__ __device__ inline float synth_value(uint64_t seed, uint64_t row, uint32_t col) {
    uint64_t z = seed ^ (row * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)col * 0xC2B2AE3D27D4EB4Full);
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    int32_t s = (int32_t)(z & 0xffff) + (int32_t)((z >> 16) & 0xffff) + (int32_t)((z >> 32) & 0xffff) +
                (int32_t)((z >> 48) & 0xffff) - 131070;
    return (float)s * (1.0f / 65536.0f);
}

template <typename T>
__global__ void synth_rows_kernel(T* __restrict__ rows, uint64_t first_local, uint64_t count, uint32_t d, uint32_t ld,
                                  uint64_t seed, int unit_norm, ScanParams map) {
    const uint64_t octet = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 3;
    const int L = threadIdx.x & 7;
    const uint64_t lr = first_local + (octet < count ? octet : count - 1);
    const uint64_t grow = scan_global_row(map, lr);
    float inv = 1.0f;
    bool scale = false;
    if (unit_norm) {
        float acc = 0.0f;
        const uint32_t chunks = d >> 3;
        for (uint32_t i = 0; i < chunks; ++i) { float x = synth_value(seed, grow, 8 * i + L); acc = fma_rn(x, x, acc); }
        acc = hsum8_ref_order(acc);
        float t = 0.0f;
        if (L == 0) for (uint32_t i = chunks * 8; i < d; ++i) { float x = synth_value(seed, grow, i); t = add_rn(t, mul_rn(x, x)); }
        float nsq = add_rn(acc, t);
        nsq = __shfl_sync(0xffffffffu, nsq, threadIdx.x & 31 & ~7);
        if (nsq != 0.0f) { inv = div_rn(1.0f, sqrt_rn(nsq)); scale = true; }
    }
    if (octet >= count) return;
    T* v = rows + lr * ld;
    for (uint32_t i = L; i < ld; i += 8) {
        float x = (i < d) ? synth_value(seed, grow, i) : 0.0f;
        if (scale) x = mul_rn(x, inv);
        st_elem(v + i, x);
    }
}

// ---- K4 -------------------------------------------------------------------------------------------
// out[i] = formula(query, row[local_rows[i]]).  One octet per pair; sequential formulas run on octet lane 0
// (their order is inherently serial: un-fused left-to-right sums, search.rs:524-526 / simd_ops.rs:266-270).
template <typename T>
__global__ void rescore_kernel(const T* __restrict__ rows, uint32_t d, uint32_t ld, const float* __restrict__ q,
                               const uint64_t* __restrict__ local_rows, uint32_t n, int metric, int formula,
                               float* __restrict__ out) {
    const uint32_t octet = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int L = threadIdx.x & 7;
    const uint32_t pi = octet < n ? octet : n - 1;
    const T* row = rows + local_rows[pi] * ld;
    float res = 0.0f;
    if (formula == FORM_SIMD) {
        float na = 0.0f, nb = 0.0f;
        if (metric == METRIC_COSINE) {
            na = sqnorm_octet(q, d, L);
            nb = sqnorm_octet(row, d, L);
        }
        if (metric == METRIC_COSINE) score_row_octet<T, METRIC_COSINE, 1>(row, q, 0, d, L, &na, nb, &res);
        else if (metric == METRIC_DOT) score_row_octet<T, METRIC_DOT, 1>(row, q, 0, d, L, &na, nb, &res);
        else score_row_octet<T, METRIC_L2, 1>(row, q, 0, d, L, &na, nb, &res);
    } else if (L == 0) {
        float dp = 0.0f, na = 0.0f, nb = 0.0f;
        if (formula == FORM_SCALAR) {                            // simd_ops.rs:262-277 (one interleaved loop)
            for (uint32_t i = 0; i < d; ++i) {
                float va = q[i], vb = ldf(row + i);
                dp = add_rn(dp, mul_rn(va, vb));
                na = add_rn(na, mul_rn(va, va));
                nb = add_rn(nb, mul_rn(vb, vb));
            }
            if (metric == METRIC_DOT) res = dp;
            else { float np = sqrt_rn(mul_rn(na, nb)); res = (np == 0.0f) ? 0.0f : div_rn(dp, np); }
        } else {                                                 // search.rs:524-532 / optimization.rs:409-417
            for (uint32_t i = 0; i < d; ++i) dp = add_rn(dp, mul_rn(q[i], ldf(row + i)));
            for (uint32_t i = 0; i < d; ++i) na = add_rn(na, mul_rn(q[i], q[i]));
            for (uint32_t i = 0; i < d; ++i) { float vb = ldf(row + i); nb = add_rn(nb, mul_rn(vb, vb)); }
            na = sqrt_rn(na);
            nb = sqrt_rn(nb);
            if (metric == METRIC_DOT) res = dp;
            else if (formula == FORM_BASELINE) res = (na == 0.0f || nb == 0.0f) ? __int_as_float(0x7f800000) : sub_rn(1.0f, div_rn(dp, mul_rn(na, nb)));
            else res = (na == 0.0f || nb == 0.0f) ? 0.0f : div_rn(dp, mul_rn(na, nb));
        }
    }
    if (L == 0 && octet < n) out[octet] = res;
}

// Turns rescored (score, global row) pairs into sort keys.
__global__ void make_keys_kernel(const float* __restrict__ scores, const uint64_t* __restrict__ grows, uint32_t n,
                                 int ascending, uint64_t* __restrict__ keys) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = make_key(scores[i], (uint32_t)grows[i], ascending != 0);
}

// ---- K5 -------------------------------------------------------------------------------------------
// in  : keys of list l of query q at in + q * q_stride + l * l_stride (list_len keys each, sorted or not — the CTA sorts), 0 = empty;
//       the contiguous [nq][n_lists][list_len] layout is q_stride = n_lists * list_len, l_stride = list_len, the gathered
//       [rank][nq][k] layout of the exchange is q_stride = k, l_stride = nq * k (no re-packing copies)
// out : [nq][n_out][k]   where CTA (x, y) merges lists [x*lists_per_cta, ...) of query y and keeps the best k.
// When out_rows/out_scores/out_counts are given (final level, n_out == 1) the winners are decoded too.
constexpr int kMergeThreads = 256;
constexpr uint32_t kTournamentMaxK = 128;
__global__ void __launch_bounds__(kMergeThreads) merge_topk_kernel(const uint64_t* __restrict__ in, uint32_t n_lists,
                                                                    uint32_t list_len, uint32_t k, uint32_t lists_per_cta,
                                                                    uint32_t sort_n, uint64_t* __restrict__ out, int ascending,
                                                                    uint64_t* __restrict__ out_rows,
                                                                    float* __restrict__ out_scores,
                                                                    uint32_t* __restrict__ out_counts, int sorted_in,
                                                                    uint64_t* trace, uint64_t q_stride, uint64_t l_stride) {
    extern __shared__ __align__(16) uint64_t s_merge_keys[];
    uint64_t* s_keys = s_merge_keys;
    if (threadIdx.x == 0) trace_begin(trace);
    // Let a programmatically-dependent successor (the next query's scan, which does not read our output) start now.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // No-op unless this kernel was itself launched programmatically (PDL chain mode 2): then the producer of `in` may
    // still be running and must have completed (memory visible) before the first read below.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const uint32_t q = blockIdx.y, n_out = gridDim.x;
    const uint32_t first = blockIdx.x * lists_per_cta;
    const uint32_t lists = min(lists_per_cta, n_lists - first);
    const uint64_t* src = in + (size_t)q * q_stride + (size_t)first * l_stride;
    const uint32_t total = lists * list_len;
    const bool dense = l_stride == list_len;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t* best;                                       // the CTA's best-k keys, descending, 0-padded
    if (sorted_in && k <= kTournamentMaxK && lists <= 32 * (kMergeThreads / 32)) {
        // tournament: every warp reduces up to 32 sorted lists to one, then warp 0 reduces the <= 8 survivors
        uint64_t* lvl = s_keys + sort_n;                        // [8][k] after the staged lists
        for (uint32_t i = threadIdx.x; i < total; i += kMergeThreads) s_keys[i] = dense ? src[i] : src[(size_t)(i / list_len) * l_stride + i % list_len];
        __syncthreads();
        const uint32_t nw = (lists + 31) / 32;
        if (warp < nw)
            warp_tournament_topk(s_keys + (size_t)warp * 32 * list_len, min(32u, lists - warp * 32), list_len, list_len, k, lvl + (size_t)warp * k, lane);
        __syncthreads();
        if (nw > 1) {
            if (warp == 0) warp_tournament_topk(lvl, nw, k, k, k, s_keys, lane);
            __syncthreads();
            best = s_keys;
        } else {
            best = lvl;
        }
    } else {
        for (uint32_t i = threadIdx.x; i < sort_n; i += kMergeThreads)
            s_keys[i] = (i < total) ? (dense ? src[i] : src[(size_t)(i / list_len) * l_stride + i % list_len]) : 0ull;
        __syncthreads();
        bitonic_sort_desc(s_keys, sort_n, threadIdx.x, kMergeThreads, 0);
        best = s_keys;
    }
    const uint32_t avail = (sorted_in && k <= kTournamentMaxK && lists <= 32 * (kMergeThreads / 32)) ? k : sort_n;
    if (out) {
        uint64_t* dst = out + ((size_t)q * n_out + blockIdx.x) * k;
        for (uint32_t i = threadIdx.x; i < k; i += kMergeThreads) dst[i] = (i < avail) ? best[i] : 0ull;
    }
    if (out_rows || out_scores || out_counts) {
        uint32_t cnt = 0;
        for (uint32_t i = threadIdx.x; i < k; i += kMergeThreads) {
            const uint64_t key = (i < avail) ? best[i] : 0ull;
            const bool valid = key != 0ull;
            if (out_rows) out_rows[(size_t)q * k + i] = valid ? (uint64_t)key_row(key) : ~0ull;
            if (out_scores) out_scores[(size_t)q * k + i] = valid ? key_score(key, ascending != 0) : 0.0f;
            cnt += valid;
        }
        if (out_counts) {
            __shared__ uint32_t s_cnt;
            if (threadIdx.x == 0) s_cnt = 0;
            __syncthreads();
            if (cnt) atomicAdd(&s_cnt, cnt);
            __syncthreads();
            if (threadIdx.x == 0) out_counts[q] = s_cnt;
        }
    }
    if (threadIdx.x == 0) trace_end(trace);
}

// cosine DISTANCE of the first `limit` rows (gpu.rs:297-322 compute_distances_cpu), one octet lane 0 per row.
template <typename T>
__global__ void distances_first_kernel(const T* __restrict__ rows, uint32_t d, uint32_t ld, const float* __restrict__ q,
                                       uint64_t limit, float* __restrict__ out) {
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= limit) return;
    const T* row = rows + i * ld;
    float dp = 0.0f, na = 0.0f, nb = 0.0f;
    for (uint32_t j = 0; j < d; ++j) dp = add_rn(dp, mul_rn(q[j], ldf(row + j)));
    for (uint32_t j = 0; j < d; ++j) na = add_rn(na, mul_rn(q[j], q[j]));
    for (uint32_t j = 0; j < d; ++j) { float vb = ldf(row + j); nb = add_rn(nb, mul_rn(vb, vb)); }
    na = sqrt_rn(na);
    nb = sqrt_rn(nb);
    out[i] = (na == 0.0f || nb == 0.0f) ? __int_as_float(0x7f800000) : sub_rn(1.0f, div_rn(dp, mul_rn(na, nb)));
}

// f16 -> f32 widening of m rows (bulk read-back): out[r * d + c] = rows[r * ld + c]
__global__ void widen_rows_kernel(const __half* __restrict__ rows, uint32_t d, uint32_t ld, uint64_t m, float* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= m * d) return;
    const uint64_t r = i / d;
    out[i] = __half2float(rows[r * ld + (i - r * d)]);
}

// f16 -> f32 widening of one row (get_embedding)
__global__ void widen_row_kernel(const __half* __restrict__ row, uint32_t d, float* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d) out[i] = __half2float(row[i]);
}

}  // namespace cgv
