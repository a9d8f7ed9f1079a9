/*
 * cgvec_oracle.c — CPU ORACLE for the embedding similarity-search hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / `--impl reference` legs may load it.  The product
 * (codegraph-rust_b200/libcgvec_b200.so) never links, loads or calls anything here.
 *
 * It is a plain-C restatement (the reference is Rust; no rustc/cargo exists in the build
 * image, so the reference itself cannot be compiled — see DESIGN.md §oracle) of the
 * arithmetic in Jakedismo/codegraph-rust @ ce5bf27a.  Every function cites the reference
 * file:line it follows (paths relative to /root/reference/crates/).  Build flags matter:
 * -ffp-contract=off -fno-fast-math so that scalar `a*b + c` stays an un-fused multiply
 * and add exactly as rustc emits it, while the AVX2 path uses explicit FMA like the
 * reference's _mm256_fmadd_ps.
 *
 * Parity status: PINNED against the reference's own known-answer tests (tests/test_oracle_golden.py
 * replays simd_ops.rs:428-472, rag/context_retriever.rs:504-512, model_optimization_tests.rs:36-58,
 * 383-424 and search.rs:178-205); the reference binary itself could not be run here.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define CG_HAVE_AVX2 1
#else
#define CG_HAVE_AVX2 0
#endif

#define CG_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * AVX2 lane model.  `_mm256_fmadd_ps` is 8 independent IEEE fmaf chains, so a plain-C loop over
 * 8 lanes with fmaf() is bit-identical to the intrinsic version; both are provided and the
 * tests check they agree bit for bit (so the oracle does not depend on the host having AVX2).
 * ------------------------------------------------------------------------------------------ */

/* simd_ops.rs:224-242 horizontal_sum_avx2: permute2f128 + add, hadd, hadd
 *   = ((l0+l4) + (l1+l5)) + ((l2+l6) + (l3+l7))                                              */
static inline float cg_hsum8(const float l[8]) {
    float a0 = l[0] + l[4], a1 = l[1] + l[5], a2 = l[2] + l[6], a3 = l[3] + l[7];
    float h0 = a0 + a1, h1 = a2 + a3;
    return h0 + h1;
}

/* simd_ops.rs:15-78 cosine_similarity_avx2, lane-emulated. */
CG_API float cg_cosine_similarity_avx2_emul(const float* a, const float* b, size_t len) {
    if (len == 0) return 0.0f;                                   /* :21-23 */
    float dp[8] = {0}, na[8] = {0}, nb[8] = {0};
    size_t chunks = len / 8;                                     /* :30 */
    for (size_t i = 0; i < chunks; ++i) {                        /* :31-47 */
        for (int l = 0; l < 8; ++l) {
            float va = a[i * 8 + l], vb = b[i * 8 + l];
            dp[l] = fmaf(va, vb, dp[l]);
            na[l] = fmaf(va, va, na[l]);
            nb[l] = fmaf(vb, vb, nb[l]);
        }
    }
    float dps = cg_hsum8(dp), nas = cg_hsum8(na), nbs = cg_hsum8(nb);   /* :50-52 */
    float dr = 0.0f, nar = 0.0f, nbr = 0.0f;                             /* :55-65 scalar tail, un-fused */
    for (size_t i = chunks * 8; i < len; ++i) {
        float va = a[i], vb = b[i];
        dr = dr + va * vb;
        nar = nar + va * va;
        nbr = nbr + vb * vb;
    }
    float fdp = dps + dr, fna = nas + nar, fnb = nbs + nbr;              /* :67-69 */
    float np = sqrtf(fna * fnb);                                         /* :72 */
    if (np == 0.0f) return 0.0f;                                         /* :73-74 */
    return fdp / np;                                                     /* :76 */
}

#if CG_HAVE_AVX2
static inline float cg_hsum_avx2(__m256 v) {                             /* simd_ops.rs:227-242 */
    __m256 vp = _mm256_permute2f128_ps(v, v, 0x01);
    __m256 a1 = _mm256_add_ps(v, vp);
    __m256 h1 = _mm256_hadd_ps(a1, a1);
    __m256 h2 = _mm256_hadd_ps(h1, h1);
    return _mm256_cvtss_f32(h2);
}
#endif

/* simd_ops.rs:15-78 with the real intrinsics (falls back to the emulation without AVX2). */
CG_API float cg_cosine_similarity_avx2(const float* a, const float* b, size_t len) {
#if CG_HAVE_AVX2
    if (len == 0) return 0.0f;
    __m256 dp = _mm256_setzero_ps(), na = _mm256_setzero_ps(), nb = _mm256_setzero_ps();
    size_t chunks = len / 8;
    for (size_t i = 0; i < chunks; ++i) {
        __m256 va = _mm256_loadu_ps(a + i * 8), vb = _mm256_loadu_ps(b + i * 8);
        dp = _mm256_fmadd_ps(va, vb, dp);
        na = _mm256_fmadd_ps(va, va, na);
        nb = _mm256_fmadd_ps(vb, vb, nb);
    }
    float dps = cg_hsum_avx2(dp), nas = cg_hsum_avx2(na), nbs = cg_hsum_avx2(nb);
    float dr = 0.0f, nar = 0.0f, nbr = 0.0f;
    for (size_t i = chunks * 8; i < len; ++i) {
        float va = a[i], vb = b[i];
        dr = dr + va * vb;
        nar = nar + va * va;
        nbr = nbr + vb * vb;
    }
    float fdp = dps + dr, fna = nas + nar, fnb = nbs + nbr;
    float np = sqrtf(fna * fnb);
    if (np == 0.0f) return 0.0f;
    return fdp / np;
#else
    return cg_cosine_similarity_avx2_emul(a, b, len);
#endif
}

CG_API int cg_have_avx2(void) { return CG_HAVE_AVX2; }

/* simd_ops.rs:149-183 dot_product_avx2 */
CG_API float cg_dot_product_avx2(const float* a, const float* b, size_t len) {
    if (len == 0) return 0.0f;
    float dp[8] = {0};
    size_t chunks = len / 8;
    for (size_t i = 0; i < chunks; ++i)
        for (int l = 0; l < 8; ++l) dp[l] = fmaf(a[i * 8 + l], b[i * 8 + l], dp[l]);
    float s = cg_hsum8(dp), r = 0.0f;
    for (size_t i = chunks * 8; i < len; ++i) r = r + a[i] * b[i];
    return s + r;
}

/* simd_ops.rs:105-143 l2_distance_avx2 */
CG_API float cg_l2_distance_avx2(const float* a, const float* b, size_t len) {
    if (len == 0) return 0.0f;
    float acc[8] = {0};
    size_t chunks = len / 8;
    for (size_t i = 0; i < chunks; ++i)
        for (int l = 0; l < 8; ++l) {
            float diff = a[i * 8 + l] - b[i * 8 + l];
            acc[l] = fmaf(diff, diff, acc[l]);
        }
    float s = cg_hsum8(acc), r = 0.0f;
    for (size_t i = chunks * 8; i < len; ++i) {
        float diff = a[i] - b[i];
        r = r + diff * diff;
    }
    return sqrtf(s + r);
}

/* simd_ops.rs:189-222 normalize_avx2 (in place) */
CG_API void cg_normalize_avx2(float* v, size_t len) {
    if (len == 0) return;
    float nsq = cg_dot_product_avx2(v, v, len);
    if (nsq == 0.0f) return;
    float norm = sqrtf(nsq);
    float inv = 1.0f / norm;
    for (size_t i = 0; i < len; ++i) v[i] = v[i] * inv;   /* mul_ps and the scalar tail are the same op */
}

/* simd_ops.rs:257-278 cosine_similarity_scalar: sequential, un-fused, sqrt of the product */
CG_API float cg_cosine_similarity_scalar(const float* a, const float* b, size_t len) {
    float dp = 0.0f, na = 0.0f, nb = 0.0f;
    for (size_t i = 0; i < len; ++i) {
        float va = a[i], vb = b[i];
        dp = dp + va * vb;
        na = na + va * va;
        nb = nb + vb * vb;
    }
    float np = sqrtf(na * nb);
    if (np == 0.0f) return 0.0f;
    return dp / np;
}

/* simd_ops.rs:281-295 adaptive_cosine_similarity: AVX2 iff available && len >= 32.
 * The oracle models an AVX2+FMA host (the only kind the reference's SIMD path targets). */
CG_API float cg_adaptive_cosine_similarity(const float* a, const float* b, size_t len) {
    if (len >= 32) return cg_cosine_similarity_avx2(a, b, len);
    return cg_cosine_similarity_scalar(a, b, len);
}

/* search.rs:519-533 cosine_similarity (same body: reranker.rs:96-110, rag/context_retriever.rs:401-415,
 * gpu.rs:324-338 (as a distance), codegraph-core integration/graph_vector.rs:500-512):
 * sequential un-fused sums, product of the two square roots, 0 when a norm is 0. */
CG_API float cg_cosine_similarity_seq(const float* a, const float* b, size_t len) {
    float dp = 0.0f, na = 0.0f, nb = 0.0f;
    for (size_t i = 0; i < len; ++i) dp = dp + a[i] * b[i];
    for (size_t i = 0; i < len; ++i) na = na + a[i] * a[i];
    for (size_t i = 0; i < len; ++i) nb = nb + b[i] * b[i];
    na = sqrtf(na);
    nb = sqrtf(nb);
    if (na == 0.0f || nb == 0.0f) return 0.0f;
    return dp / (na * nb);
}

/* optimization.rs:404-418 cosine_distance (ModelOptimizer) == gpu.rs:324-338: INFINITY when a norm is 0 */
CG_API float cg_cosine_distance_seq(const float* a, const float* b, size_t len) {
    float dp = 0.0f, na = 0.0f, nb = 0.0f;
    for (size_t i = 0; i < len; ++i) dp = dp + a[i] * b[i];
    for (size_t i = 0; i < len; ++i) na = na + a[i] * a[i];
    for (size_t i = 0; i < len; ++i) nb = nb + b[i] * b[i];
    na = sqrtf(na);
    nb = sqrtf(nb);
    if (na == 0.0f || nb == 0.0f) return INFINITY;
    return 1.0f - (dp / (na * nb));
}

/* ------------------------------------------------------------------------------------------
 * Ordering contract: the reference's comparators leave ties and NaN
 * unspecified (simd_ops.rs:379 is an UNSTABLE sort that panics on NaN; optimization.rs:393 is
 * stable with NaN==Equal).  The pinned contract is: better score first; ties -> lower row
 * index; NaN ranks after every number (and among NaNs lower index first); -0.0 == +0.0.
 * ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t idx; float score; } cg_pair;

static int cg_cmp_desc(const void* pa, const void* pb) {     /* higher score first */
    const cg_pair* a = (const cg_pair*)pa; const cg_pair* b = (const cg_pair*)pb;
    int an = isnan(a->score), bn = isnan(b->score);
    if (an != bn) return an ? 1 : -1;
    if (!an) {
        if (a->score > b->score) return -1;
        if (a->score < b->score) return 1;
    }
    return (a->idx > b->idx) - (a->idx < b->idx);
}
static int cg_cmp_asc(const void* pa, const void* pb) {      /* lower distance first */
    const cg_pair* a = (const cg_pair*)pa; const cg_pair* b = (const cg_pair*)pb;
    int an = isnan(a->score), bn = isnan(b->score);
    if (an != bn) return an ? 1 : -1;
    if (!an) {
        if (a->score < b->score) return -1;
        if (a->score > b->score) return 1;
    }
    return (a->idx > b->idx) - (a->idx < b->idx);
}

enum { CG_COSINE = 0, CG_DOT = 1, CG_L2 = 2 };
enum { CG_FORM_ADAPTIVE = 0,   /* simd_ops.rs:281-295 (AVX2 if len>=32 else scalar) */
       CG_FORM_SCALAR = 1,     /* simd_ops.rs:257-278 */
       CG_FORM_SEQ = 2 };      /* search.rs:519-533 / optimization.rs:404-418 */

static float cg_score_one(const float* q, const float* row, size_t d, int metric, int form) {
    if (metric == CG_DOT) {
        if (form == CG_FORM_ADAPTIVE) return cg_dot_product_avx2(q, row, d);
        float s = 0.0f; for (size_t i = 0; i < d; ++i) s = s + q[i] * row[i]; return s;
    }
    if (metric == CG_L2) return cg_l2_distance_avx2(q, row, d);
    if (form == CG_FORM_ADAPTIVE) return cg_adaptive_cosine_similarity(q, row, d);
    if (form == CG_FORM_SCALAR) return cg_cosine_similarity_scalar(q, row, d);
    return cg_cosine_similarity_seq(q, row, d);
}

/* All-pairs scores of one query against a flat row-major matrix (ld floats between rows). */
CG_API void cg_scores(const float* q, const float* rows, uint64_t n, size_t d, size_t ld,
                      int metric, int form, float* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = cg_score_one(q, rows + i * ld, d, metric, form);
}

/* simd_ops.rs:361-383 parallel_top_k_search: score every row with adaptive_cosine_similarity,
 * sort ALL pairs descending, truncate(k).  Single-threaded restatement used by the parity tests;
 * the ordering contract above replaces the reference's unspecified tie/NaN behaviour.
 * Returns min(k, n).  metric/form extend it to dot / L2 (ascending) with the same skeleton. */
CG_API uint64_t cg_parallel_top_k_search(const float* q, const float* rows, uint64_t n, size_t d, size_t ld,
                                         uint64_t k, int metric, int form,
                                         uint64_t* out_idx, float* out_score) {
    if (n == 0 || k == 0) return 0;
    cg_pair* p = (cg_pair*)malloc(sizeof(cg_pair) * n);
    for (uint64_t i = 0; i < n; ++i) { p[i].idx = i; p[i].score = cg_score_one(q, rows + i * ld, d, metric, form); }
    qsort(p, n, sizeof(cg_pair), metric == CG_L2 ? cg_cmp_asc : cg_cmp_desc);
    uint64_t m = k < n ? k : n;
    for (uint64_t i = 0; i < m; ++i) { out_idx[i] = p[i].idx; out_score[i] = p[i].score; }
    free(p);
    return m;
}

/* optimization.rs:376-402 search_baseline: cosine_distance per row, STABLE ascending sort, take(limit). */
CG_API uint64_t cg_search_baseline(const float* q, const float* rows, uint64_t n, size_t d, size_t ld,
                                   uint64_t limit, uint64_t* out_idx, float* out_dist) {
    if (n == 0) return 0;                                            /* :382-384 */
    cg_pair* p = (cg_pair*)malloc(sizeof(cg_pair) * n);
    for (uint64_t i = 0; i < n; ++i) { p[i].idx = i; p[i].score = cg_cosine_distance_seq(q, rows + i * ld, d); }
    qsort(p, n, sizeof(cg_pair), cg_cmp_asc);                        /* stable == (dist, idx) ascending */
    uint64_t m = limit < n ? limit : n;
    for (uint64_t i = 0; i < m; ++i) { out_idx[i] = p[i].idx; if (out_dist) out_dist[i] = p[i].score; }
    free(p);
    return m;
}

/* codegraph-core/src/integration/graph_vector.rs:479-494 InMemoryVectorStore::search_similar:
 * cosine (search.rs:519 form) per stored row, stable sort descending, take(limit). */
CG_API uint64_t cg_inmemory_search_similar(const float* q, const float* rows, uint64_t n, size_t d, size_t ld,
                                           uint64_t limit, uint64_t* out_idx, float* out_score) {
    if (n == 0 || limit == 0) return 0;
    cg_pair* p = (cg_pair*)malloc(sizeof(cg_pair) * n);
    for (uint64_t i = 0; i < n; ++i) { p[i].idx = i; p[i].score = cg_cosine_similarity_seq(rows + i * ld, q, d); }
    qsort(p, n, sizeof(cg_pair), cg_cmp_desc);
    uint64_t m = limit < n ? limit : n;
    for (uint64_t i = 0; i < m; ++i) { out_idx[i] = p[i].idx; if (out_score) out_score[i] = p[i].score; }
    free(p);
    return m;
}

/* gpu.rs:297-322 compute_distances_cpu: cosine DISTANCE of the first `limit` rows (not a top-k). */
CG_API uint64_t cg_compute_distances_cpu(const float* q, const float* rows, uint64_t n, size_t d,
                                         uint64_t limit, float* out) {
    uint64_t m = limit < n ? limit : n;
    for (uint64_t i = 0; i < m; ++i) out[i] = cg_cosine_distance_seq(q, rows + i * d, d);
    return m;
}

/* search.rs:574-592 normalize_scores: min-max to [0,1], range floored at 1e-12. */
CG_API void cg_normalize_scores(float* s, size_t n) {
    if (n == 0) return;
    float mn = INFINITY, mx = -INFINITY;
    for (size_t i = 0; i < n; ++i) { if (s[i] < mn) mn = s[i]; if (s[i] > mx) mx = s[i]; }
    float range = mx - mn; if (!(range > 1e-12f)) range = 1e-12f;    /* f32::max(range, 1e-12) */
    for (size_t i = 0; i < n; ++i) s[i] = (s[i] - mn) / range;
}

/* search.rs:113 / :276 prefetch sizes used by SemanticSearch */
CG_API uint64_t cg_prefetch_k_basic(uint64_t limit) {
    uint64_t a = limit * 3, b = limit + 10; return a > b ? a : b; }
CG_API uint64_t cg_prefetch_k_filtered(uint64_t limit) {
    uint64_t a = limit * 4, b = limit + 25; return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------
 * int8 quantised scan ("next" row f-3).
 * optimization.rs:212-224 quantize_unit_range_symmetric, :268-274 (+128 offset to u8),
 * optimization.rs:63-150 search_optimized.
 * ------------------------------------------------------------------------------------------ */
static inline float cg_round_half_away(float x) { return roundf(x); }   /* f32::round */
static inline float cg_clampf(float v, float lo, float hi) {            /* f32::clamp; NaN stays NaN */
    if (v < lo) return lo; if (v > hi) return hi; return v; }
static inline int32_t cg_f32_as_i32(float x) {                          /* Rust `as i32`: saturating, NaN -> 0 */
    if (isnan(x)) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
}

CG_API int8_t cg_quantize_unit_i8(float val) {                          /* optimization.rs:212-224, bits=8 */
    float c = cg_clampf(val, -1.0f, 1.0f);
    int32_t q = cg_f32_as_i32(cg_round_half_away(c * 127.0f));
    if (q < -127) q = -127; if (q > 127) q = 127;
    return (int8_t)q;
}
CG_API void cg_quantize_batch_u8(const float* rows, uint64_t n, size_t d, uint8_t* out) {  /* :268-274 */
    for (uint64_t i = 0; i < n * d; ++i) out[i] = (uint8_t)((int16_t)cg_quantize_unit_i8(rows[i]) + 128);
}

/* optimization.rs:63-150 search_optimized (bits == 8).  The reference keeps a running list whose
 * minimum sits at index 0 and re-sorts it (stable, ascending) on every replacement; replacement
 * needs score > min strictly.  Final order: stable sort descending.  Reproduced literally so that
 * tie behaviour is the reference's own.  Returns the number of indices written. */
static void cg_stable_sort_pairs(cg_pair* p, size_t n, int desc) {      /* insertion sort == stable; n <= limit */
    for (size_t i = 1; i < n; ++i) {
        cg_pair x = p[i]; size_t j = i;
        while (j > 0 && (desc ? (p[j - 1].score < x.score) : (p[j - 1].score > x.score))) { p[j] = p[j - 1]; --j; }
        p[j] = x;
    }
}
CG_API uint64_t cg_search_optimized_i8(const float* query, size_t qlen, const uint8_t* codes, uint64_t n, size_t d,
                                       uint64_t limit_in, uint64_t* out_idx, float* out_score) {
    uint64_t limit = limit_in < 1 ? 1 : limit_in;                        /* :64 */
    if (d == 0 || n == 0) return 0;                                      /* :70-72, :92-94 */
    int8_t* qq = (int8_t*)calloc(d, 1);
    for (size_t i = 0; i < d && i < qlen; ++i) qq[i] = cg_quantize_unit_i8(query[i]);      /* :96-106 */
    float nq = 0.0f;
    for (size_t i = 0; i < d; ++i) nq = nq + (float)qq[i] * (float)qq[i];                  /* :108-112 */
    nq = sqrtf(nq);
    if (nq == 0.0f) { free(qq); return 0; }                              /* :113-115 */
    uint64_t cap = limit < n ? limit : n;
    cg_pair* best = (cg_pair*)malloc(sizeof(cg_pair) * (cap ? cap : 1));
    uint64_t nb = 0;
    for (uint64_t idx = 0; idx < n; ++idx) {                             /* :119-147 */
        const uint8_t* row = codes + idx * d;
        int32_t dot = 0, nv = 0;
        for (size_t j = 0; j < d; ++j) {
            int32_t v = (int32_t)row[j] - 128, q = (int32_t)qq[j];
            dot += v * q; nv += v * v;
        }
        if (nv == 0) continue;
        float score = (float)dot / (nq * sqrtf((float)nv));
        if (nb < limit) {
            best[nb].idx = idx; best[nb].score = score; ++nb;
            if (nb == limit) cg_stable_sort_pairs(best, nb, 0);
        } else if (score > best[0].score) {
            best[0].idx = idx; best[0].score = score;
            cg_stable_sort_pairs(best, nb, 0);
        }
    }
    cg_stable_sort_pairs(best, nb, 1);                                   /* :149 */
    for (uint64_t i = 0; i < nb; ++i) { out_idx[i] = best[i].idx; if (out_score) out_score[i] = best[i].score; }
    free(best); free(qq);
    return nb;
}

/* ------------------------------------------------------------------------------------------
 * Deterministic fixtures used by the reference's own tests.
 * ------------------------------------------------------------------------------------------ */
/* Rust std DefaultHasher = SipHash-1-3, key (0,0).  u64/usize `hash()` feeds 8 LE bytes. */
#define ROTL(x, b) (uint64_t)(((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND do { v0 += v1; v1 = ROTL(v1, 13); v1 ^= v0; v0 = ROTL(v0, 32); v2 += v3; v3 = ROTL(v3, 16); v3 ^= v2; \
                      v0 += v3; v3 = ROTL(v3, 21); v3 ^= v0; v2 += v1; v1 = ROTL(v1, 17); v1 ^= v2; v2 = ROTL(v2, 32); } while (0)
CG_API uint64_t cg_siphash13_2xu64(uint64_t m0, uint64_t m1) {
    uint64_t v0 = 0x736f6d6570736575ULL, v1 = 0x646f72616e646f6dULL, v2 = 0x6c7967656e657261ULL, v3 = 0x7465646279746573ULL;
    uint64_t m[2] = {m0, m1};
    for (int i = 0; i < 2; ++i) { v3 ^= m[i]; SIPROUND; v0 ^= m[i]; }
    uint64_t b = ((uint64_t)16) << 56;                                   /* total length 16, no tail bytes */
    v3 ^= b; SIPROUND; v0 ^= b;
    v2 ^= 0xff; SIPROUND; SIPROUND; SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}
/* codegraph-vector/tests/model_optimization_tests.rs:36-58 generate_optimization_vectors */
CG_API void cg_generate_optimization_vectors(uint64_t count, uint64_t dim, uint64_t seed, float* out) {
    for (uint64_t i = 0; i < count; ++i) {
        uint64_t h = cg_siphash13_2xu64(seed, i);
        for (uint64_t j = 0; j < dim; ++j) {
            uint64_t hj = cg_siphash13_2xu64(h, j);
            float val = ((float)hj / 18446744073709551616.0f) - 0.5f;    /* u64::MAX as f32 == 2^64 */
            out[i * dim + j] = val * 2.0f;
        }
    }
}
/* search.rs:178-205 encode_query fallback + :535-541 simple_hash (djb2, u32 wrapping) + LCG */
CG_API void cg_hash_text_embedding(const uint8_t* text, size_t n, size_t dimension, float* out) {
    uint32_t h = 5381u;
    for (size_t i = 0; i < n; ++i) h = h * 33u + (uint32_t)text[i];
    uint32_t s = h;
    for (size_t i = 0; i < dimension; ++i) {
        s = s * 1103515245u + 12345u;
        out[i] = (((float)s / 4294967296.0f) - 0.5f) * 2.0f;             /* u32::MAX as f32 == 2^32 */
    }
    float norm = 0.0f;
    for (size_t i = 0; i < dimension; ++i) norm = norm + out[i] * out[i];
    norm = sqrtf(norm);
    if (norm > 0.0f) for (size_t i = 0; i < dimension; ++i) out[i] = out[i] / norm;
}

/* ------------------------------------------------------------------------------------------
 * IEEE binary16 <-> binary32 (software; the fp16 configs have no reference counterpart — the
 * oracle is "widen the stored halves exactly to f32, then the f32 functions above").
 * ------------------------------------------------------------------------------------------ */
CG_API float cg_half_to_float(uint16_t h) {
    uint32_t s = (uint32_t)(h >> 15) << 31, e = (h >> 10) & 0x1f, m = h & 0x3ff, u;
    if (e == 0) {
        if (m == 0) u = s;
        else { int sh = 0; while (!(m & 0x400)) { m <<= 1; ++sh; } m &= 0x3ff; u = s | ((uint32_t)(113 - sh) << 23) | (m << 13); }
    } else if (e == 31) u = s | 0x7f800000u | (m << 13);
    else u = s | ((e + 112) << 23) | (m << 13);
    float f; memcpy(&f, &u, 4); return f;
}
CG_API uint16_t cg_float_to_half(float f) {                              /* round to nearest even */
    uint32_t u; memcpy(&u, &f, 4);
    uint32_t s = (u >> 16) & 0x8000u; int32_t e = (int32_t)((u >> 23) & 0xff) - 127 + 15; uint32_t m = u & 0x7fffffu;
    if (((u >> 23) & 0xff) == 0xff) return (uint16_t)(s | 0x7c00u | (m ? (0x200u | (m >> 13)) : 0));
    if (e >= 31) return (uint16_t)(s | 0x7c00u);
    if (e <= 0) {
        if (e < -10) return (uint16_t)s;
        m |= 0x800000u; int sh = 14 - e;
        uint32_t r = m >> sh, rem = m & ((1u << sh) - 1), half = 1u << (sh - 1);
        if (rem > half || (rem == half && (r & 1))) ++r;
        return (uint16_t)(s | r);
    }
    uint32_t r = ((uint32_t)e << 10) | (m >> 13), rem = m & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1))) ++r;
    return (uint16_t)(s | r);
}
CG_API void cg_widen_f16(const uint16_t* in, uint64_t n, float* out) { for (uint64_t i = 0; i < n; ++i) out[i] = cg_half_to_float(in[i]); }
CG_API void cg_narrow_f16(const float* in, uint64_t n, uint16_t* out) { for (uint64_t i = 0; i < n; ++i) out[i] = cg_float_to_half(in[i]); }

/* ------------------------------------------------------------------------------------------
 * "ref-parallel": the timed CPU baseline — simd_ops.rs:361-383 AS WRITTEN: rows are separate heap
 * allocations (&[Vec<f32>]), every row is scored with the 3-FMA AVX2 cosine (‖q‖² recomputed
 * per row), N (idx, score) pairs are collected, ALL of them are sorted descending in parallel,
 * then truncated to k.  rayon's par_iter / par_sort_unstable_by are restated with a pthread fork-join:
 * a parallel-for over rows, then per-thread qsort of contiguous blocks + a k-bounded merge of
 * the block heads (equivalent output: the first k of the fully sorted list).
 * ------------------------------------------------------------------------------------------ */
typedef struct { float** rows; uint64_t n; size_t d; } cg_vecs;

CG_API void* cg_vecs_create(const float* flat, uint64_t n, size_t d) {   /* Vec<Vec<f32>> */
    cg_vecs* v = (cg_vecs*)malloc(sizeof(cg_vecs));
    v->rows = (float**)malloc(sizeof(float*) * (n ? n : 1)); v->n = n; v->d = d;
    for (uint64_t i = 0; i < n; ++i) {
        v->rows[i] = (float*)malloc(sizeof(float) * (d ? d : 1));
        memcpy(v->rows[i], flat + i * d, sizeof(float) * d);
    }
    return v;
}
CG_API void cg_vecs_destroy(void* h) {
    cg_vecs* v = (cg_vecs*)h; if (!v) return;
    for (uint64_t i = 0; i < v->n; ++i) free(v->rows[i]);
    free(v->rows); free(v);
}
/* Threads worth starting: online CPUs, clipped by the affinity mask AND by the cgroup CPU quota (cpu.max of cgroup v2,
 * cpu.cfs_quota_us / cpu.cfs_period_us of v1): a container limited to 32 CPUs' worth of time on a 128-thread host
 * gains nothing from 128 threads, and a baseline that reports "128 cores" there misstates what it ran on. */
static long cg_cgroup_cpu_limit(void) {
    FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r");
    if (f) {
        char a[64]; long period = 0; long lim = 0;
        if (fscanf(f, "%63s %ld", a, &period) == 2 && strcmp(a, "max") != 0 && period > 0) {
            long quota = atol(a);
            if (quota > 0) lim = (quota + period - 1) / period;
        }
        fclose(f);
        if (lim > 0) return lim;
    }
    long quota = 0, period = 0;
    f = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r");
    if (f) { if (fscanf(f, "%ld", &quota) != 1) quota = 0; fclose(f); }
    f = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r");
    if (f) { if (fscanf(f, "%ld", &period) != 1) period = 0; fclose(f); }
    if (quota > 0 && period > 0) return (quota + period - 1) / period;
    return 0;
}
CG_API int cg_affinity_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) { int c = CPU_COUNT(&set); if (c > 0 && c < n) n = c; }
    return n > 0 ? (int)n : 1;
}
CG_API int cg_max_threads(void) {
    long n = cg_affinity_threads();
    long q = cg_cgroup_cpu_limit();
    if (q > 0 && q < n) n = q;
    return n > 0 ? (int)n : 1;
}
/* minimal fork-join over T workers (stands in for rayon's pool) */
typedef void (*cg_job_fn)(int t, int T, void* arg);
typedef struct { cg_job_fn fn; int t, T; void* arg; } cg_job;
static void* cg_job_tramp(void* p) { cg_job* j = (cg_job*)p; j->fn(j->t, j->T, j->arg); return NULL; }
static void cg_fork_join(int T, cg_job_fn fn, void* arg) {
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * T);
    cg_job* jb = (cg_job*)malloc(sizeof(cg_job) * T);
    for (int t = 0; t < T; ++t) { jb[t].fn = fn; jb[t].t = t; jb[t].T = T; jb[t].arg = arg; }
    for (int t = 1; t < T; ++t) pthread_create(&th[t], NULL, cg_job_tramp, &jb[t]);
    cg_job_tramp(&jb[0]);
    for (int t = 1; t < T; ++t) pthread_join(th[t], NULL);
    free(jb); free(th);
}
typedef struct { const float* q; cg_vecs* v; const float* flat; cg_pair* p; uint64_t n, blk, m; size_t d; uint64_t* cnt; } cg_mt;
static void cg_mt_score(int t, int T, void* a) {
    cg_mt* c = (cg_mt*)a; (void)T;
    uint64_t lo = (uint64_t)t * c->blk, hi = lo + c->blk > c->n ? c->n : lo + c->blk;
    for (uint64_t i = lo; i < hi; ++i) { c->p[i].idx = i; c->p[i].score = cg_adaptive_cosine_similarity(c->q, c->v->rows[i], c->d); }
}
static void cg_mt_sort(int t, int T, void* a) {
    cg_mt* c = (cg_mt*)a; (void)T;
    uint64_t lo = (uint64_t)t * c->blk, hi = lo + c->blk > c->n ? c->n : lo + c->blk;
    if (lo < hi) qsort(c->p + lo, hi - lo, sizeof(cg_pair), cg_cmp_desc);
}
CG_API uint64_t cg_parallel_top_k_search_mt(const float* q, void* h, uint64_t k, int threads,
                                            uint64_t* out_idx, float* out_score) {
    cg_vecs* v = (cg_vecs*)h; uint64_t n = v->n; size_t d = v->d;
    if (n == 0 || k == 0) return 0;
    cg_pair* p = (cg_pair*)malloc(sizeof(cg_pair) * n);
    int T = threads > 0 ? threads : cg_max_threads();
    if ((uint64_t)T > n) T = (int)n;
    uint64_t blk = (n + T - 1) / T;
    cg_mt ctx; memset(&ctx, 0, sizeof(ctx)); ctx.q = q; ctx.v = v; ctx.p = p; ctx.n = n; ctx.blk = blk; ctx.d = d;
    cg_fork_join(T, cg_mt_score, &ctx);      /* par_iter().enumerate().map(..).collect() */
    cg_fork_join(T, cg_mt_sort, &ctx);       /* par_sort_unstable_by: block sorts ... */
    uint64_t* head = (uint64_t*)calloc(T, sizeof(uint64_t));
    uint64_t m = k < n ? k : n;
    for (uint64_t o = 0; o < m; ++o) {
        int bt = -1;
        for (int t = 0; t < T; ++t) {
            uint64_t lo = (uint64_t)t * blk, hi = lo + blk > n ? n : lo + blk;
            if (lo + head[t] >= hi) continue;
            if (bt < 0 || cg_cmp_desc(&p[lo + head[t]], &p[(uint64_t)bt * blk + head[bt]]) < 0) bt = t;
        }
        cg_pair* w = &p[(uint64_t)bt * blk + head[bt]]; head[bt]++;
        out_idx[o] = w->idx; out_score[o] = w->score;
    }
    free(head); free(p);
    return m;
}

/* "fair-cpu" (labelled NON-reference in BASELINE.md §2): contiguous matrix, ‖q‖² hoisted out of the
 * loop is NOT possible without changing rounding, so this variant keeps the reference arithmetic and
 * only removes the avoidable overheads: no per-row heap allocation, no full sort (per-thread bounded
 * selection).  Same outputs as cg_parallel_top_k_search. */
static void cg_mt_fair(int t, int T, void* a) {
    cg_mt* c = (cg_mt*)a; (void)T;
    uint64_t lo = (uint64_t)t * c->blk, hi = lo + c->blk > c->n ? c->n : lo + c->blk, m = c->m;
    cg_pair* b = c->p + (uint64_t)t * m; uint64_t cn = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        cg_pair x; x.idx = i; x.score = cg_adaptive_cosine_similarity(c->q, c->flat + i * c->d, c->d);
        if (cn == m && cg_cmp_desc(&x, &b[cn - 1]) >= 0) continue;
        uint64_t j = cn < m ? cn++ : cn - 1;
        while (j > 0 && cg_cmp_desc(&x, &b[j - 1]) < 0) { b[j] = b[j - 1]; --j; }
        b[j] = x;
    }
    c->cnt[t] = cn;
}
CG_API uint64_t cg_fair_top_k_search_mt(const float* q, const float* rows, uint64_t n, size_t d, uint64_t k,
                                        int threads, uint64_t* out_idx, float* out_score) {
    if (n == 0 || k == 0) return 0;
    int T = threads > 0 ? threads : cg_max_threads();
    if ((uint64_t)T > n) T = (int)n;
    uint64_t m = k < n ? k : n;
    cg_pair* part = (cg_pair*)malloc(sizeof(cg_pair) * m * T);
    uint64_t* cnt = (uint64_t*)calloc(T, sizeof(uint64_t));
    uint64_t blk = (n + T - 1) / T;
    cg_mt ctx; memset(&ctx, 0, sizeof(ctx)); ctx.q = q; ctx.flat = rows; ctx.p = part; ctx.n = n; ctx.blk = blk; ctx.d = d; ctx.m = m; ctx.cnt = cnt;
    cg_fork_join(T, cg_mt_fair, &ctx);
    uint64_t tot = 0; for (int t = 0; t < T; ++t) tot += cnt[t];
    cg_pair* all = (cg_pair*)malloc(sizeof(cg_pair) * (tot ? tot : 1)); uint64_t o = 0;
    for (int t = 0; t < T; ++t) { memcpy(all + o, part + (uint64_t)t * m, sizeof(cg_pair) * cnt[t]); o += cnt[t]; }
    qsort(all, tot, sizeof(cg_pair), cg_cmp_desc);
    uint64_t r = m < tot ? m : tot;
    for (uint64_t i = 0; i < r; ++i) { out_idx[i] = all[i].idx; out_score[i] = all[i].score; }
    free(all); free(cnt); free(part);
    return r;
}

/* The same function for nq queries in ONE pass over the rows (each row is scored against every query while it is in
 * cache): the bench's parity legs check several queries of a batch against multi-GB shards, and a pass per query is
 * bound by host memory bandwidth.  Outputs are cg_parallel_top_k_search's, per query: out_idx / out_score [nq][k]
 * (entries beyond out_cnt[q] are untouched). */
typedef struct { const float* qs; uint64_t nq; const float* flat; cg_pair* p; uint64_t n, blk, m; size_t d; uint64_t* cnt; int T; } cg_mtq;
static void cg_mt_fair_multi(int t, int T, void* a) {
    cg_mtq* c = (cg_mtq*)a; (void)T;
    uint64_t lo = (uint64_t)t * c->blk, hi = lo + c->blk > c->n ? c->n : lo + c->blk, m = c->m;
    for (uint64_t qi = 0; qi < c->nq; ++qi) c->cnt[qi * c->T + t] = 0;
    for (uint64_t i = lo; i < hi; ++i) {
        const float* row = c->flat + i * c->d;
        for (uint64_t qi = 0; qi < c->nq; ++qi) {
            cg_pair* b = c->p + (qi * c->T + (uint64_t)t) * m; uint64_t cn = c->cnt[qi * c->T + t];
            cg_pair x; x.idx = i; x.score = cg_adaptive_cosine_similarity(c->qs + qi * c->d, row, c->d);
            if (cn == m && cg_cmp_desc(&x, &b[cn - 1]) >= 0) continue;
            uint64_t j = cn < m ? cn++ : cn - 1;
            while (j > 0 && cg_cmp_desc(&x, &b[j - 1]) < 0) { b[j] = b[j - 1]; --j; }
            b[j] = x;
            c->cnt[qi * c->T + t] = cn;
        }
    }
}
CG_API void cg_fair_top_k_search_multi(const float* qs, uint64_t nq, const float* rows, uint64_t n, size_t d, uint64_t k,
                                       int threads, uint64_t* out_idx, float* out_score, uint64_t* out_cnt) {
    for (uint64_t qi = 0; qi < nq; ++qi) out_cnt[qi] = 0;
    if (n == 0 || k == 0 || nq == 0) return;
    int T = threads > 0 ? threads : cg_max_threads();
    if ((uint64_t)T > n) T = (int)n;
    uint64_t m = k < n ? k : n;
    cg_pair* part = (cg_pair*)malloc(sizeof(cg_pair) * m * T * nq);
    uint64_t* cnt = (uint64_t*)calloc((size_t)T * nq, sizeof(uint64_t));
    uint64_t blk = (n + T - 1) / T;
    cg_mtq ctx; memset(&ctx, 0, sizeof(ctx)); ctx.qs = qs; ctx.nq = nq; ctx.flat = rows; ctx.p = part; ctx.n = n; ctx.blk = blk; ctx.d = d; ctx.m = m; ctx.cnt = cnt; ctx.T = T;
    cg_fork_join(T, cg_mt_fair_multi, &ctx);
    cg_pair* all = (cg_pair*)malloc(sizeof(cg_pair) * m * T);
    for (uint64_t qi = 0; qi < nq; ++qi) {
        uint64_t o = 0;
        for (int t = 0; t < T; ++t) { memcpy(all + o, part + (qi * T + (uint64_t)t) * m, sizeof(cg_pair) * cnt[qi * T + t]); o += cnt[qi * T + t]; }
        qsort(all, o, sizeof(cg_pair), cg_cmp_desc);
        uint64_t r = m < o ? m : o;
        for (uint64_t i = 0; i < r; ++i) { out_idx[qi * k + i] = all[i].idx; out_score[qi * k + i] = all[i].score; }
        out_cnt[qi] = r;
    }
    free(all); free(cnt); free(part);
}

/* ------------------------------------------------------------------------------------------
 * Synthetic inputs (bench / property tests): host twin of the device generator
 * (codegraph-rust_b200/csrc/aux_kernels.cuh synth_value + normalize_avx2 arithmetic).  Not a reference
 * function: it only produces the seeded inputs both sides consume.
 * ------------------------------------------------------------------------------------------ */
static inline float cg_synth_value(uint64_t seed, uint64_t row, uint32_t col) {
    uint64_t z = seed ^ (row * 0x9E3779B97F4A7C15ULL) ^ ((uint64_t)col * 0xC2B2AE3D27D4EB4FULL);
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    int32_t s = (int32_t)(z & 0xffff) + (int32_t)((z >> 16) & 0xffff) + (int32_t)((z >> 32) & 0xffff) +
                (int32_t)((z >> 48) & 0xffff) - 131070;
    return (float)s * (1.0f / 65536.0f);
}
typedef struct { float* out; uint64_t first, n; size_t d; uint64_t seed; int unit_norm, f16; } cg_synth_ctx;
static void cg_synth_job(int t, int T, void* a) {
    cg_synth_ctx* c = (cg_synth_ctx*)a;
    uint64_t blk = (c->n + T - 1) / T, lo = (uint64_t)t * blk, hi = lo + blk > c->n ? c->n : lo + blk;
    for (uint64_t i = lo; i < hi; ++i) {
        float* v = c->out + i * c->d;
        for (size_t j = 0; j < c->d; ++j) v[j] = cg_synth_value(c->seed, c->first + i, (uint32_t)j);
        if (c->unit_norm) {                                          /* simd_ops.rs:189-222 normalize_avx2 */
            float nsq = cg_dot_product_avx2(v, v, c->d);
            if (nsq != 0.0f) { float inv = 1.0f / sqrtf(nsq); for (size_t j = 0; j < c->d; ++j) v[j] = v[j] * inv; }
        }
        if (c->f16) for (size_t j = 0; j < c->d; ++j) v[j] = cg_half_to_float(cg_float_to_half(v[j]));
    }
}
CG_API void cg_synth_rows(uint64_t seed, uint64_t first_row, uint64_t n, size_t d, int unit_norm, int round_f16,
                          int threads, float* out) {
    cg_synth_ctx c = {out, first_row, n, d, seed, unit_norm, round_f16};
    int T = threads > 0 ? threads : cg_max_threads();
    if ((uint64_t)T > n) T = n ? (int)n : 1;
    cg_fork_join(T, cg_synth_job, &c);
}
