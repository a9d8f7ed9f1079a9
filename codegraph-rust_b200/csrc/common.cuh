// common.cuh — device helpers shared by the sm_100a kernels of libcgvec_b200:
// mbarrier / bulk-async-copy (TMA engine, SASS UBLKCP) wrappers, the IEEE-exact arithmetic
// primitives that mirror the reference's AVX2 / scalar operation order, and the 64-bit sort keys
// that implement the result contract (best first, ties -> lower row, NaN last).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cgv {

enum : int { METRIC_COSINE = 0, METRIC_DOT = 1, METRIC_L2 = 2, METRIC_I8 = 3 };   // I8: the int8 cosine of search_optimized (optimization.rs:119-137)
enum : int { FORM_SIMD = 0, FORM_SCALAR = 1, FORM_SEQ = 2, FORM_BASELINE = 3 };

// ---------------------------------------------------------------------------------------------
// mbarrier + cp.async.bulk (non-tensor TMA).  One elected lane arms the barrier with the byte
// count, the copies complete_tx on it, consumers spin on try_wait.parity (HW-assisted wait).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy executed by the TMA engine; bytes % 16 == 0, both addresses 16B aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// Optional launch timeline (cgvec_set_option "trace"): every CTA folds the global nanosecond timer into a
// [first start, last end] pair of its launch; used by tools/trace_steps.py to see gaps between kernels.
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_begin(uint64_t* slot) { if (slot) atomicMin(reinterpret_cast<unsigned long long*>(slot), (unsigned long long)global_ns()); }
__device__ __forceinline__ void trace_end(uint64_t* slot) { if (slot) atomicMax(reinterpret_cast<unsigned long long*>(slot) + 1, (unsigned long long)global_ns()); }

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Element access: the resident matrix is f32 or f16; f16 widens exactly to f32 (the fp16 oracle
// is "widen, then the f32 reference arithmetic").
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __half* p) { return __half2float(*p); }

// Exact-order building blocks.  All explicit _rn intrinsics: nvcc never contracts or reorders them.
__device__ __forceinline__ float fma_rn(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }

// horizontal_sum_avx2 (reference simd_ops.rs:227-242) across the 8 lanes of an aligned lane octet:
// ((l0+l4)+(l1+l5)) + ((l2+l6)+(l3+l7)); the result is valid on octet lane 0 (and 4).
__device__ __forceinline__ float hsum8_ref_order(float v) {
    v = add_rn(v, __shfl_xor_sync(0xffffffffu, v, 4));
    v = add_rn(v, __shfl_xor_sync(0xffffffffu, v, 1));
    v = add_rn(v, __shfl_xor_sync(0xffffffffu, v, 2));
    return v;
}

// ---------------------------------------------------------------------------------------------
// Sort keys.  key = ord(score) << 32 | ~row : a larger key is a better hit; among equal scores the
// LOWER row wins; NaN maps to ord 0 (ranks after every number); key 0 means "empty slot".
// -0.0 is canonicalised to +0.0 (partial_cmp treats them as equal).  Ascending metrics (L2,
// BASELINE distance) key on the negated value.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ord_desc(float s) {
    if (s != s) return 0u;
    s = __fadd_rn(s, 0.0f);
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t row, bool ascending) {
    float s = ascending ? -score : score;
    return (static_cast<uint64_t>(ord_desc(s)) << 32) | static_cast<uint64_t>(0xffffffffu - row);
}
__device__ __forceinline__ float key_score(uint64_t key, bool ascending) {
    uint32_t hi = static_cast<uint32_t>(key >> 32);
    if (hi == 0u) return __uint_as_float(0x7fc00000u);
    uint32_t u = (hi & 0x80000000u) ? (hi ^ 0x80000000u) : ~hi;
    float s = __uint_as_float(u);
    return ascending ? __fadd_rn(-s, 0.0f) : s;
}
__device__ __forceinline__ uint32_t key_row(uint64_t key) { return 0xffffffffu - static_cast<uint32_t>(key); }

// Warp-wide maximum of 64-bit keys with two REDUX ops (hi word, then lo word among the lanes holding the max hi).
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
    const uint32_t hi = static_cast<uint32_t>(v >> 32), lo = static_cast<uint32_t>(v);
    const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    return (static_cast<uint64_t>(mh) << 32) | ml;
}

// One warp pops the best `k` keys out of up to 32 DESCENDING-sorted lists held in shared memory (lane i owns
// list i, `stride` keys apart, `len` keys long, 0 = end of list).  Keys are unique, so exactly one lane pops per
// round.  Writes out[0..k) (0-padded) from lane 0.  ~60 cycles per round instead of a block-wide bitonic sort.
__device__ __forceinline__ void warp_tournament_topk(const uint64_t* lists, uint32_t n_lists, uint32_t stride, uint32_t len,
                                                     uint32_t k, uint64_t* out, uint32_t lane) {
    const uint64_t* mine = lists + (size_t)lane * stride;
    uint32_t ptr = 0;
    uint64_t head = (lane < n_lists && len > 0) ? mine[0] : 0ull;
    for (uint32_t r = 0; r < k; ++r) {
        const uint64_t m = warp_max_u64(head);
        if (lane == 0) out[r] = m;
        if (m != 0ull && head == m) {
            ++ptr;
            head = ptr < len ? mine[ptr] : 0ull;
        }
    }
}

// Bitonic sort (descending) of n = 2^m keys in shared memory by `nthreads` threads that all call this
// with the same arguments; `bar_id`/`nthreads` name the barrier they share.
__device__ __forceinline__ void bitonic_sort_desc(uint64_t* s, uint32_t n, uint32_t tid, uint32_t nthreads,
                                                  uint32_t bar_id) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = tid; i < (n >> 1); i += nthreads) {
                uint32_t lo = 2 * i - (i & (stride - 1));
                uint32_t hi = lo + stride;
                bool desc = (lo & size) == 0;
                uint64_t a = s[lo], b = s[hi];
                if ((a < b) == desc) {
                    s[lo] = b;
                    s[hi] = a;
                }
            }
            named_bar_sync(bar_id, nthreads);
        }
    }
}

}  // namespace cgv
