// scan_i8.cuh — int8 quantised scan: the GPU twin of OptimizationResult::search_optimized
// (reference crates/codegraph-vector/src/optimization.rs:63-150) over codes produced like quantize_batch
// (:212-224 quantize_unit_range_symmetric, :268-274 "+128" offset to u8).
//
// Integer arithmetic is associative, so unlike the f32 kernels no operation order has to be reproduced: the dot
// sum((code-128)*q) = dp4a(code_u8, q_s8) - 128*sum(q) and the row norm sum((code-128)^2) are exact in int32; the
// score (dot as f32) / (norm_query * sqrt(norm_v as f32)) is finished with IEEE f32 ops exactly as the reference
// writes it, so scores are bit-identical.  1 byte per element: 4x fewer HBM bytes than the f32 scan.
#pragma once
#include "common.cuh"

namespace cgv {

constexpr int kI8Threads = 256;
constexpr int kI8RowsPerWarp = 4;

struct I8Params {
    const uint8_t* codes;     // [n][ld8] u8 codes (value + 128), rows padded with 128 (= 0) to ld8 % 16 == 0
    const int32_t* norms;     // [n] sum((code-128)^2)
    const int8_t* q;          // [ld8] quantised query, zero padded
    const float* q_norm;      // [1] sqrt(sum(q^2)) as the reference accumulates it (f32, sequential)
    const int32_t* q_sum;     // [1] sum(q)
    uint64_t* partials;       // [grid][k]
    uint64_t n_rows;
    uint32_t ld8, k, cand_cap, sync_interval;
};

__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}

// optimization.rs:212-224 (bits = 8): clamp to [-1,1], * 127, round half away from zero, `as i32` (NaN -> 0), clamp.
__device__ __forceinline__ int quantize_unit_i8(float v) {
    if (v != v) return 0;
    float c = fminf(fmaxf(v, -1.0f), 1.0f);
    int q = (int)roundf(__fmul_rn(c, 127.0f));
    return max(-127, min(127, q));
}

template <typename T>
__global__ void quantize_rows_i8_kernel(const T* __restrict__ rows, uint64_t n, uint32_t d, uint32_t ld, uint32_t ld8,
                                        uint8_t* __restrict__ codes, int32_t* __restrict__ norms) {
    const uint64_t row = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (row >= n) return;
    const T* v = rows + row * ld;
    uint8_t* c = codes + row * ld8;
    int nv = 0;
    for (uint32_t i = lane; i < ld8; i += 32) {
        int q = i < d ? quantize_unit_i8(ldf(v + i)) : 0;
        c[i] = (uint8_t)(q + 128);
        nv += q * q;
    }
    nv = __reduce_add_sync(0xffffffffu, nv);
    if (lane == 0) norms[row] = nv;
}

// query side of search_optimized (:95-115): q_i8, norm_query (sequential f32 sum of squares, then sqrt), sum(q)
__global__ void quantize_query_i8_kernel(const float* __restrict__ q, uint32_t d, uint32_t ld8, int8_t* __restrict__ out,
                                         float* __restrict__ q_norm, int32_t* __restrict__ q_sum) {
    for (uint32_t i = threadIdx.x; i < ld8; i += blockDim.x) out[i] = (int8_t)(i < d ? quantize_unit_i8(q[i]) : 0);
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        int t = 0;
        for (uint32_t i = 0; i < d; ++i) { float x = (float)out[i]; s = add_rn(s, mul_rn(x, x)); t += out[i]; }
        *q_norm = sqrt_rn(s);
        *q_sum = t;
    }
}

__global__ void __launch_bounds__(kI8Threads) scan_i8_kernel(const I8Params p) {
    extern __shared__ __align__(16) uint8_t smem_i8[];
    uint4* s_q = reinterpret_cast<uint4*>(smem_i8);                                   // ld8 bytes
    uint64_t* s_cand = reinterpret_cast<uint64_t*>(smem_i8 + ((p.ld8 + 15) & ~15u));
    __shared__ uint64_t s_thr;
    __shared__ uint32_t s_count;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t chunks = p.ld8 >> 4;
    for (uint32_t i = tid; i < chunks; i += kI8Threads) s_q[i] = reinterpret_cast<const uint4*>(p.q)[i];
    if (tid == 0) { s_thr = 0; s_count = 0; }
    __syncthreads();
    const float nq = *p.q_norm;
    const int qsum = *p.q_sum;
    const uint64_t rows_per_iter = (uint64_t)(kI8Threads / 32) * kI8RowsPerWarp;
    const uint64_t iters = (p.n_rows + rows_per_iter * gridDim.x - 1) / (rows_per_iter * gridDim.x);
    const uint32_t flush_limit = p.cand_cap - p.sync_interval * (uint32_t)rows_per_iter;

    auto compact = [&](bool force) {
        __syncthreads();
        const uint32_t cnt = s_count;
        __syncthreads();
        if (force || cnt > flush_limit) {
            for (uint32_t i = cnt + tid; i < p.cand_cap; i += kI8Threads) s_cand[i] = 0;
            __syncthreads();
            bitonic_sort_desc(s_cand, p.cand_cap, tid, kI8Threads, 0);
            if (tid == 0) {
                const uint32_t keep = min(cnt, p.k);
                s_count = keep;
                s_thr = keep >= p.k ? s_cand[p.k - 1] : 0ull;
            }
            __syncthreads();
        }
    };

    for (uint64_t it = 0; it < iters; ++it) {
        if (it && it % p.sync_interval == 0) compact(false);
        const uint64_t row0 = (it * gridDim.x + blockIdx.x) * rows_per_iter + (uint64_t)warp * kI8RowsPerWarp;
        int acc[kI8RowsPerWarp];
#pragma unroll
        for (int r = 0; r < kI8RowsPerWarp; ++r) {
            acc[r] = 0;
            const uint64_t row = row0 + r;
            if (row < p.n_rows) {
                const uint4* c = reinterpret_cast<const uint4*>(p.codes + row * p.ld8);
                for (uint32_t j = lane; j < chunks; j += 32) {
                    const uint4 a = __ldg(c + j), b = s_q[j];
                    acc[r] = dp4a_us(a.x, b.x, acc[r]);
                    acc[r] = dp4a_us(a.y, b.y, acc[r]);
                    acc[r] = dp4a_us(a.z, b.z, acc[r]);
                    acc[r] = dp4a_us(a.w, b.w, acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kI8RowsPerWarp; ++r) {
            const int s1 = __reduce_add_sync(0xffffffffu, acc[r]);
            const uint64_t row = row0 + r;
            if (lane == 0 && row < p.n_rows) {
                const int nv = p.norms[row];
                if (nv != 0) {                                                      // optimization.rs:132-134
                    const int dot = s1 - 128 * qsum;
                    const float score = div_rn((float)dot, mul_rn(nq, sqrt_rn((float)nv)));   // :136
                    const uint64_t key = make_key(score, (uint32_t)row, false);
                    if (key > *reinterpret_cast<volatile uint64_t*>(&s_thr)) {
                        const uint32_t pos = atomicAdd(&s_count, 1u);
                        s_cand[pos] = key;
                    }
                }
            }
        }
    }
    compact(true);
    const uint32_t cnt = s_count;
    uint64_t* out = p.partials + (size_t)blockIdx.x * p.k;
    for (uint32_t i = tid; i < p.k; i += kI8Threads) out[i] = i < cnt ? s_cand[i] : 0ull;
}

}  // namespace cgv
