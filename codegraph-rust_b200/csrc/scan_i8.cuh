// scan_i8.cuh — int8 quantised scan: the GPU twin of OptimizationResult::search_optimized
// (reference crates/codegraph-vector/src/optimization.rs:63-150) over codes produced like quantize_batch
// (:212-224 quantize_unit_range_symmetric, :268-274 "+128" offset to u8).
//
// Integer arithmetic is associative, so unlike the f32 kernels no operation order has to be reproduced: the dot
// sum((code-128)*q) = dp4a(code_u8, q_s8) - 128*sum(q) and the row norm sum((code-128)^2) are exact in int32; the
// score (dot as f32) / (norm_query * sqrt(norm_v as f32)) is finished with IEEE f32 ops exactly as the reference
// writes it, so scores are bit-identical.  1 byte per element: 4x fewer HBM bytes than the f32 scan.
// This file holds the quantisers; the scan itself is scan_exact_kernel<uint32_t, METRIC_I8, 1> (scan_exact.cuh): the same
// persistent TMA-staged ring as the f32 / f16 scan with dp4a consumers.
#pragma once
#include "common.cuh"

namespace cgv {

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

}  // namespace cgv
