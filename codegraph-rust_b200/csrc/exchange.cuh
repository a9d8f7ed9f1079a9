// exchange.cuh — the top-k exchange of a row-sharded index as ONE kernel over NVLink peer memory.
//
// After the shard scan (K1) every rank holds `n_lists` per-CTA best-k lists.  xchg_merge_kernel (one CTA per query)
//   0. reduces them to this shard's best k (warp tournament, common.cuh),
//   1. STORES those k keys straight into every peer's exchange buffer (peer-mapped memory, st.global over
//      NVLink / NVSwitch), fences system-wide and publishes a per-(source, query) sequence flag with st.release.sys,
//   2. spins (ld.acquire.sys) until all `world` sources have published this step's flag in ITS OWN buffer,
//   3. merges the `world` lists and decodes (row, score, count) — every rank ends with the global answer.
// This replaces merge -> ncclAllGather -> merge (three launches and NCCL's small-message latency) on the
// batch-1 path; NCCL stays the transport for large batches (tensor path) and the fallback when peer access is
// unavailable.  The kernel releases its programmatic dependents at once, so the next query's scan overlaps it.
//
// Buffer per rank (cudaMalloc'ed, shared through CUDA IPC): flags[world][4] u32 (first 512 bytes) followed by
// data[2 parities][world sources][4 queries][128 keys].  Steps alternate parity; a source can run at most one step
// ahead of a reader (it needs the reader's flag of step s to finish step s), so two parities never collide.
#pragma once
#include "common.cuh"

namespace cgv {

constexpr uint32_t kXchgMaxWorld = 8;
constexpr uint32_t kXchgMaxQ = 4;
constexpr uint32_t kXchgMaxK = 128;
constexpr uint32_t kXchgFlagWords = 128;                                     // 512 bytes of flags
constexpr size_t kXchgBytes = kXchgFlagWords * 4 + (size_t)2 * kXchgMaxWorld * kXchgMaxQ * kXchgMaxK * 8;
constexpr int kXchgThreads = 256;

struct XchgParams {
    const uint64_t* partials;      // [nq][n_lists][k] sorted per-CTA lists from the scan
    uint32_t n_lists, k, nq;
    int ascending;
    uint32_t rank, world, seq;
    uint8_t* peer[kXchgMaxWorld];  // exchange buffer of every rank as mapped in THIS process (peer[rank] = own)
    uint64_t* out_rows;
    float* out_scores;
    uint32_t* out_counts;
    uint64_t* trace;
    uint32_t* err;                 // host-mapped error word of the index: set to 1 + source rank when that rank's list never arrived
    uint64_t timeout_ns;           // bound on the flag spin (%globaltimer): a dead or absent rank becomes CGVEC_ERR_NCCL, not a hang
};

__device__ __forceinline__ uint32_t* xchg_flags(uint8_t* base) { return reinterpret_cast<uint32_t*>(base); }
__device__ __forceinline__ uint64_t* xchg_slot(uint8_t* base, uint32_t parity, uint32_t src, uint32_t q) {
    return reinterpret_cast<uint64_t*>(base + kXchgFlagWords * 4) + (((size_t)parity * kXchgMaxWorld + src) * kXchgMaxQ + q) * kXchgMaxK;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(kXchgThreads) xchg_merge_kernel(const XchgParams p) {
    extern __shared__ __align__(16) uint64_t s_x[];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");          // PDL chain mode 2: the scan that wrote `partials` has completed
    if (threadIdx.x == 0) trace_begin(p.trace);
    const uint32_t q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t parity = p.seq & 1u;
    // ---- 0. this shard's best k of query q
    const uint32_t total = p.n_lists * p.k;
    uint64_t* staged = s_x;                                     // [n_lists * k]
    uint64_t* lvl = s_x + total;                                // [8][k]
    uint64_t* best = lvl + 8 * p.k;                             // [k]
    const uint64_t* src = p.partials + (size_t)q * total;
    for (uint32_t i = tid; i < total; i += kXchgThreads) staged[i] = src[i];
    __syncthreads();
    const uint32_t nw = (p.n_lists + 31) / 32;
    if (warp < nw) warp_tournament_topk(staged + (size_t)warp * 32 * p.k, min(32u, p.n_lists - warp * 32), p.k, p.k, p.k, lvl + (size_t)warp * p.k, lane);
    __syncthreads();
    if (warp == 0) warp_tournament_topk(lvl, nw, p.k, p.k, p.k, best, lane);
    __syncthreads();
    // ---- 1. push to every rank (own buffer included), then publish
    for (uint32_t i = tid; i < p.world * p.k; i += kXchgThreads) {
        const uint32_t r = i / p.k, j = i - r * p.k;
        xchg_slot(p.peer[r], parity, p.rank, q)[j] = best[j];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < p.world) st_release_sys(&xchg_flags(p.peer[tid])[p.rank * kXchgMaxQ + q], p.seq);
    // ---- 2. wait for every source's list of this step
    if (tid < p.world) {
        const uint32_t* f = &xchg_flags(p.peer[p.rank])[tid * kXchgMaxQ + q];
        const uint64_t t0 = global_ns();
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(f) - p.seq) < 0) {
            if ((++spins & 1023u) == 0 && global_ns() - t0 > p.timeout_ns) {
                if (p.err) atomicExch_system(p.err, 1u + tid);
                break;
            }
        }
    }
    __syncthreads();
    // ---- 3. global merge (world <= 8 lists, one warp) + decode.  The lists were written by other GPUs: read them past
    //         L1 (ld.global.cg) into shared memory first.
    for (uint32_t i = tid; i < p.world * p.k; i += kXchgThreads) {
        const uint32_t r = i / p.k, j = i - r * p.k;
        staged[i] = __ldcg(xchg_slot(p.peer[p.rank], parity, r, q) + j);
    }
    __syncthreads();
    if (warp == 0) warp_tournament_topk(staged, p.world, p.k, p.k, p.k, best, lane);
    __syncthreads();
    uint32_t cnt = 0;
    for (uint32_t i = tid; i < p.k; i += kXchgThreads) {
        const uint64_t key = best[i];
        const bool valid = key != 0ull;
        if (p.out_rows) p.out_rows[(size_t)q * p.k + i] = valid ? (uint64_t)key_row(key) : ~0ull;
        if (p.out_scores) p.out_scores[(size_t)q * p.k + i] = valid ? key_score(key, p.ascending != 0) : 0.0f;
        cnt += valid;
    }
    if (p.out_counts) {
        __shared__ uint32_t s_cnt;
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        if (cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        if (tid == 0) p.out_counts[q] = s_cnt;
    }
    if (tid == 0) trace_end(p.trace);
}

}  // namespace cgv
