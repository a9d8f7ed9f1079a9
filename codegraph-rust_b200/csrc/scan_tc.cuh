// scan_tc.cuh — K2: batched-query scan on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), fp16 rows.
//
// The reference has no batched kernel: its "batch" shapes are a per-query loop
// (batch_cosine_similarity_avx2, crates/codegraph-vector/src/simd_ops.rs:85-99) and concurrent
// search_similar calls (multi_vector_search, src/search.rs:347-361).  For nq >= ~8 queries the scan is a
// genuine dense contraction S[row, q] = sum_k R[row,k] * Q[q,k], so it goes to tcgen05.mma.
//
// Shape.  One persistent CTA per SM.  The query block B (N = nq padded to 16, K = d padded to 64) is TMA-loaded
// ONCE and stays resident in shared memory as SWIZZLE_128B K-major tiles; the row stream A (128 rows x 64
// halves = 16 KB per stage) flows through a ring of TMA stages.  Warp 4 = TMA producer, warp 5 = MMA issuer
// (one elected lane issues 4 x tcgen05.mma.cta_group::1.kind::f16 M=128,N,K=16 per stage, then
// tcgen05.commit frees the stage), warp 6 owns the TMEM allocation, warps 0-3 are the epilogue: they pull
// the 128 x N fp32 accumulator tile out of TMEM (tcgen05.ld 32x32b), scale by 1/|row|, compare against the
// per-query threshold and append the rare survivors (key = ord(score) << 32 | ~row) to per-query candidate
// lists in global memory with warp-aggregated atomics.  Two TMEM accumulator buffers let the epilogue of
// tile t overlap the MMAs of tile t+1.  The Q x N score matrix is never written.
//
// Exactness.  Tensor-core scores are only an ORDERING HEURISTIC here (fp16 x fp16 products are exact in fp32,
// accumulation order differs from the oracle, queries are rounded to fp16).  The host side (cgvec_api.cu,
// search_tensor) scans the shard in a few geometrically growing row ranges, tightening each query's
// threshold to its current kp-th best between ranges (tc_select_kernel), then re-scores the kp survivors per
// query in the reference's exact order (K4), re-ranks, and PROVES with a rigorous error bound that no
// discarded row can reach the top-k; an unproven query is re-run on the exact-order kernel (scan_exact.cuh).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "scan_exact.cuh"

namespace cgv {

constexpr int kTcThreads = 384;               // warps 0-3 and 8-11: epilogue (two per TMEM lane quarter); 4: TMA; 5: MMA; 6: TMEM alloc; 7: spare
constexpr int kTcTileRows = 128;
constexpr int kTcKBlock = 64;                 // halves per 128-byte swizzle row
constexpr int kTcKBlockBytes = kTcTileRows * 128;   // one K-block of a 128-row tile: 16 KB
constexpr int kTcMaxN = 128;                // tc_scan_kernel (resident query block)
constexpr int kTc2MaxN = 256;               // tc2_scan_kernel (CTA pairs, streamed query block)
constexpr int kTcMaxStages = 12;

struct TcParams {
    uint64_t n_rows;            // local rows of the shard (rows >= n_rows are TMA zero fill and masked)
    uint64_t row_begin, row_end;   // this launch scans local rows [row_begin, row_end); row_begin % 256 == 0
    const float* norms;         // squared row norms (reference order); cosine only
    float* thr;                 // [N] per-query thresholds in dot/|row| units (tightened in place by the selector warps)
    uint64_t* cand;             // [N][cap] candidate keys: [0, kp) = the query's current best (sorted), [kp, count) = appended survivors
    uint32_t* cand_count;       // [N]
    uint32_t* consumed;         // [N] appended entries below this index are already folded into the best list
    uint32_t* overflow;         // set when a list would exceed cap
    uint32_t cap;
    uint32_t kp;                // survivors kept per query (k + margin)
    uint32_t sel_on;            // 1: warps 6-7 of every CTA refresh thresholds while the scan runs (main range); 0: thresholds fixed
    uint32_t boot;              // 1: bootstrap range: EVERY score is kept, stored at cand[q][row - row_begin] (no thresholds, no atomics)
    uint32_t sel_cap;           // keys in the selector's sort buffer (power of two, >= 2 * kp)
    uint32_t nq;                // real queries (<= N)
    uint32_t N;                 // MMA N: nq rounded up to 16
    uint32_t nkb;               // 128-byte K-blocks per row: ceil(d / 64) halves or ceil(d / 32) floats
    uint32_t kbs;               // K-blocks per ring stage: one TMA box, one tcgen05.commit
    uint32_t groups;            // ceil(nkb / kbs) stages per row tile
    uint32_t box4d;             // 1: 4-D boxes {128 B, 8 rows, kbs K-blocks, row groups} (row pitch % 128 == 0); 0: 2-D boxes, kbs == 1
    uint32_t stages;
    uint32_t metric;            // METRIC_COSINE or METRIC_DOT
    uint32_t tmem_cols;         // power of two >= 2*N, >= 32
    uint32_t tf32;              // 0: f16 rows/queries (kind::f16, 64 elements per 128-byte K-block); 1: f32 rows/queries as TF32 (kind::tf32, 32)
    uint32_t debug;             // bit 0: skip the MMAs, bit 1: skip the epilogue body, bit 2: K2b streams the query block for the first tile only (triage only; results invalid)
    uint64_t row_offset;
    uint32_t blk_rows, n_shards, shard_id;
};

// Shared-memory plan of tc_scan_kernel: resident query block | ring of row stages | barriers | misc | thresholds
struct TcSmemLayout { uint32_t off_b, off_a, stage_bytes, off_bars, off_misc, off_thr, off_sel, total; };
__host__ __device__ inline TcSmemLayout tc_smem_layout(uint32_t N, uint32_t nkb, uint32_t stages, uint32_t kbs, uint32_t sel_cap) {
    TcSmemLayout L;
    L.off_b = 0;
    L.off_a = nkb * N * 128;                                   // multiple of 1024 because N % 8 == 0
    L.stage_bytes = kbs * kTcKBlockBytes;
    L.off_bars = L.off_a + stages * L.stage_bytes;
    L.off_misc = L.off_bars + (2 * stages + 1 + 4) * 8;        // tmem base address
    L.off_thr = (L.off_misc + 16 + 15) & ~15u;                 // negated thresholds, 16-byte aligned for LDS.128
    L.off_sel = L.off_thr + N * 4;                             // selector sort buffer
    L.total = L.off_sel + 2 * sel_cap * 8;                     // one buffer per selector warp
    return L;
}

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(pol)
                 : "memory");
}
// 4-D box {128 bytes of K, 8 rows, kbs K-blocks, row groups}: one instruction lands kbs K-blocks of a row tile as kbs
// consecutive SWIZZLE_128B atoms per 8-row group, i.e. K-major operand tiles with a stride of kbs*1024 bytes between
// 8-row groups (the SBO of the UMMA descriptor).
__device__ __forceinline__ void tma_load_4d_hint(void* dst, const CUtensorMap* map, int c2, int c3, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %3, %4, %5}], [%2], %6;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(0), "r"(c2), "r"(c3), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// One lane of the (converged) warp: the single-thread roles run their loops warp-wide and issue under this predicate, so
// that ring positions, descriptors and barrier addresses stay in uniform registers.  (Issuing from `if (lane == 0)` makes
// every operand thread-divergent for the compiler: it then wraps each tcgen05.mma in an R2UR waterfall loop, ~16
// instructions per MMA on the kernel's critical thread.)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t uniform_warp_id() { return __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start>>4 [0,14),
// LBO>>4 [16,30) (ignored for swizzled K-major, 1), SBO>>4 [32,46) = bytes between 8-row groups, version 1
// [46,48), layout type SWIZZLE_128B = 2 at [61,64).  The MMA issuer keeps the two halves apart: the high word is a
// per-launch constant, the low word (start address) advances by 2 per 32 bytes of K and by 64 per K-block.
__device__ __forceinline__ uint32_t umma_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14) | (2u << 29); }
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3fffu) | (1u << 16); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) { return umma_desc(umma_desc_lo(smem_addr), umma_desc_hi(1024)); }
// kind::f16 instruction descriptor: D=F32 (bits 4-5 = 1), A=B=F16 (0), both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ inline uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
    return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// kind::tf32: A = B = TF32 (format 2 at bits [7,10) and [10,13)); f32 operands in shared memory, low mantissa bits ignored
__host__ __device__ inline uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T.  TF32 / PAIR are compile-time: the issuing thread's loop carries no branches.
template <bool TF32, bool PAIR>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (TF32 && PAIR)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else if (TF32)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else if (PAIR)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {     // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
template <bool PAIR>
__device__ __forceinline__ void umma_commit_t(uint64_t* bar) { if (PAIR) umma_commit_2sm(bar); else umma_commit(bar); }

__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// tcgen05.wait::ld with the loaded registers as in/out operands: every later use of r[] is data-dependent on the wait,
// so the compiler cannot schedule arithmetic on them ahead of it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// ---- epilogue of one 128-row accumulator tile (shared by both kernels) ----------------------------------------------
// Each thread owns one row (TMEM lane) and walks its share of the N query columns 16 at a time through two
// register buffers that are indexed statically (a dynamically indexed buffer would live in local memory: the
// first version of this epilogue spent a quarter of the kernel in STL/LDL round trips).  Fast path per column: one FFMA
// (acc * 1/|row| - thr, thresholds pre-negated in shared memory, LDS.128) and a share of a 3-input AND of sign bits; a
// single ballot per 16 columns decides whether ANY lane has a survivor.  Slow path (rare once thresholds have
// tightened): the surviving lane reserves a list slot with one atomicAdd and PARKS the key; the store that needs the
// atomic's result is issued at the lane's next survivor or at the end of the tile, so no warp ever waits out the
// atomic's round trip to L2 (that wait, ~1 us per slow path, was what kept the C3 epilogue from hiding behind the MMAs).
struct TcPending { uint32_t q, pos; uint64_t key; };
constexpr uint32_t kTcNoPending = 0xffffffffu;
__device__ __forceinline__ void tc_flush_pending(const TcParams& p, TcPending& pd) {
    if (pd.q != kTcNoPending) {
        if (pd.pos < p.cap) p.cand[(size_t)pd.q * p.cap + pd.pos] = pd.key;
        else *p.overflow = 1u;
        pd.q = kTcNoPending;
    }
}
__device__ __forceinline__ void tc_epilogue_chunk(const TcParams& p, const uint32_t (&r)[16], const float* __restrict__ s_nthr, uint32_t c0,
                                                  uint32_t grow, uint32_t boot_slot, bool valid, float inv, TcPending& pd) {
    float t[16];
    uint32_t all_neg = 0x80000000u;                                   // sign bit survives iff every t[j] is negative
#pragma unroll
    for (uint32_t j4 = 0; j4 < 4; ++j4) {
        const float4 nt = *reinterpret_cast<const float4*>(s_nthr + c0 + 4 * j4);
        t[4 * j4 + 0] = __fmaf_rn(__uint_as_float(r[4 * j4 + 0]), inv, nt.x);
        t[4 * j4 + 1] = __fmaf_rn(__uint_as_float(r[4 * j4 + 1]), inv, nt.y);
        t[4 * j4 + 2] = __fmaf_rn(__uint_as_float(r[4 * j4 + 2]), inv, nt.z);
        t[4 * j4 + 3] = __fmaf_rn(__uint_as_float(r[4 * j4 + 3]), inv, nt.w);
        all_neg &= __float_as_uint(t[4 * j4 + 0]) & __float_as_uint(t[4 * j4 + 1]);
        all_neg &= __float_as_uint(t[4 * j4 + 2]) & __float_as_uint(t[4 * j4 + 3]);
    }
    if (p.boot) {                                                     // bootstrap range: slot = row, one coalesced store per column
#pragma unroll
        for (uint32_t j = 0; j < 16; ++j)
            if (boot_slot < p.cap && c0 + j < p.nq)
                p.cand[(size_t)(c0 + j) * p.cap + boot_slot] = valid ? make_key(__uint_as_float(r[j]) * inv, grow, false) : 0ull;
        return;
    }
    const bool maybe = valid && !(all_neg & 0x80000000u);
    if (__ballot_sync(0xffffffffu, maybe) == 0u) return;
    if (maybe) {
#pragma unroll
        for (uint32_t j = 0; j < 16; ++j) {
            if (t[j] >= 0.0f) {
                tc_flush_pending(p, pd);
                pd.q = c0 + j;
                pd.key = make_key(__uint_as_float(r[j]) * inv, grow, false);
                pd.pos = atomicAdd(&p.cand_count[c0 + j], 1u);
            }
        }
    }
    __syncwarp();
}

// `pd` is the lane's parked survivor: it lives across tiles and is flushed by the caller at the start of the NEXT tile's
// epilogue (by then the atomic that reserved its slot has long returned) and once after the last tile.
__device__ __forceinline__ void tc_epilogue_tile(const TcParams& p, uint32_t taddr, const float* __restrict__ s_nthr, uint64_t row,
                                                 bool valid, float inv, uint32_t col_begin, uint32_t col_end, TcPending& pd) {
    if ((p.debug & 2u) || col_begin >= col_end) return;
    const uint32_t row32 = (uint32_t)row;                                 // shards hold < 2^32 rows: 32-bit division
    const uint32_t b = row32 / p.blk_rows, rr = row32 - b * p.blk_rows;
    const uint32_t grow = (uint32_t)(((uint64_t)b * p.n_shards + p.shard_id) * p.blk_rows + rr + p.row_offset);   // global row of this lane
    const uint32_t boot_slot = (uint32_t)(row - p.row_begin);
    uint32_t ra[16], rb[16];
    tmem_ld_x16(taddr + col_begin, ra);
    for (uint32_t c0 = col_begin; c0 < col_end; c0 += 32) {
        tmem_ld_wait(ra);
        const bool more = c0 + 16 < col_end;
        if (more) tmem_ld_x16(taddr + c0 + 16, rb);                       // next 16 columns stream in behind the compares
        tc_epilogue_chunk(p, ra, s_nthr, c0, grow, boot_slot, valid, inv, pd);
        if (more) {
            tmem_ld_wait(rb);
            if (c0 + 32 < col_end) tmem_ld_x16(taddr + c0 + 32, ra);
            tc_epilogue_chunk(p, rb, s_nthr, c0 + 16, grow, boot_slot, valid, inv, pd);
        }
    }
}

// Negated thresholds of this launch -> shared memory (padding columns get -inf so they can never survive).
__device__ __forceinline__ void tc_stage_thresholds(const TcParams& p, float* s_nthr, uint32_t tid, uint32_t nthreads) {
    for (uint32_t j = tid; j < p.N; j += nthreads) s_nthr[j] = j < p.nq ? -p.thr[j] : __int_as_float(0xff800000);
}

// ---- selector (warp 7 of every CTA, main range only) ------------------------------------------------------------------
// Tightens the per-query thresholds WHILE the scan runs, so one launch covers the whole shard (the first version
// relaunched the scan six times with a sort kernel in between).  Query q belongs to CTA q mod gridDim.  Its list is
// cand[q]: [0, kp) the best kp approximate keys seen so far (sorted, owned by the selector), [kp, count) survivors
// appended by every CTA's epilogue.  A round of the selector folds the newly appended keys [consumed, ...) into the best
// list with a warp-wide bitonic sort in shared memory and publishes the kp-th best as the new threshold; all selectors
// copy the thresholds of ALL queries into their CTA's shared memory once per round.  A slot whose key is still 0 has been
// reserved (atomicAdd) but not written yet: the round stops in front of it.  Any threshold ever published is the kp-th
// best of a subset of the rows seen, hence <= the final kp-th best: the true top-kp by approximate score always survive,
// which is all the exact re-score + proof that follow need.
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* ptr) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ float ld_volatile_f32(const float* ptr) {
    float v;
    asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_f32(float* ptr, float v) { asm volatile("st.volatile.global.f32 [%0], %1;" ::"l"(ptr), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_volatile_u32(const uint32_t* ptr) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(ptr)) : "memory");
    return v;
}
__device__ __forceinline__ void warp_bitonic_sort_desc(uint64_t* s, uint32_t n, uint32_t lane) {
    for (uint32_t size = 2; size <= n; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t i = lane; i < (n >> 1); i += 32) {
                const uint32_t lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const uint64_t a = s[lo], b = s[hi];
                if ((a < b) == desc) { s[lo] = b; s[hi] = a; }
            }
            __syncwarp();
        }
    }
}
constexpr uint32_t kTcEpilogueWarps = 8;
constexpr uint32_t kTcSelStep = 256;              // appended keys examined per step of a selector round (8 loads in flight per lane)
__device__ __noinline__ void tc_selector_loop(const TcParams& p, uint64_t* s_sort, const uint32_t* s_epi_done, uint32_t sel, uint32_t nsel,
                                              uint32_t lane) {
    const uint32_t kp = p.kp, new_max = p.sel_cap - kp;
    const uint32_t lt_mask = (1u << lane) - 1u;
    while (lds_volatile_u32(s_epi_done) < kTcEpilogueWarps) {
        for (uint32_t q = sel; q < p.nq; q += nsel) {
            const uint32_t cons = __ldcg(p.consumed + q);
            uint32_t cnt = ld_volatile_u32(p.cand_count + q);
            if (cnt > p.cap) cnt = p.cap;
            if (cnt <= cons) continue;
            uint64_t* c = p.cand + (size_t)q * p.cap;
            // Appended keys were admitted under whatever threshold their CTA had at the time; only those above the CURRENT
            // kp-th best can change the best list.  Filtering is one compare per key, so a backlog built up under a loose
            // early threshold is worked off at L2-read speed and only the few keys that matter reach the sort.
            const uint64_t kth_cur = __ldcg(c + kp - 1);
            uint32_t pos = cons, nb = 0;
            bool blocked = false;                                         // ran into a reserved-but-unwritten slot
            while (!blocked && pos < cnt && nb + kTcSelStep <= new_max) {
                uint64_t key[kTcSelStep / 32];
#pragma unroll
                for (uint32_t u = 0; u < kTcSelStep / 32; ++u) {
                    const uint32_t i = pos + u * 32 + lane;
                    key[u] = i < cnt ? __ldcg(c + i) : 0ull;
                }
                uint32_t examined = 0;
#pragma unroll
                for (uint32_t u = 0; u < kTcSelStep / 32; ++u) {
                    if (blocked) break;
                    const uint32_t i = pos + u * 32 + lane;
                    const bool inr = i < cnt;
                    const uint32_t unwritten = __ballot_sync(0xffffffffu, inr && key[u] == 0ull);
                    uint32_t span = min(32u, cnt > pos + u * 32 ? cnt - (pos + u * 32) : 0u);
                    if (unwritten) { span = __ffs(unwritten) - 1; blocked = true; }
                    const bool pass = lane < span && key[u] > kth_cur;
                    const uint32_t mask = __ballot_sync(0xffffffffu, pass);
                    if (pass) s_sort[kp + nb + __popc(mask & lt_mask)] = key[u];
                    nb += __popc(mask);
                    examined += span;
                }
                pos += examined;
            }
            if (pos == cons) continue;
            if (nb) {
                for (uint32_t i = lane; i < kp; i += 32) s_sort[i] = __ldcg(c + i);
                uint32_t n = 2;
                while (n < kp + nb) n <<= 1;
                for (uint32_t i = kp + nb + lane; i < n; i += 32) s_sort[i] = 0ull;
                __syncwarp();
                warp_bitonic_sort_desc(s_sort, n, lane);
                for (uint32_t i = lane; i < kp; i += 32) __stcg(c + i, s_sort[i]);
            }
            if (lane == 0) {
                __stcg(p.consumed + q, pos);
                if (nb) {
                    const uint64_t kth = s_sort[kp - 1];
                    if (kth != 0ull) st_volatile_f32(p.thr + q, key_score(kth, false));
                }
            }
            __syncwarp();
            if (lds_volatile_u32(s_epi_done) >= kTcEpilogueWarps) return;
        }
    }
}
// Epilogue warps re-read their columns' thresholds from global memory once per tile (behind the accumulator wait), so a
// threshold published by any selector takes effect everywhere within one tile time.
// The loads are ISSUED before the warp waits for its accumulator and CONSUMED (negated into shared memory) after it has handed
// the accumulator back, so their L2 round trip (the single most-sampled stall of the first version of this kernel) hides
// behind the tile instead of sitting in front of it; the thresholds a tile is filtered with are one tile older.
constexpr uint32_t kTcThrRegs = kTc2MaxN / 2 / 32;                     // columns per epilogue warp / lanes
__device__ __forceinline__ void tc_thresholds_issue(const TcParams& p, float (&pre)[kTcThrRegs], uint32_t col_begin, uint32_t col_end, uint32_t lane) {
#pragma unroll
    for (uint32_t u = 0; u < kTcThrRegs; ++u) {
        const uint32_t j = col_begin + lane + 32 * u;
        pre[u] = (j < col_end && j < p.nq) ? ld_volatile_f32(p.thr + j) : 0.0f;
    }
}
__device__ __forceinline__ void tc_thresholds_commit(const TcParams& p, const float (&pre)[kTcThrRegs], float* s_nthr, uint32_t col_begin, uint32_t col_end, uint32_t lane) {
#pragma unroll
    for (uint32_t u = 0; u < kTcThrRegs; ++u) {
        const uint32_t j = col_begin + lane + 32 * u;
        if (j < col_end && j < p.nq) s_nthr[j] = -pre[u];
    }
    __syncwarp();
}

// ---- MMA issue loop (one elected thread) ------------------------------------------------------------------------------
// Per ring stage: wait for the bytes, 4 MMAs (K = 16 halves / 8 floats each) per K-block of the stage, ONE commit.  The
// thread's instruction stream is the bottleneck of the whole kernel when it is not lean (the first version rebuilt both
// 64-bit descriptors and branched on the operand kind per MMA and issued one commit per K-block: the tensor pipe sat at
// 50 %), so: operand kind and pairing are template parameters, descriptor high words are loop constants, low words are
// advanced with adds, and the ring position is a counter (no division).
//   a_lo0            descriptor low word of stage 0's A tile; stage s is (s * stage_bytes) >> 4 further
//   b_resident_lo0   resident query block (tc_scan_kernel): low word of K-block 0, K-block kb is (kb * N * 128) >> 4 further
//   b_in_stage_off   streamed query block (tc2_scan_kernel): byte offset of the B part inside a stage
template <bool TF32, bool PAIR>
__device__ __forceinline__ void tc_issue_loop(const TcParams& p, uint64_t my_tiles, uint32_t tmem_base, uint32_t a_lo0, uint32_t stage_bytes,
                                              uint32_t b_resident_lo0, uint32_t b_in_stage_off, uint64_t* full_bar, uint64_t* empty_bar,
                                              uint64_t* tfull_bar, uint64_t* tempty_bar) {
    const uint32_t M = PAIR ? 2 * kTcTileRows : kTcTileRows;
    const uint32_t idesc = TF32 ? umma_idesc_tf32(M, p.N) : umma_idesc_f16(M, p.N);
    const uint32_t hi_stage = umma_desc_hi(p.kbs * 1024), hi_res = umma_desc_hi(1024);
    const uint32_t stage_step = stage_bytes >> 4, b_kb_step = PAIR ? 64u : (p.N * 128u) >> 4;
    const uint32_t skip_mma = p.debug & 1u;
    uint32_t s = 0, ph = 0, a_lo = a_lo0;
    for (uint64_t t = 0; t < my_tiles; ++t) {
        const uint32_t buf = t & 1;
        mbar_wait(&tempty_bar[buf], ((t >> 1) & 1) ^ 1);             // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * p.N;
        uint32_t kb = 0;
        for (uint32_t g = 0; g < p.groups; ++g) {
            mbar_wait(&full_bar[s], ph);                              // TMA bytes have landed
            tc_fence_after();
            const uint32_t nk = min(p.kbs, p.nkb - kb);
            uint32_t al = a_lo;
            uint32_t bl = PAIR ? a_lo + (b_in_stage_off >> 4) : b_resident_lo0 + kb * b_kb_step;
            const uint32_t bh = PAIR ? hi_stage : hi_res;
            if (elect_one()) {
                if (!skip_mma) {
                    for (uint32_t i = 0; i < nk; ++i, al += 64, bl += b_kb_step) {
#pragma unroll
                        for (uint32_t k = 0; k < 4; ++k)
                            umma_ss<TF32, PAIR>(d_tmem, umma_desc(al + 2 * k, hi_stage), umma_desc(bl + 2 * k, bh), idesc, (kb + i) | k);
                    }
                }
                umma_commit_t<PAIR>(&empty_bar[s]);                    // stage free once these MMAs retire
                if (g + 1 == p.groups) umma_commit_t<PAIR>(&tfull_bar[buf]);   // accumulator complete
            }
            __syncwarp();
            kb += nk;
            a_lo += stage_step;
            if (++s == p.stages) { s = 0; ph ^= 1u; a_lo = a_lo0; }
        }
    }
}

// ---- the kernel -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
tc_scan_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_tc_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment in the shared window; the launch reserves 1 KB of slack for this
    uint8_t* smem = smem_tc_raw + ((1024u - (smem_u32(smem_tc_raw) & 1023u)) & 1023u);
    const TcSmemLayout lay = tc_smem_layout(p.N, p.nkb, p.stages, p.kbs, p.sel_cap);
    uint8_t* sB = smem + lay.off_b;
    uint8_t* sA = smem + lay.off_a;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + lay.off_bars);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* b_bar = empty_bar + p.stages;
    uint64_t* tfull_bar = b_bar + 1;          // [2]
    uint64_t* tempty_bar = tfull_bar + 2;     // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + lay.off_misc);

    const uint32_t tid = threadIdx.x, warp = uniform_warp_id(), lane = tid & 31;
    const uint64_t first_tile = p.row_begin / kTcTileRows;
    const uint64_t num_tiles = (p.row_end - p.row_begin + kTcTileRows - 1) / kTcTileRows;
    const uint64_t my_tiles = (num_tiles > blockIdx.x) ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    uint32_t* s_epi_done = s_tmem + 1;
    uint64_t* s_sort = reinterpret_cast<uint64_t*>(smem + lay.off_sel);
    if (tid == 0) {
        *s_epi_done = 0;
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(b_bar, 1);
        mbar_init(&tfull_bar[0], 1); mbar_init(&tfull_bar[1], 1);
        mbar_init(&tempty_bar[0], 8); mbar_init(&tempty_bar[1], 8);
        fence_mbar_init();
    }
    // Warp roles.  The SM's arbiter favours HIGHER warp ids, so the two latency-critical single-thread roles (TMA
    // producer, MMA issuer) get warps 4 and 5 and the four epilogue warps (which mostly wait) get warps 0-3.
    if (warp == 6) tmem_alloc(s_tmem, p.tmem_cols);
    float* s_nthr = reinterpret_cast<float*>(smem + lay.off_thr);
    tc_stage_thresholds(p, s_nthr, tid, kTcThreads);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 4) {
        // ===================== TMA producer (warp-wide loop, one elected lane issues) =====================
        const uint64_t pol = l2_policy_evict_first();
        const int kbe = p.tf32 ? kTcKBlock / 2 : kTcKBlock;           // elements per 128-byte K-block
        if (elect_one()) {
            tma_prefetch_desc(&tmA);
            tma_prefetch_desc(&tmB);
            mbar_arrive_expect_tx(b_bar, p.nkb * p.N * 128);
            for (uint32_t kb = 0; kb < p.nkb; ++kb) tma_load_2d(sB + (size_t)kb * p.N * 128, &tmB, kb * kbe, 0, b_bar);
        }
        __syncwarp();
        uint32_t s = 0, ph = 1;
        for (uint64_t t = 0; t < my_tiles; ++t) {
            const int row0 = (int)((first_tile + blockIdx.x + t * gridDim.x) * kTcTileRows);
            for (uint32_t g = 0; g < p.groups; ++g) {
                mbar_wait(&empty_bar[s], ph);                            // passes at once during the first trip round the ring
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[s], lay.stage_bytes);
                    uint8_t* dst = sA + (size_t)s * lay.stage_bytes;
                    if (p.box4d) tma_load_4d_hint(dst, &tmA, (int)(g * p.kbs), row0 >> 3, &full_bar[s], pol);
                    else tma_load_2d_hint(dst, &tmA, (int)g * kbe, row0, &full_bar[s], pol);
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer (warp-wide loop, one elected lane issues) =====================
        mbar_wait(b_bar, 0);
        tc_fence_after();
        if (p.tf32) tc_issue_loop<true, false>(p, my_tiles, tmem_base, umma_desc_lo(smem_u32(sA)), lay.stage_bytes, umma_desc_lo(smem_u32(sB)), 0, full_bar, empty_bar, tfull_bar, tempty_bar);
        else tc_issue_loop<false, false>(p, my_tiles, tmem_base, umma_desc_lo(smem_u32(sA)), lay.stage_bytes, umma_desc_lo(smem_u32(sB)), 0, full_bar, empty_bar, tfull_bar, tempty_bar);
    } else if (warp < 4 || warp >= 8) {
        // ===================== epilogue: TMEM -> registers -> threshold filter -> candidate lists =====================
        // Two warps per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31); they split the query columns.
        const uint32_t q4 = warp & 3;
        const uint32_t col_split = ((p.N / 16 + 1) / 2) * 16;
        const uint32_t col_begin = warp < 4 ? 0u : col_split, col_end = warp < 4 ? col_split : p.N;
        TcPending pd;
        pd.q = kTcNoPending; pd.pos = 0; pd.key = 0;
        const bool cosine = p.metric == METRIC_COSINE;
        auto tile_row = [&](uint64_t t) { return (first_tile + blockIdx.x + t * gridDim.x) * kTcTileRows + q4 * 32 + lane; };   // local row
        auto row_ok = [&](uint64_t row) { return row < p.n_rows && row < p.row_end; };
        float nb_next = (cosine && my_tiles && row_ok(tile_row(0))) ? p.norms[tile_row(0)] : 0.0f;
        for (uint64_t t = 0; t < my_tiles; ++t) {
            const uint32_t buf = t & 1;
            const uint64_t row = tile_row(t);
            const bool valid = row_ok(row);
            const float nb = nb_next;
            if (cosine && t + 1 < my_tiles) { const uint64_t rn = tile_row(t + 1); nb_next = row_ok(rn) ? p.norms[rn] : 0.0f; }   // one tile ahead
            const float inv = cosine ? (nb > 0.0f ? rsqrtf(nb) : 0.0f) : 1.0f;    // zero-norm row -> cosine 0 (simd_ops.rs:73-74)
            float thr_pre[kTcThrRegs];
            if (p.sel_on) tc_thresholds_issue(p, thr_pre, col_begin, col_end, lane);
            mbar_wait(&tfull_bar[buf], (t >> 1) & 1);
            tc_fence_after();
            tc_flush_pending(p, pd);                                     // survivor parked by the previous tile
            tc_epilogue_tile(p, tmem_base + buf * p.N + ((q4 * 32u) << 16), s_nthr, row, valid, inv, col_begin, col_end, pd);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
            if (p.sel_on) tc_thresholds_commit(p, thr_pre, s_nthr, col_begin, col_end, lane);
        }
        tc_flush_pending(p, pd);
        if (lane == 0) atomicAdd(s_epi_done, 1u);
    } else if (p.sel_on) {                                           // warps 6 and 7: selectors
        tc_selector_loop(p, s_sort + (size_t)(warp - 6) * p.sel_cap, s_epi_done, 2 * blockIdx.x + (warp - 6), 2 * gridDim.x, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 6) tmem_dealloc(tmem_base, p.tmem_cols);
}

// =====================================================================================================================
// K2b — the same scan as tc_scan_kernel on CTA PAIRS (cta_group::2): two SMs of a TPC share one MMA of M = 256 rows.
// Each CTA stages its own 128 rows of A and HALF of the query block per ring stage (the tensor core reads B from both
// CTAs' shared memory), so a stage costs kbs x (16 KB + N/2 x 128 B) per CTA instead of keeping the whole query block
// resident: up to 256 queries in ONE pass over HBM (config C3).
// Leader CTA (cluster rank 0) issues the MMAs; both CTAs run a TMA producer whose transactions complete on the
// LEADER's full barrier (peer bit of the barrier address cleared); tcgen05.commit multicasts the "stage free" and
// "accumulator ready" arrivals to both CTAs; both CTAs' epilogue warps release the accumulator on the leader's
// barrier (remote mbarrier.arrive through mapa).
// =====================================================================================================================
struct Tc2SmemLayout { uint32_t a_bytes, stage_bytes, off_bars, off_misc, off_thr, off_sel, total; };
__host__ __device__ inline Tc2SmemLayout tc2_smem_layout(uint32_t N, uint32_t stages, uint32_t kbs, uint32_t sel_cap) {
    Tc2SmemLayout L;
    L.a_bytes = kbs * kTcKBlockBytes;
    L.stage_bytes = L.a_bytes + kbs * (N / 2) * 128;             // A tile + this CTA's half of the B K-blocks; multiple of 1024
    L.off_bars = stages * L.stage_bytes;
    L.off_misc = L.off_bars + (2 * stages + 4) * 8;
    L.off_thr = (L.off_misc + 16 + 15) & ~15u;
    L.off_sel = L.off_thr + N * 4;
    L.total = L.off_sel + 2 * sel_cap * 8;                     // one buffer per selector warp
    return L;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;                   // shared::cluster address of the same offset in CTA 0
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* leader_bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, int c2, int c3, uint64_t* leader_bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %3, %4, %5}], [%2], %6;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(0), "r"(c2), "r"(c3), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {   // arrive on the same-offset barrier of cluster CTA `cta`
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(cta)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
tc2_scan_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_tc2_raw[];
    uint8_t* smem = smem_tc2_raw + ((1024u - (smem_u32(smem_tc2_raw) & 1023u)) & 1023u);
    const Tc2SmemLayout lay = tc2_smem_layout(p.N, p.stages, p.kbs, p.sel_cap);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + lay.off_bars);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tfull_bar = empty_bar + p.stages;    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;          // [2] (leader's copy is the live one)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + lay.off_misc);

    const uint32_t tid = threadIdx.x, warp = uniform_warp_id(), lane = tid & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    constexpr uint32_t kPairRows = 2 * kTcTileRows;
    const uint64_t first_tile = p.row_begin / kPairRows;
    const uint64_t num_tiles = (p.row_end - p.row_begin + kPairRows - 1) / kPairRows;
    const uint64_t my_tiles = (num_tiles > pair) ? (num_tiles - pair + npairs - 1) / npairs : 0;
    const uint32_t half_n = p.N >> 1;

    uint32_t* s_epi_done = s_tmem + 1;
    uint64_t* s_sort = reinterpret_cast<uint64_t*>(smem + lay.off_sel);
    if (tid == 0) {
        *s_epi_done = 0;
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tfull_bar[0], 1); mbar_init(&tfull_bar[1], 1);
        mbar_init(&tempty_bar[0], 16); mbar_init(&tempty_bar[1], 16);
        fence_mbar_init();
    }
    if (warp == 6) tmem_alloc_2sm(s_tmem, p.tmem_cols);
    float* s_nthr = reinterpret_cast<float*>(smem + lay.off_thr);
    tc_stage_thresholds(p, s_nthr, tid, kTcThreads);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // peer barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    if (warp == 4) {
        // ===================== TMA producer (both CTAs; warp-wide loop, one elected lane issues) =====================
        const uint64_t pol = l2_policy_evict_first();
        uint64_t pol_keep;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        const int kbe = p.tf32 ? kTcKBlock / 2 : kTcKBlock;
        if (elect_one()) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
        __syncwarp();
        uint32_t s = 0, ph = 1;
        for (uint64_t t = 0; t < my_tiles; ++t) {
            const int row0 = (int)((first_tile + pair + t * npairs) * kPairRows + rank * kTcTileRows);
            for (uint32_t g = 0; g < p.groups; ++g) {
                mbar_wait(&empty_bar[s], ph);
                if (elect_one()) {
                    uint8_t* st = smem + (size_t)s * lay.stage_bytes;
                    const bool skip_b = (p.debug & 4u) && t > 0;                  // triage: the query block is not re-streamed (results invalid)
                    if (leader) mbar_arrive_expect_tx(&full_bar[s], skip_b ? 2 * lay.a_bytes : 2 * lay.stage_bytes);   // both CTAs' bytes land on the leader
                    if (p.box4d) {
                        tma_load_4d_2sm(st, &tmA, (int)(g * p.kbs), row0 >> 3, &full_bar[s], pol);
                        if (!skip_b) tma_load_4d_2sm(st + lay.a_bytes, &tmB, (int)(g * p.kbs), (int)((rank * half_n) >> 3), &full_bar[s], pol_keep);
                    } else {
                        tma_load_2d_2sm(st, &tmA, (int)g * kbe, row0, &full_bar[s], pol);
                        if (!skip_b) tma_load_2d_2sm(st + lay.a_bytes, &tmB, (int)g * kbe, (int)(rank * half_n), &full_bar[s], pol_keep);
                    }
                }
                __syncwarp();
                if (++s == p.stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer (leader CTA only; warp-wide loop, one elected lane issues) =====================
        if (leader) {
            if (p.tf32) tc_issue_loop<true, true>(p, my_tiles, tmem_base, umma_desc_lo(smem_u32(smem)), lay.stage_bytes, 0, lay.a_bytes, full_bar, empty_bar, tfull_bar, tempty_bar);
            else tc_issue_loop<false, true>(p, my_tiles, tmem_base, umma_desc_lo(smem_u32(smem)), lay.stage_bytes, 0, lay.a_bytes, full_bar, empty_bar, tfull_bar, tempty_bar);
        }
    } else if (warp < 4 || warp >= 8) {
        // ===================== epilogue (both CTAs; each owns its 128 rows of the pair tile) =====================
        const uint32_t q4 = warp & 3;
        const uint32_t col_split = ((p.N / 16 + 1) / 2) * 16;
        const uint32_t col_begin = warp < 4 ? 0u : col_split, col_end = warp < 4 ? col_split : p.N;
        TcPending pd;
        pd.q = kTcNoPending; pd.pos = 0; pd.key = 0;
        const bool cosine = p.metric == METRIC_COSINE;
        auto tile_row = [&](uint64_t t) { return (first_tile + pair + t * npairs) * kPairRows + rank * kTcTileRows + q4 * 32 + lane; };
        auto row_ok = [&](uint64_t row) { return row < p.n_rows && row < p.row_end; };
        float nb_next = (cosine && my_tiles && row_ok(tile_row(0))) ? p.norms[tile_row(0)] : 0.0f;
        for (uint64_t t = 0; t < my_tiles; ++t) {
            const uint32_t buf = t & 1;
            const uint64_t row = tile_row(t);
            const bool valid = row_ok(row);
            const float nb = nb_next;
            if (cosine && t + 1 < my_tiles) { const uint64_t rn = tile_row(t + 1); nb_next = row_ok(rn) ? p.norms[rn] : 0.0f; }   // one tile ahead
            const float inv = cosine ? (nb > 0.0f ? rsqrtf(nb) : 0.0f) : 1.0f;
            float thr_pre[kTcThrRegs];
            if (p.sel_on) tc_thresholds_issue(p, thr_pre, col_begin, col_end, lane);
            mbar_wait(&tfull_bar[buf], (t >> 1) & 1);
            tc_fence_after();
            tc_flush_pending(p, pd);                                     // survivor parked by the previous tile
            tc_epilogue_tile(p, tmem_base + buf * p.N + ((q4 * 32u) << 16), s_nthr, row, valid, inv, col_begin, col_end, pd);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&tempty_bar[buf]);
                else mbar_arrive_cta(&tempty_bar[buf], 0);
            }
            if (p.sel_on) tc_thresholds_commit(p, thr_pre, s_nthr, col_begin, col_end, lane);
        }
        tc_flush_pending(p, pd);
        if (lane == 0) atomicAdd(s_epi_done, 1u);
    } else if (p.sel_on) {                                           // warps 6 and 7: selectors
        tc_selector_loop(p, s_sort + (size_t)(warp - 6) * p.sel_cap, s_epi_done, 2 * blockIdx.x + (warp - 6), 2 * gridDim.x, lane);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                              // nobody exits (or frees TMEM) while the peer can still signal it
    if (warp == 6) tmem_dealloc_2sm(tmem_base, p.tmem_cols);
}

// ---- per-query selection around the scan launches: keep each query's best kp candidates, publish the threshold --------
// grid = nq, block = 1024, dynamic smem = sort_cap * 8.
//   mode 0 (after the bootstrap range, before the main range): cand[q][0..count) -> sorted best kp in [0, kp) (0-padded),
//          [kp, count) cleared, count = consumed = kp: the main range appends behind the best list.
//   mode 1 (after the main range): [0, kp) U [consumed, count) -> sorted best kp in [0, kp), count = live keys among them.
//   mode 2 (the bootstrap range was the whole shard): like mode 0 but count = live keys.
__global__ void __launch_bounds__(1024) tc_select_kernel(uint64_t* __restrict__ cand, uint32_t* __restrict__ cand_count,
                                                         uint32_t* __restrict__ consumed, float* __restrict__ thr, uint32_t cap,
                                                         uint32_t kp, uint32_t sort_cap, uint32_t mode, uint32_t fixed_cnt) {
    extern __shared__ __align__(16) uint64_t s_sel_keys[];
    __shared__ uint32_t s_live;
    uint64_t* s_keys = s_sel_keys;
    const uint32_t q = blockIdx.x;
    uint32_t cnt = fixed_cnt ? fixed_cnt : cand_count[q];             // bootstrap range: one slot per row, no counter
    if (cnt > cap) cnt = cap;
    uint32_t cons = 0, total = cnt;
    if (mode == 1) {
        cons = consumed[q];
        if (cons < kp) cons = kp;
        if (cons > cnt) cons = cnt;
        total = (cnt >= kp) ? kp + (cnt - cons) : cnt;
    }
    uint32_t n = 2;
    while (n < total) n <<= 1;
    if (n > sort_cap) n = sort_cap;
    uint64_t* c = cand + (size_t)q * cap;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        uint64_t key = 0ull;
        if (i < total) key = (mode == 1 && i >= kp) ? c[cons + (i - kp)] : c[i];
        s_keys[i] = key;
    }
    if (threadIdx.x == 0) s_live = 0;
    __syncthreads();
    bitonic_sort_desc(s_keys, n, threadIdx.x, blockDim.x, 0);
    for (uint32_t i = threadIdx.x; i < kp && i < n; i += blockDim.x)
        if (s_keys[i] != 0ull && (i + 1 == kp || i + 1 == n || s_keys[i + 1] == 0ull)) s_live = i + 1;
    __syncthreads();
    const uint32_t keep = s_live;
    for (uint32_t i = threadIdx.x; i < kp; i += blockDim.x) c[i] = (i < keep) ? s_keys[i] : 0ull;
    if (mode == 0) for (uint32_t i = kp + threadIdx.x; i < cnt; i += blockDim.x) c[i] = 0ull;
    if (threadIdx.x == 0) {
        cand_count[q] = (mode == 0) ? kp : keep;
        if (mode == 0) consumed[q] = kp;
        thr[q] = (keep >= kp) ? key_score(s_keys[kp - 1], false) : __int_as_float(0xff800000);   // -inf until kp found
    }
}

// ---- query preparation: f32 queries -> B operand (fp16, or f32 read as TF32; zero padded), exact |q|^2, and the
// operand-rounding ratio rho[q] >= ||q - rounded(q)|| / ||q|| (inflated 1%) that feeds the error bound.  One octet per query.
template <typename TB>
__global__ void tc_prep_queries_kernel(const float* __restrict__ q, uint32_t qstride, uint32_t nq, uint32_t N, uint32_t d, uint32_t dpad,
                                       TB* __restrict__ B, float* __restrict__ na, float* __restrict__ rho, float* __restrict__ thr,
                                       uint32_t* __restrict__ cand_count, uint32_t* __restrict__ overflow) {
    const uint32_t octet = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int L = threadIdx.x & 7;
    const uint32_t qi = octet < N ? octet : N - 1;
    const bool real = octet < nq;
    const float* v = q + (size_t)(qi < nq ? qi : 0) * qstride;
    float nsq = sqnorm_octet(v, d, L);                     // reference order (adaptive) -> exact na
    float e2 = 0.0f, s2 = 0.0f;
    for (uint32_t i = L; i < dpad; i += 8) {
        float x = (real && i < d) ? v[i] : 0.0f;
        float xr;
        if (sizeof(TB) == 2) {
            __half h = __float2half_rn(x);
            if (octet < N) reinterpret_cast<__half*>(B)[(size_t)qi * dpad + i] = h;
            xr = __half2float(h);
        } else {
            if (octet < N) reinterpret_cast<float*>(B)[(size_t)qi * dpad + i] = x;
            xr = __uint_as_float(__float_as_uint(x) & 0xffffe000u);      // what kind::tf32 sees: low 13 mantissa bits dropped
        }
        float e = x - xr;
        e2 += e * e;
        s2 += x * x;
    }
    for (int o = 4; o; o >>= 1) { e2 += __shfl_xor_sync(0xffffffffu, e2, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if (L == 0 && octet < N) {
        na[qi] = real ? nsq : 0.0f;
        rho[qi] = (real && s2 > 0.0f) ? 1.01f * sqrtf(e2 / s2) + 1e-7f : 0.0f;
        thr[qi] = __int_as_float(0xff800000);
        cand_count[qi] = 0;
    }
    if (octet == 0 && L == 0) *overflow = 0;
}

// ---- exact re-scoring of the survivors (K4, batched): keys_in[q][i] (approx) -> keys_out[q][i] (exact score in `formula`) ----
// FORM_SIMD: adaptive_cosine_similarity order (one octet per pair).  FORM_SCALAR / FORM_SEQ: the reference's sequential
// cosines (simd_ops.rs:262-277, search.rs:524-532), inherently serial, run on octet lane 0.
template <typename T>
__global__ void tc_rescore_kernel(const T* __restrict__ rows, uint32_t d, uint32_t ld, const float* __restrict__ q, uint32_t qstride,
                                  const float* __restrict__ na, const float* __restrict__ norms, const uint64_t* __restrict__ keys_in,
                                  uint32_t cap, uint32_t kp, uint32_t nq, int metric, int formula, uint64_t row_offset,
                                  uint32_t blk_rows, uint32_t n_shards, uint64_t* __restrict__ keys_out) {
    const uint32_t octet = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int L = threadIdx.x & 7;
    const uint32_t total = nq * kp;
    const uint32_t pi = octet < total ? octet : total - 1;
    const uint32_t qi = pi / kp, ci = pi - qi * kp;
    const uint64_t key = keys_in[(size_t)qi * cap + ci];
    const bool present = key != 0ull;
    const uint32_t grow = key_row(key);
    // inverse of scan_global_row: contiguous shards (n_shards == 1) are global - offset, block-dealt shards drop the other
    // shards' blocks
    const uint64_t gl = present ? (uint64_t)grow - row_offset : 0;
    const uint64_t lrow = (gl / blk_rows / n_shards) * blk_rows + gl % blk_rows;
    const T* row = rows + lrow * ld;
    const float* qv = q + (size_t)qi * qstride;
    float res = 0.0f;
    if (formula == FORM_SIMD) {
        const float naq = na[qi];
        const float nb = (metric == METRIC_COSINE) ? norms[lrow] : 0.0f;
        if (metric == METRIC_COSINE) score_row_octet<T, METRIC_COSINE, 1>(row, qv, 0, d, L, &naq, nb, &res);
        else score_row_octet<T, METRIC_DOT, 1>(row, qv, 0, d, L, &naq, nb, &res);
    } else if (L == 0 && present) {
        float dp = 0.0f, sa = 0.0f, sb = 0.0f;
        if (formula == FORM_SCALAR) {                            // simd_ops.rs:262-277 (one interleaved loop)
            for (uint32_t i = 0; i < d; ++i) {
                const float va = qv[i], vb = ldf(row + i);
                dp = add_rn(dp, mul_rn(va, vb));
                sa = add_rn(sa, mul_rn(va, va));
                sb = add_rn(sb, mul_rn(vb, vb));
            }
            const float np = sqrt_rn(mul_rn(sa, sb));
            res = (np == 0.0f) ? 0.0f : div_rn(dp, np);
        } else {                                                 // search.rs:524-532
            for (uint32_t i = 0; i < d; ++i) dp = add_rn(dp, mul_rn(qv[i], ldf(row + i)));
            for (uint32_t i = 0; i < d; ++i) sa = add_rn(sa, mul_rn(qv[i], qv[i]));
            for (uint32_t i = 0; i < d; ++i) { const float vb = ldf(row + i); sb = add_rn(sb, mul_rn(vb, vb)); }
            sa = sqrt_rn(sa);
            sb = sqrt_rn(sb);
            res = (sa == 0.0f || sb == 0.0f) ? 0.0f : div_rn(dp, mul_rn(sa, sb));
        }
    }
    if (L == 0 && octet < total) keys_out[(size_t)qi * kp + ci] = present ? make_key(res, grow, false) : 0ull;
}

// ---- proof obligation per query: exact k-th score must beat (weakest kept approximate score + eps) ----------------
// approx list: cand[q][0..kp) sorted by approximate value v = dot/|row| (cosine) ; exact list: sorted exact keys.
// proven[q] = 1 when fewer than kp rows survived in total (nothing was discarded by rank) or
//             exact_k > v_kp / |q| + eps_q   with eps_q = rho_q + acc_bound(d).
__global__ void tc_verify_kernel(const uint64_t* __restrict__ approx, uint32_t cap, const uint32_t* __restrict__ counts, uint32_t kp,
                                 const uint64_t* __restrict__ exact_sorted, uint32_t k, const float* __restrict__ na,
                                 const float* __restrict__ rho, float acc_bound, int metric, uint32_t nq, uint32_t* __restrict__ proven,
                                 const uint32_t* __restrict__ overflow) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint32_t ok = 0;
    if (*overflow == 0) {
        if (counts[q] < kp) ok = 1;
        else {
            const float v_kp = key_score(approx[(size_t)q * cap + kp - 1], false);
            const uint64_t ek = exact_sorted[(size_t)q * kp + (k < kp ? k : kp) - 1];
            const float s_k = key_score(ek, false);
            if (metric == METRIC_COSINE) {
                const float nq_ = sqrtf(na[q]);
                if (nq_ > 0.0f && ek != 0ull && s_k == s_k) {
                    const float a = v_kp / nq_;
                    ok = (s_k > a + (rho[q] + acc_bound) * 1.0001f) ? 1u : 0u;
                }                                                // zero / NaN queries stay unproven -> exact-order fallback
            }
        }
    }
    proven[q] = ok;
}

}  // namespace cgv
