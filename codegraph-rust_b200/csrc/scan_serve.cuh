// scan_serve.cuh — K1s: the exact-order batch-1 scan as a RESIDENT kernel ("server").
//
// scan_exact_kernel (scan_exact.cuh) is launched once per query: at 8 GPUs the shard streams in ~53 us and every launch
// pays ~10 us on top (barrier init, query load, first-tile latency at the head; forced compaction, list write and a
// second launch for merge + exchange at the tail).  Here the CTAs stay resident across queries:
//
//   * the host submits a query by writing a descriptor into a pinned, device-mapped ring and bumping a doorbell word; the
//     POLLER warp of CTA 0 reads it over PCIe, stages host-resident queries into device memory and publishes the sequence
//     number on one 128-byte line per CTA;
//   * the PRODUCER warp of every CTA is query-agnostic: stage n always carries tile blockIdx + (n mod tiles) * grid (K1's
//     static assignment), so the shard is streamed round and round through the TMA ring — the first tiles of the next
//     query are already in shared memory while the consumers finish the current one;
//   * the 8 CONSUMER warps score rows in the reference's exact order with the very same code as K1 (score_row_octet) and
//     derive query and tile of a stage from its number alone; per query they only flip between two query buffers and two
//     candidate buffers;
//   * the HELPER warp prefetches the next query (and its squared norm, in the reference's order) into the other buffer;
//     at the end of a query it sorts the CTA's candidates, writes the CTA's list and counts it in; on the CTA that
//     delivered the LAST list of the query it merges all lists, runs the NVLink peer exchange on a sharded index (same
//     protocol and buffers as exchange.cuh), decodes the result straight into the submitter's buffers (pinned host memory
//     for host I/O) and publishes the completion word the host spins on.
//
// No launch, no memcpy node and no stream synchronisation per query.  The kernel leaves on its own after `idle_ns` without a
// doorbell (a resident grid owns every SM) and the host transparently relaunches it; every spin is bounded by a watchdog
// (`abort_ns`, __trap) so a protocol bug or a dead peer cannot wedge the GPU.
#pragma once
#include "common.cuh"
#include "exchange.cuh"
#include "scan_exact.cuh"

namespace cgv {

constexpr int kServeConsumerWarps = 8;
constexpr int kServeConsumerThreads = 32 * kServeConsumerWarps;
constexpr int kServeThreads = 32 * (kServeConsumerWarps + 3);      // + producer (8), helper (9) and poller (10, CTA 0 only) warps
constexpr uint32_t kServeSlots = 8;                                // descriptor ring depth = most queries in flight
constexpr uint32_t kServeMaxK = 64;
constexpr uint32_t kServeMaxGrid = 192;                            // CTAs of a session (one per SM)

struct ServeDesc {                   // one submitted query; written by the host BEFORE the doorbell is bumped
    uint64_t q_ptr;                  // device-accessible address of the query (qstride f32); pinned host memory when q_on_host
    uint64_t out_rows, out_scores, out_counts;   // device-accessible result addresses: u64[k], f32[k], u32[1]
    uint32_t q_on_host;
    uint32_t xseq;                   // sequence number of the peer exchange step (sharded index)
    uint32_t pad[6];
};
static_assert(sizeof(ServeDesc) == 64, "descriptor is one 64-byte line");

struct ServeHostBlock {              // pinned + mapped into the device address space
    volatile uint32_t doorbell;      // host -> device: highest sequence number submitted
    volatile uint32_t stop;          // host -> device: leave once everything submitted has been published
    uint32_t pad0[30];
    volatile uint32_t completed;     // device -> host: highest sequence number whose results are visible
    volatile uint32_t exit_seq;      // device -> host: nonzero once the kernel has decided to leave: first sequence NOT served
    volatile uint32_t error;         // device -> host: 1 + rank whose exchange list never arrived, or 0xdead0000 | code
    uint32_t pad1[29];
    ServeDesc desc[kServeSlots];
};

struct ServeCtrl {                   // device memory, reset by the host before every launch
    unsigned long long reserved0;
    uint32_t go;                     // highest sequence number published to the CTAs
    uint32_t exit_seq;               // 0 while running, else the first sequence number this launch does not serve
    uint32_t completed;              // highest sequence number finalised
    uint32_t pad[3];
    uint32_t done[kServeSlots];      // lists delivered, per sequence slot
    ServeDesc ddesc[kServeSlots];    // device copy of the descriptors (q_ptr already pointing into device memory)
    // One 128-byte line per CTA: [0] = go, [1] = exit_seq as seen by THAT CTA's helper warp.  148 helpers polling one word
    // would make a hot spot of a single L2 slice — and every TMA row load of the scan crosses every slice.
    uint32_t cta_line[kServeMaxGrid][32];
};

struct ServeParams {
    ScanParams sp;                   // rows, norms, n_rows, d, ld, row_words, tile_rows, stages, active_groups, k, cand_cap, row mapping
    uint32_t epoch_rounds;           // stages per consumer group between threshold syncs
    uint32_t start_seq;              // first sequence number this launch serves
    uint32_t off_ctl, off_merge, smem_total;   // extra shared-memory regions behind K1's layout: barriers + control words, merge staging
    uint32_t merge_lists;            // per-CTA lists the finishing CTA stages at once (multiple of 32, <= 256)
    uint32_t contig;                 // 1: every CTA streams one contiguous run of row tiles instead of K1's interleaved ones
    ServeCtrl* ctrl;
    ServeHostBlock* host;            // device pointer of the mapped block
    float* qbuf;                     // [kServeSlots][qstride] device staging of host-resident queries
    uint64_t* lists;                 // [kServeSlots][grid][k] per-CTA lists
    uint64_t idle_ns, life_ns, abort_ns;
    uint32_t rank, world;
    uint8_t* peer[kXchgMaxWorld];
    uint64_t xchg_timeout_ns;
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_volatile_sys(const volatile uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_volatile(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile(uint32_t* p, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
// watchdog of every spin in this kernel: a bug or a dead peer must end in a failed launch, never in a wedged GPU
// (%globaltimer is read only every 16384 spins: `t0` starts as 0 and is latched at the first check, so a wait that succeeds at
// once — every stage of the steady-state stream — never touches the timer)
#define CGV_SERVE_SPIN_GUARD(spins, t0, limit)                                             \
    if (((++(spins)) & 0x3fffu) == 0u) {                                                   \
        const uint64_t now_ = global_ns();                                                 \
        if ((t0) == 0ull) (t0) = now_;                                                     \
        else if (now_ - (t0) > (limit)) { __trap(); }                                      \
    }

// A warp-wide bitonic sort (descending) of n = 2^m keys in shared memory.
__device__ __forceinline__ void serve_warp_sort_desc(uint64_t* s, uint32_t n, uint32_t lane) {
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

// Shared-memory control words behind the mbarriers of the session (s_ctl[])
enum : uint32_t { kCtlExit = 0, kCtlProdDone = 1, kCtlIssued = 2, kCtlWords = 8 };

// Warp roles: 0-7 consumers, 8 producer, 9 helper (query prefetch, end-of-query sort / hand-over / finish), 10 poller (CTA 0).
//
// Per query the CONSUMERS only flip buffers: the query (and its squared norm, in the reference's order) was put into the other
// query buffer by the helper while the previous query was being scanned; at the end marker they leave their unsorted
// candidates in the current candidate buffer for the helper and move on to the next query's stages, which the producer has
// been streaming all along (static tile assignment blockIdx + i*grid, exactly K1's).  The HELPER sorts the candidates, writes
// the CTA's list, counts it in, and — on the CTA that delivered the last list of the query — merges all lists, runs the peer
// exchange, decodes into the submitter's buffers and publishes the completion.  None of that is on the scan's critical path.
template <typename T, int METRIC>
__global__ void __launch_bounds__(kServeThreads, 1) scan_serve_kernel(const ServeParams P) {
    extern __shared__ __align__(128) uint8_t smem_sv[];
    const ScanParams& p = P.sp;
    uint8_t* smem = smem_sv;
    const ScanSmemLayout lay = scan_smem_layout(p.row_words, p.tile_rows, p.stages, p.d, 2, p.cand_cap);   // two query / candidate buffers
    uint8_t* s_rows = smem + lay.off_rows;
    float* s_norms = reinterpret_cast<float*>(smem + lay.off_norms);
    float* s_q = reinterpret_cast<float*>(smem + lay.off_q);                    // [2][qstride]
    uint64_t* s_cand = reinterpret_cast<uint64_t*>(smem + lay.off_cand);        // [2][cand_cap]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + lay.off_bars);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* s_thr = reinterpret_cast<uint64_t*>(smem + lay.off_misc);         // [2]
    uint32_t* s_count = reinterpret_cast<uint32_t*>(smem + lay.off_misc + 16);  // [2]
    uint64_t* q_ready = reinterpret_cast<uint64_t*>(smem + P.off_ctl);          // [2] helper -> consumers: query buffer filled
    uint64_t* q_free = q_ready + 2;                                             // [2] consumers -> helper: query buffer no longer read
    uint64_t* cand_full = q_free + 2;                                           // [2] consumers -> helper: candidates of the query complete
    uint64_t* cand_free = cand_full + 2;                                        // [2] helper -> consumers: candidate buffer reset
    uint32_t* s_ctl = reinterpret_cast<uint32_t*>(cand_free + 2);               // [kCtlWords]
    float* s_na = reinterpret_cast<float*>(s_ctl + kCtlWords);                  // [2] squared query norms
    uint64_t* s_merge = reinterpret_cast<uint64_t*>(smem + P.off_merge);        // [merge_lists*k | 8*k | k]

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t warps_per_stage = p.tile_rows >> 2;
    const uint32_t ngroups = p.active_groups;                                   // == 8 / warps_per_stage (no idle groups in this kernel)
    const uint32_t qstride = ((p.d * 4 + 15) & ~15u) >> 2;
    const uint32_t stage_bytes = p.tile_rows * p.row_words * 4;
    const uint32_t row_bytes = p.ld * sizeof(T);
    const bool ascending = (METRIC == METRIC_L2);
    const uint64_t num_tiles = (p.n_rows + p.tile_rows - 1) / p.tile_rows;
    // Tile assignment: interleaved (CTA c takes tiles c, c + grid, ... — K1's) or, option `contig`, one contiguous run per CTA.
    const uint64_t per_cta = (num_tiles + gridDim.x - 1) / gridDim.x;
    const uint64_t my_tiles = P.contig ? (num_tiles > blockIdx.x * per_cta ? min(per_cta, num_tiles - blockIdx.x * per_cta) : 0)
                                       : ((num_tiles > blockIdx.x) ? (num_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0);
    const uint64_t tile_base = P.contig ? blockIdx.x * per_cta : blockIdx.x, tile_step = P.contig ? 1 : gridDim.x;
    const uint64_t t_launch = global_ns();

    if (tid == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], warps_per_stage); }
        for (int i = 0; i < 8; ++i) mbar_init(&q_ready[i], 1);
        s_thr[0] = 0; s_thr[1] = 0; s_count[0] = 0; s_count[1] = 0;
        for (uint32_t i = 0; i < kCtlWords; ++i) s_ctl[i] = 0;
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kServeConsumerWarps + 2) {
        // ===================== poller (CTA 0): host doorbell -> device-wide "go" =====================
        if (blockIdx.x != 0) return;
        uint32_t pub = P.start_seq - 1;
        uint64_t t_idle = global_ns();
        while (true) {
            uint32_t db = 0, stop = 0;
            if (lane == 0) { db = ld_volatile_sys(&P.host->doorbell); stop = ld_volatile_sys(&P.host->stop); }
            db = __shfl_sync(0xffffffffu, db, 0); stop = __shfl_sync(0xffffffffu, stop, 0);
            if ((int32_t)(db - pub) > 0) {
                for (uint32_t s = pub + 1; (int32_t)(db - s) >= 0; ++s) {
                    const uint32_t slot = s % kServeSlots;
                    uint32_t w = 0;                              // the descriptor: 16 words, one per lane
                    if (lane < 16) w = ld_volatile_sys(reinterpret_cast<const volatile uint32_t*>(&P.host->desc[slot]) + lane);
                    const uint32_t q_lo = __shfl_sync(0xffffffffu, w, 0), q_hi = __shfl_sync(0xffffffffu, w, 1);
                    const uint32_t on_host = __shfl_sync(0xffffffffu, w, 8);
                    uint64_t q_ptr = ((uint64_t)q_hi << 32) | q_lo;
                    if (on_host) {                               // stage the query into device memory: every CTA reads it from L2
                        const float4* src = reinterpret_cast<const float4*>(q_ptr);
                        float4* dst = reinterpret_cast<float4*>(P.qbuf + (size_t)slot * qstride);
                        const uint32_t n4 = qstride >> 2;
                        for (uint32_t i0 = 0; i0 < n4; i0 += 128) {          // 4 x 16-byte PCIe reads in flight per lane
                            float4 v[4];
#pragma unroll
                            for (uint32_t u = 0; u < 4; ++u) { const uint32_t i = i0 + u * 32 + lane; v[u] = i < n4 ? __ldcv(src + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
                            for (uint32_t u = 0; u < 4; ++u) { const uint32_t i = i0 + u * 32 + lane; if (i < n4) dst[i] = v[u]; }
                        }
                        q_ptr = reinterpret_cast<uint64_t>(dst);
                    }
                    uint32_t* dd = reinterpret_cast<uint32_t*>(&P.ctrl->ddesc[slot]);
                    if (lane == 0) dd[0] = (uint32_t)q_ptr;
                    else if (lane == 1) dd[1] = (uint32_t)(q_ptr >> 32);
                    else if (lane < 16) dd[lane] = w;
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) st_release_gpu(&P.ctrl->go, s);
                    for (uint32_t c = lane; c < gridDim.x; c += 32) st_release_gpu(&P.ctrl->cta_line[c][0], s);
                }
                pub = db;
                t_idle = global_ns();
            } else {
                const uint64_t now = global_ns();
                uint32_t comp = 0;
                if (lane == 0) comp = ld_acquire_gpu(&P.ctrl->completed);
                comp = __shfl_sync(0xffffffffu, comp, 0);
                if (comp != pub) t_idle = now;                   // work in flight is not idleness
                if ((stop && comp == pub) || now - t_idle > P.idle_ns || (now - t_launch > P.life_ns && comp == pub)) {
                    for (uint32_t c = lane; c < gridDim.x; c += 32) st_release_gpu(&P.ctrl->cta_line[c][1], pub + 1);
                    if (lane == 0) {
                        st_release_gpu(&P.ctrl->exit_seq, pub + 1);
                        __threadfence_system();
                        P.host->exit_seq = pub + 1;
                    }
                    break;
                }
            }
        }
        return;
    }

    if (warp == kServeConsumerWarps) {
        // ===================== producer: the row stream, round and round (K1's static tile assignment) =====================
        const uint64_t policy = l2_policy_evict_first();
        uint64_t n = 0;                                          // stages posted
        bool stopped = false;
        auto slot_free = [&]() -> bool {                         // false = the session is ending
            const uint32_t s = (uint32_t)(n % p.stages);
            if (n >= p.stages) {
                const uint32_t par = (uint32_t)(((n / p.stages) - 1) & 1);
                uint32_t spins = 0;
                uint64_t t0 = 0;
                while (!mbar_try_wait(&empty_bar[s], par)) {
                    if (lds_volatile(&s_ctl[kCtlExit])) return false;
                    CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns)
                }
            }
            return lds_volatile(&s_ctl[kCtlExit]) == 0;
        };
        // Stage n of this CTA carries tile blockIdx + (n mod my_tiles) * grid: the consumers derive the query and the tile of a
        // stage from its number alone, so nothing but rows travels through the ring.
        uint64_t i = 0;                                          // tile index within the current pass
        while (!stopped && my_tiles) {
            if (!slot_free()) { stopped = true; break; }
            const uint32_t s = (uint32_t)(n % p.stages);
            const uint64_t tile = tile_base + i * tile_step;
            const uint64_t row0 = tile * p.tile_rows;
            const uint32_t rows = (uint32_t)min((uint64_t)p.tile_rows, p.n_rows - row0);
            const bool with_norms = (METRIC == METRIC_COSINE);
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], rows * row_bytes + (with_norms ? p.tile_rows * 4 : 0));
            __syncwarp();
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.rows) + row0 * row_bytes;
            uint8_t* dst = s_rows + (size_t)s * stage_bytes;
            for (uint32_t r = lane; r < rows; r += 32) {
                if (p.use_l2_hint) bulk_g2s_hint(dst + (size_t)r * p.row_words * 4, src + (size_t)r * row_bytes, row_bytes, &full_bar[s], policy);
                else bulk_g2s(dst + (size_t)r * p.row_words * 4, src + (size_t)r * row_bytes, row_bytes, &full_bar[s]);
            }
            if (with_norms && lane == 0) bulk_g2s(s_norms + s * 32, p.norms + row0, p.tile_rows * 4, &full_bar[s]);
            ++n;
            if (++i == my_tiles) i = 0;
        }
        if (!my_tiles) {                                         // a CTA without rows only waits for the end of the session
            uint32_t spins = 0;
            uint64_t t0 = 0;
            while (lds_volatile(&s_ctl[kCtlExit]) == 0u) { __nanosleep(1000); CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns) }
        }
        if (lane == 0) { sts_volatile(&s_ctl[kCtlIssued], (uint32_t)n); __threadfence_block(); sts_volatile(&s_ctl[kCtlProdDone], 1u); }
        return;
    }

    if (warp == kServeConsumerWarps + 1) {
        // ===================== helper: query prefetch, end-of-query sort + hand-over, finishing a query =====================
        uint32_t jf = 0, jp = 0;                                 // queries fetched / post-processed by this CTA (index from start_seq)
        uint32_t go_seen = P.start_seq - 1, ex_seen = 0;         // last values read from this CTA's publication line
        const uint32_t* my_line = P.ctrl->cta_line[blockIdx.x];
        uint32_t spins = 0;
        uint64_t t0 = 0;
        while (true) {
            bool progressed = false;
            // ---- (1) the consumers have finished query jp: sort, deliver, maybe finish
            if (jp < jf && mbar_try_wait(&cand_full[jp & 1], (jp >> 1) & 1)) {
                const uint32_t b = jp & 1, seq = P.start_seq + jp, slot = seq % kServeSlots;
                uint64_t* cand = s_cand + (size_t)b * p.cand_cap;
                const uint32_t cnt = min(lds_volatile(&s_count[b]), p.cand_cap);
                for (uint32_t i = cnt + lane; i < p.cand_cap; i += 32) cand[i] = 0ull;
                __syncwarp();
                serve_warp_sort_desc(cand, p.cand_cap, lane);
                uint64_t* out = P.lists + ((size_t)slot * gridDim.x + blockIdx.x) * p.k;
                for (uint32_t i = lane; i < p.k; i += 32) out[i] = (i < cnt) ? cand[i] : 0ull;
                __threadfence();
                __syncwarp();
                uint32_t last = 0;
                if (lane == 0) {
                    last = atomicAdd(&P.ctrl->done[slot], 1u) == gridDim.x - 1 ? 1u : 0u;
                    s_count[b] = 0; s_thr[b] = 0ull;
                    mbar_arrive(&cand_free[b]);                  // the consumers may reuse this candidate buffer
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    __threadfence();
                    {   // queries are finished in order (the exchange's two-parity slots and the host's completion word rely on it)
                        uint32_t sp = 0;
                        uint64_t tw = 0;
                        while ((int32_t)(ld_acquire_gpu(&P.ctrl->completed) + 1 - seq) < 0) { CGV_SERVE_SPIN_GUARD(sp, tw, P.abort_ns) }
                    }
                    const uint32_t n_lists = gridDim.x;
                    uint64_t* staged = s_merge;                  // [merge_lists * k]
                    uint64_t* lvl = s_merge + (size_t)P.merge_lists * p.k;   // [8][k] winners of every group of 32 lists
                    uint64_t* best = lvl + 8 * p.k;              // [k]
                    const uint64_t* src = P.lists + (size_t)slot * gridDim.x * p.k;
                    const uint32_t nw = (n_lists + 31) / 32;
                    for (uint32_t l0 = 0; l0 < n_lists; l0 += P.merge_lists) {
                        const uint32_t ln = min(P.merge_lists, n_lists - l0), tot = ln * p.k;
                        for (uint32_t i0 = 0; i0 < tot; i0 += 512) {             // 16 loads in flight per lane
                            uint64_t v[16];
#pragma unroll
                            for (uint32_t u = 0; u < 16; ++u) { const uint32_t i = i0 + u * 32 + lane; v[u] = i < tot ? __ldcg(src + (size_t)l0 * p.k + i) : 0ull; }
#pragma unroll
                            for (uint32_t u = 0; u < 16; ++u) { const uint32_t i = i0 + u * 32 + lane; if (i < tot) staged[i] = v[u]; }
                        }
                        __syncwarp();
                        for (uint32_t g = 0; g * 32 < ln; ++g) {
                            warp_tournament_topk(staged + (size_t)g * 32 * p.k, min(32u, ln - g * 32), p.k, p.k, p.k, lvl + (size_t)(l0 / 32 + g) * p.k, lane);
                            __syncwarp();
                        }
                    }
                    warp_tournament_topk(lvl, nw, p.k, p.k, p.k, best, lane);
                    __syncwarp();
                    const ServeDesc* dd = &P.ctrl->ddesc[slot];
                    if (P.world > 1) {
                        // the peer exchange of exchange.cuh, executed by this warp: push, publish, wait, gather, merge
                        const uint32_t xseq = __ldcg(&dd->xseq), parity = xseq & 1u;
                        for (uint32_t i = lane; i < P.world * p.k; i += 32) {
                            const uint32_t r = i / p.k, j = i - r * p.k;
                            xchg_slot(P.peer[r], parity, P.rank, 0)[j] = best[j];
                        }
                        __threadfence_system();
                        __syncwarp();
                        if (lane < P.world) st_release_sys(&xchg_flags(P.peer[lane])[P.rank * kXchgMaxQ], xseq);
                        if (lane < P.world) {
                            const uint32_t* f = &xchg_flags(P.peer[P.rank])[lane * kXchgMaxQ];
                            const uint64_t tw = global_ns();
                            uint32_t sp = 0;
                            while ((int32_t)(ld_acquire_sys(f) - xseq) < 0) {
                                if ((++sp & 1023u) == 0 && global_ns() - tw > P.xchg_timeout_ns) { atomicExch_system(const_cast<uint32_t*>(&P.host->error), 1u + lane); break; }
                            }
                        }
                        __syncwarp();
                        for (uint32_t i = lane; i < P.world * p.k; i += 32) {
                            const uint32_t r = i / p.k, j = i - r * p.k;
                            staged[i] = __ldcg(xchg_slot(P.peer[P.rank], parity, r, 0) + j);
                        }
                        __syncwarp();
                        warp_tournament_topk(staged, P.world, p.k, p.k, p.k, best, lane);
                        __syncwarp();
                    }
                    uint64_t* o_rows = reinterpret_cast<uint64_t*>(__ldcg(reinterpret_cast<const unsigned long long*>(&dd->out_rows)));
                    float* o_scores = reinterpret_cast<float*>(__ldcg(reinterpret_cast<const unsigned long long*>(&dd->out_scores)));
                    uint32_t* o_counts = reinterpret_cast<uint32_t*>(__ldcg(reinterpret_cast<const unsigned long long*>(&dd->out_counts)));
                    uint32_t valid_n = 0;
                    for (uint32_t i = lane; i < p.k; i += 32) {
                        const uint64_t key = best[i];
                        const bool valid = key != 0ull;
                        if (o_rows) o_rows[i] = valid ? (uint64_t)key_row(key) : ~0ull;
                        if (o_scores) o_scores[i] = valid ? key_score(key, ascending) : 0.0f;
                        valid_n += valid;
                    }
                    valid_n = __reduce_add_sync(0xffffffffu, valid_n);
                    if (lane == 0 && o_counts) *o_counts = valid_n;
                    __threadfence_system();
                    __syncwarp();
                    if (lane == 0) {
                        P.ctrl->done[slot] = 0;
                        __threadfence_system();
                        st_release_gpu(&P.ctrl->completed, seq);
                        st_release_sys(const_cast<uint32_t*>(&P.host->completed), seq);
                    }
                    __syncwarp();
                }
                ++jp;
                progressed = true;
            }
            // ---- (2) the next query has been published and its buffer is free: fetch it, compute its squared norm
            if (jf - jp < 2) {
                const uint32_t seq = P.start_seq + jf, b = jf & 1;
                if ((int32_t)(go_seen - seq) < 0) {              // poll L2 only while the next query is not known to be published
                    uint32_t go = 0, ex = 0;
                    if (lane == 0) { go = ld_acquire_gpu(my_line); ex = ld_acquire_gpu(my_line + 1); }
                    go_seen = __shfl_sync(0xffffffffu, go, 0); ex_seen = __shfl_sync(0xffffffffu, ex, 0);
                }
                const uint32_t go = go_seen, ex = ex_seen;
                if ((int32_t)(go - seq) >= 0) {
                    if (mbar_try_wait(&q_free[b], ((jf >> 1) & 1) ^ 1)) {
                        const ServeDesc* dd = &P.ctrl->ddesc[seq % kServeSlots];
                        const float4* q = reinterpret_cast<const float4*>(__ldcg(reinterpret_cast<const unsigned long long*>(&dd->q_ptr)));
                        float4* dst = reinterpret_cast<float4*>(s_q + (size_t)b * qstride);
                        const uint32_t n4 = qstride >> 2;
                        for (uint32_t i0 = 0; i0 < n4; i0 += 256) {
                            float4 v[8];
#pragma unroll
                            for (uint32_t u = 0; u < 8; ++u) { const uint32_t i = i0 + u * 32 + lane; v[u] = i < n4 ? __ldcg(q + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
                            for (uint32_t u = 0; u < 8; ++u) { const uint32_t i = i0 + u * 32 + lane; if (i < n4) dst[i] = v[u]; }
                        }
                        __syncwarp();
                        const float na = (METRIC == METRIC_COSINE) ? sqnorm_octet(s_q + (size_t)b * qstride, p.d, (int)(lane & 7)) : 0.0f;
                        if (lane == 0) { s_na[b] = na; mbar_arrive(&q_ready[b]); }
                        __syncwarp();
                        ++jf;
                        progressed = true;
                    }
                } else if (ex != 0 && (int32_t)(seq - ex) >= 0 && jp == jf) {
                    if (lane == 0) sts_volatile(&s_ctl[kCtlExit], 1u);      // nothing more will be published to this launch
                    break;
                }
            }
            if (progressed) { spins = 0; t0 = 0; }
            else { __nanosleep(400); CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns) }
        }
        return;
    }

    // ===================== consumers: exact-order scoring and candidate filter =====================
    const uint32_t group = warp / warps_per_stage, sub = warp % warps_per_stage;
    const uint32_t ctid = tid, nct = kServeConsumerThreads;
    const int L = lane & 7;
    const uint32_t lrow = sub * 4 + (lane >> 3);
    const uint32_t epoch_tiles = P.epoch_rounds * ngroups;                      // tiles between threshold syncs (K1's sync_interval)
    const uint32_t flush_limit = p.cand_cap - epoch_tiles * p.tile_rows;

    uint64_t n = group;                                          // next stage of my group (stage numbers run on across queries)
    for (uint32_t j = 0;; ++j) {
        const uint32_t b = j & 1;
        uint64_t* cand = s_cand + (size_t)b * p.cand_cap;
        {   // the query buffer is filled and the candidate buffer reset, or the session is over
            uint32_t spins = 0;
            uint64_t t0 = 0;
            bool leave = false;
            while (!mbar_try_wait(&q_ready[b], (j >> 1) & 1)) {
                if (lds_volatile(&s_ctl[kCtlExit])) { leave = true; break; }
                CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns)
            }
            if (leave) break;
            while (!mbar_try_wait(&cand_free[b], ((j >> 1) & 1) ^ 1)) { CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns) }
        }
        const float* qv = s_q + (size_t)b * qstride;
        const float na = s_na[b];

        auto sync_and_maybe_compact = [&]() {
            named_bar_sync(1, nct);
            const uint32_t cnt = s_count[b];
            named_bar_sync(1, nct);
            if (cnt > flush_limit) {
                for (uint32_t i = cnt + ctid; i < p.cand_cap; i += nct) cand[i] = 0;
                named_bar_sync(1, nct);
                bitonic_sort_desc(cand, p.cand_cap, ctid, nct, 1);
                if (ctid == 0) {
                    const uint32_t keep = min(cnt, p.k);
                    s_count[b] = keep;
                    s_thr[b] = (keep >= p.k) ? cand[p.k - 1] : 0ull;
                }
                named_bar_sync(1, nct);
            }
        };

        // my group's stages of query j: global stage numbers [j * my_tiles, (j + 1) * my_tiles) congruent to `group`
        const uint64_t q_begin = (uint64_t)j * my_tiles, q_end = q_begin + my_tiles;
        uint64_t next_boundary = epoch_tiles;                    // threshold syncs happen at the same local tile indices in every group
        for (; n < q_end; n += ngroups) {
            const uint64_t i = n - q_begin;                      // tile index within this pass
            while (next_boundary <= i) { sync_and_maybe_compact(); next_boundary += epoch_tiles; }
            const uint32_t s = (uint32_t)(n % p.stages);
            {
                uint32_t spins = 0;
                uint64_t t0 = 0;
                const uint32_t par = (uint32_t)((n / p.stages) & 1);
                while (!mbar_try_wait(&full_bar[s], par)) { CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns) }
            }
            const uint64_t tile = tile_base + i * tile_step;
            const uint64_t row = tile * p.tile_rows + lrow;
            const T* rp = reinterpret_cast<const T*>(s_rows + (size_t)s * stage_bytes + (size_t)lrow * p.row_words * 4);
            const float nb = (METRIC == METRIC_COSINE) ? s_norms[s * 32 + lrow] : 0.0f;
            float sc[1];
            score_row_octet<T, METRIC, 1>(rp, qv, qstride, p.d, L, &na, nb, sc);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (L == 0 && row < p.n_rows) {
                const uint64_t key = make_key(sc[0], (uint32_t)scan_global_row(p, row), ascending);
                if (key > *reinterpret_cast<volatile uint64_t*>(&s_thr[b])) {
                    const uint32_t pos = atomicAdd(&s_count[b], 1u);
                    cand[pos] = key;
                }
            }
        }
        while (next_boundary < my_tiles) { sync_and_maybe_compact(); next_boundary += epoch_tiles; }
        // hand the candidates to the helper and move straight on to the next query's rows
        named_bar_sync(1, nct);
        if (ctid == 0) { mbar_arrive(&cand_full[b]); mbar_arrive(&q_free[b]); }
    }
    // ---- leaving: wait for every copy the producer has issued (shared memory must outlive the TMA writes)
    {
        uint32_t spins = 0;
        uint64_t t0 = 0;
        while (lds_volatile(&s_ctl[kCtlProdDone]) == 0u) { CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns) }
        __threadfence_block();
        const uint64_t issued = lds_volatile(&s_ctl[kCtlIssued]);    // stage counts stay far below 2^32 within one launch (life_ns)
        for (; n < issued; n += ngroups) {
            const uint32_t s = (uint32_t)(n % p.stages), par = (uint32_t)((n / p.stages) & 1);
            while (!mbar_try_wait(&full_bar[s], par)) { CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns) }
        }
    }
}

}  // namespace cgv
