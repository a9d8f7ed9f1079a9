// scan_serve.cuh — K1s: the exact-order batch-1 scan as a RESIDENT kernel ("server").
//
// scan_exact_kernel (scan_exact.cuh) is launched once per query: at 8 GPUs the shard streams in ~53 us and every launch
// pays ~10 us on top (barrier init, query load, first-tile latency at the head; forced compaction, list write and a
// second launch for merge + exchange at the tail).  Here the CTAs stay resident across queries:
//
//   * the host submits a query by writing a descriptor into a pinned, device-mapped ring and bumping a doorbell word; the
//     poller warp of CTA 0 reads it over PCIe, stages host-resident queries into device memory and publishes the sequence
//     number to the other CTAs (one L2 word);
//   * the PRODUCER warp of every CTA is query-agnostic: it claims chunks of row tiles from one monotonic ticket counter
//     (ticket -> query = ticket / chunks_per_query, so the matrix is simply streamed round and round) and keeps its TMA
//     ring full — the first tiles of the next query are already in shared memory while the consumers finish the current
//     one, and a CTA that falls behind (it merged the previous query) just claims fewer chunks;
//   * the 8 CONSUMER warps score rows in the reference's exact order with the very same code as K1 (score_row_octet),
//     keep the CTA's best k, and at the end-of-query marker write their list; the CTA that delivers the LAST list of a
//     query merges all of them, runs the NVLink peer exchange on a sharded index (same protocol and buffers as
//     exchange.cuh), decodes the result straight into the submitter's buffers (pinned host memory for host I/O) and
//     publishes the completion word the host spins on.
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
constexpr int kServeThreads = 32 * (kServeConsumerWarps + 2);      // + producer warp (8) + poller warp (9, CTA 0 only)
constexpr uint32_t kServeSlots = 8;                                // descriptor ring depth = most queries in flight
constexpr uint32_t kServeMaxK = 64;
constexpr uint32_t kServeMarkEnd = 0xffffffffu, kServeMarkEmpty = 0xfffffffeu;

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
    unsigned long long ticket;       // next chunk of row tiles
    uint32_t go;                     // highest sequence number published to the CTAs
    uint32_t exit_seq;               // 0 while running, else the first sequence number this launch does not serve
    uint32_t completed;              // highest sequence number finalised
    uint32_t pad[3];
    uint32_t done[kServeSlots];      // lists delivered, per sequence slot
    ServeDesc ddesc[kServeSlots];    // device copy of the descriptors (q_ptr already pointing into device memory)
};

struct ServeParams {
    ScanParams sp;                   // rows, norms, n_rows, d, ld, row_words, tile_rows, stages, active_groups, k, cand_cap, row mapping
    uint32_t epoch_rounds;           // stages per consumer group between threshold syncs
    uint32_t chunk_tiles;            // row tiles per ticket (multiple of the group count)
    uint32_t start_seq;              // first sequence number this launch serves
    uint32_t off_tile, off_ctl, off_merge, smem_total;   // extra shared-memory regions behind K1's layout
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
#define CGV_SERVE_SPIN_GUARD(spins, t0, limit)                                             \
    if (((++(spins)) & 0x3fffu) == 0u && global_ns() - (t0) > (limit)) { __trap(); }

enum : uint32_t { kServeRun = 1, kServeExit = 2 };

template <typename T, int METRIC>
__global__ void __launch_bounds__(kServeThreads, 1) scan_serve_kernel(const ServeParams P) {
    extern __shared__ __align__(128) uint8_t smem_sv[];
    const ScanParams& p = P.sp;
    uint8_t* smem = smem_sv;
    const ScanSmemLayout lay = scan_smem_layout(p.row_words, p.tile_rows, p.stages, p.d, 1, p.cand_cap);
    uint8_t* s_rows = smem + lay.off_rows;
    float* s_norms = reinterpret_cast<float*>(smem + lay.off_norms);
    float* s_q = reinterpret_cast<float*>(smem + lay.off_q);
    uint64_t* s_cand = reinterpret_cast<uint64_t*>(smem + lay.off_cand);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + lay.off_bars);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* s_thr = reinterpret_cast<uint64_t*>(smem + lay.off_misc);
    uint32_t* s_count = reinterpret_cast<uint32_t*>(smem + lay.off_misc + 8);
    uint32_t* s_tile = reinterpret_cast<uint32_t*>(smem + P.off_tile);          // [stages] row tile of the stage, or a marker
    uint32_t* s_ctl = reinterpret_cast<uint32_t*>(smem + P.off_ctl);            // [0] consumer broadcast, [1] exit request to the producer,
                                                                                // [2] producer stopped, [3] stages issued in total (lo), [4] last-list flag
    uint64_t* s_merge = reinterpret_cast<uint64_t*>(smem + P.off_merge);        // [grid*k | 8*k | k]

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t warps_per_stage = p.tile_rows >> 2;
    const uint32_t ngroups = p.active_groups;                                   // == 8 / warps_per_stage (no idle groups in this kernel)
    const uint32_t qstride = ((p.d * 4 + 15) & ~15u) >> 2;
    const uint32_t stage_bytes = p.tile_rows * p.row_words * 4;
    const uint32_t row_bytes = p.ld * sizeof(T);
    const bool ascending = (METRIC == METRIC_L2);
    const uint64_t num_tiles = (p.n_rows + p.tile_rows - 1) / p.tile_rows;
    const uint64_t chunks_per_q = (num_tiles + P.chunk_tiles - 1) / P.chunk_tiles;
    const uint64_t t_launch = global_ns();

    if (tid == 0) {
        for (uint32_t s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], warps_per_stage); }
        s_thr[0] = 0; s_count[0] = 0;
        for (int i = 0; i < 8; ++i) s_ctl[i] = 0;
        fence_mbar_init();
    }
    __syncthreads();

    if (warp == kServeConsumerWarps + 1) {
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
                    // the descriptor: 16 words, one per lane
                    uint32_t w = 0;
                    if (lane < 16) w = ld_volatile_sys(reinterpret_cast<const volatile uint32_t*>(&P.host->desc[slot]) + lane);
                    const uint32_t q_lo = __shfl_sync(0xffffffffu, w, 0), q_hi = __shfl_sync(0xffffffffu, w, 1);
                    const uint32_t on_host = __shfl_sync(0xffffffffu, w, 8);
                    uint64_t q_ptr = ((uint64_t)q_hi << 32) | q_lo;
                    if (on_host) {                               // stage the query into device memory: every CTA reads it from L2
                        const float* src = reinterpret_cast<const float*>(q_ptr);
                        float* dst = P.qbuf + (size_t)slot * qstride;
                        for (uint32_t i = lane; i < qstride; i += 32) dst[i] = __ldcv(src + i);
                        q_ptr = reinterpret_cast<uint64_t>(dst);
                    }
                    uint32_t* dd = reinterpret_cast<uint32_t*>(&P.ctrl->ddesc[slot]);
                    if (lane == 0) dd[0] = (uint32_t)q_ptr;
                    else if (lane == 1) dd[1] = (uint32_t)(q_ptr >> 32);
                    else if (lane < 16) dd[lane] = w;
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) st_release_gpu(&P.ctrl->go, s);
                }
                pub = db;
                t_idle = global_ns();
            } else {
                const uint64_t now = global_ns();
                if (stop || now - t_idle > P.idle_ns || now - t_launch > P.life_ns) {
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
        // ===================== producer: query-agnostic row stream =====================
        const uint64_t policy = l2_policy_evict_first();
        uint64_t n = 0;                                          // stages posted
        uint32_t cur_q = P.start_seq;                            // oldest query whose end marker is still to be posted
        bool stopped = false;
        // waits for stage slot (n % stages) to be free; false = the consumers asked us to stop
        auto slot_free = [&]() -> bool {
            const uint32_t s = (uint32_t)(n % p.stages);
            if (n >= p.stages) {
                const uint32_t par = (uint32_t)(((n / p.stages) - 1) & 1);
                uint32_t spins = 0;
                const uint64_t t0 = global_ns();
                while (!mbar_try_wait(&empty_bar[s], par)) {
                    if (lds_volatile(&s_ctl[1])) return false;
                    CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns)
                }
            }
            return lds_volatile(&s_ctl[1]) == 0;
        };
        auto post_marker = [&](uint32_t marker) -> bool {
            if (!slot_free()) return false;
            const uint32_t s = (uint32_t)(n % p.stages);
            if (lane == 0) { sts_volatile(&s_tile[s], marker); mbar_arrive(&full_bar[s]); }
            __syncwarp();
            ++n;
            return true;
        };
        auto claim = [&]() -> unsigned long long {
            unsigned long long t = 0;
            if (lane == 0) t = atomicAdd(&P.ctrl->ticket, 1ull);
            return __shfl_sync(0xffffffffu, t, 0);
        };
        unsigned long long tk = claim();
        while (!stopped) {
            const unsigned long long tk_next = claim();          // its round trip hides behind this chunk
            const uint32_t q_t = P.start_seq + (uint32_t)(tk / chunks_per_q);
            const uint64_t chunk = tk % chunks_per_q;
            while (cur_q != q_t && !stopped) {                   // close every query older than this ticket's (also ones we got no chunk of)
                while ((n % ngroups) && !stopped) stopped = !post_marker(kServeMarkEmpty);
                for (uint32_t g = 0; g < ngroups && !stopped; ++g) stopped = !post_marker(kServeMarkEnd);
                ++cur_q;
            }
            const uint64_t tile0 = chunk * P.chunk_tiles;
            for (uint32_t c = 0; c < P.chunk_tiles && !stopped; ++c) {
                const uint64_t tile = tile0 + c;
                if (tile >= num_tiles) { stopped = !post_marker(kServeMarkEmpty); continue; }   // keeps chunks group-aligned
                if (!slot_free()) { stopped = true; break; }
                const uint32_t s = (uint32_t)(n % p.stages);
                const uint64_t row0 = tile * p.tile_rows;
                const uint32_t rows = (uint32_t)min((uint64_t)p.tile_rows, p.n_rows - row0);
                const bool with_norms = (METRIC == METRIC_COSINE);
                if (lane == 0) {
                    sts_volatile(&s_tile[s], (uint32_t)tile);
                    mbar_arrive_expect_tx(&full_bar[s], rows * row_bytes + (with_norms ? p.tile_rows * 4 : 0));
                }
                __syncwarp();
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.rows) + row0 * row_bytes;
                uint8_t* dst = s_rows + (size_t)s * stage_bytes;
                for (uint32_t r = lane; r < rows; r += 32)
                    bulk_g2s_hint(dst + (size_t)r * p.row_words * 4, src + (size_t)r * row_bytes, row_bytes, &full_bar[s], policy);
                if (with_norms && lane == 0) bulk_g2s(s_norms + s * 32, p.norms + row0, p.tile_rows * 4, &full_bar[s]);
                ++n;
            }
            tk = tk_next;
        }
        if (lane == 0) { sts_volatile(&s_ctl[3], (uint32_t)n); __threadfence_block(); sts_volatile(&s_ctl[2], 1u); }
        return;
    }

    // ===================== consumers: exact-order scoring, candidate filter, per-query hand-over =====================
    const uint32_t group = warp / warps_per_stage, sub = warp % warps_per_stage;
    const uint32_t ctid = tid, nct = kServeConsumerThreads;
    const int L = lane & 7;
    const uint32_t lrow = sub * 4 + (lane >> 3);
    const uint32_t flush_limit = p.cand_cap - P.epoch_rounds * ngroups * p.tile_rows;

    auto sync_and_maybe_compact = [&](bool force) {
        named_bar_sync(1, nct);
        const uint32_t cnt = s_count[0];
        named_bar_sync(1, nct);
        if (force || cnt > flush_limit) {
            for (uint32_t i = cnt + ctid; i < p.cand_cap; i += nct) s_cand[i] = 0;
            named_bar_sync(1, nct);
            bitonic_sort_desc(s_cand, p.cand_cap, ctid, nct, 1);
            if (ctid == 0) {
                const uint32_t keep = min(cnt, p.k);
                s_count[0] = keep;
                s_thr[0] = (keep >= p.k) ? s_cand[p.k - 1] : 0ull;
            }
            named_bar_sync(1, nct);
        }
    };

    uint64_t n = group;                                          // next stage of my group
    uint32_t cur = P.start_seq;
    while (true) {
        // ---- wait until query `cur` is published (and the query two before it is out of the way), or for the exit decision
        if (ctid == 0) {
            uint32_t spins = 0, flag = 0;
            const uint64_t t0 = global_ns();
            while (true) {
                const uint32_t go = ld_acquire_gpu(&P.ctrl->go);
                if ((int32_t)(go - cur) >= 0) {
                    const uint32_t comp = ld_acquire_gpu(&P.ctrl->completed);
                    if ((int32_t)(comp + 2 - cur) >= 0) { flag = kServeRun; break; }
                } else {
                    const uint32_t ex = ld_acquire_gpu(&P.ctrl->exit_seq);
                    if (ex != 0 && (int32_t)(cur - ex) >= 0) { flag = kServeExit; break; }
                }
                CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns)
            }
            sts_volatile(&s_ctl[0], flag);
        }
        named_bar_sync(1, nct);
        const uint32_t flag = lds_volatile(&s_ctl[0]);
        if (flag != kServeRun) break;
        const uint32_t slot = cur % kServeSlots;
        {
            const ServeDesc* dd = &P.ctrl->ddesc[slot];
            const float* q = reinterpret_cast<const float*>(__ldcg(reinterpret_cast<const unsigned long long*>(&dd->q_ptr)));
            for (uint32_t i = ctid; i < qstride; i += nct) s_q[i] = __ldcg(q + i);
            if (ctid == 0) { s_thr[0] = 0; s_count[0] = 0; }
        }
        named_bar_sync(1, nct);
        float na = (METRIC == METRIC_COSINE) ? sqnorm_octet(s_q, p.d, L) : 0.0f;
        na = __shfl_sync(0xffffffffu, na, lane & ~7);

        // ---- my group's stages of this query, up to its end marker
        uint32_t rounds = 0;
        while (true) {
            const uint32_t s = (uint32_t)(n % p.stages);
            {
                uint32_t spins = 0;
                const uint64_t t0 = global_ns();
                const uint32_t par = (uint32_t)((n / p.stages) & 1);
                while (!mbar_try_wait(&full_bar[s], par)) { CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns) }
            }
            const uint32_t mark = lds_volatile(&s_tile[s]);
            if (mark == kServeMarkEnd || mark == kServeMarkEmpty) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                n += ngroups;
                if (mark == kServeMarkEnd) break;
            } else {
                const uint64_t row = (uint64_t)mark * p.tile_rows + lrow;
                const T* rp = reinterpret_cast<const T*>(s_rows + (size_t)s * stage_bytes + (size_t)lrow * p.row_words * 4);
                const float nb = (METRIC == METRIC_COSINE) ? s_norms[s * 32 + lrow] : 0.0f;
                float sc[1];
                score_row_octet<T, METRIC, 1>(rp, s_q, qstride, p.d, L, &na, nb, sc);
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
                if (L == 0 && row < p.n_rows) {
                    const uint64_t key = make_key(sc[0], (uint32_t)scan_global_row(p, row), ascending);
                    if (key > *reinterpret_cast<volatile uint64_t*>(&s_thr[0])) {
                        const uint32_t pos = atomicAdd(&s_count[0], 1u);
                        s_cand[pos] = key;
                    }
                }
                n += ngroups;
            }
            if (++rounds % P.epoch_rounds == 0) sync_and_maybe_compact(false);
        }
        sync_and_maybe_compact(true);
        // ---- deliver this CTA's list; the CTA that delivers the last one finishes the query
        {
            const uint32_t cnt = s_count[0];
            uint64_t* out = P.lists + ((size_t)slot * gridDim.x + blockIdx.x) * p.k;
            for (uint32_t i = ctid; i < p.k; i += nct) out[i] = (i < cnt) ? s_cand[i] : 0ull;
        }
        __threadfence();
        named_bar_sync(1, nct);
        if (ctid == 0) {
            const uint32_t old = atomicAdd(&P.ctrl->done[slot], 1u);
            sts_volatile(&s_ctl[4], old == gridDim.x - 1 ? 1u : 0u);
        }
        named_bar_sync(1, nct);
        if (lds_volatile(&s_ctl[4])) {
            __threadfence();
            // queries are finalised in order (the exchange's two-parity slots and the host's completion word rely on it)
            if (ctid == 0) {
                uint32_t spins = 0;
                const uint64_t t0 = global_ns();
                while ((int32_t)(ld_acquire_gpu(&P.ctrl->completed) + 1 - cur) < 0) { CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns) }
            }
            named_bar_sync(1, nct);
            const uint32_t cwarp = warp;
            const uint32_t n_lists = gridDim.x, total = n_lists * p.k;
            uint64_t* staged = s_merge;                          // [n_lists * k]
            uint64_t* lvl = s_merge + total;                     // [8][k]
            uint64_t* best = lvl + 8 * p.k;                      // [k]
            const uint64_t* src = P.lists + (size_t)slot * gridDim.x * p.k;
            for (uint32_t i = ctid; i < total; i += nct) staged[i] = __ldcg(src + i);
            named_bar_sync(1, nct);
            const uint32_t nw = (n_lists + 31) / 32;
            if (cwarp < nw) warp_tournament_topk(staged + (size_t)cwarp * 32 * p.k, min(32u, n_lists - cwarp * 32), p.k, p.k, p.k, lvl + (size_t)cwarp * p.k, lane);
            named_bar_sync(1, nct);
            if (cwarp == 0) warp_tournament_topk(lvl, nw, p.k, p.k, p.k, best, lane);
            named_bar_sync(1, nct);
            const ServeDesc* dd = &P.ctrl->ddesc[slot];
            if (P.world > 1) {
                // the peer exchange of exchange.cuh, executed by this CTA: push, publish, wait, gather, merge
                const uint32_t xseq = __ldcg(&dd->xseq), parity = xseq & 1u;
                for (uint32_t i = ctid; i < P.world * p.k; i += nct) {
                    const uint32_t r = i / p.k, j = i - r * p.k;
                    xchg_slot(P.peer[r], parity, P.rank, 0)[j] = best[j];
                }
                __threadfence_system();
                named_bar_sync(1, nct);
                if (ctid < P.world) st_release_sys(&xchg_flags(P.peer[ctid])[P.rank * kXchgMaxQ], xseq);
                if (ctid < P.world) {
                    const uint32_t* f = &xchg_flags(P.peer[P.rank])[ctid * kXchgMaxQ];
                    const uint64_t t0 = global_ns();
                    uint32_t spins = 0;
                    while ((int32_t)(ld_acquire_sys(f) - xseq) < 0) {
                        if ((++spins & 1023u) == 0 && global_ns() - t0 > P.xchg_timeout_ns) { atomicExch_system(const_cast<uint32_t*>(&P.host->error), 1u + ctid); break; }
                    }
                }
                named_bar_sync(1, nct);
                for (uint32_t i = ctid; i < P.world * p.k; i += nct) {
                    const uint32_t r = i / p.k, j = i - r * p.k;
                    staged[i] = __ldcg(xchg_slot(P.peer[P.rank], parity, r, 0) + j);
                }
                named_bar_sync(1, nct);
                if (cwarp == 0) warp_tournament_topk(staged, P.world, p.k, p.k, p.k, best, lane);
                named_bar_sync(1, nct);
            }
            uint64_t* o_rows = reinterpret_cast<uint64_t*>(__ldcg(reinterpret_cast<const unsigned long long*>(&dd->out_rows)));
            float* o_scores = reinterpret_cast<float*>(__ldcg(reinterpret_cast<const unsigned long long*>(&dd->out_scores)));
            uint32_t* o_counts = reinterpret_cast<uint32_t*>(__ldcg(reinterpret_cast<const unsigned long long*>(&dd->out_counts)));
            uint32_t cnt = 0;
            for (uint32_t i = ctid; i < p.k; i += nct) {
                const uint64_t key = best[i];
                const bool valid = key != 0ull;
                if (o_rows) o_rows[i] = valid ? (uint64_t)key_row(key) : ~0ull;
                if (o_scores) o_scores[i] = valid ? key_score(key, ascending) : 0.0f;
                cnt += valid;
            }
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (lane == 0 && cnt) atomicAdd(&s_ctl[5], cnt);
            __threadfence_system();
            named_bar_sync(1, nct);
            if (ctid == 0) {
                if (o_counts) *o_counts = lds_volatile(&s_ctl[5]);
                sts_volatile(&s_ctl[5], 0u);
                P.ctrl->done[slot] = 0;
                __threadfence_system();
                st_release_gpu(&P.ctrl->completed, cur);
                st_release_sys(const_cast<uint32_t*>(&P.host->completed), cur);
            }
            named_bar_sync(1, nct);
        }
        ++cur;
    }
    // ---- leaving: tell the producer, then wait for every copy it has issued (shared memory must outlive the TMA writes)
    if (ctid == 0) sts_volatile(&s_ctl[1], 1u);
    {
        uint32_t spins = 0;
        const uint64_t t0 = global_ns();
        while (lds_volatile(&s_ctl[2]) == 0u) { CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns) }
        __threadfence_block();
        const uint64_t issued = lds_volatile(&s_ctl[3]);         // stage counts stay far below 2^32 within one launch (life_ns)
        for (; n < issued; n += ngroups) {
            const uint32_t s = (uint32_t)(n % p.stages), par = (uint32_t)((n / p.stages) & 1);
            while (!mbar_try_wait(&full_bar[s], par)) { CGV_SERVE_SPIN_GUARD(spins, t0, P.abort_ns) }
        }
    }
}

}  // namespace cgv
