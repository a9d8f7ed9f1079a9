// cgvec_api.cu — the C ABI of libcgvec_b200.so (include/cgvec.h): index lifecycle, the write side
// (store_embeddings), the exported search / re-score / int8 / file entry points and the host-side helpers.
// The orchestration lives in three includes: host_scan.inl (exact-order path: plan, launch, merge, exchange),
// host_tensor.inl (tcgen05 path: row ranges, thresholds, exact re-score + proof) and multi_device.inl
// (one process driving several GPUs).  No CPU scoring path exists: without an sm_100 device every compute
// entry point returns CGVEC_ERR_NO_DEVICE.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/cgvec.h"
#include "aux_kernels.cuh"
#include "common.cuh"
#include "exchange.cuh"
#include "nccl_dyn.h"
#include "scan_exact.cuh"
#include "scan_i8.cuh"
#include "scan_serve.cuh"
#include "scan_tc.cuh"

#define CGVEC_EXPORT extern "C" __attribute__((visibility("default")))

namespace {

using namespace cgv;

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            int code_ = (e_ == cudaErrorMemoryAllocation) ? CGVEC_ERR_OOM                                \
                        : (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? CGVEC_ERR_NO_DEVICE \
                                                                                          : CGVEC_ERR_CUDA; \
            return fail(code_, "%s failed: %s", #expr, cudaGetErrorString(e_));                          \
        }                                                                                                \
    } while (0)

#define NCCL_TRY(expr)                                                                           \
    do {                                                                                         \
        int r_ = (expr);                                                                         \
        if (r_ != kNcclSuccess) return fail(CGVEC_ERR_NCCL, "%s failed: %s", #expr, nccl_api().GetErrorString(r_)); \
    } while (0)

constexpr uint32_t kMaxK = 1024;
constexpr uint32_t kSmemBudget = 227 * 1024;
constexpr uint32_t kMergeMaxKeys = 8192;

struct IdKey {
    uint64_t lo, hi;
    bool operator==(const IdKey& o) const { return lo == o.lo && hi == o.hi; }
};
struct IdHash {
    size_t operator()(const IdKey& k) const {
        uint64_t x = k.lo * 0x9E3779B97F4A7C15ull ^ (k.hi + 0x7F4A7C15ull + (k.lo << 6) + (k.lo >> 2));
        return (size_t)(x ^ (x >> 29));
    }
};
IdKey id_key(const uint8_t id[16]) {
    IdKey k;
    memcpy(&k.lo, id, 8);
    memcpy(&k.hi, id + 8, 8);
    return k;
}

uint32_t next_pow2(uint32_t v) {
    uint32_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

// Per-call scratch: a stream, device buffers and pinned host mirrors.  Pooled so that search() is
// reentrant from many host threads (multi_vector_search fans out concurrent calls, search.rs:358-361).
struct SearchCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    float* d_q = nullptr;        size_t q_cap = 0;       // floats
    uint64_t* d_part[2] = {nullptr, nullptr}; size_t part_cap = 0;   // keys (merge ping-pong)
    uint64_t* d_scan[2] = {nullptr, nullptr}; size_t scan_cap = 0; uint32_t scan_flip = 0;   // per-CTA partials, double buffered (PDL)
    uint64_t* d_gather = nullptr; size_t gather_cap = 0;  // keys
    uint64_t* d_rows = nullptr;  float* d_scores = nullptr; uint32_t* d_counts = nullptr; size_t out_cap = 0, cnt_cap = 0;
    uint64_t* d_tmp_rows = nullptr; float* d_tmp_scores = nullptr; size_t tmp_cap = 0;
    // tensor path (scan_tc.cuh)
    __half* d_B = nullptr;       size_t B_cap = 0;
    float* d_tc_f = nullptr;     size_t tcf_cap = 0;      // thr[N] | na[N] | rho[N]
    uint32_t* d_tc_u = nullptr;  size_t tcu_cap = 0;      // count[N] | proven[N] | overflow
    uint64_t* d_cand = nullptr;  size_t cand_cap = 0;     // [N][cap]
    uint64_t* d_exact = nullptr; size_t exact_cap = 0;    // [N][kp] exact keys
    uint32_t* h_proven = nullptr; size_t hprov_cap = 0;
    float* h_q = nullptr;        size_t hq_cap = 0;
    uint64_t* h_rows = nullptr;  float* h_scores = nullptr; uint32_t* h_counts = nullptr; size_t hout_cap = 0, hcnt_cap = 0;
};

struct ScanGeom {
    uint32_t tile_rows, stages, groups, sync_interval, cand_cap, row_words, grid, smem, pdl;
};

struct Index {
    uint32_t dim = 0, ld = 0, esize = 4;
    cgvec_dtype dtype = CGVEC_F32;
    int device = 0;
    int rank = 0, world = 1;
    uint64_t row_offset = 0;
    uint32_t blk_rows = 1u << 30, n_shards = 1, shard_id = 0;   // global row = ((local/blk)*n_shards + shard_id)*blk + local%blk + row_offset
    std::vector<Index*> parts;    // non-empty on the parent of a single-process multi-device index (one shard per device)
    int sm_count = 0;

    void* d_rows = nullptr;
    float* d_norms = nullptr;
    uint64_t n = 0, cap = 0;
    // int8 quantised copy (scan_i8.cuh), built by cgvec_quantize_i8
    uint8_t* d_codes = nullptr;
    int32_t* d_inorms = nullptr;
    uint64_t n_codes = 0;
    uint32_t ld8 = 0;
    std::mutex i8_mu;
    std::unordered_map<uint32_t, uint64_t> i8_fill_end;   // limit -> row of the limit-th non-zero-norm code row (valid while the codes are)

    std::vector<uint8_t> ids;     // 16 bytes per local row
    std::vector<uint8_t> has_id;  // 1 per local row
    std::unordered_map<IdKey, uint64_t, IdHash> id2row;

    nccl_comm_t comm = nullptr;
    std::mutex comm_mu;           // collectives must be issued in the same order on every rank
    // peer-memory exchange (exchange.cuh): own buffer + every rank's buffer mapped through CUDA IPC
    uint8_t* xbuf = nullptr;
    uint8_t* xpeer[kXchgMaxWorld] = {nullptr};
    bool p2p = false;
    uint32_t xseq = 0;
    int opt_p2p = 1;
    uint32_t* h_xerr = nullptr;   // pinned, device-visible: the exchange kernel reports a rank whose list never arrived
    int opt_xchg_timeout_ms = 5000;

    // rank-invariant view of a sharded index: smallest / largest shard, agreed with one all-gather after the rows change
    uint64_t agreed_min_n = 0, agreed_max_n = 0;
    bool agreed_valid = false;

    std::mutex pool_mu;
    std::vector<SearchCtx*> pool;
    cudaStream_t main_stream = nullptr;

    // knobs (cgvec_set_option)
    int opt_tile_rows = 0, opt_stages = 0, opt_sync = 0, opt_l2_hint = 0, opt_grid = 0, opt_timing = 0, opt_max_nq = kScanMaxQ;
    int opt_pdl = 2;            // 0: plain launches, 1: merge releases the next scan early, 2: full programmatic chain
    // launch timeline (option "trace"): [kind, start ns, end ns] per traced launch, device-resident until read back
    uint64_t* d_trace = nullptr;
    std::vector<uint32_t> trace_kinds;      // 1 = scan, 2 = merge, 3 = exchange
    static constexpr uint32_t kTraceCap = 16384;
    int opt_tc_target = 0, opt_tc_l2promo = 2, opt_tc_prefetch = 0, opt_tc_first = 0, opt_tc_kernel = 0, opt_tc_debug = 0, opt_tc2_max_n = kTc2MaxN;
    int opt_tc_min_nq = 0, opt_tc_stages = 0, opt_tc_max_n = kTcMaxN, opt_tc_margin = 0, opt_tc_kbs = 0, opt_tc_flow = 0;
    std::atomic<uint64_t> tc_batches{0}, tc_fallbacks{0};

    // stats
    std::atomic<uint64_t> launches{0}, searches{0};
    ScanGeom last_geom{};
    std::mutex ev_mu;
    struct TimedLaunch { cudaEvent_t e0, e1; int kind; };     // kind 0: exact scan kernel, 1: main range of the tensor scan
    std::vector<TimedLaunch> timed;                           // kernel brackets awaiting readout
    double scan_ms_total = 0.0, tc_main_ms_total = 0.0;
    uint64_t scan_timed = 0, tc_main_timed = 0;
    float last_scan_ms = 0.0f;
    std::atomic<uint32_t> last_exchange{0};                   // 0 none (unsharded), 1 fused peer-memory kernel, 2 NCCL all-gather

    // reader-writer exclusion between the write side (add / reserve / normalize / fill / load / quantize: may reallocate the
    // device matrix and rehash the id map) and everything that reads them; taken by the exported entry points (RwGuard)
    std::shared_mutex rw_mu;
    std::atomic<int> open_streams{0};                         // cgvec_stream_* sessions keep device work in flight between calls

    // group commit of concurrent batch-1 callers (coalesced_search below)
    struct PendingSearch;
    std::mutex co_mu;
    std::condition_variable co_cv;
    std::deque<PendingSearch*> co_queue;
    bool co_leader = false;
    int opt_coalesce = 1, opt_coalesce_max = 16;
    std::atomic<uint64_t> co_batches{0}, co_queries{0};       // batches of >= 2 callers, callers served through them
};

// Scoped reader/writer lock of an index for one exported call.  Exported functions call each other (cgvec_get -> cgvec_get_row,
// the multi-device parent -> its parts); a thread that already holds an index does not lock it again.
thread_local std::vector<const void*> tl_held_indexes;
template <typename IndexT>
struct RwGuardT {
    IndexT* ix; int mode = 0;                                  // 0 not taken (NULL index or nested), 1 shared, 2 exclusive
    RwGuardT(const IndexT* cix, bool write) : ix(const_cast<IndexT*>(cix)) {
        if (!ix) return;
        for (const void* h : tl_held_indexes) if (h == ix) return;
        if (write) ix->rw_mu.lock(); else ix->rw_mu.lock_shared();
        mode = write ? 2 : 1;
        tl_held_indexes.push_back(ix);
    }
    ~RwGuardT() {
        if (!mode) return;
        for (size_t i = tl_held_indexes.size(); i-- > 0;) if (tl_held_indexes[i] == ix) { tl_held_indexes.erase(tl_held_indexes.begin() + i); break; }
        if (mode == 2) ix->rw_mu.unlock(); else ix->rw_mu.unlock_shared();
    }
    RwGuardT(const RwGuardT&) = delete;
    RwGuardT& operator=(const RwGuardT&) = delete;
};
using RwGuard = RwGuardT<Index>;
#define CGVEC_WRITE_GUARD(ix)                                                                                              \
    RwGuard rw_guard_(ix, true);                                                                                           \
    if ((ix) && (ix)->open_streams.load() > 0) return fail(CGVEC_ERR_UNSUPPORTED, "close the index's cgvec_stream sessions before writing to it")

// NVTX range around a host-side phase (SURVEY §5 aux: tracing).  Header-only NVTX3: a no-op costing one predicted branch unless
// a tool (nsys, ncu --nvtx) injected itself into the process.  Names: cgvec.<phase>.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

int check_device(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(CGVEC_ERR_NO_DEVICE, "no CUDA device available (%s); libcgvec_b200 has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= count) return fail(CGVEC_ERR_BAD_ARG, "device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(CGVEC_ERR_NO_DEVICE, "device %d is sm_%d%d; libcgvec_b200 is built for sm_100a only", device, prop.major,
                    prop.minor);
    return CGVEC_OK;
}

template <typename T>
int ensure(T** p, size_t* cap, size_t need, bool pinned = false) {
    if (need <= *cap && *p) return CGVEC_OK;
    if (*p) { if (pinned) cudaFreeHost(*p); else cudaFree(*p); *p = nullptr; }
    size_t c = need < 64 ? 64 : need;
    if (pinned) CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(p), c * sizeof(T)));
    else CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(p), c * sizeof(T)));
    *cap = c;
    return CGVEC_OK;
}

int ctx_acquire(Index* ix, SearchCtx** out) {
    {
        std::lock_guard<std::mutex> g(ix->pool_mu);
        if (!ix->pool.empty()) { *out = ix->pool.back(); ix->pool.pop_back(); return CGVEC_OK; }
    }
    auto* c = new SearchCtx();
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming);
    if (e != cudaSuccess) { delete c; return fail(CGVEC_ERR_CUDA, "stream/event create failed: %s", cudaGetErrorString(e)); }
    *out = c;
    return CGVEC_OK;
}
void ctx_release(Index* ix, SearchCtx* c) {
    std::lock_guard<std::mutex> g(ix->pool_mu);
    ix->pool.push_back(c);
}
void ctx_free(SearchCtx* c) {
    cudaFree(c->d_q); cudaFree(c->d_part[0]); cudaFree(c->d_part[1]); cudaFree(c->d_scan[0]); cudaFree(c->d_scan[1]); cudaFree(c->d_gather);
    cudaFree(c->d_rows); cudaFree(c->d_scores); cudaFree(c->d_counts); cudaFree(c->d_tmp_rows); cudaFree(c->d_tmp_scores);
    cudaFree(c->d_B); cudaFree(c->d_tc_f); cudaFree(c->d_tc_u); cudaFree(c->d_cand); cudaFree(c->d_exact); cudaFreeHost(c->h_proven);
    cudaFreeHost(c->h_q); cudaFreeHost(c->h_rows); cudaFreeHost(c->h_scores); cudaFreeHost(c->h_counts);
    if (c->done) cudaEventDestroy(c->done);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int grow(Index* ix, uint64_t need, bool exact = false) {
    if (need <= ix->cap) return CGVEC_OK;
    uint64_t newcap = ix->cap ? ix->cap * 2 : 1024;
    if (newcap < need || exact) newcap = need;
    void* nrows = nullptr;
    float* nnorms = nullptr;
    size_t row_bytes = (size_t)ix->ld * ix->esize;
    const size_t slack = 8 * row_bytes + 256;                    // 4-D TMA boxes read whole 8-row groups: keep the last group mapped
    cudaError_t e = cudaMalloc(&nrows, newcap * row_bytes + slack);
    if (e != cudaSuccess && newcap > need) {   // doubling did not fit: fall back to the exact size
        cudaGetLastError();
        newcap = need;
        e = cudaMalloc(&nrows, newcap * row_bytes + slack);
    }
    if (e != cudaSuccess) { cudaGetLastError(); return fail(CGVEC_ERR_OOM, "cudaMalloc of %zu bytes for %llu rows failed: %s", newcap * row_bytes, (unsigned long long)newcap, cudaGetErrorString(e)); }
    e = cudaMalloc(reinterpret_cast<void**>(&nnorms), (newcap + 64) * sizeof(float));
    if (e != cudaSuccess) { cudaGetLastError(); cudaFree(nrows); return fail(CGVEC_ERR_OOM, "cudaMalloc for norms failed: %s", cudaGetErrorString(e)); }
    e = cudaMemsetAsync(nnorms, 0, (newcap + 64) * sizeof(float), ix->main_stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(static_cast<uint8_t*>(nrows) + newcap * row_bytes, 0, slack, ix->main_stream);
    if (e == cudaSuccess && ix->n) {
        e = cudaMemcpyAsync(nrows, ix->d_rows, ix->n * row_bytes, cudaMemcpyDeviceToDevice, ix->main_stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nnorms, ix->d_norms, ix->n * sizeof(float), cudaMemcpyDeviceToDevice, ix->main_stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->main_stream);
    if (e != cudaSuccess) {                                       // the old matrix stays in place; nothing leaks
        cudaGetLastError(); cudaFree(nrows); cudaFree(nnorms);
        return fail(CGVEC_ERR_CUDA, "growing the matrix failed: %s", cudaGetErrorString(e));
    }
    cudaFree(ix->d_rows);
    cudaFree(ix->d_norms);
    ix->d_rows = nrows;
    ix->d_norms = nnorms;
    ix->cap = newcap;
    return CGVEC_OK;
}

uint64_t* trace_slot(Index* ix, uint32_t kind) {
    if (!ix->d_trace || ix->trace_kinds.size() >= Index::kTraceCap) return nullptr;
    ix->trace_kinds.push_back(kind);
    return ix->d_trace + 2 * (ix->trace_kinds.size() - 1);
}

ScanParams map_params(const Index* ix) {
    ScanParams p{};
    p.row_offset = ix->row_offset;
    p.blk_rows = ix->blk_rows;
    p.n_shards = ix->n_shards;
    p.shard_id = ix->shard_id;
    return p;
}

int launch_norms(Index* ix, uint64_t first, uint64_t count, cudaStream_t st) {
    if (!count) return CGVEC_OK;
    const int threads = 256;
    uint64_t blocks = (count * 8 + threads - 1) / threads;
    if (ix->dtype == CGVEC_F32)
        row_sqnorm_kernel<float><<<(unsigned)blocks, threads, 0, st>>>(static_cast<const float*>(ix->d_rows), first, count, ix->dim, ix->ld, ix->d_norms);
    else
        row_sqnorm_kernel<__half><<<(unsigned)blocks, threads, 0, st>>>(static_cast<const __half*>(ix->d_rows), first, count, ix->dim, ix->ld, ix->d_norms);
    ix->launches++;
    CUDA_TRY(cudaGetLastError());
    return CGVEC_OK;
}

// Every decision that changes the sequence of collectives (kernel family -> transport, empty-shard errors) must be the same
// on all ranks of a sharded index, so it is taken on the smallest / largest shard size, agreed with one NCCL all-gather the
// first time a search follows a change of the rows (all ranks call search collectively, so they all get here together).
int agree_shard_sizes(Index* ix) {
    if (ix->world == 1) { ix->agreed_min_n = ix->agreed_max_n = ix->n; ix->agreed_valid = true; return CGVEC_OK; }
    std::lock_guard<std::mutex> lk(ix->comm_mu);
    if (ix->agreed_valid) return CGVEC_OK;
    uint64_t* d_all = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_all), sizeof(uint64_t) * ix->world));
    const uint64_t mine = ix->n;
    cudaError_t e = cudaMemcpyAsync(d_all + ix->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, ix->main_stream);
    int nr = kNcclSuccess;
    if (e == cudaSuccess) nr = nccl_api().AllGather(d_all + ix->rank, d_all, 1, kNcclUint64, ix->comm, ix->main_stream);
    std::vector<uint64_t> all(ix->world, 0);
    if (e == cudaSuccess && nr == kNcclSuccess) e = cudaMemcpyAsync(all.data(), d_all, sizeof(uint64_t) * ix->world, cudaMemcpyDeviceToHost, ix->main_stream);
    if (e == cudaSuccess && nr == kNcclSuccess) e = cudaStreamSynchronize(ix->main_stream);
    cudaFree(d_all);
    if (nr != kNcclSuccess) return fail(CGVEC_ERR_NCCL, "ncclAllGather (shard sizes) failed: %s", nccl_api().GetErrorString(nr));
    if (e != cudaSuccess) return fail(CGVEC_ERR_CUDA, "shard size agreement failed: %s", cudaGetErrorString(e));
    ix->agreed_min_n = *std::min_element(all.begin(), all.end());
    ix->agreed_max_n = *std::max_element(all.begin(), all.end());
    ix->agreed_valid = true;
    return CGVEC_OK;
}

int search_formula(Index* ix, SearchCtx* c, const float* d_q, uint32_t k, int formula, cudaStream_t st, uint64_t* h_rows, float* h_scores,
                   uint32_t* h_count);

#include "host_scan.inl"
#include "host_tensor.inl"
}  // namespace

struct cgvec_index : Index {};

namespace {
#include "multi_device.inl"
}  // namespace

// ================================================================================================
// lifecycle
// ================================================================================================
static int create_common(uint32_t dim, cgvec_dtype storage, int device, int rank, int world, const void* uid,
                         uint64_t row_offset, cgvec_index** out) {
    if (!out) return fail(CGVEC_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (dim == 0) return fail(CGVEC_ERR_BAD_DIM, "dimension must be > 0");
    if (storage != CGVEC_F32 && storage != CGVEC_F16) return fail(CGVEC_ERR_BAD_ARG, "unknown storage dtype %d", (int)storage);
    if (world < 1 || rank < 0 || rank >= world) return fail(CGVEC_ERR_BAD_ARG, "bad rank %d / world %d", rank, world);
    int rc = check_device(device);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    std::unique_ptr<cgvec_index> ix(new cgvec_index());
    ix->dim = dim;
    ix->dtype = storage;
    ix->esize = storage == CGVEC_F32 ? 4 : 2;
    const uint32_t align_elems = 16 / ix->esize;                 // rows start on 16-byte boundaries (bulk copies)
    ix->ld = (dim + align_elems - 1) / align_elems * align_elems;
    ix->device = device;
    ix->rank = rank;
    ix->world = world;
    ix->row_offset = row_offset;
    CUDA_TRY(cudaDeviceGetAttribute(&ix->sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_TRY(cudaStreamCreateWithFlags(&ix->main_stream, cudaStreamNonBlocking));
    if (world > 1) {
        if (!uid) return fail(CGVEC_ERR_BAD_ARG, "world > 1 needs the NCCL unique id from cgvec_nccl_unique_id()");
        if (!nccl_api().load()) return fail(CGVEC_ERR_NCCL, "%s", nccl_api().load_error);
        NcclUniqueId id;
        memcpy(&id, uid, sizeof(id));
        NCCL_TRY(nccl_api().CommInitRank(&ix->comm, world, id, rank));
        CUDA_TRY(cudaHostAlloc(reinterpret_cast<void**>(&ix->h_xerr), sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable));
        *ix->h_xerr = 0;
        setup_peer_exchange(ix.get());            // best effort: NCCL remains the transport if peer mapping fails
    }
    *out = ix.release();
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_create(uint32_t dim, cgvec_dtype storage, const int* device_ids, int n_devices, cgvec_index** out) {
    if (n_devices < 1) return fail(CGVEC_ERR_BAD_ARG, "n_devices must be >= 1");
    if (n_devices > 1) {
        if (!out) return fail(CGVEC_ERR_BAD_ARG, "out is NULL");
        *out = nullptr;
        if (dim == 0) return fail(CGVEC_ERR_BAD_DIM, "dimension must be > 0");
        if (storage != CGVEC_F32 && storage != CGVEC_F16) return fail(CGVEC_ERR_BAD_ARG, "unknown storage dtype %d", (int)storage);
        Index* mx = nullptr;
        int rc = multi_create(dim, storage, device_ids, n_devices, &mx);
        if (rc) return rc;
        *out = static_cast<cgvec_index*>(mx);
        return CGVEC_OK;
    }
    return create_common(dim, storage, device_ids ? device_ids[0] : 0, 0, 1, nullptr, 0, out);
}

// Deployment switch: the reference's PerformanceConfig.enable_gpu (codegraph-core/src/config_manager.rs:362-364,
// default false :419) decides whether the host wires this library in at all; CODEGRAPH_ENABLE_GPU overrides it the way the
// reference's other CODEGRAPH_* variables override their config fields (config_manager.rs:696-811), and CODEGRAPH_B200_DEVICES
// picks the GPUs: "all", a count ("4" -> devices 0..3) or an explicit list ("0,2,5").  Unset -> device 0.
static bool env_truthy(const char* v, bool dflt) {
    if (!v || !*v) return dflt;
    std::string t(v);
    for (auto& ch : t) ch = (char)tolower((unsigned char)ch);
    if (t == "1" || t == "true" || t == "yes" || t == "on") return true;
    if (t == "0" || t == "false" || t == "no" || t == "off") return false;
    return dflt;
}
CGVEC_EXPORT int cgvec_create_from_env(uint32_t dim, cgvec_dtype storage, int enable_gpu, cgvec_index** out) {
    if (!out) return fail(CGVEC_ERR_BAD_ARG, "out is NULL");
    *out = nullptr;
    if (!env_truthy(getenv("CODEGRAPH_ENABLE_GPU"), enable_gpu != 0))
        return fail(CGVEC_ERR_DISABLED, "GPU acceleration is switched off (performance.enable_gpu / CODEGRAPH_ENABLE_GPU)");
    const char* spec = getenv("CODEGRAPH_B200_DEVICES");
    std::vector<int> devs;
    if (!spec || !*spec) devs.push_back(0);
    else {
        std::string t(spec);
        for (auto& ch : t) ch = (char)tolower((unsigned char)ch);
        if (t == "all") {
            int n = 0;
            if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) { cudaGetLastError(); return fail(CGVEC_ERR_NO_DEVICE, "CODEGRAPH_B200_DEVICES=all but no CUDA device is visible"); }
            if (n > (int)kXchgMaxWorld) n = (int)kXchgMaxWorld;
            for (int i = 0; i < n; ++i) devs.push_back(i);
        } else {
            std::vector<int> vals;
            size_t pos = 0;
            bool list = t.find(',') != std::string::npos;
            while (pos <= t.size()) {
                size_t e = t.find(',', pos);
                if (e == std::string::npos) e = t.size();
                std::string tok = t.substr(pos, e - pos);
                size_t a = tok.find_first_not_of(" \t"), b = tok.find_last_not_of(" \t");
                if (a == std::string::npos) return fail(CGVEC_ERR_BAD_ARG, "CODEGRAPH_B200_DEVICES='%s': empty entry", spec);
                tok = tok.substr(a, b - a + 1);
                if (tok.find_first_not_of("0123456789") != std::string::npos || tok.size() > 4)
                    return fail(CGVEC_ERR_BAD_ARG, "CODEGRAPH_B200_DEVICES='%s': '%s' is not a device number", spec, tok.c_str());
                vals.push_back(atoi(tok.c_str()));
                pos = e + 1;
            }
            if (!list) {                                         // a single number is a COUNT of devices (0..n-1)
                if (vals[0] < 1) return fail(CGVEC_ERR_BAD_ARG, "CODEGRAPH_B200_DEVICES='%s': device count must be >= 1", spec);
                for (int i = 0; i < vals[0]; ++i) devs.push_back(i);
            } else devs = vals;
        }
    }
    return cgvec_create(dim, storage, devs.data(), (int)devs.size(), out);
}

CGVEC_EXPORT int cgvec_create_rank(uint32_t dim, cgvec_dtype storage, int device, int rank, int world, const void* nccl_unique_id,
                                   uint64_t row_offset, cgvec_index** out) {
    return create_common(dim, storage, device, rank, world, nccl_unique_id, row_offset, out);
}

CGVEC_EXPORT int cgvec_nccl_unique_id(void* out_128_bytes) {
    if (!out_128_bytes) return fail(CGVEC_ERR_BAD_ARG, "out is NULL");
    if (!nccl_api().load()) return fail(CGVEC_ERR_NCCL, "%s", nccl_api().load_error);
    NcclUniqueId id;
    NCCL_TRY(nccl_api().GetUniqueId(&id));
    memcpy(out_128_bytes, &id, sizeof(id));
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_destroy(cgvec_index* ix) {
    if (!ix) return CGVEC_OK;
    if (ix->open_streams.load() > 0)                             // they hold pointers into this index (and maybe a resident kernel)
        return fail(CGVEC_ERR_UNSUPPORTED, "close the index's cgvec_stream / cgvec_serve sessions before destroying it");
    if (!ix->parts.empty()) { multi_destroy(ix); return CGVEC_OK; }
    cudaSetDevice(ix->device);
    cudaDeviceSynchronize();
    drain_timings(ix);
    for (auto* c : ix->pool) ctx_free(c);
    for (int r = 0; r < ix->world && r < (int)kXchgMaxWorld; ++r)
        if (ix->xpeer[r] && r != ix->rank) cudaIpcCloseMemHandle(ix->xpeer[r]);
    cudaFree(ix->xbuf);
    cudaFreeHost(ix->h_xerr);
    if (ix->comm) nccl_api().CommDestroy(ix->comm);
    cudaFree(ix->d_rows);
    cudaFree(ix->d_norms);
    cudaFree(ix->d_codes);
    cudaFree(ix->d_inorms);
    cudaFree(ix->d_trace);
    if (ix->main_stream) cudaStreamDestroy(ix->main_stream);
    delete ix;
    return CGVEC_OK;
}

// ================================================================================================
// write side
// ================================================================================================
CGVEC_EXPORT int cgvec_reserve(cgvec_index* ix, uint64_t n_rows) {
    CGVEC_WRITE_GUARD(ix);
    if (!ix) return fail(CGVEC_ERR_BAD_ARG, "index is NULL");
    if (!ix->parts.empty()) return multi_reserve(ix, n_rows);
    CUDA_TRY(cudaSetDevice(ix->device));
    return grow(ix, n_rows, /*exact=*/true);
}

static int add_impl(cgvec_index* ix, const uint8_t (*ids)[16], const void* rows, uint64_t n, uint32_t src_esize) {
    NvtxRange nvtx_("cgvec.add");
    CGVEC_WRITE_GUARD(ix);
    if (!ix) return fail(CGVEC_ERR_BAD_ARG, "index is NULL");
    if (n == 0) return CGVEC_OK;
    if (!rows) return fail(CGVEC_ERR_BAD_ARG, "rows is NULL");
    if (!ix->parts.empty()) return multi_add(ix, ids, rows, n, src_esize);
    if (src_esize != ix->esize)
        return fail(CGVEC_ERR_BAD_ARG, "index stores %s rows; use %s", ix->esize == 4 ? "f32" : "f16", ix->esize == 4 ? "cgvec_add" : "cgvec_add_f16");
    CUDA_TRY(cudaSetDevice(ix->device));
    // classify: append vs overwrite (InMemoryVectorStore insert semantics).  New id -> row entries are STAGED and only
    // committed to id2row once every copy has succeeded: a failed add (row limit, out of memory, copy error) must not leave
    // ids that point at rows which were never stored.
    std::vector<uint64_t> target(n);
    std::unordered_map<IdKey, uint64_t, IdHash> staged;
    uint64_t next = ix->n;
    for (uint64_t i = 0; i < n; ++i) {
        if (ids) {
            IdKey key = id_key(ids[i]);
            auto it = ix->id2row.find(key);
            if (it != ix->id2row.end()) { target[i] = it->second; continue; }
            auto st = staged.find(key);
            if (st != staged.end()) { target[i] = st->second; continue; }      // same new id twice in one call: last row wins
            staged.emplace(key, next);
        }
        target[i] = next++;
    }
    if (ix->row_offset + next > 0xfffffffeull) return fail(CGVEC_ERR_UNSUPPORTED, "more than 2^32-2 rows are not supported");
    int rc = grow(ix, next);
    if (rc) return rc;
    ix->ids.resize(next * 16, 0);
    ix->has_id.resize(next, 0);
    const size_t src_pitch = (size_t)ix->dim * ix->esize, dst_pitch = (size_t)ix->ld * ix->esize;
    const uint8_t* src = static_cast<const uint8_t*>(rows);
    uint8_t* dst = static_cast<uint8_t*>(ix->d_rows);
    uint64_t i = 0;
    while (i < n) {                     // copy maximal runs with consecutive targets in one 2D copy
        uint64_t j = i + 1;
        while (j < n && target[j] == target[j - 1] + 1) ++j;
        if (dst_pitch != src_pitch) CUDA_TRY(cudaMemset2DAsync(dst + target[i] * dst_pitch, dst_pitch, 0, dst_pitch, j - i, ix->main_stream));
        CUDA_TRY(cudaMemcpy2DAsync(dst + target[i] * dst_pitch, dst_pitch, src + i * src_pitch, src_pitch, src_pitch, j - i,
                                   cudaMemcpyHostToDevice, ix->main_stream));
        rc = launch_norms(ix, target[i], j - i, ix->main_stream);
        if (rc) return rc;
        if (ids) for (uint64_t r = i; r < j; ++r) { memcpy(&ix->ids[target[r] * 16], ids[r], 16); ix->has_id[target[r]] = 1; }
        i = j;
    }
    CUDA_TRY(cudaStreamSynchronize(ix->main_stream));
    for (auto& kv : staged) ix->id2row.emplace(kv.first, kv.second);
    ix->n = next;
    ix->agreed_valid = false;
    ix->n_codes = 0;                                            // int8 codes are stale after any write: cgvec_quantize_i8 again
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_add(cgvec_index* ix, const uint8_t (*ids)[16], const float* rows_f32, uint64_t n) {
    return add_impl(ix, ids, rows_f32, n, 4);
}
CGVEC_EXPORT int cgvec_add_f16(cgvec_index* ix, const uint8_t (*ids)[16], const uint16_t* rows_f16, uint64_t n) {
    return add_impl(ix, ids, rows_f16, n, 2);
}

CGVEC_EXPORT int cgvec_normalize_rows(cgvec_index* ix) {
    CGVEC_WRITE_GUARD(ix);
    if (!ix) return fail(CGVEC_ERR_BAD_ARG, "index is NULL");
    if (!ix->parts.empty()) {
        for (Index* p : ix->parts) { int rc = cgvec_normalize_rows(static_cast<cgvec_index*>(p)); if (rc) return rc; }
        return CGVEC_OK;
    }
    if (!ix->n) return CGVEC_OK;
    CUDA_TRY(cudaSetDevice(ix->device));
    const int threads = 256;
    uint64_t blocks = (ix->n * 8 + threads - 1) / threads;
    if (ix->dtype == CGVEC_F32) normalize_rows_kernel<float><<<(unsigned)blocks, threads, 0, ix->main_stream>>>(static_cast<float*>(ix->d_rows), ix->n, ix->dim, ix->ld);
    else normalize_rows_kernel<__half><<<(unsigned)blocks, threads, 0, ix->main_stream>>>(static_cast<__half*>(ix->d_rows), ix->n, ix->dim, ix->ld);
    ix->launches++;
    CUDA_TRY(cudaGetLastError());
    int rc = launch_norms(ix, 0, ix->n, ix->main_stream);
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(ix->main_stream));
    ix->n_codes = 0;
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_fill_synthetic(cgvec_index* ix, uint64_t n, uint64_t seed, int unit_norm) {
    CGVEC_WRITE_GUARD(ix);
    if (!ix) return fail(CGVEC_ERR_BAD_ARG, "index is NULL");
    if (!n) return CGVEC_OK;
    if (!ix->parts.empty()) return multi_fill_synthetic(ix, n, seed, unit_norm);
    CUDA_TRY(cudaSetDevice(ix->device));
    uint64_t next = ix->n + n;
    if (ix->row_offset + next > 0xfffffffeull) return fail(CGVEC_ERR_UNSUPPORTED, "more than 2^32-2 rows are not supported");
    int rc = grow(ix, next);
    if (rc) return rc;
    ix->ids.resize(next * 16, 0);
    ix->has_id.resize(next, 0);
    const int threads = 256;
    ScanParams map = map_params(ix);
    const uint64_t chunk = 1ull << 22;          // keep each grid < 2^31 blocks
    for (uint64_t done = 0; done < n; done += chunk) {
        uint64_t cnt = n - done < chunk ? n - done : chunk;
        uint64_t blocks = (cnt * 8 + threads - 1) / threads;
        if (ix->dtype == CGVEC_F32)
            synth_rows_kernel<float><<<(unsigned)blocks, threads, 0, ix->main_stream>>>(static_cast<float*>(ix->d_rows), ix->n + done, cnt, ix->dim, ix->ld, seed, unit_norm, map);
        else
            synth_rows_kernel<__half><<<(unsigned)blocks, threads, 0, ix->main_stream>>>(static_cast<__half*>(ix->d_rows), ix->n + done, cnt, ix->dim, ix->ld, seed, unit_norm, map);
        ix->launches++;
        CUDA_TRY(cudaGetLastError());
        rc = launch_norms(ix, ix->n + done, cnt, ix->main_stream);
        if (rc) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(ix->main_stream));
    ix->n = next;
    ix->agreed_valid = false;
    ix->n_codes = 0;
    return CGVEC_OK;
}

// ================================================================================================
// read side
// ================================================================================================
CGVEC_EXPORT uint64_t cgvec_len(const cgvec_index* ix) { return ix ? ix->n : 0; }
CGVEC_EXPORT uint32_t cgvec_dim(const cgvec_index* ix) { return ix ? ix->dim : 0; }

// formula != SIMD: over-fetch with the exact SIMD-order scan, re-score the candidates in the requested
// formula on the device, re-rank, and PROVE no outside row can enter the top-k (else widen and retry).

// The search proper (arguments validated by cgvec_search_ex): multi-device parent, or one device / one rank of a sharded index.
static int search_ex_body(Index* ix, const float* queries, uint32_t nq, uint32_t k, const cgvec_search_opts& o, uint64_t* out_rows,
                          uint8_t (*out_ids)[16], float* out_scores, uint32_t* out_counts) {
    NvtxRange nvtx_("cgvec.search");
    if (!ix->parts.empty()) return multi_search(ix, queries, nq, k, o, out_rows, out_ids, out_scores, out_counts);
    CUDA_TRY(cudaSetDevice(ix->device));
    const uint64_t n_total_hint = ix->n;                       // local rows; sharded ranks may be empty individually
    if (ix->world == 1 && n_total_hint == 0) {
        if (o.device_io) { if (out_counts) CUDA_TRY(cudaMemsetAsync(out_counts, 0, nq * sizeof(uint32_t), (cudaStream_t)o.stream)); }
        else if (out_counts) for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
        return CGVEC_OK;
    }
    if (k > kMaxK) return fail(CGVEC_ERR_UNSUPPORTED, "k = %u exceeds the fused top-k limit of %u", k, kMaxK);
    if (ix->world > 1) {
        if (ix->h_xerr && *ix->h_xerr) return fail(CGVEC_ERR_NCCL, "an earlier peer exchange timed out waiting for rank %u; the sharded index is out of step and must be rebuilt", *ix->h_xerr - 1);
        if (!ix->agreed_valid) { int arc = agree_shard_sizes(ix); if (arc) return arc; }
        if (ix->agreed_min_n == 0) return fail(CGVEC_ERR_UNSUPPORTED, "a rank of a sharded index holds no rows (every rank fails this call alike)");
    }

    SearchCtx* c = nullptr;
    int rc = ctx_acquire(ix, &c);
    if (rc) return rc;
    cudaStream_t st = o.stream ? (cudaStream_t)o.stream : c->stream;
    cudaStreamWaitEvent(st, c->done, 0);                       // scratch reuse: the last user of this ctx may have been an asynchronous (device_io) call on another stream
    const uint32_t qstride = (ix->dim + 3) & ~3u;
    ix->searches++;

    auto finish = [&](int code) { cudaEventRecord(c->done, st); ctx_release(ix, c); return code; };

    if (o.device_io) {
        if (qstride != ix->dim) return finish(fail(CGVEC_ERR_UNSUPPORTED, "device_io needs dim %% 4 == 0"));
        rc = run_queries(ix, c, queries, qstride, nq, k, o.metric, o.path, st, out_rows, out_scores, out_counts);
        if (rc) return finish(rc);
        return finish(CGVEC_OK);
    }

    // host I/O: stage queries through pinned memory, run, read results back
    rc = ensure(&c->h_q, &c->hq_cap, (size_t)nq * qstride, true); if (rc) return finish(rc);
    rc = ensure(&c->d_q, &c->q_cap, (size_t)nq * qstride); if (rc) return finish(rc);
    for (uint32_t q = 0; q < nq; ++q) {
        memcpy(c->h_q + (size_t)q * qstride, queries + (size_t)q * ix->dim, ix->dim * sizeof(float));
        for (uint32_t i = ix->dim; i < qstride; ++i) c->h_q[(size_t)q * qstride + i] = 0.0f;
    }
    {
        cudaError_t e = cudaMemcpyAsync(c->d_q, c->h_q, (size_t)nq * qstride * sizeof(float), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return finish(fail(CGVEC_ERR_CUDA, "query upload failed: %s", cudaGetErrorString(e)));
    }
    {
        size_t oc = c->out_cap, oc2 = c->out_cap;
        rc = ensure(&c->d_rows, &oc, (size_t)nq * k); if (rc) return finish(rc);
        rc = ensure(&c->d_scores, &oc2, (size_t)nq * k); if (rc) return finish(rc);
        c->out_cap = oc < oc2 ? oc : oc2;
        rc = ensure(&c->d_counts, &c->cnt_cap, nq); if (rc) return finish(rc);
        size_t hc = c->hout_cap, hc2 = c->hout_cap;
        rc = ensure(&c->h_rows, &hc, (size_t)nq * k, true); if (rc) return finish(rc);
        rc = ensure(&c->h_scores, &hc2, (size_t)nq * k, true); if (rc) return finish(rc);
        c->hout_cap = hc < hc2 ? hc : hc2;
        rc = ensure(&c->h_counts, &c->hcnt_cap, nq, true); if (rc) return finish(rc);
    }
    if (o.formula == CGVEC_FORMULA_SIMD) {
        // Results are decoded straight into the pinned host mirrors (device-accessible under UVA): ~130 bytes of
        // posted PCIe writes from the last kernel replace three device-to-host copies per call.
        rc = run_queries(ix, c, c->d_q, qstride, nq, k, o.metric, o.path, st, c->h_rows, c->h_scores, c->h_counts);
        if (rc) return finish(rc);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) return finish(fail(CGVEC_ERR_CUDA, "search failed on the device: %s", cudaGetErrorString(e)));
        if (ix->h_xerr && *ix->h_xerr) return finish(fail(CGVEC_ERR_NCCL, "peer exchange timed out after %d ms waiting for rank %u", ix->opt_xchg_timeout_ms, *ix->h_xerr - 1));
    } else {
        if (ix->world > 1 || ix->row_offset != 0) return finish(fail(CGVEC_ERR_UNSUPPORTED, "non-SIMD formulas are not available on sharded indexes yet"));
        // A batch under a sequential cosine (the indexer's symbol resolver: U references x S symbols, indexer.rs:2827-2843) is a
        // dense contraction like any other: ONE tensor-core pass orders the rows, the survivors are re-scored in the formula's own
        // order, and the proof's bound grows by |formula - simd|.  Small batches and the distance form keep the per-query path.
        const bool batched = o.formula != CGVEC_FORMULA_BASELINE && o.path != CGVEC_PATH_EXACT && tensor_auto_ok(ix, o.metric, nq, k) &&
                             tc_batch_limit(ix, nq) > 0;
        if (batched) {
            rc = run_queries(ix, c, c->d_q, qstride, nq, k, o.metric, CGVEC_PATH_TENSOR, st, c->h_rows, c->h_scores, c->h_counts, o.formula);
            if (rc) return finish(rc);
            cudaError_t e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return finish(fail(CGVEC_ERR_CUDA, "search failed on the device: %s", cudaGetErrorString(e)));
        }
        for (uint32_t q = 0; q < nq && !batched; ++q) {
            rc = search_formula(ix, c, c->d_q + (size_t)q * qstride, k, o.formula, st, c->h_rows + (size_t)q * k, c->h_scores + (size_t)q * k,
                                c->h_counts + q);
            if (rc) return finish(rc);
        }
    }
    for (uint32_t q = 0; q < nq; ++q) {
        uint32_t cnt = c->h_counts[q];
        if (out_counts) out_counts[q] = cnt;
        for (uint32_t i = 0; i < k; ++i) {
            size_t o_i = (size_t)q * k + i;
            bool valid = i < cnt;
            uint64_t grow = valid ? c->h_rows[o_i] : ~0ull;
            if (out_rows) out_rows[o_i] = grow;
            if (out_scores) out_scores[o_i] = valid ? c->h_scores[o_i] : 0.0f;
            if (out_ids) {
                memset(out_ids[o_i], 0, 16);
                if (valid && grow >= ix->row_offset && grow - ix->row_offset < ix->n && ix->has_id[grow - ix->row_offset])
                    memcpy(out_ids[o_i], &ix->ids[(grow - ix->row_offset) * 16], 16);
            }
        }
    }
    return finish(CGVEC_OK);
}

// Group commit of concurrent batch-1 callers (SURVEY §8b "Threading"; multi_vector_search, search.rs:347-361, fans out one
// search_similar per query with try_join_all).  Every batch-1 scan is a full pass over the matrix, so callers that arrive
// while a pass is in flight are not given passes of their own: the first caller to find no pass in flight becomes the
// leader, takes every queued request compatible with the oldest one (same k, metric, path) up to `coalesce_max`, serves them as ONE
// multi-query call (<= 4 queries per exact-order launch, the tensor path for larger groups; bit-exact either way) and hands
// the results out.  A lone caller pays one uncontended mutex; nobody waits for a batch to fill.
struct Index::PendingSearch {
    const float* query; uint32_t k; cgvec_search_opts o;
    uint64_t* out_rows; uint8_t (*out_ids)[16]; float* out_scores; uint32_t* out_count;
    int rc = CGVEC_OK; bool done = false; std::string err;
};
static int coalesced_search(Index* ix, const float* query, uint32_t k, const cgvec_search_opts& o, uint64_t* out_rows,
                            uint8_t (*out_ids)[16], float* out_scores, uint32_t* out_counts) {
    Index::PendingSearch me;
    me.query = query; me.k = k; me.o = o; me.out_rows = out_rows; me.out_ids = out_ids; me.out_scores = out_scores; me.out_count = out_counts;
    std::unique_lock<std::mutex> lk(ix->co_mu);
    ix->co_queue.push_back(&me);
    while (true) {
        ix->co_cv.wait(lk, [&] { return me.done || !ix->co_leader; });
        if (me.done) break;
        ix->co_leader = true;                                    // lead one batch: the oldest request and everything compatible with it
        std::vector<Index::PendingSearch*> batch;
        const Index::PendingSearch* head = ix->co_queue.front();
        const uint32_t cap = (uint32_t)std::max(1, ix->opt_coalesce_max);
        for (auto it = ix->co_queue.begin(); it != ix->co_queue.end() && batch.size() < cap;) {
            Index::PendingSearch* r = *it;
            if (r->k == head->k && r->o.metric == head->o.metric && r->o.path == head->o.path) { batch.push_back(r); it = ix->co_queue.erase(it); }
            else ++it;
        }
        lk.unlock();
        const uint32_t b = (uint32_t)batch.size(), kk = batch[0]->k;
        NvtxRange nvtx_(b > 1 ? "cgvec.coalesced_batch" : "cgvec.single_caller");
        int rc;
        if (b == 1) {
            Index::PendingSearch* r = batch[0];
            rc = search_ex_body(ix, r->query, 1, kk, r->o, r->out_rows, r->out_ids, r->out_scores, r->out_count);
            r->rc = rc; if (rc) r->err = g_err;
        } else {
            std::vector<float> q((size_t)b * ix->dim);
            std::vector<uint64_t> rows((size_t)b * kk);
            std::vector<float> scores((size_t)b * kk);
            std::vector<uint32_t> counts(b);
            std::vector<uint8_t> ids;
            bool want_ids = false;
            for (uint32_t i = 0; i < b; ++i) { memcpy(&q[(size_t)i * ix->dim], batch[i]->query, ix->dim * sizeof(float)); want_ids = want_ids || batch[i]->out_ids; }
            if (want_ids) ids.resize((size_t)b * kk * 16);
            rc = search_ex_body(ix, q.data(), b, kk, batch[0]->o, rows.data(), want_ids ? reinterpret_cast<uint8_t(*)[16]>(ids.data()) : nullptr,
                                scores.data(), counts.data());
            for (uint32_t i = 0; i < b; ++i) {
                Index::PendingSearch* r = batch[i];
                r->rc = rc;
                if (rc) { r->err = g_err; continue; }
                if (r->out_rows) memcpy(r->out_rows, &rows[(size_t)i * kk], kk * sizeof(uint64_t));
                if (r->out_scores) memcpy(r->out_scores, &scores[(size_t)i * kk], kk * sizeof(float));
                if (r->out_ids) memcpy(r->out_ids, &ids[(size_t)i * kk * 16], (size_t)kk * 16);
                if (r->out_count) *r->out_count = counts[i];
            }
            ix->co_batches++; ix->co_queries += b;
        }
        lk.lock();
        for (auto* r : batch) r->done = true;
        ix->co_leader = false;
        ix->co_cv.notify_all();
        if (me.done) break;                                      // else: my request was not compatible with that batch's head; go again
    }
    lk.unlock();
    if (me.rc) return fail(me.rc, "%s", me.err.c_str());
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_search_ex(const cgvec_index* cix, const float* queries, uint32_t nq, uint32_t k,
                                 const cgvec_search_opts* opts, uint64_t* out_rows, uint8_t (*out_ids)[16], float* out_scores,
                                 uint32_t* out_counts) {
    RwGuard rw_guard_(cix, false);
    Index* ix = const_cast<cgvec_index*>(cix);
    if (!ix) return fail(CGVEC_ERR_BAD_ARG, "index is NULL");
    cgvec_search_opts o{};
    o.struct_size = sizeof(o);
    if (opts) {
        if (opts->struct_size < sizeof(uint32_t) * 4) return fail(CGVEC_ERR_BAD_ARG, "opts->struct_size is not set");
        memcpy(&o, opts, opts->struct_size < sizeof(o) ? opts->struct_size : sizeof(o));
    }
    if (o.metric != CGVEC_COSINE && o.metric != CGVEC_DOT && o.metric != CGVEC_L2) return fail(CGVEC_ERR_BAD_ARG, "unknown metric %d", (int)o.metric);
    if (o.formula < CGVEC_FORMULA_SIMD || o.formula > CGVEC_FORMULA_BASELINE) return fail(CGVEC_ERR_BAD_ARG, "unknown formula %d", (int)o.formula);
    if (o.formula != CGVEC_FORMULA_SIMD && o.metric != CGVEC_COSINE)
        return fail(CGVEC_ERR_UNSUPPORTED, "formulas other than SIMD exist for cosine only (the reference has no scalar dot / L2)");
    if (o.device_io && out_ids) return fail(CGVEC_ERR_BAD_ARG, "out_ids cannot be produced with device_io");
    if (o.device_io && o.formula != CGVEC_FORMULA_SIMD) return fail(CGVEC_ERR_UNSUPPORTED, "device_io supports the SIMD formula only");
    if (nq == 0 || k == 0) {                                   // surreal_store.rs:62-64
        if (out_counts && !o.device_io) for (uint32_t q = 0; q < nq; ++q) out_counts[q] = 0;
        return CGVEC_OK;
    }
    if (!queries) return fail(CGVEC_ERR_BAD_ARG, "queries is NULL");
    if (nq == 1 && ix->opt_coalesce && !o.device_io && !o.stream && o.formula == CGVEC_FORMULA_SIMD && ix->world == 1)
        return coalesced_search(ix, queries, k, o, out_rows, out_ids, out_scores, out_counts);
    return search_ex_body(ix, queries, nq, k, o, out_rows, out_ids, out_scores, out_counts);
}

CGVEC_EXPORT int cgvec_search(const cgvec_index* ix, const float* queries, uint32_t nq, uint32_t k, cgvec_metric metric,
                              uint64_t* out_rows, uint8_t (*out_ids)[16], float* out_scores, uint32_t* out_counts) {
    cgvec_search_opts o{};
    o.struct_size = sizeof(o);
    o.metric = metric;
    o.formula = CGVEC_FORMULA_SIMD;
    o.path = CGVEC_PATH_AUTO;
    return cgvec_search_ex(ix, queries, nq, k, &o, out_rows, out_ids, out_scores, out_counts);
}

#include "stream_search.inl"
#include "serve.inl"

template <typename T>
static void launch_rescore_t(Index* ix, const float* d_q, const uint64_t* d_local_rows, uint32_t n, int metric, int formula,
                             float* d_out, cudaStream_t st) {
    const int threads = 128;
    unsigned blocks = (n * 8 + threads - 1) / threads;
    rescore_kernel<T><<<blocks, threads, 0, st>>>(static_cast<const T*>(ix->d_rows), ix->dim, ix->ld, d_q, d_local_rows, n, metric, formula, d_out);
    ix->launches++;
}

namespace {
int search_formula(Index* ix, SearchCtx* c, const float* d_q, uint32_t k, int formula, cudaStream_t st, uint64_t* h_rows,
                   float* h_scores, uint32_t* h_count) {
    NvtxRange nvtx_("cgvec.search_formula");
    const int ascending = (formula == CGVEC_FORMULA_BASELINE);
    const uint64_t n = ix->n;
    const uint32_t want = (uint32_t)(k < n ? k : n);
    // |formula(x) - simd(x)| <= eps for every row: both evaluate the same cosine with <= (d+8) roundings of
    // relative size 2^-24 each on the dot and on the norms (standard recursive-summation bound), so twice that.
    const float eps = 4.0f * (float)(ix->dim + 8) * 5.9604645e-8f;
    uint32_t kp = want + (want / 4 > 16 ? want / 4 : 16);
    std::vector<uint64_t> cand_rows;
    std::vector<float> cand_simd, cand_new;
    while (true) {
        if (kp > n) kp = (uint32_t)n;
        if (kp > kMaxK) kp = kMaxK;
        size_t oc = c->tmp_cap, oc2 = c->tmp_cap;
        int rc = ensure(&c->d_tmp_rows, &oc, (size_t)kMaxK); if (rc) return rc;
        rc = ensure(&c->d_tmp_scores, &oc2, (size_t)kMaxK * 2); if (rc) return rc;
        c->tmp_cap = oc < oc2 ? oc : oc2;
        size_t cc = c->cnt_cap;
        rc = ensure(&c->d_counts, &cc, 4); if (rc) return rc;
        c->cnt_cap = cc;
        rc = scan_batch(ix, c, d_q, 1, kp, CGVEC_COSINE, st, c->d_tmp_rows, c->d_tmp_scores, c->d_counts);
        if (rc) return rc;
        // global -> local rows happen to coincide on an unsharded index (row_offset == 0 checked by caller path)
        if (ix->dtype == CGVEC_F32) launch_rescore_t<float>(ix, d_q, c->d_tmp_rows, kp, CGVEC_COSINE, formula, c->d_tmp_scores + kMaxK, st);
        else launch_rescore_t<__half>(ix, d_q, c->d_tmp_rows, kp, CGVEC_COSINE, formula, c->d_tmp_scores + kMaxK, st);
        CUDA_TRY(cudaGetLastError());
        cand_rows.resize(kp); cand_simd.resize(kp); cand_new.resize(kp);
        CUDA_TRY(cudaMemcpyAsync(cand_rows.data(), c->d_tmp_rows, kp * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(cand_simd.data(), c->d_tmp_scores, kp * sizeof(float), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaMemcpyAsync(cand_new.data(), c->d_tmp_scores + kMaxK, kp * sizeof(float), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        // re-rank the (<= 1024) candidates under the contract: a handful of comparisons, not a scoring path
        std::vector<uint32_t> order(kp);
        for (uint32_t i = 0; i < kp; ++i) order[i] = i;
        auto better = [&](uint32_t a, uint32_t b) {
            float x = cand_new[a], y = cand_new[b];
            bool xn = std::isnan(x), yn = std::isnan(y);
            if (xn != yn) return yn;
            if (!xn && x != y) return ascending ? x < y : x > y;
            return cand_rows[a] < cand_rows[b];
        };
        std::sort(order.begin(), order.end(), better);
        bool proven = (kp >= n);
        if (!proven) {
            // rows outside the candidate set have simd score <= tau, hence formula score <= tau + eps
            // (distance >= 1 - tau - eps for BASELINE).  NaN candidates make the bound meaningless -> widen.
            float tau = cand_simd[kp - 1];
            float kth = cand_new[order[want - 1]];
            if (!std::isnan(tau) && !std::isnan(kth)) {
                if (!ascending) proven = kth > tau + eps;
                else proven = kth < (1.0f - tau) - eps;
            }
        }
        if (proven || kp >= kMaxK) {
            if (!proven) return fail(CGVEC_ERR_UNSUPPORTED, "could not separate the top-%u under formula %d within %u candidates", k, formula, kp);
            for (uint32_t i = 0; i < want; ++i) { h_rows[i] = cand_rows[order[i]]; h_scores[i] = cand_new[order[i]]; }
            *h_count = want;
            return CGVEC_OK;
        }
        kp *= 2;
    }
}

}  // namespace

CGVEC_EXPORT int cgvec_row_of_id(const cgvec_index* ix, const uint8_t id[16], uint64_t* out_local_row) {
    RwGuard rw_guard_(ix, false);
    if (!ix || !id) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    auto it = ix->id2row.find(id_key(id));
    if (it == ix->id2row.end()) return fail(CGVEC_ERR_NOT_FOUND, "id not present");
    if (out_local_row) *out_local_row = it->second;
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_get_row(const cgvec_index* cix, uint64_t local_row, float* out_row) {
    RwGuard rw_guard_(cix, false);
    Index* ix = const_cast<cgvec_index*>(cix);
    if (!ix || !out_row) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    if (local_row >= ix->n) return fail(CGVEC_ERR_NOT_FOUND, "row %llu out of range", (unsigned long long)local_row);
    if (!ix->parts.empty()) return multi_get_rows(ix, local_row, 1, out_row);
    CUDA_TRY(cudaSetDevice(ix->device));
    const uint8_t* src = static_cast<const uint8_t*>(ix->d_rows) + local_row * (size_t)ix->ld * ix->esize;
    if (ix->dtype == CGVEC_F32) {
        CUDA_TRY(cudaMemcpy(out_row, src, ix->dim * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
        SearchCtx* c = nullptr;
        int rc = ctx_acquire(ix, &c);
        if (rc) return rc;
        rc = ensure(&c->d_q, &c->q_cap, (size_t)ix->dim);
        if (!rc) {
            widen_row_kernel<<<(ix->dim + 127) / 128, 128, 0, c->stream>>>(reinterpret_cast<const __half*>(src), ix->dim, c->d_q);
            ix->launches++;
            cudaError_t e = cudaMemcpyAsync(out_row, c->d_q, ix->dim * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) rc = fail(CGVEC_ERR_CUDA, "row read-back failed: %s", cudaGetErrorString(e));
        }
        ctx_release(ix, c);
        return rc;
    }
    return CGVEC_OK;
}

// Bulk read-back of rows [first, first+n) widened to f32 (row-major n x dim): snapshot / parity checks.
CGVEC_EXPORT int cgvec_get_rows(const cgvec_index* cix, uint64_t first, uint64_t n, float* out) {
    RwGuard rw_guard_(cix, false);
    Index* ix = const_cast<cgvec_index*>(cix);
    if (!ix) return fail(CGVEC_ERR_BAD_ARG, "index is NULL");
    if (n == 0) return CGVEC_OK;
    if (!out) return fail(CGVEC_ERR_BAD_ARG, "out is NULL");
    if (!ix->parts.empty()) return multi_get_rows(ix, first, n, out);
    if (first + n > ix->n) return fail(CGVEC_ERR_NOT_FOUND, "rows [%llu, %llu) out of range", (unsigned long long)first, (unsigned long long)(first + n));
    CUDA_TRY(cudaSetDevice(ix->device));
    const size_t pitch = (size_t)ix->ld * ix->esize;
    const uint8_t* src = static_cast<const uint8_t*>(ix->d_rows) + first * pitch;
    if (ix->dtype == CGVEC_F32) {
        CUDA_TRY(cudaMemcpy2D(out, (size_t)ix->dim * 4, src, pitch, (size_t)ix->dim * 4, n, cudaMemcpyDeviceToHost));
        return CGVEC_OK;
    }
    // f16 storage: widen on the device (exact), then copy f32 rows out in chunks
    const uint64_t chunk = 1u << 16;
    float* d_tmp = nullptr;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&d_tmp), (size_t)(n < chunk ? n : chunk) * ix->dim * sizeof(float)));
    cudaError_t e = cudaSuccess;
    for (uint64_t r = 0; r < n && e == cudaSuccess; r += chunk) {
        const uint64_t m = n - r < chunk ? n - r : chunk;
        const uint64_t total = m * ix->dim;
        widen_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ix->main_stream>>>(reinterpret_cast<const __half*>(src + r * pitch), ix->dim, ix->ld, m, d_tmp);
        ix->launches++;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out + r * ix->dim, d_tmp, total * sizeof(float), cudaMemcpyDeviceToHost, ix->main_stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ix->main_stream);
    }
    cudaFree(d_tmp);
    if (e != cudaSuccess) return fail(CGVEC_ERR_CUDA, "row read-back failed: %s", cudaGetErrorString(e));
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_get(const cgvec_index* ix, const uint8_t id[16], float* out_row) {
    RwGuard rw_guard_(ix, false);
    uint64_t row = 0;
    int rc = cgvec_row_of_id(ix, id, &row);
    if (rc) return rc;
    return cgvec_get_row(ix, row, out_row);
}

CGVEC_EXPORT int cgvec_rescore(const cgvec_index* cix, const float* query, const uint64_t* local_rows, uint32_t n, cgvec_metric metric,
                               cgvec_formula formula, float* out_scores) {
    NvtxRange nvtx_("cgvec.rescore");
    RwGuard rw_guard_(cix, false);
    Index* ix = const_cast<cgvec_index*>(cix);
    if (!ix) return fail(CGVEC_ERR_BAD_ARG, "index is NULL");
    if (n == 0) return CGVEC_OK;
    if (!query || !local_rows || !out_scores) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    if (formula != CGVEC_FORMULA_SIMD && metric == CGVEC_L2) return fail(CGVEC_ERR_UNSUPPORTED, "L2 exists in SIMD form only (simd_ops.rs:105-143)");
    if (!ix->parts.empty()) return multi_rescore(ix, query, local_rows, n, metric, formula, out_scores);
    for (uint32_t i = 0; i < n; ++i)
        if (local_rows[i] >= ix->n) return fail(CGVEC_ERR_NOT_FOUND, "row %llu out of range", (unsigned long long)local_rows[i]);
    CUDA_TRY(cudaSetDevice(ix->device));
    SearchCtx* c = nullptr;
    int rc = ctx_acquire(ix, &c);
    if (rc) return rc;
    auto done = [&](int code) { ctx_release(ix, c); return code; };
    const uint32_t qstride = (ix->dim + 3) & ~3u;
    rc = ensure(&c->d_q, &c->q_cap, (size_t)qstride); if (rc) return done(rc);
    size_t oc = c->tmp_cap, oc2 = c->tmp_cap;
    size_t need = n > kMaxK ? n : kMaxK;
    rc = ensure(&c->d_tmp_rows, &oc, need); if (rc) return done(rc);
    rc = ensure(&c->d_tmp_scores, &oc2, need * 2); if (rc) return done(rc);
    c->tmp_cap = oc < oc2 ? oc : oc2;
    cudaError_t e = cudaMemcpyAsync(c->d_q, query, ix->dim * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->d_tmp_rows, local_rows, n * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return done(fail(CGVEC_ERR_CUDA, "H2D failed: %s", cudaGetErrorString(e)));
    if (ix->dtype == CGVEC_F32) launch_rescore_t<float>(ix, c->d_q, c->d_tmp_rows, n, metric, formula, c->d_tmp_scores, c->stream);
    else launch_rescore_t<__half>(ix, c->d_q, c->d_tmp_rows, n, metric, formula, c->d_tmp_scores, c->stream);
    e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_scores, c->d_tmp_scores, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return done(fail(CGVEC_ERR_CUDA, "rescore failed: %s", cudaGetErrorString(e)));
    return done(CGVEC_OK);
}

CGVEC_EXPORT int cgvec_distances_first(const cgvec_index* cix, const float* query, uint64_t limit, float* out, uint64_t* out_n) {
    RwGuard rw_guard_(cix, false);
    Index* ix = const_cast<cgvec_index*>(cix);
    if (!ix || !query) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    uint64_t m = limit < ix->n ? limit : ix->n;
    if (out_n) *out_n = m;
    if (m == 0) return CGVEC_OK;
    if (!out) return fail(CGVEC_ERR_BAD_ARG, "out is NULL");
    if (!ix->parts.empty()) {                       // optimization.rs:404-418 form == the BASELINE formula of the first m rows
        std::vector<uint64_t> rows(m);
        for (uint64_t i = 0; i < m; ++i) rows[i] = i;
        return multi_rescore(ix, query, rows.data(), (uint32_t)m, CGVEC_COSINE, CGVEC_FORMULA_BASELINE, out);
    }
    CUDA_TRY(cudaSetDevice(ix->device));
    SearchCtx* c = nullptr;
    int rc = ctx_acquire(ix, &c);
    if (rc) return rc;
    auto done = [&](int code) { ctx_release(ix, c); return code; };
    rc = ensure(&c->d_q, &c->q_cap, (size_t)ix->dim + 4); if (rc) return done(rc);
    float* d_out = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&d_out), m * sizeof(float));
    if (e != cudaSuccess) return done(fail(CGVEC_ERR_OOM, "cudaMalloc failed: %s", cudaGetErrorString(e)));
    e = cudaMemcpyAsync(c->d_q, query, ix->dim * sizeof(float), cudaMemcpyHostToDevice, c->stream);
    const int threads = 128;
    if (e == cudaSuccess) {
        if (ix->dtype == CGVEC_F32) distances_first_kernel<float><<<(unsigned)((m + threads - 1) / threads), threads, 0, c->stream>>>(static_cast<const float*>(ix->d_rows), ix->dim, ix->ld, c->d_q, m, d_out);
        else distances_first_kernel<__half><<<(unsigned)((m + threads - 1) / threads), threads, 0, c->stream>>>(static_cast<const __half*>(ix->d_rows), ix->dim, ix->ld, c->d_q, m, d_out);
        ix->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, m * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return done(fail(CGVEC_ERR_CUDA, "distances_first failed: %s", cudaGetErrorString(e)));
    return done(CGVEC_OK);
}

// ================================================================================================
// int8 quantised scan: quantize_batch (optimization.rs:212-224,268-274) + search_optimized (:63-150)
// ================================================================================================
CGVEC_EXPORT int cgvec_quantize_i8(cgvec_index* ix) {
    CGVEC_WRITE_GUARD(ix);
    if (!ix) return fail(CGVEC_ERR_BAD_ARG, "index is NULL");
    if (!ix->parts.empty() || ix->world > 1) return fail(CGVEC_ERR_UNSUPPORTED, "the int8 scan serves single-GPU indexes");
    CUDA_TRY(cudaSetDevice(ix->device));
    cudaFree(ix->d_codes); cudaFree(ix->d_inorms);
    ix->d_codes = nullptr; ix->d_inorms = nullptr; ix->n_codes = 0;
    { std::lock_guard<std::mutex> lk(ix->i8_mu); ix->i8_fill_end.clear(); }
    ix->ld8 = (ix->dim + 15) & ~15u;
    if (ix->n == 0) return CGVEC_OK;
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ix->d_codes), ix->n * (size_t)ix->ld8));
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&ix->d_inorms), (ix->n + 64) * sizeof(int32_t)));   // the scan's norm tiles read whole 32-row slots
    CUDA_TRY(cudaMemsetAsync(ix->d_inorms, 0, (ix->n + 64) * sizeof(int32_t), ix->main_stream));
    const int threads = 256;
    const uint64_t blocks = (ix->n * 32 + threads - 1) / threads;
    if (ix->dtype == CGVEC_F32)
        quantize_rows_i8_kernel<float><<<(unsigned)blocks, threads, 0, ix->main_stream>>>(static_cast<const float*>(ix->d_rows), ix->n, ix->dim, ix->ld, ix->ld8, ix->d_codes, ix->d_inorms);
    else
        quantize_rows_i8_kernel<__half><<<(unsigned)blocks, threads, 0, ix->main_stream>>>(static_cast<const __half*>(ix->d_rows), ix->n, ix->dim, ix->ld, ix->ld8, ix->d_codes, ix->d_inorms);
    ix->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(ix->main_stream));
    ix->n_codes = ix->n;
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_get_codes_i8(const cgvec_index* cix, uint64_t first, uint64_t n, uint8_t* out) {
    RwGuard rw_guard_(cix, false);
    Index* ix = const_cast<cgvec_index*>(cix);
    if (!ix || (!out && n)) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    if (first + n > ix->n_codes) return fail(CGVEC_ERR_NOT_FOUND, "codes [%llu, %llu) out of range (call cgvec_quantize_i8 first)", (unsigned long long)first, (unsigned long long)(first + n));
    if (!n) return CGVEC_OK;
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaMemcpy2D(out, ix->dim, ix->d_codes + first * ix->ld8, ix->ld8, ix->dim, n, cudaMemcpyDeviceToHost));
    return CGVEC_OK;
}

// One pass of scan_i8_kernel + merge: the best k keys (tie_mode 0: by score then lower row; tie_mode 1: the k lowest rows with
// score >= vstar) read back to the host.
static int i8_pass(Index* ix, SearchCtx* c, const int8_t* d_q8, const float* d_qn, const int32_t* d_qs, uint32_t k, uint32_t tie_mode, float vstar,
                   cudaStream_t st, std::vector<uint64_t>* rows, std::vector<float>* scores, uint32_t* cnt_out) {
    // The int8 codes go through the same TMA-staged, persistent scan kernel as the f32 / f16 matrix (scan_exact.cuh, METRIC_I8):
    // rows of ld8 bytes are bulk-copied into the shared-memory ring as ld8/4 words, 8 threads per row run dp4a over them.
    const uint32_t words = ix->ld8 / 4;
    ScanGeom g;
    int rc = plan_scan(ix, k, 1, &g, words, 4, words, ix->n_codes);
    if (rc) return rc;
    rc = ensure_parts(c, (size_t)g.grid * k); if (rc) return rc;
    {
        size_t need = (size_t)g.grid * k;
        if (need > c->scan_cap) {
            size_t c0 = c->scan_cap, c1 = c->scan_cap;
            rc = ensure(&c->d_scan[0], &c0, need); if (rc) return rc;
            rc = ensure(&c->d_scan[1], &c1, need); if (rc) return rc;
            c->scan_cap = c0 < c1 ? c0 : c1;
        }
    }
    g.pdl = 0;
    ScanParams p = map_params(ix);
    p.rows = ix->d_codes; p.norms = reinterpret_cast<const float*>(ix->d_inorms); p.queries = reinterpret_cast<const float*>(d_q8);
    p.partials = c->d_scan[0]; p.n_rows = ix->n_codes; p.d = words; p.ld = words; p.row_words = g.row_words; p.tile_rows = g.tile_rows;
    p.stages = g.stages; p.active_groups = g.groups; p.k = k; p.cand_cap = g.cand_cap; p.sync_interval = g.sync_interval; p.use_l2_hint = ix->opt_l2_hint;
    p.i8_q_norm = d_qn; p.i8_q_sum = d_qs; p.i8_tie_mode = tie_mode; p.i8_tie_vstar = vstar;
    p.trace = nullptr; p.early_trigger = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ix->opt_timing) { CUDA_TRY(cudaEventCreate(&e0)); CUDA_TRY(cudaEventCreate(&e1)); CUDA_TRY(cudaEventRecord(e0, st)); }
    rc = launch_scan_t<uint32_t, METRIC_I8, 1>(p, g, st);
    if (rc) return rc;
    ix->launches++;
    if (ix->opt_timing) {
        CUDA_TRY(cudaEventRecord(e1, st));
        std::lock_guard<std::mutex> lk(ix->ev_mu);
        ix->timed.push_back({e0, e1, 0});
    }
    const uint32_t grid = g.grid;
    cudaError_t e = cudaSuccess;
    {
        size_t oc = c->out_cap, oc2 = c->out_cap;
        rc = ensure(&c->d_rows, &oc, (size_t)k); if (rc) return rc;
        rc = ensure(&c->d_scores, &oc2, (size_t)k); if (rc) return rc;
        c->out_cap = oc < oc2 ? oc : oc2;
        rc = ensure(&c->d_counts, &c->cnt_cap, 4); if (rc) return rc;
    }
    rc = merge_lists(ix, c, c->d_scan[0], 1, grid, k, 0, nullptr, c->d_rows, c->d_scores, c->d_counts, st, (size_t)grid * k, k);
    if (rc) return rc;
    rows->assign(k, 0); scores->assign(k, 0.0f);
    uint32_t cnt = 0;
    e = cudaMemcpyAsync(rows->data(), c->d_rows, k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(scores->data(), c->d_scores, k * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, c->d_counts, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return fail(CGVEC_ERR_CUDA, "int8 search failed: %s", cudaGetErrorString(e));
    rows->resize(cnt); scores->resize(cnt);
    *cnt_out = cnt;
    return CGVEC_OK;
}

// Row index of the `limit`-th row with a non-zero int8 norm (the last row of search_optimized's initial fill), or ~0 if the
// index holds fewer such rows.
static int i8_initial_fill_end(Index* ix, uint32_t limit, uint64_t* out) {
    {
        std::lock_guard<std::mutex> lk(ix->i8_mu);
        auto it = ix->i8_fill_end.find(limit);
        if (it != ix->i8_fill_end.end()) { *out = it->second; return CGVEC_OK; }
    }
    *out = ~0ull;
    uint64_t seen = 0, pos = 0;
    std::vector<int32_t> buf;
    while (pos < ix->n_codes) {
        const uint64_t m = std::min<uint64_t>(ix->n_codes - pos, (uint64_t)limit + 4096);
        buf.resize(m);
        CUDA_TRY(cudaMemcpy(buf.data(), ix->d_inorms + pos, m * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < m && *out == ~0ull; ++i)
            if (buf[i] != 0 && ++seen == limit) *out = pos + i;
        if (*out != ~0ull) break;
        pos += m;
    }
    std::lock_guard<std::mutex> lk(ix->i8_mu);
    ix->i8_fill_end[limit] = *out;
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_search_i8(const cgvec_index* cix, const float* query, uint32_t limit, uint64_t* out_rows, float* out_scores,
                                 uint32_t* out_count) {
    NvtxRange nvtx_("cgvec.search_i8");
    RwGuard rw_guard_(cix, false);
    Index* ix = const_cast<cgvec_index*>(cix);
    if (!ix || !query) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    if (out_count) *out_count = 0;
    const uint32_t k = limit < 1 ? 1 : limit;                   // optimization.rs:64 `_limit.max(1)`
    if (ix->n_codes == 0) return CGVEC_OK;                       // :70-72 empty -> empty
    if (k + 1 > kMaxK) return fail(CGVEC_ERR_UNSUPPORTED, "limit = %u exceeds the fused top-k limit of %u", k, kMaxK - 1);
    CUDA_TRY(cudaSetDevice(ix->device));
    SearchCtx* c = nullptr;
    int rc = ctx_acquire(ix, &c);
    if (rc) return rc;
    auto done = [&](int code) { ctx_release(ix, c); return code; };
    cudaStream_t st = c->stream;
    cudaStreamWaitEvent(st, c->done, 0);
    const uint32_t ld8 = ix->ld8;
    rc = ensure(&c->d_q, &c->q_cap, (size_t)ix->dim + ld8 / 4 + 16); if (rc) return done(rc);
    int8_t* d_q8 = reinterpret_cast<int8_t*>(c->d_q + ((ix->dim + 3) & ~3u));
    float* d_qn = reinterpret_cast<float*>(d_q8 + ld8);
    int32_t* d_qs = reinterpret_cast<int32_t*>(d_qn + 1);
    cudaError_t e = cudaMemcpyAsync(c->d_q, query, ix->dim * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return done(fail(CGVEC_ERR_CUDA, "H2D failed: %s", cudaGetErrorString(e)));
    quantize_query_i8_kernel<<<1, 256, 0, st>>>(c->d_q, ix->dim, ld8, d_q8, d_qn, d_qs);
    ix->launches++;

    // Pass 1: the k + 1 best rows under (score, lower row).  That settles everything unless the (k+1)-th score TIES with the
    // k-th: search_optimized (:139-149) keeps a running list with strict `>` replacement of its first element and stable sorts,
    // so which of several equal boundary scores it ends up holding, and the order among equal scores, depend on arrival order.
    // Its outcome has a closed form (validated against the sequential restatement in oracle/ on tie-heavy inputs):
    //   I      = the first k rows with a non-zero norm (the initial fill)
    //   v*     = k-th best score;  A = rows with score > v*;  T = k-th lowest row among those with score >= v*
    //   B      = rows with score == v* and row <= T, in list order: non-initial rows by DESCENDING row, then initial rows ascending
    //   result = A  U  B minus its first P entries, P = #{a in A : a > T}
    //   order  = score descending; equal scores: non-initial rows by descending row first, then initial rows ascending.
    std::vector<uint64_t> r1, r2;
    std::vector<float> s1, s2;
    uint32_t c1 = 0, c2 = 0;
    rc = i8_pass(ix, c, d_q8, d_qn, d_qs, k + 1, 0, 0.0f, st, &r1, &s1, &c1);
    if (rc) return done(rc);
    float qn = 0.0f;
    cudaMemcpy(&qn, d_qn, sizeof(float), cudaMemcpyDeviceToHost);
    ix->searches++;
    if (qn == 0.0f || c1 == 0) return done(CGVEC_OK);            // optimization.rs:113-115: zero query -> empty
    uint64_t init_end = ~0ull;
    rc = i8_initial_fill_end(ix, k, &init_end);
    if (rc) return done(rc);
    auto initial = [&](uint64_t row) { return init_end == ~0ull || row <= init_end; };
    struct Hit { uint64_t row; float score; };
    std::vector<Hit> keep;
    if (c1 <= k || s1[k] < s1[k - 1]) {                          // no tie across the boundary: the k best are the result set
        for (uint32_t i = 0; i < c1 && i < k; ++i) keep.push_back({r1[i], s1[i]});
    } else {
        const float vstar = s1[k - 1];
        rc = i8_pass(ix, c, d_q8, d_qn, d_qs, k, 1, vstar, st, &r2, &s2, &c2);     // the k lowest rows with score >= v*
        if (rc) return done(rc);
        std::vector<Hit> A;
        for (uint32_t i = 0; i < c1; ++i) if (s1[i] > vstar) A.push_back({r1[i], s1[i]});
        uint64_t T = 0;
        for (uint64_t r : r2) T = std::max(T, r);
        auto in_A = [&](uint64_t row) { for (auto& a : A) if (a.row == row) return true; return false; };
        std::vector<uint64_t> B_rep, B_init;
        for (uint64_t r : r2) if (!in_A(r)) (initial(r) ? B_init : B_rep).push_back(r);
        std::sort(B_rep.begin(), B_rep.end(), std::greater<uint64_t>());
        std::sort(B_init.begin(), B_init.end());
        std::vector<uint64_t> B(B_rep);
        B.insert(B.end(), B_init.begin(), B_init.end());
        size_t P = 0;
        for (auto& a : A) if (a.row > T) ++P;
        keep = A;
        for (size_t i = P; i < B.size(); ++i) keep.push_back({B[i], vstar});
    }
    std::sort(keep.begin(), keep.end(), [&](const Hit& x, const Hit& y) {
        if (x.score != y.score) return x.score > y.score;
        const bool xi = initial(x.row), yi = initial(y.row);
        if (xi != yi) return !xi;                                // rows that entered by replacement sit in front of the initial fill
        return xi ? x.row < y.row : x.row > y.row;
    });
    for (size_t i = 0; i < keep.size(); ++i) { if (out_rows) out_rows[i] = keep[i].row; if (out_scores) out_scores[i] = keep[i].score; }
    if (out_count) *out_count = (uint32_t)keep.size();
    return done(CGVEC_OK);
}

// ================================================================================================
// flat matrix file: MemoryOptimizer::save_to_mmap / load_from_mmap, codegraph-vector/src/memory.rs:241-374
//   [u64 vector_count][u64 dimension][vector_count * dimension f32, row-major], native endian, exact file size
// ================================================================================================
CGVEC_EXPORT int cgvec_save_flat(const cgvec_index* cix, const char* path) {
    RwGuard rw_guard_(cix, false);
    Index* ix = const_cast<cgvec_index*>(cix);
    if (!ix || !path) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(CGVEC_ERR_BAD_ARG, "Failed to create mmap file: %s", path);
    uint64_t hdr[2] = {ix->n, ix->dim};
    bool ok = fwrite(hdr, sizeof(hdr), 1, f) == 1;
    const uint64_t chunk = 1u << 16;
    std::vector<float> buf;
    for (uint64_t r = 0; ok && r < ix->n; r += chunk) {
        uint64_t m = ix->n - r < chunk ? ix->n - r : chunk;
        buf.resize((size_t)m * ix->dim);
        int rc = cgvec_get_rows(cix, r, m, buf.data());      // widens f16 storage exactly
        if (rc) { fclose(f); return rc; }
        ok = fwrite(buf.data(), sizeof(float), buf.size(), f) == buf.size();
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(CGVEC_ERR_BAD_ARG, "Failed to write to file: %s", path);
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_load_flat(cgvec_index* ix, const char* path, uint64_t* out_rows_loaded) {
    CGVEC_WRITE_GUARD(ix);
    if (!ix || !path) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(CGVEC_ERR_BAD_ARG, "Failed to open mmap file: %s", path);
    auto bail = [&](int code) { fclose(f); return code; };
    uint64_t hdr[2];
    if (fseek(f, 0, SEEK_END) != 0) return bail(fail(CGVEC_ERR_BAD_ARG, "Failed to seek in file"));
    const long long size = ftell(f);
    rewind(f);
    if (size < (long long)sizeof(hdr) || fread(hdr, sizeof(hdr), 1, f) != 1) return bail(fail(CGVEC_ERR_BAD_ARG, "Invalid mmap file: too small"));   // memory.rs:322-326
    if (hdr[1] != ix->dim) return bail(fail(CGVEC_ERR_BAD_DIM, "Dimension mismatch: expected %u, found %llu", ix->dim, (unsigned long long)hdr[1]));      // :333-338
    const unsigned long long expect = sizeof(hdr) + (unsigned long long)hdr[0] * hdr[1] * sizeof(float);
    if ((unsigned long long)size != expect) return bail(fail(CGVEC_ERR_BAD_ARG, "Invalid mmap file size: expected %llu, got %lld", expect, size));       // :345-351
    int rc = cgvec_reserve(ix, ix->n + hdr[0]);
    if (rc) return bail(rc);
    const uint64_t chunk = 1u << 16;
    std::vector<float> buf;
    std::vector<uint16_t> half;
    for (uint64_t r = 0; r < hdr[0]; r += chunk) {
        uint64_t m = hdr[0] - r < chunk ? hdr[0] - r : chunk;
        buf.resize((size_t)m * ix->dim);
        if (fread(buf.data(), sizeof(float), buf.size(), f) != buf.size()) return bail(fail(CGVEC_ERR_BAD_ARG, "Failed to read matrix data"));
        if (ix->dtype == CGVEC_F32) rc = cgvec_add(ix, nullptr, buf.data(), m);
        else {
            half.resize(buf.size());
            for (size_t i = 0; i < buf.size(); ++i) half[i] = __half_as_ushort(__float2half_rn(buf[i]));
            rc = cgvec_add_f16(ix, nullptr, half.data(), m);
        }
        if (rc) return bail(rc);
    }
    fclose(f);
    if (out_rows_loaded) *out_rows_loaded = hdr[0];
    return CGVEC_OK;
}

// ================================================================================================
// host-side helpers (pure CPU)
// ================================================================================================
CGVEC_EXPORT int cgvec_shard_range(uint64_t n, int world, int rank, uint64_t* begin, uint64_t* end) {
    if (world < 1 || rank < 0 || rank >= world) return fail(CGVEC_ERR_BAD_ARG, "bad rank %d / world %d", rank, world);
    uint64_t per = (n + (uint64_t)world - 1) / (uint64_t)world;   // GPU g owns [g*ceil(N/G), (g+1)*ceil(N/G))
    uint64_t b = per * (uint64_t)rank, e = b + per;
    if (b > n) b = n;
    if (e > n) e = n;
    if (begin) *begin = b;
    if (end) *end = e;
    return CGVEC_OK;
}

// Placement rule of the single-process multi-device index (multi_device.inl): rows are dealt in kMultiBlk-row blocks
// round-robin.  Exposed so that host code (and the CPU tests) can reason about where a global row lives.
CGVEC_EXPORT int cgvec_multi_locate(uint32_t n_devices, uint64_t global_row, uint32_t* out_shard, uint64_t* out_local_row) {
    if (n_devices == 0) return fail(CGVEC_ERR_BAD_ARG, "n_devices must be >= 1");
    const uint64_t b = global_row / kMultiBlk;
    if (out_shard) *out_shard = (uint32_t)(b % n_devices);
    if (out_local_row) *out_local_row = (b / n_devices) * kMultiBlk + global_row % kMultiBlk;
    return CGVEC_OK;
}
CGVEC_EXPORT uint64_t cgvec_multi_local_count(uint32_t n_devices, uint32_t shard, uint64_t n_rows) {
    if (n_devices == 0 || shard >= n_devices) return 0;
    const uint64_t full = n_rows / kMultiBlk, rem = n_rows % kMultiBlk;
    uint64_t c = (full / n_devices) * kMultiBlk + ((full % n_devices) > shard ? kMultiBlk : 0);
    if (full % n_devices == shard) c += rem;
    return c;
}

CGVEC_EXPORT int cgvec_merge_topk_host(const uint64_t* rows, const float* scores, const uint32_t* counts, uint32_t parts, uint32_t k,
                                       int ascending, uint64_t* out_rows, float* out_scores, uint32_t* out_count) {
    if ((!rows || !scores || !counts) && parts && k) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    std::vector<uint32_t> head(parts, 0);
    auto better = [&](float x, uint64_t xr, float y, uint64_t yr) {
        bool xn = std::isnan(x), yn = std::isnan(y);
        if (xn != yn) return yn;
        if (!xn && x != y) return ascending ? x < y : x > y;
        return xr < yr;
    };
    uint32_t o = 0;
    for (; o < k; ++o) {
        int best = -1;
        for (uint32_t p = 0; p < parts; ++p) {
            if (head[p] >= counts[p] || head[p] >= k) continue;
            size_t i = (size_t)p * k + head[p];
            if (best < 0) { best = (int)p; continue; }
            size_t j = (size_t)best * k + head[best];
            if (better(scores[i], rows[i], scores[j], rows[j])) best = (int)p;
        }
        if (best < 0) break;
        size_t j = (size_t)best * k + head[best]++;
        if (out_rows) out_rows[o] = rows[j];
        if (out_scores) out_scores[o] = scores[j];
    }
    if (out_count) *out_count = o;
    return CGVEC_OK;
}

// AUTO's cost model as a pure function (milliseconds per call on the exact-order kernel and on the tensor path).
CGVEC_EXPORT int cgvec_path_cost_model(cgvec_dtype storage, uint32_t dim, uint64_t rows, uint32_t nq, uint32_t tensor_batch_limit,
                                       double* out_exact_ms, double* out_tensor_ms) {
    if (!out_exact_ms || !out_tensor_ms || dim == 0 || nq == 0) return fail(CGVEC_ERR_BAD_ARG, "bad argument");
    tensor_cost_model(storage == CGVEC_F32, dim, rows, nq, tensor_batch_limit, out_exact_ms, out_tensor_ms);
    return CGVEC_OK;
}

CGVEC_EXPORT uint64_t cgvec_prefetch_k_basic(uint64_t limit) {       // search.rs:113
    uint64_t a = limit > UINT64_MAX / 3 ? UINT64_MAX : limit * 3, b = limit + 10;
    return a > b ? a : b;
}
CGVEC_EXPORT uint64_t cgvec_prefetch_k_filtered(uint64_t limit) {    // search.rs:276
    uint64_t a = limit > UINT64_MAX / 4 ? UINT64_MAX : limit * 4, b = limit + 25;
    return a > b ? a : b;
}
CGVEC_EXPORT void cgvec_normalize_scores(float* s, size_t n) {       // search.rs:574-592
    if (!s || n == 0) return;
    float mn = INFINITY, mx = -INFINITY;
    for (size_t i = 0; i < n; ++i) { if (s[i] < mn) mn = s[i]; if (s[i] > mx) mx = s[i]; }
    float range = mx - mn;
    if (!(range > 1e-12f)) range = 1e-12f;
    for (size_t i = 0; i < n; ++i) s[i] = (s[i] - mn) / range;
}

// ================================================================================================
// introspection
// ================================================================================================
CGVEC_EXPORT int cgvec_get_stats(const cgvec_index* cix, cgvec_stats* out) {
    Index* ix = const_cast<cgvec_index*>(cix);
    if (!ix || !out) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    if (!ix->parts.empty()) {
        int rc = cgvec_get_stats(static_cast<const cgvec_index*>(ix->parts[0]), out);
        if (rc) return rc;
        out->rows = ix->n;
        out->coalesced_batches = ix->co_batches.load(); out->coalesced_queries = ix->co_queries.load();
        for (size_t s = 1; s < ix->parts.size(); ++s) {
            cgvec_stats t;
            rc = cgvec_get_stats(static_cast<const cgvec_index*>(ix->parts[s]), &t);
            if (rc) return rc;
            out->kernel_launches += t.kernel_launches;
            out->bytes_resident += t.bytes_resident;
        }
        return CGVEC_OK;
    }
    drain_timings(ix);
    memset(out, 0, sizeof(*out));
    out->kernel_launches = ix->launches.load();
    out->searches = ix->searches.load();
    out->rows = ix->n;
    out->bytes_resident = ix->n * (uint64_t)ix->ld * ix->esize + ix->n * sizeof(float);
    out->sm_count = (uint32_t)ix->sm_count;
    out->grid = ix->last_geom.grid;
    out->block = kScanThreads;
    out->smem_bytes = ix->last_geom.smem;
    out->stages = ix->last_geom.stages;
    out->tile_rows = ix->last_geom.tile_rows;
    out->last_scan_ms = ix->last_scan_ms;
    out->scan_ms_total = ix->scan_ms_total;
    out->scans_timed = ix->scan_timed;
    out->tc_batches = ix->tc_batches.load();
    out->tc_fallbacks = ix->tc_fallbacks.load();
    out->exchange_mode = ix->last_exchange.load();
    out->tc_main_ms_total = ix->tc_main_ms_total;
    out->tc_main_timed = ix->tc_main_timed;
    out->coalesced_batches = ix->co_batches.load(); out->coalesced_queries = ix->co_queries.load();
    return CGVEC_OK;
}

CGVEC_EXPORT int cgvec_set_option(cgvec_index* ix, const char* key, int64_t value) {
    if (!ix || !key) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    for (Index* part : ix->parts) { int rc = cgvec_set_option(static_cast<cgvec_index*>(part), key, value); if (rc) return rc; }
    std::string k(key);
    if (k == "tile_rows") ix->opt_tile_rows = (int)value;
    else if (k == "stages") ix->opt_stages = (int)value;
    else if (k == "sync_interval") ix->opt_sync = (int)value;
    else if (k == "l2_hint") ix->opt_l2_hint = (int)value;
    else if (k == "grid") ix->opt_grid = (int)value;
    else if (k == "timing") { ix->opt_timing = (int)value; if (!value) drain_timings(ix); }
    else if (k == "reset_timing") { drain_timings(ix); ix->scan_ms_total = 0; ix->scan_timed = 0; ix->tc_main_ms_total = 0; ix->tc_main_timed = 0; }
    else if (k == "max_batch") ix->opt_max_nq = (int)value;
    else if (k == "pdl") ix->opt_pdl = (int)value;
    else if (k == "coalesce") ix->opt_coalesce = (int)value;
    else if (k == "coalesce_max") ix->opt_coalesce_max = (int)value;
    else if (k == "trace") {
        cudaSetDevice(ix->device);
        if (value && !ix->d_trace) {
            if (cudaMalloc(reinterpret_cast<void**>(&ix->d_trace), Index::kTraceCap * 16) != cudaSuccess) { ix->d_trace = nullptr; return fail(CGVEC_ERR_OOM, "trace buffer"); }
        }
        if (ix->d_trace) {
            cudaDeviceSynchronize();
            std::vector<uint64_t> init(Index::kTraceCap * 2);
            for (uint32_t i = 0; i < Index::kTraceCap; ++i) { init[2 * i] = ~0ull; init[2 * i + 1] = 0; }
            cudaMemcpy(ix->d_trace, init.data(), init.size() * 8, cudaMemcpyHostToDevice);
            ix->trace_kinds.clear();
        }
        if (!value && ix->d_trace) { cudaFree(ix->d_trace); ix->d_trace = nullptr; }
    }
    else if (k == "p2p") ix->opt_p2p = (int)value;
    else if (k == "xchg_timeout_ms") ix->opt_xchg_timeout_ms = (int)value;
    else if (k == "tc_min_batch") ix->opt_tc_min_nq = (int)value;
    else if (k == "tc_stages") ix->opt_tc_stages = (int)value;
    else if (k == "tc_kbs") ix->opt_tc_kbs = (int)value;
    else if (k == "tc_flow") ix->opt_tc_flow = (int)value;
    else if (k == "tc_max_n") ix->opt_tc_max_n = (int)value;
    else if (k == "tc_margin") ix->opt_tc_margin = (int)value;
    else if (k == "tc_target") ix->opt_tc_target = (int)value;
    else if (k == "tc_l2promo") ix->opt_tc_l2promo = (int)value;
    else if (k == "tc_prefetch") ix->opt_tc_prefetch = (int)value;
    else if (k == "tc_first") ix->opt_tc_first = (int)value;
    else if (k == "tc_kernel") ix->opt_tc_kernel = (int)value;
    else if (k == "tc_debug") ix->opt_tc_debug = (int)value;
    else if (k == "tc2_max_n") ix->opt_tc2_max_n = (int)value;
    else return fail(CGVEC_ERR_BAD_ARG, "unknown option '%s'", key);
    return CGVEC_OK;
}

// Launch timeline recorded since option "trace" was set: out[3*i] = kind (1 scan, 2 merge, 3 exchange), start ns, end ns.
CGVEC_EXPORT int cgvec_get_trace(const cgvec_index* cix, uint64_t* out, uint32_t max_entries, uint32_t* out_n) {
    Index* ix = const_cast<cgvec_index*>(cix);
    if (!ix || !out_n) return fail(CGVEC_ERR_BAD_ARG, "NULL argument");
    *out_n = 0;
    if (!ix->d_trace) return CGVEC_OK;
    CUDA_TRY(cudaSetDevice(ix->device));
    CUDA_TRY(cudaDeviceSynchronize());
    uint32_t n = (uint32_t)ix->trace_kinds.size();
    if (n > max_entries) n = max_entries;
    std::vector<uint64_t> raw((size_t)n * 2);
    if (n) CUDA_TRY(cudaMemcpy(raw.data(), ix->d_trace, raw.size() * 8, cudaMemcpyDeviceToHost));
    for (uint32_t i = 0; i < n && out; ++i) { out[3 * i] = ix->trace_kinds[i]; out[3 * i + 1] = raw[2 * i]; out[3 * i + 2] = raw[2 * i + 1]; }
    *out_n = n;
    return CGVEC_OK;
}

CGVEC_EXPORT const char* cgvec_last_error(void) { return g_err; }
CGVEC_EXPORT const char* cgvec_version(void) { return "cgvec_b200 0.1.0 (sm_100a)"; }

