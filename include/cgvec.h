/*
 * cgvec.h — C ABI of the B200-native embedding similarity-search path (libcgvec_b200.so).
 *
 * This is the drop-in boundary for the `codegraph-vector` crate of Jakedismo/codegraph-rust
 * (reference paths below are relative to /root/reference/crates/).  The reference is Rust and has
 * no FFI for this path today (its crates/codegraph-vector/src/gpu.rs is a mock), so the entry
 * points are exactly what a Rust `extern "C"` block implementing the reference's own traits would
 * bind — see INTEGRATION.md for that binding (`rust/` holds the shim source):
 *
 *   trait VectorStore                (codegraph-core/src/traits.rs:11-16)
 *       store_embeddings(&mut self, &[CodeNode])          -> cgvec_add / cgvec_add_f16
 *       search_similar(&self, &[f32], limit) -> Vec<NodeId> -> cgvec_search
 *       get_embedding(&self, NodeId) -> Option<Vec<f32>>   -> cgvec_get
 *   trait SurrealVectorBackend       (codegraph-vector/src/surreal_store.rs:11-22)
 *       vector_knn(column, Vec<f32>, limit, ef) -> Vec<(String, f32)>  -> cgvec_search (+ 1 - score)
 *   ParallelVectorOps::parallel_top_k_search (codegraph-vector/src/simd_ops.rs:361-383) -> cgvec_search
 *   SIMDVectorOps::{cosine,dot,l2}_*_avx2     (simd_ops.rs:15-183)   -> cgvec_metric
 *   ParallelVectorOps::parallel_normalize_vectors (simd_ops.rs:386-419) -> cgvec_normalize_rows
 *   SemanticSearch::calculate_similarity_score (codegraph-vector/src/search.rs:207-217,519-533) -> cgvec_rescore
 *   ModelOptimizer::search_baseline  (codegraph-vector/src/optimization.rs:376-418) -> cgvec_search_ex(formula=BASELINE)
 *   GpuAcceleration::{upload_vectors, compute_distances} (codegraph-vector/src/gpu.rs:221-291)
 *                                                            -> cgvec_add (flat N x d) / cgvec_distances_first
 *
 * Rules of the boundary: plain pointers and sizes only; the library never keeps a caller pointer
 * after a call returns; outputs go into caller-allocated buffers; every function returns a
 * cgvec_status (0 = ok, < 0 = error) and never aborts, throws or panics across the boundary;
 * cgvec_last_error() gives the thread-local message that the Rust shim maps to
 * CodeGraphError::Vector(msg) (codegraph-core/src/error.rs:18-19).  There is NO CPU fallback: on a
 * machine without a usable sm_100 device every compute entry point fails with CGVEC_ERR_NO_DEVICE.
 *
 * Result contract: best score first; ties -> lower row index; NaN scores rank
 * last; k > N returns N results; k == 0 or nq == 0 returns nothing.  With formula
 * CGVEC_FORMULA_SIMD the returned scores are bit-identical to the reference's
 * adaptive_cosine_similarity / dot_product_avx2 / l2_distance_avx2 on an AVX2+FMA host.
 */
#ifndef CGVEC_H
#define CGVEC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgvec_index cgvec_index;

typedef enum { CGVEC_F32 = 0, CGVEC_F16 = 1 } cgvec_dtype;   /* storage type of the resident matrix */

typedef enum {
    CGVEC_COSINE = 0,   /* simd_ops.rs:15-78 / :257-278 via :281-295; higher is better            */
    CGVEC_DOT = 1,      /* simd_ops.rs:149-183; higher is better                                   */
    CGVEC_L2 = 2        /* simd_ops.rs:105-143; LOWER is better (results ascend)                   */
} cgvec_metric;

typedef enum {
    CGVEC_FORMULA_SIMD = 0,      /* adaptive_cosine_similarity: AVX2 lane order if d >= 32 else scalar (simd_ops.rs:281-295) */
    CGVEC_FORMULA_SCALAR = 1,    /* cosine_similarity_scalar (simd_ops.rs:257-278)                                          */
    CGVEC_FORMULA_SEQ = 2,       /* search.rs:519-533 cosine_similarity: dot/(sqrt(na)*sqrt(nb)), sequential sums           */
    CGVEC_FORMULA_BASELINE = 3   /* optimization.rs:404-418 cosine_distance = 1 - SEQ, INFINITY on zero norm; ascending     */
} cgvec_formula;

typedef enum {
    CGVEC_OK = 0,
    CGVEC_ERR_BAD_ARG = -1,
    CGVEC_ERR_BAD_DIM = -2,      /* dimension mismatch (persistent.rs:1046-1052, graph_vector.rs:396-402) */
    CGVEC_ERR_OOM = -3,
    CGVEC_ERR_CUDA = -4,
    CGVEC_ERR_NCCL = -5,
    CGVEC_ERR_NOT_FOUND = -6,    /* get_embedding -> None */
    CGVEC_ERR_NO_DEVICE = -7,    /* no sm_100 GPU: the product has no CPU fallback */
    CGVEC_ERR_UNSUPPORTED = -8,
    CGVEC_ERR_DISABLED = -9      /* cgvec_create_from_env: performance.enable_gpu is false (config_manager.rs:362-364); keep the CPU store */
} cgvec_status;

/* Which scan kernel family serves a search. AUTO picks EXACT for small batches and TENSOR for large ones. */
typedef enum { CGVEC_PATH_AUTO = 0, CGVEC_PATH_EXACT = 1, CGVEC_PATH_TENSOR = 2 } cgvec_path;

typedef struct {
    uint32_t struct_size;        /* = sizeof(cgvec_search_opts); forward compatibility */
    cgvec_metric metric;
    cgvec_formula formula;
    cgvec_path path;
    void* stream;                /* cudaStream_t to run on; NULL = a stream owned by the index */
    int device_io;               /* 1: queries/out_rows/out_scores/out_counts are DEVICE pointers on the index's
                                    (first) device, nothing is copied or synchronised; out_ids must be NULL */
} cgvec_search_opts;

/* ---- lifecycle ------------------------------------------------------------------------------ */

/* Single-process index over n_devices GPUs: row blocks are dealt round-robin to the devices; a search scans on every
 * device at once and merges over NVLink peer memory (fused exchange kernel for k <= 128 and <= 4 queries per step, peer
 * copies into the first device + one merge launch for larger k and tensor-core batches; no NCCL in this mode).
 * Serves every path (exact, tensor), host and device I/O (buffers on the first device), k <= 1024.
 * device_ids == NULL -> devices 0..n_devices-1. */
int cgvec_create(uint32_t dim, cgvec_dtype storage, const int* device_ids, int n_devices, cgvec_index** out);
/* Deployment switch.  `enable_gpu` is the host's PerformanceConfig.enable_gpu (codegraph-core/src/config_manager.rs:362-364,
 * default false); the environment overrides it: CODEGRAPH_ENABLE_GPU = 1/true/0/false.  Disabled -> CGVEC_ERR_DISABLED (the host
 * keeps its CPU vector store).  CODEGRAPH_B200_DEVICES = "all" | a device COUNT ("4" -> 0..3) | a list ("0,2,5"); unset -> device 0.
 * More than one device gives the single-process multi-device index. */
int cgvec_create_from_env(uint32_t dim, cgvec_dtype storage, int enable_gpu, cgvec_index** out);

/* One rank of a multi-process row-sharded index (one process per GPU, e.g. under torchrun).
 * All ranks must pass the same 128-byte id from cgvec_nccl_unique_id() (rank 0 creates it and
 * broadcasts it out of band).  This rank's rows carry global indices row_offset + local row.
 * world == 1 needs no id (pass NULL). */
int cgvec_create_rank(uint32_t dim, cgvec_dtype storage, int device, int rank, int world,
                      const void* nccl_unique_id, uint64_t row_offset, cgvec_index** out);
int cgvec_nccl_unique_id(void* out_128_bytes);

int cgvec_destroy(cgvec_index* idx);   /* CGVEC_ERR_UNSUPPORTED while cgvec_stream / cgvec_serve sessions of the index are open */

/* ---- write side: VectorStore::store_embeddings ----------------------------------------------- */

int cgvec_reserve(cgvec_index* idx, uint64_t n_rows);             /* capacity hint (exact-size allocation) */
/* Append n rows (row-major, n x dim).  ids may be NULL (rows are then addressable by row index only).
 * An id already present overwrites that row in place (InMemoryVectorStore insert semantics,
 * codegraph-core/src/integration/graph_vector.rs:470-477).  Requires external exclusion (&mut self). */
int cgvec_add(cgvec_index* idx, const uint8_t (*ids)[16], const float* rows_f32, uint64_t n);
int cgvec_add_f16(cgvec_index* idx, const uint8_t (*ids)[16], const uint16_t* rows_f16, uint64_t n);
/* In-place L2 normalisation of every stored row (parallel_normalize_vectors, simd_ops.rs:386-419). */
int cgvec_normalize_rows(cgvec_index* idx);
/* Deterministic synthetic rows generated on the device (bench / large parity properties): appends n rows
 * whose values depend only on (seed, global row, column); see DESIGN.md "synthetic inputs".  unit_norm != 0
 * L2-normalises each row (reference arithmetic) before storing. */
int cgvec_fill_synthetic(cgvec_index* idx, uint64_t n, uint64_t seed, int unit_norm);

/* ---- read side ------------------------------------------------------------------------------- */

uint64_t cgvec_len(const cgvec_index* idx);                       /* rows held by THIS process */
uint32_t cgvec_dim(const cgvec_index* idx);

/* VectorStore::search_similar / parallel_top_k_search for nq queries (row-major nq x dim, f32).
 * out_rows   [nq*k] global row indices, best first          (nullable)
 * out_ids    [nq*k] 16-byte NodeId (Uuid) of each hit       (nullable; zero for rows added without ids)
 * out_scores [nq*k] similarity (or L2 distance)              (nullable)
 * out_counts [nq]   number of valid results per query = min(k, N)   (nullable)
 * Reentrant: may be called concurrently from many host threads on the same index. */
int cgvec_search(const cgvec_index* idx, const float* queries, uint32_t nq, uint32_t k, cgvec_metric metric,
                 uint64_t* out_rows, uint8_t (*out_ids)[16], float* out_scores, uint32_t* out_counts);
int cgvec_search_ex(const cgvec_index* idx, const float* queries, uint32_t nq, uint32_t k,
                    const cgvec_search_opts* opts,
                    uint64_t* out_rows, uint8_t (*out_ids)[16], float* out_scores, uint32_t* out_counts);

/* ---- streaming batches (BASELINE config 5): many query embeddings against one store, the shape of
 * SemanticSearch::multi_vector_search (codegraph-vector/src/search.rs:347-361), fed batch after batch from HOST memory.
 * cgvec_stream_submit stages batch i+1 (pinned copy + asynchronous upload on a copy stream) and then runs the search of
 * batch i, so uploads travel behind the scan; results therefore lag one submit (*out_nq = 0 on the first call) and
 * cgvec_stream_flush drains the last batch.  Outputs are host buffers of max_batch*k (rows, scores) and max_batch
 * (counts) entries.  On a sharded index every rank submits the same batches (the exchange inside is collective). */
typedef struct cgvec_stream cgvec_stream;
int cgvec_stream_open(cgvec_index* idx, uint32_t max_batch, uint32_t k, cgvec_metric metric, cgvec_path path, cgvec_stream** out);
int cgvec_stream_submit(cgvec_stream* s, const float* queries /* nq x dim, host */, uint32_t nq, uint64_t* out_rows,
                        float* out_scores, uint32_t* out_counts, uint32_t* out_nq);
int cgvec_stream_flush(cgvec_stream* s, uint64_t* out_rows, float* out_scores, uint32_t* out_counts, uint32_t* out_nq);
int cgvec_stream_close(cgvec_stream* s);

/* ---- resident batch-1 sessions ------------------------------------------------------------------
 * The reference serves one search_similar call per query (VectorStore::search_similar, codegraph-core/src/traits.rs:11-16;
 * SemanticSearch::search_by_embedding, codegraph-vector/src/search.rs:91-144).  A session keeps the exact-order scan kernel
 * RESIDENT between those calls: a submit is a 64-byte descriptor and a doorbell word in pinned memory, a completion is a word
 * the host spins on — no kernel launch, copy node or stream synchronisation per query.  The kernel is started on demand, leaves
 * by itself after `idle_us` without work (it owns every SM while resident) and is restarted transparently.  Results are
 * bit-identical to cgvec_search (same kernels' arithmetic and keys).  1 <= k <= 64; cosine / dot / L2; single-device indexes
 * and the ranks of a sharded index (every rank opens a session and submits the same queries in the same order).
 * cgvec_serve_search: host buffers, blocking.  cgvec_serve_submit + cgvec_serve_wait: pipelined; with device_io the query
 * (dim rounded up to 4 floats) and the result buffers are device memory.  At most 6 tickets may be outstanding.
 * The write side of the index refuses to run while a session is open. */
typedef struct cgvec_server cgvec_server;
int cgvec_serve_open(cgvec_index* idx, uint32_t k, cgvec_metric metric, cgvec_server** out);
int cgvec_serve_search(cgvec_server* s, const float* query, uint64_t* out_rows, uint8_t (*out_ids)[16], float* out_scores, uint32_t* out_count);
int cgvec_serve_submit(cgvec_server* s, const float* query, int device_io, uint64_t* out_rows, float* out_scores, uint32_t* out_count,
                       uint32_t* out_ticket);
int cgvec_serve_wait(cgvec_server* s, uint32_t ticket);
int cgvec_serve_pause(cgvec_server* s);                          /* ask the resident kernel to leave now; the next submit restarts it */
/* CUDA-event bracket on the session's launch stream: start pauses the session and records; the kernel launch of the next submit,
 * every query and the kernel's exit lie inside; stop drains, makes the kernel leave, records and returns the elapsed ms. */
int cgvec_serve_timer_start(cgvec_server* s);
int cgvec_serve_timer_stop(cgvec_server* s, float* out_ms);
int cgvec_serve_stats(const cgvec_server* s, uint64_t* out_kernel_launches, uint64_t* out_queries_served);
int cgvec_serve_set(cgvec_server* s, const char* key, int64_t value);   /* idle_us, life_ms, abort_ms, wait_ms, max_inflight, l2_hint, contig */
int cgvec_serve_close(cgvec_server* s);

/* VectorStore::get_embedding: copies the row (widened to f32) into out_row[dim]; CGVEC_ERR_NOT_FOUND -> None. */
int cgvec_get(const cgvec_index* idx, const uint8_t id[16], float* out_row);
int cgvec_get_row(const cgvec_index* idx, uint64_t local_row, float* out_row);
int cgvec_get_rows(const cgvec_index* idx, uint64_t first_local_row, uint64_t n, float* out /* n x dim */);
int cgvec_row_of_id(const cgvec_index* idx, const uint8_t id[16], uint64_t* out_local_row);

/* SemanticSearch::calculate_similarity_score for a list of stored rows (search.rs:207-217):
 * out_scores[i] = formula(query, row[rows[i]]) computed on the device in the reference's exact order. */
int cgvec_rescore(const cgvec_index* idx, const float* query, const uint64_t* local_rows, uint32_t n,
                  cgvec_metric metric, cgvec_formula formula, float* out_scores);

/* GpuAcceleration::compute_distances as the reference's CPU twin defines it (gpu.rs:297-322):
 * cosine DISTANCE (optimization.rs:404-418 form) of the first `limit` stored rows. Returns count in *out_n. */
int cgvec_distances_first(const cgvec_index* idx, const float* query, uint64_t limit, float* out, uint64_t* out_n);

/* ---- int8 quantised scan: ModelOptimizer::quantize_batch + OptimizationResult::search_optimized
 * (codegraph-vector/src/optimization.rs:63-150, :212-224, :268-274).  cgvec_quantize_i8 builds u8 codes (value + 128)
 * of the currently stored rows; cgvec_search_i8 returns up to max(limit,1) rows by the int8 cosine, best first, with
 * scores bit-identical to the reference's f32 expression and EXACTLY the rows and the order the reference's running list
 * (strict `>` replacement + stable sorts, :139-149) produces, ties included.  Codes go stale on any write (quantize again).
 * Single-GPU indexes; limit <= 1023. */
int cgvec_quantize_i8(cgvec_index* idx);
int cgvec_get_codes_i8(const cgvec_index* idx, uint64_t first_row, uint64_t n, uint8_t* out /* n x dim */);
int cgvec_search_i8(const cgvec_index* idx, const float* query, uint32_t limit, uint64_t* out_rows, float* out_scores,
                    uint32_t* out_count);

/* ---- flat matrix file: MemoryOptimizer::save_to_mmap / load_from_mmap (codegraph-vector/src/memory.rs:241-374):
 * [u64 vector_count][u64 dimension][count*dimension f32 row-major].  load appends the file's rows (no ids) and
 * rejects a wrong dimension or file size exactly like the reference loader. */
int cgvec_save_flat(const cgvec_index* idx, const char* path);
int cgvec_load_flat(cgvec_index* idx, const char* path, uint64_t* out_rows_loaded);

/* ---- host-side helpers (pure CPU, no device needed; used by the shim and by the gloo tests) ---- */

/* Row range [begin, end) owned by `rank` of `world` for a contiguous split of n rows. */
int cgvec_shard_range(uint64_t n, int world, int rank, uint64_t* begin, uint64_t* end);
/* Placement of a global row in a single-process multi-device index (1024-row blocks dealt round-robin), and the number
 * of rows device `shard` holds when the index has n_rows. */
int cgvec_multi_locate(uint32_t n_devices, uint64_t global_row, uint32_t* out_shard, uint64_t* out_local_row);
uint64_t cgvec_multi_local_count(uint32_t n_devices, uint32_t shard, uint64_t n_rows);
/* Merge `parts` partial top-k lists (each k entries: global row + score, best first, counts[p] valid)
 * into the global top-k under the result contract.  ascending != 0 for L2 / BASELINE. */
int cgvec_merge_topk_host(const uint64_t* rows, const float* scores, const uint32_t* counts, uint32_t parts,
                          uint32_t k, int ascending, uint64_t* out_rows, float* out_scores, uint32_t* out_count);
/* search.rs:113 and :276 over-fetch sizes; search.rs:574-592 min-max normalisation. */
/* CGVEC_PATH_AUTO's cost model (DESIGN.md §5 "AUTO"): estimated milliseconds of one call with nq queries on the exact-order kernel
 * and on the tensor path (tensor_batch_limit = queries one tensor pass takes: 128, or 256 with the paired kernel).  Pure function. */
int cgvec_path_cost_model(cgvec_dtype storage, uint32_t dim, uint64_t rows, uint32_t nq, uint32_t tensor_batch_limit,
                          double* out_exact_ms, double* out_tensor_ms);
uint64_t cgvec_prefetch_k_basic(uint64_t limit);
uint64_t cgvec_prefetch_k_filtered(uint64_t limit);
void cgvec_normalize_scores(float* scores, size_t n);

/* ---- introspection --------------------------------------------------------------------------- */

typedef struct {
    uint64_t kernel_launches;    /* kernels of THIS library launched since create (all streams)        */
    uint64_t searches;
    uint64_t rows;               /* local rows                                                          */
    uint64_t bytes_resident;     /* matrix + norms bytes in HBM (local)                                 */
    uint32_t sm_count;
    uint32_t grid, block, smem_bytes, stages, tile_rows;   /* geometry of the last scan launch          */
    float last_scan_ms;          /* device time of the last scan kernel when option "timing" is on      */
    double scan_ms_total;        /* sum / count of scan-kernel device times since "reset_timing"         */
    uint64_t scans_timed;
    uint64_t tc_batches;         /* query batches served by the tensor-core path                         */
    uint64_t tc_fallbacks;       /* queries it could not prove exact and re-ran on the exact-order kernel */
    uint32_t exchange_mode;      /* last top-k exchange of a sharded index: 0 none, 1 fused NVLink peer-memory kernel, 2 NCCL all-gather */
    uint32_t reserved0;
    double tc_main_ms_total;     /* sum / count of device times of the tensor scan's main-range kernel ("timing" on) */
    uint64_t tc_main_timed;
    uint64_t coalesced_batches;  /* multi-query launches formed from concurrent batch-1 cgvec_search calls (group commit) */
    uint64_t coalesced_queries;  /* callers served through them                                                          */
} cgvec_stats;
int cgvec_get_stats(const cgvec_index* idx, cgvec_stats* out);
int cgvec_set_option(cgvec_index* idx, const char* key, int64_t value);   /* tuning knobs, see DESIGN.md */
/* Device-side launch timeline (after cgvec_set_option(idx, "trace", 1)): out[3*i] = kind (1 scan, 2 merge, 3 exchange),
 * out[3*i+1] = first CTA start, out[3*i+2] = last CTA end, both in %globaltimer nanoseconds. */
int cgvec_get_trace(const cgvec_index* idx, uint64_t* out, uint32_t max_entries, uint32_t* out_n);

const char* cgvec_last_error(void);
const char* cgvec_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CGVEC_H */
