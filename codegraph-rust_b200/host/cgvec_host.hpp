// cgvec_host.hpp — C++ host-side mirror of the reference interfaces that sit on the similarity-search path,
// written over the C ABI (include/cgvec.h).  The reference host code is Rust; no Rust toolchain exists in the
// build image, so (per the task rules) the compiled host layer is C++ with the same names, argument meaning
// and error behaviour.  The Rust shim a maintainer would add lives in ../rust/ (source only).
//
//   trait VectorStore                 codegraph-core/src/traits.rs:11-16
//   trait SurrealVectorBackend        codegraph-vector/src/surreal_store.rs:11-22
//   SurrealVectorStore::search_similar  codegraph-vector/src/surreal_store.rs:61-85
//   SemanticSearch::search_by_embedding codegraph-vector/src/search.rs:91-144
//
// Header-only; link with -lcgvec_b200.  Errors become cgvec::Error (CodeGraphError::Vector(msg), error.rs:18-19).
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "cgvec.h"

namespace cgvec {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("Vector error: " + m), code(c) {}
};
inline void check(int rc) {
    if (rc != CGVEC_OK) throw Error(rc, cgvec_last_error());
}

// NodeId = Uuid (codegraph-core/src/types.rs:8)
struct NodeId {
    std::array<uint8_t, 16> bytes{};
    bool operator==(const NodeId& o) const { return bytes == o.bytes; }
    std::string to_string() const {
        char buf[37];
        const uint8_t* b = bytes.data();
        std::snprintf(buf, sizeof(buf), "%02x%02x%02x%02x-%02x%02x-%02x%02x-%02x%02x-%02x%02x%02x%02x%02x%02x", b[0], b[1], b[2], b[3],
                      b[4], b[5], b[6], b[7], b[8], b[9], b[10], b[11], b[12], b[13], b[14], b[15]);
        return buf;
    }
    static std::optional<NodeId> parse_str(const std::string& s) {          // Uuid::parse_str (hyphenated or simple)
        NodeId id;
        int n = 0;
        auto hex = [](char c) { return c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1; };
        for (size_t i = 0; i < s.size();) {
            if (s[i] == '-') { ++i; continue; }
            if (i + 1 >= s.size() || n >= 16) return std::nullopt;
            int h = hex(s[i]), l = hex(s[i + 1]);
            if (h < 0 || l < 0) return std::nullopt;
            id.bytes[n++] = (uint8_t)(h * 16 + l);
            i += 2;
        }
        if (n != 16) return std::nullopt;
        return id;
    }
    static NodeId from_u64(uint64_t v) { NodeId id; for (int i = 0; i < 8; ++i) id.bytes[15 - i] = (uint8_t)(v >> (8 * i)); return id; }
};

// The two CodeNode fields the vector path reads (codegraph-core/src/node.rs:4-16).
struct CodeNode {
    NodeId id;
    std::optional<std::vector<float>> embedding;
};

// trait VectorStore (traits.rs:11-16)
class VectorStore {
public:
    virtual ~VectorStore() = default;
    virtual void store_embeddings(const std::vector<CodeNode>& nodes) = 0;
    virtual std::vector<NodeId> search_similar(const std::vector<float>& query_embedding, size_t limit) const = 0;
    virtual std::optional<std::vector<float>> get_embedding(const NodeId& node_id) const = 0;
};

class B200VectorStore : public VectorStore {
public:
    explicit B200VectorStore(uint32_t dimension, cgvec_dtype storage = CGVEC_F32, int device = 0) : dim_(dimension) {
        check(cgvec_create(dimension, storage, &device, 1, &idx_));
    }
    // one host process driving several GPUs: rows are dealt to the devices in blocks, searches merge over NVLink
    B200VectorStore(uint32_t dimension, const std::vector<int>& devices, cgvec_dtype storage = CGVEC_F32) : dim_(dimension) {
        check(cgvec_create(dimension, storage, devices.data(), (int)devices.size(), &idx_));
    }
    // Deployment switch: `enable_gpu` is PerformanceConfig.enable_gpu (config_manager.rs:362-364); CODEGRAPH_ENABLE_GPU overrides
    // it, CODEGRAPH_B200_DEVICES ("all" | count | "0,2,5") picks the GPUs.  nullptr = switched off: keep the CPU vector store.
    static std::shared_ptr<B200VectorStore> from_env(uint32_t dimension, bool enable_gpu, cgvec_dtype storage = CGVEC_F32) {
        cgvec_index* idx = nullptr;
        const int rc = cgvec_create_from_env(dimension, storage, enable_gpu ? 1 : 0, &idx);
        if (rc == CGVEC_ERR_DISABLED) return nullptr;
        check(rc);
        return std::shared_ptr<B200VectorStore>(new B200VectorStore(dimension, idx));
    }
    ~B200VectorStore() override { cgvec_destroy(idx_); }
    B200VectorStore(const B200VectorStore&) = delete;
    B200VectorStore& operator=(const B200VectorStore&) = delete;

    void store_embeddings(const std::vector<CodeNode>& nodes) override {
        std::vector<float> rows;
        std::vector<std::array<uint8_t, 16>> ids;
        for (const auto& n : nodes) {
            if (!n.embedding) continue;                                      // graph_vector.rs:472-476: None is skipped
            if (n.embedding->size() != dim_)                                  // persistent.rs:1046-1052
                throw Error(CGVEC_ERR_BAD_DIM, "Vector dimension " + std::to_string(n.embedding->size()) + " doesn't match expected " + std::to_string(dim_));
            rows.insert(rows.end(), n.embedding->begin(), n.embedding->end());
            ids.push_back(n.id.bytes);
        }
        if (ids.empty()) return;
        check(cgvec_add(idx_, reinterpret_cast<const uint8_t(*)[16]>(ids.data()), rows.data(), ids.size()));
    }

    std::vector<NodeId> search_similar(const std::vector<float>& q, size_t limit) const override {
        if (q.empty() || limit == 0) return {};                                // surreal_store.rs:62-64
        if (q.size() != dim_) throw Error(CGVEC_ERR_BAD_DIM, "Query dimension mismatch");   // graph_vector.rs:396-402
        std::vector<std::array<uint8_t, 16>> ids(limit);
        uint32_t count = 0;
        check(cgvec_search(idx_, q.data(), 1, (uint32_t)limit, CGVEC_COSINE, nullptr, reinterpret_cast<uint8_t(*)[16]>(ids.data()), nullptr, &count));
        std::vector<NodeId> out(count);
        for (uint32_t i = 0; i < count; ++i) out[i].bytes = ids[i];
        return out;
    }

    std::optional<std::vector<float>> get_embedding(const NodeId& id) const override {
        std::vector<float> row(dim_);
        int rc = cgvec_get(idx_, id.bytes.data(), row.data());
        if (rc == CGVEC_ERR_NOT_FOUND) return std::nullopt;
        check(rc);
        return row;
    }

    // ParallelVectorOps::parallel_top_k_search shape: (row index, similarity) pairs (simd_ops.rs:361-383)
    std::vector<std::pair<uint64_t, float>> top_k(const std::vector<float>& q, size_t k, cgvec_metric metric = CGVEC_COSINE) const {
        std::vector<uint64_t> rows(k);
        std::vector<float> scores(k);
        uint32_t count = 0;
        check(cgvec_search(idx_, q.data(), 1, (uint32_t)k, metric, rows.data(), nullptr, scores.data(), &count));
        std::vector<std::pair<uint64_t, float>> out(count);
        for (uint32_t i = 0; i < count; ++i) out[i] = {rows[i], scores[i]};
        return out;
    }

    cgvec_index* handle() const { return idx_; }
    uint32_t dimension() const { return dim_; }
    size_t len() const { return cgvec_len(idx_); }

private:
    B200VectorStore(uint32_t dimension, cgvec_index* adopted) : idx_(adopted), dim_(dimension) {}
    cgvec_index* idx_ = nullptr;
    uint32_t dim_;
};

// The serving loop of the reference is one search_similar per query (traits.rs:11-16; SemanticSearch::search_by_embedding,
// search.rs:91-144).  A ResidentSession serves those calls with the scan kernel RESIDENT on the GPU between them (cgvec_serve_*):
// same results as B200VectorStore::search_similar / top_k, no kernel launch per query.  The store must outlive the session and
// may not be written to while the session is open (the library refuses both).
class ResidentSession {
public:
    ResidentSession(std::shared_ptr<B200VectorStore> store, size_t limit, cgvec_metric metric = CGVEC_COSINE)
        : store_(std::move(store)), k_(limit) {
        check(cgvec_serve_open(store_->handle(), (uint32_t)limit, metric, &s_));
    }
    ~ResidentSession() { cgvec_serve_close(s_); }
    ResidentSession(const ResidentSession&) = delete;
    ResidentSession& operator=(const ResidentSession&) = delete;

    std::vector<NodeId> search_similar(const std::vector<float>& q) {
        if (q.size() != store_->dimension()) throw Error(CGVEC_ERR_BAD_DIM, "Query dimension mismatch");
        std::vector<std::array<uint8_t, 16>> ids(k_);
        uint32_t count = 0;
        check(cgvec_serve_search(s_, q.data(), nullptr, reinterpret_cast<uint8_t(*)[16]>(ids.data()), nullptr, &count));
        std::vector<NodeId> out(count);
        for (uint32_t i = 0; i < count; ++i) out[i].bytes = ids[i];
        return out;
    }
    std::vector<std::pair<uint64_t, float>> top_k(const std::vector<float>& q) {
        if (q.size() != store_->dimension()) throw Error(CGVEC_ERR_BAD_DIM, "Query dimension mismatch");
        std::vector<uint64_t> rows(k_);
        std::vector<float> scores(k_);
        uint32_t count = 0;
        check(cgvec_serve_search(s_, q.data(), rows.data(), nullptr, scores.data(), &count));
        std::vector<std::pair<uint64_t, float>> out(count);
        for (uint32_t i = 0; i < count; ++i) out[i] = {rows[i], scores[i]};
        return out;
    }
    void pause() { check(cgvec_serve_pause(s_)); }               // free the SMs now; the next call restarts the kernel

private:
    std::shared_ptr<B200VectorStore> store_;
    size_t k_;
    cgvec_server* s_ = nullptr;
};

// trait SurrealVectorBackend (surreal_store.rs:11-22)
class SurrealVectorBackend {
public:
    virtual ~SurrealVectorBackend() = default;
    virtual void upsert_nodes(const std::vector<CodeNode>& nodes) = 0;
    virtual std::vector<std::pair<std::string, float>> vector_knn(const std::string& column, const std::vector<float>& query_embedding,
                                                                   size_t limit, size_t ef_search) = 0;
    virtual std::optional<std::vector<float>> get_node_embedding(const NodeId& id) = 0;
};

// Exact brute-force KNN standing where SurrealDB's HNSW `<|limit,ef|>` stage stands
// (codegraph-graph/src/surrealdb_storage.rs:271-328): ids as "nodes:<uuid>", cosine DISTANCE ascending.
class B200Backend : public SurrealVectorBackend {
public:
    explicit B200Backend(std::shared_ptr<B200VectorStore> store) : store_(std::move(store)) {}
    void upsert_nodes(const std::vector<CodeNode>& nodes) override { store_->store_embeddings(nodes); }
    std::vector<std::pair<std::string, float>> vector_knn(const std::string& column, const std::vector<float>& q, size_t limit,
                                                           size_t /*ef_search: exact search has no beam*/) override {
        last_column = column;
        if (q.empty() || limit == 0) return {};
        std::vector<std::array<uint8_t, 16>> ids(limit);
        std::vector<float> scores(limit);
        uint32_t count = 0;
        check(cgvec_search(store_->handle(), q.data(), 1, (uint32_t)limit, CGVEC_COSINE, nullptr, reinterpret_cast<uint8_t(*)[16]>(ids.data()),
                           scores.data(), &count));
        std::vector<std::pair<std::string, float>> out;
        for (uint32_t i = 0; i < count; ++i) {
            NodeId id; id.bytes = ids[i];
            out.emplace_back("nodes:" + id.to_string(), 1.0f - scores[i]);
        }
        return out;
    }
    std::optional<std::vector<float>> get_node_embedding(const NodeId& id) override { return store_->get_embedding(id); }
    std::string last_column;

private:
    std::shared_ptr<B200VectorStore> store_;
};

// SurrealVectorStore (surreal_store.rs:24-90): VectorStore over any backend; parses "table:<uuid>" ids (:123-128).
class SurrealVectorStore : public VectorStore {
public:
    SurrealVectorStore(std::shared_ptr<SurrealVectorBackend> backend, size_t ef_search) : backend_(std::move(backend)), ef_(ef_search) {}
    void store_embeddings(const std::vector<CodeNode>& nodes) override { backend_->upsert_nodes(nodes); }
    std::vector<NodeId> search_similar(const std::vector<float>& q, size_t limit) const override {
        if (q.empty() || limit == 0) return {};
        auto neighbors = backend_->vector_knn(column_for_dimension(q.size()), q, limit, ef_);
        std::vector<NodeId> out;
        for (auto& [raw, dist] : neighbors) {
            (void)dist;
            auto pos = raw.rfind(':');
            std::string tail = pos == std::string::npos ? raw : raw.substr(pos + 1);
            auto id = NodeId::parse_str(tail);
            if (!id) throw Error(CGVEC_ERR_BAD_ARG, "Invalid node id '" + raw + "' returned by Surreal search");
            out.push_back(*id);
        }
        return out;
    }
    std::optional<std::vector<float>> get_embedding(const NodeId& id) const override { return backend_->get_node_embedding(id); }
    // surreal_embedding_column_for_dimension (codegraph-graph/src/surrealdb_storage.rs:1933-1954)
    static std::string column_for_dimension(size_t d) {
        switch (d) {
            case 384: case 768: case 1024: case 1536: case 2048: case 2560: case 3072: case 4096: return "embedding_" + std::to_string(d);
            default: return "embedding_2048";
        }
    }

private:
    std::shared_ptr<SurrealVectorBackend> backend_;
    size_t ef_;
};

// AI symbol resolver arg-max (codegraph-mcp/src/indexer.rs:2827-2843; cosine :2965-2979 == search.rs:519-533): among the
// candidates the caller's trigram prefilter kept, the first one with the highest cosine strictly above the threshold.
inline std::optional<std::pair<NodeId, float>> resolve_symbol(const B200VectorStore& store, const std::vector<float>& target_embedding,
                                                              const std::vector<NodeId>& candidates, float threshold = 0.75f) {
    if (candidates.empty()) return std::nullopt;
    std::vector<uint64_t> rows(candidates.size());
    for (size_t i = 0; i < candidates.size(); ++i) check(cgvec_row_of_id(store.handle(), candidates[i].bytes.data(), &rows[i]));
    std::vector<float> sims(rows.size());
    check(cgvec_rescore(store.handle(), target_embedding.data(), rows.data(), (uint32_t)rows.size(), CGVEC_COSINE, CGVEC_FORMULA_SEQ, sims.data()));
    std::optional<std::pair<NodeId, float>> best;
    for (size_t i = 0; i < rows.size(); ++i)
        if (sims[i] > threshold && (!best || sims[i] > best->second)) best = std::make_pair(candidates[i], sims[i]);
    return best;
}

// GpuAcceleration (codegraph-vector/src/gpu.rs:109-381): the reference ships this API as a mock (upload sleeps, distances
// are i*0.1 + query[0]*0.01); here the same two calls do the real thing.
class GpuVectorData {
public:
    GpuVectorData(cgvec_index* idx, size_t count, size_t dimension) : idx_(idx), count_(count), dim_(dimension) {}
    ~GpuVectorData() { cgvec_destroy(idx_); }
    GpuVectorData(const GpuVectorData&) = delete;
    GpuVectorData& operator=(const GpuVectorData&) = delete;
    size_t vector_count() const { return count_; }
    size_t dimension() const { return dim_; }
    bool is_uploaded() const { return true; }
    cgvec_index* handle() const { return idx_; }
private:
    cgvec_index* idx_;
    size_t count_, dim_;
};

class GpuAcceleration {
public:
    explicit GpuAcceleration(int device = 0) : device_(device) {}
    // upload_vectors(&[f32] flat, dimension) (gpu.rs:221-246)
    std::unique_ptr<GpuVectorData> upload_vectors(const std::vector<float>& vectors, size_t dimension) const {
        if (dimension == 0 || vectors.size() % dimension != 0) throw Error(CGVEC_ERR_BAD_ARG, "Vector data length not divisible by dimension");
        cgvec_index* idx = nullptr;
        int dev = device_;
        check(cgvec_create((uint32_t)dimension, CGVEC_F32, &dev, 1, &idx));
        const size_t n = vectors.size() / dimension;
        int rc = cgvec_add(idx, nullptr, vectors.data(), n);
        if (rc) { std::string msg = cgvec_last_error(); cgvec_destroy(idx); throw Error(rc, msg); }
        return std::make_unique<GpuVectorData>(idx, n, dimension);
    }
    // compute_distances(query, data, limit) with the semantics of its CPU twin compute_distances_cpu (gpu.rs:297-322):
    // cosine distance of the first `limit` uploaded vectors
    std::vector<float> compute_distances(const std::vector<float>& query, const GpuVectorData& data, size_t limit) const {
        if (query.size() != data.dimension())
            throw Error(CGVEC_ERR_BAD_DIM, "Query dimension " + std::to_string(query.size()) + " doesn't match GPU data dimension " + std::to_string(data.dimension()));
        std::vector<float> out(std::min(limit, data.vector_count()));
        uint64_t n = 0;
        check(cgvec_distances_first(data.handle(), query.data(), limit, out.data(), &n));
        out.resize(n);
        return out;
    }
private:
    int device_;
};

struct SearchResult {
    NodeId node_id;
    float score;
};

// SemanticSearch::search_by_embedding (search.rs:91-144) over a B200VectorStore: over-fetch max(3k, k+10) ids through
// the trait, exact re-score of each (search.rs:519-533 arithmetic, on the device), stable sort descending
// (NaN == Equal, :132-136), truncate, min-max normalise (:574-592).
class SemanticSearch {
public:
    explicit SemanticSearch(std::shared_ptr<B200VectorStore> store) : store_(std::move(store)) {}
    std::vector<SearchResult> search_by_embedding(const std::vector<float>& q, size_t limit) const {
        const size_t prefetch_k = (size_t)cgvec_prefetch_k_basic(limit);
        std::vector<NodeId> ids = store_->search_similar(q, prefetch_k);
        if (ids.empty()) return {};
        std::vector<uint64_t> rows(ids.size());
        for (size_t i = 0; i < ids.size(); ++i) check(cgvec_row_of_id(store_->handle(), ids[i].bytes.data(), &rows[i]));
        std::vector<float> raw(ids.size());
        check(cgvec_rescore(store_->handle(), q.data(), rows.data(), (uint32_t)rows.size(), CGVEC_COSINE, CGVEC_FORMULA_SEQ, raw.data()));
        std::vector<size_t> order(ids.size());
        for (size_t i = 0; i < order.size(); ++i) order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return raw[a] > raw[b]; });
        if (order.size() > limit) order.resize(limit);
        std::vector<float> s(order.size());
        for (size_t i = 0; i < order.size(); ++i) s[i] = raw[order[i]];
        cgvec_normalize_scores(s.data(), s.size());
        std::vector<SearchResult> out(order.size());
        for (size_t i = 0; i < order.size(); ++i) out[i] = {ids[order[i]], s[i]};
        return out;
    }

private:
    std::shared_ptr<B200VectorStore> store_;
};

// Steps 1-2 of fn::semantic_search_nodes_via_chunks (schema/codegraph.surql:318-417), the vector stage behind
// execute_semantic_code_search (codegraph-mcp-tools/src/graph_tool_executor.rs:578-591): the 100 nearest chunk embeddings
// (`<|100,200|>`), minus chunks without a parent node, cut to 3 x safe_limit rows (safe_limit = limit when 1 <= limit <= 100,
// else 10), each mapped to (parent node, 1 - cosine distance).  The KNN is this library's exact scan (B200Backend::vector_knn)
// in place of SurrealDB's approximate HNSW; BM25, the 0.9/0.1 blend and the graph enrichment stay in SurrealDB.
struct ChunkRecord {
    NodeId id;
    std::optional<NodeId> parent_node;
    std::vector<float> embedding;
};
struct ChunkCandidate {
    NodeId node_id;        // <string> parent_node AS node_id (codegraph.surql:403)
    float vector_score;    // 1f - distance (:412)
    NodeId chunk_id;
};
class ChunkCandidateStage {
public:
    static constexpr size_t kKnn = 100;                        // the literal of `<|100,200|>`
    explicit ChunkCandidateStage(uint32_t dimension, cgvec_dtype storage = CGVEC_F32, int device = 0)
        : store_(std::make_shared<B200VectorStore>(dimension, storage, device)), backend_(std::make_shared<B200Backend>(store_)) {}
    void upsert_chunks(const std::vector<ChunkRecord>& chunks) {
        std::vector<CodeNode> nodes;
        for (auto& c : chunks) nodes.push_back({c.id, c.embedding});
        backend_->upsert_nodes(nodes);
        for (auto& c : chunks) parent_[c.id.to_string()] = c.parent_node;
    }
    std::vector<ChunkCandidate> candidates(const std::vector<float>& query_embedding, long limit) const {
        const size_t safe_limit = (limit > 0 && limit <= 100) ? (size_t)limit : 10;      // codegraph.surql:325
        const size_t chunk_limit = safe_limit * 3;                                        // :326
        auto hits = backend_->vector_knn("embedding_" + std::to_string(store_->dimension()), query_embedding, kKnn, 200);
        std::vector<ChunkCandidate> out;
        for (auto& [raw, distance] : hits) {
            auto cid = NodeId::parse_str(raw.substr(raw.rfind(':') + 1));
            if (!cid) throw Error(CGVEC_ERR_BAD_ARG, "Invalid chunk id '" + raw + "'");
            auto it = parent_.find(cid->to_string());
            if (it == parent_.end() || !it->second) continue;                              // parent_node != NONE
            out.push_back({*it->second, 1.0f - distance, *cid});
            if (out.size() == chunk_limit) break;                                          // LIMIT $chunk_limit
        }
        return out;
    }
    std::shared_ptr<B200VectorStore> store() const { return store_; }

private:
    std::shared_ptr<B200VectorStore> store_;
    std::shared_ptr<B200Backend> backend_;
    std::unordered_map<std::string, std::optional<NodeId>> parent_;
};

}  // namespace cgvec
