// mock_backend_test.cpp — CPU-only replay of the reference's own unit tests for the SurrealVectorStore seam
// (crates/codegraph-vector/src/surreal_store.rs:130-206) against the C++ host mirror: a MockBackend records the column
// it was asked for and returns canned ("nodes:<uuid>", score) pairs.  No GPU call is made (the mirror's
// SurrealVectorStore only talks to the backend interface), so this runs in the CPU test tier.
#include <cassert>
#include <iostream>

#include "cgvec_host.hpp"

struct MockBackend : cgvec::SurrealVectorBackend {
    std::vector<std::pair<std::string, float>> results;
    std::vector<std::string> columns;
    explicit MockBackend(std::vector<std::pair<std::string, float>> r) : results(std::move(r)) {}
    void upsert_nodes(const std::vector<cgvec::CodeNode>&) override {}
    std::vector<std::pair<std::string, float>> vector_knn(const std::string& column, const std::vector<float>&, size_t, size_t) override {
        columns.push_back(column);
        return results;
    }
    std::optional<std::vector<float>> get_node_embedding(const cgvec::NodeId&) override { return std::nullopt; }
};

#define CHECK(cond) do { if (!(cond)) { std::cout << "FAIL " #cond " (line " << __LINE__ << ")\n"; return 1; } } while (0)

int main() {
    const std::string uuid = "018f3b7d-a82d-4f40-9127-2db4beefabcd";
    // strips_table_prefix_from_ids / keeps_clean_ids_intact (surreal_store.rs:138-149) through the public path
    {
        auto a = cgvec::NodeId::parse_str(uuid);
        CHECK(a && a->to_string() == uuid);
        CHECK(!cgvec::NodeId::parse_str("not-a-uuid"));
    }
    // search_similar_uses_surreal_backend (surreal_store.rs:151-165)
    {
        auto backend = std::make_shared<MockBackend>(std::vector<std::pair<std::string, float>>{{"nodes:" + uuid, 0.42f}});
        cgvec::SurrealVectorStore store(backend, 128);
        std::vector<float> embedding(2560, 0.0f);
        auto results = store.search_similar(embedding, 3);
        CHECK(results.size() == 1);
        CHECK(results[0].to_string() == uuid);
        CHECK(backend->columns.size() == 1 && backend->columns[0] == "embedding_2560");
        // clean ids (no table prefix) pass through, empty query / zero limit short-circuit (surreal_store.rs:62-64)
        backend->results = {{uuid, 0.1f}};
        CHECK(store.search_similar(embedding, 1)[0].to_string() == uuid);
        CHECK(store.search_similar({}, 3).empty());
        CHECK(store.search_similar(embedding, 0).empty());
        CHECK(backend->columns.size() == 2);
        // a malformed id surfaces as an error, not a crash (surreal_store.rs:75-81)
        backend->results = {{"nodes:zzz", 0.1f}};
        bool threw = false;
        try { store.search_similar(embedding, 1); } catch (const cgvec::Error&) { threw = true; }
        CHECK(threw);
        CHECK(!store.get_embedding(*cgvec::NodeId::parse_str(uuid)));
    }
    // surreal_embedding_column_for_dimension (codegraph-graph/src/surrealdb_storage.rs:1933-1954): unknown dims fall back to 2048
    CHECK(cgvec::SurrealVectorStore::column_for_dimension(384) == "embedding_384");
    CHECK(cgvec::SurrealVectorStore::column_for_dimension(4096) == "embedding_4096");
    CHECK(cgvec::SurrealVectorStore::column_for_dimension(123) == "embedding_2048");
    std::cout << "ok\n";
    return 0;
}
