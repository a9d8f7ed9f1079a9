// demo_main.cpp — drives the C++ host mirror end to end on the GPU and prints results for tests/test_host_cpp_gpu.py.
// Embeddings come from the reference's deterministic text embedding (search.rs:178-205: djb2 + LCG + L2 normalise),
// re-implemented here only to produce inputs the Python side can rebuild.
#include <cmath>
#include <iostream>

#include "cgvec_host.hpp"

static std::vector<float> hash_text_embedding(const std::string& text, size_t dimension) {
    uint32_t h = 5381u;
    for (unsigned char c : text) h = h * 33u + c;
    std::vector<float> e(dimension);
    uint32_t s = h;
    for (size_t i = 0; i < dimension; ++i) {
        s = s * 1103515245u + 12345u;
        e[i] = (((float)s / 4294967296.0f) - 0.5f) * 2.0f;
    }
    float norm = 0.0f;
    for (float x : e) norm = norm + x * x;
    norm = std::sqrt(norm);
    if (norm > 0.0f) for (float& x : e) x = x / norm;
    return e;
}

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? std::stoul(argv[1]) : 2000, dim = argc > 2 ? std::stoul(argv[2]) : 384, limit = argc > 3 ? std::stoul(argv[3]) : 7;
    const int n_devices = argc > 4 ? std::stoi(argv[4]) : 1;            // > 1: the single-process multi-device store
    try {
        std::vector<int> devices;
        for (int i = 0; i < n_devices; ++i) devices.push_back(i);
        auto store = n_devices > 1 ? std::make_shared<cgvec::B200VectorStore>((uint32_t)dim, devices)
                                   : std::make_shared<cgvec::B200VectorStore>((uint32_t)dim);
        std::vector<cgvec::CodeNode> nodes;
        for (size_t i = 0; i < n; ++i) nodes.push_back({cgvec::NodeId::from_u64(i + 1), hash_text_embedding("fn item_" + std::to_string(i) + "() {}", dim)});
        nodes.push_back({cgvec::NodeId::from_u64(999999), std::nullopt});          // skipped
        store->store_embeddings(nodes);
        std::cout << "len " << store->len() << "\n";
        auto q = hash_text_embedding("fn item_42() { }", dim);
        std::cout << "search_similar";
        for (auto& id : store->search_similar(q, limit)) std::cout << " " << id.to_string();
        std::cout << "\n";
        std::cout << "top_k";
        for (auto& [row, score] : store->top_k(q, limit)) std::cout << " " << row << ":" << std::hexfloat << score << std::defaultfloat;
        std::cout << "\n";
        auto backend = std::make_shared<cgvec::B200Backend>(store);
        cgvec::SurrealVectorStore surreal(backend, 100);
        std::cout << "surreal";
        for (auto& id : surreal.search_similar(q, limit)) std::cout << " " << id.to_string();
        std::cout << "\ncolumn " << backend->last_column << "\n";
        cgvec::SemanticSearch sem(store);
        std::cout << "semantic";
        for (auto& r : sem.search_by_embedding(q, limit)) std::cout << " " << r.node_id.to_string() << ":" << std::hexfloat << r.score << std::defaultfloat;
        std::cout << "\n";
        {
            cgvec::GpuAcceleration gpu(0);
            std::vector<float> flat;
            for (size_t i = 0; i < 20; ++i) { auto e = hash_text_embedding("v" + std::to_string(i), dim); flat.insert(flat.end(), e.begin(), e.end()); }
            auto data = gpu.upload_vectors(flat, dim);
            std::cout << "gpu_distances";
            for (float x : gpu.compute_distances(q, *data, 5)) std::cout << " " << std::hexfloat << x << std::defaultfloat;
            std::cout << "\n";
        }
        {   // candidate stage of semantic_code_search: chunk i belongs to node i / 3; every 5th chunk is an orphan
            cgvec::ChunkCandidateStage stage((uint32_t)dim);
            std::vector<cgvec::ChunkRecord> chunks;
            for (size_t i = 0; i < n; ++i) {
                std::optional<cgvec::NodeId> parent;
                if (i % 5 != 4) parent = cgvec::NodeId::from_u64(1000000 + i / 3);
                chunks.push_back({cgvec::NodeId::from_u64(i + 1), parent, hash_text_embedding("fn item_" + std::to_string(i) + "() {}", dim)});
            }
            stage.upsert_chunks(chunks);
            std::cout << "candidates";
            for (auto& c : stage.candidates(q, (long)limit))
                std::cout << " " << c.node_id.to_string() << "/" << c.chunk_id.to_string() << "/" << std::hexfloat << c.vector_score << std::defaultfloat;
            std::cout << "\ncandidates_bad_limit " << stage.candidates(q, 0).size() << " " << stage.candidates(q, 1000).size() << "\n";
        }
        if (n_devices == 1) {   // the same queries through a resident session (kernel stays on the GPU between the calls)
            cgvec::ResidentSession session(store, limit);
            session.top_k(q);
            std::cout << "session_top_k";
            for (auto& [row, score] : session.top_k(q)) std::cout << " " << row << ":" << std::hexfloat << score << std::defaultfloat;
            std::cout << "\nsession_similar";
            for (auto& id : session.search_similar(q)) std::cout << " " << id.to_string();
            std::cout << "\n";
        }
        auto missing = store->get_embedding(cgvec::NodeId::from_u64(123456789));
        std::cout << "missing " << (missing ? "some" : "none") << "\n";
        try {
            store->search_similar(std::vector<float>(dim + 1, 1.0f), 3);
            std::cout << "baddim no-error\n";
        } catch (const cgvec::Error& e) { std::cout << "baddim " << e.code << "\n"; }
    } catch (const cgvec::Error& e) {
        std::cout << "error " << e.code << " " << e.what() << "\n";
        return e.code == CGVEC_ERR_NO_DEVICE ? 3 : 1;
    }
    return 0;
}
