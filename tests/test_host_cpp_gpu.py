"""The compiled C++ host mirror (codegraph-rust_b200/host/cgvec_host.hpp: VectorStore, SurrealVectorBackend,
SurrealVectorStore, SemanticSearch over the C ABI) driven end to end on the GPU and checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "codegraph-rust_b200", "host", "cgvec_host_demo")


def test_host_mirror_compiles_and_refuses_without_gpu(cg):
    """CPU-side: the header-only mirror builds with g++ against the C ABI; without a device it reports NO_DEVICE."""
    cg._build.build_host_demo()
    assert os.path.exists(DEMO)
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        p = subprocess.run([DEMO, "10", "64", "3"], capture_output=True, text=True)
        assert p.returncode == 3 and "error -7" in p.stdout


@pytest.mark.gpu
def test_host_mirror_matches_oracle(cg, oracle):
    cg._build.build_host_demo()
    n, dim, limit = 2000, 384, 7
    out = subprocess.run([DEMO, str(n), str(dim), str(limit)], capture_output=True, text=True, check=True).stdout
    lines = dict(l.split(" ", 1) for l in out.strip().splitlines())
    embs = np.stack([oracle.hash_text_embedding(f"fn item_{i}() {{}}", dim) for i in range(n)])
    q = oracle.hash_text_embedding("fn item_42() { }", dim)
    assert lines["len"] == str(n)
    wi, ws = oracle.parallel_top_k_search(q, embs, limit)
    want_ids = [f"00000000-0000-0000-0000-{int(i) + 1:012x}" for i in wi]
    assert lines["search_similar"].split() == want_ids
    assert lines["surreal"].split() == want_ids
    assert lines["column"] == "embedding_384"
    got = [t.split(":") for t in lines["top_k"].split()]
    assert [int(r) for r, _ in got] == wi.tolist()
    assert np.float32([float.fromhex(s) for _, s in got]).tobytes() == ws.tobytes()
    si, snorm, _ = oracle.search_by_embedding(q, embs, limit)
    sem = [t.rsplit(":", 1) for t in lines["semantic"].split()]
    assert [u for u, _ in sem] == [f"00000000-0000-0000-0000-{int(i) + 1:012x}" for i in si]
    assert np.float32([float.fromhex(s) for _, s in sem]).tobytes() == snorm.tobytes()
    flat = np.stack([oracle.hash_text_embedding(f"v{i}", dim) for i in range(20)])
    want_d = oracle.compute_distances_cpu(q, flat.reshape(-1), dim, 5)
    assert np.float32([float.fromhex(x) for x in lines["gpu_distances"].split()]).tobytes() == want_d.tobytes()
    # candidate stage (codegraph.surql:318-417): 100 nearest chunks, orphans dropped, LIMIT 3 * safe_limit, 1 - distance
    w100, s100 = oracle.parallel_top_k_search(q, embs, 100)
    want_c = [(int(i), np.float32(1.0) - (np.float32(1.0) - s)) for i, s in zip(w100, s100) if int(i) % 5 != 4][: 3 * limit]
    got_c = [t.split("/") for t in lines["candidates"].split()]
    assert [c[1] for c in got_c] == [f"00000000-0000-0000-0000-{i + 1:012x}" for i, _ in want_c]
    assert [c[0] for c in got_c] == [f"00000000-0000-0000-0000-{1000000 + i // 3:012x}" for i, _ in want_c]
    assert np.float32([float.fromhex(c[2]) for c in got_c]).tobytes() == np.float32([s for _, s in want_c]).tobytes()
    assert lines["candidates_bad_limit"].split() == ["30", "30"]                   # limit outside 1..100 -> safe_limit 10
    # ResidentSession (cgvec_serve_*) gives the same answers as the launch-per-query calls
    assert lines["session_top_k"] == lines["top_k"]
    assert lines["session_similar"] == lines["search_similar"]
    assert lines["missing"] == "none"
    assert lines["baddim"] == str(cg.ERR_BAD_DIM)


def test_reference_mock_backend_unit_tests_on_the_cpp_mirror(cg):
    """surreal_store.rs:130-206 (MockBackend: id normalisation, embedding_2560 column, short-circuits) replayed on the
    C++ SurrealVectorStore mirror — CPU only."""
    exe = cg._build.build_mock_test()
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.strip() == "ok", p.stdout + p.stderr
