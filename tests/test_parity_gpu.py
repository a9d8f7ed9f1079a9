"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.
Bar: indices bit-exact; SIMD-formula scores bit-identical (stricter than north_star's 1e-5 relative)."""
import uuid

import numpy as np
import pytest

from tests import synth

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5      # north_star's tolerance for fp32 distances; SIMD-formula scores are checked bit-for-bit instead


def _metric(cg, oracle, name):
    return {"cosine": (cg.COSINE, oracle.COSINE), "dot": (cg.DOT, oracle.DOT), "l2": (cg.L2, oracle.L2)}[name]


def _check_exact(cg, oracle, rows, queries, k, metric="cosine", dtype=None):
    dtype = cg.F32 if dtype is None else dtype
    gm, om = _metric(cg, oracle, metric)
    ix = cg.Index(rows.shape[1], dtype)
    try:
        ix.add(rows)
        ref_rows = rows if dtype == cg.F32 else rows.astype(np.float16).astype(np.float32)
        got_r, got_s, cnt = ix.search(queries, k, gm)
        q2 = np.atleast_2d(queries)
        for qi in range(q2.shape[0]):
            wi, ws = oracle.parallel_top_k_search(q2[qi], ref_rows, k, metric=om)
            assert int(cnt[qi]) == len(wi) == min(k, rows.shape[0])
            assert got_r[qi, :len(wi)].tolist() == wi.tolist(), (metric, qi, got_r[qi, :len(wi)], wi)
            assert np.array_equal(got_s[qi, :len(wi)], ws, equal_nan=True), (metric, qi, got_s[qi, :len(wi)], ws)
    finally:
        ix.close()


@pytest.mark.parametrize("d", [1, 7, 8, 24, 31, 32, 33, 100, 384, 768, 1024, 1027])
def test_dimensions_incl_ragged_tails(cg, oracle, d):
    """adaptive_cosine_similarity switches at len 32 (simd_ops.rs:284); d % 8 != 0 exercises the scalar tail (:59-65)."""
    rng = np.random.default_rng(d)
    rows = rng.standard_normal((1500, d)).astype(np.float32)
    q = rng.standard_normal(d).astype(np.float32)
    for metric in ("cosine", "dot", "l2"):
        _check_exact(cg, oracle, rows, q, 10, metric)


@pytest.mark.parametrize("n,k", [(1, 1), (1, 10), (3, 10), (17, 17), (33, 5), (1000, 1), (1000, 100), (5000, 1024), (20000, 10)])
def test_row_counts_and_k(cg, oracle, n, k):
    rng = np.random.default_rng(n * 31 + k)
    rows = rng.standard_normal((n, 64)).astype(np.float32)
    q = rng.standard_normal(64).astype(np.float32)
    _check_exact(cg, oracle, rows, q, k, "cosine")


def test_reference_test_parallel_operations(cg, oracle):
    """simd_ops.rs:461-472: q=[1.0;256], row_i[j]=i+j, N=1000, k=10 -> forced answer [999..990]."""
    q = np.ones(256, np.float32)
    rows = (np.arange(1000)[:, None] + np.arange(256)[None, :]).astype(np.float32)
    res = cg.ParallelVectorOps.parallel_top_k_search(q, rows, 10)
    assert len(res) == 10
    assert [i for i, _ in res] == list(range(999, 989, -1))
    wi, ws = oracle.parallel_top_k_search(q, rows, 10)
    assert np.array_equal(np.float32([s for _, s in res]), ws)


def test_config1_10k_x_768(cg, oracle):
    """BASELINE config 1: 10k x 768 f32, 1 query, top-10 cosine."""
    rows = synth.synth_rows(0xC0DE6A9F, np.arange(10_000), 768)
    q = synth.synth_rows(0x5EED0001, [0], 768)[0]
    _check_exact(cg, oracle, rows, q, 10, "cosine")


@pytest.mark.parametrize("nq", [1, 2, 3, 4, 5, 7, 9])
def test_query_batches(cg, oracle, nq):
    rng = np.random.default_rng(nq)
    rows = rng.standard_normal((3000, 96)).astype(np.float32)
    qs = rng.standard_normal((nq, 96)).astype(np.float32)
    for metric in ("cosine", "l2"):
        _check_exact(cg, oracle, rows, qs, 12, metric)


@pytest.mark.parametrize("d", [64, 768, 1000])
def test_fp16_storage_oracle_is_widened_rows(cg, oracle, d):
    rng = np.random.default_rng(d + 1)
    rows = (rng.standard_normal((4000, d)) / np.sqrt(d)).astype(np.float32)
    qs = rng.standard_normal((2, d)).astype(np.float32)
    for metric in ("cosine", "dot", "l2"):
        _check_exact(cg, oracle, rows, qs, 10, metric, dtype=cg.F16)


def test_tie_nan_zero_contract(cg, oracle):
    """ties -> lower row; NaN last; zero-norm row -> 0.0; k > N -> N; k == 0 -> empty."""
    rows = np.zeros((6, 32), np.float32)
    rows[:, 0] = 1.0
    rows[4] = 0.0
    rows[2, 1] = np.nan
    q = np.zeros(32, np.float32); q[0] = 1.0
    ix = cg.Index(32)
    ix.add(rows)
    r, s, c = ix.search(q, 10)
    assert int(c[0]) == 6
    assert r[0, :6].tolist() == [0, 1, 3, 5, 4, 2]
    assert np.isnan(s[0, 5]) and s[0, 4] == 0.0
    wi, ws = oracle.parallel_top_k_search(q, rows, 10)
    assert r[0, :6].tolist() == wi.tolist() and np.array_equal(s[0, :6], ws, equal_nan=True)
    r0, s0, c0 = ix.search(q, 0)
    assert int(c0[0]) == 0
    # zero query -> every score 0.0 -> rows in index order
    r1, s1, c1 = ix.search(np.zeros(32, np.float32), 3)
    assert r1[0].tolist() == [0, 1, 3] or r1[0].tolist() == oracle.parallel_top_k_search(np.zeros(32, np.float32), rows, 3)[0].tolist()
    ix.close()


def test_many_exact_ties_across_ctas(cg, oracle):
    """Duplicate rows everywhere: the k winners must be the k lowest row indices among equals."""
    rng = np.random.default_rng(2)
    base = rng.standard_normal((4, 128)).astype(np.float32)
    rows = np.repeat(base, 5000, axis=0)[rng.permutation(20000)]
    q = base[1]
    _check_exact(cg, oracle, rows, q, 64, "cosine")
    _check_exact(cg, oracle, rows, q, 64, "l2")


def test_formulas_rescore_bit_exact(cg, oracle):
    """cgvec_rescore reproduces each reference formula bit-for-bit (simd_ops.rs:257-278, search.rs:519-533,
    optimization.rs:404-418)."""
    rng = np.random.default_rng(8)
    rows = rng.standard_normal((300, 200)).astype(np.float32)
    rows[7] = 0.0
    q = rng.standard_normal(200).astype(np.float32)
    ix = cg.Index(200)
    ix.add(rows)
    sel = np.arange(300, dtype=np.uint64)
    got = ix.rescore(q, sel, cg.COSINE, cg.FORMULA_SCALAR)
    assert got.tobytes() == np.float32([oracle.cosine_similarity_scalar(q, r) for r in rows]).tobytes()
    got = ix.rescore(q, sel, cg.COSINE, cg.FORMULA_SEQ)
    assert got.tobytes() == np.float32([oracle.cosine_similarity_seq(q, r) for r in rows]).tobytes()
    got = ix.rescore(q, sel, cg.COSINE, cg.FORMULA_BASELINE)
    assert got.tobytes() == np.float32([oracle.cosine_distance_seq(q, r) for r in rows]).tobytes()
    got = ix.rescore(q, sel, cg.COSINE, cg.FORMULA_SIMD)
    assert got.tobytes() == oracle.scores(q, rows).tobytes()
    got = ix.rescore(q, sel, cg.L2, cg.FORMULA_SIMD)
    assert got.tobytes() == oracle.scores(q, rows, metric=oracle.L2).tobytes()
    ix.close()


def test_search_baseline_formula_on_reference_vectors(cg, oracle):
    """model_optimization_tests.rs:36-58,383-424 vectors (N=1000, d=128, seed 11223, query=row 0):
    formula BASELINE == ModelOptimizer::search_baseline (optimization.rs:376-402): same indices, same distances."""
    vecs = oracle.generate_optimization_vectors(1000, 128, 11223)
    q = vecs[0]
    want_i, want_d = oracle.search_baseline(q, vecs, 10)
    ix = cg.Index(128)
    ix.add(vecs)
    r, s, c = ix.search(q, 10, cg.COSINE, formula=cg.FORMULA_BASELINE)
    assert int(c[0]) == 10 and r[0, 0] == 0
    assert r[0].tolist() == want_i.tolist()
    assert s[0].tobytes() == want_d.tobytes()
    # and the trait-level semantics of InMemoryVectorStore (graph_vector.rs:479-494) via formula SEQ
    wi, ws = oracle.inmemory_search_similar(q, vecs, 25)
    r2, s2, _ = ix.search(q, 25, cg.COSINE, formula=cg.FORMULA_SEQ)
    assert r2[0].tolist() == wi.tolist() and s2[0].tobytes() == ws.tobytes()
    ix.close()


def test_vector_store_trait_mirror(cg, oracle):
    """trait VectorStore round trip (traits.rs:11-16) incl. skip-None, upsert-by-id and get_embedding -> None."""
    rng = np.random.default_rng(12)
    embs = rng.standard_normal((200, 384)).astype(np.float32)
    nodes = [cg.CodeNode(uuid.UUID(int=i + 1), embs[i].tolist()) for i in range(200)]
    nodes.insert(50, cg.CodeNode(uuid.UUID(int=10_000), None))                 # no embedding -> skipped
    store = cg.B200VectorStore(384)
    store.store_embeddings(nodes)
    assert len(store.index) == 200
    q = embs[17] + 0.05 * rng.standard_normal(384).astype(np.float32)
    ids = store.search_similar(q, 5)
    wi, _ = oracle.parallel_top_k_search(q, embs, 5)
    assert [i.int - 1 for i in ids] == wi.tolist()
    assert ids[0] == uuid.UUID(int=18)
    assert store.search_similar(q, 0) == [] and store.search_similar([], 5) == []
    assert store.get_embedding(uuid.UUID(int=999_999)) is None
    assert np.float32(store.get_embedding(uuid.UUID(int=18))).tobytes() == embs[17].tobytes()
    # upsert: same id, new embedding -> overwritten in place, still 200 rows
    store.store_embeddings([cg.CodeNode(uuid.UUID(int=18), (-embs[17]).tolist())])
    assert len(store.index) == 200
    assert store.search_similar(q, 1)[0] != uuid.UUID(int=18)
    # backend seam: ("nodes:<uuid>", cosine distance) ascending
    be = cg.B200Backend(store)
    knn = be.vector_knn("embedding_384", q, 4, 100)
    assert be.last_column == "embedding_384" and len(knn) == 4
    assert all(k.startswith("nodes:") for k, _ in knn)
    assert [d for _, d in knn] == sorted(d for _, d in knn)


def test_semantic_search_by_embedding_mirror(cg, oracle):
    """search.rs:91-144: over-fetch max(3k, k+10), exact rescore, stable sort, truncate, min-max normalise."""
    rng = np.random.default_rng(13)
    embs = rng.standard_normal((500, 128)).astype(np.float32)
    store = cg.B200VectorStore(128)
    store.store_embeddings([cg.CodeNode(uuid.UUID(int=i + 1), embs[i]) for i in range(500)])
    q = rng.standard_normal(128).astype(np.float32)
    res = cg.SemanticSearch(store).search_by_embedding(q, 7)
    wi, wnorm, _ = oracle.search_by_embedding(q, embs, 7)
    assert [r.node_id.int - 1 for r in res] == wi.tolist()
    assert np.float32([r.score for r in res]).tobytes() == wnorm.tobytes()
    assert res[0].score == 1.0 and res[-1].score == 0.0


def test_gpu_acceleration_api_mirror(cg, oracle):
    """gpu.rs:221-322: upload_vectors(flat, dim) + compute_distances == compute_distances_cpu (first `limit` rows)."""
    rng = np.random.default_rng(14)
    flat = rng.standard_normal(40 * 96).astype(np.float32)
    q = rng.standard_normal(96).astype(np.float32)
    gpu = cg.GpuAcceleration()
    data = gpu.upload_vectors(flat, 96)
    got = gpu.compute_distances(q, data, 25)
    assert got.tobytes() == oracle.compute_distances_cpu(q, flat, 96, 25).tobytes()
    assert len(gpu.compute_distances(q, data, 1000)) == 40
    with pytest.raises(cg.CgvecError):
        gpu.upload_vectors(flat[:-1], 96)
    with pytest.raises(cg.CgvecError):
        gpu.compute_distances(q[:10], data, 5)


def test_normalize_rows_matches_normalize_avx2(cg, oracle):
    """parallel_normalize_vectors (simd_ops.rs:386-419) -> normalize_avx2 (:189-222), bit-exact, zero rows untouched."""
    rng = np.random.default_rng(15)
    rows = (rng.standard_normal((100, 77)) * 3).astype(np.float32)
    rows[5] = 0.0
    ix = cg.Index(77)
    ix.add(rows)
    ix.normalize_rows()
    got = ix.get_rows(0, 100)
    want = np.stack([oracle.normalize_avx2(r) for r in rows])
    assert got.tobytes() == want.tobytes()
    ix.close()


def test_device_synthetic_rows_equal_host_mirror(cg):
    for dtype, f16 in ((cg.F32, False), (cg.F16, True)):
        ix = cg.Index(768, dtype)
        ix.fill_synthetic(3000, 0xC0DE6A9F, True)
        got = ix.get_rows(0, 3000)
        want = synth.synth_rows(0xC0DE6A9F, np.arange(3000), 768, True, f16)
        assert got.tobytes() == want.tobytes()
        ix.close()


def test_errors_cross_the_boundary_as_codes(cg):
    ix = cg.Index(16)
    with pytest.raises(cg.CgvecError) as e:
        ix.add(np.zeros((3, 15), np.float32))
    assert e.value.code == cg.ERR_BAD_DIM
    with pytest.raises(cg.CgvecError) as e:
        ix.search(np.zeros(17, np.float32), 3)
    assert e.value.code == cg.ERR_BAD_DIM
    r, s, c = ix.search(np.zeros(16, np.float32), 3)      # empty index -> empty result, not an error
    assert int(c[0]) == 0
    ix.add(np.ones((4, 16), np.float32))
    with pytest.raises(cg.CgvecError) as e:
        ix.search(np.ones(16, np.float32), 5000)
    assert e.value.code == cg.ERR_UNSUPPORTED
    assert ix.get_row(2).tolist() == [1.0] * 16
    with pytest.raises(cg.CgvecError):
        ix.get_row(99)
    ix.close()


def test_concurrent_searches_are_reentrant(cg, oracle):
    """multi_vector_search fans out concurrent &self calls (search.rs:358-361)."""
    import threading
    rng = np.random.default_rng(16)
    rows = rng.standard_normal((20000, 64)).astype(np.float32)
    qs = rng.standard_normal((16, 64)).astype(np.float32)
    ix = cg.Index(64)
    ix.add(rows)
    want = [oracle.parallel_top_k_search(q, rows, 10)[0].tolist() for q in qs]
    got = [None] * 16

    def work(i):
        for _ in range(5):
            got[i] = ix.search(qs[i], 10)[0][0].tolist()
    th = [threading.Thread(target=work, args=(i,)) for i in range(16)]
    [t.start() for t in th]; [t.join() for t in th]
    assert got == want
    ix.close()


def test_concurrent_batch1_callers_are_coalesced(cg, oracle):
    """Group commit (SURVEY §8b Threading; multi_vector_search, search.rs:347-361): 16 threads issuing batch-1 searches against
    one index are served by shared multi-query passes — every caller still gets exactly the oracle's answer (rows, score bytes,
    ids), mixed k / metric requests are never merged, and the pass count drops below one per caller."""
    import threading, uuid
    rng = np.random.default_rng(116)
    n, d = 300_000, 128
    rows = rng.standard_normal((n, d)).astype(np.float32)
    ids = [uuid.UUID(int=i + 1) for i in range(2000)]
    qs = rng.standard_normal((16, d)).astype(np.float32)
    ix = cg.Index(d)
    ix.add(rows[:2000], ids); ix.add(rows[2000:])
    ks = [10, 10, 10, 25]                                       # thread i uses k = ks[i % 4]; thread 7 uses L2
    want = []
    for i in range(16):
        metric = oracle.L2 if i == 7 else oracle.COSINE
        want.append(oracle.parallel_top_k_search(qs[i], rows, ks[i % 4], metric=metric))
    reps = 12
    got = [[None] * reps for _ in range(16)]
    start = threading.Barrier(16)

    def work(i):
        start.wait()
        for r in range(reps):
            got[i][r] = ix.search(qs[i], ks[i % 4], cg.L2 if i == 7 else cg.COSINE, want_ids=True)
    th = [threading.Thread(target=work, args=(i,)) for i in range(16)]
    [t.start() for t in th]; [t.join() for t in th]
    for i in range(16):
        wi, ws = want[i]
        for r in range(reps):
            gr, gs, gc, gids = got[i][r]
            assert gr[0].tolist() == wi.tolist(), (i, r)
            assert gs[0].tobytes() == ws.tobytes(), (i, r)
            assert int(gc[0]) == len(wi)
            for j, row in enumerate(wi):
                want_id = ids[int(row)].bytes if int(row) < 2000 else bytes(16)
                assert gids[0, j].tobytes() == want_id
    st = ix.stats()
    assert st.coalesced_batches > 0 and st.coalesced_queries >= 2 * st.coalesced_batches
    # switched off: every caller scans on its own again, same answers
    ix.set_option("coalesce", 0)
    b0 = ix.stats().coalesced_batches
    th = [threading.Thread(target=work, args=(i,)) for i in range(16)]
    start.reset()
    [t.start() for t in th]; [t.join() for t in th]
    assert ix.stats().coalesced_batches == b0
    for i in range(16):
        assert got[i][-1][0][0].tolist() == want[i][0].tolist()
    ix.close()


def test_config2_full_size_1m_x_768(cg, oracle):
    """BASELINE config 2 at full size: 1M x 768 f32, batch-1, top-10, checked against the oracle on the rows
    read back from HBM; plus size-independent properties (planted neighbour, self-match, idempotence)."""
    n, d = 1_000_000, 768
    ix = cg.Index(d)
    ix.fill_synthetic(n, 0xC0DE6A9F, True)
    assert len(ix) == n
    rows = ix.get_rows(0, n)
    q = synth.synth_rows(0x5EED0001, [0], d)[0]
    r, s, c = ix.search(q, 10)
    wi, ws = oracle.parallel_top_k_search(q, rows, 10)
    assert r[0].tolist() == wi.tolist()
    assert s[0].tobytes() == ws.tobytes()
    r2, s2, _ = ix.search(q, 10)
    assert r2.tobytes() == r.tobytes() and s2.tobytes() == s.tobytes()          # idempotent
    rng = np.random.default_rng(99)
    for planted in (0, 123_456, n - 1):
        pq = rows[planted] + np.float32(0.05 / np.sqrt(d)) * rng.standard_normal(d).astype(np.float32)
        rr, ss, _ = ix.search(pq, 5)
        assert rr[0, 0] == planted and ss[0, 0] > 0.99
        rr, ss, _ = ix.search(rows[planted], 1, cg.L2)
        assert rr[0, 0] == planted and ss[0, 0] == 0.0
    # dot and L2 at full size against the oracle too
    for gm, om in ((cg.DOT, oracle.DOT), (cg.L2, oracle.L2)):
        r, s, _ = ix.search(q, 10, gm)
        wi, ws = oracle.parallel_top_k_search(q, rows, 10, metric=om)
        assert r[0].tolist() == wi.tolist() and s[0].tobytes() == ws.tobytes()
    st = ix.stats()
    assert st.grid == st.sm_count
    ix.close()


@pytest.mark.parametrize("d", [384, 1024, 1536, 2048, 4096])
def test_deep_pipeline_every_reference_dimension(cg, oracle, d):
    """The SurrealDB path whitelists these dims (surrealdb_storage.rs:1933-1954).  Enough rows that every CTA
    wraps its stage ring several times (mbarrier phases flip) whatever stage count the planner picks."""
    rng = np.random.default_rng(d)
    n = 148 * 16 * 10 + 7
    rows = rng.standard_normal((n, d)).astype(np.float32)
    q = rng.standard_normal(d).astype(np.float32)
    _check_exact(cg, oracle, rows, q, 10, "cosine")


@pytest.mark.parametrize("tile,stages", [(4, 8), (8, 4), (8, 8), (16, 2), (16, 4), (16, 6), (32, 2), (32, 3)])
def test_every_launch_geometry_is_exact(cg, oracle, tile, stages):
    rng = np.random.default_rng(tile * 10 + stages)
    n, d = 148 * 32 * 9 + 3, 256
    rows = rng.standard_normal((n, d)).astype(np.float32)
    qs = rng.standard_normal((2, d)).astype(np.float32)
    ix = cg.Index(d)
    ix.add(rows)
    ix.set_option("tile_rows", tile); ix.set_option("stages", stages)
    for hint in (0, 1):
        ix.set_option("l2_hint", hint)
        r, s, c = ix.search(qs, 33)
        st = ix.stats()
        assert st.tile_rows == tile and st.stages == stages
        for qi in range(2):
            wi, ws = oracle.parallel_top_k_search(qs[qi], rows, 33)
            assert r[qi].tolist() == wi.tolist() and s[qi].tobytes() == ws.tobytes()
    # a stage count that no full grouping divides still runs exactly (fewer active warp groups)
    for tile, stages in ((8, 7), (8, 3), (16, 3), (16, 5), (4, 6)):
        ix.set_option("tile_rows", tile); ix.set_option("stages", stages)
        r, s, c = ix.search(qs[0], 5)
        assert ix.stats().stages == stages
        wi, ws = oracle.parallel_top_k_search(qs[0], rows, 5)
        assert r[0].tolist() == wi.tolist() and s[0].tobytes() == ws.tobytes()
    ix.close()


def test_flat_matrix_file_round_trip(cg, oracle, tmp_path):
    """memory.rs:241-374: [u64 n][u64 d][n*d f32].  A file written the reference's way loads; a saved index reads back
    the reference's way; wrong dimension / truncated files are rejected like load_from_mmap does."""
    rng = np.random.default_rng(77)
    rows = rng.standard_normal((3000, 96)).astype(np.float32)
    ref_file = tmp_path / "ref.bin"
    with open(ref_file, "wb") as f:
        f.write(np.array([3000, 96], np.uint64).tobytes()); f.write(rows.tobytes())
    ix = cg.Index(96)
    assert ix.load_flat(str(ref_file)) == 3000 and len(ix) == 3000
    q = rng.standard_normal(96).astype(np.float32)
    wi, ws = oracle.parallel_top_k_search(q, rows, 10)
    r, s, _ = ix.search(q, 10)
    assert r[0].tolist() == wi.tolist() and s[0].tobytes() == ws.tobytes()
    out_file = tmp_path / "out.bin"
    ix.save_flat(str(out_file))
    assert open(out_file, "rb").read() == open(ref_file, "rb").read()
    bad = cg.Index(64)
    with pytest.raises(cg.CgvecError) as e:
        bad.load_flat(str(ref_file))
    assert e.value.code == cg.ERR_BAD_DIM
    trunc = tmp_path / "trunc.bin"
    trunc.write_bytes(open(ref_file, "rb").read()[:-4])
    with pytest.raises(cg.CgvecError):
        ix.load_flat(str(trunc))
    ix.close(); bad.close()


def test_int8_quantised_scan_matches_search_optimized(cg, oracle):
    """SURVEY §8f-3.  Codes == quantize_batch (optimization.rs:212-224,268-274) byte for byte; search_optimized
    (:63-150) scores bit-identical AND indices identical, ties included (the reference's running list with strict `>`
    replacement and stable sorts decides which equal scores stay and in what order; the library reproduces that order
    in closed form); and the reference's own property (model_optimization_tests.rs:383-424): >= 80 % position-wise
    agreement with search_baseline."""
    vecs = oracle.generate_optimization_vectors(1000, 128, 11223)
    ix = cg.Index(128)
    ix.add(vecs)
    ix.quantize_i8()
    want_codes = oracle.quantize_batch_u8(vecs)
    assert ix.codes_i8(0, 1000).tobytes() == want_codes.tobytes()
    q = vecs[0]
    got_i, got_s = ix.search_optimized(q, 10)
    want_i, want_s = oracle.search_optimized_i8(q, want_codes, 10)
    assert got_s.tobytes() == want_s.tobytes()
    assert got_i.tolist() == want_i.tolist()
    base, _ = oracle.search_baseline(q, vecs, 10)
    assert float(np.mean(got_i == base)) >= 0.8
    # larger, ragged dimension, values outside [-1, 1] (clamped), NaN element (-> code 128), zero row (skipped)
    rng = np.random.default_rng(5)
    rows = (rng.standard_normal((20_000, 100)) * 0.7).astype(np.float32)
    rows[17, 3] = np.nan; rows[99] = 0.0
    ix2 = cg.Index(100); ix2.add(rows); ix2.quantize_i8()
    codes = oracle.quantize_batch_u8(rows)
    assert ix2.codes_i8(0, 20_000).tobytes() == codes.tobytes()
    for limit in (1, 7, 50):
        q = rng.standard_normal(100).astype(np.float32) * 0.5
        gi, gs = ix2.search_optimized(q, limit)
        wi, ws = oracle.search_optimized_i8(q, codes, limit)
        assert gs.tobytes() == ws.tobytes()
        assert gi.tolist() == wi.tolist()
    assert len(ix2.search_optimized(np.zeros(100, np.float32), 5)[0]) == 0          # zero query -> empty (:113-115)
    ix.close(); ix2.close()


@pytest.mark.parametrize("seed", range(6))
def test_int8_scan_reproduces_the_reference_tie_order(cg, oracle, seed):
    """Ties are the hard part of integer parity: a tiny alphabet makes most int8 scores collide (and some rows all-zero,
    which search_optimized skips, :132-134); the reference's answer then depends on arrival order (:139-149).  Indices
    must still match exactly, for limits below, at and above the number of distinct scores and of valid rows."""
    rng = np.random.default_rng(100 + seed)
    n, d = [37, 300, 5000, 5000, 20_000, 64][seed], [4, 8, 8, 16, 32, 4][seed]
    levels = np.array([-2, -1, 0, 0, 1, 2], np.float32) / 127.0            # codes 126..130
    rows = rng.choice(levels, (n, d)).astype(np.float32)
    rows[rng.integers(0, n, max(n // 20, 1))] = 0.0                          # zero rows are skipped by the reference
    ix = cg.Index(d)
    try:
        ix.add(rows)
        ix.quantize_i8()
        codes = oracle.quantize_batch_u8(rows)
        assert ix.codes_i8(0, n).tobytes() == codes.tobytes()
        for limit in (1, 2, 5, 17, 60, 200):
            for _ in range(3):
                q = rng.choice(np.array([-1.0, -0.5, 0.25, 0.5, 1.0], np.float32), d)
                gi, gs = ix.search_optimized(q, limit)
                wi, ws = oracle.search_optimized_i8(q, codes, limit)
                assert gi.tolist() == wi.tolist(), (limit, gi[:12], wi[:12])
                assert gs.tobytes() == ws.tobytes()
    finally:
        ix.close()


@pytest.mark.parametrize("pdl", [0, 1, 2])
def test_back_to_back_device_resident_searches(cg, oracle, pdl):
    """The throughput path: hundreds of batch-1 searches enqueued on one stream with device-resident I/O and no host
    sync in between.  pdl=1 overlaps each merge with the next scan; pdl=2 chains every kernel programmatically
    (early release + griddepcontrol.wait before the first conflicting access).  Every result must still be exact."""
    import torch
    rng = np.random.default_rng(60 + pdl)
    n, d, k, steps = 60_000, 256, 10, 300
    rows = rng.standard_normal((n, d)).astype(np.float32)
    qs = rng.standard_normal((steps, d)).astype(np.float32)
    ix = cg.Index(d)
    ix.add(rows)
    ix.set_option("pdl", pdl)
    dq = torch.from_numpy(qs).cuda()
    o_r = torch.zeros((steps, k), dtype=torch.int64, device="cuda")
    o_s = torch.zeros((steps, k), dtype=torch.float32, device="cuda")
    o_c = torch.zeros((steps,), dtype=torch.int32, device="cuda")
    st = torch.cuda.Stream()
    for rep in range(2):
        for i in range(steps):
            ix.search_device(dq[i].data_ptr(), 1, k, o_r[i].data_ptr(), o_s[i].data_ptr(), o_c[i].data_ptr(), cg.COSINE, st.cuda_stream)
    torch.cuda.synchronize()
    got_r, got_s = o_r.cpu().numpy(), o_s.cpu().numpy()
    assert int(o_c.min()) == k
    for i in range(0, steps, 3):
        wi, ws = oracle.parallel_top_k_search(qs[i], rows, k)
        assert got_r[i].tolist() == wi.tolist() and got_s[i].tobytes() == ws.tobytes(), i
    ix.close()


def test_symbol_resolver_argmax(cg, oracle):
    """indexer.rs:2827-2843 — best candidate strictly above 0.75 by the search.rs:519-533 cosine."""
    rng = np.random.default_rng(88)
    embs = rng.standard_normal((400, 384)).astype(np.float32)
    target = embs[123] + 0.3 * rng.standard_normal(384).astype(np.float32)
    ix = cg.Index(384)
    ix.add(embs)
    cands = np.array([5, 123, 77, 300, 123], np.uint64)
    got = cg.resolve_symbol(ix, target, cands)
    sims = [oracle.cosine_similarity_seq(target, embs[int(c)]) for c in cands]
    assert got is not None and got[0] == 123 and np.float32(got[1]) == np.float32(sims[1]) and got[1] > 0.75
    assert cg.resolve_symbol(ix, target, np.array([5, 77, 300], np.uint64)) is None       # nothing above the threshold
    assert cg.resolve_symbol(ix, target, np.array([], np.uint64)) is None
    ix.close()


def test_symbol_resolver_batched_over_the_whole_index(cg, oracle):
    """indexer.rs:2827-2843 arg-max for several unresolved references in one call, all symbols as candidates."""
    rng = np.random.default_rng(2843)
    embs = np.stack([oracle.hash_text_embedding(f"sym_{i}", 384) for i in range(3000)])
    embs[1700] = embs[40]                                                       # duplicate symbol: the first one wins
    targets = np.stack([embs[40], embs[2999] + np.float32(0.02) * rng.standard_normal(384).astype(np.float32),
                        rng.standard_normal(384).astype(np.float32)])          # exact, near, unrelated
    ix = cg.Index(384)
    ix.add(embs)
    got = cg.resolve_symbols(ix, targets)
    for u in range(3):
        wi, ws = oracle.inmemory_search_similar(targets[u], embs, 1)
        if ws[0] > 0.75:
            assert got[u] is not None and got[u][0] == int(wi[0]) and np.float32(got[u][1]).tobytes() == np.float32(ws[0]).tobytes()
        else:
            assert got[u] is None
    assert got[0][0] == 40 and got[1][0] == 2999 and got[2] is None
    assert cg.resolve_symbols(ix, np.zeros((0, 384), np.float32)) == []
    ix.close()


def test_parallel_vector_ops_mirror(cg, oracle):
    """ParallelVectorOps::{parallel_batch_similarity, parallel_normalize_vectors} (simd_ops.rs:347-358, 386-419)."""
    rng = np.random.default_rng(91)
    emb = rng.standard_normal((500, 70)).astype(np.float32)
    emb[3] = 0.0
    q = rng.standard_normal(70).astype(np.float32)
    sims = cg.ParallelVectorOps.parallel_batch_similarity(q, emb)
    assert sims.tobytes() == oracle.scores(q, emb).tobytes()
    norm = cg.ParallelVectorOps.parallel_normalize_vectors(emb)
    assert norm.tobytes() == np.stack([oracle.normalize_avx2(r) for r in emb]).tobytes()


def test_query_stream_double_buffered_batches(cg, oracle):
    """cgvec_stream_*: batches submitted one after the other (upload of batch i+1 behind the scan of batch i) return
    exactly what one-shot searches return, in order, including a short last batch; flush drains the pipeline."""
    rng = np.random.default_rng(91)
    n, d, k = 50_000, 256, 10
    rows = (rng.standard_normal((n, d)) / 16).astype(np.float32)
    ix = cg.Index(d, cg.F16)
    try:
        ix.add(rows)
        ref = rows.astype(np.float16).astype(np.float32)
        batches = [rng.standard_normal((m, d)).astype(np.float32) for m in (64, 64, 17, 64)]
        stream = ix.stream(64, k, cg.COSINE, cg.PATH_AUTO)
        got = []
        assert stream.submit(batches[0]) is None                 # nothing to return yet
        for b in batches[1:]:
            got.append(stream.submit(b))
        got.append(stream.flush())
        assert stream.flush() is None
        stream.close()
        for b, (r, s, c) in zip(batches, got):
            assert r.shape == (len(b), k)
            for qi in range(0, len(b), 5):
                wi, ws = oracle.parallel_top_k_search(b[qi], ref, k)
                assert r[qi].tolist() == wi.tolist() and s[qi].tobytes() == ws.tobytes() and int(c[qi]) == k
        with pytest.raises(cg.CgvecError):
            ix.stream(64, k).submit(np.zeros((65, d), np.float32))
    finally:
        ix.close()


def test_chunk_candidate_stage_mirror(cg, oracle):
    """steps 1-2 of fn::semantic_search_nodes_via_chunks (schema/codegraph.surql:318-417) behind
    execute_semantic_code_search (graph_tool_executor.rs:578-591) over the exact scan: the 100 nearest chunks, orphan
    chunks dropped AFTER the KNN, LIMIT 3 * safe_limit, vector_score = 1 - distance, parents as node ids."""
    import uuid
    rng = np.random.default_rng(331)
    n, d = 5000, 384
    emb = rng.standard_normal((n, d)).astype(np.float32)
    emb /= np.linalg.norm(emb, axis=1, keepdims=True)
    chunk_ids = [uuid.UUID(int=i + 1) for i in range(n)]
    parents = [None if i % 7 == 3 else uuid.UUID(int=10_000_000 + i // 4) for i in range(n)]
    stage = cg.ChunkCandidateStage(d)
    stage.upsert_chunks([cg.ChunkRecord(chunk_ids[i], parents[i], emb[i]) for i in range(n)])
    q = (emb[123] + 0.2 * rng.standard_normal(d).astype(np.float32) / np.sqrt(d)).astype(np.float32)
    w100, s100 = oracle.parallel_top_k_search(q, emb, 100)
    for limit, safe in ((5, 5), (40, 40), (0, 10), (101, 10)):
        got = stage.candidates(q, limit)
        want = [(int(i), np.float32(1.0) - (np.float32(1.0) - s)) for i, s in zip(w100, s100) if parents[int(i)] is not None][: 3 * safe]
        assert [c.chunk_id for c in got] == [chunk_ids[i] for i, _ in want]
        assert [c.node_id for c in got] == [parents[i] for i, _ in want]
        assert np.float32([c.vector_score for c in got]).tobytes() == np.float32([s for _, s in want]).tobytes()
    assert stage.backend.last_column == "embedding_384"
    stage.store.index.close()
